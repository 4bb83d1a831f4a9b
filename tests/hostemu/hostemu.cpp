// hostemu.cpp — compiles the DEVICE trace code (csrc/rt_trace.cuh, rt_scene.cuh) as plain C++ so that its logic can
// be checked against the oracle on a machine without a GPU (`pytest -m "not gpu"`).
// TEST-ONLY: this is not a CPU fallback — librtb200.so never contains or loads it, and nothing outside tests/ does.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../uu-infogr-raytracer_b200/csrc/rt_scene.cuh"

using namespace rtb;

extern "C" int emu_render(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl,
                          const float* ambient, const float* cam15, int w, int h, int max_depth, int spp, uint32_t seed,
                          int use_tiny, int32_t* pixels, uint32_t* hash, int32_t* aov_id, float* aov_t, uint64_t* counters) {
    std::vector<f4> sg((size_t)ns); std::vector<MatRec> sm((size_t)ns);
    std::vector<PlaneRec> pl((size_t)np); std::vector<LightRec> li((size_t)nl);
    for (int i = 0; i < ns; i++) {
        const float* f = spheres + 18 * (size_t)i;
        sg[i].x = f[0]; sg[i].y = f[1]; sg[i].z = f[2]; sg[i].w = f[17];
        sm[i] = make_mat(f + 4);
    }
    for (int i = 0; i < np; i++) pl[i] = make_plane(planes + 20 * (size_t)i);
    for (int i = 0; i < nl; i++) li[i] = make_light(lights + 4 * (size_t)i);
    CamRec cam;
    cam.pos = mk3(cam15[0], cam15[1], cam15[2]); cam.right = mk3(cam15[3], cam15[4], cam15[5]);
    cam.up = mk3(cam15[6], cam15[7], cam15[8]); cam.fwd = mk3(cam15[9], cam15[10], cam15[11]);
    cam.view = mk3(cam15[12], cam15[13], cam15[14]);

    GlobalSceneData g; memset(&g, 0, sizeof(g));
    g.ns = ns; g.np = np; g.nl = nl; g.amb = mk3(ambient[0], ambient[1], ambient[2]);
    g.sgeom = sg.data(); g.smat = sm.data(); g.planes = pl.data(); g.lights = li.data();
    TinySceneData t; memset(&t, 0, sizeof(t));
    if (use_tiny) {
        if (ns > TINY_MAX_SPHERES || np > TINY_MAX_PLANES || nl > TINY_MAX_LIGHTS) return -1;
        t.ns = ns; t.np = np; t.nl = nl; t.amb = g.amb;
        for (int i = 0; i < ns; i++) { t.sgeom[i] = sg[i]; t.smat[i] = sm[i]; }
        for (int i = 0; i < np; i++) t.planes[i] = pl[i];
        for (int i = 0; i < nl; i++) t.lights[i] = li[i];
    }
    uint64_t cnt[10] = {0};
    const bool dbg_mode = hash || aov_id || aov_t || counters;
#pragma omp parallel
    {
        HitRec stack[33];
        uint64_t lc[10] = {0};
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                size_t p = (size_t)y * w + x;
                if (dbg_mode) {
                    FullDbg dbg;
                    uint32_t c = use_tiny ? trace_pixel(TinyScene(t), cam, x, y, w, h, max_depth, spp, seed, stack, dbg)
                                          : trace_pixel(GlobalScene(g), cam, x, y, w, h, max_depth, spp, seed, stack, dbg);
                    pixels[p] = (int32_t)c;
                    if (hash) hash[p] = dbg.hash;
                    if (aov_id) aov_id[p] = dbg.aov_id;
                    if (aov_t) aov_t[p] = dbg.aov_t;
                    lc[0] += dbg.primary; lc[1] += dbg.n_shadow; lc[2] += dbg.secondary; lc[3] += dbg.sphere_tests;
                    lc[4] += dbg.sphere_disc_pos; lc[5] += dbg.plane_tests; lc[6] += dbg.shade_diffuse; lc[7] += dbg.shade_specular;
                    lc[8] += dbg.shade_mirror; lc[9] += dbg.shaded_hits;
                } else {
                    NoDbg dbg;
                    uint32_t c = use_tiny ? trace_pixel(TinyScene(t), cam, x, y, w, h, max_depth, spp, seed, stack, dbg)
                                          : trace_pixel(GlobalScene(g), cam, x, y, w, h, max_depth, spp, seed, stack, dbg);
                    pixels[p] = (int32_t)c;
                }
            }
#pragma omp critical
        for (int i = 0; i < 10; i++) cnt[i] += lc[i];
    }
    if (counters) for (int i = 0; i < 10; i++) counters[i] = cnt[i];
    return 0;
}
