// hostemu.cpp — compiles the DEVICE trace code (csrc/rt_trace.cuh, rt_scene.cuh) as plain C++ so that its logic can
// be checked against the oracle on a machine without a GPU (`pytest -m "not gpu"`).
// TEST-ONLY: this is not a CPU fallback — librtb200.so never contains or loads it, and nothing outside tests/ does.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../uu-infogr-raytracer_b200/csrc/rt_scene.cuh"
#include "../../uu-infogr-raytracer_b200/csrc/rt_primary_bins.cuh"
#include "../../uu-infogr-raytracer_b200/csrc/rt_gate.cuh"
#include "../../uu-infogr-raytracer_b200/csrc/rt_tiles.cuh"

using namespace rtb;

// Host-side LBVH build with the same pieces the CUDA build kernels use (morton_key, karras_node, sphere_box, box_union).
#include <algorithm>
#include <random>
struct HostBvh {
    std::vector<BvhNode> nodes, nodes_cam; std::vector<f4> sorted; std::vector<int> orig; float r2max = 0; int n = 0;
    BvhView view() const {
        BvhView v; v.nodes = nodes.data(); v.nodes_cam = nodes_cam.empty() ? nullptr : nodes_cam.data();
        v.sgeom_sorted = sorted.data(); v.orig = orig.data(); v.n = n; v.r2max = r2max; return v;
    }
};
// camera copy: same topology, leaf boxes inflated for rays starting at the camera (k_lbvh_refit_cam on the device)
static BvhBox host_refit_cam(HostBvh& b, f3 cam, int child) {
    if (child < 0) {
        f4 g = b.sorted[~child];
        float dx = cam.x - g.x, dy = cam.y - g.y, dz = cam.z - g.z;
        return sphere_box(g, inflated_radius(g.w, dx * dx + dy * dy + dz * dz));
    }
    BvhNode& nd = b.nodes_cam[child];
    BvhBox b0 = host_refit_cam(b, cam, nd.c[0]), b1 = host_refit_cam(b, cam, nd.c[1]);
    node_set_child_box(nd, 0, b0); node_set_child_box(nd, 1, b1);
    return box_union(b0, b1);
}
static BvhBox host_refit(HostBvh& b, const std::vector<float>& reff, int child) {
    if (child < 0) return sphere_box(b.sorted[~child], reff[b.orig[~child]]);
    BvhNode& nd = b.nodes[child];
    BvhBox b0 = host_refit(b, reff, nd.c[0]), b1 = host_refit(b, reff, nd.c[1]);
    node_set_child_box(nd, 0, b0); node_set_child_box(nd, 1, b1);
    return box_union(b0, b1);
}
static void host_build(const std::vector<f4>& sg, HostBvh& b) {
    const int n = (int)sg.size();
    b.n = n; b.nodes.clear(); b.sorted.clear(); b.orig.clear(); b.r2max = 0;
    if (n < 2) return;
    float bmin[3] = {sg[0].x, sg[0].y, sg[0].z}, bmax[3] = {sg[0].x, sg[0].y, sg[0].z};
    std::vector<float> reff(n);
    for (int i = 0; i < n; i++) {
        bmin[0] = fminf(bmin[0], sg[i].x); bmin[1] = fminf(bmin[1], sg[i].y); bmin[2] = fminf(bmin[2], sg[i].z);
        bmax[0] = fmaxf(bmax[0], sg[i].x); bmax[1] = fmaxf(bmax[1], sg[i].y); bmax[2] = fmaxf(bmax[2], sg[i].z);
        float r2 = sg[i].w > 0 ? sg[i].w : 0; reff[i] = sqrtf(r2) * 1.000001f + 1e-30f; if (r2 > b.r2max) b.r2max = r2;
    }
    float binv[3];
    morton_scale(bmin, bmax, binv);
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; i++) keys[i] = morton_key(sg[i].x, sg[i].y, sg[i].z, bmin, binv, (uint32_t)i);
    std::sort(keys.begin(), keys.end());
    b.sorted.resize(n); b.orig.resize(n); b.nodes.resize(n - 1);
    for (int j = 0; j < n; j++) { int idx = (int)(keys[j] & 0xFFFFFFFFull); b.sorted[j] = sg[idx]; b.orig[j] = idx; }
    for (int i = 0; i < n - 1; i++) { int l, r; karras_node(keys.data(), n, i, &l, &r); memset(&b.nodes[i], 0, sizeof(BvhNode)); b.nodes[i].c[0] = l; b.nodes[i].c[1] = r; }
    host_refit(b, reff, 0);
}

// The uninstrumented tiny-scene path at one sample per pixel runs like the single-sample kernels: with the host's primary-ray gate.
static bool g_gate_on = false;
static FrameGates g_gates;
template <class SC, class DBG>
static uint32_t px_of(const SC& sc, const CamRec& cam, int x, int y, int w, int h, int d, int spp, uint32_t seed, HitRec* st, DBG& dbg) {
    scene_begin_pixel(sc, x, y, spp, 0);          // as render_loop / debug_loop (rtb200.cu)
    if constexpr (!DBG::count_tests) {          // NoDbg (the shipped kernels) and ProdDbg (k_debug_tiny_prod): gated
        if (g_gate_on) {
            // as render_loop (rtb200.cu): gates per span of 4 consecutive pixels (linear index aligned to 4), only inside one row
            const int p = y * w + x, ps = p & ~3, ys = ps / w, xs = ps - ys * w;
            uint32_t bits = 0u;
            if (xs + 4 <= w && ps + 4 <= w * h) {
                bool black = false;
                bits = gate_bits_span(g_gates, xs, xs + 3, ys, sc.n_lights(), &black);
                if (black) {                                  // render_loop's black-span branch: the reference's primary ray hits nothing
                    if constexpr (DBG::enabled) dbg.ray(0u, 1u, 0xFFFFFFFFu, 0.0f);
                    return 0u;
                }
            }
            return trace_pixel<true>(sc, cam, x, y, w, h, d, 1, seed, st, dbg, 0.0f, 0.0f, bits);
        }
    }
    return trace_pixel(sc, cam, x, y, w, h, d, spp, seed, st, dbg);
}
// use_tiny: 0 global-memory policy, 1 TinyScene<-1> (run-time count), 2 TinyScene<NS> with the exact compile-time count,
//           3 LBVH policy, 4 shared-memory-staged policy (here: an ordinary host array), 6 LBVH policy + per-frame primary bins
static const HostBvh* g_bvh = nullptr;
static PrimaryBinsHost g_pbh;
static int g_pb_capacity = -1;                    // emu_set_primary_bins_capacity: < 0 = the device build's rule (16 per sphere ...)
static int pb_capacity_for(int n, const PbCam& c) {
    if (g_pb_capacity >= 0) return g_pb_capacity;
    return pb_list_capacity(n, (long long)c.tiles_x * c.tiles_y);
}
extern "C" void emu_set_primary_bins_capacity(int capacity) { g_pb_capacity = capacity; }
// The DEVICE build's algorithm (rt_primary_bins_build.cuh: k_pb_bin<false>, k_pb_alloc, k_pb_bin<true>) replayed on the host with the
// freedom its atomics have: spheres are counted and filled in two independent random orders (the order in which warps reach their
// atomics), "everywhere" slots are handed out in count order, runs of the list array are handed out per group of 32 consecutive tiles
// (k_pb_alloc's warp: prefix sum inside, one cursor atomic per warp) in a random group order, the rectangle of the count pass is kept for
// the fill pass, and every sphere walks its rectangle with the kernel's own step -> tile arithmetic (pb_walk_tile).
static uint64_t g_pb_shuffle = 0;                 // emu_set_primary_bins_shuffle: 0 = primary_bins_build_host, else the seed of the replay
extern "C" void emu_set_primary_bins_shuffle(uint64_t seed) { g_pb_shuffle = seed; }
static void primary_bins_build_device_order(const PbCam& c, const f4* sgeom, int n, int capacity, uint64_t seed, PrimaryBinsHost* out) {
    out->valid = false;
    memset(&out->hdr, 0, sizeof(out->hdr));
    if (c.eps < 0 || n <= 0) return;
    out->tiles_x = c.tiles_x; out->tiles_y = c.tiles_y;
    const size_t nt = (size_t)c.tiles_x * (size_t)c.tiles_y;
    std::vector<int> count(nt, 0), fill(nt, 0);
    out->tiles.assign(nt, PbTile{-1, 0});
    out->geom.assign((size_t)capacity, f4()); out->orig.assign((size_t)capacity, -1);
    struct Rect { int x0, y0, x1, y1; };
    std::vector<Rect> rects((size_t)n, Rect{0, 0, -1, -1});
    std::mt19937_64 rng(seed);
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; i++) order[(size_t)i] = i;
    std::shuffle(order.begin(), order.end(), rng);
    for (int i : order) {                                                    // k_pb_bin<false>
        int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
        const int kind = pb_sphere_tiles(c, sgeom[i], &x0, &y0, &x1, &y1);
        if (kind == 2) { const int s = out->hdr.n_everywhere++; if (s < PB_MAX_EVERYWHERE) { out->hdr.ev_geom[s] = sgeom[i]; out->hdr.ev_orig[s] = i; } }
        if (kind != 1) continue;
        rects[(size_t)i] = Rect{x0, y0, x1, y1};
        const int tw = x1 - x0 + 1, cells = tw * (y1 - y0 + 1);
        for (int k = 0; k < cells; k++) count[(size_t)pb_walk_tile(k, tw, x0, y0, c.tiles_x)]++;
    }
    std::vector<size_t> groups((nt + 31) / 32);                              // k_pb_alloc
    for (size_t g = 0; g < groups.size(); g++) groups[g] = g;
    std::shuffle(groups.begin(), groups.end(), rng);
    for (size_t g : groups) {
        int incl = 0;
        const int base = out->hdr.cursor;
        for (size_t t = g * 32; t < g * 32 + 32 && t < nt; t++) {
            const int cnt = count[t] > PB_CAP ? 0 : count[t];
            incl += cnt;
            out->tiles[t] = pb_tile_decide(count[t], base + incl - cnt, capacity);
        }
        out->hdr.cursor += incl;
    }
    std::shuffle(order.begin(), order.end(), rng);
    for (int i : order) {                                                    // k_pb_bin<true>
        const Rect r = rects[(size_t)i];
        if (r.x1 < r.x0 || r.y1 < r.y0) continue;
        const int tw = r.x1 - r.x0 + 1, cells = tw * (r.y1 - r.y0 + 1);
        for (int k = 0; k < cells; k++) {
            const size_t t = (size_t)pb_walk_tile(k, tw, r.x0, r.y0, c.tiles_x);
            const PbTile tl = out->tiles[t];
            if (tl.n < 0) continue;
            const int s = tl.start + fill[t]++;
            out->geom[(size_t)s] = sgeom[i]; out->orig[(size_t)s] = i;
        }
    }
    out->valid = true;
}
static void build_primary_bins(const PbCam& c, const f4* sgeom, int n, PrimaryBinsHost* out) {
    if (g_pb_shuffle) primary_bins_build_device_order(c, sgeom, n, pb_capacity_for(n, c), g_pb_shuffle, out);
    else primary_bins_build_host(c, sgeom, n, pb_capacity_for(n, c), out);
}
static ShadowGridsHost g_sgh;
static ShadowGridsView sg_view() {
    ShadowGridsView v; memset(&v, 0, sizeof(v));
    if (!g_sgh.empty() && !g_sgh.cell_start.empty()) { v.grids = g_sgh.grids.data(); v.cell_start = g_sgh.cell_start.data(); v.items = g_sgh.items.data(); v.lo = g_sgh.lo; v.hi = g_sgh.hi; }
    return v;
}
template <class DBG>
static uint32_t dispatch(int use_tiny, const TinySceneData& t, const GlobalSceneData& g, const CamRec& cam, int x, int y, int w, int h,
                         int d, int spp, uint32_t seed, HitRec* st, DBG& dbg) {
    if (use_tiny == 0) return px_of(GlobalScene(g), cam, x, y, w, h, d, spp, seed, st, dbg);
    if (use_tiny == 3) return px_of(LbvhScene(g, g_bvh->view(), sg_view()), cam, x, y, w, h, d, spp, seed, st, dbg);
    if (use_tiny == 4) return px_of(StagedScene(g, g.sgeom), cam, x, y, w, h, d, spp, seed, st, dbg);
    if (use_tiny == 6) return px_of(LbvhBinsScene(g, g_bvh->view(), sg_view(), g_pbh.view()), cam, x, y, w, h, d, spp, seed, st, dbg);
    if (use_tiny == 5) {
        // park / resume (the compacting kernel's two passes, sequentially): trace to the second hit, copy the state out the
        // way k_render_tiny_compact parks it, wipe the stack, restore, finish. spp == 1 only.
        TinyScene<-1, -1, -1> sc(t);
        f3 o, dir, C; int bounce = 0, top = 0;
        primary_ray(cam, (float)x, (float)y, (float)w, (float)h, 0.0f, 0.0f, &o, &dir);
        if (!trace_chain(sc, d, o, dir, bounce, top, st, 2, &C, dbg)) {
            struct { int bounce; f3 o, dir; HitRec rec[2]; } e = {bounce, o, dir, {st[0], st[1]}};
            for (int i = 0; i < 33; i++) memset(&st[i], 0xCD, sizeof(HitRec));
            o = e.o; dir = e.dir; bounce = e.bounce; top = 2; st[0] = e.rec[0]; st[1] = e.rec[1];
            bool done = trace_chain(sc, d, o, dir, bounce, top, st, -1, &C, dbg);
            if (!done) return 0xDEADBEEFu;
        }
        return pack_color(C);
    }
    if (use_tiny == 2 && t.np == 1 && (t.nl == 2 || t.nl == 4)) {
        // the device's exact (spheres, lights, 1 plane) instantiations with an even light count: the packed two-light pass
        // (shade_light_pair) is compiled in when the library is built with -DRT_EMULATE_F32X2
#define EMU_EXACT(NS, NL) if (t.ns == NS && t.nl == NL) return px_of(TinyScene<NS, NL, 1>(t), cam, x, y, w, h, d, spp, seed, st, dbg)
        EMU_EXACT(0, 2); EMU_EXACT(1, 2); EMU_EXACT(2, 2); EMU_EXACT(3, 2); EMU_EXACT(4, 2);
        EMU_EXACT(0, 4); EMU_EXACT(1, 4); EMU_EXACT(2, 4); EMU_EXACT(3, 4); EMU_EXACT(4, 4);
#undef EMU_EXACT
    }
    if (use_tiny == 2) {
        switch (t.ns) {
            case 0: return px_of(TinyScene<0>(t), cam, x, y, w, h, d, spp, seed, st, dbg);
            case 1: return px_of(TinyScene<1>(t), cam, x, y, w, h, d, spp, seed, st, dbg);
            case 2: return px_of(TinyScene<2>(t), cam, x, y, w, h, d, spp, seed, st, dbg);
            case 3: return px_of(TinyScene<3>(t), cam, x, y, w, h, d, spp, seed, st, dbg);
            case 4: return px_of(TinyScene<4>(t), cam, x, y, w, h, d, spp, seed, st, dbg);
            default: break;
        }
    }
    return px_of(TinyScene<-1>(t), cam, x, y, w, h, d, spp, seed, st, dbg);
}

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
    HostBvh bvh;
    if (use_tiny == 3 || use_tiny == 6) {
        if (ns < 2) return -2;
        host_build(sg, bvh);
        bvh.nodes_cam = bvh.nodes;
        host_refit_cam(bvh, mk3(cam15[0], cam15[1], cam15[2]), 0);
        g_bvh = &bvh;
        std::vector<f3> lp((size_t)nl);
        for (int i = 0; i < nl; i++) lp[i] = li[i].p;
        shadow_grids_build(sg, lp, &g_sgh);
        if (use_tiny == 6) { const PbCam pc = make_pb_cam(cam, w, h); build_primary_bins(pc, sg.data(), ns, &g_pbh); }
    }
    TinySceneData t; memset(&t, 0, sizeof(t));
    if (use_tiny == 1 || use_tiny == 2 || use_tiny == 5 || use_tiny == 11 || use_tiny == 12) {
        if (ns > TINY_MAX_SPHERES || np > TINY_MAX_PLANES || nl > TINY_MAX_LIGHTS) return -1;
        t.ns = ns; t.np = np; t.nl = nl; t.amb = g.amb;
        for (int i = 0; i < ns; i++) { t.sgeom[i] = sg[i]; t.smat[i] = sm[i]; }
        for (int i = 0; i < np; i++) t.planes[i] = pl[i];
        for (int i = 0; i < nl; i++) t.lights[i] = li[i];
        tiny_fill_pairs(t);                       // sphere / light pairs of the packed-fp32 paths (used with -DRT_EMULATE_F32X2)
    }
    const bool prod_dbg = use_tiny >= 10;         // 11 / 12: TinyScene<-1> / <exact> with the events-only policy of k_debug_tiny_prod
    if (prod_dbg) use_tiny -= 10;
    g_gate_on = (use_tiny == 1 || use_tiny == 2) && spp == 1;
    if (g_gate_on) g_gates = compute_frame_gates(cam, w, h, sg.data(), ns, pl.data(), np, li.data(), nl);
    uint64_t cnt[14] = {0};
    const bool dbg_mode = hash || aov_id || aov_t || counters;
#pragma omp parallel
    {
        HitRec stack[33];
        uint64_t lc[14] = {0};
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                size_t p = (size_t)y * w + x;
                if (dbg_mode && prod_dbg) {
                    ProdDbg dbg;
                    uint32_t c = dispatch(use_tiny, t, g, cam, x, y, w, h, max_depth, spp, seed, stack, dbg);
                    pixels[p] = (int32_t)c;
                    if (hash) hash[p] = dbg.hash;
                    if (aov_id) aov_id[p] = dbg.aov_id;
                    if (aov_t) aov_t[p] = dbg.aov_t;
                    lc[0] += dbg.primary; lc[1] += dbg.n_shadow; lc[2] += dbg.secondary;
                } else if (dbg_mode) {
                    FullDbg dbg;
                    uint32_t c = dispatch(use_tiny, t, g, cam, x, y, w, h, max_depth, spp, seed, stack, dbg);
                    pixels[p] = (int32_t)c;
                    if (hash) hash[p] = dbg.hash;
                    if (aov_id) aov_id[p] = dbg.aov_id;
                    if (aov_t) aov_t[p] = dbg.aov_t;
                    lc[0] += dbg.primary; lc[1] += dbg.n_shadow; lc[2] += dbg.secondary; lc[3] += dbg.sphere_tests;
                    lc[4] += dbg.sphere_disc_pos; lc[5] += dbg.plane_tests; lc[6] += dbg.shade_diffuse; lc[7] += dbg.shade_specular;
                    lc[8] += dbg.shade_mirror; lc[9] += dbg.shaded_hits;
                    lc[10] += dbg.node_visits[0]; lc[11] += dbg.node_visits[1]; lc[12] += dbg.node_visits[2]; lc[13] += dbg.fallbacks;
                } else {
                    NoDbg dbg;
                    uint32_t c = dispatch(use_tiny, t, g, cam, x, y, w, h, max_depth, spp, seed, stack, dbg);
                    pixels[p] = (int32_t)c;
                }
            }
#pragma omp critical
        for (int i = 0; i < 14; i++) cnt[i] += lc[i];
    }
    if (counters) for (int i = 0; i < 14; i++) counters[i] = cnt[i];
    return 0;
}

// The host's frame gates (rt_gate.cuh) for a scene and camera. rects: 6 x 4 ints (spheres, mirror, shadow[0..3]), inclusive,
// empty = {w, h, w, h}; affine: 2 x 3 floats (sky, deep); bits (nullable): w*h bytes = gate_bits() per pixel, bit 7 = black.
extern "C" int emu_gates(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl, const float* cam15,
                         int w, int h, int* rects, float* affine, unsigned char* bits) {
    std::vector<f4> sg((size_t)ns); std::vector<PlaneRec> pl((size_t)np); std::vector<LightRec> li((size_t)nl);
    for (int i = 0; i < ns; i++) { const float* f = spheres + 18 * (size_t)i; sg[i].x = f[0]; sg[i].y = f[1]; sg[i].z = f[2]; sg[i].w = f[17]; }
    for (int i = 0; i < np; i++) pl[i] = make_plane(planes + 20 * (size_t)i);
    for (int i = 0; i < nl; i++) li[i] = make_light(lights + 4 * (size_t)i);
    CamRec cam;
    cam.pos = mk3(cam15[0], cam15[1], cam15[2]); cam.right = mk3(cam15[3], cam15[4], cam15[5]);
    cam.up = mk3(cam15[6], cam15[7], cam15[8]); cam.fwd = mk3(cam15[9], cam15[10], cam15[11]);
    cam.view = mk3(cam15[12], cam15[13], cam15[14]);
    const FrameGates g = compute_frame_gates(cam, w, h, sg.data(), ns, pl.data(), np, li.data(), nl);
    const GateRect* rs[6] = {&g.spheres, &g.mirror, &g.shadow[0], &g.shadow[1], &g.shadow[2], &g.shadow[3]};
    for (int i = 0; i < 6; i++) { rects[4 * i] = rs[i]->x0; rects[4 * i + 1] = rs[i]->y0; rects[4 * i + 2] = rs[i]->x1; rects[4 * i + 3] = rs[i]->y1; }
    affine[0] = g.sky.a; affine[1] = g.sky.bx; affine[2] = g.sky.by; affine[3] = g.deep.a; affine[4] = g.deep.bx; affine[5] = g.deep.by;
    if (bits)
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            bool black = false;
            const uint32_t b = gate_bits(g, x, y, &black);
            bits[(size_t)y * w + x] = (unsigned char)(b | (black ? 0x80u : 0u));
        }
    return 0;
}

// Row plan of the sparse device -> host return (plan_rows, rt_gate.cuh): kind[h] (0 copy, 1 rectangle columns only, 2 black), rx[2].
extern "C" int emu_row_plan(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl, const float* cam15,
                            int w, int h, unsigned char* kind, int* rx) {
    std::vector<f4> sg((size_t)ns); std::vector<PlaneRec> pl((size_t)np); std::vector<LightRec> li((size_t)nl);
    for (int i = 0; i < ns; i++) { const float* f = spheres + 18 * (size_t)i; sg[i].x = f[0]; sg[i].y = f[1]; sg[i].z = f[2]; sg[i].w = f[17]; }
    for (int i = 0; i < np; i++) pl[i] = make_plane(planes + 20 * (size_t)i);
    for (int i = 0; i < nl; i++) li[i] = make_light(lights + 4 * (size_t)i);
    CamRec cam;
    cam.pos = mk3(cam15[0], cam15[1], cam15[2]); cam.right = mk3(cam15[3], cam15[4], cam15[5]);
    cam.up = mk3(cam15[6], cam15[7], cam15[8]); cam.fwd = mk3(cam15[9], cam15[10], cam15[11]);
    cam.view = mk3(cam15[12], cam15[13], cam15[14]);
    const FrameGates g = compute_frame_gates(cam, w, h, sg.data(), ns, pl.data(), np, li.data(), nl);
    RowPlan rp;
    plan_rows(g, w, h, &rp);
    for (int y = 0; y < h; y++) kind[y] = rp.kind[(size_t)y];
    rx[0] = rp.rx0; rx[1] = rp.rx1;
    return rp.sparse ? 1 : 0;
}

// The body of k_ray_log (rtb200.cu) on the host: LogDbg + the global-memory scene policy, one listed pixel after the other.
extern "C" int emu_ray_log(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl,
                           const float* ambient, const float* cam15, int w, int h, int max_depth,
                           const uint32_t* pixels, int n_pixels, void* out_records, int max_records) {
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
    const int slots = (max_depth + 2) + (max_depth + 1) * nl;
    std::vector<RayRec> buf((size_t)slots);
    HitRec stack[33];
    RayRec* out = (RayRec*)out_records;
    long long total = 0;
    for (int i = 0; i < n_pixels; i++) {
        const uint32_t p = pixels[i];
        LogDbg dbg; dbg.out = buf.data(); dbg.cap = (uint32_t)slots; dbg.pixel = p;
        const int y = (int)(p / (uint32_t)w), x = (int)(p - (uint32_t)y * (uint32_t)w);
        trace_pixel<true>(GlobalScene(g), cam, x, y, w, h, max_depth, 1, 0u, stack, dbg);
        if (dbg.n > (uint32_t)slots) return -3;                    // the slot bound rt_ray_log relies on
        for (uint32_t k = 0; k < dbg.n; k++, total++)
            if (total < max_records) out[total] = buf[k];
    }
    return (int)total;
}

// Single-ray sphere queries through the device query code: accel 1 brute, 2 LBVH (same semantics as rt_query_spheres).
extern "C" int emu_query(const float* spheres, int ns, const float* rays6, int n_rays, int kind, int accel, int32_t* out_id, float* out_t) {
    std::vector<f4> sg((size_t)ns); std::vector<MatRec> sm((size_t)ns);
    for (int i = 0; i < ns; i++) {
        const float* f = spheres + 18 * (size_t)i;
        sg[i].x = f[0]; sg[i].y = f[1]; sg[i].z = f[2]; sg[i].w = f[17]; sm[i] = make_mat(f + 4);
    }
    GlobalSceneData g; memset(&g, 0, sizeof(g));
    g.ns = ns; g.sgeom = sg.data(); g.smat = sm.data();
    HostBvh bvh;
    if (accel == 2) { if (ns < 2) return -2; host_build(sg, bvh); }
    BvhView bv = bvh.view();
#pragma omp parallel for schedule(dynamic, 256)
    for (int r = 0; r < n_rays; r++) {
        NoDbg dbg;
        f3 o = mk3(rays6[6 * r], rays6[6 * r + 1], rays6[6 * r + 2]);
        f3 d = mk3(rays6[6 * r + 3], rays6[6 * r + 4], rays6[6 * r + 5]);
        float a = dot3(d, d), a2 = 2 * a, a4 = 4 * a;
        int sel = -1; float t = 0.0f;
        auto run = [&](auto sc) {
            if (kind == 0) { sc.nearest(o, d, a2, a4, 0.0f, &sel, &t, dbg); if (sel < 0) t = 0.0f; }
            else if (kind == 1) { sc.nearest(o, d, a2, a4, 0.01f, &sel, &t, dbg); if (sel < 0) t = 0.0f; }
            else { sel = sc.shadow_any(-1, o, d, a2, a4, dbg) ? 1 : 0; t = 0.0f; }
        };
        ShadowGridsView nosg; memset(&nosg, 0, sizeof(nosg));
        if (accel == 2) run(LbvhScene(g, bv, nosg)); else run(GlobalScene(g));
        out_id[r] = sel; out_t[r] = t;
    }
    return 0;
}

// Shadow-bin soundness (rt_shadow_grid.cuh): for every query point and every light, the answer of the binned query must equal the
// reference's loop over ALL spheres (RayTracer.cs:573-582). out[0] = queries decided by the bins, out[1] = mismatches among them,
// out[2] = queries left to the traversal (outside the validity box / no grid), out[3] = occluded answers, out[4] = sphere tests run.
extern "C" int emu_shadow_bins_check(const float* spheres, int ns, const float* lights, int nl, const float* points3, int n_points, uint64_t* out) {
    std::vector<f4> sg((size_t)ns); std::vector<LightRec> li((size_t)nl); std::vector<f3> lp((size_t)nl);
    for (int i = 0; i < ns; i++) { const float* f = spheres + 18 * (size_t)i; sg[i].x = f[0]; sg[i].y = f[1]; sg[i].z = f[2]; sg[i].w = f[17]; }
    for (int i = 0; i < nl; i++) { li[i] = make_light(lights + 4 * (size_t)i); lp[i] = li[i].p; }
    ShadowGridsHost h;
    shadow_grids_build(sg, lp, &h);
    ShadowGridsView v; memset(&v, 0, sizeof(v));
    if (!h.empty() && !h.cell_start.empty()) { v.grids = h.grids.data(); v.cell_start = h.cell_start.data(); v.items = h.items.data(); v.lo = h.lo; v.hi = h.hi; }
    uint64_t decided_n = 0, bad = 0, undecided = 0, occ_n = 0, tests = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : decided_n, bad, undecided, occ_n, tests)
    for (int p = 0; p < n_points; p++) {
        const f3 P = mk3(points3[3 * p], points3[3 * p + 1], points3[3 * p + 2]);
        for (int l = 0; l < nl; l++) {
            FullDbg dbg;
            bool decided = false;
            const bool got = shadow_grid_any(v, l, P, li[l].p, li[l].a2, li[l].a4, &decided, dbg);
            if (!decided) { undecided++; continue; }
            decided_n++; tests += dbg.sphere_tests;
            bool want = false; NoDbg nd;
            for (int i = 0; i < ns; i++) {
                float t;
                if (sphere_hit(sub3(P, mk3(sg[i].x, sg[i].y, sg[i].z)), li[l].p, sg[i].w, li[l].a2, li[l].a4, 0.001f, &t, nd)) want = true;
            }
            if (want != got) bad++;
            if (got) occ_n++;
        }
    }
    out[0] = decided_n; out[1] = bad; out[2] = undecided; out[3] = occ_n; out[4] = tests;
    return 0;
}

// Primary bins (rt_primary_bins.cuh), checked directly: for every pixel of the frame the kernel's own primary ray (primary_ray) is
// tested against ALL spheres with the reference's test (:613-642, :975-981); every sphere that reports a hit must be in the list of
// the pixel's tile (or in the everywhere list) unless the tile keeps no list — and the fold over the list must select what the
// brute-force fold selects. out[0] = pixels answered by the bins, out[1] = pixels left to the traversal, out[2] = reported hits
// missing from a list (must be 0), out[3] = pixels whose fold differs (id or t bits; must be 0), out[4] = sphere tests run by the
// bins, out[5] = tiles without a list, out[6] = list entries, out[7] = spheres in the everywhere list, out[8] = bins valid (0 / 1).
extern "C" int emu_primary_bins_check(const float* spheres, int ns, const float* cam15, int w, int h, uint64_t* out) {
    std::vector<f4> sg((size_t)ns);
    for (int i = 0; i < ns; i++) { const float* f = spheres + 18 * (size_t)i; sg[i].x = f[0]; sg[i].y = f[1]; sg[i].z = f[2]; sg[i].w = f[17]; }
    CamRec cam;
    cam.pos = mk3(cam15[0], cam15[1], cam15[2]); cam.right = mk3(cam15[3], cam15[4], cam15[5]);
    cam.up = mk3(cam15[6], cam15[7], cam15[8]); cam.fwd = mk3(cam15[9], cam15[10], cam15[11]);
    cam.view = mk3(cam15[12], cam15[13], cam15[14]);
    PrimaryBinsHost pbh;
    const PbCam pc = make_pb_cam(cam, w, h);
    build_primary_bins(pc, sg.data(), ns, &pbh);
    const PrimaryBinsView v = pbh.view();
    uint64_t by_bins = 0, by_tree = 0, missing = 0, differ = 0, tests = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : by_bins, by_tree, missing, differ, tests)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            f3 o, dir;
            primary_ray(cam, (float)x, (float)y, (float)w, (float)h, 0.0f, 0.0f, &o, &dir);
            const float a = dot3(dir, dir), a2 = 2 * a, a4 = 4 * a;
            FullDbg dbg;
            int sel = -1; float t_sel = RT_INF;
            if (!primary_bins_nearest(v, x, y, o, dir, a2, a4, &sel, &t_sel, dbg)) { by_tree++; continue; }
            by_bins++; tests += dbg.sphere_tests;
            const PbTile tl = v.tiles[(size_t)(y >> PB_TILE_SHIFT) * v.tiles_x + (x >> PB_TILE_SHIFT)];
            int want = -1; float t_want = RT_INF; NoDbg nd;
            for (int i = 0; i < ns; i++) {
                float t;
                if (!(sphere_hit(sub3(o, mk3(sg[i].x, sg[i].y, sg[i].z)), dir, sg[i].w, a2, a4, 0.0f, &t, nd) && t > 0)) continue;
                if (t < t_want) { t_want = t; want = i; }                     // :977 strict '>' in array order
                bool listed = false;
                for (int k = 0; k < tl.n && !listed; k++) listed = v.orig[tl.start + k] == i;
                for (int k = 0; k < v.hdr->n_everywhere && k < PB_MAX_EVERYWHERE && !listed; k++) listed = v.hdr->ev_orig[k] == i;
                if (!listed) missing++;
            }
            if (want != sel || f2bits(t_want) != f2bits(t_sel)) differ++;
        }
    uint64_t no_list = 0, entries = 0;
    for (size_t t = 0; t < pbh.tiles.size(); t++) { if (pbh.tiles[t].n < 0) no_list++; else entries += (uint64_t)pbh.tiles[t].n; }
    out[0] = by_bins; out[1] = by_tree; out[2] = missing; out[3] = differ; out[4] = tests; out[5] = no_list; out[6] = entries;
    out[7] = pbh.valid ? (uint64_t)pbh.hdr.n_everywhere : 0; out[8] = pbh.valid ? 1 : 0;
    return 0;
}

// 2-D pixel blocks (rt_tiles.cuh): runs every (tile, item, thread) of a frame through the kernels' own mapping and counts how often
// each pixel is owned. kind 0: render_loop (ppt 1 or 4), kind 1: k_render_tiny_pack (ppt 4, w % 128 == 0). cover: w*h bytes.
// Returns the number of spans that leave the frame or their tile (must be 0).
extern "C" int emu_tile_cover(int kind, int ppt, int w, int h, int tile_rows, unsigned char* cover) {
    int bad = 0;
    const int tiles = (h + tile_rows - 1) / tile_rows;
    const int items = kind == 0 ? tile2d_items_per_tile(ppt, w, tile_rows) : (tile_rows / 4) * (w / 128);
    if (kind == 1 && items != tile2d_items_per_tile(4, w, tile_rows)) return -1;      // the launch sizes both from chunks_per_tile
    for (int t = 0; t < tiles; t++)
        for (int b = 0; b < items; b++)
            for (int tid = 0; tid < 128; tid++) {
                int x, y, cb = 0, yf = 0;
                const bool ok = kind == 0 ? tile2d_span(ppt, w, h, t, tile_rows, b, tid, &x, &y) : pack2d_span(w, h, t, tile_rows, b, tid, &x, &y, &cb, &yf);
                if (!ok) continue;
                if (y < t * tile_rows || y >= (t + 1) * tile_rows || y >= h || x < 0) { bad++; continue; }
                if (kind == 1 && (cb != x / 128 || yf > y || y - yf > 3 || (yf - t * tile_rows) % 4 != 0)) bad++;
                for (int q = 0; q < ppt; q++) {
                    if (x + q >= w) { if (kind == 1 || ppt == 4) bad++; continue; }   // ragged spans only exist on the 1-pixel paths' right edge (none there)
                    cover[(size_t)y * w + x + q]++;
                }
            }
    return bad;
}
