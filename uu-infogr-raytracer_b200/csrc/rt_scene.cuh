// rt_scene.cuh — scene policies (where the SoA fp32 records live) for rt_trace.cuh.
//   TinyScene   : whole scene by value in the kernel parameter block (constant bank; uniform broadcast reads).
//                 The reference's default scene (3 spheres, 1 plane, 2 lights, RayTracer.cs:441-465) is 0.5 KB.
//   GlobalScene : records in global memory (L1/L2-resident), any size; base of the shared-memory staged and LBVH paths.
#pragma once
#include "rt_trace.cuh"
#include "rt_lbvh.cuh"
#include "rt_shadow_grid.cuh"

namespace rtb {

enum { TINY_MAX_SPHERES = 8, TINY_MAX_PLANES = 4, TINY_MAX_LIGHTS = 8 };

// Sphere geometry of spheres 2k and 2k+1 side by side, NEGATED (o - c == o + (-c) and x - r^2 == x + (-r^2) exactly), so that one
// packed-fp32 instruction (Blackwell FADD2 / FMUL2, rt_trace.cuh) serves both spheres with a 64-bit constant-bank operand.
// An odd sphere count is padded with a sphere that can never pass the exact test (-r^2 = +1e30).
struct SpherePair { float ncx[2], ncy[2], ncz[2], nr2[2]; };
struct TinySceneData {
    int ns, np, nl, pad;
    f3 amb; float pad1;
    SpherePair pairs[TINY_MAX_SPHERES / 2];
    f4 sgeom[TINY_MAX_SPHERES];
    MatRec smat[TINY_MAX_SPHERES];
    PlaneRec planes[TINY_MAX_PLANES];
    LightRec lights[TINY_MAX_LIGHTS];
    LightPair lpairs[TINY_MAX_LIGHTS / 2];
};

// NS >= 0: the sphere count is a compile-time constant — the sphere loops unroll and every record is addressed
// statically in the constant bank (operands fold into the FMUL/FADD, no load instructions). NS < 0: run-time count.
// NL likewise for the light loop of the shading code, NP for the plane loop.
template <int NS, int NL = -1, int NP = -1>
struct TinyScene {
    const TinySceneData& s;
    RT_HD explicit TinyScene(const TinySceneData& d) : s(d) {}
    RT_HD int n_spheres() const { return NS >= 0 ? NS : s.ns; }
    RT_HD int n_planes() const { return NP >= 0 ? NP : s.np; }
    RT_HD int n_lights() const { return NL >= 0 ? NL : s.nl; }
    RT_HD f3 ambient() const { return s.amb; }
    RT_HD f4 sphere_geom(int i) const { return s.sgeom[i]; }
    RT_HD const SpherePair& sphere_pair(int k) const { return s.pairs[k]; }
    static constexpr bool has_pairs = true;
    static constexpr int static_lights = NL;      // >= 0: compile-time light count (an even count takes the packed two-light pass)
    RT_HD const LightPair& light_pair(int k) const { return s.lpairs[k]; }
    RT_HD MatRec sphere_mat(int i) const { return s.smat[i]; }
    RT_HD uint32_t sphere_flags(int i) const { return s.smat[i].flags; }
    RT_HD f4 plane_n(int i) const { f4 r; r.x = s.planes[i].n.x; r.y = s.planes[i].n.y; r.z = s.planes[i].n.z; r.w = s.planes[i].cn; return r; }
    RT_HD f3 plane_e1(int i) const { return s.planes[i].e1; }
    RT_HD f3 plane_e2(int i) const { return s.planes[i].e2; }
    RT_HD MatRec plane_mat(int i) const { return s.planes[i].m; }
    RT_HD uint32_t plane_flags(int i) const { return s.planes[i].m.flags; }
    RT_HD LightRec light(int i) const { return s.lights[i]; }
    template <class DBG> RT_HD void nearest(f3 o, f3 d, float a2, float a4, float off, int* sel, float* t, DBG& dbg) const {
        brute_nearest<NS>(*this, o, d, a2, a4, off, sel, t, dbg);
    }
    template <class DBG> RT_HD bool shadow_any(int, f3 hit, f3 lp, float a2, float a4, DBG& dbg) const {
        return brute_shadow_any<NS>(*this, hit, lp, a2, a4, dbg);
    }
};

struct GlobalSceneData {
    int ns, np, nl, pad;
    f3 amb; float pad1;
    const f4* sgeom;
    const MatRec* smat;
    const PlaneRec* planes;
    const LightRec* lights;
};

RT_HD f4 load_f4(const f4* p) {
#if defined(__CUDA_ARCH__)
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return *p;
#endif
}

RT_HD MatRec load_mat(const MatRec* p) {
#if defined(__CUDA_ARCH__)
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    MatRec m;
    m.kd = mk3(a.x, a.y, a.z); m.n = a.w; m.ka = mk3(b.x, b.y, b.z); m.flags = __float_as_uint(b.w);
    m.ks = mk3(c.x, c.y, c.z); m.pad0 = 0; m.km = mk3(d.x, d.y, d.z); m.pad1 = 0;
    return m;
#else
    return *p;
#endif
}

struct GlobalScene {
    const GlobalSceneData& s;
    RT_HD explicit GlobalScene(const GlobalSceneData& d) : s(d) {}
    RT_HD int n_spheres() const { return s.ns; }
    RT_HD int n_planes() const { return s.np; }
    RT_HD int n_lights() const { return s.nl; }
    static constexpr int static_lights = -1;
    RT_HD f3 ambient() const { return s.amb; }
    RT_HD f4 sphere_geom(int i) const { return load_f4(s.sgeom + i); }
    RT_HD MatRec sphere_mat(int i) const { return load_mat(s.smat + i); }
    RT_HD uint32_t sphere_flags(int i) const { return s.smat[i].flags; }
    RT_HD f4 plane_n(int i) const { return load_f4(reinterpret_cast<const f4*>(&s.planes[i].n)); }
    RT_HD f3 plane_e1(int i) const { return s.planes[i].e1; }
    RT_HD f3 plane_e2(int i) const { return s.planes[i].e2; }
    RT_HD MatRec plane_mat(int i) const { return load_mat(&s.planes[i].m); }
    RT_HD uint32_t plane_flags(int i) const { return s.planes[i].m.flags; }
    RT_HD LightRec light(int i) const { return s.lights[i]; }
    template <class DBG> RT_HD void nearest(f3 o, f3 d, float a2, float a4, float off, int* sel, float* t, DBG& dbg) const {
        brute_nearest<-1>(*this, o, d, a2, a4, off, sel, t, dbg);
    }
    template <class DBG> RT_HD bool shadow_any(int, f3 hit, f3 lp, float a2, float a4, DBG& dbg) const {
        return brute_shadow_any<-1>(*this, hit, lp, a2, a4, dbg);
    }
};

// Brute force with the sphere geometry staged in shared memory (LDS.128 broadcast per test): BASELINE.json configs[2]
// ("1,024 random spheres ... shared-memory staged").  Materials / planes / lights stay in global memory.
struct StagedScene : GlobalScene {
    const f4* sm_geom;          // ns records in shared memory
    RT_HD StagedScene(const GlobalSceneData& d, const f4* smem) : GlobalScene(d), sm_geom(smem) {}
    RT_HD f4 sphere_geom(int i) const { return sm_geom[i]; }
    template <class DBG> RT_HD void nearest(f3 o, f3 d, float a2, float a4, float off, int* sel, float* t, DBG& dbg) const {
        brute_nearest<-1>(*this, o, d, a2, a4, off, sel, t, dbg);
    }
    template <class DBG> RT_HD bool shadow_any(int, f3 hit, f3 lp, float a2, float a4, DBG& dbg) const {
        return brute_shadow_any<-1>(*this, hit, lp, a2, a4, dbg);
    }
};

// LBVH over the spheres (rt_lbvh.cuh) for nearest-hit queries, per-light projected bins (rt_shadow_grid.cuh) for shadow rays;
// planes stay in the brute-force side list (they are infinite).
struct LbvhScene : GlobalScene {
    BvhView bv;
    ShadowGridsView sg;
    RT_HD LbvhScene(const GlobalSceneData& d, const BvhView& v, const ShadowGridsView& g) : GlobalScene(d), bv(v), sg(g) {}
    template <class DBG> RT_HD void nearest(f3 o, f3 d, float a2, float a4, float off, int* sel, float* t, DBG& dbg) const {
        const GlobalScene& base = *this;
        bvh_nearest(bv, o, d, a2, a4, off, sel, t, dbg,
                    [&](int* s2, float* t2) { NoDbg nd; brute_nearest<-1>(base, o, d, a2, a4, off, s2, t2, nd); });
    }
    template <class DBG> RT_HD bool shadow_any(int li, f3 hit, f3 lp, float a2, float a4, DBG& dbg) const {
        bool decided;
        const bool occ = shadow_grid_any(sg, li, hit, lp, a2, a4, &decided, dbg);
        if (decided) return occ;
        const GlobalScene& base = *this;
        return bvh_shadow_any(bv, hit, lp, a2, a4, dbg, [&]() { NoDbg nd; return brute_shadow_any<-1>(base, hit, lp, a2, a4, nd); });
    }
};

inline void tiny_fill_pairs(TinySceneData& t) {
    for (int k = 0; k < TINY_MAX_SPHERES / 2; k++)
        for (int h = 0; h < 2; h++) {
            const int i = 2 * k + h;
            SpherePair& p = t.pairs[k];
            if (i < t.ns) { p.ncx[h] = -t.sgeom[i].x; p.ncy[h] = -t.sgeom[i].y; p.ncz[h] = -t.sgeom[i].z; p.nr2[h] = -t.sgeom[i].w; }
            else { p.ncx[h] = 0.0f; p.ncy[h] = 0.0f; p.ncz[h] = 0.0f; p.nr2[h] = 1e30f; }
        }
    for (int k = 0; k < TINY_MAX_LIGHTS / 2; k++)
        for (int h = 0; h < 2; h++) {
            const int i = 2 * k + h;
            LightPair& p = t.lpairs[k];
            if (i < t.nl) { p.px[h] = t.lights[i].p.x; p.py[h] = t.lights[i].p.y; p.pz[h] = t.lights[i].p.z; p.intensity[h] = t.lights[i].intensity; }
            else { p.px[h] = 0.0f; p.py[h] = 0.0f; p.pz[h] = 0.0f; p.intensity[h] = 0.0f; }
        }
}

// ---------------------------------------------------------------------------------------------------------
// Upload-time record construction (host, strict fp32: this file is compiled with -ffp-contract=off on the host side)
// ---------------------------------------------------------------------------------------------------------
inline MatRec make_mat(const float* f) {                      // 13 floats Kd,Ka,Ks,n,Km (RayTracer.cs:64-80)
    MatRec m;
    m.kd = mk3(f[0], f[1], f[2]); m.ka = mk3(f[3], f[4], f[5]); m.ks = mk3(f[6], f[7], f[8]); m.n = f[9];
    m.km = mk3(f[10], f[11], f[12]); m.pad0 = 0; m.pad1 = 0;
    auto is_zero = [](f3 v) { return v.x == 0 && v.y == 0 && v.z == 0; };          // VecUtil.IsZero :52
    m.flags = (!is_zero(m.km) ? MAT_MIRROR : 0u)                                   // IsMirror :85
            | (!is_zero(m.kd) ? MAT_DIFFUSE : 0u)                                  // IsDiffuse :89
            | ((!is_zero(m.ks) && m.n > 0.0f) ? MAT_SPEC : 0u);                    // HasSpecularity :93
    return m;
}
inline PlaneRec make_plane(const float* f) {                  // 20 floats center, normal, Material, isTiled (ignored, :289)
    PlaneRec p;
    f3 c = mk3(f[0], f[1], f[2]);
    p.n = mk3(f[3], f[4], f[5]);
    p.cn = dot3(c, p.n);                                                           // :594
    f3 e1 = normalize3(cross3(p.n, mk3(1.0f, 0.0f, 0.0f)));                        // :760
    if (e1.x == 0.0f && e1.y == 0.0f && e1.z == 0.0f)                              // :761 (cannot fire: 0*inf = NaN)
        e1 = normalize3(cross3(p.n, mk3(0, 0, 1)));                                // :762
    p.e1 = e1;
    p.e2 = normalize3(cross3(p.n, e1));                                            // :765
    p.pad0 = p.pad1 = 0;
    p.m = make_mat(f + 6);
    return p;
}
inline LightRec make_light(const float* f) {
    LightRec l;
    l.p = mk3(f[0], f[1], f[2]); l.intensity = f[3];
    l.a = dot3(l.p, l.p);                                                          // :617 with direction = light.position (:574)
    l.a2 = 2 * l.a;                                                                // :624
    l.a4 = 4 * l.a;                                                                // :621
    l.pad = 0;
    return l;
}

}  // namespace rtb
