// rt_scene.cuh — scene policies (where the SoA fp32 records live) for rt_trace.cuh.
//   TinyScene   : whole scene by value in the kernel parameter block (constant bank; uniform broadcast reads).
//                 The reference's default scene (3 spheres, 1 plane, 2 lights, RayTracer.cs:441-465) is 0.5 KB.
//   GlobalScene : records in global memory (L1/L2-resident), any size; base of the shared-memory staged and LBVH paths.
#pragma once
#include "rt_trace.cuh"

namespace rtb {

enum { TINY_MAX_SPHERES = 16, TINY_MAX_PLANES = 4, TINY_MAX_LIGHTS = 8 };

struct TinySceneData {
    int ns, np, nl, pad;
    f3 amb; float pad1;
    f4 sgeom[TINY_MAX_SPHERES];
    MatRec smat[TINY_MAX_SPHERES];
    PlaneRec planes[TINY_MAX_PLANES];
    LightRec lights[TINY_MAX_LIGHTS];
};

struct TinyScene {
    const TinySceneData& s;
    RT_HD explicit TinyScene(const TinySceneData& d) : s(d) {}
    RT_HD int n_spheres() const { return s.ns; }
    RT_HD int n_planes() const { return s.np; }
    RT_HD int n_lights() const { return s.nl; }
    RT_HD f3 ambient() const { return s.amb; }
    RT_HD f4 sphere_geom(int i) const { return s.sgeom[i]; }
    RT_HD const MatRec& sphere_mat(int i) const { return s.smat[i]; }
    RT_HD const PlaneRec& plane(int i) const { return s.planes[i]; }
    RT_HD const LightRec& light(int i) const { return s.lights[i]; }
    template <class DBG> RT_HD void nearest_primary(f3 o, f3 d, float a2, float a4, int* sel, float* t, DBG& dbg) const {
        brute_nearest_primary(*this, o, d, a2, a4, sel, t, dbg);
    }
    template <class DBG> RT_HD void nearest_secondary(f3 o, f3 d, float a2, float a4, int* sel, float* t, DBG& dbg) const {
        brute_nearest_secondary(*this, o, d, a2, a4, sel, t, dbg);
    }
    template <class DBG> RT_HD bool shadow_any(f3 hit, const LightRec& l, DBG& dbg) const {
        return brute_shadow_any(*this, hit, l, dbg);
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

struct GlobalScene {
    const GlobalSceneData& s;
    RT_HD explicit GlobalScene(const GlobalSceneData& d) : s(d) {}
    RT_HD int n_spheres() const { return s.ns; }
    RT_HD int n_planes() const { return s.np; }
    RT_HD int n_lights() const { return s.nl; }
    RT_HD f3 ambient() const { return s.amb; }
    RT_HD f4 sphere_geom(int i) const { return load_f4(s.sgeom + i); }
    RT_HD const MatRec& sphere_mat(int i) const { return s.smat[i]; }
    RT_HD const PlaneRec& plane(int i) const { return s.planes[i]; }
    RT_HD const LightRec& light(int i) const { return s.lights[i]; }
    template <class DBG> RT_HD void nearest_primary(f3 o, f3 d, float a2, float a4, int* sel, float* t, DBG& dbg) const {
        brute_nearest_primary(*this, o, d, a2, a4, sel, t, dbg);
    }
    template <class DBG> RT_HD void nearest_secondary(f3 o, f3 d, float a2, float a4, int* sel, float* t, DBG& dbg) const {
        brute_nearest_secondary(*this, o, d, a2, a4, sel, t, dbg);
    }
    template <class DBG> RT_HD bool shadow_any(f3 hit, const LightRec& l, DBG& dbg) const {
        return brute_shadow_any(*this, hit, l, dbg);
    }
};

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
