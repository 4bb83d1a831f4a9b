// rt_trace.cuh — the per-sample trace of RayTracer.cs restructured for a GPU thread:
//   descent : geometry only — resolve the nearest hit, reflect, repeat (one ray chain)
//   unwind  : shade the recorded hits back-to-front, accumulating colour in the reference's exact order
// The reference shades every intersected primitive and then keeps the nearest (RayTracer.cs:975-993, :792-825);
// all its trace functions are pure, so select-then-shade returns the same bits (oracle test
// test_faithful_equals_nearest).  Every arithmetic expression below keeps the reference's operation order; the
// only liberties taken are hoists of loop-invariant subexpressions, merged call sites (one sphere loop for the
// primary and secondary folds, one light loop for spheres and planes — this keeps the kernel inside the
// instruction cache) and early-outs proven equivalent in DESIGN.md.
//
// Templated on a scene policy SC that supplies the records and the two sphere queries, so the brute-force loops
// and the LBVH traversal share all shading code:
//   int  n_spheres()/n_planes()/n_lights();  f3 ambient();
//   f4   sphere_geom(i)  -> (cx, cy, cz, r^2);      MatRec sphere_mat(i) / uint32_t sphere_flags(i)
//   PlaneRec access: plane_n(i) (normal + Dot(center,normal)), plane_e1/e2(i), plane_mat(i), plane_flags(i)
//   LightRec light(i)
//   void nearest   (o, dir, a2, a4, off, &sel, &d, dbg)   RayTracer.cs:975-981 (off = 0) / :792-808 (off = 0.01f)
//   bool shadow_any(li, hit, lp, a2, a4, dbg)              RayTracer.cs:573-582 (li = light index, -1 for free queries)
#pragma once
#include "rt_math.cuh"

namespace rtb {

enum { MAT_MIRROR = 1u, MAT_DIFFUSE = 2u, MAT_SPEC = 4u };
// per-pixel proof bits computed from the host's frame gates (rt_gate.cuh); 0 = nothing proven, trace everything
enum : uint32_t { RT_GATE_SPHERES = 1u, RT_GATE_MIRROR = 2u, RT_GATE_SHADOW_SHIFT = 2u, RT_GATE_MAX_LIGHTS = 4u };

struct MatRec {            // 16 words; Material RayTracer.cs:60-93
    f3 kd; float n;        // diffuseColor, specularity
    f3 ka; uint32_t flags; // ambientColor, IsMirror/IsDiffuse/HasSpecularity (:85-93) evaluated at upload
    f3 ks; float pad0;     // specularColor
    f3 km; float pad1;     // mirrorColor
};
struct PlaneRec {          // Plane RayTracer.cs:260-303 + upload-time invariants
    f3 n; float cn;        // normal (as given, not normalised), Dot(center, normal) (:594)
    f3 e1; float pad0;     // checkerboard basis (:760-765), depends on the plane only
    f3 e2; float pad1;
    MatRec m;
};
struct LightRec {          // Light RayTracer.cs:236-255 + invariants of the shadow ray whose DIRECTION is `position` (:574)
    f3 p; float intensity;
    float a, a2, a4, pad;  // Dot(p,p), 2*a, 4*a  (:617, :624, :621)
};
// Position and intensity of lights 2k and 2k+1 side by side: 64-bit constant-bank operands of the packed two-light Phong pass
// (shade_light_pair).
struct LightPair { float px[2], py[2], pz[2], intensity[2]; };
struct CamRec { f3 pos, right, up, fwd, view; };

struct HitRec { f3 o; f3 dir; float d; int prim; };   // prim >= 0: sphere index; prim < 0: ~plane index

#ifndef RT_INF
#define RT_INF (bits2f(0x7f800000u))
#endif

// ---------------------------------------------------------------------------------------------------------
// Debug policies
// ---------------------------------------------------------------------------------------------------------
struct NoDbg {
    static constexpr bool enabled = false;
    static constexpr bool count_tests = false;   // true: per-sphere test counters wanted -> the scalar (unpacked) sphere loops
    RT_HD void sphere_test(bool) {}
    RT_HD void plane_test() {}
    RT_HD void ray(uint32_t, uint32_t, uint32_t, float) {}
    RT_HD void shadow(uint32_t, uint32_t, bool) {}
    RT_HD void shaded(bool) {}
    RT_HD void spec() {}
    RT_HD void primary_aov(int, float) {}
    RT_HD void node_visit(int) {}
    RT_HD void fallback() {}
    template <class SC> RT_HD void ray_geom(const SC&, uint32_t, uint32_t, uint32_t, f3, f3, float) {}
    template <class SC> RT_HD void shadow_geom(const SC&, uint32_t, uint32_t, f3, f3, float, float, bool) {}
};
struct FullDbg {
    static constexpr bool enabled = true;
    static constexpr bool count_tests = true;
    uint32_t hash = 0;
    uint32_t primary = 0, n_shadow = 0, secondary = 0, sphere_tests = 0, sphere_disc_pos = 0, plane_tests = 0;
    uint32_t shade_diffuse = 0, shade_specular = 0, shade_mirror = 0, shaded_hits = 0;
    uint32_t node_visits[3] = {0, 0, 0}, fallbacks = 0;    // LBVH: node (box-pair) visits by ray kind, brute-force fallbacks
    int aov_id = -1; float aov_t = 0.0f; bool aov_set = false;
    RT_HD void sphere_test(bool disc_pos) { sphere_tests++; if (disc_pos) sphere_disc_pos++; }
    RT_HD void plane_test() { plane_tests++; }
    RT_HD void ray(uint32_t level, uint32_t kind, uint32_t code, float d) {
        hash += event_hash(level, kind, code, f2bits(d));
        if (kind == 1) primary++; else secondary++;
    }
    RT_HD void shadow(uint32_t level, uint32_t li, bool occluded) {
        hash += event_hash(level, 3, li, occluded ? 1u : 0u); n_shadow++; shade_diffuse++;
    }
    RT_HD void shaded(bool mirror) { shaded_hits++; if (mirror) shade_mirror++; }
    RT_HD void spec() { shade_specular++; }
    RT_HD void primary_aov(int id, float t) { if (!aov_set) { aov_id = id; aov_t = t; aov_set = true; } }
    RT_HD void node_visit(int kind) { node_visits[kind]++; }
    RT_HD void fallback() { fallbacks++; }
    template <class SC> RT_HD void ray_geom(const SC&, uint32_t, uint32_t, uint32_t, f3, f3, float) {}
    template <class SC> RT_HD void shadow_geom(const SC&, uint32_t, uint32_t, f3, f3, float, float, bool) {}
};

// Events only (chain hash, ray counters, primary AOV) — no per-test counters, so the code it instruments is the PRODUCTION code:
// packed-fp32 sphere pairs, exact-count unrolled loops, frame gates, fast pixel division. The shipped kernel instantiated with this
// policy (k_debug_tiny_prod, rt_set_option(RT_OPT_DEBUG_SHIPPED)) certifies hit ids / t bits / shadow results of the gated path
// against the oracle's chain hash; work a gate skips is reported as the event the reference would have produced (a reflection
// ray that hits nothing, an unoccluded shadow ray, a primary ray that hits nothing).
struct ProdDbg {
    static constexpr bool enabled = true;
    static constexpr bool count_tests = false;
    uint32_t hash = 0, primary = 0, n_shadow = 0, secondary = 0;
    int aov_id = -1; float aov_t = 0.0f; bool aov_set = false;
    RT_HD void sphere_test(bool) {}
    RT_HD void plane_test() {}
    RT_HD void ray(uint32_t level, uint32_t kind, uint32_t code, float d) {
        hash += event_hash(level, kind, code, f2bits(d));
        if (kind == 1) primary++; else secondary++;
    }
    RT_HD void shadow(uint32_t level, uint32_t li, bool occluded) { hash += event_hash(level, 3, li, occluded ? 1u : 0u); n_shadow++; }
    RT_HD void shaded(bool) {}
    RT_HD void spec() {}
    RT_HD void primary_aov(int id, float t) { if (!aov_set) { aov_id = id; aov_t = t; aov_set = true; } }
    RT_HD void node_visit(int) {}
    RT_HD void fallback() {}
    template <class SC> RT_HD void ray_geom(const SC&, uint32_t, uint32_t, uint32_t, f3, f3, float) {}
    template <class SC> RT_HD void shadow_geom(const SC&, uint32_t, uint32_t, f3, f3, float, float, bool) {}
};

// ---------------------------------------------------------------------------------------------------------
// Sphere test — IntersectsSphere RayTracer.cs:613-642.
// Given per-ray a2 = 2*Dot(d,d), a4 = 4*Dot(d,d) (hoisted; `4 * a * c` associates as (4*a)*c).
// Returns true and *t = distance iff the reference reports a collision with this epsilon.
// Equivalences used (DESIGN.md §sphere test):  with s = sqrt(D) >= 0 and a2 >= 0,  t1 = (-b-s)/a2 <= t2 = (-b+s)/a2
// (monotone rounding) or t1 is NaN, hence  distanceEps > 0  <=>  t1 - eps > 0,  and then distance = min(t1,t2) = t1;
// and t1 > 0 requires b < 0, so b >= 0 (or NaN) is a miss without a square root or a division.
// ---------------------------------------------------------------------------------------------------------
template <class DBG>
RT_HD bool sphere_hit(f3 oc, f3 dir, float r2, float a2, float a4, float eps, float* t, DBG& dbg) {
    float b = 2 * dot3(oc, dir);                                  // :618
    float c = dot3(oc, oc) - r2;                                  // :619
    float D = b * b - a4 * c;                                     // :621
    dbg.sphere_test(D >= 0);
    if (!(b < 0 && D >= 0)) return false;                         // :622 (+ the b >= 0 early-out)
    float s = sqrtf(D);                                           // :623  (float)Math.Sqrt((double)D) == sqrtf(D)
    float t1 = (-b - s) / a2;                                     // :627
    if (!(t1 - eps > 0)) return false;                            // :629-635
    *t = t1;                                                      // :632
    return true;
}

// The slow half of sphere_hit once b and D are known and b < 0 && D >= 0.
RT_HD bool sphere_root(float b, float D, float a2, float eps, float* t) {
    float s = sqrtf(D);                                           // :623
    float t1 = (-b - s) / a2;                                     // :627
    if (!(t1 - eps > 0)) return false;                            // :629-635
    *t = t1;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Ray log (rt_ray_log): the GPU-side counterpart of the reference's DEBUG_ENABLE TracedRay log, RayTracer.cs:424-435 — one
// record per RAY of the nearest-first chain (the reference logs one per intersection TEST, :601, :639, :801, and then draws a
// random sample of them, :914-933). Layout == rt_ray_record (include/rtb200.h).
//   primary / secondary: the selected hit (prim = sphere index, n_spheres + plane index or -1; d = its distance, 0 on a miss)
//   shadow: the NEAREST occluding sphere (lowest index on ties) by a brute-force pass with the shadow epsilon, or -1 / d = 0
//   hit = o + dir * d in every case (TracedRay.hitPoint).
// Record order within a pixel: the rays of the descent, then the shadow rays deepest level first — the order in which the
// reference's recursion creates them (the mirror term :857/:746 is evaluated before the light loop :863/:751).
// ---------------------------------------------------------------------------------------------------------
struct RayRec {
    float o[3], dir[3], hit[3], d;
    int32_t prim;
    uint32_t kind, pixel, level, light, reserved;
};
static_assert(sizeof(RayRec) == 64, "RayRec layout");

struct LogDbg {
    static constexpr bool enabled = true;
    static constexpr bool count_tests = true;
    RayRec* out = nullptr; uint32_t cap = 0, n = 0, pixel = 0;
    RT_HD void sphere_test(bool) {}
    RT_HD void plane_test() {}
    RT_HD void ray(uint32_t, uint32_t, uint32_t, float) {}
    RT_HD void shadow(uint32_t, uint32_t, bool) {}
    RT_HD void shaded(bool) {}
    RT_HD void spec() {}
    RT_HD void primary_aov(int, float) {}
    RT_HD void node_visit(int) {}
    RT_HD void fallback() {}
    RT_HD void push(uint32_t kind, uint32_t level, uint32_t light, int32_t prim, f3 o, f3 dir, float d) {
        if (n < cap) {
            RayRec& r = out[n];
            f3 h = add3(o, mulf3(dir, d));                        // :601 / :639 / :801
            r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.dir[0] = dir.x; r.dir[1] = dir.y; r.dir[2] = dir.z;
            r.hit[0] = h.x; r.hit[1] = h.y; r.hit[2] = h.z; r.d = d; r.prim = prim;
            r.kind = kind; r.pixel = pixel; r.level = level; r.light = light; r.reserved = 0;
        }
        n++;                                                      // keeps counting past cap: the caller sees the overflow
    }
    template <class SC> RT_HD void ray_geom(const SC&, uint32_t level, uint32_t kind, uint32_t code, f3 o, f3 dir, float d) {
        push(kind, level, 0, (int32_t)code, o, dir, d);
    }
    template <class SC> RT_HD void shadow_geom(const SC& sc, uint32_t level, uint32_t li, f3 hit, f3 lp, float a2, float a4, bool) {
        NoDbg nd;
        int sel = -1; float best = RT_INF;
        const int ns = sc.n_spheres();
        for (int i = 0; i < ns; i++) {                            // :577, keeping the nearest collision instead of a flag
            f4 g = sc.sphere_geom(i);
            float t;
            if (sphere_hit(sub3(hit, mk3(g.x, g.y, g.z)), lp, g.w, a2, a4, 0.001f, &t, nd) && t < best) { best = t; sel = i; }
        }
        push(2, level, li, sel, hit, lp, sel >= 0 ? best : 0.0f);
    }
};

// b and the discriminant of spheres i0 and i0+1 in one pass of packed-fp32 instructions (sm_100 FADD2 / FMUL2: two IEEE
// round-to-nearest fp32 operations per lane per instruction — bit-identical to the scalar sequence, half the issue slots; the
// kernel is issue-bound with the FMA pipe ~35 % busy). Operation order per element is exactly :614-621.
#if defined(RT_HAVE_F32X2)
template <class PAIR>
RT_D void sphere_pair_bd(const PAIR& p, float2 ox, float2 oy, float2 oz, float2 dx, float2 dy, float2 dz,
                                               float2 na4, float* b_out, float* D_out) {
    const float2 ocx = rt_add2(ox, make_float2(p.ncx[0], p.ncx[1]));                      // :614  o - c
    const float2 ocy = rt_add2(oy, make_float2(p.ncy[0], p.ncy[1]));
    const float2 ocz = rt_add2(oz, make_float2(p.ncz[0], p.ncz[1]));
    const float2 px = rt_mul2(ocx, dx), py = rt_mul2(ocy, dy), pz = rt_mul2(ocz, dz);
    const float2 qx = rt_mul2(ocx, ocx), qy = rt_mul2(ocy, ocy), qz = rt_mul2(ocz, ocz);
    const float2 bd = make_float2((px.x + py.x) + pz.x, (px.y + py.y) + pz.y);            // Dot: scalar adds (see above)
    const float2 cd = make_float2((qx.x + qy.x) + qz.x, (qx.y + qy.y) + qz.y);
    const float2 b = rt_mul2(make_float2(2.0f, 2.0f), bd);                                // :618
    const float2 c = rt_add2(cd, make_float2(p.nr2[0], p.nr2[1]));                        // :619  dot - r^2
    const float2 bb = rt_mul2(b, b), ac = rt_mul2(na4, c);
    const float2 D = make_float2(bb.x + ac.x, bb.y + ac.y);                               // :621  b*b - (4a)*c
    b_out[0] = b.x; b_out[1] = b.y; D_out[0] = D.x; D_out[1] = D.y;
}
#endif

// The reference's loops over every sphere, in array order.  One loop serves both folds:
//   primary   (:977)  `d > 0 && nearest > d`                    == key > 0 && key < closest with key = d - 0
//   secondary (:804)  `d - 0.01f > 0 && d - 0.01f < closest`    == the same with key = d - 0.01f
// (d - 0.0f == d bit for bit; the stored distance is the un-offset d in both: the secondary fold is order dependent).
// NS >= 0: compile-time sphere count — fully unrolled, records addressed statically, and the b / discriminant of ALL
// spheres are evaluated branch-free first so that the common all-miss case costs one branch instead of NS.
template <int NS, class SC, class DBG>
RT_HD void brute_nearest(const SC& sc, f3 o, f3 dir, float a2, float a4, float off, int* sel, float* dsel, DBG& dbg) {
    int best = -1; float closest = RT_INF;
    if constexpr (NS > 0) {
        // `gate` = max_i min(D_i, -b_i) >= 0 is a cheap SUPERSET of "some sphere has b < 0 && D >= 0" (two FMNMX per sphere;
        // fminf drops a NaN, b == 0 passes): it only decides whether the exact per-sphere tests below run at all.
        float bs[NS + 1], Ds[NS + 1]; float gate = -1.0f;
#if defined(RT_HAVE_F32X2)
        if constexpr (SC::has_pairs && NS >= 2 && !DBG::count_tests) {
            const float2 ox = make_float2(o.x, o.x), oy = make_float2(o.y, o.y), oz = make_float2(o.z, o.z);
            const float2 dx = make_float2(dir.x, dir.x), dy = make_float2(dir.y, dir.y), dz = make_float2(dir.z, dir.z);
            const float2 na4 = make_float2(-a4, -a4);
#pragma unroll
            for (int k = 0; k < NS / 2; k++) sphere_pair_bd(sc.sphere_pair(k), ox, oy, oz, dx, dy, dz, na4, bs + 2 * k, Ds + 2 * k);
            if (NS & 1) {                                         // odd one out: scalar
                f4 g = sc.sphere_geom(NS - 1);
                f3 oc = sub3(o, mk3(g.x, g.y, g.z));
                bs[NS - 1] = 2 * dot3(oc, dir);
                Ds[NS - 1] = bs[NS - 1] * bs[NS - 1] - a4 * (dot3(oc, oc) - g.w);
            }
#pragma unroll
            for (int i = 0; i < NS; i++) gate = fmaxf(gate, fminf(Ds[i], -bs[i]));
        } else
#endif
        {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                f4 g = sc.sphere_geom(i);
                f3 oc = sub3(o, mk3(g.x, g.y, g.z));              // :614
                bs[i] = 2 * dot3(oc, dir);                        // :618
                float c = dot3(oc, oc) - g.w;                     // :619
                Ds[i] = bs[i] * bs[i] - a4 * c;                   // :621
                dbg.sphere_test(Ds[i] >= 0);
                gate = fmaxf(gate, fminf(Ds[i], -bs[i]));
            }
        }
        if (gate >= 0) {
#pragma unroll
            for (int i = 0; i < NS; i++) {                        // :975 / :792, array order
                float t;
                if (bs[i] < 0 && Ds[i] >= 0 && sphere_root(bs[i], Ds[i], a2, 0.0f, &t)) {
                    float key = t - off;
                    if (key > 0 && key < closest) { closest = t; best = i; }
                }
            }
        }
    } else {
        const int ns = NS >= 0 ? NS : sc.n_spheres();
        for (int i = 0; i < ns; i++) {                            // :975 / :792
            f4 g = sc.sphere_geom(i);
            float t;
            if (sphere_hit(sub3(o, mk3(g.x, g.y, g.z)), dir, g.w, a2, a4, 0.0f, &t, dbg)) {
                float key = t - off;
                if (key > 0 && key < closest) { closest = t; best = i; }
            }
        }
    }
    *sel = best; *dsel = closest;
}
template <int NS, class SC, class DBG>
RT_HD bool brute_shadow_any(const SC& sc, f3 hit, f3 lp, float a2, float a4, DBG& dbg) {
    bool occluded = false;
    if constexpr (NS > 0) {
        float bs[NS + 1], Ds[NS + 1]; float gate = -1.0f;
#if defined(RT_HAVE_F32X2)
        if constexpr (SC::has_pairs && NS >= 2 && !DBG::count_tests) {
            const float2 ox = make_float2(hit.x, hit.x), oy = make_float2(hit.y, hit.y), oz = make_float2(hit.z, hit.z);
            const float2 dx = make_float2(lp.x, lp.x), dy = make_float2(lp.y, lp.y), dz = make_float2(lp.z, lp.z);
            const float2 na4 = make_float2(-a4, -a4);
#pragma unroll
            for (int k = 0; k < NS / 2; k++) sphere_pair_bd(sc.sphere_pair(k), ox, oy, oz, dx, dy, dz, na4, bs + 2 * k, Ds + 2 * k);
            if (NS & 1) {                                         // odd one out: scalar
                f4 g = sc.sphere_geom(NS - 1);
                f3 oc = sub3(hit, mk3(g.x, g.y, g.z));
                bs[NS - 1] = 2 * dot3(oc, lp);
                Ds[NS - 1] = bs[NS - 1] * bs[NS - 1] - a4 * (dot3(oc, oc) - g.w);
            }
#pragma unroll
            for (int i = 0; i < NS; i++) gate = fmaxf(gate, fminf(Ds[i], -bs[i]));
        } else
#endif
        {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                f4 g = sc.sphere_geom(i);
                f3 oc = sub3(hit, mk3(g.x, g.y, g.z));
                bs[i] = 2 * dot3(oc, lp);
                float c = dot3(oc, oc) - g.w;
                Ds[i] = bs[i] * bs[i] - a4 * c;
                dbg.sphere_test(Ds[i] >= 0);
                gate = fmaxf(gate, fminf(Ds[i], -bs[i]));
            }
        }
        if (gate >= 0) {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                float t;
                if (bs[i] < 0 && Ds[i] >= 0 && sphere_root(bs[i], Ds[i], a2, 0.001f, &t)) occluded = true;   // :578
            }
        }
    } else {
        const int ns = NS >= 0 ? NS : sc.n_spheres();
        for (int i = 0; i < ns; i++) {                            // :577
            f4 g = sc.sphere_geom(i);
            float t;
            if (sphere_hit(sub3(hit, mk3(g.x, g.y, g.z)), lp, g.w, a2, a4, 0.001f, &t, dbg))   // :578
                occluded = true;                                  // boolean OR over all spheres (no early-out in the reference)
        }
    }
    return occluded;
}

// (float)Math.Pow((double)base, (double)n) — RayTracer.cs:691.  pow(x,1) == x exactly; pow(x,.5) rounds as sqrtf(x)
// (SURVEY A.12); everything else takes the f64 pow.
RT_HD float spec_pow(float base, float n) {
    // base comes from cs_max0: +0 whenever the highlight faces away. pow(+0, n) = +0 for every n > 0 (HasSpecularity, :93) —
    // answering it here also keeps sqrtf(0) out of its slow path (0.6 % of all instructions on the default scene).
    if (base == 0.0f) return 0.0f;
    if (n == 1.0f) return base;
    if (n == 0.5f) return sqrtf(base);
#if defined(__CUDA_ARCH__)
    // The f64 pow is ~1 K instructions with a long side-effect-free prologue that ptxas otherwise hoists above the two
    // tests and executes for every light sample.  Routing the operand through a volatile asm pins it inside this branch.
    float pinned;
    asm volatile("mov.f32 %0, %1;" : "=f"(pinned) : "f"(base));
    return (float)pow((double)pinned, (double)n);
#else
    return (float)pow((double)base, (double)n);
#endif
}

#if defined(RT_HAVE_F32X2)
// ShapePhongShading :665-695 and the light term :866-869 / :754-775 for lights 2k and 2k+1 in ONE pass of packed fp32: half .x
// of every pair is light 2k, half .y light 2k+1; hit point, normal, view vector and material are scalar operands broadcast to
// both halves. Per half the operations and their order are exactly the scalar loop's (shade_hit below). As in sphere_pair_bd,
// only instructions that cannot be contracted are packed: differences of inputs, products, the two normalisations (rt_inv_len2);
// every sum that consumes a product — the dot products, L - N*(2 L.N), ph + ks*pow — stays a scalar FADD on the register halves,
// because ptxas fuses a packed multiply into a following packed add even under -fmad=false. ~130 instead of ~190 instructions
// per shaded hit with two lights. I0 / I1: the lights' intensities after the shadow test (:581).
template <class DBG>
RT_D void shade_light_pair(const LightPair& lp, const MatRec& m, f3 hit, f3 N, f3 V, float att, float tile,
                                                 bool is_plane, float I0, float I1, f3* tA, f3* tB, DBG& dbg) {
    float2 Lx = rt_sub2(make_float2(lp.px[0], lp.px[1]), rt_splat2(hit.x));                   // :667  light.position - hit
    float2 Ly = rt_sub2(make_float2(lp.py[0], lp.py[1]), rt_splat2(hit.y));
    float2 Lz = rt_sub2(make_float2(lp.pz[0], lp.pz[1]), rt_splat2(hit.z));
    {
        const float2 xx = rt_mul2(Lx, Lx), yy = rt_mul2(Ly, Ly), zz = rt_mul2(Lz, Lz);
        const float2 sc = rt_inv_len2(make_float2((xx.x + yy.x) + zz.x, (xx.y + yy.y) + zz.y));   // Normalize: 1f / Length
        Lx = rt_mul2(Lx, sc); Ly = rt_mul2(Ly, sc); Lz = rt_mul2(Lz, sc);
    }
    const float2 nlx = rt_mul2(rt_splat2(N.x), Lx), nly = rt_mul2(rt_splat2(N.y), Ly), nlz = rt_mul2(rt_splat2(N.z), Lz);
    const float2 ndl = make_float2((nlx.x + nly.x) + nlz.x, (nlx.y + nly.y) + nlz.y);         // Dot(N, L) == Dot(L, N) bit for bit
    const float2 diff = make_float2(cs_max0(ndl.x), cs_max0(ndl.y));                          // :678
    float2 phx = rt_mul2(rt_splat2(m.kd.x), diff), phy = rt_mul2(rt_splat2(m.kd.y), diff), phz = rt_mul2(rt_splat2(m.kd.z), diff);   // :672-678
    if (m.flags & MAT_SPEC) {                                                                 // :682
        const float2 two = rt_mul2(rt_splat2(2.0f), ndl);                                     // :684  2 * Dot(L, N)
        const float2 qx = rt_mul2(rt_splat2(N.x), two), qy = rt_mul2(rt_splat2(N.y), two), qz = rt_mul2(rt_splat2(N.z), two);
        float2 rx = make_float2(Lx.x - qx.x, Lx.y - qx.y), ry = make_float2(Ly.x - qy.x, Ly.y - qy.y), rz = make_float2(Lz.x - qz.x, Lz.y - qz.y);   // :683
        {
            const float2 xx = rt_mul2(rx, rx), yy = rt_mul2(ry, ry), zz = rt_mul2(rz, rz);
            const float2 sc = rt_inv_len2(make_float2((xx.x + yy.x) + zz.x, (xx.y + yy.y) + zz.y));   // :687 Normalize
            rx = rt_mul2(rx, sc); ry = rt_mul2(ry, sc); rz = rt_mul2(rz, sc);
        }
        const float2 vx = rt_mul2(rt_splat2(V.x), rx), vy = rt_mul2(rt_splat2(V.y), ry), vz = rt_mul2(rt_splat2(V.z), rz);   // :685-688 Dot(V, Rv)
        const float2 pw = make_float2(spec_pow(cs_max0((vx.x + vy.x) + vz.x), m.n), spec_pow(cs_max0((vx.y + vy.y) + vz.y), m.n));   // :691
        const float2 kx = rt_mul2(rt_splat2(m.ks.x), pw), ky = rt_mul2(rt_splat2(m.ks.y), pw), kz = rt_mul2(rt_splat2(m.ks.z), pw);
        phx = make_float2(phx.x + kx.x, phx.y + kx.y); phy = make_float2(phy.x + ky.x, phy.y + ky.y); phz = make_float2(phz.x + kz.x, phz.y + kz.y);   // :690-694
        dbg.spec(); dbg.spec();
    }
    const float2 ia = rt_mul2(make_float2(I0, I1), rt_splat2(att));                           // (I,I,I) * att
    const float2 tl = rt_splat2(tile);
    const float2 tx = rt_mul2(rt_mul2(ia, phx), tl), ty = rt_mul2(rt_mul2(ia, phy), tl), tz = rt_mul2(rt_mul2(ia, phz), tl);   // :868-869 / :774
    if (is_plane) {                                                                           // :775 (.Max(0))
        *tA = mk3(cs_max0(tx.x), cs_max0(ty.x), cs_max0(tz.x)); *tB = mk3(cs_max0(tx.y), cs_max0(ty.y), cs_max0(tz.y));
    } else {
        *tA = mk3(tx.x, ty.x, tz.x); *tB = mk3(tx.y, ty.y, tz.y);
    }
}
#endif

// ---------------------------------------------------------------------------------------------------------
// Colour of one recorded hit given the colour Cin seen by its reflection ray:
// TraceSphere :846-875 / TracePlane :736-779 with ShapePhongShading :665-695 inlined into ONE light loop.
//   sphere: col += ((I,I,I) * att) * phong                      att = (1/d)*d                       (:866-869)
//   plane : col += Max((((I,I,I) * att) * phong) * tile, 0)     att = (float)(1/Math.Pow(d,2))      (:754-775)
// `x * tile` with tile = (1,1,1) is exact, so spheres run the plane expression with tile = 1 and skip only the Max.
// ---------------------------------------------------------------------------------------------------------
template <class SC, class DBG>
RT_HD f3 shade_hit(const SC& sc, const HitRec& h, f3 Cin, uint32_t level, DBG& dbg, uint32_t gates = 0u) {
    const f3 hit = add3(h.o, mulf3(h.dir, h.d));                  // :846 / :736
    const bool is_plane = h.prim < 0;
    MatRec m; f3 N; float att, tile;
    if (!is_plane) {
        m = sc.sphere_mat(h.prim);
        f4 g = sc.sphere_geom(h.prim);
        N = normalize3(sub3(hit, mk3(g.x, g.y, g.z)));            // :706
        att = 1 / h.d * h.d;                                      // :866  ((1/d)*d)
        tile = 1.0f;
    } else {
        const int pi = ~h.prim;
        m = sc.plane_mat(pi);
        f4 pn = sc.plane_n(pi);
        N = mk3(pn.x, pn.y, pn.z);                                // :653 plane.normal as given
        double dd = (double)h.d;
        att = (float)(1.0 / (dd * dd));                           // :754  Math.Pow(d,2) == d*d exactly in f64
        float u = dot3(sc.plane_e1(pi), hit);                     // :766
        float v = dot3(sc.plane_e2(pi), hit);                     // :767
        tile = (float)(int32_t)(((uint32_t)cs_f2i(u) + (uint32_t)cs_f2i(v)) & 1u);   // :769-770  ((int)u + (int)v) & 1
    }
    dbg.shaded((m.flags & MAT_MIRROR) != 0);
    f3 col = mk3(0, 0, 0);
    if (m.flags & MAT_MIRROR) col = mulv3(Cin, m.km);             // :857-858 / :746-747   (0 + x == x)
    if (m.flags & MAT_DIFFUSE) {                                  // :862 / :750
        const f3 V = normalize3(h.dir);                           // :668
        const int nl = sc.n_lights();
#if defined(RT_HAVE_F32X2)
        // Exact even light counts (the reference scene has two): lights are shaded two at a time (shade_light_pair). The shadow
        // tests stay scalar and come first; the two terms are added to the colour in light order, as the loop below does.
        if constexpr (SC::static_lights >= 2 && (SC::static_lights & 1) == 0 && !DBG::count_tests) {
#pragma unroll
            for (int li = 0; li < SC::static_lights; li += 2) {
                float I[2];
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const LightRec l = sc.light(li + k);
                    const bool proven_clear = li + k < (int)RT_GATE_MAX_LIGHTS && ((gates >> (RT_GATE_SHADOW_SHIFT + li + k)) & 1u) && level == 0 && is_plane;
                    const bool occ = proven_clear ? false : sc.shadow_any(li + k, hit, l.p, l.a2, l.a4, dbg);   // :864 / :752
                    dbg.shadow(level, (uint32_t)(li + k), occ);
                    dbg.shadow_geom(sc, level, (uint32_t)(li + k), hit, l.p, l.a2, l.a4, occ);
                    I[k] = occ ? 0.0f : l.intensity;              // :581
                }
                f3 tA, tB;
                shade_light_pair(sc.light_pair(li >> 1), m, hit, N, V, att, tile, is_plane, I[0], I[1], &tA, &tB, dbg);
                col = add3(add3(col, tA), tB);
            }
        } else
#endif
#pragma unroll
        for (int li = 0; li < nl; li++) {                         // :863 / :751
            const LightRec l = sc.light(li);
            // gates bit 2 + li: the shadow ray of light li from the PRIMARY hit on the (single) plane meets no sphere (rt_gate.cuh)
            const bool proven_clear = li < (int)RT_GATE_MAX_LIGHTS && ((gates >> (RT_GATE_SHADOW_SHIFT + li)) & 1u) && level == 0 && is_plane;
            bool occ = proven_clear ? false : sc.shadow_any(li, hit, l.p, l.a2, l.a4, dbg);  // :864 / :752
            dbg.shadow(level, (uint32_t)li, occ);
            dbg.shadow_geom(sc, level, (uint32_t)li, hit, l.p, l.a2, l.a4, occ);
            float I = occ ? 0.0f : l.intensity;                   // :581
            // ShapePhongShading :665-695
            f3 L = normalize3(sub3(l.p, hit));                    // :667
            f3 ph = mulf3(m.kd, cs_max0(dot3(N, L)));       // :672-678 (IsDiffuse holds inside this loop)
            if (m.flags & MAT_SPEC) {                             // :682
                f3 rv = sub3(L, mulf3(N, 2 * dot3(L, N)));        // :683-684
                float s = dot3(V, normalize3(rv));                // :685-688
                float pw = spec_pow(cs_max0(s), m.n);       // :691
                ph = add3(ph, mulv3(m.ks, splat3(pw)));           // :690-694
                dbg.spec();
            }
            float ia = I * att;                                   // (I,I,I) * att
            f3 t = mulf3(mulv3(splat3(ia), ph), tile);            // :868-869 / :774
            if (is_plane) t = mk3(cs_max0(t.x), cs_max0(t.y), cs_max0(t.z));   // :775 (.Max(0))
            col = add3(col, t);
        }
    }
    return add3(col, mulv3(sc.ambient(), m.ka));                  // :873 / :778
}

// ---------------------------------------------------------------------------------------------------------
// One sample: TracePixel RayTracer.cs:962-1002 with (fx, fy) in place of (x, y).
// ---------------------------------------------------------------------------------------------------------
// FASTDIV (device, single-sample kernels, frame sides <= RT_FASTDIV_MAX): rw = 1/fw, rh = 1/fh correctly rounded on the host,
// the two divisions become rt_div_rcp — same bits (rt_math.cuh), a quarter of the instructions.
template <bool FASTDIV = false>
RT_HD void primary_ray(const CamRec& cam, float fx, float fy, float fw, float fh, float rw, float rh, f3* o, f3* dir) {
#if defined(__CUDA_ARCH__)
    float u = (FASTDIV ? rt_div_rcp(fx, fw, rw) : fx / fw) - 0.5f;                         // :964
    float v = (FASTDIV ? rt_div_rcp(fy, fh, rh) : fy / fh) - 0.5f;
#else
    (void)rw; (void)rh;
    float u = fx / fw - 0.5f;                                                              // :964
    float v = fy / fh - 0.5f;
#endif
    f3 local = mulv3(mk3(u, v, 1.0f), cam.view);                                           // :965
    f3 vp = add3(add3(add3(cam.pos, mulf3(cam.right, local.x)), mulf3(cam.up, local.y)), mulf3(cam.fwd, local.z));   // :967-969
    *o = cam.pos;
    *dir = normalize3(sub3(vp, cam.pos));                                                  // :971
}

// Numerator of IntersectPlane (:591-594) — also evaluated on the host for the primary rays' common origin (rt_gate.cuh).
RT_HD float plane_num(f3 o, f4 pn) { return -o.x * pn.x - o.y * pn.y - o.z * pn.z + pn.w; }

// The ray chain of one sample, resumable.  Continues the descent from the state (o, dir, bounce, top; stack[0..top) filled);
// when the chain ends it unwinds the whole stack, stores the colour in *C and returns true.  If defer_at >= 0 and the chain is
// about to trace the ray of level `defer_at` it returns false instead, leaving the state ready for a later call (used by the
// compacting kernel to hand deep mirror chains to fully populated warps).  `stack` must hold cap+1 records.
template <class SC, class DBG>
RT_HD bool trace_chain(const SC& sc, int cap, f3& o, f3& dir, int& bounce, int& top, HitRec* stack, int defer_at, f3* Cout, DBG& dbg,
                       uint32_t gates = 0u) {
    f3 C = mk3(0, 0, 0);
    const int np = sc.n_planes();      // compile-time constant in the exact-count kernels
    bool proven_no_sphere = false;     // debug policies only, see the mirror gate below
    for (;;) {
        if (top == defer_at) return false;
        float a = dot3(dir, dir);                                                          // :617
        float a2 = 2 * a;                                                                  // :624
        float a4 = 4 * a;                                                                  // :621
        // gates (rt_gate.cuh, bits RT_GATE_*): what the host PROVED about this pixel. Bit 0: its primary ray cannot be reported
        // as hitting any sphere.
        int sel_s = -1; float d_s = RT_INF;
        if (!(((gates & RT_GATE_SPHERES) && bounce == 0) || (DBG::enabled && proven_no_sphere)))
            sc.nearest(o, dir, a2, a4, bounce == 0 ? 0.0f : 0.01f, &sel_s, &d_s, dbg);     // :975-981 / :792-808
        int sel_p = -1; float d_p = RT_INF;
#pragma unroll
        for (int i = 0; i < np; i++) {                                                     // :985 / :812
            f4 pn = sc.plane_n(i);
            float num = plane_num(o, pn);                                                  // :591-594
            dbg.plane_test();
            // num == 0 (a ray leaving the plane it starts on: every floor reflection re-tests the floor) gives
            // t = 0/den = +-0 or NaN: never `> 0`. Skipping the division is exact and avoids the IEEE-divide slow path that a zero
            // numerator takes (2.4 % of all instructions on the default scene).
            if (num != 0.0f) {
                float t = num / dot3(dir, mk3(pn.x, pn.y, pn.z));                          // :595-596
                if (t > 0 && t < d_p) { d_p = t; sel_p = i; }                              // :598 + :987 / :819
            }
        }
        bool pick_s = d_s < d_p;                                                           // :993 / :825
        bool none = !pick_s && sel_p < 0;
        float d = pick_s ? d_s : (none ? 0.0f : d_p);
        if (DBG::enabled) {
            uint32_t code = pick_s ? (uint32_t)sel_s : (none ? 0xFFFFFFFFu : (uint32_t)(sc.n_spheres() + sel_p));
            dbg.ray((uint32_t)top, bounce == 0 ? 1u : 2u, code, d);
            dbg.ray_geom(sc, (uint32_t)top, bounce == 0 ? 0u : 1u, code, o, dir, d);
            if (bounce == 0) dbg.primary_aov((int)code, d);
        }
        if (none) break;                                                                   // nothing hit: black
        if (d - 0.01f <= 0) break;                                                         // :839 / :731 (distance kept, colour black)
        if (bounce > cap) { if (!pick_s) C = mk3(1, 1, 1); break; }                        // :843 black / :734 white
        HitRec& h = stack[top++];
        h.o = o; h.dir = dir; h.d = d; h.prim = pick_s ? sel_s : ~sel_p;
        uint32_t flags = pick_s ? sc.sphere_flags(sel_s) : sc.plane_flags(sel_p);
        if (!(flags & MAT_MIRROR)) break;                                                  // :850 / :739
        // Bit 1: if the primary ray hit the (single) plane, its reflection ray hits nothing — the mirror term is (0,0,0), C as is.
        if ((gates & RT_GATE_MIRROR) && bounce == 0 && !pick_s) {
            if (!DBG::enabled) break;
            // Debug policies report the ray the reference traces here. The gate proves it meets no sphere and can hit its own plane
            // only inside the 0.01 cut-off (:731) — which of the two decides the event, so the ray is followed against
            // the planes alone; it ends the chain at the `none` / `d - 0.01f <= 0` exits above (anything else would be an unsound
            // gate and shows up as a pixel that differs from the shipped kernel's).
            proven_no_sphere = true;
        }
        bounce++;                                                                          // :851 / :740
        f3 hit = add3(o, mulf3(dir, d));                                                   // :846 / :736
        f3 N;
        if (pick_s) { f4 g = sc.sphere_geom(sel_s); N = normalize3(sub3(hit, mk3(g.x, g.y, g.z))); }   // :854
        else { f4 pn = sc.plane_n(sel_p); N = mk3(pn.x, pn.y, pn.z); }                     // :743
        dir = sub3(dir, mulf3(N, 2 * dot3(dir, N)));                                       // :719
        o = hit;
    }
    while (top > 0) {
        --top;
        C = shade_hit(sc, stack[top], C, (uint32_t)top, dbg, gates);
    }
    *Cout = C;
    return true;
}

template <bool FASTDIV = false, class SC, class DBG>
RT_HD f3 trace_sample(const SC& sc, const CamRec& cam, float fx, float fy, float fw, float fh, float rw, float rh, int cap,
                      HitRec* stack, DBG& dbg, uint32_t gates = 0u) {
    f3 o, dir, C;
    primary_ray<FASTDIV>(cam, fx, fy, fw, fh, rw, rh, &o, &dir);
    int bounce = 0, top = 0;
    trace_chain(sc, cap, o, dir, bounce, top, stack, -1, &C, dbg, gates);
    return C;
}

// One pixel: spp == 1 is the reference; spp > 1 is the jittered extension (BASELINE.json configs[4]).
// A single call site of trace_sample: with spp == 1 the jitter is +0.0f and the average is *1.0f, both exact.
// SPP1 = true: compile-time single sample (the reference): no sample loop, no jitter, no average.
// FASTDIV (only with SPP1, w and h <= RT_FASTDIV_MAX): rw / rh = correctly rounded 1/w, 1/h; see primary_ray.
template <bool SPP1 = false, bool FASTDIV = false, class SC, class DBG>
RT_HD uint32_t trace_pixel(const SC& sc, const CamRec& cam, int x, int y, int w, int h, int cap, int spp, uint32_t seed,
                           HitRec* stack, DBG& dbg, float rw = 0.0f, float rh = 0.0f, uint32_t gates = 0u) {
    const float fw = (float)w, fh = (float)h;
    if (SPP1) return pack_color(trace_sample<FASTDIV>(sc, cam, (float)x, (float)y, fw, fh, rw, rh, cap, stack, dbg, gates));   // :1000 -> :1038
    f3 acc = mk3(0, 0, 0);
    for (int s = 0; s < spp; s++) {
        float jx = 0.0f, jy = 0.0f;
        if (spp > 1) {
            uint32_t k = ((uint32_t)y * (uint32_t)w + (uint32_t)x) * (uint32_t)spp + (uint32_t)s;
            uint32_t h1 = pcg_hash(k ^ seed), h2 = pcg_hash(h1);
            jx = (float)(h1 >> 8) * 5.9604644775390625e-08f;
            jy = (float)(h2 >> 8) * 5.9604644775390625e-08f;
        }
        acc = add3(acc, trace_sample(sc, cam, (float)x + jx, (float)y + jy, fw, fh, 0.0f, 0.0f, cap, stack, dbg));
    }
    return pack_color(mulf3(acc, 1.0f / (float)spp));                                      // :1000 -> :1038, :1046-1052
}

}  // namespace rtb
