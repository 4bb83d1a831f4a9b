// rt_lbvh.cuh — LBVH over the spheres (Morton codes -> radix sort -> Karras 2012 hierarchy -> bottom-up refit) and a
// traversal that returns EXACTLY what the reference's brute-force loops return (RayTracer.cs:577, :792, :975).
//
// Exactness contract:
//  * Leaves run the reference's own sphere test (sphere_hit, bit-exact), so traversal can never invent a hit; it can
//    only lose one by culling a node.  Culling is therefore made conservative against the fp32 NOISE of the reference
//    test, not against the exact sphere: the reference's discriminant b*b - 4*a*c carries an absolute error of up to
//    68u * a * (|oc|^2 + r^2) (u = 2^-24; derivation in DESIGN.md), so it reports hits for rays that pass as far as
//    sqrt(r^2 + ~1.0e-6 (|oc|^2 + r^2)) from the centre, and the reported point o + t*d can lie that far outside too.
//    Every node box is inflated, per ray, by  pad = 2.02e-3 * sqrt(max|oc|^2 + max r^2)  (>= 2x the bound, plus 1 % for
//    the rounding of the slab test itself), with max|oc| taken to the farthest corner of the box.
//  * primary fold (:977, strict '>'): lexicographic min over (t, original index) — ties keep the lower index.
//  * secondary fold (:804-805) is ORDER DEPENDENT (offset distance compared with stored un-offset distance): the
//    traversal collects every candidate within a window above the running minimum and replays the reference's fold
//    over them in original index order; if the replay could have reached beyond the window (chains of hits < 0.01
//    apart) or the candidate buffer overflows, the ray falls back to the brute-force loop.  Both fallbacks are exact.
//  * shadow any-hit (:573-582): boolean OR over all spheres, unbounded t, so the first confirmed hit ends the search.
#pragma once
#include "rt_trace.cuh"

namespace rtb {

// 64 B: both children's boxes live in the parent (one fetch = two box tests), interleaved per coordinate — [child 0, child 1] —
// so that each pair sits in an aligned 64-bit register pair and the two slab tests of a visit run as ONE pass of packed fp32
// (box_entry2: FADD2 / FFMA2; the box test is not reference arithmetic, only its conservativeness matters).
struct alignas(16) BvhNode {
    float lox[2], loy[2];
    float loz[2], hix[2];
    float hiy[2], hiz[2];
    int c[2];                           // child >= 0: internal node index; child < 0: ~leaf (sorted sphere position)
    int pad[2];
};
struct BvhBox { float lox, loy, loz, hix, hiy, hiz; };

// ---- build pieces (shared by the CUDA build kernels and the host-emulation tests) ------------------------------------
RT_HD uint32_t expand_bits10(uint32_t v) {          // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
// 30-bit Morton code of the sphere centre inside the centre bounds; key = code << 32 | original index (unique keys).
RT_HD uint64_t morton_key(float cx, float cy, float cz, const float* bmin, const float* binv, uint32_t index) {
    float fx = (cx - bmin[0]) * binv[0], fy = (cy - bmin[1]) * binv[1], fz = (cz - bmin[2]) * binv[2];
    fx = fx < 0.0f ? 0.0f : (fx > 1023.0f ? 1023.0f : fx);
    fy = fy < 0.0f ? 0.0f : (fy > 1023.0f ? 1023.0f : fy);
    fz = fz < 0.0f ? 0.0f : (fz > 1023.0f ? 1023.0f : fz);
    uint32_t code = (expand_bits10((uint32_t)fx) << 2) | (expand_bits10((uint32_t)fy) << 1) | expand_bits10((uint32_t)fz);
    return ((uint64_t)code << 32) | (uint64_t)index;
}
// One cell size for all three axes (the largest extent / 1024): a flat scene (a carpet of spheres on the floor) then
// spends its Morton bits on the two long axes instead of slicing the thin one into 1024 overlapping layers.
inline void morton_scale(const float bmin[3], const float bmax[3], float binv[3]) {
    float ext = 0.0f;
    for (int k = 0; k < 3; k++) { float e = bmax[k] - bmin[k]; if (e > ext) ext = e; }
    for (int k = 0; k < 3; k++) binv[k] = ext > 0.0f ? 1023.0f / ext : 0.0f;
}
RT_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
// Karras 2012: children of internal node i over n sorted unique keys. Leaves are encoded as ~position.
RT_HD void karras_node(const uint64_t* keys, int n, int i, int* left, int* right) {
    auto delta = [&](int a, int b) -> int { return (b < 0 || b >= n) ? -1 : clz64(keys[a] ^ keys[b]); };
    int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(i, i - d);
    int lmax = 2;
    while (delta(i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    *left = (lo == gamma) ? ~gamma : gamma;
    *right = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
}
RT_HD BvhBox sphere_box(f4 g, float radius) {
    BvhBox b;
    b.lox = g.x - radius; b.loy = g.y - radius; b.loz = g.z - radius;
    b.hix = g.x + radius; b.hiy = g.y + radius; b.hiz = g.z + radius;
    return b;
}
RT_HD BvhBox box_union(const BvhBox& a, const BvhBox& b) {
    BvhBox r;
    r.lox = fminf(a.lox, b.lox); r.loy = fminf(a.loy, b.loy); r.loz = fminf(a.loz, b.loz);
    r.hix = fmaxf(a.hix, b.hix); r.hiy = fmaxf(a.hiy, b.hiy); r.hiz = fmaxf(a.hiz, b.hiz);
    return r;
}
RT_HD void node_set_child_box(BvhNode& nd, int which, const BvhBox& b) {
    nd.lox[which] = b.lox; nd.loy[which] = b.loy; nd.loz[which] = b.loz; nd.hix[which] = b.hix; nd.hiy[which] = b.hiy; nd.hiz[which] = b.hiz;
}

// ---- traversal ---------------------------------------------------------------------------------------------------------
struct BvhView {
    const BvhNode* nodes;       // n-1 internal nodes, root = 0; boxes = exact sphere boxes (inflated per ray while traversing)
    const BvhNode* nodes_cam;   // same topology, boxes pre-inflated for rays that start at the frame's camera (nullable)
    const f4* sgeom_sorted;     // leaf position -> (cx, cy, cz, r^2)
    const int* orig;            // leaf position -> original sphere index
    int n;
    float r2max;                // max radiusSquared (and radius^2) over the scene, for the pad
};

RT_HD float approx_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;                                 // one MUFU.RSQ (2 ulp): the 1 % pad head-room covers it. x >= r2max > 0 here.
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x + 1e-30f));
    return x * r;
#else
    return sqrtf(x);
#endif
}

#ifndef RT_BVH_PAD_K
#define RT_BVH_PAD_K 2.02e-3f                // see header comment; overridable only to demonstrate that it is needed
#endif
constexpr float BVH_PAD_K = RT_BVH_PAD_K;
constexpr float BVH_T_SLACK = 1.0f + 4e-6f; // relative head-room on the pruning bound (t values carry ~3u error)
constexpr int BVH_STACK = 72;               // Karras tree depth <= key bits (64) + 1
constexpr int BVH_CAND = 6;                 // secondary-fold candidate buffer
constexpr float BVH_WINDOW = 0.045f;        // candidates within this distance above the minimum are replayed

// Radius of the ball around a sphere centre that contains every point the reference can report as a hit on it for rays
// starting at distance^2 `oc2` from the centre (see header): sqrt(r^2 + K^2 (|oc|^2 + r^2)), rounded up.
RT_HD float inflated_radius(float r2, float oc2) {
    float rr = r2 > 0.0f ? r2 : 0.0f;
    return sqrtf(rr + (BVH_PAD_K * BVH_PAD_K) * (oc2 + rr)) * 1.000001f + 1e-30f;
}

// Entry parameter of the ray into the box inflated by pad (conservative), or +inf when it misses t in [0, tmax].
// PAD = false: the box is already inflated for this ray origin (nodes_cam).
// The box test is NOT part of the reference's arithmetic — only its conservativeness matters — so it may use FMA:
// t = fma(bound, 1/d, -o/d) (noi = -o * inv, per ray). A NaN (inf - inf on an axis with d == 0) is dropped by
// fminf / fmaxf, i.e. that axis stops constraining: never a false miss.
RT_HD float rt_fmaf(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
template <bool PAD>
RT_HD float box_entry(f3 o, f3 inv, f3 noi, float lox, float loy, float loz, float hix, float hiy, float hiz, float r2max, float tmax) {
    float pad = 0.0f;
    if (PAD) {
        float dx = fmaxf(fabsf(lox - o.x), fabsf(hix - o.x));
        float dy = fmaxf(fabsf(loy - o.y), fabsf(hiy - o.y));
        float dz = fmaxf(fabsf(loz - o.z), fabsf(hiz - o.z));
        pad = BVH_PAD_K * approx_sqrt(rt_fmaf(dx, dx, rt_fmaf(dy, dy, rt_fmaf(dz, dz, r2max))));
    }
    float t0x = rt_fmaf(lox - pad, inv.x, noi.x), t1x = rt_fmaf(hix + pad, inv.x, noi.x);
    float t0y = rt_fmaf(loy - pad, inv.y, noi.y), t1y = rt_fmaf(hiy + pad, inv.y, noi.y);
    float t0z = rt_fmaf(loz - pad, inv.z, noi.z), t1z = rt_fmaf(hiz + pad, inv.z, noi.z);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
    return (tn <= tf * BVH_T_SLACK) ? tn : RT_INF;
}

// Both children of a node at once. Device: one pass of packed fp32 — per half exactly the operations of box_entry (an IEEE fma
// per half is the scalar fma), so e0 / e1 carry the same bits as two scalar calls; only the min / max trees stay scalar.
template <bool PAD>
RT_HD void box_entry2(const BvhNode& nd, f3 o, f3 inv, f3 noi, float r2max, float tmax, float* e0, float* e1) {
#if defined(RT_HAVE_F32X2)
    float2 lox = make_float2(nd.lox[0], nd.lox[1]), loy = make_float2(nd.loy[0], nd.loy[1]), loz = make_float2(nd.loz[0], nd.loz[1]);
    float2 hix = make_float2(nd.hix[0], nd.hix[1]), hiy = make_float2(nd.hiy[0], nd.hiy[1]), hiz = make_float2(nd.hiz[0], nd.hiz[1]);
    if (PAD) {
        const float2 ax = rt_sub2(lox, rt_splat2(o.x)), bx = rt_sub2(hix, rt_splat2(o.x));
        const float2 ay = rt_sub2(loy, rt_splat2(o.y)), by = rt_sub2(hiy, rt_splat2(o.y));
        const float2 az = rt_sub2(loz, rt_splat2(o.z)), bz = rt_sub2(hiz, rt_splat2(o.z));
        const float2 dx = make_float2(fmaxf(fabsf(ax.x), fabsf(bx.x)), fmaxf(fabsf(ax.y), fabsf(bx.y)));
        const float2 dy = make_float2(fmaxf(fabsf(ay.x), fabsf(by.x)), fmaxf(fabsf(ay.y), fabsf(by.y)));
        const float2 dz = make_float2(fmaxf(fabsf(az.x), fabsf(bz.x)), fmaxf(fabsf(az.y), fabsf(bz.y)));
        const float2 q = rt_fma2(dx, dx, rt_fma2(dy, dy, rt_fma2(dz, dz, rt_splat2(r2max))));
        const float2 pad = make_float2(BVH_PAD_K * approx_sqrt(q.x), BVH_PAD_K * approx_sqrt(q.y));
        lox = rt_sub2(lox, pad); loy = rt_sub2(loy, pad); loz = rt_sub2(loz, pad);
        hix = rt_add2(hix, pad); hiy = rt_add2(hiy, pad); hiz = rt_add2(hiz, pad);
    }
    const float2 t0x = rt_fma2(lox, rt_splat2(inv.x), rt_splat2(noi.x)), t1x = rt_fma2(hix, rt_splat2(inv.x), rt_splat2(noi.x));
    const float2 t0y = rt_fma2(loy, rt_splat2(inv.y), rt_splat2(noi.y)), t1y = rt_fma2(hiy, rt_splat2(inv.y), rt_splat2(noi.y));
    const float2 t0z = rt_fma2(loz, rt_splat2(inv.z), rt_splat2(noi.z)), t1z = rt_fma2(hiz, rt_splat2(inv.z), rt_splat2(noi.z));
    const float tn0 = fmaxf(fmaxf(fminf(t0x.x, t1x.x), fminf(t0y.x, t1y.x)), fmaxf(fminf(t0z.x, t1z.x), 0.0f));
    const float tf0 = fminf(fminf(fmaxf(t0x.x, t1x.x), fmaxf(t0y.x, t1y.x)), fminf(fmaxf(t0z.x, t1z.x), tmax));
    const float tn1 = fmaxf(fmaxf(fminf(t0x.y, t1x.y), fminf(t0y.y, t1y.y)), fmaxf(fminf(t0z.y, t1z.y), 0.0f));
    const float tf1 = fminf(fminf(fmaxf(t0x.y, t1x.y), fmaxf(t0y.y, t1y.y)), fminf(fmaxf(t0z.y, t1z.y), tmax));
    *e0 = (tn0 <= tf0 * BVH_T_SLACK) ? tn0 : RT_INF;
    *e1 = (tn1 <= tf1 * BVH_T_SLACK) ? tn1 : RT_INF;
#else
    *e0 = box_entry<PAD>(o, inv, noi, nd.lox[0], nd.loy[0], nd.loz[0], nd.hix[0], nd.hiy[0], nd.hiz[0], r2max, tmax);
    *e1 = box_entry<PAD>(o, inv, noi, nd.lox[1], nd.loy[1], nd.loz[1], nd.hix[1], nd.hiy[1], nd.hiz[1], r2max, tmax);
#endif
}

// Reciprocal direction for the slab test, kept FINITE: a component below 1e-12 of the largest one (incl. exact zeros:
// axis-parallel rays, e.g. a light with a zero coordinate used as shadow direction, RayTracer.cs:574) is treated as that
// threshold. With an infinite reciprocal the FMA form fma(bound, inv, -o*inv) yields inf - inf = NaN on one face and +inf on
// the other, which would cull a box the origin is inside. The substitution only affects culling and stays conservative: it
// tilts the ray by <= 1e-12 rad, far inside the pad. Directions whose largest component is outside [1e-12, 1e12] (zero
// vectors, denormal or non-finite directions) do not traverse at all: bvh_* hands them to the brute-force loop.
RT_HD float dir_scale(f3 d) { return fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fabsf(d.z)); }
RT_HD bool dir_in_envelope(float m) { return m >= 1e-12f && m <= 1e12f; }      // false for NaN
RT_HD float finite_rcp(float d, float thr) {
    float ad = fabsf(d);
    float r = 1.0f / (ad < thr ? thr : ad);
    return (f2bits(d) >> 31) ? -r : r;
}
RT_HD f3 safe_inv(f3 d, float m) { float thr = m * 1e-12f; return mk3(finite_rcp(d.x, thr), finite_rcp(d.y, thr), finite_rcp(d.z, thr)); }
RT_HD f3 neg_o_inv(f3 o, f3 inv) { return mk3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z); }

// Nearest fold.  off == 0: primary (:977).  off == 0.01f: secondary (:804-805), exact incl. its order dependence.
// BRUTE is a callable fallback  void(int* sel, float* t)  running the reference loop.
template <class DBG, class BRUTE>
RT_HD void bvh_nearest(const BvhView& bv, f3 o, f3 dir, float a2, float a4, float off, int* sel, float* dsel, DBG& dbg, BRUTE brute) {
    const bool secondary = off != 0.0f;
    const float window = secondary ? BVH_WINDOW : 0.0f;
    const float dscale = dir_scale(dir);
    if (dscale == 0.0f) { *sel = -1; *dsel = RT_INF; return; }      // zero direction: a = 0 -> every test is 0/0 = NaN -> miss
    if (!dir_in_envelope(dscale)) { dbg.fallback(); brute(sel, dsel); return; }
    const f3 inv = safe_inv(dir, dscale);
    const f3 noi = neg_o_inv(o, inv);
    int best = -1; float best_t = RT_INF;                // lexicographic min over (t, original index)
    int cand_i[BVH_CAND]; float cand_t[BVH_CAND]; int ncand = 0; bool overflow = false;
    int stack[BVH_STACK]; int sp = 0;
    int node = 0;
    // Primary rays (off == 0 inside a frame) all start at the camera: they traverse the per-frame pre-inflated copy with a
    // plain slab test. Everything else inflates each box for its own origin.
    const bool cam_boxes = !secondary && bv.nodes_cam != nullptr;
    const BvhNode* const nodes = cam_boxes ? bv.nodes_cam : bv.nodes;
    for (;;) {
        const BvhNode nd = nodes[node];
        dbg.node_visit(secondary ? 1 : 0);
        const float bound = (best_t + window) * BVH_T_SLACK;
        float e0, e1;
        if (cam_boxes) box_entry2<false>(nd, o, inv, noi, bv.r2max, bound, &e0, &e1);
        else box_entry2<true>(nd, o, inv, noi, bv.r2max, bound, &e0, &e1);
        int c0 = nd.c[0], c1 = nd.c[1];
        if (e1 < e0) { float te = e0; e0 = e1; e1 = te; int tc = c0; c0 = c1; c1 = tc; }   // near child first
        int next = -1;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int c = k == 0 ? c0 : c1;
            const float e = k == 0 ? e0 : e1;
            if (!(e < RT_INF)) continue;
            if (c >= 0) {
                if (next < 0) next = c; else stack[sp++] = c;
                continue;
            }
            const int leaf = ~c;
            const f4 g = bv.sgeom_sorted[leaf];
            float t;
            if (sphere_hit(sub3(o, mk3(g.x, g.y, g.z)), dir, g.w, a2, a4, 0.0f, &t, dbg)) {
                const float key = t - off;
                if (key > 0) {
                    const int oi = bv.orig[leaf];
                    if (secondary) {
                        if (t <= best_t + window) {
                            if (ncand < BVH_CAND) { cand_i[ncand] = oi; cand_t[ncand] = t; ncand++; }
                            else overflow = true;
                        }
                    }
                    if (t < best_t || (t == best_t && oi < best)) { best_t = t; best = oi; }
                }
            }
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
    if (!secondary || best < 0) { *sel = best; *dsel = best_t; return; }
    if (overflow) { dbg.fallback(); brute(sel, dsel); return; }
    // Replay the reference's fold (:792-808) over the candidates near the minimum, in original index order.
    int s_sel = -1; float closest = RT_INF, cmax = 0.0f;
    for (int done = 0; done < ncand; done++) {
        int k = -1;
        for (int q = 0; q < ncand; q++) if (cand_i[q] >= 0 && (k < 0 || cand_i[q] < cand_i[k])) k = q;   // next lowest index
        const float t = cand_t[k]; const int oi = cand_i[k]; cand_i[k] = -1;
        if (t > best_t + window) continue;              // collected against an older, larger minimum
        const float key = t - off;
        if (key > 0 && key < closest) { closest = t; s_sel = oi; if (oi >= best && t > cmax) cmax = t; }
    }
    // A sphere outside the window (t > best_t + window) could only have been accepted after the minimum if its offset
    // distance were below some running `closest` >= best_t; cmax bounds those.  Otherwise: exact fallback.
    if (cmax + 0.0101f + cmax * 2.4e-7f >= best_t + window) { dbg.fallback(); brute(sel, dsel); return; }
    *sel = s_sel; *dsel = closest;
}

// Shadow any-hit (:573-582): eps = 0.001, unbounded t, boolean result.
template <class DBG, class BRUTE>
RT_HD bool bvh_shadow_any(const BvhView& bv, f3 hit, f3 lp, float a2, float a4, DBG& dbg, BRUTE brute) {
    const float dscale = dir_scale(lp);
    if (dscale == 0.0f) return false;                                   // e.g. a light at the origin (:574): never occluded
    if (!dir_in_envelope(dscale)) { dbg.fallback(); return brute(); }
    const f3 inv = safe_inv(lp, dscale);
    const f3 noi = neg_o_inv(hit, inv);
    int stack[BVH_STACK]; int sp = 0;
    int node = 0;
    bool occluded = false;
    for (;;) {
        const BvhNode nd = bv.nodes[node];
        dbg.node_visit(2);
        float e0, e1;
        box_entry2<true>(nd, hit, inv, noi, bv.r2max, RT_INF, &e0, &e1);
        int next = -1;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int c = nd.c[k];
            const float e = k == 0 ? e0 : e1;
            if (!(e < RT_INF)) continue;
            if (c >= 0) { if (next < 0) next = c; else stack[sp++] = c; continue; }
            const f4 g = bv.sgeom_sorted[~c];
            float t;
            if (sphere_hit(sub3(hit, mk3(g.x, g.y, g.z)), lp, g.w, a2, a4, 0.001f, &t, dbg)) occluded = true;
        }
        if (occluded && !DBG::count_tests) return true;      // boolean OR: the first hit decides
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return occluded;
}

}  // namespace rtb
