// rt_gate.cuh — per-frame, host-side "gates": pixel-space regions outside of which a part of the reference's per-pixel work
// is PROVEN to find nothing, so the single-sample tiny-scene kernels skip it and keep the exact result. None of this is
// reference arithmetic — only conservativeness matters — and everything here is evaluated in double on the host (a few
// microseconds per frame). Any doubt (odd camera, several planes, unbounded projection, non-finite input) yields the gate that
// never skips. rt_set_option(RT_OPT_PRIMARY_GATE, 0) turns all of them off; frames are identical either way (tests).
//
// Facts used.
//  (F1) Sphere test, RayTracer.cs:613-642: the fp32 discriminant b^2 - 4ac can only be >= 0 for a ray whose exact line passes
//       within R' = sqrt(r^2 + K^2 (|oc|^2 + r^2)) of the centre, K = RT_BVH_PAD_K (derivation: rt_lbvh.cuh, DESIGN.md §5 — the
//       bound the LBVH boxes rely on), |oc| = distance from the ray origin to the centre.
//  (F2) All primary rays start at the camera position P, and the direction the kernel really uses, normalize(fl(vp - P))
//       (:964-971), is within an angle eps of the ideal direction D(x, y) = R (x/w - .5) pw + U (y/h - .5) ph + F near of its
//       pixel (primary_dir_eps: rounding of u, v, the scaled basis vectors and the running sum, which grows with |P|; the basis
//       must be orthonormal to 1e-5).
//  (F3) Plane test, :590-604: t = num / den > 0 with num = plane_num(origin) — ONE fp32 number per frame for primary rays,
//       evaluated here with the kernel's own expression — and den = Dot(direction, normal), whose sign is that of D . n outside
//       a margin. D . n is affine in (x, y).
//
// Gates (FrameGates), for pixel (x, y):
//  spheres  rectangle. Outside: no sphere can be hit by the primary ray -> the sphere loop :975-981 is skipped.
//           [cone of directions within R'' of a centre, R'' = R' + eps (|oc| + R'), projected per axis by the tangent planes
//            k^2 (Z^2 - R^2) - 2 X Z k + (X^2 - R^2) = 0, +-2 pixels]
//  sky      affine, > 0: den has the wrong sign -> the plane cannot be hit (single plane only). Outside `spheres` AND sky:
//           the pixel is 0x00000000 (:993, :1000) without tracing.
//  deep     affine, > 0: the primary ray meets the plane at a slope >= sin_min, hence within dmax = H / sin_min of P (H = distance
//           of P from the plane). Precondition of the two gates below, which concern a primary ray that HIT the plane
//           (checked at run time).
//  mirror   rectangle. deep AND outside: the reflection ray of the plane hit (:741-746) hits nothing, so the mirror term is
//           (0,0,0) and the ray is not traced. [the exact reflection of the primary line passes through the mirror image P' of P:
//           the spheres seen from P', inflated by the fp32 deviations of the hit point and of the reflected direction; and the
//           reflection cannot re-hit its own plane beyond the 0.01 cut-off :731 — t_self bound below; needs |n| = 1 to 1e-6]
//  shadow[l] rectangle. deep AND outside: the shadow ray of light l from the plane hit (:752, direction = light POSITION :574)
//           meets no sphere -> unoccluded, IntersectShadowLight is not evaluated. [the points of the plane whose shadow line
//           passes within R'' of a centre form an ellipse (cylinder around the line centre + s * light, cut by the plane); a
//           circumscribed polygon of it, clipped to the front of the camera, is projected to pixels; + the pixel margin that
//           covers eps]
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "rt_lbvh.cuh"

namespace rtb {

constexpr int RT_GATE_LIGHTS = (int)RT_GATE_MAX_LIGHTS;   // shadow gates exist for the first 4 lights; further lights are always tested

struct GateRect { int x0, y0, x1, y1; };      // inclusive pixel ranges; empty = {w, h, w, h}
struct GateAffine { float a, bx, by; };       // value(x, y) = fma(bx, x, fma(by, y, a)); "never" = {-1, 0, 0}, "always" = {+1, 0, 0}
// Gate SHAPES, second generation (shipped since round 2: the whole GPU parity suite, incl. the chain hashes of the production kernel,
// passes with them; +3.2 % on the bench frame, profiles/r02/) — tighter shapes for the two gates whose union rectangles leave the
// most on the table:
//   mirror_s[i]   one rectangle per sphere (scenes of <= 4 spheres): the reflection is skipped if the span is outside ALL of them
//   shadow_out[l] two half-planes per light, taken from the convex hull of the projected shadow polygons (the two hull edges that
//                 cut the most off the rectangle): > 0 on the whole span => outside the hull => unoccluded. Fits the diagonal
//                 streaks of low lights, which a rectangle cannot.
// -DRT_GATES_V1 compiles the first-generation shapes (union rectangles only); the tests use it to show V2 never skips less.
#ifndef RT_GATES_V1
#define RT_GATES_V2 1
#endif
#ifdef RT_GATES_V2
constexpr int RT_GATE_MIRROR_RECTS = 4;
#endif
struct FrameGates {
    GateRect spheres;
    GateAffine sky;
    GateAffine deep;
    GateRect mirror;
    GateRect shadow[RT_GATE_LIGHTS];
#ifdef RT_GATES_V2
    int n_mirror_s;                                  // 0: only the union rectangle `mirror`
    GateRect mirror_s[RT_GATE_MIRROR_RECTS];
    GateAffine shadow_out[RT_GATE_LIGHTS][2];        // "never" = {-1, 0, 0}
#endif
};
enum : uint32_t { GATE_SKIP_SPHERES = RT_GATE_SPHERES, GATE_SKIP_MIRROR = RT_GATE_MIRROR, GATE_SKIP_SHADOW0 = 1u << RT_GATE_SHADOW_SHIFT };   // bits of gate_bits() (rt_trace.cuh)

inline GateRect gate_full(int w, int h) { GateRect g = {0, 0, w - 1, h - 1}; return g; }
inline GateRect gate_empty(int w, int h) { GateRect g = {w, h, w, h}; return g; }
inline FrameGates gates_off(int w, int h) {
    FrameGates g;
    g.spheres = gate_full(w, h); g.mirror = gate_full(w, h);
    for (int i = 0; i < RT_GATE_LIGHTS; i++) g.shadow[i] = gate_full(w, h);
    g.sky.a = -1.0f; g.sky.bx = 0.0f; g.sky.by = 0.0f; g.deep = g.sky;
#ifdef RT_GATES_V2
    g.n_mirror_s = 0;
    for (int i = 0; i < RT_GATE_MIRROR_RECTS; i++) g.mirror_s[i] = gate_full(w, h);
    for (int l = 0; l < RT_GATE_LIGHTS; l++) g.shadow_out[l][0] = g.shadow_out[l][1] = g.sky;
#endif
    return g;
}

namespace gate_detail {

constexpr double U32 = 5.9604644775390625e-08;          // 2^-24
constexpr double PI = 3.14159265358979323846;
inline double dot3d(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline bool finite3d(const double* a) { return std::isfinite(a[0]) && std::isfinite(a[1]) && std::isfinite(a[2]); }

struct Cam {                 // a pinhole in double: position, orthonormal basis, view-plane size, frame size
    double P[3], R[3], U[3], F[3], pw, ph, nearp;
    int w, h;
};
inline Cam load_cam(const CamRec& c, int w, int h) {
    Cam k = {{c.pos.x, c.pos.y, c.pos.z}, {c.right.x, c.right.y, c.right.z}, {c.up.x, c.up.y, c.up.z}, {c.fwd.x, c.fwd.y, c.fwd.z},
             c.view.x, c.view.y, c.view.z, w, h};
    return k;
}

// (F2) bound on the angle between the fp32 primary direction and the ideal one, or -1 if the camera is not one the gates
// are derived for.
inline double primary_dir_eps(const Cam& c) {
    if (!finite3d(c.P) || !finite3d(c.R) || !finite3d(c.U) || !finite3d(c.F) || !std::isfinite(c.pw) || !std::isfinite(c.ph) || !std::isfinite(c.nearp)) return -1;
    if (!(c.pw > 1e-6) || !(c.ph > 1e-6) || !(c.nearp > 1e-6) || c.pw > 1e6 || c.ph > 1e6 || c.nearp > 1e6) return -1;
    const double tol = 1e-5;
    if (std::fabs(dot3d(c.R, c.R) - 1) > tol || std::fabs(dot3d(c.U, c.U) - 1) > tol || std::fabs(dot3d(c.F, c.F) - 1) > tol ||
        std::fabs(dot3d(c.R, c.U)) > tol || std::fabs(dot3d(c.R, c.F)) > tol || std::fabs(dot3d(c.U, c.F)) > tol) return -1;
    const double pmax = std::fmax(std::fabs(c.P[0]), std::fmax(std::fabs(c.P[1]), std::fabs(c.P[2])));
    const double L = 0.5 * c.pw + 0.5 * c.ph + c.nearp;
    const double E = U32 * (8.0 * (pmax + L) + 4.0 * (c.pw + c.ph));          // per-component error of fl(vp - P), factor 2 in hand
    const double eps = 2.0 * (2.0 * std::sqrt(3.0) * E / c.nearp + 1e-6) + 4e-5;   // + normalisation, + basis tolerance
    return eps < 1e-2 ? eps : -1;
}

// an angle `ang` moves a pixel by at most ang * sec^2 * (pixels per unit of slope); + 2 pixels
inline double pixel_margin(const Cam& c, double ang) {
    const double sec2 = 1.0 + 0.25 * (c.pw * c.pw + c.ph * c.ph) / (c.nearp * c.nearp);
    const double ppu = std::fmax(c.w * c.nearp / c.pw, c.h * c.nearp / c.ph);
    return std::ceil(ang * sec2 * ppu) + 2.0;
}

// Accumulates into `out` (init {w, h, -1, -1}) the pixel bounds of the directions from c.P that pass within Rg of `ctr`.
// Returns false when the projection is unbounded / undefined (caller must fall back to the full frame).
inline bool add_ball_rect(const Cam& c, const double* ctr, double Rg, double margin_px, GateRect* out) {
    const double oc[3] = {ctr[0] - c.P[0], ctr[1] - c.P[1], ctr[2] - c.P[2]};
    const double ocl = std::sqrt(dot3d(oc, oc));
    if (!std::isfinite(ocl) || !std::isfinite(Rg)) return false;
    if (!(ocl > Rg * 1.001)) return false;                  // eye inside (or on) the ball
    const double X = dot3d(oc, c.R), Y = dot3d(oc, c.U), Z = dot3d(oc, c.F);
    if (Z < -Rg * 1.001) return true;                       // wholly behind the eye plane: every ray of the frame has z > 0
    if (!(Z > Rg * 1.001)) return false;                    // crosses the eye plane: unbounded projection
    const double den = Z * Z - Rg * Rg;
    int lim[2][2];
    const double ctr2[2] = {X, Y}, size[2] = {c.pw, c.ph};
    const int npx[2] = {c.w, c.h};
    for (int a = 0; a < 2; a++) {
        const double A = ctr2[a];
        const double root = Rg * std::sqrt(A * A + Z * Z - Rg * Rg);
        const double k1 = (A * Z - root) / den, k2 = (A * Z + root) / den;
        const double p1 = (k1 * c.nearp / size[a] + 0.5) * npx[a], p2 = (k2 * c.nearp / size[a] + 0.5) * npx[a];   // (x/w - .5) pw = k near
        if (!std::isfinite(p1) || !std::isfinite(p2)) return false;
        const double lo = std::floor(std::fmin(p1, p2)) - margin_px, hi = std::ceil(std::fmax(p1, p2)) + margin_px;
        lim[a][0] = lo < 0 ? 0 : (lo > npx[a] ? npx[a] : (int)lo);
        lim[a][1] = hi > npx[a] - 1 ? npx[a] - 1 : (hi < -1 ? -1 : (int)hi);
    }
    if (lim[0][1] < lim[0][0] || lim[1][1] < lim[1][0]) return true;       // projects outside the frame
    if (lim[0][0] < out->x0) out->x0 = lim[0][0];
    if (lim[1][0] < out->y0) out->y0 = lim[1][0];
    if (lim[0][1] > out->x1) out->x1 = lim[0][1];
    if (lim[1][1] > out->y1) out->y1 = lim[1][1];
    return true;
}
inline GateRect finish_rect(const GateRect& acc, int w, int h) {
    if (acc.x1 < acc.x0 || acc.y1 < acc.y0) return gate_empty(w, h);
    return acc;
}
inline double inflated(double r2, double oc_max) {          // (F1)
    const double K = (double)RT_BVH_PAD_K;
    return std::sqrt(r2 + K * K * (oc_max * oc_max + r2)) * (1.0 + 1e-6) + 1e-30;
}

// unit circle sampled for the circumscribed shadow polygons (built once)
constexpr int GATE_POLY = 32;
struct PolyTable { double cs[GATE_POLY], sn[GATE_POLY], circ; };
inline const PolyTable& poly_table() {
    static const PolyTable t = [] {
        PolyTable p;
        for (int v = 0; v < GATE_POLY; v++) { p.cs[v] = std::cos(2.0 * PI * v / GATE_POLY); p.sn[v] = std::sin(2.0 * PI * v / GATE_POLY); }
        p.circ = 1.0 / std::cos(PI / GATE_POLY);
        return p;
    }();
    return t;
}

// float coefficients of an affine gate that is > 0 only where sgn * (a + bx x + by y) > margin, with the evaluation error
// of the two device FMAs and of the double -> float conversions folded in
inline GateAffine make_affine(double sgn, double a, double bx, double by, double margin, int w, int h) {
    const double S = std::fabs(a) + std::fabs(bx) * w + std::fabs(by) * h;
    GateAffine g;
    g.bx = (float)(sgn * bx); g.by = (float)(sgn * by);
    g.a = std::nextafter((float)(sgn * a - margin - 16.0 * U32 * S), -INFINITY);
    if (!std::isfinite(g.a) || !std::isfinite(g.bx) || !std::isfinite(g.by)) { g.a = -1.0f; g.bx = 0.0f; g.by = 0.0f; }
    return g;
}

#ifdef RT_GATES_V2
struct P2 { double x, y; };
// convex hull (Andrew's monotone chain), counter-clockwise, no repeated end point; n <= a few hundred
inline int convex_hull(P2* pts, int n, P2* hull) {
    for (int i = 1; i < n; i++) {                    // insertion sort by (x, y)
        P2 v = pts[i]; int j = i - 1;
        while (j >= 0 && (pts[j].x > v.x || (pts[j].x == v.x && pts[j].y > v.y))) { pts[j + 1] = pts[j]; j--; }
        pts[j + 1] = v;
    }
    auto cross = [](const P2& o, const P2& a, const P2& b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
    int k = 0;
    for (int i = 0; i < n; i++) { while (k >= 2 && cross(hull[k - 2], hull[k - 1], pts[i]) <= 0) k--; hull[k++] = pts[i]; }
    for (int i = n - 2, t = k + 1; i >= 0; i--) { while (k >= t && cross(hull[k - 2], hull[k - 1], pts[i]) <= 0) k--; hull[k++] = pts[i]; }
    return k > 1 ? k - 1 : k;
}
// area of the part of rectangle r on the side nx x + ny y + c < 0 of a line (Sutherland-Hodgman against one half-plane)
inline double rect_area_outside(const GateRect& r, double nx, double ny, double cc) {
    const P2 q[4] = {{(double)r.x0, (double)r.y0}, {(double)r.x1 + 1, (double)r.y0}, {(double)r.x1 + 1, (double)r.y1 + 1}, {(double)r.x0, (double)r.y1 + 1}};
    P2 out[8]; int m = 0;
    for (int i = 0; i < 4; i++) {
        const P2 a = q[i], b = q[(i + 1) & 3];
        const double da = nx * a.x + ny * a.y + cc, db = nx * b.x + ny * b.y + cc;
        if (da < 0) out[m++] = a;
        if ((da < 0) != (db < 0)) { const double t = da / (da - db); out[m++] = {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)}; }
    }
    double area = 0;
    for (int i = 0; i < m; i++) { const P2 a = out[i], b = out[(i + 1) % m]; area += a.x * b.y - a.y * b.x; }
    return 0.5 * std::fabs(area);
}
#endif

}  // namespace gate_detail

// All gates of one frame. planes: PlaneRec as uploaded (n, cn); lights: LightRec (p).
inline FrameGates compute_frame_gates(const CamRec& camrec, int w, int h, const f4* sgeom, int ns, const PlaneRec* planes, int np,
                                      const LightRec* lights, int nl) {
    using namespace gate_detail;
    FrameGates g = gates_off(w, h);
    const Cam c = load_cam(camrec, w, h);
    const double eps = primary_dir_eps(c);
    if (eps < 0) { if (ns <= 0) g.spheres = gate_empty(w, h); return g; }
    for (int i = 0; i < ns; i++) {
        const double ctr[3] = {sgeom[i].x, sgeom[i].y, sgeom[i].z};
        if (!finite3d(ctr) || !std::isfinite((double)sgeom[i].w)) return g;
    }
    // ---- spheres --------------------------------------------------------------------------------------------------------
    {
        GateRect acc = {w, h, -1, -1};
        bool ok = true;
        for (int i = 0; i < ns && ok; i++) {
            const double ctr[3] = {sgeom[i].x, sgeom[i].y, sgeom[i].z};
            const double r2 = sgeom[i].w > 0.0f ? (double)sgeom[i].w : 0.0;
            const double oc[3] = {ctr[0] - c.P[0], ctr[1] - c.P[1], ctr[2] - c.P[2]};
            const double ocl = std::sqrt(dot3d(oc, oc));
            const double Rp = inflated(r2, ocl);
            ok = add_ball_rect(c, ctr, Rp + eps * (ocl + Rp), 2.0, &acc);
        }
        g.spheres = ok ? finish_rect(acc, w, h) : gate_full(w, h);
    }
    // ---- the plane ------------------------------------------------------------------------------------------------------
    if (np <= 0) { g.sky.a = 1.0f; return g; }                   // nothing but spheres can be hit
    if (np > 1) return g;
    const PlaneRec& pl = planes[0];
    f4 pn; pn.x = pl.n.x; pn.y = pl.n.y; pn.z = pl.n.z; pn.w = pl.cn;
    const float num = plane_num(camrec.pos, pn);                 // (F3) the kernel's own fp32 value (same expression, no contraction)
    if (!(num == num)) return g;
    if (num == 0.0f) { g.sky.a = 1.0f; return g; }               // 0 / den is never > 0 (:598)
    const double n[3] = {pl.n.x, pl.n.y, pl.n.z};
    const double nlen = std::sqrt(dot3d(n, n));
    if (!std::isfinite(nlen) || !(nlen > 1e-12) || !std::isfinite((double)pl.cn)) return g;
    const double nh[3] = {n[0] / nlen, n[1] / nlen, n[2] / nlen};
    // ideal D . n_hat = a + bx x + by y
    const double Rn = dot3d(c.R, nh), Un = dot3d(c.U, nh), Fn = dot3d(c.F, nh);
    const double a = -0.5 * c.pw * Rn - 0.5 * c.ph * Un + c.nearp * Fn, bx = c.pw * Rn / w, by = c.ph * Un / h;
    const double Dmax = std::sqrt(c.nearp * c.nearp + 0.25 * c.pw * c.pw + 0.25 * c.ph * c.ph) * 1.0001;
    const double away = num > 0 ? -1.0 : 1.0;                    // sign of D . n for which den has the wrong sign (or is zero)
    const double den_margin = 2.0 * (eps + 24.0 * U32) * std::sqrt(3.0);          // |den_fp / |n| - D_hat . n_hat| is below this
    g.sky = make_affine(away, a, bx, by, den_margin * Dmax, w, h);
    // ---- deep: bounded plane hits ---------------------------------------------------------------------------------------
    const double c0 = (double)pl.cn / nlen;                      // plane: x . n_hat = c0
    const double hP = dot3d(c.P, nh) - c0;                       // signed distance of the eye
    const double H = std::fabs(hP);
    if (!(H > 1e-3) || !(eps < 2e-3)) return g;
    int kmax = 0;
    for (int k = 1; k < 3; k++) if (std::fabs(nh[k]) > std::fabs(nh[kmax])) kmax = k;
    double T = 0;                                                // sum of the non-dominant |n_hat| components (0 if axis-aligned)
    for (int k = 0; k < 3; k++) if (k != kmax) T += std::fabs(nh[k]);
    const double Pl = std::sqrt(dot3d(c.P, c.P));
    double PnAbs = 0;
    for (int k = 0; k < 3; k++) PnAbs += std::fabs(c.P[k] * nh[k]);
    double sin_min = 0.02, dmax = 0, delta_h = 0;
    bool mirror_ok = std::fabs(nlen - 1.0) < 1e-6;               // :741-744 reflects with the normal as given: a mirror only if unit
    for (;; sin_min *= 2.0) {
        if (sin_min > 0.5) { mirror_ok = false; sin_min = 0.02; }
        const double s_lo = sin_min - den_margin;                // slope the fp32 direction is guaranteed to have
        dmax = H / s_lo * 1.001;
        // relative error of d = num / den: cancellation in num, rounding of den (only the non-dominant normal components add
        // absolute error), the division
        const double rel_d = 4.0 * U32 * (PnAbs + std::fabs(c0)) / H + 3.0 * U32 * (1.0 + 2.0 * T / s_lo) + 2.0 * U32;
        // |h_k| bounds: within dmax of P; along the dominant normal axis the plane equation pins it
        double hk[3], side = 0;
        for (int k = 0; k < 3; k++) { hk[k] = std::fabs(c.P[k]) + dmax; if (k != kmax) side += std::fabs(nh[k]) * hk[k]; }
        hk[kmax] = std::fmin(hk[kmax], (std::fabs(c0) + side + 1e-3) / std::fabs(nh[kmax]));
        double Bnd = 0;
        for (int k = 0; k < 3; k++) Bnd += std::fabs(nh[k]) * hk[k];
        // displacement of the fp32 hit point from the exact plane hit of the fp32 primary line
        delta_h = 2.0 * U32 * std::sqrt(3.0) * (Pl + dmax) + dmax * rel_d;
        if (!mirror_ok) break;
        // the reflection re-tests its own plane (:812-819): |num_self| / |n| <= off-plane distance of the fp32 hit point
        // + rounding of the expression; den_self / |n| >= s_lo - 1e-5. It must stay below the 0.01 cut-off (:731) with room.
        const double off_plane = H * rel_d + 2.0 * U32 * Bnd;
        const double num_self = off_plane + 4.0 * U32 * (Bnd + std::fabs(c0));
        const double t_self = 2.0 * num_self / (s_lo - 1e-5);
        if (t_self < 0.004) break;
    }
    g.deep = make_affine(-away, a, bx, by, sin_min * Dmax, w, h);      // toward the plane: -away * D_hat . n_hat >= sin_min (|D| <= Dmax)
    const double mpx = pixel_margin(c, eps);
    // ---- mirror: spheres seen from the mirror image of the eye --------------------------------------------------------------
    if (mirror_ok) {
        Cam m = c;
        for (int k = 0; k < 3; k++) {
            m.P[k] = c.P[k] - 2.0 * hP * nh[k];
            m.R[k] = c.R[k] - 2.0 * Rn * nh[k]; m.U[k] = c.U[k] - 2.0 * Un * nh[k]; m.F[k] = c.F[k] - 2.0 * Fn * nh[k];
        }
        const double eps_r = 5e-6 + 8.0 * U32;                   // |n| = 1 to 1e-6, rounding of the reflection formula
        GateRect acc = {w, h, -1, -1};
        bool ok = true;
        for (int i = 0; i < ns && ok; i++) {
            const double ctr[3] = {sgeom[i].x, sgeom[i].y, sgeom[i].z};
            const double r2 = sgeom[i].w > 0.0f ? (double)sgeom[i].w : 0.0;
            const double pc[3] = {ctr[0] - c.P[0], ctr[1] - c.P[1], ctr[2] - c.P[2]};
            const double mc[3] = {ctr[0] - m.P[0], ctr[1] - m.P[1], ctr[2] - m.P[2]};
            const double Rp = inflated(r2, std::sqrt(dot3d(pc, pc)) + dmax);          // |oc| of the reflection ray <= |P c| + dmax
            const double ml = std::sqrt(dot3d(mc, mc));
            const double Rg = Rp + (eps + 2.0 * eps_r) * (ml + Rp) + 2.0 * delta_h;
            ok = add_ball_rect(m, ctr, Rg, 2.0, &acc);
#ifdef RT_GATES_V2
            if (ok && ns <= RT_GATE_MIRROR_RECTS) {
                GateRect one = {w, h, -1, -1};
                add_ball_rect(m, ctr, Rg, 2.0, &one);            // cannot fail: the union call above just succeeded
                g.mirror_s[i] = finish_rect(one, w, h);
            }
#endif
        }
        g.mirror = ok ? finish_rect(acc, w, h) : gate_full(w, h);
#ifdef RT_GATES_V2
        g.n_mirror_s = (ok && ns <= RT_GATE_MIRROR_RECTS) ? ns : 0;
#endif
    }
    // ---- shadow[l]: shadow ellipses on the plane, seen from the eye ---------------------------------------------------------
    for (int l = 0; l < nl && l < RT_GATE_LIGHTS; l++) {
        const double lp[3] = {lights[l].p.x, lights[l].p.y, lights[l].p.z};
        const double ll = std::sqrt(dot3d(lp, lp));
        if (!finite3d(lp) || !(ll > 1e-12) || !std::isfinite(ll)) continue;                 // stays "full": always tested
        const double lh[3] = {lp[0] / ll, lp[1] / ll, lp[2] / ll};
        const double ln = dot3d(lh, nh);
        if (std::fabs(ln) < 1e-3) continue;                      // light direction (almost) in the plane: unbounded shadows
        // in-plane axes: e1 along the projected light direction (long axis), e2 across
        double e1[3], e2[3];
        for (int k = 0; k < 3; k++) e1[k] = lh[k] - ln * nh[k];
        const double e1l = std::sqrt(dot3d(e1, e1));
        if (e1l < 1e-9) {                                        // light along the normal: a circle; any in-plane basis
            const double t[3] = {std::fabs(nh[0]) < 0.9 ? 1.0 : 0.0, std::fabs(nh[0]) < 0.9 ? 0.0 : 1.0, 0.0};
            const double tn = dot3d(t, nh);
            for (int k = 0; k < 3; k++) e1[k] = t[k] - tn * nh[k];
            const double l1 = std::sqrt(dot3d(e1, e1));
            for (int k = 0; k < 3; k++) e1[k] /= l1;
        } else {
            for (int k = 0; k < 3; k++) e1[k] /= e1l;
        }
        e2[0] = nh[1] * e1[2] - nh[2] * e1[1]; e2[1] = nh[2] * e1[0] - nh[0] * e1[2]; e2[2] = nh[0] * e1[1] - nh[1] * e1[0];
        GateRect acc = {w, h, -1, -1};
        bool ok = true;
        constexpr int M = GATE_POLY;
        const PolyTable& tab = poly_table();
        const double circ = tab.circ;                            // circumscribed polygon
        const double zc = 1e-4 * c.nearp;
#ifdef RT_GATES_V2
        constexpr int MAXP = 8 * 2 * GATE_POLY;                  // TINY_MAX_SPHERES polygons, clipping adds at most 2 points each
        P2 pts[MAXP]; int npts = 0;
#endif
        auto project = [&](double x, double y, double z) {
            const double px = (x / z * c.nearp / c.pw + 0.5) * w, py = (y / z * c.nearp / c.ph + 0.5) * h;
#ifdef RT_GATES_V2
            if (npts < MAXP) { pts[npts].x = px; pts[npts].y = py; }
            npts++;
#endif
            const double cx = std::fmax(-1e6, std::fmin(1e6, px)), cy = std::fmax(-1e6, std::fmin(1e6, py));   // before the int conversion
            const int x0 = (int)std::floor(cx - mpx), x1 = (int)std::ceil(cx + mpx), y0 = (int)std::floor(cy - mpx), y1 = (int)std::ceil(cy + mpx);
            if (x0 < acc.x0) acc.x0 = x0;
            if (y0 < acc.y0) acc.y0 = y0;
            if (x1 > acc.x1) acc.x1 = x1;
            if (y1 > acc.y1) acc.y1 = y1;
        };
        for (int i = 0; i < ns && ok; i++) {
            const double ctr[3] = {sgeom[i].x, sgeom[i].y, sgeom[i].z};
            const double r2 = sgeom[i].w > 0.0f ? (double)sgeom[i].w : 0.0;
            const double pc[3] = {ctr[0] - c.P[0], ctr[1] - c.P[1], ctr[2] - c.P[2]};
            const double Rg = (inflated(r2, std::sqrt(dot3d(pc, pc)) + dmax) + 2.0 * delta_h) * 1.0001;
            const double s0 = (dot3d(ctr, nh) - c0) / ln;        // ctr - s0 * l_hat lies on the plane
            double e0[3];
            for (int k = 0; k < 3; k++) e0[k] = ctr[k] - s0 * lh[k];
            const double ax1 = Rg / std::fabs(ln) * circ, ax2 = Rg * circ;
            if (!std::isfinite(ax1) || !finite3d(e0)) { ok = false; break; }
            double X[M], Y[M], Z[M];                             // polygon vertices in eye coordinates
            for (int v = 0; v < M; v++) {
                const double cs = tab.cs[v], sn = tab.sn[v];
                double q[3];
                for (int k = 0; k < 3; k++) q[k] = e0[k] + ax1 * cs * e1[k] + ax2 * sn * e2[k] - c.P[k];
                X[v] = dot3d(q, c.R); Y[v] = dot3d(q, c.U); Z[v] = dot3d(q, c.F);
                if (!std::isfinite(X[v]) || !std::isfinite(Y[v]) || !std::isfinite(Z[v])) ok = false;
            }
            if (!ok) break;
            for (int v = 0; v < M; v++) {                        // clipped to z >= zc, projected
                const int u2 = (v + 1) % M;
                if (Z[v] >= zc) project(X[v], Y[v], Z[v]);
                if ((Z[v] >= zc) != (Z[u2] >= zc)) {             // the edge crosses the clip plane: add the crossing point
                    const double t = (zc - Z[v]) / (Z[u2] - Z[v]);
                    project(X[v] + t * (X[u2] - X[v]), Y[v] + t * (Y[u2] - Y[v]), zc);
                }
            }
        }
        if (!ok) continue;
        if (acc.x1 < acc.x0 || acc.y1 < acc.y0) { g.shadow[l] = gate_empty(w, h); continue; }
        GateRect r = {acc.x0 < 0 ? 0 : acc.x0, acc.y0 < 0 ? 0 : acc.y0, acc.x1 > w - 1 ? w - 1 : acc.x1, acc.y1 > h - 1 ? h - 1 : acc.y1};
        g.shadow[l] = (r.x1 < r.x0 || r.y1 < r.y0) ? gate_empty(w, h) : r;
#ifdef RT_GATES_V2
        if (!(r.x1 < r.x0 || r.y1 < r.y0) && npts >= 3 && npts <= MAXP) {
            // Every possibly-shadowed pixel lies inside the convex hull of the projected (clipped) polygon vertices, grown by the
            // pixel margin. Keep the two hull edges that cut the most area off the rectangle; coordinates are taken relative
            // to the rectangle's centre so that the float coefficients stay small.
            bool fin = true;
            for (int i = 0; i < npts; i++) fin = fin && std::isfinite(pts[i].x) && std::isfinite(pts[i].y);
            P2 hull[MAXP + 1];
            const int nh_ = fin ? convex_hull(pts, npts, hull) : 0;
            double best[2] = {0, 0}; double coef[2][3] = {{0, 0, 0}, {0, 0, 0}};
            for (int i = 0; i < nh_ && nh_ >= 3; i++) {
                const P2 a2 = hull[i], b2 = hull[(i + 1) % nh_];
                const double ex = b2.x - a2.x, ey = b2.y - a2.y, el = std::sqrt(ex * ex + ey * ey);
                if (!(el > 1e-9)) continue;
                const double nx = -ey / el, ny = ex / el;         // inward normal of a counter-clockwise hull
                const double cc = -(nx * a2.x + ny * a2.y) + mpx; // inside (grown by the margin): nx x + ny y + cc >= 0
                if (!std::isfinite(cc) || std::fabs(nx * (0.5 * w) + ny * (0.5 * h) + cc) > 1e5) continue;    // a line far from the frame
                const double area = rect_area_outside(r, nx, ny, cc);
                int slot = area > best[0] ? 0 : (area > best[1] ? 1 : -1);
                if (slot == 0) { best[1] = best[0]; for (int k = 0; k < 3; k++) coef[1][k] = coef[0][k]; }
                if (slot >= 0) { best[slot] = area; coef[slot][0] = cc; coef[slot][1] = nx; coef[slot][2] = ny; }
            }
            for (int e = 0; e < 2; e++)
                if (best[e] > 0.02 * (double)(r.x1 - r.x0 + 1) * (double)(r.y1 - r.y0 + 1))                    // worth two FMAs
                    g.shadow_out[l][e] = make_affine(-1.0, coef[e][0], coef[e][1], coef[e][2], 0.0, w, h);      // > 0 => outside the hull
        }
#endif
    }
    return g;
}

// ---- device side ------------------------------------------------------------------------------------------------------------
// Gates are evaluated once per SPAN of consecutive pixels of one row, xa..xb (the pixels one thread renders): a skip bit is set
// only if it holds for every pixel of the span. Regions are large, so little is lost at their borders, and the per-pixel cost
// of the gates (each rectangle is four constant-bank loads and four integer instructions) is divided by the span length.
RT_HD bool span_outside(const GateRect& g, int xa, int xb, int y) {
    return xb < g.x0 || xa > g.x1 || (unsigned)(y - g.y0) > (unsigned)(g.y1 - g.y0);
}
RT_HD bool span_positive(const GateAffine& g, float fxa, float fxb, float fy) {     // affine: extreme at an end of the span
    const float base = rt_fmaf(g.by, fy, g.a);
    return rt_fmaf(g.bx, fxa, base) > 0.0f && rt_fmaf(g.bx, fxb, base) > 0.0f;
}
// The skip bits (RT_GATE_*, rt_trace.cuh) common to pixels (xa..xb, y), nl = number of lights; *black: all of them are
// 0x00000000 without tracing.
RT_HD uint32_t gate_bits_span(const FrameGates& g, int xa, int xb, int y, int nl, bool* black) {
    const float fxa = (float)xa, fxb = (float)xb, fy = (float)y;
    uint32_t bits = span_outside(g.spheres, xa, xb, y) ? (uint32_t)GATE_SKIP_SPHERES : 0u;
    *black = bits != 0u && span_positive(g.sky, fxa, fxb, fy);
    if (!*black && span_positive(g.deep, fxa, fxb, fy)) {
#ifdef RT_GATES_V2
        bool no_mirror = span_outside(g.mirror, xa, xb, y);
        if (!no_mirror && g.n_mirror_s > 0) {
            no_mirror = true;
#pragma unroll
            for (int i = 0; i < RT_GATE_MIRROR_RECTS; i++) if (i < g.n_mirror_s && !span_outside(g.mirror_s[i], xa, xb, y)) no_mirror = false;
        }
        if (no_mirror) bits |= GATE_SKIP_MIRROR;
#pragma unroll
        for (int l = 0; l < RT_GATE_LIGHTS; l++)
            if (l < nl && (span_outside(g.shadow[l], xa, xb, y) || span_positive(g.shadow_out[l][0], fxa, fxb, fy) ||
                           span_positive(g.shadow_out[l][1], fxa, fxb, fy))) bits |= (uint32_t)GATE_SKIP_SHADOW0 << l;
#else
        if (span_outside(g.mirror, xa, xb, y)) bits |= GATE_SKIP_MIRROR;
#pragma unroll
        for (int l = 0; l < RT_GATE_LIGHTS; l++)
            if (l < nl && span_outside(g.shadow[l], xa, xb, y)) bits |= (uint32_t)GATE_SKIP_SHADOW0 << l;
#endif
    }
    return bits;
}
// Only the `black` verdict of gate_bits_span (same two predicates, same arithmetic): what the fill / expand passes of the multi-GPU
// gathers need — they never look at the mirror and shadow gates.
RT_HD bool gate_black_span(const FrameGates& g, int xa, int xb, int y) {
    return span_outside(g.spheres, xa, xb, y) && span_positive(g.sky, (float)xa, (float)xb, (float)y);
}
RT_HD uint32_t gate_bits(const FrameGates& g, int x, int y, bool* black) { return gate_bits_span(g, x, x, y, RT_GATE_LIGHTS, black); }

// ---- sparse device -> host return (host side, used by render_frames in rtb200.cu) ------------------------------------------------
// What the frame gates (rt_gate.cuh) prove black need not cross PCIe. Per frame the rows are classified on the host:
//   ROW_COPY   the row is copied whole
//   ROW_RECT   the whole row lies on the sky side of the plane's horizon and inside the sphere rectangle's row range: only columns
//              rx0..rx1 (the rectangle, rounded out to 64-byte granules) can be non-black, the rest of the row is zero-filled
//   ROW_BLACK  sky side and outside the rectangle's rows: every pixel is 0x00000000 (RayTracer.cs:993) — nothing is copied
// and the zero fill of `Surface.pixels` is done by a small pool of library threads while the copies are in flight (FillPool).
// Soundness: a pixel is skipped only if the gate predicate holds for that very pixel. The kernel evaluates the same predicate per
// 4-pixel span; here it is evaluated per row at x = 0 and x = w-1 with the kernel's own two FMAs — fma(bx, x, base) is monotone in x
// under rounding, so "positive at both ends" covers the row. 35 % of the bench frame never crosses PCIe.
enum : uint8_t { ROW_COPY = 0, ROW_RECT = 1, ROW_BLACK = 2 };
struct RowPlan {
    std::vector<uint8_t> kind; int rx0 = 0, rx1 = -1; bool sparse = false;
};
inline void plan_rows(const FrameGates& g, int w, int h, RowPlan* rp) {
    rp->kind.assign((size_t)h, (uint8_t)ROW_COPY); rp->sparse = false;
    const bool rect_empty = g.spheres.x1 < g.spheres.x0 || g.spheres.y1 < g.spheres.y0 || g.spheres.x0 >= w || g.spheres.x1 < 0;
    const int rx0 = rect_empty ? 0 : ((g.spheres.x0 < 0 ? 0 : g.spheres.x0) & ~15);
    int rx1 = rect_empty ? -1 : (g.spheres.x1 | 15); if (rx1 > w - 1) rx1 = w - 1;
    // a strided (2-D) copy must save enough bytes to beat the whole-row copy: rows whose rectangle covers more than `max_frac`
    // of the width are copied whole (RTB200_RECT_MAX_FRAC, percent; measured on B200: profiles/r02/)
    static const int max_frac = [] { const char* e = getenv("RTB200_RECT_MAX_FRAC"); return e ? atoi(e) : 90; }();
    const bool rect_wide = !rect_empty && (long long)(rx1 - rx0 + 1) * 100 > (long long)w * max_frac;
    rp->rx0 = rx0; rp->rx1 = rx1;
    const float fx1 = (float)(w - 1);
    for (int y = 0; y < h; y++) {
        const float base = rt_fmaf(g.sky.by, (float)y, g.sky.a);                                  // span_positive(), rt_gate.cuh
        if (!(rt_fmaf(g.sky.bx, 0.0f, base) > 0.0f && rt_fmaf(g.sky.bx, fx1, base) > 0.0f)) continue;
        uint8_t k = ROW_BLACK;
        if (!(rect_empty || y < g.spheres.y0 || y > g.spheres.y1)) k = rect_wide ? ROW_COPY : ROW_RECT;
        rp->kind[(size_t)y] = k;
        if (k != ROW_COPY) rp->sparse = true;
    }
}

}  // namespace rtb
