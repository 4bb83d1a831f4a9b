// rt_gate.cuh — host-side, per frame: the pixel rectangle outside of which NO primary ray can be reported as hitting ANY sphere by
// the reference's test (RayTracer.cs:613-642), so the kernel may skip the sphere loop of those primary rays (:975-981) and keep
// the exact result `no sphere` (the plane loop still runs). Only conservativeness matters here, none of this is reference
// arithmetic; everything is evaluated in double.
//
// Why a rectangle exists.  All primary rays start at the camera position P.  The reference's fp32 discriminant can only be >= 0
// for a ray whose exact line passes within R' = sqrt(r^2 + K^2 (|oc|^2 + r^2)) of the centre (K = RT_BVH_PAD_K, the bound the
// LBVH boxes use — derivation in rt_lbvh.cuh / DESIGN.md §5), and a sphere wholly behind the camera plane gives b >= 0.  The
// directions that pass within R' of a centre form a cone; its projection on the view plane is bounded, per axis, by the two planes
// through the camera's other axis that touch the ball: with the centre at (X, Y, Z) in camera coordinates the slopes k = x/z of
// those planes solve  k^2 (Z^2 - R^2) - 2 X Z k + (X^2 - R^2) = 0.
// What the fp32 pipeline adds.  The direction the kernel really uses is normalize(fl(vp - P)) with vp accumulated in fp32
// (:964-971); its angle to the ideal direction of the pixel is bounded by eps below (rounding of u, v, of the three scaled basis
// vectors and of the running sum, which scales with |P|), plus 4e-5 for a basis that is orthonormal only to 1e-5. A ray that is
// off by eps passes at most eps (|oc| + R') further from the centre, so R' grows by that much. Then +-2 pixels.
// Anything unusual (basis not orthonormal, non-positive view-plane sizes, camera inside an inflated ball, a ball crossing the
// camera plane, eps > 1e-2, non-finite numbers) returns the full frame: the gate then never skips anything.
#pragma once
#include <cmath>
#include <cstdint>

#include "rt_lbvh.cuh"

namespace rtb {

struct GateRect { int x0, y0, x1, y1; };      // inclusive pixel ranges; empty (no primary ray can hit a sphere) = {w, h, w, h}

// Second gate, same spirit: the side of the (single) plane's horizon on which no primary ray can hit the plane. All primary rays
// share the origin, so the numerator of IntersectPlane (:591-594) is one number per frame (evaluated here with the kernel's own
// fp32 expression, plane_num); `t = num / den > 0` (:598) then needs den = Dot(direction, normal) to have num's sign, and
// den(x, y) is — up to the direction error eps and the rounding of the dot product — an affine function of the pixel
// coordinates. sky(x, y) = fma(gx, x, fma(gy, y, ga)) > 0  ==>  the plane cannot be hit by pixel (x, y)'s primary ray.
// A pixel that is outside the sphere rectangle AND on the sky side hits nothing: its colour is 0x00000000 (:993, :1000) without
// tracing. More than one plane, or anything unusual: ga = -1, gx = gy = 0 (never skips). No plane or num == 0: ga = +1.
struct SkyGate { float ga, gx, gy; };

inline GateRect gate_full(int w, int h) { GateRect g = {0, 0, w - 1, h - 1}; return g; }

// Bound on the angle between the fp32 primary direction of any pixel and its ideal direction (header), or -1 when the camera is
// not one the gates are derived for.
inline double primary_dir_eps(const CamRec& cam) {
    const double P[3] = {cam.pos.x, cam.pos.y, cam.pos.z};
    const double R[3] = {cam.right.x, cam.right.y, cam.right.z}, U[3] = {cam.up.x, cam.up.y, cam.up.z}, F[3] = {cam.fwd.x, cam.fwd.y, cam.fwd.z};
    const double pw = cam.view.x, ph = cam.view.y, nearp = cam.view.z;
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    auto finite3 = [](const double* a) { return std::isfinite(a[0]) && std::isfinite(a[1]) && std::isfinite(a[2]); };
    if (!finite3(P) || !finite3(R) || !finite3(U) || !finite3(F) || !std::isfinite(pw) || !std::isfinite(ph) || !std::isfinite(nearp)) return -1;
    if (!(pw > 1e-6) || !(ph > 1e-6) || !(nearp > 1e-6) || pw > 1e6 || ph > 1e6 || nearp > 1e6) return -1;
    const double tol = 1e-5;
    if (std::fabs(dot(R, R) - 1) > tol || std::fabs(dot(U, U) - 1) > tol || std::fabs(dot(F, F) - 1) > tol ||
        std::fabs(dot(R, U)) > tol || std::fabs(dot(R, F)) > tol || std::fabs(dot(U, F)) > tol) return -1;
    // rounding of u, v, of the three scaled basis vectors and of the running sum (scales with |P|), with a factor 2 in hand
    const double u32 = 5.9604644775390625e-08;           // 2^-24
    const double pmax = std::fmax(std::fabs(P[0]), std::fmax(std::fabs(P[1]), std::fabs(P[2])));
    const double L = 0.5 * pw + 0.5 * ph + nearp;
    const double E = u32 * (8.0 * (pmax + L) + 4.0 * (pw + ph));
    const double eps = 2.0 * (2.0 * std::sqrt(3.0) * E / nearp + 1e-6) + 4e-5;
    return eps < 1e-2 ? eps : -1;
}

inline GateRect primary_gate_rect(const CamRec& cam, int w, int h, const f4* sgeom, int ns) {
    const GateRect full = gate_full(w, h);
    const GateRect empty = {w, h, w, h};
    GateRect out = {w, h, -1, -1};
    if (ns <= 0) return empty;
    const double eps = primary_dir_eps(cam);
    if (eps < 0) return full;
    const double P[3] = {cam.pos.x, cam.pos.y, cam.pos.z};
    const double R[3] = {cam.right.x, cam.right.y, cam.right.z}, U[3] = {cam.up.x, cam.up.y, cam.up.z}, F[3] = {cam.fwd.x, cam.fwd.y, cam.fwd.z};
    const double pw = cam.view.x, ph = cam.view.y, nearp = cam.view.z;
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    auto finite3 = [](const double* a) { return std::isfinite(a[0]) && std::isfinite(a[1]) && std::isfinite(a[2]); };
    const double K = (double)RT_BVH_PAD_K;
    for (int i = 0; i < ns; i++) {
        const double c[3] = {sgeom[i].x, sgeom[i].y, sgeom[i].z};
        const double r2 = sgeom[i].w > 0.0f ? (double)sgeom[i].w : 0.0;
        if (!finite3(c) || !std::isfinite((double)sgeom[i].w)) return full;
        const double oc[3] = {c[0] - P[0], c[1] - P[1], c[2] - P[2]};
        const double oc2 = dot(oc, oc), ocl = std::sqrt(oc2);
        const double Rp = std::sqrt(r2 + K * K * (oc2 + r2)) * (1.0 + 1e-6) + 1e-30;
        const double Rg = Rp + eps * (ocl + Rp);
        if (!(ocl > Rg * 1.001)) return full;             // camera inside (or on) the inflated ball
        const double X = dot(oc, R), Y = dot(oc, U), Z = dot(oc, F);
        if (Z < -Rg * 1.001) continue;                    // wholly behind the camera plane: b >= 0 for every primary ray
        if (!(Z > Rg * 1.001)) return full;               // crosses the camera plane: the projection is unbounded
        const double den = Z * Z - Rg * Rg;
        int lim[2][2];
        const double ctr[2] = {X, Y}, size[2] = {pw, ph};
        const int npx[2] = {w, h};
        for (int a = 0; a < 2; a++) {
            const double A = ctr[a];
            const double disc = A * A + Z * Z - Rg * Rg;  // > 0 because Z > Rg
            const double root = Rg * std::sqrt(disc);
            const double k1 = (A * Z - root) / den, k2 = (A * Z + root) / den;
            // slope k <-> pixel: (x / w - 0.5) * pw = k * near
            const double p1 = (k1 * nearp / size[a] + 0.5) * npx[a], p2 = (k2 * nearp / size[a] + 0.5) * npx[a];
            if (!std::isfinite(p1) || !std::isfinite(p2)) return full;
            const double lo = std::floor(std::fmin(p1, p2)) - 2.0, hi = std::ceil(std::fmax(p1, p2)) + 2.0;
            lim[a][0] = lo < 0 ? 0 : (lo > npx[a] ? npx[a] : (int)lo);
            lim[a][1] = hi > npx[a] - 1 ? npx[a] - 1 : (hi < -1 ? -1 : (int)hi);
        }
        if (lim[0][1] < lim[0][0] || lim[1][1] < lim[1][0]) continue;      // projects outside the frame
        if (lim[0][0] < out.x0) out.x0 = lim[0][0];
        if (lim[1][0] < out.y0) out.y0 = lim[1][0];
        if (lim[0][1] > out.x1) out.x1 = lim[0][1];
        if (lim[1][1] > out.y1) out.y1 = lim[1][1];
    }
    if (out.x1 < out.x0 || out.y1 < out.y0) return empty;
    return out;
}

// planes: PlaneRec array of the scene (n, cn as uploaded). See SkyGate.
inline SkyGate primary_sky_gate(const CamRec& cam, int w, int h, const PlaneRec* planes, int np) {
    const SkyGate never = {-1.0f, 0.0f, 0.0f}, always = {1.0f, 0.0f, 0.0f};
    if (np <= 0) return always;
    if (np > 1) return never;
    const double eps = primary_dir_eps(cam);
    if (eps < 0) return never;
    const PlaneRec& pl = planes[0];
    f4 pn; pn.x = pl.n.x; pn.y = pl.n.y; pn.z = pl.n.z; pn.w = pl.cn;
    const float num = plane_num(cam.pos, pn);            // the kernel's own fp32 value (same expression, no contraction)
    if (!(num == num)) return never;
    if (num == 0.0f) return always;                      // 0 / den is never > 0 (:598)
    const double n[3] = {pl.n.x, pl.n.y, pl.n.z};
    const double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!std::isfinite(nl) || !(nl > 0)) return never;
    const double R[3] = {cam.right.x, cam.right.y, cam.right.z}, U[3] = {cam.up.x, cam.up.y, cam.up.z}, F[3] = {cam.fwd.x, cam.fwd.y, cam.fwd.z};
    const double pw = cam.view.x, ph = cam.view.y, nearp = cam.view.z;
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    // ideal direction of pixel (x, y): D = A + x Bx + y By;  den has the sign of D . n up to the margin below
    const double Rn = dot(R, n), Un = dot(U, n), Fn = dot(F, n);
    const double a = -0.5 * pw * Rn - 0.5 * ph * Un + nearp * Fn, bx = pw * Rn / w, by = ph * Un / h;
    const double sgn = num > 0 ? -1.0 : 1.0;              // sky side: den of the opposite sign to num (or zero)
    const double u32 = 5.9604644775390625e-08;
    const double Dmax = std::sqrt(nearp * nearp + 0.25 * pw * pw + 0.25 * ph * ph) * 1.0001;
    const double margin = 2.0 * (eps + 24.0 * u32) * std::sqrt(3.0) * nl * Dmax;
    const double S = std::fabs(a) + std::fabs(bx) * w + std::fabs(by) * h;
    SkyGate g;
    g.gx = (float)(sgn * bx); g.gy = (float)(sgn * by);
    g.ga = (float)(sgn * a - margin - 16.0 * u32 * S);
    // the float conversion of ga may round up by half an ulp: step it down once more
    g.ga = std::nextafter(g.ga, -INFINITY);
    if (!std::isfinite(g.ga) || !std::isfinite(g.gx) || !std::isfinite(g.gy)) return never;
    return g;
}
RT_HD bool sky_skips(const SkyGate& g, float fx, float fy) { return rt_fmaf(g.gx, fx, rt_fmaf(g.gy, fy, g.ga)) > 0.0f; }

// true: the pixel's primary ray may skip the sphere loop
RT_HD bool gate_skips(const GateRect& g, int x, int y) {
    return (unsigned)(x - g.x0) > (unsigned)(g.x1 - g.x0) || (unsigned)(y - g.y0) > (unsigned)(g.y1 - g.y0);
}

}  // namespace rtb
