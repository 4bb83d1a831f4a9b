// rt_shadow_grid.cuh — per-light 2-D bins for shadow rays (used together with the LBVH on scaled scenes).
//
// The reference's shadow ray uses the light POSITION as its direction (RayTracer.cs:574): all shadow rays of one light are
// parallel, whatever the shaded point.  "Is any sphere hit along that direction" is therefore a 2-D problem: project every sphere
// along the light vector onto the perpendicular plane (a disc), bin the discs in a uniform grid, and a shadow query only has to
// run the reference's own sphere test (sphere_hit, eps = 0.001 — bit-exact) on the spheres of ONE cell instead of walking a
// tree.  The answer is a boolean OR over spheres, so any superset of the spheres that can report a hit gives the exact result.
//
// Conservativeness: a ray from P can be reported as hitting sphere i only if its line passes within
// R_i = sqrt(r^2 + K^2 (|oc|^2 + r^2)) of the centre (fp32 noise bound of the reference test, rt_lbvh.cuh / DESIGN.md §5); the
// distance between the projections of P and of the centre IS that line distance.  Discs are inserted with R_i evaluated for
// |oc| = Dg, the largest origin-to-centre distance possible for query points inside the grid's validity box `lo..hi` (the
// sphere bounds grown by a margin), plus slack for the fp32 projection and cell arithmetic of the query.  Query points outside
// the box (far floor points) take the LBVH traversal instead.
#pragma once
#include <cmath>
#include <cstdlib>
#include <vector>

#include "rt_lbvh.cuh"

namespace rtb {

struct ShadowGrid {             // one per light
    f3 eu; float s0;            // basis vector u of the plane perpendicular to the light vector; grid origin (s)
    f3 ev; float t0;            // basis vector v; grid origin (t)
    float inv_cell; int dim_s, dim_t, cell_base;   // 1 / cell size; grid dimensions; offset of this light in cell_start
    int valid, pad0, pad1, pad2;                   // 0: no grid (zero / out-of-envelope light vector) -> LBVH traversal
};
// Cell lists hold spheres two at a time in the negated pair layout of rt_scene.cuh (SpherePair semantics: -cx[2], -cy[2], -cz[2],
// -r^2[2]; an odd list is padded with a sphere that cannot pass the exact test), so that the query tests two spheres per pass of
// packed fp32 instructions (sphere_pair_bd).
struct GridPair { float ncx[2], ncy[2], ncz[2], nr2[2]; };
struct ShadowGridsView {
    const ShadowGrid* grids;    // n_lights records, or nullptr
    const int* cell_start;      // per light dim_s*dim_t + 1 offsets into items (in pairs)
    const GridPair* items;      // sphere pairs per cell
    f3 lo, hi;                  // query points inside this box may use the grids
};

// Query: true/false if decided by the grid, `*decided = false` if the caller must traverse the LBVH instead.
template <class DBG>
RT_HD bool shadow_grid_any(const ShadowGridsView& sg, int li, f3 hit, f3 lp, float a2, float a4, bool* decided, DBG& dbg) {
    *decided = false;
    if (!sg.grids || li < 0) return false;
    if (!(hit.x >= sg.lo.x && hit.x <= sg.hi.x && hit.y >= sg.lo.y && hit.y <= sg.hi.y && hit.z >= sg.lo.z && hit.z <= sg.hi.z)) return false;
    const ShadowGrid g = sg.grids[li];
    if (!g.valid) return false;
    *decided = true;
    const float fs = (dot3(g.eu, hit) - g.s0) * g.inv_cell, ft = (dot3(g.ev, hit) - g.t0) * g.inv_cell;
    if (!(fs >= 0.0f && ft >= 0.0f && fs < (float)g.dim_s && ft < (float)g.dim_t)) return false;   // no disc reaches this point
    const int cell = g.cell_base + (int)ft * g.dim_s + (int)fs;
    const int b = sg.cell_start[cell], e = sg.cell_start[cell + 1];
    bool occluded = false;
#if defined(RT_HAVE_F32X2)
    const float2 ox = make_float2(hit.x, hit.x), oy = make_float2(hit.y, hit.y), oz = make_float2(hit.z, hit.z);
    const float2 dx = make_float2(lp.x, lp.x), dy = make_float2(lp.y, lp.y), dz = make_float2(lp.z, lp.z);
    const float2 na4 = make_float2(-a4, -a4);
#endif
    for (int k = b; k < e; k++) {
        float bs[2], Ds[2];
#if defined(RT_HAVE_F32X2)
        sphere_pair_bd(sg.items[k], ox, oy, oz, dx, dy, dz, na4, bs, Ds);                   // RayTracer.cs:614-621, two spheres
        dbg.sphere_test(Ds[0] >= 0); dbg.sphere_test(Ds[1] >= 0);
#else
        const GridPair p = sg.items[k];
        for (int h = 0; h < 2; h++) {
            const f3 oc = mk3(hit.x + p.ncx[h], hit.y + p.ncy[h], hit.z + p.ncz[h]);       // o - c == o + (-c)
            bs[h] = 2 * dot3(oc, lp);
            Ds[h] = bs[h] * bs[h] - a4 * (dot3(oc, oc) + p.nr2[h]);                        // x - r^2 == x + (-r^2)
            dbg.sphere_test(Ds[h] >= 0);
        }
#endif
        for (int h = 0; h < 2; h++) {
            float t;
            if (bs[h] < 0 && Ds[h] >= 0 && sphere_root(bs[h], Ds[h], a2, 0.001f, &t)) occluded = true;   // :622-635 (:578)
        }
        if (occluded && !DBG::count_tests) break;      // boolean OR
    }
    return occluded;
}

// ---- build ---------------------------------------------------------------------------------------------------------------------
// The same geometry code serves the host build (shadow_grids_build: hostemu tests, RTB200_SG_HOST=1) and the device build
// (rt_shadow_grid_build.cuh). Everything is double precision without FMA contraction (device: -fmad=false; host: x86-64 without
// -mfma), sqrt / division IEEE on both sides, so the two builds put the same spheres into the same cells.
//
// Disc radius. A shadow ray from P along the light vector w can be reported as hitting sphere (c, r^2) only if its line passes
// within rho of c with  rho^2 <= r^2 + K^2 (|P - c|^2 + r^2)  (K = BVH_PAD_K, the fp32 noise bound of the reference's discriminant
// with head-room, rt_lbvh.cuh).  In the orthonormal frame (eu, ev, w):  |P - c|^2 = rho^2 + dw^2  with dw the distance ALONG the
// light vector, hence  rho^2 <= (r^2 (1 + K^2) + K^2 dw^2) / (1 - K^2).  dw is bounded per sphere: query points lie inside the
// validity box, and their projection lies within R0 (the bound for the whole box diagonal) of the centre's, so dw <= the extent
// along w of  box ∩ {|s - cs| <= R0, |t - ct| <= R0}  seen from the centre (sg_w_range, interval arithmetic per box axis).  For a
// flat scene under an oblique light that is (box height) / |w_y| instead of the scene diameter: on BASELINE configs[3]
// (100 k spheres over 300 x 300) the pad shrinks from 1.3 to 0.05 and a query scans ~5 spheres instead of ~40.
constexpr int SG_MAX_DIM = 4096;
struct SgBox { double lo[3], hi[3], Dg, cmax; };      // validity box (float-rounded bounds), its diagonal * 1.001, largest |corner|
struct SgLight {                                        // build-time view of one light's grid
    double eu[3], ev[3], wv[3];                         // eu / ev: the ROUNDED basis the query uses; wv: unit light vector
    double s0, t0, ic, cell, slack;                     // rounded origin / inverse cell size as the query uses them; border slack
    int dim_s, dim_t, cell_base, valid;
};
struct SgDisc { double cs, ct, R; };

RT_HD bool sg_finite(double x) { return x - x == 0.0; }
RT_HD double sg_min(double a, double b) { return a < b ? a : b; }
RT_HD double sg_max(double a, double b) { return a > b ? a : b; }
RT_HD double sg_abs(double a) { return a < 0.0 ? -a : a; }

// Range [*wlo, *whi] of the coordinate along wv over the points  s*eu + t*ev + w*wv  of the box with s in [sa, sb], t in [ta, tb]
// (a superset: per-axis interval hull). Returns false when the range is empty or unbounded (caller keeps the global bound).
RT_HD bool sg_w_range(const SgBox& bx, const SgLight& L, double sa, double sb, double ta, double tb, double* wlo, double* whi) {
    double lo = -1e300, hi = 1e300; bool bounded_lo = false, bounded_hi = false;
    for (int k = 0; k < 3; k++) {
        const double qa = sg_min(sa * L.eu[k], sb * L.eu[k]) + sg_min(ta * L.ev[k], tb * L.ev[k]);
        const double qb = sg_max(sa * L.eu[k], sb * L.eu[k]) + sg_max(ta * L.ev[k], tb * L.ev[k]);
        const double wk = L.wv[k];
        if (sg_abs(wk) < 1e-9) continue;                 // this axis does not constrain w
        const double a = (bx.lo[k] - qb) / wk, b = (bx.hi[k] - qa) / wk;
        const double l = sg_min(a, b), h = sg_max(a, b);
        if (l > lo) lo = l;
        if (h < hi) hi = h;
        bounded_lo = bounded_hi = true;
    }
    *wlo = lo; *whi = hi;
    return bounded_lo && bounded_hi && lo <= hi && sg_finite(lo) && sg_finite(hi);
}

// Projection and radius of the disc that sphere g is binned with for light L; false: the sphere is left out (non-finite record —
// it can never pass the reference's test, every comparison of RayTracer.cs:622-635 sees a NaN or the wrong-signed infinity).
RT_HD bool sg_disc(const SgBox& bx, const SgLight& L, f4 g, SgDisc* d) {
    const double c[3] = {(double)g.x, (double)g.y, (double)g.z};
    if (!(sg_finite(c[0]) && sg_finite(c[1]) && sg_finite(c[2]) && sg_finite((double)g.w))) return false;
    const double K2 = (double)BVH_PAD_K * (double)BVH_PAD_K;
    const double cs = c[0] * L.eu[0] + c[1] * L.eu[1] + c[2] * L.eu[2];
    const double ct = c[0] * L.ev[0] + c[1] * L.ev[1] + c[2] * L.ev[2];
    const double wc = c[0] * L.wv[0] + c[1] * L.wv[1] + c[2] * L.wv[2];
    const double r2 = g.w > 0.0f ? (double)g.w : 0.0;
    const double cn = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    const double proj = 2e-6 * (cn + bx.Dg) + 1e-6;                         // fp32 projection of the query point, rounded basis
    const double R0 = sqrt(r2 + K2 * (bx.Dg * bx.Dg + r2)) * 1.02 + proj;   // |P - c| <= box diagonal
    double R = R0, wlo, whi;
    if (sg_w_range(bx, L, cs - R0, cs + R0, ct - R0, ct + R0, &wlo, &whi)) {
        double dw = sg_max(sg_max(wc - wlo, whi - wc), 0.0) * 1.001 + 1e-6 * (cn + bx.Dg);   // + the basis is orthonormal to ~1e-7 only
        if (dw < bx.Dg) {
            const double R1 = sqrt((r2 * (1.0 + K2) + K2 * dw * dw) / (1.0 - K2)) * 1.02 + proj;
            if (R1 < R) R = R1;
        }
    }
    d->cs = cs; d->ct = ct; d->R = R;
    return true;
}
// Inclusive cell range covered by [c - r - slack, c + r + slack] on one axis; an empty range (i0 > i1) inserts nothing.
RT_HD void sg_range(double c, double r, double slack, double o, double ic, int dimk, int* i0, int* i1) {
    double a0 = floor((c - r - slack - o) * ic), a1 = floor((c + r + slack - o) * ic);
    if (!(a0 >= 0)) a0 = 0;                               // also catches NaN
    if (!(a1 <= dimk - 1)) a1 = dimk - 1;
    *i0 = (int)a0; *i1 = (int)a1;
}
// Does the disc reach cell (x, y) — the cell's rectangle grown by the border slack?
RT_HD bool sg_cell_touched(const SgLight& L, const SgDisc& d, int x, int y) {
    const double x0 = L.s0 + (double)x * L.cell - L.slack, x1 = L.s0 + (double)(x + 1) * L.cell + L.slack;
    const double y0 = L.t0 + (double)y * L.cell - L.slack, y1 = L.t0 + (double)(y + 1) * L.cell + L.slack;
    const double dx = d.cs < x0 ? x0 - d.cs : (d.cs > x1 ? d.cs - x1 : 0.0);
    const double dy = d.ct < y0 ? y0 - d.ct : (d.ct > y1 ? d.ct - y1 : 0.0);
    return dx * dx + dy * dy <= d.R * d.R;
}
RT_HD void sg_store_item(GridPair* items, int pair_index, int half, f4 g) {
    GridPair& p = items[pair_index];
    p.ncx[half] = -g.x; p.ncy[half] = -g.y; p.ncz[half] = -g.z; p.nr2[half] = -g.w;
}
RT_HD GridPair sg_never_pair() {                          // a pair that cannot pass the exact test (pads odd lists)
    GridPair p;
    for (int h = 0; h < 2; h++) { p.ncx[h] = 0.0f; p.ncy[h] = 0.0f; p.ncz[h] = 0.0f; p.nr2[h] = 1e30f; }
    return p;
}

// O(n) host pass over the sphere records: validity box (per-axis margins), r^2 max; then per light the basis and the grid geometry
// from the projected bounds of the centres — no per-sphere work, so the device build needs nothing from the GPU to size its grids.
struct SgSetup {
    SgBox box; f3 lo, hi;                                 // lo / hi: the float box the query compares against
    std::vector<SgLight> lights; std::vector<ShadowGrid> grids;
    int total_cells = 0; bool ok = false;
};
inline double sg_cells_per_sphere() { return getenv("RTB200_SG_CELLS_PER_SPHERE") ? atof(getenv("RTB200_SG_CELLS_PER_SPHERE")) : 4.0; }
inline void sg_setup(const f4* sg, int n, const std::vector<f3>& lights, SgSetup* out) {
    out->lights.clear(); out->grids.clear(); out->total_cells = 0; out->ok = false;
    const int nl = (int)lights.size();
    if (n < 2 || nl == 0) return;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, r2max = 0.0; int n_ok = 0;
    for (int i = 0; i < n; i++) {
        const f4& g = sg[(size_t)i];
        if (!(std::isfinite(g.x) && std::isfinite(g.y) && std::isfinite(g.z) && std::isfinite(g.w))) continue;
        const double c[3] = {g.x, g.y, g.z};
        for (int k = 0; k < 3; k++) { lo[k] = n_ok ? std::fmin(lo[k], c[k]) : c[k]; hi[k] = n_ok ? std::fmax(hi[k], c[k]) : c[k]; }
        if (g.w > r2max) r2max = g.w;
        n_ok++;
    }
    if (n_ok == 0) return;
    const double clo[3] = {lo[0], lo[1], lo[2]}, chi[3] = {hi[0], hi[1], hi[2]};     // bounds of the centres
    const double rmax = std::sqrt(r2max);
    // Per-axis margins: a thin axis (a carpet of spheres on the floor) keeps a thin box, which is what bounds dw above. Points
    // outside (far floor points; a floor far below the spheres) take the LBVH traversal — slower, equally exact.
    for (int k = 0; k < 3; k++) { const double m = 0.25 * (hi[k] - lo[k]) + 4.0 * rmax + 1.0; lo[k] -= m; hi[k] += m; }
    out->lo = mk3((float)lo[0], (float)lo[1], (float)lo[2]); out->hi = mk3((float)hi[0], (float)hi[1], (float)hi[2]);
    if (!(std::isfinite(out->lo.x) && std::isfinite(out->lo.y) && std::isfinite(out->lo.z) &&
          std::isfinite(out->hi.x) && std::isfinite(out->hi.y) && std::isfinite(out->hi.z))) return;      // float overflow of the box: no grids
    SgBox& bx = out->box;
    bx.lo[0] = out->lo.x; bx.lo[1] = out->lo.y; bx.lo[2] = out->lo.z; bx.hi[0] = out->hi.x; bx.hi[1] = out->hi.y; bx.hi[2] = out->hi.z;
    bx.Dg = std::sqrt((bx.hi[0] - bx.lo[0]) * (bx.hi[0] - bx.lo[0]) + (bx.hi[1] - bx.lo[1]) * (bx.hi[1] - bx.lo[1]) +
                      (bx.hi[2] - bx.lo[2]) * (bx.hi[2] - bx.lo[2])) * 1.001;
    bx.cmax = 0.0;
    for (int k = 0; k < 3; k++) bx.cmax += std::fmax(bx.lo[k] * bx.lo[k], bx.hi[k] * bx.hi[k]);
    bx.cmax = std::sqrt(bx.cmax);
    if (!std::isfinite(bx.Dg) || !std::isfinite(bx.cmax)) return;
    const double K2 = (double)BVH_PAD_K * (double)BVH_PAD_K;
    const double Rmax = std::sqrt(r2max + K2 * (bx.Dg * bx.Dg + r2max)) * 1.02 + 2e-6 * (bx.cmax + bx.Dg) + 1e-6;   // >= every disc radius
    out->lights.resize((size_t)nl); out->grids.resize((size_t)nl);
    const double cells = sg_cells_per_sphere();
    long long total = 0;
    for (int li = 0; li < nl; li++) {
        SgLight& L = out->lights[(size_t)li]; ShadowGrid& g = out->grids[(size_t)li];
        memset(&L, 0, sizeof(L)); memset(&g, 0, sizeof(g));
        L.cell_base = g.cell_base = (int)total;
        const double d[3] = {lights[(size_t)li].x, lights[(size_t)li].y, lights[(size_t)li].z};
        const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double dmax = std::fmax(std::fabs(d[0]), std::fmax(std::fabs(d[1]), std::fabs(d[2])));
        if (!(dmax >= 1e-12 && dmax <= 1e12) || !(len > 0.0)) continue;      // same envelope as the traversal: no grid
        const double wv[3] = {d[0] / len, d[1] / len, d[2] / len};
        int ax = 0; if (std::fabs(wv[1]) < std::fabs(wv[ax])) ax = 1; if (std::fabs(wv[2]) < std::fabs(wv[ax])) ax = 2;
        double a[3] = {0, 0, 0}; a[ax] = 1.0;
        double eu[3] = {wv[1] * a[2] - wv[2] * a[1], wv[2] * a[0] - wv[0] * a[2], wv[0] * a[1] - wv[1] * a[0]};
        const double eul = std::sqrt(eu[0] * eu[0] + eu[1] * eu[1] + eu[2] * eu[2]);
        for (int k = 0; k < 3; k++) eu[k] /= eul;
        const double ev[3] = {wv[1] * eu[2] - wv[2] * eu[1], wv[2] * eu[0] - wv[0] * eu[2], wv[0] * eu[1] - wv[1] * eu[0]};
        g.eu = mk3((float)eu[0], (float)eu[1], (float)eu[2]); g.ev = mk3((float)ev[0], (float)ev[1], (float)ev[2]);
        // project with the ROUNDED basis (what the query uses); it need not be exactly orthonormal for conservativeness as long as
        // the slacks cover it: |eu_f - eu| <= 6e-8 per component
        L.eu[0] = g.eu.x; L.eu[1] = g.eu.y; L.eu[2] = g.eu.z; L.ev[0] = g.ev.x; L.ev[1] = g.ev.y; L.ev[2] = g.ev.z;
        for (int k = 0; k < 3; k++) L.wv[k] = wv[k];
        // grid extents: the projected corners of the centres' bounds, grown by the largest disc radius (every disc lies inside)
        double smin = 1e300, smax = -1e300, tmin = 1e300, tmax = -1e300;
        for (int q = 0; q < 8; q++) {
            const double c[3] = {(q & 1) ? chi[0] : clo[0], (q & 2) ? chi[1] : clo[1], (q & 4) ? chi[2] : clo[2]};
            const double s = c[0] * L.eu[0] + c[1] * L.eu[1] + c[2] * L.eu[2], t = c[0] * L.ev[0] + c[1] * L.ev[1] + c[2] * L.ev[2];
            smin = std::fmin(smin, s); smax = std::fmax(smax, s); tmin = std::fmin(tmin, t); tmax = std::fmax(tmax, t);
        }
        smin -= Rmax; smax += Rmax; tmin -= Rmax; tmax += Rmax;
        if (!(std::isfinite(smin) && std::isfinite(smax) && std::isfinite(tmin) && std::isfinite(tmax))) continue;   // no grid: LBVH traversal
        // ~4 cells per sphere over the bounding rectangle (RTB200_SG_CELLS_PER_SPHERE): finer cells cut the list a query scans
        // (tests per query ~ density * (cell + 2R)^2) at the price of more (cell, sphere) entries
        int dim = (int)std::ceil(std::sqrt((double)n * cells)); if (dim < 1) dim = 1; if (dim > SG_MAX_DIM) dim = SG_MAX_DIM;
        double cell = std::sqrt((smax - smin) * (tmax - tmin)) / dim;
        const double cell_min = std::fmax(smax - smin, tmax - tmin) / SG_MAX_DIM;
        if (cell < cell_min) cell = cell_min;
        if (!(cell > 1e-9)) cell = 1e-9;
        cell *= 1.0001;
        const double slack = 2e-3 * cell;                    // fp32 rounding of (s - s0) * inv_cell near a cell border
        smin -= slack; tmin -= slack;
        g.dim_s = (int)std::ceil((smax + slack - smin) / cell); if (g.dim_s < 1) g.dim_s = 1;
        g.dim_t = (int)std::ceil((tmax + slack - tmin) / cell); if (g.dim_t < 1) g.dim_t = 1;
        g.s0 = (float)smin; g.t0 = (float)tmin; g.inv_cell = (float)(1.0 / cell);
        if (g.dim_s > 2 * SG_MAX_DIM || g.dim_t > 2 * SG_MAX_DIM || total + (long long)g.dim_s * g.dim_t > 0x3FFFFFFFll) continue;
        // the disc-to-cell assignment uses the rounded origin / cell size of the query
        L.s0 = g.s0; L.t0 = g.t0; L.ic = g.inv_cell; L.cell = 1.0 / L.ic; L.slack = slack;
        L.dim_s = g.dim_s; L.dim_t = g.dim_t; L.valid = g.valid = 1;
        total += (long long)g.dim_s * g.dim_t;
    }
    out->total_cells = (int)total;
    out->ok = true;
}

// Host-side build. sg: sphere geometry in original order; lights: n x (px, py, pz).
struct ShadowGridsHost {
    std::vector<ShadowGrid> grids; std::vector<int> cell_start; std::vector<GridPair> items; f3 lo, hi;
    bool empty() const { return grids.empty(); }
};
inline void shadow_grids_build(const std::vector<f4>& sg, const std::vector<f3>& lights, ShadowGridsHost* out) {
    out->grids.clear(); out->cell_start.clear(); out->items.clear();
    const int n = (int)sg.size();
    SgSetup su;
    sg_setup(sg.data(), n, lights, &su);
    if (!su.ok) return;
    out->lo = su.lo; out->hi = su.hi; out->grids = su.grids;
    const int total = su.total_cells;
    std::vector<int> count((size_t)total + 1, 0);
    auto each_cell = [&](auto&& fn) {
        for (size_t li = 0; li < su.lights.size(); li++) {
            const SgLight& L = su.lights[li];
            if (!L.valid) continue;
            for (int i = 0; i < n; i++) {
                SgDisc d;
                if (!sg_disc(su.box, L, sg[(size_t)i], &d)) continue;
                int x0, x1, y0, y1;
                sg_range(d.cs, d.R, L.slack, L.s0, L.ic, L.dim_s, &x0, &x1); sg_range(d.ct, d.R, L.slack, L.t0, L.ic, L.dim_t, &y0, &y1);
                for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++)
                    if (sg_cell_touched(L, d, x, y)) fn(L.cell_base + y * L.dim_s + x, i);
            }
        }
    };
    each_cell([&](int c, int) { count[(size_t)c]++; });
    // per-cell sphere counts -> pair counts -> offsets (in pairs); one offset array over the cells of all lights
    out->cell_start.assign((size_t)total + 1, 0);
    for (int c = 0; c < total; c++) out->cell_start[(size_t)c + 1] = out->cell_start[(size_t)c] + (count[(size_t)c] + 1) / 2;
    out->items.assign((size_t)out->cell_start[(size_t)total], sg_never_pair());
    std::fill(count.begin(), count.end(), 0);
    each_cell([&](int c, int i) {
        const int slot = count[(size_t)c]++;
        sg_store_item(out->items.data(), out->cell_start[(size_t)c] + slot / 2, slot & 1, sg[(size_t)i]);
    });
}

}  // namespace rtb
