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

// Host-side build (scene upload). sg: sphere geometry in original order; lights: n x (px, py, pz).
struct ShadowGridsHost {
    std::vector<ShadowGrid> grids; std::vector<int> cell_start; std::vector<GridPair> items; f3 lo, hi;
    bool empty() const { return grids.empty(); }
};
inline void shadow_grids_build(const std::vector<f4>& sg, const std::vector<f3>& lights, ShadowGridsHost* out) {
    out->grids.clear(); out->cell_start.clear(); out->items.clear();
    const int n = (int)sg.size(), nl = (int)lights.size();
    if (n < 2 || nl == 0) return;
    // A sphere with a non-finite centre or radiusSquared can never pass the reference's test (every comparison of :622-635 sees a NaN
    // or the wrong-signed infinity), so it is simply left out of the bins; it must not reach the cell arithmetic below, where
    // floor(NaN) -> int would index out of bounds.
    std::vector<char> ok((size_t)n, 0);
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, r2max = 0.0; int n_ok = 0;
    for (int i = 0; i < n; i++) {
        const f4& g = sg[(size_t)i];
        if (!(std::isfinite(g.x) && std::isfinite(g.y) && std::isfinite(g.z) && std::isfinite(g.w))) continue;
        ok[(size_t)i] = 1;
        const double c[3] = {g.x, g.y, g.z};
        for (int k = 0; k < 3; k++) { lo[k] = n_ok ? std::fmin(lo[k], c[k]) : c[k]; hi[k] = n_ok ? std::fmax(hi[k], c[k]) : c[k]; }
        if (g.w > r2max) r2max = g.w;
        n_ok++;
    }
    if (n_ok == 0) return;
    const double rmax = std::sqrt(r2max);
    double ext = 0.0; for (int k = 0; k < 3; k++) ext = std::fmax(ext, hi[k] - lo[k]);
    const double margin = 0.25 * ext + 4.0 * rmax + 1.0;
    for (int k = 0; k < 3; k++) { lo[k] -= margin; hi[k] += margin; }
    out->lo = mk3((float)lo[0], (float)lo[1], (float)lo[2]); out->hi = mk3((float)hi[0], (float)hi[1], (float)hi[2]);
    // largest |oc| between a query point in the (float-rounded) box and a sphere centre inside it
    const double Dg = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2])) * 1.001;
    if (!std::isfinite(Dg) || !std::isfinite(out->lo.x) || !std::isfinite(out->lo.y) || !std::isfinite(out->lo.z) ||
        !std::isfinite(out->hi.x) || !std::isfinite(out->hi.y) || !std::isfinite(out->hi.z)) return;       // float overflow of the box: no grids
    const double K2 = (double)BVH_PAD_K * (double)BVH_PAD_K;
    out->grids.resize((size_t)nl);
    for (int li = 0; li < nl; li++) {
        ShadowGrid& g = out->grids[(size_t)li];
        memset(&g, 0, sizeof(g));
        const double d[3] = {lights[(size_t)li].x, lights[(size_t)li].y, lights[(size_t)li].z};
        const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double dmax = std::fmax(std::fabs(d[0]), std::fmax(std::fabs(d[1]), std::fabs(d[2])));
        g.cell_base = (int)out->cell_start.size();
        if (!(dmax >= 1e-12 && dmax <= 1e12) || !(len > 0.0)) continue;      // same envelope as the traversal: no grid
        const double wv[3] = {d[0] / len, d[1] / len, d[2] / len};
        int ax = 0; if (std::fabs(wv[1]) < std::fabs(wv[ax])) ax = 1; if (std::fabs(wv[2]) < std::fabs(wv[ax])) ax = 2;
        double a[3] = {0, 0, 0}; a[ax] = 1.0;
        double eu[3] = {wv[1] * a[2] - wv[2] * a[1], wv[2] * a[0] - wv[0] * a[2], wv[0] * a[1] - wv[1] * a[0]};
        const double eul = std::sqrt(eu[0] * eu[0] + eu[1] * eu[1] + eu[2] * eu[2]);
        for (int k = 0; k < 3; k++) eu[k] /= eul;
        const double ev[3] = {wv[1] * eu[2] - wv[2] * eu[1], wv[2] * eu[0] - wv[0] * eu[2], wv[0] * eu[1] - wv[1] * eu[0]};
        g.eu = mk3((float)eu[0], (float)eu[1], (float)eu[2]); g.ev = mk3((float)ev[0], (float)ev[1], (float)ev[2]);
        // project with the ROUNDED basis (what the query uses); the basis need not be exactly orthonormal for conservativeness
        // as long as the slack below covers it: |eu_f - eu| <= 6e-8 per component
        std::vector<double> cs((size_t)n), ct((size_t)n), R((size_t)n);
        double smin = 1e300, smax = -1e300, tmin = 1e300, tmax = -1e300;
        for (int i = 0; i < n; i++) {
            if (!ok[(size_t)i]) { cs[(size_t)i] = ct[(size_t)i] = R[(size_t)i] = 0.0; continue; }
            const double c[3] = {sg[(size_t)i].x, sg[(size_t)i].y, sg[(size_t)i].z};
            cs[(size_t)i] = c[0] * g.eu.x + c[1] * g.eu.y + c[2] * g.eu.z;
            ct[(size_t)i] = c[0] * g.ev.x + c[1] * g.ev.y + c[2] * g.ev.z;
            const double r2 = sg[(size_t)i].w > 0 ? sg[(size_t)i].w : 0.0;
            const double cn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            R[(size_t)i] = std::sqrt(r2 + K2 * (Dg * Dg + r2)) * 1.02 + 2e-6 * (cn + Dg) + 1e-6;   // noise bound + fp32 projection slack
            smin = std::fmin(smin, cs[(size_t)i] - R[(size_t)i]); smax = std::fmax(smax, cs[(size_t)i] + R[(size_t)i]);
            tmin = std::fmin(tmin, ct[(size_t)i] - R[(size_t)i]); tmax = std::fmax(tmax, ct[(size_t)i] + R[(size_t)i]);
        }
        // ~4 cells per sphere (measured: 0.5 -> 2.37 ms, 2 -> 2.07, 8 -> 1.95 ms on configs[3]; upload 0.21 / 0.24 / 0.56 s): the discs are wide (noise pad for the whole scene diameter), so finer cells cut the list a query
        // scans (tests per query ~ density * (cell + 2R)^2) at the price of more (cell, sphere) entries
        double cells = getenv("RTB200_SG_CELLS_PER_SPHERE") ? atof(getenv("RTB200_SG_CELLS_PER_SPHERE")) : 4.0;
        int dim = (int)std::ceil(std::sqrt((double)n * cells)); if (dim < 1) dim = 1; if (dim > 2048) dim = 2048;
        if (!(std::isfinite(smin) && std::isfinite(smax) && std::isfinite(tmin) && std::isfinite(tmax))) continue;   // no grid: LBVH traversal
        double cell = std::fmax(smax - smin, tmax - tmin) / dim; if (!(cell > 1e-9)) cell = 1e-9;
        cell *= 1.0001;
        const double slack = 2e-3 * cell;                    // fp32 rounding of (s - s0) * inv_cell near a cell border
        smin -= slack; tmin -= slack;
        g.dim_s = (int)std::ceil((smax + slack - smin) / cell); if (g.dim_s < 1) g.dim_s = 1;
        g.dim_t = (int)std::ceil((tmax + slack - tmin) / cell); if (g.dim_t < 1) g.dim_t = 1;
        g.s0 = (float)smin; g.t0 = (float)tmin; g.inv_cell = (float)(1.0 / cell);
        // use the rounded origin / cell size the query uses when assigning discs to cells
        const double s0 = g.s0, t0 = g.t0, ic = g.inv_cell;
        const int ncell = g.dim_s * g.dim_t;
        std::vector<int> count((size_t)ncell + 1, 0);
        auto range = [&](double c, double r, double o, int dimk, int* i0, int* i1) {
            double a0 = std::floor((c - r - slack - o) * ic), a1 = std::floor((c + r + slack - o) * ic);
            if (!(a0 >= 0)) a0 = 0;                           // also catches NaN
            if (!(a1 <= dimk - 1)) a1 = dimk - 1;
            *i0 = (int)a0; *i1 = (int)a1;                     // an empty range (i0 > i1) inserts nothing
        };
        for (int i = 0; i < n; i++) {
            if (!ok[(size_t)i]) continue;
            int x0, x1, y0, y1; range(cs[(size_t)i], R[(size_t)i], s0, g.dim_s, &x0, &x1); range(ct[(size_t)i], R[(size_t)i], t0, g.dim_t, &y0, &y1);
            for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) count[(size_t)(y * g.dim_s + x) + 1]++;
        }
        // per-cell sphere counts -> pair counts -> offsets (in pairs)
        std::vector<int> pstart((size_t)ncell + 1, 0);
        for (int c = 0; c < ncell; c++) pstart[(size_t)c + 1] = pstart[(size_t)c] + (count[(size_t)c + 1] + 1) / 2;
        const int item_base = (int)out->items.size();
        GridPair never; for (int h = 0; h < 2; h++) { never.ncx[h] = 0.0f; never.ncy[h] = 0.0f; never.ncz[h] = 0.0f; never.nr2[h] = 1e30f; }
        out->items.resize((size_t)item_base + (size_t)pstart[(size_t)ncell], never);
        std::vector<int> fill((size_t)ncell, 0);
        for (int i = 0; i < n; i++) {
            if (!ok[(size_t)i]) continue;
            int x0, x1, y0, y1; range(cs[(size_t)i], R[(size_t)i], s0, g.dim_s, &x0, &x1); range(ct[(size_t)i], R[(size_t)i], t0, g.dim_t, &y0, &y1);
            for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
                const int c = y * g.dim_s + x, slot = fill[(size_t)c]++;
                GridPair& p = out->items[(size_t)item_base + (size_t)pstart[(size_t)c] + (size_t)(slot / 2)];
                const int hh = slot & 1;
                p.ncx[hh] = -sg[(size_t)i].x; p.ncy[hh] = -sg[(size_t)i].y; p.ncz[hh] = -sg[(size_t)i].z; p.nr2[hh] = -sg[(size_t)i].w;
            }
        }
        for (int c = 0; c <= ncell; c++) out->cell_start.push_back(item_base + pstart[(size_t)c]);
        g.valid = 1;
    }
}

}  // namespace rtb
