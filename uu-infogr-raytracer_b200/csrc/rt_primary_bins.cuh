// rt_primary_bins.cuh — per-frame screen-space bins for PRIMARY rays on the LBVH path (BASELINE configs[2], [3]).
//
// All primary rays of a frame start at the camera, so "which spheres can the ray of pixel (x, y) be reported to hit"
// (RayTracer.cs:975-981) is a 2-D question, exactly as the per-light shadow bins (rt_shadow_grid.cuh) are for the parallel shadow
// rays: per frame every sphere is projected to the pixel rectangle outside of which no primary ray can be reported as hitting it,
// and entered into the 8 x 8-pixel tiles the rectangle touches. A primary ray then tests the spheres of ITS tile with the
// reference's own sphere test (sphere_hit, bit-exact) and folds them as the reference does — lexicographic minimum over
// (t, original index) = strict `>` in array order (:977) — instead of walking the tree (11.5 node visits per primary ray at 100 k
// spheres). Any superset of the spheres that can report a hit gives the reference's result.
//
// The rectangle is the one the frame gates use for the tiny scenes (rt_gate.cuh: add_ball_rect, derivation (F1), (F2) there),
// restated here as host/device code: ball of the noise-inflated radius R' = sqrt(r^2 + K^2 (|oc|^2 + r^2)) (the reference's
// discriminant reports hits that far out), grown by eps (|oc| + R') for the angle eps between the fp32 primary direction
// and the ideal direction of its pixel (primary_dir_eps, computed on the host), projected per axis by its tangent planes, + 2 pixels.
// Tiles whose list would exceed PB_CAP (the horizon rows, where hundreds of spheres project into one tile) keep no list: their
// pixels traverse the LBVH as before, and so does every pixel of a frame whose camera the derivation does not cover (eps < 0), or with
// more than PB_MAX_EVERYWHERE spheres with the eye inside their inflated ball or with a rectangle of more than PB_MAX_TILES tiles —
// up to that many such spheres are tested for every pixel.
// The bins are rebuilt on the GPU whenever the camera or the frame size changes (rt_primary_bins_build.cuh: two memsets, three small
// launches, nothing read back).
#pragma once
#include <vector>

#include "rt_scene.cuh"
#include "rt_gate.cuh"

namespace rtb {

#ifndef RT_PB_MARGIN_PX
#define RT_PB_MARGIN_PX 2.0                  // overridable only to demonstrate that the margin is needed (tests/test_primary_bins.py)
#endif
constexpr int PB_TILE_SHIFT = 3;            // 8 x 8-pixel tiles: the LBVH kernels' warps are 8 x 4 pixels, so a warp reads one list
#ifndef RT_PB_CAP
#define RT_PB_CAP 32                         // measured alternatives: profiles/r02/tuning.md
#endif
constexpr int PB_CAP = RT_PB_CAP;            // longest list a tile keeps
constexpr int PB_MAX_EVERYWHERE = 8;        // spheres tested for every pixel (unbounded / huge projections)
constexpr int PB_MAX_TILES = 4096;          // a sphere touching more tiles than this counts as "everywhere"

struct PbCam {                              // the frame's pinhole in double (host-filled; eps from gate_detail::primary_dir_eps)
    double P[3], R[3], U[3], F[3], pw, ph, nearp, eps;
    int w, h, tiles_x, tiles_y;
};
struct PbHeader {                           // device memory, written by the build
    int n_everywhere;                       // may exceed PB_MAX_EVERYWHERE: then the bins are not used this frame
    int cursor;                             // device build: list entries handed out so far (k_pb_alloc)
    int pad[2];
    f4 ev_geom[PB_MAX_EVERYWHERE];
    int ev_orig[PB_MAX_EVERYWHERE];
};
struct alignas(8) PbTile { int n, start; };             // n < 0: the tile keeps no list (its pixels traverse the LBVH); else geom[start .. start + n)
struct PrimaryBinsView {
    const PbHeader* hdr;                    // nullptr: no bins (every primary ray traverses the LBVH)
    const PbTile* tiles;                    // tiles_x * tiles_y
    const f4* geom;                         // list entries: (cx, cy, cz, r^2) ...
    const int* orig;                        // ... and the sphere's original index
    int tiles_x, tiles_y;
};
// What a tile with `count` spheres and a run starting at `start` of a list array of `capacity` entries keeps (build, both twins).
RT_HD PbTile pb_tile_decide(int count, int start, int capacity) {
    PbTile t;
    t.n = (count > PB_CAP || start + count > capacity) ? -1 : count;
    t.start = start;
    return t;
}

// Step k of the walk over a sphere's tile rectangle (tw tiles wide, top-left tile (sx0, sy0), row-major inside the rectangle) -> tile
// index of the frame. The device build's lanes take the steps 32 apart (rt_primary_bins_build.cuh); tests/hostemu replays them.
RT_HD int pb_walk_tile(int k, int tw, int sx0, int sy0, int tiles_x) {
    const int ry = k / tw;
    return (sy0 + ry) * tiles_x + sx0 + (k - ry * tw);
}

RT_HD PbTile pb_load_tile(const PbTile* p) {
#if defined(__CUDA_ARCH__)
    const int2 v = __ldg(reinterpret_cast<const int2*>(p));
    PbTile t; t.n = v.x; t.start = v.y; return t;
#else
    return *p;
#endif
}
RT_HD double pb_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
RT_HD bool pb_finite(double x) { return x - x == 0.0; }

// Pixel bounds [x0, x1] x [y0, y1] (inclusive, clamped to the frame) of the directions d = (x, y, z), z > 0, from the camera whose
// half-line passes within Rg of ctr. 0: no pixel of the frame; 1: the rectangle; 2: the eye is inside the ball (or non-finite input).
// Per axis a in {x, y} the half-line's projection onto the (a, z) plane passes within Rg of the projected centre (A, Z) (projections
// do not increase distances), i.e. its slope k = a / z lies between the tangents from the origin to that circle:
//   k^2 (Z^2 - Rg^2) - 2 A Z k + (A^2 - Rg^2) <= 0,  roots k = (A Z -+ Rg sqrt(A^2 + Z^2 - Rg^2)) / (Z^2 - Rg^2).
//  * Z > Rg (the ball in front of the eye plane; rt_gate.cuh add_ball_rect): between the roots.
//  * |Z| < Rg (the ball crosses the eye plane; the origin outside the circle): the circle meets the line z = 0 on the side of A only,
//    exactly one tangent points forward (z > 0), and the forward directions of the wedge are those from that tangent to the
//    direction (sign A, 0): a half-line of slopes — k >= the larger root for A > 0, k <= the smaller root for A < 0.
//  * the origin inside the circle (A^2 + Z^2 <= Rg^2), or Z within 0.1 % of +-Rg (the leading coefficient cancels): the axis does
//    not constrain.
// Pixel coordinate of a slope: (x / w - .5) pw = k near. Margin +-2 pixels as add_ball_rect (the eps of (F2) is part of Rg).
RT_HD int pb_ball_rect(const PbCam& c, const double* ctr, double Rg, int* rx0, int* ry0, int* rx1, int* ry1) {
    const double oc[3] = {ctr[0] - c.P[0], ctr[1] - c.P[1], ctr[2] - c.P[2]};
    const double ocl = sqrt(pb_dot(oc, oc));
    if (!pb_finite(ocl) || !pb_finite(Rg)) return 2;
    if (!(ocl > Rg * 1.001)) return 2;                      // eye inside (or on) the ball
    const double X = pb_dot(oc, c.R), Y = pb_dot(oc, c.U), Z = pb_dot(oc, c.F);
    if (Z < -Rg * 1.001) return 0;                          // wholly behind the eye plane: every ray of the frame has z > 0
    const double den = Z * Z - Rg * Rg;
    const bool den_ok = fabs(den) > 2e-3 * Rg * Rg;
    int lim[2][2];
    const double ctr2[2] = {X, Y}, size[2] = {c.pw, c.ph};
    const int npx[2] = {c.w, c.h};
    for (int a = 0; a < 2; a++) {
        const double A = ctr2[a];
        const double disc = A * A + den;
        double lo = 0.0, hi = (double)npx[a] - 1.0;         // pixel bounds of the axis; full range unless constrained below
        if (den_ok && disc > 0.0) {
            const double root = Rg * sqrt(disc);
            const double k1 = (A * Z - root) / den, k2 = (A * Z + root) / den;
            const double p1 = (k1 * c.nearp / size[a] + 0.5) * npx[a], p2 = (k2 * c.nearp / size[a] + 0.5) * npx[a];
            if (pb_finite(p1) && pb_finite(p2)) {
                const double pmin = p1 < p2 ? p1 : p2, pmax = p1 > p2 ? p1 : p2;
                if (den > 0.0) { lo = floor(pmin) - RT_PB_MARGIN_PX; hi = ceil(pmax) + RT_PB_MARGIN_PX; }
                else if (A > 0.0) lo = floor(pmax) - RT_PB_MARGIN_PX;
                else if (A < 0.0) hi = ceil(pmin) + RT_PB_MARGIN_PX;
            }
        }
        lim[a][0] = lo < 0 ? 0 : (lo > npx[a] ? npx[a] : (int)lo);
        lim[a][1] = hi > npx[a] - 1 ? npx[a] - 1 : (hi < -1 ? -1 : (int)hi);
    }
    if (lim[0][1] < lim[0][0] || lim[1][1] < lim[1][0]) return 0;          // projects outside the frame
    *rx0 = lim[0][0]; *ry0 = lim[1][0]; *rx1 = lim[0][1]; *ry1 = lim[1][1];
    return 1;
}

// Tile range of sphere g for this frame. 0: in no tile (never hit by a primary ray of the frame, or a non-finite record, which can
// never pass the reference's test); 1: tiles [tx0, tx1] x [ty0, ty1]; 2: everywhere.
RT_HD int pb_sphere_tiles(const PbCam& c, f4 g, int* tx0, int* ty0, int* tx1, int* ty1) {
    const double ctr[3] = {(double)g.x, (double)g.y, (double)g.z};
    if (!(pb_finite(ctr[0]) && pb_finite(ctr[1]) && pb_finite(ctr[2]) && pb_finite((double)g.w))) return 0;
    const double r2 = g.w > 0.0f ? (double)g.w : 0.0;
    const double oc[3] = {ctr[0] - c.P[0], ctr[1] - c.P[1], ctr[2] - c.P[2]};
    const double ocl = sqrt(pb_dot(oc, oc));
    const double K = (double)BVH_PAD_K;
    const double Rp = sqrt(r2 + K * K * (ocl * ocl + r2)) * (1.0 + 1e-6) + 1e-30;        // gate_detail::inflated (F1)
    int x0, y0, x1, y1;
    const int kind = pb_ball_rect(c, ctr, Rp + c.eps * (ocl + Rp), &x0, &y0, &x1, &y1);
    if (kind != 1) return kind;
    *tx0 = x0 >> PB_TILE_SHIFT; *ty0 = y0 >> PB_TILE_SHIFT; *tx1 = x1 >> PB_TILE_SHIFT; *ty1 = y1 >> PB_TILE_SHIFT;
    if ((long long)(*tx1 - *tx0 + 1) * (long long)(*ty1 - *ty0 + 1) > PB_MAX_TILES) return 2;
    return 1;
}

// The primary fold (:975-981) for the ray of pixel (px, py) over the spheres of its tile. false: this pixel has no usable list —
// the caller traverses the LBVH.
template <class DBG>
RT_HD bool primary_bins_nearest(const PrimaryBinsView& pb, int px, int py, f3 o, f3 dir, float a2, float a4, int* sel, float* dsel, DBG& dbg) {
    if (!pb.hdr || px < 0) return false;
    const int tx = px >> PB_TILE_SHIFT, ty = py >> PB_TILE_SHIFT;
    if (tx >= pb.tiles_x || ty >= pb.tiles_y) return false;
    const PbTile tile = pb_load_tile(pb.tiles + (ty * pb.tiles_x + tx));
    if (tile.n < 0) return false;
    const int n_ev = pb.hdr->n_everywhere;
    if (n_ev > PB_MAX_EVERYWHERE) return false;
    int best = -1; float best_t = RT_INF;                  // lexicographic min over (t, original index)
    for (int k = -n_ev; k < tile.n; k++) {
        const f4 g = k < 0 ? pb.hdr->ev_geom[k + n_ev] : load_f4(pb.geom + tile.start + k);
        float t;
        if (sphere_hit(sub3(o, mk3(g.x, g.y, g.z)), dir, g.w, a2, a4, 0.0f, &t, dbg) && t > 0) {
            const int oi = k < 0 ? pb.hdr->ev_orig[k + n_ev] : pb.orig[tile.start + k];
            if (t < best_t || (t == best_t && oi < best)) { best_t = t; best = oi; }
        }
    }
    *sel = best; *dsel = best_t;
    return true;
}

// The frame's pinhole for the bins: the gates' camera in double + the (F2) direction bound (eps < 0: no bins for this camera).
inline PbCam make_pb_cam(const CamRec& cam, int w, int h) {
    const gate_detail::Cam c = gate_detail::load_cam(cam, w, h);
    PbCam p; memset(&p, 0, sizeof(p));
    for (int k = 0; k < 3; k++) { p.P[k] = c.P[k]; p.R[k] = c.R[k]; p.U[k] = c.U[k]; p.F[k] = c.F[k]; }
    p.pw = c.pw; p.ph = c.ph; p.nearp = c.nearp; p.eps = gate_detail::primary_dir_eps(c);
    p.w = w; p.h = h;
    p.tiles_x = (w + (1 << PB_TILE_SHIFT) - 1) >> PB_TILE_SHIFT; p.tiles_y = (h + (1 << PB_TILE_SHIFT) - 1) >> PB_TILE_SHIFT;
    return p;
}
// List array: room for 16 entries per sphere (a small sphere touches 1-9 tiles) or 6 per tile (near spheres cover thousands of
// tiles each), whichever is more; at least 64 Ki, at most 8 Mi entries (160 MB). Whatever does not fit falls back to the traversal
// tile by tile.
inline int pb_list_capacity(int n, long long n_tiles) {
    long long c = 16LL * n;
    if (c < 6 * n_tiles) c = 6 * n_tiles;
    if (c < (1 << 16)) c = 1 << 16;
    if (c > (8 << 20)) c = 8 << 20;
    return (int)c;
}

// ---- host build (tests/hostemu; the device build is rt_primary_bins_build.cuh) ------------------------------------------------
struct PrimaryBinsHost {
    PbHeader hdr; std::vector<PbTile> tiles; std::vector<f4> geom; std::vector<int> orig; int tiles_x = 0, tiles_y = 0; bool valid = false;
    PrimaryBinsView view() const {
        PrimaryBinsView v; memset(&v, 0, sizeof(v));
        if (!valid) return v;
        v.hdr = &hdr; v.tiles = tiles.data(); v.geom = geom.data(); v.orig = orig.data();
        v.tiles_x = tiles_x; v.tiles_y = tiles_y;
        return v;
    }
};
// `capacity`: entries of the list array (the device build's is fixed per scene size; tests pass small ones to force the fallback).
inline void primary_bins_build_host(const PbCam& c, const f4* sgeom, int n, int capacity, PrimaryBinsHost* out) {
    out->valid = false;
    memset(&out->hdr, 0, sizeof(out->hdr));
    if (c.eps < 0 || n <= 0) return;
    out->tiles_x = c.tiles_x; out->tiles_y = c.tiles_y;
    const size_t nt = (size_t)c.tiles_x * (size_t)c.tiles_y;
    std::vector<int> count(nt, 0), fill(nt, 0);
    out->tiles.assign(nt, PbTile{-1, 0});
    out->geom.assign((size_t)capacity, f4()); out->orig.assign((size_t)capacity, -1);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1)
            for (size_t t = 0; t < nt; t++) {               // k_pb_alloc: runs in tile order here, in atomic order on the device
                const int cnt = count[t] > PB_CAP ? 0 : count[t];
                out->tiles[t] = pb_tile_decide(count[t], out->hdr.cursor, capacity);
                out->hdr.cursor += cnt;
            }
        for (int i = 0; i < n; i++) {
            int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
            const int kind = pb_sphere_tiles(c, sgeom[i], &x0, &y0, &x1, &y1);
            if (kind == 0) continue;
            if (kind == 2) {
                if (pass == 0) { const int s = out->hdr.n_everywhere++; if (s < PB_MAX_EVERYWHERE) { out->hdr.ev_geom[s] = sgeom[i]; out->hdr.ev_orig[s] = i; } }
                continue;
            }
            for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
                const size_t t = (size_t)y * c.tiles_x + x;
                if (pass == 0) { count[t]++; continue; }
                const PbTile tl = out->tiles[t];
                if (tl.n < 0) continue;
                const int s = tl.start + fill[t]++;
                out->geom[(size_t)s] = sgeom[i]; out->orig[(size_t)s] = i;
            }
        }
    }
    out->valid = true;
}

// The LBVH scene policy with the frame's primary bins: primary rays of single-sample frames (begin_pixel announces the pixel) fold
// over their tile's list; everything else — secondary and shadow rays, pixels of tiles without a list, jittered samples, free
// single-ray queries — is LbvhScene.
struct LbvhBinsScene : LbvhScene {
    PrimaryBinsView pb;
    mutable int px = -1, py = -1;
    RT_HD LbvhBinsScene(const GlobalSceneData& d, const BvhView& v, const ShadowGridsView& g, const PrimaryBinsView& p) : LbvhScene(d, v, g), pb(p) {}
    RT_HD void begin_pixel(int x, int y, int spp) const { px = spp == 1 ? x : -1; py = y; }
    template <class DBG> RT_HD void nearest(f3 o, f3 d, float a2, float a4, float off, int* sel, float* t, DBG& dbg) const {
        if (off == 0.0f && primary_bins_nearest(pb, px, py, o, d, a2, a4, sel, t, dbg)) return;
        LbvhScene::nearest(o, d, a2, a4, off, sel, t, dbg);
    }
};
// render / debug loops: tell the scene policy which pixel is traced next (only policies that care define begin_pixel)
template <class SC> RT_HD auto scene_begin_pixel(const SC& sc, int x, int y, int spp, int) -> decltype(sc.begin_pixel(x, y, spp), void()) { sc.begin_pixel(x, y, spp); }
template <class SC> RT_HD void scene_begin_pixel(const SC&, int, int, int, long) {}

}  // namespace rtb
