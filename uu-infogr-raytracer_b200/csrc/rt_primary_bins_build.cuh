// rt_primary_bins_build.cuh — device-side construction of the per-frame primary-ray bins (included by rtb200.cu only).
//
// Rebuilt whenever the camera or the frame size changes, asynchronously on the stream of the frame, nothing read back:
//   memset            header (everywhere count, list cursor) and the per-tile count / fill counters
//   k_pb_bin<false>   one thread per sphere computes its tile rectangle (pb_sphere_tiles: the host build's code, double precision,
//                     no FMA contraction on either side) and keeps it for the fill pass; the warp then walks the rectangles of its
//                     1..32 spheres one after the other, 32 tiles at a time, counting the sphere into every tile (atomicAdd) — a
//                     sphere that covers 4 000 tiles costs its warp 125 steps, not one thread 4 000
//   k_pb_alloc        per tile with 1..PB_CAP spheres: a run of the list array (warp-aggregated atomicAdd on the cursor: the order of
//                     the runs in memory is immaterial). A tile with more spheres, or whose run would end beyond the array, keeps
//                     no list (PbTile.n = -1): its pixels traverse the LBVH.
//   k_pb_bin<true>    same walk; each (tile, sphere) takes a slot of the tile's run and stores the sphere record + original index
// The order of the spheres inside a list depends on the atomics; the query folds with the lexicographic minimum over (t, original
// index), which does not. Host twin: primary_bins_build_host (rt_primary_bins.cuh; tests/hostemu), same lists as sets; the three passes
// below are also replayed on the host with their atomics resolved in random orders (tests/hostemu/hostemu.cpp:
// primary_bins_build_device_order, tests/test_primary_bins.py::test_device_build_replay_*).
#pragma once
#include "rt_primary_bins.cuh"

namespace rtb {

// `spw` spheres per warp (a power of two, 1..32; pb_spheres_per_warp): lanes < spw each own one sphere. The walk over the rectangles
// is serial inside a warp, so small scenes spread over many warps (1 024 spheres, one per warp: the near spheres of BASELINE
// configs[2] cover 2 000+ tiles each and used to queue up 32 to a warp), large ones amortise the warp over several spheres.
// k_pb_bin<false> stores each sphere's rectangle (rects[i]; x1 < x0: none) and k_pb_bin<true> reads it back: the two passes agree by
// construction and the double-precision projection runs once per sphere and frame.
template <bool FILL>
__global__ void __launch_bounds__(256) k_pb_bin(const f4* __restrict__ sgeom, int n, int spw, const __grid_constant__ PbCam cam, PbHeader* __restrict__ hdr,
                                                int4* __restrict__ rects, int* __restrict__ count, const PbTile* __restrict__ tiles, int* __restrict__ fill,
                                                f4* __restrict__ geom, int* __restrict__ orig) {
    const int lane = (int)(threadIdx.x & 31u);
    const long long warp = (long long)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const long long i64 = warp * spw + lane;
    const bool mine = lane < spw && i64 < (long long)n;
    const int i = mine ? (int)i64 : 0;
    int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    f4 g; g.x = 0.0f; g.y = 0.0f; g.z = 0.0f; g.w = 0.0f;
    if (mine) {
        if (!FILL) {
            g = sgeom[i];
            const int kind = pb_sphere_tiles(cam, g, &x0, &y0, &x1, &y1);
            if (kind == 2) {
                const int s = atomicAdd(&hdr->n_everywhere, 1);
                if (s < PB_MAX_EVERYWHERE) { hdr->ev_geom[s] = g; hdr->ev_orig[s] = i; }
            }
            if (kind != 1) { x0 = 0; y0 = 0; x1 = -1; y1 = -1; }
            rects[i] = make_int4(x0, y0, x1, y1);
        } else {
            const int4 r = rects[i];
            x0 = r.x; y0 = r.y; x1 = r.z; y1 = r.w;
            if (x1 >= x0) g = sgeom[i];
        }
    }
    unsigned todo = __ballot_sync(0xffffffffu, x1 >= x0 && y1 >= y0);   // every thread of the warp is here (no early return above)
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
        const int sx1 = __shfl_sync(0xffffffffu, x1, src), sy1 = __shfl_sync(0xffffffffu, y1, src);
        const int si = __shfl_sync(0xffffffffu, i, src);
        f4 sg;
        sg.x = __shfl_sync(0xffffffffu, g.x, src); sg.y = __shfl_sync(0xffffffffu, g.y, src);
        sg.z = __shfl_sync(0xffffffffu, g.z, src); sg.w = __shfl_sync(0xffffffffu, g.w, src);
        const int tw = sx1 - sx0 + 1, cells = tw * (sy1 - sy0 + 1);
        // four tiles per lane and step, in phases, so that a lane has four atomics in flight (the slot of a list entry is the
        // atomic's return value: one round trip per step, not four)
        for (int k0 = lane; k0 < cells; k0 += 128) {
            int t[4], s[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = k0 + 32 * j;
                t[j] = k < cells ? pb_walk_tile(k, tw, sx0, sy0, cam.tiles_x) : -1;
            }
            if (!FILL) {
#pragma unroll
                for (int j = 0; j < 4; j++) if (t[j] >= 0) atomicAdd(count + t[j], 1);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    PbTile tl; tl.n = -1; tl.start = 0;
                    if (t[j] >= 0) tl = pb_load_tile(tiles + t[j]);
                    s[j] = tl.n >= 0 ? tl.start : -1;                     // -1: no list kept for this tile (too many spheres, or its run did not fit)
                }
#pragma unroll
                for (int j = 0; j < 4; j++) if (s[j] >= 0) s[j] += atomicAdd(fill + t[j], 1);
#pragma unroll
                for (int j = 0; j < 4; j++) if (s[j] >= 0) {
                    *reinterpret_cast<float4*>(geom + s[j]) = make_float4(sg.x, sg.y, sg.z, sg.w);      // cudaMalloc'd array of 16-byte records
                    orig[s[j]] = si;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_pb_alloc(const int* __restrict__ count, PbTile* __restrict__ tiles, int n_tiles, int capacity, PbHeader* hdr) {
    const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int lane = (int)(threadIdx.x & 31u);
    int c = 0, c_all = 0;
    if (t < n_tiles) { c_all = count[t]; c = c_all > PB_CAP ? 0 : c_all; }
    int incl = c;                                                     // inclusive prefix sum over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&hdr->cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (t < n_tiles) tiles[t] = pb_tile_decide(c_all, base + incl - c, capacity);
}

// Spheres per warp of k_pb_bin: the largest power of two that still leaves 4 096 warps (148 SMs x 8 warps x 3-4 waves), 1..32.
inline int pb_spheres_per_warp(int n) {
    int spw = 1;
    while (spw < 32 && (long long)n / (2 * spw) >= 4096) spw *= 2;
    return spw;
}

struct PrimaryBinsDevice {
    PbHeader* hdr = nullptr; int* counters = nullptr; PbTile* tiles = nullptr; f4* geom = nullptr; int* orig = nullptr; int4* rects = nullptr;
    size_t cap_counters = 0, cap_tiles = 0, cap_geom = 0, cap_orig = 0, cap_rects = 0;
    int tiles_x = 0, tiles_y = 0, capacity = 0;
    void release() {
        cudaFree(hdr); cudaFree(counters); cudaFree(tiles); cudaFree(geom); cudaFree(orig); cudaFree(rects);
        hdr = nullptr; counters = nullptr; tiles = nullptr; geom = nullptr; orig = nullptr; rects = nullptr;
        cap_counters = cap_tiles = cap_geom = cap_orig = cap_rects = 0; tiles_x = tiles_y = capacity = 0;
    }
    template <class T> static cudaError_t grow(T** p, size_t* cap, size_t need) {      // buffers only ever grow
        if (need <= *cap && *p) return cudaSuccess;
        cudaFree(*p); *p = nullptr; *cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * need);
        if (e == cudaSuccess) *cap = need;
        return e;
    }
    PrimaryBinsView view() const {
        PrimaryBinsView v; memset(&v, 0, sizeof(v));
        v.hdr = hdr; v.tiles = tiles; v.geom = geom; v.orig = orig;
        v.tiles_x = tiles_x; v.tiles_y = tiles_y;
        return v;
    }
    // Asynchronous on `stream`. cam.eps >= 0 (the caller does not build for cameras the derivation does not cover).
    cudaError_t build(const f4* sgeom_dev, int n, const PbCam& cam, cudaStream_t stream, uint64_t* launches) {
        cudaError_t e;
        const size_t nt = (size_t)cam.tiles_x * (size_t)cam.tiles_y;
        const int cap = pb_list_capacity(n, (long long)nt);
        if (!hdr && (e = cudaMalloc(reinterpret_cast<void**>(&hdr), sizeof(PbHeader))) != cudaSuccess) return e;
        if ((e = grow(&counters, &cap_counters, 2 * nt)) != cudaSuccess) return e;      // [0, nt): count, [nt, 2 nt): fill
        if ((e = grow(&tiles, &cap_tiles, nt)) != cudaSuccess) return e;
        if ((e = grow(&geom, &cap_geom, (size_t)cap)) != cudaSuccess) return e;
        if ((e = grow(&orig, &cap_orig, (size_t)cap)) != cudaSuccess) return e;
        if ((e = grow(&rects, &cap_rects, (size_t)n)) != cudaSuccess) return e;
        tiles_x = cam.tiles_x; tiles_y = cam.tiles_y; capacity = cap;
        if ((e = cudaMemsetAsync(hdr, 0, sizeof(PbHeader), stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(counters, 0, sizeof(int) * 2 * nt, stream)) != cudaSuccess) return e;
        const int B = 256;
        const int spw = pb_spheres_per_warp(n);
        const long long warps = ((long long)n + spw - 1) / spw;
        const unsigned Gs = (unsigned)((warps * 32 + B - 1) / B), Gt = (unsigned)((nt + B - 1) / B);
        k_pb_bin<false><<<Gs, B, 0, stream>>>(sgeom_dev, n, spw, cam, hdr, rects, counters, nullptr, nullptr, nullptr, nullptr);
        k_pb_alloc<<<Gt, B, 0, stream>>>(counters, tiles, (int)nt, cap, hdr);
        k_pb_bin<true><<<Gs, B, 0, stream>>>(sgeom_dev, n, spw, cam, hdr, rects, counters, tiles, counters + nt, geom, orig);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (launches) *launches += 3;
        return cudaSuccess;
    }
};

}  // namespace rtb
