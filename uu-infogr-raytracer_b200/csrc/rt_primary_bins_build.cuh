// rt_primary_bins_build.cuh — device-side construction of the per-frame primary-ray bins (included by rtb200.cu only).
//
// Rebuilt whenever the camera or the frame size changes, asynchronously on the stream of the frame, nothing read back:
//   memset            header (everywhere count, list cursor) and the per-tile count / fill counters
//   k_pb_bin<false>   one thread per sphere computes its tile rectangle (pb_sphere_tiles: the host build's code, double precision,
//                     no FMA contraction on either side); the warp then walks the rectangles of its 32 spheres one after the other,
//                     32 tiles at a time, counting the sphere into every tile (atomicAdd) — a sphere that covers 4 000 tiles costs
//                     its warp 125 steps, not one thread 4 000
//   k_pb_alloc        per tile with 1..PB_CAP spheres: a run of the list array (warp-aggregated atomicAdd on the cursor: the order of
//                     the runs in memory is immaterial). A tile with more spheres, or whose run would end beyond the array, keeps
//                     no list (PbTile.n = -1): its pixels traverse the LBVH.
//   k_pb_bin<true>    same walk; each (tile, sphere) takes a slot of the tile's run and stores the sphere record + original index
// The order of the spheres inside a list depends on the atomics; the query folds with the lexicographic minimum over (t, original
// index), which does not. Host twin: primary_bins_build_host (rt_primary_bins.cuh; tests/hostemu), same lists as sets.
#pragma once
#include "rt_primary_bins.cuh"

namespace rtb {

template <bool FILL>
__global__ void __launch_bounds__(256) k_pb_bin(const f4* __restrict__ sgeom, int n, const __grid_constant__ PbCam cam, PbHeader* hdr,
                                                int* count, const PbTile* __restrict__ tiles, int* fill, f4* geom, int* orig) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int lane = (int)(threadIdx.x & 31u);
    int kind = 0, x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    f4 g; g.x = 0.0f; g.y = 0.0f; g.z = 0.0f; g.w = 0.0f;
    if (i < n) {
        g = sgeom[i];
        kind = pb_sphere_tiles(cam, g, &x0, &y0, &x1, &y1);
    }
    if (!FILL && kind == 2) {
        const int s = atomicAdd(&hdr->n_everywhere, 1);
        if (s < PB_MAX_EVERYWHERE) { hdr->ev_geom[s] = g; hdr->ev_orig[s] = i; }
    }
    unsigned todo = __ballot_sync(0xffffffffu, kind == 1);          // every thread of the warp is here (no early return above)
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
        for (int k = lane; k < cells; k += 32) {
            const int ry = k / tw;
            const int t = (sy0 + ry) * cam.tiles_x + sx0 + (k - ry * tw);
            if (!FILL) {
                atomicAdd(count + t, 1);
            } else {
                const PbTile tl = tiles[t];
                if (tl.n < 0) continue;                               // no list kept for this tile (too many spheres, or its run did not fit)
                const int s = tl.start + atomicAdd(fill + t, 1);
                geom[s] = sg; orig[s] = si;
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

struct PrimaryBinsDevice {
    PbHeader* hdr = nullptr; int* counters = nullptr; PbTile* tiles = nullptr; f4* geom = nullptr; int* orig = nullptr;
    size_t cap_counters = 0, cap_tiles = 0, cap_geom = 0, cap_orig = 0;
    int tiles_x = 0, tiles_y = 0, capacity = 0;
    void release() {
        cudaFree(hdr); cudaFree(counters); cudaFree(tiles); cudaFree(geom); cudaFree(orig);
        hdr = nullptr; counters = nullptr; tiles = nullptr; geom = nullptr; orig = nullptr;
        cap_counters = cap_tiles = cap_geom = cap_orig = 0; tiles_x = tiles_y = capacity = 0;
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
        tiles_x = cam.tiles_x; tiles_y = cam.tiles_y; capacity = cap;
        if ((e = cudaMemsetAsync(hdr, 0, sizeof(PbHeader), stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(counters, 0, sizeof(int) * 2 * nt, stream)) != cudaSuccess) return e;
        const int B = 256;
        const unsigned Gs = (unsigned)((n + B - 1) / B), Gt = (unsigned)((nt + B - 1) / B);
        k_pb_bin<false><<<Gs, B, 0, stream>>>(sgeom_dev, n, cam, hdr, counters, nullptr, nullptr, nullptr, nullptr);
        k_pb_alloc<<<Gt, B, 0, stream>>>(counters, tiles, (int)nt, cap, hdr);
        k_pb_bin<true><<<Gs, B, 0, stream>>>(sgeom_dev, n, cam, hdr, counters, tiles, counters + nt, geom, orig);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (launches) *launches += 3;
        return cudaSuccess;
    }
};

}  // namespace rtb
