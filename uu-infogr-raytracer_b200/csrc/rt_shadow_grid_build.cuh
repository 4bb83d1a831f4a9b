// rt_shadow_grid_build.cuh — device-side construction of the per-light shadow bins (included by rtb200.cu only).
//
// The host sizes the grids from the bounds of the sphere centres alone (sg_setup, rt_shadow_grid.cuh: O(n) min/max, no
// per-sphere projection), so nothing has to come back from the GPU before the bins can be laid out:
//   k_sg_bin<false> : one thread per (sphere, light) — disc of the sphere for that light (sg_disc), count it into every cell it
//                     touches (circle / rectangle test per candidate cell, atomicAdd)
//   k_sg_pairs      : per-cell sphere count -> pair count (the query tests two spheres per pass of packed fp32)
//   cub scan        : exclusive prefix sum over the cells of all lights -> cell_start (in pairs); its last element, the total
//                     number of pairs, is the one value read back (4 bytes) to size the item array
//   k_sg_never      : pre-fill the item array with a pair that cannot pass the exact test (pads odd lists)
//   k_sg_bin<true>  : same walk, each (cell, sphere) takes a slot with atomicAdd and stores the negated sphere record
// The order of the spheres inside a cell depends on the atomics; the query is a boolean OR over the cell (RayTracer.cs:577-581),
// so its result does not. The geometry code is the host build's (double precision, no FMA contraction on either side): both
// put the same spheres into the same cells (tests/test_gpu_lbvh.py::test_device_shadow_bins_equal_host_bins).
// Replaces a host build that took 0.2-0.6 s at 100 k spheres and was re-run by every rt_update_spheres.
#pragma once
#include <cub/device/device_scan.cuh>

#include "rt_shadow_grid.cuh"

namespace rtb {

template <bool FILL>
__global__ void __launch_bounds__(256) k_sg_bin(const f4* __restrict__ sgeom, int n, const SgLight* __restrict__ lights, int nl,
                                                const __grid_constant__ SgBox box, int* counters, const int* __restrict__ cell_start,
                                                GridPair* items) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * nl) return;
    const int li = (int)(idx / n), i = (int)(idx - (long long)li * n);
    const SgLight L = lights[li];
    if (!L.valid) return;
    const f4 g = sgeom[i];
    SgDisc d;
    if (!sg_disc(box, L, g, &d)) return;
    int x0, x1, y0, y1;
    sg_range(d.cs, d.R, L.slack, L.s0, L.ic, L.dim_s, &x0, &x1);
    sg_range(d.ct, d.R, L.slack, L.t0, L.ic, L.dim_t, &y0, &y1);
    for (int y = y0; y <= y1; y++)
        for (int x = x0; x <= x1; x++) {
            if (!sg_cell_touched(L, d, x, y)) continue;
            const int c = L.cell_base + y * L.dim_s + x;
            const int slot = atomicAdd(counters + c, 1);
            if (FILL) sg_store_item(items, cell_start[c] + (slot >> 1), slot & 1, g);
        }
}
__global__ void k_sg_pairs(const int* __restrict__ count, int total, int* __restrict__ pairs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= total) pairs[c] = c < total ? (count[c] + 1) >> 1 : 0;
}
__global__ void k_sg_never(GridPair* items, int n_pairs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_pairs) items[k] = sg_never_pair();
}

struct ShadowGridsDevice {
    ShadowGrid* grids = nullptr; int* cells = nullptr; GridPair* items = nullptr;     // what the query reads (ShadowGridsView)
    SgLight* lights = nullptr; int* counters = nullptr; int* pairs = nullptr; void* scan_tmp = nullptr;
    size_t cap_grids = 0, cap_lights = 0, cap_cells = 0, cap_counters = 0, cap_pairs = 0, cap_items = 0, cap_tmp = 0;
    int n_pairs = 0;
    void release() {
        cudaFree(grids); cudaFree(cells); cudaFree(items); cudaFree(lights); cudaFree(counters); cudaFree(pairs); cudaFree(scan_tmp);
        grids = nullptr; cells = nullptr; items = nullptr; lights = nullptr; counters = nullptr; pairs = nullptr; scan_tmp = nullptr;
        cap_grids = cap_lights = cap_cells = cap_counters = cap_pairs = cap_items = cap_tmp = 0; n_pairs = 0;
    }
    // buffers only ever grow (a rebuild after rt_update_spheres re-uses them: no allocation in the steady state)
    template <class T> static cudaError_t grow(T** p, size_t* cap, size_t need) {
        if (need <= *cap && *p) return cudaSuccess;
        cudaFree(*p); *p = nullptr; *cap = 0;
        const size_t want = need + need / 4 + 16;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * want);
        if (e == cudaSuccess) *cap = want;
        return e;
    }
    // Host-built bins (RTB200_SG_HOST=1, tests): plain upload.
    cudaError_t upload(const ShadowGridsHost& h, cudaStream_t stream) {
        cudaError_t e;
        if ((e = grow(&grids, &cap_grids, h.grids.size())) != cudaSuccess) return e;
        if ((e = grow(&cells, &cap_cells, h.cell_start.size())) != cudaSuccess) return e;
        if ((e = grow(&items, &cap_items, h.items.size() + 1)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(grids, h.grids.data(), sizeof(ShadowGrid) * h.grids.size(), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(cells, h.cell_start.data(), sizeof(int) * h.cell_start.size(), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if (!h.items.empty() &&
            (e = cudaMemcpyAsync(items, h.items.data(), sizeof(GridPair) * h.items.size(), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        n_pairs = (int)h.items.size();
        return cudaStreamSynchronize(stream);
    }
    // Device build from the sphere records already on this device (original order) and the host's grid layout.
    cudaError_t build(const f4* sgeom_dev, int n, const SgSetup& su, cudaStream_t stream, uint64_t* launches) {
        cudaError_t e;
        const int nl = (int)su.lights.size(), total = su.total_cells;
        if ((e = grow(&grids, &cap_grids, (size_t)nl)) != cudaSuccess) return e;
        if ((e = grow(&lights, &cap_lights, (size_t)nl)) != cudaSuccess) return e;
        if ((e = grow(&cells, &cap_cells, (size_t)total + 1)) != cudaSuccess) return e;
        if ((e = grow(&counters, &cap_counters, (size_t)total + 1)) != cudaSuccess) return e;
        if ((e = grow(&pairs, &cap_pairs, (size_t)total + 1)) != cudaSuccess) return e;
        size_t tmp_bytes = 0;
        if ((e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pairs, cells, total + 1, stream)) != cudaSuccess) return e;
        if (tmp_bytes > cap_tmp || !scan_tmp) {
            cudaFree(scan_tmp); scan_tmp = nullptr; cap_tmp = 0;
            if ((e = cudaMalloc(&scan_tmp, tmp_bytes + 256)) != cudaSuccess) return e;
            cap_tmp = tmp_bytes + 256;
        }
        if ((e = cudaMemcpyAsync(grids, su.grids.data(), sizeof(ShadowGrid) * (size_t)nl, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(lights, su.lights.data(), sizeof(SgLight) * (size_t)nl, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(counters, 0, sizeof(int) * ((size_t)total + 1), stream)) != cudaSuccess) return e;
        const long long work = (long long)n * nl;
        const int B = 256; const unsigned G = (unsigned)((work + B - 1) / B), Gc = (unsigned)((total + 1 + B - 1) / B);
        k_sg_bin<false><<<G, B, 0, stream>>>(sgeom_dev, n, lights, nl, su.box, counters, nullptr, nullptr);
        k_sg_pairs<<<Gc, B, 0, stream>>>(counters, total, pairs);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        size_t tb = cap_tmp;
        if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, tb, pairs, cells, total + 1, stream)) != cudaSuccess) return e;
        int np = 0;
        if ((e = cudaMemcpyAsync(&np, cells + total, sizeof(int), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;        // the one read-back: total number of pairs
        if (np < 0) return cudaErrorUnknown;
        if ((e = grow(&items, &cap_items, (size_t)np + 1)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(counters, 0, sizeof(int) * ((size_t)total + 1), stream)) != cudaSuccess) return e;
        if (np) k_sg_never<<<(unsigned)((np + B - 1) / B), B, 0, stream>>>(items, np);
        k_sg_bin<true><<<G, B, 0, stream>>>(sgeom_dev, n, lights, nl, su.box, counters, cells, items);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        n_pairs = np;
        if (launches) *launches += np ? 5 : 4;
        return cudaStreamSynchronize(stream);
    }
};

}  // namespace rtb
