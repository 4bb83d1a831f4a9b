// rt_lbvh_build.cuh — device-side LBVH construction (included by rtb200.cu only).
//   k_lbvh_keys   : 30-bit Morton code of each sphere centre, key = code << 32 | index
//   cub radix sort: 64-bit keys (unique by construction, so the Karras hierarchy is well defined and deterministic)
//   k_lbvh_leaves : gather sphere geometry into sorted order, remember the original index of every leaf
//   k_lbvh_karras : one thread per internal node (Karras 2012), records parents
//   k_lbvh_refit  : bottom-up union of child boxes; the second thread to arrive at a node continues upwards
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "rt_lbvh.cuh"

namespace rtb {

__global__ void k_lbvh_keys(const f4* sgeom, int n, float bminx, float bminy, float bminz, float binvx, float binvy, float binvz,
                            uint64_t* keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float bmin[3] = {bminx, bminy, bminz}, binv[3] = {binvx, binvy, binvz};
    f4 g = sgeom[i];
    keys[i] = morton_key(g.x, g.y, g.z, bmin, binv, (uint32_t)i);
}

__global__ void k_lbvh_leaves(const uint64_t* keys, const f4* sgeom, int n, f4* sorted, int* orig) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int idx = (int)(uint32_t)(keys[j] & 0xFFFFFFFFull);
    sorted[j] = sgeom[idx];
    orig[j] = idx;
}

// parent encoding: (parent_node << 1) | side ; root's parent = -1
__global__ void k_lbvh_karras(const uint64_t* keys, int n, BvhNode* nodes, int* parent_node, int* parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int l, r;
    karras_node(keys, n, i, &l, &r);
    nodes[i].c[0] = l; nodes[i].c[1] = r; nodes[i].pad[0] = 0; nodes[i].pad[1] = 0;
    if (l >= 0) parent_node[l] = (i << 1) | 0; else parent_leaf[~l] = (i << 1) | 0;
    if (r >= 0) parent_node[r] = (i << 1) | 1; else parent_leaf[~r] = (i << 1) | 1;
    if (i == 0) parent_node[0] = -1;
}

// node words: lox[2] loy[2] loz[2] hix[2] hiy[2] hiz[2] (child `which` = the odd / even word of each pair)
__device__ __forceinline__ BvhBox load_child_box_volatile(const BvhNode* nd, int which) {
    const volatile float* p = reinterpret_cast<const volatile float*>(nd) + which;
    BvhBox b;
    b.lox = p[0]; b.loy = p[2]; b.loz = p[4]; b.hix = p[6]; b.hiy = p[8]; b.hiz = p[10];
    return b;
}
__device__ __forceinline__ void store_child_box_volatile(BvhNode* nd, int which, const BvhBox& b) {
    volatile float* p = reinterpret_cast<volatile float*>(nd) + which;
    p[0] = b.lox; p[2] = b.loy; p[4] = b.loz; p[6] = b.hix; p[8] = b.hiy; p[10] = b.hiz;
}

// Effective radius of a leaf box: the reference's test only ever sees radiusSquared (:619), so sqrt(r^2) rounded up.
__device__ __forceinline__ float effective_radius(float r2) { return sqrtf(r2 > 0.0f ? r2 : 0.0f) * 1.000001f + 1e-30f; }

// Bottom-up refit. `sgeom` != nullptr: first re-gather the leaf geometry from the (updated) original-order array.
__global__ void k_lbvh_refit(f4* sorted, const f4* sgeom, const int* orig, int n, BvhNode* nodes,
                             const int* parent_node, const int* parent_leaf, int* arrivals) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (sgeom) sorted[j] = sgeom[orig[j]];
    const f4 g = sorted[j];
    BvhBox box = sphere_box(g, effective_radius(g.w));
    int enc = parent_leaf[j];
    while (enc >= 0) {
        const int node = enc >> 1, side = enc & 1;
        store_child_box_volatile(nodes + node, side, box);
        __threadfence();
        if (atomicAdd(arrivals + node, 1) == 0) return;        // the sibling subtree is not finished yet
        __threadfence();
        box = box_union(box, load_child_box_volatile(nodes + node, side ^ 1));
        enc = parent_node[node];
    }
}

// Per-frame refit of the camera copy: leaf boxes = centre +- inflated_radius(r^2, |cam - centre|^2), unions above.
__global__ void k_lbvh_refit_cam(const f4* sorted, int n, BvhNode* nodes_cam, const int* parent_node, const int* parent_leaf,
                                 int* arrivals, float cx, float cy, float cz) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const f4 g = sorted[j];
    const float dx = cx - g.x, dy = cy - g.y, dz = cz - g.z;
    BvhBox box = sphere_box(g, inflated_radius(g.w, dx * dx + dy * dy + dz * dz));
    int enc = parent_leaf[j];
    while (enc >= 0) {
        const int node = enc >> 1, side = enc & 1;
        store_child_box_volatile(nodes_cam + node, side, box);
        __threadfence();
        if (atomicAdd(arrivals + node, 1) == 0) return;
        __threadfence();
        box = box_union(box, load_child_box_volatile(nodes_cam + node, side ^ 1));
        enc = parent_node[node];
    }
}

struct LbvhDevice {
    BvhNode* nodes = nullptr; BvhNode* nodes_cam = nullptr; f4* sorted = nullptr; int* orig = nullptr;
    int *parent_node = nullptr, *parent_leaf = nullptr, *arrivals = nullptr;
    float r2max = 0.0f; int n = 0;
    void release() {
        cudaFree(nodes); cudaFree(nodes_cam); cudaFree(sorted); cudaFree(orig); cudaFree(parent_node); cudaFree(parent_leaf); cudaFree(arrivals);
        nodes = nullptr; nodes_cam = nullptr; sorted = nullptr; orig = nullptr; parent_node = nullptr; parent_leaf = nullptr; arrivals = nullptr; n = 0;
    }
    // Refit after the sphere records changed (rt_update_spheres): same topology, new leaf geometry and boxes.
    cudaError_t refit_geometry(const f4* sgeom_dev, cudaStream_t stream) {
        if (n < 2) return cudaSuccess;
        cudaError_t e = cudaMemsetAsync(arrivals, 0, sizeof(int) * (size_t)n, stream);
        if (e != cudaSuccess) return e;
        const int B = 256, G = (n + B - 1) / B;
        k_lbvh_refit<<<G, B, 0, stream>>>(sorted, sgeom_dev, orig, n, nodes, parent_node, parent_leaf, arrivals);
        return cudaGetLastError();
    }
    // Re-inflates nodes_cam for rays starting at (cx, cy, cz). Asynchronous on `stream`.
    cudaError_t refit_for_camera(float cx, float cy, float cz, cudaStream_t stream) {
        if (n < 2) return cudaSuccess;
        cudaError_t e = cudaMemsetAsync(arrivals, 0, sizeof(int) * (size_t)n, stream);
        if (e != cudaSuccess) return e;
        const int B = 256, G = (n + B - 1) / B;
        k_lbvh_refit_cam<<<G, B, 0, stream>>>(sorted, n, nodes_cam, parent_node, parent_leaf, arrivals, cx, cy, cz);
        return cudaGetLastError();
    }
};

// Builds the LBVH on the current device. sgeom_dev: n records (cx,cy,cz,r2) in ORIGINAL order; bmin/bmax: bounds of the
// centres. Returns cudaSuccess or the failing error.
inline cudaError_t lbvh_build(const f4* sgeom_dev, int n, const float bmin[3], const float bmax[3],
                              float r2max, cudaStream_t stream, LbvhDevice* out, uint64_t* launches) {
    out->release();
    if (n < 2) return cudaSuccess;
    cudaError_t e;
    uint64_t *keys = nullptr, *keys_sorted = nullptr; int *pn = nullptr, *pl = nullptr, *arr = nullptr;
    (void)pn; (void)pl; (void)arr;
    void* tmp = nullptr; size_t tmp_bytes = 0;
#define LB_TRY(x) do { e = (x); if (e != cudaSuccess) goto fail; } while (0)
    LB_TRY(cudaMalloc(&keys, sizeof(uint64_t) * (size_t)n));
    LB_TRY(cudaMalloc(&keys_sorted, sizeof(uint64_t) * (size_t)n));
    LB_TRY(cudaMalloc(&out->parent_node, sizeof(int) * (size_t)n));
    LB_TRY(cudaMalloc(&out->parent_leaf, sizeof(int) * (size_t)n));
    LB_TRY(cudaMalloc(&out->arrivals, sizeof(int) * (size_t)n));
    pn = out->parent_node; pl = out->parent_leaf; arr = out->arrivals;
    LB_TRY(cudaMalloc(&out->nodes, sizeof(BvhNode) * (size_t)(n - 1)));
    LB_TRY(cudaMalloc(&out->nodes_cam, sizeof(BvhNode) * (size_t)(n - 1)));
    LB_TRY(cudaMalloc(&out->sorted, sizeof(f4) * (size_t)n));
    LB_TRY(cudaMalloc(&out->orig, sizeof(int) * (size_t)n));
    LB_TRY(cudaMemsetAsync(arr, 0, sizeof(int) * (size_t)n, stream));
    {
        float binv[3];
        morton_scale(bmin, bmax, binv);
        const int B = 256, G = (n + B - 1) / B;
        k_lbvh_keys<<<G, B, 0, stream>>>(sgeom_dev, n, bmin[0], bmin[1], bmin[2], binv[0], binv[1], binv[2], keys);
        LB_TRY(cudaGetLastError());
        LB_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys_sorted, n, 0, 64, stream));
        LB_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
        LB_TRY(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, keys_sorted, n, 0, 64, stream));
        k_lbvh_leaves<<<G, B, 0, stream>>>(keys_sorted, sgeom_dev, n, out->sorted, out->orig);
        LB_TRY(cudaGetLastError());
        k_lbvh_karras<<<G, B, 0, stream>>>(keys_sorted, n, out->nodes, pn, pl);
        LB_TRY(cudaGetLastError());
        k_lbvh_refit<<<G, B, 0, stream>>>(out->sorted, nullptr, out->orig, n, out->nodes, pn, pl, arr);
        LB_TRY(cudaGetLastError());
        // the camera copy shares the topology; its boxes are written by refit_for_camera before every frame
        LB_TRY(cudaMemcpyAsync(out->nodes_cam, out->nodes, sizeof(BvhNode) * (size_t)(n - 1), cudaMemcpyDeviceToDevice, stream));
        if (launches) *launches += 4;
    }
    LB_TRY(cudaStreamSynchronize(stream));
    out->n = n; out->r2max = r2max;
    e = cudaSuccess;
    goto done;
fail:
    out->release();
done:
    cudaFree(keys); cudaFree(keys_sorted); cudaFree(tmp);
#undef LB_TRY
    return e;
}

}  // namespace rtb
