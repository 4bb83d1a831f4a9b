// rtb200.cu — render kernels and the C ABI (include/rtb200.h) of the B200 ray-cast backend.
//
// Replaces the pixel loop of RayTracer.Tick() (Raytracer/RayTracer.cs:898-901) and everything it calls
// (TracePixel :962-1002, TraceSphere :835-876, TracePlane :729-780, TraceSecondaryRay :789-826,
// IntersectsSphere :613-642, IntersectPlane :590-604, IntersectShadowLight :573-582, ShapePhongShading :665-695,
// ShiftColor/SetPixel :1037-1052).  Strict fp32 (see rt_math.cuh); no tensor cores (no dense contraction on this path).
//
// Work decomposition: a frame is cut into ROW TILES of `tile_rows` rows (contiguous pixel ranges in the row-major
// framebuffer); tile t belongs to rank t % world (multi-GPU row interleave).  A rank's tiles are cut into work items of
// 128 threads x PPT pixels, one CTA per (item, tile, frame).  Each thread owns a span of PPT=4 adjacent pixels and writes
// it with one 128-bit store; the store address may be a peer (NVLink) mapping of rank 0's framebuffer — the gather is
// fused into the render kernel — or the caller's page-locked Surface.pixels.  Items are 2-D pixel blocks (rt_tiles.cuh: a warp
// covers 32 x 4 pixels, 8 x 4 on the one-pixel-per-thread paths) whenever tile_rows % 8 == 0 and a row is whole spans; linear
// runs of the pixel index otherwise.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
// cuda_gl_interop.h needs <GL/gl.h>, which a headless build image does not have; the one entry point used is declared here
// (GLuint is `unsigned int`; the symbol lives in the CUDA runtime).
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource, unsigned int buffer, unsigned int flags);

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "rt_scene.cuh"
#include "rt_lbvh_build.cuh"
#include "rt_shadow_grid_build.cuh"
#include "rt_primary_bins_build.cuh"
#include "rt_gate.cuh"
#include "rt_gather.cuh"
#include "rt_tiles.cuh"

using namespace rtb;

namespace {

#ifndef RT_BLOCK
#define RT_BLOCK 128              // 128-thread CTAs (512 / 128 pixels per CTA): +2.4 % over 256 on B200 (profiles/r01/tuning.md)
#endif
#ifndef RT_PPT_TINY
#define RT_PPT_TINY 4
#endif
constexpr int BLOCK = RT_BLOCK;
constexpr int PPT_TINY = RT_PPT_TINY;       // pixels per thread on the tiny-scene path: one 128-bit store per thread
constexpr int PPT_HEAVY = 1;                // staged / global / LBVH paths: a pixel costs 10-1000x more, so short CTAs (256
                                            // pixels) keep the slowest CTA off the critical path (multi-GPU: 2.4 -> see tuning.md)
constexpr int STACK_RECS = RT_MAX_DEPTH + 1;
constexpr int INLINE_CAMS = 16;
constexpr int N_DEBUG_COUNTERS = 16;
#ifndef RT_PRIMARY_BINS_DEFAULT
#define RT_PRIMARY_BINS_DEFAULT 1      // RT_OPT_PRIMARY_BINS of a new context (environment RTB200_PRIMARY_BINS overrides it at rt_create)
#endif
#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 8      // resident CTAs per SM the register allocator must allow (tuned on B200, profiles/)
#endif

// Zero-copy host return: the render kernel stores straight into the caller's page-locked Surface.pixels over PCIe, and the pixels
// the frame gates prove black are zero-filled by host threads instead of being stored. The host's fill set (plan_rows, rt_gate.cuh)
// is passed down in this form: rows [yb0, yb1) and [yc0, yc1) are filled whole, rows [yr0, yr1) outside the columns rx0..rx1 (the largest
// runs of the plan; whatever does not fit this shape is simply stored by the kernel). Empty ranges = nothing.
struct HostSkip { int yb0, yb1, yc0, yc1, yr0, yr1, rx0, rx1; };     // two whole-row ranges, one rectangle-rows range
__host__ __device__ __forceinline__ bool host_fills_span(const HostSkip& k, int xa, int xb, int y) {
    return (y >= k.yb0 && y < k.yb1) || (y >= k.yc0 && y < k.yc1) || (y >= k.yr0 && y < k.yr1 && (xb < k.rx0 || xa > k.rx1));
}

constexpr int PART_MAX_PERIOD = 64, PART_MAX_CNT = 16;
struct FrameParams {
    int w, h, cap, spp;
    uint32_t seed;
    int n_frames;
    int rank, world, tile_rows;
    // Row-tile partition: tiles are dealt out in PERIODS of part_period tiles; slot j of a period belongs to rank part_owner[j]; this
    // rank owns part_cnt slots per period, part_pos[0..part_cnt). Equal shares: period = world, owner[j] = j (tile t -> rank t % world).
    // Weighted (packed gather): GPU 0, the sink that also runs the expand pass, gets fewer slots than the others (PartTable).
    int part_period, part_cnt;
    unsigned char part_pos[PART_MAX_CNT];
    unsigned char part_owner[PART_MAX_PERIOD];
    GatherParams gather;        // packed multi-GPU gather (rt_gather.cuh); area == nullptr: plain stores into `out`
    int tiles_total;            // ceil(h / tile_rows)
    int tiles_mine;             // tiles of this rank rendered by this launch ...
    int k_begin;                // ... starting at the rank's k_begin-th tile (band pipelining, see render_frames)
    int chunks_per_tile;        // ceil(tile_rows * w / (BLOCK * pixels per thread)); tile2d: (tile_rows / 8) * ceil(w / 16)
    int tile2d;                 // heavy paths (one pixel per thread), tile_rows a multiple of 8: a CTA is 16 x 8 pixels, a warp 8 x 4
                                // (coherent rays for the LBVH traversal) instead of 128 / 32 consecutive pixels of a row
    int tile_rot;               // this rank's tiles are visited starting at its tile_rot-th one, wrapping around (launch_render: tail of the launch)
    int skip_black_store;       // sparse gather (launch_render): this rank does not store proven-black spans, rank 0 fills them locally
    float rcp_w, rcp_h;         // 1/w, 1/h correctly rounded (host): pixel-coordinate divisions of the single-sample kernels (rt_div_rcp)
    long long frame_stride;     // pixels between consecutive frames in `out`
    uint32_t* out;              // framebuffer(s): 0x00RRGGBB, row-major (Surface.pixels, surface.cs:9-20)
    HostSkip hskip[INLINE_CAMS];      // per frame: pixels the HOST zero-fills itself (zero-copy return, render_frames): not stored
    FrameGates gates[INLINE_CAMS];    // per frame: what the host proved pixel regions cannot hit (rt_gate.cuh; tiny single-sample kernels)
    CamRec cam_inline[INLINE_CAMS];   // the launch's cameras travel in the parameter block (constant bank): no upload, no host
                                      // sync; batches of more than INLINE_CAMS frames are split into several launches
};

struct DebugOut {
    uint32_t* hash; int32_t* aov_id; float* aov_t; unsigned long long* counters;
};

// this rank's k-th tile
__device__ __forceinline__ int tile_of(const FrameParams& fp, int k) {
    if (fp.gather.area == nullptr) return k * fp.world + fp.rank;            // equal shares (any world size)
    if (fp.part_cnt == 1) return k * fp.part_period + fp.part_pos[0];
    const int q = k / fp.part_cnt;
    return q * fp.part_period + fp.part_pos[k - q * fp.part_cnt];
}


// One CTA per work item: blockIdx = (chunk inside the tile, this rank's tile, frame) — no index divisions, and the hardware
// block scheduler balances sky / floor / mirror chunks (measured 8 % faster than a one-wave persistent grid-stride loop,
// profiles/r01/tuning.md). gridDim.y is folded when a launch has more than 65535 tiles.
// DBG = NoDbg: the shipped kernels. DBG = ProdDbg (k_debug_tiny_prod): the SAME loop, gates and trace instantiation, additionally
// writing the per-pixel chain hash / primary AOV / ray counters of frame 0 to `dout` — the certificate that the gated, packed,
// exact-count production path takes the reference's hits (tests/test_gpu_parity.py::test_shipped_kernel_chain_hashes).
template <int PPT, bool SPP1 = false, class DBG = NoDbg, class SC>
__device__ __forceinline__ void render_loop(const SC& sc, const FrameParams& fp, const DebugOut* dout = nullptr) {
    constexpr int CHUNK = BLOCK * PPT;          // pixels per CTA work item
    HitRec stack[STACK_RECS];
    NoDbg nodbg;
    unsigned long long n_primary = 0, n_shadow = 0, n_secondary = 0;     // DBG::enabled only
    const int npix = fp.w * fp.h;
    const int tile_pix = fp.tile_rows * fp.w;
    const int frame = blockIdx.z;
    for (int kk = blockIdx.y; kk < fp.tiles_mine; kk += gridDim.y) {
        int kr = kk + fp.tile_rot; if (kr >= fp.tiles_mine) kr -= fp.tiles_mine;
        const int k = fp.k_begin + kr;                         // my k-th tile
        const int tile = tile_of(fp, k);
        const int base = tile * tile_pix;
        int end = base + tile_pix; if (end > npix || end < base) end = npix;
        int p0, x, y;
        if ((PPT == 1 || PPT == 4) && fp.tile2d) {
            // 2-D pixel blocks (rt_tiles.cuh): blockIdx.x = (row group of 8 rows inside the tile, block of 16 * PPT columns); a warp
            // covers 8 * PPT x 4 pixels (8 threads side by side, 4 rows), the CTA's four warps 2 x 2 of those
            if (!tile2d_span(PPT, fp.w, fp.h, tile, fp.tile_rows, (int)blockIdx.x, (int)threadIdx.x, &x, &y)) continue;
            p0 = y * fp.w + x;
            end = p0 + (fp.w - x < PPT ? fp.w - x : PPT);      // the thread's span ends with its row
        } else {
            p0 = base + (int)blockIdx.x * CHUNK + (int)threadIdx.x * PPT;
            if (p0 >= end) continue;
            // row of the span: frames below 2^24 pixels without an integer division (p0 is exact in fp32 and 1/w is correctly
            // rounded, so the estimate is off by one at most; as in k_gather_expand); then step along the row
            if (npix < (1 << 24)) {
                y = (int)((float)p0 * fp.rcp_w);
                if (y * fp.w > p0) y--; else if ((y + 1) * fp.w <= p0) y++;
            } else {
                y = p0 / fp.w;
            }
            x = p0 - y * fp.w;
        }
        const CamRec& cam = fp.cam_inline[frame];               // constant bank (LDC); batches > INLINE_CAMS are split on the host
        const FrameGates& gates = fp.gates[frame];
        uint32_t* out = fp.out + (long long)frame * fp.frame_stride;
        uint32_t px[PPT];
        // What the host proved about this thread's span of pixels (rt_gate.cuh): evaluated once, for spans inside one row.
        uint32_t bits = 0u;
        if (SPP1 && x + PPT <= fp.w && p0 + PPT <= end) {
            bool black = false;
            bits = gate_bits_span(gates, x, x + PPT - 1, y, sc.n_lights(), &black);
            if (black) {                                       // nothing can be hit anywhere in the span: 0x00000000 (:993) untraced
                if constexpr (DBG::enabled) {                  // what the reference's primary ray does there: one ray, no hit
                    for (int q = 0; q < PPT; q++) {
                        if (dout->hash) dout->hash[p0 + q] = event_hash(0u, 1u, 0xFFFFFFFFu, 0u);
                        if (dout->aov_id) dout->aov_id[p0 + q] = -1;
                        if (dout->aov_t) dout->aov_t[p0 + q] = 0.0f;
                    }
                    n_primary += PPT;
                }
                if (fp.skip_black_store) continue;             // ... and rank 0 writes those zeros itself (k_fill_black): nothing crosses NVLink
                if (host_fills_span(fp.hskip[frame], x, x + PPT - 1, y)) continue;   // ... or the host does (zero-copy return): nothing crosses PCIe
                if (PPT == 4 && ((reinterpret_cast<uintptr_t>(out + p0) & 15) == 0)) *reinterpret_cast<uint4*>(out + p0) = make_uint4(0u, 0u, 0u, 0u);
                else for (int q = 0; q < PPT; q++) out[p0 + q] = 0u;
                continue;
            }
        }
#pragma unroll 1
        for (int q = 0; q < PPT; q++) {
            uint32_t c = 0u;
            if (p0 + q < end) {
                scene_begin_pixel(sc, x, y, fp.spp, 0);        // policies with per-pixel state (LbvhBinsScene); nothing for the others
                if constexpr (DBG::enabled) {
                    DBG dbg;
                    c = trace_pixel<SPP1, SPP1>(sc, cam, x, y, fp.w, fp.h, fp.cap, fp.spp, fp.seed, stack, dbg, fp.rcp_w, fp.rcp_h, bits);
                    if (dout->hash) dout->hash[p0 + q] = dbg.hash;
                    if (dout->aov_id) dout->aov_id[p0 + q] = dbg.aov_id;
                    if (dout->aov_t) dout->aov_t[p0 + q] = dbg.aov_t;
                    n_primary += dbg.primary; n_shadow += dbg.n_shadow; n_secondary += dbg.secondary;
                } else {
                    c = trace_pixel<SPP1, SPP1>(sc, cam, x, y, fp.w, fp.h, fp.cap, fp.spp, fp.seed, stack, nodbg, fp.rcp_w, fp.rcp_h, bits);
                }
            }
            if (++x == fp.w) { x = 0; ++y; }
#pragma unroll
            for (int z = 0; z + 1 < PPT; z++) px[z] = px[z + 1];   // shift register: after PPT iterations px[] is in pixel order
            px[PPT - 1] = c;
        }
        if (PPT == 4 && p0 + PPT <= end && ((reinterpret_cast<uintptr_t>(out + p0) & 15) == 0)) {
            *reinterpret_cast<uint4*>(out + p0) = make_uint4(px[0], px[PPT > 1 ? 1 : 0], px[PPT > 2 ? 2 : 0], px[PPT > 3 ? 3 : 0]);   // 128-bit coalesced store
        } else {
            for (int q = 0; q < PPT; q++) if (p0 + q < end) out[p0 + q] = px[q];
        }
    }
    if constexpr (DBG::enabled) {
        if (dout->counters) {
            if (n_primary) atomicAdd(dout->counters + 0, n_primary);
            if (n_shadow) atomicAdd(dout->counters + 1, n_shadow);
            if (n_secondary) atomicAdd(dout->counters + 2, n_secondary);
        }
    }
}

template <int NS, int NL, int NP, bool SPP1>
__global__ void __launch_bounds__(BLOCK, RT_MIN_BLOCKS) k_render_tiny(const __grid_constant__ TinySceneData scd, const __grid_constant__ FrameParams fp) {
    render_loop<PPT_TINY, SPP1>(TinyScene<NS, NL, NP>(scd), fp);
}
// Sparse gather, rank 0's half: the proven-black spans of the tiles that belong to the OTHER ranks are zeroed here, in local HBM,
// so that those ranks need not send zeros over NVLink (35 % of the bench frame). Same span geometry and the same gates as
// render_loop<PPT_TINY, true> — both sides evaluate the identical predicate on identical data, so every span is written by exactly
// one of them. fp is rank 0's own FrameParams (k_begin / tiles_mine = its band; it owns at least as many tiles as any other rank).
__global__ void __launch_bounds__(BLOCK) k_fill_black(const __grid_constant__ FrameParams fp) {
    constexpr int PPT = PPT_TINY, CHUNK = BLOCK * PPT;
    const int npix = fp.w * fp.h;
    const int tile_pix = fp.tile_rows * fp.w;
    const int frame = blockIdx.z;
    const FrameGates& gates = fp.gates[frame];
    uint32_t* out = fp.out + (long long)frame * fp.frame_stride;
    for (int kk = blockIdx.y; kk < fp.tiles_mine; kk += gridDim.y) {
        const int k = fp.k_begin + kk;
        for (int r = 1; r < fp.world; r++) {
            const int tile = k * fp.world + r;
            if (tile >= fp.tiles_total) break;
            const int base = tile * tile_pix;
            int end = base + tile_pix; if (end > npix || end < base) end = npix;
            const int p0 = base + (int)blockIdx.x * CHUNK + (int)threadIdx.x * PPT;
            if (p0 >= end) continue;
            const int y = p0 / fp.w, x = p0 - y * fp.w;
            if (!(x + PPT <= fp.w && p0 + PPT <= end)) continue;
            if (!gate_black_span(gates, x, x + PPT - 1, y)) continue;
            if ((reinterpret_cast<uintptr_t>(out + p0) & 15) == 0) *reinterpret_cast<uint4*>(out + p0) = make_uint4(0u, 0u, 0u, 0u);
            else for (int q = 0; q < PPT; q++) out[p0 + q] = 0u;
        }
    }
}

using TinyKernel = void (*)(const TinySceneData, const FrameParams);
// ---------------------------------------------------------------------------------------------------------------------
// Packed multi-GPU gather (rt_gather.cuh): the render kernel of ranks 1..N-1 and the expand pass of rank 0.
// Requirements checked on the host (gather_applicable): single-sample gated tiny-scene launch, width a multiple of 128 (a warp's
// 128 pixels then never straddle a row, a tile or the end of the frame, so validity and quad membership are warp-uniform).
// ---------------------------------------------------------------------------------------------------------------------
template <int NS, int NL, int NP>
__global__ void __launch_bounds__(BLOCK, RT_MIN_BLOCKS) k_render_tiny_pack(const __grid_constant__ TinySceneData scd, const __grid_constant__ FrameParams fp) {
    constexpr int PPT = PPT_TINY, CHUNK = BLOCK * PPT;
    TinyScene<NS, NL, NP> sc(scd);
    HitRec stack[STACK_RECS];
    NoDbg nodbg;
    const GatherParams& ga = fp.gather;
    GatherCtl* ctl = reinterpret_cast<GatherCtl*>(ga.area);
    const int frame = blockIdx.z, lane = threadIdx.x & 31;
    // write-after-read: rank 0 must have consumed this slot's planes of the previous epoch before anything is stored into them
    if (threadIdx.x == 0 && ga.epoch > 1) {
        volatile unsigned long long* seen = &ga.local->seen_freed[frame];
        if (*seen < ga.epoch - 1) {
            gather_wait_ge(&ctl->freed[frame], ga.epoch - 1, ctl);           // over NVLink; rare: rank 0 trails by less than a frame
            *seen = ga.epoch - 1;
        }
    }
    __syncthreads();
    unsigned char* plane_c = ga.area + ga.off_c + (unsigned long long)frame * ga.stride_c;
    unsigned char* plane_g = ga.area + ga.off_g + (unsigned long long)frame * ga.stride_g;
    unsigned char* plane_f = ga.area + ga.off_f + (unsigned long long)frame * ga.stride_f;
    const int npix = fp.w * fp.h;
    const int tile_pix = fp.tile_rows * fp.w;
    const CamRec& cam = fp.cam_inline[frame];
    const FrameGates& gates = fp.gates[frame];
    const unsigned qmask = 0xFu << (lane & 28);
    const int q = lane >> 2, ql = lane & 3;
    __shared__ unsigned int flag_bits[2][4];                       // 2-D blocks: the flag bytes of the item's four rows (double-buffered)
    int flag_par = 0;
    if (threadIdx.x < 8) flag_bits[threadIdx.x >> 2][threadIdx.x & 3] = 0u;
    __syncthreads();
    // gridDim.x CTAs per frame, each striding over the (tile, chunk) items of this rank: a few hundred fat CTAs instead of one per
    // chunk, because every CTA ends with a system-scope fence that waits for its NVLink stores to be acknowledged (measured at N = 2:
    // 16 200 fences per frame cost 0.4 ms per step, profiles/r02/)
    const int n_items = fp.tiles_mine * fp.chunks_per_tile;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int kk = item / fp.chunks_per_tile, chunk = item - kk * fp.chunks_per_tile;
        int kr = kk + fp.tile_rot; if (kr >= fp.tiles_mine) kr -= fp.tiles_mine;
        const int tile = tile_of(fp, fp.k_begin + kr);
        const int base = tile * tile_pix;
        int end = base + tile_pix; if (end > npix || end < base) end = npix;
        int p0, x, y;
        bool black = false;
        const int wp = (int)threadIdx.x >> 5;
        int cb = 0, y_first = 0;
        if (fp.tile2d) {
            // 2-D pixel blocks as in render_loop: a warp is 32 x 4 pixels (a quad still 16 consecutive pixels of one row); here the
            // CTA's four warps lie side by side — 128 x 4 pixels — so that the CTA owns whole flag bytes (one per row and 128-pixel
            // group, assembled from the warps' quad bits in shared memory below). Rows below the frame (last tile) count as black:
            // nothing traced or stored. No warp leaves the item early: the item loop and its barriers are CTA-uniform.
            black = !pack2d_span(fp.w, fp.h, tile, fp.tile_rows, chunk, (int)threadIdx.x, &x, &y, &cb, &y_first);     // rt_tiles.cuh; w % 128 == 0
            p0 = y * fp.w + x;
        } else {
            p0 = base + chunk * CHUNK + (int)threadIdx.x * PPT;
            if (p0 >= end) continue;                               // warp-uniform (see above)
            y = p0 / fp.w; x = p0 - y * fp.w;
        }
        uint32_t bits = 0u;
        if (!black) bits = gate_bits_span(gates, x, x + PPT - 1, y, sc.n_lights(), &black);
        uint32_t px[PPT] = {0u, 0u, 0u, 0u};
        if (!black) {
#pragma unroll 1
            for (int i = 0; i < PPT; i++) {
                const uint32_t c = trace_pixel<true, true>(sc, cam, x + i, y, fp.w, fp.h, fp.cap, 1, fp.seed, stack, nodbg, fp.rcp_w, fp.rcp_h, bits);
#pragma unroll
                for (int z = 0; z + 1 < PPT; z++) px[z] = px[z + 1];
                px[PPT - 1] = c;
            }
        }
        __syncwarp();
        const uint32_t qblack = gather_quad_all(black);
        const bool grey4 = gather_is_grey(px[0]) && gather_is_grey(px[1]) && gather_is_grey(px[2]) && gather_is_grey(px[3]);
        const uint32_t qgrey = ga.grey ? gather_quad_all(grey4) : 0u;
        if (!((qblack >> q) & 1u)) {                               // quad-uniform from here on
            if ((qgrey >> q) & 1u) {
                *reinterpret_cast<uint32_t*>(plane_g + p0) = (px[0] & 0xFFu) | ((px[1] & 0xFFu) << 8) | ((px[2] & 0xFFu) << 16) | ((px[3] & 0xFFu) << 24);
            } else {
                // 4 lanes x 12 bytes -> 3 lanes x 16 bytes: quad word m = 3 * lane + slot goes to output lane m / 4, word m % 4
                uint32_t wd[3];
                gather_pack_rgb(px, wd);
                const int l0 = lane & 28;
                const uint32_t a = ql == 0 ? wd[0] : (ql == 1 ? wd[1] : wd[2]);
                const uint32_t b = ql == 0 ? wd[1] : (ql == 1 ? wd[2] : wd[0]);
                const uint32_t c = ql == 0 ? wd[2] : (ql == 2 ? wd[0] : wd[1]);
                const uint32_t d = ql == 1 ? wd[0] : (ql == 2 ? wd[1] : wd[2]);
                const uint32_t o0 = __shfl_sync(qmask, a, l0 + (ql < 3 ? ql : 0));
                const uint32_t o1 = __shfl_sync(qmask, b, l0 + (ql == 0 ? 0 : (ql == 1 ? 1 : 3)));
                const uint32_t o2 = __shfl_sync(qmask, c, l0 + (ql == 0 ? 0 : (ql == 1 ? 2 : 3)));
                const uint32_t o3 = __shfl_sync(qmask, d, l0 + (ql < 3 ? ql + 1 : 3));
                if (ql < 3)
                    *reinterpret_cast<uint4*>(plane_c + 3ull * (unsigned long long)(p0 - ql * PPT) + 16u * (unsigned)ql) = make_uint4(o0, o1, o2, o3);
            }
        }
        if (ga.grey) {
            if (fp.tile2d) {
                // quad 2 r + h of the warp = row r, half h of the warp's 32 pixels -> bits 2 wp + h of the flag byte of (row r, group cb)
                unsigned int* fl = flag_bits[flag_par];
                if (lane == 0)
                    for (int r = 0; r < 4; r++) atomicOr(&fl[r], ((qgrey >> (2 * r)) & 3u) << (2 * wp));
                __syncthreads();
                if (threadIdx.x < 4) {
                    const int yr = y_first + (int)threadIdx.x;
                    if (yr < fp.h) plane_f[((long long)yr * fp.w >> 7) + cb] = (unsigned char)fl[threadIdx.x];
                    fl[threadIdx.x] = 0u;                          // this buffer is used again two items on, behind the next item's barrier
                }
                flag_par ^= 1;
            } else if (lane == 0 && qblack != 0xFFu) {
                plane_f[p0 >> 7] = (unsigned char)qgrey;
            }
        }
    }
    // this rank's planes of the slot have landed once its LAST CTA is through
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(&ga.local->cta_count[frame], 1u);
        if (prev == gridDim.x - 1u) {
            ga.local->cta_count[frame] = 0u;
            __threadfence_system();
            gather_st_release(&ctl->done[fp.rank][frame], ga.epoch);
        }
    }
}
TinyKernel tiny_kernel_pack(int ns, int nl, int np) {
    return (ns == 3 && nl == 2 && np == 1) ? k_render_tiny_pack<3, 2, 1> : k_render_tiny_pack<-1, -1, -1>;
}

// Rank 0: expands the other ranks' tiles from the planes into the framebuffer (local HBM), zero-fills their proven-black quads,
// and hands the slot back. One CTA per (chunk, tile, frame) over ALL tiles; CTAs of rank 0's own tiles have nothing to do.
__global__ void __launch_bounds__(BLOCK) k_gather_expand(const __grid_constant__ FrameParams fp) {
    constexpr int PPT = PPT_TINY, CHUNK = BLOCK * PPT;
    const GatherParams& ga = fp.gather;
    GatherCtl* ctl = reinterpret_cast<GatherCtl*>(ga.area);
    const int frame = blockIdx.z, lane = threadIdx.x & 31, q = lane >> 2;
    const unsigned char* plane_c = ga.area + ga.off_c + (unsigned long long)frame * ga.stride_c;
    const unsigned char* plane_g = ga.area + ga.off_g + (unsigned long long)frame * ga.stride_g;
    const unsigned char* plane_f = ga.area + ga.off_f + (unsigned long long)frame * ga.stride_f;
    const int npix = fp.w * fp.h;
    const int tile_pix = fp.tile_rows * fp.w;
    const FrameGates& gates = fp.gates[frame];
    uint32_t* out = fp.out + (long long)frame * fp.frame_stride;
    // gridDim.x CTAs per frame striding over the (tile, chunk) items of the frame; a rank's `done` flag is awaited once per CTA
    unsigned ready = 1u;                                                     // bit r: rank r's planes of this slot have landed (CTA-uniform)
    const int n_items = fp.tiles_total * fp.chunks_per_tile;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {         // CTA-uniform
        const int t = item / fp.chunks_per_tile, chunk = item - t * fp.chunks_per_tile;
        const int owner = fp.part_owner[t % fp.part_period];
        if (owner == 0) continue;
        if (!((ready >> owner) & 1u)) {
            if (threadIdx.x == 0) gather_wait_ge(&ctl->done[owner][frame], ga.epoch, ctl);
            __syncthreads();
            ready |= 1u << owner;
        }
        const int base = t * tile_pix;
        int end = base + tile_pix; if (end > npix || end < base) end = npix;
        const int p0 = base + chunk * CHUNK + (int)threadIdx.x * PPT;
        if (p0 >= end) continue;                                             // warp-uniform; the barrier above was CTA-wide
        // row of the span without an integer division: p0 < 2^24 is exact in fp32, 1/w is correctly rounded -> off by one at most
        int y = (int)((float)p0 * fp.rcp_w);
        if (y * fp.w > p0) y--; else if ((y + 1) * fp.w <= p0) y++;
        const int x = p0 - y * fp.w;
        const bool black = gate_black_span(gates, x, x + PPT - 1, y);
        const uint32_t qblack = gather_quad_all(black);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (qblack != 0xFFu) {                                               // warp-uniform
            // The planes were written by another GPU: read them from L2 (the coherence point), never through L1 — a line of the F
            // plane can hold flags of two tiles, i.e. of two ranks that finish at different times. The flag byte and the G-plane
            // word are fetched together (one round trip for grey quads); only coloured quads pay a second one for the C plane.
            const uint32_t qgrey = ga.grey ? (uint32_t)__ldcg(plane_f + (p0 >> 7)) : 0u;
            const uint32_t gword = ga.grey ? __ldcg(reinterpret_cast<const uint32_t*>(plane_g + p0)) : 0u;
            if (!((qblack >> q) & 1u)) {
                if ((qgrey >> q) & 1u) {
                    v = gather_unpack_grey(gword);
                } else {
                    const uint32_t* c = reinterpret_cast<const uint32_t*>(plane_c + 3ull * (unsigned long long)p0);
                    v = gather_unpack_rgb(__ldcg(c), __ldcg(c + 1), __ldcg(c + 2));
                }
            }
        }
        *reinterpret_cast<uint4*>(out + p0) = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(&ctl->expand_count[frame], 1u);
        if (prev == gridDim.x - 1u) {
            ctl->expand_count[frame] = 0u;
            __threadfence_system();
            gather_st_release(&ctl->freed[frame], ga.epoch);                 // the other ranks poll this over NVLink
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Opt-in variant (rt_set_option(RT_OPT_COMPACTION, 1)): warp-ballot ray compaction between bounces.
// Pass 1 traces every pixel of the chunk up to its second hit. A chain that goes on (a mirror seen in a mirror) is not followed by
// its lane: the lanes that want to continue are counted with __ballot_sync / __popc, one lane per warp reserves that many slots
// of a per-CTA shared-memory queue, and each parks its ray state there (ray, bounce count, the two hit records needed for the
// exact back-to-front colour sum). After the barrier pass 2 hands the parked rays to the first threads of the CTA, so deep chains
// run in fully populated warps instead of one or two lanes of many. Both passes go through ONE call site of trace_chain (a
// second inlined copy would double the kernel and push it out of the instruction cache). Pixels are identical to the default
// kernel's (tests/test_gpu_parity.py::test_compaction_variant_is_identical); measured speed: profiles/r01/tuning.md.
// spp == 1 only (a parked sample could not be averaged in order); other launches take the default kernel.
// ---------------------------------------------------------------------------------------------------------------------
struct ParkedRay { int p; int bounce; f3 o; f3 dir; HitRec rec[2]; };     // 96 B
constexpr int PARK_CAP = 128;          // per CTA; a full queue simply makes further lanes follow their chain inline
constexpr int DEFER_LEVEL = 2;

template <int NS, int NL, int NP>
__global__ void __launch_bounds__(BLOCK, RT_MIN_BLOCKS) k_render_tiny_compact(const __grid_constant__ TinySceneData scd,
                                                                              const __grid_constant__ FrameParams fp) {
    constexpr int PPT = PPT_TINY, CHUNK = BLOCK * PPT;
    __shared__ ParkedRay queue[PARK_CAP];
    __shared__ int qcount;
    TinyScene<NS, NL, NP> sc(scd);
    HitRec stack[STACK_RECS];
    NoDbg dbg;
    if (threadIdx.x == 0) qcount = 0;
    __syncthreads();
    const int npix = fp.w * fp.h;
    const int tile_pix = fp.tile_rows * fp.w;
    const int frame = blockIdx.z;
    const int lane = threadIdx.x & 31;
    const CamRec& cam = fp.cam_inline[frame];
    uint32_t* out = fp.out + (long long)frame * fp.frame_stride;
    const float fw = (float)fp.w, fh = (float)fp.h;
    for (int kk = blockIdx.y; kk < fp.tiles_mine; kk += gridDim.y) {       // uniform per CTA
        const int k = fp.k_begin + kk;
        const int tile = tile_of(fp, k);
        const int base = tile * tile_pix;
        int end = base + tile_pix; if (end > npix || end < base) end = npix;
        const int p0 = base + (int)blockIdx.x * CHUNK + (int)threadIdx.x * PPT;
        uint32_t px[PPT];
        int y = p0 / fp.w, x = p0 - y * fp.w;
#pragma unroll 1
        for (int it = 0; it <= PPT; it++) {                                // it < PPT: my pixels; it == PPT: parked rays
            if (it == PPT) __syncthreads();
            bool active; int p, bounce, top, defer_at; f3 o, dir;
            if (it < PPT) {
                p = p0 + it; active = p < end; bounce = 0; top = 0; defer_at = DEFER_LEVEL;
                if (active) primary_ray<true>(cam, (float)x, (float)y, fw, fh, fp.rcp_w, fp.rcp_h, &o, &dir);
                if (++x == fp.w) { x = 0; ++y; }
            } else {
                const int n = qcount < PARK_CAP ? qcount : PARK_CAP;
                active = (int)threadIdx.x < n; defer_at = -1; top = DEFER_LEVEL; p = 0; bounce = 0;
                if (active) {
                    const ParkedRay& e = queue[threadIdx.x];
                    p = e.p; bounce = e.bounce; o = e.o; dir = e.dir; stack[0] = e.rec[0]; stack[1] = e.rec[1];
                }
            }
            f3 C = mk3(0, 0, 0);
            bool done = !active, parked = false;
#pragma unroll 1
            for (int attempt = 0; attempt < 2; attempt++) {
                if (active && !done) done = trace_chain(sc, fp.cap, o, dir, bounce, top, stack, defer_at, &C, dbg);
                if (attempt == 0 && it < PPT) {                            // every lane of the warp is here: convergent ballot
                    const bool want = active && !done;
                    const unsigned m = __ballot_sync(0xffffffffu, want);
                    if (m) {
                        const int leader = __ffs((int)m) - 1;
                        int slot0 = 0;
                        if (lane == leader) slot0 = atomicAdd(&qcount, __popc(m));
                        slot0 = __shfl_sync(0xffffffffu, slot0, leader);
                        const int slot = slot0 + __popc(m & ((1u << lane) - 1u));
                        if (want && slot < PARK_CAP) {
                            ParkedRay& e = queue[slot];
                            e.p = p; e.bounce = bounce; e.o = o; e.dir = dir; e.rec[0] = stack[0]; e.rec[1] = stack[1];
                            parked = true;
                        }
                    }
                }
                if (done || parked || !active) break;
                defer_at = -1;                                             // queue full: follow the chain inline
            }
            if (it < PPT) {
                const uint32_t c = (active && done) ? pack_color(C) : 0u;  // parked pixels are written by pass 2
#pragma unroll
                for (int z = 0; z + 1 < PPT; z++) px[z] = px[z + 1];
                px[PPT - 1] = c;
                if (it == PPT - 1 && p0 < end) {
                    if (p0 + PPT <= end && ((reinterpret_cast<uintptr_t>(out + p0) & 15) == 0)) {
                        *reinterpret_cast<uint4*>(out + p0) = make_uint4(px[0], px[PPT > 1 ? 1 : 0], px[PPT > 2 ? 2 : 0], px[PPT > 3 ? 3 : 0]);
                    } else {
                        for (int q = 0; q < PPT; q++) if (p0 + q < end) out[p0 + q] = px[q];
                    }
                }
            } else if (active) {
                out[p] = pack_color(C);                                    // after the barrier: overwrites the placeholder
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) qcount = 0;
        __syncthreads();
    }
}

// Exact-count instantiations (sphere and light loops unrolled, records addressed statically) for scenes of the
// reference's size; everything else up to the TINY_MAX_* limits takes the run-time-count instantiation.
constexpr int EXACT_MAX = 4;
template <int NS, int NL> struct TinyTable {
    static TinyKernel get(int ns, int nl) {
        if (ns == NS && nl == NL) return k_render_tiny<NS, NL, 1, true>;
        if constexpr (NL < EXACT_MAX) return TinyTable<NS, NL + 1>::get(ns, nl);
        else if constexpr (NS < EXACT_MAX) return TinyTable<NS + 1, 0>::get(ns, nl);
        else return k_render_tiny<-1, -1, -1, true>;
    }
};
// Exact kernels exist for 0..4 spheres x 0..4 lights x exactly 1 plane x one sample per pixel (the reference is 3 x 2 x 1 x 1).
// Supersampled frames (extension) take the reference scene's own multi-sample instantiation or the run-time-count kernel.
// Frames with a side above RT_FASTDIV_MAX (outside the exhaustively verified range of rt_div_rcp) also take the multi-sample kernels,
// which divide with the IEEE sequence.
TinyKernel tiny_kernel(int ns, int nl, int np, int spp, bool fastdiv_ok) {
    if (spp == 1 && fastdiv_ok) return np == 1 ? TinyTable<0, 0>::get(ns, nl) : k_render_tiny<-1, -1, -1, true>;
    return (ns == 3 && nl == 2 && np == 1) ? k_render_tiny<3, 2, 1, false> : k_render_tiny<-1, -1, -1, false>;
}
// compacting variant: exact instantiation for the reference scene's shape, run-time counts otherwise
TinyKernel tiny_kernel_compact(int ns, int nl, int np) {
    return (ns == 3 && nl == 2 && np == 1) ? k_render_tiny_compact<3, 2, 1> : k_render_tiny_compact<-1, -1, -1>;
}
__global__ void __launch_bounds__(BLOCK, RT_MIN_BLOCKS) k_render_global(const __grid_constant__ GlobalSceneData scd, const __grid_constant__ FrameParams fp) {
    render_loop<PPT_HEAVY>(GlobalScene(scd), fp);
}
// Brute force with the sphere geometry staged in shared memory (dynamic: 16 B per sphere).
__device__ __forceinline__ const f4* stage_spheres(const GlobalSceneData& scd) {
    extern __shared__ float4 sm_raw[];
    f4* sm = reinterpret_cast<f4*>(sm_raw);
    for (int i = threadIdx.x; i < scd.ns; i += blockDim.x) sm[i] = load_f4(scd.sgeom + i);
    __syncthreads();
    return sm;
}
__global__ void __launch_bounds__(BLOCK, RT_MIN_BLOCKS) k_render_staged(const __grid_constant__ GlobalSceneData scd, const __grid_constant__ FrameParams fp) {
    render_loop<PPT_HEAVY>(StagedScene(scd, stage_spheres(scd)), fp);
}
struct LbvhSceneData { GlobalSceneData g; BvhView bv; ShadowGridsView sg; };
#ifndef RT_LBVH_MIN_BLOCKS
#define RT_LBVH_MIN_BLOCKS 10    // 48 registers: the traversal hides its node fetches with resident warps — measured on B200 (profiles/r02/tuning.md):
                                 // 8 CTAs / SM 1.399 ms, 9 1.332, 10 1.292, 11 1.316, 12 1.303 on configs[3]; fewer than 8 is slower still
#endif
__global__ void __launch_bounds__(BLOCK, RT_LBVH_MIN_BLOCKS) k_render_lbvh(const __grid_constant__ LbvhSceneData scd, const __grid_constant__ FrameParams fp) {
    render_loop<PPT_HEAVY>(LbvhScene(scd.g, scd.bv, scd.sg), fp);
}
// The same with the frame's primary bins (rt_primary_bins.cuh; RT_OPT_PRIMARY_BINS): primary rays fold over the sphere list of their
// 8 x 8-pixel tile — one list per warp of the 2-D pixel blocks — instead of walking the tree.
struct LbvhBinsSceneData { LbvhSceneData l; PrimaryBinsView pb; };
__global__ void __launch_bounds__(BLOCK, RT_LBVH_MIN_BLOCKS) k_render_lbvh_bins(const __grid_constant__ LbvhBinsSceneData scd, const __grid_constant__ FrameParams fp) {
    render_loop<PPT_HEAVY>(LbvhBinsScene(scd.l.g, scd.l.bv, scd.l.sg, scd.pb), fp);
}

// Instrumented kernel: one thread per pixel, writes hash / AOVs / counters.
template <class SC>
__device__ __forceinline__ void debug_loop(const SC& sc, const FrameParams& fp, const DebugOut& dout) {
    HitRec stack[STACK_RECS];
    const int npix = fp.w * fp.h;
    unsigned long long cnt[N_DEBUG_COUNTERS] = {0};
    for (long long pl = (long long)blockIdx.x * blockDim.x + threadIdx.x; pl < npix; pl += (long long)gridDim.x * blockDim.x) {
        const int p = (int)pl;
        FullDbg dbg;
        const int y = p / fp.w, x = p - y * fp.w;
        scene_begin_pixel(sc, x, y, fp.spp, 0);
        uint32_t c = trace_pixel(sc, fp.cam_inline[0], x, y, fp.w, fp.h, fp.cap, fp.spp, fp.seed, stack, dbg);
        fp.out[p] = c;
        if (dout.hash) dout.hash[p] = dbg.hash;
        if (dout.aov_id) dout.aov_id[p] = dbg.aov_id;
        if (dout.aov_t) dout.aov_t[p] = dbg.aov_t;
        cnt[0] += dbg.primary; cnt[1] += dbg.n_shadow; cnt[2] += dbg.secondary; cnt[3] += dbg.sphere_tests;
        cnt[4] += dbg.sphere_disc_pos; cnt[5] += dbg.plane_tests; cnt[6] += dbg.shade_diffuse; cnt[7] += dbg.shade_specular;
        cnt[8] += dbg.shade_mirror; cnt[9] += dbg.shaded_hits;
        cnt[10] += dbg.node_visits[0]; cnt[11] += dbg.node_visits[1]; cnt[12] += dbg.node_visits[2]; cnt[13] += dbg.fallbacks;
    }
    if (dout.counters)
        for (int i = 0; i < N_DEBUG_COUNTERS; i++) if (cnt[i]) atomicAdd(dout.counters + i, cnt[i]);
}
__global__ void __launch_bounds__(BLOCK) k_debug_tiny(const __grid_constant__ TinySceneData scd, const __grid_constant__ FrameParams fp, DebugOut dout) {
    debug_loop(TinyScene<-1, -1, -1>(scd), fp, dout);
}
// The shipped tiny-scene kernel with the events-only debug policy: same template arguments, same gates, same grid as k_render_tiny.
template <int NS, int NL, int NP>
__global__ void __launch_bounds__(BLOCK) k_debug_tiny_prod(const __grid_constant__ TinySceneData scd, const __grid_constant__ FrameParams fp, DebugOut dout) {
    render_loop<PPT_TINY, true, ProdDbg>(TinyScene<NS, NL, NP>(scd), fp, &dout);
}
using TinyDebugKernel = void (*)(const TinySceneData, const FrameParams, DebugOut);
// exact-count instantiations for a few scene shapes incl. the reference's (3 spheres, 2 lights, 1 plane); every other shape takes the
// run-time-count instantiation (which is also what production uses for scenes outside its exact table)
TinyDebugKernel tiny_debug_prod_kernel(int ns, int nl, int np) {
    if (np == 1) {
        if (ns == 3 && nl == 2) return k_debug_tiny_prod<3, 2, 1>;
        if (ns == 0 && nl == 0) return k_debug_tiny_prod<0, 0, 1>;
        if (ns == 1 && nl == 1) return k_debug_tiny_prod<1, 1, 1>;
        if (ns == 2 && nl == 2) return k_debug_tiny_prod<2, 2, 1>;
        if (ns == 4 && nl == 4) return k_debug_tiny_prod<4, 4, 1>;
    }
    return k_debug_tiny_prod<-1, -1, -1>;
}
__global__ void __launch_bounds__(BLOCK) k_debug_global(const __grid_constant__ GlobalSceneData scd, const __grid_constant__ FrameParams fp, DebugOut dout) {
    debug_loop(GlobalScene(scd), fp, dout);
}
__global__ void __launch_bounds__(BLOCK) k_debug_staged(const __grid_constant__ GlobalSceneData scd, const __grid_constant__ FrameParams fp, DebugOut dout) {
    debug_loop(StagedScene(scd, stage_spheres(scd)), fp, dout);
}
__global__ void __launch_bounds__(BLOCK) k_debug_lbvh(const __grid_constant__ LbvhSceneData scd, const __grid_constant__ FrameParams fp, DebugOut dout) {
    debug_loop(LbvhScene(scd.g, scd.bv, scd.sg), fp, dout);
}
__global__ void __launch_bounds__(BLOCK) k_debug_lbvh_bins(const __grid_constant__ LbvhBinsSceneData scd, const __grid_constant__ FrameParams fp, DebugOut dout) {
    debug_loop(LbvhBinsScene(scd.l.g, scd.l.bv, scd.l.sg, scd.pb), fp, dout);
}

// Sphere-query kernels (LBVH == brute equality harness, rt_query_spheres).
template <class SC>
__device__ __forceinline__ void query_loop(const SC& sc, const float* rays6, int n, int kind, int32_t* out_id, float* out_t) {
    NoDbg dbg;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        f3 o = mk3(rays6[6 * r], rays6[6 * r + 1], rays6[6 * r + 2]);
        f3 d = mk3(rays6[6 * r + 3], rays6[6 * r + 4], rays6[6 * r + 5]);
        float a = dot3(d, d), a2 = 2 * a, a4 = 4 * a;
        int sel = -1; float t = 0.0f;
        if (kind == 0) { sc.nearest(o, d, a2, a4, 0.0f, &sel, &t, dbg); if (sel < 0) t = 0.0f; }
        else if (kind == 1) { sc.nearest(o, d, a2, a4, 0.01f, &sel, &t, dbg); if (sel < 0) t = 0.0f; }
        else { sel = sc.shadow_any(-1, o, d, a2, a4, dbg) ? 1 : 0; t = 0.0f; }
        out_id[r] = sel; out_t[r] = t;
    }
}
__global__ void __launch_bounds__(BLOCK) k_query_brute(const __grid_constant__ GlobalSceneData scd, const float* rays6, int n, int kind,
                                                        int32_t* out_id, float* out_t) {
    query_loop(GlobalScene(scd), rays6, n, kind, out_id, out_t);
}
__global__ void __launch_bounds__(BLOCK) k_query_lbvh(const __grid_constant__ LbvhSceneData scd, const float* rays6, int n, int kind,
                                                       int32_t* out_id, float* out_t) {
    query_loop(LbvhScene(scd.g, scd.bv, scd.sg), rays6, n, kind, out_id, out_t);
}

// Self-test kernels (rt_selftest): the hand-scheduled fp32 sequences of rt_math.cuh against the compiler's IEEE code, exhaustively.
__global__ void __launch_bounds__(256) k_selftest_inv_len(uint32_t first_bits, uint32_t count, int variant, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const float s = __uint_as_float(first_bits + i);
        float got;
        if (variant == 0) got = rt_inv_len(s);
        else {                                                 // experiment: reuse the rsqrt estimate as the reciprocal seed
            float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
            const float g = __fmul_rn(s, y), h = __fmul_rn(y, 0.5f);
            const float len = __fmaf_rn(__fmaf_rn(-g, g, s), h, g);
            got = __fmaf_rn(y, __fmaf_rn(-len, y, 1.0f), y);
        }
        if (__float_as_uint(got) != __float_as_uint(rt_inv_len_ieee(s))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}
// The packed pair version (rt_inv_len2): every float of the range in half .x, and — through a bijection of the range — in half .y.
__global__ void __launch_bounds__(256) k_selftest_inv_len_pair(uint32_t first_bits, uint32_t count, unsigned long long* mismatches) {
#if defined(RT_HAVE_F32X2)
    unsigned long long bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(((unsigned long long)i * 2654435761ull) % count);       // 2654435761 is prime: a permutation
        const float s0 = __uint_as_float(first_bits + i), s1 = __uint_as_float(first_bits + j);
        const float2 got = rt_inv_len2(make_float2(s0, s1));
        if (__float_as_uint(got.x) != __float_as_uint(rt_inv_len_ieee(s0))) bad++;
        if (__float_as_uint(got.y) != __float_as_uint(rt_inv_len_ieee(s1))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
#endif
}
__global__ void __launch_bounds__(256) k_selftest_pixel_div(int max_side, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (int w = 1 + blockIdx.x; w <= max_side; w += gridDim.x) {
        const float fw = (float)w, rw = 1.0f / fw;
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            const float fx = (float)x;
            if (__float_as_uint(rt_div_rcp(fx, fw, rw)) != __float_as_uint(fx / fw)) bad++;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// L2 read-bandwidth probe (rt_measure_l2_read): every CTA streams the whole working set (sized to stay L2-resident) with 128-bit
// L1-bypassing loads, starting at a different offset so that the CTAs do not move in lock step; the xor keeps the loads alive.
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n_vec, int passes, unsigned int* sink) {
    unsigned int acc = 0;
    const size_t start = ((size_t)blockIdx.x * 7919u * 256u) % n_vec;
    for (int p = 0; p < passes; p++) {
        size_t i = start + threadIdx.x;
#pragma unroll 4
        for (size_t k = 0; k < n_vec / 256; k++) {
            if (i >= n_vec) i -= n_vec;
            const uint4 v = __ldcg(buf + i);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
            i += 256;
        }
    }
    if (acc == 0x12345u) *sink = acc;          // never true for the zero-filled buffer; defeats dead-code elimination
}

// Ray-log kernel (rt_ray_log): one thread per listed pixel, `slots` records reserved per pixel, count[i] = records produced.
static_assert(sizeof(RayRec) == sizeof(rt_ray_record), "RayRec must mirror rt_ray_record");
__global__ void __launch_bounds__(BLOCK) k_ray_log(const __grid_constant__ GlobalSceneData scd, const __grid_constant__ FrameParams fp,
                                                    const uint32_t* pixels, int n_pixels, int slots, RayRec* recs, uint32_t* count) {
    HitRec stack[STACK_RECS];
    GlobalScene sc(scd);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        const uint32_t p = pixels[i];
        LogDbg dbg; dbg.out = recs + (size_t)i * slots; dbg.cap = (uint32_t)slots; dbg.pixel = p;
        const int y = (int)(p / (uint32_t)fp.w), x = (int)(p - (uint32_t)y * (uint32_t)fp.w);
        trace_pixel<true>(sc, fp.cam_inline[0], x, y, fp.w, fp.h, fp.cap, 1, 0u, stack, dbg);
        count[i] = dbg.n;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
struct DeviceState {
    int dev = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // scene
    f4* sgeom = nullptr; MatRec* smat = nullptr; PlaneRec* planes = nullptr; LightRec* lights = nullptr;
    LbvhDevice bvh;
    ShadowGridsDevice sg;                                         // per-light shadow bins (rt_shadow_grid.cuh), built on this device
    float bvh_cam[3] = {0, 0, 0}; bool bvh_cam_valid = false;    // camera the nodes_cam copy is currently inflated for ...
    cudaStream_t bvh_cam_stream = nullptr;                        // ... by a refit issued on this stream
    // nodes_cam / the refit arrival counters are ONE buffer per device: a refit issued on another stream than the last LBVH launch
    // first waits for that launch (rt_render_device on several user streams, rt_update_spheres while user-stream frames are in flight)
    cudaEvent_t bvh_done = nullptr; cudaStream_t bvh_last_stream = nullptr; bool bvh_done_valid = false;
    // per-frame primary bins (rt_primary_bins.cuh): one set of buffers per device, ordered like nodes_cam
    PrimaryBinsDevice pb; CamRec pb_cam; int pb_w = 0, pb_h = 0; bool pb_valid = false, pb_usable = false; cudaStream_t pb_stream = nullptr;
    // framebuffer ring (device 0 of the context only, unless partitioned multi-process)
    uint32_t* fb = nullptr; size_t fb_pixels = 0;
    // band pipelining (render_frames): copy stream on device 0, per-segment "band rendered" events on every device
    cudaStream_t copy_stream = nullptr;
    // sparse gather, rank 0: k_fill_black runs on its own stream beside the render kernel (it only touches the OTHER ranks' tiles)
    cudaStream_t fill_stream = nullptr; cudaEvent_t fill_go = nullptr, fill_done = nullptr;
    cudaEvent_t evc0 = nullptr, evc1 = nullptr;
    std::vector<cudaEvent_t> band_events;
};

enum ScenePath { PATH_TINY = 0, PATH_STAGED = 1, PATH_GLOBAL = 2, PATH_LBVH = 3 };
constexpr int STAGED_MAX_SPHERES = 12288;      // 192 KB of dynamic shared memory
constexpr int AUTO_LBVH_MIN_SPHERES = 48;      // below this the brute-force loop is cheaper than a traversal

}  // namespace

// One enqueue worker per extra device: a multi-device frame needs ~5 runtime calls per device (refit, launch, events, copies);
// issued from one thread they skew the last device's start by ~20 us per device. The workers issue them concurrently.
struct DeviceWorkers {
    struct W { std::thread th; std::mutex m; std::condition_variable cv; std::function<int()> job; bool has_job = false, quit = false; int rc = 0; };
    std::vector<W*> ws;
    void start(int n) {
        for (int i = 0; i < n; i++) {
            W* w = new W();
            w->th = std::thread([w]() {
                std::unique_lock<std::mutex> lk(w->m);
                for (;;) {
                    w->cv.wait(lk, [w] { return w->has_job || w->quit; });
                    if (w->quit) return;
                    std::function<int()> j = std::move(w->job);
                    lk.unlock();
                    int rc = j();
                    lk.lock();
                    w->rc = rc; w->has_job = false;
                    w->cv.notify_all();
                }
            });
            ws.push_back(w);
        }
    }
    void post(int i, std::function<int()> j) { W* w = ws[(size_t)i]; { std::lock_guard<std::mutex> lk(w->m); w->job = std::move(j); w->has_job = true; } w->cv.notify_all(); }
    int wait(int i) { W* w = ws[(size_t)i]; std::unique_lock<std::mutex> lk(w->m); w->cv.wait(lk, [w] { return !w->has_job; }); return w->rc; }
    void stop() {
        for (W* w : ws) { { std::lock_guard<std::mutex> lk(w->m); w->quit = true; } w->cv.notify_all(); w->th.join(); delete w; }
        ws.clear();
    }
};

// Host-side zero fill of the parts of `Surface.pixels` that are not copied (sparse D2H, render_frames): a few library threads
// zero the segments while the copies of the rest are in flight. Threads are started on first use (RTB200_FILL_THREADS, default 6,
// at most hardware_concurrency - 1; 0 = the calling thread does it all in wait()).
struct FillPool {
    struct Seg { char* p; size_t bytes; };
    std::vector<std::thread> th;
    std::mutex m; std::condition_variable cv, cv_done;
    std::vector<Seg> segs; std::atomic<size_t> next{0};
    uint64_t gen = 0; int active = 0; bool quit = false, started = false;
    // Non-temporal zero fill: the lines are not read for ownership and do not evict the host's caches (the frame is consumed by
    // the display path, not by this thread). SSE2 is x86-64 baseline.
    static void zero_nt(char* p, size_t bytes) {
#if defined(__SSE2__)
        const size_t head = (size_t)(-(intptr_t)p & 15);
        if (head >= bytes) { memset(p, 0, bytes); return; }
        memset(p, 0, head); p += head; bytes -= head;
        const __m128i z = _mm_setzero_si128();
        size_t n = bytes / 64;
        for (size_t i = 0; i < n; i++, p += 64) {
            _mm_stream_si128((__m128i*)p, z); _mm_stream_si128((__m128i*)(p + 16), z);
            _mm_stream_si128((__m128i*)(p + 32), z); _mm_stream_si128((__m128i*)(p + 48), z);
        }
        _mm_sfence();
        memset(p, 0, bytes & 63);
#else
        memset(p, 0, bytes);
#endif
    }
    void drain() { for (;;) { const size_t i = next.fetch_add(1); if (i >= segs.size()) break; zero_nt(segs[i].p, segs[i].bytes); } }
    void start() {
        started = true;
        int n = 6;
        if (const char* e = getenv("RTB200_FILL_THREADS")) n = atoi(e);
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && n > hw - 1) n = hw - 1;
        if (n < 0) n = 0;
        for (int i = 0; i < n; i++)
            th.emplace_back([this]() {
                uint64_t seen = 0;
                std::unique_lock<std::mutex> lk(m);
                for (;;) {
                    cv.wait(lk, [&] { return quit || gen != seen; });
                    if (quit) return;
                    seen = gen;
                    lk.unlock();
                    drain();
                    lk.lock();
                    if (--active == 0) cv_done.notify_all();
                }
            });
    }
    void run(std::vector<Seg>&& v) {             // asynchronous; pair with wait()
        if (!started) start();
        std::lock_guard<std::mutex> lk(m);
        segs = std::move(v); next = 0; active = (int)th.size(); gen++;
        cv.notify_all();
    }
    void wait() {                                // the caller helps, then waits for the workers' last segments
        drain();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return active == 0; });
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
};

struct rt_context {
    std::vector<DeviceState> devs;
    std::string err;
    std::mutex err_mu;              // fail() may be called from the enqueue workers
    DeviceWorkers workers;          // devs.size() - 1 threads (device 0 is served by the calling thread)
    bool has_scene = false;
    bool tiny = false;              // scene fits the kernel-parameter block (TinySceneData)
    int path = PATH_TINY;           // how rt_render traces spheres
    bool has_bvh = false;
    bool has_shadow_grids = false;
    bool host_shadow_bins = false;       // RT_OPT_HOST_SHADOW_BINS: build the bins on the host (tests: device build == host build)
    bool primary_bins = RT_PRIMARY_BINS_DEFAULT != 0;   // RT_OPT_PRIMARY_BINS
    std::atomic<uint64_t> pb_builds{0};  // primary-bin builds so far (rt_get_info)
    uint64_t sg_build_ns = 0;            // wall time of the last shadow-bin (re)build (rt_get_info)
    std::vector<f4> host_sgeom; std::vector<f3> host_lights;     // kept for rt_update_spheres (shadow bins are rebuilt on the host)
    f3 sg_lo, sg_hi;
    TinySceneData tiny_data;
    GlobalSceneData gdata_host;     // counts + ambient (pointers per device filled at launch)
    int accel = RT_ACCEL_BRUTE;
    int rank = 0, world = 1, tile_rows = 8;
    bool compaction = false;        // RT_OPT_COMPACTION
    bool host_via_gpu0 = false;     // RT_OPT_HOST_VIA_GPU0
    bool primary_gate = true;       // RT_OPT_PRIMARY_GATE
    int shared_target = 0;          // RT_OPT_SHARED_TARGET (0 / 1 = above 4 ranks / 2 = always)
    // last frame gates computed (a camera that does not move, batches of equal cameras): guarded, launches may come from workers
    std::mutex gate_mutex; bool gate_valid = false; CamRec gate_cam; int gate_w = 0, gate_h = 0; FrameGates gate_last;
    bool peer_ok = false;
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> gate_host_ns{0}, gate_computes{0};   // host time of compute_frame_gates, number of evaluations (rt_get_info)
    bool debug_shipped = false;     // RT_OPT_DEBUG_SHIPPED
    bool sparse_d2h = true;         // RT_OPT_SPARSE_D2H
    bool host_precleared = false;   // RT_OPT_HOST_PRECLEARED
    // packed multi-GPU gather (rt_gather.cuh)
    int gather_mode = -1;           // RT_OPT_GATHER_MODE: -1 auto (2 from 8 ranks on), 0 off, 1 RGB24, 2 RGB24 + grey quads
    int sink_tiles = 0, peer_tiles = 0;   // RT_OPT_SINK_TILES / RT_OPT_PEER_TILES: 0 = automatic per world size
    unsigned char* gather_area = nullptr; uint64_t gather_bytes = 0; bool gather_owned = false;    // as this context addresses it
    std::vector<GatherLocal*> gather_local;     // per device, in its own memory
    uint64_t gather_epoch = 0;
    bool gather_active = false;     // the last launch group used the packed gather (rt_get_info)
    int zero_copy = 1;              // RT_OPT_HOST_ZERO_COPY: 0 never, 1 frames up to 40 MB, 2 always
    FillPool fill_pool;
    uint64_t last_d2h_bytes = 0;    // bytes the last host-returning render really copied device -> host (rt_get_info)
    uint64_t last_enqueue_ns = 0, last_total_ns = 0;       // host time of the last render_frames: entry -> everything enqueued, entry -> return
    uint64_t last_fill_bytes = 0, last_fill_wait_ns = 0;   // host zero fill of that render: bytes, and time the caller waited for it
    std::vector<void*> gl_resources;   // cudaGraphicsResource* registered through rt_gl_register_buffer
};

namespace {

std::string g_create_err;
std::mutex g_err_mu;

int fail(rt_context* ctx, int code, const std::string& msg) {
    if (ctx) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = msg; }
    else { std::lock_guard<std::mutex> lk(g_err_mu); g_create_err = msg; }
    if (getenv("RT_LOG")) fprintf(stderr, "[rtb200] error %d: %s\n", code, msg.c_str());
    return code;
}
// Scratch device allocation of one ABI call: released on every return path (CU_TRY returns early).
template <class T> struct DevMem {
    T* p = nullptr;
    DevMem() = default;
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
    ~DevMem() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, count * sizeof(T)); }
};

#define CU_TRY(ctx, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, RT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
    } while (0)

void free_scene(DeviceState& d) {
    cudaSetDevice(d.dev);
    d.bvh.release(); d.bvh_cam_valid = false;
    d.pb.release(); d.pb_valid = false;
    d.sg.release();
    cudaFree(d.sgeom); cudaFree(d.smat); cudaFree(d.planes); cudaFree(d.lights);
    d.sgeom = nullptr; d.smat = nullptr; d.planes = nullptr; d.lights = nullptr;
}

int ensure_fb(rt_context* ctx, size_t pixels, int dev_index = 0) {
    DeviceState& d0 = ctx->devs[(size_t)dev_index];
    if (d0.fb_pixels >= pixels) return RT_OK;
    CU_TRY(ctx, cudaSetDevice(d0.dev));
    if (d0.fb) CU_TRY(ctx, cudaFree(d0.fb));
    d0.fb = nullptr; d0.fb_pixels = 0;
    CU_TRY(ctx, cudaMalloc(&d0.fb, pixels * sizeof(uint32_t)));
    d0.fb_pixels = pixels;
    return RT_OK;
}

CamRec to_cam(const rt_camera& c) {
    CamRec r;
    r.pos = mk3(c.pos[0], c.pos[1], c.pos[2]); r.right = mk3(c.right[0], c.right[1], c.right[2]);
    r.up = mk3(c.up[0], c.up[1], c.up[2]); r.fwd = mk3(c.forward[0], c.forward[1], c.forward[2]);
    r.view = mk3(c.view_params[0], c.view_params[1], c.view_params[2]);
    return r;
}

int check_frame_args(rt_context* ctx, const void* cam, int w, int h, int depth, int spp) {
    if (!ctx) return RT_ERR_INVALID;
    if (!cam) return fail(ctx, RT_ERR_INVALID, "camera is NULL");
    if (w <= 0 || h <= 0) return fail(ctx, RT_ERR_INVALID, "width/height must be positive");
    if ((long long)w * h > 0x40000000LL) return fail(ctx, RT_ERR_UNSUPPORTED, "frames above 2^30 pixels are not supported");
    if (depth < 0 || depth > RT_MAX_DEPTH) return fail(ctx, RT_ERR_UNSUPPORTED, "max_depth must be in [0, 32]");
    if (spp < 1 || spp > 1024) return fail(ctx, RT_ERR_INVALID, "spp must be in [1, 1024]");
    if (!ctx->has_scene) return fail(ctx, RT_ERR_NO_SCENE, "rt_set_scene has not been called");
    return RT_OK;
}

// Partition table of a launch: sink_tiles slots per period for rank 0, peer_tiles for every other rank (1, 1 = equal shares,
// period = world, tile t -> rank t % world). Rank 0's slots are spread evenly over the period (Bresenham), the others' dealt round robin.
struct PartTable { int period; int cnt[GATHER_MAX_RANKS]; unsigned char owner[PART_MAX_PERIOD]; };
PartTable make_part_table(int world, int sink_tiles, int peer_tiles) {
    PartTable t; memset(&t, 0, sizeof(t));
    if (world > GATHER_MAX_RANKS || sink_tiles < 0 || peer_tiles < 1 || sink_tiles > PART_MAX_CNT || peer_tiles > PART_MAX_CNT ||
        sink_tiles + peer_tiles * (world - 1) > PART_MAX_PERIOD || world < 2) { sink_tiles = 1; peer_tiles = 1; }
    if (world > PART_MAX_PERIOD) { t.period = 0; return t; }         // not representable: the caller falls back to t % world
    t.period = world < 2 ? 1 : sink_tiles + peer_tiles * (world - 1);
    int next_peer = 1;
    for (int j = 0; j < t.period; j++) {
        const bool sink = world < 2 || ((long long)(j + 1) * sink_tiles) / t.period > ((long long)j * sink_tiles) / t.period;
        int r = 0;
        if (!sink) { r = next_peer; next_peer = next_peer + 1 < world ? next_peer + 1 : 1; }
        t.owner[j] = (unsigned char)r;
        if (r < GATHER_MAX_RANKS) t.cnt[r]++;
    }
    return t;
}

FrameParams make_params(const rt_context* ctx, int w, int h, int depth, int spp, uint32_t seed, int n_frames,
                        int rank, int world, uint32_t* out, long long frame_stride, const PartTable* pt = nullptr) {
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.w = w; fp.h = h; fp.cap = depth; fp.spp = spp; fp.seed = seed; fp.n_frames = n_frames;
    fp.rank = rank; fp.world = world; fp.tile_rows = ctx->tile_rows;
    fp.tiles_total = (h + fp.tile_rows - 1) / fp.tile_rows;
    if (pt && pt->period > 0 && rank < GATHER_MAX_RANKS) {
        fp.part_period = pt->period; fp.part_cnt = 0;
        for (int j = 0; j < pt->period; j++) {
            fp.part_owner[j] = pt->owner[j];
            if (pt->owner[j] == rank && fp.part_cnt < PART_MAX_CNT) fp.part_pos[fp.part_cnt++] = (unsigned char)j;
        }
        const int full = fp.tiles_total / pt->period, rem = fp.tiles_total % pt->period;
        fp.tiles_mine = full * fp.part_cnt;
        for (int i = 0; i < fp.part_cnt; i++) if (fp.part_pos[i] < rem) fp.tiles_mine++;
        if (fp.part_cnt == 0) { fp.part_cnt = 1; fp.part_pos[0] = 0; fp.tiles_mine = 0; }     // a rank without slots renders nothing
    } else {                                                         // equal shares: tile t -> rank t % world
        fp.part_period = world; fp.part_cnt = 1; fp.part_pos[0] = (unsigned char)rank;       // (part_owner is not needed: no expand pass)
        fp.tiles_mine = fp.tiles_total > rank ? (fp.tiles_total - rank + world - 1) / world : 0;
    }
    const int chunk = BLOCK * (ctx->path == PATH_TINY ? PPT_TINY : PPT_HEAVY);
    fp.chunks_per_tile = (int)(((long long)fp.tile_rows * w + chunk - 1) / chunk);
    // 2-D pixel blocks (render_loop): heavy paths always; the tiny path when a row is a whole number of 4-pixel spans (the spans —
    // and with them every black-span / host-fill decision, k_fill_black, the packed gather — are then the same as in linear order)
    static const bool no_tile2d = getenv("RTB200_NO_TILE2D") != nullptr;
    const int ppt = ctx->path == PATH_TINY ? PPT_TINY : PPT_HEAVY;
    if (BLOCK == 128 && fp.tile_rows % 8 == 0 && w % ppt == 0 && (ppt == 1 || ppt == 4) && !no_tile2d) {
        fp.tile2d = 1;
        fp.chunks_per_tile = tile2d_items_per_tile(ppt, w, fp.tile_rows);      // == (tile_rows / 4) * (w / 128) for the packed gather's items
    }
    fp.rcp_w = 1.0f / (float)w; fp.rcp_h = 1.0f / (float)h;      // host fp32 division: IEEE
    fp.frame_stride = frame_stride;
    fp.out = out;
    return fp;
}

// (Re)builds the per-light shadow bins from ctx->host_sgeom / host_lights on every device. The host only sizes the grids (sg_setup:
// bounds of the centres); binning runs on the GPU from the sphere records already there (rt_shadow_grid_build.cuh) — 100 k spheres
// x 4 lights in well under a millisecond, so rt_update_spheres can run per frame. RTB200_SG_HOST=1 (or RT_OPT_HOST_SHADOW_BINS)
// selects the host build (same geometry code, same bins) and uploads it.
int upload_shadow_grids(rt_context* ctx) {
    ctx->has_shadow_grids = false;
    if (!ctx->has_bvh || getenv("RTB200_NO_SHADOW_GRID")) return RT_OK;
    const auto t0 = std::chrono::steady_clock::now();
    if (ctx->host_shadow_bins || getenv("RTB200_SG_HOST")) {
        ShadowGridsHost h;
        shadow_grids_build(ctx->host_sgeom, ctx->host_lights, &h);
        if (h.empty() || h.cell_start.empty()) return RT_OK;
        for (auto& d : ctx->devs) {
            CU_TRY(ctx, cudaSetDevice(d.dev));
            CU_TRY(ctx, d.sg.upload(h, d.stream));
        }
        ctx->sg_lo = h.lo; ctx->sg_hi = h.hi;
    } else {
        SgSetup su;
        sg_setup(ctx->host_sgeom.data(), (int)ctx->host_sgeom.size(), ctx->host_lights, &su);
        if (!su.ok || su.total_cells <= 0) return RT_OK;
        for (auto& d : ctx->devs) {
            CU_TRY(ctx, cudaSetDevice(d.dev));
            uint64_t nl = 0;
            cudaError_t e = d.sg.build(d.sgeom, (int)ctx->host_sgeom.size(), su, d.stream, &nl);
            if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("shadow bins: ") + cudaGetErrorString(e));
            ctx->launches += nl;
        }
        ctx->sg_lo = su.lo; ctx->sg_hi = su.hi;
    }
    ctx->sg_build_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    ctx->has_shadow_grids = true;
    return RT_OK;
}

GlobalSceneData global_data(const rt_context* ctx, const DeviceState& d) {
    GlobalSceneData g = ctx->gdata_host;
    g.sgeom = d.sgeom; g.smat = d.smat; g.planes = d.planes; g.lights = d.lights;
    return g;
}
LbvhSceneData lbvh_data(const rt_context* ctx, const DeviceState& d, bool with_cam_boxes) {
    LbvhSceneData l;
    memset(&l, 0, sizeof(l));
    l.g = global_data(ctx, d);
    l.bv.nodes = d.bvh.nodes; l.bv.nodes_cam = with_cam_boxes ? d.bvh.nodes_cam : nullptr; l.bv.sgeom_sorted = d.bvh.sorted; l.bv.orig = d.bvh.orig; l.bv.n = d.bvh.n; l.bv.r2max = d.bvh.r2max;
    const bool grids = ctx->has_shadow_grids && with_cam_boxes;      // free single-ray queries always traverse
    l.sg.grids = grids ? d.sg.grids : nullptr; l.sg.cell_start = d.sg.cells; l.sg.items = d.sg.items;
    l.sg.lo = ctx->sg_lo; l.sg.hi = ctx->sg_hi;
    return l;
}

LbvhBinsSceneData lbvh_bins_data(const rt_context* ctx, const DeviceState& d) {
    LbvhBinsSceneData b;
    memset(&b, 0, sizeof(b));
    b.l = lbvh_data(ctx, d, true);
    b.pb = d.pb.view();
    return b;
}
// Per-frame state of the LBVH path on device d, (re)built on `stream` when the camera changed: the camera-inflated copy of the node
// boxes (depends on the camera position) and, with RT_OPT_PRIMARY_BINS on single-sample frames, the primary bins (depend on the whole
// camera and the frame size). Both are one set of buffers per device: a rebuild on another stream than the last LBVH launch first
// waits for that launch. *use_bins: the frame's primary rays may use the bins.
int prepare_lbvh_frame(rt_context* ctx, DeviceState& d, const CamRec& c, int w, int h, int spp, cudaStream_t stream, bool* use_bins) {
    *use_bins = false;
    const bool boxes_ok = d.bvh_cam_valid && d.bvh_cam_stream == stream && d.bvh_cam[0] == c.pos.x && d.bvh_cam[1] == c.pos.y && d.bvh_cam[2] == c.pos.z;
    const bool want_bins = ctx->primary_bins && spp == 1 && d.bvh.n >= 2;
    const bool bins_ok = want_bins && d.pb_valid && d.pb_stream == stream && d.pb_w == w && d.pb_h == h && memcmp(&d.pb_cam, &c, sizeof(CamRec)) == 0;
    if ((!boxes_ok || (want_bins && !bins_ok)) && d.bvh_done_valid && d.bvh_last_stream != stream)
        CU_TRY(ctx, cudaStreamWaitEvent(stream, d.bvh_done, 0));
    if (!boxes_ok) {
        cudaError_t e = d.bvh.refit_for_camera(c.pos.x, c.pos.y, c.pos.z, stream);
        if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("lbvh refit: ") + cudaGetErrorString(e));
        d.bvh_cam[0] = c.pos.x; d.bvh_cam[1] = c.pos.y; d.bvh_cam[2] = c.pos.z; d.bvh_cam_valid = true; d.bvh_cam_stream = stream;
        ctx->launches++;
    }
    if (want_bins && !bins_ok) {
        const PbCam pc = make_pb_cam(c, w, h);
        d.pb_usable = pc.eps >= 0;                       // a camera the gate derivation does not cover: this frame traverses
        if (d.pb_usable) {
            uint64_t launches = 0;
            cudaError_t e = d.pb.build(d.sgeom, ctx->gdata_host.ns, pc, stream, &launches);
            if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("primary bins: ") + cudaGetErrorString(e));
            ctx->launches += launches;
            ctx->pb_builds++;
        }
        d.pb_cam = c; d.pb_w = w; d.pb_h = h; d.pb_valid = true; d.pb_stream = stream;
    }
    *use_bins = want_bins && d.pb_usable;
    return RT_OK;
}

// Per-frame gates of a tiny-scene launch (rt_gate.cuh): computed on the host for each distinct camera; a camera that does not move
// (and batches of equal cameras) reuse the last result. Host time spent here is accumulated in ctx->gate_host_ns (rt_get_info).
FrameGates gates_for(rt_context* ctx, const CamRec& cam, int w, int h) {
    const TinySceneData& t = ctx->tiny_data;
    std::lock_guard<std::mutex> lock(ctx->gate_mutex);
    if (!(ctx->gate_valid && ctx->gate_w == w && ctx->gate_h == h && memcmp(&ctx->gate_cam, &cam, sizeof(CamRec)) == 0)) {
        const auto t0 = std::chrono::steady_clock::now();
        ctx->gate_last = compute_frame_gates(cam, w, h, t.sgeom, t.ns, t.planes, t.np, t.lights, t.nl);
        ctx->gate_host_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        ctx->gate_computes++;
        ctx->gate_cam = cam; ctx->gate_w = w; ctx->gate_h = h; ctx->gate_valid = true;
    }
    return ctx->gate_last;
}
void fill_gates(rt_context* ctx, FrameParams& gp) {
    for (int f = 0; f < gp.n_frames; f++)
        gp.gates[f] = (!ctx->primary_gate || gp.spp != 1) ? gates_off(gp.w, gp.h) : gates_for(ctx, gp.cam_inline[f], gp.w, gp.h);
}

// Does a launch that stores into rank 0's framebuffer (shared target) use the packed gather? If so: the partition table to render
// with and the GatherParams of this device for `epoch`. Every rank evaluates this on identical inputs and must agree.
struct GatherPlan { bool on = false; PartTable pt; GatherParams gp; };
GatherPlan plan_gather(rt_context* ctx, int dev_index, int w, int h, int spp, int world, int shared_target, uint64_t epoch) {
    GatherPlan g;
    memset(&g.gp, 0, sizeof(g.gp)); memset(&g.pt, 0, sizeof(g.pt));
    int mode = ctx->gather_mode;
    // automatic (measured: profiles/r02/multi4/, push_gather/): below 8 ranks off — at N = 4 the plain stores (659 Grays/s) beat
    // RGB24 (607; its render kernel keeps the 2-D pixel blocks) and RGB24 + grey (575; 128-pixel strips for the per-warp flag
    // byte): the expand pass and the per-CTA fences cost more than the link saves; 8 ranks: RGB24 + grey quads — GPU 0's NVLink
    // ingress is the limit there
    if (mode < 0) mode = world >= 8 ? 2 : 0;
    if (getenv("RTB200_GATHER_MODE")) mode = atoi(getenv("RTB200_GATHER_MODE"));
    if (mode <= 0 || !shared_target || world < 2 || world > GATHER_MAX_RANKS || ctx->path != PATH_TINY || !ctx->primary_gate || spp != 1 ||
        w > RT_FASTDIV_MAX || h > RT_FASTDIV_MAX || (w % 128) != 0 || !ctx->gather_area ||
        ctx->gather_bytes < gather_area_bytes(w, h) || (size_t)dev_index >= ctx->gather_local.size() || !ctx->gather_local[(size_t)dev_index])
        return g;
    int st = ctx->sink_tiles, ptl = ctx->peer_tiles;
    if (getenv("RTB200_SINK_TILES")) st = atoi(getenv("RTB200_SINK_TILES"));
    if (getenv("RTB200_PEER_TILES")) ptl = atoi(getenv("RTB200_PEER_TILES"));
    if (ptl <= 0) {                                               // automatic: rank 0's share shrinks as its expand work grows
        if (world >= 8) { st = 1; ptl = 2; }                      // 1 / 15 of the tiles
        else if (world >= 4) { st = 4; ptl = 5; }                 // 4 / 19
        else { st = 1; ptl = 1; }
    }
    g.pt = make_part_table(world, st, ptl);
    if (g.pt.period <= 0) return g;
    const unsigned long long npix = (unsigned long long)w * (unsigned long long)h;
    g.gp.area = ctx->gather_area; g.gp.local = ctx->gather_local[(size_t)dev_index]; g.gp.epoch = epoch;
    g.gp.stride_c = 3 * npix; g.gp.stride_g = npix; g.gp.stride_f = (npix + 127) / 128 + 64;
    g.gp.off_c = GATHER_CTL_BYTES; g.gp.off_g = g.gp.off_c + GATHER_SLOTS * g.gp.stride_c; g.gp.off_f = g.gp.off_g + GATHER_SLOTS * g.gp.stride_g;
    g.gp.grey = mode >= 2 ? 1 : 0;
    g.on = true;
    return g;
}
void free_gather_local(rt_context* ctx) {
    for (size_t i = 0; i < ctx->gather_local.size(); i++)
        if (ctx->gather_local[i]) { cudaSetDevice(ctx->devs[i].dev); cudaFree(ctx->gather_local[i]); }
    ctx->gather_local.clear();
}
// Per-device GatherLocal blocks (zeroed) for the context's devices.
int ensure_gather_local(rt_context* ctx) {
    if (ctx->gather_local.size() == ctx->devs.size()) return RT_OK;
    ctx->gather_local.assign(ctx->devs.size(), nullptr);
    for (size_t i = 0; i < ctx->devs.size(); i++) {
        CU_TRY(ctx, cudaSetDevice(ctx->devs[i].dev));
        CU_TRY(ctx, cudaMalloc(&ctx->gather_local[i], sizeof(GatherLocal)));
        CU_TRY(ctx, cudaMemset(ctx->gather_local[i], 0, sizeof(GatherLocal)));
    }
    return RT_OK;
}

// Launches the render kernel for one device's share. Asynchronous on `stream`.
int launch_render(rt_context* ctx, DeviceState& d, const FrameParams& fp, cudaStream_t stream) {
    if (ctx->path == PATH_LBVH && fp.n_frames > 1) {
        // the camera-inflated BVH copy is per frame: one refit + one launch per frame, back to back on the stream
        for (int f = 0; f < fp.n_frames; f++) {
            FrameParams one = fp;
            one.n_frames = 1;
            one.cam_inline[0] = fp.cam_inline[f];
            one.out = fp.out + (long long)f * fp.frame_stride;
            int rc = launch_render(ctx, d, one, stream);
            if (rc) return rc;
        }
        return RT_OK;
    }
    const bool gather_sink = fp.gather.area != nullptr && fp.rank == 0;      // rank 0 of a packed gather runs the expand pass even without own tiles
    if ((fp.tiles_mine <= 0 && !gather_sink) || fp.chunks_per_tile <= 0 || fp.n_frames <= 0) return RT_OK;
    if (fp.n_frames > 65535) return fail(ctx, RT_ERR_UNSUPPORTED, "more than 65535 frames in one launch");
    bool use_bins = false;
    if (ctx->path == PATH_LBVH) {
        int rc = prepare_lbvh_frame(ctx, d, fp.cam_inline[0], fp.w, fp.h, fp.spp, stream, &use_bins);
        if (rc) return rc;
    }
    const size_t smem = ctx->path == PATH_STAGED ? sizeof(f4) * (size_t)ctx->gdata_host.ns : 0;
    const dim3 grid((unsigned)fp.chunks_per_tile, (unsigned)(fp.tiles_mine < 65535 ? fp.tiles_mine : 65535), (unsigned)fp.n_frames);
    switch (ctx->path) {
        case PATH_TINY: {
            const TinySceneData& t = ctx->tiny_data;
            const bool fastdiv_ok = fp.w <= RT_FASTDIV_MAX && fp.h <= RT_FASTDIV_MAX;
            FrameParams gp = fp;                      // + per-frame gates (a few microseconds of host work per distinct camera)
            fill_gates(ctx, gp);
            // Tail of the launch: CTAs are dispatched in grid order and a chunk on the mirror sphere costs ~10x an average one, so a
            // launch that ends with expensive chunks leaves most SMs idle while the last ones finish — a fixed ~45 us per launch,
            // 15 % of a 1/8 share of the 16-frame step (profiles/r02/gather_probe_launches.csv). Tiles are therefore visited starting
            // at the top of the sphere rectangle of the launch's LAST frame and wrapping around: spheres first, then the floor below
            // them (plentiful, uniform), the sky above last.
            if (ctx->primary_gate && fp.spp == 1 && fp.tiles_mine > 1 && fp.tiles_total > 0 && !getenv("RTB200_NO_TILE_ORDER")) {
                const GateRect& r = gp.gates[fp.n_frames - 1].spheres;
                if (r.y0 > 0 && r.y0 < fp.h && r.y1 >= r.y0) {
                    const long long t0 = r.y0 / fp.tile_rows;
                    long long rot = t0 * (fp.k_begin + fp.tiles_mine) / fp.tiles_total - fp.k_begin;      // ~ this rank's first tile at / below t0
                    if (rot < 0) rot = 0;
                    if (rot >= fp.tiles_mine) rot = 0;
                    gp.tile_rot = (int)rot;
                }
            }
            // the compacting variant knows no black spans: a launch that takes part in a sparse gather uses the default kernel on
            // EVERY rank, so that a rank with RT_OPT_COMPACTION set differently cannot leave spans nobody writes
            const bool compact = ctx->compaction && fp.spp == 1 && fastdiv_ok && !fp.skip_black_store;
            TinyKernel kern = compact ? tiny_kernel_compact(t.ns, t.nl, t.np) : tiny_kernel(t.ns, t.nl, t.np, fp.spp, fastdiv_ok);
            // Sparse gather: all ranks store into ONE framebuffer owned by rank 0 (fp.skip_black_store = the caller's promise).
            // Only with the gated single-sample kernels, which are the ones that know black spans.
            // It pays off only where rank 0's NVLink ingress is the bottleneck (measured: N = 8); with fewer ranks rank 0 itself is
            // the critical path and its fill pass (55 us per 16 4K frames) costs more than the link saves -> automatic above 4 ranks,
            // promise value 2 forces it (tests).
            if (fp.gather.area) {
                // Packed gather (rt_gather.cuh; plan_gather has checked that it applies): ranks != 0 render into the planes in GPU 0's
                // memory; rank 0 renders its own (smaller) share straight into the framebuffer, then expands the others' tiles.
                // a few hundred CTAs per frame (each ends with a system-scope fence): ~6 items per CTA, between 1 and 4 per SM
                auto fat_grid = [&](long long items) {
                    long long c = items / 6; const long long lo = d.sm_count, hi = 4LL * d.sm_count;
                    if (c < lo) c = lo; if (c > hi) c = hi; if (c > items) c = items; if (c < 1) c = 1;
                    return dim3((unsigned)c, 1u, (unsigned)fp.n_frames);
                };
                if (fp.rank != 0) {
                    tiny_kernel_pack(t.ns, t.nl, t.np)<<<fat_grid((long long)fp.tiles_mine * fp.chunks_per_tile), BLOCK, 0, stream>>>(t, gp);
                } else {
                    gp.skip_black_store = 0;
                    if (fp.tiles_mine > 0) {
                        kern<<<grid, BLOCK, 0, stream>>>(t, gp);
                        CU_TRY(ctx, cudaGetLastError());
                        ctx->launches++;
                    }
                    k_gather_expand<<<fat_grid((long long)fp.tiles_total * fp.chunks_per_tile), BLOCK, 0, stream>>>(gp);
                }
                break;
            }
            static const int sparse_min_world = getenv("RTB200_SPARSE_MIN_WORLD") ? atoi(getenv("RTB200_SPARSE_MIN_WORLD")) : 5;
            const bool sparse = fp.skip_black_store && (fp.world >= sparse_min_world || fp.skip_black_store == 2) && fp.world > 1 && ctx->primary_gate &&
                                fp.spp == 1 && fastdiv_ok && !compact;
            gp.skip_black_store = sparse && fp.rank != 0;
            if (sparse && fp.rank == 0 && fp.tiles_mine > 0) {
                // The fill touches only the other ranks' tiles, the render kernel only rank 0's: they run side by side (one is bound
                // by HBM writes, the other by issue slots). Stream order is kept on both sides: the fill starts after everything
                // already queued on `stream` (the previous consumer of the framebuffer), and `stream` continues after it.
                CU_TRY(ctx, cudaEventRecord(d.fill_go, stream));
                CU_TRY(ctx, cudaStreamWaitEvent(d.fill_stream, d.fill_go, 0));
                k_fill_black<<<grid, BLOCK, 0, d.fill_stream>>>(gp);
                CU_TRY(ctx, cudaGetLastError());
                CU_TRY(ctx, cudaEventRecord(d.fill_done, d.fill_stream));
                ctx->launches++;
                kern<<<grid, BLOCK, 0, stream>>>(t, gp);
                CU_TRY(ctx, cudaStreamWaitEvent(stream, d.fill_done, 0));
            } else {
                kern<<<grid, BLOCK, 0, stream>>>(t, gp);
            }
            break;
        }
        case PATH_STAGED: k_render_staged<<<grid, BLOCK, smem, stream>>>(global_data(ctx, d), fp); break;
        case PATH_GLOBAL: k_render_global<<<grid, BLOCK, 0, stream>>>(global_data(ctx, d), fp); break;
        default:
            if (use_bins) k_render_lbvh_bins<<<grid, BLOCK, 0, stream>>>(lbvh_bins_data(ctx, d), fp);
            else k_render_lbvh<<<grid, BLOCK, 0, stream>>>(lbvh_data(ctx, d, true), fp);
            break;
    }
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (ctx->path == PATH_LBVH) {
        CU_TRY(ctx, cudaEventRecord(d.bvh_done, stream));
        d.bvh_done_valid = true; d.bvh_last_stream = stream;
    }
    return RT_OK;
}

}  // namespace

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

int rt_abi_version(void) { return RTB200_ABI_VERSION; }

const char* rt_last_error(const rt_context* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mu);
    return g_create_err.c_str();
}

int rt_create(rt_context** out, const int* device_ids, int n_devices) {
    if (!out) return fail(nullptr, RT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!(n_devices == 1 || n_devices == 2 || n_devices == 4 || n_devices == 8))
        return fail(nullptr, RT_ERR_INVALID, "n_devices must be 1, 2, 4 or 8");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, RT_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (librtb200 has no CPU fallback)");
    rt_context* ctx = new rt_context();
    if (getenv("RTB200_PRIMARY_BINS")) ctx->primary_bins = atoi(getenv("RTB200_PRIMARY_BINS")) != 0;
    ctx->devs.resize((size_t)n_devices);
    for (int i = 0; i < n_devices; i++) {
        int dev = device_ids ? device_ids[i] : i;
        if (dev < 0 || dev >= count) { delete ctx; return fail(nullptr, RT_ERR_INVALID, "device id out of range"); }
        DeviceState& d = ctx->devs[(size_t)i];
        d.dev = dev;
        cudaDeviceProp prop;
        if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&d.fill_stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.fill_go, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.fill_done, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreate(&d.evc0)) != cudaSuccess || (e = cudaEventCreate(&d.evc1)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.bvh_done, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreate(&d.ev0)) != cudaSuccess || (e = cudaEventCreate(&d.ev1)) != cudaSuccess) {
            std::string m = std::string("device init: ") + cudaGetErrorString(e);
            delete ctx; return fail(nullptr, RT_ERR_CUDA, m);
        }
        d.sm_count = prop.multiProcessorCount;
    }
    // Peer access so that devices 1..n-1 can store straight into device 0's framebuffer (fused gather over NVLink).
    ctx->peer_ok = true;
    for (int i = 1; i < n_devices; i++) {
        // a device id may repeat: several partitions of one context on ONE physical GPU (each with its own streams, scene copy and
        // launch thread) — how the single-GPU test run exercises the multi-device code; memory of the same device needs no peer mapping
        if (ctx->devs[(size_t)i].dev == ctx->devs[0].dev) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->devs[(size_t)i].dev, ctx->devs[0].dev);
        if (!can) { ctx->peer_ok = false; break; }
        cudaSetDevice(ctx->devs[(size_t)i].dev);
        e = cudaDeviceEnablePeerAccess(ctx->devs[0].dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->peer_ok = false; break; }
        cudaGetLastError();
    }
    if (n_devices > 1 && !ctx->peer_ok) { delete ctx; return fail(nullptr, RT_ERR_PEER, "peer access to device 0 unavailable"); }
    if (n_devices > 1) ctx->workers.start(n_devices - 1);
    *out = ctx;
    return RT_OK;
}

int rt_destroy(rt_context* ctx) {
    if (!ctx) return RT_ERR_INVALID;
    ctx->workers.stop();
    ctx->fill_pool.stop();
    for (auto& d : ctx->devs) { cudaSetDevice(d.dev); cudaDeviceSynchronize(); }
    free_gather_local(ctx);
    if (ctx->gather_owned && ctx->gather_area) { cudaSetDevice(ctx->devs[0].dev); cudaFree(ctx->gather_area); }
    for (void* r : ctx->gl_resources) cudaGraphicsUnregisterResource((cudaGraphicsResource_t)r);
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.dev);
        cudaStreamSynchronize(d.stream);
        free_scene(d);
        cudaFree(d.fb);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.evc0) cudaEventDestroy(d.evc0);
        if (d.evc1) cudaEventDestroy(d.evc1);
        if (d.bvh_done) cudaEventDestroy(d.bvh_done);
        for (cudaEvent_t ev : d.band_events) cudaEventDestroy(ev);
        if (d.fill_go) cudaEventDestroy(d.fill_go);
        if (d.fill_done) cudaEventDestroy(d.fill_done);
        if (d.fill_stream) cudaStreamDestroy(d.fill_stream);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
    return RT_OK;
}

int rt_set_scene(rt_context* ctx, const float* spheres, int ns, const float* planes, int np, const float* lights, int nl,
                 const float ambient[3], int accel) {
    if (!ctx) return RT_ERR_INVALID;
    if (ns < 0 || np < 0 || nl < 0 || (ns > 0 && !spheres) || (np > 0 && !planes) || (nl > 0 && !lights) || !ambient)
        return fail(ctx, RT_ERR_INVALID, "bad scene arrays");
    if (accel < RT_ACCEL_AUTO || accel > RT_ACCEL_LBVH) return fail(ctx, RT_ERR_INVALID, "bad accel");
    ctx->has_scene = false;
    std::vector<f4> sg((size_t)ns); std::vector<MatRec> sm((size_t)ns);
    std::vector<PlaneRec> pl((size_t)np); std::vector<LightRec> li((size_t)nl);
    for (int i = 0; i < ns; i++) {
        const float* f = spheres + 18 * (size_t)i;
        sg[(size_t)i].x = f[0]; sg[(size_t)i].y = f[1]; sg[(size_t)i].z = f[2]; sg[(size_t)i].w = f[17];   // radiusSquared passed through (:336)
        sm[(size_t)i] = make_mat(f + 4);
    }
    for (int i = 0; i < np; i++) pl[(size_t)i] = make_plane(planes + 20 * (size_t)i);
    for (int i = 0; i < nl; i++) li[(size_t)i] = make_light(lights + 4 * (size_t)i);

    ctx->tiny = (ns <= TINY_MAX_SPHERES && np <= TINY_MAX_PLANES && nl <= TINY_MAX_LIGHTS);
    memset(&ctx->tiny_data, 0, sizeof(ctx->tiny_data));
    if (ctx->tiny) {
        TinySceneData& t = ctx->tiny_data;
        t.ns = ns; t.np = np; t.nl = nl; t.amb = mk3(ambient[0], ambient[1], ambient[2]);
        for (int i = 0; i < ns; i++) { t.sgeom[i] = sg[(size_t)i]; t.smat[i] = sm[(size_t)i]; }
        for (int i = 0; i < np; i++) t.planes[i] = pl[(size_t)i];
        for (int i = 0; i < nl; i++) t.lights[i] = li[(size_t)i];
        tiny_fill_pairs(t);
    }
    memset(&ctx->gdata_host, 0, sizeof(ctx->gdata_host));
    ctx->gdata_host.ns = ns; ctx->gdata_host.np = np; ctx->gdata_host.nl = nl;
    ctx->gdata_host.amb = mk3(ambient[0], ambient[1], ambient[2]);
    // The global-memory copy is always uploaded (rt_query_spheres and the debug kernels of non-tiny scenes use it).
    for (auto& d : ctx->devs) {
        CU_TRY(ctx, cudaSetDevice(d.dev));
        free_scene(d);
        CU_TRY(ctx, cudaMalloc(&d.sgeom, sizeof(f4) * (size_t)(ns > 0 ? ns : 1)));
        CU_TRY(ctx, cudaMalloc(&d.smat, sizeof(MatRec) * (size_t)(ns > 0 ? ns : 1)));
        CU_TRY(ctx, cudaMalloc(&d.planes, sizeof(PlaneRec) * (size_t)(np > 0 ? np : 1)));
        CU_TRY(ctx, cudaMalloc(&d.lights, sizeof(LightRec) * (size_t)(nl > 0 ? nl : 1)));
        // Stream-ordered uploads: the context's streams are non-blocking, so a default-stream cudaMemcpy from pageable
        // memory (which may return before its DMA lands) would NOT be ordered before kernels launched on d.stream.
        if (ns) CU_TRY(ctx, cudaMemcpyAsync(d.sgeom, sg.data(), sizeof(f4) * (size_t)ns, cudaMemcpyHostToDevice, d.stream));
        if (ns) CU_TRY(ctx, cudaMemcpyAsync(d.smat, sm.data(), sizeof(MatRec) * (size_t)ns, cudaMemcpyHostToDevice, d.stream));
        if (np) CU_TRY(ctx, cudaMemcpyAsync(d.planes, pl.data(), sizeof(PlaneRec) * (size_t)np, cudaMemcpyHostToDevice, d.stream));
        if (nl) CU_TRY(ctx, cudaMemcpyAsync(d.lights, li.data(), sizeof(LightRec) * (size_t)nl, cudaMemcpyHostToDevice, d.stream));
        CU_TRY(ctx, cudaStreamSynchronize(d.stream));
    }
    // ---- how spheres are traced ----
    const bool want_bvh = ns >= 2 && (accel == RT_ACCEL_LBVH || (accel == RT_ACCEL_AUTO && ns >= AUTO_LBVH_MIN_SPHERES));
    ctx->has_bvh = false;
    if (want_bvh) {
        // bounds of the centres, effective radii (the reference's test only ever sees radiusSquared, :619) — O(n) on the host
        float bmin[3] = {sg[0].x, sg[0].y, sg[0].z}, bmax[3] = {sg[0].x, sg[0].y, sg[0].z}, r2max = 0.0f;
        for (int i = 0; i < ns; i++) {
            const f4& g = sg[(size_t)i];
            bmin[0] = fminf(bmin[0], g.x); bmin[1] = fminf(bmin[1], g.y); bmin[2] = fminf(bmin[2], g.z);
            bmax[0] = fmaxf(bmax[0], g.x); bmax[1] = fmaxf(bmax[1], g.y); bmax[2] = fmaxf(bmax[2], g.z);
            if (g.w > r2max) r2max = g.w;
        }
        for (auto& d : ctx->devs) {
            CU_TRY(ctx, cudaSetDevice(d.dev));
            uint64_t nl = 0;
            cudaError_t e = lbvh_build(d.sgeom, ns, bmin, bmax, r2max, d.stream, &d.bvh, &nl);
            if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("lbvh_build: ") + cudaGetErrorString(e));
            ctx->launches += nl;
        }
        ctx->has_bvh = true;
    }
    ctx->host_sgeom = sg;
    ctx->host_lights.resize((size_t)nl);
    for (int i = 0; i < nl; i++) ctx->host_lights[(size_t)i] = li[(size_t)i].p;
    { int rc = upload_shadow_grids(ctx); if (rc) return rc; }
    if (ctx->has_bvh) ctx->path = PATH_LBVH;
    else if (ctx->tiny) ctx->path = PATH_TINY;
    else if (ns <= STAGED_MAX_SPHERES) ctx->path = PATH_STAGED;
    else ctx->path = PATH_GLOBAL;
    if (ctx->path == PATH_STAGED) {
        for (auto& d : ctx->devs) {
            CU_TRY(ctx, cudaSetDevice(d.dev));
            CU_TRY(ctx, cudaFuncSetAttribute(k_render_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(f4) * STAGED_MAX_SPHERES)));
            CU_TRY(ctx, cudaFuncSetAttribute(k_debug_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(f4) * STAGED_MAX_SPHERES)));
        }
    }
    ctx->accel = accel;
    ctx->has_scene = true;
    { std::lock_guard<std::mutex> lock(ctx->gate_mutex); ctx->gate_valid = false; }
    return RT_OK;
}

int rt_update_spheres(rt_context* ctx, const float* spheres, int first, int count) {
    if (!ctx) return RT_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, RT_ERR_NO_SCENE, "rt_set_scene has not been called");
    const int ns = ctx->gdata_host.ns;
    if (count < 0 || first < 0 || first > ns || count > ns - first || (count > 0 && !spheres)) return fail(ctx, RT_ERR_INVALID, "sphere range out of bounds");
    if (count == 0) return RT_OK;
    std::vector<f4> sg((size_t)count); std::vector<MatRec> sm((size_t)count);
    float r2max = 0.0f;
    for (int i = 0; i < count; i++) {
        const float* f = spheres + 18 * (size_t)i;
        sg[(size_t)i].x = f[0]; sg[(size_t)i].y = f[1]; sg[(size_t)i].z = f[2]; sg[(size_t)i].w = f[17];
        sm[(size_t)i] = make_mat(f + 4);
        if (f[17] > r2max) r2max = f[17];
    }
    if (ctx->tiny) {
        for (int i = 0; i < count; i++) { ctx->tiny_data.sgeom[first + i] = sg[(size_t)i]; ctx->tiny_data.smat[first + i] = sm[(size_t)i]; }
        tiny_fill_pairs(ctx->tiny_data);
    }
    for (int i = 0; i < count; i++) ctx->host_sgeom[(size_t)(first + i)] = sg[(size_t)i];
    for (auto& d : ctx->devs) {
        CU_TRY(ctx, cudaSetDevice(d.dev));
        if (d.bvh_done_valid && d.bvh_last_stream != d.stream) CU_TRY(ctx, cudaStreamWaitEvent(d.stream, d.bvh_done, 0));
        CU_TRY(ctx, cudaMemcpyAsync(d.sgeom + first, sg.data(), sizeof(f4) * (size_t)count, cudaMemcpyHostToDevice, d.stream));
        CU_TRY(ctx, cudaMemcpyAsync(d.smat + first, sm.data(), sizeof(MatRec) * (size_t)count, cudaMemcpyHostToDevice, d.stream));
        if (ctx->has_bvh) {
            // Same topology, new leaf geometry and boxes (bottom-up refit). The pad bound only ever grows (conservative).
            if (r2max > d.bvh.r2max) d.bvh.r2max = r2max;
            cudaError_t e = d.bvh.refit_geometry(d.sgeom, d.stream);
            if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("lbvh refit: ") + cudaGetErrorString(e));
            d.bvh_cam_valid = false; d.pb_valid = false;
            ctx->launches++;
        }
        CU_TRY(ctx, cudaStreamSynchronize(d.stream));     // sg / sm go out of scope; later launches may use another stream
    }
    { std::lock_guard<std::mutex> lock(ctx->gate_mutex); ctx->gate_valid = false; }
    return upload_shadow_grids(ctx);                       // the bins depend on the sphere positions: rebuilt on the host
}

int rt_set_option(rt_context* ctx, int option, int value) {
    if (!ctx) return RT_ERR_INVALID;
    switch (option) {
        case RT_OPT_COMPACTION: ctx->compaction = value != 0; return RT_OK;
        case RT_OPT_HOST_VIA_GPU0: ctx->host_via_gpu0 = value != 0; return RT_OK;
        case RT_OPT_PRIMARY_GATE: ctx->primary_gate = value != 0; return RT_OK;
        case RT_OPT_SHARED_TARGET: ctx->shared_target = value < 0 ? 0 : (value > 2 ? 2 : value); return RT_OK;
        case RT_OPT_DEBUG_SHIPPED: ctx->debug_shipped = value != 0; return RT_OK;
        case RT_OPT_SPARSE_D2H: ctx->sparse_d2h = value != 0; return RT_OK;
        case RT_OPT_HOST_PRECLEARED: ctx->host_precleared = value != 0; return RT_OK;
        case RT_OPT_GATHER_MODE: ctx->gather_mode = value < -1 ? -1 : (value > 2 ? 2 : value); return RT_OK;
        case RT_OPT_SINK_TILES: ctx->sink_tiles = value < 0 ? 0 : value; return RT_OK;
        case RT_OPT_PEER_TILES: ctx->peer_tiles = value < 0 ? 0 : value; return RT_OK;
        case RT_OPT_HOST_SHADOW_BINS: ctx->host_shadow_bins = value != 0; return RT_OK;
        case RT_OPT_PRIMARY_BINS: ctx->primary_bins = value != 0; return RT_OK;
        case RT_OPT_HOST_ZERO_COPY: ctx->zero_copy = value < 0 ? 0 : (value > 2 ? 2 : value); return RT_OK;
        default: return fail(ctx, RT_ERR_INVALID, "unknown option");
    }
}

int rt_set_partition(rt_context* ctx, int rank, int world, int tile_rows) {
    if (!ctx) return RT_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world || tile_rows < 1 || tile_rows > 65536) return fail(ctx, RT_ERR_INVALID, "bad partition");
    if (ctx->devs.size() != 1 && world != 1) return fail(ctx, RT_ERR_INVALID, "rt_set_partition needs a single-device context");
    ctx->rank = rank; ctx->world = world; ctx->tile_rows = tile_rows;
    return RT_OK;
}

int rt_render_device(rt_context* ctx, const rt_camera* cams, int n_frames, int w, int h, int depth, int spp, uint32_t seed,
                     void* dev_pixels, void* cuda_stream) {
    int rc = check_frame_args(ctx, cams, w, h, depth, spp);
    if (rc) return rc;
    if (n_frames < 1 || !dev_pixels) return fail(ctx, RT_ERR_INVALID, "n_frames/dev_pixels");
    if (ctx->devs.size() != 1) return fail(ctx, RT_ERR_INVALID, "rt_render_device needs a single-device context");
    DeviceState& d = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    for (int f0 = 0; f0 < n_frames; f0 += INLINE_CAMS) {          // <= INLINE_CAMS frames per launch (cameras in the parameter block)
        const int nf = n_frames - f0 < INLINE_CAMS ? n_frames - f0 : INLINE_CAMS;
        const GatherPlan gp = plan_gather(ctx, 0, w, h, spp, ctx->world, ctx->shared_target, ctx->gather_epoch + 1);
        if (gp.on) ctx->gather_epoch++;
        ctx->gather_active = gp.on;
        FrameParams fp = make_params(ctx, w, h, depth, spp, seed, nf, ctx->rank, ctx->world,
                                     (uint32_t*)dev_pixels + (size_t)f0 * w * h, (long long)w * h, gp.on ? &gp.pt : nullptr);
        if (gp.on) fp.gather = gp.gp;
        for (int i = 0; i < nf; i++) fp.cam_inline[i] = to_cam(cams[f0 + i]);
        fp.skip_black_store = ctx->shared_target;
        rc = launch_render(ctx, d, fp, st);
        if (rc) return rc;
    }
    return RT_OK;
}

// HostSkip of a row plan: its two longest runs of ROW_BLACK rows and its longest run of ROW_RECT rows.
static HostSkip host_skip_of(const RowPlan& rp, int h) {
    HostSkip k; memset(&k, 0, sizeof(k));
    if (!rp.sparse) return k;
    int best_b[2][2] = {{0, 0}, {0, 0}}, best_r[2] = {0, 0};
    for (int y = 0; y < h;) {
        const uint8_t kind = rp.kind[(size_t)y];
        int y2 = y + 1;
        while (y2 < h && rp.kind[(size_t)y2] == kind) y2++;
        if (kind == ROW_BLACK) {
            if (y2 - y > best_b[0][1] - best_b[0][0]) { best_b[1][0] = best_b[0][0]; best_b[1][1] = best_b[0][1]; best_b[0][0] = y; best_b[0][1] = y2; }
            else if (y2 - y > best_b[1][1] - best_b[1][0]) { best_b[1][0] = y; best_b[1][1] = y2; }
        } else if (kind == ROW_RECT && y2 - y > best_r[1] - best_r[0]) { best_r[0] = y; best_r[1] = y2; }
        y = y2;
    }
    k.yb0 = best_b[0][0]; k.yb1 = best_b[0][1]; k.yc0 = best_b[1][0]; k.yc1 = best_b[1][1];
    k.yr0 = best_r[0]; k.yr1 = best_r[1]; k.rx0 = rp.rx0; k.rx1 = rp.rx1;
    return k;
}

// Zero-copy host return (RT_OPT_HOST_ZERO_COPY, default on): when host_pixels is page-locked memory the devices can address
// (rt_host_register, cudaHostAlloc), the render kernel's 128-bit stores go straight into it over PCIe — no device framebuffer, no copy
// engine, no bands, one launch per device and <= 16 frames — and what the frame gates prove black is neither stored nor copied: host
// threads zero-fill those rows / row parts while the kernel runs (HostSkip). Every device (or rank of a partition) writes its own row
// tiles over its own PCIe link. `dev_out` is the device-side address of host_pixels.
static int render_frames_zero_copy(rt_context* ctx, const rt_camera* cams, int n_frames, int w, int h, int depth, int spp, uint32_t seed,
                                   int32_t* host_pixels, uint32_t* dev_out, rt_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    auto since_enter_ns = [&]() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_enter).count(); };
    const size_t npix = (size_t)w * h;
    const int G = (int)ctx->devs.size();
    const int world = G > 1 ? G : ctx->world;
    const int tile_rows = ctx->tile_rows;
    std::vector<HostSkip> hs((size_t)n_frames);
    memset(hs.data(), 0, sizeof(HostSkip) * (size_t)n_frames);
    const bool gated = ctx->sparse_d2h && ctx->path == PATH_TINY && ctx->primary_gate && spp == 1 && w <= RT_FASTDIV_MAX && h <= RT_FASTDIV_MAX &&
                       !ctx->compaction;
    if (gated) {
        RowPlan rp;
        for (int f = 0; f < n_frames; f++) {
            plan_rows(gates_for(ctx, to_cam(cams[f]), w, h), w, h, &rp);
            hs[(size_t)f] = host_skip_of(rp, h);
        }
    }
    bool fill_started = false;
    struct FillGuard { FillPool* p; const bool* started; ~FillGuard() { if (p && *started) p->wait(); } } fill_guard{&ctx->fill_pool, &fill_started};
    ctx->last_fill_bytes = 0; ctx->last_fill_wait_ns = 0;
    auto start_fill = [&]() {
        if (!gated || ctx->host_precleared) return;
        std::vector<FillPool::Seg> segs;
        uint64_t total = 0;
        const bool own_only = G == 1 && world > 1;                      // one rank of a partition fills the skipped rows of ITS tiles
        auto add_rows = [&](int32_t* dst, int ya, int yb, int xa, int xb) {     // columns [xa, xb) of rows [ya, yb)
            if (xb <= xa) return;
            for (int y = ya; y < yb;) {
                if (own_only && (y / tile_rows) % world != ctx->rank) { y = (y / tile_rows + 1) * tile_rows; continue; }
                int y2 = yb;
                if (own_only) { const int te = (y / tile_rows + 1) * tile_rows; if (te < y2) y2 = te; }
                if (xa == 0 && xb == w) {
                    const size_t piece_rows = ((size_t)1 << 18) / (size_t)w + 1;            // ~1 MB pieces: load balance across the pool
                    for (int yy = y; yy < y2; yy += (int)piece_rows) {
                        const int ye = yy + (int)piece_rows < y2 ? yy + (int)piece_rows : y2;
                        segs.push_back({(char*)(dst + (size_t)yy * w), (size_t)(ye - yy) * w * 4}); total += (uint64_t)(ye - yy) * w * 4;
                    }
                } else {
                    for (int yy = y; yy < y2; yy++) { segs.push_back({(char*)(dst + (size_t)yy * w + xa), (size_t)(xb - xa) * 4}); total += (uint64_t)(xb - xa) * 4; }
                }
                y = y2;
            }
        };
        for (int f = 0; f < n_frames; f++) {
            const HostSkip& k = hs[(size_t)f];
            int32_t* dst = host_pixels + (size_t)f * npix;
            add_rows(dst, k.yb0, k.yb1, 0, w); add_rows(dst, k.yc0, k.yc1, 0, w);
            add_rows(dst, k.yr0, k.yr1, 0, k.rx0); add_rows(dst, k.yr0, k.yr1, k.rx1 + 1, w);
        }
        ctx->last_fill_bytes = total;
        if (!segs.empty()) { ctx->fill_pool.run(std::move(segs)); fill_started = true; }
    };
    auto device_job = [&](int g) -> int {
        DeviceState& d = ctx->devs[(size_t)g];
        CU_TRY(ctx, cudaSetDevice(d.dev));
        CU_TRY(ctx, cudaEventRecord(d.ev0, d.stream));
        for (int f0 = 0; f0 < n_frames; f0 += INLINE_CAMS) {
            const int nf = n_frames - f0 < INLINE_CAMS ? n_frames - f0 : INLINE_CAMS;
            FrameParams fp = make_params(ctx, w, h, depth, spp, seed, nf, G > 1 ? g : ctx->rank, world, dev_out + (size_t)f0 * npix, (long long)npix);
            for (int i = 0; i < nf; i++) { fp.cam_inline[i] = to_cam(cams[f0 + i]); fp.hskip[i] = hs[(size_t)(f0 + i)]; }
            int rc2 = launch_render(ctx, d, fp, d.stream); if (rc2) return rc2;
            if (g == 0 && f0 == 0) start_fill();                        // after the first launch: the GPU and the link start first
        }
        CU_TRY(ctx, cudaEventRecord(d.ev1, d.stream));
        return RT_OK;
    };
    int rc = RT_OK;
    for (int g = 1; g < G; g++) ctx->workers.post(g - 1, [&device_job, g]() { return device_job(g); });
    rc = device_job(0);
    for (int g = 1; g < G; g++) { int rc2 = ctx->workers.wait(g - 1); if (!rc) rc = rc2; }
    if (rc) return rc;
    ctx->last_enqueue_ns = since_enter_ns();
    if (fill_started) {
        const auto t0 = std::chrono::steady_clock::now();
        ctx->fill_pool.wait();
        fill_started = false;
        ctx->last_fill_wait_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    }
    float kernel_ms = 0.0f;
    for (int g = 0; g < G; g++) {
        DeviceState& d = ctx->devs[(size_t)g];
        CU_TRY(ctx, cudaSetDevice(d.dev));
        CU_TRY(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0.0f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        if (ms > kernel_ms) kernel_ms = ms;
    }
    // bytes that crossed PCIe: every pixel of this context's tiles that the host did not fill
    uint64_t sent = 0;
    for (int f = 0; f < n_frames; f++) {
        const HostSkip& k = hs[(size_t)f];
        for (int y = 0; y < h; y++) {
            if (G == 1 && world > 1 && (y / tile_rows) % world != ctx->rank) continue;
            if ((y >= k.yb0 && y < k.yb1) || (y >= k.yc0 && y < k.yc1)) continue;
            sent += (uint64_t)((y >= k.yr0 && y < k.yr1) ? (k.rx1 - k.rx0 + 1) : w) * 4;
        }
    }
    ctx->last_d2h_bytes = sent;
    ctx->last_total_ns = since_enter_ns();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->kernel_ms = kernel_ms; stats->d2h_ms = kernel_ms;        // the return IS the kernel's stores
    }
    return RT_OK;
}

// Renders n_frames frames and, if host_pixels != NULL, returns them to the host.
//  * headless: ONE launch per device covers (up to 16) frames; with several devices every device stores its row tiles straight
//    into device 0's framebuffer ring (peer stores over NVLink: the gather is fused into the kernel).
//  * with a host buffer: every frame is cut into BANDS of row tiles; band s+1 is rendered while band s crosses PCIe on a
//    separate copy stream, so the device->host copy (33 MB per 4K frame, ~0.6 ms) hides the kernel instead of following it.
//    Bands are multiples of `world` tiles so that every device owns the same share of each band.
//    With several devices the frame is NOT gathered on device 0 first: each device renders into a local framebuffer and sends
//    its own tiles to the host over ITS OWN PCIe link (one strided 2-D copy per band), so host bandwidth scales with the device
//    count. rt_set_option(RT_OPT_HOST_VIA_GPU0, 1) restores gather-on-GPU-0-then-copy. A single-device context that is one rank of
//    a partition (rt_set_partition, one process per GPU) returns ITS OWN tiles only — the ranks of a torchrun job all pass the same
//    shared, page-locked host frame and fill it together, each over its own PCIe link.
//    Pixels the frame gates prove black are not copied at all (RT_OPT_SPARSE_D2H, see plan_rows above).
static int render_frames(rt_context* ctx, const rt_camera* cams, int n_frames, int w, int h, int depth, int spp, uint32_t seed,
                         int32_t* host_pixels, rt_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    auto since_enter_ns = [&]() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_enter).count(); };
    int rc = check_frame_args(ctx, cams, w, h, depth, spp);
    if (rc) return rc;
    if (n_frames < 1) return fail(ctx, RT_ERR_INVALID, "n_frames");
    const size_t npix = (size_t)w * h;
    const int G = (int)ctx->devs.size();
    // zero copy: 2 = always, 1 = frames up to 40 MB. Measured on B200 (profiles/r02/e2e_ab.json, in-process A/B, ms per frame, copy engine
    // vs the kernel's own PCIe stores): 1280x720 0.149 / 0.105, 1920x1080 0.21 / 0.165, 2560x1440 0.30 / 0.25, 3840x2160 0.517 / 0.508,
    // 7680x4320 1.88 / 1.89 — the stores reach ~47 GB/s against the copy engine's 56, but need no copy hand-off and a tenth of the
    // API calls; beyond 4K the higher DMA rate catches up.
    if (host_pixels && !ctx->host_via_gpu0 && (ctx->zero_copy == 2 || (ctx->zero_copy == 1 && npix * 4 <= ((size_t)40 << 20)))) {
        // page-locked, device-addressable host memory? (first and last byte in one registered / cudaHostAlloc'd range)
        cudaPointerAttributes a0, a1;
        const char* last = (const char*)host_pixels + npix * (size_t)n_frames * 4 - 1;
        if (cudaPointerGetAttributes(&a0, host_pixels) == cudaSuccess && cudaPointerGetAttributes(&a1, last) == cudaSuccess &&
            a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost && a0.devicePointer && a1.devicePointer &&
            (const char*)a1.devicePointer - (const char*)a0.devicePointer == last - (const char*)host_pixels)
            return render_frames_zero_copy(ctx, cams, n_frames, w, h, depth, spp, seed, host_pixels, (uint32_t*)a0.devicePointer, stats);
        cudaGetLastError();       // pageable memory: cudaPointerGetAttributes may leave an error behind on older drivers
    }
    const bool direct = host_pixels != nullptr && G > 1 && !ctx->host_via_gpu0;     // per-device D2H of own tiles
    const bool tiles_mode = direct || (host_pixels != nullptr && G == 1 && ctx->world > 1);   // copies go tile by tile (strided)
    rc = ensure_fb(ctx, npix * (size_t)n_frames); if (rc) return rc;
    if (direct) for (int g = 1; g < G; g++) { rc = ensure_fb(ctx, npix * (size_t)n_frames, g); if (rc) return rc; }
    DeviceState& d0 = ctx->devs[0];
    const int world = G > 1 ? G : ctx->world;
    const int tile_rows = ctx->tile_rows;
    const int tiles_total = (h + tile_rows - 1) / tile_rows;
    const bool pipelined = host_pixels != nullptr;
    const long long tile_pix = (long long)tile_rows * w;
    // ---- what need not be copied (per frame), and the host-side zero fill that replaces it ----
    std::vector<RowPlan> plans;
    std::atomic<uint64_t> d2h_bytes{0};
    bool fill_started = false;
    if (pipelined && ctx->sparse_d2h && ctx->path == PATH_TINY && ctx->primary_gate && spp == 1 &&
        w <= RT_FASTDIV_MAX && h <= RT_FASTDIV_MAX && !ctx->compaction) {
        plans.resize((size_t)n_frames);
        bool any = false;
        for (int f = 0; f < n_frames; f++) {
            plan_rows(gates_for(ctx, to_cam(cams[f]), w, h), w, h, &plans[(size_t)f]);
            any = any || plans[(size_t)f].sparse;
        }
        if (!any) plans.clear();
    }
    // tile kind in tiles mode: black only if every row of the tile is
    auto tile_black = [&](const RowPlan& rp, long long tile) -> bool {
        const long long ya = tile * tile_rows; long long yb = ya + tile_rows; if (yb > h) yb = h;
        for (long long y = ya; y < yb; y++) if (rp.kind[(size_t)y] != ROW_BLACK) return false;
        return true;
    };
    // ---- band layout, per frame ----
    // Bands are ranges of tile GROUPS (group q = tiles q*world .. q*world + world-1, one tile per rank), so that every device owns
    // the same share of each band. Groups that are entirely proven black at either end of the frame are not rendered at all when
    // the frame goes to the host (nobody reads those rows of the device framebuffer), and the number of bands follows the bytes that
    // really cross PCIe (~4 MB per band and link).
    constexpr int MAX_BANDS = 16;
    struct Bands { int g_lo = 0, g_hi = 0, n = 0; int bound[MAX_BANDS + 1] = {0}; };      // band i = groups [bound[i], bound[i + 1])
    const int groups_total = (tiles_total + world - 1) / world;
    std::vector<Bands> bands((size_t)n_frames);
    for (int f = 0; f < n_frames; f++) {
        Bands& b = bands[(size_t)f];
        b.g_lo = 0; b.g_hi = groups_total;
        size_t copy_rows_n = (size_t)h;
        if (!plans.empty() && plans[(size_t)f].sparse) {
            const RowPlan& rp = plans[(size_t)f];
            auto group_black = [&](int q) { for (int r = 0; r < world; r++) { const long long t = (long long)q * world + r; if (t < tiles_total && !tile_black(rp, t)) return false; } return true; };
            while (b.g_lo < b.g_hi && group_black(b.g_lo)) b.g_lo++;
            while (b.g_hi > b.g_lo && group_black(b.g_hi - 1)) b.g_hi--;
            copy_rows_n = 0;
            for (int y = 0; y < h; y++) if (rp.kind[(size_t)y] != ROW_BLACK) copy_rows_n++;
        }
        if (!pipelined) { b.n = 1; b.bound[0] = 0; b.bound[1] = groups_total; continue; }
        if (b.g_hi <= b.g_lo) { b.n = 0; continue; }
        const char* be = getenv("RTB200_BANDS");
        const size_t link_bytes = copy_rows_n * (size_t)w * 4 / (size_t)(tiles_mode ? world : 1);      // bytes one PCIe link carries
        int nb = be ? atoi(be) : (int)((link_bytes + (4u << 20) - 1) / (4u << 20));                   // ~4 MB per band and link
        if (nb > MAX_BANDS - 1) nb = MAX_BANDS - 1;
        if (nb < 1) nb = 1;
        const int range = b.g_hi - b.g_lo;
        const int per = (range + nb - 1) / nb;
        // a short head band (a quarter of the others): its kernel is done within microseconds, so the first copy starts almost at
        // once and the link — the bottleneck of the whole call — is busy from then on (measured: profiles/r02/e2e_probe*.jsonl)
        int head = (nb > 1 && !getenv("RTB200_NO_HEAD_BAND")) ? per / 4 : 0;
        b.n = 0; b.bound[0] = b.g_lo;
        if (head > 0) b.bound[++b.n] = b.g_lo + head;
        const int rest = range - head, per2 = (rest + nb - 1) / nb;
        for (int q = b.g_lo + head; q < b.g_hi; q += per2) { const int e = q + per2 < b.g_hi ? q + per2 : b.g_hi; b.bound[++b.n] = e; }
    }
    const int n_groups = (n_frames + INLINE_CAMS - 1) / INLINE_CAMS;      // headless: one launch per <= INLINE_CAMS frames
    const int n_segments = pipelined ? n_frames * MAX_BANDS : n_groups;
    if (pipelined) {
        for (int g = 0; g < G; g++) {
            DeviceState& d = ctx->devs[(size_t)g];
            CU_TRY(ctx, cudaSetDevice(d.dev));
            while ((int)d.band_events.size() < n_segments) {
                cudaEvent_t ev; CU_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                d.band_events.push_back(ev);
            }
        }
    }
    // The host-side zero fill is prepared and handed to the pool AFTER the first band is enqueued (the GPU and the link start first).
    auto start_fill = [&]() {
        ctx->last_fill_bytes = 0;
        if (plans.empty() || ctx->host_precleared) return;
        std::vector<FillPool::Seg> segs;
        auto add = [&](int32_t* p, size_t npx) {
            const size_t piece = (size_t)1 << 18;                                       // 1 MB pieces: load balance across the pool
            for (size_t o = 0; o < npx; o += piece) segs.push_back({(char*)(p + o), (npx - o < piece ? npx - o : piece) * 4});
        };
        for (int f = 0; f < n_frames; f++) {
            const RowPlan& rp = plans[(size_t)f];
            if (!rp.sparse) continue;
            int32_t* dst = host_pixels + (size_t)f * npix;
            if (tiles_mode) {
                for (long long t = 0; t < tiles_total; t++) {
                    if (G == 1 && (int)(t % world) != ctx->rank) continue;               // partitioned: own tiles only
                    if (!tile_black(rp, t)) continue;
                    const long long p0 = t * tile_pix; long long p1 = p0 + tile_pix; if (p1 > (long long)npix) p1 = (long long)npix;
                    add(dst + p0, (size_t)(p1 - p0));
                }
            } else {
                for (int y = 0; y < h;) {
                    const uint8_t k = rp.kind[(size_t)y];
                    int y2 = y + 1;
                    while (y2 < h && rp.kind[(size_t)y2] == k) y2++;
                    if (k == ROW_BLACK) add(dst + (size_t)y * w, (size_t)(y2 - y) * w);
                    else if (k == ROW_RECT)
                        for (int yy = y; yy < y2; yy++) {
                            if (rp.rx0 > 0) segs.push_back({(char*)(dst + (size_t)yy * w), (size_t)rp.rx0 * 4});
                            if (rp.rx1 < w - 1) segs.push_back({(char*)(dst + (size_t)yy * w + rp.rx1 + 1), (size_t)(w - 1 - rp.rx1) * 4});
                        }
                    y = y2;
                }
            }
        }
        uint64_t fb = 0; for (const auto& sg : segs) fb += sg.bytes;
        ctx->last_fill_bytes = fb;
        if (!segs.empty()) { ctx->fill_pool.run(std::move(segs)); fill_started = true; }
    };
    ctx->last_fill_wait_ns = 0;
    struct FillGuard { FillPool* p; const bool* started; ~FillGuard() { if (p && *started) p->wait(); } } fill_guard{&ctx->fill_pool, &fill_started};   // every return path
    std::vector<char> first_copy_dev((size_t)G, 1);
    // Device g's D2H copies for rows/tiles of one band of one frame, on its copy stream (which already waits for the band's event).
    auto copy_tiles = [&](DeviceState& d, int rank, int frame, int k_begin, int k_count) -> int {
        // tiles k*world + rank for k in [k_begin, k_begin + k_count): blocks of tile_pix pixels every world*tile_pix pixels -> one
        // strided 2-D copy per run of tiles that need copying (+ a 1-D copy if the frame's last tile is short)
        const RowPlan* rp = plans.empty() ? nullptr : &plans[(size_t)frame];
        uint32_t* src = d.fb + (size_t)frame * npix; int32_t* dst = host_pixels + (size_t)frame * npix;
        for (int k = k_begin; k < k_begin + k_count;) {
            const long long t = (long long)k * world + rank;
            if (rp && rp->sparse && tile_black(*rp, t)) { k++; continue; }
            int k2 = k + 1;
            while (k2 < k_begin + k_count && !(rp && rp->sparse && tile_black(*rp, (long long)k2 * world + rank))) k2++;
            const long long first_px = t * tile_pix;
            const long long last_tile = (long long)(k2 - 1) * world + rank;
            const bool last_short = (last_tile + 1) * tile_pix > (long long)npix;
            const int full = (k2 - k) - (last_short ? 1 : 0);
            if (full > 0) {
                CU_TRY(ctx, cudaMemcpy2DAsync(dst + first_px, (size_t)world * tile_pix * 4, src + first_px, (size_t)world * tile_pix * 4,
                                              (size_t)tile_pix * 4, (size_t)full, cudaMemcpyDeviceToHost, d.copy_stream));
                d2h_bytes += (uint64_t)full * (uint64_t)tile_pix * 4;
            }
            if (last_short) {
                const long long p0 = last_tile * tile_pix;
                CU_TRY(ctx, cudaMemcpyAsync(dst + p0, src + p0, (size_t)((long long)npix - p0) * 4, cudaMemcpyDeviceToHost, d.copy_stream));
                d2h_bytes += (uint64_t)((long long)npix - p0) * 4;
            }
            k = k2;
        }
        return RT_OK;
    };
    auto copy_rows = [&](DeviceState& d, int frame, int ya, int yb) -> int {       // contiguous rows [ya, yb) of a complete frame on d
        const RowPlan* rp = plans.empty() ? nullptr : &plans[(size_t)frame];
        uint32_t* src = d.fb + (size_t)frame * npix; int32_t* dst = host_pixels + (size_t)frame * npix;
        for (int y = ya; y < yb;) {
            const uint8_t k = (rp && rp->sparse) ? rp->kind[(size_t)y] : (uint8_t)ROW_COPY;
            int y2 = y + 1;
            while (y2 < yb && ((rp && rp->sparse) ? rp->kind[(size_t)y2] : (uint8_t)ROW_COPY) == k) y2++;
            if (k == ROW_COPY) {
                CU_TRY(ctx, cudaMemcpyAsync(dst + (size_t)y * w, src + (size_t)y * w, (size_t)(y2 - y) * w * 4, cudaMemcpyDeviceToHost, d.copy_stream));
                d2h_bytes += (uint64_t)(y2 - y) * (uint64_t)w * 4;
            } else if (k == ROW_RECT) {
                const size_t wb = (size_t)(rp->rx1 - rp->rx0 + 1) * 4;
                CU_TRY(ctx, cudaMemcpy2DAsync(dst + (size_t)y * w + rp->rx0, (size_t)w * 4, src + (size_t)y * w + rp->rx0, (size_t)w * 4, wb,
                                              (size_t)(y2 - y), cudaMemcpyDeviceToHost, d.copy_stream));
                d2h_bytes += (uint64_t)(y2 - y) * wb;
            }
            y = y2;
        }
        return RT_OK;
    };
    uint64_t gather_epoch0 = ctx->gather_epoch;
    if (G > 1 && !pipelined) {
        // the context's own gather area on device 0 (the other devices reach it through peer access, like the framebuffer)
        const uint64_t need = gather_area_bytes(w, h);
        if (ctx->gather_owned && ctx->gather_bytes < need) {
            CU_TRY(ctx, cudaSetDevice(d0.dev)); CU_TRY(ctx, cudaFree(ctx->gather_area));
            ctx->gather_area = nullptr; ctx->gather_bytes = 0; ctx->gather_owned = false;
        }
        if (!ctx->gather_area && (w % 128) == 0 && ctx->path == PATH_TINY) {
            CU_TRY(ctx, cudaSetDevice(d0.dev));
            if (cudaMalloc(&ctx->gather_area, (size_t)need) == cudaSuccess) {
                CU_TRY(ctx, cudaMemset(ctx->gather_area, 0, GATHER_CTL_BYTES));
                ctx->gather_bytes = need; ctx->gather_owned = true; ctx->gather_epoch = 0; gather_epoch0 = 0;
                free_gather_local(ctx);                               // epochs restart with the area: fresh (zeroed) per-device blocks
            } else { cudaGetLastError(); ctx->gather_area = nullptr; }
        }
        rc = ensure_gather_local(ctx); if (rc) return rc;
        if (plan_gather(ctx, 0, w, h, spp, world, 1, 1).on) ctx->gather_epoch += (uint64_t)n_groups;
    }
    // Everything device g has to enqueue for segment s (its launch, its band event, and in tiles mode its own D2H copies).
    auto enqueue = [&](int g, int s) -> int {
        DeviceState& d = ctx->devs[(size_t)g];
        const int frame = pipelined ? s / MAX_BANDS : s * INLINE_CAMS, band = pipelined ? s % MAX_BANDS : 0;
        if (pipelined && band >= bands[(size_t)frame].n) return RT_OK;                       // this frame has fewer bands
        const int nf = pipelined ? 1 : (n_frames - frame < INLINE_CAMS ? n_frames - frame : INLINE_CAMS);
        const int rank = G > 1 ? g : ctx->rank;
        // several devices, headless: every device stores into device 0's framebuffer — through the packed gather when it applies
        // (one epoch per launch group, the same on every device)
        GatherPlan gpl;
        if (G > 1 && !pipelined) gpl = plan_gather(ctx, g, w, h, spp, world, 1, gather_epoch0 + (uint64_t)s + 1);
        FrameParams fp = make_params(ctx, w, h, depth, spp, seed, nf, rank, world, (direct ? d.fb : d0.fb) + (size_t)frame * npix, (long long)npix,
                                     gpl.on ? &gpl.pt : nullptr);
        if (gpl.on) fp.gather = gpl.gp;
        fp.skip_black_store = (G > 1 && !direct) ? (ctx->shared_target == 2 ? 2 : 1) : 0;   // every device stores into device 0's framebuffer
        if (pipelined) {
            fp.cam_inline[0] = to_cam(cams[frame]);
            const Bands& b = bands[(size_t)frame];
            const int k0 = b.bound[band], k1 = b.bound[band + 1];                             // this rank's tiles of the band
            const int mine = fp.tiles_mine;
            fp.k_begin = k0 < mine ? k0 : mine;
            fp.tiles_mine = (k1 < mine ? k1 : mine) - fp.k_begin;
        } else {
            for (int i = 0; i < nf; i++) fp.cam_inline[i] = to_cam(cams[frame + i]);
        }
        int rc2 = launch_render(ctx, d, fp, d.stream); if (rc2) return rc2;
        if (pipelined) CU_TRY(ctx, cudaEventRecord(d.band_events[(size_t)s], d.stream));
        if (tiles_mode && fp.tiles_mine > 0) {
            CU_TRY(ctx, cudaStreamWaitEvent(d.copy_stream, d.band_events[(size_t)s], 0));
            if (first_copy_dev[(size_t)g]) { CU_TRY(ctx, cudaEventRecord(d.evc0, d.copy_stream)); first_copy_dev[(size_t)g] = 0; }
            rc2 = copy_tiles(d, rank, frame, fp.k_begin, fp.tiles_mine); if (rc2) return rc2;
        }
        return RT_OK;
    };
    if (G > 1 && (direct || !pipelined)) {
        // No cross-device dependency at enqueue time: every device's whole frame (all segments) is issued by its own thread.
        auto device_job = [&](int g) -> int {
            DeviceState& d = ctx->devs[(size_t)g];
            CU_TRY(ctx, cudaSetDevice(d.dev));
            CU_TRY(ctx, cudaEventRecord(d.ev0, d.stream));
            bool fill_pending = g == 0;
            for (int s = 0; s < n_segments; s++) {
                int rc2 = enqueue(g, s); if (rc2) return rc2;
                if (fill_pending && (!pipelined || s % MAX_BANDS < bands[(size_t)(s / MAX_BANDS)].n)) { start_fill(); fill_pending = false; }
            }
            if (fill_pending) start_fill();
            CU_TRY(ctx, cudaEventRecord(d.ev1, d.stream));
            return RT_OK;
        };
        for (int g = 1; g < G; g++) ctx->workers.post(g - 1, [&device_job, g]() { return device_job(g); });
        rc = device_job(0);
        for (int g = 1; g < G; g++) { int rc2 = ctx->workers.wait(g - 1); if (!rc) rc = rc2; }
        if (rc) return rc;
    } else {
        for (int g = 0; g < G; g++) {
            DeviceState& d = ctx->devs[(size_t)g];
            CU_TRY(ctx, cudaSetDevice(d.dev));
            CU_TRY(ctx, cudaEventRecord(d.ev0, d.stream));
        }
        bool fill_pending_serial = true;
        // Frame complete on device 0 (one device, or gather-on-GPU-0): per frame, the first band's launch and copy are enqueued first
        // (the link starts as early as possible), then the launches of ALL other bands (the kernels run ahead at full speed, no band
        // waits for the host), then their copies — each on the copy stream behind its band's event.
        auto copy_band = [&](int s) -> int {
            const int frame = s / MAX_BANDS, band = s % MAX_BANDS;
            CU_TRY(ctx, cudaSetDevice(d0.dev));
            for (int g = 0; g < G; g++) CU_TRY(ctx, cudaStreamWaitEvent(d0.copy_stream, ctx->devs[(size_t)g].band_events[(size_t)s], 0));
            if (first_copy_dev[0]) { CU_TRY(ctx, cudaEventRecord(d0.evc0, d0.copy_stream)); first_copy_dev[0] = 0; }
            const Bands& b = bands[(size_t)frame];
            const int q0 = b.bound[band], q1 = b.bound[band + 1];
            const long long ya = (long long)q0 * world * tile_rows; long long yb = (long long)q1 * world * tile_rows;
            if (yb > h) yb = h;
            if (yb > ya) return copy_rows(d0, frame, (int)ya, (int)yb);
            return RT_OK;
        };
        auto launch_band = [&](int s) -> int {
            for (int g = 0; g < G; g++) {
                CU_TRY(ctx, cudaSetDevice(ctx->devs[(size_t)g].dev));
                int rc2 = enqueue(g, s); if (rc2) return rc2;
            }
            return RT_OK;
        };
        if (pipelined && !tiles_mode) {
            for (int f = 0; f < n_frames; f++) {
                const int nb = bands[(size_t)f].n;
                if (nb <= 0) continue;
                const int s0 = f * MAX_BANDS;
                rc = launch_band(s0); if (rc) return rc;
                rc = copy_band(s0); if (rc) return rc;
                if (fill_pending_serial) { start_fill(); fill_pending_serial = false; }
                for (int band = 1; band < nb; band++) { rc = launch_band(s0 + band); if (rc) return rc; }
                for (int band = 1; band < nb; band++) { rc = copy_band(s0 + band); if (rc) return rc; }
            }
        } else {
            for (int s = 0; s < n_segments; s++) {
                const int frame = pipelined ? s / MAX_BANDS : s * INLINE_CAMS, band = pipelined ? s % MAX_BANDS : 0;
                if (pipelined && band >= bands[(size_t)frame].n) continue;
                rc = launch_band(s); if (rc) return rc;
                if (fill_pending_serial) { start_fill(); fill_pending_serial = false; }
            }
        }
        if (fill_pending_serial) start_fill();
        for (int g = 0; g < G; g++) {
            DeviceState& d = ctx->devs[(size_t)g];
            CU_TRY(ctx, cudaSetDevice(d.dev));
            CU_TRY(ctx, cudaEventRecord(d.ev1, d.stream));
        }
    }
    ctx->last_enqueue_ns = since_enter_ns();
    if (fill_started) {      // everything is enqueued: help with the zero fill instead of idling in the stream synchronisation below
        const auto t0 = std::chrono::steady_clock::now();
        ctx->fill_pool.wait();
        fill_started = false;
        ctx->last_fill_wait_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    }
    float kernel_ms = 0.0f, d2h_ms = 0.0f;
    for (int g = 0; g < G; g++) {
        DeviceState& d = ctx->devs[(size_t)g];
        CU_TRY(ctx, cudaSetDevice(d.dev));
        CU_TRY(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0.0f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        if (ms > kernel_ms) kernel_ms = ms;
    }
    if (pipelined) {
        for (int g = 0; g < G; g++) {
            DeviceState& d = ctx->devs[(size_t)g];
            if (first_copy_dev[(size_t)g]) continue;             // this device issued no copy
            CU_TRY(ctx, cudaSetDevice(d.dev));
            CU_TRY(ctx, cudaEventRecord(d.evc1, d.copy_stream));
            CU_TRY(ctx, cudaStreamSynchronize(d.copy_stream));
            float ms = 0.0f;
            CU_TRY(ctx, cudaEventElapsedTime(&ms, d.evc0, d.evc1));
            if (ms > d2h_ms) d2h_ms = ms;
        }
    }
    if (pipelined) ctx->last_d2h_bytes = d2h_bytes.load();
    ctx->last_total_ns = since_enter_ns();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->kernel_ms = kernel_ms; stats->gather_ms = 0.0f; stats->d2h_ms = d2h_ms;   // d2h overlaps the kernel when banded
    }
    return RT_OK;
}

int rt_render(rt_context* ctx, const rt_camera* cam, int w, int h, int depth, int spp, uint32_t seed, int32_t* host_pixels, rt_stats* stats) {
    return render_frames(ctx, cam, 1, w, h, depth, spp, seed, host_pixels, stats);
}

int rt_render_batch(rt_context* ctx, const rt_camera* cams, int n_frames, int w, int h, int depth, int spp, uint32_t seed,
                    int32_t* host_pixels, rt_stats* stats) {
    return render_frames(ctx, cams, n_frames, w, h, depth, spp, seed, host_pixels, stats);
}

int rt_render_debug(rt_context* ctx, const rt_camera* cam, int w, int h, int depth, int spp, uint32_t seed, int32_t* host_pixels,
                    uint32_t* host_hash, int32_t* host_aov_id, float* host_aov_t, uint64_t* counters, rt_stats* stats) {
    int rc = check_frame_args(ctx, cam, w, h, depth, spp);
    if (rc) return rc;
    const size_t npix = (size_t)w * h;
    rc = ensure_fb(ctx, npix); if (rc) return rc;
    DeviceState& d = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    DebugOut dout; memset(&dout, 0, sizeof(dout));
    DevMem<unsigned long long> m_cnt; DevMem<uint32_t> m_hash; DevMem<int32_t> m_id; DevMem<float> m_t;
    CU_TRY(ctx, m_cnt.alloc(N_DEBUG_COUNTERS));
    unsigned long long* dcnt = m_cnt.p;
    CU_TRY(ctx, cudaMemsetAsync(dcnt, 0, N_DEBUG_COUNTERS * sizeof(unsigned long long), d.stream));
    dout.counters = dcnt;
    if (host_hash) { CU_TRY(ctx, m_hash.alloc(npix)); dout.hash = m_hash.p; }
    if (host_aov_id) { CU_TRY(ctx, m_id.alloc(npix)); dout.aov_id = m_id.p; }
    if (host_aov_t) { CU_TRY(ctx, m_t.alloc(npix)); dout.aov_t = m_t.p; }
    FrameParams fp = make_params(ctx, w, h, depth, spp, seed, 1, 0, 1, d.fb, (long long)npix);
    fp.cam_inline[0] = to_cam(*cam);
    long long grid = ((long long)npix + BLOCK - 1) / BLOCK;
    long long maxgrid = (long long)d.sm_count * 8 * 4;
    if (grid > maxgrid) grid = maxgrid;
    CU_TRY(ctx, cudaEventRecord(d.ev0, d.stream));
    switch (ctx->path) {
        case PATH_TINY: {
            const bool fastdiv_ok = w <= RT_FASTDIV_MAX && h <= RT_FASTDIV_MAX;
            if (ctx->debug_shipped && spp == 1 && fastdiv_ok) {
                // the PRODUCTION instantiation (exact counts, packed fp32, fast division, frame gates) with the events-only policy,
                // on the production grid
                FrameParams gp = fp;
                fill_gates(ctx, gp);
                const TinySceneData& t = ctx->tiny_data;
                const dim3 pgrid((unsigned)gp.chunks_per_tile, (unsigned)(gp.tiles_mine < 65535 ? gp.tiles_mine : 65535), 1u);
                tiny_debug_prod_kernel(t.ns, t.nl, t.np)<<<pgrid, BLOCK, 0, d.stream>>>(t, gp, dout);
            } else {
                k_debug_tiny<<<(unsigned)grid, BLOCK, 0, d.stream>>>(ctx->tiny_data, fp, dout);
            }
            break;
        }
        case PATH_STAGED: k_debug_staged<<<(unsigned)grid, BLOCK, sizeof(f4) * (size_t)ctx->gdata_host.ns, d.stream>>>(global_data(ctx, d), fp, dout); break;
        case PATH_GLOBAL: k_debug_global<<<(unsigned)grid, BLOCK, 0, d.stream>>>(global_data(ctx, d), fp, dout); break;
        default: {
            // the same per-frame state as the production launch (launch_render), incl. the primary bins when they are switched on: the
            // chain hashes / counters of this kernel then certify the binned primary fold (node_visits_primary counts what still traverses)
            bool use_bins = false;
            int prc = prepare_lbvh_frame(ctx, d, fp.cam_inline[0], w, h, spp, d.stream, &use_bins);
            if (prc) return prc;
            if (use_bins) k_debug_lbvh_bins<<<(unsigned)grid, BLOCK, 0, d.stream>>>(lbvh_bins_data(ctx, d), fp, dout);
            else k_debug_lbvh<<<(unsigned)grid, BLOCK, 0, d.stream>>>(lbvh_data(ctx, d, true), fp, dout);
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaEventRecord(d.bvh_done, d.stream));
            d.bvh_done_valid = true; d.bvh_last_stream = d.stream;
            break;
        }
    }
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CU_TRY(ctx, cudaEventRecord(d.ev1, d.stream));
    CU_TRY(ctx, cudaStreamSynchronize(d.stream));
    float ms = 0; CU_TRY(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
    if (host_pixels) CU_TRY(ctx, cudaMemcpy(host_pixels, d.fb, npix * 4, cudaMemcpyDeviceToHost));
    if (host_hash) CU_TRY(ctx, cudaMemcpy(host_hash, dout.hash, npix * 4, cudaMemcpyDeviceToHost));
    if (host_aov_id) CU_TRY(ctx, cudaMemcpy(host_aov_id, dout.aov_id, npix * 4, cudaMemcpyDeviceToHost));
    if (host_aov_t) CU_TRY(ctx, cudaMemcpy(host_aov_t, dout.aov_t, npix * 4, cudaMemcpyDeviceToHost));
    unsigned long long hc[N_DEBUG_COUNTERS];
    CU_TRY(ctx, cudaMemcpy(hc, dcnt, sizeof(hc), cudaMemcpyDeviceToHost));
    if (counters) for (int i = 0; i < N_DEBUG_COUNTERS; i++) counters[i] = hc[i];
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->primary = hc[0]; stats->shadow = hc[1]; stats->secondary = hc[2]; stats->kernel_ms = ms;
    }
    return RT_OK;
}

int rt_query_spheres(rt_context* ctx, const float* rays6, int n_rays, int kind, int accel, int32_t* out_id, float* out_t) {
    if (!ctx) return RT_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, RT_ERR_NO_SCENE, "rt_set_scene has not been called");
    if (!rays6 || n_rays < 0 || kind < 0 || kind > 2 || !out_id || !out_t) return fail(ctx, RT_ERR_INVALID, "bad query args");
    if (accel != RT_ACCEL_BRUTE && accel != RT_ACCEL_LBVH) return fail(ctx, RT_ERR_INVALID, "accel must be RT_ACCEL_BRUTE or RT_ACCEL_LBVH");
    if (accel == RT_ACCEL_LBVH && !ctx->has_bvh) return fail(ctx, RT_ERR_UNSUPPORTED, "the scene was uploaded without an LBVH");
    if (n_rays == 0) return RT_OK;
    DeviceState& d = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    DevMem<float> m_r, m_t; DevMem<int32_t> m_i;
    CU_TRY(ctx, m_r.alloc((size_t)n_rays * 6));
    CU_TRY(ctx, m_i.alloc((size_t)n_rays));
    CU_TRY(ctx, m_t.alloc((size_t)n_rays));
    float* dr = m_r.p; int32_t* di = m_i.p; float* dt = m_t.p;
    CU_TRY(ctx, cudaMemcpyAsync(dr, rays6, (size_t)n_rays * 24, cudaMemcpyHostToDevice, d.stream));   // stream-ordered before the kernel
    int grid = (n_rays + BLOCK - 1) / BLOCK; if (grid > d.sm_count * 32) grid = d.sm_count * 32;
    if (accel == RT_ACCEL_LBVH) k_query_lbvh<<<grid, BLOCK, 0, d.stream>>>(lbvh_data(ctx, d, false), dr, n_rays, kind, di, dt);
    else k_query_brute<<<grid, BLOCK, 0, d.stream>>>(global_data(ctx, d), dr, n_rays, kind, di, dt);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CU_TRY(ctx, cudaStreamSynchronize(d.stream));
    CU_TRY(ctx, cudaMemcpy(out_id, di, (size_t)n_rays * 4, cudaMemcpyDeviceToHost));
    CU_TRY(ctx, cudaMemcpy(out_t, dt, (size_t)n_rays * 4, cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_selftest(rt_context* ctx, int test, uint64_t* n_checked, uint64_t* n_mismatch) {
    if (!ctx) return RT_ERR_INVALID;
    if (!n_checked || !n_mismatch) return fail(ctx, RT_ERR_INVALID, "bad selftest args");
    DeviceState& d = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    DevMem<unsigned long long> m_bad;
    CU_TRY(ctx, m_bad.alloc(1));
    unsigned long long* dbad = m_bad.p;
    CU_TRY(ctx, cudaMemsetAsync(dbad, 0, sizeof(unsigned long long), d.stream));
    uint64_t checked = 0;
    if (test == RT_SELFTEST_INV_LEN || test == RT_SELFTEST_INV_LEN_RSQ_SEED) {
        // every float of the fast path's range [2^-64, 2^64) plus a binade on either side (those take the IEEE branch)
        const uint32_t first = 0x1F000000u, count = 0x41000000u;
        k_selftest_inv_len<<<d.sm_count * 16, 256, 0, d.stream>>>(first, count, test == RT_SELFTEST_INV_LEN ? 0 : 1, dbad);
        checked = count;
    } else if (test == RT_SELFTEST_INV_LEN_PAIR) {
        const uint32_t first = 0x1F000000u, count = 0x41000000u;
        k_selftest_inv_len_pair<<<d.sm_count * 16, 256, 0, d.stream>>>(first, count, dbad);
        checked = 2ull * count;
    } else if (test == RT_SELFTEST_PIXEL_DIV) {
        k_selftest_pixel_div<<<d.sm_count * 8, 256, 0, d.stream>>>(RT_FASTDIV_MAX, dbad);
        checked = (uint64_t)RT_FASTDIV_MAX * (RT_FASTDIV_MAX + 1) / 2;
    } else {
        return fail(ctx, RT_ERR_INVALID, "unknown selftest");
    }
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CU_TRY(ctx, cudaStreamSynchronize(d.stream));
    unsigned long long bad = 0;
    CU_TRY(ctx, cudaMemcpy(&bad, dbad, sizeof(bad), cudaMemcpyDeviceToHost));
    *n_checked = checked; *n_mismatch = bad;
    return RT_OK;
}

int rt_ray_log(rt_context* ctx, const rt_camera* cam, int w, int h, int depth, const uint32_t* pixels, int n_pixels,
               rt_ray_record* out, int max_records, int* n_records) {
    int rc = check_frame_args(ctx, cam, w, h, depth, 1);
    if (rc) return rc;
    if (n_pixels < 0 || max_records < 0 || !n_records || (n_pixels > 0 && !pixels) || (max_records > 0 && !out))
        return fail(ctx, RT_ERR_INVALID, "bad ray-log args");
    const uint64_t npix = (uint64_t)w * (uint64_t)h;
    for (int i = 0; i < n_pixels; i++)
        if (pixels[i] >= npix) return fail(ctx, RT_ERR_INVALID, "ray-log pixel index outside the frame");
    *n_records = 0;
    if (n_pixels == 0) return RT_OK;
    // a chain has at most depth + 2 rays and depth + 1 shaded hits, each casting one shadow ray per light
    const int slots = (depth + 2) + (depth + 1) * ctx->gdata_host.nl;
    if ((uint64_t)slots * (uint64_t)n_pixels * sizeof(RayRec) > (1ull << 31))
        return fail(ctx, RT_ERR_UNSUPPORTED, "ray log too large: list fewer pixels per call");
    DeviceState& d = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    DevMem<uint32_t> m_pix, m_count; DevMem<RayRec> m_recs;
    CU_TRY(ctx, m_pix.alloc((size_t)n_pixels));
    CU_TRY(ctx, m_count.alloc((size_t)n_pixels));
    CU_TRY(ctx, m_recs.alloc((size_t)n_pixels * slots));
    uint32_t* dpix = m_pix.p; uint32_t* dcount = m_count.p; RayRec* drecs = m_recs.p;
    CU_TRY(ctx, cudaMemcpyAsync(dpix, pixels, (size_t)n_pixels * 4, cudaMemcpyHostToDevice, d.stream));
    FrameParams fp = make_params(ctx, w, h, depth, 1, 0u, 1, 0, 1, nullptr, (long long)npix);
    fp.cam_inline[0] = to_cam(*cam);
    int grid = (n_pixels + BLOCK - 1) / BLOCK; if (grid > d.sm_count * 32) grid = d.sm_count * 32;
    k_ray_log<<<grid, BLOCK, 0, d.stream>>>(global_data(ctx, d), fp, dpix, n_pixels, slots, drecs, dcount);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CU_TRY(ctx, cudaStreamSynchronize(d.stream));
    std::vector<uint32_t> hcount((size_t)n_pixels);
    CU_TRY(ctx, cudaMemcpy(hcount.data(), dcount, (size_t)n_pixels * 4, cudaMemcpyDeviceToHost));
    long long total = 0, written = 0;
    for (int i = 0; i < n_pixels; i++) {
        const uint32_t c = hcount[(size_t)i];
        if (c > (uint32_t)slots) return fail(ctx, RT_ERR_CUDA, "ray log: slot bound violated");
        const long long room = (long long)max_records - written;
        const long long take = (long long)c < room ? (long long)c : (room > 0 ? room : 0);
        if (take > 0)
            CU_TRY(ctx, cudaMemcpy(out + written, drecs + (size_t)i * slots, (size_t)take * sizeof(RayRec), cudaMemcpyDeviceToHost));
        written += take; total += c;
    }
    if (total > 0x7fffffffLL) return fail(ctx, RT_ERR_UNSUPPORTED, "ray log: record count overflows int");
    *n_records = (int)total;
    return RT_OK;
}

int rt_ipc_export(rt_context* ctx, void* dev_ptr, void* handle64) {
    if (!ctx || !dev_ptr || !handle64) return RT_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    cudaIpcMemHandle_t h;
    CU_TRY(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return RT_OK;
}
int rt_ipc_open(rt_context* ctx, const void* handle64, void** out_dev_ptr) {
    if (!ctx || !handle64 || !out_dev_ptr) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(out_dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(ctx, RT_ERR_PEER, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    return RT_OK;
}
int rt_ipc_close(rt_context* ctx, void* dev_ptr) {
    if (!ctx || !dev_ptr) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return RT_OK;
}

int rt_dev_alloc(rt_context* ctx, uint64_t bytes, void** out_dev_ptr) {
    if (!ctx || !out_dev_ptr || bytes == 0) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaMalloc(out_dev_ptr, (size_t)bytes));
    return RT_OK;
}
int rt_dev_free(rt_context* ctx, void* dev_ptr) {
    if (!ctx) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaFree(dev_ptr));
    return RT_OK;
}
int rt_dev_to_host(rt_context* ctx, void* host_dst, const void* dev_src, uint64_t bytes) {
    if (!ctx || !host_dst || !dev_src) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaMemcpy(host_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return RT_OK;
}
int rt_dev_memset(rt_context* ctx, void* dev_dst, int byte_value, uint64_t bytes) {
    if (!ctx || !dev_dst) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaMemset(dev_dst, byte_value, (size_t)bytes));
    CU_TRY(ctx, cudaDeviceSynchronize());
    return RT_OK;
}
int rt_sync(rt_context* ctx) {
    if (!ctx) return RT_ERR_INVALID;
    for (auto& d : ctx->devs) {
        CU_TRY(ctx, cudaSetDevice(d.dev));
        CU_TRY(ctx, cudaDeviceSynchronize());
    }
    return RT_OK;
}
int rt_host_register(rt_context* ctx, void* host_ptr, uint64_t bytes) {
    if (!ctx || !host_ptr || bytes == 0) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return RT_OK;
}
int rt_host_unregister(rt_context* ctx, void* host_ptr) {
    if (!ctx || !host_ptr) return RT_ERR_INVALID;
    CU_TRY(ctx, cudaHostUnregister(host_ptr));
    return RT_OK;
}
uint64_t rt_launch_count(const rt_context* ctx) { return ctx ? ctx->launches.load() : 0; }

// Measured L2 -> SM read bandwidth (GB/s) for a working set of `bytes` (rounded down to 4 KB; 1 MB .. 96 MB): the denominator of the
// LBVH kernels' roofline (node and leaf fetches are L2-resident: SURVEY §8(d)). Best of 3 timed launches after a warm-up pass.
int rt_measure_l2_read(rt_context* ctx, uint64_t bytes, double* gbs) {
    if (!ctx || !gbs) return RT_ERR_INVALID;
    if (bytes < (1u << 20) || bytes > (96u << 20)) return fail(ctx, RT_ERR_INVALID, "working set must be 1 MB .. 96 MB");
    DeviceState& d = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    const size_t n_vec = (size_t)(bytes / 4096) * 256;          // uint4 elements, a multiple of 256
    DevMem<uint4> buf; DevMem<unsigned int> sink;
    CU_TRY(ctx, buf.alloc(n_vec));
    CU_TRY(ctx, sink.alloc(1));
    CU_TRY(ctx, cudaMemsetAsync(buf.p, 0, n_vec * sizeof(uint4), d.stream));
    const int grid = d.sm_count * 4, passes = 2;
    k_l2_read<<<grid, 256, 0, d.stream>>>(buf.p, n_vec, 1, sink.p);                       // warm-up: pulls the set into L2
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CU_TRY(ctx, cudaEventRecord(d.ev0, d.stream));
        k_l2_read<<<grid, 256, 0, d.stream>>>(buf.p, n_vec, passes, sink.p);
        CU_TRY(ctx, cudaEventRecord(d.ev1, d.stream));
        CU_TRY(ctx, cudaStreamSynchronize(d.stream));
        CU_TRY(ctx, cudaGetLastError());
        float ms = 0.0f; CU_TRY(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        if (ms < best) best = ms;
        ctx->launches++;
    }
    *gbs = (double)grid * passes * (double)n_vec * 16.0 / ((double)best * 1e-3) / 1e9;
    return RT_OK;
}

uint64_t rt_gather_bytes(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    return gather_area_bytes(width, height);
}
int rt_gather_attach(rt_context* ctx, void* area_dev_ptr, uint64_t bytes) {
    if (!ctx) return RT_ERR_INVALID;
    if (ctx->devs.size() != 1) return fail(ctx, RT_ERR_INVALID, "rt_gather_attach needs a single-device context (multi-device contexts own their gather area)");
    if (area_dev_ptr && bytes < GATHER_CTL_BYTES) return fail(ctx, RT_ERR_INVALID, "gather area too small");
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CU_TRY(ctx, cudaDeviceSynchronize());
    if (ctx->gather_owned && ctx->gather_area) cudaFree(ctx->gather_area);
    ctx->gather_owned = false;
    free_gather_local(ctx);
    ctx->gather_area = (unsigned char*)area_dev_ptr; ctx->gather_bytes = area_dev_ptr ? bytes : 0;
    ctx->gather_epoch = 0; ctx->gather_active = false;
    if (area_dev_ptr) return ensure_gather_local(ctx);
    return RT_OK;
}

int rt_get_info(const rt_context* ctx, int what, uint64_t* value) {
    if (!ctx || !value) return RT_ERR_INVALID;
    switch (what) {
        case RT_INFO_GATHER_ACTIVE: *value = ctx->gather_active ? 1u : 0u; return RT_OK;
        case RT_INFO_GATHER_TIMEOUTS: {
            *value = 0;
            if (!ctx->gather_area) return RT_OK;
            GatherCtl c;
            if (cudaSetDevice(ctx->devs[0].dev) != cudaSuccess || cudaMemcpy(&c, ctx->gather_area, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) {
                cudaGetLastError();
                return RT_ERR_CUDA;
            }
            *value = c.timeouts;
            return RT_OK;
        }
        case RT_INFO_GATE_HOST_NS: *value = ctx->gate_host_ns.load(); return RT_OK;
        case RT_INFO_GATE_COMPUTES: *value = ctx->gate_computes.load(); return RT_OK;
        case RT_INFO_LAST_D2H_BYTES: *value = ctx->last_d2h_bytes; return RT_OK;
        case RT_INFO_SCENE_PATH: *value = (uint64_t)ctx->path; return RT_OK;
        case RT_INFO_LAST_FILL_BYTES: *value = ctx->last_fill_bytes; return RT_OK;
        case RT_INFO_LAST_FILL_WAIT_NS: *value = ctx->last_fill_wait_ns; return RT_OK;
        case RT_INFO_LAST_ENQUEUE_NS: *value = ctx->last_enqueue_ns; return RT_OK;
        case RT_INFO_LAST_TOTAL_NS: *value = ctx->last_total_ns; return RT_OK;
        case RT_INFO_SHADOW_BINS_NS: *value = ctx->sg_build_ns; return RT_OK;
        case RT_INFO_PRIMARY_BIN_BUILDS: *value = ctx->pb_builds.load(); return RT_OK;
        case RT_INFO_PRIMARY_BINS: *value = ctx->primary_bins ? 1u : 0u; return RT_OK;
        case RT_INFO_SHADOW_BIN_PAIRS: *value = (ctx->has_shadow_grids && !ctx->devs.empty()) ? (uint64_t)ctx->devs[0].sg.n_pairs : 0u; return RT_OK;
        default: return RT_ERR_INVALID;
    }
}

// ---- zero-copy display ---------------------------------------------------------------------------------------
// The frame is rendered straight into a device buffer the DISPLAY owns (a CUDA-mapped OpenGL pixel-unpack buffer) and never
// crosses PCIe: replaces the upload of Surface.pixels from host memory (template.cs:81, :188-193).
int rt_render_mapped(rt_context* ctx, const rt_camera* cam, int w, int h, int depth, int spp, uint32_t seed, void* mapped_dev_pixels,
                     uint64_t mapped_bytes, rt_stats* stats) {
    int rc = check_frame_args(ctx, cam, w, h, depth, spp);
    if (rc) return rc;
    if (!mapped_dev_pixels) return fail(ctx, RT_ERR_INVALID, "mapped_dev_pixels is NULL");
    const uint64_t need = (uint64_t)w * (uint64_t)h * 4u;
    if (mapped_bytes < need) return fail(ctx, RT_ERR_INVALID, "mapped buffer smaller than 4*width*height bytes");
    if ((reinterpret_cast<uintptr_t>(mapped_dev_pixels) & 3u) != 0) return fail(ctx, RT_ERR_INVALID, "mapped buffer not 4-byte aligned");
    DeviceState& d0 = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d0.dev));
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, mapped_dev_pixels);
    if (e != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(ctx, RT_ERR_INVALID, "mapped_dev_pixels is not a device pointer");
    }
    if (attr.type == cudaMemoryTypeDevice && attr.device != d0.dev)
        return fail(ctx, RT_ERR_INVALID, "the mapped buffer must live on the context's first device");
    if (ctx->devs.size() == 1) {
        // one device: the render kernel's 128-bit stores go straight into the mapped buffer
        FrameParams fp = make_params(ctx, w, h, depth, spp, seed, 1, ctx->rank, ctx->world, (uint32_t*)mapped_dev_pixels, (long long)w * h);
        fp.cam_inline[0] = to_cam(*cam);
        CU_TRY(ctx, cudaEventRecord(d0.ev0, d0.stream));
        rc = launch_render(ctx, d0, fp, d0.stream); if (rc) return rc;
        CU_TRY(ctx, cudaEventRecord(d0.ev1, d0.stream));
        CU_TRY(ctx, cudaStreamSynchronize(d0.stream));
        float ms = 0.0f; CU_TRY(ctx, cudaEventElapsedTime(&ms, d0.ev0, d0.ev1));
        if (stats) { memset(stats, 0, sizeof(*stats)); stats->kernel_ms = ms; }
        return RT_OK;
    }
    // several devices: gather on device 0 as always (peer stores into the context's framebuffer; graphics-interop memory is not
    // guaranteed to be peer-mappable), then one device-to-device copy into the mapped buffer
    rt_stats st;
    rc = render_frames(ctx, cam, 1, w, h, depth, spp, seed, nullptr, &st); if (rc) return rc;
    CU_TRY(ctx, cudaSetDevice(d0.dev));
    CU_TRY(ctx, cudaEventRecord(d0.ev0, d0.stream));
    CU_TRY(ctx, cudaMemcpyAsync(mapped_dev_pixels, d0.fb, (size_t)need, cudaMemcpyDeviceToDevice, d0.stream));
    CU_TRY(ctx, cudaEventRecord(d0.ev1, d0.stream));
    CU_TRY(ctx, cudaStreamSynchronize(d0.stream));
    float ms = 0.0f; CU_TRY(ctx, cudaEventElapsedTime(&ms, d0.ev0, d0.ev1));
    if (stats) { *stats = st; stats->gather_ms += ms; }
    return RT_OK;
}

int rt_gl_register_buffer(rt_context* ctx, unsigned int gl_buffer, void** out_resource) {
    if (!ctx || !out_resource) return RT_ERR_INVALID;
    *out_resource = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    cudaGraphicsResource_t res = nullptr;
    cudaError_t e = cudaGraphicsGLRegisterBuffer(&res, gl_buffer, cudaGraphicsRegisterFlagsWriteDiscard);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, RT_ERR_CUDA, std::string("cudaGraphicsGLRegisterBuffer: ") + cudaGetErrorString(e) +
                                          " (needs a current OpenGL context on the calling thread)");
    }
    ctx->gl_resources.push_back((void*)res);
    *out_resource = (void*)res;
    return RT_OK;
}
int rt_gl_unregister_buffer(rt_context* ctx, void* resource) {
    if (!ctx || !resource) return RT_ERR_INVALID;
    for (size_t i = 0; i < ctx->gl_resources.size(); i++)
        if (ctx->gl_resources[i] == resource) {
            ctx->gl_resources.erase(ctx->gl_resources.begin() + (long)i);
            CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
            CU_TRY(ctx, cudaGraphicsUnregisterResource((cudaGraphicsResource_t)resource));
            return RT_OK;
        }
    return fail(ctx, RT_ERR_INVALID, "unknown graphics resource");
}
int rt_render_gl(rt_context* ctx, const rt_camera* cam, int w, int h, int depth, int spp, uint32_t seed, void* resource, rt_stats* stats) {
    if (!ctx) return RT_ERR_INVALID;
    if (!resource) return fail(ctx, RT_ERR_INVALID, "resource is NULL");
    bool known = false;
    for (void* r : ctx->gl_resources) known = known || r == resource;
    if (!known) return fail(ctx, RT_ERR_INVALID, "unknown graphics resource");
    DeviceState& d0 = ctx->devs[0];
    CU_TRY(ctx, cudaSetDevice(d0.dev));
    cudaGraphicsResource_t res = (cudaGraphicsResource_t)resource;
    CU_TRY(ctx, cudaGraphicsMapResources(1, &res, d0.stream));
    void* ptr = nullptr; size_t bytes = 0;
    cudaError_t e = cudaGraphicsResourceGetMappedPointer(&ptr, &bytes, res);
    int rc = e == cudaSuccess ? rt_render_mapped(ctx, cam, w, h, depth, spp, seed, ptr, (uint64_t)bytes, stats)
                              : fail(ctx, RT_ERR_CUDA, std::string("cudaGraphicsResourceGetMappedPointer: ") + cudaGetErrorString(e));
    cudaError_t e2 = cudaGraphicsUnmapResources(1, &res, d0.stream);      // always unmap: GL may not touch a mapped buffer
    if (rc == RT_OK && e2 != cudaSuccess) rc = fail(ctx, RT_ERR_CUDA, std::string("cudaGraphicsUnmapResources: ") + cudaGetErrorString(e2));
    return rc;
}

}  // extern "C"
