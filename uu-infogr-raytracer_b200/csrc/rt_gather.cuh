// rt_gather.cuh — the packed multi-GPU gather: row tiles rendered on GPUs 1..N-1 cross NVLink to GPU 0 in a compressed wire format
// and are expanded into the 0x00RRGGBB framebuffer (Surface.pixels layout, surface.cs:9-20) by a pass on GPU 0, with all
// synchronisation done by the kernels themselves through flags in GPU 0's memory (no host round trip, no library collective).
//
// Why: with 8 ranks the plain fused gather (every rank's 128-bit stores straight into GPU 0's framebuffer, rtb200.cu render_loop)
// is bound by GPU 0's NVLink ingress — 302 MB per 16-frame step at ~0.70 TB/s = 0.43 ms against 0.30 ms of rendering
// (SCALE_r01: efficiency 0.68). The alpha byte of every pixel is 0 and most floor pixels are grey (R = G = B), so the wire format is
//   * nothing                 for a quad (16 pixels = 4 lanes x 4 pixels) the frame gates prove black (as before),
//   * 1 byte per pixel        for a quad whose 16 pixels are all grey       -> G plane (16 B per quad)
//   * 3 bytes per pixel       otherwise (RGB, little-endian B G R)          -> C plane (48 B per quad, three 128-bit stores)
//   * 1 flag byte per warp    (bit q: quad q is grey)                       -> F plane (one byte per 128 pixels)
// i.e. ~1.7 B per non-black pixel on the bench frame instead of 4. Planes live in a "gather area" in GPU 0's memory (allocated by
// rank 0, mapped by the others through CUDA IPC or peer access), one slot per frame of a launch (<= 16).
//
// Synchronisation (all counters 64-bit epochs, one epoch per launch group, identical on every rank):
//   done[r][f]  written by rank r's LAST CTA of frame slot f after a system-scope fence: rank r's planes of slot f have landed.
//               GPU 0's expand CTAs spin on it (local memory) before they read a tile of rank r.
//   freed[f]    written by GPU 0's last expand CTA of slot f: the slot may be overwritten. Rank r's CTAs of the NEXT epoch spin on
//               it (over NVLink; the value is cached in rank r's own memory) before their first store.
// Every spin is bounded (2 s of %globaltimer): on a time-out the CTA counts it in GatherCtl::timeouts, raises `abort` so that
// nobody else waits, and goes on — the frame is then wrong, which rt_get_info(RT_INFO_GATHER_TIMEOUTS) reports, but nothing hangs.
#pragma once
#include <stdint.h>

namespace rtb {

constexpr int GATHER_MAX_RANKS = 8;
constexpr int GATHER_SLOTS = 16;            // frame slots = INLINE_CAMS
constexpr int GATHER_CTL_BYTES = 4096;

struct GatherCtl {                          // first GATHER_CTL_BYTES of the gather area (GPU 0's memory)
    unsigned long long done[GATHER_MAX_RANKS][GATHER_SLOTS];
    unsigned long long freed[GATHER_SLOTS];
    unsigned int expand_count[GATHER_SLOTS];    // GPU 0 only: expand CTAs finished per slot (reset by the last one)
    unsigned int timeouts;
    unsigned int abort;
};
static_assert(sizeof(GatherCtl) <= GATHER_CTL_BYTES, "GatherCtl must fit its page");

struct GatherLocal {                        // in each rank's OWN memory
    unsigned long long seen_freed[GATHER_SLOTS];   // last value of GatherCtl::freed[f] this rank has observed
    unsigned int cta_count[GATHER_SLOTS];          // this rank's render CTAs finished per slot (reset by the last one)
};

struct GatherParams {                       // travels in the kernel parameter block; area == nullptr: plain 4 B / pixel gather
    unsigned char* area;                    // the gather area as THIS rank addresses it
    GatherLocal* local;
    unsigned long long epoch;
    unsigned long long off_c, off_g, off_f; // byte offsets of the C / G / F planes of slot 0 inside the area
    unsigned long long stride_c, stride_g, stride_f;   // bytes per slot
    int grey;                               // 1: grey quads go to the G plane; 0: every non-black quad goes to the C plane
    int pad;
};

// bytes of a gather area for frames of w x h pixels (16 slots): control page + C (3 B/px) + G (1 B/px) + F (1 B / 128 px) planes
inline unsigned long long gather_area_bytes(int w, int h) {
    const unsigned long long npix = (unsigned long long)w * (unsigned long long)h;
    return GATHER_CTL_BYTES + GATHER_SLOTS * (3 * npix + npix + (npix + 127) / 128 + 64);
}

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long gather_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gather_st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gather_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spins until *p >= want (acquire, system scope). Bounded: see the header comment. `ctl` is in GPU 0's memory.
__device__ __forceinline__ bool gather_wait_ge(const unsigned long long* p, unsigned long long want, GatherCtl* ctl) {
    if (gather_ld_acquire(p) >= want) return true;
    const unsigned long long t0 = gather_timer_ns();
    for (;;) {
        __nanosleep(500);
        if (gather_ld_acquire(p) >= want) return true;
        if (*(volatile unsigned int*)&ctl->abort) return false;
        if (gather_timer_ns() - t0 > 2000000000ull) {
            atomicAdd(&ctl->timeouts, 1u);
            *(volatile unsigned int*)&ctl->abort = 1u;
            return false;
        }
    }
}

// 4 pixels 0x00RRGGBB -> 12 bytes B G R B G R ... (three little-endian words) and back
__device__ __forceinline__ void gather_pack_rgb(const uint32_t px[4], uint32_t w[3]) {
    w[0] = (px[0] & 0x00FFFFFFu) | (px[1] << 24);
    w[1] = ((px[1] >> 8) & 0x0000FFFFu) | (px[2] << 16);
    w[2] = ((px[2] >> 16) & 0x000000FFu) | (px[3] << 8);
}
__device__ __forceinline__ uint4 gather_unpack_rgb(uint32_t w0, uint32_t w1, uint32_t w2) {
    return make_uint4(w0 & 0x00FFFFFFu, (w0 >> 24) | ((w1 & 0x0000FFFFu) << 8), (w1 >> 16) | ((w2 & 0x000000FFu) << 16), w2 >> 8);
}
__device__ __forceinline__ bool gather_is_grey(uint32_t p) { return ((p ^ (p >> 8)) & 0xFFFFu) == 0u; }     // R == G == B (alpha is 0)
__device__ __forceinline__ uint4 gather_unpack_grey(uint32_t g) {
    return make_uint4((g & 0xFFu) * 0x010101u, ((g >> 8) & 0xFFu) * 0x010101u, ((g >> 16) & 0xFFu) * 0x010101u, (g >> 24) * 0x010101u);
}
// bit q of the result = all four lanes of quad q have `pred` (called by all 32 lanes)
__device__ __forceinline__ uint32_t gather_quad_all(bool pred) {
    uint32_t b = __ballot_sync(0xffffffffu, pred);
    b &= b >> 1; b &= b >> 2;                         // bit 4q = AND of bits 4q..4q+3
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) r |= ((b >> (4 * q)) & 1u) << q;
    return r;
}
#endif

}  // namespace rtb
