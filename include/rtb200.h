/* rtb200.h — C ABI of librtb200.so, the B200 (sm_100a) render backend for the per-pixel ray-cast path of
 * TobiasDeBruijn/UU-INFOGR-Raytracer, Raytracer/RayTracer.cs.
 *
 * The reference exposes no plugin / FFI interface: its trace methods are private members of `RayTracer`.
 * The seam this library fills is the pixel loop in `RayTracer.Tick()` (RayTracer.cs:898-901): the C# host keeps
 * its scene arrays (:441-465), camera state (:494-523), `Surface` (surface.cs:7-20) and display path, and replaces
 * that loop by one P/Invoke call to rt_render() per frame.  INTEGRATION.md shows the C# binding.
 *
 * Conventions: every function returns 0 on success or a negative rt_status; nothing throws or aborts across the
 * ABI; rt_last_error() returns the message of the last failing call on that context.  A context is
 * single-caller (the reference renders from one thread, template.cs:175-179); distinct contexts are independent.
 * There is NO CPU fallback: without a CUDA device every entry point fails with RT_ERR_CUDA.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB200_ABI_VERSION 2

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID = -1,    /* bad argument */
    RT_ERR_CUDA = -2,       /* CUDA runtime error / no device */
    RT_ERR_NO_SCENE = -3,   /* rt_render before rt_set_scene */
    RT_ERR_UNSUPPORTED = -4,/* e.g. max_depth above RT_MAX_DEPTH */
    RT_ERR_PEER = -5        /* peer access / IPC mapping failed */
} rt_status;

#define RT_MAX_DEPTH 32     /* the reference's ReflectionRecursionLimit, RayTracer.cs:490 */

/* accel for rt_set_scene */
#define RT_ACCEL_AUTO 0
#define RT_ACCEL_BRUTE 1    /* the reference's loop over every sphere (RayTracer.cs:577,792,975) */
#define RT_ACCEL_LBVH 2     /* LBVH over spheres; returns the same hits as RT_ACCEL_BRUTE (tests/test_lbvh*.py) */

typedef struct rt_context rt_context;   /* opaque, library-owned; one per C# `RayTracer` instance (RayTracer.cs:437) */

/* Camera of one frame, by value.  Replaces the per-pixel evaluation of _cameraPosition, CameraRightDirection,
 * CameraUpDirection, CameraForwardDirection (RayTracer.cs:494-523, used at :967-971) and `viewParams`
 * (planeWidth, planeHeight, NearClip; RayTracer.cs:892-896).  All f64 trig stays on the C# side. */
typedef struct rt_camera {
    float pos[3];
    float right[3];
    float up[3];
    float forward[3];
    float view_params[3];
} rt_camera;

/* Ray counters use the nearest-first deterministic accounting: primary = w*h*spp; secondary = number of
 * TraceSecondaryRay calls on the selected hit chain (RayTracer.cs:746,857); shadow = IntersectShadowLight calls on
 * the selected chain (:752,864).  Counters are filled only by rt_render_debug (instrumented kernel); rt_render
 * fills the timings and leaves the counters 0. */
typedef struct rt_stats {
    uint64_t primary, shadow, secondary;
    float kernel_ms;   /* device time of the render kernel(s), max over devices */
    float gather_ms;   /* extra device time for the row-tile gather to device 0 (0 when fused into the kernel) */
    float d2h_ms;      /* device->host copy into host_pixels (0 when headless) */
} rt_stats;

/* Creates a context on `n_devices` CUDA devices (1, 2, 4 or 8; device_ids NULL => 0..n-1).  With n > 1 every frame is
 * partitioned by interleaved row tiles and gathered on device_ids[0] over NVLink by peer stores fused into the
 * render kernel.  Replaces `new RayTracer(screen)` (RayTracer.cs:535, template.cs:80).
 * A device id may repeat: the context then runs several of its partitions on one physical GPU, each with its own streams, scene
 * copy and launch thread (how a single-GPU machine exercises the multi-device code; no performance meaning). */
int rt_create(rt_context** out, const int* device_ids, int n_devices);

/* Uploads (copies) the scene: the arrays RayTracer.cs:441-465 and `_ambientLightColor` :469, in C# field order.
 *   spheres: n x 18 floats  center[3], radius, Kd[3], Ka[3], Ks[3], n, Km[3], radiusSquared   (:308-338, :60-80)
 *   planes : n x 20 floats  center[3], normal[3], Kd[3], Ka[3], Ks[3], n, Km[3], isTiled(0/1)  (:260-303)
 *            isTiled is accepted but ignored: the reference's constructor forces it to true (:289).
 *   lights : n x 4 floats   position[3], intensity                                             (:236-255)
 * radiusSquared is passed through, not recomputed (:336). */
int rt_set_scene(rt_context* ctx, const float* spheres, int n_spheres, const float* planes, int n_planes,
                 const float* lights, int n_lights, const float ambient[3], int accel);

/* Replaces the records of spheres [first, first + count) (same 18-float layout) without rebuilding the acceleration
 * structure: an LBVH keeps its topology and is refitted bottom-up (results stay exact; traversal slows down if spheres move
 * far from where they were when rt_set_scene built the tree — call rt_set_scene again to rebuild).  The reference's scene
 * arrays are readonly (RayTracer.cs:441-465); this is the scene-mutation hook of SURVEY §8(f). */
int rt_update_spheres(rt_context* ctx, const float* spheres, int first, int count);

/* Renders one frame — the replacement for the loop RayTracer.cs:898-901 (TracePixel for every x,y).
 *   max_depth : the reference's ReflectionRecursionLimit (32)   spp : 1 in the reference; >1 = jittered extension
 *   host_pixels: w*h int32, 0x00RRGGBB, row-major (`Surface.pixels`, surface.cs:9-20; written as :1038); NULL => headless
 * Synchronous: returns after the frame is in host_pixels (or, headless, after the kernel finished). */
int rt_render(rt_context* ctx, const rt_camera* cam, int width, int height, int max_depth, int spp, uint32_t seed,
              int32_t* host_pixels, rt_stats* stats);

/* Headless batch: renders n_frames frames (one camera each) in one persistent launch per device into a ring of
 * device framebuffers (benchmarks / offline fly-throughs).  host_pixels: n_frames*w*h int32 or NULL. */
int rt_render_batch(rt_context* ctx, const rt_camera* cams, int n_frames, int width, int height, int max_depth,
                    int spp, uint32_t seed, int32_t* host_pixels, rt_stats* stats);

/* Instrumented render (separate kernel instantiation; not for timing).  All outputs nullable:
 *   host_hash  w*h uint32 : order-independent hash of every (ray kind, hit id, t bits, shadow result) on the chain
 *   host_aov_id/t w*h     : primary hit id (sphere index, n_spheres+plane index, -1 none) and its distance
 *   counters[16]          : primary, shadow, secondary, sphere_tests, sphere_disc_pos, plane_tests, shade_diffuse,
 *                           shade_specular, shade_mirror, shaded_hits (same order as the oracle's), then LBVH statistics:
 *                           node visits of primary / secondary / shadow rays, brute-force fallbacks, 0, 0 */
int rt_render_debug(rt_context* ctx, const rt_camera* cam, int width, int height, int max_depth, int spp, uint32_t seed,
                    int32_t* host_pixels, uint32_t* host_hash, int32_t* host_aov_id, float* host_aov_t,
                    uint64_t* counters, rt_stats* stats);

/* Single-ray sphere queries against the uploaded scene (LBVH == brute-force equality harness).
 *   rays6: n x 6 floats origin[3], direction[3] (direction need not be normalised)
 *   kind : 0 primary fold (:975-981), 1 secondary fold (:792-808), 2 shadow any-hit eps=0.001 (:573-582; id = 1 occluded / 0 clear)
 *   accel: RT_ACCEL_BRUTE or RT_ACCEL_LBVH */
int rt_query_spheres(rt_context* ctx, const float* rays6, int n_rays, int kind, int accel, int32_t* out_id, float* out_t);

/* Ray log — the data behind the reference's DEBUG_ENABLE overlay: `TracedRay` RayTracer.cs:424-435, filled at :601, :639, :801 and
 * drawn (a random sample of 500) at :914-933. One record per RAY of the nearest-first chain of each listed pixel (spp = 1):
 *   kind      RayKind :343-362: 0 primary, 1 secondary, 2 shadow
 *   origin, direction   Ray.origin / Ray.direction (shadow rays: direction = the light POSITION, :574)
 *   hit       primary / secondary: the selected hit — sphere index, n_spheres + plane index, -1 none;
 *             shadow: the nearest occluding sphere (eps 0.001, lowest index on ties), -1 if the light is visible
 *   distance  of that hit, 0 when hit = -1;   hit_point = origin + direction * distance (TracedRay.hitPoint)
 *   pixel     y * width + x;   level = bounce level of the surface point the ray starts from / was cast for;   light = light index
 * Records are grouped by pixel in the order of `pixels`; within a pixel: the rays of the chain, then the shadow rays deepest
 * level first (the order the reference's recursion creates them). Always computed by brute force on device 0. */
typedef struct rt_ray_record {
    float origin[3];
    float direction[3];
    float hit_point[3];
    float distance;
    int32_t hit;
    uint32_t kind;
    uint32_t pixel;
    uint32_t level;
    uint32_t light;
    uint32_t reserved;
} rt_ray_record;   /* 64 bytes */
/* pixels: n_pixels linear indices (< width*height). out: room for max_records records (may be NULL with max_records = 0 to
 * size the log). *n_records receives the number of records the pixels produce; at most max_records of them are written. */
int rt_ray_log(rt_context* ctx, const rt_camera* cam, int width, int height, int max_depth, const uint32_t* pixels, int n_pixels,
               rt_ray_record* out, int max_records, int* n_records);

/* Device self-test of the library's hand-scheduled fp32 sequences against the compiler's IEEE code, exhaustive over their domain:
 *   RT_SELFTEST_INV_LEN    1/sqrt as two correctly rounded operations (OpenTK Vector3.Normalize, used at RayTracer.cs:667, :668,
 *                          :687, :706, :854, :971): every float in [2^-65, 2^65)
 *   RT_SELFTEST_PIXEL_DIV  x / w (:964) for every integer 0 <= x < w <= 16384
 *   RT_SELFTEST_INV_LEN_RSQ_SEED  a rejected cheaper variant, kept to document WHY it is rejected (it has mismatches)
 * n_mismatch must come back 0 for all but the rejected variant. */
#define RT_SELFTEST_INV_LEN 0
#define RT_SELFTEST_PIXEL_DIV 1
#define RT_SELFTEST_INV_LEN_RSQ_SEED 2
#define RT_SELFTEST_INV_LEN_PAIR 3   /* the packed two-at-a-time version used by the two-light Phong pass: every float of the range in either half */
int rt_selftest(rt_context* ctx, int test, uint64_t* n_checked, uint64_t* n_mismatch);

/* Tuning options. RT_OPT_COMPACTION (default 0): tiny-scene kernel variant that parks rays needing a third or later bounce in a
 * shared-memory queue (warp-ballot compaction) and finishes them in fully populated warps; identical pixels, spp == 1 only. */
#define RT_OPT_COMPACTION 1
/* RT_OPT_HOST_VIA_GPU0 (default 0): multi-device contexts normally send each device's row tiles to the host over that device's
 * own PCIe link; 1 = gather the frame on device 0 first (NVLink peer stores) and copy it from there. */
#define RT_OPT_HOST_VIA_GPU0 2
/* RT_OPT_PRIMARY_GATE (default 1): tiny-scene kernels skip the sphere loop of primary rays (RayTracer.cs:975-981) for pixels outside
 * a per-frame rectangle the host proves no sphere can be hit in (csrc/rt_gate.cuh); 0 = test every sphere for every pixel. Same pixels. */
#define RT_OPT_PRIMARY_GATE 3
/* RT_OPT_SHARED_TARGET (default 0), one process per GPU: a promise that the `dev_pixels` every rank of the partition passes to
 * rt_render_device alias ONE framebuffer owned by rank 0 (rt_ipc_export / rt_ipc_open). Ranks other than 0 then do not store the
 * spans the frame gates prove black, and rank 0 zero-fills those spans of all tiles locally: 35 % of the default scene's frame
 * never crosses NVLink. 1 = used when the partition has more than 4 ranks (where rank 0's NVLink ingress is the bottleneck; with
 * fewer ranks the fill pass costs rank 0 more than the link saves), 2 = always. Every rank must set the same value (and the same
 * RT_OPT_PRIMARY_GATE). Leave 0 when ranks render into separate buffers. Multi-device contexts (one process) do this by
 * themselves (above 4 devices; 2 forces it there too). */
#define RT_OPT_SHARED_TARGET 4
/* RT_OPT_DEBUG_SHIPPED (default 0): rt_render_debug of a tiny scene (<= 8 spheres) with spp == 1 runs the PRODUCTION kernel
 * instantiation — exact-count unrolled loops, packed fp32 sphere pairs, fast pixel division, frame gates, 4-pixel spans — with an
 * events-only debug policy, so that the per-pixel chain hash / primary AOV / ray counters certify the shipped code path itself.
 * Work a gate skips is reported as the event the reference produces there (a reflection ray that hits nothing, an unoccluded
 * shadow ray, a primary ray that hits nothing). The per-test counters (sphere_tests ... shaded_hits) stay 0 in this mode. */
#define RT_OPT_DEBUG_SHIPPED 5
/* RT_OPT_SPARSE_D2H (default 1): rt_render / rt_render_batch with a host buffer do not copy what the frame gates prove black
 * (whole rows above the horizon outside the sphere rectangle, and the parts of sky rows beside it); those pixels of host_pixels are
 * zero-filled by library threads (RTB200_FILL_THREADS, default 6) while the rest crosses PCIe. Same pixels; 35 % fewer PCIe bytes on
 * the reference's default frame. rt_get_info(RT_INFO_LAST_D2H_BYTES) reports what was really copied. */
#define RT_OPT_SPARSE_D2H 6
/* RT_OPT_HOST_PRECLEARED (default 0): a promise that host_pixels is already all zero when rt_render is called — the reference's
 * Tick() does `screen.Clear(0)` (RayTracer.cs:890) right before the pixel loop — so the library skips its own zero fill of the
 * pixels it does not copy. */
#define RT_OPT_HOST_PRECLEARED 7
/* RT_OPT_HOST_ZERO_COPY (default 1): when host_pixels is page-locked memory the devices can address (rt_host_register, or memory from
 * cudaHostAlloc), rt_render / rt_render_batch can let the render kernel store the frame straight into it over PCIe — no device
 * framebuffer, no copy engine, one launch per device — and with RT_OPT_SPARSE_D2H what the frame gates prove black is neither stored
 * nor copied but zero-filled by host threads. 0 = never (device framebuffer + band-pipelined copies), 1 = for frames up to 40 MB
 * (the reference's 1280x720 window: 0.105 instead of 0.149 ms per frame on B200, 4K: 0.508 vs 0.517; beyond, the copy engine's higher
 * PCIe rate catches up), 2 = always. Pageable host memory always takes the copy path. */
#define RT_OPT_HOST_ZERO_COPY 8
/* Packed multi-GPU gather (csrc/rt_gather.cuh): when every rank stores into rank 0's framebuffer (RT_OPT_SHARED_TARGET, or a multi-device
 * context) and a gather area is attached (rt_gather_attach; multi-device contexts own one), ranks != 0 send their tiles in a
 * compressed wire format — nothing for quads the frame gates prove black, 1 byte per pixel for grey quads, 3 bytes per pixel
 * otherwise — into planes in rank 0's memory, and rank 0 expands them into the 0x00RRGGBB framebuffer; the kernels synchronise among
 * themselves through flags in rank 0's memory. Rank 0 gets a smaller share of the tiles (it also runs the expand pass).
 *   RT_OPT_GATHER_MODE  -1 automatic (2 from 8 ranks on, else 0), 0 off (plain 4 B / pixel stores), 1 RGB24 only, 2 RGB24 + grey quads
 *   RT_OPT_SINK_TILES / RT_OPT_PEER_TILES  tiles per period for rank 0 / for every other rank (0 = automatic: 1/1 for 2 ranks,
 *                        4/5 for 4, 1/2 for 8); period = sink + peer * (world - 1) <= 64
 * Applies to single-sample frames of tiny scenes whose width is a multiple of 128; everything else takes the plain gather. All three
 * options must be identical on every rank. Frames are byte-identical either way (tests/test_gpu_shipped_path.py). */
#define RT_OPT_GATHER_MODE 9
#define RT_OPT_SINK_TILES 10
#define RT_OPT_PEER_TILES 11
/* RT_OPT_HOST_SHADOW_BINS (default 0): build the per-light shadow bins of LBVH scenes on the host instead of on the GPU (same
 * geometry code, same bins; 0.1-0.3 s at 100 k spheres instead of ~2 ms). Takes effect at the next rt_set_scene / rt_update_spheres.
 * Exists for the test that shows both builds agree. */
#define RT_OPT_HOST_SHADOW_BINS 12
/* RT_OPT_PRIMARY_BINS (default 1; environment RTB200_PRIMARY_BINS overrides the default at rt_create): LBVH scenes, single-sample frames — per frame (whenever the camera or the frame size changes) every sphere is
 * projected to the pixel rectangle outside of which no primary ray can be reported as hitting it (the rectangle of the tiny scenes'
 * frame gates, csrc/rt_gate.cuh) and entered into the 8 x 8-pixel tiles it touches, on the GPU (three small launches, nothing read
 * back); a primary ray then runs the reference's sphere test (RayTracer.cs:613-642) over the list of ITS tile and folds as :975-981
 * does, instead of walking the tree. Tiles with more than 32 spheres (the horizon of the 100 k-sphere scene) keep no list and
 * traverse as before. Same pixels, hit ids and t bits either way (csrc/rt_primary_bins.cuh). */
#define RT_OPT_PRIMARY_BINS 13
/* Options that change WHICH rank writes a pixel (RT_OPT_SHARED_TARGET, RT_OPT_PRIMARY_GATE) must be set identically on every rank
 * of a partition. RT_OPT_COMPACTION may differ: a launch that takes part in a sparse gather always uses the default kernel. */
int rt_set_option(rt_context* ctx, int option, int value);

/* Counters / facts about the context. */
#define RT_INFO_GATE_HOST_NS 1      /* host nanoseconds spent computing frame gates (csrc/rt_gate.cuh) so far */
#define RT_INFO_GATE_COMPUTES 2     /* number of frame-gate evaluations so far (a camera that does not move is cached) */
#define RT_INFO_LAST_D2H_BYTES 3    /* bytes the last rt_render / rt_render_batch with a host buffer copied device -> host */
#define RT_INFO_SCENE_PATH 4        /* how the uploaded scene is traced: 0 tiny (constant bank), 1 staged (shared memory), 2 global, 3 LBVH */
#define RT_INFO_LAST_FILL_BYTES 5   /* bytes of host_pixels that render zero-filled on the host instead of copying them */
#define RT_INFO_LAST_FILL_WAIT_NS 6 /* nanoseconds its calling thread spent in (helping with) that fill after enqueuing the GPU work */
#define RT_INFO_GATHER_TIMEOUTS 9   /* packed gather: spin waits that gave up after 2 s (must be 0; a non-zero value means ranks were out of step) */
#define RT_INFO_GATHER_ACTIVE 10    /* 1 if the last rt_render_device launch group used the packed gather */
#define RT_INFO_LAST_ENQUEUE_NS 7   /* host nanoseconds of the last rt_render / rt_render_batch from entry until all GPU work was enqueued */
#define RT_INFO_LAST_TOTAL_NS 8     /* ... from entry to return */
#define RT_INFO_SHADOW_BINS_NS 11   /* host wall-clock nanoseconds of the last shadow-bin (re)build (rt_set_scene / rt_update_spheres; LBVH scenes) */
#define RT_INFO_SHADOW_BIN_PAIRS 12 /* sphere pairs stored in the bins of device 0 (32 bytes each) */
#define RT_INFO_PRIMARY_BIN_BUILDS 13 /* per-frame primary-bin builds so far, all devices (a camera that does not move is cached) */
#define RT_INFO_PRIMARY_BINS 14     /* 1 if RT_OPT_PRIMARY_BINS is on */
int rt_get_info(const rt_context* ctx, int what, uint64_t* value);

/* ---- device-pointer / multi-process interface (torchrun: one process per GPU) --------------------------------- */

/* Row-tile partition of this context inside a `world` of cooperating contexts (default rank 0 of 1). Tile t of
 * `tile_rows` rows belongs to rank t % world. rt_render / rt_render_batch of a partitioned context return ONLY this rank's tiles
 * to host_pixels (the ranks of a job pass one shared page-locked frame and fill it together, each over its own PCIe link). */
int rt_set_partition(rt_context* ctx, int rank, int world, int tile_rows);

/* Renders this context's row tiles of one frame (or of n_frames frames, frame f at dev_pixels + f*w*h) into a
 * caller-provided DEVICE buffer, asynchronously on `cuda_stream` (a cudaStream_t, NULL = the legacy default stream).
 * LBVH scenes keep one camera-inflated node copy per device: launches on different streams are ordered by the library (a refit
 * waits for the previous LBVH launch), so they serialise on the device but never read a half-refitted tree.
 * dev_pixels may be a peer mapping of another GPU's framebuffer (CUDA IPC): the gather is then the kernel's own stores. */
int rt_render_device(rt_context* ctx, const rt_camera* cams, int n_frames, int width, int height, int max_depth, int spp,
                     uint32_t seed, void* dev_pixels, void* cuda_stream);

/* Gather area of the packed multi-GPU gather (see RT_OPT_GATHER_MODE). One process per GPU: rank 0 allocates rt_gather_bytes(w, h)
 * bytes (rt_dev_alloc), ZEROES them (rt_dev_memset) and exports them (rt_ipc_export); every rank — rank 0 with its own pointer, the
 * others with the pointer rt_ipc_open returned — then calls rt_gather_attach, which also restarts the epoch counter: all ranks must
 * attach before any of them renders, and must issue the same sequence of rt_render_device calls afterwards. NULL detaches. */
uint64_t rt_gather_bytes(int width, int height);
int rt_gather_attach(rt_context* ctx, void* area_dev_ptr, uint64_t bytes);

/* CUDA IPC plumbing so that ranks 1..N-1 can store straight into rank 0's framebuffer. handle64 = 64 bytes. */
int rt_ipc_export(rt_context* ctx, void* dev_ptr, void* handle64);
int rt_ipc_open(rt_context* ctx, const void* handle64, void** out_dev_ptr);
int rt_ipc_close(rt_context* ctx, void* dev_ptr);

/* Device memory owned by the context (so that non-torch hosts need no CUDA runtime of their own). */
int rt_dev_alloc(rt_context* ctx, uint64_t bytes, void** out_dev_ptr);
int rt_dev_free(rt_context* ctx, void* dev_ptr);
int rt_dev_to_host(rt_context* ctx, void* host_dst, const void* dev_src, uint64_t bytes);
int rt_dev_memset(rt_context* ctx, void* dev_dst, int byte_value, uint64_t bytes);   /* synchronous; tests poison framebuffers with it */
int rt_sync(rt_context* ctx);

/* Page-locks and maps (cudaHostRegister, portable + mapped) a host buffer the caller keeps alive — e.g. the pinned `Surface.pixels`
 * array — so that rt_render can return the frame at full PCIe rate: by the render kernel's own stores (RT_OPT_HOST_ZERO_COPY) or by
 * copy-engine transfers, instead of through a pageable staging copy. Optional. */
int rt_host_register(rt_context* ctx, void* host_ptr, uint64_t bytes);
int rt_host_unregister(rt_context* ctx, void* host_ptr);

/* ---- zero-copy display ---------------------------------------------------------------------------------
 * Replaces the per-frame upload of Surface.pixels from HOST memory, template.cs:81 (`GL.TexImage2D(..., screen.pixels)`) and
 * :188-193: the frame is rendered straight into a device buffer the display owns and never crosses PCIe.
 *
 * rt_render_mapped: `mapped_dev_pixels` is a DEVICE pointer to at least 4*width*height bytes on the context's first device — the
 * pointer cudaGraphicsResourceGetMappedPointer returns for a mapped OpenGL pixel-unpack buffer (or any cudaMalloc'd buffer).
 * Synchronous: when it returns the frame (0x00RRGGBB words, row-major, as Surface.pixels) is in the buffer and the host may unmap
 * it and call glTexSubImage2D from the bound PBO. One device: the render kernel's 128-bit stores go straight into the buffer;
 * several devices: the frame is gathered on device 0 (peer stores) and copied device-to-device. */
int rt_render_mapped(rt_context* ctx, const rt_camera* cam, int width, int height, int max_depth, int spp, uint32_t seed,
                     void* mapped_dev_pixels, uint64_t mapped_bytes, rt_stats* stats);
/* Convenience wrappers so that the C# host needs no CUDA binding of its own. The calling thread must have the OpenGL context
 * current (template.cs runs everything on the render thread). gl_buffer: a GL buffer object (PixelUnpackBuffer) of 4*w*h bytes.
 *   rt_gl_register_buffer   cudaGraphicsGLRegisterBuffer(write-discard), once per buffer (re-register after a resize)
 *   rt_render_gl            map -> rt_render_mapped -> unmap
 *   rt_gl_unregister_buffer before the GL buffer is deleted (rt_destroy unregisters what is left)
 * Without an OpenGL context rt_gl_register_buffer fails with RT_ERR_CUDA (nothing crashes). */
int rt_gl_register_buffer(rt_context* ctx, unsigned int gl_buffer, void** out_resource);
int rt_gl_unregister_buffer(rt_context* ctx, void* resource);
int rt_render_gl(rt_context* ctx, const rt_camera* cam, int width, int height, int max_depth, int spp, uint32_t seed,
                 void* resource, rt_stats* stats);

/* Measured L2 -> SM read bandwidth in GB/s on the context's first device for a working set of `bytes` (1 MB .. 96 MB, L2-resident):
 * every CTA of a full grid streams the set with 128-bit L1-bypassing loads; best of 3 launches. bench.py uses it as the peak of the
 * LBVH kernels' roofline (their node / leaf fetches are L2 traffic). */
int rt_measure_l2_read(rt_context* ctx, uint64_t bytes, double* gbs);

/* Number of render-kernel launches issued by this context so far (bench.py reports it as gpu_launches). */
uint64_t rt_launch_count(const rt_context* ctx);

int rt_destroy(rt_context* ctx);
const char* rt_last_error(const rt_context* ctx);   /* valid until the next call on ctx; ctx NULL => last create error */
int rt_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
