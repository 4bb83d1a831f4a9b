#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ray-cast backend (contract: see the task brief / DESIGN.md §measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (librtb200.so through its C ABI)
  python bench.py --impl reference [...]                          # the reference's CPU path (oracle port, host cores)

Workload = BASELINE.json configs[1]: the reference's default scene (RayTracer.cs:441-469) at 3840x2160, recursion cap 8.
A STEP is one pass of the hot path over one batch of `--frames` (default 16) such frames, each written to its own
framebuffer of a ring (16 x 33.2 MB = 531 MB > the 126 MB L2, so no frame's stores hit lines left by the previous one).
Metric: Mrays/s (primary + shadow + secondary, nearest-first accounting — DESIGN.md), whole job over all N GPUs.

N > 1 (torchrun, one process per GPU): every frame is cut into interleaved row tiles (tile t -> rank t % N); each rank's
kernel stores its tiles straight into rank 0's framebuffer through a CUDA-IPC peer mapping (gather fused into the render
kernel, over NVLink).  Total work is fixed as N grows => "scaling": "strong".

torch is plumbing only here (process group, events, pinned memory); every kernel in the timed region is ours.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))

import scenes  # noqa: E402

W, H, DEPTH = 3840, 2160, 8
WORKLOAD = "default_scene_3840x2160_depth8"       # BASELINE.json configs[1]
METRIC = "Mrays/s (primary+shadow+secondary) at 4K"
UNIT = "Mrays/s"


def algorithmic_flops(c: dict) -> float:
    """SURVEY §8d flop model (each + - * / sqrt min max = 1): sphere test 24 (+14 when the discriminant >= 0), plane test
    17, primary-ray generation 38, per shaded hit 6 (hit point) + 35 per diffuse light term + 32 more with specular +
    9 per reflection."""
    return (24.0 * c["sphere_tests"] + 14.0 * c["sphere_disc_pos"] + 17.0 * c["plane_tests"] + 38.0 * c["primary"]
            + 6.0 * c["shaded_hits"] + 35.0 * c["shade_diffuse"] + 32.0 * c["shade_specular"] + 9.0 * c["shade_mirror"])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.p = None
        # nvidia-smi numbers physical GPUs; map the CUDA ordinal through CUDA_VISIBLE_DEVICES (indices or UUIDs) if it is set
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        self.idx = vis[gpu_index] if gpu_index < len(vis) else gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.06)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill(); out, _ = self.p.communicate()
        sm, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))


def cpu_reference_run(frames: int, warm: int, threads: int = 0):
    """Times the CPU oracle in FAITHFUL mode (shade every intersected primitive, per-column fork/join: the work the
    reference actually does, RayTracer.cs:898-901) on all host cores. Returns (best_seconds_per_frame, all_times, rays)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if threads <= 0:    # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on explicitly
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=W, height=H)
    cnt = O.render(sc, cam, W, H, DEPTH, mode="nearest", threads=threads)["counters"]
    rays = cnt["primary"] + cnt["shadow"] + cnt["secondary"]
    times = []
    for i in range(warm + frames):
        t0 = time.perf_counter()
        O.render(sc, cam, W, H, DEPTH, mode="faithful", threads=threads)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return min(times), times, rays, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    # one "step" = a bounded sample of the workload: ONE of the step's identical 4K frames, faithful mode
    best, times, rays, cores = cpu_reference_run(frames=max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    mean = float(np.mean(times))
    val = rays / mean / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "1 frame per step (our arm: %d identical frames per step)" % args.frames,
                   "mode": "faithful (shade-all-then-select, per-column fork/join)", "rays_per_frame": rays},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d full 3840x2160 frames, C++ strict-fp restatement of RayTracer.cs (no .NET in this image)" % len(times)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def moving_cameras(n: int):
    """n distinct cameras around the default one: +-1-pixel yaw steps (one 4K pixel = view-plane width / 3840 / near clip radians),
    so that every frame of a step needs its own host-side frame gates (csrc/rt_gate.cuh) — what a moving camera pays."""
    px = float(scenes.view_params(W, H)[0]) / W / float(scenes.NEAR_CLIP)
    return np.stack([scenes.make_camera(yaw=(k - n // 2) * px, width=W, height=H) for k in range(n)])


def time_steps(torch, stream, fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def per_config_extras(rtb200, local_rank, peaks):
    """BASELINE.json configs[0], [2], [3], [4] on ONE GPU, after the headline loop, in this process: single-frame launches through
    rt_render (headless), kernel time from CUDA events inside the library (best of 5 after 2 warm-ups), ray / LBVH counters from the
    instrumented kernel, and the returned frame compared with the instrumented render of the same frame. LBVH configs also carry
    `primary_bins`: RT_OPT_PRIMARY_BINS off and on in the same context, camera standing still and moving every frame."""
    import zlib
    out = {}
    l2_gbs = None
    specs = [
        ("configs[0] default scene 1280x720 cap 32", scenes.default_scene, {}, 1280, 720, 32, 1, rtb200.RT_ACCEL_AUTO),
        ("configs[2] 1024 spheres 4K cap 8, LBVH + shadow bins", scenes.config3_scene, scenes.SCALED_CAMERA, W, H, 8, 1, rtb200.RT_ACCEL_LBVH),
        ("configs[2] 1024 spheres 4K cap 8, brute force staged in shared memory", scenes.config3_scene, scenes.SCALED_CAMERA, W, H, 8, 1, rtb200.RT_ACCEL_BRUTE),
        ("configs[3] 100k spheres 4K cap 8, LBVH + shadow bins", scenes.config4_scene, scenes.SCALED_CAMERA, W, H, 8, 1, rtb200.RT_ACCEL_AUTO),
        ("configs[4] default scene 7680x4320 x 16 spp cap 8", scenes.default_scene, {}, 7680, 4320, 8, 16, rtb200.RT_ACCEL_AUTO),
    ]
    for name, mk, camkw, w, h, depth, spp, accel in specs:
        sc = mk()
        cam = scenes.make_camera(width=w, height=h, **camkw)
        ctx = rtb200.Context([local_rank])
        t0 = time.perf_counter()
        ctx.set_scene(sc, accel)
        upload_s = time.perf_counter() - t0
        dbg = ctx.render_debug(cam, w, h, depth, spp, 0, arrays=False)
        cnt = dbg["counters"]
        rays = cnt["primary"] + cnt["shadow"] + cnt["secondary"]
        for _ in range(2):
            ctx.render(cam, w, h, depth, spp, 0, headless=True)
        ms = [ctx.render(cam, w, h, depth, spp, 0, headless=True)[1].kernel_ms for _ in range(5)]
        px, _ = ctx.render(cam, w, h, depth, spp, 0)
        rec = {"kernel_ms": min(ms), "kernel_ms_mean": float(np.mean(ms)), "mrays_per_s": rays / (min(ms) * 1e-3) / 1e6,
               "rays": rays, "counters": cnt, "frame_crc32": zlib.crc32(px.tobytes()) & 0xFFFFFFFF,
               "frame_equals_instrumented_render": bool(np.array_equal(px, dbg["pixels"])), "scene_upload_s": upload_s,
               "path": ["tiny", "staged", "global", "lbvh"][ctx.get_info(rtb200.RT_INFO_SCENE_PATH)]}
        if rec["path"] == "lbvh":
            # RT_OPT_PRIMARY_BINS (on by default) both ways in this context: a camera that stands still (per-frame LBVH state cached) and
            # one that moves every frame (camera-inflated boxes refit + primary bins rebuilt inside the timed region); frames compared
            default_on = ctx.get_info(rtb200.RT_INFO_PRIMARY_BINS)
            pbs = {"default": bool(default_on)}
            try:
                cam_b = scenes.make_camera(width=w, height=h, **dict(camkw, yaw=camkw.get("yaw", 0.0) + 1e-3))
                for setting in (0, 1):
                    ctx.set_option(rtb200.RT_OPT_PRIMARY_BINS, setting)
                    for _ in range(2):
                        ctx.render(cam, w, h, depth, spp, 0, headless=True)
                    still = [ctx.render(cam, w, h, depth, spp, 0, headless=True)[1].kernel_ms for _ in range(5)]
                    moving = [ctx.render(cam_b if k % 2 == 0 else cam, w, h, depth, spp, 0, headless=True)[1].kernel_ms for k in range(6)]
                    px2, _ = ctx.render(cam, w, h, depth, spp, 0)
                    nv = ctx.render_debug(cam, w, h, depth, spp, 0, arrays=False)["lbvh"]["node_visits_primary"]
                    pbs["on" if setting else "off"] = {"kernel_ms": min(still), "kernel_ms_moving_camera": min(moving), "node_visits_primary": nv,
                                                       "frame_equals_instrumented_render": bool(np.array_equal(px2, dbg["pixels"]))}
            except Exception as e:                        # an auxiliary A/B: never lose the bench line over it
                pbs["error"] = repr(e)
            finally:
                ctx.set_option(rtb200.RT_OPT_PRIMARY_BINS, default_on)
            rec["primary_bins"] = pbs
            lb = dbg["lbvh"]
            visits = lb["node_visits_primary"] + lb["node_visits_secondary"] + lb["node_visits_shadow"]
            nbytes = visits * 64 + cnt["sphere_tests"] * 16
            if l2_gbs is None:
                l2_gbs = ctx.measure_l2_read(32 << 20)
            ach = nbytes / (min(ms) * 1e-3) / 1e9
            rec["lbvh"] = lb
            rec["roofline"] = {"bound": "L2", "achieved": ach, "peak": l2_gbs, "unit": "GB/s", "frac": ach / l2_gbs,
                               "bytes": nbytes, "note": "algorithmic bytes = node visits x 64 B + sphere tests x 16 B (SURVEY §8d); peak = L2->SM read "
                                                        "bandwidth measured in this run by rt_measure_l2_read (every CTA of a full grid streams a 32 MB "
                                                        "L2-resident set with 128-bit L1-bypassing loads, best of 3)"}
        else:
            fl = algorithmic_flops(cnt)
            fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
            rec["roofline"] = {"bound": "fp32", "achieved": fl / (min(ms) * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                               "frac": fl / (min(ms) * 1e-3) / 1e12 / fp32_peak}
        out[name] = rec
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=16, help="4K frames per step (ring of framebuffers > L2)")
    ap.add_argument("--tile-rows", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip per_config / single-frame / moving-camera / gates-off measurements")
    ap.add_argument("--compaction", type=int, default=0, help="1: opt-in warp-ballot compaction kernel variant (tuning)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import rtb200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — librtb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    F = args.frames
    assert 1 <= F <= 16
    npix = W * H
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=W, height=H)
    cams = np.repeat(cam[None], F, 0)
    cams_move = moving_cameras(F)

    ctx = rtb200.Context([local_rank])
    ctx.set_scene(sc)
    ctx.set_partition(rank, world, args.tile_rows)
    if args.compaction:
        ctx.set_option(rtb200.RT_OPT_COMPACTION, 1)

    # ray / flop accounting from the instrumented kernel (not timed); identical on every rank
    dbg = ctx.render_debug(cam, W, H, DEPTH)
    cnt = dbg["counters"]
    rays_per_frame = cnt["primary"] + cnt["shadow"] + cnt["secondary"]
    flops_per_frame = algorithmic_flops(cnt)
    # ... and of the moving-camera frames (e2e, value_moving_camera); the last one is kept to check the frame that comes back
    rays_move, dbg_move_last = 0, None
    for f in range(F):
        d = ctx.render_debug(cams_move[f], W, H, DEPTH, arrays=False)
        rays_move += d["counters"]["primary"] + d["counters"]["shadow"] + d["counters"]["secondary"]
        dbg_move_last = d["pixels"]

    # framebuffer ring on rank 0; other ranks map it through CUDA IPC (peer stores over NVLink)
    fb_bytes = F * npix * 4
    if rank == 0:
        fb = ctx.dev_alloc(fb_bytes)
        handle = [ctx.ipc_export(fb)] if world > 1 else None
    else:
        handle = [None]
    if world > 1:
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            fb = ctx.ipc_open(handle[0])

    # gather area of the packed multi-GPU gather (csrc/rt_gather.cuh), in rank 0's memory like the framebuffer ring
    ga = None
    if world > 1:
        ga_bytes = ctx.gather_bytes(W, H)
        if rank == 0:
            ga = ctx.dev_alloc(ga_bytes)
            ctx.dev_memset(ga, 0, ga_bytes)
            h2 = [ctx.ipc_export(ga)]
        else:
            h2 = [None]
        dist.broadcast_object_list(h2, src=0)
        if rank != 0:
            ga = ctx.ipc_open(h2[0])
        ctx.gather_attach(ga, ga_bytes)
        dist.barrier()

    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()

    if world > 1:
        # all ranks store into rank 0's framebuffer: spans the frame gates prove black are not sent over NVLink, rank 0 zero-fills
        # them locally (RT_OPT_SHARED_TARGET). The ring is poisoned first so that a span nobody writes would show up in the
        # frame comparisons below.
        ctx.set_option(rtb200.RT_OPT_SHARED_TARGET, 1)
        if rank == 0:
            ctx.dev_memset(fb, 0x5A, fb_bytes)
        barrier()

    def step():
        ctx.render_device(cams, W, H, DEPTH, 1, 0, fb, sh)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize(); barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    launches_before = ctx.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize(); barrier()
    launches_timed = ctx.launch_count() - launches_before
    gather_active = ctx.get_info(rtb200.RT_INFO_GATHER_ACTIVE) if world > 1 else 0
    gather_timeouts = ctx.get_info(rtb200.RT_INFO_GATHER_TIMEOUTS) if world > 1 else 0
    assert gather_timeouts == 0, "packed gather: %d spin waits timed out (ranks out of step)" % gather_timeouts
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    launches_all = torch.tensor([float(launches_timed)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches_all, op=dist.ReduceOp.SUM)      # rank 0 also launches the sparse-gather fill kernel above 4 ranks
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and not clocks.get("samples"):
        # the timed region was shorter than nvidia-smi's sampling period: sample the same step loop again, untimed
        sampler = ClockSampler(local_rank); sampler.start()
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            for _ in range(8):
                step()
            torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks["source"] = "untimed repeat of the step loop (timed region shorter than the sampling period)"
    if world > 1:
        barrier()

    # ---- N = 1 only: the same kernel seen three other ways (not the headline; each explains a part of it) ---------------------
    extras = {}
    if world == 1 and not args.no_extras:
        k = max(3, min(args.steps, 20))
        # (1) one frame per launch — what Tick() does (RayTracer.cs:886-901): launch latency and the tail of every frame included
        def single():
            for f in range(F):
                ctx.render_device(cam[None], W, H, DEPTH, 1, 0, fb + f * npix * 4, sh)
        single()
        t = time_steps(torch, stream, single, k)
        extras["value_single_frame"] = {"value": rays_per_frame * F * k / (t * 1e-3) / 1e6, "unit": UNIT, "ms_per_frame": t / (k * F),
                                        "note": "one 4K frame per launch, back to back on one stream (what Tick() does)"}
        # (2) a moving camera: 16 DISTINCT cameras per step, so every frame's gates are computed on the host inside the timed region
        def moving():
            ctx.render_device(cams_move, W, H, DEPTH, 1, 0, fb, sh)
        moving()
        g_ns0, g_n0 = ctx.get_info(rtb200.RT_INFO_GATE_HOST_NS), ctx.get_info(rtb200.RT_INFO_GATE_COMPUTES)
        t = time_steps(torch, stream, moving, k)
        g_ns1, g_n1 = ctx.get_info(rtb200.RT_INFO_GATE_HOST_NS), ctx.get_info(rtb200.RT_INFO_GATE_COMPUTES)
        extras["value_moving_camera"] = {"value": rays_move * k / (t * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": t / k,
                                         "gate_host_us": (g_ns1 - g_ns0) / max(1, g_n1 - g_n0) / 1e3, "gate_computes": g_n1 - g_n0,
                                         "note": "16 distinct cameras per step (+-1-pixel yaw steps): frame gates recomputed for every frame"}
        # (3) frame gates off: every pixel traces everything the reference traces (the rays actually traced)
        ctx.set_option(rtb200.RT_OPT_PRIMARY_GATE, 0)
        step()
        t = time_steps(torch, stream, step, k)
        ctx.set_option(rtb200.RT_OPT_PRIMARY_GATE, 1)
        extras["value_gates_off"] = {"value": rays_per_frame * F * k / (t * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": t / k,
                                     "note": "RT_OPT_PRIMARY_GATE = 0: no host-proven skips, all %d accounted rays are traced" % rays_per_frame}
        step(); torch.cuda.synchronize()

    # ---- N > 1 only: the same step with a library collective instead of the fused peer stores (comparison, not the product):
    # every rank renders its tiles into a LOCAL framebuffer, packs its rows, NCCL-gathers them to rank 0, rank 0 scatters
    # them into the frame (partition.pack_rows / assemble — the code the gloo test covers on CPU).
    nccl_ms = None
    if world > 1:
        import partition
        local = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
        rows = partition.rows_of_rank(H, args.tile_rows, rank, world)
        pad = partition.max_rows_per_rank(H, args.tile_rows, world)
        rows_t = torch.as_tensor(rows, device="cuda")
        all_rows = [torch.as_tensor(partition.rows_of_rank(H, args.tile_rows, r, world), device="cuda") for r in range(world)]
        payload = torch.zeros((F, pad, W), dtype=torch.int32, device="cuda")
        gathered = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
        frame0 = torch.empty((F, H, W), dtype=torch.int32, device="cuda") if rank == 0 else None

        ctx.set_option(rtb200.RT_OPT_SHARED_TARGET, 0)       # separate local framebuffers: every rank writes all of its tiles

        def nccl_step():
            ctx.render_device(cams, W, H, DEPTH, 1, 0, local.data_ptr(), sh)
            payload[:, : len(rows)] = local.index_select(1, rows_t)
            dist.gather(payload, gathered, dst=0)
            if rank == 0:
                for r in range(world):
                    frame0[:, all_rows[r]] = gathered[r][:, : len(all_rows[r])]

        nccl_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            nccl_step()
        torch.cuda.synchronize(); barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for _ in range(nccl_steps):
            nccl_step()
        n1.record(stream)
        torch.cuda.synchronize(); barrier()
        t = torch.tensor([n0.elapsed_time(n1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nccl_ms = float(t.item()) / nccl_steps
        if rank == 0:     # both gathers must produce the same frames
            fused = torch.from_numpy(ctx.dev_to_host(fb, fb_bytes).reshape(F, H, W)).cuda()
            assert torch.equal(fused, frame0), "NCCL-gathered frames differ from the fused peer-store frames"
            del fused
        del local, payload, gathered, frame0
        ctx.set_option(rtb200.RT_OPT_SHARED_TARGET, 1)

    # ---- e2e: the reference-facing call, rt_render, with HOST buffers: per frame the camera goes down, the kernel runs and the
    # frame comes back into a page-locked Surface.pixels. 16 DISTINCT cameras per step (the host-side gate computation is inside
    # the timing). N = 1: one context. N > 1: every rank's context is one rank of the row-tile partition and returns ITS OWN tiles
    # over ITS OWN PCIe link into ONE frame in shared host memory (POSIX shm, page-locked by every rank) — no gather on GPU 0 at all.
    e2e_steps = max(2, min(args.steps, 5))
    shm = None
    if world == 1:
        host = torch.empty((F, npix), dtype=torch.int32, pin_memory=True)
        host_np = host.numpy()
    else:
        from multiprocessing import shared_memory
        name = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=fb_bytes)
            name[0] = shm.name
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
        host_np = np.ndarray((F, npix), dtype=np.int32, buffer=shm.buf)
        if rank == 0:
            host_np[...] = 0x5A5A5A5A
        barrier()
        ctx.host_register(host_np)
    d2h_step = [0]
    rt_cams_move = [rtb200.to_rt_camera(c) for c in cams_move]        # the C# host passes a blittable struct: no per-call conversion
    host_frames = [host_np[f].reshape(H, W) for f in range(F)]

    fill_wait_us = [0.0]

    def e2e_step():
        n, fw = 0, 0
        for f in range(F):
            ctx.render(rt_cams_move[f], W, H, DEPTH, out=host_frames[f])
            n += ctx.get_info(rtb200.RT_INFO_LAST_D2H_BYTES)
            fw += ctx.get_info(rtb200.RT_INFO_LAST_FILL_WAIT_NS)
        d2h_step[0] = n
        fill_wait_us[0] = fw / F / 1e3
        barrier()           # N > 1: the step is done when every rank's tiles of every frame have landed

    e2e_step()
    torch.cuda.synchronize(); barrier()
    g_ns0, g_n0 = ctx.get_info(rtb200.RT_INFO_GATE_HOST_NS), ctx.get_info(rtb200.RT_INFO_GATE_COMPUTES)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(); barrier()
    e2e_s = time.perf_counter() - t0
    g_ns1, g_n1 = ctx.get_info(rtb200.RT_INFO_GATE_HOST_NS), ctx.get_info(rtb200.RT_INFO_GATE_COMPUTES)
    e2e_t = torch.tensor([e2e_s], device="cuda")
    d2h_t = torch.tensor([float(d2h_step[0])], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(d2h_t, op=dist.ReduceOp.SUM)
    e2e_s = float(e2e_t.item())
    d2h_bytes_step = int(d2h_t.item())
    # sanity: the frame that came back is the frame the oracle-checked debug kernel produced
    if rank == 0:
        assert np.array_equal(host_np[F - 1].reshape(H, W), dbg_move_last), "e2e frame differs from the instrumented render"
    fill_wait_sparse = fill_wait_us[0]
    # the dense return (RT_OPT_SPARSE_D2H = 0), for comparison: every byte of every frame crosses PCIe
    ctx.set_option(rtb200.RT_OPT_SPARSE_D2H, 0)
    e2e_step(); torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(); barrier()
    dense_t = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(dense_t, op=dist.ReduceOp.MAX)
    e2e_dense_s = float(dense_t.item())
    ctx.set_option(rtb200.RT_OPT_SPARSE_D2H, 1)
    if world > 1:
        ctx.host_unregister(host_np)

    if rank == 0:
        peaks = measured_peaks()
        rays_step = rays_per_frame * F
        value = rays_step * args.steps / (total_ms * 1e-3) / 1e6
        ms_per_step = total_ms / args.steps
        kernel_ms = ms_per_step                        # one launch per step per GPU
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12        # TFLOP/s, FMA counted as 2
        achieved_tflops = flops_per_frame * F / world / (kernel_ms * 1e-3) / 1e12
        hbm_ach = npix * 4 * F / world / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("frames_per_launch") == F and world == 1:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": F, "rays_per_frame": rays_per_frame,
                       "l2": "ring of %d framebuffers (%.0f MB) larger than the 126 MB L2; the path reads no input from HBM" % (F, fb_bytes / 1e6),
                       "partition": ("interleaved row tiles of %d rows into rank 0's framebuffer over NVLink (CUDA IPC): " % args.tile_rows +
                                     ("packed gather — ranks != 0 send nothing / 1 B / 3 B per pixel for black / grey / coloured quads into planes on GPU 0, "
                                      "rank 0 renders a smaller share and expands them; device-side flags, no collective" if gather_active else
                                      "tile t -> rank t % N, plain 128-bit peer stores"))
                       if world > 1 else "single GPU", "packed_gather": bool(gather_active)},
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp32_peak, "traffic": traffic,
                         "note": "kernel is fp32-instruction bound, not HBM or tensor: peak = 148 SM x 128 lanes x 2 (FMA) x sm_max_mhz "
                                 "(%s); parity forbids FMA contraction, so the attainable ceiling is peak/2; achieved = SURVEY §8d "
                                 "algorithmic flops (%.3e per frame: every ray tests every sphere, as the reference does) / CUDA-event kernel time; the "
                                 "kernel skips tests the host's frame gates prove fruitless (csrc/rt_gate.cuh), so it executes fewer — see "
                                 "value_gates_off" % (peaks["source"], flops_per_frame),
                         "hbm_write": {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"]}},
            "e2e": {"value": rays_move * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 60 * F * world,
                    "d2h_bytes_per_step": d2h_bytes_step, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "frame_bytes_per_step": fb_bytes,
                    "gate_host_us": (g_ns1 - g_ns0) / max(1, g_n1 - g_n0) / 1e3, "host_fill_wait_us_per_frame": fill_wait_sparse,
                    "dense_return": {"value": rays_move * e2e_steps / e2e_dense_s / 1e6, "d2h_bytes_per_step": fb_bytes,
                                     "note": "RT_OPT_SPARSE_D2H = 0: every byte of every frame crosses PCIe"},
                    "path": ("rt_render per frame into a page-locked host Surface.pixels, 16 distinct cameras per step; pixels the frame gates prove "
                             "black are not copied but zero-filled by library threads (RT_OPT_SPARSE_D2H)") if world == 1 else
                            ("rt_render per frame on every rank (each a rank of the row-tile partition): every GPU returns its own tiles over its own "
                             "PCIe link into ONE frame in shared page-locked host memory; 16 distinct cameras per step; sparse return")},
            "gpu_launches": int(launches_all.item()),
            "gather_compare": None if nccl_ms is None else {
                "fused_peer_stores_ms_per_step": ms_per_step, "nccl_gather_ms_per_step": nccl_ms,
                "note": "same step; NCCL path = render to a local framebuffer, pack rows, torch.distributed.gather, scatter on rank 0"},
            "clocks": clocks,
        }
        line.update(extras)
        if world == 1 and not args.no_extras:
            line["per_config"] = per_config_extras(rtb200, local_rank, peaks)
        if not args.no_cpu_baseline and world == 1:
            best, times, rays, cores = cpu_reference_run(frames=5, warm=1)
            line["cpu_baseline"] = {"value": rays / best / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "best of 5 full 3840x2160 frames after 1 warm-up, faithful mode, C++ strict-fp "
                                              "restatement of RayTracer.cs (no .NET in this image; likely faster than the C# JIT)"}
        print(json.dumps(line), flush=True)

    barrier()
    if world > 1:
        ctx.gather_attach(None)
    if rank != 0 and world > 1:
        ctx.ipc_close(fb)
        ctx.ipc_close(ga)
    barrier()
    if rank == 0:
        ctx.dev_free(fb)
        if ga is not None:
            ctx.dev_free(ga)
    ctx.close()
    if shm is not None:
        host_frames.clear()
        del host_np
        import gc
        gc.collect()
        shm.close()
        barrier()
        if rank == 0:
            shm.unlink()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
