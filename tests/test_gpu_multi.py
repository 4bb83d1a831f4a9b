"""Multi-GPU tests: interleaved row tiles across GPUs with the gather fused into the render kernel as peer stores over NVLink.
G-GPU output must be byte-identical to 1-GPU output.

With fewer physical GPUs than a test wants, its partitions are mapped onto the GPUs that exist (device r % N_GPUS; rt_create accepts
repeated device ids, CUDA IPC works between processes on one device): the same partition / launch-thread / gather / per-device return
code runs, only the wire is local memory instead of NVLink. The one thing that needs real concurrency between ranks — the packed
gather's kernel-to-kernel flags in the multi-process test — still needs >= 2 GPUs."""
import os
import sys

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if lib.cuInit(0) != 0:
            return 0
        lib.cuDeviceGetCount(ctypes.byref(n))
        return n.value
    except OSError:
        return 0


N_GPUS = _n_gpus()
need1 = pytest.mark.skipif(N_GPUS < 1, reason="needs a GPU")
need2 = pytest.mark.skipif(N_GPUS < 2, reason="needs >= 2 GPUs (kernels of different ranks must run concurrently)")


def _devices(g):
    """g partitions on the GPUs that exist"""
    return [i % N_GPUS for i in range(g)]


@pytest.fixture(scope="module")
def rt(built):
    import rtb200
    return rtb200


@need1
@pytest.mark.parametrize("g", [2, 4, 8])
def test_in_library_multi_device_equals_single(rt, g):
    w, h = 1280, 720
    for sc, camkw, accel in ((scenes.default_scene(), dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), rt.RT_ACCEL_AUTO),
                             (scenes.config3_scene(), scenes.SCALED_CAMERA, rt.RT_ACCEL_LBVH)):
        cam = scenes.make_camera(width=w, height=h, **camkw)
        one = rt.Context([0]); one.set_scene(sc, accel)
        ref, _ = one.render(cam, w, h, 8)
        one.close()
        multi = rt.Context(_devices(g)); multi.set_scene(sc, accel)
        if g > N_GPUS:
            # partitions sharing a GPU: the packed gather (automatic from 8 devices on) makes rank 0's expand kernel wait for flags the
            # other partitions' kernels write — on one GPU those may not be resident at the same time. Plain / sparse gather there.
            multi.set_option(rt.RT_OPT_GATHER_MODE, 0)
        got, st = multi.render(cam, w, h, 8)
        assert np.array_equal(got, ref)
        cams = np.stack([scenes.make_camera(pos=(0.1 * i, 0.4, -1.0), yaw=0.05 * i, width=w, height=h) for i in range(3)])
        batch, _ = multi.render_batch(cams, w, h, 8, headless=False)
        multi.close()
        one = rt.Context([0]); one.set_scene(sc, accel)
        for i in range(3):
            assert np.array_equal(batch[i], one.render(cams[i], w, h, 8)[0])
        one.close()


@need1
@pytest.mark.parametrize("w,h,tile_rows", [(1280, 723, 8), (640, 97, 16), (96, 5, 8)])
def test_multi_device_host_output_paths(rt, w, h, tile_rows):
    """Host output from a multi-device context: (a) default — every device sends its own row tiles over its own PCIe link
    (strided 2-D copies, incl. a short last tile), (b) RT_OPT_HOST_VIA_GPU0 — gather on device 0, then copy. Both must equal the
    single-device frame, into pageable and into page-locked host memory."""
    g = 4 if N_GPUS != 2 else 2
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)
    one = rt.Context([0]); one.set_scene(sc)
    ref, _ = one.render(cam, w, h, 8)
    one.close()
    multi = rt.Context(_devices(g)); multi.set_scene(sc); multi.set_partition(0, 1, tile_rows)
    for via0 in (0, 1):
        multi.set_option(rt.RT_OPT_HOST_VIA_GPU0, via0)
        got, st = multi.render(cam, w, h, 8)
        assert np.array_equal(got, ref), via0
        pinned = np.full((h, w), 0x55555555, dtype=np.int32)
        multi.host_register(pinned)
        multi.render(cam, w, h, 8, out=pinned)
        multi.host_unregister(pinned)
        assert np.array_equal(pinned, ref), via0
        cams = np.stack([scenes.make_camera(pos=(0.1 * i, 0.4, -1.0), yaw=0.05 * i, width=w, height=h) for i in range(3)])
        batch, _ = multi.render_batch(cams, w, h, 8, headless=False)
        for i in range(3):
            one = rt.Context([0]); one.set_scene(sc)
            assert np.array_equal(batch[i], one.render(cams[i], w, h, 8)[0])
            one.close()
    multi.close()


@need1
def test_sparse_gather_leaves_no_stale_pixels(rt):
    """Gather on device 0 (RT_OPT_HOST_VIA_GPU0): devices other than 0 do not send the spans the frame gates prove black, device 0
    zero-fills them. Alternating a camera that sees only floor with one that sees mostly sky makes a missing fill visible as
    stale floor pixels."""
    g = 4 if N_GPUS != 2 else 2
    sc = scenes.default_scene()
    w, h = 1024, 600
    down = scenes.make_camera(pos=(0.0, 3.0, 2.0), pitch=1.3, width=w, height=h)       # floor everywhere
    up = scenes.make_camera(pos=(0.0, 0.5, 0.0), pitch=-0.6, width=w, height=h)        # mostly sky
    level = scenes.make_camera(width=w, height=h)
    one = rt.Context([0]); one.set_scene(sc)
    refs = [one.render(c, w, h, 8)[0].copy() for c in (down, up, level)]
    one.close()
    assert (refs[0] != 0).mean() > 0.95 and (refs[1] == 0).mean() > 0.5
    multi = rt.Context(_devices(g)); multi.set_scene(sc)
    multi.set_option(rt.RT_OPT_HOST_VIA_GPU0, 1)
    multi.set_option(rt.RT_OPT_SHARED_TARGET, 2)         # force the sparse gather (automatic only above 4 devices)
    for rep in range(2):
        for c, ref in zip((down, up, level), refs):
            got, _ = multi.render(c, w, h, 8)
            assert np.array_equal(got, ref), "%d pixels differ" % (got != ref).sum()
    multi.close()


def _ipc_worker(rank, world, w, h, tile_rows, q_handle, q_done, q_go, shared_target=0):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "uu-infogr-raytracer_b200"))
    import rtb200
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)
    ctx = rtb200.Context([rank % max(N_GPUS, 1)]); ctx.set_scene(sc); ctx.set_partition(rank, world, tile_rows)
    ctx.set_option(rtb200.RT_OPT_SHARED_TARGET, shared_target)
    if rank == 0:
        fb = ctx.dev_alloc(w * h * 4)
        ctx.dev_memset(fb, 0x5A, w * h * 4)          # poison: a pixel nobody writes cannot pass for black
        handle = ctx.ipc_export(fb)
        for _ in range(world - 1):
            q_handle.put(handle)
    else:
        fb = ctx.ipc_open(q_handle.get(timeout=60))
    ctx.render_device(cam[None], w, h, 8, 1, 0, fb)
    ctx.sync()
    if rank != 0:
        ctx.ipc_close(fb)
        q_done.put(rank)
        q_go.get(timeout=60)
    else:
        for _ in range(world - 1):
            q_done.get(timeout=60)
        got = ctx.dev_to_host(fb, w * h * 4).reshape(h, w)
        one = rtb200.Context([0]); one.set_scene(sc)
        ref, _ = one.render(cam, w, h, 8)
        ok = bool(np.array_equal(got, ref))
        for _ in range(world - 1):
            q_go.put(ok)
        one.close()
        ctx.dev_free(fb)
        q_done.put("ok" if ok else "MISMATCH")
    ctx.close()


@need1
@pytest.mark.parametrize("shared_target,w,h,tile_rows", [(0, 1000, 563, 8), (2, 1000, 563, 8), (2, 1283, 97, 3)])
def test_multi_process_ipc_peer_stores(built, shared_target, w, h, tile_rows):
    """One process per GPU (the torchrun shape): rank 1 stores its row tiles straight into rank 0's framebuffer (CUDA IPC).
    shared_target = 2 (forced; 1 = only above 4 ranks): sparse gather — rank 1 does not send the spans the frame gates prove black, rank 0 zero-fills them
    (RT_OPT_SHARED_TARGET); the frame must still equal the single-GPU frame, also with ragged spans (odd width, 3-row tiles)."""
    import torch.multiprocessing as mp
    world = 2
    mctx = mp.get_context("spawn")
    qh, qd, qg = mctx.Queue(), mctx.Queue(), mctx.Queue()
    procs = [mctx.Process(target=_ipc_worker, args=(r, world, w, h, tile_rows, qh, qd, qg, shared_target)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = []
    while not qd.empty():
        res.append(qd.get())
    assert "ok" in res, res


def _packed_worker(rank, world, w, h, mode, steps, q_handle, q_done, q_go):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "uu-infogr-raytracer_b200"))
    import rtb200
    sc = scenes.default_scene()
    F = 4
    cam_sets = [np.stack([scenes.make_camera(pos=(0.05 * i + 0.1 * s, 0.3 * (s % 2), -0.2 * i), yaw=0.02 * i - 0.1 * s, pitch=0.05 * s - 0.1,
                                             width=w, height=h) for i in range(F)]) for s in range(3)]
    ctx = rtb200.Context([rank]); ctx.set_scene(sc); ctx.set_partition(rank, world, 8)
    ctx.set_option(rtb200.RT_OPT_SHARED_TARGET, 1)
    ctx.set_option(rtb200.RT_OPT_GATHER_MODE, mode)
    nbytes = ctx.gather_bytes(w, h)
    if rank == 0:
        fb = ctx.dev_alloc(F * w * h * 4)
        ctx.dev_memset(fb, 0x5A, F * w * h * 4)
        ga = ctx.dev_alloc(nbytes)
        ctx.dev_memset(ga, 0, nbytes)
        hs = (ctx.ipc_export(fb), ctx.ipc_export(ga))
        for _ in range(world - 1):
            q_handle.put(hs)
    else:
        hs = q_handle.get(timeout=60)
        fb, ga = ctx.ipc_open(hs[0]), ctx.ipc_open(hs[1])
    ctx.gather_attach(ga, nbytes)
    q_done.put(("attached", rank))
    q_go.get(timeout=60)                      # everybody attached: go
    # `steps` launch groups back to back, NO host synchronisation in between: the kernels' own flags order producers and consumer
    for s in range(steps):
        ctx.render_device(cam_sets[s % 3], w, h, 8, 1, 0, fb)
    ctx.sync()
    active = ctx.get_info(rtb200.RT_INFO_GATHER_ACTIVE)
    timeouts = ctx.get_info(rtb200.RT_INFO_GATHER_TIMEOUTS)
    q_done.put(("rendered", rank, active, timeouts))
    ok = q_go.get(timeout=120)
    if rank == 0:
        got = ctx.dev_to_host(fb, F * w * h * 4).reshape(F, h, w)
        one = rtb200.Context([0]); one.set_scene(sc)
        last = cam_sets[(steps - 1) % 3]
        same = all(np.array_equal(got[i], one.render(last[i], w, h, 8)[0]) for i in range(F))
        one.close()
        q_done.put(("checked", bool(same)))
        q_go.get(timeout=60)
        ctx.gather_attach(None)
        ctx.dev_free(fb); ctx.dev_free(ga)
    else:
        ctx.gather_attach(None)
        ctx.ipc_close(fb); ctx.ipc_close(ga)
    ctx.close()


@need2
@pytest.mark.parametrize("mode,w,h", [(2, 1280, 720), (1, 1024, 600), (2, 3840, 2160)])
def test_multi_process_packed_gather(built, mode, w, h):
    """One process per GPU, the packed gather over real NVLink: every rank issues 7 launch groups of 4 frames back to back without
    any host synchronisation; producers (ranks != 0) and the consumer (rank 0's expand pass) are ordered only by the flags the kernels
    write into rank 0's memory. The last group's frames on rank 0 must equal single-GPU frames; no spin may time out."""
    import torch.multiprocessing as mp
    world = min(N_GPUS, 4)
    mctx = mp.get_context("spawn")
    qh, qd, qg = mctx.Queue(), mctx.Queue(), mctx.Queue()
    procs = [mctx.Process(target=_packed_worker, args=(r, world, w, h, mode, 7, qh, qd, qg)) for r in range(world)]
    for p in procs: p.start()
    try:
        for _ in range(world): assert qd.get(timeout=180)[0] == "attached"
        for _ in range(world): qg.put(True)
        rendered = [qd.get(timeout=300) for _ in range(world)]
        assert all(r[0] == "rendered" for r in rendered), rendered
        assert all(r[2] == 1 for r in rendered), "packed gather not active: %r" % (rendered,)
        assert all(r[3] == 0 for r in rendered), "spin time-outs: %r" % (rendered,)
        for _ in range(world): qg.put(True)
        checked = qd.get(timeout=300)
        assert checked == ("checked", True), checked
        qg.put(True)
    finally:
        for p in procs: p.join(timeout=120)
        for p in procs:
            if p.is_alive(): p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
