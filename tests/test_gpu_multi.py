"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): interleaved row tiles across GPUs with the gather fused into
the render kernel as peer stores over NVLink. G-GPU output must be byte-identical to 1-GPU output."""
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
need2 = pytest.mark.skipif(N_GPUS < 2, reason="needs >= 2 GPUs")


@pytest.fixture(scope="module")
def rt(built):
    import rtb200
    return rtb200


@need2
@pytest.mark.parametrize("g", [2, 4, 8])
def test_in_library_multi_device_equals_single(rt, g):
    if g > N_GPUS:
        pytest.skip("only %d GPUs" % N_GPUS)
    w, h = 1280, 720
    for sc, camkw, accel in ((scenes.default_scene(), dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), rt.RT_ACCEL_AUTO),
                             (scenes.config3_scene(), scenes.SCALED_CAMERA, rt.RT_ACCEL_LBVH)):
        cam = scenes.make_camera(width=w, height=h, **camkw)
        one = rt.Context([0]); one.set_scene(sc, accel)
        ref, _ = one.render(cam, w, h, 8)
        one.close()
        multi = rt.Context(list(range(g))); multi.set_scene(sc, accel)
        got, st = multi.render(cam, w, h, 8)
        assert np.array_equal(got, ref)
        cams = np.stack([scenes.make_camera(pos=(0.1 * i, 0.4, -1.0), yaw=0.05 * i, width=w, height=h) for i in range(3)])
        batch, _ = multi.render_batch(cams, w, h, 8, headless=False)
        multi.close()
        one = rt.Context([0]); one.set_scene(sc, accel)
        for i in range(3):
            assert np.array_equal(batch[i], one.render(cams[i], w, h, 8)[0])
        one.close()


@need2
@pytest.mark.parametrize("w,h,tile_rows", [(1280, 723, 8), (640, 97, 16), (96, 5, 8)])
def test_multi_device_host_output_paths(rt, w, h, tile_rows):
    """Host output from a multi-device context: (a) default — every device sends its own row tiles over its own PCIe link
    (strided 2-D copies, incl. a short last tile), (b) RT_OPT_HOST_VIA_GPU0 — gather on device 0, then copy. Both must equal the
    single-device frame, into pageable and into page-locked host memory."""
    g = min(N_GPUS, 4) if N_GPUS >= 4 else 2
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)
    one = rt.Context([0]); one.set_scene(sc)
    ref, _ = one.render(cam, w, h, 8)
    one.close()
    multi = rt.Context(list(range(g))); multi.set_scene(sc); multi.set_partition(0, 1, tile_rows)
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


def _ipc_worker(rank, world, w, h, tile_rows, q_handle, q_done, q_go):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "uu-infogr-raytracer_b200"))
    import rtb200
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)
    ctx = rtb200.Context([rank]); ctx.set_scene(sc); ctx.set_partition(rank, world, tile_rows)
    if rank == 0:
        fb = ctx.dev_alloc(w * h * 4)
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


@need2
def test_multi_process_ipc_peer_stores(built):
    """One process per GPU (the torchrun shape): rank 1 stores its row tiles straight into rank 0's framebuffer (CUDA IPC)."""
    import torch.multiprocessing as mp
    world = 2
    mctx = mp.get_context("spawn")
    qh, qd, qg = mctx.Queue(), mctx.Queue(), mctx.Queue()
    procs = [mctx.Process(target=_ipc_worker, args=(r, world, 1000, 563, 8, qh, qd, qg)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = []
    while not qd.empty():
        res.append(qd.get())
    assert "ok" in res, res
