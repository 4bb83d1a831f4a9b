"""GPU parity of the code paths the driver's 1-GPU run can see (VERDICT r01 "parity holes"):
  * the SHIPPED tiny-scene kernel instantiation (exact counts, packed fp32, fast division, frame gates) certified by chain hashes,
  * the multi-GPU partition + sparse-gather logic with the ranks emulated as several single-device contexts on device 0,
  * BASELINE configs[4] at its stated size (7680x4320 x 16 spp) on a pixel subset, configs[3] through a world-4 partition,
  * the sparse device->host return, partitioned host returns into one shared frame, the zero-copy display entry points,
  * the ADVICE r01 fixes (non-finite spheres on the LBVH path, range overflow in rt_update_spheres, LBVH launches on two streams).
Everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest

import oracle_lib as O
import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built):
    import rtb200
    return rtb200


def _default_with(ns, nl):
    """The reference scene cut down / extended to ns spheres and nl lights, one plane (shapes the exact-count kernels exist for)."""
    sc = scenes.default_scene()
    sph = sc.spheres
    if ns <= len(sph):
        sph = sph[:ns]
    else:
        extra = sph[0:1].copy(); extra[0, 0:3] = np.float32([0.0, 0.5, 7.0])
        sph = np.concatenate([sph, extra])[:ns]
    lig = sc.lights
    if nl <= len(lig):
        lig = lig[:nl]
    else:
        more = np.float32([[-4.0, 6.0, 2.0, 1.0], [5.0, 3.0, -1.0, 1.0]])
        lig = np.concatenate([lig, more])[:nl]
    return scenes.Scene(np.ascontiguousarray(sph), sc.planes, np.ascontiguousarray(lig), sc.ambient)


CAMS = [dict(), dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), dict(pos=(-2.0, 2.5, 3.0), yaw=-0.4, pitch=0.5),
        dict(pos=(0.0, 3.0, 2.0), pitch=1.3), dict(pos=(0.0, 0.5, 0.0), pitch=-0.6), dict(pos=(0.0, 0.0, 6.0), yaw=3.1)]


@pytest.mark.parametrize("ns,nl", [(3, 2), (0, 0), (1, 1), (2, 2), (4, 4), (3, 1)])
def test_shipped_kernel_chain_hashes(rt, ns, nl):
    """RT_OPT_DEBUG_SHIPPED: rt_render_debug runs k_debug_tiny_prod = the production instantiation + gates with the events-only
    policy. Chain hash (every hit id, t bit pattern, shadow result), primary AOVs, ray counters and pixels must equal the oracle's —
    and the pixels those of rt_render. (3, 1) has no exact debug instantiation: run-time counts, still gated."""
    sc = _default_with(ns, nl)
    ctx = rt.Context([0]); ctx.set_scene(sc)
    ctx.set_option(rt.RT_OPT_DEBUG_SHIPPED, 1)
    for camkw, (w, h), depth in zip(CAMS, [(640, 360), (1280, 720), (333, 187), (512, 288), (640, 360), (250, 131)], [32, 8, 8, 3, 8, 32]):
        cam = scenes.make_camera(width=w, height=h, **camkw)
        a = O.render(sc, cam, w, h, depth, want_hash=True, want_aov=True)
        d = ctx.render_debug(cam, w, h, depth)
        assert np.array_equal(d["hash"], a["hash"]), "%d chain hashes differ" % int((d["hash"] != a["hash"]).sum())
        assert np.array_equal(d["aov_id"], a["aov_id"])
        assert np.array_equal(d["aov_t"].view(np.uint32), a["aov_t"].view(np.uint32))
        assert np.array_equal(d["pixels"], a["pixels"])
        for k in ("primary", "shadow", "secondary"):
            assert d["counters"][k] == a["counters"][k], k
        px, _ = ctx.render(cam, w, h, depth)
        assert np.array_equal(px, d["pixels"])
    ctx.close()


def test_shipped_kernel_chain_hashes_4k(rt):
    """The bench workload itself (BASELINE configs[1]): 3840x2160, cap 8, production kernel, hashes vs the oracle."""
    sc = scenes.default_scene()
    w, h = 3840, 2160
    cam = scenes.make_camera(width=w, height=h)
    ctx = rt.Context([0]); ctx.set_scene(sc)
    ctx.set_option(rt.RT_OPT_DEBUG_SHIPPED, 1)
    d = ctx.render_debug(cam, w, h, 8)
    a = O.render(sc, cam, w, h, 8, want_hash=True, want_aov=True)
    assert np.array_equal(d["hash"], a["hash"])
    assert np.array_equal(d["aov_id"], a["aov_id"])
    assert np.array_equal(d["aov_t"].view(np.uint32), a["aov_t"].view(np.uint32))
    assert np.array_equal(d["pixels"], a["pixels"])
    assert [d["counters"][k] for k in ("primary", "shadow", "secondary")] == [a["counters"][k] for k in ("primary", "shadow", "secondary")]
    ctx.close()


# ---- multi-GPU logic on ONE device: ranks = several single-device contexts storing into one poisoned buffer ---------------------
@pytest.mark.parametrize("world,tile_rows,w,h", [(2, 8, 1024, 600), (4, 3, 1000, 563), (8, 16, 1283, 397), (8, 1, 250, 131), (4, 8, 1280, 720)])
@pytest.mark.parametrize("shared_target", [0, 2])
def test_emulated_ranks_sparse_gather(rt, world, tile_rows, w, h, shared_target):
    """rt_set_partition(r, world) + RT_OPT_SHARED_TARGET = 2 (forced sparse gather): ranks != 0 do not store the spans the gates
    prove black, rank 0 zero-fills exactly those (k_fill_black). Three alternating cameras (floor only / mostly sky / level) into
    the SAME poisoned framebuffer: a span nobody writes shows up as poison or as stale pixels of the previous camera."""
    sc = scenes.default_scene()
    cams = [scenes.make_camera(pos=(0.0, 3.0, 2.0), pitch=1.3, width=w, height=h),
            scenes.make_camera(pos=(0.0, 0.5, 0.0), pitch=-0.6, width=w, height=h),
            scenes.make_camera(width=w, height=h)]
    base = rt.Context([0]); base.set_scene(sc)
    refs = [base.render(c, w, h, 8)[0].copy() for c in cams]
    assert (refs[0] != 0).mean() > 0.95 and (refs[1] == 0).mean() > 0.5
    fb = base.dev_alloc(w * h * 4)
    base.dev_memset(fb, 0x5A, w * h * 4)
    ranks = []
    for r in range(world):
        c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, world, tile_rows)
        c.set_option(rt.RT_OPT_SHARED_TARGET, shared_target)
        ranks.append(c)
    for rep in range(2):
        for cam, ref in zip(cams, refs):
            for r in list(range(1, world)) + [0]:             # the peers first, rank 0 (the sink) last
                ranks[r].render_device(cam[None], w, h, 8, 1, 0, fb)
                ranks[r].sync()
            got = base.dev_to_host(fb, w * h * 4).reshape(h, w)
            assert np.array_equal(got, ref), "%d pixels differ (world %d)" % (int((got != ref).sum()), world)
    for c in ranks:
        c.close()
    base.dev_free(fb); base.close()


def test_emulated_ranks_batch_of_distinct_cameras(rt):
    """16 frames per launch (the bench shape) with 16 DIFFERENT cameras through a world-8 partition with the sparse gather."""
    sc = scenes.default_scene()
    w, h, world, F = 640, 360, 8, 16
    cams = np.stack([scenes.make_camera(pos=(0.02 * i, 0.01 * i, -0.03 * i), yaw=0.004 * i, pitch=0.002 * i - 0.01, width=w, height=h) for i in range(F)])
    base = rt.Context([0]); base.set_scene(sc)
    fb = base.dev_alloc(F * w * h * 4)
    base.dev_memset(fb, 0x5A, F * w * h * 4)
    for r in list(range(1, world)) + [0]:
        c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, world, 8)
        c.set_option(rt.RT_OPT_SHARED_TARGET, 2)
        c.render_device(cams, w, h, 8, 1, 0, fb); c.sync(); c.close()
    got = base.dev_to_host(fb, F * w * h * 4).reshape(F, h, w)
    for i in range(F):
        assert np.array_equal(got[i], base.render(cams[i], w, h, 8)[0]), i
    base.dev_free(fb); base.close()


def test_config4_100k_through_world4_partition(rt):
    """BASELINE configs[3] (100 k spheres, LBVH, 4K) through a world-4 row-tile partition: union of the four ranks' tiles == the
    unpartitioned frame == the oracle's brute force on the fixed 65,536-pixel subset (seed 7)."""
    sc = scenes.config4_scene()
    w, h = 3840, 2160
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    base = rt.Context([0]); base.set_scene(sc)
    assert base.get_info(rt.RT_INFO_SCENE_PATH) == 3
    full, _ = base.render(cam, w, h, 8)
    fb = base.dev_alloc(w * h * 4)
    base.dev_memset(fb, 0x5A, w * h * 4)
    for r in range(4):
        c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, 4, 8)
        c.render_device(cam[None], w, h, 8, 1, 0, fb); c.sync(); c.close()
    got = base.dev_to_host(fb, w * h * 4).reshape(h, w)
    assert np.array_equal(got, full)
    idx = np.random.default_rng(7).choice(w * h, 65536, replace=False).astype(np.int32)
    ref = O.render(sc, cam, w, h, 8, subset=idx)["pixels"]
    assert np.array_equal(got.reshape(-1)[idx], ref)
    base.dev_free(fb); base.close()


def test_config5_8k_16spp_subset_vs_oracle(rt):
    """BASELINE configs[4] at its stated size: default scene, 7680x4320, 16 jittered samples per pixel, cap 8 — a fixed pixel subset
    (seed 11) against the oracle, single device and through a world-8 partition (the jitter stream is keyed by the pixel index, so
    the partition cannot change it)."""
    sc = scenes.default_scene()
    w, h, spp, seed = 7680, 4320, 16, 5
    cam = scenes.make_camera(width=w, height=h)
    idx = np.random.default_rng(11).choice(w * h, 32768, replace=False).astype(np.int32)
    ref = O.render(sc, cam, w, h, 8, spp=spp, seed=seed, subset=idx)["pixels"]
    base = rt.Context([0]); base.set_scene(sc)
    px, _ = base.render(cam, w, h, 8, spp=spp, seed=seed)
    assert np.array_equal(px.reshape(-1)[idx], ref)
    fb = base.dev_alloc(w * h * 4)
    base.dev_memset(fb, 0x5A, w * h * 4)
    for r in range(8):
        c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, 8, 8)
        c.render_device(cam[None], w, h, 8, spp, seed, fb); c.sync(); c.close()
    got = base.dev_to_host(fb, w * h * 4).reshape(h, w)
    assert np.array_equal(got, px)
    base.dev_free(fb); base.close()


# ---- sparse device -> host return ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h", [(3840, 2160), (1280, 720), (1000, 563), (37, 23)])
def test_sparse_d2h_equals_dense(rt, w, h):
    """RT_OPT_SPARSE_D2H on (default) / off: identical frames into poisoned pageable and page-locked buffers, for cameras with a lot /
    nothing / everything to skip; fewer bytes cross PCIe when there is sky."""
    sc = scenes.default_scene()
    ctx = rt.Context([0]); ctx.set_scene(sc)
    pinned = np.empty((h, w), np.int32)
    ctx.host_register(pinned)
    for camkw in CAMS:
        cam = scenes.make_camera(width=w, height=h, **camkw)
        ctx.set_option(rt.RT_OPT_SPARSE_D2H, 0)
        ref, _ = ctx.render(cam, w, h, 8)
        assert ctx.get_info(rt.RT_INFO_LAST_D2H_BYTES) == w * h * 4
        ctx.set_option(rt.RT_OPT_SPARSE_D2H, 1)
        # pageable memory: copy engine; page-locked memory: the kernel's own stores over PCIe (zero copy), or the copy engine
        for buf, zero_copy in ((np.full((h, w), 0x5A5A5A5A, np.int32), 1), (pinned, 1), (pinned, 0)):
            ctx.set_option(rt.RT_OPT_HOST_ZERO_COPY, 2 * zero_copy)
            buf[...] = 0x5A5A5A5A
            ctx.render(cam, w, h, 8, out=buf)
            assert np.array_equal(buf, ref), "%d pixels differ (zero_copy %d)" % (int((buf != ref).sum()), zero_copy)
            copied = ctx.get_info(rt.RT_INFO_LAST_D2H_BYTES)
            assert copied <= w * h * 4
            if camkw == {} and w >= 1000:
                assert copied < 0.72 * w * h * 4, copied / (w * h * 4)      # a third of the default frame is proven black
        ctx.set_option(rt.RT_OPT_HOST_ZERO_COPY, 2)
        ctx.set_option(rt.RT_OPT_SPARSE_D2H, 0)                             # zero copy without the sparse return: every pixel is stored
        pinned[...] = 0x5A5A5A5A
        ctx.render(cam, w, h, 8, out=pinned)
        assert np.array_equal(pinned, ref) and ctx.get_info(rt.RT_INFO_LAST_D2H_BYTES) == w * h * 4
        ctx.set_option(rt.RT_OPT_SPARSE_D2H, 1); ctx.set_option(rt.RT_OPT_HOST_ZERO_COPY, 1)
    # the caller's promise that the buffer is already zero (the reference's screen.Clear(0), RayTracer.cs:890): no fill by the library
    ctx.set_option(rt.RT_OPT_HOST_PRECLEARED, 1)
    cam = scenes.make_camera(width=w, height=h)
    pinned[...] = 0
    ctx.render(cam, w, h, 8, out=pinned)
    ctx.set_option(rt.RT_OPT_HOST_PRECLEARED, 0); ctx.set_option(rt.RT_OPT_SPARSE_D2H, 0)
    assert np.array_equal(pinned, ctx.render(cam, w, h, 8)[0])
    ctx.host_unregister(pinned)
    ctx.close()


def test_sparse_d2h_batch_and_other_paths(rt):
    """Batches with distinct cameras; scenes without gates (LBVH / staged / supersampled) take the dense copy and stay correct."""
    w, h = 640, 360
    sc = scenes.default_scene()
    ctx = rt.Context([0]); ctx.set_scene(sc)
    cams = np.stack([scenes.make_camera(width=w, height=h, **kw) for kw in CAMS])
    batch, _ = ctx.render_batch(cams, w, h, 8, headless=False)
    ctx.set_option(rt.RT_OPT_SPARSE_D2H, 0)
    for i in range(len(cams)):
        assert np.array_equal(batch[i], ctx.render(cams[i], w, h, 8)[0]), i
    ctx.set_option(rt.RT_OPT_SPARSE_D2H, 1)
    ss, _ = ctx.render(cams[0], w, h, 8, spp=4, seed=3)
    assert np.array_equal(ss, O.render(sc, cams[0], w, h, 8, spp=4, seed=3)["pixels"])
    ctx.close()
    sc3 = scenes.config3_scene()
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    for accel in (rt.RT_ACCEL_LBVH, rt.RT_ACCEL_BRUTE):
        c = rt.Context([0]); c.set_scene(sc3, accel)
        px, _ = c.render(cam, w, h, 8)
        assert c.get_info(rt.RT_INFO_LAST_D2H_BYTES) == w * h * 4
        assert np.array_equal(px, c.render_debug(cam, w, h, 8)["pixels"])
        c.close()


@pytest.mark.parametrize("world,tile_rows,w,h", [(2, 8, 1280, 720), (4, 3, 1000, 563), (8, 8, 3840, 2160), (8, 16, 333, 187)])
def test_partitioned_contexts_fill_one_host_frame(rt, world, tile_rows, w, h):
    """The e2e shape of `bench.py --gpus N`: every rank's rt_render returns ITS OWN row tiles into one shared host frame (here: one
    numpy buffer, the ranks one after the other); together they must write every pixel — poison must not survive, with and without
    the sparse return."""
    sc = scenes.default_scene()
    base = rt.Context([0]); base.set_scene(sc)
    for camkw in (dict(), dict(pos=(0.0, 0.5, 0.0), pitch=-0.6), dict(pos=(0.0, 3.0, 2.0), pitch=1.3)):
        cam = scenes.make_camera(width=w, height=h, **camkw)
        ref, _ = base.render(cam, w, h, 8)
        for sparse, pin in ((1, 0), (0, 0), (1, 1), (0, 1)):
            frame = np.full((h, w), 0x5A5A5A5A, np.int32)
            if pin:
                base.host_register(frame)         # page-locked: every rank's kernel stores its tiles straight into the frame
            total = 0
            for r in range(world):
                c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, world, tile_rows)
                c.set_option(rt.RT_OPT_SPARSE_D2H, sparse)
                c.set_option(rt.RT_OPT_HOST_ZERO_COPY, 2 if pin else 0)
                c.render(cam, w, h, 8, out=frame)
                total += c.get_info(rt.RT_INFO_LAST_D2H_BYTES)
                c.close()
            if pin:
                base.host_unregister(frame)
            assert np.array_equal(frame, ref), "%d pixels differ (sparse %d, pinned %d)" % (int((frame != ref).sum()), sparse, pin)
            assert total <= w * h * 4
            if not sparse:
                assert total == w * h * 4
    base.close()


# ---- zero-copy display entry points -----------------------------------------------------------------------------
def test_render_mapped_into_a_display_owned_buffer(rt):
    """rt_render_mapped: the frame lands in a device buffer the caller owns (stand-in for a CUDA-mapped GL pixel-unpack buffer,
    template.cs:81,188-193) without touching host memory, identical to rt_render; argument errors are reported, not crashed on."""
    sc = scenes.default_scene()
    w, h = 1280, 720
    ctx = rt.Context([0]); ctx.set_scene(sc)
    pbo = ctx.dev_alloc(w * h * 4 + 256)
    for camkw in CAMS[:3]:
        cam = scenes.make_camera(width=w, height=h, **camkw)
        ctx.dev_memset(pbo, 0x5A, w * h * 4 + 256)
        st = ctx.render_mapped(cam, w, h, 8, pbo, w * h * 4 + 256)
        assert st.kernel_ms > 0 and st.d2h_ms == 0
        got = ctx.dev_to_host(pbo, w * h * 4 + 256)
        assert np.array_equal(got[: w * h].reshape(h, w), ctx.render(cam, w, h, 8)[0])
        assert (got[w * h:] == 0x5A5A5A5A).all()                    # nothing written past the frame
    cam = scenes.make_camera(width=w, height=h)
    with pytest.raises(rt.RtError) as e:
        ctx.render_mapped(cam, w, h, 8, pbo, w * h * 4 - 4)         # buffer too small
    assert e.value.code == -1
    host = np.zeros(w * h, np.int32)
    with pytest.raises(rt.RtError) as e:
        ctx.render_mapped(cam, w, h, 8, host.ctypes.data, w * h * 4)    # a host pointer is not a mapped device buffer
    assert e.value.code == -1
    with pytest.raises(rt.RtError) as e:
        ctx.gl_register_buffer(1)                                   # no OpenGL context on this box: a clean error
    assert e.value.code == -2 and "OpenGL" in str(e.value)
    with pytest.raises(rt.RtError):
        ctx.render_gl(cam, w, h, 8, 0xDEAD0000)                     # unknown resource handle
    ctx.dev_free(pbo); ctx.close()


# ---- ADVICE r01 ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [56, 72, 200])
def test_lbvh_path_with_non_finite_spheres(rt, n):
    """A NaN / Inf centre or radiusSquared on the LBVH path (shadow bins included): such a sphere can never be hit in the reference
    (every comparison of RayTracer.cs:622-635 fails), so the frame equals the oracle's and nothing indexes out of bounds."""
    sc = scenes.small_random_scene(n, 5)
    sph = sc.spheres.copy()
    sph[3, 0] = np.nan; sph[7, 17] = np.inf; sph[11, 2] = -np.inf; sph[n - 1, 17] = np.nan; sph[20, 1] = np.inf
    sc = scenes.Scene(sph, sc.planes, sc.lights, sc.ambient)
    w, h = 200, 120
    cam = scenes.make_camera(pos=(0, 1.5, -4.0), pitch=0.1, width=w, height=h)
    ref = O.render(sc, cam, w, h, 8)["pixels"]
    for accel in (rt.RT_ACCEL_LBVH, rt.RT_ACCEL_AUTO, rt.RT_ACCEL_BRUTE):
        c = rt.Context([0]); c.set_scene(sc, accel)
        assert np.array_equal(c.render(cam, w, h, 8)[0], ref)
        moved = sph[:4].copy(); moved[0, 0:3] = np.float32([np.nan, np.inf, 0.0])
        c.update_spheres(moved, 0)
        sc2 = scenes.Scene(np.concatenate([moved, sph[4:]]), sc.planes, sc.lights, sc.ambient)
        assert np.array_equal(c.render(cam, w, h, 8)[0], O.render(sc2, cam, w, h, 8)["pixels"])
        c.close()


def test_update_spheres_range_check_does_not_overflow(rt):
    sc = scenes.default_scene()
    c = rt.Context([0]); c.set_scene(sc)
    one = sc.spheres[:1].copy()
    for first in (2**31 - 1, 3, -1):
        with pytest.raises(rt.RtError) as e:
            c.update_spheres(one, first)
        assert e.value.code == -1
    c.update_spheres(one, 2)
    c.close()


def test_lbvh_launches_on_two_streams_are_ordered(rt):
    """One camera-inflated node copy per device: rt_render_device on two user streams with different cameras must not let the
    second refit overwrite the boxes the first launch is still traversing."""
    import torch
    sc = scenes.config4_scene(20000)
    w, h = 1280, 720
    cams = [scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA), scenes.make_camera(pos=(40.0, 6.0, 20.0), yaw=0.9, pitch=0.3, width=w, height=h)]
    c = rt.Context([0]); c.set_scene(sc, rt.RT_ACCEL_LBVH)
    refs = [c.render(cam, w, h, 8)[0].copy() for cam in cams]
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs = [torch.empty((h, w), dtype=torch.int32, device="cuda") for _ in range(2)]
    for rep in range(4):
        for i in range(2):
            c.render_device(cams[i][None], w, h, 8, 1, 0, bufs[i].data_ptr(), s[i].cuda_stream)
        sph = sc.spheres[:64].copy()
        torch.cuda.synchronize()
        for i in range(2):
            assert np.array_equal(bufs[i].cpu().numpy(), refs[i]), (rep, i)
        if rep == 1:          # a scene update between frames that were launched on user streams
            c.update_spheres(sph, 0)
    c.close()


# ---- packed multi-GPU gather (csrc/rt_gather.cuh) with the ranks emulated on ONE device ------------------------------------------
def _packed_ranks(rt, sc, world, tile_rows, w, h, mode, sink, peer, shared_target=1):
    base = rt.Context([0]); base.set_scene(sc)
    nbytes = base.gather_bytes(w, h)
    area = base.dev_alloc(nbytes)
    base.dev_memset(area, 0, nbytes)
    ranks = []
    for r in range(world):
        c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, world, tile_rows)
        c.set_option(rt.RT_OPT_SHARED_TARGET, shared_target)
        c.set_option(rt.RT_OPT_GATHER_MODE, mode); c.set_option(rt.RT_OPT_SINK_TILES, sink); c.set_option(rt.RT_OPT_PEER_TILES, peer)
        c.gather_attach(area, nbytes)
        ranks.append(c)
    return base, area, ranks


@pytest.mark.parametrize("world,tile_rows,w,h,mode,sink,peer", [
    (2, 8, 1024, 600, 2, 1, 1), (4, 8, 1280, 720, 2, 0, 0), (8, 8, 1280, 720, 2, 0, 0), (8, 8, 1280, 720, 1, 0, 0), (8, 3, 640, 97, 2, 1, 2),
    (4, 16, 1152, 333, 1, 2, 3), (8, 1, 256, 131, 2, 3, 5), (3, 8, 1280, 720, 2, 0, 4), (8, 8, 3840, 2160, 2, 0, 0), (4, 8, 3840, 2160, 2, 0, 0),
    (4, 8, 3840, 2160, 1, 0, 0), (4, 8, 1280, 724, 1, 0, 0), (5, 24, 896, 500, 1, 1, 2),
    (4, 8, 1280, 724, 2, 0, 0), (8, 16, 1152, 333, 2, 1, 2)])   # tile_rows % 8 == 0: 2-D pixel blocks (flag bytes assembled per CTA); else row strips
def test_packed_gather_emulated_ranks(rt, world, tile_rows, w, h, mode, sink, peer):
    """Ranks != 0 write the wire format (nothing / grey bytes / RGB24 + flag bytes) into the planes of the gather area, rank 0 renders
    its (smaller, weighted) share and expands the rest; flags and epochs as between real GPUs. Four alternating cameras, two rounds
    (eight epochs) into one poisoned framebuffer must reproduce the single-context frames byte for byte, with no spin time-outs.
    sink = 0 with peer > 0: rank 0 renders nothing and only expands."""
    sc = scenes.default_scene()
    cams = [scenes.make_camera(pos=(0.0, 3.0, 2.0), pitch=1.3, width=w, height=h),
            scenes.make_camera(pos=(0.0, 0.5, 0.0), pitch=-0.6, width=w, height=h),
            scenes.make_camera(width=w, height=h), scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)]
    base, area, ranks = _packed_ranks(rt, sc, world, tile_rows, w, h, mode, sink, peer)
    refs = [base.render(c, w, h, 8)[0].copy() for c in cams]
    fb = base.dev_alloc(w * h * 4)
    base.dev_memset(fb, 0x5A, w * h * 4)
    for rep in range(2):
        for cam, ref in zip(cams, refs):
            for r in list(range(1, world)) + [0]:             # the peers first: on ONE device rank 0's expand pass would spin on them
                ranks[r].render_device(cam[None], w, h, 8, 1, 0, fb)
                ranks[r].sync()
            assert all(c.get_info(rt.RT_INFO_GATHER_ACTIVE) == 1 for c in ranks)
            got = base.dev_to_host(fb, w * h * 4).reshape(h, w)
            assert np.array_equal(got, ref), "%d pixels differ (world %d, mode %d)" % (int((got != ref).sum()), world, mode)
    assert ranks[0].get_info(rt.RT_INFO_GATHER_TIMEOUTS) == 0
    for c in ranks:
        c.close()
    base.dev_free(fb); base.dev_free(area); base.close()


def test_packed_gather_batches_colours_and_fallbacks(rt):
    """16 distinct cameras per launch (the bench shape) through world 8; a scene with coloured planes and every material class
    (run-time-count kernels, gates that never skip); and frames the packed gather does not apply to (width not a multiple of 128,
    supersampling) must fall back to the plain gather on every rank and stay correct."""
    w, h, world, F = 640, 360, 8, 16
    sc = scenes.default_scene()
    cams = np.stack([scenes.make_camera(pos=(0.02 * i, 0.01 * i, -0.03 * i), yaw=0.004 * i, pitch=0.002 * i - 0.01, width=w, height=h) for i in range(F)])
    base, area, ranks = _packed_ranks(rt, sc, world, 8, w, h, 2, 0, 0)
    fb = base.dev_alloc(F * w * h * 4)
    base.dev_memset(fb, 0x5A, F * w * h * 4)
    for rep in range(3):
        for r in list(range(1, world)) + [0]:
            ranks[r].render_device(cams, w, h, 8, 1, 0, fb); ranks[r].sync()
        got = base.dev_to_host(fb, F * w * h * 4).reshape(F, h, w)
        for i in range(F):
            assert np.array_equal(got[i], base.render(cams[i], w, h, 8)[0]), (rep, i)
    # not applicable: odd width / supersampled -> plain gather, still exact
    for (w2, h2, spp) in ((1000, 563, 1), (640, 360, 4)):
        cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w2, height=h2)
        base.dev_memset(fb, 0x5A, w2 * h2 * 4)
        for r in list(range(1, world)) + [0]:
            ranks[r].render_device(cam[None], w2, h2, 8, spp, 3, fb); ranks[r].sync()
            assert ranks[r].get_info(rt.RT_INFO_GATHER_ACTIVE) == 0
        assert np.array_equal(base.dev_to_host(fb, w2 * h2 * 4).reshape(h2, w2), base.render(cam, w2, h2, 8, spp=spp, seed=3)[0])
    assert ranks[0].get_info(rt.RT_INFO_GATHER_TIMEOUTS) == 0
    for c in ranks:
        c.close()
    base.dev_free(fb); base.dev_free(area); base.close()
    # coloured planes, 5 spheres of every material class
    sc2 = scenes.small_random_scene(5, 3)
    w, h = 768, 432
    cam = scenes.make_camera(pos=(0, 1.5, -4.0), pitch=0.1, width=w, height=h)
    base, area, ranks = _packed_ranks(rt, sc2, 4, 8, w, h, 2, 0, 0)
    fb = base.dev_alloc(w * h * 4)
    base.dev_memset(fb, 0x5A, w * h * 4)
    for rep in range(2):
        for r in (1, 2, 3, 0):
            ranks[r].render_device(cam[None], w, h, 8, 1, 0, fb); ranks[r].sync()
        assert np.array_equal(base.dev_to_host(fb, w * h * 4).reshape(h, w), O.render(sc2, cam, w, h, 8)["pixels"])
    assert ranks[0].get_info(rt.RT_INFO_GATHER_TIMEOUTS) == 0
    for c in ranks:
        c.close()
    base.dev_free(fb); base.dev_free(area); base.close()
