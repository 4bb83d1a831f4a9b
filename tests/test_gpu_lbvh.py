"""GPU tests of the scaled-scene paths: shared-memory-staged brute force and the LBVH must both equal the oracle, and each
other, bit for bit (hit index, t bits, image) — BASELINE.json configs[2] and configs[3]."""
import numpy as np
import pytest

import oracle_lib as O
import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built):
    import rtb200
    return rtb200


def mk(c, r):
    return scenes.sphere(c, r, scenes.mat_diffuse((1, 1, 1)))


def _scene_of(spheres):
    d = scenes.default_scene()
    return scenes.Scene(spheres, d.planes, d.lights, d.ambient)


def _check_queries(rt, spheres, rays):
    ctx = rt.Context([0])
    ctx.set_scene(_scene_of(spheres), rt.RT_ACCEL_LBVH)
    for kind in (0, 1, 2):
        oi, ot = O.query_spheres(spheres, rays, kind)
        for accel in (rt.RT_ACCEL_BRUTE, rt.RT_ACCEL_LBVH):
            gi, gt = ctx.query_spheres(rays, kind, accel)
            assert np.array_equal(gi, oi), "kind %d accel %d: %d id mismatches" % (kind, accel, (gi != oi).sum())
            assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32)), "kind %d accel %d: t bits" % (kind, accel)
    ctx.close()


def test_queries_random_and_chains(rt):
    rng = np.random.default_rng(11)
    sph = []
    for _ in range(60):
        c = rng.uniform(-5, 5, 3); c[2] += 12
        for _ in range(rng.integers(2, 9)):
            sph.append(mk(c + rng.normal(size=3) * 0.004, 1.0 + rng.normal() * 0.003))
    sph = np.stack(sph)[rng.permutation(len(sph))]
    n = 50000
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    tgt = sph[rng.integers(0, len(sph), n), :3] + rng.normal(size=(n, 3)).astype(np.float32) * 0.7
    d = (tgt - o).astype(np.float32) * rng.uniform(0.05, 20, (n, 1)).astype(np.float32)
    _check_queries(rt, sph, np.concatenate([o, d], 1))


def test_queries_far_spheres_fp_noise(rt):
    rng = np.random.default_rng(11)
    n = 100000
    sph = np.stack([mk((rng.uniform(-400, 400), rng.uniform(0, 30), rng.uniform(300, 900)), rng.uniform(0.05, 0.3)) for _ in range(3000)])
    o = np.tile(np.array([[0, 3, -6]], np.float32), (n, 1))
    k = rng.integers(0, len(sph), n)
    v = sph[k, :3] - o; v /= np.linalg.norm(v, axis=1, keepdims=True)
    perp = np.cross(v, rng.normal(size=(n, 3))); perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    tgt = sph[k, :3] + perp * sph[k, 3:4] * rng.uniform(0.9, 1.15, (n, 1))
    d = (tgt - o); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 2] *= rng.uniform(1e-3, 1e3, (n // 2, 1))
    _check_queries(rt, sph, np.concatenate([o, d.astype(np.float32)], 1))


def test_queries_degenerate(rt):
    rng = np.random.default_rng(5)
    n = 20000
    sph = np.stack([mk((1.0, 0.5, 5 + 0.5 * (i % 7)), 0.4) for i in range(50)])
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d[:, 2] = np.abs(d[:, 2])
    d[:n // 3, 0] = 0; d[n // 3:2 * n // 3, 1] = 0; d[:100] = 0
    o[:500] = sph[rng.integers(0, 50, 500), :3]
    _check_queries(rt, sph, np.concatenate([o, d], 1))
    _check_queries(rt, sph[:2], np.concatenate([o, d], 1))


def test_config3_reduced_vs_oracle(rt):
    """1,024 spheres + plane + 4 lights: staged brute force and LBVH vs the oracle at 480x270 (oracle = brute force)."""
    sc = scenes.config3_scene()
    w, h = 480, 270
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    ref = O.render(sc, cam, w, h, 8, want_hash=True, want_aov=True)
    for accel in (rt.RT_ACCEL_BRUTE, rt.RT_ACCEL_LBVH, rt.RT_ACCEL_AUTO):
        ctx = rt.Context([0]); ctx.set_scene(sc, accel)
        got = ctx.render_debug(cam, w, h, 8)
        assert np.array_equal(got["hash"], ref["hash"]), accel
        assert np.array_equal(got["aov_id"], ref["aov_id"]) and np.array_equal(got["aov_t"].view(np.uint32), ref["aov_t"].view(np.uint32))
        for k in ("primary", "shadow", "secondary", "plane_tests", "shade_diffuse", "shade_specular", "shade_mirror", "shaded_hits"):
            assert got["counters"][k] == ref["counters"][k], (accel, k)
        px, _ = ctx.render(cam, w, h, 8)
        assert np.array_equal(px, ref["pixels"]), accel
        ctx.close()


def test_config3_4k_brute_equals_lbvh(rt):
    """BASELINE configs[2] at full size: identical (hit id, t bits) per ray — via the chain hash — and identical image."""
    sc = scenes.config3_scene()
    w, h = 3840, 2160
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    out = {}
    for name, accel in (("brute", rt.RT_ACCEL_BRUTE), ("lbvh", rt.RT_ACCEL_LBVH)):
        ctx = rt.Context([0]); ctx.set_scene(sc, accel)
        dbg = ctx.render_debug(cam, w, h, 8)
        px, st = ctx.render(cam, w, h, 8)
        assert np.array_equal(px, dbg["pixels"])
        out[name] = (px, dbg["hash"], dbg["aov_id"], dbg["aov_t"])
        ctx.close()
    assert np.array_equal(out["brute"][0], out["lbvh"][0])
    assert np.array_equal(out["brute"][1], out["lbvh"][1])
    assert np.array_equal(out["brute"][2], out["lbvh"][2])
    assert np.array_equal(out["brute"][3].view(np.uint32), out["lbvh"][3].view(np.uint32))
    # and a pixel subset against the oracle at full resolution
    idx = np.random.default_rng(7).choice(w * h, 4096, replace=False).astype(np.int32)
    ref = O.render(sc, cam, w, h, 8, subset=idx)["pixels"]
    assert np.array_equal(out["lbvh"][0].reshape(-1)[idx], ref)


def test_config4_100k_lbvh_vs_oracle_subset(rt):
    """BASELINE configs[3]: 100k spheres, LBVH, 4K; CPU equality on a fixed 65,536-pixel subset (seed 7) with brute force."""
    sc = scenes.config4_scene()
    w, h = 3840, 2160
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    ctx = rt.Context([0]); ctx.set_scene(sc, rt.RT_ACCEL_AUTO)
    px, st = ctx.render(cam, w, h, 8)
    idx = np.random.default_rng(7).choice(w * h, 65536, replace=False).astype(np.int32)
    ref = O.render(sc, cam, w, h, 8, subset=idx, want_hash=True)
    assert np.array_equal(px.reshape(-1)[idx], ref["pixels"])
    dbg = ctx.render_debug(cam, w, h, 8)
    assert np.array_equal(dbg["hash"].reshape(-1)[idx], ref["hash"])
    assert np.array_equal(dbg["pixels"], px)
    ctx.close()


@pytest.mark.parametrize("n,accel_name", [(3, "tiny"), (300, "brute"), (300, "lbvh"), (20000, "lbvh")])
def test_update_spheres_refit_stays_exact(rt, n, accel_name):
    """rt_update_spheres (scene mutation + LBVH refit): after moving / resizing / re-materialising spheres the
    frame must equal the oracle on the new scene and a freshly built context, on every scene path."""
    accel = {"tiny": rt.RT_ACCEL_AUTO, "brute": rt.RT_ACCEL_BRUTE, "lbvh": rt.RT_ACCEL_LBVH}[accel_name]
    sc = scenes.default_scene() if n == 3 else scenes.Scene(scenes.random_spheres_scene(n, 77, 30.0, 4.0, 64.0, "m")[0],
                                                           scenes.default_scene().planes, scenes.config3_scene().lights, scenes.REF_AMBIENT)
    w, h = 320, 180
    cam = scenes.make_camera(width=w, height=h, **(dict() if n == 3 else scenes.SCALED_CAMERA))
    ctx = rt.Context([0]); ctx.set_scene(sc, accel)
    before, _ = ctx.render(cam, w, h, 8)
    rng = np.random.default_rng(5)
    moved = sc.spheres.copy()
    k = max(1, n // 3)
    idx0 = n // 4
    moved[idx0:idx0 + k, 0:3] += rng.normal(size=(k, 3)).astype(np.float32) * np.float32(0.8)     # move
    moved[idx0:idx0 + k, 3] *= rng.uniform(0.6, 1.5, k).astype(np.float32)                         # resize
    moved[idx0:idx0 + k, 17] = moved[idx0:idx0 + k, 3] * moved[idx0:idx0 + k, 3]
    moved[idx0, 4:17] = scenes.mat_mirror((0.8, 0.8, 0.8))                                         # new material
    ctx.update_spheres(moved[idx0:idx0 + k], first=idx0)
    sc2 = scenes.Scene(moved, sc.planes, sc.lights, sc.ambient)
    got, _ = ctx.render(cam, w, h, 8)
    fresh = rt.Context([0]); fresh.set_scene(sc2, accel)
    ref_gpu, _ = fresh.render(cam, w, h, 8)
    fresh.close()
    assert not np.array_equal(got, before)
    assert np.array_equal(got, ref_gpu)
    if n <= 300:
        assert np.array_equal(got, O.render(sc2, cam, w, h, 8)["pixels"])
    else:
        sub = np.random.default_rng(7).choice(w * h, 3000, replace=False).astype(np.int32)
        assert np.array_equal(got.reshape(-1)[sub], O.render(sc2, cam, w, h, 8, subset=sub)["pixels"])
    with pytest.raises(rt.RtError):
        ctx.update_spheres(moved[:2], first=n - 1)          # out of range
    ctx.close()


@pytest.mark.parametrize("which", ["config3", "noise", "non_finite"])
def test_device_shadow_bins_equal_host_bins(rt, which):
    """The per-light shadow bins are built on the GPU (csrc/rt_shadow_grid_build.cuh) with the geometry code of the host build
    (rt_shadow_grid.cuh, checked against the reference's loop over all spheres in tests/test_shadow_bins.py): same frame, same chain
    hashes, and the same NUMBER of sphere tests — every shadow query scans a cell with the same population under both builds."""
    if which == "config3":
        sc = scenes.config3_scene(); cam_kw = scenes.SCALED_CAMERA
    elif which == "noise":
        rng = np.random.default_rng(21)
        sph = np.stack([scenes.sphere((rng.uniform(-400, 400), rng.uniform(-0.8, 6), rng.uniform(20, 700)), rng.uniform(0.05, 0.35),
                                      scenes.mat_diffuse((0.9, 0.6, 0.3))) for _ in range(1500)])
        lights = np.stack([scenes.light((-300, 40, 100), 1.0), scenes.light((0, 500, 0), 1.0), scenes.light((250, 3, 650), 1.0), scenes.light((0, 0, 0), 1.0)])
        sc = scenes.Scene(sph, scenes.reference_plane()[None], lights, scenes.REF_AMBIENT); cam_kw = dict(pos=(0, 8, -5), pitch=0.12)
    else:
        sc = scenes.small_random_scene(200, 5)
        sph = sc.spheres.copy()
        sph[3, 0] = np.nan; sph[7, 17] = np.inf; sph[11, 2] = -np.inf; sph[199, 17] = np.nan; sph[20, 1] = np.inf
        sc = scenes.Scene(sph, sc.planes, sc.lights, sc.ambient); cam_kw = dict(pos=(0, 1.5, -4.0), pitch=0.1)
    w, h = 640, 360
    cam = scenes.make_camera(width=w, height=h, **cam_kw)
    res = []
    for host in (1, 0):
        ctx = rt.Context([0])
        ctx.set_option(rt.RT_OPT_HOST_SHADOW_BINS, host)
        ctx.set_scene(sc, rt.RT_ACCEL_LBVH)
        px, _ = ctx.render(cam, w, h, 8)
        dbg = ctx.render_debug(cam, w, h, 8)
        res.append((px, dbg, ctx.get_info(rt.RT_INFO_SHADOW_BIN_PAIRS)))
        ctx.close()
    (pa, da, na), (pb, db, nb) = res
    assert na > 0 and na == nb                                   # same number of (cell, sphere) pairs
    assert np.array_equal(pa, pb) and np.array_equal(da["hash"], db["hash"])
    assert da["counters"] == db["counters"] and da["lbvh"] == db["lbvh"]
    assert np.array_equal(da["pixels"], pa)
    if which == "config3":
        assert da["lbvh"]["node_visits_shadow"] < da["counters"]["shadow"]       # the bins, not the traversal, answered most shadow rays


def test_update_spheres_rebuilds_the_bins_on_the_gpu(rt):
    """rt_update_spheres at 100 k spheres (BASELINE configs[3]): refit + shadow-bin rebuild on the GPU stay exact (same frame as a
    freshly built context) and take milliseconds, not the 0.2-0.6 s of the former host build."""
    import time
    sc = scenes.config4_scene()
    w, h = 480, 270
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    ctx = rt.Context([0]); ctx.set_scene(sc, rt.RT_ACCEL_LBVH)
    moved = sc.spheres.copy()
    rng = np.random.default_rng(3)
    moved[:, 0] += rng.normal(size=len(moved)).astype(np.float32) * np.float32(0.3)
    moved[:, 1] += np.abs(rng.normal(size=len(moved))).astype(np.float32) * np.float32(0.2)
    ctx.update_spheres(moved, first=0)                          # warm-up (buffers grow once)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ctx.update_spheres(moved, first=0); ts.append(time.perf_counter() - t0)
    bins_ms = ctx.get_info(rt.RT_INFO_SHADOW_BINS_NS) / 1e6
    got, _ = ctx.render(cam, w, h, 8)
    fresh = rt.Context([0]); fresh.set_scene(scenes.Scene(moved, sc.planes, sc.lights, sc.ambient), rt.RT_ACCEL_LBVH)
    ref, _ = fresh.render(cam, w, h, 8)
    fresh.close(); ctx.close()
    assert np.array_equal(got, ref)
    print("rt_update_spheres(100k): %.2f ms total, shadow bins %.2f ms" % (min(ts) * 1e3, bins_ms))
    assert bins_ms < 20.0


# ---- per-frame primary bins (csrc/rt_primary_bins.cuh, RT_OPT_PRIMARY_BINS) -------------------------------------------------------
def _ctx(rt, sc, accel, bins):
    ctx = rt.Context([0])
    ctx.set_option(rt.RT_OPT_PRIMARY_BINS, bins)
    ctx.set_scene(sc, accel)
    return ctx


def _same_debug(a, b):
    assert np.array_equal(a["pixels"], b["pixels"])
    assert np.array_equal(a["hash"], b["hash"])
    assert np.array_equal(a["aov_id"], b["aov_id"]) and np.array_equal(a["aov_t"].view(np.uint32), b["aov_t"].view(np.uint32))
    for k in ("primary", "shadow", "secondary", "plane_tests", "shade_diffuse", "shade_specular", "shade_mirror", "shaded_hits"):
        assert a["counters"][k] == b["counters"][k], k


def test_primary_bins_config3_4k_same_hits_as_the_tree(rt):
    """BASELINE configs[2] at full size: with the bins the primary rays take the same hits (chain hash = ids + t bits of every ray,
    primary AOVs) and the frame is the same, while (almost) no primary ray walks the tree; a pixel subset against the oracle."""
    sc = scenes.config3_scene()
    w, h = 3840, 2160
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    res = {}
    for bins in (0, 1):
        ctx = _ctx(rt, sc, rt.RT_ACCEL_LBVH, bins)
        dbg = ctx.render_debug(cam, w, h, 8)
        px, _ = ctx.render(cam, w, h, 8)
        assert np.array_equal(px, dbg["pixels"])
        builds = ctx.get_info(rt.RT_INFO_PRIMARY_BIN_BUILDS)
        res[bins] = dbg
        ctx.close()
    _same_debug(res[0], res[1])
    assert builds == 1                                       # (the context with bins) debug and render share the frame's bins
    assert res[0]["lbvh"]["node_visits_primary"] > 4 * w * h
    assert res[1]["lbvh"]["node_visits_primary"] < 0.05 * res[0]["lbvh"]["node_visits_primary"]
    assert res[1]["lbvh"]["node_visits_secondary"] == res[0]["lbvh"]["node_visits_secondary"]
    idx = np.random.default_rng(7).choice(w * h, 4096, replace=False).astype(np.int32)
    ref = O.render(sc, cam, w, h, 8, subset=idx, want_hash=True)
    assert np.array_equal(res[1]["pixels"].reshape(-1)[idx], ref["pixels"])
    assert np.array_equal(res[1]["hash"].reshape(-1)[idx], ref["hash"])


def test_primary_bins_config4_100k_same_frame_and_oracle_subset(rt):
    """BASELINE configs[3]: 100k spheres at 4K. Crowded horizon tiles keep no list and traverse; everything else folds over its list."""
    sc = scenes.config4_scene()
    w, h = 3840, 2160
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    res = {}
    for bins in (0, 1):
        ctx = _ctx(rt, sc, rt.RT_ACCEL_AUTO, bins)
        px, _ = ctx.render(cam, w, h, 8)
        dbg = ctx.render_debug(cam, w, h, 8)
        assert np.array_equal(px, dbg["pixels"])
        res[bins] = dbg
        ctx.close()
    _same_debug(res[0], res[1])
    assert res[1]["lbvh"]["node_visits_primary"] < 0.5 * res[0]["lbvh"]["node_visits_primary"]
    idx = np.random.default_rng(7).choice(w * h, 8192, replace=False).astype(np.int32)
    ref = O.render(sc, cam, w, h, 8, subset=idx, want_hash=True)
    assert np.array_equal(res[1]["pixels"].reshape(-1)[idx], ref["pixels"])
    assert np.array_equal(res[1]["hash"].reshape(-1)[idx], ref["hash"])


def test_primary_bins_follow_the_camera_the_frame_size_and_the_scene(rt):
    """The bins are per (camera, frame size, sphere records): a moving / turning camera, alternating frame sizes, a batch of distinct
    cameras, rt_update_spheres — each frame equals the frame of a context without bins; a camera that stands still does not rebuild."""
    sc = scenes.config3_scene()
    a = _ctx(rt, sc, rt.RT_ACCEL_LBVH, 1)
    b = _ctx(rt, sc, rt.RT_ACCEL_LBVH, 0)
    rng = np.random.default_rng(3)
    frames = [((480, 270), dict(pos=(0.0, 3.0, -6.0), yaw=0.0, pitch=0.25)), ((480, 270), dict(pos=(0.0, 3.0, -6.0), yaw=0.0, pitch=0.25)),
              ((480, 270), dict(pos=(0.0, 3.0, -6.0), yaw=0.01, pitch=0.25)), ((333, 187), dict(pos=(0.0, 3.0, -6.0), yaw=0.01, pitch=0.25)),
              ((480, 270), dict(pos=(7.5, 1.25, 9.0), yaw=0.9, pitch=-0.2)), ((640, 360), dict(pos=(-11.0, 6.0, 30.0), yaw=2.8, pitch=0.6)),
              ((640, 360), dict(pos=(0.3, 0.2, 20.0), yaw=-1.3, pitch=0.05)), ((200, 120), dict(pos=(3.0, 40.0, 25.0), yaw=0.2, pitch=1.45))]
    builds = []
    for (w, h), kw in frames:
        cam = scenes.make_camera(width=w, height=h, **kw)
        pa, _ = a.render(cam, w, h, 8)
        pb, _ = b.render(cam, w, h, 8)
        assert np.array_equal(pa, pb), kw
        builds.append(a.get_info(rt.RT_INFO_PRIMARY_BIN_BUILDS))
    # a batch of distinct cameras (one launch per frame on the LBVH path, each with its own bins)
    w, h = 320, 180
    cams = np.stack([scenes.make_camera(width=w, height=h, pos=(0.0, 3.0, -6.0 + 0.5 * k), yaw=0.02 * k, pitch=0.25) for k in range(4)])
    fa, _ = a.render_batch(cams, w, h, 8, headless=False)
    fb, _ = b.render_batch(cams, w, h, 8, headless=False)
    assert np.array_equal(fa, fb) and not np.array_equal(fa[0], fa[3])
    # the scene changes under a standing camera
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    before, _ = a.render(cam, w, h, 8)
    moved = sc.spheres.copy()
    moved[100:400, 0:3] += rng.normal(size=(300, 3)).astype(np.float32) * np.float32(1.5)
    moved[100:400, 3] *= rng.uniform(0.5, 2.0, 300).astype(np.float32); moved[100:400, 17] = moved[100:400, 3] * moved[100:400, 3]
    for c in (a, b):
        c.update_spheres(moved[100:400], first=100)
    pa, _ = a.render(cam, w, h, 8)
    pb, _ = b.render(cam, w, h, 8)
    assert np.array_equal(pa, pb) and not np.array_equal(pa, before)
    assert np.array_equal(pa, O.render(scenes.Scene(moved, sc.planes, sc.lights, sc.ambient), cam, w, h, 8)["pixels"])
    assert builds == [1, 1, 2, 3, 4, 5, 6, 7] and b.get_info(rt.RT_INFO_PRIMARY_BIN_BUILDS) == 0
    a.close(); b.close()


def test_primary_bins_fallbacks_supersampling_and_partitions(rt):
    """Frames the bins do not serve are the tree's: a camera outside the gate derivation (basis not orthonormal), the eye inside spheres,
    non-finite sphere records, supersampled frames; and a row-tile partition renders its tiles from the frame's bins."""
    rng = np.random.default_rng(12)
    sph, _ = scenes.random_spheres_scene(600, 5, 12.0, 2.0, 40.0, "pb")
    sph = sph.copy()
    sph[0, 0:3] = (0.0, 3.0, -6.0); sph[0, 3] = 2.0; sph[0, 17] = 4.0                 # the eye of SCALED_CAMERA is inside this one
    sph[7, 0] = np.nan; sph[9, 17] = np.inf; sph[11, 2] = -np.inf
    d = scenes.config3_scene()
    sc = scenes.Scene(sph, d.planes, d.lights, d.ambient)
    w, h = 400, 225
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    skew = cam.copy(); skew[3:6] *= np.float32(1.01)
    a = _ctx(rt, sc, rt.RT_ACCEL_LBVH, 1)
    b = _ctx(rt, sc, rt.RT_ACCEL_LBVH, 0)
    for c_, spp in ((cam, 1), (skew, 1), (cam, 4)):
        pa, _ = a.render(c_, w, h, 8, spp, 9)
        pb, _ = b.render(c_, w, h, 8, spp, 9)
        assert np.array_equal(pa, pb)
    assert np.array_equal(a.render(cam, w, h, 8)[0], O.render(sc, cam, w, h, 8)["pixels"])
    _same_debug(a.render_debug(cam, w, h, 8), b.render_debug(cam, w, h, 8))
    full, _ = b.render(cam, w, h, 8)
    b.close()
    # two ranks of a partition on this GPU store their tiles into one poisoned buffer
    buf = a.dev_alloc(w * h * 4)
    try:
        a.dev_memset(buf, 0xAB, w * h * 4)
        r1 = _ctx(rt, sc, rt.RT_ACCEL_LBVH, 1)
        a.set_partition(0, 2, 8); r1.set_partition(1, 2, 8)
        a.render_device(cam[None], w, h, 8, 1, 0, buf); r1.render_device(cam[None], w, h, 8, 1, 0, buf)
        a.sync(); r1.sync()
        got = a.dev_to_host(buf, w * h * 4).reshape(h, w)
        assert np.array_equal(got, full)
        r1.close()
    finally:
        a.dev_free(buf)
    a.close()
