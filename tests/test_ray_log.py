"""Ray log (rt_ray_log): the per-ray digest of the reference's DEBUG_ENABLE TracedRay log (RayTracer.cs:424-435,
:601, :639, :801, drawn at :914-933).  CPU part: the device logging code (LogDbg, compiled as plain C++) must equal the
oracle's log byte for byte, and the log must be consistent with everything else the oracle reports (ray counters, primary AOVs,
shadow queries).  GPU part (-m gpu): librtb200.so's rt_ray_log must equal the oracle's log byte for byte."""
import numpy as np
import pytest

import hostemu_lib as E
import oracle_lib as O
import scenes

CASES = [
    ("default_d32", scenes.default_scene, dict(), 160, 90, 32),
    ("default_moved_d8", scenes.default_scene, dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 192, 108, 8),
    ("default_above_d0", scenes.default_scene, dict(pos=(0, 4.0, 6.0), pitch=1.2), 128, 128, 0),
    ("small12_d8", lambda: scenes.small_random_scene(12, 1), dict(pos=(0, 1.5, -4.0), pitch=0.1), 160, 100, 8),
    ("small40_d3", lambda: scenes.small_random_scene(40, 3), dict(pos=(0, 1.5, -4.0), pitch=0.1), 160, 100, 3),
    ("degenerate_d4", scenes.degenerate_scene, dict(), 96, 64, 4),
    ("empty_d2", lambda: scenes.small_random_scene(0, 1), dict(pos=(0, 1.5, -4.0), pitch=0.1), 64, 48, 2),
]


def _sample_pixels(w, h, n, seed):
    rng = np.random.default_rng(seed)
    return rng.choice(w * h, size=min(n, w * h), replace=False).astype(np.uint32)


def _case(name):
    for c in CASES:
        if c[0] == name:
            _, mk, camkw, w, h, depth = c
            return mk(), scenes.make_camera(width=w, height=h, **camkw), w, h, depth
    raise KeyError(name)


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_device_log_code_equals_oracle(built, name):
    sc, cam, w, h, depth = _case(name)
    px = _sample_pixels(w, h, 700, 11)
    ref = O.ray_log(sc, cam, w, h, depth, px)
    got = E.ray_log(sc, cam, w, h, depth, px)
    assert len(ref) == len(got) and len(ref) >= len(px)
    assert ref.tobytes() == got.tobytes()


@pytest.mark.parametrize("name", ["default_d32", "default_moved_d8", "small12_d8", "small40_d3"])
def test_log_is_consistent_with_counters_and_aovs(built, name):
    sc, cam, w, h, depth = _case(name)
    px = _sample_pixels(w, h, 900, 5)
    log = O.ray_log(sc, cam, w, h, depth, px)
    sub = O.render(sc, cam, w, h, depth, subset=px.astype(np.int32), want_hash=True, want_aov=True)
    kinds = np.bincount(log["kind"], minlength=3)
    assert kinds[0] == sub["counters"]["primary"] == len(px)
    assert kinds[1] == sub["counters"]["secondary"]
    assert kinds[2] == sub["counters"]["shadow"]
    # records are grouped by pixel in list order, each group opens with the primary ray, and that ray is the AOV hit
    first = np.flatnonzero(log["kind"] == 0)
    assert np.array_equal(log["pixel"][first], px)
    assert np.array_equal(log["hit"][first], sub["aov_id"])
    assert np.array_equal(log["distance"][first].view(np.uint32), sub["aov_t"].view(np.uint32))
    bounds = np.append(first, len(log))
    for a, b in zip(bounds[:-1], bounds[1:]):
        assert (log["pixel"][a:b] == log["pixel"][a]).all()
    # hit_point = origin + direction * distance in fp32 (TracedRay.hitPoint), distance 0 on a miss
    hp = log["origin"] + log["direction"] * log["distance"][:, None]
    assert np.array_equal(hp.view(np.uint32), log["hit_point"].view(np.uint32))
    assert (log["distance"][log["hit"] < 0] == 0).all()
    # chain structure: the rays of the descent carry levels 0, 1, 2, ...; a secondary ray starts at the previous hit point
    for a, b in zip(bounds[:-1], bounds[1:]):
        g = log[a:b]
        chain = g[g["kind"] < 2]
        assert np.array_equal(chain["level"], np.arange(len(chain)))
        assert len(chain) <= depth + 2
        for k in range(1, len(chain)):
            assert np.array_equal(chain["origin"][k].view(np.uint32), chain["hit_point"][k - 1].view(np.uint32))
        sh = g[g["kind"] == 2]
        if len(sh):
            assert (np.diff(sh["level"].astype(np.int64)) <= 0).all()        # deepest level first
            for lv in np.unique(sh["level"]):
                assert np.array_equal(sh["light"][sh["level"] == lv], np.arange(len(sc.lights)))
                # shadow rays of level lv start at the hit point of the chain ray of that level, towards the light POSITION (:574)
                o = sh["origin"][sh["level"] == lv]
                assert (o.view(np.uint32) == chain["hit_point"][lv].view(np.uint32)).all()
            assert np.array_equal(sh["direction"], sc.lights[sh["light"], 0:3])
    # shadow records agree with the any-hit query of the oracle on the same rays
    sh = log[log["kind"] == 2]
    if len(sh) and len(sc.spheres):
        ids, _ = O.query_spheres(sc.spheres, np.concatenate([sh["origin"], sh["direction"]], axis=1), 2)
        assert np.array_equal(ids == 1, sh["hit"] >= 0)


def test_known_answers_default_scene(built):
    """the centre pixel's ray (0, +-0, 1) misses every sphere (centres at x = 2.5, 3, -3, r = 1,
    RayTracer.cs:441-457) and is parallel to the floor: one record, no hit. Sphere 0's centre (2.5, 0, 8) projects to
    ~(0.652 w, 0.5 h): that pixel hits sphere 0 almost head-on, just beyond |c| - r."""
    sc = scenes.default_scene()
    w, h = 1000, 500
    cam = scenes.make_camera(width=w, height=h)
    log = O.ray_log(sc, cam, w, h, 2, np.array([(h // 2) * w + w // 2], np.uint32))
    assert len(log) == 1
    p = log[0]
    assert p["kind"] == 0 and p["level"] == 0 and p["hit"] == -1 and p["distance"] == 0
    assert tuple(p["origin"]) == (0.0, 0.0, 0.0) and tuple(np.abs(p["direction"])) == (0.0, 0.0, 1.0)
    assert tuple(p["hit_point"]) == (0.0, 0.0, 0.0)
    log = O.ray_log(sc, cam, w, h, 2, np.array([(h // 2) * w + 652], np.uint32))
    p = log[0]
    c = sc.spheres[0, 0:3]
    assert p["kind"] == 0 and p["hit"] == 0
    assert 0 <= p["distance"] - (np.linalg.norm(c) - 1.0) < 0.05                # |c| - r is the nearest any ray can hit it
    assert abs(np.linalg.norm(p["hit_point"] - c) - 1.0) < 1e-4               # the hit point lies on the sphere
    # sphere 0 is Diffuse (:442): no secondary ray, one shadow ray per light from the hit point towards the light POSITION (:574)
    assert [int(k) for k in log["kind"]] == [0, 2, 2]
    assert np.array_equal(log["direction"][1:], sc.lights[:, 0:3])
    assert np.array_equal(log["origin"][1:].view(np.uint32), np.stack([p["hit_point"]] * 2).view(np.uint32))


def test_empty_pixel_list_and_bad_pixel(built):
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=32, height=32)
    assert len(O.ray_log(sc, cam, 32, 32, 4, np.zeros(0, np.uint32))) == 0
    assert len(E.ray_log(sc, cam, 32, 32, 4, np.zeros(0, np.uint32))) == 0
    with pytest.raises(RuntimeError):
        O.ray_log(sc, cam, 32, 32, 4, np.array([32 * 32], np.uint32))


# ---- GPU: the library itself -------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_gpu_ray_log_equals_oracle(built, name):
    import rtb200
    sc, cam, w, h, depth = _case(name)
    px = _sample_pixels(w, h, 2000, 3)
    ref = O.ray_log(sc, cam, w, h, depth, px)
    with rtb200.Context([0]) as ctx:
        ctx.set_scene(sc)
        got = ctx.ray_log(cam, w, h, depth, px)
    assert len(got) == len(ref)
    assert got.tobytes() == ref.tobytes()


@pytest.mark.gpu
def test_gpu_ray_log_on_lbvh_scene_matches_render(built):
    """A scene rendered through the LBVH: the log (always brute force) must name the same primary hits as the LBVH render's AOVs."""
    import rtb200
    sc = scenes.small_random_scene(300, 9)
    w, h, depth = 320, 200, 6
    cam = scenes.make_camera(pos=(0, 1.5, -4.0), pitch=0.1, width=w, height=h)
    px = _sample_pixels(w, h, 3000, 8)
    ref = O.ray_log(sc, cam, w, h, depth, px)
    with rtb200.Context([0]) as ctx:
        ctx.set_scene(sc, rtb200.RT_ACCEL_LBVH)
        got = ctx.ray_log(cam, w, h, depth, px)
        dbg = ctx.render_debug(cam, w, h, depth, 1, 0)
    assert got.tobytes() == ref.tobytes()
    first = got[got["kind"] == 0]
    assert np.array_equal(first["hit"], dbg["aov_id"].reshape(-1)[px])
    assert np.array_equal(first["distance"].view(np.uint32), dbg["aov_t"].reshape(-1)[px].view(np.uint32))


@pytest.mark.gpu
def test_gpu_ray_log_sizing_truncation_and_errors(built):
    import ctypes as C
    import rtb200
    sc = scenes.default_scene()
    w, h, depth = 128, 72, 8
    cam = scenes.make_camera(width=w, height=h)
    px = _sample_pixels(w, h, 500, 2)
    with rtb200.Context([0]) as ctx:
        ctx.set_scene(sc)
        full = ctx.ray_log(cam, w, h, depth, px)
        camrec = rtb200.to_rt_camera(cam)
        n = C.c_int(-1)
        half = np.zeros(len(full) // 2, dtype=rtb200.RAY_RECORD)
        rc = ctx.lib.rt_ray_log(ctx.h, C.byref(camrec), w, h, depth, px.ctypes.data_as(C.POINTER(C.c_uint32)), len(px),
                                C.c_void_p(half.ctypes.data), len(half), C.byref(n))
        assert rc == 0 and n.value == len(full)                   # the full count is reported, only max_records are written
        assert half.tobytes() == full[:len(half)].tobytes()
        assert len(ctx.ray_log(cam, w, h, depth, np.zeros(0, np.uint32))) == 0
        with pytest.raises(RuntimeError):
            ctx.ray_log(cam, w, h, depth, np.array([w * h], np.uint32))
        with pytest.raises(RuntimeError):
            ctx.ray_log(cam, w, h, 33, px)
