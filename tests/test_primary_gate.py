"""The host's primary-ray sphere gate (csrc/rt_gate.cuh): outside its rectangle the kernels skip the sphere loop of primary rays,
so the rectangle must contain EVERY pixel whose primary ray the oracle reports as hitting a sphere — for any camera — and it
must degrade to the full frame whenever its derivation does not apply."""
import numpy as np
import pytest

import hostemu_lib as E
import oracle_lib as O
import scenes


def _sphere_hit_mask(sc, cam, w, h):
    a = O.render(sc, cam, w, h, 0, want_aov=True)
    return (a["aov_id"] >= 0) & (a["aov_id"] < len(sc.spheres))


def _check(sc, cam, w, h):
    x0, y0, x1, y1 = E.gate_rect(sc, cam, w, h)
    m = _sphere_hit_mask(sc, cam, w, h)
    ys, xs = np.nonzero(m)
    if len(xs):
        assert x0 <= xs.min() and xs.max() <= x1 and y0 <= ys.min() and ys.max() <= y1, ((x0, y0, x1, y1), (xs.min(), ys.min(), xs.max(), ys.max()))
    return (x0, y0, x1, y1), m


def test_default_scene_rect_is_tight():
    sc = scenes.default_scene()
    w, h = 640, 360
    (x0, y0, x1, y1), m = _check(sc, scenes.make_camera(width=w, height=h), w, h)
    ys, xs = np.nonzero(m)
    assert xs.min() - x0 <= 4 and x1 - xs.max() <= 4 and ys.min() - y0 <= 4 and y1 - ys.max() <= 4
    assert (x1 - x0 + 1) * (y1 - y0 + 1) < 0.35 * w * h


@pytest.mark.parametrize("seed", range(12))
def test_random_cameras_contain_every_sphere_hit(built, seed):
    rng = np.random.default_rng(seed)
    sc = scenes.default_scene() if seed % 3 == 0 else scenes.small_random_scene(int(rng.integers(1, 9)), seed)
    w, h = 200, 120
    for _ in range(6):
        pos = tuple(rng.uniform(-6, 6, 3) * np.array([1, 0.5, 1]) + np.array([0, 1.0, -3]))
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.2, 1.2)), width=w, height=h)
        _check(sc, cam, w, h)


def test_far_camera_large_coordinates(built):
    """|P| ~ 1e4: the fp32 primary directions are coarse (vp - P cancels); the gate must widen accordingly or give up."""
    sc = scenes.default_scene()
    sc.spheres[:, 0:3] += np.float32([10000.0, 0.0, 10000.0])
    w, h = 240, 135
    cam = scenes.make_camera(pos=(10000.0, 0.0, 10000.0), width=w, height=h)
    _check(sc, cam, w, h)


def test_degenerate_inputs_disable_the_gate(built):
    sc = scenes.default_scene()
    w, h = 64, 48
    full = (0, 0, w - 1, h - 1)
    cam = scenes.make_camera(width=w, height=h)
    inside = np.array(cam, np.float32); inside[0:3] = sc.spheres[0, 0:3]                   # camera inside a sphere
    assert E.gate_rect(sc, inside, w, h) == full
    skew = np.array(cam, np.float32); skew[3:6] = (1.0, 0.2, 0.0)                          # basis not orthonormal
    assert E.gate_rect(sc, skew, w, h) == full
    nan = np.array(cam, np.float32); nan[1] = np.nan
    assert E.gate_rect(sc, nan, w, h) == full
    flat = np.array(cam, np.float32); flat[12] = 0.0                                       # zero-width view plane
    assert E.gate_rect(sc, flat, w, h) == full
    behind = scenes.make_camera(yaw=3.14159, width=w, height=h)                            # every sphere behind the camera
    assert E.gate_rect(sc, behind, w, h) == (w, h, w, h)
    assert E.gate_rect(scenes.small_random_scene(0, 1), cam, w, h) == (w, h, w, h)         # no spheres: always skip
    for c in (inside, skew, behind):                                                       # and rendering stays exact either way
        a = O.render(sc, c, w, h, 4)
        b = E.render(sc, c, w, h, 4, tiny=2)
        assert np.array_equal(a["pixels"], b["pixels"])


def test_crossing_camera_plane_is_full(built):
    sc = scenes.default_scene()
    w, h = 64, 48
    cam = scenes.make_camera(pos=(2.5, 0.0, 7.5), yaw=1.5708, width=w, height=h)            # sphere 0 straddles the camera plane
    r = E.gate_rect(sc, cam, w, h)
    assert r == (0, 0, w - 1, h - 1)


# ---- sky gate: the side of the plane's horizon on which no primary ray can hit the plane ----------------------------------

def _plane_hit_mask(sc, cam, w, h):
    import copy
    bare = copy.copy(sc)
    bare.spheres = sc.spheres[:0]
    a = O.render(bare, cam, w, h, 0, want_aov=True)
    return a["aov_id"] >= 0


@pytest.mark.parametrize("seed", range(10))
def test_sky_gate_never_claims_a_plane_hit(built, seed):
    rng = np.random.default_rng(1000 + seed)
    sc = scenes.default_scene()
    w, h = 200, 120
    claimed = missed = 0
    for k in range(8):
        pos = tuple(rng.uniform(-6, 6, 3) * np.array([1, 0.6, 1]) + np.array([0, 0.5, -3]))      # above and below the floor
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.5, 1.5)), width=w, height=h)
        sky, _ = E.sky_mask(sc, cam, w, h)
        hit = _plane_hit_mask(sc, cam, w, h)
        assert not (sky & hit).any(), "camera %d: %d plane hits inside the sky mask" % (k, (sky & hit).sum())
        claimed += int(sky.sum()); missed += int((~hit).sum())
    assert claimed > 0.9 * missed       # and it is tight: it claims > 90 % of the pixels that really miss the plane


def test_sky_gate_tilted_plane_and_rolled_camera(built):
    """A plane whose normal is not axis-aligned (and not unit length) seen by a rolled camera: the horizon is an oblique line."""
    sc = scenes.default_scene()
    sc.planes[0, 3:6] = np.float32([0.3, 1.7, -0.4])
    w, h = 160, 120
    for roll in (0.0, 0.4, -1.1, 2.5):
        cam = np.array(scenes.make_camera(pos=(0.5, 1.0, -2.0), yaw=0.3, pitch=0.2, width=w, height=h), np.float32)
        r, u = cam[3:6].copy(), cam[6:9].copy()
        cam[3:6] = np.float32(np.cos(roll) * r + np.sin(roll) * u)
        cam[6:9] = np.float32(-np.sin(roll) * r + np.cos(roll) * u)
        sky, co = E.sky_mask(sc, cam, w, h)
        hit = _plane_hit_mask(sc, cam, w, h)
        assert not (sky & hit).any()
        a = O.render(sc, cam, w, h, 4)
        b = E.render(sc, cam, w, h, 4, tiny=2)
        assert np.array_equal(a["pixels"], b["pixels"])
    assert sky.any() and co[1] != 0.0 and co[2] != 0.0


def test_sky_gate_degenerate_cases(built):
    sc = scenes.default_scene()
    w, h = 64, 48
    cam = scenes.make_camera(width=w, height=h)
    two = scenes.default_scene()
    two.planes = np.concatenate([sc.planes, sc.planes]); two.planes[1, 1] = 5.0; two.planes[1, 4] = -1.0     # a ceiling: 2 planes
    sky, co = E.sky_mask(two, cam, w, h)
    assert not sky.any() and co[0] < 0                                     # more than one plane: the gate is off
    a = O.render(two, cam, w, h, 3); b = E.render(two, cam, w, h, 3, tiny=1)
    assert np.array_equal(a["pixels"], b["pixels"])
    none = scenes.default_scene(); none.planes = sc.planes[:0]
    sky, _ = E.sky_mask(none, cam, w, h)
    assert sky.all()                                                       # no plane: nothing to hit
    a = O.render(none, cam, w, h, 3); b = E.render(none, cam, w, h, 3, tiny=1)
    assert np.array_equal(a["pixels"], b["pixels"])
    on_plane = scenes.make_camera(pos=(0.0, -1.0, 0.0), width=w, height=h)                                  # camera ON the floor: num == 0
    sky, _ = E.sky_mask(sc, on_plane, w, h)
    assert sky.all() and not _plane_hit_mask(sc, on_plane, w, h).any()
    a = O.render(sc, on_plane, w, h, 3); b = E.render(sc, on_plane, w, h, 3, tiny=2)
    assert np.array_equal(a["pixels"], b["pixels"])
    skew = np.array(cam, np.float32); skew[3:6] = (1.0, 0.2, 0.0)
    sky, _ = E.sky_mask(sc, skew, w, h)
    assert not sky.any()


# ---- all gates against the oracle's ray log: every skipped piece of work must be one the oracle finds fruitless --------------

def _check_bits_against_log(sc, cam, w, h, depth=3):
    g = E.gates(sc, cam, w, h)
    bits = g["bits"].reshape(-1)
    ns = len(sc.spheres)
    log = O.ray_log(sc, cam, w, h, depth, np.arange(w * h, dtype=np.uint32))
    prim = log[log["kind"] == 0]
    assert np.array_equal(prim["pixel"], np.arange(w * h))
    phit = prim["hit"]
    on_plane = phit >= ns                                                       # primary ray hit the plane
    assert (phit[(bits & E.GATE_BLACK) != 0] == -1).all()
    assert not ((phit >= 0) & (phit < ns))[(bits & E.GATE_SPHERES) != 0].any()
    # mirror bit: the reflection ray of a primary plane hit contributes nothing
    sec = log[(log["kind"] == 1) & (log["level"] == 1)]
    sel = ((bits[sec["pixel"]] & E.GATE_MIRROR) != 0) & on_plane[sec["pixel"]]
    s = sec[sel]
    fruitless = (s["hit"] == -1) | ((s["hit"] >= ns) & (s["distance"] - np.float32(0.01) <= 0))
    assert fruitless.all(), "%d reflection rays the mirror gate would skip do hit something" % (~fruitless).sum()
    # shadow bits: the shadow ray of light l from the primary plane hit is unoccluded
    sh = log[(log["kind"] == 2) & (log["level"] == 0)]
    for l in range(min(4, len(sc.lights))):
        q = sh[(sh["light"] == l)]
        sel = ((bits[q["pixel"]] & (E.GATE_SHADOW0 << l)) != 0) & on_plane[q["pixel"]]
        assert (q["hit"][sel] == -1).all(), "light %d: %d occluded shadow rays inside the gate" % (l, (q["hit"][sel] != -1).sum())
    return g, bits, on_plane, sec, sh


def test_default_scene_gates_are_sound_and_effective(built):
    sc = scenes.default_scene()
    w, h = 384, 216
    g, bits, on_plane, sec, sh = _check_bits_against_log(sc, scenes.make_camera(width=w, height=h), w, h)
    floor = on_plane.mean()
    assert ((bits & E.GATE_BLACK) != 0).mean() > 0.30
    assert ((bits & E.GATE_MIRROR) != 0).mean() > 0.35 * floor
    assert ((bits & E.GATE_SHADOW0) != 0).mean() > 0.75 * floor


@pytest.mark.parametrize("seed", range(16))
def test_random_cameras_and_scenes_all_gates(built, seed):
    rng = np.random.default_rng(500 + seed)
    sc = scenes.default_scene() if seed % 4 == 0 else scenes.small_random_scene(int(rng.integers(1, 9)), 700 + seed)
    if seed % 4 == 3:                                                           # a tilted (but unit-normal) mirror floor
        n = np.float32([0.2, 1.0, -0.1]); n = (n / np.float32(np.linalg.norm(n))).astype(np.float32)
        sc.planes[0, 3:6] = n
    w, h = 144, 96
    for _ in range(4):
        pos = tuple(rng.uniform(-6, 6, 3) * np.array([1, 0.5, 1]) + np.array([0, 1.0, -3]))
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.4, 1.4)), width=w, height=h)
        _check_bits_against_log(sc, cam, w, h)
        a = O.render(sc, cam, w, h, 5)
        b = E.render(sc, cam, w, h, 5, tiny=1 if len(sc.spheres) > 4 else 2)
        assert np.array_equal(a["pixels"], b["pixels"])


def test_high_camera_and_grazing_light(built):
    """A camera high above the floor (long primary rays: the self-hit bound must raise sin_min or drop the mirror gate) and a light
    almost in the plane (unbounded shadows: that light's gate must stay off)."""
    sc = scenes.default_scene()
    sc.lights[1, 0:3] = np.float32([40.0, 0.0001, 5.0])
    w, h = 160, 120
    for pos, pitch in (((0.0, 60.0, -20.0), 1.0), ((3.0, 400.0, 2.0), 1.5), ((0.0, 0.2, 0.0), 0.05)):
        cam = scenes.make_camera(pos=pos, pitch=pitch, width=w, height=h)
        g, *_ = _check_bits_against_log(sc, cam, w, h)
        assert g["shadow"][1] == (0, 0, w - 1, h - 1)
        a = O.render(sc, cam, w, h, 5)
        b = E.render(sc, cam, w, h, 5, tiny=2)
        assert np.array_equal(a["pixels"], b["pixels"])


def test_non_unit_normal_disables_only_the_mirror_gate(built):
    sc = scenes.default_scene()
    sc.planes[0, 3:6] = np.float32([0.0, 2.0, 0.0])          # :741-744 reflects about the normal AS GIVEN: not a mirror image
    w, h = 160, 90
    cam = scenes.make_camera(width=w, height=h)
    g, bits, *_ = _check_bits_against_log(sc, cam, w, h)
    assert g["mirror"] == (0, 0, w - 1, h - 1) and not (bits & E.GATE_MIRROR).any()
    assert (bits & E.GATE_SHADOW0).any()
    a = O.render(sc, cam, w, h, 5)
    b = E.render(sc, cam, w, h, 5, tiny=2)
    assert np.array_equal(a["pixels"], b["pixels"])


def test_4k_gate_borders_against_the_oracle(built):
    """At the bench resolution the margins are a few pixels wide: check every gated pixel within 5 pixels of a gate border (where
    the skip bits change) plus a random sample, for the bench camera and a moved one, against the oracle's ray log."""
    sc = scenes.default_scene()
    w, h = 3840, 2160
    ns = len(sc.spheres)
    for camkw in (dict(), dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15)):
        cam = scenes.make_camera(width=w, height=h, **camkw)
        bits2 = E.gates(sc, cam, w, h)["bits"]
        chg = np.zeros((h, w), bool)
        chg[:, 1:] |= bits2[:, 1:] != bits2[:, :-1]
        chg[1:, :] |= bits2[1:, :] != bits2[:-1, :]
        ys, xs = np.nonzero(chg)
        near = np.zeros((h, w), bool)
        for dy in range(-5, 6):
            for dx in range(-5, 6):
                near[np.clip(ys + dy, 0, h - 1), np.clip(xs + dx, 0, w - 1)] = True
        near &= bits2 != 0
        b = bits2.reshape(-1)
        rng = np.random.default_rng(3)
        idx = np.flatnonzero(near.reshape(-1))
        if len(idx) > 150000:
            idx = rng.choice(idx, 150000, replace=False)
        idx = np.unique(np.concatenate([idx, rng.choice(np.flatnonzero(b != 0), 50000, replace=False)])).astype(np.uint32)
        log = O.ray_log(sc, cam, w, h, 3, idx)
        prim = log[log["kind"] == 0]
        assert np.array_equal(prim["pixel"], idx)
        ph, pb = prim["hit"], b[idx]
        assert (ph[(pb & E.GATE_BLACK) != 0] == -1).all()
        assert not ((ph >= 0) & (ph < ns))[(pb & E.GATE_SPHERES) != 0].any()
        on_plane = np.zeros(w * h, bool); on_plane[idx[ph >= ns]] = True
        sec = log[(log["kind"] == 1) & (log["level"] == 1)]
        s = sec[((b[sec["pixel"]] & E.GATE_MIRROR) != 0) & on_plane[sec["pixel"]]]
        assert ((s["hit"] == -1) | ((s["hit"] >= ns) & (s["distance"] - np.float32(0.01) <= 0))).all()
        sh = log[(log["kind"] == 2) & (log["level"] == 0)]
        for l in range(len(sc.lights)):
            q = sh[sh["light"] == l]
            assert (q["hit"][((b[q["pixel"]] & (E.GATE_SHADOW0 << l)) != 0) & on_plane[q["pixel"]]] == -1).all()


@pytest.mark.parametrize("seed", range(6))
def test_arbitrary_planes_lights_and_cameras(built, seed):
    """Walls, ceilings, slopes (unit normals in any direction through a random point), random light vectors, cameras at any
    distance and attitude, above or below the plane."""
    rng = np.random.default_rng(9000 + seed)
    for _ in range(10):
        sc = scenes.default_scene() if rng.integers(0, 3) == 0 else scenes.small_random_scene(int(rng.integers(1, 9)), int(rng.integers(0, 10000)))
        n = rng.normal(size=3)
        n = (n / np.linalg.norm(n)).astype(np.float32)
        sc.planes[0, 0:3] = (rng.normal(size=3) * 3).astype(np.float32)
        sc.planes[0, 3:6] = n
        if rng.integers(0, 2) and len(sc.lights):
            sc.lights[:, 0:3] = rng.uniform(-40, 40, (len(sc.lights), 3)).astype(np.float32)
        w, h = int(rng.integers(40, 140)), int(rng.integers(30, 100))
        pos = tuple(rng.normal(size=3) * 10.0 ** rng.uniform(-0.5, 1.5))
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.55, 1.55)), width=w, height=h)
        _check_bits_against_log(sc, cam, w, h)
        a = O.render(sc, cam, w, h, 4)
        b = E.render(sc, cam, w, h, 4, tiny=1)
        assert np.array_equal(a["pixels"], b["pixels"])


# ---- second-generation gate shapes (per-sphere mirror rectangles, shadow hull half-planes; shipped since round 2): same soundness
# bar, and never skipping less than the first-generation shapes (libhostemu_v1.so, -DRT_GATES_V1). ---------------------------------

@pytest.fixture()
def v2(built):
    E.use_variant("")
    yield
    E.use_variant("")


def test_v2_default_scene_is_sound_and_tighter(v2):
    sc = scenes.default_scene()
    w, h = 384, 216
    cam = scenes.make_camera(width=w, height=h)
    g, bits, on_plane, sec, sh = _check_bits_against_log(sc, cam, w, h)
    E.use_variant("_v1")
    base = E.gates(sc, cam, w, h)["bits"].reshape(-1)
    E.use_variant("")
    assert ((base & ~bits) == 0).all()                                    # never skips less than the shipped gates
    assert ((bits & E.GATE_MIRROR) != 0).mean() > 1.4 * ((base & E.GATE_MIRROR) != 0).mean()
    assert ((bits & (E.GATE_SHADOW0 << 1)) != 0).mean() > 1.8 * ((base & (E.GATE_SHADOW0 << 1)) != 0).mean()
    a = O.render(sc, cam, w, h, 8)
    b = E.render(sc, cam, w, h, 8, tiny=2)
    assert np.array_equal(a["pixels"], b["pixels"])


@pytest.mark.parametrize("seed", range(8))
def test_v2_random_scenes_planes_lights_cameras(v2, seed):
    rng = np.random.default_rng(4000 + seed)
    for _ in range(8):
        sc = scenes.default_scene() if rng.integers(0, 3) == 0 else scenes.small_random_scene(int(rng.integers(1, 9)), int(rng.integers(0, 10000)))
        if rng.integers(0, 2):
            n = rng.normal(size=3)
            sc.planes[0, 3:6] = (n / np.linalg.norm(n)).astype(np.float32)
            sc.planes[0, 0:3] = (rng.normal(size=3) * 3).astype(np.float32)
        if rng.integers(0, 2) and len(sc.lights):
            sc.lights[:, 0:3] = rng.uniform(-40, 40, (len(sc.lights), 3)).astype(np.float32)
        w, h = int(rng.integers(40, 160)), int(rng.integers(30, 110))
        pos = tuple(rng.normal(size=3) * 10.0 ** rng.uniform(-0.5, 1.5) + np.array([0, 1.0, -2.0]))
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.55, 1.55)), width=w, height=h)
        _check_bits_against_log(sc, cam, w, h)
        a = O.render(sc, cam, w, h, 4)
        b = E.render(sc, cam, w, h, 4, tiny=1)
        assert np.array_equal(a["pixels"], b["pixels"])
