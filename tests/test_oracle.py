"""CPU tests of the oracle (test infrastructure): self-consistency, known answers derivable from the reference source
(the reference has no tests or golden vectors of its own) and the committed golden fixtures."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
from common import GOLDEN_SCENES, golden_names, load_golden


@pytest.fixture(scope="module")
def default_720(built):
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=1280, height=720)
    return sc, cam, O.render(sc, cam, 1280, 720, 32, want_hash=True, want_aov=True)


@pytest.mark.parametrize("camkw,depth", [(dict(), 32), (dict(), 8), (dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 32),
                                         (dict(pos=(-2.0, 2.5, 3.0), yaw=-0.4, pitch=0.5), 5), (dict(pos=(0, 4.0, 6.0), pitch=1.2), 32)])
def test_faithful_equals_nearest(built, camkw, depth):
    """A.11: shade-all-then-select (the reference as written) == select-then-shade, bit for bit."""
    sc = scenes.default_scene()
    w, h = 320, 180
    cam = scenes.make_camera(width=w, height=h, **camkw)
    a = O.render(sc, cam, w, h, depth, mode="faithful")
    b = O.render(sc, cam, w, h, depth, mode="nearest")
    assert np.array_equal(a["pixels"], b["pixels"])
    assert a["counters"]["faithful_secondary"] >= b["counters"]["secondary"]
    assert a["counters"]["faithful_shadow"] >= b["counters"]["shadow"]


def test_faithful_equals_nearest_random_scene(built):
    sc = scenes.small_random_scene(12, 5)
    cam = scenes.make_camera(pos=(0, 1.5, -4.0), pitch=0.1, width=160, height=100)
    a = O.render(sc, cam, 160, 100, 6, mode="faithful")
    b = O.render(sc, cam, 160, 100, 6, mode="nearest")
    assert np.array_equal(a["pixels"], b["pixels"])


def test_known_answers_default_scene(default_720):
    """Appendix B: facts derivable from RayTracer.cs alone."""
    sc, cam, r = default_720
    px, ids = r["pixels"], r["aov_id"]
    # every primary miss is exactly 0x00000000 (A.9); ~43.6 % of the frame
    assert np.all(px[ids == -1] == 0)
    assert 0.43 < (ids == -1).mean() < 0.445
    # centre pixel: dir = (0, +-0, 1) misses all spheres, plane denominator is +-0 => black
    assert px[360, 640] == 0
    # sphere 1 = Diffuse(1,0,0) (:442): G = B = 0 exactly, R >= 43 (ambient 43/255 * Ka)
    s1 = px[ids == 0]
    assert s1.size > 0 and np.all((s1 & 0xFFFF) == 0) and np.all(((s1 >> 16) & 255) >= 43)
    # sphere 2 = Plastic(0,1,0) (:443): R == B (grey specular only), G >= 43
    s2 = px[ids == 1]
    assert np.all(((s2 >> 16) & 255) == (s2 & 255)) and np.all(((s2 >> 8) & 255) >= 43)
    # floor share and sphere shares (SURVEY §6 planning numbers)
    assert 0.45 < (ids == 3).mean() < 0.465
    assert 0.06 < (ids == 1).mean() < 0.07 and 0.02 < (ids == 2).mean() < 0.025 and 0.017 < (ids == 0).mean() < 0.02
    # sphere 1 centre projects to ~(0.652 w, 0.5 h)
    ys, xs = np.nonzero(ids == 0)
    assert abs(xs.mean() / 1280 - 0.652) < 0.03
    # alpha byte is always 0 (:1051)
    assert np.all((px >> 24) == 0)


def test_floor_has_no_self_hit_artefact_at_default_camera(default_720):
    """A.12: with exact reference arithmetic the floor mirror never re-hits the floor at the default camera, so floor
    pixels whose reflection leaves the scene keep their ambient term: no floor pixel is black."""
    sc, cam, r = default_720
    floor = r["pixels"][r["aov_id"] == 3]
    assert np.all(floor != 0)
    assert np.all((floor & 255) >= 21)       # ambient 43/255 * Ka 0.5 -> floor(21.5)


def test_ray_accounting(default_720):
    sc, cam, r = default_720
    c = r["counters"]
    assert c["primary"] == 1280 * 720
    assert c["shadow"] == 2 * (c["shade_diffuse"] // 2) == c["shade_diffuse"]
    assert c["sphere_tests"] == 3 * (c["primary"] + c["secondary"] + c["shadow"])
    assert c["plane_tests"] == c["primary"] + c["secondary"]
    rays_per_pixel = (c["primary"] + c["shadow"] + c["secondary"]) / c["primary"]
    assert 2.5 < rays_per_pixel < 3.0


def test_depth_cap_semantics(built):
    """:734 / :843 — a plane hit beyond the cap is WHITE, a sphere hit beyond the cap is black; cap 0 still traces
    one secondary ray (bounce 1 > 0 terminates at the NEXT hit)."""
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=160, height=90)
    r0 = O.render(sc, cam, 160, 90, 0)
    r32 = O.render(sc, cam, 160, 90, 32)
    assert r0["counters"]["secondary"] > 0
    assert r0["counters"]["secondary"] <= r32["counters"]["secondary"]
    assert not np.array_equal(r0["pixels"], r32["pixels"])


def test_plane_is_always_tiled(built):
    """:289 — isTiled is forced true, so a plane passed with tiled = 0 renders the same checkerboard."""
    sc = scenes.default_scene()
    sc2 = scenes.default_scene()
    sc2.planes[0, 19] = 0.0
    cam = scenes.make_camera(width=96, height=54)
    assert np.array_equal(O.render(sc, cam, 96, 54)["pixels"], O.render(sc2, cam, 96, 54)["pixels"])


def test_pack_color(built):
    """ShiftColor :1046-1052."""
    assert O.pack_color(0.0, 0.0, 0.0) == 0
    assert O.pack_color(1.0, 1.0, 1.0) == 0xFFFFFF
    assert O.pack_color(2.0, -1.0, 0.5) == (255 << 16) | (0 << 8) | 127
    assert O.pack_color(float("nan"), 0.999, 43.0 / 255.0) == (0 << 16) | (254 << 8) | 43 or \
        O.pack_color(float("nan"), 0.999, 43.0 / 255.0) == (0 << 16) | (254 << 8) | 42
    assert O.pack_color(float("inf"), float("-inf"), 1e-9) == 255 << 16


def test_subset_render_matches_full(built):
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=200, height=120)
    full = O.render(sc, cam, 200, 120, 8)["pixels"].reshape(-1)
    idx = np.random.default_rng(7).choice(200 * 120, 500, replace=False).astype(np.int32)
    sub = O.render(sc, cam, 200, 120, 8, subset=idx)["pixels"]
    assert np.array_equal(sub, full[idx])


def test_secondary_fold_is_order_dependent(built):
    """:804-805 compares the OFFSET distance with the stored un-offset one: of two spheres whose hits are < 0.01 apart
    the LATER one in array order wins, whichever is nearer."""
    mk = lambda z: scenes.sphere((0, 0, z), 1.0, scenes.mat_diffuse((1, 1, 1)))
    rays = np.array([[0, 0, 0, 0, 0, 1]], dtype=np.float32)
    ids, ts = O.query_spheres(np.stack([mk(5.0), mk(5.005)]), rays, 1)
    assert ids[0] == 1 and abs(ts[0] - 4.005) < 1e-5
    ids, ts = O.query_spheres(np.stack([mk(5.005), mk(5.0)]), rays, 1)
    assert ids[0] == 1 and abs(ts[0] - 4.0) < 1e-5
    # the primary fold (:977, strict '>') keeps the truly nearest, first index on ties
    ids, ts = O.query_spheres(np.stack([mk(5.0), mk(5.005)]), rays, 0)
    assert ids[0] == 0
    ids, ts = O.query_spheres(np.stack([mk(5.0), mk(5.0)]), rays, 0)
    assert ids[0] == 0


def test_sphere_invisible_from_inside_and_shadow_semantics(built):
    """:632-635 both roots must be positive; :574 the shadow ray's direction is the light POSITION."""
    s = np.stack([scenes.sphere((0, 0, 0), 2.0, scenes.mat_diffuse((1, 1, 1)))])
    ids, _ = O.query_spheres(s, np.array([[0, 0, 0, 0, 0, 1]], np.float32), 0)
    assert ids[0] == -1
    # origin (0,0,-5), "direction" = light position (0,0,3): hits the sphere although the light is at z=3 "behind" it
    ids, _ = O.query_spheres(s, np.array([[0, 0, -5, 0, 0, 3]], np.float32), 2)
    assert ids[0] == 1
    ids, _ = O.query_spheres(s, np.array([[0, 0, -5, 0, 0, -3]], np.float32), 2)
    assert ids[0] == 0
    # zero direction: a = 0 -> 0/0 = NaN -> miss (no crash)
    ids, _ = O.query_spheres(s, np.array([[0, 0, -5, 0, 0, 0]], np.float32), 2)
    assert ids[0] == 0


@pytest.mark.parametrize("name", golden_names())
def test_golden_fixtures(built, name):
    g = load_golden(name)
    sc = GOLDEN_SCENES[name]()
    w, h = int(g["w"]), int(g["h"])
    r = O.render(sc, g["cam"], w, h, int(g["depth"]), int(g["spp"]), int(g["seed"]), want_hash=True, want_aov=True)
    assert np.array_equal(r["pixels"], g["pixels"])
    assert np.array_equal(r["hash"], g["hash"])
    assert np.array_equal(r["aov_id"], g["aov_id"])
    assert np.array_equal(r["aov_t"].view(np.uint32), g["aov_t"].view(np.uint32))
    assert [r["counters"][k] for k in O.COUNTER_NAMES[:10]] == [int(v) for v in g["counters"]]


def test_normalize_variant_sensitivity(built):
    """OpenTK's Normalize is restated as reciprocal-multiply; the true-division variant must stay within the
    image tolerance (so the unpinned choice cannot move the picture visibly)."""
    from common import channels
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=320, height=180)
    a = O.render(sc, cam, 320, 180, 32)["pixels"]
    b = O.render(sc, cam, 320, 180, 32, variant="truediv")["pixels"]
    d = np.abs(channels(a) - channels(b)).max(-1)
    assert (d > 1).mean() < 0.02
