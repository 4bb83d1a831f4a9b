"""The DEVICE trace code (csrc/rt_trace.cuh, rt_scene.cuh) compiled as plain C++ must equal the oracle bit for bit:
pixels, per-pixel chain hash (every hit id / t bits / shadow result), primary-hit AOVs and ray counters.
This checks the kernel's logic on a box without a GPU; the -m gpu tests check the compiled kernels themselves."""
import numpy as np
import pytest

import hostemu_lib as E
import oracle_lib as O
import scenes
from common import GOLDEN_SCENES, golden_names, load_golden


def _compare(sc, cam, w, h, depth, spp=1, seed=0, tiny=False):
    a = O.render(sc, cam, w, h, depth, spp, seed, want_hash=True, want_aov=True)
    b = E.render(sc, cam, w, h, depth, spp, seed, tiny=tiny, debug=True)
    c = E.render(sc, cam, w, h, depth, spp, seed, tiny=tiny, debug=False)
    assert np.array_equal(a["pixels"], b["pixels"])
    assert np.array_equal(a["pixels"], c["pixels"])          # fast path (early-outs enabled) == instrumented path
    assert np.array_equal(a["hash"], b["hash"])
    assert np.array_equal(a["aov_id"], b["aov_id"])
    assert np.array_equal(a["aov_t"].view(np.uint32), b["aov_t"].view(np.uint32))
    assert [a["counters"][k] for k in O.COUNTER_NAMES[:10]] == b["counters"]


@pytest.mark.parametrize("tiny", [0, 1, 2])
@pytest.mark.parametrize("camkw,depth", [(dict(), 32), (dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 8),
                                         (dict(pos=(-2.0, 2.5, 3.0), yaw=-0.4, pitch=0.5), 0), (dict(pos=(0, 4.0, 6.0), pitch=1.2), 32)])
def test_default_scene(built, tiny, camkw, depth):
    _compare(scenes.default_scene(), scenes.make_camera(width=320, height=180, **camkw), 320, 180, depth, tiny=tiny)


def test_supersampling(built):
    _compare(scenes.default_scene(), scenes.make_camera(pos=(-2, 2.5, 3), yaw=-0.4, pitch=0.5, width=128, height=96), 128, 96, 3, spp=4, seed=7, tiny=2)


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 2), (2, 7), (3, 8), (4, 9), (7, 5), (8, 6), (12, 1), (16, 2), (40, 3), (200, 4)])
def test_random_scenes(built, n, seed):
    sc = scenes.small_random_scene(n, seed)
    cam = scenes.make_camera(pos=(0, 1.5, -4.0), pitch=0.1, width=160, height=100)
    _compare(sc, cam, 160, 100, 8, tiny=(2 if n <= 4 else 1 if n <= 8 else 0))


def test_empty_scene(built):
    sc = scenes.Scene(np.zeros((0, 18), np.float32), np.zeros((0, 20), np.float32), np.zeros((0, 4), np.float32), scenes.REF_AMBIENT)
    cam = scenes.make_camera(width=33, height=17)
    r = E.render(sc, cam, 33, 17, 32, tiny=True)
    assert np.all(r["pixels"] == 0)
    _compare(sc, cam, 33, 17, 32, tiny=True)


def test_no_lights_and_odd_sizes(built):
    sc = scenes.default_scene()
    sc = scenes.Scene(sc.spheres, sc.planes, np.zeros((0, 4), np.float32), sc.ambient)
    _compare(sc, scenes.make_camera(width=37, height=23), 37, 23, 4, tiny=True)
    _compare(scenes.default_scene(), scenes.make_camera(width=1, height=1), 1, 1, 4, tiny=False)


@pytest.mark.parametrize("name", golden_names())
def test_golden(built, name):
    g = load_golden(name)
    sc = GOLDEN_SCENES[name]()
    r = E.render(sc, g["cam"], int(g["w"]), int(g["h"]), int(g["depth"]), int(g["spp"]), int(g["seed"]),
                 tiny=len(sc.spheres) <= 8, debug=True)
    assert np.array_equal(r["pixels"], g["pixels"])
    assert np.array_equal(r["hash"], g["hash"])


@pytest.mark.parametrize("camkw,depth", [(dict(), 32), (dict(pos=(-2.0, 1.2, 5.0), yaw=-0.3, pitch=0.2), 8), (dict(pos=(0, 0.4, 2.0)), 2)])
def test_chain_park_and_resume(built, camkw, depth):
    """trace_chain is resumable: parking a chain at its third ray and finishing it later (what the compacting kernel does with
    deep mirror chains) gives the same pixels, hashes and counters as following it in one go."""
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=240, height=135, **camkw)
    a = O.render(sc, cam, 240, 135, depth, want_hash=True)
    b = E.render(sc, cam, 240, 135, depth, tiny=5, debug=True)
    assert np.array_equal(a["pixels"], b["pixels"]) and np.array_equal(a["hash"], b["hash"])
    assert [a["counters"][k] for k in O.COUNTER_NAMES[:10]] == b["counters"]
    sc2 = scenes.small_random_scene(8, 1)         # several mirror classes incl. DiffuseMirror
    cam2 = scenes.make_camera(pos=(0, 1.5, -4.0), pitch=0.1, width=160, height=100)
    assert np.array_equal(O.render(sc2, cam2, 160, 100, 8)["pixels"], E.render(sc2, cam2, 160, 100, 8, tiny=5)["pixels"])


@pytest.mark.parametrize("policy", [0, 1, 3, 4, 5])
@pytest.mark.parametrize("camkw", [dict(), dict(pos=(0.0, 0.0, 6.0)), dict(pos=(0.5, -1.0, 2.0), yaw=0.1, pitch=-0.2)])
def test_degenerate_scene(built, policy, camkw):
    """scenes.degenerate_scene(): zero / negative radiusSquared, light at the origin, zero-normal plane, camera inside a sphere
    and on the floor, negative colours, general pow exponents, far sphere, overlapping spheres — every scene policy."""
    sc = scenes.degenerate_scene()
    w, h = 160, 96
    cam = scenes.make_camera(width=w, height=h, **camkw)
    a = O.render(sc, cam, w, h, 8, want_hash=True)
    assert np.array_equal(a["pixels"], O.render(sc, cam, w, h, 8, mode="faithful")["pixels"])
    b = E.render(sc, cam, w, h, 8, tiny=policy, debug=True)
    assert np.array_equal(a["pixels"], b["pixels"]) and np.array_equal(a["hash"], b["hash"])


@pytest.mark.parametrize("tiny", [11, 12])
@pytest.mark.parametrize("camkw,depth", [(dict(), 32), (dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 8),
                                         (dict(pos=(0, 4.0, 6.0), pitch=1.2), 8), (dict(pos=(0.0, 0.5, 0.0), pitch=-0.6), 3)])
def test_production_debug_policy(built, tiny, camkw, depth):
    """ProdDbg (events only) on the GATED tiny-scene path — the logic of k_debug_tiny_prod: work a gate skips must be reported as the
    event the reference produces there, so chain hashes, primary AOVs and ray counters still equal the oracle's."""
    sc = scenes.default_scene()
    w, h = 320, 180
    cam = scenes.make_camera(width=w, height=h, **camkw)
    a = O.render(sc, cam, w, h, depth, want_hash=True, want_aov=True)
    b = E.render(sc, cam, w, h, depth, tiny=tiny, debug=True)
    assert np.array_equal(a["pixels"], b["pixels"])
    assert np.array_equal(a["hash"], b["hash"])
    assert np.array_equal(a["aov_id"], b["aov_id"])
    assert np.array_equal(a["aov_t"].view(np.uint32), b["aov_t"].view(np.uint32))
    assert [a["counters"][k] for k in ("primary", "shadow", "secondary")] == b["counters"][:3]
    g = E.gates(sc, cam, w, h)["bits"]
    assert (g != 0).mean() > 0.2          # the gates really were active on this frame


@pytest.mark.parametrize("kind,ppt,w,h,tile_rows", [
    (0, 4, 3840, 2160, 8), (0, 4, 1280, 720, 8), (0, 4, 1000, 563, 8), (0, 4, 1284, 397, 16), (0, 4, 4, 1, 8), (0, 4, 68, 19, 24),
    (0, 1, 3840, 2160, 8), (0, 1, 1283, 397, 16), (0, 1, 250, 131, 8), (0, 1, 37, 23, 8), (0, 1, 1, 1, 8), (0, 1, 17, 9, 32),
    (1, 4, 3840, 2160, 8), (1, 4, 1280, 724, 8), (1, 4, 1152, 333, 16), (1, 4, 128, 3, 8), (1, 4, 896, 500, 24)])
def test_2d_pixel_blocks_cover_every_pixel_once(kind, ppt, w, h, tile_rows):
    """csrc/rt_tiles.cuh — the thread -> pixel mapping of the render kernels (render_loop: 16 ppt x 8-pixel CTAs; the packed gather's
    render kernel: 128 x 4): over all tiles, items and threads every pixel of the frame is owned exactly once, no span leaves its
    tile, its row or the frame — for ragged widths / heights and several tile heights."""
    cover, bad = E.tile_cover(kind, ppt, w, h, tile_rows)
    assert bad == 0
    assert cover.min() == 1 and cover.max() == 1, (int(cover.min()), int(cover.max()))


@pytest.fixture
def packed():
    """libhostemu_packed.so: the device headers with -DRT_EMULATE_F32X2 — the PACKED fp32 code paths (two spheres, two lights, two bin
    entries, two child boxes per pass), each packed operation emulated as two scalar IEEE operations."""
    E.use_variant("_packed")
    try:
        yield
    finally:
        E.use_variant("")


def _scene_with_lights(n_spheres, n_lights, seed):
    sc = scenes.small_random_scene(n_spheres, seed)
    rng = np.random.default_rng(seed)
    lights = np.stack([scenes.light((rng.uniform(-6, 6), rng.uniform(2, 8), rng.uniform(-4, 10)), float(rng.uniform(0.4, 1.2))) for _ in range(n_lights)])
    return scenes.Scene(sc.spheres, sc.planes[:1], lights, sc.ambient)      # one plane: the exact-count instantiations


@pytest.mark.parametrize("ns,nl,seed", [(3, 2, 1), (2, 2, 2), (4, 4, 3), (4, 2, 4), (1, 4, 5), (3, 3, 6), (0, 2, 7)])
def test_packed_sphere_and_light_pairs_equal_the_oracle(built, packed, ns, nl, seed):
    """The production code path of the exact-count kernels — sphere_pair_bd, shade_light_pair (even light counts), xy-packed vector
    products, frame gates — as the shipped policy (no per-test counters: tiny = 2) and as the events-only policy of k_debug_tiny_prod
    (tiny = 12): pixels and chain hashes must equal the oracle's. Odd light counts take the scalar loop inside the same build."""
    sc = _scene_with_lights(ns, nl, seed)
    for camkw in (dict(pos=(0, 1.5, -4.0), pitch=0.1), dict(pos=(2.0, 3.0, 1.0), yaw=-0.5, pitch=0.6)):
        cam = scenes.make_camera(width=200, height=120, **camkw)
        a = O.render(sc, cam, 200, 120, 8, want_hash=True)
        c = E.render(sc, cam, 200, 120, 8, tiny=2, debug=False)
        assert np.array_equal(a["pixels"], c["pixels"]), "%d pixels differ" % (a["pixels"] != c["pixels"]).sum()
        p = E.render(sc, cam, 200, 120, 8, tiny=12, debug=True)
        assert np.array_equal(a["pixels"], p["pixels"]) and np.array_equal(a["hash"], p["hash"])


@pytest.mark.parametrize("camkw,depth", [(dict(), 32), (dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 8), (dict(pos=(0, 4.0, 6.0), pitch=1.2), 32)])
def test_packed_default_scene(built, packed, camkw, depth):
    sc = scenes.default_scene()
    cam = scenes.make_camera(width=320, height=180, **camkw)
    a = O.render(sc, cam, 320, 180, depth, want_hash=True)
    assert np.array_equal(a["pixels"], E.render(sc, cam, 320, 180, depth, tiny=2)["pixels"])
    p = E.render(sc, cam, 320, 180, depth, tiny=12, debug=True)
    assert np.array_equal(a["pixels"], p["pixels"]) and np.array_equal(a["hash"], p["hash"])
    d = scenes.degenerate_scene()                 # NaN / zero / negative records through the run-time-count packed-free path
    assert np.array_equal(O.render(d, cam, 320, 180, 8)["pixels"], E.render(d, cam, 320, 180, 8, tiny=1)["pixels"])


def test_packed_lbvh_paths(built, packed):
    """LBVH policy in the packed build: two-child slab test (box_entry2) and the shadow bins' pair loop (sphere_pair_bd on GridPair)."""
    sc = scenes.Scene(scenes.random_spheres_scene(600, 11, 20.0, 4.0, 44.0, "m")[0], scenes.default_scene().planes, scenes.config3_scene().lights, scenes.REF_AMBIENT)
    w, h = 192, 108
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    a = O.render(sc, cam, w, h, 8, want_hash=True)
    b = E.render(sc, cam, w, h, 8, tiny=3, debug=False)       # NoDbg: packed bins, early-outs
    assert np.array_equal(a["pixels"], b["pixels"]), "%d pixels differ" % (a["pixels"] != b["pixels"]).sum()
    c = E.render(sc, cam, w, h, 8, tiny=3, debug=True)        # FullDbg: scalar bins, packed slab test
    assert np.array_equal(a["pixels"], c["pixels"]) and np.array_equal(a["hash"], c["hash"])
