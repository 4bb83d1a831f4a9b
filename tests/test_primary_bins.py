"""Per-frame primary bins of the LBVH path (csrc/rt_primary_bins.cuh) against the reference's loop over ALL spheres for the primary
ray of every pixel (TracePixel, RayTracer.cs:975-981) — CPU only: the device query, the device scene policy (LbvhBinsScene) and the
host twin of the device build compiled as C++ (tests/hostemu).

Two kinds of checks:
 * direct (emu_primary_bins_check): the kernel's own primary ray of every pixel is tested against every sphere with the reference's
   test; each sphere that REPORTS a hit — including the hits the reference reports out of fp32 noise of its discriminant outside the
   exact sphere — must be in the list of the pixel's tile, and the fold over the list must select the same sphere
   with the same t bits as the fold over all spheres;
 * end to end: the frame / chain hashes / AOVs rendered through the binned policy equal the oracle's.
(Sensitivity, checked by hand while developing with -DRT_PB_MARGIN_PX / -DRT_BVH_PAD_K builds of tests/hostemu: rectangles shrunk
by 3 pixels lose 6 775 reported hits on the 384 x 216 config3 frame below, by 6 pixels 20 046 — the check sees an unsound rectangle;
with the margin at -1 pixel AND the noise pad removed it still reports none: the shipped rectangle has 2-3 pixels of slack on these
scenes, most of it the conservative direction bound eps of rt_gate.cuh (F2).)"""
import numpy as np
import pytest

import hostemu_lib as E
import oracle_lib as O
import scenes

LBVH, BINS = 3, 6          # hostemu policies


def mk(c, r, mat=None):
    return scenes.sphere(tuple(float(v) for v in c), float(r), mat if mat is not None else scenes.mat_diffuse((1, 1, 1)))


def check(spheres, cam, w, h, capacity=-1, want_valid=True, expect_lists=True):
    r = E.primary_bins_check(spheres, cam, w, h, capacity)
    assert r["missing"] == 0 and r["differ"] == 0, r
    assert r["by_bins"] + r["by_tree"] == w * h
    assert r["valid"] == (1 if want_valid else 0), r
    if expect_lists:
        assert r["by_bins"] > 0, r
    return r


CAMERAS = [dict(pos=(0.0, 3.0, -6.0), yaw=0.0, pitch=0.25), dict(pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0),
           dict(pos=(7.5, 1.25, 9.0), yaw=0.9, pitch=-0.2), dict(pos=(-11.0, 6.0, 30.0), yaw=2.8, pitch=0.6),
           dict(pos=(0.3, 0.2, 20.0), yaw=-1.3, pitch=0.05), dict(pos=(3.0, 40.0, 25.0), yaw=0.2, pitch=1.45)]


@pytest.mark.parametrize("ci", range(len(CAMERAS)))
def test_config3_scene_every_reported_hit_is_listed(built, ci):
    sc = scenes.config3_scene()
    w, h = 384, 216
    cam = scenes.make_camera(width=w, height=h, **CAMERAS[ci])
    r = check(sc.spheres, cam, w, h)
    assert r["by_bins"] >= 0.9 * w * h, r            # the bins, not the fallback, answer (lists of <= 32 spheres at this density)


def test_ragged_frame_sizes_and_single_tiles(built):
    sc = scenes.small_random_scene(60, 3)
    for (w, h) in [(1, 1), (7, 5), (8, 8), (9, 17), (33, 31), (130, 71)]:
        cam = scenes.make_camera(width=w, height=h, pos=(0.0, 1.0, -2.0), yaw=0.1, pitch=0.1)
        check(sc.spheres, cam, w, h, expect_lists=w * h > 64)      # a frame of one or two tiles may see more than 32 spheres per tile


def test_silhouettes_and_fp_noise_hits_far_away(built):
    """Small spheres 300-900 units away: the reference's discriminant reports hits for rays passing outside the exact sphere; the
    rectangle is built from the noise-inflated radius. Every pixel around every silhouette is checked by the exhaustive loop."""
    rng = np.random.default_rng(11)
    sph = np.stack([mk((rng.uniform(-300, 300), rng.uniform(0, 200), rng.uniform(300, 900)), rng.uniform(0.05, 0.6)) for _ in range(1500)])
    w, h = 640, 360
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 3.0, -6.0), yaw=0.0, pitch=-0.15)
    r = check(sph, cam, w, h)
    assert r["by_bins"] == w * h
    # far from the origin the rounding of vp - P grows (eps of rt_gate.cuh (F2) grows with |P|)
    cam = scenes.make_camera(width=w, height=h, pos=(400.0, 3.0, -250.0), yaw=0.0, pitch=-0.15)
    sph2 = sph.copy(); sph2[:, 0] += 400.0; sph2[:, 2] -= 250.0
    check(sph2, cam, w, h)


def test_eye_inside_spheres_huge_spheres_and_spheres_behind(built):
    rng = np.random.default_rng(2)
    sph = [mk((0, 0, 0), 2.0), mk((0.5, 0.2, 1.0), 3.0),           # the eye is inside both: tested for every pixel
           mk((0, 0, 40), 30.0),                                    # fills most of the frame
           mk((0, 0, -10), 3.0), mk((3, 1, -0.5), 1.0)]             # behind / across the eye plane
    sph += [mk((rng.uniform(-8, 8), rng.uniform(-3, 5), rng.uniform(3, 30)), rng.uniform(0.1, 1.0)) for _ in range(80)]
    sph = np.stack(sph)
    w, h = 320, 180
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0)
    r = check(sph, cam, w, h)
    assert r["everywhere"] == 2, r                             # 920 tiles: the huge sphere is simply in (almost) every list
    w2, h2 = 1280, 720                                          # 14 400 tiles: its rectangle exceeds PB_MAX_TILES -> tested for every pixel
    r = check(sph, scenes.make_camera(width=w2, height=h2, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0), w2, h2)
    assert r["everywhere"] == 3 and r["by_bins"] == w2 * h2, r
    # more "everywhere" spheres than the header holds: no bins this frame, everything traverses
    many = np.stack([mk((0.01 * i, 0, 0), 2.0 + 0.1 * i) for i in range(12)] + list(sph))
    r = check(many, cam, w, h, expect_lists=False)
    assert r["by_bins"] == 0 and r["everywhere"] >= 12, r


def test_crowded_tiles_and_a_full_list_array_fall_back_tile_by_tile(built):
    rng = np.random.default_rng(9)
    # 400 spheres behind one another in a narrow cone: the central tiles see far more than 32
    sph = np.stack([mk((rng.normal() * 0.2, rng.normal() * 0.2, 5 + 0.2 * i), 0.3) for i in range(400)] +
                   [mk((rng.uniform(-20, 20), rng.uniform(-5, 9), rng.uniform(8, 60)), rng.uniform(0.2, 0.8)) for _ in range(300)])
    w, h = 320, 180
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0)
    r = check(sph, cam, w, h)
    assert r["tiles_without_list"] > 0 and r["by_tree"] > 0 and r["by_bins"] > 0, r
    full = r["entries"]
    for cap in (0, 1, 37, full // 2, full - 1, full):
        r2 = check(sph, cam, w, h, capacity=cap, expect_lists=False)
        assert r2["entries"] <= cap
    assert check(sph, cam, w, h, capacity=full)["tiles_without_list"] == r["tiles_without_list"]


def test_non_finite_and_degenerate_records(built):
    rng = np.random.default_rng(4)
    sph = np.stack([mk((rng.uniform(-8, 8), rng.uniform(-3, 5), rng.uniform(3, 30)), rng.uniform(0.1, 1.0)) for _ in range(64)])
    sph[3, 0] = np.nan; sph[5, 17] = np.inf; sph[7, 2] = -np.inf; sph[9, 17] = np.nan; sph[11, 17] = -1.0; sph[13, 17] = 0.0
    sph[15, 1] = 1e30; sph[17, 17] = 1e-40; sph[19, 17] = 3e38
    w, h = 200, 120
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 1.0, -2.0), yaw=0.05, pitch=0.1)
    check(sph, cam, w, h)


def test_cameras_outside_the_derivation_build_no_bins(built):
    sc = scenes.small_random_scene(60, 3)
    w, h = 64, 48
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 1.0, -2.0))
    bad = cam.copy(); bad[3:6] *= 1.01                      # basis not orthonormal to 1e-5
    r = check(sc.spheres, bad, w, h, want_valid=False, expect_lists=False)
    assert r["by_bins"] == 0
    nanpos = cam.copy(); nanpos[0] = np.nan
    assert check(sc.spheres, nanpos, w, h, want_valid=False, expect_lists=False)["by_bins"] == 0
    far = cam.copy(); far[0:3] = (3e7, 0, 0)               # eps beyond 1e-2
    assert check(sc.spheres, far, w, h, want_valid=False, expect_lists=False)["by_bins"] == 0


@pytest.mark.parametrize("which", ["small", "config3", "mirrors_two_planes"])
def test_binned_policy_renders_the_oracles_frame(built, which):
    """LbvhBinsScene end to end: pixels, chain hashes (hit ids + t bits of every ray of the chain), primary AOVs and ray counters
    equal the oracle's; the primary rays no longer walk the tree (node_visits_primary ~ 0) while secondary rays still do."""
    if which == "small":
        sc, (w, h), camkw = scenes.small_random_scene(120, 8), (200, 120), dict(pos=(0.0, 1.5, -3.0), yaw=0.1, pitch=0.15)
    elif which == "config3":
        sc, (w, h), camkw = scenes.config3_scene(), (320, 180), scenes.SCALED_CAMERA
    else:
        sc, (w, h), camkw = scenes.small_random_scene(300, 21), (240, 136), dict(pos=(1.0, 2.0, 1.0), yaw=-0.4, pitch=0.3)
    cam = scenes.make_camera(width=w, height=h, **camkw)
    ref = O.render(sc, cam, w, h, 8, want_hash=True, want_aov=True)
    tree = E.render(sc, cam, w, h, 8, tiny=LBVH, debug=True)
    bins = E.render(sc, cam, w, h, 8, tiny=BINS, debug=True)
    for got in (tree, bins):
        assert np.array_equal(got["pixels"], ref["pixels"])
        assert np.array_equal(got["hash"], ref["hash"])
        assert np.array_equal(got["aov_id"], ref["aov_id"]) and np.array_equal(got["aov_t"].view(np.uint32), ref["aov_t"].view(np.uint32))
    names = ("primary", "shadow", "secondary", "sphere_tests", "sphere_disc_pos", "plane_tests", "shade_diffuse", "shade_specular", "shade_mirror", "shaded_hits")
    for k in (0, 1, 2, 5, 6, 7, 8, 9):
        assert bins["counters"][k] == ref["counters"][names[k]], names[k]
    assert tree["lbvh"][0] > 2 * w * h                        # the tree walk: several node visits per primary ray
    assert bins["lbvh"][0] < 0.3 * tree["lbvh"][0]            # the bins: only the crowded tiles (more than 32 spheres) still traverse
    assert bins["lbvh"][1] == tree["lbvh"][1] and bins["lbvh"][2] == tree["lbvh"][2]      # secondary / shadow traversal untouched
    # uninstrumented path too (NoDbg is what the render kernels instantiate)
    assert np.array_equal(E.render(sc, cam, w, h, 8, tiny=BINS)["pixels"], ref["pixels"])


def test_supersampled_frames_do_not_use_the_bins(built):
    """Jittered samples leave their pixel's rectangle margin: begin_pixel switches the bins off for spp != 1."""
    sc = scenes.small_random_scene(120, 8)
    w, h = 96, 54
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 1.5, -3.0), yaw=0.1, pitch=0.15)
    a = E.render(sc, cam, w, h, 8, spp=4, seed=5, tiny=LBVH, debug=True)
    b = E.render(sc, cam, w, h, 8, spp=4, seed=5, tiny=BINS, debug=True)
    assert np.array_equal(a["pixels"], b["pixels"]) and a["lbvh"] == b["lbvh"]
    assert np.array_equal(a["pixels"], O.render(sc, cam, w, h, 8, spp=4, seed=5)["pixels"])


def test_packed_slab_build_of_the_emulation_agrees(built):
    """The same frame through the library built with the packed-fp32 code paths emulated (RT_EMULATE_F32X2)."""
    sc = scenes.small_random_scene(120, 8)
    w, h = 160, 90
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 1.5, -3.0), yaw=0.1, pitch=0.15)
    ref = O.render(sc, cam, w, h, 8)["pixels"]
    E.use_variant("_packed")
    try:
        got = E.render(sc, cam, w, h, 8, tiny=BINS)["pixels"]
    finally:
        E.use_variant("")
    assert np.array_equal(got, ref)


# ---- the DEVICE build's algorithm, replayed on the host with its atomics resolved in random orders -----------------------------------
@pytest.mark.parametrize("seed", [1, 2, 0xC0FFEE])
def test_device_build_replay_is_order_independent(built, seed):
    """rt_primary_bins_build.cuh hands out list slots, "everywhere" slots and runs of the list array in whatever order its atomics
    resolve. Replayed on the host (hostemu.cpp: primary_bins_build_device_order — random sphere orders for the count and the fill pass,
    random order of k_pb_alloc's 32-tile groups, the kernel's own step -> tile arithmetic): every reported hit is listed, every fold
    selects the reference's sphere, and the lists hold as many entries in as many tiles as the sequential host build's."""
    sc = scenes.config3_scene()
    w, h = 384, 216
    for ci in (0, 2, 5):
        cam = scenes.make_camera(width=w, height=h, **CAMERAS[ci])
        seq = check(sc.spheres, cam, w, h)
        r = E.primary_bins_check(sc.spheres, cam, w, h, shuffle=seed)
        assert r["missing"] == 0 and r["differ"] == 0 and r["valid"] == 1, r
        for k in ("by_bins", "by_tree", "tiles_without_list", "entries", "everywhere", "sphere_tests"):
            assert r[k] == seq[k], (k, r, seq)


def test_device_build_replay_with_a_full_list_array_and_everywhere_spheres(built):
    """Which tiles lose their list when the array is full, and which slot an "everywhere" sphere takes, depends on the order — the
    answers must not: crowded tiles, a list array too small for the frame, the eye inside spheres, non-finite records."""
    rng = np.random.default_rng(9)
    sph = np.stack([mk((rng.normal() * 0.2, rng.normal() * 0.2, 5 + 0.2 * i), 0.3) for i in range(400)] +
                   [mk((rng.uniform(-20, 20), rng.uniform(-5, 9), rng.uniform(8, 60)), rng.uniform(0.2, 0.8)) for _ in range(300)] +
                   [mk((0.0, 0.0, 0.0), 1.5), mk((0.2, 0.1, -0.3), 3.0), mk((0.0, -50.0, 0.0), 49.0)])
    sph[17, 0] = np.nan; sph[23, 17] = np.inf
    w, h = 320, 180
    cam = scenes.make_camera(width=w, height=h, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0)
    seq = check(sph, cam, w, h)
    assert seq["everywhere"] >= 2 and seq["tiles_without_list"] > 0, seq
    for seed in (3, 4, 5):
        r = E.primary_bins_check(sph, cam, w, h, shuffle=seed)
        assert r["missing"] == 0 and r["differ"] == 0, r
        assert r["entries"] == seq["entries"] and r["everywhere"] == seq["everywhere"]
        for cap in (0, 1, 37, seq["entries"] // 2, seq["entries"] - 1):
            r2 = E.primary_bins_check(sph, cam, w, h, capacity=cap, shuffle=seed)
            assert r2["missing"] == 0 and r2["differ"] == 0 and r2["entries"] <= cap, r2


def test_device_build_replay_renders_the_oracles_frame(built):
    """The whole LbvhBinsScene policy over bins built in device order: pixels and chain hashes equal the oracle's."""
    sc, (w, h) = scenes.config3_scene(), (320, 180)
    cam = scenes.make_camera(width=w, height=h, **scenes.SCALED_CAMERA)
    ref = O.render(sc, cam, w, h, 8, want_hash=True)
    try:
        for seed in (7, 8):
            E.set_primary_bins_shuffle(seed)
            got = E.render(sc, cam, w, h, 8, tiny=BINS, debug=True)
            assert np.array_equal(got["pixels"], ref["pixels"]) and np.array_equal(got["hash"], ref["hash"])
    finally:
        E.set_primary_bins_shuffle(0)


def test_fuzz_slice(built):
    """A short slice of tests/fuzz_primary_bins.py (the campaign itself: 28 000 cases in round 2, none bad)."""
    import fuzz_primary_bins
    bad, valid, with_lists = fuzz_primary_bins.run(5, 150, verbose=False)
    assert bad == 0 and with_lists >= 60
