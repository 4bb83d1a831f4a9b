"""The whole LbvhBinsScene policy (LBVH traversal + per-light shadow bins + per-frame primary bins: the device headers compiled for the
host, tests/hostemu) against the oracle on random scenes: 60 .. 1 024 spheres (the small parity scenes with two planes and mirrors, or
carpets of three extents), random cameras and frame shapes, primary bins from the host build or the replay of the device build; pixels
and chain hashes (hit ids + t bits of every ray) must be equal.   python tests/fuzz_lbvh_policy.py SEED CASES
(round 2: 1 800 cases, none bad; ~0.08 s per case on 8 cores)."""
import os
import sys
import time

_T = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _T); sys.path.insert(0, os.path.join(os.path.dirname(_T), 'uu-infogr-raytracer_b200'))
import numpy as np
import hostemu_lib as E
import oracle_lib as O
import scenes
bad = 0; t0 = time.time()
for case in range(int(sys.argv[2])):
    rng = np.random.default_rng(int(sys.argv[1]) * 1000 + case)
    n = int(rng.choice([60, 150, 400, 1024]))
    if rng.random() < 0.5:
        sc = scenes.small_random_scene(n, int(rng.integers(1, 10**6)))
        pos = (float(rng.uniform(-4, 4)), float(rng.uniform(0.2, 5)), float(rng.uniform(-6, 2)))
    else:
        xh = float(rng.choice([6.0, 24.0, 80.0]))
        sph, _ = scenes.random_spheres_scene(n, int(rng.integers(1, 10**6)), xh, 4.0, 4.0 + 2 * xh, "f")
        d = scenes.config3_scene()
        sc = scenes.Scene(sph, d.planes, d.lights, d.ambient)
        pos = (float(rng.uniform(-xh, xh) * 0.5), float(rng.uniform(0.2, 8)), float(rng.uniform(-6, xh)))
    w, h = [(160, 90), (200, 120), (97, 61)][rng.integers(0, 3)]
    cam = scenes.make_camera(width=w, height=h, pos=pos, yaw=float(rng.uniform(-0.8, 0.8)), pitch=float(rng.uniform(-0.1, 0.6)))
    ref = O.render(sc, cam, w, h, 8, want_hash=True)
    E.set_primary_bins_shuffle(int(rng.integers(0, 3)))
    got = E.render(sc, cam, w, h, 8, tiny=6, debug=True)
    E.set_primary_bins_shuffle(0)
    ok = np.array_equal(got["pixels"], ref["pixels"]) and np.array_equal(got["hash"], ref["hash"])
    if not ok:
        bad += 1; print("BAD", case, int((got["pixels"] != ref["pixels"]).sum()), flush=True)
    if case % 25 == 0:
        print(case, n, w, h, "visits primary", got["lbvh"][0], "%.0fs" % (time.time() - t0), flush=True)
print("done bad", bad)
