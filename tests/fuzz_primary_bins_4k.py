"""The per-pixel primary-bins check of tests/fuzz_primary_bins.py at the frame sizes of BASELINE.json (3840x2160, 7680x4320, and a
4096x1000 strip): 5 / 40 / 150 spheres, cameras looking at a silhouette.   python tests/fuzz_primary_bins_4k.py SEED CASES
(round 2: 1 560 cases, none bad; ~0.2 s per case on 8 cores)."""
import os
import sys
import time

_T = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _T); sys.path.insert(0, os.path.join(os.path.dirname(_T), 'uu-infogr-raytracer_b200'))
import numpy as np
import hostemu_lib as E
import scenes
def mk(c, r):
    return scenes.sphere(tuple(float(v) for v in c), float(r), scenes.mat_diffuse((1, 1, 1)))
bad = 0; t0 = time.time()
for case in range(int(sys.argv[2])):
    rng = np.random.default_rng(int(sys.argv[1]) * 1000 + case)
    n = int(rng.choice([5, 40, 150]))
    scale = float(rng.choice([1.0, 30.0, 300.0]))
    origin = rng.uniform(-1, 1, 3) * float(rng.choice([0.0, 50.0, 500.0]))
    sph = np.stack([mk(origin + rng.normal(size=3) * scale, float(rng.choice([1e-3, 0.02, 0.2, 1.0])) * scale) for _ in range(n)])
    cpos = origin + rng.normal(size=3) * scale * float(rng.choice([1.0, 3.0, 10.0]))
    k = rng.integers(0, n)
    tgt = sph[k, :3] + rng.normal(size=3) * sph[k, 3]
    d = tgt - cpos; d /= np.linalg.norm(d)
    pitch = float(-np.arcsin(np.clip(d[1], -0.9999, 0.9999))); yaw = float(np.arctan2(d[0], d[2]))
    w, h = [(3840, 2160), (7680, 4320), (4096, 1000)][rng.integers(0, 3)] if n <= 40 else (3840, 2160)
    cam = scenes.make_camera(width=w, height=h, pos=tuple(float(v) for v in cpos), yaw=yaw, pitch=pitch)
    r = E.primary_bins_check(sph, cam, w, h, shuffle=int(rng.integers(0, 2)))
    if r["missing"] or r["differ"]:
        bad += 1; print("BAD", case, r, flush=True)
    print(case, w, h, n, r, "%.0fs" % (time.time() - t0), flush=True)
print("done bad", bad)
