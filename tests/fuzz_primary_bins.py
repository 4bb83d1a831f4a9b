"""Fuzz campaign of the per-frame primary bins (csrc/rt_primary_bins.cuh) on the CPU: random scenes (clouds, carpets, huge + tiny mixes,
lines of spheres; scales 1 .. 1000, up to 1000 units from the origin), random cameras (80 % looking at / just past a sphere, 20 % next to
or inside one, any yaw / pitch), five frame shapes, host build or the replay of the device build. For every pixel the kernel's own primary
ray is tested against ALL spheres with the reference's test: every reported hit must be listed for the pixel's tile and the fold over the
list must select the reference's sphere (hostemu.cpp: emu_primary_bins_check).
  python tests/fuzz_primary_bins.py SEED CASES        (round 2: seeds 11-14 x 4000 cases + 12 400 cases of an earlier variant, 0 bad)
tests/test_primary_bins.py runs a short slice of it."""
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

def run(seed0, ncases, verbose=True):
    bad = 0
    nvalid = 0
    nbins = 0
    t0 = time.time()
    for case in range(ncases):
        rng = np.random.default_rng(seed0 * 100000 + case)
        n = int(rng.choice([3, 20, 100, 400, 1500]))
        scale = float(rng.choice([1.0, 10.0, 100.0, 1000.0]))
        origin = rng.uniform(-1, 1, 3) * float(rng.choice([0.0, 10.0, 100.0, 1000.0]))
        kind = rng.integers(0, 4)
        sph = []
        for i in range(n):
            if kind == 0:   # cloud
                c = origin + rng.normal(size=3) * scale; r = abs(rng.normal()) * scale * 0.1 + 1e-3
            elif kind == 1: # carpet
                c = origin + np.array([rng.uniform(-1, 1) * scale, rng.uniform(0, 0.02) * scale, rng.uniform(-1, 1) * scale]); r = rng.uniform(0.002, 0.03) * scale
            elif kind == 2: # mixed huge + tiny
                c = origin + rng.normal(size=3) * scale; r = float(rng.choice([1e-4, 1e-2, 0.3, 2.0])) * scale
            else:           # line of spheres away from origin
                d = rng.normal(size=3); d /= np.linalg.norm(d)
                c = origin + d * rng.uniform(0.1, 50) * scale + rng.normal(size=3) * 0.01 * scale; r = rng.uniform(0.01, 0.2) * scale
            sph.append(mk(c, r))
        sph = np.stack(sph)
        # camera: near the scene, looking roughly at it or randomly
        cpos = origin + rng.normal(size=3) * scale * float(rng.choice([0.1, 1.0, 3.0, 30.0]))
        if rng.random() < 0.2:
            cpos = sph[rng.integers(0, n), :3] + rng.normal(size=3) * sph[rng.integers(0, n), 3] * float(rng.choice([0.5, 1.0, 1.001, 1.5]))
        yaw = rng.uniform(-np.pi, np.pi); pitch = rng.uniform(-1.55, 1.55)
        if rng.random() < 0.8:                       # look at a sphere (its centre, its silhouette, or just past it)
            k = rng.integers(0, n)
            tgt = sph[k, :3] + rng.normal(size=3) * sph[k, 3] * float(rng.choice([0.0, 1.0, 1.0, 3.0]))
            d = tgt - cpos; L = np.linalg.norm(d)
            if L > 0:
                d /= L; pitch = float(-np.arcsin(np.clip(d[1], -0.9999, 0.9999))); yaw = float(np.arctan2(d[0], d[2]))
        w, h = [(160, 90), (213, 117), (384, 216), (64, 200), (500, 40)][rng.integers(0, 5)]
        cam = scenes.make_camera(width=w, height=h, pos=tuple(float(v) for v in cpos), yaw=float(yaw), pitch=float(pitch))
        r = E.primary_bins_check(sph, cam, w, h, shuffle=int(rng.integers(0, 3)))
        nvalid += r['valid']; nbins += r['by_bins'] > 0
        if r["missing"] or r["differ"]:
            bad += 1
            print("BAD", seed0, case, r, flush=True)
        if verbose and case % 1000 == 0:
            print("case", case, "n", n, "scale", scale, r, "%.0fs" % (time.time() - t0), flush=True)
    print("done seed", seed0, "cases", ncases, "bad", bad, "valid", nvalid, "with lists", nbins)
    return bad, nvalid, nbins


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]))
