"""Fuzz campaign of the per-frame gates of the tiny-scene path (csrc/rt_gate.cuh, shipped shapes) on the CPU, beyond the seeds of
tests/test_primary_gate.py: random small scenes, arbitrary plane positions / unit normals, random light vectors, cameras at any distance
and attitude, random frame shapes. Every piece of work a gate bit lets the kernel skip must be one the oracle's ray log shows fruitless
(test_primary_gate._check_bits_against_log), and the gated emulation of the kernel must render the oracle's frame.
  python tests/fuzz_frame_gates.py SEED CASES"""
import os
import sys
import time

_T = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _T); sys.path.insert(0, os.path.join(os.path.dirname(_T), 'uu-infogr-raytracer_b200'))
import numpy as np
import hostemu_lib as E
import oracle_lib as O
import scenes
from test_primary_gate import _check_bits_against_log


def run(seed0, ncases, verbose=True):
    bad = 0
    t0 = time.time()
    for case in range(ncases):
        rng = np.random.default_rng(seed0 * 100000 + case)
        sc = scenes.default_scene() if rng.integers(0, 3) == 0 else scenes.small_random_scene(int(rng.integers(1, 9)), int(rng.integers(0, 10**6)))
        if rng.integers(0, 2):
            n = rng.normal(size=3)
            sc.planes[0, 3:6] = (n / np.linalg.norm(n)).astype(np.float32)
            sc.planes[0, 0:3] = (rng.normal(size=3) * 3).astype(np.float32)
        if rng.integers(0, 2) and len(sc.lights):
            sc.lights[:, 0:3] = rng.uniform(-40, 40, (len(sc.lights), 3)).astype(np.float32)
        w, h = int(rng.integers(40, 200)), int(rng.integers(30, 130))
        pos = tuple(rng.normal(size=3) * 10.0 ** rng.uniform(-0.5, 1.7) + np.array([0, 1.0, -2.0]))
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.55, 1.55)), width=w, height=h)
        try:
            _check_bits_against_log(sc, cam, w, h)
            a = O.render(sc, cam, w, h, 4)
            b = E.render(sc, cam, w, h, 4, tiny=1)
            assert np.array_equal(a["pixels"], b["pixels"]), "pixels differ"
        except AssertionError as e:
            bad += 1
            print("BAD", seed0, case, str(e)[:200], flush=True)
        if verbose and case % 250 == 0:
            print("case", case, "%.0fs" % (time.time() - t0), flush=True)
    print("done seed", seed0, "cases", ncases, "bad", bad)
    return bad


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]))
