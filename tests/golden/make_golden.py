"""Generates tests/golden/*.npz from the CPU oracle (oracle/rt_oracle.cpp).

The reference holds no golden vectors for this path and cannot be executed here (no .NET), so these fixtures pin the
ORACLE's output (regression anchor + GPU parity target), not the C# program's: parity stays "unpinned" (DESIGN.md).
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O   # noqa: E402
import scenes            # noqa: E402

CASES = {
    # name: (scene factory, camera kwargs, w, h, depth, spp, seed)
    "default_160x90_d32": (scenes.default_scene, dict(), 160, 90, 32, 1, 0),
    "default_moved_192x108_d8": (scenes.default_scene, dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 192, 108, 8, 1, 0),
    "default_above_128x128_d3_spp4": (scenes.default_scene, dict(pos=(-2.0, 2.5, 3.0), yaw=-0.4, pitch=0.5), 128, 128, 3, 4, 7),
    "small12_seed1_160x100_d8": (lambda: scenes.small_random_scene(12, 1), dict(pos=(0, 1.5, -4.0), pitch=0.1), 160, 100, 8, 1, 0),
    "small40_seed3_160x100_d8": (lambda: scenes.small_random_scene(40, 3), dict(pos=(0, 1.5, -4.0), pitch=0.1), 160, 100, 8, 1, 0),
}


def main():
    for name, (mk, camkw, w, h, depth, spp, seed) in CASES.items():
        sc = mk()
        cam = scenes.make_camera(width=w, height=h, **camkw)
        r = O.render(sc, cam, w, h, depth, spp, seed, mode="nearest", want_hash=True, want_aov=True)
        f = O.render(sc, cam, w, h, depth, spp, seed, mode="faithful")
        assert np.array_equal(r["pixels"], f["pixels"]), name
        cnt = np.array([r["counters"][k] for k in O.COUNTER_NAMES[:10]], dtype=np.uint64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pixels=r["pixels"], hash=r["hash"], aov_id=r["aov_id"],
                            aov_t=r["aov_t"], counters=cnt, cam=cam, w=w, h=h, depth=depth, spp=spp, seed=seed)
        print(name, "ok", r["counters"]["primary"], r["counters"]["shadow"], r["counters"]["secondary"])


if __name__ == "__main__":
    main()
