"""CPU-oracle timings for every BASELINE.json config on this box's host cores (bounded samples; rays counted by the oracle).
usage: python profiles/cpu_configs.py [out.json]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import numpy as np
import oracle_lib as O
import scenes

cores = len(os.sched_getaffinity(0))
res = []


def run(name, sc, cam, w, h, depth, spp, mode, subset=None, reps=3, note=""):
    cnt = O.render(sc, cam, w, h, depth, spp, 1, mode="nearest", threads=cores, subset=subset)["counters"]
    rays = cnt["primary"] + cnt["shadow"] + cnt["secondary"]
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); O.render(sc, cam, w, h, depth, spp, 1, mode=mode, threads=cores, subset=subset); ts.append(time.perf_counter() - t0)
    out = dict(config=name, mode=mode, cores=cores, rays=rays, best_s=min(ts), mrays_per_s=rays / min(ts) / 1e6, sample=note)
    print(json.dumps(out), flush=True); res.append(out)


d = scenes.default_scene()
run("1_default_1280x720_d32", d, scenes.make_camera(width=1280, height=720), 1280, 720, 32, 1, "faithful", reps=5, note="full frame")
run("2_default_4k_d8", d, scenes.make_camera(width=3840, height=2160), 3840, 2160, 8, 1, "faithful", reps=5, note="full frame")
c3 = scenes.config3_scene()
run("3_1024spheres", c3, scenes.make_camera(width=960, height=540, **scenes.SCALED_CAMERA), 960, 540, 8, 1, "nearest", reps=2, note="960x540 (1/16 of the 4K pixels), brute force")
c4 = scenes.config4_scene()
idx = np.random.default_rng(7).choice(3840 * 2160, 65536, replace=False).astype(np.int32)
run("4_100kspheres", c4, scenes.make_camera(width=3840, height=2160, **scenes.SCALED_CAMERA), 3840, 2160, 8, 1, "nearest", subset=idx, reps=1, note="65,536-pixel subset (seed 7) of the 4K frame, brute force")
run("5_default_8k_16spp", d, scenes.make_camera(width=7680, height=4320), 7680, 4320, 8, 16, "faithful", subset=np.arange(0, 7680 * 4320, 64, dtype=np.int32), reps=2, note="every 64th pixel of the 8K frame, 16 spp")
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
