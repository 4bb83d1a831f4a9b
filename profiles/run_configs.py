"""Measures every BASELINE.json config on one GPU through the C ABI (headless kernel time, best of N) and prints one JSON
object per config. Ray counts come from the instrumented kernel (nearest-first accounting).
usage: python profiles/run_configs.py [out.json] [--skip-brute3]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import numpy as np
import rtb200
import scenes


N_DEV = 1


def measure(name, sc, cam, w, h, depth, spp, accel, reps=5, count=True):
    ctx = rtb200.Context(list(range(N_DEV)))
    t0 = time.perf_counter(); ctx.set_scene(sc, accel); t_scene = time.perf_counter() - t0
    if count:
        dbg = ctx.render_debug(cam, w, h, depth, spp, 1)
        c = dict(dbg["counters"]); rays = c["primary"] + c["shadow"] + c["secondary"]; c.update(dbg["lbvh"])
    else:
        c, rays = {}, 0
    ms = []
    for i in range(reps + 1):
        _, st = ctx.render(cam, w, h, depth, spp, 1, headless=True)
        if i: ms.append(st.kernel_ms)
    px, st = ctx.render(cam, w, h, depth, spp, 1)
    ctx.close()
    best = min(ms)
    out = dict(config=name, n_devices=N_DEV, width=w, height=h, depth=depth, spp=spp, accel=accel, n_spheres=len(sc.spheres), rays=rays,
               kernel_ms_best=best, kernel_ms_all=ms, mrays_per_s=(rays / best / 1e3) if rays else None,
               scene_upload_s=t_scene, d2h_ms=st.d2h_ms, counters=c, checksum=int(np.bitwise_xor.reduce(px.reshape(-1).astype(np.uint32))))
    print(json.dumps(out), flush=True)
    return out


def main():
    global N_DEV
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
    for a in sys.argv:
        if a.startswith("--devices="):
            N_DEV = int(a.split("=")[1])
    res = []
    d = scenes.default_scene()
    res.append(measure("1_default_1280x720_d32", d, scenes.make_camera(width=1280, height=720), 1280, 720, 32, 1, rtb200.RT_ACCEL_AUTO))
    res.append(measure("2_default_4k_d8", d, scenes.make_camera(width=3840, height=2160), 3840, 2160, 8, 1, rtb200.RT_ACCEL_AUTO))
    c3 = scenes.config3_scene(); cam4k = scenes.make_camera(width=3840, height=2160, **scenes.SCALED_CAMERA)
    if "--only45" in sys.argv:
        c4 = scenes.config4_scene(); cam4k = scenes.make_camera(width=3840, height=2160, **scenes.SCALED_CAMERA)
        res.append(measure("4_100kspheres_4k_lbvh", c4, cam4k, 3840, 2160, 8, 1, rtb200.RT_ACCEL_LBVH, count=(N_DEV == 1)))
        res.append(measure("5_default_8k_16spp", d, scenes.make_camera(width=7680, height=4320), 7680, 4320, 8, 16, rtb200.RT_ACCEL_AUTO, reps=3, count=(N_DEV == 1)))
        if out_path: json.dump(res, open(out_path, "w"), indent=1)
        return
    if "--skip-brute3" not in sys.argv:
        res.append(measure("3_1024spheres_4k_brute_staged", c3, cam4k, 3840, 2160, 8, 1, rtb200.RT_ACCEL_BRUTE, reps=2))
    res.append(measure("3_1024spheres_4k_lbvh", c3, cam4k, 3840, 2160, 8, 1, rtb200.RT_ACCEL_LBVH))
    c4 = scenes.config4_scene()
    res.append(measure("4_100kspheres_4k_lbvh", c4, cam4k, 3840, 2160, 8, 1, rtb200.RT_ACCEL_LBVH))
    res.append(measure("5_default_8k_16spp", d, scenes.make_camera(width=7680, height=4320), 7680, 4320, 8, 16, rtb200.RT_ACCEL_AUTO, reps=3))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
