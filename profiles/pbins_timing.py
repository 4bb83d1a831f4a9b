"""RT_OPT_PRIMARY_BINS off / on, BASELINE configs[2] and configs[3] at 4K through the C ABI (headless rt_render, kernel time from the
library's CUDA events, which bracket everything a frame enqueues: refit of the camera-inflated boxes, bins build, render kernel).
usage: python profiles/pbins_timing.py [config3] [config4]   (RTB200_LIB selects the library)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import numpy as np
import rtb200
import scenes

W, H = 3840, 2160
which = sys.argv[1:] or ["config3", "config4"]
for name in which:
    sc = scenes.config3_scene() if name == "config3" else scenes.config4_scene()
    cam = scenes.make_camera(width=W, height=H, **scenes.SCALED_CAMERA)
    cam_b = scenes.make_camera(width=W, height=H, **dict(scenes.SCALED_CAMERA, yaw=1e-3))
    ctx = rtb200.Context([0])
    ctx.set_scene(sc, rtb200.RT_ACCEL_LBVH)
    rec = {"config": name, "lib": os.environ.get("RTB200_LIB", "default")}
    frames = {}
    for setting in (0, 1):
        ctx.set_option(rtb200.RT_OPT_PRIMARY_BINS, setting)
        for _ in range(2):
            ctx.render(cam, W, H, 8, headless=True)
        still = [ctx.render(cam, W, H, 8, headless=True)[1].kernel_ms for _ in range(7)]
        b0 = ctx.get_info(rtb200.RT_INFO_PRIMARY_BIN_BUILDS)
        moving = [ctx.render(cam_b if k % 2 == 0 else cam, W, H, 8, headless=True)[1].kernel_ms for k in range(8)]
        builds = ctx.get_info(rtb200.RT_INFO_PRIMARY_BIN_BUILDS) - b0
        frames[setting] = ctx.render(cam, W, H, 8)[0]
        rec["on" if setting else "off"] = {"kernel_ms": round(min(still), 4), "kernel_ms_all": [round(v, 4) for v in still],
                                           "kernel_ms_moving_camera": round(min(moving), 4), "bin_builds_while_moving": builds}
    rec["frames_equal"] = bool(np.array_equal(frames[0], frames[1]))
    print(json.dumps(rec), flush=True)
    ctx.close()
