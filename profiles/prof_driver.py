"""Minimal driver for ncu captures: N headless 4K frames of the bench workload (default scene, cap 8) through the C ABI.
usage: python profiles/prof_driver.py [frames_per_launch] [launches] [scene: default|config3|config4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import numpy as np
import rtb200
import scenes

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
which = sys.argv[3] if len(sys.argv) > 3 else "default"
W, H = 3840, 2160
if which == "default":
    sc, cam = scenes.default_scene(), scenes.make_camera(width=W, height=H)
elif which == "config3":
    sc, cam = scenes.config3_scene(), scenes.make_camera(width=W, height=H, **scenes.SCALED_CAMERA)
else:
    sc, cam = scenes.config4_scene(), scenes.make_camera(width=W, height=H, **scenes.SCALED_CAMERA)
ctx = rtb200.Context([0])
ctx.set_scene(sc)
cams = np.repeat(cam[None], F, 0)
for i in range(L):
    _, st = ctx.render_batch(cams, W, H, 8, headless=True)
    print("launch %d: %.3f ms for %d frame(s)" % (i, st.kernel_ms, F))
ctx.close()
