"""Default kernel vs the opt-in compacting variant (RT_OPT_COMPACTION) on the bench scene and on a mirror-heavy scene."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import numpy as np
import rtb200, scenes

W, H = 3840, 2160
d = scenes.default_scene()
mirrors = scenes.default_scene()
for i in range(3):
    mirrors.spheres[i, 4:17] = scenes.mat_mirror((0.9, 0.9, 0.9))
mirrors.spheres = np.concatenate([mirrors.spheres, scenes.sphere((0, 0.2, 4), 1.2, scenes.mat_mirror((0.95, 0.95, 0.95)))[None]])
for name, sc, depth, camkw in (("default scene, cap 8", d, 8, dict()), ("default scene, cap 32", d, 32, dict()),
                               ("4 mirror spheres + mirror floor, cap 32", mirrors, 32, dict(pos=(0, 0.5, -1.0)))):
    cam = scenes.make_camera(width=W, height=H, **camkw)
    ctx = rtb200.Context([0]); ctx.set_scene(sc)
    dbg = ctx.render_debug(cam, W, H, depth); c = dbg["counters"]; rays = c["primary"] + c["shadow"] + c["secondary"]
    res = {}
    for comp in (0, 1):
        ctx.set_option(rtb200.RT_OPT_COMPACTION, comp)
        ms = [ctx.render(cam, W, H, depth, headless=True)[1].kernel_ms for _ in range(6)][1:]
        res[comp] = min(ms)
        px, _ = ctx.render(cam, W, H, depth)
        assert np.array_equal(px, dbg["pixels"])
    print("%-42s rays/px %.2f  secondary/px %.2f | default %.3f ms (%.1f Grays/s) | compacting %.3f ms (%.1f Grays/s)" % (
        name, rays / (W * H), c["secondary"] / (W * H), res[0], rays / res[0] / 1e6, res[1], rays / res[1] / 1e6))
    ctx.close()
