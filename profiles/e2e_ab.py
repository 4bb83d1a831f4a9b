"""In-process A/B of the two host-return paths of rt_render at several frame sizes: copy engine (band-pipelined, RT_OPT_HOST_ZERO_COPY 0)
vs the kernel's own PCIe stores (2), alternating, 4 rounds of 48 frames each; median ms per frame of every round."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import rtb200, scenes, torch

out = {}
for (W, H) in ((1280, 720), (1920, 1080), (2560, 1440), (3840, 2160), (7680, 4320)):
    ctx = rtb200.Context([0]); ctx.set_scene(scenes.default_scene())
    px = float(scenes.view_params(W, H)[0]) / W / float(scenes.NEAR_CLIP)
    cams = [rtb200.to_rt_camera(scenes.make_camera(yaw=(k - 8) * px, width=W, height=H)) for k in range(16)]
    host = torch.empty((4, H, W), dtype=torch.int32, pin_memory=True).numpy()
    frames = [host[f] for f in range(4)]
    res = {0: [], 2: []}
    for rnd in range(4):
        for mode in (0, 2):
            ctx.set_option(rtb200.RT_OPT_HOST_ZERO_COPY, mode)
            for i in range(8):
                ctx.render(cams[i % 16], W, H, 8, out=frames[i % 4])
            ts = []
            for i in range(48):
                t0 = time.perf_counter()
                ctx.render(cams[i % 16], W, H, 8, out=frames[i % 4])
                ts.append((time.perf_counter() - t0) * 1e3)
            res[mode].append(round(float(np.median(ts)), 4))
    out["%dx%d" % (W, H)] = {"copy_engine_ms": res[0], "zero_copy_ms": res[2]}
    ctx.close()
print(json.dumps(out))
