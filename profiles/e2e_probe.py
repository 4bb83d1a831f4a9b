"""Breakdown of one rt_render call with a host buffer (the e2e path of bench.py): wall time per frame in Python, host time inside the
library (entry -> enqueued -> return), copy-stream time, kernel time, bytes copied. Knobs through the environment:
RTB200_BANDS, RTB200_RECT_MAX_FRAC, RTB200_FILL_THREADS, and --sparse 0/1 --precleared 0/1."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import rtb200, scenes, torch

ap = argparse.ArgumentParser()
ap.add_argument("--sparse", type=int, default=1)
ap.add_argument("--precleared", type=int, default=0)
ap.add_argument("--zero-copy", type=int, default=1)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--w", type=int, default=3840)
ap.add_argument("--h", type=int, default=2160)
a = ap.parse_args()
W, H = a.w, a.h
ctx = rtb200.Context([0]); ctx.set_scene(scenes.default_scene())
ctx.set_option(rtb200.RT_OPT_SPARSE_D2H, a.sparse); ctx.set_option(rtb200.RT_OPT_HOST_PRECLEARED, a.precleared)
ctx.set_option(rtb200.RT_OPT_HOST_ZERO_COPY, a.zero_copy)
px = float(scenes.view_params(W, H)[0]) / W / float(scenes.NEAR_CLIP)
cams = [rtb200.to_rt_camera(scenes.make_camera(yaw=(k - 8) * px, width=W, height=H)) for k in range(16)]
host = torch.empty((16, H, W), dtype=torch.int32, pin_memory=True).numpy()
frames = [host[f] for f in range(16)]
for f in range(16):
    ctx.render(cams[f], W, H, 8, out=frames[f])
rec = dict(wall=[], total=[], enq=[], fillw=[], d2h=[], kern=[], bytes=[])
t_all0 = time.perf_counter()
for i in range(a.frames):
    f = i % 16
    t0 = time.perf_counter()
    _, st = ctx.render(cams[f], W, H, 8, out=frames[f])
    rec["wall"].append((time.perf_counter() - t0) * 1e3)
    rec["total"].append(ctx.get_info(rtb200.RT_INFO_LAST_TOTAL_NS) / 1e6); rec["enq"].append(ctx.get_info(rtb200.RT_INFO_LAST_ENQUEUE_NS) / 1e6)
    rec["fillw"].append(ctx.get_info(rtb200.RT_INFO_LAST_FILL_WAIT_NS) / 1e6); rec["d2h"].append(st.d2h_ms); rec["kern"].append(st.kernel_ms)
    rec["bytes"].append(ctx.get_info(rtb200.RT_INFO_LAST_D2H_BYTES))
t_all = (time.perf_counter() - t_all0) * 1e3 / a.frames
out = {k: float(np.median(v)) for k, v in rec.items()}
out["loop_ms_per_frame"] = t_all
out["env"] = {k: os.environ.get(k) for k in ("RTB200_BANDS", "RTB200_RECT_MAX_FRAC", "RTB200_FILL_THREADS")}
out["sparse"] = a.sparse; out["precleared"] = a.precleared; out["zero_copy"] = a.zero_copy
print(json.dumps(out))
