"""End-to-end rt_render (host buffer, page-locked) through the in-library multi-device context: frames/s and Grays/s for the
bench workload (default scene, 4K, cap 8), per-device PCIe return vs gather-on-GPU-0. usage: python profiles/e2e_multi.py <max_devices>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import numpy as np
import rtb200, scenes

maxd = int(sys.argv[1]) if len(sys.argv) > 1 else 1
W, H, DEPTH = 3840, 2160, 8
sc = scenes.default_scene(); cam = scenes.make_camera(width=W, height=H)
c0 = rtb200.Context([0]); c0.set_scene(sc)
cnt = c0.render_debug(cam, W, H, DEPTH); rays = sum(cnt["counters"][k] for k in ("primary", "shadow", "secondary")); ref = cnt["pixels"]; c0.close()
for n in (1, 2, 4, 8):
    if n > maxd: break
    ctx = rtb200.Context(list(range(n))); ctx.set_scene(sc)
    host = np.zeros((H, W), np.int32); ctx.host_register(host)
    for via0 in ((0, 1) if n > 1 else (0,)):
        ctx.set_option(rtb200.RT_OPT_HOST_VIA_GPU0, via0)
        for _ in range(3): ctx.render(cam, W, H, DEPTH, out=host)
        t0 = time.perf_counter()
        for _ in range(20): _, st = ctx.render(cam, W, H, DEPTH, out=host)
        dt = (time.perf_counter() - t0) / 20
        assert np.array_equal(host, ref)
        print("N=%d %-28s %.3f ms/frame  %.1f Grays/s e2e  (kernel %.3f ms, d2h %.3f ms)" % (
            n, "via GPU 0 (NVLink gather)" if via0 else "per-device PCIe return", dt * 1e3, rays / dt / 1e9, st.kernel_ms, st.d2h_ms))
    ctx.host_unregister(host); ctx.close()
