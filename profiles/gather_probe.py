"""Cost of the packed gather's pieces on ONE GPU (ranks emulated as contexts on device 0, so no link effects): for world W,
the 16-frame 4K launch of (a) a peer with the plain gather, (b) the same peer with the packed wire format, (c) rank 0's own share +
expand pass (planes already written). CUDA-event times in ms."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import rtb200 as rt, scenes, torch

W, H, F = 3840, 2160, 16
sc = scenes.default_scene()
cams = np.repeat(scenes.make_camera(width=W, height=H)[None], F, 0)
stream = torch.cuda.current_stream(); sh = stream.cuda_stream

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
base = rt.Context([0]); base.set_scene(sc)
fb = base.dev_alloc(F * W * H * 4)
nbytes = base.gather_bytes(W, H)
area = base.dev_alloc(nbytes); base.dev_memset(area, 0, nbytes)
out["single_gpu_16_frames_ms"] = timed(lambda: base.render_device(cams, W, H, 8, 1, 0, fb, sh))
for world, sink, peer in ((2, 1, 1), (4, 1, 1), (4, 4, 5), (8, 1, 1), (8, 1, 2)):
    for mode in (1, 2):
        base.dev_memset(area, 0, nbytes)
        ranks = []
        for r in range(world):
            c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, world, 8)
            c.set_option(rt.RT_OPT_SHARED_TARGET, 1); c.set_option(rt.RT_OPT_GATHER_MODE, mode)
            c.set_option(rt.RT_OPT_SINK_TILES, sink); c.set_option(rt.RT_OPT_PEER_TILES, peer)
            c.gather_attach(area, nbytes)
            ranks.append(c)
        plain = rt.Context([0]); plain.set_scene(sc); plain.set_partition(1, world, 8)
        t_plain = timed(lambda: plain.render_device(cams, W, H, 8, 1, 0, fb, sh))
        # one epoch at a time: peers, then rank 0 (its expand pass finds every flag set) — time each launch with events
        tp, t0 = [], []
        for rep in range(4):
            for r in range(1, world):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); ranks[r].render_device(cams, W, H, 8, 1, 0, fb, sh); e1.record(stream); torch.cuda.synchronize()
                if r == 1: tp.append(e0.elapsed_time(e1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ranks[0].render_device(cams, W, H, 8, 1, 0, fb, sh); e1.record(stream); torch.cuda.synchronize()
            t0.append(e0.elapsed_time(e1))
        assert ranks[0].get_info(rt.RT_INFO_GATHER_TIMEOUTS) == 0
        out["world%d_sink%d_peer%d_mode%d" % (world, sink, peer, mode)] = dict(peer_plain_ms=t_plain, peer_packed_ms=min(tp[1:]), rank0_render_plus_expand_ms=min(t0[1:]))
        for c in ranks: c.close()
        plain.close()
print(json.dumps(out, indent=1))
