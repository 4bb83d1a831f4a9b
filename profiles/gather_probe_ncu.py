"""One epoch of the packed gather per world size on ONE GPU (emulated ranks), meant to run under
`ncu --metrics gpu__time_duration.sum`: the launch list then gives the duration of every piece (peer pack kernel, rank 0's own
render kernel, the expand pass) without host effects."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))
import rtb200 as rt, scenes

W, H, F = 3840, 2160, 16
sc = scenes.default_scene()
cams = np.repeat(scenes.make_camera(width=W, height=H)[None], F, 0)
base = rt.Context([0]); base.set_scene(sc)
fb = base.dev_alloc(F * W * H * 4)
nbytes = base.gather_bytes(W, H)
area = base.dev_alloc(nbytes)
for world, sink, peer in ((8, 1, 2), (4, 4, 5)):
    base.dev_memset(area, 0, nbytes)
    ranks = []
    for r in range(world):
        c = rt.Context([0]); c.set_scene(sc); c.set_partition(r, world, 8)
        c.set_option(rt.RT_OPT_SHARED_TARGET, 1); c.set_option(rt.RT_OPT_GATHER_MODE, 2)
        c.set_option(rt.RT_OPT_SINK_TILES, sink); c.set_option(rt.RT_OPT_PEER_TILES, peer)
        c.gather_attach(area, nbytes)
        ranks.append(c)
    for rep in range(2):
        for r in list(range(1, world)) + [0]:
            ranks[r].render_device(cams, W, H, 8, 1, 0, fb); ranks[r].sync()
    plain = rt.Context([0]); plain.set_scene(sc); plain.set_partition(1, world, 8)
    for rep in range(2):
        plain.render_device(cams, W, H, 8, 1, 0, fb); plain.sync()
    plain.close()
    for c in ranks: c.close()
for rep in range(2):
    base.render_device(cams, W, H, 8, 1, 0, fb); base.sync()
