import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'uu-infogr-raytracer_b200'))
import numpy as np
import oracle_lib as O, hostemu_lib as E, scenes, rtb200
def mk(c, r): return scenes.sphere(c, r, scenes.mat_diffuse((1, 1, 1)))
rng = np.random.default_rng(11)
n = 100000
sph = np.stack([mk((rng.uniform(-400, 400), rng.uniform(0, 30), rng.uniform(300, 900)), rng.uniform(0.05, 0.3)) for _ in range(3000)])
o = np.tile(np.array([[0, 3, -6]], np.float32), (n, 1))
k = rng.integers(0, len(sph), n)
v = sph[k, :3] - o; v /= np.linalg.norm(v, axis=1, keepdims=True)
perp = np.cross(v, rng.normal(size=(n, 3))); perp /= np.linalg.norm(perp, axis=1, keepdims=True)
tgt = sph[k, :3] + perp * sph[k, 3:4] * rng.uniform(0.9, 1.15, (n, 1))
d = (tgt - o); d /= np.linalg.norm(d, axis=1, keepdims=True)
d[: n // 2] *= rng.uniform(1e-3, 1e3, (n // 2, 1))
rays = np.concatenate([o, d.astype(np.float32)], 1)
dd = scenes.default_scene()
ctx = rtb200.Context([0]); ctx.set_scene(scenes.Scene(sph, dd.planes, dd.lights, dd.ambient), rtb200.RT_ACCEL_LBVH)
dn = np.linalg.norm(rays[:, 3:], axis=1)
for kind in (0, 1, 2):
    oi, ot = O.query_spheres(sph, rays, kind)
    gi, gt = ctx.query_spheres(rays, kind, rtb200.RT_ACCEL_LBVH)
    hi, ht = E.query(sph, rays, kind, 2)
    bad = np.nonzero((gi != oi) | (gt.view(np.uint32) != ot.view(np.uint32)))[0]
    print("kind", kind, "gpu-vs-oracle mismatches", len(bad), "hostemu-vs-oracle", int((hi != oi).sum()))
    if len(bad):
        b = bad[:8]
        print("  idx", b, "|d|", dn[b]); print("  oracle id", oi[b], "t", ot[b]); print("  gpu    id", gi[b], "t", gt[b])
        print("  |d| range of bad rays: min %.3g max %.3g ; frac with gpu==-1: %.3f" % (dn[bad].min(), dn[bad].max(), (gi[bad] < 0).mean()))
        print("  bad in scaled half:", int((bad < n // 2).sum()), "unscaled half:", int((bad >= n // 2).sum()))
