"""Per-light shadow bins (csrc/rt_shadow_grid.cuh) against the reference's loop over ALL spheres (IntersectShadowLight,
RayTracer.cs:573-582) — CPU only: the device query and the host build compiled as C++ (tests/hostemu).

The bins may only ever hold a SUPERSET of the spheres a query can be reported to hit, including the hits the reference reports
out of fp32 noise of its discriminant far outside the exact sphere. Query points are drawn where that matters:
on the silhouette of a sphere as seen along the light vector, up to a scene diameter away, inside the noise band
rho^2 - r^2 ~ 1e-6 |oc|^2.  (Sensitivity, checked by hand while developing: with the noise pad removed from sg_disc the
far-tiny-cube case below reports 10 mismatches; with it, none.)"""
import os

import numpy as np
import pytest

import hostemu_lib as E
import scenes


def mk(c, r):
    return scenes.sphere(tuple(float(v) for v in c), float(r), scenes.mat_diffuse((1, 1, 1)))


def L(*p):
    return scenes.light(p, 1.0)


LIGHTS = np.stack([L(-20, 12, 10), L(3, 50, 1), L(40, 0.7, 5), L(0, 9, 4), L(-7, -3, 2), L(1e-3, 2e-3, 5e-4)])


def grazing_points(rng, sph, lights, n, ext):
    pts = []
    for _ in range(n):
        s = sph[rng.integers(len(sph))]; l = lights[rng.integers(len(lights))]
        c = s[0:3].astype(np.float64); r = np.sqrt(max(float(s[17]), 0.0))
        w = l[0:3].astype(np.float64); w = w / np.linalg.norm(w)
        perp = np.cross(w, rng.normal(size=3)); perp /= np.linalg.norm(perp)
        t = rng.uniform(0, ext) * rng.choice([1.0, 0.1, 0.01])
        if rng.uniform() < 0.5:      # inside the fp32 noise band of the discriminant
            rho = np.sqrt(r * r + 10 ** rng.uniform(-3, 0.3) * 1.0e-6 * (t * t + r * r))
        else:                        # on / just inside / just outside the exact silhouette
            rho = r * (1 + rng.choice([0.0, 1e-7, -1e-7, 1e-5, -1e-5, 1e-3, -1e-3, 3e-2, -3e-2, 0.3]) * rng.uniform(0, 1))
        pts.append(c - w * t + perp * rho)
    return np.array(pts, np.float32)


def check(sph, lights, rng, ext, npts):
    ok = np.isfinite(sph[:, 0:3]).all(1)
    lo = sph[ok, 0:3].min(0) - 2; hi = sph[ok, 0:3].max(0) + 2
    pts = np.concatenate([grazing_points(rng, sph[ok], lights, npts, ext), rng.uniform(lo, hi, (npts // 2, 3)).astype(np.float32)])
    r = E.shadow_bins_check(sph, lights, pts)
    assert r["mismatches"] == 0, r
    assert r["decided"] > 0.3 * (r["decided"] + r["undecided"]), r          # the bins, not the traversal fallback, were exercised
    return r


@pytest.fixture(scope="module")
def built():
    E.load()


@pytest.mark.parametrize("seed", [1, 2])
def test_bins_equal_the_loop_over_all_spheres(built, seed):
    rng = np.random.default_rng(seed)
    off = float(rng.choice([0, 300, 2000])); n = int(rng.choice([60, 500, 3000])); half = float(rng.choice([5, 50, 400]))
    sph = np.stack([mk((off + rng.uniform(-half, half), rng.uniform(-0.8, 1.5), off + rng.uniform(-half, half)), rng.uniform(0.03, 0.5)) for _ in range(n)])
    check(sph, LIGHTS, rng, 3 * half, 3000)                                   # flat carpet, possibly far from the origin
    sph = np.stack([mk((rng.uniform(-2, 2), rng.uniform(0, 600), rng.uniform(-2, 2) + 30), rng.uniform(0.05, 0.6)) for _ in range(400)])
    check(sph, LIGHTS, rng, 600, 3000)                                        # tall column
    sph = np.stack([mk(rng.normal(size=3) * 40 + (0, 0, 100), rng.choice([0.01, 0.3, 5.0])) for _ in range(800)])
    sph[5, 17] = 0.0; sph[6, 17] = -0.3; sph[7, 0] = np.nan                   # zero / negative radiusSquared, non-finite centre
    check(sph, LIGHTS, rng, 300, 3000)
    t = rng.uniform(0, 1, 700)                                                # thin slab along a diagonal, ~1 km out
    sph = np.stack([mk((600 * ti + rng.uniform(-1, 1), 2 * ti + rng.uniform(0, 1), 900 * ti + rng.uniform(-1, 1)), rng.uniform(0.02, 0.2)) for ti in t])
    check(sph, LIGHTS, rng, 1200, 3000)


def test_far_tiny_spheres_with_fine_cells(built):
    """Tiny spheres in a 200-unit cube with cells much smaller than the noise pad (64 cells per sphere): here a pad that is too
    small puts grazing query points into cells that do not list the sphere."""
    rng = np.random.default_rng(5)
    sph = np.stack([mk(rng.uniform(-100, 100, 3) + (0, 0, 150), rng.choice([0.02, 0.005, 0.1])) for _ in range(4000)])
    old = os.environ.get("RTB200_SG_CELLS_PER_SPHERE")
    os.environ["RTB200_SG_CELLS_PER_SPHERE"] = "64"
    try:
        r = check(sph, LIGHTS[:4], rng, 300, 30000)
    finally:
        if old is None: del os.environ["RTB200_SG_CELLS_PER_SPHERE"]
        else: os.environ["RTB200_SG_CELLS_PER_SPHERE"] = old
    assert r["occluded"] > 5000


def test_flat_scene_keeps_short_lists(built):
    """The pad follows the distance ALONG the light vector that query points inside the (thin) validity box can have, not the
    scene diameter: a carpet of 20 000 spheres over 134 x 134 units (the density of BASELINE configs[3]) scans a handful of
    spheres per query, not dozens."""
    sph, _ = scenes.random_spheres_scene(20000, 7, 67.0, 4.0, 138.0, "carpet")
    lights = np.stack([L(-20, 12, 10), L(20, 12, 10), L(-20, 12, 40), L(20, 12, 40)])
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(-67, 67, 4000), rng.choice([-1.0, 0.0, 0.5], 4000), rng.uniform(4, 138, 4000)], 1).astype(np.float32)
    r = E.shadow_bins_check(sph, lights, pts)
    assert r["mismatches"] == 0 and r["undecided"] == 0, r
    assert r["sphere_tests"] / r["decided"] < 8.0, r
