"""ctypes binding of the CPU oracle (oracle/librt_oracle.so).  TEST INFRASTRUCTURE ONLY — the product
package never imports this module (see oracle/rt_oracle.cpp header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
sys.path.insert(0, os.path.join(ROOT, "uu-infogr-raytracer_b200"))

COUNTER_NAMES = ["primary", "shadow", "secondary", "sphere_tests", "sphere_disc_pos", "plane_tests",
                 "shade_diffuse", "shade_specular", "shade_mirror", "shaded_hits", "faithful_secondary", "faithful_shadow"]

_libs = {}


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def load(variant: str = ""):
    name = "librt_oracle%s.so" % (("_" + variant) if variant else "")
    if name in _libs:
        return _libs[name]
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    lib.orc_render.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                               C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                               C.POINTER(C.c_int32), fp, C.POINTER(C.c_int32), C.c_int]
    lib.orc_render.restype = C.c_int
    lib.orc_query_spheres.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.POINTER(C.c_int32), fp]
    lib.orc_query_spheres.restype = C.c_int
    lib.orc_ray_log.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_uint32), C.c_int, C.c_void_p, C.c_int]
    lib.orc_ray_log.restype = C.c_int
    lib.orc_pack_color.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.orc_pack_color.restype = C.c_int
    lib.orc_max_threads.restype = C.c_int
    _libs[name] = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None and a.size else None


def render(scene, cam, w, h, max_depth=32, spp=1, seed=0, mode="nearest", threads=0, subset=None,
           want_hash=False, want_aov=False, variant=""):
    """Returns dict(pixels=int32[h,w] (or [n_subset]), counters=dict, hash=uint32[...], aov_id, aov_t)."""
    lib = load(variant)
    n = int(len(subset)) if subset is not None else w * h
    pixels = np.zeros(n, dtype=np.int32)
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    hsh = np.zeros(n, dtype=np.uint32) if want_hash else None
    aid = np.zeros(n, dtype=np.int32) if want_aov else None
    at = np.zeros(n, dtype=np.float32) if want_aov else None
    sub = np.ascontiguousarray(subset, dtype=np.int32) if subset is not None else None
    cam = np.ascontiguousarray(cam, dtype=np.float32)
    rc = lib.orc_render(_fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes),
                        _fp(scene.lights), len(scene.lights), _fp(scene.ambient), _fp(cam), w, h, max_depth, spp, seed,
                        0 if mode == "faithful" else 1, threads,
                        pixels.ctypes.data_as(C.POINTER(C.c_int32)), counters.ctypes.data_as(C.POINTER(C.c_uint64)),
                        hsh.ctypes.data_as(C.POINTER(C.c_uint32)) if want_hash else None,
                        aid.ctypes.data_as(C.POINTER(C.c_int32)) if want_aov else None, _fp(at),
                        sub.ctypes.data_as(C.POINTER(C.c_int32)) if sub is not None else None, n if sub is not None else 0)
    if rc != 0:
        raise RuntimeError("orc_render failed rc=%d" % rc)
    shape = (n,) if subset is not None else (h, w)
    return dict(pixels=pixels.reshape(shape), counters=dict(zip(COUNTER_NAMES, (int(v) for v in counters))),
                hash=hsh.reshape(shape) if want_hash else None,
                aov_id=aid.reshape(shape) if want_aov else None, aov_t=at.reshape(shape) if want_aov else None)


def query_spheres(spheres, rays6, kind):
    lib = load()
    spheres = np.ascontiguousarray(spheres, dtype=np.float32)
    rays6 = np.ascontiguousarray(rays6, dtype=np.float32).reshape(-1, 6)
    n = len(rays6)
    ids = np.zeros(n, dtype=np.int32)
    ts = np.zeros(n, dtype=np.float32)
    lib.orc_query_spheres(_fp(spheres), len(spheres), _fp(rays6), n, kind, ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(ts))
    return ids, ts


RAY_RECORD = np.dtype([("origin", np.float32, 3), ("direction", np.float32, 3), ("hit_point", np.float32, 3), ("distance", np.float32),
                       ("hit", np.int32), ("kind", np.uint32), ("pixel", np.uint32), ("level", np.uint32), ("light", np.uint32),
                       ("reserved", np.uint32)])


def ray_log(scene, cam, w, h, max_depth, pixels):
    """Oracle ray log of the listed pixels (nearest-first chain, one record per ray): the checker of rt_ray_log."""
    lib = load()
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).reshape(-1)
    cam = np.ascontiguousarray(cam, dtype=np.float32)
    args = (_fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes), _fp(scene.lights), len(scene.lights),
            _fp(scene.ambient), _fp(cam), w, h, max_depth, pixels.ctypes.data_as(C.POINTER(C.c_uint32)) if len(pixels) else None, len(pixels))
    n = lib.orc_ray_log(*args, None, 0)
    if n < 0:
        raise RuntimeError("orc_ray_log failed rc=%d" % n)
    out = np.zeros(n, dtype=RAY_RECORD)
    if n:
        lib.orc_ray_log(*args, C.c_void_p(out.ctypes.data), n)
    return out


def pack_color(r, g, b):
    return load().orc_pack_color(r, g, b)


def max_threads():
    return load().orc_max_threads()
