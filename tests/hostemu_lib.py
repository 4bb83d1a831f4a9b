"""ctypes binding of tests/hostemu/libhostemu.so: the DEVICE trace code compiled as plain C++ (test-only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
_libs = {}
VARIANT = ""          # "" = the shipped configuration; "_v1" = built with -DRT_GATES_V1, the first-generation gate shapes (use_variant)


def use_variant(name):
    """Selects the library the module-level helpers call: "" (default build) or "_v1" (RT_GATES_V1)."""
    global VARIANT
    VARIANT = name


def load():
    _lib = _libs.get(VARIANT)
    if _lib is None:
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(os.path.join(HERE, "libhostemu%s.so" % VARIANT))
        _libs[VARIANT] = _lib
        fp = C.POINTER(C.c_float)
        _lib.emu_render.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                    C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), fp, C.POINTER(C.c_uint64)]
        _lib.emu_render.restype = C.c_int
        _lib.emu_query.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), fp]
        _lib.emu_query.restype = C.c_int
        _lib.emu_ray_log.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_uint32), C.c_int, C.c_void_p, C.c_int]
        _lib.emu_ray_log.restype = C.c_int
        _lib.emu_gates.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.POINTER(C.c_int), fp, C.POINTER(C.c_ubyte)]
        _lib.emu_gates.restype = C.c_int
        _lib.emu_row_plan.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_int)]
        _lib.emu_row_plan.restype = C.c_int
        _lib.emu_shadow_bins_check.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, C.POINTER(C.c_uint64)]
        _lib.emu_shadow_bins_check.restype = C.c_int
        _lib.emu_tile_cover.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
        _lib.emu_tile_cover.restype = C.c_int
        _lib.emu_primary_bins_check.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        _lib.emu_primary_bins_check.restype = C.c_int
        _lib.emu_set_primary_bins_capacity.argtypes = [C.c_int]
        _lib.emu_set_primary_bins_capacity.restype = None
        _lib.emu_set_primary_bins_shuffle.argtypes = [C.c_uint64]
        _lib.emu_set_primary_bins_shuffle.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None and a.size else None


def render(scene, cam, w, h, max_depth=32, spp=1, seed=0, tiny=0, debug=False):
    """tiny: 0 global-memory policy, 1 TinyScene<-1>, 2 TinyScene<exact NS>, 3 LBVH, 4 staged, 6 LBVH + primary bins (see hostemu.cpp)"""
    lib = load()
    n = w * h
    px = np.zeros(n, np.int32)
    hsh = np.zeros(n, np.uint32); aid = np.zeros(n, np.int32); at = np.zeros(n, np.float32); cnt = np.zeros(14, np.uint64)
    cam = np.ascontiguousarray(cam, np.float32)
    rc = lib.emu_render(_fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes), _fp(scene.lights),
                        len(scene.lights), _fp(scene.ambient), _fp(cam), w, h, max_depth, spp, seed, int(tiny),
                        px.ctypes.data_as(C.POINTER(C.c_int32)),
                        hsh.ctypes.data_as(C.POINTER(C.c_uint32)) if debug else None,
                        aid.ctypes.data_as(C.POINTER(C.c_int32)) if debug else None, _fp(at) if debug else None,
                        cnt.ctypes.data_as(C.POINTER(C.c_uint64)) if debug else None)
    assert rc == 0
    return dict(pixels=px.reshape(h, w), hash=hsh.reshape(h, w), aov_id=aid.reshape(h, w), aov_t=at.reshape(h, w),
                counters=[int(v) for v in cnt[:10]], lbvh=[int(v) for v in cnt[10:]])


def query(spheres, rays6, kind, accel):
    """accel: 1 brute, 2 LBVH"""
    lib = load()
    spheres = np.ascontiguousarray(spheres, np.float32); rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
    n = len(rays6)
    ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
    rc = lib.emu_query(_fp(spheres), len(spheres), _fp(rays6), n, kind, accel, ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(ts))
    assert rc == 0, rc
    return ids, ts


def ray_log(scene, cam, w, h, max_depth, pixels):
    """The device ray-log code (LogDbg + GlobalScene, the body of k_ray_log) on the host."""
    from oracle_lib import RAY_RECORD
    lib = load()
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).reshape(-1)
    cam = np.ascontiguousarray(cam, np.float32)
    args = (_fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes), _fp(scene.lights), len(scene.lights),
            _fp(scene.ambient), _fp(cam), w, h, max_depth, pixels.ctypes.data_as(C.POINTER(C.c_uint32)) if len(pixels) else None, len(pixels))
    n = lib.emu_ray_log(*args, None, 0)
    if n < 0:
        raise RuntimeError("emu_ray_log failed rc=%d" % n)
    out = np.zeros(n, dtype=RAY_RECORD)
    if n:
        lib.emu_ray_log(*args, C.c_void_p(out.ctypes.data), n)
    return out


GATE_SPHERES, GATE_MIRROR, GATE_SHADOW0, GATE_BLACK = 1, 2, 4, 0x80


def gates(scene, cam, w, h, want_bits=True):
    """The host's frame gates (csrc/rt_gate.cuh): dict(spheres, mirror, shadow[4] rects (x0, y0, x1, y1) inclusive, empty =
    (w, h, w, h); sky, deep affine (a, bx, by); bits uint8[h, w] = per-pixel skip bits, 0x80 = black untraced)."""
    lib = load()
    cam = np.ascontiguousarray(cam, np.float32)
    r = (C.c_int * 24)(); aff = np.zeros(6, np.float32)
    bits = np.zeros(w * h, np.uint8) if want_bits else None
    lib.emu_gates(_fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes), _fp(scene.lights), len(scene.lights),
                  _fp(cam), w, h, r, _fp(aff), bits.ctypes.data_as(C.POINTER(C.c_ubyte)) if want_bits else None)
    rr = [tuple(int(v) for v in r[4 * i:4 * i + 4]) for i in range(6)]
    return dict(spheres=rr[0], mirror=rr[1], shadow=rr[2:6], sky=tuple(float(v) for v in aff[0:3]), deep=tuple(float(v) for v in aff[3:6]),
                bits=bits.reshape(h, w) if want_bits else None)


def gate_rect(scene, cam, w, h):
    return gates(scene, cam, w, h, want_bits=False)["spheres"]


def sky_mask(scene, cam, w, h):
    """(mask bool[h,w], (a, bx, by)): where the sky gate says the plane cannot be hit by the pixel's primary ray."""
    g = gates(scene, cam, w, h, want_bits=False)
    a, bx, by = (np.float32(v) for v in g["sky"])
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    # fma(bx, x, fma(by, y, a)) evaluated exactly (float64 holds the float32 products and sums of these magnitudes to within 2^-53)
    inner = np.float32(np.float64(by) * ys + np.float64(a))
    val = np.float32(np.float64(bx) * xs + np.float64(inner))
    return val > 0, g["sky"]


ROW_COPY, ROW_RECT, ROW_BLACK = 0, 1, 2


def row_plan(scene, cam, w, h):
    """plan_rows (csrc/rt_gate.cuh) — what rt_render's sparse device->host return copies: (kind uint8[h], rx0, rx1, sparse)."""
    lib = load()
    cam = np.ascontiguousarray(cam, np.float32)
    kind = np.zeros(h, np.uint8); rx = (C.c_int * 2)()
    sparse = lib.emu_row_plan(_fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes), _fp(scene.lights), len(scene.lights),
                              _fp(cam), w, h, kind.ctypes.data_as(C.POINTER(C.c_ubyte)), rx)
    return kind, int(rx[0]), int(rx[1]), bool(sparse)


def shadow_bins_check(spheres, lights, points):
    """Binned shadow query (rt_shadow_grid.cuh) against the reference's loop over all spheres for every (point, light):
    dict(decided, mismatches, undecided, occluded, sphere_tests)."""
    lib = load()
    spheres = np.ascontiguousarray(spheres, np.float32); lights = np.ascontiguousarray(lights, np.float32)
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros(5, np.uint64)
    rc = lib.emu_shadow_bins_check(_fp(spheres), len(spheres), _fp(lights), len(lights), _fp(points), len(points), out.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0
    return dict(zip(("decided", "mismatches", "undecided", "occluded", "sphere_tests"), (int(v) for v in out)))


def set_primary_bins_shuffle(seed):
    """0: the bins of the emulation are built by primary_bins_build_host; else by the replay of the DEVICE build's algorithm with its
    atomics resolved in a random order drawn from `seed` (hostemu.cpp: primary_bins_build_device_order)."""
    load().emu_set_primary_bins_shuffle(int(seed))


def primary_bins_check(spheres, cam, w, h, capacity=-1, shuffle=0):
    """Per-frame primary bins (rt_primary_bins.cuh) against the reference's loop over all spheres for every pixel's primary ray.
    capacity: entries of the list array (-1 = the device build's rule); shuffle: see set_primary_bins_shuffle."""
    lib = load()
    spheres = np.ascontiguousarray(spheres, np.float32); cam = np.ascontiguousarray(cam, np.float32)
    out = np.zeros(9, np.uint64)
    lib.emu_set_primary_bins_capacity(int(capacity))
    lib.emu_set_primary_bins_shuffle(int(shuffle))
    try:
        rc = lib.emu_primary_bins_check(_fp(spheres), len(spheres), _fp(cam), w, h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    finally:
        lib.emu_set_primary_bins_capacity(-1)
        lib.emu_set_primary_bins_shuffle(0)
    assert rc == 0
    return dict(zip(("by_bins", "by_tree", "missing", "differ", "sphere_tests", "tiles_without_list", "entries", "everywhere", "valid"),
                    (int(v) for v in out)))


def tile_cover(kind, ppt, w, h, tile_rows):
    """How often each pixel is owned under the kernels' 2-D block mapping (csrc/rt_tiles.cuh): (cover uint8[h, w], bad spans)."""
    lib = load()
    cover = np.zeros(w * h, np.uint8)
    bad = lib.emu_tile_cover(kind, ppt, w, h, tile_rows, cover.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return cover.reshape(h, w), bad
