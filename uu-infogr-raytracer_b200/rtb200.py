"""ctypes binding of librtb200.so (include/rtb200.h) plus a thin mirror of the reference's host classes.

The library is the product; this module is the Python-side harness the tests and bench.py call it through.
`RayTracer` / `Surface` mirror the C# classes of the same name (Raytracer/RayTracer.cs:437-1062, surface.cs:7-46):
same scene, same camera state and input handlers, `Tick()` = one frame into `screen.pixels` — with the pixel loop
(:898-901) replaced by rt_render().  There is no CPU fallback: if the shared library or a CUDA device is missing,
everything here raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import numpy as np

import scenes as _scenes

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB200_LIB", os.path.join(_HERE, "librtb200.so"))   # RTB200_LIB: tuning variants only

RT_ACCEL_AUTO, RT_ACCEL_BRUTE, RT_ACCEL_LBVH = 0, 1, 2
RT_OPT_COMPACTION = 1
RT_OPT_HOST_VIA_GPU0 = 2
RT_OPT_PRIMARY_GATE = 3
RT_OPT_SHARED_TARGET = 4
RT_OPT_DEBUG_SHIPPED = 5
RT_OPT_SPARSE_D2H = 6
RT_OPT_HOST_PRECLEARED = 7
RT_OPT_HOST_ZERO_COPY = 8
RT_OPT_GATHER_MODE, RT_OPT_SINK_TILES, RT_OPT_PEER_TILES = 9, 10, 11
RT_OPT_HOST_SHADOW_BINS = 12
RT_OPT_PRIMARY_BINS = 13
RT_INFO_PRIMARY_BIN_BUILDS, RT_INFO_PRIMARY_BINS = 13, 14
RT_INFO_SHADOW_BINS_NS, RT_INFO_SHADOW_BIN_PAIRS = 11, 12
RT_INFO_GATHER_TIMEOUTS, RT_INFO_GATHER_ACTIVE = 9, 10
RT_INFO_GATE_HOST_NS, RT_INFO_GATE_COMPUTES, RT_INFO_LAST_D2H_BYTES, RT_INFO_SCENE_PATH = 1, 2, 3, 4
RT_INFO_LAST_FILL_BYTES, RT_INFO_LAST_FILL_WAIT_NS, RT_INFO_LAST_ENQUEUE_NS, RT_INFO_LAST_TOTAL_NS = 5, 6, 7, 8
COUNTER_NAMES = ["primary", "shadow", "secondary", "sphere_tests", "sphere_disc_pos", "plane_tests",
                 "shade_diffuse", "shade_specular", "shade_mirror", "shaded_hits"]

# every symbol include/rtb200.h declares
ABI_SYMBOLS = ["rt_create", "rt_set_scene", "rt_update_spheres", "rt_render", "rt_render_batch", "rt_render_debug", "rt_query_spheres", "rt_ray_log", "rt_selftest",
               "rt_set_option", "rt_set_partition", "rt_render_device", "rt_ipc_export", "rt_ipc_open", "rt_ipc_close", "rt_dev_alloc",
               "rt_dev_free", "rt_dev_to_host", "rt_dev_memset", "rt_sync", "rt_host_register", "rt_host_unregister", "rt_launch_count", "rt_destroy", "rt_last_error", "rt_abi_version",
               "rt_get_info", "rt_gather_bytes", "rt_gather_attach", "rt_measure_l2_read", "rt_render_mapped", "rt_gl_register_buffer", "rt_gl_unregister_buffer", "rt_render_gl"]


class RtCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3), ("forward", C.c_float * 3),
                ("view_params", C.c_float * 3)]


# rt_ray_record (include/rtb200.h) as a numpy record type: one entry per ray, TracedRay RayTracer.cs:424-435
RAY_RECORD = np.dtype([("origin", np.float32, 3), ("direction", np.float32, 3), ("hit_point", np.float32, 3), ("distance", np.float32),
                       ("hit", np.int32), ("kind", np.uint32), ("pixel", np.uint32), ("level", np.uint32), ("light", np.uint32),
                       ("reserved", np.uint32)])
assert RAY_RECORD.itemsize == 64
RT_SELFTEST_INV_LEN, RT_SELFTEST_PIXEL_DIV, RT_SELFTEST_INV_LEN_RSQ_SEED, RT_SELFTEST_INV_LEN_PAIR = 0, 1, 2, 3
RAY_PRIMARY, RAY_SECONDARY, RAY_SHADOW = 0, 1, 2      # RayKind RayTracer.cs:343-362


class RtStats(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("shadow", C.c_uint64), ("secondary", C.c_uint64),
                ("kernel_ms", C.c_float), ("gather_ms", C.c_float), ("d2h_ms", C.c_float)]


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rtb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """Loads librtb200.so and declares prototypes. Raises if the library was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError("%s not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    camp, statp = C.POINTER(RtCamera), C.POINTER(RtStats)
    lib.rt_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    lib.rt_set_scene.argtypes = [vp, fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_int]
    lib.rt_update_spheres.argtypes = [vp, fp, C.c_int, C.c_int]
    lib.rt_render.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, ip, statp]
    lib.rt_render_batch.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, ip, statp]
    lib.rt_render_debug.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, ip, C.POINTER(C.c_uint32), ip, fp,
                                    C.POINTER(C.c_uint64), statp]
    lib.rt_query_spheres.argtypes = [vp, fp, C.c_int, C.c_int, C.c_int, ip, fp]
    lib.rt_ray_log.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    lib.rt_selftest.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.rt_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.rt_set_partition.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.rt_render_device.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, vp, vp]
    lib.rt_ipc_export.argtypes = [vp, vp, vp]
    lib.rt_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.rt_ipc_close.argtypes = [vp, vp]
    lib.rt_dev_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    lib.rt_dev_free.argtypes = [vp, vp]
    lib.rt_dev_to_host.argtypes = [vp, vp, vp, C.c_uint64]
    lib.rt_dev_memset.argtypes = [vp, vp, C.c_int, C.c_uint64]
    lib.rt_sync.argtypes = [vp]
    lib.rt_host_register.argtypes = [vp, vp, C.c_uint64]
    lib.rt_host_unregister.argtypes = [vp, vp]
    lib.rt_launch_count.argtypes = [vp]
    lib.rt_get_info.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
    lib.rt_gather_bytes.argtypes = [C.c_int, C.c_int]
    lib.rt_gather_bytes.restype = C.c_uint64
    lib.rt_gather_attach.argtypes = [vp, vp, C.c_uint64]
    lib.rt_measure_l2_read.argtypes = [vp, C.c_uint64, C.POINTER(C.c_double)]
    lib.rt_render_mapped.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, vp, C.c_uint64, statp]
    lib.rt_gl_register_buffer.argtypes = [vp, C.c_uint, C.POINTER(vp)]
    lib.rt_gl_unregister_buffer.argtypes = [vp, vp]
    lib.rt_render_gl.argtypes = [vp, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, vp, statp]
    lib.rt_launch_count.restype = C.c_uint64
    lib.rt_destroy.argtypes = [vp]
    lib.rt_last_error.argtypes = [vp]
    lib.rt_last_error.restype = C.c_char_p
    lib.rt_abi_version.restype = C.c_int
    for name in ABI_SYMBOLS:
        if name not in ("rt_launch_count", "rt_last_error", "rt_gather_bytes"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None and a.size else None


def to_rt_camera(cam15) -> RtCamera:
    """15 floats -> rt_camera; an RtCamera is passed through (hosts that render many frames convert once)."""
    if isinstance(cam15, RtCamera):
        return cam15
    cam15 = np.asarray(cam15, dtype=np.float32).reshape(15)
    c = RtCamera()
    for k, name in enumerate(["pos", "right", "up", "forward", "view_params"]):
        getattr(c, name)[:] = [float(v) for v in cam15[3 * k:3 * k + 3]]
    return c


class Context:
    """One rt_context. `devices`: list of CUDA device ids (1, 2, 4 or 8 of them)."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self.lib = load_library()
        devices = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.rt_create(C.byref(h), arr, len(devices))
        if rc != 0:
            raise RtError(rc, (self.lib.rt_last_error(None) or b"").decode())
        self.h = h
        self.n_spheres = 0

    def _check(self, rc):
        if rc != 0:
            raise RtError(rc, (self.lib.rt_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def set_scene(self, scene, accel=RT_ACCEL_AUTO):
        self._scene = scene
        self.n_spheres = len(scene.spheres)
        self._check(self.lib.rt_set_scene(self.h, _fp(scene.spheres), len(scene.spheres), _fp(scene.planes), len(scene.planes),
                                          _fp(scene.lights), len(scene.lights), _fp(scene.ambient), accel))

    def update_spheres(self, spheres, first=0):
        spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 18)
        self._check(self.lib.rt_update_spheres(self.h, _fp(spheres), first, len(spheres)))

    def render(self, cam15, w, h, max_depth=32, spp=1, seed=0, headless=False, out: Optional[np.ndarray] = None):
        """One frame through rt_render (host buffer, D2H inside). Returns (pixels int32[h,w] or None, RtStats)."""
        cam = to_rt_camera(cam15)
        st = RtStats()
        if headless:
            self._check(self.lib.rt_render(self.h, C.byref(cam), w, h, max_depth, spp, seed, None, C.byref(st)))
            return None, st
        px = out if out is not None else np.empty((h, w), dtype=np.int32)
        assert px.dtype == np.int32 and px.size == w * h and px.flags.c_contiguous
        self._check(self.lib.rt_render(self.h, C.byref(cam), w, h, max_depth, spp, seed,
                                       px.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return px, st

    def render_batch(self, cams15, w, h, max_depth=32, spp=1, seed=0, headless=True):
        cams15 = np.asarray(cams15, dtype=np.float32).reshape(-1, 15)
        n = len(cams15)
        arr = (RtCamera * n)(*[to_rt_camera(c) for c in cams15])
        st = RtStats()
        px = None if headless else np.empty((n, h, w), dtype=np.int32)
        self._check(self.lib.rt_render_batch(self.h, arr, n, w, h, max_depth, spp, seed,
                                             None if headless else px.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return px, st

    def render_debug(self, cam15, w, h, max_depth=32, spp=1, seed=0, arrays=True):
        """Instrumented render. arrays=False: pixels and counters only (no per-pixel hash / AOV arrays)."""
        cam = to_rt_camera(cam15)
        st = RtStats()
        n = w * h
        px = np.empty(n, np.int32)
        hsh = np.empty(n if arrays else 0, np.uint32); aid = np.empty(n if arrays else 0, np.int32); at = np.empty(n if arrays else 0, np.float32)
        cnt = np.zeros(16, np.uint64)
        self._check(self.lib.rt_render_debug(self.h, C.byref(cam), w, h, max_depth, spp, seed,
                                             px.ctypes.data_as(C.POINTER(C.c_int32)),
                                             hsh.ctypes.data_as(C.POINTER(C.c_uint32)) if arrays else None,
                                             aid.ctypes.data_as(C.POINTER(C.c_int32)) if arrays else None, _fp(at) if arrays else None,
                                             cnt.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(st)))
        if not arrays:
            return dict(pixels=px.reshape(h, w), hash=None, aov_id=None, aov_t=None,
                        counters=dict(zip(COUNTER_NAMES, (int(v) for v in cnt[:10]))), stats=st,
                        lbvh=dict(node_visits_primary=int(cnt[10]), node_visits_secondary=int(cnt[11]), node_visits_shadow=int(cnt[12]),
                                  brute_fallbacks=int(cnt[13])))
        return dict(pixels=px.reshape(h, w), hash=hsh.reshape(h, w), aov_id=aid.reshape(h, w), aov_t=at.reshape(h, w),
                    counters=dict(zip(COUNTER_NAMES, (int(v) for v in cnt[:10]))), stats=st,
                    lbvh=dict(node_visits_primary=int(cnt[10]), node_visits_secondary=int(cnt[11]), node_visits_shadow=int(cnt[12]),
                              brute_fallbacks=int(cnt[13])))

    def query_spheres(self, rays6, kind, accel=RT_ACCEL_BRUTE):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float32).reshape(-1, 6)
        n = len(rays6)
        ids = np.empty(n, np.int32); ts = np.empty(n, np.float32)
        self._check(self.lib.rt_query_spheres(self.h, _fp(rays6), n, kind, accel, ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(ts)))
        return ids, ts

    def ray_log(self, cam15, w, h, max_depth, pixels):
        """Ray records (RAY_RECORD array) of the listed pixels (linear indices y*w+x) — the data of the reference's debug overlay."""
        pixels = np.ascontiguousarray(pixels, dtype=np.uint32).reshape(-1)
        cam = to_rt_camera(cam15)
        n = C.c_int(0)
        pp = pixels.ctypes.data_as(C.POINTER(C.c_uint32)) if len(pixels) else None
        self._check(self.lib.rt_ray_log(self.h, C.byref(cam), w, h, max_depth, pp, len(pixels), None, 0, C.byref(n)))    # size it
        out = np.zeros(n.value, dtype=RAY_RECORD)
        if n.value:
            self._check(self.lib.rt_ray_log(self.h, C.byref(cam), w, h, max_depth, pp, len(pixels), C.c_void_p(out.ctypes.data), n.value, C.byref(n)))
        return out

    def selftest(self, test):
        """(n_checked, n_mismatch) of one exhaustive device self-test (RT_SELFTEST_*)."""
        n, bad = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.rt_selftest(self.h, test, C.byref(n), C.byref(bad)))
        return n.value, bad.value

    def set_option(self, option, value):
        self._check(self.lib.rt_set_option(self.h, option, value))

    # ---- device-pointer / multi-process interface --------------------------------------------------------
    def set_partition(self, rank, world, tile_rows=8):
        self._check(self.lib.rt_set_partition(self.h, rank, world, tile_rows))

    def render_device(self, cams15, w, h, max_depth, spp, seed, dev_ptr: int, stream: int = 0):
        cams15 = np.asarray(cams15, dtype=np.float32).reshape(-1, 15)
        n = len(cams15)
        arr = (RtCamera * n)(*[to_rt_camera(c) for c in cams15])
        self._check(self.lib.rt_render_device(self.h, arr, n, w, h, max_depth, spp, seed, C.c_void_p(dev_ptr), C.c_void_p(stream)))

    def ipc_export(self, dev_ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.rt_ipc_export(self.h, C.c_void_p(dev_ptr), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        out = C.c_void_p()
        buf = C.create_string_buffer(handle, 64)
        self._check(self.lib.rt_ipc_open(self.h, buf, C.byref(out)))
        return out.value

    def ipc_close(self, dev_ptr: int):
        self._check(self.lib.rt_ipc_close(self.h, C.c_void_p(dev_ptr)))

    def dev_alloc(self, nbytes: int) -> int:
        out = C.c_void_p()
        self._check(self.lib.rt_dev_alloc(self.h, nbytes, C.byref(out)))
        return out.value

    def dev_free(self, ptr: int):
        self._check(self.lib.rt_dev_free(self.h, C.c_void_p(ptr)))

    def dev_to_host(self, ptr: int, nbytes: int, dtype=np.int32) -> np.ndarray:
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self.lib.rt_dev_to_host(self.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), nbytes))
        return out

    def dev_to_host_into(self, arr: np.ndarray, ptr: int, nbytes: int):
        assert arr.nbytes >= nbytes and arr.flags.c_contiguous
        self._check(self.lib.rt_dev_to_host(self.h, arr.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), nbytes))

    def dev_memset(self, ptr: int, byte_value: int, nbytes: int):
        self._check(self.lib.rt_dev_memset(self.h, C.c_void_p(ptr), byte_value, nbytes))

    def host_register(self, arr: np.ndarray):
        self._check(self.lib.rt_host_register(self.h, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._check(self.lib.rt_host_unregister(self.h, arr.ctypes.data_as(C.c_void_p)))

    def sync(self):
        self._check(self.lib.rt_sync(self.h))

    def gather_bytes(self, w: int, h: int) -> int:
        return int(self.lib.rt_gather_bytes(w, h))

    def gather_attach(self, dev_ptr, nbytes: int = 0):
        """Attaches (or with None detaches) the gather area of the packed multi-GPU gather; see rt_gather_attach."""
        self._check(self.lib.rt_gather_attach(self.h, C.c_void_p(dev_ptr) if dev_ptr else None, nbytes))

    def launch_count(self) -> int:
        return int(self.lib.rt_launch_count(self.h))

    def get_info(self, what: int) -> int:
        v = C.c_uint64(0)
        self._check(self.lib.rt_get_info(self.h, what, C.byref(v)))
        return int(v.value)

    def measure_l2_read(self, nbytes: int = 32 << 20) -> float:
        """Measured L2 -> SM read bandwidth in GB/s for an L2-resident working set (roofline peak of the LBVH kernels)."""
        v = C.c_double(0.0)
        self._check(self.lib.rt_measure_l2_read(self.h, nbytes, C.byref(v)))
        return float(v.value)

    # ---- zero-copy display ----------------------------------------------------------------
    def render_mapped(self, cam15, w, h, max_depth, dev_ptr: int, nbytes: int, spp=1, seed=0):
        """One frame straight into a device buffer the display owns (a CUDA-mapped GL pixel-unpack buffer); synchronous."""
        cam = to_rt_camera(cam15)
        st = RtStats()
        self._check(self.lib.rt_render_mapped(self.h, C.byref(cam), w, h, max_depth, spp, seed, C.c_void_p(dev_ptr), nbytes, C.byref(st)))
        return st

    def gl_register_buffer(self, gl_buffer: int) -> int:
        out = C.c_void_p()
        self._check(self.lib.rt_gl_register_buffer(self.h, gl_buffer, C.byref(out)))
        return out.value

    def gl_unregister_buffer(self, resource: int):
        self._check(self.lib.rt_gl_unregister_buffer(self.h, C.c_void_p(resource)))

    def render_gl(self, cam15, w, h, max_depth, resource: int, spp=1, seed=0):
        cam = to_rt_camera(cam15)
        st = RtStats()
        self._check(self.lib.rt_render_gl(self.h, C.byref(cam), w, h, max_depth, spp, seed, C.c_void_p(resource), C.byref(st)))
        return st


# --------------------------------------------------------------------------------------------------------
# Mirror of the reference host classes
# --------------------------------------------------------------------------------------------------------
class Surface:
    """surface.cs:7-20,43-46 — linear framebuffer `pixels[y*width + x] = 0x00RRGGBB`."""

    def __init__(self, width: int, height: int):
        self.width, self.height = width, height
        self.pixels = np.zeros(width * height, dtype=np.int32)

    def Clear(self, c: int):
        self.pixels[:] = c


class RayTracer:
    """RayTracer.cs:437-1062 with the trace methods replaced by the native backend."""
    ReflectionRecursionLimit = _scenes.REFLECTION_RECURSION_LIMIT          # :490

    def __init__(self, screen: Surface, devices: Optional[Sequence[int]] = None, scene=None):
        self.screen = screen                                                # :535-537
        self._scene = scene if scene is not None else _scenes.default_scene()   # :441-469
        self._cameraPosition = np.zeros(3, dtype=np.float32)                # :494
        self._yaw = np.float32(0.0)                                         # :498
        self._pitch = np.float32(0.0)                                       # :502
        self._ctx = Context(devices)
        self._ctx.set_scene(self._scene)
        self.last_stats = None

    def OnKeyPress(self, key: str):                                         # :543-554
        right, up, fwd = _scenes.camera_basis(float(self._yaw), float(self._pitch))
        s = np.float32(0.05)
        p = self._cameraPosition
        if key == "W": p = p + fwd * s
        elif key == "A": p = p - right * s
        elif key == "S": p = p - fwd * s
        elif key == "D": p = p + right * s
        elif key == "Space": p = p - up * s
        elif key in ("LeftShift", "RightShift"): p = p + up * s
        self._cameraPosition = p.astype(np.float32)

    def OnMouseMove(self, delta_x: float, delta_y: float):                  # :1058-1061
        self._yaw = np.float32(self._yaw + np.float32(delta_x) / np.float32(360))
        self._pitch = np.float32(self._pitch + np.float32(delta_y) / np.float32(360))

    def camera(self) -> np.ndarray:
        return _scenes.make_camera(self._cameraPosition, float(self._yaw), float(self._pitch), self.screen.width, self.screen.height)

    def Tick(self):                                                         # :886-901
        # screen.Clear(0) (:890) is subsumed: the backend writes every pixel of the frame.
        px, st = self._ctx.render(self.camera(), self.screen.width, self.screen.height, self.ReflectionRecursionLimit, 1, 0,
                                  out=self.screen.pixels.reshape(self.screen.height, self.screen.width))
        self.last_stats = st

    def close(self):
        self._ctx.close()
