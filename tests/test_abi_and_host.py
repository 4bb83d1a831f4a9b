"""CPU tests of the boundary and the host-side logic: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), error behaviour without a device, camera / scene packing."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import scenes
from conftest import HAS_GPU, ROOT


def test_library_exports_every_declared_symbol(built):
    import rtb200
    lib = rtb200.load_library()
    header = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    declared = sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    assert sorted(declared) == sorted(rtb200.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rt_abi_version() == 2


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device error path")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product fails loudly (RT_ERR_CUDA); it never renders on the CPU."""
    import rtb200
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Context([0])
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "uu-infogr-raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".cs", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "librt_oracle" not in txt and "libhostemu" not in txt and "hostemu_lib" not in txt, os.path.join(dirpath, f)


def test_camera_basis_default():
    """:511-523 at yaw = pitch = 0: R = (1,0,-0), U = (0,-1,-0), F = (0,-0,1) — signed zeros matter (A.1)."""
    r, u, f = scenes.camera_basis(0.0, 0.0)
    assert list(r) == [1.0, 0.0, 0.0] and math.copysign(1, r[2]) == -1
    assert list(f) == [0.0, 0.0, 1.0] and math.copysign(1, f[1]) == -1
    assert list(u) == [0.0, -1.0, 0.0]


def test_view_params():
    """:892-896 — planeHeight = 0.3 * tan(30 deg) * 2 = 0.34641016, planeWidth = planeHeight * w/h."""
    v = scenes.view_params(1280, 720)
    assert abs(float(v[1]) - 0.34641016) < 1e-7 and v[2] == np.float32(0.3)
    assert v[0] == np.float32(v[1] * np.float32(np.float32(1280) / np.float32(720)))


def test_camera_basis_orthonormal_when_moved():
    r, u, f = scenes.camera_basis(0.7, -0.3)
    for a, b in [(r, u), (r, f), (u, f)]:
        assert abs(float(np.dot(a, b))) < 1e-6
    for a in (r, u, f):
        assert abs(float(np.linalg.norm(a)) - 1) < 1e-6


def test_pcg32_known_vector():
    """PCG32 reference demo: seed 42, stream 54 -> 0xa15c02b7, 0x7b47f409, 0xba1d3330, ..."""
    rng = scenes.PCG32(42, 54)
    assert [rng.next_u32() for _ in range(3)] == [0xA15C02B7, 0x7B47F409, 0xBA1D3330]


def test_default_scene_records():
    sc = scenes.default_scene()
    assert sc.spheres.shape == (3, 18) and sc.planes.shape == (1, 20) and sc.lights.shape == (2, 4)
    assert list(sc.spheres[1, :4]) == [3.0, 0.0, 5.0, 1.0] and sc.spheres[1, 17] == 1.0
    assert list(sc.spheres[1, 10:13]) == [np.float32(0.4)] * 3 and sc.spheres[1, 13] == 1.0      # Plastic :127-129
    assert list(sc.spheres[2, 14:17]) == [1.0, 1.0, 1.0] and not sc.spheres[2, 4:14].any()        # Mirror :146-148
    assert list(sc.planes[0, :6]) == [0, -1, 0, 0, 1, 0] and sc.planes[0, 15] == np.float32(0.5)
    assert sc.ambient[0] == np.float32(43.0) / np.float32(255.0)


def test_synthetic_scene_is_deterministic():
    a, _ = scenes.random_spheres_scene(64, 1024, 24.0, 4.0, 52.0, "x")
    b, _ = scenes.random_spheres_scene(64, 1024, 24.0, 4.0, 52.0, "x")
    assert np.array_equal(a, b)
    assert np.all(a[:, 3] >= 0.15) and np.all(a[:, 3] <= 0.6) and np.all(a[:, 1] >= -1 + a[:, 3] - 1e-6)
    assert np.all(a[:, 17] == a[:, 3] * a[:, 3])


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device error path")
def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu(built, tmp_path):
    """host/rt_demo (C++ mirror of RayTracer/Surface over the C ABI) links against librtb200.so; without a device it must
    exit non-zero with the library's error message instead of rendering on the CPU."""
    import subprocess
    exe = os.path.join(ROOT, "uu-infogr-raytracer_b200", "host", "rt_demo")
    assert os.path.exists(exe)
    r = subprocess.run([exe, str(tmp_path / "x.ppm"), "32", "18"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert not (tmp_path / "x.ppm").exists()


def test_sass_has_packed_fp32_but_no_contracted_packed_fma(built, tmp_path):
    """The exact-count kernels use Blackwell's packed fp32 instructions (FADD2 / FMUL2) for two spheres / two lights at a time,
    but a packed FMA must never appear in REFERENCE arithmetic: ptxas contracts packed mul+add into FFMA2 even under --fmad=false,
    which would round dot products and discriminants once instead of twice (csrc/rt_trace.cuh). The only FFMA2 allowed are the
    ones written on purpose with rt_fma2 (csrc/rt_math.cuh) — the correctly rounded sqrt / reciprocal cores of rt_inv_len2 and the
    LBVH slab tests (box_entry2, not reference arithmetic): every FFMA2 of the library must carry the source line of that asm
    statement; a contracted one would carry the line of rt_mul2 / rt_add2. Scalar FFMA only comes from IEEE div / sqrt sequences."""
    import re
    import shutil
    import subprocess
    if not (shutil.which("cuobjdump") and shutil.which("nvdisasm")):
        pytest.skip("cuobjdump / nvdisasm not available")
    lib = os.path.join(ROOT, "uu-infogr-raytracer_b200", "librtb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert sass.count("FADD2") > 0 and sass.count("FMUL2") > 0
    assert "STG.E.128" in sass                      # 128-bit framebuffer stores
    for mnem in ("HMMA", "UTCHMMA", "UTCQMMA"):      # no tensor cores on this path (north_star)
        assert mnem not in sass
    # where does each FFMA2 come from?
    math_src = open(os.path.join(ROOT, "uu-infogr-raytracer_b200", "csrc", "rt_math.cuh")).read().splitlines()
    fma_lines = {i + 1 + k for i, l in enumerate(math_src) if "fma.rn.f32x2" in l for k in (0, 1)}    # the asm statement spans two lines
    assert len(fma_lines) == 2
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp_path, capture_output=True, check=True)
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert cubins
    n_ffma2 = 0
    for cb in cubins:
        dis = subprocess.run(["nvdisasm", "-g", "-c", str(tmp_path / cb)], capture_output=True, text=True).stdout
        cur = None
        for l in dis.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            elif "FFMA2" in l:
                n_ffma2 += 1
                assert cur is not None and cur[0] == "rt_math.cuh" and cur[1] in fma_lines, (cur, l)
    assert n_ffma2 == sass.count("FFMA2") and n_ffma2 > 0
