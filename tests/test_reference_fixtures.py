"""The pin of the oracle (and of the CUDA path) to frames rendered by the UNMODIFIED reference.

`tests/golden/ref/GenGolden.cs` is dropped into the reference's `Raytracer/` project and run under `dotnet` (README.md there); it
renders the cases below with the reference's own `RayTracer.Tick()` and writes `<name>.bin` + `ref_math.bin` + `manifest.json` into
`tests/golden/ref/`. No .NET exists in the build image, so until someone commits those files this module cannot compare anything:

  * default runs (`-m "not gpu"`, `-m gpu`): the comparisons are SKIPPED with the reason "UNPINNED";
  * `pytest -m refpin`: they FAIL with "UNPINNED" — the state of the parity claim is then visible in a test result.

What always runs: the generator is committed, lists the same cases as this file, and touches only members the reference has.
"""
import json
import os
import re

import numpy as np
import pytest

import oracle_lib as O
import scenes
from common import assert_image_parity

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "golden", "ref")
GEN = os.path.join(REF_DIR, "GenGolden.cs")
DEPTH = 32                    # the reference's private const ReflectionRecursionLimit (RayTracer.cs:490)

# name, w, h, pos, yaw, pitch — the table `Cases` of GenGolden.cs
CASES = [
    ("ref_default_160x90", 160, 90, (0.0, 0.0, 0.0), 0.0, 0.0),
    ("ref_default_192x108", 192, 108, (0.0, 0.0, 0.0), 0.0, 0.0),
    ("ref_moved_192x108", 192, 108, (0.3, 0.5, -1.0), 0.2, 0.15),
    ("ref_above_160x90", 160, 90, (-2.0, 2.5, 3.0), -0.4, 0.5),
    ("ref_down_160x90", 160, 90, (0.0, 3.0, 2.0), 0.0, 1.3),
    ("ref_behind_160x90", 160, 90, (0.0, 0.0, 6.0), 3.1, 0.0),
    ("ref_sky_160x90", 160, 90, (0.0, 0.5, 0.0), 0.0, -0.6),
    ("ref_default_1280x720", 1280, 720, (0.0, 0.0, 0.0), 0.0, 0.0),
]


def _have(name):
    return os.path.exists(os.path.join(REF_DIR, name))


def _unpinned(request, what):
    msg = ("UNPINNED: %s not found — run tests/golden/ref/GenGolden.cs inside the unmodified reference (tests/golden/ref/README.md) "
           "and commit its output; until then parity is against our restatement of RayTracer.cs only" % what)
    if "refpin" in (request.config.getoption("-m") or ""):
        pytest.fail(msg)
    pytest.skip(msg)


def _load_frame(name, w, h):
    px = np.fromfile(os.path.join(REF_DIR, name + ".bin"), dtype="<i4")
    assert px.size == w * h, "%s.bin holds %d pixels, expected %dx%d" % (name, px.size, w, h)
    return px.reshape(h, w)


def test_generator_is_committed_and_lists_the_same_cases():
    src = open(GEN).read()
    rows = re.findall(r'\("(ref_[a-z0-9_]+)",\s*(\d+),\s*(\d+),\s*([-0-9.]+)f,\s*([-0-9.]+)f,\s*([-0-9.]+)f,\s*([-0-9.]+)f,\s*([-0-9.]+)f\)', src)
    got = [(n, int(w), int(h), (float(x), float(y), float(z)), float(yaw), float(pitch)) for n, w, h, x, y, z, yaw, pitch in rows]
    assert got == CASES
    # it drives the reference through members that exist there, unmodified (RayTracer.cs:494-502, :535, :886; surface.cs:9-20)
    for needle in ('GetField("_cameraPosition"', 'GetField("_yaw"', 'GetField("_pitch"', "new RayTracer(screen)", "new Surface(c.w, c.h)",
                   "rt.Tick()", "screen.pixels", "namespace Template"):
        assert needle in src, needle


def test_generator_matches_the_reference_source_when_it_is_present():
    """In the build container (/root/reference present): every member GenGolden.cs reaches into exists in the unmodified
    RayTracer.cs / surface.cs with the type it assumes, and the project is the one it documents. Skipped where the checkout is
    absent (the GPU box)."""
    ref = "/root/reference/Raytracer"
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout on this machine")
    rt = open(os.path.join(ref, "RayTracer.cs")).read()
    sf = open(os.path.join(ref, "surface.cs"), encoding="utf-8-sig").read()
    proj = open(os.path.join(ref, "InfogrRaytracer.csproj")).read()
    assert re.search(r"private Vector3 _cameraPosition\b", rt) and re.search(r"private float _yaw;", rt) and re.search(r"private float _pitch;", rt)
    assert "public RayTracer(Surface screen)" in rt and "public void Tick()" in rt and "namespace Template;" in rt
    assert "private const int ReflectionRecursionLimit = 32;" in rt and DEPTH == 32
    assert "public int[] pixels;" in sf and "public Surface(int w, int h)" in sf
    assert 'Include="OpenTK" Version="4.7.1"' in proj and "<RootNamespace>Template</RootNamespace>" in proj and "net6.0" in proj
    assert "public static void Main()" in open(os.path.join(ref, "template.cs")).read()       # hence -p:StartupObject=Template.GenGolden


@pytest.mark.refpin
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_reference_frame(request, built, case):
    name, w, h, pos, yaw, pitch = case
    if not _have(name + ".bin"):
        _unpinned(request, "tests/golden/ref/%s.bin" % name)
    ref = _load_frame(name, w, h)
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=pos, yaw=yaw, pitch=pitch, width=w, height=h)
    diffs = {}
    for variant in ("", "truediv"):
        got = O.render(sc, cam, w, h, DEPTH, variant=variant)["pixels"]
        diffs[variant or "rcp_mul"] = int((got != ref).sum())
    print("%s: differing pixels vs the reference — Normalize as v*(1/len): %d, as v/len: %d" % (name, diffs["rcp_mul"], diffs["truediv"]))
    got = O.render(sc, cam, w, h, DEPTH)["pixels"]
    assert_image_parity(got, ref, name)          # BASELINE.json tolerance: <= 1/255 on >= 99.9 % of pixels, none > 4/255
    assert diffs["rcp_mul"] <= diffs["truediv"], "the reference's OpenTK normalises by true division: rebuild the oracle with ORC_NORMALIZE_TRUE_DIV"


@pytest.mark.refpin
def test_reference_math_probe(request):
    """OpenTK 4.7.1 / System.Math as the reference really evaluates them (row a19), operation by operation, against the fp32
    restatement used by oracle/rt_oracle.cpp and csrc/rt_math.cuh."""
    if not _have("ref_math.bin"):
        _unpinned(request, "tests/golden/ref/ref_math.bin")
    raw = np.fromfile(os.path.join(REF_DIR, "ref_math.bin"), dtype=np.uint8)
    n = int(raw[:4].view("<i4")[0])
    inp = raw[4:4 + n * 24].view("<f4").reshape(n, 6)
    out = raw[4 + n * 24:].view("<f4").reshape(n, 16)
    a, b = inp[:, :3].astype(np.float32), inp[:, 3:].astype(np.float32)
    f32 = np.float32
    s = (a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]) + a[:, 2] * a[:, 2]
    length = np.sqrt(s).astype(f32)
    n_rcp = a * (f32(1.0) / length)[:, None]
    n_div = a / length[:, None]
    m_rcp = int((n_rcp.view(np.uint32) != out[:, 0:3].view(np.uint32)).any(1).sum())
    m_div = int((n_div.view(np.uint32) != out[:, 0:3].view(np.uint32)).any(1).sum())
    print("Vector3.Normalize: v*(1/len) mismatches %d, v/len mismatches %d of %d" % (m_rcp, m_div, n))
    assert m_rcp == 0, "OpenTK's Normalize is not v * (1f / Length): %d of %d probes differ (true division: %d)" % (m_rcp, n, m_div)
    dot = (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]
    assert np.array_equal(dot.view(np.uint32), out[:, 3].view(np.uint32)), "Vector3.Dot association differs"
    cr = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)
    assert np.array_equal(cr.view(np.uint32), out[:, 4:7].view(np.uint32)), "Vector3.Cross differs"
    rng = np.where(np.arange(n) % 3 == 0, 1.0, np.where(np.arange(n) % 3 == 1, 40.0, 1000.0)).astype(f32)
    assert np.array_equal(np.sqrt(np.abs(a[:, 2]).astype(np.float64)).astype(f32).view(np.uint32), out[:, 10].view(np.uint32))
    p05 = (np.abs(a[:, 1] / rng).astype(np.float64) ** 0.5).astype(f32)
    assert np.array_equal(p05.view(np.uint32), out[:, 9].view(np.uint32)), "(float)Math.Pow(x, 0.5) != sqrt"
    assert np.array_equal(np.where(a[:, 0] <= 0, f32(0), a[:, 0]).view(np.uint32) & 0x7FFFFFFF, out[:, 11].view(np.uint32) & 0x7FFFFFFF)
    assert np.array_equal(np.trunc(a[:, 1] * f32(1e3)).astype(f32), out[:, 12]), "(int)float does not truncate toward zero"
    inv_sq = (1.0 / (a[:, 2].astype(np.float64) ** 2)).astype(f32)
    assert np.array_equal(inv_sq.view(np.uint32), out[:, 13].view(np.uint32)), "(float)(1 / Math.Pow(d, 2)) differs from 1 / (d*d) in f64"
    assert np.array_equal((a[:, 0] * b[:, 0]).view(np.uint32), out[:, 14].view(np.uint32))
    assert np.array_equal(((f32(1.0) / a[:, 0]) * a[:, 0]).view(np.uint32), out[:, 15].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.refpin
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_equals_reference_frame(request, built, case):
    import rtb200
    name, w, h, pos, yaw, pitch = case
    if not _have(name + ".bin"):
        _unpinned(request, "tests/golden/ref/%s.bin" % name)
    ref = _load_frame(name, w, h)
    ctx = rtb200.Context([0]); ctx.set_scene(scenes.default_scene())
    got, _ = ctx.render(scenes.make_camera(pos=pos, yaw=yaw, pitch=pitch, width=w, height=h), w, h, DEPTH)
    ctx.close()
    assert_image_parity(got, ref, name)
