"""bench.py output contract: one JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"]


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line(built):
    """`bench.py --impl reference`: the reference's CPU path (oracle port, faithful mode) on the host cores — no GPU needed."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "default_scene_3840x2160_depth8" and d["config"]["rays_per_frame"] == 23447045
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_b200_arm_line(built):
    d = _run(["--steps", "3", "--warmup", "3"])
    for k in BASE_KEYS + ["roofline", "cpu_baseline", "gpu_launches", "clocks"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "strong"
    assert d["value"] > 1000.0                          # north_star: >= 1 Grays/s on one B200
    assert d["gpu_launches"] == 3                       # one launch per step
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    e = d["e2e"]
    frame_bytes = 3840 * 2160 * 4 * d["config"]["frames_per_step"]
    # sparse device->host return: what the frame gates prove black (about a third of this frame) does not cross PCIe
    assert 0.5 * frame_bytes < e["d2h_bytes_per_step"] < 0.75 * frame_bytes and e["frame_bytes_per_step"] == frame_bytes
    assert e["dense_return"]["d2h_bytes_per_step"] == frame_bytes and e["h2d_bytes_per_step"] > 0
    assert e["value"] > e["dense_return"]["value"] * 0.9
    for k in ("value_single_frame", "value_moving_camera", "value_gates_off", "per_config"):
        assert k in d, k
    assert d["value_gates_off"]["value"] < d["value"]
    for name, rec in d["per_config"].items():
        assert rec["frame_equals_instrumented_render"], name
        assert 0 < rec["roofline"]["frac"] < 1, name
    assert 0 < e["value"] < d["value"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "Mrays/s"


def test_per_config_extras_against_a_stub_library():
    """bench.per_config_extras walks every BASELINE.json config through the binding's API; run here against a stub of that API (no GPU)
    so that a slip in the harness cannot cost the bench line: keys, the primary-bins A/B of the LBVH configs, option restored."""
    import types
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench

    class Stats:
        kernel_ms = 1.0

    class Ctx:
        options = []

        def __init__(self, devs):
            self.bins = 1
            self.path = 0

        def set_scene(self, sc, accel):
            self.path = 3 if len(sc.spheres) >= 48 and accel != 1 else (1 if len(sc.spheres) >= 48 else 0)

        def render_debug(self, cam, w, h, depth, spp, seed, arrays=True):
            cnt = dict(primary=w * h * spp, shadow=w * h, secondary=w * h // 2, sphere_tests=10 * w * h, sphere_disc_pos=w * h, plane_tests=w * h,
                       shade_diffuse=w * h, shade_specular=w * h // 2, shade_mirror=w * h // 4, shaded_hits=w * h)
            return dict(pixels=np.zeros((h, w), np.int32), counters=cnt, stats=Stats(),
                        lbvh=dict(node_visits_primary=w * h * (1 if self.bins else 11), node_visits_secondary=w * h, node_visits_shadow=0, brute_fallbacks=0))

        def render(self, cam, w, h, depth, spp=1, seed=0, headless=False):
            return (None if headless else np.zeros((h, w), np.int32)), Stats()

        def get_info(self, what):
            return {stub.RT_INFO_SCENE_PATH: self.path, stub.RT_INFO_PRIMARY_BINS: self.bins}[what]

        def set_option(self, opt, value):
            assert opt == stub.RT_OPT_PRIMARY_BINS
            self.bins = int(value); Ctx.options.append(int(value))

        def measure_l2_read(self, nbytes):
            return 18400.0

        def close(self):
            pass

    stub = types.SimpleNamespace(Context=Ctx, RT_ACCEL_AUTO=0, RT_ACCEL_BRUTE=1, RT_ACCEL_LBVH=2, RT_INFO_SCENE_PATH=4, RT_INFO_PRIMARY_BINS=14,
                                 RT_OPT_PRIMARY_BINS=13)
    out = bench.per_config_extras(stub, 0, dict(hbm_gbs=6548.0, sm_max_mhz=1965.0))
    assert len(out) == 5
    n_lbvh = 0
    for name, rec in out.items():
        assert rec["frame_equals_instrumented_render"] and rec["kernel_ms"] == 1.0 and "roofline" in rec, name
        if rec["path"] == "lbvh":
            n_lbvh += 1
            pb = rec["primary_bins"]
            assert "error" not in pb, pb
            assert pb["default"] is True and pb["off"]["node_visits_primary"] > pb["on"]["node_visits_primary"]
            assert rec["roofline"]["bound"] == "L2"
    assert n_lbvh == 2 and Ctx.options == [0, 1, 1, 0, 1, 1]          # off, on, restored — per LBVH config
