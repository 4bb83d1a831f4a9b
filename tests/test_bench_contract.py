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
