"""CPU suite: the bench lines committed under profiles/ carry every key of the bench contract, with consistent
numbers (value = states / time, roofline.frac = achieved / peak, e2e has its byte counts ...).  Guards bench.py's output
format without needing a GPU; the lines themselves were produced on a B200 by tools/run_round_capture.sh."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _load(name):
    with open(os.path.join(PROF, name)) as fh:
        return json.loads(fh.read())


@pytest.mark.parametrize("name", ["r01_bench_1gpu.json", "r01_bench_2gpu.json"])
def test_our_arm_line(name):
    d = _load(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    B = d["config"]["batch_per_gpu"]
    assert d["value"] == pytest.approx(d["n_gpus"] * B / (d["ms_per_step"] * 1e-3), rel=1e-6)
    assert d["gpu_launches"] >= d["steps"] > 0 and d["warmup"] >= 3
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] == B * 18 * 8 and e["d2h_bytes_per_step"] == B * 6 * 8
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
    if d["n_gpus"] == 1:
        assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
        assert r["traffic"] is None or r["traffic"] > 0
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        assert c["parity_vs_gpu_p99_rel"] < 1e-9  # the CPU sample and the GPU path agree in the same run
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _load("r01_bench_reference_arm.json")
    ours = _load("r01_bench_1gpu.json")
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better", "dtype"):
        assert d[k] == ours[k]
    assert d["config"]["workload"] == ours["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
