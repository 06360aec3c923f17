"""CPU suite: the committed B200 evidence under profiles/ is well formed and agrees with itself — the bench line carries
every key of the bench contract, the roofline figures are the quotients they claim to be, the ncu traffic table names the
kernels bench.py looks up, and the launch list's shares add up."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def _load(name):
    with open(os.path.join(P, name)) as fh:
        return json.load(fh)


def test_bench_line_of_our_arm():
    d = _load("r02_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["gpu_launches"] >= d["steps"] > 0
    B = d["config"]["batch_per_gpu"]
    assert d["value"] == pytest.approx(B / (d["ms_per_step"] * 1e-3), rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["achieved"] == pytest.approx(B * r["algorithmic_bytes_per_state"] / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)
    assert r["traffic"] and 0 < r["fp_pipe_frac"] < 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == B * 18 * 8 and e["d2h_bytes_per_step"] == B * 6 * 8 and e["calls"] >= 100
    assert 0 < e["sync_value"] < e["value"] < d["value"]  # host buffers cost something; overlapping the calls recovers part
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] < e["value"] and c["sample"]
    assert c["parity_vs_gpu_p99_rel"] < 1e-9
    k = d["clocks"]
    assert k["sm_mhz"] > 0.9 * k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line_matches_our_config():
    ours, ref = _load("r02_bench_1gpu.json"), _load("r02_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["config"] == ours["config"] and ref["metric"] == ours["metric"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["as_shipped_python"]["evals_per_s_1_core"] < ref["generated_c"]["evals_per_s"] < ours["e2e"]["value"]


def test_multi_gpu_line_has_the_collective_record():
    d = _load("r02_bench_8gpu.json")
    assert d["n_gpus"] == 8 and d["scaling"] == "weak"
    one = _load("r02_bench_1gpu.json")
    assert d["value"] > 7 * 0.9 * one["value"]  # weak scaling without a data-path collective
    c = d["collective"]
    assert c["fused_matches_nccl"] is True
    assert c["kernel_only_us"] < c["kernel_with_fused_peer_store_us"] < c["kernel_then_nccl_allgather_us"]
    cfg = d["configs"]
    assert cfg["config5_jaco2_avoid_f32"]["global_states"] == 1048576 and cfg["config5_jaco2_avoid_f32"]["gather_ok"] is True
    assert cfg["config4_ur5_rollout_f64"]["global_trajectories"] == 4096 and cfg["config4_ur5_rollout_f64"]["horizon"] == 128


def test_ncu_tables():
    t = _load("ncu_traffic.json")
    for prefix in ("osc:osc_kernel<double, 6", "osc_cfg3:", "osc_cfg5:", "osc_ur5_f32:", "rollout:", "rbd_JMg:", "rbd_JMgC:",
                   "rbd_JMg_B262144:", "rbd_JMgC_B262144:"):
        rows = [v for k, v in t.items() if k.startswith(prefix)]
        assert len(rows) == 1 and rows[0]["dram_mb_per_launch"] > 0 and 0 < rows[0]["fp_pipe_frac"] < 1, prefix
    shares = []
    with open(os.path.join(P, "r02_launches.txt")) as fh:
        for ln in fh:
            m = re.match(r"\s*([0-9.]+)%\s+n=", ln)
            if m:
                shares.append(float(m.group(1)))
    assert shares and sum(shares) == pytest.approx(100.0, abs=0.1)
    for name in ("r02_osc.txt", "r02_osc_cfg3.txt", "r02_osc_cfg5.txt", "r02_rollout.txt", "r02_rbd_JMg.txt"):
        txt = open(os.path.join(P, name)).read()
        assert "stalled_no_instruction" in txt and "dram__bytes_read.sum" in txt and "launch__registers_per_thread" in txt
    assert "ERROR SUMMARY: 0 errors" in open(os.path.join(P, "r02_sanitizer.txt")).read()
