"""CPU suite: bench.py's behaviour where it can run without a GPU.

* the reference arm (`python bench.py --impl reference`) is RUN here on a small sample and its JSON line checked against
  the contract: one line on stdout, `impl`, the metric / unit / config of our arm (the driver compares the two `config`
  dicts), `cpu_baseline` describing the run, zero-byte `e2e`, and the `generated_c` / `as_shipped_python` pair;
* our arm must refuse to run without a CUDA device (there is no CPU fallback to time);
* the hidden worker of the as-shipped arm is only reachable inside the reference's environment.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _run(*args, env=None, timeout=900):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env)


def test_reference_arm_runs_and_keeps_the_contract():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1  # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["steps"] == 2 and d["warmup"] == 1
    assert d["config"] == bench.bench_config(1)  # key-compatible with our arm's config
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["value"] == pytest.approx(d["sample_per_step"] / (d["ms_per_step"] * 1e-3), rel=1e-6)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["value"] == d["value"] and c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["sample"]
    assert d["generated_c"]["evals_per_s"] == d["value"]
    a = d["as_shipped_python"]
    if "unavailable" not in a:  # baseline/_ref is built in the development container (oracle/ref_harness/install_baseline.sh)
        assert a["cython_path"] is True and a["workers"] >= 1
        assert 0 < a["evals_per_s_1_core"] <= a["evals_per_s_all_cores"] * 1.5
        assert a["evals_per_s_1_core"] < d["value"]  # the Python loop is far below its own generated C


def test_our_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "3", "--warmup", "3")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]  # and prints no bench line


def test_bench_config_is_shared_by_both_arms():
    c1, c8 = bench.bench_config(1), bench.bench_config(8)
    assert c1["workload"] == bench.WORKLOAD and c1["batch_per_gpu"] == bench.B_PER_GPU
    assert c8["global_batch"] == 8 * bench.B_PER_GPU and set(c1) == set(c8)
