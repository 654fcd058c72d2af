"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       env={**os.environ, "OMP_NUM_THREADS": "1"})      # torchrun's default must not throttle the CPU arm
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 600)
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("loglik evals/sec") and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--cpu-budget", "1"], 900)
    assert COMMON <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 1e6 and d["gpu_launches"] >= 4 * 3
    e = d["e2e"]
    assert e["value"] > 1e6 and e["h2d_bytes_per_step"] == 65536 * 9 * 8 and e["d2h_bytes_per_step"] == 65536 * 12
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert {k["name"].split("<")[0] for k in r["kernels"]} == {"model_kernel", "cr_kernel", "assemble_kernel", "kalman_kernel"}
    assert abs(sum(r["kernel_share_of_step"].values()) - 1.0) < 0.05
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1 and "sample" in cb
