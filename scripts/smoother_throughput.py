#!/usr/bin/env python
"""Throughput of the batched calibsmoother! (dyb_smoother_batch) on prior draws: python scripts/smoother_throughput.py [model] [N]."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import BatchSSWs   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fs2000"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.npz")))
th = np.resize(g["theta"][g["status"] == 0], (N, g["theta"].shape[1]))
ws = BatchSSWs(name, g["Y"])
ws.smoother(th[:64])
t0 = time.perf_counter()
out = ws.smoother(th)
dt = time.perf_counter() - t0
n, nobs = ws.dims["n"], g["Y"].shape[1]
print(f"{name}: {N} draws x {nobs} periods x {n} variables smoothed in {dt * 1e3:.1f} ms = {N / dt:.0f} draws/s "
      f"({(out['status'] == 0).sum()} ok; outputs {out['alphah'].nbytes / 1e6:.0f} MB to the host)")
