#!/usr/bin/env python
"""Device-side throughput of the model-specific likelihood path for one model (not a bench line): kernel split and
evaluations per second for a batch of prior draws.  Usage: python scripts/model_throughput.py [ls2003] [65536]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import BatchSSWs, model_description   # noqa: E402
from dynare_jl_b200.priors import PriorSet                       # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ls2003"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.npz")))
pri = PriorSet.from_description(model_description(name)["priors"])
th = pri.sample(np.random.default_rng(0), n)
ws = BatchSSWs(name, g["Y"])
ws.loglikelihood(th)
ws.timers_reset()
t = time.perf_counter()
for _ in range(5):
    ll, st = ws.loglikelihood(th)
dt = (time.perf_counter() - t) / 5
m, c, a = ws.timers_read_solve_split()
k, s, f = ws.timers_read()
dev = (s + f) / k
print(f"{name} {n} draws: host-call {dt * 1e3:.3f} ms; kernels ms: model {m / k:.3f} cr {c / k:.3f} assemble {a / k:.3f} "
      f"filter {f / k:.3f} -> {n / dev / 1e3:.1f} M evals/s device-side; ok {(st == 0).sum()}; redone by the Householder pass "
      f"{ws.solver_redo_count()} of {7 * n} draw evaluations")
