#!/usr/bin/env python
"""Kernel split of the medium-scale model path (mc4): one batched log-likelihood of prior draws; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel times.  Usage: python scripts/medium_model_split.py [8192]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import BatchSSWs, model_description   # noqa: E402
from dynare_jl_b200.priors import PriorSet                       # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mc4_golden.npz")))
pri = PriorSet.from_description(model_description("mc4")["priors"])
th = pri.sample(np.random.default_rng(0), n)
ws = BatchSSWs("mc4", g["Y"])
res = {}
for solver in (0, 1):            # 0: register-blocked kernel (solve_block.cuh), 1: shared-memory kernel (solve_generic.cuh)
    ws.set_options(solver=solver)
    ll, st = ws.loglikelihood(th)
    t = time.perf_counter()
    ll, st = ws.loglikelihood(th)
    dt = time.perf_counter() - t
    t = time.perf_counter()
    fo = ws.first_order(th)
    dtf = time.perf_counter() - t
    res[solver] = (ll, st, fo)
    print(f"mc4 {n} draws solver={solver}: loglikelihood {dt * 1e3:.2f} ms = {n / dt / 1e3:.1f} k evals/s; first_order call "
          f"{dtf * 1e3:.2f} ms; ok {(st == 0).sum()}; status histogram {np.bincount(st, minlength=7).tolist()}")
ok = (res[0][1] == 0) & (res[1][1] == 0)
print("status equal:", np.array_equal(res[0][1], res[1][1]), "max rel loglik diff:",
      np.max(np.abs(res[0][0][ok] - res[1][0][ok]) / np.abs(res[1][0][ok])), "max g1 diff:",
      max(np.abs(res[0][2][k][ok] - res[1][2][k][ok]).max() for k in ("g1_1", "g1_2")))
