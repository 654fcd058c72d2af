#!/usr/bin/env python
"""Accuracy of the cooperative solver's two paths on widened-prior stress draws: policy matrices of solver 0 (verified fast
pass without pivoting + Householder pass) and solver 2 (Householder pass alone) against the QZ oracle.
Usage (GPU box): python -W ignore scripts/fast_pass_accuracy.py [model] [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynare_jl_b200.engine import BatchSSWs, model_description
from dynare_jl_b200.priors import PriorSet
from oracle import models, pipeline as pl

name = sys.argv[1] if len(sys.argv) > 1 else "fs2000"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.npz")))
m = models.get_model(name)
pri = PriorSet.from_description(model_description(name)["priors"])
rng = np.random.default_rng(77)
th = pri.sample(rng, n)
mu = th.mean(0)
wide = rng.random(n) < 0.5
th[wide] = mu + 2.5 * (th[wide] - mu)
ws = BatchSSWs(name, g["Y"])
res = {}
for solver in (2, 0):
    ws.set_options(solver=solver)
    r0 = ws.solver_redo_count()
    fo = ws.first_order(th)
    res[solver] = (fo["status"].copy(), np.concatenate([fo["g1_1"], fo["g1_2"]], axis=2), ws.solver_redo_count() - r0)
ref = np.full(res[0][1].shape, np.nan)
rst = np.zeros(n, dtype=int)
todo = range(n)
if n > 5000:          # large runs: the oracle only for the draws on which the two solvers differ most
    both = (res[0][0] == 0) & (res[2][0] == 0)
    dd = np.where(both, np.abs(res[0][1] - res[2][1]).reshape(n, -1).max(1) / np.maximum(1.0, np.abs(res[2][1]).reshape(n, -1).max(1)), 0.0)
    todo = np.argsort(-dd)[:40]
    rst[:] = -1
    rst[todo] = 0
with np.errstate(all="ignore"):
    for i in todo:
        try:
            d = pl.solve_draw(m, th[i])
            ref[i] = np.concatenate([d["sol"]["g1_1"], d["sol"]["g1_2"]], axis=1)
        except pl.DrawFailure as f:
            rst[i] = f.status
print(name, n, "draws; statuses equal between the two solvers:", np.array_equal(res[0][0], res[2][0]), "; redone by the second pass:", res[0][2])
ok = (res[0][0] == 0) & (res[2][0] == 0) & (rst == 0)
scale = np.maximum(1.0, np.abs(ref[ok]).reshape(ok.sum(), -1).max(1))
e0 = np.abs(res[0][1][ok] - ref[ok]).reshape(ok.sum(), -1).max(1) / scale
e2 = np.abs(res[2][1][ok] - ref[ok]).reshape(ok.sum(), -1).max(1) / scale
d02 = np.abs(res[0][1][ok] - res[2][1][ok]).reshape(ok.sum(), -1).max(1) / scale
q = [0.5, 0.9, 0.99, 0.999, 1.0]
print("accepted by all three:", int(ok.sum()))
print("  error vs QZ oracle, fast pass      quantiles", q, np.quantile(e0, q))
print("  error vs QZ oracle, Householder    quantiles", q, np.quantile(e2, q))
print("  fast pass vs Householder           quantiles", q, np.quantile(d02, q))
w = np.argsort(-d02)[:5]
idx = np.flatnonzero(ok)[w]
for i, j in zip(idx, w):
    print(f"  draw {i}: fast-vs-QR {d02[j]:.2e}, fast-vs-oracle {e0[j]:.2e}, QR-vs-oracle {e2[j]:.2e}, max|g1| {np.abs(ref[i]).max():.2e}")
