#!/usr/bin/env python
"""Stress check beyond the test tolerances: draws from WIDENED priors (deviations from the prior mean x 2.5, policy
parameters pushed across the determinacy bound) through the engine and through the QZ oracle; prints the cross-table of
status codes, the number of OK-vs-failed disagreements (must be 0) and the worst relative log-likelihood errors with
the spectral radius / size of the initial covariance of those draws (ill-conditioned draws are where 1e-8 is exceeded).
Usage (GPU box): python -W ignore scripts/stress_status.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynare_jl_b200.engine import BatchSSWs, model_description
from dynare_jl_b200.priors import PriorSet
from oracle import models, pipeline as pl
for name, n in (("ls2003", 3000), ("hs_nk", 3000), ("fs2000", 1500)):
    g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"{name}_golden.npz")))
    m = models.get_model(name)
    pri = PriorSet.from_description(model_description(name)["priors"])
    rng = np.random.default_rng(123)
    th = pri.sample(rng, n)
    # widen: multiply deviations from the prior mean by 2.5 for a random half of the draws (keeps supports mostly)
    mu = th.mean(0)
    wide = rng.random(n) < 0.5
    th[wide] = mu + 2.5 * (th[wide] - mu)
    if name == "hs_nk": th[: n // 3, 5] = rng.uniform(0.5, 1.3, n // 3)      # psi1 around the Taylor-principle bound
    if name == "ls2003": th[: n // 3, 0] = rng.uniform(-1.2, 0.3, n // 3)    # psi1 (1 + psi1 multiplies inflation)
    ws = BatchSSWs(name, g["Y"])
    ll, st = ws.loglikelihood(th)
    t = time.time()
    ref = np.empty(n); rst = np.empty(n, dtype=int)
    with np.errstate(all="ignore"):
        for i in range(n):
            ref[i], rst[i] = pl.loglikelihood(m, th[i], g["Y"])
    dis = np.flatnonzero(st != rst)
    both = (st == 0) & (rst == 0)
    rel = np.abs(ll[both] - ref[both]) / np.abs(ref[both])
    print(name, "draws", n, "oracle status hist", np.bincount(rst, minlength=7), "disagreements", dis.size,
          "max rel err", rel.max() if both.any() else None, f"({time.time()-t:.0f}s oracle)")
    ct = np.zeros((7, 7), dtype=int)
    for a, b in zip(st, rst):
        ct[a, b] += 1
    print("   cross-table, rows = engine status, columns = oracle status:")
    print(ct)
    okfail = np.flatnonzero((st == 0) != (rst == 0))
    print("   OK-vs-failed disagreements:", okfail.size)
    worst = np.argsort(-np.where(both, np.abs(ll - ref) / np.abs(ref), 0))[:5]
    for i in worst:
        d = pl.solve_draw(m, th[i])
        rho = np.max(np.abs(np.linalg.eigvals(d["sol"]["g1_1"][m.i_bkwrd_b, :])))
        print("     worst rel err draw", i, abs(ll[i] - ref[i]) / abs(ref[i]), "loglik", ref[i], "rho(G)", rho, "max|P0|", np.abs(d["ss"]["P0"]).max())
    ws.close()
