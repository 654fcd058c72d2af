#!/usr/bin/env python
"""Small invocations of every kernel family in one process (a quick all-kernels check; also usable under
compute-sanitizer where that tool is available: compute-sanitizer --tool memcheck python scripts/all_kernels_smoke.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import (BatchSSWs, GenericSSWs, kalman_batch, kalman_smoother_batch, lyapunov_batch,   # noqa: E402
                                   model_description, set_kalman_variant)
from dynare_jl_b200.priors import PriorSet                                                                        # noqa: E402
from dynare_jl_b200.samplers import SMC, rwmh                                                                     # noqa: E402
from oracle import models, pipeline as pl                                                                         # noqa: E402

for name in ("fs2000", "ls2003", "hs_nk"):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.npz")))
    pri = PriorSet.from_description(model_description(name)["priors"])
    ws = BatchSSWs(name, g["Y"])
    th = g["theta"][:77]
    for solver in (1, 2):
        ws.set_options(solver=solver)
        ll, st = ws.loglikelihood(th)
        ok = g["status"][:77] == 0
        assert np.array_equal(st == 0, ok) and np.allclose(ll[ok], g["loglik"][:77][ok], rtol=1e-8)
    ws.set_options(solver=0)
    ws.first_order(th); ws.state_space(th); ws.prior_predictive(th[:9], periods=8); ws.smoother(th[:5])
    ws.set_priors(pri); ws.logposterior(th)
    s = SMC(ws, pri, npart=1100, nphi=16, seed=1)
    s.initialise(np.random.default_rng(0)); s.run(); s.stats(); s.particles(); s.close()
    y0 = ws.transform(th[ok][:3], inverse=True)
    rwmh(ws, pri, y0, np.diag(np.full(ws.np, 1e-6)), mcmc_replic=3, mcmc_chains=3)
    m = models.get_model(name)
    gs = GenericSSWs(m.n, m.nx, m.i_static, m.i_bkwrd_b, m.i_fwrd_b, m.obs_idx, observations=g["Y"])
    J = []; yb = []; Se = []; Sm = []
    for t in th[ok][:6]:
        p, se, sm_ = pl.set_estimated_parameters(m, t)
        y = m.steady_state(p); J.append(m.compressed_jacobian(y, p)); yb.append(y); Se.append(se); Sm.append(sm_)
    gs.first_order(np.array(J), np.array(Se)); gs.loglikelihood(np.array(J), np.array(yb), np.array(Se), np.array(Sm))
    gs.close(); ws.close()
    print(name, "ok")

rng = np.random.default_rng(0)
import scipy.linalg as sla
for ns, ny, variant in ((12, 3, 1), (24, 5, 2), (40, 7, 2), (97, 6, 0)):
    n = 3
    T = np.empty((n, ns, ns)); Q = np.empty((n, ns, ns)); P = np.empty((n, ns, ns))
    for i in range(n):
        Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        T[i] = Qo @ np.diag(rng.uniform(0.3, 0.9, ns)) @ Qo.T
        R = 0.1 * rng.standard_normal((ns, ny)); Q[i] = R @ R.T; P[i] = sla.solve_discrete_lyapunov(T[i], Q[i])
    Y = rng.standard_normal((ny, 7)); Y[0, 2] = np.nan
    set_kalman_variant(variant)
    kalman_batch(Y, np.arange(ny), T, Q, P, H=np.tile(0.01 * np.eye(ny), (n, 1, 1)))
    set_kalman_variant(0)
    if ns <= 40:
        kalman_smoother_batch(Y, np.arange(ny), T, np.tile(np.eye(ns)[:, :ny] * 0.1, (n, 1, 1)), np.tile(np.eye(ny), (n, 1, 1)), P)
    print("dense", ns, "ok")
A = 0.5 * np.tile(np.eye(6), (4, 1, 1)); lyapunov_batch(A, np.tile(np.eye(6), (4, 1, 1)))
# exact-diffuse filter / smoother: a random walk plus noise observed through a second state
from dynare_jl_b200.engine import diffuse_initial_covariances, kalman_diffuse_smoother_batch   # noqa: E402
Td = np.tile(np.array([[1.0, 0.0], [1.0, 0.0]]), (2, 1, 1)); Rd = np.tile(np.array([[1.0, 0.0], [1.0, 0.3]]), (2, 1, 1))
Qd = np.tile(np.eye(2), (2, 1, 1))
Pinf, Pstar, _ = diffuse_initial_covariances(Td, Rd, Qd)
kalman_diffuse_smoother_batch(rng.standard_normal((1, 9)), [1], Td, Rd, Qd, Pinf, Pstar)
print("all ok")
