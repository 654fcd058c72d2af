"""Size-independent properties at the sizes BASELINE.json names for configs 3, 4 and 5 (config 2 is covered in
test_gpu_loglik.py): results do not depend on the position in the batch, repeats are bit-identical, a sub-sample agrees
with the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg as sla

from conftest import ROOT
from oracle import pipeline as pl

pytestmark = pytest.mark.gpu


def _state_spaces(rng, n, ns, ny):
    T = np.empty((n, ns, ns)); Q = np.empty((n, ns, ns)); P = np.empty((n, ns, ns))
    for i in range(n):
        Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        T[i] = Qo @ np.diag(rng.uniform(0.3, 0.97, ns)) @ Qo.T
        R = 0.1 * rng.standard_normal((ns, ny))
        Q[i] = R @ R.T
        P[i] = sla.solve_discrete_lyapunov(T[i], Q[i])
    return T, Q, P


@pytest.mark.parametrize("ns,ny,nobs,n,distinct", [(40, 7, 200, 32768, 16),      # config 4: 32 768 chains
                                                    (200, 20, 500, 296, 2)])      # config 5 shape, two draws per SM
def test_dense_filter_at_config_sizes(ns, ny, nobs, n, distinct):
    from dynare_jl_b200.engine import kalman_batch
    rng = np.random.default_rng(ns)
    T, Q, P = _state_spaces(rng, distinct, ns, ny)
    H = 0.01 * np.eye(ny)
    Y = rng.standard_normal((ny, nobs))
    idx = rng.integers(0, distinct, n); idx[:distinct] = np.arange(distinct)
    ll, st = kalman_batch(Y, np.arange(ny), T[idx], Q[idx], P[idx], H=np.tile(H, (n, 1, 1)))
    assert np.all(st == 0)
    for d in range(distinct):                      # every copy of a draw gives the same bits, wherever it sits
        assert np.unique(ll[idx == d]).size == 1
    Z = np.eye(ny, ns)
    for d in range(min(distinct, 3)):
        ref = pl.kalman_likelihood(Y, Z, H, T[d], np.eye(ns), Q[d], np.zeros(ns), P[d])
        assert abs(ll[d] - ref) <= 1e-8 * abs(ref)
    ll2, _ = kalman_batch(Y, np.arange(ny), T[idx], Q[idx], P[idx], H=np.tile(H, (n, 1, 1)))
    assert np.array_equal(ll, ll2)


def test_smc_config3_properties(golden):
    """ls2003, 16 384 particles (config 3), the first 25 of the 400 stages of the reference's schedule."""
    from dynare_jl_b200.engine import BatchSSWs, model_description
    from dynare_jl_b200.priors import PriorSet
    from dynare_jl_b200.samplers import SMC
    g = golden("ls2003")
    pri = PriorSet.from_description(model_description("ls2003")["priors"])
    ws = BatchSSWs("ls2003", g["Y"])
    runs = []
    for _ in range(2):
        s = SMC(ws, pri, npart=16384, nphi=400, seed=7)
        s.initialise(np.random.default_rng(1))
        s.run(25)
        runs.append((s.particles(), s.stats()))
        phi = s.phi
        s.close()
    p, st = runs[0]
    assert np.array_equal(p["para"], runs[1][0]["para"]) and np.array_equal(st["stats"], runs[1][1]["stats"])
    assert abs(p["w"].sum() - 1.0) < 1e-12 and np.all(p["w"] >= 0)
    ess = st["ESS"][1:26]
    assert np.all(ess > 1.0) and np.all(ess <= 16384 * (1 + 1e-12))
    assert np.all((st["acpt"][1:26] >= 0) & (st["acpt"][1:26] <= 1)) and st["clamped"][25] == 0
    # bookkeeping of the tempered posterior kernel: logpost = phi_25 * loglh + logprior for every particle
    lp, ll, status = ws.logposterior(p["para"])
    assert np.all(status == 0)
    assert np.allclose(p["loglh"], ll, rtol=1e-12, atol=0)
    assert np.allclose(p["logpost"], phi[25] * ll + (lp - ll), rtol=1e-9, atol=1e-8)
    ws.close()


def test_smc_is_independent_of_the_number_of_gpus():
    """scripts/smc_multigpu.py --check on two GPUs when the box has them: particles and records bit-identical."""
    from dynare_jl_b200 import _lib as L
    import ctypes
    c = ctypes.c_int(0)
    L.load().dyb_device_count(ctypes.byref(c))
    if c.value < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29591", os.path.join(ROOT, "scripts", "smc_multigpu.py"), "--npart", "4096", "--nphi", "30", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"identical_to_single_gpu": true' in r.stdout
