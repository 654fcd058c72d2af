"""Batched exact-diffuse filter/smoother on the GPU against the oracle restatement (oracle/diffuse.py) and, through it,
the closed-form GLS second opinion.  Tolerance 1e-9 relative (fp64; summation order differs)."""
import numpy as np
import pytest

from diffuse_cases import trend_cycle_case
from oracle import diffuse as D

pytestmark = pytest.mark.gpu

CASES = [dict(seed=1), dict(seed=2, n_trend=1, ny=2), dict(seed=3, i2=True, ny=2), dict(seed=4, missing=False, meas_err=False),
         dict(seed=5, n_trend=3, n_cycle=4, ny=4, nobs=24), dict(seed=8, n_trend=4, n_cycle=30, ny=6, nobs=60)]


def _batch(kw, n):
    cs = [trend_cycle_case(**{**kw, "seed": kw["seed"] * 100 + i}) for i in range(n)]
    Y = cs[0]["Y"]                                    # one data set, N state spaces
    return Y, cs[0]["z"], np.stack([c["T"] for c in cs]), np.stack([c["R"] for c in cs]), np.stack([c["Q"] for c in cs]), \
        np.stack([c["Hd"] for c in cs])


@pytest.mark.parametrize("kw", CASES)
def test_diffuse_smoother_matches_oracle(kw):
    from dynare_jl_b200.engine import diffuse_initial_covariances, kalman_diffuse_smoother_batch
    n = 5
    Y, z, T, R, Q, Hd = _batch(kw, n)
    ns = T.shape[1]
    Pinf, Pstar, ks = diffuse_initial_covariances(T, R, Q)
    out = kalman_diffuse_smoother_batch(Y, z, T, R, Q, Pinf, Pstar, Hdiag=Hd)
    for i in range(n):
        Pi_o, Ps_o, k_o = D.diffuse_initial_covariances(T[i], R[i], Q[i])
        assert ks[i] == k_o
        assert np.abs(Pinf[i] - Pi_o).max() <= 1e-12 and np.abs(Pstar[i] - Ps_o).max() <= 1e-9 * max(1.0, np.abs(Ps_o).max())
        ref = D.diffuse_kalman_smoother(Y, z, Hd[i], T[i], R[i], Q[i], np.zeros(ns), Pi_o, Ps_o)
        assert out["status"][i] == 0 and out["n_diffuse"][i] == ref["d"]
        assert abs(out["loglik"][i] - ref["loglik"]) <= 1e-9 * abs(ref["loglik"])
        for key in ("alphah", "etah", "epsilonh", "att", "a_pred"):
            sc = max(1.0, np.abs(ref[key]).max())
            assert np.abs(out[key][i].T - ref[key]).max() <= 1e-9 * sc, key


def test_diffuse_smoother_with_nothing_diffuse_equals_the_ordinary_one():
    from dynare_jl_b200.engine import kalman_diffuse_smoother_batch, kalman_smoother_batch
    import scipy.linalg as sla
    rng = np.random.default_rng(11)
    n, ns, ny, nx, nobs = 4, 12, 3, 4, 35
    T = np.empty((n, ns, ns)); R = np.empty((n, ns, nx)); Q = np.empty((n, nx, nx)); P0 = np.empty((n, ns, ns))
    for i in range(n):
        Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        T[i] = Qo @ np.diag(rng.uniform(0.2, 0.95, ns)) @ Qo.T
        R[i] = 0.3 * rng.standard_normal((ns, nx))
        Q[i] = np.diag(rng.uniform(0.2, 1.0, nx))
        P0[i] = sla.solve_discrete_lyapunov(T[i], R[i] @ Q[i] @ R[i].T)
    Hd = rng.uniform(0.01, 0.1, (n, ny))
    z = rng.permutation(ns)[:ny]
    Y = rng.standard_normal((ny, nobs)); Y[1, 2] = np.nan; Y[:, 9] = np.nan
    a = kalman_diffuse_smoother_batch(Y, z, T, R, Q, np.zeros_like(P0), P0, Hdiag=Hd)
    b = kalman_smoother_batch(Y, z, T, R, Q, P0, H=np.stack([np.diag(h) for h in Hd]))
    assert (a["status"] == 0).all() and (a["n_diffuse"] == 0).all()
    assert np.abs(a["loglik"] - b["loglik"]).max() <= 1e-10 * np.abs(b["loglik"]).max()
    for key in ("alphah", "etah", "epsilonh", "att", "a_pred"):
        assert np.abs(a[key] - b[key]).max() <= 1e-9 * max(1.0, np.abs(b[key]).max()), key


def test_diffuse_smoother_rejects_bad_arguments():
    from dynare_jl_b200 import _lib as L
    from dynare_jl_b200.engine import kalman_diffuse_smoother_batch
    c = trend_cycle_case(1)
    T, R, Q = c["T"][None], c["R"][None], c["Q"][None]
    P = np.zeros_like(T)
    with pytest.raises(L.DybError):
        kalman_diffuse_smoother_batch(c["Y"], c["z"], T, R, Q, P, P, tol=0.0)
    with pytest.raises(L.DybError):
        kalman_diffuse_smoother_batch(c["Y"], np.array([0, 1, 99]), T, R, Q, P, P)
