"""The diffuse (unit-root) filter/smoother restatement against an independent closed form: generalised least squares in
the diffuse initial effects (no recursion), and against the ordinary smoother when nothing is diffuse."""
import numpy as np
import pytest

from diffuse_cases import trend_cycle_case
from oracle import diffuse as D
from oracle import pipeline as pl

CASES = [dict(seed=1), dict(seed=2, n_trend=1, ny=2), dict(seed=3, i2=True, ny=2), dict(seed=4, missing=False, meas_err=False),
         dict(seed=5, n_trend=3, n_cycle=4, ny=4, nobs=24)]


@pytest.mark.parametrize("kw", CASES)
def test_diffuse_smoother_matches_gls_closed_form(kw):
    c = trend_cycle_case(**kw)
    ns = c["T"].shape[0]
    Pinf, Pstar, k = D.diffuse_initial_covariances(c["T"], c["R"], c["Q"])
    assert k == (2 if kw.get("i2") else kw.get("n_trend", 2))
    assert np.linalg.matrix_rank(Pinf, tol=1e-9) == k
    out = D.diffuse_kalman_smoother(c["Y"], c["z"], c["Hd"], c["T"], c["R"], c["Q"], np.zeros(ns), Pinf, Pstar)
    ll, ah = D.diffuse_bruteforce(c["Y"], c["z"], c["Hd"], c["T"], c["R"], c["Q"], np.zeros(ns), Pinf, Pstar)
    assert out["d"] >= 1
    assert abs(out["loglik"] - ll) <= 1e-10 * abs(ll)
    assert np.abs(out["alphah"] - ah).max() <= 1e-9 * max(1.0, np.abs(ah).max())
    # states that are observed without error are reproduced exactly
    for i, z in enumerate(c["z"]):
        if c["Hd"][i] == 0.0:
            m = ~np.isnan(c["Y"][i])
            assert np.abs(out["alphah"][z, m] - c["Y"][i, m]).max() <= 1e-9 * max(1.0, np.abs(c["Y"][i, m]).max())


def test_initial_covariances_satisfy_the_reference_construction():
    """kalman.jl:163-195: Pinf0 projects on the unit-root invariant subspace, Pstar0 solves the Lyapunov equation
    there restricted to the stable block."""
    c = trend_cycle_case(7)
    T, R, Q = c["T"], c["R"], c["Q"]
    Pinf, Pstar, k = D.diffuse_initial_covariances(T, R, Q)
    assert np.allclose(Pinf @ Pinf, Pinf, atol=1e-12) and np.allclose(Pinf @ Pstar, 0, atol=1e-10)
    Pi_s = np.eye(T.shape[0]) - Pinf
    assert np.allclose(Pstar, Pi_s @ (T @ Pstar @ T.T + R @ Q @ R.T) @ Pi_s, atol=1e-10)


def test_without_unit_roots_it_is_the_ordinary_smoother():
    c = trend_cycle_case(6)
    T = c["T"].copy()
    nb = T.shape[0] - 3
    Lo = c["R"][nb:]
    T[0, 0] = 0.9; T[1, 1] = 0.7
    T[nb:, :nb] = Lo @ T[:nb, :nb]
    ns = T.shape[0]
    Pinf, Pstar, k = D.diffuse_initial_covariances(T, c["R"], c["Q"])
    assert k == 0 and np.abs(Pinf).max() == 0.0
    out = D.diffuse_kalman_smoother(c["Y"], c["z"], c["Hd"], T, c["R"], c["Q"], np.zeros(ns), Pinf, Pstar)
    Z = np.zeros((3, ns)); Z[np.arange(3), c["z"]] = 1
    ref = pl.kalman_smoother(c["Y"], Z, np.diag(c["Hd"]), T, c["R"], c["Q"], np.zeros(ns), Pstar)
    assert out["d"] == 0
    assert abs(out["loglik"] - ref["loglik"]) <= 1e-11 * abs(ref["loglik"])
    for key in ("alphah", "etah", "epsilonh", "att", "a_pred"):
        assert np.abs(out[key] - ref[key]).max() <= 1e-10 * max(1.0, np.abs(ref[key]).max()), key
