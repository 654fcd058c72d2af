"""Batched Kalman smoother (dense state spaces and the model-level calibsmoother! for N draws) against the oracle."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import models, pipeline as pl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ns,ny,nx,nobs,n", [(5, 2, 3, 25, 9), (16, 3, 4, 40, 5), (40, 7, 7, 30, 3)])
def test_dense_smoother_matches_oracle(ns, ny, nx, nobs, n):
    from dynare_jl_b200.engine import kalman_smoother_batch
    rng = np.random.default_rng(ns)
    T = np.empty((n, ns, ns)); R = np.empty((n, ns, nx)); Q = np.empty((n, nx, nx)); P0 = np.empty((n, ns, ns))
    H = np.empty((n, ny, ny))
    for i in range(n):
        Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        T[i] = Qo @ np.diag(rng.uniform(0.2, 0.95, ns)) @ Qo.T
        R[i] = 0.3 * rng.standard_normal((ns, nx))
        A = rng.standard_normal((nx, nx)); Q[i] = A @ A.T / nx + 0.1 * np.eye(nx)
        P0[i] = sla.solve_discrete_lyapunov(T[i], R[i] @ Q[i] @ R[i].T)
        B = 0.1 * rng.standard_normal((ny, ny)); H[i] = B @ B.T + 0.02 * np.eye(ny)
    z = rng.permutation(ns)[:ny]
    Y = rng.standard_normal((ny, nobs))
    Y[0, 3] = np.nan; Y[:, 11] = np.nan
    out = kalman_smoother_batch(Y, z, T, R, Q, P0, H=H)
    Z = np.zeros((ny, ns)); Z[np.arange(ny), z] = 1
    for i in range(n):
        ref = pl.kalman_smoother(Y, Z, H[i], T[i], R[i], Q[i], np.zeros(ns), P0[i])
        assert out["status"][i] == 0
        assert abs(out["loglik"][i] - ref["loglik"]) <= 1e-9 * abs(ref["loglik"])
        sc = max(1.0, np.abs(ref["alphah"]).max())
        assert np.abs(out["alphah"][i].T - ref["alphah"]).max() <= 1e-9 * sc
        assert np.abs(out["etah"][i].T - ref["etah"]).max() <= 1e-9 * max(1.0, np.abs(ref["etah"]).max())
        assert np.abs(out["epsilonh"][i].T - ref["epsilonh"]).max() <= 1e-9
        assert np.abs(out["att"][i].T - ref["att"]).max() <= 1e-9 * sc
        assert np.abs(out["a_pred"][i].T - ref["a_pred"]).max() <= 1e-9 * sc


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_model_smoother_matches_oracle(name, golden):
    """calibsmoother! for a batch of draws: every endogenous variable is a state, outputs in levels."""
    from dynare_jl_b200.engine import BatchSSWs
    g = golden(name)
    m = models.get_model(name)
    Y = g["Y"].copy()
    Y[0, 5] = np.nan
    ws = BatchSSWs(name, Y)
    theta = g["theta"][:24]
    out = ws.smoother(theta)
    ok = g["status"][:24] == 0
    assert np.array_equal(out["status"] == 0, ok)
    Z = np.zeros((m.ny, m.n)); Z[np.arange(m.ny), m.obs_idx] = 1
    checked = 0
    for i in np.flatnonzero(ok)[:10]:
        d = pl.solve_draw(m, theta[i])
        T = np.zeros((m.n, m.n)); T[:, m.i_bkwrd_b] = d["sol"]["g1_1"]
        R = d["sol"]["g1_2"]
        Yd = Y - d["ybar"][m.obs_idx][:, None]
        ref = pl.kalman_smoother(Yd, Z, d["sigma_m"], T, R, d["sigma_e"], np.zeros(m.n), d["Sigma_y"])
        lev = ref["alphah"] + d["ybar"][:, None]
        sc = max(1.0, np.abs(lev).max())
        assert np.abs(out["alphah"][i].T - lev).max() <= 1e-8 * sc, (i, np.abs(out["alphah"][i].T - lev).max())
        assert np.abs(out["etah"][i].T - ref["etah"]).max() <= 1e-7 * max(1.0, np.abs(ref["etah"]).max())
        assert abs(out["loglik"][i] - ref["loglik"]) <= 1e-8 * abs(ref["loglik"])
        # the smoothed observables reproduce the data wherever it is observed (up to the smoothed measurement error)
        fit = out["alphah"][i][:, m.obs_idx].T + out["epsilonh"][i].T
        obs = ~np.isnan(Y)
        assert np.abs((fit - Y)[obs]).max() <= 1e-7 * max(1.0, np.abs(Y[obs]).max())
        checked += 1
    assert checked >= 5
    ws.close()


@pytest.mark.parametrize("name", ["fs2000", "hs_nk"])
def test_forecast_from_last_smoothed_state(name, golden):
    """forecasting_ in calibsmoother mode (forecast.jl:54-115): y0 = last smoothed state, zero future shocks."""
    from dynare_jl_b200.engine import BatchSSWs
    g = golden(name)
    m = models.get_model(name)
    ws = BatchSSWs(name, g["Y"])
    theta = g["theta"][:16]
    sm = ws.smoother(theta)
    y0 = np.where(np.isfinite(sm["alphah"][:, -1, :]), sm["alphah"][:, -1, :], 0.0)
    Y, st = ws.forecast(theta, y0, 12)
    assert np.array_equal(st != 0, sm["status"] != 0) or np.all((st != 0) <= (sm["status"] != 0))
    checked = 0
    for i in range(16):
        if st[i] != 0:
            assert np.all(np.isnan(Y[i]))
            continue
        d = pl.solve_draw(m, theta[i])
        y = y0[i].copy()
        for h in range(13):
            assert np.allclose(Y[i, h], y, rtol=1e-10, atol=1e-12)
            y = d["ybar"] + d["sol"]["g1_1"] @ (y - d["ybar"])[m.i_bkwrd_b]
        checked += 1
    assert checked >= 8
    ws.close()
