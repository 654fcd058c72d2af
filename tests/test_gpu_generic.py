"""Run-time-size path for models without a generated device function: the oracle's complex-step Jacobians go in,
policy matrices / state space / log-likelihood are compared with the QZ oracle; plus the batched Lyapunov solver."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import models, pipeline as pl

pytestmark = pytest.mark.gpu


def _host_side(m, theta):
    """What the Julia host computes per draw: parameter scatter, steady state (+ check), compressed Jacobian."""
    n = theta.shape[0]
    ncol = m.k + m.n + m.nf + m.nx
    J = np.zeros((n, m.n, ncol)); yb = np.zeros((n, m.n)); Se = np.zeros((n, m.nx, m.nx)); Sm = np.zeros((n, m.ny, m.ny))
    st = np.zeros(n, dtype=np.int32)
    for i in range(n):
        params, Se[i], Sm[i] = pl.set_estimated_parameters(m, theta[i])
        with np.errstate(all="ignore"):
            y = m.steady_state(params)
            if not np.all(np.isfinite(y)) or not (np.linalg.norm(m.static_resid(y, params)) <= pl.STEADY_TOL):
                st[i] = pl.ST_STEADY
                continue
            yb[i] = y
            J[i] = m.compressed_jacobian(y, params)
    return J, yb, Se, Sm, st


@pytest.mark.parametrize("compiled", [False, True])
@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_generic_path_matches_oracle(name, compiled, golden):
    """compiled = False: the run-time-size kernels; True: the register-resident kernels compiled for the model's sizes and
    index sets alone (a "sizes-only" model, build.build_sizes_library), fed with the same host-evaluated Jacobians."""
    from dynare_jl_b200.engine import GenericSSWs
    g = golden(name)
    m = models.get_model(name)
    theta = g["theta"][:96]
    J, yb, Se, Sm, st_host = _host_side(m, theta)
    ws = GenericSSWs(m.n, m.nx, m.i_static, m.i_bkwrd_b, m.i_fwrd_b, m.obs_idx, observations=g["Y"], compile=compiled)
    assert ws.compiled == compiled
    assert ws.dims["ns"] == m.ns and ws.dims["k"] == m.k and ws.dims["ncol"] == J.shape[2]
    fo = ws.first_order(J, Se)
    ll, st = ws.loglikelihood(J, yb, Se, Sm, status_in=st_host)
    assert np.array_equal(st, g["status"][:96])
    nchecked = 0
    for i in range(theta.shape[0]):
        if g["status"][i] != 0:
            assert ll[i] == -np.inf
            continue
        d = pl.solve_draw(m, theta[i])
        assert np.abs(fo["g1_1"][i] - d["sol"]["g1_1"]).max() < 1e-10
        assert np.abs(fo["g1_2"][i] - d["sol"]["g1_2"]).max() < 1e-10
        ss = d["ss"]
        assert np.abs(fo["T"][i] - ss["T"]).max() < 1e-10
        rqr = ss["R"] @ ss["Q"] @ ss["R"].T
        assert np.allclose(fo["RQR"][i], rqr, rtol=1e-9, atol=1e-13 * max(1.0, np.abs(rqr).max()))
        sc = max(1.0, np.abs(ss["P0"]).max())
        assert np.abs(fo["P0"][i] - ss["P0"]).max() <= 1e-9 * sc
        assert abs(ll[i] - g["loglik"][i]) <= 1e-8 * abs(g["loglik"][i])
        nchecked += 1
    assert nchecked > 40
    ws.close()


def test_generic_handle_rejects_theta_entry_points(golden):
    from dynare_jl_b200 import _lib as L
    from dynare_jl_b200.engine import GenericSSWs
    m = models.get_model("fs2000")
    ws = GenericSSWs(m.n, m.nx, m.i_static, m.i_bkwrd_b, m.i_fwrd_b, m.obs_idx)
    th = np.zeros((1, 9)); ll = np.zeros(1); st = np.zeros(1, dtype=np.int32)
    assert ws.lib.dyb_loglik_batch(ws._h, L.ptr(th), 1, L.ptr(ll), L.ptr(st)) == 5       # DYB_ERR_UNSUPPORTED
    with pytest.raises(L.DybError):
        GenericSSWs(4, 1, [0, 1], [1], [2], [0])            # a static variable cannot have a lag
    ws.close()


@pytest.mark.parametrize("k", [1, 4, 20, 60])
def test_lyapunov_batch(k):
    from dynare_jl_b200.engine import lyapunov_batch
    rng = np.random.default_rng(k)
    n = 17
    A = np.empty((n, k, k)); B = np.empty((n, k, k))
    for i in range(n):
        Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        A[i] = Q @ np.diag(rng.uniform(-0.95, 0.95, k)) @ Q.T + 0.02 * rng.standard_normal((k, k)) / k
        R = rng.standard_normal((k, max(1, k // 2)))
        B[i] = R @ R.T
    A[3] = 1.2 * np.eye(k)                                 # explosive: no finite variance
    X, it, st = lyapunov_batch(A, B)
    for i in range(n):
        if i == 3:
            assert st[i] == pl.ST_NONSTATIONARY
            continue
        assert st[i] == 0 and it[i] >= 1
        ref = sla.solve_discrete_lyapunov(A[i], B[i])
        assert np.allclose(X[i], ref, rtol=1e-9, atol=1e-11)
        assert np.abs(X[i] - (A[i] @ X[i] @ A[i].T + B[i])).max() < 1e-10 * max(1.0, np.abs(X[i]).max())   # test_smc.jl:41-42
