"""Known-answer tests that pin the CPU oracle (the reference itself pins nothing)."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import models, pipeline as pl


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_quadratic_residual_and_stability(name):
    m = models.get_model(name)
    d = pl.solve_draw(m, m.theta_mode)
    A, B, C, Fu = m.abcf(d["ybar"], d["params"])
    G = d["sol"]["G"]
    assert np.abs(A @ G @ G + B @ G + C).max() < 1e-12
    assert np.max(np.abs(np.linalg.eigvals(G))) < 1.0
    # columns of G outside the predetermined variables are structurally zero
    other = [j for j in range(m.n) if j not in m.i_bkwrd_b]
    assert np.abs(G[:, other]).max() < 1e-12
    # shock response solves (A G + B) g_u = -F_u
    assert np.abs((A @ G + B) @ d["sol"]["g1_2"] + Fu).max() < 1e-12


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_qz_and_cyclic_reduction_agree(name):
    m = models.get_model(name)
    a = pl.solve_draw(m, m.theta_mode, "GS")["sol"]
    b = pl.solve_draw(m, m.theta_mode, "CR")["sol"]
    assert np.abs(a["g1_1"] - b["g1_1"]).max() < 1e-11
    assert np.abs(a["g1_2"] - b["g1_2"]).max() < 1e-11


def test_fs2000_eigenvalues_at_calibration():
    """fs2000 at the .mod calibration: finite non-zero generalised eigenvalue moduli 0.7, 0.9467, 1.0670;
    7 unstable (incl. infinite) roots for 7 forward variables."""
    m = models.get_model("fs2000")
    theta = np.array([0.33, 0.99, 0.003, 1.011, 0.7, 0.787, 0.02, 0.014, 0.005])
    d = pl.solve_draw(m, theta)
    eig = d["sol"]["eig"]
    finite = eig[(eig > 1e-8) & (eig < 1e6)]
    assert np.allclose(finite, [0.7, 0.9467, 1.0670], atol=5e-4)
    assert int(np.sum(eig >= pl.QZ_CRITERIUM)) == 2 * m.n - m.n      # n unstable of 2n in the companion form
    assert int(np.sum((eig >= pl.QZ_CRITERIUM))) - (m.n - m.nf) == m.nf


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_lyapunov_identity(name):
    """P0 = T P0 T' + R Q R'  (pattern of reference test/test_smc.jl:29-42)."""
    m = models.get_model(name)
    ss = pl.solve_draw(m, m.theta_mode)["ss"]
    rhs = ss["T"] @ ss["P0"] @ ss["T"].T + ss["R"] @ ss["Q"] @ ss["R"].T
    assert np.abs(ss["P0"] - rhs).max() < 1e-10 * max(1.0, np.abs(ss["P0"]).max())


def test_kalman_matches_bruteforce_gaussian():
    m = models.get_model("ls2003")
    d = pl.solve_draw(m, m.theta_mode)
    ss = d["ss"]
    Y = pl.simulate_observations(m, m.theta_mode, 14, 7)
    Yd = Y - d["ybar"][m.obs_idx][:, None]
    a = pl.kalman_likelihood(Yd, ss["Z"], ss["H"], ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"])
    b = pl.gaussian_loglik_bruteforce(Yd, ss["Z"], ss["H"], ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"])
    assert abs(a - b) < 1e-9 * abs(b)


def test_kalman_bruteforce_with_measurement_error_fs2000():
    m = models.get_model("fs2000")
    d = pl.solve_draw(m, m.theta_mode)
    ss = d["ss"]
    H = np.diag([1e-5, 2e-5])
    Y = pl.simulate_observations(m, m.theta_mode, 10, 3)
    Yd = Y - d["ybar"][m.obs_idx][:, None]
    a = pl.kalman_likelihood(Yd, ss["Z"], H, ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"])
    b = pl.gaussian_loglik_bruteforce(Yd, ss["Z"], H, ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"])
    assert abs(a - b) < 1e-8 * abs(b)


def test_kalman_missing_observations_match_marginal_density():
    """Dropping an observation must equal integrating it out of the joint Gaussian."""
    m = models.get_model("ls2003")
    d = pl.solve_draw(m, m.theta_mode)
    ss = d["ss"]
    Y = pl.simulate_observations(m, m.theta_mode, 6, 11)
    Yd = Y - d["ybar"][m.obs_idx][:, None]
    Ym = Yd.copy()
    Ym[1, 2] = np.nan
    Ym[:, 4] = np.nan
    a = pl.kalman_likelihood(Ym, ss["Z"], ss["H"], ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"])
    # marginal of the stacked Gaussian on the observed entries
    ny, nobs = Yd.shape
    QQ = ss["R"] @ ss["Q"] @ ss["R"].T
    V = [ss["P0"]]
    for t in range(1, nobs):
        V.append(ss["T"] @ V[-1] @ ss["T"].T + QQ)
    big = np.zeros((ny * nobs, ny * nobs))
    for s_ in range(nobs):
        Cm = V[s_].copy()
        for t in range(s_, nobs):
            if t > s_:
                Cm = ss["T"] @ Cm
            blk = ss["Z"] @ Cm @ ss["Z"].T + (ss["H"] if t == s_ else 0)
            big[t * ny:(t + 1) * ny, s_ * ny:(s_ + 1) * ny] = blk
            big[s_ * ny:(s_ + 1) * ny, t * ny:(t + 1) * ny] = blk.T
    yv = Ym.T.reshape(-1)
    keep = ~np.isnan(yv)
    S = big[np.ix_(keep, keep)]
    cF = np.linalg.cholesky(S)
    w = sla.solve_triangular(cF, yv[keep], lower=True)
    b = -0.5 * (keep.sum() * np.log(2 * np.pi) + 2 * np.log(np.diag(cF)).sum() + w @ w)
    assert abs(a - b) < 1e-9 * abs(b)


def test_status_classification():
    ls = models.get_model("ls2003")
    th = ls.theta_mode.copy()
    th[0] = -1.5                      # (1 + psi1) < 0: Taylor principle violated -> indeterminacy
    ll, st = pl.loglikelihood(ls, th, np.zeros((5, 4)))
    assert st == pl.ST_INDETERMINATE and ll == -np.inf
    fs = models.get_model("fs2000")
    th = fs.theta_mode.copy()
    th[0] = 1.5                       # alp > 1: log / power of a negative number in the steady state
    ll, st = pl.loglikelihood(fs, th, np.zeros((2, 4)))
    assert st == pl.ST_STEADY and ll == -np.inf
    th = ls.theta_mode.copy()
    th[8] = 1.2                       # rho_q > 1: explosive exogenous process -> no stable solution
    ll, st = pl.loglikelihood(ls, th, np.zeros((5, 4)))
    assert st == pl.ST_UNSTABLE and ll == -np.inf


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_golden_matches_live_oracle(name, golden):
    g = golden(name)
    m = models.get_model(name)
    for i in [0, 1, 17, 100, 255]:
        ll, st = pl.loglikelihood(m, g["theta"][i], g["Y"])
        assert st == g["status"][i]
        assert abs(ll - g["loglik"][i]) <= 1e-12 * abs(ll)


def test_smoother_oracle_matches_bruteforce_conditional_expectation():
    """The recursive smoother against E[alpha | Y] from the stacked joint Gaussian law; and structural identities
    with missing observations (states consistent with smoothed shocks)."""
    rng = np.random.default_rng(4)
    ns, ny, nx, nobs = 5, 2, 3, 12
    Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
    T = Qo @ np.diag(rng.uniform(0.2, 0.9, ns)) @ Qo.T
    R = rng.standard_normal((ns, nx)) * 0.3
    Q = np.diag(rng.uniform(0.5, 1.5, nx))
    H = 0.05 * np.eye(ny)
    Z = np.zeros((ny, ns)); Z[0, 3] = 1; Z[1, 1] = 1
    import scipy.linalg as sla
    P0 = sla.solve_discrete_lyapunov(T, R @ Q @ R.T)
    Y = rng.standard_normal((ny, nobs))
    sm = pl.kalman_smoother(Y, Z, H, T, R, Q, np.zeros(ns), P0)
    ref = pl.gaussian_smoother_bruteforce(Y, Z, H, T, R, Q, np.zeros(ns), P0)
    assert np.abs(sm["alphah"] - ref).max() < 1e-10
    assert abs(sm["loglik"] - pl.kalman_likelihood(Y, Z, H, T, R, Q, np.zeros(ns), P0)) < 1e-10
    # smoothed states obey the transition equation with the smoothed shocks, also with missing data
    Y[0, 4] = np.nan; Y[:, 7] = np.nan
    sm = pl.kalman_smoother(Y, Z, H, T, R, Q, np.zeros(ns), P0)
    for t in range(nobs - 1):
        assert np.abs(sm["alphah"][:, t + 1] - (T @ sm["alphah"][:, t] + R @ sm["etah"][:, t])).max() < 1e-10
    obs = ~np.isnan(Y)
    fit = Z @ sm["alphah"] + sm["epsilonh"]
    assert np.abs((fit - Y)[obs]).max() < 1e-10                  # y = Z alpha_hat + eps_hat where observed


def test_univariate_fallback_of_the_oracle_filter():
    """KalmanFilterTools' fallback as restated in oracle/pipeline.py: (i) all periods univariate == multivariate when H is
    diagonal and F is positive definite; (ii) an exactly singular F rejects the draw with the fallback off and scores the
    non-degenerate observation with it on (closed form)."""
    rng = np.random.default_rng(11)
    ns, ny, nobs = 5, 3, 30
    Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
    T = Qo @ np.diag(rng.uniform(0.2, 0.9, ns)) @ Qo.T
    R = 0.3 * rng.standard_normal((ns, ny)); Q = np.eye(ny)
    H = np.diag([0.01, 0.0, 0.02])
    Z = np.eye(ny, ns)
    import scipy.linalg as sla
    P0 = sla.solve_discrete_lyapunov(T, R @ R.T)
    Y = rng.standard_normal((ny, nobs)); Y[1, 4] = np.nan
    a = pl.kalman_likelihood(Y, Z, H, T, R, Q, np.zeros(ns), P0)
    b = pl.kalman_likelihood(Y, Z, H, T, R, Q, np.zeros(ns), P0, univariate_fallback=2)
    c = pl.kalman_likelihood(Y, Z, H, T, R, Q, np.zeros(ns), P0, univariate_fallback=1)
    assert abs(a - b) <= 1e-10 * abs(a) and a == c
    # two observables that are the same state plus no measurement error: F = [[1, 1], [1, 1]] exactly
    T0 = np.zeros((3, 3)); QQ = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    Y2 = np.vstack([np.arange(1, 9) / 8.0] * 2)
    Z2 = np.eye(2, 3)
    with pytest.raises(pl.DrawFailure):
        pl.kalman_likelihood(Y2, Z2, np.zeros((2, 2)), T0, np.eye(3), QQ, np.zeros(3), QQ)
    got = pl.kalman_likelihood(Y2, Z2, np.zeros((2, 2)), T0, np.eye(3), QQ, np.zeros(3), QQ, univariate_fallback=1)
    want = -0.5 * (16 * np.log(2 * np.pi) + np.sum(Y2[0] ** 2))
    assert abs(got - want) <= 1e-13 * abs(want)
