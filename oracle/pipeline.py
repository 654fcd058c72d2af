"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product path.

numpy/scipy restatement of the reference's per-draw likelihood path
``loglikelihood(theta, context, observations, ssws)`` (reference
``src/estimation/estimation.jl:603-702``) and of the un-vendored packages it calls
(LinearRationalExpectations 0.5.x, PolynomialMatrixEquations 0.2.x, KalmanFilterTools 0.1.x:
``Project.toml:19,27,30`` -- their source is NOT in the reference tree, so their *published
algorithms* are restated: generalised-Schur solution with the ``1+1e-6`` criterion, cyclic
reduction, discrete Lyapunov equation, prediction-error-decomposition Kalman filter).

PARITY UNPINNED: the reference ships no golden likelihood / policy matrix for this path and
cannot be run here (no Julia).  What pins this oracle instead (tests/test_oracle_*.py):
  * A G^2 + B G + C = 0 residual, rho(G) < 1, QZ vs cyclic reduction agreement;
  * Lyapunov identity (the pattern of reference ``test/test_smc.jl:29-42``);
  * Kalman likelihood vs a brute-force evaluation of the stacked Gaussian density;
  * fs2000 eigenvalue moduli at the ``.mod`` calibration.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

QZ_CRITERIUM = 1.0 + 1e-6          # reference src/perturbations.jl (StochSimulOptions.dr qz_criterium)
STEADY_TOL = np.cbrt(np.finfo(float).eps)   # reference src/steady_state/SteadyState.jl:183

# per-draw status codes (shared with include/dynare_b200.h)
OK, ST_STEADY, ST_INDETERMINATE, ST_UNSTABLE, ST_SOLVER, ST_NONSTATIONARY, ST_F_NOT_PD = range(7)


class DrawFailure(Exception):
    def __init__(self, status, msg=""):
        super().__init__(msg)
        self.status = status


# -----------------------------------------------------------------------------------------
# a-1  parameter scatter  (reference src/estimation/estimation.jl:517-601)
# -----------------------------------------------------------------------------------------
def set_estimated_parameters(model, theta):
    params = np.array(model.params0, dtype=float)
    sigma_e = np.array(model.sigma_e0, dtype=float)
    sm0 = getattr(model, "sigma_m0", None)                  # calibrated measurement errors (shocks block)
    sigma_m = np.zeros((model.ny, model.ny)) if sm0 is None else np.array(sm0, dtype=float)
    for spec, v in zip(model.est, theta):
        kind = spec[0]
        if kind == "param":
            params[model.params.index(spec[1])] = v
        elif kind == "stderr":
            if spec[1] in model.exo:
                i = model.exo.index(spec[1]); sigma_e[i, i] = v * v
            else:
                i = model.varobs.index(spec[1]); sigma_m[i, i] = v * v
        elif kind == "var":
            if spec[1] in model.exo:
                i = model.exo.index(spec[1]); sigma_e[i, i] = v
            else:
                i = model.varobs.index(spec[1]); sigma_m[i, i] = v
        elif kind == "corr":
            if spec[1] in model.exo:
                i, j = model.exo.index(spec[1]), model.exo.index(spec[2]); S = sigma_e
            else:
                i, j = model.varobs.index(spec[1]), model.varobs.index(spec[2]); S = sigma_m
            S[i, j] = S[j, i] = v * np.sqrt(S[i, i] * S[j, j])
    return params, sigma_e, sigma_m


# -----------------------------------------------------------------------------------------
# a-5  first-order solution
# -----------------------------------------------------------------------------------------
def solve_qz(A, B, C, criterium=QZ_CRITERIUM):
    """Stable solution G of A G^2 + B G + C = 0 via the ordered generalised Schur form of the
    companion pencil  [I 0; 0 A] x(t+1) = [0 I; -C -B] x(t),  x(t) = [y(t-1); y(t)].
    Blanchard-Kahn: the number of generalised eigenvalues with modulus < criterium must be n.
    Returns (G, eigenvalue moduli sorted ascending)."""
    n = A.shape[0]
    E = np.block([[np.eye(n), np.zeros((n, n))], [np.zeros((n, n)), A]])
    F = np.block([[np.zeros((n, n)), np.eye(n)], [-C, -B]])
    if not (np.all(np.isfinite(E)) and np.all(np.isfinite(F))):
        raise DrawFailure(ST_STEADY, "non-finite Jacobian")
    sel = lambda a, b: np.abs(a) < criterium * np.abs(b)
    try:
        S, T, alpha, beta, Q, Z = sla.ordqz(F, E, sort=sel, output="real")
    except Exception as exc:          # LAPACKException analogue
        raise DrawFailure(ST_SOLVER, str(exc))
    with np.errstate(divide="ignore", invalid="ignore"):
        mod = np.where(np.abs(beta) > 0, np.abs(alpha) / np.abs(beta), np.inf)
    nstable = int(np.sum(np.abs(alpha) < criterium * np.abs(beta)))
    if nstable > n:
        raise DrawFailure(ST_INDETERMINATE, f"{nstable} stable roots for {n} predetermined")
    if nstable < n:
        raise DrawFailure(ST_UNSTABLE, f"{nstable} stable roots for {n} predetermined")
    Z11, Z21 = Z[:n, :n], Z[n:, :n]
    try:
        G = np.linalg.solve(Z11.T, Z21.T).T
    except np.linalg.LinAlgError as exc:
        raise DrawFailure(ST_SOLVER, str(exc))
    return G, np.sort(mod)


CR_UNIT_ROOT_EPS = 1e-6


def solve_cr(A0, A1, A2, tol=1e-13, maxit=100):
    """Cyclic reduction for A0 + A1 X + A2 X^2 = 0 (Bini-Latouche-Meini; the algorithm of
    Dynare's ``cycle_reduction`` which PolynomialMatrixEquations mirrors).
    Returns (X, Xdual, iterations) where Xdual solves A2 + A1 H + A0 H^2 = 0."""
    m = A0.shape[0]
    A0_0, A2_0 = A0.copy(), A2.copy()
    A0, A1, A2 = A0.copy(), A1.copy(), A2.copy()
    A1hat, A1chk = A1.copy(), A1.copy()
    n0 = n2 = np.inf
    stalled = False
    for it in range(1, maxit + 1):
        try:
            with np.errstate(all="ignore"):
                X = np.linalg.solve(A1, np.hstack([A0, A2]))
                W = np.vstack([A0, A2]) @ X
        except np.linalg.LinAlgError:
            if it == 1:
                raise DrawFailure(ST_SOLVER, "singular A1 in cyclic reduction")
            stalled = True
            break
        with np.errstate(all="ignore"):
            m0 = np.abs(W[:m, :m]).sum(0).max()
            m2 = np.abs(W[m:, m:]).sum(0).max()
        if not (m0 <= 1e150) or not (m2 <= 1e150):
            if it == 1:
                raise DrawFailure(ST_SOLVER, "NaN in cyclic reduction")
            stalled = True
            break
        A1 = A1 - W[:m, m:] - W[m:, :m]
        A1hat = A1hat - W[m:, :m]
        A1chk = A1chk - W[:m, m:]
        A0 = -W[:m, :m]
        A2 = -W[m:, m:]
        n0_prev = n0
        n0, n2 = m0, m2
        # unit roots among the stable roots: A0(j) settles on a non-zero limit while A2(j) vanishes
        if (n0 < tol and n2 < tol) or (n2 < tol * tol and abs(n0 - n0_prev) <= CR_UNIT_ROOT_EPS * n0):
            break
    else:
        # A0(j) ~ lambda_m^(2^j) not vanishing: too few stable roots; A2(j) not vanishing: too many
        raise DrawFailure(ST_UNSTABLE if n0 >= tol else ST_INDETERMINATE, "cyclic reduction did not converge")
    if stalled:
        raise DrawFailure(ST_INDETERMINATE if n0 < n2 else ST_UNSTABLE, "cyclic reduction diverged")
    return -np.linalg.solve(A1hat, A0_0), -np.linalg.solve(A1chk, A2_0), it


def first_order(model, J, algo="GS"):
    """Reference ``LRE.first_order_solver!`` call (``src/perturbations.jl:663-664``).
    J = (A, B, C, Fu) full-width blocks.  Returns dict(g1_1 n x k, g1_2 n x nx, G n x n, eig)."""
    A, B, C, Fu = J
    n = model.n
    if algo == "GS":
        G, eig = solve_qz(A, B, C)
    else:
        G, H, _ = solve_cr(C, B, A)
        rho_g = np.max(np.abs(np.linalg.eigvals(G)))
        rho_h = np.max(np.abs(np.linalg.eigvals(H)))
        if rho_g >= QZ_CRITERIUM:
            raise DrawFailure(ST_UNSTABLE)
        if rho_h * QZ_CRITERIUM > 1.0:
            raise DrawFailure(ST_INDETERMINATE)
        eig = np.sort(np.abs(np.linalg.eigvals(G)))
    try:
        g1_2 = -np.linalg.solve(A @ G + B, Fu)
    except np.linalg.LinAlgError:
        raise DrawFailure(ST_SOLVER)
    return dict(g1_1=G[:, model.i_bkwrd_b], g1_2=g1_2, G=G, eig=eig)


# -----------------------------------------------------------------------------------------
# a-6  unconditional variance  (reference ``compute_variance!`` call, estimation.jl:630)
# -----------------------------------------------------------------------------------------
def compute_variance(model, sol, sigma_e):
    ib = model.i_bkwrd_b
    Gs = sol["g1_1"][ib, :]
    Bs = sol["g1_2"][ib, :]
    if np.max(np.abs(np.linalg.eigvals(Gs))) > 1.0 - 1e-6:      # the reference's split (kalman.jl:172-180)
        raise DrawFailure(ST_NONSTATIONARY)
    Sig_s = sla.solve_discrete_lyapunov(Gs, Bs @ sigma_e @ Bs.T)
    Sig_s = 0.5 * (Sig_s + Sig_s.T)
    g1, g2 = sol["g1_1"], sol["g1_2"]
    return g1 @ Sig_s @ g1.T + g2 @ sigma_e @ g2.T


# -----------------------------------------------------------------------------------------
# a-8  state-space assembly  (reference estimation.jl:640-658)
# -----------------------------------------------------------------------------------------
def state_space(model, sol, sigma_e, sigma_m, Sigma_y):
    sid = model.state_ids
    ns, ny, nx = model.ns, model.ny, model.nx
    Z = np.zeros((ny, ns))
    for i, o in enumerate(model.obs_idx_state):
        Z[i, o] = 1.0
    T = np.zeros((ns, ns))
    T[:, model.lagged_state_ids] = sol["g1_1"][sid, :]
    R = sol["g1_2"][sid, :]
    P0 = Sigma_y[np.ix_(sid, sid)]
    return dict(Z=Z, H=sigma_m.copy(), T=T, R=R, Q=sigma_e.copy(), a0=np.zeros(ns), P0=P0)


# -----------------------------------------------------------------------------------------
# a-9  Kalman log-likelihood  (reference ``kalman_likelihood`` call, estimation.jl:671-700)
# -----------------------------------------------------------------------------------------
KALMAN_TOL = 1e-10          # KalmanFilterTools' `kalman_tol` [recalled from the package source, not vendored]


def kalman_likelihood(Y, Z, H, T, R, Q, a0, P0, start=0, last=None, presample=0, univariate_fallback=0):
    """-1/2 [ nobs*ny*log(2 pi) + sum_t ( log det F_t + v_t' F_t^-1 v_t ) ].
    Missing observations are NaN in Y: rows are dropped for that period (the ``data_pattern``
    variant, estimation.jl:665-685) and the constant counts only observed entries.

    univariate_fallback (KalmanFilterTools' behaviour behind the call at estimation.jl:671-700, [recalled]: when the
    Cholesky factorisation of F_t fails the period is processed one observation at a time, an observation whose
    conditional variance is below ``kalman_tol`` being skipped): 0 = off (a failed factorisation rejects the draw),
    1 = on, 2 = every period univariate (diagnostic).  The univariate step needs a diagonal H."""
    ny, nobs = Y.shape
    last = nobs if last is None else last
    a = a0.astype(float).copy()
    P = P0.astype(float).copy()
    QQ = R @ Q @ R.T
    lik = 0.0
    ncount = 0
    for t in range(start, last):
        obs = np.flatnonzero(~np.isnan(Y[:, t]))
        if obs.size:
            Zt = Z[obs, :]
            v = Y[obs, t] - Zt @ a
            F = Zt @ P @ Zt.T + H[np.ix_(obs, obs)]
            F = 0.5 * (F + F.T)
            cF = None
            if univariate_fallback != 2:
                try:
                    cF = np.linalg.cholesky(F)
                except np.linalg.LinAlgError:
                    if not univariate_fallback:
                        raise DrawFailure(ST_F_NOT_PD)
            if cF is None:
                Hd = H[np.ix_(obs, obs)]
                if np.any(Hd - np.diag(np.diag(Hd)) != 0):
                    raise DrawFailure(ST_F_NOT_PD)
                for q, i in enumerate(obs):
                    z = Z[i, :]
                    Fi = z @ P @ z + H[i, i]
                    if Fi > KALMAN_TOL:
                        vi = Y[i, t] - z @ a
                        K = (P @ z) / Fi
                        a = a + K * vi
                        P = P - np.outer(K, K) * Fi
                        if t >= start + presample:
                            lik += np.log(Fi) + vi * vi / Fi
                if t >= start + presample:
                    ncount += obs.size
                a = T @ a
                P = T @ P @ T.T + QQ
                P = 0.5 * (P + P.T)
                continue
            w = sla.solve_triangular(cF, v, lower=True)
            PZt = P @ Zt.T
            U = sla.solve_triangular(cF, PZt.T, lower=True)          # L^-1 Z P
            if t >= start + presample:
                lik += 2.0 * np.sum(np.log(np.diag(cF))) + w @ w
                ncount += obs.size
            a = a + U.T @ w
            P = P - U.T @ U
        a = T @ a
        P = T @ P @ T.T + QQ
        P = 0.5 * (P + P.T)
    return -0.5 * (ncount * np.log(2 * np.pi) + lik)


def gaussian_loglik_bruteforce(Y, Z, H, T, R, Q, a0, P0):
    """Independent evaluation: stack all observations, build the (nobs*ny)^2 covariance of the
    linear Gaussian state space explicitly and evaluate the multivariate normal density."""
    ny, nobs = Y.shape
    ns = T.shape[0]
    QQ = R @ Q @ R.T
    # Var(alpha_t) and Cov(alpha_t, alpha_s)
    V = [P0.copy()]
    for t in range(1, nobs):
        V.append(T @ V[-1] @ T.T + QQ)
    big = np.zeros((nobs * ny, nobs * ny))
    for s in range(nobs):
        C = V[s].copy()
        for t in range(s, nobs):
            if t > s:
                C = T @ C                    # Cov(alpha_t, alpha_s) = T^(t-s) Var(alpha_s)
            blk = Z @ C @ Z.T
            if t == s:
                blk = blk + H
            big[t * ny:(t + 1) * ny, s * ny:(s + 1) * ny] = blk
            big[s * ny:(s + 1) * ny, t * ny:(t + 1) * ny] = blk.T
    mean = np.concatenate([Z @ np.linalg.matrix_power(T, t) @ a0 for t in range(nobs)])
    y = Y.T.reshape(-1) - mean
    cF = np.linalg.cholesky(big)
    w = sla.solve_triangular(cF, y, lower=True)
    return -0.5 * (nobs * ny * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(cF))) + w @ w)


# -----------------------------------------------------------------------------------------
# the whole path
# -----------------------------------------------------------------------------------------
def solve_draw(model, theta, algo="GS"):
    """theta -> everything up to the state space.  Raises DrawFailure(status)."""
    params, sigma_e, sigma_m = set_estimated_parameters(model, theta)
    ybar = model.steady_state(params)
    if not np.all(np.isfinite(ybar)):
        raise DrawFailure(ST_STEADY, "non-finite steady state")
    res = model.static_resid(ybar, params)
    if not (np.linalg.norm(res) <= STEADY_TOL):            # SteadyState.jl:476-482
        raise DrawFailure(ST_STEADY, "steady state residuals")
    J = model.abcf(ybar, params)
    if not all(np.all(np.isfinite(x)) for x in J):
        raise DrawFailure(ST_STEADY, "non-finite Jacobian")
    sol = first_order(model, J, algo)
    Sigma_y = compute_variance(model, sol, sigma_e)
    ss = state_space(model, sol, sigma_e, sigma_m, Sigma_y)
    return dict(params=params, sigma_e=sigma_e, sigma_m=sigma_m, ybar=ybar, sol=sol,
                Sigma_y=Sigma_y, ss=ss)


def loglikelihood(model, theta, Y, algo="GS"):
    """Reference ``loglikelihood`` (estimation.jl:603-702).  Y is ny x nobs (NaN = missing).
    Returns (loglik, status): loglik = -inf whenever status != 0 (estimation.jl:503-515)."""
    try:
        d = solve_draw(model, theta, algo)
        Yd = Y - d["ybar"][model.obs_idx][:, None]          # data.jl:355-363
        ss = d["ss"]
        ll = kalman_likelihood(Yd, ss["Z"], ss["H"], ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"])
        if not np.isfinite(ll):
            return -np.inf, ST_F_NOT_PD
        return ll, OK
    except DrawFailure as f:
        return -np.inf, f.status


def simulate_observations(model, theta, nobs, seed):
    """Synthetic data of SURVEY.md section 8d: alpha_0 ~ N(0, P0), then nobs steps; Y = Z alpha + ybar_obs."""
    d = solve_draw(model, theta)
    ss = d["ss"]
    rng = np.random.default_rng(seed)
    ns = model.ns
    w, V = np.linalg.eigh(ss["P0"])
    a = V @ (np.sqrt(np.clip(w, 0, None)) * rng.standard_normal(ns))
    cQ = np.linalg.cholesky(ss["Q"])
    Y = np.zeros((model.ny, nobs))
    H = ss["H"]
    cH = np.linalg.cholesky(H + 1e-300 * np.eye(model.ny)) if np.any(H) else None
    for t in range(nobs):
        Y[:, t] = ss["Z"] @ a + d["ybar"][model.obs_idx]
        if cH is not None:                                   # measurement errors (models that declare them)
            Y[:, t] += cH @ rng.standard_normal(model.ny)
        a = ss["T"] @ a + ss["R"] @ (cQ @ rng.standard_normal(model.nx))
    return Y


# -----------------------------------------------------------------------------------------
# Kalman smoother (reference `kalman_smoother!` call, src/filters/kalman/kalman.jl:52-160; KalmanFilterTools)
# -----------------------------------------------------------------------------------------
def kalman_smoother(Y, Z, H, T, R, Q, a0, P0):
    """Fixed-interval state / disturbance smoother (Durbin & Koopman 2012, sections 4.3-4.5) with rows of missing
    observations (NaN) dropped period by period.  Returns dict(a_pred (ns, nobs+1), att (ns, nobs), alphah (ns, nobs),
    etah (nx, nobs), epsilonh (ny, nobs), loglik)."""
    ny, nobs = Y.shape
    ns = T.shape[0]
    nx = R.shape[1]
    a = np.array(a0, dtype=float); P = np.array(P0, dtype=float)
    QQ = R @ Q @ R.T
    A = np.zeros((ns, nobs + 1)); att = np.zeros((ns, nobs))
    store = []
    ll = 0.0
    for t in range(nobs):
        A[:, t] = a
        obs = ~np.isnan(Y[:, t])
        Zt = Z[obs]
        if obs.any():
            v = Y[obs, t] - Zt @ a
            F = Zt @ P @ Zt.T + H[np.ix_(obs, obs)]
            Fi = np.linalg.inv(F)
            G = P @ Zt.T @ Fi
            K = T @ G
            ll += -0.5 * (obs.sum() * np.log(2 * np.pi) + np.linalg.slogdet(F)[1] + v @ Fi @ v)
            att[:, t] = a + G @ v
            L = T - K @ Zt
            store.append((obs, Zt, Fi @ v, K, L, P.copy()))
            a = T @ a + K @ v
            P = T @ P @ L.T + QQ
        else:
            att[:, t] = a
            store.append((obs, Zt, np.zeros(0), np.zeros((ns, 0)), T.copy(), P.copy()))
            a = T @ a
            P = T @ P @ T.T + QQ
        P = 0.5 * (P + P.T)
    A[:, nobs] = a
    r = np.zeros(ns)
    alphah = np.zeros((ns, nobs)); etah = np.zeros((nx, nobs)); eps = np.zeros((ny, nobs))
    for t in range(nobs - 1, -1, -1):
        obs, Zt, u, K, L, Pt = store[t]
        etah[:, t] = Q @ R.T @ r
        e = u - K.T @ r
        full = np.zeros(ny); full[obs] = e
        eps[:, t] = H @ full
        r = Zt.T @ u + L.T @ r
        alphah[:, t] = A[:, t] + Pt @ r
    return dict(a_pred=A, att=att, alphah=alphah, etah=etah, epsilonh=eps, loglik=ll)


def gaussian_smoother_bruteforce(Y, Z, H, T, R, Q, a0, P0):
    """Independent second opinion for the smoothed states: E[alpha_t | all observations] from the explicit joint
    Gaussian law of (alpha_0..alpha_{n-1}, y_0..y_{n-1}) (no missing observations)."""
    ny, nobs = Y.shape
    ns = T.shape[0]
    QQ = R @ Q @ R.T
    V = [np.array(P0, dtype=float)]
    for t in range(1, nobs):
        V.append(T @ V[-1] @ T.T + QQ)
    Saa = np.zeros((nobs * ns, nobs * ns))
    for s in range(nobs):
        C = V[s].copy()
        for t in range(s, nobs):
            if t > s:
                C = T @ C
            Saa[t * ns:(t + 1) * ns, s * ns:(s + 1) * ns] = C
            Saa[s * ns:(s + 1) * ns, t * ns:(t + 1) * ns] = C.T
    Zb = np.kron(np.eye(nobs), Z)
    Syy = Zb @ Saa @ Zb.T + np.kron(np.eye(nobs), H)
    mean_a = np.concatenate([np.linalg.matrix_power(T, t) @ a0 for t in range(nobs)])
    y = Y.T.reshape(-1) - Zb @ mean_a
    ah = mean_a + Saa @ Zb.T @ np.linalg.solve(Syy, y)
    return ah.reshape(nobs, ns).T
