"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the diffuse (unit-root) branch of the reference's smoother.

What the reference does (src/filters/kalman/kalman.jl:163-223): when some endogenous variables are non-stationary it
takes the ordered real Schur form of T (eigenvalues with modulus > 1 - 1e-6 first), solves the Lyapunov equation on the
stable block only, puts an identity on the unit-root block of Pinf and calls KalmanFilterTools'
`diffuse_kalman_smoother!` (0.1.x, not vendored) with tolerance 1e-8 in the rotated basis, rotating the results back.
The filter is equivariant under that orthogonal rotation, so this restatement stays in the ORIGINAL basis (Z remains a
selection): Pinf0 = Z1 Z1', Pstar0 = Z2 P22 Z2' with [Z1 Z2] the Schur vectors.

Numerics: exact diffuse initialisation in its univariate form (Koopman & Durbin 2000; Durbin & Koopman 2012,
sections 5.2-5.4 and 6.4), which covers rank-deficient Finf -- the usual DSGE case of several observables loading on
one stochastic trend -- and needs a diagonal H.  Parity unpinned by the reference (no stored outputs, no Julia); the
second opinion is `diffuse_bruteforce`, the closed-form limit of the Gaussian law as the diffuse variance grows.
"""
import numpy as np
import scipy.linalg as sla

DIFFUSE_TOL = 1e-8          # kalman.jl:211 (the tolerance argument of diffuse_kalman_smoother!)
UNIT_ROOT_CRITERIUM = 1 - 1e-6   # kalman.jl:172, 180


def diffuse_initial_covariances(T, R, Q, criterium=UNIT_ROOT_CRITERIUM):
    """kalman.jl:163-195 restated in the original basis.  Returns (Pinf0, Pstar0, k) with k unit/explosive roots."""
    ns = T.shape[0]
    S, Zs, k = sla.schur(T, output="real", sort=lambda re, im: np.hypot(re, im) > criterium)
    tR = Zs.T @ R
    P22 = np.zeros((ns - k, ns - k))
    if ns > k:
        P22 = sla.solve_discrete_lyapunov(S[k:, k:], tR[k:] @ Q @ tR[k:].T)
    Z1, Z2 = Zs[:, :k], Zs[:, k:]
    Pinf0 = Z1 @ Z1.T
    Pstar0 = Z2 @ P22 @ Z2.T
    return 0.5 * (Pinf0 + Pinf0.T), 0.5 * (Pstar0 + Pstar0.T), int(k)


def diffuse_kalman_smoother(Y, z_idx, Hdiag, T, R, Q, a0, Pinf0, Pstar0, tol=DIFFUSE_TOL):
    """Univariate exact-diffuse filter and smoother.  Y (ny, nobs) with NaN = missing; z_idx[i] = state observed by
    row i; Hdiag (ny,) measurement-error variances.  A period is treated as diffuse while max diag(Pinf) > tol.
    Returns dict(loglik, d (number of diffuse periods), a_pred (ns, nobs+1), att, alphah (ns, nobs), etah (nx, nobs),
    epsilonh (ny, nobs))."""
    ny, nobs = Y.shape
    ns = T.shape[0]
    nx = R.shape[1]
    QQ = R @ Q @ R.T
    a = np.array(a0, dtype=float)
    Ps = np.array(Pstar0, dtype=float)
    Pi = np.array(Pinf0, dtype=float)
    A = np.zeros((ns, nobs + 1)); att = np.zeros((ns, nobs))
    Pstore = []; Pistore = []; steps = []
    acc = 0.0; nseen = 0; d = 0
    diffuse = Pi.diagonal().max(initial=0.0) > tol
    for t in range(nobs):
        A[:, t] = a
        Pstore.append(Ps.copy()); Pistore.append(Pi.copy() if diffuse else None)
        if diffuse:
            d = t + 1
        rec = []
        for i in range(ny):
            if np.isnan(Y[i, t]):
                continue
            z = z_idx[i]
            v = Y[i, t] - a[z]
            Ks = Ps[:, z].copy(); Fs = Ps[z, z] + Hdiag[i]
            if diffuse:
                Ki = Pi[:, z].copy(); Fi = Pi[z, z]
            else:
                Ki = None; Fi = 0.0
            if diffuse and Fi > tol:
                a = a + Ki * (v / Fi)
                Ps = Ps + np.outer(Ki, Ki) * (Fs / (Fi * Fi)) - (np.outer(Ks, Ki) + np.outer(Ki, Ks)) / Fi
                Pi = Pi - np.outer(Ki, Ki) / Fi
                acc += np.log(Fi); nseen += 1
                rec.append((i, z, 1, v, Fi, Fs, Ki, Ks))
            elif Fs > tol:
                a = a + Ks * (v / Fs)
                Ps = Ps - np.outer(Ks, Ks) / Fs
                acc += np.log(Fs) + v * v / Fs; nseen += 1
                rec.append((i, z, 0, v, Fi, Fs, Ki, Ks))
        steps.append(rec)
        att[:, t] = a
        a = T @ a
        Ps = T @ Ps @ T.T + QQ; Ps = 0.5 * (Ps + Ps.T)
        if diffuse:
            Pi = T @ Pi @ T.T; Pi = 0.5 * (Pi + Pi.T)
            diffuse = Pi.diagonal().max(initial=0.0) > tol
    A[:, nobs] = a
    loglik = -0.5 * (nseen * np.log(2 * np.pi) + acc)
    r0 = np.zeros(ns); r1 = np.zeros(ns)
    alphah = np.zeros((ns, nobs)); etah = np.zeros((nx, nobs)); eps = np.zeros((ny, nobs))
    for t in range(nobs - 1, -1, -1):
        etah[:, t] = Q @ R.T @ r0
        r0 = T.T @ r0
        r1 = T.T @ r1
        for (i, z, kind, v, Fi, Fs, Ki, Ks) in reversed(steps[t]):
            if kind == 1:
                n1 = r1.copy()
                n1[z] += v / Fi + ((Ki * (Fs / Fi) - Ks) @ r0) / Fi - (Ki @ r1) / Fi
                n0 = r0.copy()
                n0[z] -= (Ki @ r0) / Fi
                r0, r1 = n0, n1
            else:
                n0 = r0.copy()
                n0[z] += v / Fs - (Ks @ r0) / Fs
                n1 = r1.copy()
                n1[z] -= (Ks @ r1) / Fs
                r0, r1 = n0, n1
        x = A[:, t] + Pstore[t] @ r0
        if Pistore[t] is not None:
            x = x + Pistore[t] @ r1
        alphah[:, t] = x
        for i in range(ny):
            if not np.isnan(Y[i, t]) and Hdiag[i] > 0.0:
                eps[i, t] = Y[i, t] - x[z_idx[i]]
    return dict(loglik=loglik, d=d, a_pred=A, att=att, alphah=alphah, etah=etah, epsilonh=eps)


def diffuse_bruteforce(Y, z_idx, Hdiag, T, R, Q, a0, Pinf0, Pstar0):
    """Independent second opinion (no recursion): with alpha_0 = a0 + A delta + N(0, Pstar0), Pinf0 = A A', the limit
    of the Gaussian law as var(delta) -> infinity is generalised least squares in delta.  Returns (diffuse log-likelihood
    = lim [log L + (k/2) log kappa], smoothed states (ns, nobs)).  Missing observations are dropped."""
    ny, nobs = Y.shape
    ns = T.shape[0]
    QQ = R @ Q @ R.T
    w, U = np.linalg.eigh(Pinf0)
    keep = w > 1e-10
    A = U[:, keep] * np.sqrt(w[keep])
    k = A.shape[1]
    V = [np.array(Pstar0, dtype=float)]
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
    Z = np.zeros((ny, ns)); Z[np.arange(ny), z_idx] = 1.0
    Zb = np.kron(np.eye(nobs), Z)
    Tp = [np.eye(ns)]
    for t in range(1, nobs):
        Tp.append(T @ Tp[-1])
    mean_a = np.concatenate([Tp[t] @ a0 for t in range(nobs)])
    Phi = np.vstack([Tp[t] @ A for t in range(nobs)])
    yfull = Y.T.reshape(-1)
    obs = ~np.isnan(yfull)
    Zo = Zb[obs]
    Syy = Zo @ Saa @ Zo.T + np.diag(np.tile(Hdiag, nobs)[obs])
    y = yfull[obs] - Zo @ mean_a
    X = Zo @ Phi
    Si_y = np.linalg.solve(Syy, y)
    Si_X = np.linalg.solve(Syy, X)
    XtSX = X.T @ Si_X
    delta = np.linalg.solve(XtSX, X.T @ Si_y) if k else np.zeros(0)
    resid = y - X @ delta
    quad = resid @ np.linalg.solve(Syy, resid)
    ll = -0.5 * (obs.sum() * np.log(2 * np.pi) + np.linalg.slogdet(Syy)[1]
                 + (np.linalg.slogdet(XtSX)[1] if k else 0.0) + quad)
    ah = mean_a + Phi @ delta + Saa @ Zo.T @ np.linalg.solve(Syy, resid)
    return ll, ah.reshape(nobs, ns).T
