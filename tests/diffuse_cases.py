"""State spaces with unit roots shared by the oracle and GPU tests of the diffuse smoother (Z is a selection, so the
observed series are states: y = trend loadings + cycles, written as their own rows of T and R)."""
import numpy as np


def trend_cycle_case(seed, n_trend=2, n_cycle=3, ny=3, nobs=30, i2=False, missing=True, meas_err=True):
    """n_trend random walks (or one local linear trend when i2), a stable VAR(1) in n_cycle cycles, ny observed sums."""
    rng = np.random.default_rng(seed)
    nb = n_trend + n_cycle
    B = np.zeros((nb, nb))
    B[:n_trend, :n_trend] = np.eye(n_trend)
    if i2:
        B[0, 1] = 1.0                                    # level_t = level_{t-1} + slope_{t-1} + e ; slope a random walk
    A = rng.uniform(-0.3, 0.3, (n_cycle, n_cycle)) + np.diag(rng.uniform(0.3, 0.8, n_cycle))
    A *= 0.85 / max(0.85, np.abs(np.linalg.eigvals(A)).max())
    B[n_trend:, n_trend:] = A
    Lo = np.zeros((ny, nb))
    for i in range(ny):
        Lo[i, 0 if i2 else i % n_trend] = 1.0
        if not i2 and i == ny - 1 and n_trend > 1:
            Lo[i, 0] = 0.5
        Lo[i, n_trend + i % n_cycle] = 1.0
    ns, nx = nb + ny, nb
    T = np.zeros((ns, ns)); T[:nb, :nb] = B; T[nb:, :nb] = Lo @ B
    R = np.zeros((ns, nx)); R[:nb] = np.eye(nb); R[nb:] = Lo
    Qh = 0.2 * rng.standard_normal((nx, nx)); Q = Qh @ Qh.T + np.diag(rng.uniform(0.1, 0.6, nx))
    z = np.arange(nb, ns)
    Hd = np.zeros(ny)
    if meas_err:
        Hd[ny // 2] = 0.07
    a = np.zeros(ns); a[:n_trend] = 4.0 * rng.standard_normal(n_trend)
    Lq = np.linalg.cholesky(Q)
    Y = np.zeros((ny, nobs))
    for t in range(nobs):
        a = T @ a + R @ (Lq @ rng.standard_normal(nx))
        Y[:, t] = a[z] + np.sqrt(Hd) * rng.standard_normal(ny)
    if missing:
        Y[0, 0] = np.nan
        if ny > 1:
            Y[1, 1] = np.nan
        Y[:, 4] = np.nan
        Y[ny - 1, nobs // 2] = np.nan
    return dict(Y=Y, z=z, Hd=Hd, T=T, R=R, Q=Q)
