"""Model-free entry points: generic batched Kalman filter and the SMC weight primitives."""
import numpy as np
import pytest

from oracle import pipeline as pl
from oracle import smc as osmc

pytestmark = pytest.mark.gpu


def _random_state_space(rng, n, ns, ny, nx):
    T = np.empty((n, ns, ns)); RQR = np.empty((n, ns, ns)); P0 = np.empty((n, ns, ns)); H = np.empty((n, ny, ny))
    for i in range(n):
        Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        lam = rng.uniform(0.3, 0.97, ns)
        T[i] = Qo @ np.diag(lam) @ Qo.T
        R = 0.1 * rng.standard_normal((ns, nx))
        RQR[i] = R @ R.T
        import scipy.linalg as sla
        P0[i] = sla.solve_discrete_lyapunov(T[i], RQR[i])
        H[i] = 0.01 * np.eye(ny)
    return T, RQR, P0, H


@pytest.mark.parametrize("ns,ny,nobs,n,variant", [
    (6, 2, 50, 40, 0), (6, 2, 50, 40, 1), (3, 1, 20, 11, 0), (40, 7, 60, 12, 1), (40, 7, 60, 12, 2), (24, 5, 30, 9, 2), (37, 9, 25, 7, 2),
    (64, 12, 20, 5, 2), (88, 20, 10, 3, 2), (110, 20, 12, 3, 0), (110, 20, 12, 3, 1), (97, 3, 9, 4, 0),
    (200, 20, 6, 2, 0), (200, 20, 6, 2, 1), (256, 7, 4, 2, 0),
    # large kernel, bulk-staged path: ny a multiple of 8 (no spare column in the last tile), ny > 32 (shared-memory
    # Cholesky and substitution), ny = 1, 13 warps with an odd strip count; variant 3 = register-staged panels
    (104, 24, 8, 3, 0), (104, 36, 6, 2, 0), (96, 1, 12, 3, 0), (136, 16, 7, 2, 0), (200, 20, 6, 2, 3), (104, 36, 6, 2, 3),
    # four / eight lanes per draw (variant 4; the automatic choice up to ns = 16): every register-array size, ny up to 8,
    # batch sizes that leave idle draw slots in the last CTA
    (10, 5, 79, 37, 0), (10, 5, 79, 37, 2), (12, 8, 30, 16, 4), (9, 8, 25, 5, 4), (16, 8, 20, 33, 4), (13, 1, 15, 4, 4),
    (1, 1, 12, 3, 0), (8, 3, 40, 100, 4)])
def test_generic_kalman_matches_oracle(ns, ny, nobs, n, variant):
    """variant 1 = scalar kernel, 2 = tensor-core (DMMA) kernel, 4 = four lanes per draw, 0 = automatic choice (four lanes
    per draw for ns <= 12, eight for ns <= 16, shared-memory DMMA kernel up to ns = 88, L2-scratch DMMA kernel for 88 < ns <= 256); sizes of BASELINE configs 4 (ns 40, ny 7) and 5 (ns 200,
    ny 20) included."""
    from dynare_jl_b200.engine import kalman_batch, set_kalman_variant
    set_kalman_variant(variant)
    rng = np.random.default_rng(ns)
    T, RQR, P0, H = _random_state_space(rng, n, ns, ny, ny)
    z = np.arange(ny)[::-1].copy() if ns < 20 else np.arange(ny)
    Y = rng.standard_normal((ny, nobs))
    Y[0, 3] = np.nan
    if nobs > 8:
        Y[:, 5] = np.nan                       # a period without any observation
    ll, st = kalman_batch(Y, z, T, RQR, P0, H=H)
    set_kalman_variant(0)
    Z = np.zeros((ny, ns)); Z[np.arange(ny), z] = 1
    for i in range(n):
        ref = pl.kalman_likelihood(Y, Z, H[i], T[i], np.eye(ns), RQR[i], np.zeros(ns), P0[i])
        assert st[i] == 0
        assert abs(ll[i] - ref) <= 1e-8 * abs(ref), (i, ll[i], ref)


def test_large_kalman_flags_non_pd_draw_and_keeps_the_others():
    from dynare_jl_b200.engine import kalman_batch
    rng = np.random.default_rng(5)
    ns, ny, n = 96, 4, 3
    T, RQR, P0, H = _random_state_space(rng, n, ns, ny, ny)
    P0[1] = -P0[1]
    Y = rng.standard_normal((ny, 6))
    ll, st = kalman_batch(Y, np.arange(ny), T, RQR, P0, H=H)
    assert st[1] == pl.ST_F_NOT_PD and ll[1] == -np.inf
    Z = np.zeros((ny, ns)); Z[np.arange(ny), np.arange(ny)] = 1
    for i in (0, 2):
        ref = pl.kalman_likelihood(Y, Z, H[i], T[i], np.eye(ns), RQR[i], np.zeros(ns), P0[i])
        assert st[i] == 0 and abs(ll[i] - ref) <= 1e-8 * abs(ref)


def test_generic_kalman_flags_non_pd():
    from dynare_jl_b200.engine import kalman_batch
    ns, ny = 4, 2
    T = 0.5 * np.eye(ns)[None]; RQR = np.zeros((1, ns, ns)); P0 = -np.eye(ns)[None]
    ll, st = kalman_batch(np.ones((ny, 5)), [0, 1], T, RQR, P0)
    assert st[0] == pl.ST_F_NOT_PD and ll[0] == -np.inf


@pytest.mark.parametrize("n", [1, 777, 1024, 16384, 100000])
def test_resampling_indices_bit_exact(n):
    from dynare_jl_b200.engine import smc_resample_multinomial, smc_resample_systematic
    rng = np.random.default_rng(n)
    w = rng.random(n) ** 3
    w /= w.sum()
    u = rng.random(n)
    if n > 10:
        cw = osmc.canonical_cumsum(w)
        u[0] = cw[5]; u[1] = 0.0; u[2] = np.nextafter(cw[-1], 2.0); u[3] = cw[-1]
    idx, ncl = smc_resample_multinomial(w, u)
    ref, rcl = osmc.resample_indices(w, u)
    assert np.array_equal(idx, ref) and ncl == rcl
    idx, ncl = smc_resample_systematic(w, 0.37)
    ref, rcl = osmc.resample_indices(w, osmc.systematic_uniforms(n, 0.37))
    assert np.array_equal(idx, ref) and ncl == rcl


def test_covariance_factor_projects_on_the_psd_cone():
    """smc.jl:161-165: prox!(RPSD, IndPSD(), R) + cholesky, against the oracle's eigen-decomposition (oracle/sampler.py)."""
    from dynare_jl_b200.engine import smc_covariance_factor
    from oracle.sampler import project_psd
    rng = np.random.default_rng(4)
    for np_ in (1, 5, 13, 36):
        A = rng.normal(size=(np_, np_ + 3))
        R = A @ A.T / (np_ + 3)
        Rp, Lc, flag = smc_covariance_factor(R)                      # positive definite: untouched, plain Cholesky
        assert flag == 0 and np.array_equal(Rp, R)
        assert np.allclose(Lc, np.linalg.cholesky(R), rtol=1e-12, atol=1e-14)
        if np_ == 1:
            continue
        # indefinite: two eigenvalues pushed below zero -> set to zero; the projection is singular, so its Cholesky factorisation
        # fails as the reference's would (flag 2), but RPSD itself must be the oracle's
        lam, V = np.linalg.eigh(R)
        lam2 = lam.copy(); lam2[0] = -0.3 * lam[-1]; lam2[1] = -1e-3 * lam[-1]
        Rind = (V * lam2) @ V.T
        Rind = 0.5 * (Rind + Rind.T)
        Rp, Lc, flag = smc_covariance_factor(Rind)
        ref = project_psd(Rind)
        assert flag in (1, 2)           # 1 if rounding left every pivot of the singular projection positive
        assert np.allclose(Rp, ref, rtol=0, atol=1e-12 * lam[-1])
        assert np.linalg.eigvalsh(Rp).min() >= -1e-12 * lam[-1]
        # projecting the projection changes nothing
        Rp2, _, _ = smc_covariance_factor(Rp)
        assert np.allclose(Rp2, Rp, rtol=0, atol=1e-12 * lam[-1])


def test_reweight_ess_and_moments():
    from dynare_jl_b200.engine import smc_reweight, smc_moments
    rng = np.random.default_rng(9)
    n = 16384
    w = rng.random(n); w /= w.sum()
    ll = rng.normal(-820, 4, n)
    wn, zhat, ess = smc_reweight(w, ll, 0.0123)
    rw, rz, re = osmc.reweight(w, ll, 0.0123)
    assert abs(zhat - rz) <= 1e-13 * abs(rz) and abs(ess - re) <= 1e-12 * re
    assert np.allclose(wn, rw, rtol=1e-13, atol=0)
    # with identical (oracle) weights the canonical sums are bit-exact
    para = rng.normal(size=(n, 17))
    mu, R = smc_moments(para, rw)
    rmu, rR = osmc.moments(para, rw)
    assert np.array_equal(mu, rmu)
    assert np.array_equal(R, rR)


@pytest.mark.parametrize("ns,ny", [(5, 2), (8, 8), (11, 4), (12, 1), (15, 6)])
def test_dense_filter_kernels_agree_with_initial_state_presample_and_correlated_noise(ns, ny):
    """The same inputs through the scalar kernel (1), the tensor-core kernel (2) and the four-lanes-per-draw kernel (4):
    non-zero a0, a presample, a full measurement-error covariance, missing observations."""
    from dynare_jl_b200.engine import kalman_batch, set_kalman_variant
    rng = np.random.default_rng(100 * ns + ny)
    n, nobs = 21, 33
    T, RQR, P0, _ = _random_state_space(rng, n, ns, ny, ny)
    H = np.empty((n, ny, ny))
    for i in range(n):
        B = 0.2 * rng.standard_normal((ny, ny)); H[i] = B @ B.T + 0.05 * np.eye(ny)
    a0 = rng.standard_normal((n, ns))
    z = rng.permutation(ns)[:ny]
    Y = rng.standard_normal((ny, nobs)); Y[0, 1] = np.nan; Y[:, 7] = np.nan; Y[ny - 1, 20] = np.nan
    res = {}
    for variant in (1, 2, 4):
        set_kalman_variant(variant)
        res[variant] = kalman_batch(Y, z, T, RQR, P0, a0=a0, H=H, presample=3)
    set_kalman_variant(0)
    Z = np.zeros((ny, ns)); Z[np.arange(ny), z] = 1
    for i in range(n):
        ref = pl.kalman_likelihood(Y, Z, H[i], T[i], np.eye(ns), RQR[i], a0[i], P0[i], presample=3)
        for variant in (1, 2, 4):
            ll, st = res[variant]
            assert st[i] == 0 and abs(ll[i] - ref) <= 1e-8 * abs(ref), (variant, i, ll[i], ref)


def test_univariate_fallback_dense():
    """dyb_kalman_batch_fallback: mode 2 (every period one observation at a time) reproduces the multivariate filter for
    diagonal H; an exactly singular innovation covariance is status 6 without the fallback and the oracle's value with it."""
    import scipy.linalg as sla
    from dynare_jl_b200.engine import kalman_batch
    from oracle import pipeline as pl
    rng = np.random.default_rng(21)
    for ns, ny, nobs in ((3, 2, 40), (10, 5, 79), (40, 7, 30)):
        n = 12
        T = np.empty((n, ns, ns)); QQ = np.empty_like(T); P0 = np.empty_like(T)
        for i in range(n):
            Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
            T[i] = Qo @ np.diag(rng.uniform(0.3, 0.95, ns)) @ Qo.T
            R = 0.2 * rng.standard_normal((ns, ny)); QQ[i] = R @ R.T
            P0[i] = sla.solve_discrete_lyapunov(T[i], QQ[i])
        H = np.tile(np.diag(rng.uniform(0.0, 0.05, ny)), (n, 1, 1))
        Y = rng.standard_normal((ny, nobs)); Y[0, 3] = np.nan
        a, sa = kalman_batch(Y, np.arange(ny), T, QQ, P0, H=H)
        b, sb = kalman_batch(Y, np.arange(ny), T, QQ, P0, H=H, univariate_fallback=2)
        c, sc = kalman_batch(Y, np.arange(ny), T, QQ, P0, H=H, univariate_fallback=1)
        assert np.all(sa == 0) and np.all(sb == 0) and np.all(sc == 0)
        assert np.max(np.abs(a - b) / np.abs(a)) <= 1e-10 and np.array_equal(a, c)
        ref = pl.kalman_likelihood(Y, np.eye(ny, ns), H[0], T[0], np.eye(ns), QQ[0], np.zeros(ns), P0[0], univariate_fallback=2)
        assert abs(b[0] - ref) <= 1e-10 * abs(ref)
    T0 = np.zeros((2, 3, 3)); Q1 = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    QQ = np.stack([Q1, np.eye(3)])
    Y2 = np.vstack([np.arange(1, 9) / 8.0] * 2)
    ll0, st0 = kalman_batch(Y2, np.arange(2), T0, QQ, QQ)
    assert st0[0] == 6 and st0[1] == 0 and ll0[0] == -np.inf
    ll1, st1 = kalman_batch(Y2, np.arange(2), T0, QQ, QQ, univariate_fallback=1)
    want = pl.kalman_likelihood(Y2, np.eye(2, 3), np.zeros((2, 2)), T0[0], np.eye(3), Q1, np.zeros(3), Q1, univariate_fallback=1)
    assert np.all(st1 == 0) and abs(ll1[0] - want) <= 1e-13 * abs(want) and ll1[1] == ll0[1]


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_univariate_fallback_model_path(name, golden):
    """dyb_options.univariate_fallback on the model-specific path: 1 leaves every golden draw as it is (their F is positive
    definite); 2 re-evaluates every draw one observation at a time and agrees wherever H is diagonal."""
    from dynare_jl_b200.engine import BatchSSWs
    g = golden(name)
    theta = g["theta"].copy()
    if name == "hs_nk":
        theta[::2, -1] = 0.0          # corr(YGR, INT) = 0 on every other draw: a diagonal measurement-error covariance
    ws = BatchSSWs(name, g["Y"])
    base, sb = ws.loglikelihood(theta)
    ws.set_options(univariate_fallback=1)
    one, s1 = ws.loglikelihood(theta)
    assert np.array_equal(base, one) and np.array_equal(sb, s1)
    ws.set_options(univariate_fallback=2)
    two, s2 = ws.loglikelihood(theta)
    _, sm = ws.shock_covariances(theta)
    diag_h = np.array([np.count_nonzero(m - np.diag(np.diag(m))) == 0 for m in sm])
    ok = (sb == 0) & diag_h
    assert ok.sum() >= 5 and np.all(s2[ok] == 0)
    assert np.max(np.abs(two[ok] - base[ok]) / np.abs(base[ok])) <= 1e-9
    assert np.all(s2[(sb == 0) & ~diag_h] == 6) and np.array_equal(s2[sb != 0], sb[sb != 0])
    ws.close()
