"""SMC oracle primitives: the specified summation order and the reference's resampling loop."""
import numpy as np

from oracle import smc


def test_canonical_cumsum_is_blocked_sequential():
    rng = np.random.default_rng(0)
    w = rng.random(2500)
    w /= w.sum()
    cw = smc.canonical_cumsum(w)
    # first block: plain sequential cumsum
    s = 0.0
    for i in range(1024):
        s += w[i]
        assert cw[i] == s
    assert np.all(np.diff(cw) >= 0)
    assert abs(cw[-1] - 1.0) < 1e-12
    assert smc.canonical_sum(w) == cw[-1] or abs(smc.canonical_sum(w) - cw[-1]) < 1e-15


def test_bisect_equals_reference_linear_search_and_edges():
    rng = np.random.default_rng(1)
    w = rng.random(300)
    w /= w.sum()
    u = rng.random(300)
    cw = smc.canonical_cumsum(w)
    u[0] = cw[10]              # exactly on a boundary: strict '<' moves to the next particle
    u[1] = 0.0
    u[2] = np.nextafter(cw[-1], 2.0)   # beyond the last cumulative weight: reference returns n+1 (OOB)
    idx, ncl = smc.resample_indices(w, u)
    assert np.array_equal(idx, smc.resample_indices_linear(w, u))
    assert idx[0] == 11 and idx[1] == 0 and idx[2] == 299 and ncl == 1


def test_reweight_and_moments():
    rng = np.random.default_rng(2)
    n = 3000
    w = np.full(n, 1.0 / n)
    ll = rng.normal(-800, 3, n)
    wn, zhat, ess = smc.reweight(w, ll - ll.max(), 0.01)
    assert abs(wn.sum() - 1) < 1e-12 and 1 <= ess <= n
    para = rng.normal(size=(n, 4))
    mu, R = smc.moments(para, wn)
    assert np.allclose(mu, para.T @ wn)
    z = para - mu
    assert np.allclose(R, (z * wn[:, None]).T @ z)
