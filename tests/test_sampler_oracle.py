"""The samplers' CPU oracle (oracle/sampler.py) pinned by identities: priors against the host prior code,
transformations round trip, and the whole SMC recursion on a conjugate Gaussian problem whose posterior and
marginal data density are known in closed form."""
import math

import numpy as np

from oracle import rng as orng
from oracle import sampler as osam
from dynare_jl_b200.priors import make_prior, PriorSet

ALL = [("beta", 0.4, 0.1), ("gamma", 2.0, 0.5), ("normal", 0.3, 0.2), ("inv_gamma", 0.05, float("inf")),
       ("inv_gamma", 1.2, 0.6), ("uniform", 0.5, 0.25), ("inv_gamma2", 0.8, 0.3), ("weibull", 1.5, 0.7)]


def _abi(ps):
    code, a, b = ps.abi_arrays()
    return list(zip(code.tolist(), a.tolist(), b.tolist()))


def test_prior_oracle_matches_host_priors():
    ps = PriorSet([make_prior(*p) for p in ALL])
    x = ps.sample(np.random.default_rng(0), 50)
    ref = ps.logpriordensity(x)
    got = np.array([osam.logprior(_abi(ps), xi) for xi in x])
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    # outside the support
    bad = x[0].copy(); bad[0] = 1.5
    assert osam.logprior(_abi(ps), bad) == -math.inf


def test_transform_round_trip_and_jacobian():
    ps = PriorSet([make_prior(*p) for p in ALL])
    pri = _abi(ps)
    y = np.random.default_rng(1).normal(size=len(ALL))
    th, lj = osam.transform(pri, y)
    assert np.allclose(osam.inverse_transform(pri, th), y, atol=1e-12)
    # log-Jacobian by central differences
    h = 1e-6
    num = 0.0
    for k in range(len(ALL)):
        e = np.zeros(len(ALL)); e[k] = h
        num += math.log((osam.transform(pri, y + e)[0][k] - osam.transform(pri, y - e)[0][k]) / (2 * h))
    assert abs(num - lj) < 1e-7


def test_smc_oracle_recovers_conjugate_gaussian_posterior():
    """theta ~ N(m0, s0^2 I), data_k ~ N(theta, s^2 I): posterior and log marginal density are closed-form."""
    npar, N, nphi = 2, 2048, 40
    m0, s0, s = np.array([0.3, -0.2]), 0.8, 0.5
    data = np.array([[0.9, 0.1], [1.1, -0.3], [0.7, 0.2]])
    pri = [(3, m0[k], s0) for k in range(npar)]

    def loglik(th):
        ll = np.zeros(th.shape[0])
        for d in data:
            ll += -0.5 * np.sum((d - th) ** 2, axis=1) / s ** 2 - npar * math.log(s) - 0.5 * npar * math.log(2 * math.pi)
        return ll, np.zeros(th.shape[0], dtype=np.int32)

    g = np.random.default_rng(3)
    para = m0 + s0 * g.standard_normal((N, npar))
    ll, _ = loglik(para)
    phi = (np.arange(nphi) / (nphi - 1)) ** 2.0
    state = dict(para=para, loglh=ll, logpost=np.array([phi[0] * l + osam.logprior(pri, p) for l, p in zip(ll, para)]),
                 w=np.full(N, 1.0 / N))
    tune = dict(c=0.5, acpt=0.25, trgt=0.25, alp=0.9)
    logmdd = 0.0
    idx = np.arange(N)
    nres = 0
    for i in range(1, nphi):
        rand = dict(u_res=orng.uniform(11, idx, i, orng.STREAM_RESAMPLE), u_mix=orng.uniform(11, idx, i, orng.STREAM_MIX),
                    u_acc=orng.uniform(11, idx, i, orng.STREAM_ACCEPT), z=orng.normals(11, idx, i, npar))
        state, tune, info = osam.smc_stage(state, i, phi, tune, rand, loglik, lambda t: osam.logprior(pri, t))
        logmdd += math.log(info["zhat"])
        nres += info["resampled"]
    n = data.shape[0]
    prec = 1 / s0 ** 2 + n / s ** 2
    post_mean = (m0 / s0 ** 2 + data.sum(0) / s ** 2) / prec
    mean = state["w"] @ state["para"]
    var = state["w"] @ (state["para"] - mean) ** 2
    assert np.all(np.abs(mean - post_mean) < 0.04)
    assert np.all(np.abs(var - 1 / prec) < 0.03)
    # log marginal density of the data: product over coordinates of N(data_k; m0 1, s^2 I + s0^2 11')
    from scipy.stats import multivariate_normal as mvn
    exact = sum(mvn.logpdf(data[:, k], mean=np.full(n, m0[k]), cov=s ** 2 * np.eye(n) + s0 ** 2 * np.ones((n, n)))
                for k in range(npar))
    assert abs(logmdd - exact) < 0.1
    assert nres >= 1 and 0.05 < tune["acpt"] < 0.9
    # posterior kernel bookkeeping: logpost = phi_end * loglh + logprior for every particle
    llf, _ = loglik(state["para"])
    ref = np.array([phi[-1] * l + osam.logprior(pri, p) for l, p in zip(llf, state["para"])])
    assert np.allclose(state["logpost"], ref, atol=1e-9)


def test_rwmh_oracle_targets_the_right_law():
    pri = [(3, 0.0, 1.0)]
    target = lambda y: osam.logprior(pri, y)
    n = 20000
    z = orng.normals(5, np.arange(n), 0, 1)
    u = orng.uniform(5, np.arange(n), 0, orng.STREAM_ACCEPT)
    hist, hlp, acc = osam.rwmh_chain(np.array([0.1]), np.array([[2.4]]), n, z, u, target)
    assert abs(hist[2000:].mean()) < 0.1 and abs(hist[2000:].var() - 1.0) < 0.15
    assert 0.2 < acc / n < 0.7


def test_prior_draws_follow_their_distributions():
    """The counter-based prior sampler restated in oracle/sampler.py (Marsaglia-Tsang Gammas etc., the algorithm the device
    runs) against scipy.stats: Kolmogorov-Smirnov on 40 000 draws per shape, hyper-parameters as the reference's
    distributions take them (priors.jl:482-529)."""
    import scipy.stats as st
    from oracle.sampler import prior_draws
    codes = [1, 1, 2, 2, 3, 4, 5, 6, 8]
    a = [2.5, 0.6, 3.0, 0.4, 0.3, 3.2, -1.0, 4.0, 1.7]
    b = [4.0, 0.8, 0.5, 2.0, 1.5, 0.9, 2.5, 2.0, 0.8]
    n = 40000
    x = prior_draws(codes, a, b, seed=1234, first=10, n=n)
    cdfs = [st.beta(2.5, 4.0).cdf, st.beta(0.6, 0.8).cdf, st.gamma(3.0, scale=0.5).cdf, st.gamma(0.4, scale=2.0).cdf,
            st.norm(0.3, 1.5).cdf, lambda v: st.invgamma(3.2, scale=0.9).cdf(v * v), st.uniform(-1.0, 3.5).cdf,
            st.invgamma(4.0, scale=2.0).cdf, st.weibull_min(1.7, scale=0.8).cdf]
    for k, cdf in enumerate(cdfs):
        d, pval = st.kstest(x[:, k], cdf)
        assert pval > 1e-3, (k, codes[k], d, pval)
    # a draw depends on (seed, particle, round, parameter) only
    y = prior_draws(codes, a, b, seed=1234, first=10 + 100, n=50)
    assert np.array_equal(y, x[100:150])
    assert not np.array_equal(prior_draws(codes, a, b, seed=1234, first=10, n=50, rnd=1), x[:50])
    assert not np.array_equal(prior_draws(codes, a, b, seed=1235, first=10, n=50), x[:50])
