"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product path.

CPU restatement of the reference's samplers, one particle / chain at a time as the reference loops:
  prior log-densities    src/estimation/estimation.jl:455-462, src/distributions/inversegamma1.jl:134-141
                         (scipy.stats distributions as an independent second opinion)
  transformations        src/estimation/estimation.jl:465-487 (TransformVariables as𝕀, asℝ₊, asℝ, as(Real,a,b))
  SMC stage              src/estimation/smc.jl:110-243, proposal densities :326-338
  RWMH                   src/estimation/estimation.jl:964-989 (AdvancedMH RWMH, symmetric MvNormal proposal)
Random numbers are inputs (oracle/rng.py restates the engine's generator), sums over particles use the canonical
blocked order of oracle/smc.py.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import stats

from . import smc as osmc

LOG2PI = math.log(2.0 * math.pi)


# ---- priors (shape codes of src/estimation/priors.jl:482-529) -----------------------------------------
def prior_logpdf(code, a, b, x):
    if code == 1:
        return float(stats.beta.logpdf(x, a, b))
    if code == 2:
        return float(stats.gamma.logpdf(x, a, scale=b))
    if code == 3:
        return float(stats.norm.logpdf(x, a, b))
    if code == 4:                       # InverseGamma1 (inversegamma1.jl:134-141)
        if not x > 0:
            return -math.inf
        return math.log(2.0) + a * math.log(b) - math.lgamma(a) - (2 * a + 1) * math.log(x) - b / (x * x)
    if code == 5:
        return float(stats.uniform.logpdf(x, a, b - a))
    if code == 6:
        return float(stats.invgamma.logpdf(x, a, scale=b))
    if code == 8:
        return float(stats.weibull_min.logpdf(x, a, scale=b))
    raise ValueError(code)


def logprior(priors, theta):
    """priors: list of (code, a, b); left-to-right sum as estimation.jl:455-462."""
    s = 0.0
    for (code, a, b), x in zip(priors, theta):
        s += prior_logpdf(code, a, b, float(x))
    return s


TRANSFORM_OF_SHAPE = {1: 2, 2: 1, 3: 0, 4: 1, 5: 3, 6: 1, 8: 1}


def _logistic_logjac(y):
    return -(np.logaddexp(0.0, -y) + np.logaddexp(0.0, y))


def transform(priors, y):
    """R^np -> parameters; returns (theta, log-Jacobian)."""
    th = np.empty(len(priors)); lj = 0.0
    for k, ((code, a, b), v) in enumerate(zip(priors, y)):
        t = TRANSFORM_OF_SHAPE[code]
        if t == 0:
            th[k] = v
        elif t == 1:
            th[k] = math.exp(v); lj += v
        elif t == 2:
            th[k] = 1.0 / (1.0 + math.exp(-v)); lj += _logistic_logjac(v)
        else:
            th[k] = (b - a) / (1.0 + math.exp(-v)) + a; lj += math.log(b - a) + _logistic_logjac(v)
    return th, float(lj)


def inverse_transform(priors, theta):
    y = np.empty(len(priors))
    for k, ((code, a, b), x) in enumerate(zip(priors, theta)):
        t = TRANSFORM_OF_SHAPE[code]
        if t == 0:
            y[k] = x
        elif t == 1:
            y[k] = math.log(x)
        elif t == 2:
            y[k] = math.log(x / (1.0 - x))
        else:
            u = (x - a) / (b - a); y[k] = math.log(u / (1.0 - u))
    return y


# ---- SMC ---------------------------------------------------------------------------------------------------
def _mvn_logpdf(x, mean, cov):
    """logpdf(MvNormal(mean, cov), x) = -(d log 2pi + logdet cov)/2 - sqmahal/2 (Distributions' mvnormal_c0 form).
    Written out (scipy's multivariate_normal refuses badly scaled but positive definite covariances)."""
    d = np.asarray(x, dtype=np.float64) - mean
    sign, logdet = np.linalg.slogdet(cov)
    if sign <= 0:
        raise np.linalg.LinAlgError("covariance not positive definite")
    return float(-0.5 * (d.size * LOG2PI + logdet) - 0.5 * d @ np.linalg.solve(cov, d))


def logpdf_mixture_mh(px, p0, c, mu, RPSD, Rdiag, mixsel):
    """smc.jl:326-338.  For mixsel == 2 the reference passes the VECTOR c^2 * Rdiag to MvNormal, which
    Distributions (0.25) reads as standard deviations."""
    if mixsel == 1:
        return _mvn_logpdf(px, p0, c * c * RPSD)
    if mixsel == 2:
        sd = c * c * Rdiag
        return _mvn_logpdf(px, p0, np.diag(sd * sd))
    return _mvn_logpdf(px, mu, c * c * RPSD)


def project_psd(R):
    """prox!(RPSD, IndPSD(), R): eigen-decomposition of the symmetric matrix with negative eigenvalues set to 0."""
    Rs = np.tril(R) + np.tril(R, -1).T
    lam, V = np.linalg.eigh(Rs)
    return (V * np.maximum(lam, 0.0)) @ V.T


def smc_stage(state, i, phi, tune, rand, loglik, logprior_fn, systematic=False):
    """One tempering stage i (0-based, i >= 1) on the whole particle system.
    state: dict(para (N, np), loglh, logpost, w);  tune: dict(c, acpt, trgt, alp);
    rand: dict(u_res (N,) or scalar u0 if systematic, u_mix (N,), u_acc (N,), z (N, np));
    loglik(theta (N, np)) -> (ll, status); logprior_fn(theta) -> float.  Returns (state, tune, info)."""
    para, loglh, logpost, w = (np.array(state[k], dtype=np.float64) for k in ("para", "loglh", "logpost", "w"))
    N, npar = para.shape
    dphi = phi[i] - phi[i - 1]
    # correction (smc.jl:118-123), ESS (:129)
    w, zhat, ess = osmc.reweight(w, loglh, dphi)
    resampled = ess < N / 2
    nclamp = 0
    if resampled:
        u = osmc.systematic_uniforms(N, rand["u_res"]) if systematic else rand["u_res"]
        idx, nclamp = osmc.resample_indices(w, u)
        para, loglh, logpost = para[idx], loglh[idx], logpost[idx]
        w = np.full(N, 1.0 / N)
    # scale update (smc.jl:150-151)
    e = math.exp(16.0 * (tune["acpt"] - tune["trgt"]))
    c = tune["c"] * (0.95 + 0.10 * e / (1.0 + e))
    # moments (smc.jl:154-165)
    mu, R = osmc.moments(para, w)
    RPSD = project_psd(R)
    Rdiag = np.diag(RPSD).copy()
    Rchol = np.linalg.cholesky(RPSD)
    Rchol2 = np.sqrt(Rdiag)
    # mutation (smc.jl:172-233)
    prop = np.empty_like(para); qx = np.empty(N); q0 = np.empty(N); sel = np.empty(N, dtype=int)
    alp = tune["alp"]
    for j in range(N):
        p0 = para[j]
        z = rand["z"][j]
        m = rand["u_mix"][j]
        if m < alp:
            px = p0 + (c * Rchol) @ z; s = 1
        elif m < alp + (1 - alp) / 2:
            px = p0 + (c * Rchol2) * z; s = 2
        else:
            px = mu + (c * Rchol) @ z; s = 3
        prop[j] = px; sel[j] = s
        qx[j] = logpdf_mixture_mh(px, p0, c, mu, RPSD, Rdiag, s)
        q0[j] = logpdf_mixture_mh(p0, px, c, mu, RPSD, Rdiag, s)
    lx, st = loglik(prop)
    lx = np.where(st != 0, -np.inf, lx)
    accepted = np.zeros(N, dtype=bool)
    with np.errstate(all="ignore"):
        for j in range(N):
            post0 = logpost[j] + dphi * loglh[j]
            prior = logprior_fn(prop[j])
            postx = phi[i] * lx[j] + prior
            a = math.exp(min((postx - qx[j]) - (post0 - q0[j]), 700.0)) if not math.isnan((postx - qx[j]) - (post0 - q0[j])) else math.nan
            if rand["u_acc"][j] < a:
                para[j] = prop[j]; loglh[j] = lx[j]; logpost[j] = postx; accepted[j] = True
            else:
                logpost[j] = post0
    acpt = float(np.mean(accepted))
    new_state = dict(para=para, loglh=loglh, logpost=logpost, w=w)
    info = dict(zhat=zhat, ESS=ess, resampled=bool(resampled), c=c, acpt=acpt, clamped=nclamp, mu=mu, R=R, prop=prop,
                qx=qx, q0=q0, mixsel=sel, accepted=accepted, lx=lx)
    return new_state, dict(tune, c=c, acpt=acpt), info


# ---- RWMH ---------------------------------------------------------------------------------------------------
def rwmh_chain(y0, Lchol, n_iter, z, u, logtarget):
    """One chain: y' = y + L z_t; accept if log(u_t) < logtarget(y') - logtarget(y).
    z (n_iter, np), u (n_iter,).  Returns (history (n_iter, np), logp history, accepted count)."""
    y = np.array(y0, dtype=np.float64)
    lp = logtarget(y)
    hist = np.empty((n_iter, y.size)); hlp = np.empty(n_iter); acc = 0
    for t in range(n_iter):
        yp = y + Lchol @ z[t]
        lpx = logtarget(yp)
        with np.errstate(all="ignore"):
            if math.log(u[t]) < lpx - lp:
                y, lp = yp, lpx; acc += 1
        hist[t] = y; hlp[t] = lp
    return hist, hlp, acc


def prior_draws(codes, a, b, seed, first, n, rnd=0):
    """prior_draw! (src/estimation/estimation.jl:329-335: pdraw[i] = rand(prior_i)) with the engine's counter-based stream
    (sampler_kernels.cuh prior_draw): (n, np) draws of particles first .. first + n - 1 in round `rnd`.  codes: the
    reference's shape codes (1 Beta, 2 Gamma, 3 Normal, 4 InverseGamma1, 5 Uniform, 6 InverseGamma, 8 Weibull)."""
    from . import rng as R
    idx = np.arange(first, first + n, dtype=np.uint64)
    out = np.empty((n, len(codes)))
    for k, (code, ak, bk) in enumerate(zip(codes, a, b)):
        kbase = k << 12
        if code == 1:
            x = R.gamma_mt(seed, idx, rnd, k, 0, ak); y = R.gamma_mt(seed, idx, rnd, k, 1, bk)
            out[:, k] = x / (x + y)
        elif code == 2:
            out[:, k] = bk * R.gamma_mt(seed, idx, rnd, k, 0, ak)
        elif code == 3:
            out[:, k] = ak + bk * R.normal_first(seed, idx, rnd, R.STREAM_PRIOR, kbase)
        elif code == 4:
            out[:, k] = np.sqrt(bk / R.gamma_mt(seed, idx, rnd, k, 0, ak))
        elif code == 5:
            out[:, k] = ak + (bk - ak) * R.uniform(seed, idx, rnd, R.STREAM_PRIOR, kbase)
        elif code == 6:
            out[:, k] = bk / R.gamma_mt(seed, idx, rnd, k, 0, ak)
        elif code == 8:
            out[:, k] = bk * (-np.log(1.0 - R.uniform(seed, idx, rnd, R.STREAM_PRIOR, kbase))) ** (1.0 / ak)
        else:
            raise ValueError(code)
    return out
