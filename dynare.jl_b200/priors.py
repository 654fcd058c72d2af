"""Prior densities of the estimated parameters -- host side (SURVEY.md section 8f rank 1 moves them
to the device later).  Mirrors reference ``src/estimation/priors.jl:147-178`` (shape -> distribution),
``src/distributions/distribution_parameters.jl:4-76`` (mean/std -> hyper-parameters, reproduced as
coded, including the ``inv_gamma_pdf`` with ``std = inf`` rule) and
``src/distributions/inversegamma1.jl:134-141`` (InverseGamma1 log-density); the sum over parameters is
``logpriordensity`` (``src/estimation/estimation.jl:455-462``).  Vectorised over draws.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
from scipy import optimize, special


def beta_specification(mu, var, lb=0.0, ub=1.0):
    ln = ub - lb
    mu = (mu - lb) / ln
    var = var / (ln * ln)
    if var > (1 - mu) * mu:
        raise ValueError("Beta prior: std too large for the given mean")
    a = (1 - mu) * mu * mu / var - mu
    b = a * (1 / mu - 1)
    return a, b


def gamma_specification(mu, var):
    theta = var / mu
    return mu / theta, theta


def inverse_gamma_1_specification(mu, var):
    mu2 = mu * mu
    if math.isinf(var):
        return 1.0, mu2 * math.gamma(1.0) / math.gamma(0.5)
    f = lambda x: (math.log(x - 1) + math.log(var + mu2) - math.log(mu2)
                   - 2 * (special.gammaln(x) - special.gammaln(x - 0.5)))
    a = optimize.brentq(f, 1.0 + 1e-12, 100.0, xtol=1e-14, rtol=1e-14)
    return a, (a - 1) * (var + mu2)


def inverse_gamma_2_specification(mu, var):
    a = 2.0 if math.isinf(var) else 2.0 + mu * mu / var
    return a, (a - 1) * mu


def uniform_specification(mu, var):
    b = math.sqrt(12 * var) / 2 + mu
    return 2 * mu - b, b


def weibull_specification(mu, var):
    # the reference's root function (distribution_parameters.jl:70-75) has no zero as written;
    # this is the moment match it intends: var/mu^2 = G(1+2/a)/G(1+1/a)^2 - 1
    f = lambda a: special.gamma(1 + 2 / a) / special.gamma(1 + 1 / a) ** 2 - 1 - var / (mu * mu)
    a = optimize.brentq(f, 0.05, 100.0)
    return a, mu / special.gamma(1 + 1 / a)


# mod-file shape codes of the reference (src/estimation/priors.jl:482-529), used by the C ABI
SHAPE_CODES = {"beta": 1, "gamma": 2, "normal": 3, "inv_gamma": 4, "uniform": 5, "inv_gamma2": 6, "weibull": 8}


@dataclass
class Prior:
    shape: str
    a: float
    b: float

    @property
    def code(self) -> int:
        return SHAPE_CODES[self.shape]

    def logpdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        a, b = self.a, self.b
        out = np.full(x.shape, -np.inf)
        with np.errstate(all="ignore"):
            if self.shape == "beta":
                ok = (x > 0) & (x < 1)
                out[ok] = ((a - 1) * np.log(x[ok]) + (b - 1) * np.log1p(-x[ok]) - special.betaln(a, b))
            elif self.shape == "gamma":            # shape a, scale b
                ok = x > 0
                out[ok] = (a - 1) * np.log(x[ok]) - x[ok] / b - special.gammaln(a) - a * math.log(b)
            elif self.shape == "normal":           # mean a, std b
                out = -0.5 * ((x - a) / b) ** 2 - math.log(b) - 0.5 * math.log(2 * math.pi)
            elif self.shape == "inv_gamma":        # InverseGamma1(alpha=a, theta=b)
                ok = x > 0
                out[ok] = (math.log(2) + a * math.log(b) - special.gammaln(a)
                           - (2 * a + 1) * np.log(x[ok]) - b / (x[ok] * x[ok]))
            elif self.shape == "inv_gamma2":       # InverseGamma(alpha=a, theta=b)
                ok = x > 0
                out[ok] = a * math.log(b) - special.gammaln(a) - (a + 1) * np.log(x[ok]) - b / x[ok]
            elif self.shape == "uniform":          # on [a, b]
                ok = (x >= a) & (x <= b)
                out[ok] = -math.log(b - a)
            elif self.shape == "weibull":          # shape a, scale b
                ok = x > 0
                z = x[ok] / b
                out[ok] = math.log(a / b) + (a - 1) * np.log(z) - z ** a
            else:
                raise ValueError(self.shape)
        return out

    def sample(self, rng: np.random.Generator, size):
        a, b = self.a, self.b
        if self.shape == "beta":
            return rng.beta(a, b, size)
        if self.shape == "gamma":
            return rng.gamma(a, b, size)
        if self.shape == "normal":
            return rng.normal(a, b, size)
        if self.shape == "inv_gamma":              # sqrt of InverseGamma(a, b)  (inversegamma1.jl:167)
            return np.sqrt(b / rng.gamma(a, 1.0, size))
        if self.shape == "inv_gamma2":
            return b / rng.gamma(a, 1.0, size)
        if self.shape == "uniform":
            return rng.uniform(a, b, size)
        if self.shape == "weibull":
            return b * rng.weibull(a, size)
        raise ValueError(self.shape)


def make_prior(shape: str, mean: float, stdev: float, domain=None) -> Prior:
    """``prior_`` (priors.jl:147-178).  ``domain`` = (a, b) gives the bounds of a Uniform prior directly
    (priors.jl:165-170; ``p3, p4`` of an ``estimated_params`` row with empty mean and std, priors.jl:508-517)."""
    if shape == "uniform" and domain is not None and len(domain) == 2:
        return Prior(shape, float(domain[0]), float(domain[1]))
    var = stdev * stdev
    if shape == "beta":
        return Prior(shape, *beta_specification(mean, var))
    if shape == "gamma":
        return Prior(shape, *gamma_specification(mean, var))
    if shape == "normal":
        return Prior(shape, mean, stdev)
    if shape == "inv_gamma":
        return Prior(shape, *inverse_gamma_1_specification(mean, var))
    if shape == "inv_gamma2":
        return Prior(shape, *inverse_gamma_2_specification(mean, var))
    if shape == "uniform":
        return Prior(shape, *uniform_specification(mean, var))
    if shape == "weibull":
        return Prior(shape, *weibull_specification(mean, var))
    raise ValueError(f"unknown prior shape {shape}")


class PriorSet:
    """The ``prior`` vector of ``EstimatedParameters`` (dynare_containers.jl:800-841)."""

    def __init__(self, priors: Sequence[Prior]):
        self.priors: List[Prior] = list(priors)

    @classmethod
    def from_description(cls, desc: Sequence[dict]):
        out = []
        for d in desc:
            sd = float("inf") if d["stdev"] == "inf" else float(d["stdev"])
            out.append(make_prior(d["shape"], float(d["mean"]), sd, d.get("domain")))
        return cls(out)

    def __len__(self):
        return len(self.priors)

    def abi_arrays(self):
        """(shape codes int32, a, b) as dyb_set_priors takes them."""
        return (np.array([p.code for p in self.priors], dtype=np.int32),
                np.array([p.a for p in self.priors], dtype=np.float64),
                np.array([p.b for p in self.priors], dtype=np.float64))

    def logpriordensity(self, theta) -> np.ndarray:
        """theta: (N, np) -> (N,)  sum_k logpdf(prior_k, theta_k)  (estimation.jl:455-462)."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        out = np.zeros(theta.shape[0])
        for k, p in enumerate(self.priors):
            out += p.logpdf(theta[:, k])
        return out

    def sample(self, rng: np.random.Generator, n: int) -> np.ndarray:
        return np.column_stack([p.sample(rng, n) for p in self.priors])
