"""Host-side priors against scipy.stats (reference src/estimation/priors.jl:147-178 and
src/distributions/*.jl)."""
import math

import numpy as np
import scipy.stats as ss

from dynare_jl_b200.priors import PriorSet, make_prior


def test_beta_gamma_normal_match_scipy():
    x = np.array([0.05, 0.3, 0.72])
    p = make_prior("beta", 0.356, 0.02)
    assert np.allclose(p.logpdf(x), ss.beta(p.a, p.b).logpdf(x))
    assert math.isclose(ss.beta(p.a, p.b).mean(), 0.356, rel_tol=1e-12)
    assert math.isclose(ss.beta(p.a, p.b).std(), 0.02, rel_tol=1e-10)
    g = make_prior("gamma", 2.5, 1.0)
    assert np.allclose(g.logpdf(x * 5), ss.gamma(g.a, scale=g.b).logpdf(x * 5))
    assert math.isclose(ss.gamma(g.a, scale=g.b).mean(), 2.5, rel_tol=1e-12)
    n = make_prior("normal", 1.0002, 0.007)
    assert np.allclose(n.logpdf(x + 0.5), ss.norm(1.0002, 0.007).logpdf(x + 0.5))


def test_inverse_gamma1_infinite_std_rule_as_coded():
    """distribution_parameters.jl:42-44: alpha = 1, theta = mu^2 Gamma(1)/Gamma(1/2) -- the code, not the intent."""
    p = make_prior("inv_gamma", 0.035449, float("inf"))
    assert p.a == 1.0 and math.isclose(p.b, 0.035449 ** 2 / math.sqrt(math.pi), rel_tol=1e-14)
    x = np.array([0.01, 0.03, 0.2])
    ref = math.log(2) + p.a * math.log(p.b) - math.lgamma(p.a) - (2 * p.a + 1) * np.log(x) - p.b / x ** 2
    assert np.allclose(p.logpdf(x), ref)
    assert np.all(np.isneginf(p.logpdf(np.array([-1.0, 0.0]))))


def test_inverse_gamma1_finite_std_moments():
    p = make_prior("inv_gamma", 1.2533, 0.6551)
    # X = sqrt(Y), Y ~ InverseGamma(a, b):  E X = sqrt(b) G(a-1/2)/G(a),  E X^2 = b/(a-1)
    mean = math.sqrt(p.b) * math.gamma(p.a - 0.5) / math.gamma(p.a)
    var = p.b / (p.a - 1) - mean * mean
    assert math.isclose(mean, 1.2533, rel_tol=1e-8)
    assert math.isclose(math.sqrt(var), 0.6551, rel_tol=1e-6)
    # density integrates to one
    xs = np.linspace(1e-3, 60, 400001)
    assert abs(np.trapezoid(np.exp(p.logpdf(xs)), xs) - 1.0) < 1e-4


def test_prior_set_sum_and_sampling_support():
    from dynare_jl_b200.engine import model_description
    ps = PriorSet.from_description(model_description("fs2000")["priors"])
    rng = np.random.default_rng(3)
    th = ps.sample(rng, 1000)
    assert th.shape == (1000, 9)
    lp = ps.logpriordensity(th)
    assert np.all(np.isfinite(lp))
    one = sum(p.logpdf(np.array([th[5, k]]))[0] for k, p in enumerate(ps.priors))
    assert math.isclose(lp[5], one, rel_tol=1e-13)


def test_uniform_prior_from_domain():
    """`prior!(:x, shape=Uniform, domain=[a, b])` takes (a, b) as the bounds and lets mean / stdev be missing
    (src/estimation/priors.jl:165-170); an `estimated_params` row with empty mean and std takes p3, p4
    (parse_prior_distribution(::Val{5}), priors.jl:508-517); the lb / ub columns are not read."""
    from dynare_jl_b200.modfile import parse_mod
    text = """
    var y; varexo e; parameters rho sig tau;
    rho = 0.5; sig = 1; tau = 0.2;
    model; y = rho*y(-1) + sig*tau*e; end;
    shocks; var e; stderr 1; end;
    prior!(:rho, shape=Uniform, domain=[0, 1])
    prior!(:sig, shape=Uniform, mean=1.0, stdev=0.5, domain=[0.2, 3.0])
    prior!(:tau, shape=Uniform, mean=1.0, stdev=0.5)
    varobs y;
    """
    mf = parse_mod(text)
    eps = {e.name1: e for e in mf.estimated_params}
    assert eps["rho"].domain == (0.0, 1.0) and eps["rho"].mean == 0.5 and abs(eps["rho"].stdev - 1 / 12 ** 0.5) < 1e-15
    assert eps["sig"].domain == (0.2, 3.0) and eps["tau"].domain is None
    pr = {k: make_prior(e.shape, e.mean, e.stdev, e.domain) for k, e in eps.items()}
    assert (pr["rho"].a, pr["rho"].b) == (0.0, 1.0)
    assert (pr["sig"].a, pr["sig"].b) == (0.2, 3.0)                     # the domain wins over the moments
    assert abs(pr["tau"].a - (1.0 - 0.5 * 3 ** 0.5)) < 1e-15 and abs(pr["tau"].b - (1.0 + 0.5 * 3 ** 0.5)) < 1e-15
    assert pr["rho"].logpdf(np.array([0.3]))[0] == 0.0 and pr["rho"].logpdf(np.array([1.2]))[0] == -np.inf
    # the spec / JSON route keeps the domain
    ps = PriorSet.from_description([dict(shape="uniform", mean=0.5, stdev=0.1, domain=[0.0, 1.0]),
                                    dict(shape="uniform", mean=0.5, stdev=0.1)])
    assert (ps.priors[0].a, ps.priors[0].b) == (0.0, 1.0) and ps.priors[1].a > 0.3
    block = """
    var y; varexo e; parameters rho sig;
    rho = 0.5; sig = 1;
    model; y = rho*y(-1) + sig*e; end;
    shocks; var e; stderr 1; end;
    estimated_params;
    rho, 0.4, 0.01, 0.99, uniform_pdf, , , 0.1, 0.9;
    sig, uniform_pdf, 1.0, 0.2;
    end;
    varobs y;
    """
    mf2 = parse_mod(block)
    e0, e1 = mf2.estimated_params
    assert e0.domain == (0.1, 0.9) and e0.init == 0.4 and e1.domain is None
    p0 = make_prior(e0.shape, e0.mean, e0.stdev, e0.domain)
    assert (p0.a, p0.b) == (0.1, 0.9)
