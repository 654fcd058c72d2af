"""A user model compiled at run time from its .mod text into its own shared object and registered with the library
(the run-time analogue of `use_dll`).  The model text below is a two-equation stand-in written for the test."""
import os
import shutil

import numpy as np
import pytest

MOD = """
var x pi_obs;
varexo e u;
parameters rho kap sig;
rho = 0.8; kap = 0.3; sig = 0.5;
model;
  x = rho*x(-1) + sig*e;
  pi_obs = kap*x + 0.5*pi_obs(+1) + 0.1*u;
end;
steady_state_model;
  x = 0; pi_obs = 0;
end;
shocks;
  var e; stderr 1;
  var u; stderr 1;
end;
varobs pi_obs;
estimated_params;
  rho, beta_pdf, 0.7, 0.1;
  kap, gamma_pdf, 0.3, 0.1;
  stderr e, inv_gamma_pdf, 1.0, inf;
end;
"""


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("no nvcc")
    from dynare_jl_b200 import build
    d = tmp_path_factory.mktemp("usermodel")
    mod = d / "toy2.mod"
    mod.write_text(MOD)
    return build.build_model_library("toy2", str(mod), str(d))


def test_user_model_compiles_and_registers(built):
    from dynare_jl_b200.engine import available_models, load_model_library, model_description
    assert os.path.exists(built)
    load_model_library(built)
    assert "toy2" in available_models()
    d = model_description("toy2")
    assert d["n"] == 2 and d["nx"] == 2 and d["np_est"] == 3 and d["varobs"] == ["pi_obs"]


@pytest.mark.gpu
def test_user_model_likelihood_matches_closed_form_state_space(built):
    """x is AR(1), pi solves forward: pi = kap/(1 - 0.5 rho) x + 0.1 u, so the likelihood is that of
    y_t = c x_t + 0.1 u_t; checked against the oracle's Kalman filter on the closed-form state space."""
    from dynare_jl_b200.engine import BatchSSWs, load_model_library
    from oracle import pipeline as pl
    load_model_library(built)
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((1, 60))
    ws = BatchSSWs("toy2", Y)
    theta = np.array([[0.8, 0.3, 1.0], [0.5, 0.6, 0.7], [0.95, 0.1, 2.0]])
    ll, st = ws.loglikelihood(theta)
    assert np.all(st == 0)
    for (rho, kap, se), got in zip(theta, ll):
        c = kap / (1 - 0.5 * rho)
        # states (x, pi): x' = rho x + 0.5 se e ; pi' = c x' + 0.1 u
        T = np.array([[rho, 0.0], [c * rho, 0.0]])
        R = np.array([[0.5, 0.0], [c * 0.5, 0.1]])
        Q = np.diag([se * se, 1.0])
        P0 = __import__("scipy.linalg", fromlist=["x"]).solve_discrete_lyapunov(T, R @ Q @ R.T)
        ref = pl.kalman_likelihood(Y, np.array([[0.0, 1.0]]), np.zeros((1, 1)), T, R, Q, np.zeros(2), P0)
        assert abs(got - ref) <= 1e-8 * abs(ref)
    ws.close()


# ---- a model with a unit root: x is a random walk, pi = 2 kap x + 0.1 u --------------------------------------------
MOD_UR = """
var x pi_obs;
varexo e u;
parameters kap sig;
kap = 0.3; sig = 0.5;
model;
  x = x(-1) + sig*e;
  pi_obs = kap*x + 0.5*pi_obs(+1) + 0.1*u;
end;
steady_state_model;
  x = 0; pi_obs = 0;
end;
shocks;
  var e; stderr 1;
  var u; stderr 1;
end;
varobs pi_obs;
estimated_params;
  kap, gamma_pdf, 0.3, 0.1;
  stderr e, inv_gamma_pdf, 1.0, inf;
end;
"""


@pytest.fixture(scope="module")
def built_ur(tmp_path_factory):
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("no nvcc")
    from dynare_jl_b200 import build
    d = tmp_path_factory.mktemp("urmodel")
    mod = d / "toyur.mod"
    mod.write_text(MOD_UR)
    return build.build_model_library("toyur", str(mod), str(d))


@pytest.mark.gpu
def test_unit_root_model_is_solved_and_flagged_non_stationary(built_ur):
    """The reference's generalised Schur solver counts |lambda| < 1 + 1e-6 as stable (src/perturbations.jl:663-664), so
    a random walk has a first-order solution; cyclic reduction accepts it when A2(j) vanishes and A0(j) settles.  The
    likelihood needs a finite unconditional variance and is -Inf (status 5), as in the reference's estimation path."""
    from dynare_jl_b200.engine import BatchSSWs, load_model_library
    load_model_library(built_ur)
    Y = np.random.default_rng(0).standard_normal((1, 30))
    ws = BatchSSWs("toyur", Y)
    theta = np.array([[0.3, 1.0], [0.6, 0.7]])
    fo = ws.first_order(theta)
    assert np.all(fo["status"] == 5) and np.all(fo["cr_iters"] <= 8)
    for (kap, se), g11, g12 in zip(theta, fo["g1_1"], fo["g1_2"]):
        assert np.abs(g11 - np.array([[1.0], [2 * kap]])).max() <= 1e-12
        assert np.abs(g12 - np.array([[0.5, 0.0], [2 * kap * 0.5, 0.1]])).max() <= 1e-12
    ll, st = ws.loglikelihood(theta)
    assert np.all(st == 5) and np.all(ll == -np.inf)
    ws.close()


@pytest.mark.gpu
def test_unit_root_model_smoother_matches_diffuse_oracle(built_ur):
    """calibsmoother!'s non-stationary branch (kalman.jl:163-223) at the model level against oracle/diffuse.py on the
    closed-form state space (states x, pi; pi observed)."""
    from dynare_jl_b200.engine import BatchSSWs, load_model_library
    from oracle import diffuse as D
    load_model_library(built_ur)
    rng = np.random.default_rng(1)
    nobs = 40
    theta = np.array([[0.3, 1.0], [0.6, 0.7], [0.2, 1.5]])
    x = np.cumsum(0.5 * rng.standard_normal(nobs))
    Y = (0.6 * x + 0.1 * rng.standard_normal(nobs))[None, :]
    Y[0, 7] = np.nan
    ws = BatchSSWs("toyur", Y)
    out = ws.smoother_nonstationary(theta, Y)
    assert np.all(out["status"] == 0) and np.all(out["n_unit_roots"] == 1) and np.all(out["n_diffuse"] == 1)
    for i, (kap, se) in enumerate(theta):
        T = np.array([[1.0, 0.0], [2 * kap, 0.0]])
        R = np.array([[0.5, 0.0], [kap, 0.1]])
        Q = np.diag([se * se, 1.0])
        Pinf, Pstar, k = D.diffuse_initial_covariances(T, R, Q)
        ref = D.diffuse_kalman_smoother(Y, np.array([1]), np.zeros(1), T, R, Q, np.zeros(2), Pinf, Pstar)
        assert abs(out["loglik"][i] - ref["loglik"]) <= 1e-9 * abs(ref["loglik"])
        assert np.abs(out["alphah"][i].T - ref["alphah"]).max() <= 1e-9 * max(1.0, np.abs(ref["alphah"]).max())
        assert np.abs(out["etah"][i].T - ref["etah"]).max() <= 1e-9 * max(1.0, np.abs(ref["etah"]).max())
    ws.close()


# ---- a model WITHOUT a steady_state_model block: Newton's method on the static model, on the device ------------------
MOD_NUM = """
var c k a y_obs;
varexo e;
parameters alpha beta delta rho sig;
alpha = 0.33; beta = 0.99; delta = 0.025; rho = 0.9; sig = 0.01;
model;
  1/c = beta*(1/c(+1))*(alpha*exp(a(+1))*k^(alpha-1) + 1 - delta);
  c + k = exp(a)*k(-1)^alpha + (1-delta)*k(-1);
  a = rho*a(-1) + sig*e;
  y_obs = exp(a)*k(-1)^alpha;
end;
initval;
  k = 10; c = 1; a = 0; y_obs = 2;
end;
shocks;
  var e; stderr 1;
end;
varobs y_obs;
estimated_params;
  alpha, beta_pdf, 0.33, 0.05;
  rho, beta_pdf, 0.8, 0.1;
  stderr e, inv_gamma_pdf, 1.0, inf;
end;
"""


def _rbc_oracle():
    """The same model written by hand for the oracle, with its closed-form steady state."""
    from oracle.models import HandModel

    class RBC(HandModel):
        name = "rbcnum"
        endo = ["c", "k", "a", "y_obs"]
        exo = ["e"]
        params = ["alpha", "beta", "delta", "rho", "sig"]
        params0 = np.array([0.33, 0.99, 0.025, 0.9, 0.01])
        sigma_e0 = np.array([[1.0]])
        varobs = ["y_obs"]
        lagged = ["k", "a"]
        led = ["c", "a"]
        est = [("param", "alpha"), ("param", "rho"), ("stderr", "e")]

        def steady_state(self, p):
            alpha, beta, delta = p[0], p[1], p[2]
            k = ((1 / beta - 1 + delta) / alpha) ** (1 / (alpha - 1))
            return np.array([k ** alpha - delta * k, k, 0.0, k ** alpha])

        def dynamic_resid(self, ym, y, yp, u, p):
            alpha, beta, delta, rho, sig = p
            c, k, a, yo = y
            return np.array([
                1 / c - beta * (1 / yp[0]) * (alpha * np.exp(yp[2]) * k ** (alpha - 1) + 1 - delta),
                c + k - np.exp(a) * ym[1] ** alpha - (1 - delta) * ym[1],
                a - rho * ym[2] - sig * u[0],
                yo - np.exp(a) * ym[1] ** alpha])
    return RBC()


@pytest.fixture(scope="module")
def built_num(tmp_path_factory):
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("no nvcc")
    from dynare_jl_b200 import build
    d = tmp_path_factory.mktemp("nummodel")
    mod = d / "rbcnum.mod"
    mod.write_text(MOD_NUM)
    return build.build_model_library("rbcnum", str(mod), str(d))


def test_numeric_steady_state_host_mirror_finds_the_closed_form():
    from dynare_jl_b200 import modelgen
    from dynare_jl_b200.modfile import parse_mod
    spec = modelgen.build_spec("rbcnum", parse_mod(MOD_NUM))
    assert spec.ss_numeric and spec.describe()["steady_state"] == "numeric"
    hm = modelgen.HostModel(spec)
    m = _rbc_oracle()
    for alpha in (0.33, 0.2, 0.45):
        p = spec.param_values.copy(); p[0] = alpha
        assert np.abs(hm.steady_state(p) - m.steady_state(p)).max() <= 1e-11 * np.abs(m.steady_state(p)).max()
    p = spec.param_values.copy(); p[0] = 0.0                    # 1 = beta (1 - delta): no steady state, reported as NaN
    assert not np.all(np.isfinite(hm.steady_state(p)))


@pytest.mark.gpu
def test_numeric_steady_state_model_matches_oracle(built_num):
    """Steady state by Newton's method on the device (no steady_state_model block), then the usual path: policy
    matrices and likelihood against the oracle run on the hand-written model with its closed-form steady state."""
    from dynare_jl_b200.engine import BatchSSWs, load_model_library, model_description
    from oracle import pipeline as pl
    load_model_library(built_num)
    assert model_description("rbcnum")["steady_state"] == "numeric"
    m = _rbc_oracle()
    rng = np.random.default_rng(4)
    theta = np.column_stack([rng.uniform(0.2, 0.45, 40), rng.uniform(0.3, 0.97, 40), rng.uniform(0.5, 2.0, 40)])
    Y = m.steady_state(m.params0)[3] + 0.05 * rng.standard_normal((1, 50))
    ws = BatchSSWs("rbcnum", Y)
    fo = ws.first_order(theta)
    ll, st = ws.loglikelihood(theta)
    assert np.all(st == 0) and np.all(fo["status"] == 0)
    for i, th in enumerate(theta):
        d = pl.solve_draw(m, th)
        assert np.abs(fo["ybar"][i] - d["ybar"]).max() <= 1e-10 * np.abs(d["ybar"]).max()
        assert np.abs(fo["g1_1"][i] - d["sol"]["g1_1"]).max() <= 1e-10 * max(1.0, np.abs(d["sol"]["g1_1"]).max())
        assert np.abs(fo["g1_2"][i] - d["sol"]["g1_2"]).max() <= 1e-10 * max(1.0, np.abs(d["sol"]["g1_2"]).max())
        ref, rst = pl.loglikelihood(m, th, Y)
        assert rst == 0 and abs(ll[i] - ref) <= 1e-8 * abs(ref)
    # a draw without a steady state is flagged, the others are unaffected
    bad = theta.copy(); bad[3, 0] = 0.0
    ll2, st2 = ws.loglikelihood(bad)
    assert st2[3] == 1 and ll2[3] == -np.inf and np.all(st2[np.arange(40) != 3] == 0)
    ws.close()
