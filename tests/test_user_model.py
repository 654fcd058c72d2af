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
