"""A medium-scale model (BASELINE configs[3]: ~40 states, 7 observables, T = 200) through the whole path: mc4, four linked
Lubik-Schorfheide economies (44 endogenous variables, dynamic block m = 28, 30 Kalman states, 33 estimated parameters).
It is beyond the register-resident kernels (m <= 12), so its generated model function feeds the run-time-size solver and
the tensor-core dense filter on the device (ModelOps::jacobian_batch, csrc/generic.cu) -- no host hop."""
import numpy as np
import pytest

from oracle import models, pipeline as pl

pytestmark = pytest.mark.gpu


def test_mc4_loglik_policy_and_status_match_the_oracle(golden):
    from dynare_jl_b200.engine import BatchSSWs
    g = golden("mc4")
    ws = BatchSSWs("mc4", g["Y"])
    assert ws.dims["n"] == 44 and ws.dims["ns"] == 30 and ws.dims["ny"] == 7 and ws.np == 33
    ll, st = ws.loglikelihood(g["theta"])
    assert np.array_equal(st, g["status"]) and (g["status"] != 0).sum() >= 1
    ok = g["status"] == 0
    assert np.max(np.abs(ll[ok] - g["loglik"][ok]) / np.abs(g["loglik"][ok])) <= 1e-8
    assert np.all(np.isneginf(ll[~ok]))
    fo = ws.first_order(g["theta"])
    assert np.array_equal(fo["status"], g["status"])
    g1 = np.concatenate([fo["g1_1"], fo["g1_2"]], axis=2)
    assert np.abs(g1[ok] - g["g1"][ok]).max() <= 1e-10                 # vs the generalised-Schur oracle
    # a fresh draw against the live oracle, and position independence inside a larger batch
    m = models.get_model("mc4")
    th = g["theta"][ok][:3]
    big = np.tile(th, (700, 1))
    llb, stb = ws.loglikelihood(big)
    assert np.all(stb == 0) and np.array_equal(llb[:3], llb[-3:])
    ref, _ = pl.loglikelihood(m, th[1], g["Y"])
    assert abs(llb[1] - ref) <= 1e-8 * abs(ref)
    ws.close()


def test_mc4_samplers_run_on_the_device(golden):
    """RWMH on many chains and a short SMC recursion (CUDA-graph replay included) with the medium-scale model."""
    from dynare_jl_b200 import estimation as est
    g = golden("mc4")
    ctx = est.make_context("mc4")
    x0 = est.get_initial_value_or_mean(ctx)
    r = est.rwmh_compute(ctx, data=g["Y"], initial_values=x0, transformed_covariance=np.diag(np.full(33, 0.02 ** 2)),
                         mcmc_chains=128, mcmc_replic=12, mcmc_jscale=1.0, seed=4)
    assert r["chains"].shape == (12, 128, 33) and np.all(np.isfinite(r["chains"])) and 0.0 < r["acceptance_rate"].mean() < 1.0
    out = est.smc_compute(ctx, data=g["Y"], npart=8192, nphi=60, seed=2)
    assert np.isfinite(out["logmdd"]) and abs(out["w"].sum() - 1) < 1e-12 and np.all(np.isfinite(out["loglh"]))


def test_mc4_register_blocked_solver_matches_the_shared_memory_solver(golden):
    """solve_block_kernel (Gauss-Jordan with implicit pivoting in registers) against solve_generic_kernel (pivoted LU in
    shared memory): same status for every draw, policy matrices and likelihood to rounding; failing draws included."""
    from dynare_jl_b200.engine import BatchSSWs, model_description
    from dynare_jl_b200.priors import PriorSet
    g = golden("mc4")
    pri = PriorSet.from_description(model_description("mc4")["priors"])
    th = np.concatenate([g["theta"], pri.sample(np.random.default_rng(7), 600)])
    ws = BatchSSWs("mc4", g["Y"])
    out = {}
    for solver in (0, 1):
        ws.set_options(solver=solver)
        ll, st = ws.loglikelihood(th)
        fo = ws.first_order(th)
        out[solver] = (ll, st, fo)
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2]["status"], out[1][2]["status"])
    assert (out[0][1] != 0).sum() >= 5
    ok = out[0][1] == 0
    assert np.max(np.abs(out[0][0][ok] - out[1][0][ok]) / np.abs(out[1][0][ok])) <= 1e-9
    for k in ("g1_1", "g1_2"):
        assert np.abs(out[0][2][k][ok] - out[1][2][k][ok]).max() <= 1e-10
    assert np.array_equal(out[0][2]["cr_iters"][ok], out[1][2]["cr_iters"][ok])
    ws.close()


def test_mc4_wide_instantiation_of_the_register_blocked_solver(golden, monkeypatch):
    """mc4 fits the narrow instantiation solve_block_kernel<4,5,3>; the wide one <4,6,4> (models with m + k + nf > 80 or
    k + nf > 48) is forced through the library's test hook and must return the same bits: every element of the Gauss-Jordan
    array and of the products sees the same operations, only the number of (zero) padding columns differs."""
    from dynare_jl_b200.engine import BatchSSWs
    g = golden("mc4")
    ws = BatchSSWs("mc4", g["Y"])
    ll0, st0 = ws.loglikelihood(g["theta"])
    fo0 = ws.first_order(g["theta"])
    ws.close()
    monkeypatch.setenv("DYB_SB_WIDE", "1")
    ws = BatchSSWs("mc4", g["Y"])
    ll1, st1 = ws.loglikelihood(g["theta"])
    fo1 = ws.first_order(g["theta"])
    ws.close()
    assert np.array_equal(st0, st1) and np.array_equal(st0, g["status"])
    ok = st0 == 0
    assert np.array_equal(ll0[ok], ll1[ok])
    for k in ("g1_1", "g1_2"):
        assert np.array_equal(fo0[k][ok], fo1[k][ok])
