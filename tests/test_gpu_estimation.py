"""The reference's estimation entry points (smc_compute!, mode_compute!, rwmh_compute!, run_simulations) through their
host mirrors in dynare_jl_b200/estimation.py: every loop over draws is one batched call on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_entry_points_on_fs2000(golden):
    from dynare_jl_b200 import estimation as est
    from dynare_jl_b200.engine import BatchSSWs
    g = golden("fs2000")
    ctx = est.make_context("fs2000")
    Y = est.get_observations(ctx, data=g["Y"], first_obs=1, nobs=192)
    assert Y.shape == (2, 192) and np.array_equal(Y, g["Y"])
    assert est.get_observations(ctx, data=g["Y"], first_obs=11, nobs=50).shape == (2, 50)

    # mode_compute!: BFGS from the estimated_params_init point; objective decreases, Hessian gives finite standard errors
    ws = BatchSSWs("fs2000", g["Y"]); ws.set_priors(ctx.priors)
    x0 = est.get_initial_value_or_mean(ctx)
    lp0 = ws.logposterior(x0[None, :])[0][0]
    m = est.mode_compute(ctx, data=g["Y"])
    lp1 = ws.logposterior(m["mode"][None, :])[0][0]
    assert lp1 >= lp0 - 1e-9 and np.isfinite(lp1)
    assert m["std"].shape == (9,) and np.all(np.isfinite(m["std"])) and np.all(m["std"] > 0)
    assert "transformed_posterior_mode_covariance" in ctx.results and ctx.results["posterior_mode"].shape == (9,)
    ws.close()

    # rwmh_compute!: many chains from the mode with the mode's transformed covariance
    r = est.rwmh_compute(ctx, data=g["Y"], initial_values=m["mode"],
                         transformed_covariance=ctx.results["transformed_posterior_mode_covariance"],
                         mcmc_chains=256, mcmc_replic=40, mcmc_jscale=0.3, seed=3)
    assert r["chains"].shape == (40, 256, 9) and 0.05 < r["acceptance_rate"].mean() < 0.9
    assert np.all(np.isfinite(r["chains"])) and len(ctx.results["posterior_mcmc_chains"]) == 1

    # smc_compute!: Tune overridden to a small problem; same result as driving the SMC class by hand
    out = est.smc_compute(ctx, data=g["Y"], npart=1024, nphi=20, seed=5)
    assert np.isfinite(out["logmdd"]) and abs(out["w"].sum() - 1) < 1e-12 and out["posterior_mean"].shape == (9,)
    from dynare_jl_b200.samplers import SMC
    ws2 = BatchSSWs("fs2000", g["Y"])
    s = SMC(ws2, ctx.priors, npart=1024, nphi=20, seed=5)
    s.initialise(); s.run()                                 # device prior draws, the default of smc_compute
    assert np.array_equal(s.particles()["para"], out["para"]) and s.stats()["logmdd"] == out["logmdd"]
    s.close(); ws2.close()

    # run_simulations: prior-predictive moments with the reference's failure classes
    pp = est.prior_predictive_check(ctx, iterations=512)
    assert pp["steady_state"].shape == (512, 16) and pp["failure"].sum() == pp["domain"] + pp["undetermined"] + pp["unstable"] + pp["other"]
    assert np.all(np.isfinite(pp["stdev"][pp["status"] == 0]))
