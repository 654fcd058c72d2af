"""Device-resident samplers (priors, transformations, SMC recursion, RWMH) against the CPU oracle, through the C ABI."""
import math

import numpy as np
import pytest

from oracle import cbuild
from oracle import rng as orng
from oracle import sampler as osam
from oracle import smc as osmc

pytestmark = pytest.mark.gpu

ALL = [("beta", 0.4, 0.1), ("gamma", 2.0, 0.5), ("normal", 0.3, 0.2), ("inv_gamma", 0.05, float("inf")),
       ("inv_gamma", 1.2, 0.6), ("uniform", 0.5, 0.25), ("inv_gamma2", 0.8, 0.3), ("weibull", 1.5, 0.7)]


def _abi(ps):
    code, a, b = ps.abi_arrays()
    return list(zip(code.tolist(), a.tolist(), b.tolist()))


def _model_priors(name):
    from dynare_jl_b200.engine import model_description
    from dynare_jl_b200.priors import PriorSet
    return PriorSet.from_description(model_description(name)["priors"])


def test_generator_matches_oracle():
    from dynare_jl_b200.engine import rng_fill
    for first, n, step in [(0, 1000, 0), (123456789012, 257, 77)]:
        idx = np.arange(first, first + n)
        for stream in (0, 1, 2):
            u = rng_fill(42, first, n, step, stream=stream, kind=0, count=3)
            for b in range(3):
                assert np.array_equal(u[:, b], orng.uniform(42, idx, step, stream, b))       # bit-exact
        z = rng_fill(42, first, n, step, kind=1, count=9)
        assert np.abs(z - orng.normals(42, idx, step, 9)).max() < 1e-13


def test_logprior_and_transform_all_shapes(golden):
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.priors import PriorSet, make_prior
    ps = PriorSet([make_prior(*p) for p in ALL])
    g = golden("fs2000")
    ws = BatchSSWs("fs2000", g["Y"])
    ws.set_estimated_params(np.zeros(8, dtype=np.int32), np.arange(8) % 7, np.zeros(8))
    ws.set_priors(ps)
    x = ps.sample(np.random.default_rng(0), 300)
    x[0, 0] = 1.5; x[1, 1] = -0.1; x[2, 5] = 5.0           # outside the supports
    lp = ws.logprior(x)
    ref = np.array([osam.logprior(_abi(ps), xi) for xi in x])
    fin = np.isfinite(ref)
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    assert np.abs(lp[fin] - ref[fin]).max() < 1e-11
    y = np.random.default_rng(1).normal(size=(100, 8)) * 2
    th, lj = ws.transform(y)
    for i in range(100):
        rth, rlj = osam.transform(_abi(ps), y[i])
        assert np.allclose(th[i], rth, rtol=1e-14, atol=0) and abs(lj[i] - rlj) < 1e-12
    assert np.abs(ws.transform(th, inverse=True) - y).max() < 1e-9
    ws.close()


def test_prior_draws_on_device_match_oracle(golden):
    """prior_draw! on the device (dyb_prior_draw) against its numpy restatement, all seven shapes incl. Gamma / Beta shapes
    below one; the distributions themselves are checked against scipy.stats in tests/test_sampler_oracle.py."""
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.priors import PriorSet, Prior, make_prior
    ps = PriorSet([make_prior(*p) for p in ALL])
    ps.priors[0] = Prior("beta", 0.6, 0.8)                # both Gammas of the ratio take the boost path
    ps.priors[1] = Prior("gamma", 0.4, 2.0)
    g = golden("fs2000")
    ws = BatchSSWs("fs2000", g["Y"])
    ws.set_estimated_params(np.zeros(8, dtype=np.int32), np.arange(8) % 7, np.zeros(8))
    ws.set_priors(ps)
    code, a, b = ps.abi_arrays()
    for first, n, rnd, seed in [(0, 5000, 0, 7), (123456789012, 257, 3, 99)]:
        x = ws.prior_draw(n, seed=seed, first=first, round=rnd)
        ref = osam.prior_draws(code.tolist(), a.tolist(), b.tolist(), seed, first, n, rnd)
        assert x.shape == ref.shape and np.all(np.isfinite(x))
        assert np.allclose(x, ref, rtol=1e-11, atol=0), np.abs(x / ref - 1).max()
        assert np.all(np.isfinite(ws.logprior(x)))         # every draw lies inside its support
    # a shard of the cloud is the same draws
    assert np.array_equal(ws.prior_draw(100, seed=7, first=1000), ws.prior_draw(5000, seed=7)[1000:1100])
    ws.close()


def test_smc_initialised_on_the_device(golden):
    """smc.jl:84-101 on the device: every initial particle has a finite likelihood, is an oracle prior draw of SOME round, and
    the run is reproducible."""
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.samplers import SMC
    g = golden("ls2003")
    ps = _model_priors("ls2003")
    ws = BatchSSWs("ls2003", g["Y"])
    code, a, b = ps.abi_arrays()
    outs = []
    for _ in range(2):
        s = SMC(ws, ps, npart=2048, nphi=6, seed=11)
        s.initialise()
        draws = s.prior_draws
        p0 = s.particles()
        s.run()
        outs.append((p0, s.particles(), s.stats()["logmdd"]))
        s.close()
    p0 = outs[0][0]
    assert draws >= 2048 and np.all(np.isfinite(p0["loglh"])) and np.all(np.isfinite(p0["logpost"]))
    ll, st = ws.loglikelihood(p0["para"])
    assert np.all(st == 0) and np.allclose(ll, p0["loglh"], rtol=1e-12, atol=0)
    rounds = [osam.prior_draws(code.tolist(), a.tolist(), b.tolist(), 11, 0, 2048, r) for r in range(8)]
    hit = np.zeros(2048, dtype=bool)
    for r in rounds:
        hit |= np.all(np.isclose(p0["para"], r, rtol=1e-11, atol=0), axis=1)
    assert hit.all()
    # round 0 is kept wherever its likelihood is finite
    ll0, st0 = ws.loglikelihood(rounds[0])
    keep = st0 == 0
    assert np.allclose(p0["para"][keep], rounds[0][keep], rtol=1e-11, atol=0)
    assert np.array_equal(outs[0][1]["para"], outs[1][1]["para"]) and outs[0][2] == outs[1][2]
    ws.close()


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_logposterior_matches_oracle(name, golden):
    from dynare_jl_b200.engine import BatchSSWs, DSGELogPosteriorDensity
    g = golden(name)
    ps = _model_priors(name)
    ws = BatchSSWs(name, g["Y"])
    prob = DSGELogPosteriorDensity(ws, ps)
    assert prob.dimension() == g["theta"].shape[1]
    lp = prob(g["theta"])
    prior = np.array([osam.logprior(_abi(ps), t) for t in g["theta"]])
    ref = np.where(g["status"] == 0, prior + g["loglik"], -np.inf)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isneginf(lp), ~ok)
    assert np.max(np.abs(lp[ok] - ref[ok]) / np.abs(ref[ok])) < 1e-8
    ws.close()


def _oracle_loglik(name, Y):
    def f(theta):
        ll, st, _ = cbuild.loglik_batch(name, theta, Y)
        return ll, st
    return f


@pytest.mark.parametrize("name,npart,resampling", [("ls2003", 2048, "multinomial"), ("fs2000", 1500, "systematic"),
                                                   ("hs_nk", 1024, "multinomial")])
def test_smc_stages_match_oracle(name, npart, resampling, golden):
    """Every tempering stage, replayed by the CPU oracle from the device state of the previous stage with the
    same random numbers: weights, ESS, resampling decision, moments, proposals, accept decisions, particles."""
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.samplers import SMC
    g = golden(name)
    ps = _model_priors(name)
    pri = _abi(ps)
    ws = BatchSSWs(name, g["Y"])
    nphi, seed = 8, 20260101
    s = SMC(ws, ps, npart=npart, nphi=nphi, seed=seed, resampling=resampling)
    s.initialise(np.random.default_rng(5))
    state = s.particles()
    npar = state["para"].shape[1]
    # initial state: loglh / logpost against the oracle
    ll0, st0 = _oracle_loglik(name, g["Y"])(state["para"])
    assert np.all(st0 == 0)
    assert np.max(np.abs(state["loglh"] - ll0) / np.abs(ll0)) < 1e-8
    assert np.allclose(state["logpost"], [osam.logprior(pri, p) for p in state["para"]], rtol=1e-12, atol=1e-10)
    assert np.all(state["w"] == 1.0 / npart)
    tune = dict(c=0.5, acpt=0.25, trgt=0.25, alp=0.9)
    idx = np.arange(npart)
    n_resampled = 0
    for i in range(1, nphi):
        s.run(1)
        new = s.particles()
        stt = s.stats()
        if resampling == "systematic":
            u_res = float(orng.uniform(seed, np.array([0]), i, orng.STREAM_RESAMPLE)[0])
        else:
            u_res = orng.uniform(seed, idx, i, orng.STREAM_RESAMPLE)
        rand = dict(u_res=u_res, u_mix=orng.uniform(seed, idx, i, orng.STREAM_MIX),
                    u_acc=orng.uniform(seed, idx, i, orng.STREAM_ACCEPT), z=orng.normals(seed, idx, i, npar))
        ref, tune, info = osam.smc_stage(state, i, s.phi, tune, rand, _oracle_loglik(name, g["Y"]),
                                         lambda t: osam.logprior(pri, t), systematic=(resampling == "systematic"))
        row = stt["stats"][i]
        assert abs(row[0] - info["zhat"]) <= 1e-12 * abs(info["zhat"])
        assert abs(row[1] - info["ESS"]) <= 1e-10 * info["ESS"]
        assert bool(row[2]) == info["resampled"]
        assert abs(row[3] - info["c"]) <= 1e-14 and abs(row[4] - info["acpt"]) <= 1e-15
        assert np.allclose(stt["mu"], info["mu"], rtol=1e-11, atol=1e-13)
        assert np.allclose(np.tril(stt["R"]), np.tril(info["R"]), rtol=1e-9, atol=1e-15)
        assert np.allclose(new["w"], ref["w"], rtol=1e-12, atol=0)
        assert np.allclose(new["para"], ref["para"], rtol=1e-9, atol=1e-12)
        assert np.allclose(new["loglh"], ref["loglh"], rtol=1e-8, atol=0)
        assert np.allclose(new["logpost"], ref["logpost"], rtol=1e-8, atol=1e-8)
        n_resampled += info["resampled"]
        state = new                       # replay each stage from the DEVICE state: errors cannot accumulate silently
    assert n_resampled >= 1
    assert abs(stt["logmdd"] - np.sum(np.log(stt["stats"][1:, 0]))) < 1e-9
    s.close(); ws.close()


def test_smc_is_reproducible_and_seed_sensitive(golden):
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.samplers import SMC
    g = golden("ls2003")
    ps = _model_priors("ls2003")
    ws = BatchSSWs("ls2003", g["Y"])
    outs = []
    for seed in (1, 1, 2):
        s = SMC(ws, ps, npart=1024, nphi=12, seed=seed)
        s.initialise(np.random.default_rng(0))
        s.run()
        outs.append((s.particles(), s.stats()))
        s.close()
    assert np.array_equal(outs[0][0]["para"], outs[1][0]["para"]) and outs[0][1]["logmdd"] == outs[1][1]["logmdd"]
    assert not np.array_equal(outs[0][0]["para"], outs[2][0]["para"])
    st = outs[0][1]
    assert st["stats"][-1, 4] > 0.0 and np.all(st["ESS"][1:] > 0) and st["resampled"].sum() >= 1
    ws.close()


def test_rwmh_matches_oracle(golden):
    from dynare_jl_b200.engine import BatchSSWs, rng_fill
    from dynare_jl_b200.samplers import rwmh
    from oracle import models
    name = "fs2000"
    g = golden(name)
    ps = _model_priors(name)
    pri = _abi(ps)
    ws = BatchSSWs(name, g["Y"])
    ws.set_priors(ps)
    m = models.get_model(name)
    nchain, niter, seed, first = 6, 25, 99, 1000
    npar = len(pri)
    y0 = osam.inverse_transform(pri, m.theta_mode)
    rng = np.random.default_rng(2)
    y0s = y0 + 0.01 * rng.standard_normal((nchain, npar))
    cov = np.diag(np.full(npar, 0.02 ** 2)); cov[0, 1] = cov[1, 0] = 1e-4
    out = rwmh(ws, ps, y0s, cov, mcmc_replic=niter, mcmc_chains=nchain, seed=seed, first_chain=first, keep_history=True)
    Lc = np.linalg.cholesky((cov + cov.T) / 2)
    like = _oracle_loglik(name, g["Y"])

    def target(y):
        th, lj = osam.transform(pri, y)
        ll, st = like(th[None, :])
        v = osam.logprior(pri, th) + ll[0]
        return -math.inf if (st[0] != 0 or not math.isfinite(v)) else v + lj

    total_acc = 0
    for cidx in range(nchain):
        z = np.stack([orng.normals(seed, np.array([first + cidx]), t, npar)[0] for t in range(niter)])
        u = np.array([orng.uniform(seed, np.array([first + cidx]), t, orng.STREAM_ACCEPT)[0] for t in range(niter)])
        hist, hlp, acc = osam.rwmh_chain(y0s[cidx], Lc, niter, z, u, target)
        assert np.allclose(out["history"][:, cidx, :], hist, rtol=1e-10, atol=1e-12)
        assert np.allclose(out["history_logp"][:, cidx], hlp, rtol=1e-8)
        assert out["accepted"][cidx] == acc
        total_acc += acc
    assert 0 < total_acc < nchain * niter
    ws.close()


def test_rwmh_chains_can_be_sharded(golden):
    """Chains are independent and their random numbers are keyed by the global chain index: running them in two
    shards (two GPUs / two calls) reproduces the single-call result bit for bit."""
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.samplers import rwmh
    g = golden("ls2003")
    ps = _model_priors("ls2003")
    ws = BatchSSWs("ls2003", g["Y"])
    ws.set_priors(ps)
    ok = np.flatnonzero(g["status"] == 0)[:64]
    y0 = ws.transform(g["theta"][ok], inverse=True)
    cov = np.diag(np.full(ws.np, 0.03 ** 2))
    full = rwmh(ws, ps, y0, cov, mcmc_replic=15, mcmc_chains=64, seed=5, keep_history=True)
    a = rwmh(ws, ps, y0[:32], cov, mcmc_replic=15, mcmc_chains=32, seed=5, first_chain=0, keep_history=True)
    b = rwmh(ws, ps, y0[32:], cov, mcmc_replic=15, mcmc_chains=32, seed=5, first_chain=32, keep_history=True)
    assert np.array_equal(full["history"], np.concatenate([a["history"], b["history"]], axis=1))
    assert np.array_equal(full["accepted"], np.concatenate([a["accepted"], b["accepted"]]))
    # a run can be continued: 15 sweeps == 10 sweeps + 5 sweeps started at step 10
    c1 = rwmh(ws, ps, y0, cov, mcmc_replic=10, mcmc_chains=64, seed=5)
    c2 = rwmh(ws, ps, c1["y"], cov, mcmc_replic=5, mcmc_chains=64, seed=5, first_step=10)
    assert np.array_equal(c2["y"], full["y"]) and 0 < full["accepted"].sum() < 64 * 15
    ws.close()
