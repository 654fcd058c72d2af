"""Edge cases of the C ABI: empty batches, argument errors, asynchronous calls, option handling (through ctypes)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ws(golden):
    from dynare_jl_b200.engine import BatchSSWs
    g = golden("fs2000")
    w = BatchSSWs("fs2000", g["Y"])
    yield w
    w.close()


def _priors(name="fs2000"):
    from dynare_jl_b200.engine import model_description
    from dynare_jl_b200.priors import PriorSet
    return PriorSet.from_description(model_description(name)["priors"])


def test_empty_batches_are_no_ops(ws):
    from dynare_jl_b200 import _lib as L
    lib = ws.lib
    z = np.zeros((0, 9)); e = np.zeros(0); ei = np.zeros(0, dtype=np.int32)
    assert lib.dyb_loglik_batch(ws._h, L.ptr(z), 0, L.ptr(e), L.ptr(ei)) == 0
    ws.set_priors(_priors())
    assert lib.dyb_logprior_batch(ws._h, L.ptr(z), 0, L.ptr(e)) == 0
    assert lib.dyb_logposterior_batch(ws._h, L.ptr(z), 0, L.ptr(e), None, None) == 0
    assert lib.dyb_rwmh_run(ws._h, 0, 0, 5, 0, 1, None, 0, None, None, None, None, None) == 0


def test_argument_errors_are_reported_not_thrown(ws):
    from dynare_jl_b200 import _lib as L
    from dynare_jl_b200.priors import PriorSet, make_prior
    lib = ws.lib
    th = np.zeros((4, 9)); ll = np.zeros(4); st = np.zeros(4, dtype=np.int32)
    assert lib.dyb_loglik_batch(ws._h, None, 4, L.ptr(ll), L.ptr(st)) == 1            # DYB_ERR_ARG
    assert b"null" in lib.dyb_last_error()
    with pytest.raises(ValueError):
        ws.set_priors(PriorSet([make_prior("normal", 0.0, 1.0)]))                      # one prior per parameter
    code = np.array([7] * 9, dtype=np.int32); a = np.ones(9); b = np.ones(9)
    assert lib.dyb_set_priors(ws._h, 9, L.ptr(code), L.ptr(a), L.ptr(b)) == 1          # unknown shape code
    code[:] = 1; a[:] = -1.0
    assert lib.dyb_set_priors(ws._h, 9, L.ptr(code), L.ptr(a), L.ptr(b)) == 1          # Beta needs a > 0
    # the asynchronous call refuses pageable memory instead of silently synchronising
    assert lib.dyb_loglik_batch_async(ws._h, L.ptr(th), 4, L.ptr(ll), L.ptr(st)) == 1
    o = L.SmcOptions()
    lib.dyb_smc_default_options(C.byref(o))
    assert (o.npart, o.nphi, o.lam, o.c0, o.acpt0, o.trgt, o.alp) == (1000, 400, 2.0, 0.5, 0.25, 0.25, 0.9)   # Tune defaults
    o.nphi = 1
    s = C.c_void_p()
    ws.set_priors(_priors())
    assert lib.dyb_smc_create(ws._h, C.byref(o), None, C.byref(s)) == 1                # fewer than two stages
    assert lib.dyb_kalman_batch(0, 4, 5, 3, 0, None, None, None, None, None, None, None, 1, None, None) == 1   # ny > ns


def test_async_call_matches_sync_call(ws, golden):
    from dynare_jl_b200.engine import PinnedArray
    g = golden("fs2000")
    n = g["theta"].shape[0]
    pth = PinnedArray((n, 9)); pll = PinnedArray((n,)); pst = PinnedArray((n,), np.int32)
    pth.array[:] = g["theta"]
    pll.array[:] = 0.0
    ws.loglikelihood_async(pth.array, pll.array, pst.array)
    while not ws.done():                       # completion query of the asynchronous call
        pass
    assert ws.done()
    ws.sync()
    ll, st = ws.loglikelihood(g["theta"])
    assert np.array_equal(pll.array, ll) and np.array_equal(pst.array, st)             # same kernels: bit-identical
    for b in (pth, pll, pst):
        b.free()


def test_logposterior_optional_outputs_and_untransformed_rwmh(ws, golden):
    from dynare_jl_b200 import _lib as L
    from dynare_jl_b200.samplers import rwmh
    g = golden("fs2000")
    pri = _priors()
    ws.set_priors(pri)
    th = np.ascontiguousarray(g["theta"][:32])
    lp = np.empty(32)
    assert ws.lib.dyb_logposterior_batch(ws._h, L.ptr(th), 32, L.ptr(lp), None, None) == 0
    full, ll, st = ws.logposterior(th)
    assert np.array_equal(lp, full)
    # chains in the parameter space itself (transformed_parameters = false, estimation.jl:946-948)
    ok = np.flatnonzero(st == 0)[:4]
    out = rwmh(ws, pri, th[ok], np.diag(np.full(9, 1e-8)), mcmc_replic=5, mcmc_chains=4, transformed_parameters=False, seed=3)
    assert np.allclose(out["logp"][out["accepted"] == 0], full[ok][out["accepted"] == 0], rtol=1e-12)
    assert np.all(np.isfinite(out["logp"]))


def test_smc_on_one_rank_accepts_any_particle_count(golden):
    from dynare_jl_b200.engine import BatchSSWs
    from dynare_jl_b200.samplers import SMC
    g = golden("fs2000")
    w = BatchSSWs("fs2000", g["Y"])
    s = SMC(w, _priors(), npart=777, nphi=6, seed=1)
    s.initialise(np.random.default_rng(0))
    s.run()
    st = s.stats(); p = s.particles()
    assert s.stages_done() == 6 and abs(p["w"].sum() - 1.0) < 1e-12 and np.all(np.isfinite(p["loglh"]))
    assert st["stats"][0, 0] == 1.0 and np.all(st["zhat"][1:] > 0)
    s.close(); w.close()
