"""Parity of the CUDA likelihood path with the CPU oracle, through the C ABI (ctypes)."""
import numpy as np
import pytest

from oracle import models, pipeline as pl

pytestmark = pytest.mark.gpu

RTOL_LOGLIK = 1e-8       # north-star tolerance: log-likelihood within 1e-8 relative (fp64)
ATOL_POLICY = 1e-10      # north-star tolerance: policy matrices within 1e-10


@pytest.fixture(scope="module")
def ws_factory():
    from dynare_jl_b200.engine import BatchSSWs
    made = []

    def make(name, Y):
        w = BatchSSWs(name, Y)
        made.append(w)
        return w
    yield make
    for w in made:
        w.close()


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_loglik_matches_golden(name, golden, ws_factory):
    g = golden(name)
    ws = ws_factory(name, g["Y"])
    ll, st = ws.loglikelihood(g["theta"])
    assert np.array_equal(st, g["status"])
    ok = g["status"] == 0
    rel = np.abs(ll[ok] - g["loglik"][ok]) / np.abs(g["loglik"][ok])
    assert rel.max() <= RTOL_LOGLIK, rel.max()
    assert np.all(np.isneginf(ll[~ok]))


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_loglik_matches_live_oracle_on_fresh_draws(name, golden, ws_factory):
    from dynare_jl_b200.engine import model_description
    from dynare_jl_b200.priors import PriorSet
    g = golden(name)
    m = models.get_model(name)
    pri = PriorSet.from_description(model_description(name)["priors"])
    theta = pri.sample(np.random.default_rng(12345), 200)
    ws = ws_factory(name, g["Y"])
    ll, st = ws.loglikelihood(theta)
    for i in range(theta.shape[0]):
        ref, rst = pl.loglikelihood(m, theta[i], g["Y"])
        assert st[i] == rst, (i, st[i], rst)
        if rst == 0:
            assert abs(ll[i] - ref) <= RTOL_LOGLIK * abs(ref), (i, ll[i], ref)
        else:
            assert ll[i] == -np.inf


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_policy_matrices(name, golden, ws_factory):
    """g1 = [g1_1 | g1_2] against the generalised-Schur oracle (the reference's default dr_algo 'GS')."""
    g = golden(name)
    ws = ws_factory(name, g["Y"])
    out = ws.first_order(g["theta"])
    assert np.array_equal(out["status"], g["status"])
    m = models.get_model(name)
    g1 = np.concatenate([out["g1_1"], out["g1_2"]], axis=2)
    ok = g["status"] == 0
    err = np.abs(g1[ok] - g["g1"][ok]).max()
    assert err <= ATOL_POLICY, err
    assert np.all(out["cr_iters"] > 0) and np.all(out["cr_iters"] < 60)
    p, se, sm = pl.set_estimated_parameters(m, g["theta"][3])
    assert np.allclose(out["ybar"][3], m.steady_state(p), rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_state_space_assembly(name, golden, ws_factory):
    g = golden(name)
    m = models.get_model(name)
    ws = ws_factory(name, g["Y"])
    theta = g["theta"][:16]
    out = ws.state_space(theta)
    for i in range(theta.shape[0]):
        d = pl.solve_draw(m, theta[i])
        ss = d["ss"]
        sc = max(1.0, np.abs(ss["P0"]).max())
        assert np.abs(out["T"][i] - ss["T"]).max() <= 1e-10
        assert np.abs(out["RQR"][i] - ss["R"] @ ss["Q"] @ ss["R"].T).max() <= 1e-12 * sc
        assert np.abs(out["P0"][i] - ss["P0"]).max() <= 1e-9 * sc
        assert np.allclose(out["ybar_obs"][i], d["ybar"][m.obs_idx], rtol=1e-13, atol=1e-15)
        assert np.abs(out["H"][i] - ss["H"]).max() == 0.0
        # Lyapunov identity on the device result (reference test/test_smc.jl:41-42 pattern)
        P0, T = out["P0"][i], out["T"][i]
        assert np.abs(P0 - (T @ P0 @ T.T + out["RQR"][i])).max() <= 1e-9 * sc


def test_status_classification_matches_oracle(golden, ws_factory):
    gl, gf = golden("ls2003"), golden("fs2000")
    ls, fs = models.get_model("ls2003"), models.get_model("fs2000")
    cases_ls = []
    for k, v in [(0, -1.5), (8, 1.2), (10, 1.05), (3, 1.5), (0, -1.01)]:
        th = ls.theta_mode.copy(); th[k] = v; cases_ls.append(th)
    cases_ls.append(ls.theta_mode.copy())
    th = np.array(cases_ls)
    ll, st = ws_factory("ls2003", gl["Y"]).loglikelihood(th)
    for i in range(th.shape[0]):
        ref, rst = pl.loglikelihood(ls, th[i], gl["Y"])
        assert st[i] == rst, (i, th[i], st[i], rst)
        assert (ll[i] == -np.inf) if rst else abs(ll[i] - ref) <= RTOL_LOGLIK * abs(ref)
    cases_fs = []
    for k, v in [(0, 1.5), (1, 1.2), (5, 1.3), (4, 1.01), (2, np.nan), (7, 0.0)]:
        th = fs.theta_mode.copy(); th[k] = v; cases_fs.append(th)
    th = np.array(cases_fs)
    ll, st = ws_factory("fs2000", gf["Y"]).loglikelihood(th)
    for i in range(th.shape[0]):
        ref, rst = pl.loglikelihood(fs, th[i], gf["Y"])
        assert (st[i] != 0) == (rst != 0), (i, th[i], st[i], rst)
        if rst in (pl.ST_STEADY, pl.ST_INDETERMINATE, pl.ST_UNSTABLE):
            assert st[i] == rst, (i, th[i], st[i], rst)
        assert (ll[i] == -np.inf) if rst else abs(ll[i] - ref) <= RTOL_LOGLIK * abs(ref)


def test_missing_observations_and_presample(golden, ws_factory):
    g = golden("ls2003")
    m = models.get_model("ls2003")
    Y = g["Y"].copy()
    Y[1, 5] = np.nan
    Y[:, 20] = np.nan
    Y[[0, 3], 40] = np.nan
    ws = ws_factory("ls2003", Y)
    theta = g["theta"][:24]
    ll, st = ws.loglikelihood(theta)
    for i in range(theta.shape[0]):
        ref, rst = pl.loglikelihood(m, theta[i], Y)
        assert st[i] == rst
        assert abs(ll[i] - ref) <= RTOL_LOGLIK * abs(ref)
    # presample: the first periods update the state but are excluded from the sum (estimation.jl:659-661)
    ws.set_options(presample=4)
    ll2, st2 = ws.loglikelihood(theta[:4])
    for i in range(4):
        d = pl.solve_draw(m, theta[i])
        ss = d["ss"]
        Yd = Y - d["ybar"][m.obs_idx][:, None]
        ref = pl.kalman_likelihood(Yd, ss["Z"], ss["H"], ss["T"], ss["R"], ss["Q"], ss["a0"], ss["P0"], presample=4)
        assert abs(ll2[i] - ref) <= RTOL_LOGLIK * abs(ref)
    ws.set_options(presample=0)


def test_measurement_error_and_custom_estimated_parameter_map(golden, ws_factory):
    """stderr of a measurement error and a shock correlation in the estimated-parameter map
    (set_estimated_parameters!, estimation.jl:517-601, incl. corr -> cov with current variances)."""
    g = golden("fs2000")
    m = models.get_model("fs2000")
    ws = ws_factory("fs2000", g["Y"])
    from dynare_jl_b200 import modelgen as mg
    kinds = [mg.EST_PARAM, mg.EST_PARAM, mg.EST_SD_SHOCK, mg.EST_SD_SHOCK, mg.EST_CORR_SHOCK, mg.EST_SD_MEAS, mg.EST_VAR_MEAS]
    i1 = [0, 4, 0, 1, 0, 0, 1]
    i2 = [0, 0, 0, 1, 1, 0, 1]
    ws.set_estimated_params(kinds, i1, i2)
    m.est = [("param", "alp"), ("param", "rho"), ("stderr", "e_a"), ("stderr", "e_m"),
             ("corr", "e_a", "e_m"), ("stderr", "gp_obs"), ("var", "gy_obs")]
    rng = np.random.default_rng(5)
    th = np.column_stack([rng.uniform(0.3, 0.45, 20), rng.uniform(0.1, 0.9, 20), rng.uniform(0.005, 0.03, 20),
                          rng.uniform(0.002, 0.01, 20), rng.uniform(-0.6, 0.6, 20), rng.uniform(0.0005, 0.004, 20),
                          rng.uniform(1e-7, 1e-5, 20)])
    ll, st = ws.loglikelihood(th)
    for i in range(20):
        ref, rst = pl.loglikelihood(m, th[i], g["Y"])
        assert st[i] == rst == 0
        assert abs(ll[i] - ref) <= RTOL_LOGLIK * abs(ref)


def test_full_size_batch_properties(golden, ws_factory):
    """BASELINE config 2 size (65 536 draws): the batch result does not depend on batch position or
    batch size, and repeated draws give bit-identical values."""
    g = golden("fs2000")
    ws = ws_factory("fs2000", g["Y"])
    n = 65536
    base = g["theta"]
    reps = n // base.shape[0]
    theta = np.tile(base, (reps, 1))
    ll, st = ws.loglikelihood(theta)
    assert np.all(st == 0)
    llr = ll.reshape(reps, base.shape[0])
    assert np.all(llr == llr[0][None, :])                 # bit-identical across positions
    rel = np.abs(llr[0] - g["loglik"]) / np.abs(g["loglik"])
    assert rel.max() <= RTOL_LOGLIK
    perm = np.random.default_rng(0).permutation(n)[:5000]
    ll2, _ = ws.loglikelihood(theta[perm])
    assert np.array_equal(ll2, ll[perm])
    ll3, _ = ws.loglikelihood(theta[:1])                  # N = 1 is the reference's scalar call
    assert ll3[0] == ll[0]
    ll0, st0 = ws.loglikelihood(np.zeros((0, 9)))
    assert ll0.size == 0 and st0.size == 0


def test_pinned_host_buffers(golden, ws_factory):
    from dynare_jl_b200.engine import PinnedArray
    g = golden("fs2000")
    ws = ws_factory("fs2000", g["Y"])
    n = 256
    pth = PinnedArray((n, 9)); pll = PinnedArray((n,)); pst = PinnedArray((n,), np.int32)
    pth.array[:] = g["theta"]
    ll, st = ws.loglikelihood(pth.array, out=pll.array, status=pst.array)
    assert np.allclose(pll.array, g["loglik"], rtol=RTOL_LOGLIK, atol=0) and np.all(pst.array == 0)
    pth.free(); pll.free(); pst.free()


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_solver_variants_agree(name, golden, ws_factory):
    """solver = 1 (one thread per draw, pivoted LU) and solver = 2 (cooperative, Householder QR in registers)
    are two implementations of the same stages; both must meet the oracle tolerances."""
    g = golden(name)
    ws = ws_factory(name, g["Y"])
    res = {}
    for variant in (0, 1, 2):          # 0: verified fast pass (no pivoting) + Householder pass on what it cannot verify
        ws.set_options(solver=variant)
        ll, st = ws.loglikelihood(g["theta"])
        fo = ws.first_order(g["theta"])
        assert np.array_equal(st, g["status"])
        ok = g["status"] == 0
        assert np.all(np.isneginf(ll[~ok]))
        assert np.max(np.abs(ll[ok] - g["loglik"][ok]) / np.abs(g["loglik"][ok])) <= RTOL_LOGLIK
        g1 = np.concatenate([fo["g1_1"], fo["g1_2"]], axis=2)
        assert np.abs(g1[ok] - g["g1"][ok]).max() <= ATOL_POLICY
        res[variant] = (ll[ok], fo["cr_iters"][ok])
    ws.set_options(solver=0)
    # same number of cyclic-reduction iterations up to rounding at the stopping threshold
    assert np.max(np.abs(res[1][1].astype(int) - res[2][1].astype(int))) <= 1
    assert np.mean(res[1][1] != res[2][1]) < 0.05
    assert np.max(np.abs(res[1][0] - res[2][0]) / np.abs(res[1][0])) <= 1e-9


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_fast_pass_hands_over_what_it_cannot_verify(name, golden, ws_factory):
    """The cooperative solver's fast pass eliminates WITHOUT pivoting (equations pre-ordered at the calibration) and verifies
    the residuals of what it returns; everything else goes to the Householder pass.  On widened-prior stress draws -- roughly
    half of them fail Blanchard-Kahn or the steady state -- solver 0 must return the statuses of the Householder pass alone
    (solver 2), agree with it to solver accuracy on the accepted draws, and never redo a draw that failed before the solve."""
    from dynare_jl_b200.engine import model_description
    from dynare_jl_b200.priors import PriorSet
    g = golden(name)
    ws = ws_factory(name, g["Y"])
    pri = PriorSet.from_description(model_description(name)["priors"])
    rng = np.random.default_rng(77)
    n = 20000
    th = pri.sample(rng, n)
    mu = th.mean(0)
    wide = rng.random(n) < 0.5
    th[wide] = mu + 2.5 * (th[wide] - mu)
    out = {}
    for variant in (2, 0):
        ws.set_options(solver=variant)
        r0 = ws.solver_redo_count()
        ll, st = ws.loglikelihood(th)
        fo = ws.first_order(th)
        out[variant] = (ll, st, np.concatenate([fo["g1_1"], fo["g1_2"]], axis=2), ws.solver_redo_count() - r0)
    ws.set_options(solver=0)
    (ll2, st2, g2, redo2), (ll0, st0, g0, redo0) = out[2], out[0]
    assert redo2 == 0
    assert np.array_equal(st0, st2)
    ok = st2 == 0
    assert ok.sum() > n // 10
    # per draw, relative to the size of its policy matrix (scripts/fast_pass_accuracy.py: against the QZ oracle the fast pass is
    # as accurate as the Householder pass, 7.6e-13 against 2.5e-11 at worst on 3 000 such draws of fs2000)
    scale = np.maximum(1.0, np.abs(g2[ok]).reshape(int(ok.sum()), -1).max(1))
    dd = np.abs(g0[ok] - g2[ok]).reshape(int(ok.sum()), -1).max(1) / scale
    assert np.quantile(dd, 0.999) <= 1e-9 and dd.max() <= 1e-6      # the worst draws are ill-conditioned for BOTH solvers
    fin = ok & (np.abs(ll2) < 1e9)                       # beyond that the filter problem itself is ill-conditioned
    rl = np.abs(ll0[fin] - ll2[fin]) / np.abs(ll2[fin])
    assert np.quantile(rl, 0.99) <= 1e-8 and rl.max() <= 1e-3      # the filter amplifies the policy difference of the few worst draws
    # two calls (likelihood, first_order): every draw that entered the solver and did not come out accepted was redone,
    # plus the accepted ones whose residual the fast pass was not satisfied with (far from the calibration point its
    # equation order is no longer a good pivot order: 3.6 % of these widened ls2003 draws, 6 of 458 752 prior draws)
    bk_failures = int(np.isin(st2, (2, 3)).sum())        # decided inside the cyclic reduction
    solver_failures = int((st2 == 4).sum())              # some of these are raised before / after it
    assert 2 * bk_failures <= redo0 <= 2 * (bk_failures + solver_failures) + n // 5


def test_fast_pass_without_a_pivot_order(golden, ws_factory):
    """The fast pass orders the equations by partial pivoting at the CALIBRATION point.  When that point itself has no steady
    state (here: an estimated parameter calibrated to a value outside its domain, which every draw overrides) there is no
    order to use: every draw must go to the Householder pass and come back with the same results as solver = 2."""
    from dynare_jl_b200.engine import model_description
    g = golden("fs2000")
    ws = ws_factory("fs2000", g["Y"])
    d = model_description("fs2000")
    p = np.array(d["param_values"], dtype=float)
    est = [i for t, i in zip(d["est_type"], d["est_i"]) if t == 0]
    p[est[0]] = -5.0                                         # alp < 0: the steady state of the calibration is not finite
    ws.set_calibration(p)
    out = {}
    for variant in (2, 0):
        ws.set_options(solver=variant)
        r0 = ws.solver_redo_count()
        ll, st = ws.loglikelihood(g["theta"])
        out[variant] = (ll, st, ws.solver_redo_count() - r0)
    ws.set_options(solver=0)
    assert np.array_equal(out[0][1], g["status"]) and np.array_equal(out[2][1], g["status"])
    assert np.array_equal(out[0][0], out[2][0])              # the same kernel solved them: identical bits
    assert out[0][2] == int((g["status"] != 1).sum()) and out[2][2] == 0
    ok = g["status"] == 0
    assert np.max(np.abs(out[0][0][ok] - g["loglik"][ok]) / np.abs(g["loglik"][ok])) <= RTOL_LOGLIK


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_prior_predictive_moments_and_irfs(name, golden, ws_factory):
    """run_simulations (priorprediction.jl:64-98): steady state, robustsqrt(diag(variance)) and the 40-period responses of
    irfs! (perturbations.jl:670-694) for a batch of draws against the oracle's per-draw computation."""
    g = golden(name)
    m = models.get_model(name)
    ws = ws_factory(name, g["Y"])
    theta = g["theta"][:64]
    out = ws.prior_predictive(theta, periods=40)
    gs = g["status"][:64]
    assert np.array_equal(out["status"] != 0, (gs != 0) & (gs != pl.ST_F_NOT_PD))      # filter failures do not exist here
    eps = np.finfo(float).eps
    checked = 0
    for i in range(theta.shape[0]):
        if out["status"][i] != 0:
            assert np.all(np.isnan(out["stdev"][i]))
            continue
        d = pl.solve_draw(m, theta[i])
        g11, g12 = d["sol"]["g1_1"], d["sol"]["g1_2"]
        assert np.allclose(out["steady_state"][i], d["ybar"], rtol=1e-12, atol=1e-14)
        sd = np.sqrt(np.diag(d["Sigma_y"]) + eps)
        assert np.allclose(out["stdev"][i], sd, rtol=1e-8, atol=1e-12)
        for q in range(m.nx):
            y = g12[:, q] * np.sqrt(d["sigma_e"][q, q] + eps)
            for p in range(40):
                assert np.allclose(out["irfs"][i, q, p], y, rtol=1e-8, atol=1e-12), (i, q, p)
                y = g11 @ y[m.i_bkwrd_b]
        checked += 1
    assert checked > 30


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_filter_lane_variants_agree(name, golden, ws_factory):
    """The thread-per-draw filter, the 4 / 8 lanes-per-draw shared-tile filter and the 16 / 32 lanes-per-draw cooperative
    filter (small batches, SMC shards) run the same recursion with the same summation order: identical statuses and
    IDENTICAL log-likelihood bits (what keeps a sharded SMC run bit-identical to the single-GPU one); also with missing
    observations and a presample.  Each variant is also held to the golden values."""
    g = golden(name)
    Y = g["Y"].copy()
    Y[0, 3] = np.nan; Y[-1, 10] = np.nan; Y[:, 17] = np.nan
    for obs, presample in ((g["Y"], 0), (Y, 4)):
        ws = ws_factory(name, obs)
        out = {}
        for lanes in (1, 4, 8, 16, 32):
            ws.set_options(filter_lanes=lanes, presample=presample)
            out[lanes] = ws.loglikelihood(g["theta"])
        ws.set_options(filter_lanes=0, presample=0)
        ok = g["status"] == 0
        for lanes in (4, 8, 16, 32):
            assert np.array_equal(out[lanes][1], out[1][1])
            good = out[1][1] == 0
            assert np.array_equal(out[lanes][0][good], out[1][0][good]), (name, lanes, np.abs(out[lanes][0][good] - out[1][0][good]).max())
            assert np.all(np.isneginf(out[lanes][0][~good]))
        if presample == 0:
            for lanes in (1, 4, 8, 16, 32):
                rel = np.abs(out[lanes][0][ok] - g["loglik"][ok]) / np.abs(g["loglik"][ok])
                assert rel.max() <= RTOL_LOGLIK, (name, lanes, rel.max())
