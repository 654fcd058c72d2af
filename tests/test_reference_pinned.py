"""Parity pinned to the REFERENCE ITSELF.

scripts/make_ref_input.py writes tests/golden/ref_input_<model>.json (parameter vectors incl. stress draws, observations);
scripts/julia_golden.jl, run once on any box with Julia + Dynare.jl, evaluates the real `loglikelihood`
(src/estimation/estimation.jl:603-702) and `first_order_solver!` results on them and writes
tests/golden/ref_<model>.json.  The tests below compare the CPU oracle (not gpu) and the CUDA engine (gpu) with those
values and SKIP while the file is absent -- until then the repo's parity claim is "green against the oracle, unpinned by
the reference" (DESIGN.md section 5)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import models, pipeline as pl

MODELS = ["fs2000", "ls2003", "hs_nk"]
RTOL_LOGLIK = 1e-8       # north-star tolerance (prior draws)
ATOL_POLICY = 1e-10


def _stress_tol(ref_value):
    """Stress draws sit where the Lyapunov / filter problem itself is ill-conditioned (a root at 1 - 1e-6, innovation
    covariances within rounding of singular: |loglik| up to 1e13): 1e-6 relative, 1e-3 once |loglik| > 1e9."""
    return 1e-3 if abs(ref_value) > 1e9 else 1e-6


def _load(kind, name):
    p = os.path.join(GOLDEN, f"{kind}_{name}.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


@pytest.mark.parametrize("name", MODELS)
def test_reference_inputs_are_current(name, golden):
    """The inputs handed to the Julia script are the golden draws / observations (regenerate with make_ref_input.py)."""
    inp = _load("ref_input", name)
    assert inp is not None, "run scripts/make_ref_input.py"
    g = golden(name)
    th = np.array(inp["theta"], dtype=float)
    n0 = inp["n_prior_draws"]
    assert np.array_equal(th[:n0], g["theta"][:n0])
    Y = np.array([[np.nan if v is None else v for v in row] for row in inp["Y"]], dtype=float)
    assert np.array_equal(np.isnan(Y), np.isnan(g["Y"])) and np.allclose(np.nan_to_num(Y), np.nan_to_num(g["Y"]), rtol=0, atol=0)
    from dynare_jl_b200.engine import model_description
    assert inp["estimated"] == list(model_description(name)["est_names"])
    assert th.shape[0] == n0 + len(inp["stress"]) and len(inp["stress"]) >= 5


def _check_against_reference(name, ref, inp, ll, st, g1):
    """ll / st / g1 (n x k+nx per draw or None) of one implementation against the reference's record."""
    assert ref["estimated"] == inp["estimated"] or len(ref["estimated"]) == len(inp["estimated"])
    worst_ll = worst_g1 = 0.0
    n0 = inp["n_prior_draws"]
    for i, err in enumerate(ref["error"]):
        if err != "" or ref["loglik"][i] is None:
            assert st[i] != 0 and ll[i] == -np.inf, (name, i, "the reference rejects this draw", err, st[i], ll[i])
            continue
        r = ref["loglik"][i]
        assert st[i] == 0, (name, i, "the reference scores this draw", r, st[i])
        if i >= n0:
            assert abs(ll[i] - r) <= _stress_tol(r) * abs(r), (name, i, ll[i], r)
            continue
        worst_ll = max(worst_ll, abs(ll[i] - r) / abs(r))
        if g1 is not None and ref["g1_1"][i] is not None:
            rg = np.hstack([np.array(ref["g1_1"][i]), np.array(ref["g1_2"][i])])
            worst_g1 = max(worst_g1, float(np.max(np.abs(g1[i] - rg))))
    assert worst_ll <= RTOL_LOGLIK, (name, worst_ll)
    assert worst_g1 <= ATOL_POLICY, (name, worst_g1)


@pytest.mark.parametrize("name", MODELS)
def test_oracle_against_reference(name):
    ref, inp = _load("ref", name), _load("ref_input", name)
    if ref is None:
        pytest.skip(f"tests/golden/ref_{name}.json absent: run scripts/julia_golden.jl on a box with Julia + Dynare.jl")
    m = models.get_model(name)
    Y = np.array([[np.nan if v is None else v for v in row] for row in inp["Y"]], dtype=float)
    th = np.array(inp["theta"], dtype=float)
    ll = np.empty(len(th)); st = np.empty(len(th), dtype=int); g1 = []
    for i, t in enumerate(th):
        ll[i], st[i] = pl.loglikelihood(m, t, Y)
        if st[i] == 0:
            d = pl.solve_draw(m, t)
            g1.append(np.hstack([d["sol"]["g1_1"], d["sol"]["g1_2"]]))
        else:
            g1.append(None)
    _check_against_reference(name, ref, inp, ll, st, g1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MODELS)
def test_engine_against_reference(name):
    ref, inp = _load("ref", name), _load("ref_input", name)
    if ref is None:
        pytest.skip(f"tests/golden/ref_{name}.json absent: run scripts/julia_golden.jl on a box with Julia + Dynare.jl")
    from dynare_jl_b200.engine import BatchSSWs
    Y = np.array([[np.nan if v is None else v for v in row] for row in inp["Y"]], dtype=float)
    th = np.array(inp["theta"], dtype=float)
    ws = BatchSSWs(name, Y)
    ll, st = ws.loglikelihood(th)
    fo = ws.first_order(th)
    g1 = [np.hstack([fo["g1_1"][i], fo["g1_2"][i]]) if fo["status"][i] == 0 else None for i in range(len(th))]
    ws.close()
    _check_against_reference(name, ref, inp, ll, st, g1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MODELS)
def test_engine_matches_oracle_on_reference_inputs(name):
    """Runs with or without the reference's record: on the draws handed to the Julia script (stress draws included) the
    engine and the oracle accept the same draws and agree on them."""
    inp = _load("ref_input", name)
    from dynare_jl_b200.engine import BatchSSWs
    m = models.get_model(name)
    Y = np.array([[np.nan if v is None else v for v in row] for row in inp["Y"]], dtype=float)
    th = np.array(inp["theta"], dtype=float)
    ws = BatchSSWs(name, Y)
    ll, st = ws.loglikelihood(th)
    ws.close()
    n0 = inp["n_prior_draws"]
    for i, t in enumerate(th):
        r, rs = pl.loglikelihood(m, t, Y)
        assert (st[i] == 0) == (rs == 0), (name, i, st[i], rs)
        if rs == 0:
            tol = RTOL_LOGLIK if i < n0 else _stress_tol(r)
            assert abs(ll[i] - r) <= tol * abs(r), (name, i, ll[i], r)
