"""The front end that reads the REFERENCE'S preprocessor output (model/json/modfile.json + model/julia/*.jl, what
Dynare.jl's `parser` and `load_model_functions` read: src/DynareParser.jl:10-15, 169-218; src/dynare_functions.jl:10-47)
against (i) the hand-written oracle models and (ii) the descriptions shipped for the .mod route.  The samples under
tests/golden/preprocessor_<model>/ are written in the preprocessor's layout and syntax by
scripts/write_preprocessor_sample.py (the preprocessor binary is a JLL artifact that is not in the build image)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import models

MODELS = ["fs2000", "hs_nk"]


def _spec(name):
    from dynare_jl_b200 import preprocessor
    return preprocessor.spec_from_preprocessor(name, os.path.join(GOLDEN, f"preprocessor_{name}"))


@pytest.mark.parametrize("name", MODELS)
def test_bookkeeping_matches_the_mod_route(name):
    from dynare_jl_b200.engine import model_description
    S, d = _spec(name), model_description(name)
    assert S.endo == d["endo"] and S.exo == d["exo"] and S.params == d["params"] and S.varobs == d["varobs"]
    for key in ("i_bkwrd_b", "i_fwrd_b", "i_static", "i_dyn", "state_ids", "obs_idx_state", "lagged_state_ids",
                "est_type", "est_i", "est_j"):
        assert list(getattr(S, key)) == list(d[key]), key
    assert np.allclose(S.param_values, d["param_values"], rtol=0, atol=0, equal_nan=True)
    assert np.array_equal(S.sigma_e0, np.array(d["sigma_e0"])) and np.array_equal(S.sigma_m0, np.array(d["sigma_m0"]))
    pri = S.describe()["priors"]
    assert [(p["shape"], p["mean"], p["stdev"]) for p in pri] == [(p["shape"], p["mean"], p["stdev"]) for p in d["priors"]]
    assert np.allclose(S.theta_init(), d["theta_init"], rtol=0, atol=0)


@pytest.mark.parametrize("name", MODELS)
def test_model_functions_match_the_oracle(name, golden):
    """Steady state, static residuals and the compressed Jacobian translated from the generated Julia code, against the
    hand-written residual functions differentiated by complex step (oracle/models.py)."""
    from dynare_jl_b200.modelgen import HostModel
    from oracle import pipeline as pl
    S = _spec(name)
    hm = HostModel(S)
    m = models.get_model(name)
    g = golden(name)
    ok = np.flatnonzero(g["status"] == 0)[:6]
    for i in ok:
        params, _, _ = pl.set_estimated_parameters(m, g["theta"][i])
        yb = hm.steady_state(params)
        assert np.allclose(yb, m.steady_state(params), rtol=1e-13, atol=1e-15)
        assert np.max(np.abs(hm.static_residuals(params, yb))) < 1e-10
        J = hm.jacobian(params, yb)
        Jo = m.compressed_jacobian(yb, params)
        assert J.shape == Jo.shape and np.allclose(J, Jo, rtol=1e-11, atol=1e-13)


def test_generated_header_is_the_same_model(tmp_path):
    """The CUDA traits emitted from the preprocessor route declare the same sizes and index maps as the shipped header."""
    from dynare_jl_b200 import modelgen, PACKAGE_DIR
    S = _spec("fs2000")
    cuh = modelgen.emit_cuda(S)
    shipped = open(os.path.join(PACKAGE_DIR, "csrc", "generated", "model_fs2000.cuh")).read()
    for line in shipped.splitlines():
        if "static constexpr int" in line and "=" in line and "a[" not in line and "NNZ" not in line and "FMA" not in line:
            assert line.strip() in cuh, line
    for fn in ("i_bkwrd_b", "i_fwrd_b", "state_ids", "obs_idx_state", "lagged_state_ids", "est_type0", "est_i0"):
        want = [l.strip() for l in shipped.splitlines() if f"int {fn}(int i)" in l][0]
        assert want in cuh, fn


@pytest.mark.gpu
def test_preprocessor_route_gives_the_same_likelihood(golden, tmp_path):
    """fs2000 compiled at run time from the preprocessor-format sample, against the shipped (.mod route) fs2000."""
    from dynare_jl_b200 import engine, preprocessor
    lib = preprocessor.build_model_library_from_preprocessor("fs2000_pp", os.path.join(GOLDEN, "preprocessor_fs2000"),
                                                             str(tmp_path))
    engine.load_model_library(lib)
    g = golden("fs2000")
    a = engine.BatchSSWs("fs2000", g["Y"]); b = engine.BatchSSWs("fs2000_pp", g["Y"])
    lla, sta = a.loglikelihood(g["theta"]); llb, stb = b.loglikelihood(g["theta"])
    assert np.array_equal(sta, stb)
    ok = sta == 0
    assert np.max(np.abs(lla[ok] - llb[ok]) / np.abs(lla[ok])) <= 1e-11
    fa, fb = a.first_order(g["theta"][:32]), b.first_order(g["theta"][:32])
    assert np.allclose(fa["g1_1"], fb["g1_1"], rtol=0, atol=1e-11) and np.allclose(fa["g1_2"], fb["g1_2"], rtol=0, atol=1e-11)
    a.close(); b.close()
