"""The symbolic model set-up agrees with the hand-written oracle models (complex-step Jacobian)."""
import json
import os

import numpy as np
import pytest

from conftest import REFERENCE, ROOT
from oracle import models

MODS = {"fs2000": "test/models/fs2000/fs2000.mod", "ls2003": "test/models/ls2003/ls2003.mod",
        "hs_nk": "test/models/estimation/test1.mod"}


def _spec(name):
    from dynare_jl_b200 import modelgen
    return modelgen.spec_from_file(name, os.path.join(REFERENCE, MODS[name]))


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_committed_description_matches_oracle_bookkeeping(name):
    with open(os.path.join(ROOT, "dynare.jl_b200", "models", f"{name}.json")) as f:
        d = json.load(f)
    m = models.get_model(name)
    assert d["endo"] == m.endo and d["exo"] == m.exo and d["params"] == m.params
    assert d["i_bkwrd_b"] == m.i_bkwrd_b and d["i_fwrd_b"] == m.i_fwrd_b and d["i_static"] == m.i_static
    assert d["state_ids"] == m.state_ids and d["obs_idx_state"] == m.obs_idx_state
    assert d["lagged_state_ids"] == m.lagged_state_ids
    assert np.allclose(d["param_values"], m.params0)
    assert np.allclose(d["sigma_e0"], m.sigma_e0)
    assert [p["shape"] for p in d["priors"]] == [p[0] for p in m.priors]


def test_fs2000_structure():
    """Sizes of SURVEY.md section 8: n=16 (14 + 2 lead auxiliaries), k=4, s=7, m=9, ns=6, ny=2, ncol=29."""
    with open(os.path.join(ROOT, "dynare.jl_b200", "models", "fs2000.json")) as f:
        d = json.load(f)
    assert (d["n"], d["n_orig"], d["nx"], d["k"], d["s"], d["m"], d["ns"], d["ny"], d["ncol"], d["np_est"]) == \
        (16, 14, 2, 4, 7, 9, 6, 2, 29, 9)
    assert d["endo"][14:] == ["AUX_ENDO_LEAD_c_1", "AUX_ENDO_LEAD_P_1"]


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_symbolic_jacobian_matches_complex_step(name, have_reference):
    if not have_reference:
        pytest.skip("reference model files not available on this box")
    from dynare_jl_b200.modelgen import HostModel
    spec = _spec(name)
    hm = HostModel(spec)
    m = models.get_model(name)
    rng = np.random.default_rng(0)
    for trial in range(4):
        p = m.params0 * (1 + 0.02 * rng.standard_normal(m.params0.size)) if trial else m.params0
        yb = m.steady_state(p)
        assert np.allclose(hm.steady_state(p), yb, rtol=1e-12, atol=1e-13)
        assert np.linalg.norm(hm.static_residuals(p, yb)) < 1e-10
        Jo = m.compressed_jacobian(yb, p)
        Jg = hm.jacobian(p, yb)
        assert np.abs(Jo - Jg).max() <= 1e-12 * max(1.0, np.abs(Jo).max())


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_generated_sources_are_current(name, have_reference):
    if not have_reference:
        pytest.skip("reference model files not available on this box")
    from dynare_jl_b200 import modelgen
    spec = _spec(name)
    with open(os.path.join(ROOT, "dynare.jl_b200", "csrc", "generated", f"model_{name}.cuh")) as f:
        assert f.read() == modelgen.emit_cuda(spec)
    with open(os.path.join(ROOT, "oracle", "c", f"model_{name}.c")) as f:
        assert f.read() == modelgen.emit_c(spec)


def test_mod_reader_handles_prior_statements_and_commented_estimation(have_reference):
    if not have_reference:
        pytest.skip("reference model files not available on this box")
    from dynare_jl_b200.modfile import parse_mod
    with open(os.path.join(REFERENCE, MODS["ls2003"])) as f:
        mf = parse_mod(f.read())
    assert mf.linear and len(mf.estimated_params) == 17
    assert mf.estimation_options["first_obs"] == "8" and mf.estimation_options["nobs"] == "79"
    assert mf.estimated_params[12].kind == "stderr" and mf.estimated_params[12].shape == "inv_gamma"


def test_aux_variables_for_long_leads_and_lags():
    from dynare_jl_b200 import modelgen
    from dynare_jl_b200.modfile import parse_mod
    txt = """var x z; varexo e; parameters a; a = 0.5;
    model; x = a*x(-2) + 0.1*z(+3) + e; z = 0.3*z(-1) + e; end;"""
    spec = modelgen.build_spec("toy", parse_mod(txt))
    assert spec.endo == ["x", "z", "AUX_ENDO_LAG_x_1", "AUX_ENDO_LEAD_z_1", "AUX_ENDO_LEAD_z_2"]
    assert spec.n == 5 and spec.k == 3 and spec.nf == 3


def test_sizes_only_header():
    """modelgen.emit_cuda_jacobian_fed: traits of a model whose steady state and Jacobian stay on the host -- sizes, the
    reference's index sets (src/dynare_containers.jl:110-131) and the dense hand-off patterns, no model function."""
    from dynare_jl_b200 import modelgen
    from oracle import models
    m = models.get_model("ls2003")
    txt = modelgen.emit_cuda_jacobian_fed("jf_test", m.n, m.nx, m.i_static, m.i_bkwrd_b, m.i_fwrd_b, m.obs_idx)
    s, k, nf, mm = len(m.i_static), m.k, m.nf, m.n - len(m.i_static)
    ncol = k + m.n + nf + m.nx
    assert "static constexpr bool JAC_FED = true;" in txt and "struct Model_jf_test" in txt
    assert f"static constexpr int RED_NNZ = {mm * (ncol - s)};" in txt and f"static constexpr int TOP_NNZ = {s * ncol};" in txt
    assert f"static constexpr int NS = {m.ns};" in txt and f"static constexpr int NCOL = {ncol};" in txt
    # the working-matrix permutation is a permutation, static columns first
    import re
    jp = [int(x) for x in re.search(r"jperm\(int i\) \{ constexpr int a\[\d+\] = \{([^}]*)\}", txt).group(1).split(",")]
    assert sorted(jp) == list(range(ncol))
    assert sorted(jp[k + v] for v in m.i_static) == list(range(s))
    with pytest.raises(ValueError):
        modelgen.emit_cuda_jacobian_fed("bad", 4, 1, [0, 1], [1], [2], [0])        # a static variable cannot have a lag
