"""The .mod reader and the symbolic model set-up on the reference's own test models (read from /root/reference, so
these run in the build container only): every stochastic model of `test/models` that has a steady state must be read,
its steady state must satisfy the static model, the Blanchard-Kahn conditions must hold at the calibration, and the
two solvers of the oracle (ordered QZ = the reference's default, cyclic reduction = the device algorithm) must return
the same policy matrix.  Device code is generated for each (not compiled here: tests/test_user_model.py does that)."""
import os

import numpy as np
import pytest

from conftest import REFERENCE
from oracle import pipeline as pl

MODELS = ["example1/example1.mod", "example2/example2.mod", "example3/example3.mod", "example5/example5_est_a.mod",
          "example1pf/example1pf.mod", "herbst_schorfheide/nk.mod", "herbst_schorfheide/nk_est.mod",
          "herbst_schorfheide/nk_sol.mod", "estimation/fs2000.mod", "estimation/test1.mod", "fs2000/fs2000.mod",
          "ls2003/ls2003.mod"]


@pytest.mark.parametrize("rel", MODELS)
def test_reference_model_is_read_and_solved(rel, have_reference):
    if not have_reference:
        pytest.skip("reference model files not available on this box")
    path = os.path.join(REFERENCE, "test", "models", rel)
    if not os.path.exists(path):
        pytest.skip(f"{rel} not in this reference checkout")
    from dynare_jl_b200 import modelgen
    spec = modelgen.spec_from_file("m", path)
    hm = modelgen.HostModel(spec)
    p = spec.param_values
    assert np.all(np.isfinite(p)), "every parameter is calibrated"
    y = hm.steady_state(p)
    assert np.all(np.isfinite(y)) and np.abs(hm.static_residuals(p, y)).max() <= 1e-10
    n, k, nf = spec.n, spec.k, spec.nf
    Jc = hm.jacobian(p, y)                                   # [lags (k) | current (n) | leads (nf) | shocks]
    C = np.zeros((n, n)); A = np.zeros((n, n))
    C[:, spec.i_bkwrd_b] = Jc[:, :k]
    B = Jc[:, k:k + n]
    A[:, spec.i_fwrd_b] = Jc[:, k + n:k + n + nf]
    G = pl.solve_qz(A, B, C)[0]                              # raises DrawFailure when Blanchard-Kahn fails
    assert np.abs(A @ G @ G + B @ G + C).max() <= 1e-10 * max(1.0, np.abs(B).max())
    assert np.max(np.abs(np.linalg.eigvals(G))) < 1 + 1e-6
    Gcr = pl.solve_cr(C, B, A, tol=1e-13)[0]
    assert np.abs(Gcr - G).max() <= 1e-9 * max(1.0, np.abs(G).max())
    src = modelgen.emit_cuda(spec)
    assert f"struct Model_{spec.name}" in src and "static_resid_ss" in src


def test_models_outside_the_path_are_refused_loudly(have_reference):
    """A Ramsey model (planner's first-order conditions are derived by the preprocessor) is not a square system here."""
    if not have_reference:
        pytest.skip("reference model files not available on this box")
    path = os.path.join(REFERENCE, "test", "models", "nk_ramsey", "nk_ramsey_2.mod")
    if not os.path.exists(path):
        pytest.skip("not in this reference checkout")
    from dynare_jl_b200 import modelgen
    with pytest.raises(ValueError):
        modelgen.spec_from_file("m", path)
