"""The generated CUDA model headers are also valid host C++: compile them with g++ and check the
device model function (steady state, residuals, Jacobian) against the hand-written oracle models."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import models

SHIM = r"""
#include "generated/model_%(name)s.cuh"
using M = dyb::Model_%(name)s;
extern "C" {
int dims(int* out) { out[0] = M::N; out[1] = M::NCOL; out[2] = M::NPAR; return 0; }
void steady(const double* p, double* y) { M::steady_state(p, y); }
double resid(const double* p, const double* y) { return M::static_resid_ss(p, y); }
void jac(const double* p, const double* y, double* J) { M::jacobian(p, y, J); }
int jperm(int c) { return M::jperm(c); }
}
"""


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_generated_model_function_on_host(name, tmp_path):
    src = tmp_path / "shim.cpp"
    src.write_text(SHIM % dict(name=name))
    so = tmp_path / "shim.so"
    inc = os.path.join(ROOT, "dynare.jl_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-x", "c++", "-fPIC", "-shared", f"-I{inc}", str(src), "-o", str(so)],
                   check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.resid.restype = C.c_double
    d = (C.c_int * 3)()
    lib.dims(d)
    m = models.get_model(name)
    n, ncol = d[0], d[1]
    assert n == m.n
    rng = np.random.default_rng(1)
    for trial in range(5):
        p = m.params0 * (1 + 0.003 * rng.standard_normal(m.params0.size)) if trial else m.params0.copy()
        y = np.zeros(n)
        lib.steady(p.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
        yo = m.steady_state(p)
        assert np.allclose(y, yo, rtol=1e-12, atol=1e-14)
        r2 = lib.resid(p.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
        assert np.sqrt(r2) < 1e-10
        J = np.zeros(n * ncol)
        lib.jac(p.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), J.ctypes.data_as(C.c_void_p))
        J = J.reshape(ncol, n).T
        Jo = m.compressed_jacobian(yo, p)
        assert np.abs(J - Jo).max() <= 1e-12 * max(1.0, np.abs(Jo).max())
    perm = sorted(lib.jperm(c) for c in range(ncol))
    assert perm == list(range(ncol))
