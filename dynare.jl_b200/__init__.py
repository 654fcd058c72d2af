"""dynare.jl_b200 -- B200-native batched first-order DSGE log-likelihood engine.

Host-side mirror (Python, ctypes) of the part of Dynare.jl's estimation interface that sits on
the per-draw likelihood path (reference ``src/estimation/estimation.jl:358-429, 503-702`` and
``src/estimation/smc.jl``).  All arithmetic runs in ``libdynare_b200.so`` (hand-written CUDA
for sm_100a behind the C ABI declared in ``include/dynare_b200.h``).  There is no CPU fallback:
anything that computes raises if the library or a GPU is missing.
"""
import os as _os

PACKAGE_DIR = _os.path.dirname(_os.path.abspath(__file__))
REPO_ROOT = _os.path.dirname(PACKAGE_DIR)

__all__ = ["PACKAGE_DIR", "REPO_ROOT"]
