"""ctypes binding of libdynare_b200.so -- one Python function per symbol of include/dynare_b200.h.

The library is the product; this file only marshals arguments.  Loading fails loudly when the
shared object is missing: there is no Python / CPU fallback for any computation.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import PACKAGE_DIR, REPO_ROOT

LIB_PATH = os.path.join(PACKAGE_DIR, "libdynare_b200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "dynare_b200.h")


class DybError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdynare_b200 error {code}: {msg}")
        self.code = code


class ModelDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n", "nx", "npar", "k", "nf", "s", "ncol", "ns", "ny", "np_default", "ss_doubles", "mid_doubles")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Options(C.Structure):
    _fields_ = [("cr_tol", C.c_double), ("cr_maxit", C.c_int32), ("steady_tol", C.c_double),
                ("lyap_tol", C.c_double), ("lyap_maxit", C.c_int32), ("presample", C.c_int32),
                ("solver", C.c_int32), ("filter_lanes", C.c_int32),
                ("univariate_fallback", C.c_int32)]


class SmcOptions(C.Structure):
    _fields_ = [("npart", C.c_int64), ("nphi", C.c_int32), ("lam", C.c_double), ("c0", C.c_double),
                ("acpt0", C.c_double), ("trgt", C.c_double), ("alp", C.c_double), ("resampling", C.c_int32),
                ("seed", C.c_uint64)]


class ModelDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("n_static", C.c_int32),
                ("n_bkwrd_b", C.c_int32), ("n_fwrd_b", C.c_int32), ("i_static", C.c_void_p),
                ("i_bkwrd_b", C.c_void_p), ("i_fwrd_b", C.c_void_p), ("obs_idx", C.c_void_p)]


COMM_ID_BYTES = 128

_P = C.c_void_p
_D = C.c_void_p      # double* (host or device address)
_I32 = C.c_void_p
_I64 = C.c_void_p

# symbol -> (restype, argtypes); must cover every function declared in the header
SIGNATURES = {
    "dyb_version": (C.c_int, []),
    "dyb_last_error": (C.c_char_p, []),
    "dyb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "dyb_model_count": (C.c_int, []),
    "dyb_model_name": (C.c_char_p, [C.c_int]),
    "dyb_model_get_dims": (C.c_int, [C.c_char_p, C.POINTER(ModelDims)]),
    "dyb_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "dyb_destroy": (C.c_int, [_P]),
    "dyb_get_dims": (C.c_int, [_P, C.POINTER(ModelDims)]),
    "dyb_get_options": (C.c_int, [_P, C.POINTER(Options)]),
    "dyb_set_options": (C.c_int, [_P, C.POINTER(Options)]),
    "dyb_set_calibration": (C.c_int, [_P, _D, _D, _D]),
    "dyb_set_estimated_params": (C.c_int, [_P, C.c_int32, _I32, _I32, _I32]),
    "dyb_set_observations": (C.c_int, [_P, _D, C.c_int32, C.c_int32]),
    "dyb_loglik_batch": (C.c_int, [_P, _D, C.c_int64, _D, _I32]),
    "dyb_loglik_batch_async": (C.c_int, [_P, _D, C.c_int64, _D, _I32]),
    "dyb_loglik_batch_device": (C.c_int, [_P, _D, C.c_int64, _D, _I32]),
    "dyb_sync": (C.c_int, [_P]),
    "dyb_query": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "dyb_kernel_launches": (C.c_int64, [_P]),
    "dyb_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "dyb_first_order_batch": (C.c_int, [_P, _D, C.c_int64, _D, _D, _I32, _I32, _I32]),
    "dyb_set_stream": (C.c_int, [_P, C.c_void_p]),
    "dyb_kernel_timers_reset": (C.c_int, [_P]),
    "dyb_kernel_timers_read": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "dyb_kernel_timers_read_solve_split": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "dyb_solver_redo_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "dyb_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "dyb_fp64_peak_dmma": (C.c_int, [C.c_int, C.c_int32, C.POINTER(C.c_double)]),
    "dyb_state_space_batch": (C.c_int, [_P, _D, C.c_int64, _D, _D, _D, _D, _D, _I32]),
    "dyb_prior_predictive_batch": (C.c_int, [_P, _D, C.c_int64, C.c_int32, _D, _D, _D, _I32]),
    "dyb_kalman_smoother_batch": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, _D, _D, _D, _D, _D,
                                            _D, _D, C.c_int64, _D, _D, _D, _D, _D, _D, _I32]),
    "dyb_kalman_diffuse_smoother_batch": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, _D, _D, _D,
                                                    _D, _D, _D, _D, _D, _D, C.c_double, C.c_int64, _D, _D, _D, _D, _D,
                                                    _D, _I32, _I32]),
    "dyb_shock_covariances_batch": (C.c_int, [_P, _D, C.c_int64, _D, _D]),
    "dyb_smoother_batch": (C.c_int, [_P, _D, C.c_int64, _D, _D, _D, _D, _I32]),
    "dyb_forecast_batch": (C.c_int, [_P, _D, C.c_int64, C.c_int32, _D, _D, _I32]),
    "dyb_kalman_batch": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, _D,
                                   _D, _D, _D, _D, _D, C.c_int64, _D, _I32]),
    "dyb_kalman_batch_fallback": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, _D,
                                            _D, _D, _D, _D, _D, C.c_int64, C.c_int32, _D, _I32]),
    "dyb_smc_reweight": (C.c_int, [C.c_int, _D, _D, C.c_double, C.c_int64, _D,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dyb_smc_resample_multinomial": (C.c_int, [C.c_int, _D, _D, C.c_int64, _I64, C.POINTER(C.c_int64)]),
    "dyb_smc_resample_systematic": (C.c_int, [C.c_int, _D, C.c_double, C.c_int64, _I64, C.POINTER(C.c_int64)]),
    "dyb_smc_moments": (C.c_int, [C.c_int, _D, _D, C.c_int64, C.c_int32, _D, _D]),
    "dyb_smc_covariance_factor": (C.c_int, [C.c_int, _D, C.c_int32, _D, _D, C.POINTER(C.c_int32)]),
    "dyb_set_kalman_variant": (C.c_int, [C.c_int32]),
    "dyb_create_generic": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(_P)]),
    "dyb_first_order_from_jacobian_batch": (C.c_int, [_P, _D, _D, C.c_int32, C.c_int64, _D, _D, _D, _D, _I32, _I32, _I32]),
    "dyb_loglik_from_jacobian_batch": (C.c_int, [_P, _D, _D, _D, C.c_int32, _D, C.c_int32, _I32, C.c_int64, _D, _I32]),
    "dyb_lyapunov_batch": (C.c_int, [C.c_int, C.c_int32, _D, _D, C.c_int64, C.c_double, C.c_int32, _D, _I32, _I32]),
    "dyb_set_priors": (C.c_int, [_P, C.c_int32, _I32, _D, _D]),
    "dyb_logprior_batch": (C.c_int, [_P, _D, C.c_int64, _D]),
    "dyb_logposterior_batch": (C.c_int, [_P, _D, C.c_int64, _D, _D, _I32]),
    "dyb_transform_batch": (C.c_int, [_P, _D, C.c_int64, C.c_int32, _D, _D]),
    "dyb_rng_fill": (C.c_int, [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_int32,
                               C.c_int32, _D]),
    "dyb_smc_default_options": (C.c_int, [C.POINTER(SmcOptions)]),
    "dyb_smc_create": (C.c_int, [_P, C.POINTER(SmcOptions), _D, C.POINTER(_P)]),
    "dyb_smc_destroy": (C.c_int, [_P]),
    "dyb_smc_local_range": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dyb_smc_transport": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dyb_smc_init": (C.c_int, [_P, _D, _I32]),
    "dyb_smc_init_from_prior": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_int64)]),
    "dyb_prior_draw": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_int32, _D]),
    "dyb_smc_run": (C.c_int, [_P, C.c_int32]),
    "dyb_smc_stages_done": (C.c_int, [_P]),
    "dyb_smc_set_profiling": (C.c_int, [_P, C.c_int32]),
    "dyb_smc_get_timing": (C.c_int, [_P, C.POINTER(C.c_int32), _D]),
    "dyb_smc_get_particles": (C.c_int, [_P, _D, _D, _D, _D]),
    "dyb_smc_get_stats": (C.c_int, [_P, _D, _D, _D, C.POINTER(C.c_double)]),
    "dyb_rwmh_run": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_uint32, C.c_int32, _D, C.c_uint64,
                               _D, _D, _I64, _D, _D]),
    "dyb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "dyb_comm_init": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_void_p]),
    "dyb_comm_destroy": (C.c_int, [_P]),
    "dyb_comm_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dyb_comm_allreduce_sum": (C.c_int, [_P, C.c_void_p, C.c_int64, C.c_int32]),
    "dyb_host_alloc": (C.c_int, [C.POINTER(_P), C.c_int64]),
    "dyb_host_free": (C.c_int, [_P]),
}


def header_symbols():
    """Names of every function declared in include/dynare_b200.h."""
    with open(HEADER_PATH) as f:
        txt = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dyb_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DybError(rc, load().dyb_last_error().decode(errors="replace"))


def ptr(a):
    """Address of a numpy array / torch tensor / raw int address (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):                 # torch tensor (device or host)
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)
