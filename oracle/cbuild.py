"""TEST INFRASTRUCTURE: builds the C restatement of the oracle (oracle/c) into oracle/_build/liboracle.so
and binds it with ctypes.  The reference itself is Julia and cannot be compiled here, so there is no
oracle/_ref; this port is the CPU baseline (`cpu_baseline.kind = "port"`)."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liboracle.so")


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(SRC, "*.c")))
    deps = srcs + glob.glob(os.path.join(SRC, "*.h"))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(map(os.path.getmtime, deps)):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
           "-o", LIB, *srcs, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stderr}")
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.oracle_loglik_batch.restype = C.c_int
        lib.oracle_loglik_batch.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def loglik_batch(model, theta, Y, presample=0, nthreads=0, want_g1=False, est=None, cr_tol=0.0, dims=None):
    """theta (N, np); Y ny x nobs.  Returns (loglik, status, threads used[, g1 (N, n, k+nx)])."""
    lib = load()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, np_ = theta.shape
    Yf = np.asfortranarray(np.asarray(Y, dtype=np.float64))
    ll = np.empty(n); st = np.empty(n, dtype=np.int32)
    g1 = None
    if want_g1:
        nn, kk = dims
        g1 = np.zeros((n, kk, nn))
    if est is None:
        et = ei = ej = None; npass = 0
    else:
        et, ei, ej = (np.ascontiguousarray(a, dtype=np.int32) for a in est); npass = np_
    p = lambda a: None if a is None else a.ctypes.data
    used = lib.oracle_loglik_batch(model.encode(), p(theta), n, npass, p(et), p(ei), p(ej), p(Yf), Yf.shape[0],
                                   Yf.shape[1], presample, cr_tol, p(ll), p(st), p(g1), nthreads)
    if used < 0:
        raise RuntimeError("oracle_loglik_batch failed")
    if want_g1:
        return ll, st, used, g1.transpose(0, 2, 1)
    return ll, st, used
