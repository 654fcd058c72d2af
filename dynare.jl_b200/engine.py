"""Host-side mirror of the reference's per-draw likelihood interface, batched.

Reference interface                                   -> here
  SSWs(context, observations, varobs)                 -> BatchSSWs(model, observations)
     (src/estimation/estimation.jl:358-429)
  loglikelihood(theta, context, observations, ssws)   -> BatchSSWs.loglikelihood(Theta)  (N draws at once)
     (src/estimation/estimation.jl:603-702)
  DSGELogPosteriorDensity (callable, -Inf on failure) -> DSGELogPosteriorDensity(ssws, priors)(Theta)
     (src/estimation/estimation.jl:439-453, 503-515)
  LRE.first_order_solver! results g1_1 / g1_2         -> BatchSSWs.first_order(Theta)
     (src/perturbations.jl:663-664)

``Theta`` is an ``(N, np)`` C-contiguous float64 array, i.e. the ``np x N`` column-major matrix of the
C ABI.  Everything is computed by libdynare_b200.so on the GPU; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional

import numpy as np

from . import PACKAGE_DIR
from . import _lib as L
from .priors import PriorSet

STATUS_NAMES = {0: "ok", 1: "steady_state", 2: "indeterminate", 3: "unstable", 4: "solver",
                5: "nonstationary", 6: "F_not_pd"}


def available_models():
    lib = L.load()
    return [lib.dyb_model_name(i).decode() for i in range(lib.dyb_model_count())]


_MODEL_DIRS = [os.path.join(PACKAGE_DIR, "models")]
_MODEL_LIBS = []


def model_description(name: str) -> dict:
    """Plain-data description written by the model set-up (names, index sets, priors, defaults)."""
    for d in _MODEL_DIRS:
        p = os.path.join(d, f"{name}.json")
        if os.path.exists(p):
            with open(p) as f:
                return json.load(f)
    raise FileNotFoundError(f"no description for model {name!r}")


def load_model_library(path: str):
    """Load a model compiled by ``build.build_model_library``: its static initialiser registers the model with
    libdynare_b200.so (which must be loaded first, with global symbols); the .json beside it becomes visible to
    ``model_description``."""
    L.load()
    _MODEL_LIBS.append(C.CDLL(os.path.abspath(path), mode=C.RTLD_GLOBAL))
    d = os.path.dirname(os.path.abspath(path))
    if d not in _MODEL_DIRS:
        _MODEL_DIRS.append(d)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class PinnedArray:
    """A numpy view of page-locked host memory obtained from ``dyb_host_alloc``."""

    def __init__(self, shape, dtype=np.float64):
        lib = L.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        L.check(lib.dyb_host_alloc(C.byref(p), max(self.nbytes, 8)))
        self._ptr = p
        buf = (C.c_char * max(self.nbytes, 8)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            L.load().dyb_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchSSWs:
    """State-space workspace of a compiled model on one GPU (the batched ``SSWs``)."""

    def __init__(self, model: str, observations: Optional[np.ndarray] = None, device: int = 0):
        self.lib = L.load()
        self.model = model
        self.device = device
        h = C.c_void_p()
        L.check(self.lib.dyb_create(model.encode(), device, C.byref(h)))
        self._h = h
        d = L.ModelDims()
        L.check(self.lib.dyb_get_dims(self._h, C.byref(d)))
        self.dims = d.asdict()
        self.np = self.dims["np_default"]
        self.nobs = 0
        if observations is not None:
            self.set_observations(observations)

    # -- configuration --------------------------------------------------------------------------
    def set_observations(self, Y):
        """Y: ny x nobs (rows in varobs order), NaN = missing."""
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim != 2 or Y.shape[0] != self.dims["ny"]:
            raise ValueError(f"observations must be ny x nobs with ny = {self.dims['ny']}")
        Yf = np.asfortranarray(Y)
        L.check(self.lib.dyb_set_observations(self._h, L.ptr(Yf), Y.shape[0], Y.shape[1]))
        self.nobs = Y.shape[1]

    def set_calibration(self, params=None, sigma_e=None, sigma_m=None):
        d = self.dims
        p = None if params is None else _f64(params, (d["npar"],))
        se = None if sigma_e is None else np.asfortranarray(_f64(sigma_e, (d["nx"], d["nx"])))
        sm = None if sigma_m is None else np.asfortranarray(_f64(sigma_m, (d["ny"], d["ny"])))
        L.check(self.lib.dyb_set_calibration(self._h, L.ptr(p), L.ptr(se), L.ptr(sm)))

    def set_estimated_params(self, kinds, index1, index2):
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        i = np.ascontiguousarray(index1, dtype=np.int32)
        j = np.ascontiguousarray(index2, dtype=np.int32)
        if not (k.shape == i.shape == j.shape and k.ndim == 1):
            raise ValueError("kinds / index1 / index2 must be 1-d arrays of equal length")
        L.check(self.lib.dyb_set_estimated_params(self._h, k.size, L.ptr(k), L.ptr(i), L.ptr(j)))
        self.np = int(k.size)

    def options(self) -> L.Options:
        o = L.Options()
        L.check(self.lib.dyb_get_options(self._h, C.byref(o)))
        return o

    def set_options(self, **kw):
        o = self.options()
        for k, v in kw.items():
            if not hasattr(o, k):
                raise KeyError(k)
            setattr(o, k, v)
        L.check(self.lib.dyb_set_options(self._h, C.byref(o)))

    # -- the hot path -------------------------------------------------------------------------------
    def _theta(self, Theta):
        Theta = np.asarray(Theta, dtype=np.float64)
        if Theta.ndim == 1:
            Theta = Theta[None, :]
        if Theta.ndim != 2 or Theta.shape[1] != self.np:
            raise ValueError(f"Theta must be (N, {self.np})")
        return np.ascontiguousarray(Theta)

    def loglikelihood(self, Theta, out=None, status=None):
        """Log-likelihood of N draws: returns (loglik[N], status[N]); loglik = -inf where status != 0."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        ll = np.empty(n, dtype=np.float64) if out is None else out
        st = np.empty(n, dtype=np.int32) if status is None else status
        L.check(self.lib.dyb_loglik_batch(self._h, L.ptr(Theta), n, L.ptr(ll), L.ptr(st)))
        return ll, st

    def loglikelihood_async(self, Theta, out, status):
        """Enqueue H2D + kernels + D2H and return; ``Theta`` / ``out`` / ``status`` must be pinned (``PinnedArray``)
        or device memory and stay untouched until ``sync()``."""
        n = Theta.shape[0]
        L.check(self.lib.dyb_loglik_batch_async(self._h, L.ptr(Theta), n, L.ptr(out), L.ptr(status)))

    def loglikelihood_device(self, d_theta: int, n: int, d_loglik: int, d_status: int):
        """Device-pointer variant (addresses as ints); asynchronous, call ``sync()``."""
        L.check(self.lib.dyb_loglik_batch_device(self._h, L.ptr(d_theta), n, L.ptr(d_loglik), L.ptr(d_status)))

    def sync(self):
        L.check(self.lib.dyb_sync(self._h))

    def done(self) -> bool:
        """Non-blocking: has everything enqueued on this handle finished?"""
        d = C.c_int32()
        L.check(self.lib.dyb_query(self._h, C.byref(d)))
        return bool(d.value)

    def set_stream(self, cuda_stream: Optional[int]):
        """Run on a caller-owned CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        L.check(self.lib.dyb_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def timers_reset(self):
        L.check(self.lib.dyb_kernel_timers_reset(self._h))

    def timers_read(self):
        """(number of batches, total solve-kernel ms, total filter-kernel ms) since ``timers_reset``."""
        n, a, b = C.c_int32(), C.c_float(), C.c_float()
        L.check(self.lib.dyb_kernel_timers_read(self._h, C.byref(n), C.byref(a), C.byref(b)))
        return n.value, a.value, b.value

    def timers_read_solve_split(self):
        """(model set-up ms, cyclic-reduction ms, assembly ms) totals of the armed batches; call before timers_read."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        L.check(self.lib.dyb_kernel_timers_read_solve_split(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def solver_redo_count(self) -> int:
        """Draws the solver's fast pass could not verify and its Householder pass solved again (since creation)."""
        v = C.c_int64()
        L.check(self.lib.dyb_solver_redo_count(self._h, C.byref(v)))
        return v.value

    def kernel_launches(self) -> int:
        return int(self.lib.dyb_kernel_launches(self._h))

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        L.check(self.lib.dyb_last_kernel_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- priors / posterior (device) ------------------------------------------------------------------
    def set_priors(self, priors: PriorSet):
        """Upload the prior of every estimated parameter (hyper-parameters as the distributions take them)."""
        if len(priors) != self.np:
            raise ValueError("one prior per estimated parameter is required")
        code, a, b = priors.abi_arrays()
        L.check(self.lib.dyb_set_priors(self._h, self.np, L.ptr(code), L.ptr(a), L.ptr(b)))
        self.priors = priors

    def prior_draw(self, n, seed=0, first=0, round=0):
        """(n, np) draws from the priors on the device (prior_draw!, estimation.jl:329-335); draw i is particle first + i of
        the counter-based stream, so a shard of a larger cloud is reproducible on its own."""
        out = np.empty((int(n), self.np))
        L.check(self.lib.dyb_prior_draw(self._h, int(n), int(first), int(seed), int(round), L.ptr(out)))
        return out

    def logprior(self, Theta):
        Theta = self._theta(Theta)
        out = np.empty(Theta.shape[0])
        L.check(self.lib.dyb_logprior_batch(self._h, L.ptr(Theta), Theta.shape[0], L.ptr(out)))
        return out

    def logposterior(self, Theta):
        """(logpost, loglik, status): logprior + loglik, -inf for failed draws (estimation.jl:503-515)."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        lp = np.empty(n); ll = np.empty(n); st = np.empty(n, dtype=np.int32)
        L.check(self.lib.dyb_logposterior_batch(self._h, L.ptr(Theta), n, L.ptr(lp), L.ptr(ll), L.ptr(st)))
        return lp, ll, st

    def transform(self, X, inverse=False):
        """inverse=False: R^np -> parameters, returns (theta, logjac); inverse=True: parameters -> R^np."""
        X = self._theta(X)
        n = X.shape[0]
        out = np.empty_like(X); lj = np.empty(n)
        L.check(self.lib.dyb_transform_batch(self._h, L.ptr(X), n, 1 if inverse else 0, L.ptr(out), L.ptr(lj)))
        return out if inverse else (out, lj)

    # -- stage-level ----------------------------------------------------------------------------------
    def first_order(self, Theta):
        """Returns dict(g1_1 (N,n,k), g1_2 (N,n,nx), ybar (N,n), cr_iters (N,), status (N,))."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        d = self.dims
        g1 = np.zeros((n, d["k"] + d["nx"], d["n"]))          # per draw column-major n x (k+nx)
        yb = np.zeros((n, d["n"]))
        it = np.zeros(n, dtype=np.int32)
        li = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        L.check(self.lib.dyb_first_order_batch(self._h, L.ptr(Theta), n, L.ptr(g1), L.ptr(yb), L.ptr(it), L.ptr(li), L.ptr(st)))
        g1 = g1.transpose(0, 2, 1)
        return dict(g1_1=g1[:, :, :d["k"]], g1_2=g1[:, :, d["k"]:], ybar=yb, cr_iters=it, lyap_iters=li, status=st)

    def prior_predictive(self, Theta, periods=40):
        """run_simulations (priorprediction.jl:64-98) for N draws at once: dict(steady_state (N, n), stdev (N, n),
        irfs (N, nx, periods, n), status); failed draws carry NaN."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        d = self.dims
        yb = np.full((n, d["n"]), np.nan); sd = np.empty((n, d["n"])); irf = np.empty((n, d["nx"], periods, d["n"]))
        st = np.empty(n, dtype=np.int32)
        L.check(self.lib.dyb_prior_predictive_batch(self._h, L.ptr(Theta), n, periods, L.ptr(yb), L.ptr(sd), L.ptr(irf), L.ptr(st)))
        return dict(steady_state=yb, stdev=sd, irfs=irf, status=st)

    def smoother(self, Theta):
        """calibsmoother! for N draws: dict(alphah (N, nobs, n) smoothed variables in levels, etah (N, nobs, nx),
        epsilonh (N, nobs, ny), loglik, status)."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        d = self.dims
        al = np.full((n, self.nobs, d["n"]), np.nan); et = np.full((n, self.nobs, d["nx"]), np.nan)
        ep = np.full((n, self.nobs, d["ny"]), np.nan); ll = np.empty(n); st = np.empty(n, dtype=np.int32)
        L.check(self.lib.dyb_smoother_batch(self._h, L.ptr(Theta), n, L.ptr(al), L.ptr(et), L.ptr(ep), L.ptr(ll), L.ptr(st)))
        return dict(alphah=al, etah=et, epsilonh=ep, loglik=ll, status=st)

    def shock_covariances(self, Theta):
        """Sigma_e (N, nx, nx) and Sigma_m (N, ny, ny) of every draw (set_estimated_parameters!, estimation.jl:517-601)."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        d = self.dims
        se = np.empty((n, d["nx"], d["nx"])); sm = np.empty((n, d["ny"], d["ny"]))
        L.check(self.lib.dyb_shock_covariances_batch(self._h, L.ptr(Theta), n, L.ptr(se), L.ptr(sm)))
        return se.transpose(0, 2, 1), sm.transpose(0, 2, 1)

    def smoother_nonstationary(self, Theta, Y, tol=1e-8):
        """The non-stationary branch of calibsmoother! (src/filters/kalman/kalman.jl:163-223) for N draws of a model with
        unit roots: first-order solution on the device (unit roots are accepted among the stable roots), every
        endogenous variable a state (T[:, i_bkwrd_b] = g1_1, R = g1_2; kalman.jl:78-108), ordered Schur split and
        Lyapunov equation on the stable block (host LAPACK as in the reference + device doubling), exact-diffuse filter
        and smoother on the device.  Y (ny, nobs) in levels.  Returns dict(alphah (N, nobs, n) in levels, etah, epsilonh,
        loglik (diffuse), n_diffuse, n_unit_roots, status); H must be diagonal (measurement-error variances)."""
        Theta = self._theta(Theta)
        N = Theta.shape[0]
        d = self.dims
        desc = model_description(self.model)
        fo = self.first_order(Theta)
        ok = (fo["status"] == 0) | (fo["status"] == 5)          # 5: Lyapunov doubling did not converge = unit roots
        n, nx, ny = d["n"], d["nx"], d["ny"]
        Y = np.asarray(Y, dtype=np.float64)
        nobs = Y.shape[1]
        out = dict(alphah=np.full((N, nobs, n), np.nan), etah=np.full((N, nobs, nx), np.nan),
                   epsilonh=np.full((N, nobs, ny), np.nan), loglik=np.full(N, -np.inf),
                   n_diffuse=np.zeros(N, dtype=np.int32), n_unit_roots=np.zeros(N, dtype=np.int32), status=fo["status"].copy())
        idx = np.flatnonzero(ok)
        if idx.size == 0:
            return out
        se, sm = self.shock_covariances(Theta[idx])
        T = np.zeros((idx.size, n, n))
        T[:, :, desc["i_bkwrd_b"]] = fo["g1_1"][idx]
        R = np.ascontiguousarray(fo["g1_2"][idx])
        Pinf, Pstar, ku = diffuse_initial_covariances(T, R, se, device=self.device)
        ybo = fo["ybar"][idx][:, desc["obs_idx"]]
        Hd = np.ascontiguousarray(np.diagonal(sm, axis1=1, axis2=2))
        r = kalman_diffuse_smoother_batch(Y, desc["obs_idx"], T, R, se, Pinf, Pstar, Hdiag=Hd, ybar_obs=ybo, tol=tol,
                                          device=self.device)
        out["alphah"][idx] = r["alphah"] + fo["ybar"][idx][:, None, :]
        out["etah"][idx] = r["etah"]; out["epsilonh"][idx] = r["epsilonh"]
        out["loglik"][idx] = r["loglik"]; out["n_diffuse"][idx] = r["n_diffuse"]; out["n_unit_roots"][idx] = ku
        out["status"][idx] = r["status"]
        return out

    def forecast(self, Theta, y0, periods):
        """Unconditional forecast (forecast.jl:54-159) for N draws from the levels y0 (N, n): (N, periods+1, n), status."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        y0 = _f64(y0, (n, self.dims["n"]))
        Y = np.empty((n, periods + 1, self.dims["n"])); st = np.empty(n, dtype=np.int32)
        L.check(self.lib.dyb_forecast_batch(self._h, L.ptr(Theta), n, periods, L.ptr(y0), L.ptr(Y), L.ptr(st)))
        return Y, st

    def state_space(self, Theta):
        """Returns dict(T, RQR, P0 (N,ns,ns), ybar_obs (N,ny), H (N,ny,ny), status)."""
        Theta = self._theta(Theta)
        n = Theta.shape[0]
        ns, ny = self.dims["ns"], self.dims["ny"]
        T = np.zeros((n, ns, ns)); Q = np.zeros((n, ns, ns)); P = np.zeros((n, ns, ns))
        yb = np.zeros((n, ny)); H = np.zeros((n, ny, ny)); st = np.zeros(n, dtype=np.int32)
        L.check(self.lib.dyb_state_space_batch(self._h, L.ptr(Theta), n, L.ptr(T), L.ptr(Q), L.ptr(P),
                                               L.ptr(yb), L.ptr(H), L.ptr(st)))
        tr = lambda a: a.transpose(0, 2, 1)                  # column-major per draw -> numpy row-major
        return dict(T=tr(T), RQR=tr(Q), P0=tr(P), ybar_obs=yb, H=tr(H), status=st)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.dyb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GenericSSWs:
    """State-space workspace of a model WITHOUT a generated device function: the host evaluates steady state and
    Jacobian per draw (as the reference does, src/perturbations.jl:533-561) and this class runs everything after
    that on the GPU.  Index sets are 0-based, as ``Model`` holds them (src/dynare_containers.jl:110-131)."""

    def __init__(self, n, nx, i_static, i_bkwrd_b, i_fwrd_b, obs_idx, observations=None, device=0, compile=False,
                 build_dir=None):
        """``compile=True``: the register-resident kernels of the generated-model path are compiled (nvcc, once per set of
        sizes and index sets, cached in ``build_dir``) for THIS model's sizes -- a "sizes-only" model without a model
        function (build.build_sizes_library) -- and used instead of the run-time-size kernels: the same calls, an order of
        magnitude faster for small models (dynamic block <= 12).  Falls back to the run-time-size kernels when the model
        does not fit them."""
        self.lib = L.load()
        self._idx = [np.ascontiguousarray(a, dtype=np.int32) for a in (i_static, i_bkwrd_b, i_fwrd_b, obs_idx)]
        self.compiled = False
        h = C.c_void_p()
        if compile and n - len(self._idx[0]) <= 12:
            from . import build as _build
            bdir = build_dir or os.path.join(L.PACKAGE_DIR, "models", "_sizes")      # in-tree: travels with the built library
            lib, name = _build.build_sizes_library(n, nx, *[a.tolist() for a in self._idx], bdir)
            load_model_library(lib)
            L.check(self.lib.dyb_create(name.encode(), device, C.byref(h)))
            self.compiled = True
        else:
            d = L.ModelDesc(n, nx, len(self._idx[3]), len(self._idx[0]), len(self._idx[1]), len(self._idx[2]),
                            *[a.ctypes.data if a.size else None for a in self._idx])
            L.check(self.lib.dyb_create_generic(C.byref(d), device, C.byref(h)))
        self._h = h
        self.nobs = 0
        dm = L.ModelDims()
        L.check(self.lib.dyb_get_dims(self._h, C.byref(dm)))
        self.dims = dm.asdict()
        if observations is not None:
            self.set_observations(observations)

    set_observations = BatchSSWs.set_observations
    options = BatchSSWs.options
    set_options = BatchSSWs.set_options
    close = BatchSSWs.close
    __del__ = BatchSSWs.__del__
    timers_reset = BatchSSWs.timers_reset
    timers_read = BatchSSWs.timers_read
    timers_read_solve_split = BatchSSWs.timers_read_solve_split

    def _jac(self, jac):
        """(N, n, ncol) numpy -> per-draw column-major."""
        jac = np.asarray(jac, dtype=np.float64)
        if jac.ndim != 3 or jac.shape[1:] != (self.dims["n"], self.dims["ncol"]):
            raise ValueError(f"jac must be (N, {self.dims['n']}, {self.dims['ncol']})")
        return np.ascontiguousarray(jac.transpose(0, 2, 1))

    def _cov(self, S, k):
        if S is None:
            return None, 0
        S = np.asarray(S, dtype=np.float64)
        if S.shape == (k, k):
            return np.ascontiguousarray(S.T), 0
        return np.ascontiguousarray(S.transpose(0, 2, 1)), 1

    def first_order(self, jac, sigma_e):
        """dict(g1_1, g1_2, T, RQR, P0, cr_iters, lyap_iters, status) from compressed Jacobians (N, n, ncol)."""
        J = self._jac(jac)
        n = J.shape[0]
        d = self.dims
        Se, per = self._cov(sigma_e, d["nx"])
        g1 = np.zeros((n, d["k"] + d["nx"], d["n"])); ns = d["ns"]
        T = np.zeros((n, ns, ns)); Q = np.zeros((n, ns, ns)); P = np.zeros((n, ns, ns))
        it = np.zeros(n, dtype=np.int32); li = np.zeros(n, dtype=np.int32); st = np.zeros(n, dtype=np.int32)
        L.check(self.lib.dyb_first_order_from_jacobian_batch(self._h, L.ptr(J), L.ptr(Se), per, n, L.ptr(g1), L.ptr(T),
                                                             L.ptr(Q), L.ptr(P), L.ptr(it), L.ptr(li), L.ptr(st)))
        g1 = g1.transpose(0, 2, 1)
        tr = lambda a: a.transpose(0, 2, 1)
        return dict(g1_1=g1[:, :, :d["k"]], g1_2=g1[:, :, d["k"]:], T=tr(T), RQR=tr(Q), P0=tr(P), cr_iters=it,
                    lyap_iters=li, status=st)

    def loglikelihood(self, jac, ybar, sigma_e, sigma_m=None, status_in=None):
        """(loglik[N], status[N]) from host-evaluated Jacobians and steady states (N, n)."""
        J = self._jac(jac)
        n = J.shape[0]
        d = self.dims
        yb = None if ybar is None else _f64(ybar, (n, d["n"]))
        Se, pe = self._cov(sigma_e, d["nx"])
        Sm, pm = self._cov(sigma_m, d["ny"])
        si = None if status_in is None else np.ascontiguousarray(status_in, dtype=np.int32)
        ll = np.empty(n); st = np.empty(n, dtype=np.int32)
        L.check(self.lib.dyb_loglik_from_jacobian_batch(self._h, L.ptr(J), L.ptr(yb), L.ptr(Se), pe, L.ptr(Sm), pm,
                                                        L.ptr(si), n, L.ptr(ll), L.ptr(st)))
        return ll, st


def lyapunov_batch(A, B, tol=1e-16, maxit=60, device=0):
    """Batched doubling solve of X = A X A' + B: A, B (N, k, k).  Returns (X, iters, status)."""
    A = np.asarray(A, dtype=np.float64); B = np.asarray(B, dtype=np.float64)
    n, k, _ = A.shape
    cm = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
    Ac, Bc = cm(A), cm(B)                    # keep the column-major copies alive across the call
    X = np.empty((n, k, k)); it = np.empty(n, dtype=np.int32); st = np.empty(n, dtype=np.int32)
    L.check(L.load().dyb_lyapunov_batch(device, k, L.ptr(Ac), L.ptr(Bc), n, tol, maxit, L.ptr(X), L.ptr(it), L.ptr(st)))
    return X.transpose(0, 2, 1), it, st


class DSGELogPosteriorDensity:
    """Batched analogue of the reference callable (estimation.jl:439-453, 503-515):
    ``logprior(theta) + loglikelihood(theta)``, ``-inf`` for any failed draw."""

    def __init__(self, ssws: BatchSSWs, priors: PriorSet):
        ssws.set_priors(priors)
        self.ssws = ssws
        self.priors = priors

    def dimension(self):
        return self.ssws.np

    def __call__(self, Theta):
        return self.ssws.logposterior(Theta)[0]


# ---------------------------------------------------------------------------------------------------
# model-free entry points
# ---------------------------------------------------------------------------------------------------
def fp64_peak(device=0) -> float:
    """Measured fp64 FMA peak of the device in TFLOP/s (register-resident DFMA loop)."""
    t = C.c_double()
    L.check(L.load().dyb_fp64_peak(device, C.byref(t)))
    return t.value


def fp64_peak_dmma(shape=0, device=0) -> float:
    """Measured fp64 tensor-core peak (TFLOP/s): shape 0 = mma.m8n8k4, 1 = mma.m16n8k16."""
    t = C.c_double()
    L.check(L.load().dyb_fp64_peak_dmma(device, shape, C.byref(t)))
    return t.value


def set_kalman_variant(variant: int):
    """0 auto, 1 scalar dense filter kernel, 2 tensor-core (DMMA) kernel, 3 DMMA without bulk copies, 4 four lanes per draw
    (ns <= 16, ny <= 8), 5 DMMA with the unfolded ns <= 88 kernel."""
    L.check(L.load().dyb_set_kalman_variant(variant))


def kalman_batch(Y, z_idx, T, RQR, P0, a0=None, H=None, presample=0, device=0, univariate_fallback=0):
    """Generic batched Kalman log-likelihood.  T, RQR, P0: (N, ns, ns) numpy (row-major per draw);
    Y: ny x nobs; z_idx: state observed by each row.  Returns (loglik[N], status[N]).
    ``univariate_fallback``: 1 = periods whose innovation covariance fails its Cholesky factorisation are processed one
    observation at a time (KalmanFilterTools' behaviour), 2 = every period univariate (diagnostic)."""
    lib = L.load()
    T = np.asarray(T, dtype=np.float64)
    n, ns, _ = T.shape
    Y = np.asarray(Y, dtype=np.float64)
    ny, nobs = Y.shape
    cm = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).transpose(0, 2, 1))
    Tc, Qc, Pc = cm(T), cm(RQR), cm(P0)
    Hc = None if H is None else cm(H)
    a0c = None if a0 is None else _f64(a0, (n, ns))
    z = np.ascontiguousarray(z_idx, dtype=np.int32)
    Yf = np.asfortranarray(Y)
    ll = np.empty(n); st = np.empty(n, dtype=np.int32)
    if univariate_fallback:
        L.check(lib.dyb_kalman_batch_fallback(device, ns, ny, nobs, presample, L.ptr(z), L.ptr(Yf), L.ptr(Tc), L.ptr(Qc),
                                              L.ptr(Pc), L.ptr(a0c), L.ptr(Hc), n, int(univariate_fallback), L.ptr(ll), L.ptr(st)))
    else:
        L.check(lib.dyb_kalman_batch(device, ns, ny, nobs, presample, L.ptr(z), L.ptr(Yf), L.ptr(Tc), L.ptr(Qc),
                                     L.ptr(Pc), L.ptr(a0c), L.ptr(Hc), n, L.ptr(ll), L.ptr(st)))
    return ll, st


def rng_fill(seed, first, n, step, stream=0, kind=0, count=1, device=0):
    """The samplers' Philox stream: kind 0 -> (n, count) uniforms of `stream`, kind 1 -> (n, count) normals."""
    out = np.empty((n, count))
    L.check(L.load().dyb_rng_fill(device, seed, first, n, step, stream, kind, count, L.ptr(out)))
    return out


def kalman_smoother_batch(Y, z_idx, T, R, Q, P0, a0=None, H=None, device=0):
    """Dense batched smoother: T, P0 (N, ns, ns), R (N, ns, nx), Q (N, nx, nx) numpy row-major per draw.
    Returns dict(alphah (N, nobs, ns), etah (N, nobs, nx), epsilonh (N, nobs, ny), att, a_pred, loglik, status)."""
    lib = L.load()
    T = np.asarray(T, dtype=np.float64)
    n, ns, _ = T.shape
    Y = np.asarray(Y, dtype=np.float64)
    ny, nobs = Y.shape
    nx = np.asarray(R).shape[2]
    cm = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).transpose(0, 2, 1))
    Tc, Rc, Qc, Pc = cm(T), cm(R), cm(Q), cm(P0)
    Hc = None if H is None else cm(H)
    a0c = None if a0 is None else _f64(a0, (n, ns))
    z = np.ascontiguousarray(z_idx, dtype=np.int32)
    Yf = np.asfortranarray(Y)
    al = np.empty((n, nobs, ns)); et = np.empty((n, nobs, nx)); ep = np.empty((n, nobs, ny)); at = np.empty((n, nobs, ns))
    ap = np.empty((n, nobs + 1, ns)); ll = np.empty(n); st = np.empty(n, dtype=np.int32)
    L.check(lib.dyb_kalman_smoother_batch(device, ns, ny, nx, nobs, L.ptr(z), L.ptr(Yf), L.ptr(Tc), L.ptr(Rc), L.ptr(Qc),
                                          L.ptr(Pc), L.ptr(a0c), L.ptr(Hc), n, L.ptr(al), L.ptr(et), L.ptr(ep), L.ptr(at),
                                          L.ptr(ap), L.ptr(ll), L.ptr(st)))
    return dict(alphah=al, etah=et, epsilonh=ep, att=at, a_pred=ap, loglik=ll, status=st)


def diffuse_initial_covariances(T, R, Q, criterium=1 - 1e-6, device=0):
    """Host-side preparation of the diffuse smoother, mirroring src/filters/kalman/kalman.jl:163-195 for each draw:
    ordered real Schur form of T (LAPACK gees, moduli > criterium first -- host LAPACK in the reference too), Lyapunov
    solve on the stable block (on the device), results rotated back to the original basis.
    T (N, ns, ns), R (N, ns, nx), Q (N, nx, nx).  Returns (Pinf0, Pstar0 (N, ns, ns), k (N,) unit-root counts)."""
    import scipy.linalg as sla
    T = np.asarray(T, dtype=np.float64); R = np.asarray(R, dtype=np.float64); Q = np.asarray(Q, dtype=np.float64)
    n, ns, _ = T.shape
    Pinf = np.zeros((n, ns, ns)); Pstar = np.zeros((n, ns, ns)); ks = np.zeros(n, dtype=np.int32)
    groups = {}
    for d in range(n):
        S, Zs, k = sla.schur(T[d], output="real", sort=lambda re, im: np.hypot(re, im) > criterium)
        ks[d] = k
        Pinf[d] = Zs[:, :k] @ Zs[:, :k].T
        if ns > k:
            tR = Zs[:, k:].T @ R[d]
            groups.setdefault(int(k), []).append((d, S[k:, k:], tR @ Q[d] @ tR.T, Zs[:, k:]))
    for k, items in groups.items():
        X, _, st = lyapunov_batch(np.stack([it[1] for it in items]), np.stack([it[2] for it in items]), device=device)
        for (d, _, _, Z2), x, s in zip(items, X, st):
            if s != 0:
                raise L.DybError(5, f"Lyapunov solve on the stable block failed for draw {d} (status {int(s)})")
            Pstar[d] = Z2 @ (0.5 * (x + x.T)) @ Z2.T
    return Pinf, Pstar, ks


def kalman_diffuse_smoother_batch(Y, z_idx, T, R, Q, Pinf0, Pstar0, a0=None, Hdiag=None, ybar_obs=None, tol=1e-8,
                                  device=0):
    """Exact-diffuse batched filter + smoother (dyb_kalman_diffuse_smoother_batch): per-draw arrays as in
    kalman_smoother_batch, Hdiag (N, ny) or None.  Returns the same dict plus n_diffuse (N,)."""
    lib = L.load()
    T = np.asarray(T, dtype=np.float64)
    n, ns, _ = T.shape
    Y = np.asarray(Y, dtype=np.float64)
    ny, nobs = Y.shape
    nx = np.asarray(R).shape[2]
    cm = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).transpose(0, 2, 1))
    Tc, Rc, Qc, Pic, Psc = cm(T), cm(R), cm(Q), cm(Pinf0), cm(Pstar0)
    Hc = None if Hdiag is None else _f64(Hdiag, (n, ny))
    ybc = None if ybar_obs is None else _f64(ybar_obs, (n, ny))
    a0c = None if a0 is None else _f64(a0, (n, ns))
    z = np.ascontiguousarray(z_idx, dtype=np.int32)
    Yf = np.asfortranarray(Y)
    al = np.empty((n, nobs, ns)); et = np.empty((n, nobs, nx)); ep = np.empty((n, nobs, ny)); at = np.empty((n, nobs, ns))
    ap = np.empty((n, nobs + 1, ns)); ll = np.empty(n); st = np.empty(n, dtype=np.int32); nd = np.empty(n, dtype=np.int32)
    L.check(lib.dyb_kalman_diffuse_smoother_batch(device, ns, ny, nx, nobs, L.ptr(z), L.ptr(Yf), L.ptr(Tc), L.ptr(Rc),
                                                  L.ptr(Qc), L.ptr(Pic), L.ptr(Psc), L.ptr(a0c), L.ptr(Hc), L.ptr(ybc), float(tol), n,
                                                  L.ptr(al), L.ptr(et), L.ptr(ep), L.ptr(at), L.ptr(ap), L.ptr(ll),
                                                  L.ptr(nd), L.ptr(st)))
    return dict(alphah=al, etah=et, epsilonh=ep, att=at, a_pred=ap, loglik=ll, n_diffuse=nd, status=st)


def smc_reweight(w, loglh, dphi, device=0):
    lib = L.load()
    w = _f64(w); loglh = _f64(loglh)
    out = np.empty_like(w)
    z, e = C.c_double(), C.c_double()
    L.check(lib.dyb_smc_reweight(device, L.ptr(w), L.ptr(loglh), float(dphi), w.size, L.ptr(out), C.byref(z), C.byref(e)))
    return out, z.value, e.value


def smc_resample_multinomial(w, u, device=0):
    lib = L.load()
    w = _f64(w); u = _f64(u)
    idx = np.empty(w.size, dtype=np.int64)
    nc = C.c_int64()
    L.check(lib.dyb_smc_resample_multinomial(device, L.ptr(w), L.ptr(u), w.size, L.ptr(idx), C.byref(nc)))
    return idx, nc.value


def smc_resample_systematic(w, u0, device=0):
    lib = L.load()
    w = _f64(w)
    idx = np.empty(w.size, dtype=np.int64)
    nc = C.c_int64()
    L.check(lib.dyb_smc_resample_systematic(device, L.ptr(w), float(u0), w.size, L.ptr(idx), C.byref(nc)))
    return idx, nc.value


def smc_moments(para, w, device=0):
    lib = L.load()
    para = _f64(para); w = _f64(w)
    n, np_ = para.shape
    mu = np.empty(np_); R = np.empty((np_, np_))
    L.check(lib.dyb_smc_moments(device, L.ptr(para), L.ptr(w), n, np_, L.ptr(mu), L.ptr(R)))
    return mu, R.T.copy()      # library writes column-major


def smc_covariance_factor(R, device=0):
    """`prox!(RPSD, IndPSD(), R)`, `cholesky(RPSD).L` of the SMC mutation step (src/estimation/smc.jl:161-165) on the device.
    Returns (RPSD, L, projected): projected = 0 when R was positive definite (RPSD = R), 1 when it had to be projected on the
    positive semi-definite cone first, 2 when even the projection cannot be factorised (the reference's cholesky throws)."""
    lib = L.load()
    R = np.asfortranarray(np.asarray(R, dtype=np.float64))
    np_ = R.shape[0]
    Rp = np.empty((np_, np_), order="F"); Lc = np.empty((np_, np_), order="F")
    flag = C.c_int32()
    L.check(lib.dyb_smc_covariance_factor(device, L.ptr(R), np_, L.ptr(Rp), L.ptr(Lc), C.byref(flag)))
    return Rp, Lc, flag.value
