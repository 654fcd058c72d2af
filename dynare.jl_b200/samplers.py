"""Host-side mirror of the reference's samplers, driving the device-resident implementations.

Reference interface                                              -> here
  smc(pdraw!, logprior, loglikelihood, names, np)                -> SMC(ssws, priors, ...).run() / smc(...)
     (src/estimation/smc.jl:64-313; Tune :13-62)
  run_mcmc(posterior_density, initial_values, proposal_covariance,
           param_names, mcmc_replic, mcmc_chains)                -> rwmh(ssws, priors, initial_values, covariance, ...)
     (src/estimation/estimation.jl:912-989)
  several GPUs (one process per GPU)                             -> Communicator (NCCL, bootstrapped over any
                                                                    torch.distributed process group or by hand)

Particles / chains live on the GPU for the whole run; the host only draws the initial particles from the prior
(the reference's ``pdraw!`` closure) and reads the results.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .engine import BatchSSWs
from .priors import PriorSet

SMC_STAT_NAMES = ("zhat", "ESS", "resampled", "c", "acpt", "clamped")


class Communicator:
    """NCCL communicator attached to a BatchSSWs handle (one per process / GPU)."""

    def __init__(self, ssws: BatchSSWs, rank: int, nranks: int, unique_id: bytes):
        if len(unique_id) != L.COMM_ID_BYTES:
            raise ValueError("unique_id must be 128 bytes (dyb_comm_unique_id)")
        self.ssws, self.rank, self.nranks = ssws, rank, nranks
        buf = C.create_string_buffer(unique_id, L.COMM_ID_BYTES)
        L.check(ssws.lib.dyb_comm_init(ssws._h, rank, nranks, buf))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(L.COMM_ID_BYTES)
        L.check(L.load().dyb_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, ssws: BatchSSWs, group=None):
        """Bootstrap over an initialised torch.distributed group (any backend): rank 0's id is broadcast."""
        import torch.distributed as dist
        rank, nranks = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(ssws, rank, nranks, box[0])

    def close(self):
        L.check(self.ssws.lib.dyb_comm_destroy(self.ssws._h))


def shard_range(n_total: int, rank: int, nranks: int):
    """Contiguous block partition used for particles and chains: (first, n_local)."""
    if n_total % nranks:
        raise ValueError("the batch must divide evenly over the ranks")
    n = n_total // nranks
    return rank * n, n


def tempering_schedule(nphi: int, lam: float) -> np.ndarray:
    """phi = (collect(0:nphi-1) / (nphi-1)) .^ lam   (smc.jl:48)."""
    return (np.arange(nphi, dtype=np.float64) / (nphi - 1)) ** lam


class SMC:
    """The reference's ``smc`` with ``Tune`` defaults (npart 1000, nphi 400, lam 2, c 0.5, acpt/trgt 0.25, alp 0.9)."""

    def __init__(self, ssws: BatchSSWs, priors: PriorSet, npart=1000, nphi=400, lam=2.0, c=0.5, acpt=0.25, trgt=0.25,
                 alp=0.9, resampling="multinomial", seed=0, phi: Optional[np.ndarray] = None):
        self.ssws, self.priors = ssws, priors
        ssws.set_priors(priors)
        o = L.SmcOptions()
        L.check(ssws.lib.dyb_smc_default_options(C.byref(o)))
        o.npart, o.nphi, o.lam, o.c0, o.acpt0, o.trgt, o.alp = npart, nphi, lam, c, acpt, trgt, alp
        o.resampling = {"multinomial": 0, "systematic": 1}[resampling]
        o.seed = seed
        self._seed = seed
        self.phi = tempering_schedule(nphi, lam) if phi is None else np.ascontiguousarray(phi, dtype=np.float64)
        if self.phi.shape != (nphi,):
            raise ValueError("phi must have nphi entries")
        s = C.c_void_p()
        L.check(ssws.lib.dyb_smc_create(ssws._h, C.byref(o), L.ptr(self.phi), C.byref(s)))
        self._s = s
        f, n = C.c_int64(), C.c_int64()
        L.check(ssws.lib.dyb_smc_local_range(self._s, C.byref(f), C.byref(n)))
        self.first, self.n_local = f.value, n.value
        self.npart, self.nphi, self.np = npart, nphi, ssws.np
        self.prior_draws = 0

    def initialise(self, rng: Optional[np.random.Generator] = None, para: Optional[np.ndarray] = None, max_rounds=1000,
                   seed: Optional[int] = None):
        """Draw the initial particles from the prior until every one has a finite likelihood (smc.jl:84-101).
        Without ``rng`` and ``para`` the draws and the rejection loop run on the device (dyb_smc_init_from_prior; Philox
        stream keyed by ``seed``, default the sampler's seed); with ``rng`` they are numpy draws on the host; ``para`` are
        caller-supplied particles (rejected ones are replaced from ``rng``)."""
        n = self.n_local
        if rng is None and para is None:
            nd = C.c_int64()
            L.check(self.ssws.lib.dyb_smc_init_from_prior(self._s, int(self._seed if seed is None else seed), int(max_rounds),
                                                          C.byref(nd)))
            self.prior_draws += nd.value
            return None
        if rng is None:
            rng = np.random.default_rng(self._seed if seed is None else seed)
        if para is None:
            para = self.priors.sample(rng, n)
            self.prior_draws += n
        para = np.ascontiguousarray(para, dtype=np.float64)
        st = np.empty(n, dtype=np.int32)
        for _ in range(max_rounds):
            L.check(self.ssws.lib.dyb_smc_init(self._s, L.ptr(para), L.ptr(st)))
            bad = np.flatnonzero(st != 0)
            if bad.size == 0:
                return para
            para[bad] = self.priors.sample(rng, bad.size)
            self.prior_draws += bad.size
        raise RuntimeError("could not draw initial particles with a finite likelihood")

    def run(self, n_stages: Optional[int] = None):
        """Enqueue the next ``n_stages`` tempering stages (all remaining ones by default); asynchronous."""
        L.check(self.ssws.lib.dyb_smc_run(self._s, self.nphi if n_stages is None else n_stages))

    def set_profiling(self, on=True):
        L.check(self.ssws.lib.dyb_smc_set_profiling(self._s, 1 if on else 0))

    def timing(self):
        """Per-segment device time (ms, totals over the profiled stages) from CUDA events."""
        ms = np.zeros(6); n = C.c_int32()
        L.check(self.ssws.lib.dyb_smc_get_timing(self._s, C.byref(n), L.ptr(ms)))
        names = ("correction", "selection", "moments", "proposal", "likelihood", "accept")
        return dict(stages=n.value, **dict(zip(names, ms.tolist())))

    def transport(self):
        """dict(peer_stores, cuda_graph): how the stages exchange particles and whether they replay a CUDA graph."""
        a, b = C.c_int32(), C.c_int32()
        L.check(self.ssws.lib.dyb_smc_transport(self._s, C.byref(a), C.byref(b)))
        return dict(peer_stores=bool(a.value), cuda_graph=bool(b.value))

    def stages_done(self) -> int:
        return self.ssws.lib.dyb_smc_stages_done(self._s)

    def particles(self):
        """Local slices: dict(para (n_local, np), w, loglh, logpost)."""
        n = self.n_local
        para = np.empty((n, self.np)); w = np.empty(n); ll = np.empty(n); lp = np.empty(n)
        L.check(self.ssws.lib.dyb_smc_get_particles(self._s, L.ptr(para), L.ptr(w), L.ptr(ll), L.ptr(lp)))
        return dict(para=para, w=w, loglh=ll, logpost=lp)

    def stats(self):
        """dict(stats (nphi, 6), mu, R, logmdd) -- the csim / ESSsim / acptsim / rsmpsim / zhat records of smc.jl."""
        st = np.zeros((self.nphi, len(SMC_STAT_NAMES))); mu = np.empty(self.np); R = np.empty((self.np, self.np))
        lm = C.c_double()
        L.check(self.ssws.lib.dyb_smc_get_stats(self._s, L.ptr(st), L.ptr(mu), L.ptr(R), C.byref(lm)))
        return dict(stats=st, mu=mu, R=R.T.copy(), logmdd=lm.value, **{k: st[:, i] for i, k in enumerate(SMC_STAT_NAMES)})

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            self.ssws.lib.dyb_smc_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def smc(ssws: BatchSSWs, priors: PriorSet, rng: Optional[np.random.Generator] = None, **tune):
    """Run the whole recursion; returns (particles dict, stats dict)."""
    s = SMC(ssws, priors, **tune)
    s.initialise(rng)             # rng None: prior draws and the rejection loop on the device
    s.run()
    out = s.particles(), s.stats()
    s.close()
    return out


def rwmh(ssws: BatchSSWs, priors: PriorSet, initial_values, covariance, mcmc_replic=1000, mcmc_chains=1,
         transformed_parameters=True, seed=0, first_chain=0, first_step=0, keep_history=False):
    """Random-walk Metropolis on ``mcmc_chains`` chains at once (run_mcmc, estimation.jl:964-989).

    ``initial_values``: (np,) or (mcmc_chains, np) in the sampling space (R^np when ``transformed_parameters``);
    ``covariance``: proposal covariance in that space (symmetrised as the reference does).
    Returns dict(y, logp, accepted[, history (n_iter, chains, np), history_logp])."""
    ssws.set_priors(priors)
    npar = ssws.np
    y = np.array(initial_values, dtype=np.float64, order="C")        # a copy: the call updates it in place
    if y.ndim == 1:
        y = np.tile(y, (mcmc_chains, 1))
    if y.shape != (mcmc_chains, npar):
        raise ValueError("initial_values must be (np,) or (mcmc_chains, np)")
    cov = np.asarray(covariance, dtype=np.float64)
    cov = (cov + cov.T) / 2
    Lc = np.asfortranarray(np.linalg.cholesky(cov))
    logp = np.empty(mcmc_chains); acc = np.zeros(mcmc_chains, dtype=np.int64)
    hist = np.empty((mcmc_replic, mcmc_chains, npar)) if keep_history else None
    hlp = np.empty((mcmc_replic, mcmc_chains)) if keep_history else None
    L.check(ssws.lib.dyb_rwmh_run(ssws._h, mcmc_chains, first_chain, mcmc_replic, first_step,
                                  1 if transformed_parameters else 0, L.ptr(Lc), seed, L.ptr(y), L.ptr(logp),
                                  L.ptr(acc), L.ptr(hist), L.ptr(hlp)))
    out = dict(y=y, logp=logp, accepted=acc)
    if keep_history:
        out.update(history=hist, history_logp=hlp)
    return out
