"""The reference's estimation ENTRY POINTS, taking a context and routing every loop over draws to the batched engine.

Reference                                                        -> here (same keyword names)
  smc_compute!(; context, datafile, first_obs, last_obs, nobs)    -> smc_compute(context, ...)
     (src/estimation/estimation.jl:298-327 -> smc, smc.jl:64-313)
  rwmh_compute!(; context, datafile, data, initial_values, covariance, transformed_covariance, mcmc_chains,
                mcmc_jscale, mcmc_replic, transformed_parameters, ...)          -> rwmh_compute(context, ...)
     (estimation.jl:233-279 -> mh_estimation :912-962 -> run_mcmc :964-989)
  mode_compute!(; context, datafile, data, initial_values, transformed_parameters, ...) -> mode_compute(context, ...)
     (estimation.jl:150-183 -> posterior_mode! :848-919: optimiser + two finite-difference Hessians)
  prior predictive check: run_simulations(context, draws)          -> run_simulations(context, draws)
     (src/estimation/priorprediction.jl:47-100)

``EstimationContext`` is the slice of the reference's ``Context`` these functions read: the compiled model
(``context.models[1]`` + the generated functions = a registered device model), ``context.work.estimated_parameters``
(names, priors, initial values), ``context.work.observed_variables`` and the ``results`` they write
(``context.results.model_results[1].estimation``).  The Julia methods of the same names are in ``julia/DynareB200.jl``.
"""
from __future__ import annotations

import csv
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import mode as _mode
from .engine import BatchSSWs, DSGELogPosteriorDensity, model_description
from .priors import PriorSet
from .samplers import SMC, rwmh


@dataclass
class EstimationContext:
    model: str                                   # name of the registered device model
    description: dict                            # names, index sets, priors, initial values (models/<name>.json)
    priors: PriorSet
    device: int = 0
    results: Dict[str, object] = field(default_factory=dict)     # posterior_mode, posterior_mode_std, smc, mcmc chains ...

    @property
    def parameter_names(self) -> List[str]:
        return list(self.description["est_names"])

    @property
    def observed_variables(self) -> List[str]:
        return list(self.description["varobs"])


def make_context(model: str, device: int = 0) -> EstimationContext:
    d = model_description(model)
    return EstimationContext(model, d, PriorSet.from_description(d["priors"]), device)


def get_observations(context: EstimationContext, datafile: str = "", data=None, first_obs: int = 1, last_obs: int = 0,
                     nobs: int = 0) -> np.ndarray:
    """ny x nobs observations in varobs order (``get_transposed_data!``, src/data.jl:107-140; NaN = missing).
    ``data``: an (ny, T) array in varobs order or a mapping name -> series; ``datafile``: CSV with a header row."""
    names = context.observed_variables
    if data is not None:
        Y = (np.vstack([np.asarray(data[v], dtype=np.float64) for v in names]) if isinstance(data, dict)
             else np.asarray(data, dtype=np.float64))
    elif datafile:
        with open(datafile, newline="") as f:
            rows = list(csv.reader(f))
        head = [h.strip() for h in rows[0]]
        cols = [head.index(v) for v in names]
        Y = np.array([[float(r[c]) if r[c].strip() not in ("", "NaN", "NA", "missing") else np.nan for c in cols]
                      for r in rows[1:] if r], dtype=np.float64).T
    else:
        raise ValueError("Either `data` or `datafile` must be specified")
    if Y.ndim != 2 or Y.shape[0] != len(names):
        raise ValueError(f"observations must be {len(names)} x T in the order {names}")
    lo = max(first_obs, 1) - 1
    hi = Y.shape[1] if last_obs <= 0 else min(last_obs, Y.shape[1])
    if nobs > 0:
        hi = min(hi, lo + nobs)
    return np.ascontiguousarray(Y[:, lo:hi])


def get_initial_value_or_mean(context: EstimationContext) -> np.ndarray:
    """``estimated_params_init`` values where given, prior means elsewhere (estimation.jl:1054-1088)."""
    return np.array(context.description["theta_init"], dtype=np.float64)


def prior_variance(context: EstimationContext) -> np.ndarray:
    """Diagonal of prior variances capped at 1 (estimation.jl:337-344)."""
    sd = np.array([float("inf") if p["stdev"] == "inf" else float(p["stdev"]) for p in context.description["priors"]])
    v = sd * sd
    return np.diag(np.where(v > 1, 1.0, v))


# ------------------------------------------------------------------------------------------------------------------------
def smc_compute(context: EstimationContext, datafile: str = "", data=None, first_obs: int = 1, last_obs: int = 0, nobs: int = 0,
                communicator_group=None, rng: Optional[np.random.Generator] = None, **tune) -> dict:
    """``smc_compute!`` -> ``smc`` with the ``Tune`` defaults (npart 1000, nphi 400, lam 2, ...; pass other values as
    keywords).  Particles live on the device for the whole recursion; with ``communicator_group`` (a torch.distributed
    group, one process per GPU) they are sharded over the ranks.  Returns dict(para, w, loglh, logpost, stats, mu, R,
    logmdd, posterior_mean, posterior_std) and stores it in ``context.results['smc']``."""
    from .samplers import Communicator
    Y = get_observations(context, datafile, data, first_obs, last_obs, nobs)
    ws = BatchSSWs(context.model, Y, device=context.device)
    comm = Communicator.from_torch_distributed(ws, group=communicator_group) if communicator_group is not None else None
    seed = int(tune.pop("seed", 0))
    s = SMC(ws, context.priors, seed=seed, **tune)
    if rng is None:
        # prior draws and the rejection loop of smc.jl:87-101 on the device; the stream is keyed by the global particle
        # index, so the initial cloud is the same however the particles are sharded
        s.initialise()
    elif comm is not None:                                   # every rank draws the same host stream and keeps its slice
        para = context.priors.sample(rng, s.npart)[s.first:s.first + s.n_local]
        s.initialise(np.random.default_rng(seed + 1 + comm.rank), para=para)
    else:
        s.initialise(rng)
    s.run()
    st = s.stats()
    part = s.particles()
    s.close()
    if comm is not None:
        comm.close()
    ws.close()
    out = dict(part, **st)
    if comm is None:
        mean = part["w"] @ part["para"]
        out["posterior_mean"] = mean
        out["posterior_std"] = np.sqrt(np.maximum(part["w"] @ (part["para"] - mean) ** 2, 0.0))
    context.results["smc"] = out
    return out


def mode_compute(context: EstimationContext, datafile: str = "", data=None, first_obs: int = 1, last_obs: int = 0, nobs: int = 0,
                 initial_values=None, transformed_parameters: bool = True, iterations: int = 1000, show_trace: bool = False) -> dict:
    """``mode_compute!`` -> ``posterior_mode!`` (estimation.jl:848-919): BFGS on minus the (transformed) log posterior,
    then the finite-difference Hessian(s) at the minimiser as ONE batch each.  The objective's gradient is a central
    difference whose 2 np points are one batched call too.  Results go to ``context.results`` under the reference's field
    names (posterior_mode, posterior_mode_std, posterior_mode_covariance, transformed_*)."""
    from scipy import optimize
    Y = get_observations(context, datafile, data, first_obs, last_obs, nobs)
    ws = BatchSSWs(context.model, Y, device=context.device)
    post = DSGELogPosteriorDensity(ws, context.priors)
    x0 = get_initial_value_or_mean(context) if initial_values is None else np.asarray(initial_values, dtype=np.float64)
    res = context.results

    def transformed_objective(Yt):                            # -logdensity of the TransformedLogDensity (N points)
        theta, lj = ws.transform(np.atleast_2d(Yt))
        v = post(theta) + lj
        return np.where(np.isfinite(v), -v, 1e30)

    def plain_objective(T):
        v = post(np.atleast_2d(T))
        return np.where(np.isfinite(v), -v, 1e30)

    obj = transformed_objective if transformed_parameters else plain_objective
    p0 = ws.transform(x0[None, :], inverse=True)[0] if transformed_parameters else x0

    def f(p):
        return float(obj(p[None, :])[0])

    def grad(p):
        h = np.maximum(np.abs(p), 1.0) * 6.0554544523933395e-06            # cbrt(eps): central differences
        pts = np.concatenate([p + np.diag(h), p - np.diag(h)])
        v = obj(pts)
        return (v[:p.size] - v[p.size:]) / (2 * h)

    opt = optimize.minimize(f, p0, jac=grad, method="BFGS", options=dict(maxiter=iterations, gtol=1e-5, disp=show_trace))
    minimizer = opt.x
    if transformed_parameters:
        hess_t = _mode.finite_difference_hessian(transformed_objective, minimizer)
        cov_t, std_t = _mode.mode_covariance(hess_t)
        res["transformed_posterior_mode"] = minimizer.copy()
        res["transformed_posterior_mode_std"] = std_t
        res["transformed_posterior_mode_covariance"] = cov_t
        theta_mode = ws.transform(minimizer[None, :])[0][0]
    else:
        theta_mode = minimizer
    hess = _mode.finite_difference_hessian(plain_objective, theta_mode)
    cov, std = _mode.mode_covariance(hess)
    res["posterior_mode"] = theta_mode.copy()
    res["posterior_mode_std"] = std
    res["posterior_mode_covariance"] = cov
    res["mode_objective"] = float(opt.fun)
    res["mode_evaluations"] = int(opt.nfev)
    ws.close()
    return dict(minimizer=minimizer, mode=theta_mode, std=std, covariance=cov, fun=float(opt.fun), success=bool(opt.success))


def rwmh_compute(context: EstimationContext, datafile: str = "", data=None, first_obs: int = 1, last_obs: int = 0, nobs: int = 0,
                 initial_values=None, covariance=None, transformed_covariance=None, mcmc_chains: int = 1,
                 mcmc_jscale: float = 0.0, mcmc_replic: int = 0, transformed_parameters: bool = True,
                 back_transformation: bool = True, seed: int = 0, keep_history: bool = True) -> dict:
    """``rwmh_compute!`` -> ``mh_estimation`` -> ``run_mcmc``: random-walk Metropolis on ``mcmc_chains`` chains at once, proposal
    covariance ``mcmc_jscale * (transformed_)covariance``.  Defaults as the reference: ``initial_values`` = prior means,
    ``covariance`` = capped prior variances.  (The reference maps ``covariance`` to the transformed space with
    ``inverse_transform_variance``, which it only defines for untransformed (Normal-prior) parameters, estimation.jl:1135-1146;
    pass ``transformed_covariance`` -- e.g. ``context.results['transformed_posterior_mode_covariance']`` -- otherwise.)
    Returns dict(chains (mcmc_replic, mcmc_chains, np) in parameter space, logp, acceptance_rate) and appends it to
    ``context.results['posterior_mcmc_chains']``."""
    Y = get_observations(context, datafile, data, first_obs, last_obs, nobs)
    ws = BatchSSWs(context.model, Y, device=context.device)
    ws.set_priors(context.priors)
    # default: prior_mean(ep) (estimation.jl:335)
    x0 = _prior_means(context) if initial_values is None else np.asarray(initial_values, dtype=np.float64)
    cov = prior_variance(context) if covariance is None else np.asarray(covariance, dtype=np.float64)
    if transformed_parameters:
        y0 = ws.transform(np.atleast_2d(x0), inverse=True) if back_transformation else np.atleast_2d(x0)
        prop = np.asarray(transformed_covariance, dtype=np.float64) if transformed_covariance is not None else cov
    else:
        y0, prop = np.atleast_2d(x0), cov
    if y0.shape[0] == 1:
        y0 = y0[0]
    out = rwmh(ws, context.priors, y0, mcmc_jscale * prop, mcmc_replic=mcmc_replic, mcmc_chains=mcmc_chains,
               transformed_parameters=transformed_parameters, seed=seed, keep_history=keep_history)
    res = dict(logp=out["logp"], accepted=out["accepted"], acceptance_rate=out["accepted"] / max(mcmc_replic, 1))
    if keep_history:
        h = out["history"]
        if transformed_parameters and back_transformation:         # transform_chains!
            flat, _ = ws.transform(h.reshape(-1, h.shape[-1]))
            h = flat.reshape(h.shape)
        res["chains"] = h
        res["chains_logp"] = out["history_logp"]
    res["last"] = out["y"]
    context.results.setdefault("posterior_mcmc_chains", []).append(res)
    ws.close()
    return res


def _prior_means(context: EstimationContext) -> np.ndarray:
    return np.array([float(p["mean"]) for p in context.description["priors"]], dtype=np.float64)


def run_simulations(context: EstimationContext, draws: np.ndarray, periods: int = 40) -> dict:
    """``run_simulations(context, draws)`` (priorprediction.jl:64-98) for all rows of ``draws`` at once: steady state,
    standard deviations, IRFs, and the failure counts by class (domain / undetermined / unstable / other)."""
    ws = BatchSSWs(context.model, None, device=context.device)
    r = ws.prior_predictive(np.asarray(draws, dtype=np.float64), periods=periods)
    ws.close()
    st = r["status"]
    r["failure"] = (st != 0).astype(np.float64)
    r["domain"] = int((st == 1).sum()); r["undetermined"] = int((st == 2).sum()); r["unstable"] = int((st == 3).sum())
    r["other"] = int((st > 3).sum())
    return r


def prior_predictive_check(context: EstimationContext, iterations: int = 1000, rng: Optional[np.random.Generator] = None) -> dict:
    """Draw ``iterations`` parameter vectors from the priors and run ``run_simulations`` on them (priorprediction.jl:47-62)."""
    draws = context.priors.sample(rng or np.random.default_rng(0), iterations)
    out = run_simulations(context, draws)
    out["draws"] = draws
    return out
