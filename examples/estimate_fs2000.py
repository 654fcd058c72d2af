#!/usr/bin/env python
"""The reference's estimation sequence for fs2000, every loop over draws replaced by one batched call
(needs a B200 and the built library):

    prior draws -> SMC (smc.jl)                       particles on the device for all tempering stages
    particle mean -> Hessian of -log posterior        ONE batch of 1 + 4 np + 2 np (np - 1) points (estimation.jl:873-890)
    RWMH from there on many chains (run_mcmc)         proposal, prior, likelihood, accept on the device
    smoother + forecast at posterior draws            calibsmoother!, forecast_! for N draws at once

Usage:  python examples/estimate_fs2000.py [--npart 4096] [--nphi 100] [--chains 1024] [--replic 200]
Data: the synthetic 2 x 192 sample of tests/golden (the reference ships none for fs2000)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200 import mode                                                    # noqa: E402
from dynare_jl_b200.engine import BatchSSWs, DSGELogPosteriorDensity, model_description   # noqa: E402
from dynare_jl_b200.priors import PriorSet                                          # noqa: E402
from dynare_jl_b200.samplers import rwmh, smc                                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npart", type=int, default=4096)
    ap.add_argument("--nphi", type=int, default=100)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--replic", type=int, default=200)
    a = ap.parse_args()
    desc = model_description("fs2000")
    Y = dict(np.load(os.path.join(ROOT, "tests", "golden", "fs2000_golden.npz")))["Y"]
    priors = PriorSet.from_description(desc["priors"])
    ws = BatchSSWs("fs2000", Y)
    names = desc["est_names"]

    t0 = time.perf_counter()
    part, stats = smc(ws, priors, rng=np.random.default_rng(1), npart=a.npart, nphi=a.nphi, seed=1)
    t_smc = time.perf_counter() - t0
    w, theta = part["w"], part["para"]
    mean = w @ theta
    print(f"SMC: {a.npart} particles x {a.nphi} stages in {t_smc:.2f} s; log marginal data density {stats['logmdd']:.3f}")

    post = DSGELogPosteriorDensity(ws, priors)
    t0 = time.perf_counter()
    hess = mode.finite_difference_hessian(lambda T: -post(T), mean)
    cov, std = mode.mode_covariance(hess)
    print(f"Hessian at the particle mean: one batch of {mode.hessian_stencil(mean)[0].shape[0]} posterior evaluations, "
          f"{(time.perf_counter() - t0) * 1e3:.1f} ms")

    print("   posterior standard deviations from the inverse Hessian:", np.round(std, 4))

    # proposal in the sampling space R^np (DSGETransformation): scaled covariance of the transformed particle cloud
    ycloud = ws.transform(theta, inverse=True)
    y0 = w @ ycloud
    prop = 0.3 * np.cov(ycloud.T, aweights=w)
    t0 = time.perf_counter()
    chain = rwmh(ws, priors, y0, prop, mcmc_replic=a.replic, mcmc_chains=a.chains, seed=2)
    t_mh = time.perf_counter() - t0
    draws, _ = ws.transform(chain["y"])
    print(f"RWMH: {a.chains} chains x {a.replic} sweeps in {t_mh:.2f} s "
          f"({a.chains * a.replic / t_mh / 1e6:.1f} M posterior evaluations/s), acceptance "
          f"{chain['accepted'].mean() / a.replic:.2f}")
    print(f"{'parameter':>10s} {'SMC mean':>10s} {'MH mean':>10s} {'MH sd':>10s}")
    for i, nm in enumerate(names):
        print(f"{nm:>10s} {mean[i]:10.4f} {draws[:, i].mean():10.4f} {draws[:, i].std():10.4f}")

    t0 = time.perf_counter()
    sm = ws.smoother(draws[:256])
    ok = sm["status"] == 0
    fc, _ = ws.forecast(draws[:256][ok], sm["alphah"][ok][:, -1, :], 8)
    print(f"smoother + 8-period forecast for {ok.sum()} posterior draws: {(time.perf_counter() - t0) * 1e3:.0f} ms; "
          f"gy_obs forecast mean {fc[:, 1:, desc['endo'].index('gy_obs')].mean(0).round(4)}")
    ws.close()


if __name__ == "__main__":
    main()
