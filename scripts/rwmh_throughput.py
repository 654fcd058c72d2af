#!/usr/bin/env python
"""Random-walk Metropolis on many chains at once (run_mcmc, estimation.jl:964-989): chains x sweeps per second on one
GPU.  Usage: python scripts/rwmh_throughput.py [fs2000] [32768] [100]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import BatchSSWs, model_description   # noqa: E402
from dynare_jl_b200.priors import PriorSet                       # noqa: E402
from dynare_jl_b200.samplers import rwmh                         # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fs2000"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.npz")))
desc = model_description(name)
pri = PriorSet.from_description(desc["priors"])
ws = BatchSSWs(name, g["Y"])
ws.set_priors(pri)
y0 = ws.transform(np.array([desc["theta_init"]]), inverse=True)[0]
cov = np.diag(np.full(ws.np, 0.01 ** 2))
rwmh(ws, pri, y0, cov, mcmc_replic=3, mcmc_chains=chains, seed=1)
t = time.perf_counter()
out = rwmh(ws, pri, y0, cov, mcmc_replic=sweeps, mcmc_chains=chains, seed=2)
dt = time.perf_counter() - t
acc = out["accepted"].sum() / (chains * sweeps)
print(f"{name}: {chains} chains x {sweeps} sweeps in {dt:.3f} s -> {chains * sweeps / dt / 1e6:.1f} M posterior evaluations/s, "
      f"{1e3 * dt / sweeps:.3f} ms/sweep, acceptance {acc:.3f}, logp range [{out['logp'].min():.1f}, {out['logp'].max():.1f}]")
