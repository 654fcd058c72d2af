#!/usr/bin/env python
"""How many batches in flight pay?  The headline step (fs2000, 65 536 draws) round-robin over NH handles / streams, host clock
over 400 steps.  Usage: python scripts/streams_experiment.py [NH ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import BatchSSWs, model_description   # noqa: E402
from dynare_jl_b200.priors import PriorSet                       # noqa: E402

n = 65536
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "fs2000_golden.npz")))
pri = PriorSet.from_description(model_description("fs2000")["priors"])
dev = torch.device("cuda", 0)
pool = 16
th = torch.from_numpy(np.ascontiguousarray(pri.sample(np.random.default_rng(0), n * pool).reshape(pool, n, -1))).to(dev)
for nh in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    W = [BatchSSWs("fs2000", g["Y"]) for _ in range(nh)]
    ll = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(nh)]
    st = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(nh)]
    for i in range(3 * nh):
        W[i % nh].loglikelihood_device(th[i % pool].data_ptr(), n, ll[i % nh].data_ptr(), st[i % nh].data_ptr())
    for w in W:
        w.sync()
    steps = 400
    t = time.perf_counter()
    for i in range(steps):
        W[i % nh].loglikelihood_device(th[i % pool].data_ptr(), n, ll[i % nh].data_ptr(), st[i % nh].data_ptr())
    for w in W:
        w.sync()
    dt = time.perf_counter() - t
    print(f"{nh} handle(s): {dt / steps * 1e3:.4f} ms per step -> {n * steps / dt / 1e6:.1f} M evals/s")
    for w in W:
        w.close()
