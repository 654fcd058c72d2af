#!/usr/bin/env python
"""A short SMC run (ls2003, 16 384 particles, a handful of stages) for `ncu --metrics gpu__time_duration.sum`: the launch list
shows the kernels of ONE tempering stage (seven fused kernels + the likelihood batch).  Usage: python scripts/smc_stage_launches.py [stages]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import BatchSSWs, model_description   # noqa: E402
from dynare_jl_b200.priors import PriorSet                       # noqa: E402
from dynare_jl_b200.samplers import SMC                          # noqa: E402

stages = int(sys.argv[1]) if len(sys.argv) > 1 else 6
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ls2003_golden.npz")))
pri = PriorSet.from_description(model_description("ls2003")["priors"])
ws = BatchSSWs("ls2003", g["Y"])
s = SMC(ws, pri, npart=16384, nphi=400, seed=0)
s.initialise()
s.run(stages)
print("stages done", s.stages_done(), "transport", s.transport())
s.close(); ws.close()
