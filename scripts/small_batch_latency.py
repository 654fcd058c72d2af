#!/usr/bin/env python
"""Per-kernel device time of the likelihood path at small batch widths (what one SMC stage on a shard costs).
Usage: python scripts/small_batch_latency.py [--model ls2003] [--sizes 2048,4096,8192,16384,65536]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ls2003")
    ap.add_argument("--sizes", default="2048,4096,8192,16384,65536")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=0, help="dyb_options.filter_lanes (0 = automatic)")
    ap.add_argument("--solver", type=int, default=0, help="dyb_options.solver (0 = fast pass + Householder pass, 2 = Householder only)")
    a = ap.parse_args()
    import torch
    from dynare_jl_b200.engine import BatchSSWs, model_description
    from dynare_jl_b200.priors import PriorSet
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{a.model}_golden.npz")))
    pri = PriorSet.from_description(model_description(a.model)["priors"])
    ws = BatchSSWs(a.model, g["Y"])
    ws.set_options(filter_lanes=a.lanes, solver=a.solver)
    dev = torch.device("cuda", 0)
    out = {"model": a.model, "filter_lanes": a.lanes, "solver": a.solver, "rows": []}
    for n in [int(x) for x in a.sizes.split(",")]:
        th = torch.from_numpy(np.ascontiguousarray(pri.sample(np.random.default_rng(1), n))).to(dev)
        ll = torch.empty(n, dtype=torch.float64, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev)
        for _ in range(3):
            ws.loglikelihood_device(th.data_ptr(), n, ll.data_ptr(), st.data_ptr())
        ws.sync()
        ws.timers_reset()
        for _ in range(a.reps):
            ws.loglikelihood_device(th.data_ptr(), n, ll.data_ptr(), st.data_ptr())
        m, c, s_ = ws.timers_read_solve_split()
        nb, solve, filt = ws.timers_read()
        out["rows"].append({"n": n, "model_ms": m / nb, "cr_ms": c / nb, "assemble_ms": s_ / nb, "solve_ms": solve / nb,
                            "filter_ms": filt / nb, "total_ms": (solve + filt) / nb,
                            "evals_per_s": n / ((solve + filt) / nb * 1e-3)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
