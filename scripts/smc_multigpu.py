#!/usr/bin/env python
"""Config 3 of BASELINE.json: ls2003 SMC with the particles sharded over the GPUs of one node.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             scripts/smc_multigpu.py [--npart 16384] [--nphi 400] [--check]
One process per GPU; torch.distributed (gloo) only bootstraps the NCCL id and gathers results on the host.
--check also runs the same recursion on ONE GPU (rank 0) and requires bit-identical particles and records:
sums use the canonical 1024-block order and random numbers are keyed by the global particle index, so the
result does not depend on the number of GPUs.  Prints one JSON line with the per-stage time split.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ls2003")
    ap.add_argument("--npart", type=int, default=16384)
    ap.add_argument("--nphi", type=int, default=400)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--profile", action="store_true", help="record the per-stage time split with CUDA events")
    ap.add_argument("--resampling", default="multinomial")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)

    from dynare_jl_b200.engine import BatchSSWs, model_description
    from dynare_jl_b200.priors import PriorSet
    from dynare_jl_b200.samplers import SMC, Communicator

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{a.model}_golden.npz")))
    pri = PriorSet.from_description(model_description(a.model)["priors"])
    ws = BatchSSWs(a.model, g["Y"], device=local)
    comm = Communicator.from_torch_distributed(ws) if world > 1 else None

    # every rank draws the SAME initial particle set and keeps its own slice (the host-side pdraw! closure)
    def initial(n_total):
        rng = np.random.default_rng(a.seed)
        probe = BatchSSWs(a.model, g["Y"], device=local)
        para = pri.sample(rng, n_total)
        while True:
            _, st = probe.loglikelihood(para)
            bad = np.flatnonzero(st != 0)
            if bad.size == 0:
                break
            para[bad] = pri.sample(rng, bad.size)
        probe.close()
        return para

    para0 = initial(a.npart)
    s = SMC(ws, pri, npart=a.npart, nphi=a.nphi, seed=a.seed, resampling=a.resampling)
    s.initialise(np.random.default_rng(0), para=para0[s.first:s.first + s.n_local].copy())
    s.set_profiling(a.profile)
    if world > 1:
        dist.barrier()
    ws.sync()
    t0 = time.perf_counter()
    s.run()
    st = s.stats()                      # synchronises
    dt = time.perf_counter() - t0
    part = s.particles()
    if world > 1:
        boxes = [None] * world
        dist.all_gather_object(boxes, part)
        times = [None] * world
        dist.all_gather_object(times, dt)
        dt = max(times)
        full = {k: np.concatenate([b[k] for b in boxes]) for k in part}
    else:
        full = part
    out = {"workload": f"{a.model} SMC, {a.npart} particles, {a.nphi} stages, {a.resampling} resampling",
           "n_gpus": world, "seconds": dt, "ms_per_stage": 1e3 * dt / (a.nphi - 1),
           "particle_moves_per_s": a.npart * (a.nphi - 1) / dt, "logmdd": st["logmdd"],
           "resamples": int(st["resampled"].sum()), "final_acpt": float(st["acpt"][-1]), "final_c": float(st["c"][-1]),
           "posterior_mean": (full["w"] @ full["para"]).tolist() if rank == 0 else None}
    if a.profile:
        out["stage_split_ms_total_rank0"] = s.timing()
    if a.check and rank == 0:
        ws1 = BatchSSWs(a.model, g["Y"], device=local)
        s1 = SMC(ws1, pri, npart=a.npart, nphi=a.nphi, seed=a.seed, resampling=a.resampling)
        s1.initialise(np.random.default_rng(0), para=para0.copy())
        s1.run()
        st1 = s1.stats(); p1 = s1.particles()
        same = all(np.array_equal(full[k], p1[k]) for k in full) and np.array_equal(st["stats"], st1["stats"])
        out["identical_to_single_gpu"] = bool(same)
        if not same:
            out["max_abs_para_diff"] = float(np.abs(full["para"] - p1["para"]).max())
            out["stats_diff_rows"] = np.flatnonzero(np.any(st["stats"] != st1["stats"], axis=1)).tolist()[:10]
        s1.close(); ws1.close()
    if rank == 0:
        print(json.dumps(out))
    s.close()
    if comm:
        comm.close()
    ws.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if a.check and rank == 0 and not out.get("identical_to_single_gpu", True):
        sys.exit(1)


if __name__ == "__main__":
    main()
