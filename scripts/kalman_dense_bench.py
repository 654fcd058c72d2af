#!/usr/bin/env python
"""Throughput of the dense (model-free) batched Kalman filter on synthetic state spaces: BASELINE configs 4 (ns 40,
ny 7, T 200, 32768 chains) and, scaled down, 5.  State spaces are generated once for a pool of distinct draws and
tiled on the device; timing is with CUDA events around dyb_kalman_batch called with device pointers.
Usage: python scripts/kalman_dense_bench.py [--ns 40 --ny 7 --nobs 200 --n 32768 --variant 0|1|2]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", type=int, default=40); ap.add_argument("--ny", type=int, default=7)
    ap.add_argument("--nobs", type=int, default=200); ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--pool", type=int, default=64); ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import scipy.linalg as sla
    import torch
    from dynare_jl_b200 import _lib as L
    from dynare_jl_b200.engine import set_kalman_variant, fp64_peak, fp64_peak_dmma
    lib = L.load()
    rng = np.random.default_rng(2)
    ns, ny = a.ns, a.ny
    T = np.empty((a.pool, ns, ns)); Q = np.empty_like(T); P = np.empty_like(T)
    for i in range(a.pool):
        Qo, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        T[i] = Qo @ np.diag(rng.uniform(0.3, 0.97, ns)) @ Qo.T
        R = 0.1 * rng.standard_normal((ns, ny))
        Q[i] = R @ R.T
        P[i] = sla.solve_discrete_lyapunov(T[i], Q[i])
    H = np.tile(0.01 * np.eye(ny), (a.n, 1, 1))
    Y = rng.standard_normal((ny, a.nobs))
    dev = torch.device("cuda", 0)
    rep = (a.n + a.pool - 1) // a.pool
    cm = lambda x: torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).to(dev).repeat(rep, 1, 1)[:a.n].contiguous()
    dT, dQ, dP, dH = cm(T), cm(Q), cm(P), torch.from_numpy(H).to(dev)
    dY = torch.from_numpy(np.asfortranarray(Y).T.copy()).to(dev)           # ny x nobs column-major
    dz = torch.arange(ny, dtype=torch.int32, device=dev)
    dll = torch.empty(a.n, dtype=torch.float64, device=dev); dst = torch.empty(a.n, dtype=torch.int32, device=dev)
    set_kalman_variant(a.variant)

    def run():
        L.check(lib.dyb_kalman_batch(0, ns, ny, a.nobs, 0, L.ptr(dz), L.ptr(dY), L.ptr(dT), L.ptr(dQ), L.ptr(dP), None,
                                     L.ptr(dH), a.n, L.ptr(dll), L.ptr(dst)))
    run()
    torch.cuda.synchronize()
    import time
    best = 1e30
    for _ in range(a.reps):
        t = time.perf_counter(); run(); best = min(best, time.perf_counter() - t)     # the call synchronises
    w_kf = 4 * ns ** 3 + 2 * ny * ns ** 2 + 3 * ns ** 2 + 2 * ny ** 2 * ns + 2 * ny * ns + ny ** 3 / 3 + 3 * ny ** 2 + 4 * ny
    flop = w_kf * a.nobs * a.n
    out = {"workload": f"dense Kalman ns={ns} ny={ny} T={a.nobs}, {a.n} draws", "variant": a.variant, "seconds": best,
           "evals_per_s": a.n / best, "algorithmic_tflops": flop / best / 1e12, "dfma_peak": fp64_peak(), "dmma_peak": fp64_peak_dmma(0),
           "ok": int((dst == 0).sum().item()), "loglik0": float(dll[0].item())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
