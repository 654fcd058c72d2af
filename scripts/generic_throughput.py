#!/usr/bin/env python
"""Throughput of the run-time-size path (host-evaluated Jacobians): fs2000 Jacobians from the oracle, tiled to N draws;
timed through GenericSSWs.loglikelihood with host arrays (H2D of 3.7 KB per draw included)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynare_jl_b200.engine import GenericSSWs          # noqa: E402
from oracle import models, pipeline as pl              # noqa: E402

COMPILE = "--compile" in sys.argv            # the register-resident kernels compiled for the model's sizes (sizes-only model)
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
name = argv[0] if len(argv) > 0 else "fs2000"
N = int(argv[1]) if len(argv) > 1 else 16384
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.npz")))
m = models.get_model(name)
ok = np.flatnonzero(g["status"] == 0)[:64]
J = []; yb = []; Se = []; Sm = []
for t in g["theta"][ok]:
    p, se, sm_ = pl.set_estimated_parameters(m, t)
    y = m.steady_state(p); J.append(m.compressed_jacobian(y, p)); yb.append(y); Se.append(se); Sm.append(sm_)
rep = (N + len(J) - 1) // len(J)
J = np.tile(np.array(J), (rep, 1, 1))[:N]; yb = np.tile(np.array(yb), (rep, 1))[:N]
Se = np.tile(np.array(Se), (rep, 1, 1))[:N]; Sm = np.tile(np.array(Sm), (rep, 1, 1))[:N]
ws = GenericSSWs(m.n, m.nx, m.i_static, m.i_bkwrd_b, m.i_fwrd_b, m.obs_idx, observations=g["Y"], compile=COMPILE)
print("compiled sizes-only model:", ws.compiled)
ll, st = ws.loglikelihood(J, yb, Se, Sm)
ref = np.tile(g["loglik"][ok], rep)[:N]
assert np.all(st == 0) and np.allclose(ll, ref, rtol=1e-8)
t = time.perf_counter()
for _ in range(3):
    ws.loglikelihood(J, yb, Se, Sm)
dt = (time.perf_counter() - t) / 3
print(f"{name} run-time-size path: {N} draws in {dt * 1e3:.1f} ms -> {N / dt / 1e6:.2f} M evals/s (host arrays in, {J.nbytes / N:.0f} B/draw)")
t = time.perf_counter(); fo = ws.first_order(J, Se); dt1 = time.perf_counter() - t
from dynare_jl_b200.engine import kalman_batch, set_kalman_variant
z = [m.state_ids.index(i) for i in m.obs_idx]
Yd = g["Y"] - 0.0
for variant in (1, 2):
    set_kalman_variant(variant)
    kalman_batch(Yd, z, fo["T"][:256], fo["RQR"][:256], fo["P0"][:256])
    t = time.perf_counter(); kalman_batch(Yd, z, fo["T"], fo["RQR"], fo["P0"]); dt2 = time.perf_counter() - t
    print(f"   dense filter variant {variant}: {dt2 * 1e3:.1f} ms (host arrays in)")
set_kalman_variant(0)
print(f"   first_order (solve + outputs to host): {dt1 * 1e3:.1f} ms")
# pinned host inputs (what a Julia host that fills dyb_host_alloc buffers sees)
from dynare_jl_b200.engine import PinnedArray
from dynare_jl_b200 import _lib as L
pJ = PinnedArray((N, J.shape[2], J.shape[1])); pJ.array[:] = J.transpose(0, 2, 1)
pyb = PinnedArray(yb.shape); pyb.array[:] = yb
pSe = PinnedArray(Se.shape); pSe.array[:] = Se.transpose(0, 2, 1)
pSm = PinnedArray(Sm.shape); pSm.array[:] = Sm.transpose(0, 2, 1)
pll = PinnedArray((N,)); pst = PinnedArray((N,), np.int32)
def run_pinned():
    L.check(ws.lib.dyb_loglik_from_jacobian_batch(ws._h, L.ptr(pJ.array), L.ptr(pyb.array), L.ptr(pSe.array), 1, L.ptr(pSm.array), 1, None,
                                                  N, L.ptr(pll.array), L.ptr(pst.array)))
run_pinned()
t = time.perf_counter()
for _ in range(5):
    run_pinned()
dt = (time.perf_counter() - t) / 5
assert np.allclose(pll.array, ref, rtol=1e-8)
print(f"   pinned host inputs: {dt * 1e3:.2f} ms -> {N / dt / 1e6:.2f} M evals/s")
# device-resident inputs: the kernels alone
import torch
from dynare_jl_b200 import _lib as L
dev = torch.device("cuda", 0)
dJ = torch.from_numpy(np.ascontiguousarray(J.transpose(0, 2, 1))).to(dev)
dyb = torch.from_numpy(yb).to(dev); dSe = torch.from_numpy(np.ascontiguousarray(Se.transpose(0, 2, 1))).to(dev)
dSm = torch.from_numpy(np.ascontiguousarray(Sm.transpose(0, 2, 1))).to(dev)
dll = torch.empty(N, dtype=torch.float64, device=dev); dst = torch.empty(N, dtype=torch.int32, device=dev)
def run():
    L.check(ws.lib.dyb_loglik_from_jacobian_batch(ws._h, L.ptr(dJ), L.ptr(dyb), L.ptr(dSe), 1, L.ptr(dSm), 1, None, N, L.ptr(dll), L.ptr(dst)))
run(); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    run()
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
assert np.allclose(dll.cpu().numpy(), ref, rtol=1e-8)
print(f"   device-resident inputs: {dt * 1e3:.2f} ms -> {N / dt / 1e6:.2f} M evals/s")
