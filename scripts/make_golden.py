#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the CPU oracle (oracle/pipeline.py).

The reference stores no likelihood / policy-matrix values (parity unpinned), so
these vectors are outputs of the oracle restatement, pinned by the identities in tests/test_oracle.py.
Inputs follow SURVEY.md section 8d: fs2000 observations simulated at the estimated_params_init point
with numpy default_rng(20260101), T = 192; parameter draws from the priors with seed 1.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynare_jl_b200  # noqa: E402
from dynare_jl_b200.engine import model_description  # noqa: E402
from dynare_jl_b200.priors import PriorSet  # noqa: E402
from oracle import models, pipeline as pl  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def make(name, Y, n_draws, seed):
    m = models.get_model(name)
    pri = PriorSet.from_description(model_description(name)["priors"])
    rng = np.random.default_rng(seed)
    theta = pri.sample(rng, n_draws)
    theta[0] = m.theta_mode
    ll = np.empty(n_draws); st = np.empty(n_draws, dtype=np.int32)
    g1 = np.full((n_draws, m.n, m.k + m.nx), np.nan)
    for i in range(n_draws):
        ll[i], st[i] = pl.loglikelihood(m, theta[i], Y)
        if st[i] == 0:
            d = pl.solve_draw(m, theta[i])
            g1[i] = np.hstack([d["sol"]["g1_1"], d["sol"]["g1_2"]])
    np.savez_compressed(os.path.join(GOLD, f"{name}_golden.npz"), Y=Y, theta=theta, loglik=ll, status=st, g1=g1)
    print(name, "draws", n_draws, "ok", int((st == 0).sum()), "status hist", np.bincount(st, minlength=7))


def main():
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    if not only or "fs2000" in only:
        fs = models.get_model("fs2000")
        Yfs = pl.simulate_observations(fs, fs.theta_mode, 192, 20260101)
        make("fs2000", Yfs, 256, 1)
    if not only or "ls2003" in only:
        Yls = np.load(os.path.join(GOLD, "ls2003_data.npy"))
        make("ls2003", Yls, 256, 1)
    if not only or "mc4" in only:
        # mc4: four linked Lubik-Schorfheide economies (dynare.jl_b200/models/src/mc4.mod), 200 synthetic quarters
        mc = models.get_model("mc4")
        make("mc4", pl.simulate_observations(mc, mc.theta_mode, 200, 20260103), 128, 1)
    if only and "hs_nk" not in only:
        return
    # hs_nk: the reference's data file (dsge1_data.csv) is not in the tree; 80 synthetic quarters at the
    # estimated_params_init point, measurement errors included
    nk = models.get_model("hs_nk")
    make("hs_nk", pl.simulate_observations(nk, nk.theta_mode, 80, 20260102), 256, 1)


if __name__ == "__main__":
    main()
