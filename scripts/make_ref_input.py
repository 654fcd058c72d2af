#!/usr/bin/env python
"""Inputs for scripts/julia_golden.jl: for each shipped model, the parameter vectors and observations on which the REAL
Dynare.jl `loglikelihood` / `first_order_solver!` are to be evaluated (on any box that has Julia + Dynare.jl), written
as tests/golden/ref_input_<model>.json.  The Julia script writes tests/golden/ref_<model>.json beside it and
tests/test_reference_pinned.py consumes it (skipping while it is absent).

Draws: the first 32 golden prior draws (index 0 is the estimated_params_init point) plus stress draws that sit where
implementations can legitimately differ: a root pushed towards the unit circle, tiny shock standard deviations (F close
to singular), and draws on both sides of the determinacy bound.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynare_jl_b200  # noqa: E402,F401
from dynare_jl_b200.engine import model_description  # noqa: E402
from oracle import models  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MODFILES = {"fs2000": "test/models/fs2000/fs2000.mod", "ls2003": "test/models/ls2003/ls2003.mod",
            "hs_nk": "test/models/estimation/test1.mod"}
# (parameter name, value) overrides applied to the estimated_params_init point, one stress draw each
STRESS = {
    "fs2000": [[("rho", 0.999)], [("rho", 0.999999)], [("e_a", 1e-6)], [("e_m", 1e-7), ("e_a", 1e-7)],
               [("del", 0.9)], [("alp", 0.01)], [("bet", 0.5)]],
    "ls2003": [[("rho_R", 0.9999)], [("rho_A", 0.99999), ("rho_ys", 0.99999)], [("psi1", 0.5)], [("psi1", 0.99)],
               [("psi1", 1.01)], [("e_R", 1e-6)], [("rho_q", 0.0)]],
    "hs_nk": [[("rho_g", 0.99999)], [("psi1", 0.7)], [("psi1", 1.0001)], [("kappa", 1e-4)], [("ep", 1e-6)],
              [("eg:ez", 0.99)], [("INFL", 1e-8)]],
}


def est_names(desc):
    return list(desc["est_names"])


def main():
    for name, modfile in MODFILES.items():
        g = dict(np.load(os.path.join(GOLD, f"{name}_golden.npz")))
        desc = model_description(name)
        m = models.get_model(name)
        names = est_names(desc)
        theta = [g["theta"][i].tolist() for i in range(32)]
        for over in STRESS[name]:
            t = np.array(m.theta_mode, dtype=float)
            for k, v in over:
                t[names.index(k)] = v
            theta.append(t.tolist())
        Y = g["Y"]
        out = {"model": name, "modfile": modfile, "varobs": list(desc["varobs"]), "estimated": names,
               "Y": [[None if np.isnan(v) else float(v) for v in row] for row in Y], "theta": theta,
               "n_prior_draws": 32, "stress": [[list(kv) for kv in o] for o in STRESS[name]]}
        with open(os.path.join(GOLD, f"ref_input_{name}.json"), "w") as f:
            json.dump(out, f)
        print(name, len(theta), "draws,", Y.shape, "observations ->", f"tests/golden/ref_input_{name}.json")


if __name__ == "__main__":
    main()
