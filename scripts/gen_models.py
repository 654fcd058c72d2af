#!/usr/bin/env python
"""Regenerate the compiled-model sources from the reference's model files.

Runs in the build container only (it reads /root/reference, which does not exist on the GPU box);
everything it writes is committed:
  dynare.jl_b200/csrc/generated/model_<name>.cuh   device model function + traits
  dynare.jl_b200/csrc/generated/model_<name>.cu    translation unit instantiating the kernels
  dynare.jl_b200/models/<name>.json                plain-data model description (host side)
  oracle/c/model_<name>.c                          the same model function for the CPU oracle port
  tests/golden/ls2003_data.npy                     the ls2003 sample (first_obs=8, nobs=79), varobs order
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynare_jl_b200  # noqa: E402
from dynare_jl_b200 import modelgen  # noqa: E402

REF = os.environ.get("DYNARE_REFERENCE", "/root/reference")
MODELS = {
    "fs2000": "test/models/fs2000/fs2000.mod",
    "ls2003": "test/models/ls2003/ls2003.mod",
    # Herbst-Schorfheide NK model with prior!() syntax, calibrated and estimated measurement errors, correlations
    "hs_nk": "test/models/estimation/test1.mod",
}


# models written for this repository (sources under dynare.jl_b200/models/src): mc4 = four linked Lubik-Schorfheide economies,
# 44 variables / 20 shocks / 7 observables -- the medium-scale fixture (BASELINE configs[3]) for the path beyond m = 12
OWN_MODELS = {"mc4": os.path.join(ROOT, "dynare.jl_b200", "models", "src", "mc4.mod")}


def main():
    gen = os.path.join(ROOT, "dynare.jl_b200", "csrc", "generated")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(ROOT, "oracle", "c"), exist_ok=True)
    sources = {name: os.path.join(REF, rel) for name, rel in MODELS.items()}
    sources.update(OWN_MODELS)
    only = set(sys.argv[1:])
    for name, path in sources.items():
        if only and name not in only:
            continue
        spec = modelgen.spec_from_file(name, path)
        with open(os.path.join(gen, f"model_{name}.cuh"), "w") as f:
            f.write(modelgen.emit_cuda(spec))
        with open(os.path.join(gen, f"model_{name}.cu"), "w") as f:
            f.write(f"// GENERATED -- instantiates the solve and filter kernels for model '{name}'.\n"
                    f"#include \"model_{name}.cuh\"\n#define DYB_MODEL dyb::Model_{name}\n"
                    f"#include \"../model_tu.cuh\"\n")
        modelgen.write_json(spec, os.path.join(ROOT, "dynare.jl_b200", "models", f"{name}.json"))
        with open(os.path.join(ROOT, "oracle", "c", f"model_{name}.c"), "w") as f:
            f.write(modelgen.emit_c(spec))
        print(f"{name}: n={spec.n} k={spec.k} nf={spec.nf} s={spec.s} ns={spec.ns} ny={spec.ny} "
              f"np={spec.np_est} jac nnz={len(spec.jac)}")

    if only and "ls2003" not in only:
        return
    # ls2003 data: reference test/models/ls2003/data_ca1.csv, estimation(first_obs=8, nobs=79)
    import csv
    with open(os.path.join(REF, "test/models/ls2003/data_ca1.csv")) as f:
        rows = list(csv.reader(f))
    hdr = [h.strip() for h in rows[0]]
    data = np.array([[float(x) for x in r] for r in rows[1:] if r], dtype=np.float64)
    varobs = ["y_obs", "R_obs", "pie_obs", "dq", "de"]
    Y = data[7:7 + 79, [hdr.index(v) for v in varobs]].T.copy()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.save(os.path.join(ROOT, "tests", "golden", "ls2003_data.npy"), Y)
    print("ls2003 data", Y.shape)


if __name__ == "__main__":
    main()
