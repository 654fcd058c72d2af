"""Build libdynare_b200.so in-tree with nvcc for sm_100a (no torch involved)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libdynare_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return [os.path.join(CSRC, "api.cu"), os.path.join(CSRC, "sampler.cu"), os.path.join(CSRC, "generic.cu")] + sorted(glob.glob(os.path.join(CSRC, "generated", "*.cu")))


def _deps():
    return (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
            glob.glob(os.path.join(CSRC, "generated", "*.cuh")) +
            [os.path.join(os.path.dirname(PKG), "include", "dynare_b200.h")])


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    srcs = sources()
    newest = max(os.path.getmtime(p) for p in srcs + _deps())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(bdir, os.path.basename(src) + ".o")
        cmd = [NVCC, *FLAGS, *extra_flags, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_model_library(name: str, mod_path: str, out_dir: str, spec=None) -> str:
    """Compile ONE user model into its own shared object (the run-time analogue of Dynare's `use_dll`):
    .mod text -> device model function (modelgen) -> nvcc -> <out_dir>/libdynare_b200_model_<name>.so.
    Loading it (engine.load_model_library) registers the model with libdynare_b200.so, after which
    dyb_create("<name>", ...) works like for the shipped models.  Writes <out_dir>/<name>.json beside it."""
    from . import modelgen
    build_library()
    os.makedirs(out_dir, exist_ok=True)
    if spec is None:                                  # else: a ModelSpec built elsewhere (preprocessor.spec_from_preprocessor)
        spec = modelgen.spec_from_file(name, mod_path)
    cuh = os.path.join(out_dir, f"model_{name}.cuh")
    cu = os.path.join(out_dir, f"model_{name}.cu")
    with open(cuh, "w") as f:
        f.write(modelgen.emit_cuda(spec).replace('#include "../model_traits.cuh"', '#include "model_traits.cuh"'))
    with open(cu, "w") as f:
        f.write(f"// GENERATED -- instantiates the solve and filter kernels for model '{name}'.\n"
                f"#include \"model_{name}.cuh\"\n#define DYB_MODEL dyb::Model_{name}\n#include \"model_tu.cuh\"\n")
    modelgen.write_json(spec, os.path.join(out_dir, f"{name}.json"))
    lib = os.path.join(out_dir, f"libdynare_b200_model_{name}.so")
    cmd = [NVCC, *FLAGS, "-I", CSRC, "-I", out_dir, "-shared", "-o", lib, cu, "-L", PKG, "-l:libdynare_b200.so",
           "-Xlinker", f"-rpath={PKG}", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {cu}:\n{r.stdout}\n{r.stderr}")
    return lib


def build_sizes_library(n: int, nx: int, i_static, i_bkwrd_b, i_fwrd_b, obs_idx, out_dir: str, name: str = None) -> tuple:
    """Compile the register-resident kernels for the SIZES and index sets of a model whose steady state and Jacobian stay on
    the host (``engine.GenericSSWs(..., compile=True)``): modelgen.emit_cuda_jacobian_fed -> nvcc ->
    <out_dir>/libdynare_b200_model_<name>.so, self-registering like any compiled model.  The name defaults to a hash of the
    sizes and index sets, so that equal models share one library.  Returns (library path, model name)."""
    import hashlib
    from . import modelgen
    build_library()
    key = repr((int(n), int(nx), [int(v) for v in i_static], [int(v) for v in i_bkwrd_b], [int(v) for v in i_fwrd_b],
                [int(v) for v in obs_idx]))
    if name is None:
        name = "jf_" + hashlib.sha1(key.encode()).hexdigest()[:12]
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"libdynare_b200_model_{name}.so")
    cuh = os.path.join(out_dir, f"model_{name}.cuh")
    text = modelgen.emit_cuda_jacobian_fed(name, n, nx, i_static, i_bkwrd_b, i_fwrd_b, obs_idx).replace(
        '#include "../model_traits.cuh"', '#include "model_traits.cuh"')
    newest = max(os.path.getmtime(p) for p in _deps())
    if os.path.exists(lib) and os.path.exists(cuh) and open(cuh).read() == text and os.path.getmtime(lib) >= newest:
        return lib, name
    with open(cuh, "w") as f:
        f.write(text)
    cu = os.path.join(out_dir, f"model_{name}.cu")
    with open(cu, "w") as f:
        f.write(f"// GENERATED -- instantiates the solve and filter kernels for the sizes of '{name}'.\n"
                f"#include \"model_{name}.cuh\"\n#define DYB_MODEL dyb::Model_{name}\n#include \"model_tu.cuh\"\n")
    cmd = [NVCC, *FLAGS, "-I", CSRC, "-I", out_dir, "-shared", "-o", lib, cu, "-L", PKG, "-l:libdynare_b200.so",
           "-Xlinker", f"-rpath={PKG}", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {cu}:\n{r.stdout}\n{r.stderr}")
    return lib, name


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
