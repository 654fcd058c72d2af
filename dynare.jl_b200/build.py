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


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
