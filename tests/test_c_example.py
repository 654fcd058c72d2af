"""The plain-C example links against the shared library with nothing but the public header (no Python, no torch in
the boundary) and, on a GPU box, runs."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from dynare_jl_b200 import build
    build.build_library()
    exe = str(tmp_path / "c_abi_example")
    pkg = os.path.join(ROOT, "dynare.jl_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_example.c"),
           "-L", pkg, "-l:libdynare_b200.so", f"-Wl,-rpath,{pkg}", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_example_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr          # without a device it reports so and exits 0


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout and "status 1 loglik -inf" in r.stdout
