"""The C-ABI library loads and exports exactly what include/dynare_b200.h declares (no compute)."""
import ctypes as C
import json
import os

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from dynare_jl_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from dynare_jl_b200 import _lib
    declared = _lib.header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_registry(lib):
    from dynare_jl_b200 import _lib
    assert lib.dyb_version() == 100
    names = [lib.dyb_model_name(i).decode() for i in range(lib.dyb_model_count())]
    assert sorted(names) == ["fs2000", "hs_nk", "ls2003", "mc4"]
    for nm in names:
        d = _lib.ModelDims()
        assert lib.dyb_model_get_dims(nm.encode(), C.byref(d)) == 0
        with open(os.path.join(ROOT, "dynare.jl_b200", "models", f"{nm}.json")) as f:
            js = json.load(f)
        for key in ("n", "nx", "npar", "k", "nf", "s", "ncol", "ns", "ny"):
            assert getattr(d, key) == js[key], key
        assert d.np_default == js["np_est"]
    d = _lib.ModelDims()
    assert lib.dyb_model_get_dims(b"nope", C.byref(d)) != 0
    assert b"unknown model" in lib.dyb_last_error()


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device the product path must fail loudly, not fall back."""
    cnt = C.c_int(0)
    rc = lib.dyb_device_count(C.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip("a GPU is present")
    from dynare_jl_b200 import _lib, engine
    with pytest.raises(_lib.DybError):
        engine.BatchSSWs("fs2000")


def test_missing_library_raises(monkeypatch, tmp_path):
    from dynare_jl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()
