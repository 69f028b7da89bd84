"""CPU: the C-ABI library loads without a GPU, exports every symbol include/svirl_b200.h declares,
and refuses to run (loudly) when no device is present -- there is no CPU fallback."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_cuda


def declared_symbols():
    with open(os.path.join(ROOT, "include", "svirl_b200.h")) as f:
        txt = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(svl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from svirl_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    # the Python binding declares a signature for everything except svl_last_error
    assert set(syms) - {"svl_last_error"} <= set(_lib.SIGNATURES), set(syms) - set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    if has_cuda():
        pytest.skip("GPU present")
    from svirl_b200 import GLSolver, _lib
    with pytest.raises(_lib.SvirlB200Error, match="no CPU fallback"):
        GLSolver(Nx=16, Ny=16, dx=0.5, dy=0.5)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: no Python module under svirl_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "svirl_b200")):
        for fn in files:
            if fn.endswith(".py"):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert not re.search(r"^\s*(import|from)\s+(glnumpy|oracle|refrun|build_ref)", txt, flags=re.M), fn
