"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/jchemo_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge
from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    ge.build()
    import jchemo_b200
    return jchemo_b200.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "jchemo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jcb200_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    import jchemo_b200
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(jchemo_b200.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/jchemo_b200.h but not exported"
    # and the ctypes signature table binds exactly the header's surface
    assert sorted(jchemo_b200.SIGNATURES) == syms


def test_version_and_packed_len(lib):
    assert lib.jcb200_version() == 100
    p, q = 500, 10
    assert lib.jcb200_packed_len(p, q) == p * p + p * q + 2 * q + p + 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import jchemo_b200
    with pytest.raises(jchemo_b200.JchemoB200Error, match="no CPU fallback|no CUDA device"):
        jchemo_b200.plskern(np.random.rand(10, 3), np.random.rand(10, 1), nlv=2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "jchemo.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_argument_validation_shapes():
    import jchemo_b200
    with pytest.raises(ValueError, match="DimensionMismatch"):
        jchemo_b200.plskern(np.zeros((5, 2)), np.zeros((4, 1)), nlv=1)
    with pytest.raises(TypeError):
        jchemo_b200.plskern_bang(np.zeros((5, 2), order="C"), np.zeros((5, 1), order="F"), nlv=1)
