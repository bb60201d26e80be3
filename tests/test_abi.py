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
    assert lib.jcb200_version() == 200
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


def test_next_row_host_logic_without_gpu():
    """Host-side checks of the next-row mirrors that fire before any device call."""
    import jchemo_b200 as jc
    import oracle
    rng = np.random.default_rng(0)
    Xtr, Ytr, X = rng.random((30, 6)), rng.random((30, 1)), rng.random((3, 6))
    listnn = [np.arange(10), np.arange(3), np.arange(5, 20)]
    # a neighbourhood smaller than nlv throws in the reference (locwlv.jl:37); so do the oracle and the mirror
    with pytest.raises(ValueError):
        jc.locwlv(Xtr, Ytr, X, listnn=listnn, nlv=5)
    with pytest.raises(ValueError):
        oracle.locwlv(Xtr, Ytr, X, listnn=listnn, nlv=5)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        jc.locwlv(Xtr, Ytr, X[:, :4], listnn=listnn, nlv=2)
    with pytest.raises(TypeError):
        jc.locwlv(Xtr, Ytr, X, listnn=listnn, nlv=2, fun=oracle.plskern)
    fm = oracle.plskern(Xtr, Ytr, nlv=3)
    with pytest.raises(TypeError):
        jc.xfit_bang(fm, np.ascontiguousarray(rng.random((4, 6))))
    with pytest.raises(ValueError, match="DimensionMismatch"):
        jc.xresid(fm, rng.random((4, 5)))


def test_oracle_xfit_and_locwlv_properties():
    """xfit at full rank reproduces X; xresid + xfit == X; locwlv with every row as neighbour and unit weights
    equals one global fit."""
    import oracle
    rng = np.random.default_rng(1)
    X, Y = rng.random((12, 5)), rng.random((12, 2))
    fm = oracle.plskern(X, Y, nlv=5)
    assert np.allclose(oracle.xfit(fm, X), X, atol=1e-10)
    assert np.allclose(oracle.xfit(fm, X, nlv=0), np.tile(fm.xmeans, (12, 1)))
    assert np.allclose(oracle.xfit(fm, X, nlv=2) + oracle.xresid(fm, X, nlv=2), X)
    Xq = rng.random((4, 5))
    loc = oracle.locwlv(X, Y, Xq, listnn=[np.arange(12)] * 4, nlv=range(0, 4))
    glob = oracle.predict(oracle.plskern(X, Y, nlv=3), Xq, nlv=range(0, 4))
    for a in range(4):
        assert np.allclose(loc[a], glob[a], atol=1e-12)


def _header_prototypes():
    """name -> list of C parameter types (names stripped) from include/jchemo_b200.h."""
    text = open(os.path.join(ROOT, "include", "jchemo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|void|const char\s*\*|void\s*\*)\s*(jcb200_\w+)\s*\(([^)]*)\)\s*;", text):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        if params == ["void"] or params == [""]:
            params = []
        types = []
        for p in params:
            p = re.sub(r"\s+", " ", p)
            p = re.sub(r"\s*\w+$", "", p) if not p.endswith("*") else p     # drop the parameter name
            types.append(p.replace(" *", "*").strip())
        protos[m.group(1)] = types
    return protos


_C_TO_JULIA = {
    "const double*": {"Ptr{Float64}", "Ptr{Cdouble}"}, "double*": {"Ptr{Float64}", "Ptr{Cdouble}"},
    "int64_t": {"Int64"}, "int32_t": {"Int32", "Cint"}, "int": {"Cint", "Int32"},
    "int32_t*": {"Ref{Int32}", "Ptr{Int32}", "Ptr{Cint}"}, "const int*": {"Ptr{Cint}", "Ptr{Int32}"},
    "const int32_t*": {"Ptr{Cint}", "Ptr{Int32}"},
    "const int64_t*": {"Ptr{Int64}"}, "double* const*": {"Ptr{Ptr{Float64}}"}, "void*": {"Ptr{Cvoid}"},
    "const void*": {"Ptr{Cvoid}"},
}


def test_julia_ccall_signatures_match_the_header():
    """Julia cannot run in this image, so the `ccall` module is checked statically: every call names a
    declared entry point and passes exactly the header's parameter types, in order."""
    protos = _header_prototypes()
    src = open(os.path.join(ROOT, "jchemo.jl_b200", "julia", "JchemoB200", "src", "JchemoB200.jl")).read()
    calls = re.findall(r"ccall\(\(:(jcb200_\w+),\s*LIB\),\s*(\w+),\s*\(([^)]*)\)", src)
    assert len(calls) >= 14
    # the in-Jchemo binding of INTEGRATION.md (jchemo_binding/plskern_b200.jl) is held to the same check
    src2 = open(os.path.join(ROOT, "jchemo.jl_b200", "julia", "jchemo_binding", "plskern_b200.jl")).read()
    calls2 = re.findall(r"ccall\(\(:(jcb200_\w+),\s*LIBB200\),\s*(\w+),\s*\(([^)]*)\)", src2)
    assert {c[0] for c in calls2} >= {"jcb200_plskern_fit", "jcb200_transform", "jcb200_coef", "jcb200_predict_sweep",
                                      "jcb200_init", "jcb200_init_multi"}
    for name, ret, argt in calls + calls2:
        assert name in protos, f"{name} is not declared in include/jchemo_b200.h"
        jl = [t.strip() for t in argt.replace("\n", " ").split(",") if t.strip()]
        ct = protos[name]
        assert len(jl) == len(ct), f"{name}: Julia passes {len(jl)} arguments, the header declares {len(ct)}"
        for k, (j, c) in enumerate(zip(jl, ct)):
            assert c in _C_TO_JULIA, f"{name}: unmapped C type {c!r}"
            assert j in _C_TO_JULIA[c], f"{name}: argument {k + 1} is {j} in Julia but {c} in the header"
