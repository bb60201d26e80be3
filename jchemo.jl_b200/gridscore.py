"""Host-side mirror of `gridscorelv` for `fun = plskern` (next row after the hot path).

Reference: `/root/reference/src/gridscore.jl:167-221` (branch `pars === nothing`) with the scores of
`/root/reference/src/scores.jl`.  The reference materialises one prediction matrix per nlv and scores
each on the host; here one call to `jcb200_gridscore` returns the residual sums for every nlv in a
single pass over X (no prediction ever leaves the GPU) and the named score is a formula on them.
"""
import ctypes as C

import numpy as np

from . import _lib
from .plskern import _fmat, _ld, _ptr, plskern

SCORES = ("msep", "rmsep", "ssr", "bias", "sep", "r2", "rpd")


def residual_sums(obj, X, Y, nlv):
    """ssr[k, j], sum of residuals[k, j] for k in the clamped range, plus sum(Y), sum(Y^2), m."""
    X, Y = _fmat(X), _fmat(Y)
    a = obj.T.shape[1]
    ks = np.atleast_1d(np.asarray(nlv))
    k_lo, k_hi = max(0, int(ks.min())), min(a, int(ks.max()))
    m, p = X.shape
    q = Y.shape[1]
    if Y.shape[0] != m:
        raise ValueError(f"DimensionMismatch: X has {m} rows, Y has {Y.shape[0]}")
    nk = k_hi - k_lo + 1
    ssr = np.empty((nk, q), order="F")
    sres = np.empty((nk, q), order="F")
    ysum, ysumsq = np.empty(q), np.empty(q)
    R, Cm = np.asfortranarray(obj.R), np.asfortranarray(obj.C)
    rc = _lib.lib().jcb200_gridscore(_ptr(X), _ld(X), _ptr(Y), _ld(Y), m, p, q, _ptr(R) if a else None,
                                     _ptr(Cm) if a else None, a, _ptr(obj.xmeans), _ptr(obj.xscales),
                                     _ptr(obj.ymeans), _ptr(obj.yscales), k_lo, k_hi, _ptr(ssr),
                                     _ptr(sres), _ptr(ysum), _ptr(ysumsq))
    _lib.check(rc, "gridscorelv")
    return np.arange(k_lo, k_hi + 1), ssr, sres, ysum, ysumsq, m


def score_table(score, ssr, sres, ysum, ysumsq, m):
    """scores.jl formulas on the residual sums (r = Y - pred)."""
    ms = ssr / m
    bi = -sres / m
    if score == "msep":
        return ms
    if score == "rmsep":
        return np.sqrt(ms)
    if score == "ssr":
        return ssr
    if score == "bias":
        return bi
    if score == "sep":
        return np.sqrt(ms - bi ** 2)
    vary = ysumsq / m - (ysum / m) ** 2            # msep of the mean model = uncorrected variance
    if score == "r2":
        return 1.0 - ms / vary
    if score == "rpd":
        return np.sqrt(vary) / np.sqrt(ms)
    raise ValueError(f"score must be one of {SCORES}")


def gridscorelv(Xtrain, Ytrain, X, Y, *, score, nlv, fun=plskern, **kwargs):
    """gridscorelv(Xtrain, Ytrain, X, Y; score, fun = plskern, nlv) -> columns nlv, y1..yq
    (a pandas DataFrame when pandas is importable, else a dict of arrays)."""
    if fun is not plskern:
        raise TypeError("the fused path covers fun = plskern")
    Xtrain = _fmat(Xtrain)
    ks = np.atleast_1d(np.asarray(nlv))
    lo, hi = max(0, int(ks.min())), min(Xtrain.shape[1], int(ks.max()))      # gridscore.jl:170-173
    fm = plskern(Xtrain, Ytrain, nlv=hi, **kwargs)                            # :179
    name = score if isinstance(score, str) else getattr(score, "__name__", None)
    if name not in SCORES:
        raise ValueError(f"score must be one of {SCORES}")
    kk, ssr, sres, ysum, ysumsq, m = residual_sums(fm, X, Y, range(lo, hi + 1))
    res = score_table(name, ssr, sres, ysum, ysumsq, m)
    out = {"nlv": kk}
    for j in range(res.shape[1]):
        out[f"y{j + 1}"] = res[:, j]
    try:
        import pandas as pd
        return pd.DataFrame(out)
    except Exception:
        return out
