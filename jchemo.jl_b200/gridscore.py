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


def gridcvlv(X, Y, *, segm, score, nlv, fun=plskern, scal=False):
    """gridcvlv(X, Y; segm, score, fun = plskern, nlv) -> (res, res_rep), reference
    `/root/reference/src/gridcv.jl:187-228` (branch `pars === nothing`).  `segm` is a list of repetitions,
    each a list of zero-based row-index arrays (the segments; disjoint within a repetition, as `segmkf` /
    `segmts` build them).  Every repetition is ONE call of `jcb200_gridcv`: Gram down-dating on the GPU
    instead of K fits on K row-copies (next row)."""
    if fun is not plskern:
        raise TypeError("the fused path covers fun = plskern")
    name = score if isinstance(score, str) else getattr(score, "__name__", None)
    if name not in SCORES:
        raise ValueError(f"score must be one of {SCORES}")
    X, Y = _fmat(X), _fmat(Y)
    n, p = X.shape
    q = Y.shape[1]
    if Y.shape[0] != n:
        raise ValueError(f"DimensionMismatch: X has {n} rows, Y has {Y.shape[0]}")
    ks = np.atleast_1d(np.asarray(nlv))
    lo, hi = max(0, int(ks.min())), min(p, int(ks.max()))                    # gridcv.jl:193
    nk = hi - lo + 1
    lib = _lib.lib()
    rows = []
    for i, listsegm in enumerate(segm):
        segs = [np.asarray(s, dtype=np.int64).reshape(-1) for s in listsegm]
        allidx = np.concatenate(segs)
        if np.unique(allidx).size != allidx.size:
            raise ValueError("segments of a repetition must be disjoint")
        rest = np.setdiff1d(np.arange(n, dtype=np.int64), allidx, assume_unique=True)
        perm = np.ascontiguousarray(np.concatenate([allidx, rest]))
        seg_start = np.ascontiguousarray(np.concatenate([[0], np.cumsum([s.size for s in segs])]).astype(np.int64))
        K = len(segs)
        ssr = np.empty((K, q, nk))           # per segment: nk x q column-major
        sres = np.empty((K, q, nk))
        ysum, ysumsq = np.empty((K, q)), np.empty((K, q))
        rc = lib.jcb200_gridcv(_ptr(X), _ld(X), _ptr(Y), _ld(Y), n, p, q, _ptr(perm), _ptr(seg_start), K, lo,
                               hi, 1 if scal else 0, 1 if i > 0 else 0, _ptr(ssr), _ptr(sres), _ptr(ysum),
                               _ptr(ysumsq))
        _lib.check(rc, "gridcvlv")
        for j in range(K):
            tab = score_table(name, ssr[j].T, sres[j].T, ysum[j], ysumsq[j], segs[j].size)
            for t in range(nk):
                rows.append([i + 1, j + 1, lo + t] + list(tab[t]))
    arr = np.array(rows, dtype=float)
    res_rep = {"repl": arr[:, 0].astype(int), "segm": arr[:, 1].astype(int), "nlv": arr[:, 2].astype(int)}
    for c in range(q):
        res_rep[f"y{c + 1}"] = arr[:, 3 + c]
    kk = np.unique(res_rep["nlv"])
    res = {"nlv": kk}
    for c in range(q):
        res[f"y{c + 1}"] = np.array([arr[res_rep["nlv"] == k, 3 + c].mean() for k in kk])
    try:
        import pandas as pd
        res, res_rep = pd.DataFrame(res), pd.DataFrame(res_rep)
    except Exception:
        pass
    from collections import namedtuple
    return namedtuple("GridcvResult", ["res", "res_rep"])(res, res_rep)


def locwlv(Xtrain, Ytrain, X, *, listnn, listw=None, fun=plskern, nlv, scal=False):
    """locwlv(Xtrain, Ytrain, X; listnn, listw = nothing, fun = plskern, nlv) -> (pred = ...,), reference
    `/root/reference/src/locwlv.jl:9-48`: one weighted kernel-PLS fit per row of X on its neighbours
    `listnn[i]` (zero-based row indices of Xtrain) with weights `listw[i]`, predictions for every nlv of the
    clamped range.  All m tiny fits run in ONE kernel launch (`jcb200_locw_plskern`, next row SURVEY 8f-4)."""
    from .plskern import PredResult
    if fun is not plskern:
        raise TypeError("the batched path covers fun = plskern")
    Xtrain, Ytrain, X = _fmat(Xtrain), _fmat(Ytrain), _fmat(X)
    ntr, p = Xtrain.shape
    q = Ytrain.shape[1]
    m = X.shape[0]
    if X.shape[1] != p or Ytrain.shape[0] != ntr or len(listnn) != m:
        raise ValueError("DimensionMismatch in locwlv")
    ks = np.atleast_1d(np.asarray(nlv))
    lo, hi = max(0, int(ks.min())), min(p, int(ks.max()))                        # locwlv.jl:14
    nk = hi - lo + 1
    segs = [np.atleast_1d(np.asarray(s, dtype=np.int64)) for s in listnn]
    off = np.ascontiguousarray(np.concatenate([[0], np.cumsum([s.size for s in segs])]).astype(np.int64))
    idx = np.ascontiguousarray(np.concatenate(segs))
    for i, sg in enumerate(segs):
        # predict(fm; nlv = k) with k above the model's LV count gives an empty range and the reference's
        # zpred assignment throws (locwlv.jl:37, plskern.jl:229); the C entry point would clamp instead
        if min(sg.size, p) < hi and not (q == 1 and np.unique(Ytrain[sg]).size == 1):
            raise ValueError(f"DimensionMismatch: neighbourhood {i} has {sg.size} rows, fewer than nlv = {hi}")
    w = None
    if listw is not None:
        w = np.ascontiguousarray(np.concatenate([np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in listw]))
        if w.size != idx.size:
            raise ValueError("listw does not match listnn")
    pred = np.empty((m, q, nk), order="F")
    rc = _lib.lib().jcb200_locw_plskern(_ptr(Xtrain), _ld(Xtrain), _ptr(Ytrain), _ld(Ytrain), ntr, p, q, _ptr(X),
                                        _ld(X), m, _ptr(idx), _ptr(off), _ptr(w), lo, hi, 1 if scal else 0,
                                        _ptr(pred))
    _lib.check(rc, "locwlv")
    out = [np.asfortranarray(pred[:, :, a]) for a in range(nk)]
    return PredResult(out[0] if nk == 1 else out)
