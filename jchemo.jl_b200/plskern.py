"""Host-side mirror of the reference's Julia API for the kernel-PLS path.

Same names, argument meaning and error behaviour as `/root/reference/src/plskern.jl`:
`plskern` (:106-110), `plskern!` → `plskern_bang` (:112-178), `transform` (:187-195),
`coef` (:207-217), `predict` (:226-238), struct `Plsr` (:1-14).  All arithmetic happens in
libjchemo_b200.so on the GPU through the C ABI; this file only coerces shapes
(`ensure_mat`, utility.jl:544-548), allocates the caller-owned outputs and maps status codes to
exceptions.  There is no CPU fallback.
"""
import ctypes as C
import os
import weakref
from collections import namedtuple
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

CoefResult = namedtuple("CoefResult", ["B", "int"])
PredResult = namedtuple("PredResult", ["pred"])


@dataclass
class Plsr:
    """plskern.jl:1-14 — 12 fields in the reference's positional order; `V` aliases `P`."""
    T: np.ndarray
    P: np.ndarray
    R: np.ndarray
    W: np.ndarray
    C: np.ndarray
    TT: np.ndarray
    xmeans: np.ndarray
    xscales: np.ndarray
    ymeans: np.ndarray
    yscales: np.ndarray
    weights: np.ndarray
    niter: Optional[np.ndarray] = None

    @property
    def V(self):
        return self.P


def ensure_mat(X):
    """utility.jl:544-548: vector -> n x 1 matrix, number -> 1 x 1, DataFrame -> Matrix."""
    if hasattr(X, "to_numpy"):
        X = X.to_numpy()
    X = np.asarray(X)
    if X.ndim == 0:
        X = X.reshape(1, 1)
    elif X.ndim == 1:
        X = X.reshape(-1, 1)
    elif X.ndim != 2:
        raise TypeError("ensure_mat: expected a number, vector or matrix")
    return X


def _fmat(X):
    """Float64 column-major view/copy with unit row stride."""
    X = ensure_mat(X)
    if X.dtype != np.float64 or not X.flags.f_contiguous:
        X = np.asfortranarray(X, dtype=np.float64)
    return X


def _out_empty(shape):
    """Column-major Float64 output array.  Large ones come from the library's page-locked pool
    (jcb200_host_alloc): the device-to-host copy of the scores then runs at PCIe speed; the block goes
    back to the pool when the array is garbage collected.  Falls back to ordinary memory."""
    nbytes = int(np.prod(shape)) * 8
    if nbytes >= int(os.environ.get("JCB_PINNED_MIN_BYTES", 1 << 22)) and nbytes > 0:
        lib = _lib.lib()
        p = lib.jcb200_host_alloc(nbytes)
        if p:
            buf = (C.c_char * nbytes).from_address(p)
            arr = np.frombuffer(buf, dtype=np.float64).reshape(shape, order="F")
            weakref.finalize(buf, lib.jcb200_host_free, C.c_void_p(p))
            return arr
    return np.empty(shape, order="F")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _ld(a):
    return max(a.shape[0], 1) if a.shape[1] <= 1 else a.strides[1] // 8


def _fit(X, Y, weights, nlv, scal, writeback):
    lib = _lib.lib()
    n, p = X.shape
    q = Y.shape[1]
    if Y.shape[0] != n:
        raise ValueError(f"DimensionMismatch: X has {n} rows, Y has {Y.shape[0]}")
    if weights is None:
        w = None
    else:
        w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1))
        if w.shape[0] != n:
            raise ValueError(f"DimensionMismatch: weights has length {w.shape[0]}, X has {n} rows")
    nlv = int(nlv)
    a = max(0, min(n, p, nlv))                                   # plskern.jl:116
    T = _out_empty((n, a))
    P = np.empty((p, a), order="F")
    R = np.empty((p, a), order="F")
    W = np.empty((p, a), order="F")
    Cm = np.empty((q, a), order="F")
    TT = np.empty(a)
    xmeans, xscales = np.empty(p), np.empty(p)
    ymeans, yscales = np.empty(q), np.empty(q)
    w_out = _out_empty((n,))
    nlv_out = C.c_int32(0)
    rc = lib.jcb200_plskern_fit(_ptr(X), _ld(X), _ptr(Y), _ld(Y), _ptr(w), n, p, q, nlv,
                                1 if scal else 0, 1 if writeback else 0, _ptr(T), max(n, 1), _ptr(P),
                                _ptr(R), _ptr(W), _ptr(Cm), _ptr(TT), _ptr(xmeans), _ptr(xscales),
                                _ptr(ymeans), _ptr(yscales), _ptr(w_out), C.byref(nlv_out))
    _lib.check(rc, "plskern")
    assert nlv_out.value == a
    return Plsr(T, P, R, W, Cm, TT, xmeans, xscales, ymeans, yscales, w_out, None)


def last_fit_info():
    """Facts about the calling thread's last fit: `nlv_effective` = LVs that carry information (TT > 0 and C != 0);
    degenerate LVs (the reference divides 0/0 at plskern.jl:152,166) are returned inert."""
    k = C.c_int32(0)
    _lib.check(_lib.lib().jcb200_last_fit_info(C.byref(k)), "last_fit_info")
    return namedtuple("FitInfo", ["nlv_effective"])(k.value)


def _resident_mat(A):
    if not (isinstance(A, np.ndarray) and A.ndim == 2 and A.dtype == np.float64 and A.flags.f_contiguous):
        raise TypeError("resident: needs a Float64 column-major matrix (the very array later calls are handed)")
    return A


def resident_add(A):
    """Upload A once and keep the device copy (jcb200_resident_add): later calls that are handed this very array
    as X or Y skip the host-to-device transfer.  Do not modify A on the host while it is resident."""
    A = _resident_mat(A)
    _lib.check(_lib.lib().jcb200_resident_add(_ptr(A), _ld(A), A.shape[0], A.shape[1]), "resident_add")
    return A


def resident_drop(A):
    _lib.check(_lib.lib().jcb200_resident_drop(_ptr(_resident_mat(A))), "resident_drop")


class resident:
    """`with resident(X, Y): fm = plskern(X, Y, nlv=25); summary(fm, X); gridscorelv(...)` — one upload for the
    block (the reference's workflows reuse one X: gridscore.jl:179-180, plskern.jl:246-249)."""

    def __init__(self, *mats):
        self.mats = [_resident_mat(A) for A in mats]

    def __enter__(self):
        done = []
        try:
            for A in self.mats:
                resident_add(A)
                done.append(A)
        except Exception:
            for A in done:
                resident_drop(A)
            raise
        return self

    def __exit__(self, *exc):
        for A in self.mats:
            try:
                resident_drop(A)
            except _lib.JchemoB200Error:
                pass                      # dropped by a call that overwrote it on the host (xfit!, xresid!)
        return False


def plskern(X, Y, weights=None, *, nlv, scal=False):
    """plskern(X, Y, weights = ones(n); nlv, scal = false) — inputs are left untouched (:106-110)."""
    return _fit(_fmat(X), _fmat(Y), weights, nlv, scal, writeback=False)


def plskern_bang(X, Y, weights=None, *, nlv, scal=False):
    """plskern!(X::Matrix, Y::Matrix, weights; nlv, scal): X and Y leave centred (and scaled)
    in the caller's arrays (:125-129).  Like the reference's `::Matrix` signature this refuses
    anything that is not a Float64 column-major matrix (a MethodError there, a TypeError here)."""
    for name, A in (("X", X), ("Y", Y)):
        if not (isinstance(A, np.ndarray) and A.ndim == 2 and A.dtype == np.float64 and
                A.flags.f_contiguous and A.flags.writeable):
            raise TypeError(f"plskern_bang: {name} must be a writable Float64 column-major matrix")
    return _fit(X, Y, weights, nlv, scal, writeback=True)


def transform(obj, X, *, nlv=None):
    """transform(object::Plsr, X; nlv = nothing) (:187-195)."""
    X = _fmat(X)
    a = obj.T.shape[1]
    nlv = a if nlv is None else min(int(nlv), a)
    m, p = X.shape
    if p != obj.R.shape[0]:
        raise ValueError(f"DimensionMismatch: X has {p} columns, the model has {obj.R.shape[0]}")
    nlv = max(nlv, 0)
    T = _out_empty((m, nlv))
    if nlv > 0 and m > 0:
        R = np.asfortranarray(obj.R)
        rc = _lib.lib().jcb200_transform(_ptr(X), _ld(X), m, p, _ptr(obj.xmeans), _ptr(obj.xscales),
                                         _ptr(R), nlv, _ptr(T), max(m, 1))
        _lib.check(rc, "transform")
    return T


def _xfit(obj, X, nlv, resid, out):
    a = obj.T.shape[1]
    nlv = a if nlv is None else min(int(nlv), a)
    nlv = max(nlv, 0)
    m, p = X.shape
    if p != obj.R.shape[0]:
        raise ValueError(f"DimensionMismatch: X has {p} columns, the model has {obj.R.shape[0]}")
    if m > 0:
        R, P = np.asfortranarray(obj.R), np.asfortranarray(obj.P)
        rc = _lib.lib().jcb200_xfit(_ptr(X), _ld(X), m, p, _ptr(obj.xmeans), _ptr(obj.xscales),
                                    _ptr(R) if nlv else None, _ptr(P) if nlv else None, nlv, resid, _ptr(out),
                                    _ld(out))
        _lib.check(rc, "xfit")
    return out


def _bang_mat(X, name):
    if not (isinstance(X, np.ndarray) and X.ndim == 2 and X.dtype == np.float64 and X.flags.f_contiguous
            and X.flags.writeable):
        raise TypeError(f"MethodError: {name} needs a writeable column-major float64 matrix (X::Matrix)")
    return X


def xfit(obj, X, *, nlv=None):
    """xfit(object::Plsr, X; nlv = nothing), `/root/reference/src/xfit.jl:33-35`: X_fit in the original scale."""
    X = _fmat(X)
    return _xfit(obj, X, nlv, 0, _out_empty(X.shape))


def xfit_bang(obj, X, *, nlv=None):
    """xfit!(object, X::Matrix; nlv) (xfit.jl:37-56): X is overwritten by X_fit."""
    X = _bang_mat(X, "xfit!")
    return _xfit(obj, X, nlv, 0, X)


def xresid(obj, X, *, nlv=None):
    """xresid(object, X; nlv = nothing) (xfit.jl:88-90): E = X - X_fit."""
    X = _fmat(X)
    return _xfit(obj, X, nlv, 1, _out_empty(X.shape))


def xresid_bang(obj, X, *, nlv=None):
    """xresid!(object, X::Matrix; nlv) (xfit.jl:92-99): X is overwritten by the residuals."""
    X = _bang_mat(X, "xresid!")
    return _xfit(obj, X, nlv, 1, X)


def coef(obj, *, nlv=None):
    """coef(object; nlv = nothing) -> (B = B, int = int) (:207-217)."""
    a = obj.T.shape[1]
    nlv = a if nlv is None else min(int(nlv), a)
    nlv = max(nlv, 0)
    p, q = obj.R.shape[0], obj.C.shape[0]
    B = np.empty((p, q), order="F")
    intercept = np.empty((1, q), order="F")
    R, Cm = np.asfortranarray(obj.R), np.asfortranarray(obj.C)
    rc = _lib.lib().jcb200_coef(_ptr(R) if a else None, _ptr(Cm) if a else None, _ptr(obj.xmeans),
                                _ptr(obj.xscales), _ptr(obj.ymeans), _ptr(obj.yscales), p, q, nlv,
                                _ptr(B), _ptr(intercept))
    _lib.check(rc, "coef")
    return CoefResult(B, intercept)


def predict(obj, X, *, nlv=None):
    """predict(object, X; nlv = nothing) -> (pred = pred,) (:226-238): `nlv` may be an int or any
    collection; it is widened to the contiguous range max(0, min):min(a, max) as the reference does,
    and a single value is unwrapped to a bare matrix."""
    X = _fmat(X)
    a = obj.T.shape[1]
    if nlv is None:
        k_lo = k_hi = a
    else:
        ks = np.atleast_1d(np.asarray(nlv))
        k_lo, k_hi = max(0, int(ks.min())), min(a, int(ks.max()))
    m, p = X.shape
    q = obj.C.shape[0]
    if p != obj.xmeans.shape[0]:
        raise ValueError(f"DimensionMismatch: X has {p} columns, the model has {obj.xmeans.shape[0]}")
    nk = k_hi - k_lo + 1
    if nk <= 0:
        return PredResult([])
    preds = [_out_empty((m, q)) for _ in range(nk)]
    if m > 0:
        arr = (C.c_void_p * nk)(*[pm.ctypes.data for pm in preds])
        R, Cm = np.asfortranarray(obj.R), np.asfortranarray(obj.C)
        rc = _lib.lib().jcb200_predict_sweep(_ptr(X), _ld(X), m, p, q, _ptr(R) if a else None,
                                             _ptr(Cm) if a else None, a, _ptr(obj.xmeans),
                                             _ptr(obj.xscales), _ptr(obj.ymeans), _ptr(obj.yscales),
                                             k_lo, k_hi, arr)
        _lib.check(rc, "predict")
    return PredResult(preds[0] if nk == 1 else preds)


def summary(obj, X):
    """Base.summary(object::Plsr, X) (:246-260): `(explvarx = table(nlv, var, pvar, cumpvar),)` for the
    X the model was fitted on; a pandas DataFrame when pandas is importable, else a dict of arrays."""
    X = _fmat(X)
    n, a = obj.T.shape
    p = X.shape[1]
    if X.shape[0] != n:
        raise ValueError(f"DimensionMismatch: X has {X.shape[0]} rows, the model was fitted on {n}")
    xvar, pvar, cum = np.empty(a), np.empty(a), np.empty(a)
    P = np.asfortranarray(obj.P)
    rc = _lib.lib().jcb200_summary(_ptr(X), _ld(X), n, p, _ptr(obj.xmeans), _ptr(obj.xscales),
                                   _ptr(np.ascontiguousarray(obj.weights)), _ptr(P) if a else None,
                                   _ptr(obj.TT) if a else None, a, _ptr(xvar), _ptr(pvar), _ptr(cum))
    _lib.check(rc, "summary")
    out = {"nlv": np.arange(1, a + 1), "var": xvar, "pvar": pvar, "cumpvar": cum}
    try:
        import pandas as pd
        out = pd.DataFrame(out)
    except Exception:
        pass
    return namedtuple("SummaryResult", ["explvarx"])(out)
