"""NumPy restatement of Jchemo's kernel-PLS path — TEST INFRASTRUCTURE ONLY.

Follows, line by line, `/root/reference/src/plskern.jl`:
  struct Plsr ............ :1-14
  plskern ................ :106-110
  plskern! ............... :112-178   (Dayal & MacGregor improved kernel #1)
  transform .............. :187-195
  coef ................... :207-217
  predict ................ :226-238
and the helpers of `/root/reference/src/utility.jl`:
  center! :76-81 · colmean :193-195 · colstd :262-264 · colvar :312-323 ·
  cscale / cscale! :476-487 · ensure_mat :544-548 · mweight :715-723.

PARITY UNPINNED (see oracle/__init__.py): no reference golden vectors exist.
The dense arithmetic that Julia delegates to LinearAlgebra → OpenBLAS/LAPACK
(gemv/gemm/dot/nrm2/gesdd; docs/Manifest.toml pins julia 1.8.5 +
OpenBLAS_jll 0.3.20+0) is delegated here to NumPy → OpenBLAS/LAPACK; the thin
SVD is `np.linalg.svd(full_matrices=False)` = LAPACK gesdd, the routine Julia's
default `svd` calls.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class Plsr:
    """plskern.jl:1-14 — 12 positional fields, reference order."""
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


# ---------------------------------------------------------------- utility.jl
def ensure_mat(X):
    """utility.jl:544-548 — vector → n×1 matrix, number → 1×1."""
    X = np.asarray(X, dtype=np.float64) if not isinstance(X, np.ndarray) else X
    if X.ndim == 0:
        return X.reshape(1, 1)
    if X.ndim == 1:
        return X.reshape(-1, 1)
    return X


def mweight(w):
    """utility.jl:715-723 — copy, convert to Float64, divide by the sum."""
    zw = np.array(w, dtype=np.float64).reshape(-1).copy()
    zw /= zw.sum()
    return zw


def colmean(X, w):
    """utility.jl:195 — vec(mweight(w)' * X): a gemv, re-normalising w."""
    return mweight(w) @ ensure_mat(X)


def colvar(X, w):
    """utility.jl:314-323 — two-pass, uncorrected weighted variance."""
    X = ensure_mat(X)
    w = mweight(w)
    z = colmean(X, w)
    for j in range(X.shape[1]):
        d = X[:, j] - z[j]
        z[j] = np.dot(w, d * d)
    return z


def colstd(X, w):
    """utility.jl:264."""
    return np.sqrt(colvar(X, mweight(w)))


def center_bang(X, v):
    """utility.jl:76-81 — column loop, in place."""
    for j in range(X.shape[1]):
        X[:, j] = X[:, j] - v[j]


def cscale_bang(X, u, v):
    """utility.jl:482-487 — subtract then divide, per element, in place."""
    for j in range(X.shape[1]):
        X[:, j] = (X[:, j] - u[j]) / v[j]


def cscale(X, u, v):
    """utility.jl:476-480."""
    zX = np.array(ensure_mat(X), dtype=np.float64, order="F", copy=True)
    cscale_bang(zX, u, v)
    return zX


# ---------------------------------------------------------------- plskern.jl
def plskern(X, Y, weights=None, *, nlv, scal=False):
    """plskern.jl:106-110 — copies X and Y, caller's arrays untouched."""
    X = np.array(ensure_mat(X), dtype=np.float64, order="F", copy=True)
    Y = np.array(ensure_mat(Y), dtype=np.float64, order="F", copy=True)
    return plskern_bang(X, Y, weights, nlv=nlv, scal=scal)


def plskern_bang(X, Y, weights=None, *, nlv, scal=False):
    """plskern.jl:112-178 — in place: X and Y leave centred (and scaled)."""
    n, p = X.shape
    q = Y.shape[1]
    nlv = min(n, p, nlv)                                   # :116
    if weights is None:
        weights = np.ones(n)                               # default, :112
    weights = mweight(weights)                             # :117
    xmeans = colmean(X, weights)                           # :118
    ymeans = colmean(Y, weights)                           # :119
    xscales = np.ones(p)                                   # :120
    yscales = np.ones(q)                                   # :121
    if scal:                                               # :122-126
        xscales[:] = colstd(X, weights)
        yscales[:] = colstd(Y, weights)
        cscale_bang(X, xmeans, xscales)
        cscale_bang(Y, ymeans, yscales)
    else:                                                  # :128-129
        center_bang(X, xmeans)
        center_bang(Y, ymeans)
    XtY = X.T @ (weights[:, None] * Y)                     # :131-132
    T = np.empty((n, nlv), order="F")                      # :135-147
    W = np.empty((p, nlv), order="F")
    P = np.empty((p, nlv), order="F")
    R = np.empty((p, nlv), order="F")
    C = np.empty((q, nlv), order="F")
    TT = np.empty(nlv)
    for a in range(nlv):                                   # :149
        if q == 1:                                         # :150-152
            w = XtY[:, 0].copy()
            w /= np.linalg.norm(w)
        else:                                              # :154
            w = np.linalg.svd(XtY, full_matrices=False)[0][:, 0].copy()
        r = w.copy()                                       # :156
        for j in range(a):                                 # :157-161 (classical GS)
            r -= np.dot(w, P[:, j]) * R[:, j]
        t = X @ r                                          # :162
        dt = weights * t                                   # :163
        tt = np.dot(t, dt)                                 # :164
        c = (XtY.T @ r) / tt                               # :165-166
        zp = X.T @ dt                                      # :167
        XtY -= np.outer(zp, c)                             # :168
        P[:, a] = zp / tt                                  # :169
        T[:, a] = t                                        # :170
        W[:, a] = w                                        # :171
        R[:, a] = r                                        # :172
        C[:, a] = c                                        # :173
        TT[a] = tt                                         # :174
    return Plsr(T, P, R, W, C, TT, xmeans, xscales, ymeans, yscales,
                weights, None)                             # :176-177


def transform(obj, X, *, nlv=None):
    """plskern.jl:187-195."""
    X = ensure_mat(X)
    a = obj.T.shape[1]
    nlv = a if nlv is None else min(nlv, a)
    return cscale(X, obj.xmeans, obj.xscales) @ obj.R[:, :nlv]


def coef(obj, *, nlv=None):
    """plskern.jl:207-217 — returns (B, int); nlv = 0 gives zeros / ymeans'."""
    a = obj.T.shape[1]
    nlv = a if nlv is None else min(nlv, a)
    beta = obj.C[:, :nlv].T
    B = ((1.0 / obj.xscales)[:, None] * obj.R[:, :nlv]) @ beta * obj.yscales[None, :]
    intercept = obj.ymeans[None, :] - obj.xmeans[None, :] @ B
    return B, intercept


def predict(obj, X, *, nlv=None):
    """plskern.jl:226-238 — nlv widened to the contiguous range min:max, clamped
    to 0:a; a single value is unwrapped to a bare matrix."""
    X = ensure_mat(X)
    a = obj.T.shape[1]
    if nlv is None:
        ks = [a]
    else:
        ks_req = np.atleast_1d(np.asarray(nlv))
        ks = list(range(max(0, int(ks_req.min())), min(a, int(ks_req.max())) + 1))
    pred = []
    for k in ks:
        B, intercept = coef(obj, nlv=k)
        pred.append(intercept + X @ B)
    return pred[0] if len(pred) == 1 else pred


def summary(obj, X):
    """plskern.jl:246-260 — explained X-variance table (nlv, var, pvar, cumpvar)."""
    X = ensure_mat(X)
    n, nlv = obj.T.shape
    Xs = cscale(X, obj.xmeans, obj.xscales)
    sstot = np.sum(obj.weights @ (Xs ** 2))                # :252
    tt_adj = np.sum(obj.P ** 2, axis=0) * obj.TT           # :254
    pvar = tt_adj / sstot
    return {"nlv": np.arange(1, nlv + 1), "var": tt_adj / n, "pvar": pvar, "cumpvar": np.cumsum(pvar)}


def xfit(obj, X, *, nlv=None):
    """xfit.jl:33-56 — X_fit = transform(X) P' in the original scale; nlv = 0 gives rows of xmeans."""
    X = np.array(ensure_mat(X), dtype=np.float64, order="F", copy=True)
    a = obj.T.shape[1]
    nlv = a if nlv is None else min(nlv, a)                # :39
    if nlv == 0:                                           # :41-45
        X[:, :] = obj.xmeans[None, :]
        return X
    X = transform(obj, X, nlv=nlv) @ obj.P[:, :nlv].T      # :47-48
    X = X / (1.0 / obj.xscales)[None, :]                   # :50 scale!(X, 1 ./ xscales) divides by its argument
    X = X - (-obj.xmeans)[None, :]                         # :52 center!(X, -xmeans)
    return X


def xresid(obj, X, *, nlv=None):
    """xfit.jl:88-99 — E = X - xfit(object, X; nlv)."""
    X = ensure_mat(X)
    return X - xfit(obj, X, nlv=nlv)


# ---------------------------------------------------------------- parity aid
def sign_align(ref, dev):
    """Per-LV signs s_a = sign(<W_ref[:,a], W_dev[:,a]>): columns a
    of W, R, P, T, C flip together; TT and B do not."""
    s = np.sign(np.sum(ref.W * dev.W, axis=0))
    s[s == 0] = 1.0
    return s
