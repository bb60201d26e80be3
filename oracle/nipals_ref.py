"""NumPy restatement of the sibling NIPALS fit — TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/src/plsnipals.jl:37-97` (`plsnipals!`): explicit
deflation of X and Y by the score t, no Gram matrix, no r-recurrence.  It builds
the same `Plsr` as `plskern!` up to the sign of each latent variable and
rounding, and shares no recurrence with either the reference kernel algorithm
or the Gram form the CUDA path uses — an independent second opinion for the
(unpinned) oracle.
"""
import numpy as np

from .plskern_ref import (Plsr, center_bang, colmean, colstd, cscale_bang,
                          ensure_mat, mweight)


def plsnipals(X, Y, weights=None, *, nlv, scal=False):
    X = np.array(ensure_mat(X), dtype=np.float64, order="F", copy=True)
    Y = np.array(ensure_mat(Y), dtype=np.float64, order="F", copy=True)
    n, p = X.shape
    q = Y.shape[1]
    nlv = min(n, p, nlv)
    weights = mweight(np.ones(n) if weights is None else weights)
    xmeans = colmean(X, weights)
    ymeans = colmean(Y, weights)
    xscales = np.ones(p)
    yscales = np.ones(q)
    if scal:
        xscales[:] = colstd(X, weights)
        yscales[:] = colstd(Y, weights)
        cscale_bang(X, xmeans, xscales)
        cscale_bang(Y, ymeans, yscales)
    else:
        center_bang(X, xmeans)
        center_bang(Y, ymeans)
    T = np.empty((n, nlv), order="F")
    W = np.empty((p, nlv), order="F")
    P = np.empty((p, nlv), order="F")
    C = np.empty((q, nlv), order="F")
    TT = np.empty(nlv)
    for a in range(nlv):                                   # plsnipals.jl:70-92
        XtY = X.T @ (weights[:, None] * Y)
        if q == 1:
            w = XtY[:, 0].copy()
            w /= np.linalg.norm(w)
        else:
            w = np.linalg.svd(XtY, full_matrices=False)[0][:, 0].copy()
        t = X @ w
        dt = weights * t
        tt = np.dot(t, dt)
        zp = (X.T @ dt) / tt
        c = (Y.T @ dt) / tt
        X -= np.outer(t, zp)
        Y -= np.outer(t, c)
        P[:, a], T[:, a], W[:, a], C[:, a], TT[a] = zp, t, w, c, tt
    R = W @ np.linalg.inv(P.T @ W)                          # plsnipals.jl:95
    return Plsr(T, P, R, W, C, TT, xmeans, xscales, ymeans, yscales,
                weights, None)
