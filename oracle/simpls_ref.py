"""NumPy restatement of the sibling SIMPLS fit — TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/src/plssimp.jl:28-88` (`plssimp!`, de Jong 1993, Table 1): the weight vector r is
the dominant left singular vector of X'DY projected on the orthogonal complement of the previous X-loadings;
scores are not normed.  For a single response (q = 1) SIMPLS and kernel PLS give the same regression
coefficients and proportional scores, through a different recurrence (explicit projector on P, no
Gram-Schmidt on r) — a third independent opinion for the (unpinned) oracle, next to NIPALS.
"""
import numpy as np

from .plskern_ref import (Plsr, center_bang, colmean, colstd, cscale_bang,
                          ensure_mat, mweight)


def plssimp(X, Y, weights=None, *, nlv, scal=False):
    X = np.array(ensure_mat(X), dtype=np.float64, order="F", copy=True)       # plssimp.jl:22-26
    Y = np.array(ensure_mat(Y), dtype=np.float64, order="F", copy=True)
    n, p = X.shape                                                             # :30-33
    q = Y.shape[1]
    nlv = min(nlv, n, p)
    weights = mweight(np.ones(n) if weights is None else weights)             # :34
    xmeans = colmean(X, weights)                                               # :35-36
    ymeans = colmean(Y, weights)
    xscales = np.ones(p)
    yscales = np.ones(q)
    if scal:                                                                   # :39-47
        xscales[:] = colstd(X, weights)
        yscales[:] = colstd(Y, weights)
        cscale_bang(X, xmeans, xscales)
        cscale_bang(Y, ymeans, yscales)
    else:
        center_bang(X, xmeans)
        center_bang(Y, ymeans)
    XtY = X.T @ (weights[:, None] * Y)                                         # :48-49
    T = np.empty((n, nlv), order="F")
    P = np.empty((p, nlv), order="F")
    R = np.empty((p, nlv), order="F")
    C = np.empty((q, nlv), order="F")
    TT = np.empty(nlv)
    for a in range(nlv):                                                       # :65-83
        if a == 0:
            tmp = XtY.copy()
        else:
            zP = P[:, :a]
            tmp = XtY - zP @ np.linalg.inv(zP.T @ zP) @ zP.T @ XtY             # :69-70
        r = np.linalg.svd(tmp, full_matrices=False)[0][:, 0].copy()            # :72
        t = X @ r                                                              # :73
        dt = weights * t
        tt = np.dot(t, dt)
        c = (XtY.T @ r) / tt                                                   # :76-77
        zp = X.T @ dt                                                          # :78
        P[:, a], T[:, a], R[:, a], C[:, a], TT[a] = zp / tt, t, r, c, tt       # :79-83
    return Plsr(T, P, R, R.copy(), C, TT, xmeans, xscales, ymeans, yscales, weights, None)   # :88
