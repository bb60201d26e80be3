"""NumPy restatement of the sibling ROSA-type fit — TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/src/plsrosa.jl:32-96` (`plsrosa!`, Liland et al. 2016): only Y is deflated; the score
t = X w is orthogonalised against the previous scores and w against the previous weights.  It yields the same
`Plsr` as `plskern!` for any number of responses (up to the sign of each latent variable and rounding) through
yet another recurrence — explicit projectors on T and W, no Gram-Schmidt on r, no deflation of X — a further
independent opinion for the (unpinned) oracle, next to NIPALS and SIMPLS.
"""
import numpy as np

from .plskern_ref import (Plsr, center_bang, colmean, colstd, cscale_bang,
                          ensure_mat, mweight)


def plsrosa(X, Y, weights=None, *, nlv, scal=False):
    X = np.array(ensure_mat(X), dtype=np.float64, order="F", copy=True)       # plsrosa.jl:26-30
    Y = np.array(ensure_mat(Y), dtype=np.float64, order="F", copy=True)
    n, p = X.shape                                                             # :34-36
    q = Y.shape[1]
    nlv = min(nlv, n, p)
    weights = mweight(np.ones(n) if weights is None else weights)             # :37
    xmeans = colmean(X, weights)                                               # :39-40
    ymeans = colmean(Y, weights)
    xscales = np.ones(p)
    yscales = np.ones(q)
    if scal:                                                                   # :43-51
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
    for a in range(nlv):                                                       # :65-93
        XtY = X.T @ (weights[:, None] * Y)                                     # :66
        if q == 1:
            w = XtY[:, 0].copy()
            w /= np.linalg.norm(w)
        else:
            w = np.linalg.svd(XtY, full_matrices=False)[0][:, 0].copy()        # :71
        t = X @ w                                                              # :73
        if a > 0:                                                              # :74-80
            z = T[:, :a]
            t = t - z @ np.linalg.inv(z.T @ (weights[:, None] * z)) @ (z.T @ (weights * t))
            z = W[:, :a]
            w = w - z @ (z.T @ w)
            w /= np.sqrt(np.dot(w, w))
        dt = weights * t
        tt = np.dot(t, dt)
        c = (Y.T @ dt) / tt                                                    # :83-84
        zp = (X.T @ dt) / tt                                                   # :85-86
        Y -= np.outer(t, c)                                                    # :87
        P[:, a], T[:, a], W[:, a], C[:, a], TT[a] = zp, t, w, c, tt
    R = W @ np.linalg.inv(P.T @ W)                                             # :94
    return Plsr(T, P, R, W, C, TT, xmeans, xscales, ymeans, yscales, weights, None)
