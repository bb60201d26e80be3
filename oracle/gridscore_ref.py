"""NumPy restatement of `gridscorelv` and the regression scores — TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/src/gridscore.jl:167-221` (branch `pars === nothing`: one fit with
`nlv = maximum(nlv)`, one `predict` over the clamped contiguous range, `score(pred[i], Y)` per nlv) and
`/root/reference/src/scores.jl`: residreg :241, msep :155-158, rmsep :268, ssr :426-429, bias :25-28,
sep :400, r2 :190-195, rpd :332-335.  Returns a dict of columns like the reference's DataFrame
(`nlv`, `y1`..`yq`).  PARITY UNPINNED (see oracle/__init__.py).
"""
import numpy as np

from .plskern_ref import ensure_mat, plskern, predict


def residreg(pred, Y):
    return ensure_mat(Y) - pred


def msep(pred, Y):
    return np.mean(residreg(pred, Y) ** 2, axis=0).reshape(1, -1)


def rmsep(pred, Y):
    return np.sqrt(msep(pred, Y))


def ssr(pred, Y):
    return np.sum(residreg(pred, Y) ** 2, axis=0).reshape(1, -1)


def bias(pred, Y):
    return (-np.mean(residreg(pred, Y), axis=0)).reshape(1, -1)


def sep(pred, Y):
    return np.sqrt(msep(pred, Y) - bias(pred, Y) ** 2)


def r2(pred, Y):
    Y = ensure_mat(Y)
    M = np.tile(Y.mean(axis=0), (Y.shape[0], 1))
    return 1.0 - msep(pred, Y) / msep(M, Y)


def rpd(pred, Y):
    Y = ensure_mat(Y)
    return Y.std(axis=0).reshape(1, -1) / rmsep(pred, Y)


SCORES = dict(msep=msep, rmsep=rmsep, ssr=ssr, bias=bias, sep=sep, r2=r2, rpd=rpd)


def gridscorelv(Xtrain, Ytrain, X, Y, *, score, nlv, fun=plskern, **kwargs):
    Xtrain, Ytrain, X, Y = map(ensure_mat, (Xtrain, Ytrain, X, Y))
    p = Xtrain.shape[1]
    ks = np.atleast_1d(np.asarray(nlv))
    lo, hi = max(0, int(ks.min())), min(p, int(ks.max()))            # gridscore.jl:170-173
    fm = fun(Xtrain, Ytrain, nlv=hi, **kwargs)                       # :179
    pred = predict(fm, X, nlv=range(lo, hi + 1))                     # :180
    if not isinstance(pred, list):
        pred = [pred]                                                # :181
    fscore = SCORES[score] if isinstance(score, str) else score
    res = np.vstack([fscore(pr, Y) for pr in pred])                  # :183-185
    out = {"nlv": np.arange(lo, lo + len(pred))}
    for j in range(res.shape[1]):
        out[f"y{j + 1}"] = res[:, j]
    return out


def rmrow(X, s):
    """utility.jl rmrow: X without the rows s."""
    keep = np.ones(X.shape[0], dtype=bool)
    keep[np.asarray(s)] = False
    return X[keep]


def gridcvlv(X, Y, *, segm, score, nlv, fun=plskern, **kwargs):
    """gridcv.jl:187-228 (pars === nothing): for every repetition i and segment s, gridscorelv on
    (rmrow(X, s), rmrow(Y, s)) -> (X[s, :], Y[s, :]); `res_rep` has one row per (repl, segm, nlv), `res`
    is the mean over repetitions and segments per nlv.  Indices are zero-based here."""
    X, Y = ensure_mat(X), ensure_mat(Y)
    q = Y.shape[1]
    rows = []
    for i, listsegm in enumerate(segm):
        for j, s in enumerate(listsegm):
            s = np.asarray(s)
            z = gridscorelv(rmrow(X, s), rmrow(Y, s), X[s], Y[s], score=score, nlv=nlv, fun=fun, **kwargs)
            for t in range(len(z["nlv"])):
                rows.append([i + 1, j + 1, z["nlv"][t]] + [z[f"y{c + 1}"][t] for c in range(q)])
    arr = np.array(rows, dtype=float)
    res_rep = {"repl": arr[:, 0].astype(int), "segm": arr[:, 1].astype(int), "nlv": arr[:, 2].astype(int)}
    for c in range(q):
        res_rep[f"y{c + 1}"] = arr[:, 3 + c]
    ks = np.unique(res_rep["nlv"])
    res = {"nlv": ks}
    for c in range(q):
        res[f"y{c + 1}"] = np.array([arr[res_rep["nlv"] == k, 3 + c].mean() for k in ks])
    return res, res_rep


def locwlv(Xtrain, Ytrain, X, *, listnn, listw=None, nlv, fun=plskern, **kwargs):
    """locwlv.jl:9-48 — one fit per row of X on its neighbours, predictions for the clamped nlv range;
    zero-based neighbour indices.  Returns a list of m x q matrices (a single matrix for one nlv)."""
    Xtrain, Ytrain, X = ensure_mat(Xtrain), ensure_mat(Ytrain), ensure_mat(X)
    p, m, q = Xtrain.shape[1], X.shape[0], Ytrain.shape[1]
    ks = np.atleast_1d(np.asarray(nlv))
    lo, hi = max(0, int(ks.min())), min(p, int(ks.max()))                 # :14
    nk = hi - lo + 1
    zpred = np.empty((m, q, nk))
    for i in range(m):                                                    # :18
        s = np.atleast_1d(np.asarray(listnn[i]))
        zY = Ytrain[s]
        if q == 1 and np.unique(zY).size == 1:                            # :24-28
            zpred[i, :, :] = zY[0, 0]
            continue
        wi = None if listw is None else listw[i]
        fm = fun(Xtrain[s], zY, wi, nlv=hi, **kwargs)                     # :30-34
        for a in range(nk):
            zpred[i, :, a] = predict(fm, X[i:i + 1], nlv=lo + a)          # :36-38
    out = [zpred[:, :, a] for a in range(nk)]
    return out[0] if nk == 1 else out
