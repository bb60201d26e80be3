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
