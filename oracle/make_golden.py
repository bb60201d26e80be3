"""Writes tests/golden/*.npz from the NumPy oracle — TEST INFRASTRUCTURE ONLY.

Run from the repo root:  python -m oracle.make_golden
Inputs are NOT stored: they are regenerated bit-exactly from the counter-based
generator (oracle/synth.py, SURVEY.md §8 d).  Outputs of the oracle are stored
whole when small and as a strided row sample (`rows`) when n is large.

PARITY UNPINNED: these vectors come from the restatement, not from a run of
the Julia reference (Julia is not installed here; the reference holds no
vectors of its own).
"""
import os

import numpy as np

from . import synth
from .plskern_ref import coef, plskern, predict, transform

CASES = {
    # name: n, p, q, nlv, m, uniform weights, scal
    "c1": dict(n=150, p=200, q=2, nlv=5, m=50, uniform=True, scal=False),
    "c1_wscal": dict(n=150, p=200, q=2, nlv=5, m=50, uniform=False, scal=True),
    "c2_cut": dict(n=20000, p=500, q=10, nlv=25, m=1000, uniform=True, scal=False),
    "c3_cut": dict(n=20000, p=1000, q=1, nlv=30, m=1000, uniform=False, scal=True),
    "c5_cut": dict(n=20000, p=500, q=10, nlv=50, m=2000, uniform=True, scal=False),
    "edge_odd": dict(n=37, p=5, q=3, nlv=10, m=7, uniform=False, scal=False),
}


def inputs(cfg):
    X = synth.synth_matrix(synth.SEED_X, cfg["n"], cfg["p"])
    Y = synth.synth_matrix(synth.SEED_Y, cfg["n"], cfg["q"])
    w = synth.synth_weights(cfg["n"], uniform=cfg["uniform"])
    Xnew = synth.synth_matrix(synth.SEED_N, cfg["m"], cfg["p"])
    return X, Y, w, Xnew


def run_case(cfg):
    X, Y, w, Xnew = inputs(cfg)
    fm = plskern(X, Y, w, nlv=cfg["nlv"], scal=cfg["scal"])
    a = fm.T.shape[1]
    rows = np.arange(0, cfg["n"], max(1, cfg["n"] // 256))
    out = dict(
        rows=rows, T_rows=fm.T[rows], T_colnorm=np.linalg.norm(fm.T, axis=0),
        P=fm.P, R=fm.R, W=fm.W, C=fm.C, TT=fm.TT, xmeans=fm.xmeans,
        xscales=fm.xscales, ymeans=fm.ymeans, yscales=fm.yscales,
        weights_rows=fm.weights[rows],
    )
    ks = sorted({0, 1, a // 2, a})         # coefficients at selected nlv
    Bs, ints = [], []
    for k in ks:
        B, i0 = coef(fm, nlv=k)
        Bs.append(B)
        ints.append(i0)
    out["ks"] = np.array(ks)
    out["B_ks"] = np.stack(Bs)             # (len(ks), p, q)
    out["int_ks"] = np.stack(ints)         # (len(ks), 1, q)
    mrows = np.arange(0, cfg["m"], max(1, cfg["m"] // 64))
    preds = predict(fm, Xnew, nlv=range(0, a + 1))
    out["mrows"] = mrows
    out["pred_all_rows"] = np.stack(preds)[:, mrows]   # (a+1, len(mrows), q)
    out["Tnew_rows"] = transform(fm, Xnew)[mrows]
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    gold = os.path.join(os.path.dirname(here), "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    for name, cfg in CASES.items():
        out = run_case(cfg)
        np.savez_compressed(os.path.join(gold, name + ".npz"),
                            cfg=np.array(repr(cfg)), **out)
        print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
