"""End-to-end times of the streaming host paths through the C ABI (host arrays in, host arrays out) at the
C5 shape: predict sweep nlv = 0:50, transform nlv = 50, xfit — m = 1e6 held-out rows, p = 500, q = 10.
Inputs and outputs are page-locked (the library's pool), so the numbers are PCIe-bound: the row-chunk
pipeline overlaps the two directions of the link.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jchemo_b200 as jc  # noqa: E402
import importlib  # noqa: E402
from jchemo_b200 import _lib  # noqa: E402
pk = importlib.import_module("jchemo_b200.plskern")   # the module (the package re-exports the function)
import oracle  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    m, n, p, q, nlv = int(os.environ.get("M", 1_000_000)), 200_000, 500, 10, 50
    X, Y = synth.synth_matrix(1, n, p), synth.synth_matrix(2, n, q)
    fm = jc.plskern(X, Y, nlv=nlv)
    Xn = pk._out_empty((m, p))                              # page-locked input
    step = 100_000
    for r0 in range(0, m, step):
        nr = min(step, m - r0)
        Xn[r0:r0 + nr] = synth.synth_matrix(4, nr, p, row0=r0, n_global=m)
    rows = np.arange(0, m, 9973)
    out = {}
    cases = [
        ("predict_sweep_0_50", lambda: jc.predict(fm, Xn, nlv=range(0, nlv + 1)).pred,
         8.0 * m * p, 8.0 * (nlv + 1) * m * q),
        ("transform_nlv50", lambda: jc.transform(fm, Xn, nlv=nlv), 8.0 * m * p, 8.0 * m * nlv),
        ("xfit_nlv25", lambda: jc.xfit(fm, Xn, nlv=25), 8.0 * m * p, 8.0 * m * p),
    ]
    for name, fn, bin_, bout in cases:
        for _ in range(2):
            res = fn()
            del res
        t0 = time.perf_counter()
        res = fn()
        e2e = time.perf_counter() - t0
        tm = _lib.last_timings()
        if name.startswith("predict"):
            want = oracle.predict(fm, Xn[rows], nlv=range(0, nlv + 1))
            err = max(float(np.linalg.norm(a[rows] - b) / np.linalg.norm(b)) for a, b in zip(res, want))
        elif name.startswith("transform"):
            want = oracle.transform(fm, Xn[rows], nlv=nlv)
            err = float(np.linalg.norm(res[rows] - want) / np.linalg.norm(want))
        else:
            want = oracle.xfit(fm, Xn[rows], nlv=25)
            err = float(np.linalg.norm(res[rows] - want) / np.linalg.norm(want))
        out[name] = {"e2e_ms": e2e * 1e3, "abi_total_ms": tm["total"], "h2d_ms": tm["h2d"], "d2h_ms": tm["d2h"],
                     "kernel_ms": tm["scores"], "h2d_GB": bin_ * 1e-9, "d2h_GB": bout * 1e-9,
                     "link_GBps_in_plus_out": (bin_ + bout) / (tm["total"] * 1e-3) * 1e-9,
                     "rel_err_vs_oracle_on_sample": err}
        del res
    print(json.dumps({"workload": f"streaming host paths, m={m} p={p} q={q} nlv={nlv} (model fitted on n={n}), "
                                  "page-locked host arrays", "results": out}))


if __name__ == "__main__":
    main()
