"""Batched tiny fits (locwlv): m query rows, k neighbours each, one kernel launch.
Prints one JSON line: device time of the batched kernel, end-to-end time through the C-ABI, and the oracle's
loop of fits on a bounded sample of the same queries."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jchemo_b200 as jc  # noqa: E402
from jchemo_b200 import _lib  # noqa: E402
import oracle  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    ntr, p, q, m, k, nlv = 10000, 500, 1, int(os.environ.get("M", 2000)), int(os.environ.get("K", 200)), 15
    Xtr = synth.synth_matrix(1, ntr, p)
    Ytr = synth.synth_matrix(2, ntr, q) + Xtr[:, :q]
    X = synth.synth_matrix(4, m, p)
    rng = np.random.default_rng(1)
    listnn = [np.sort(rng.choice(ntr, size=k, replace=False)) for _ in range(m)]
    listw = [0.1 + rng.random(k) for _ in range(m)]
    for _ in range(2):
        res = jc.locwlv(Xtr, Ytr, X, listnn=listnn, listw=listw, nlv=range(0, nlv + 1))
    t0 = time.perf_counter()
    res = jc.locwlv(Xtr, Ytr, X, listnn=listnn, listw=listw, nlv=range(0, nlv + 1))
    e2e = time.perf_counter() - t0
    tm = _lib.last_timings()
    ms = 40
    t0 = time.perf_counter()
    ref = oracle.locwlv(Xtr, Ytr, X[:ms], listnn=listnn[:ms], listw=listw[:ms], nlv=range(0, nlv + 1))
    cpu = (time.perf_counter() - t0) / ms
    err = max(float(np.max(np.abs(res.pred[a][:ms] - ref[a])) / max(np.max(np.abs(ref[a])), 1e-300))
              for a in range(nlv + 1))
    flop = m * nlv * 4.0 * k * p
    print(json.dumps({"workload": f"locwlv ntr={ntr} p={p} q={q} m={m} k={k} nlv=0:{nlv}",
                      "kernel_ms": tm["scores"], "h2d_ms": tm["h2d"], "abi_total_ms": tm["total"],
                      "python_e2e_ms": e2e * 1e3, "fits_per_s_kernel": m / (tm["scores"] * 1e-3),
                      "kernel_gflops": flop / (tm["scores"] * 1e-3) / 1e9,
                      "oracle_ms_per_fit": cpu * 1e3, "oracle_fits_per_s": 1.0 / cpu, "max_rel_err_vs_oracle": err}))


if __name__ == "__main__":
    main()
