"""xfit / xresid at m = 1e6, p = 500, nlv = 25: device time of K5 + K10 (phase 'scores'),
end-to-end time through the C ABI, algorithmic bytes against the measured HBM peak.  One JSON line."""
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
    m, n, p, q, nlv = int(os.environ.get("M", 1_000_000)), 100_000, 500, 10, 25
    X, Y = synth.synth_matrix(1, n, p), synth.synth_matrix(2, n, q)
    fm = jc.plskern(X, Y, nlv=nlv)
    Xn = synth.synth_matrix(4, m, p)
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    out = {}
    for name, fn, by in [("xfit", jc.xfit, 8.0 * (2 * m * p + m * nlv * 2)),
                         ("xresid", jc.xresid, 8.0 * (3 * m * p + m * nlv * 2))]:
        for _ in range(2):
            res = fn(fm, Xn, nlv=nlv)
        t0 = time.perf_counter()
        res = fn(fm, Xn, nlv=nlv)
        e2e = time.perf_counter() - t0
        tm = _lib.last_timings()
        ref = getattr(oracle, name)(fm, Xn[:2000], nlv=nlv)
        err = float(np.max(np.abs(res[:2000] - ref)) / np.max(np.abs(Xn[:2000])))
        out[name] = {"device_ms_k5_plus_k10": tm["scores"], "h2d_ms": tm["h2d"], "d2h_ms": tm["d2h"],
                     "e2e_ms": e2e * 1e3, "algorithmic_GB": by * 1e-9,
                     "hbm_frac_of_measured": by / (tm["scores"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "max_err_vs_oracle_rel_to_max_x": err}
    print(json.dumps({"workload": f"xfit/xresid m={m} p={p} nlv={nlv} (model fitted on n={n})", "results": out}))


if __name__ == "__main__":
    main()
