"""gridcvlv (K-fold CV over nlv = 0:25) at the README shape through the host API; one JSON line.
Reference cost model: K fits on (K-1)/K of the rows + K predict sweeps; here: one Gram pass + K solves +
one scoring pass."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jchemo_b200 as jc
from oracle import synth
n, p, q, nlv, K = int(os.environ.get("N", 1_000_000)), 500, 10, 25, 5
X = synth.synth_matrix(1, n, p); Y = synth.synth_matrix(2, n, q)
rng = np.random.default_rng(0)
idx = rng.permutation(n)
segm = [[np.sort(s) for s in np.array_split(idx, K)], [np.sort(s) for s in np.array_split(rng.permutation(n), K)]]
jc.gridcvlv(X[:20000], Y[:20000], segm=[[np.arange(0, 20000, 3)]], score="rmsep", nlv=range(0, nlv + 1))  # warm-up
t0 = time.perf_counter()
out = jc.gridcvlv(X, Y, segm=segm, score="rmsep", nlv=range(0, nlv + 1))
dt = time.perf_counter() - t0
ph = jc.last_timings()
print(json.dumps({"workload": f"gridcvlv n={n} p={p} q={q} nlv=0:{nlv} K={K} rep=2 (pageable host arrays)",
                  "seconds_total": dt, "last_rep_device_ms": ph["total"], "last_rep_phases_ms": ph,
                  "rmsep_y1_nlv": [float(v) for v in np.asarray(out.res["y1"])[:4]]}))
