"""End-to-end fit from ordinary (pageable) host arrays, outputs pageable too — what a Julia caller hands over:
JCB_PINNED_MIN_BYTES is set huge so the Python mirror does not use the page-locked pool.  One JSON line."""
import json
import os
import sys
import time

os.environ["JCB_PINNED_MIN_BYTES"] = str(1 << 62)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import jchemo_b200 as jc  # noqa: E402

n, p, q, nlv = 1_000_000, 500, 10, 25
rng = np.random.default_rng(0)
X = np.asfortranarray(rng.random((n, p)))
Y = np.asfortranarray(rng.random((n, q)))
for _ in range(2):
    fm = jc.plskern(X, Y, nlv=nlv)
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    fm = jc.plskern(X, Y, nlv=nlv)
    ts.append(time.perf_counter() - t0)
print(json.dumps({"workload": f"plskern n={n} p={p} q={q} nlv={nlv}, pageable inputs and outputs",
                  "threads_env": os.environ.get("JCB_STAGE_THREADS"), "cpu_count": os.cpu_count(),
                  "e2e_ms_min": min(ts) * 1e3, "e2e_ms_mean": sum(ts) / len(ts) * 1e3, "phases_ms": jc.last_timings()}))
