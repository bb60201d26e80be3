"""K1 alone on the C2 shape (joint [X | Y] buffer): per-launch CUDA-event time of the Gram kernel.  Environment
switches of the scheduler (JCB_STAGE_OVERHEAD, JCB_ZONE_MB, JCB_GRAM_JOINT) are read by the library at first use,
so every setting is its own process."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jchemo_b200 import device as dev, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--p", type=int, default=500)
ap.add_argument("--q", type=int, default=10)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
X, Y = dev.colmajor_empty_xy(a.n, a.p, a.q)
dev.fill_uniform(X, a.n, 1); dev.fill_uniform(Y, a.n, 2)
pivot = torch.empty(a.p + a.q + 1, dtype=torch.float64, device="cuda")
packed = torch.empty(dev.packed_len(a.p, a.q), dtype=torch.float64, device="cuda")
dev.pivot_dev(X, Y, a.n, pivot)
for _ in range(3):
    dev.gram_dev(X, Y, None, a.n, pivot, packed)
for _ in range(a.steps):
    dev.gram_dev(X, Y, None, a.n, pivot, packed)
g = _lib.gram_timings(a.steps)
ms = sum(g) / len(g)
fl = a.n * a.p * (a.p + 1) + 2.0 * a.n * a.p * a.q
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("JCB_")}, "n": a.n, "p": a.p, "q": a.q,
                  "k1_ms": round(ms, 4), "tflops": round(fl / ms * 1e-9, 2), "frac_of_37.145": round(fl / ms * 1e-9 / 37.145, 4)}))
