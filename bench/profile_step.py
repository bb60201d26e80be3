"""Short driver for ncu: W warm-up fits + K timed device-resident fits of config C2 (no e2e, no CPU arm)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jchemo_b200 as jc
from jchemo_b200 import device as dev, sharded, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--p", type=int, default=500)
ap.add_argument("--q", type=int, default=10)
ap.add_argument("--nlv", type=int, default=25)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--weighted", action="store_true")
ap.add_argument("--scal", action="store_true")
a = ap.parse_args()
torch.cuda.set_device(0)
dev.init(0); dev.use_current_stream()
X, Y = dev.colmajor_empty_xy(a.n, a.p, a.q)      # the layout bench.py uses: Y right behind X
dev.fill_uniform(X, a.n, 1); dev.fill_uniform(Y, a.n, 2)
w = None
if a.weighted:
    w = torch.empty((1, dev.even_up(a.n)), dtype=torch.float64, device="cuda")
    dev.fill_uniform(w, a.n, 3); w = (w + 0.5).reshape(-1)
model = dev.DeviceModel(a.n, a.p, a.q, a.nlv)
for _ in range(a.warmup):
    sharded.fit_sharded(X, Y, w, a.n, model, scal=a.scal)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    sharded.fit_sharded(X, Y, w, a.n, model, scal=a.scal)
e1.record(); torch.cuda.synchronize()
print("ms_per_fit", e0.elapsed_time(e1) / a.steps, "gram_ms", _lib.gram_timings(a.steps), dev.sync_timings())
