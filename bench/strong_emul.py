"""Per-rank work of a STRONG-scaled C2 fit (n = 1e6 rows in total) on ONE GPU: the step a rank of an N-GPU
job runs on its n/N rows, without the exchange.  Shows the Amdahl terms (replicated K3/K4, launch gaps)
before GPU-minutes are spent on N real GPUs.  One JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jchemo_b200 import device as dev, sharded, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--p", type=int, default=500)
ap.add_argument("--q", type=int, default=10)
ap.add_argument("--nlv", type=int, default=25)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
out = {}
for N in (1, 2, 4, 8):
    nl = a.n // N
    X, Y = dev.colmajor_empty_xy(nl, a.p, a.q) if os.environ.get("JCB_XY", "1") != "0" else (dev.colmajor_empty(nl, a.p), dev.colmajor_empty(nl, a.q))
    dev.fill_uniform(X, nl, 1, 0, a.n); dev.fill_uniform(Y, nl, 2, 0, a.n)
    model = dev.DeviceModel(nl, a.p, a.q, a.nlv)
    pivot = torch.empty(a.p + a.q + 1, dtype=torch.float64, device="cuda")
    packed = torch.empty(dev.packed_len(a.p, a.q), dtype=torch.float64, device="cuda")
    step = lambda: sharded.fit_sharded(X, Y, None, nl, model, pivot=pivot, packed=packed)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    g = _lib.gram_timings(a.steps)
    ph = dev.sync_timings()
    out[f"N{N}"] = {"rows": nl, "ms_per_step": ms, "gram_ms": sum(g) / len(g), "phases": ph,
                    "sum_phases": sum(v for k, v in ph.items() if k != "total"),
                    "strong_eff_vs_N1": None}
    del X, Y, model
base = out["N1"]["ms_per_step"]
for N in (1, 2, 4, 8):
    out[f"N{N}"]["strong_eff_vs_N1"] = base / (N * out[f"N{N}"]["ms_per_step"])
print(json.dumps(out))
