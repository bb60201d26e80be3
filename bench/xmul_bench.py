"""K5 / K6 roofline table (device resident, one GPU): every use of the streaming (X - mu)/sigma * M kernel with its
algorithmic bytes and flops, the bound max(bytes / HBM peak, flops / FP64 peak) and the fraction of it.
  fit scores   T = Xc R on the rows of the fit (centre-free when the pivot pass allows)   C2: nlv 25
  transform    new rows (always centred)                                                  nlv 25, 50
  predict      single k (coef + narrow GEMM) and the sweep 0:50 in one pass                C5
One JSON line; JCB_LIB selects an alternative build of the library for A/B runs."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jchemo_b200 import device as dev, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=1_000_000)
ap.add_argument("--p", type=int, default=500)
ap.add_argument("--q", type=int, default=10)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--tag", default="")
ap.add_argument("--nlvs", default="25,50")
a = ap.parse_args()
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
root = os.path.join(os.path.dirname(__file__), "..")
hbm = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
fp64 = json.load(open(os.path.join(root, "profiles", "fp64_peak_r01.json")))["dmma_tflops_burst"]
m, p, q = a.m, a.p, a.q
X = dev.colmajor_empty(m, p); Y = dev.colmajor_empty(m, q)
dev.fill_uniform(X, m, 1); dev.fill_uniform(Y, m, 2)
Xn = dev.colmajor_empty(m, p); dev.fill_uniform(Xn, m, 4)
out = {}


def run(name, fn, by, fl):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    t_hbm, t_fp = by / hbm * 1e-6, fl / fp64 * 1e-9
    bound = max(t_hbm, t_fp)
    out[name] = {"ms": round(ms, 4), "bound_ms": round(bound, 4), "bound": "hbm" if t_hbm >= t_fp else "fp64",
                 "frac_of_bound": round(bound / ms, 3), "GBps": round(by / ms * 1e-6, 1), "tflops": round(fl / ms * 1e-9, 2)}


for nlv in [int(v) for v in a.nlvs.split(',')]:
    model = dev.DeviceModel(m, p, q, nlv)
    pivot = torch.empty(p + q + 1, dtype=torch.float64, device="cuda")
    sharded.fit_sharded(X, Y, None, m, model, pivot=pivot)
    by_t, fl_t = 8.0 * (m * p + m * nlv), 2.0 * m * p * nlv
    run(f"fit_scores_nlv{nlv}", lambda: dev.scores_dev(X, m, model, pivot=pivot), by_t, fl_t)
    Tn = dev.colmajor_empty(m, nlv)
    run(f"transform_nlv{nlv}", lambda: dev.scores_dev(Xn, m, model, Tn), by_t, fl_t)
    if nlv == 50:
        pred = torch.empty((nlv + 1, q, m), dtype=torch.float64, device="cuda")
        run("predict_sweep_0_50", lambda: dev.predict_sweep_dev(Xn, m, model, 0, nlv, pred),
            8.0 * (m * p + (nlv + 1) * m * q), 2.0 * m * p * nlv + 2.0 * m * q * nlv)
        run("predict_single_k", lambda: dev.predict_sweep_dev(Xn, m, model, nlv, nlv, pred[:1]),
            8.0 * (m * p + m * q), 2.0 * m * p * q)
        del pred
    del model, Tn
print(json.dumps({"tag": a.tag, "lib": os.environ.get("JCB_LIB", "default"), "m": m, "p": p, "q": q,
                  "hbm_peak_gbs": hbm, "fp64_peak_tflops": fp64, "results": out}))
