"""C5: predict / transform sweep nlv = 0:50 on m = 1e6 held-out rows (model p=500, q=10, nlv=50), device
resident; reports time, algorithmic bytes and HBM fraction as one JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jchemo_b200 import device as dev, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=1_000_000)
ap.add_argument("--n", type=int, default=200_000)
ap.add_argument("--p", type=int, default=500)
ap.add_argument("--q", type=int, default=10)
ap.add_argument("--nlv", type=int, default=50)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
X = dev.colmajor_empty(a.n, a.p); Y = dev.colmajor_empty(a.n, a.q)
dev.fill_uniform(X, a.n, 1); dev.fill_uniform(Y, a.n, 2)
model = dev.DeviceModel(a.n, a.p, a.q, a.nlv)
sharded.fit_sharded(X, Y, None, a.n, model)
Xn = dev.colmajor_empty(a.m, a.p); dev.fill_uniform(Xn, a.m, 4)
pred = torch.empty((a.nlv + 1, a.q, a.m), dtype=torch.float64, device="cuda")
Tn = dev.colmajor_empty(a.m, a.nlv)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
out = {}
for name, fn, by, fl in [
    ("predict_sweep_0_%d" % a.nlv, lambda: dev.predict_sweep_dev(Xn, a.m, model, 0, a.nlv, pred),
     8.0 * (a.m * a.p + (a.nlv + 1) * a.m * a.q), 2.0 * a.m * a.p * a.nlv + 2.0 * a.m * a.q * a.nlv),
    ("predict_single_k", lambda: dev.predict_sweep_dev(Xn, a.m, model, a.nlv, a.nlv, pred[:1]),
     8.0 * (a.m * a.p + a.m * a.q), 2.0 * a.m * a.p * a.q),
    ("transform_nlv%d" % a.nlv, lambda: dev.scores_dev(Xn, a.m, model, Tn),
     8.0 * (a.m * a.p + a.m * a.nlv), 2.0 * a.m * a.p * a.nlv),
]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    out[name] = {"ms": ms, "algorithmic_GB": by * 1e-9, "GBps": by / ms * 1e-6,
                 "hbm_frac_of_measured": by / ms * 1e-6 / peaks["hbm_gbs"], "tflops": fl / ms * 1e-9}
print(json.dumps({"config": vars(a), "hbm_peak_gbs": peaks["hbm_gbs"], "results": out}))
