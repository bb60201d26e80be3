"""Single-process multi-GPU fit (jcb200_init_multi): parity vs the oracle on a cut, then end-to-end time
of the C2 fit from pinned host arrays.  Usage: python bench/multigpu_inproc.py NGPU"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jchemo_b200 as jc
import oracle
from oracle import synth
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
jc.init_multi(list(range(ng)))
assert jc.lib().jcb200_device_count() == ng
n, p, q, nlv = 300_001, 120, 3, 10
X = synth.synth_matrix(1, n, p); Y = synth.synth_matrix(2, n, q) + X[:, :q]
w = synth.synth_weights(n, uniform=False)
fm = jc.plskern(X, Y, w, nlv=nlv, scal=True)
ref = oracle.plskern(X, Y, w, nlv=nlv, scal=True)
s = oracle.sign_align(ref, fm)
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
errs = {"T": rel(fm.T * s, ref.T), "R": rel(fm.R * s, ref.R), "B": rel(jc.coef(fm).B, oracle.coef(ref)[0]),
        "xmeans": rel(fm.xmeans, ref.xmeans), "weights": rel(fm.weights, ref.weights)}
Xb, Yb = X.copy(order="F"), Y.copy(order="F")
fmb = jc.plskern_bang(Xb, Yb, w, nlv=nlv, scal=True)
Xr, Yr = X.copy(order="F"), Y.copy(order="F"); oracle.plskern_bang(Xr, Yr, w, nlv=nlv, scal=True)
errs["writeback_X"] = rel(Xb, Xr)
ok = all(v < 1e-10 for v in errs.values())
if os.environ.get("JCB_SKIP_E2E"):
    print(json.dumps({"ngpu": ng, "parity_vs_oracle": errs, "parity_ok": ok}))
    sys.exit(0 if ok else 1)
# ---- C2 end to end from pinned host arrays
N, P, Q, NLV = 1_000_000, 500, 10, 25
hX = torch.empty((P, N), dtype=torch.float64).pin_memory(); hY = torch.empty((Q, N), dtype=torch.float64).pin_memory()
hX.numpy()[:] = synth.synth_matrix(1, N, P).T; hY.numpy()[:] = synth.synth_matrix(2, N, Q).T
Xh, Yh = hX.numpy().T, hY.numpy().T
for _ in range(3):
    f = jc.plskern(Xh, Yh, nlv=NLV)
t0 = time.perf_counter()
for _ in range(5):
    f = jc.plskern(Xh, Yh, nlv=NLV)
dt = (time.perf_counter() - t0) / 5
print(json.dumps({"ngpu": ng, "parity_vs_oracle": errs, "parity_ok": ok, "c2_e2e_fit_seconds": dt,
                  "device0_total_ms": jc.last_timings()["total"]}))
sys.exit(0 if ok else 1)
