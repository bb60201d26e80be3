"""Debug: per-stage clock64 stamps of K1 (needs a -DJCB_K1_TRACE build selected with JCB_LIB)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jchemo_b200 as jc
from jchemo_b200 import device as dev, sharded
n, p, q, nlv = 1_000_000, 500, 10, 25
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
X = dev.colmajor_empty(n, p); Y = dev.colmajor_empty(n, q)
dev.fill_uniform(X, n, 1); dev.fill_uniform(Y, n, 2)
model = dev.DeviceModel(n, p, q, nlv)
for _ in range(2):
    sharded.fit_sharded(X, Y, None, n, model)
torch.cuda.synchronize()
raw = C.CDLL(os.environ["JCB_LIB"])
N = 4 * 64 * 16 * 3 + 256
buf = (C.c_longlong * N)()
raw.jcb200_debug_trace(buf, N)
allv = np.frombuffer(buf, dtype=np.int64)
t = allv[:4 * 64 * 16 * 3].reshape(4, 64, 16, 3)
iss = allv[4 * 64 * 16 * 3:].reshape(4, 64)
for cta in (0, 1, 2, 3):
    print("CTA", cta)
    a = t[cta]
    t0 = a[a > 0].min()
    for st in range(4, 24):
        w = a[st]
        wait = w[:, 1] - w[:, 0]; comp = w[:, 2] - w[:, 1]
        print(f" st {st:2d} start {int(w[:,0].min()-t0):8d} spread_in {int(w[:,0].max()-w[:,0].min()):6d} "
              f"wait[min/med/max] {int(wait.min()):6d} {int(np.median(wait)):6d} {int(wait.max()):6d} "
              f"comp[min/med/max] {int(comp.min()):6d} {int(np.median(comp)):6d} {int(comp.max()):6d} "
              f"end_spread {int(w[:,2].max()-w[:,2].min()):6d}")
    for st in range(8, 20):
        w = a[st]
        first_need = w[:, 0].min(); landed_by = w[:, 1].min()
        print(f"   st {st}: issued {int(iss[cta, st]-t0):8d}  first warp needs it {int(first_need-t0):8d}  first warp got it {int(landed_by-t0):8d}  "
              f"issue->got {int(landed_by-iss[cta, st]):6d}  lead {int(first_need-iss[cta, st]):6d}  last release of st-2 {int(a[st-2][:,2].max()-t0):8d}")
    print(" per-warp wait of stage 10:", (a[10,:,1]-a[10,:,0]).tolist())
    print(" per-warp comp of stage 10:", (a[10,:,2]-a[10,:,1]).tolist())
