"""Debug: per-phase clock64 totals of the LV loop (needs a -DJCB_K1_TRACE build selected with JCB_LIB)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jchemo_b200 import device as dev, sharded
n, p, q, nlv = 200_000, int(os.environ.get("P", 500)), int(os.environ.get("Q", 10)), int(os.environ.get("NLV", 25))
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
X = dev.colmajor_empty(n, p); Y = dev.colmajor_empty(n, q)
dev.fill_uniform(X, n, 1); dev.fill_uniform(Y, n, 2)
model = dev.DeviceModel(n, p, q, nlv)
for _ in range(2):
    sharded.fit_sharded(X, Y, None, n, model)
torch.cuda.synchronize()
raw = C.CDLL(os.environ["JCB_LIB"])
buf = (C.c_longlong * 24)()
raw.jcb200_debug_lv_trace(buf)
names = ["A: M partial, exchange, sum", "eigenvector: wait for the Z warps (ZF)", "eigenvector + v'Mv (warp 0)", "B: sum of partials + sync (after wait)", "C: r, gather, u, wait", "D: matvec + send", "D: wait", "tt, c", "-", "deflate + store",
         "B: w~ + sync", "B: dots + send", "B: wait"]
if os.environ.get("JCB_LV_LINEAR", "1") != "0":
    n0 = ["t0: partial M,Z + send", "t0: wait A", "t0: sum A + sync", "t0: eigenvector", "t0: wait D (builders' branch)",
          "t0: sum D + sync", "t0: tt, c, slices, deflate", "t0: loop top"]
    n1 = ["t32: partial M,Z + send", "t32: wait A", "t32: sum A + sync", "t32: Rho + send", "t32: wait G + Zeta (warp 1)",
          "t32: barrier after Zeta", "t32: small products + send", "t32: wait D .. loop top",
          "   of which: request first row + wait G", "   of which: Zeta rows of warp 1"]
    for i, nm in enumerate(n0 + n1):
        print(f"{nm:40s} {buf[i]:10d} cyc  {buf[i]/1965e3:8.3f} ms")
    print("total t0", sum(buf[:8]) / 1965e3, "ms; total t32", sum(buf[8:16]) / 1965e3, "ms")
    sys.exit(0)
tot = sum(buf[i] for i in range(13))
for i, nm in enumerate(names):
    print(f"{nm:40s} {buf[i]:10d} cyc  {buf[i]/1965e3:8.3f} ms  {100*buf[i]/max(tot,1):5.1f}%")
print("total", tot/1965e3, "ms")
