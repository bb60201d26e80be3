"""Where the phases of one end-to-end fit sit inside the call: run with JCB_DEBUG_TIMELINE=1 (the library prints every
phase occurrence, ms after the begin of the call, on stderr).  NDEV=k binds the library to k GPUs (in-process row
sharding; the phases are those of device 0)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jchemo_b200 as jc
from jchemo_b200 import device as dev
n, p, q, nlv = 1_000_000, 500, 10, 25
torch.cuda.set_device(0)
ndev = int(os.environ.get("NDEV", "1"))
if ndev > 1:
    jc.init_multi(list(range(ndev)))     # single process, rows sharded over ndev GPUs inside the call
else:
    dev.init(0)
X = torch.empty((p, n), dtype=torch.float64).pin_memory(); Y = torch.empty((q, n), dtype=torch.float64).pin_memory()
X.uniform_(); Y.uniform_()
Xn = X.numpy().T; Yn = Y.numpy().T
for i in range(3):
    t0 = time.perf_counter(); fm = jc.plskern(Xn, Yn, nlv=nlv); t1 = time.perf_counter()
    print("wall ms", (t1 - t0) * 1e3, file=sys.stderr)
