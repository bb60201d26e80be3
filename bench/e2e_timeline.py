import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jchemo_b200 as jc
from jchemo_b200 import device as dev
n, p, q, nlv = 1_000_000, 500, 10, 25
torch.cuda.set_device(0); dev.init(0)
X = torch.empty((p, n), dtype=torch.float64).pin_memory(); Y = torch.empty((q, n), dtype=torch.float64).pin_memory()
X.uniform_(); Y.uniform_()
Xn = X.numpy().T; Yn = Y.numpy().T
for i in range(3):
    t0 = time.perf_counter(); fm = jc.plskern(Xn, Yn, nlv=nlv); t1 = time.perf_counter()
    print("wall ms", (t1 - t0) * 1e3, file=sys.stderr)
