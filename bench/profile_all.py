"""Driver for the final ncu captures: one fit (C2) + one predict sweep (C5 shape at m = 2e5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jchemo_b200 import device as dev, sharded
n, p, q, nlv = 1_000_000, 500, 10, 25
torch.cuda.set_device(0); dev.init(0); dev.use_current_stream()
X = dev.colmajor_empty(n, p); Y = dev.colmajor_empty(n, q)
dev.fill_uniform(X, n, 1); dev.fill_uniform(Y, n, 2)
model = dev.DeviceModel(n, p, q, nlv)
sharded.fit_sharded(X, Y, None, n, model)
m5 = dev.DeviceModel(200_000, p, q, 50)
sharded.fit_sharded(X, Y, None, 200_000, m5)
pred = torch.empty((51, q, n), dtype=torch.float64, device="cuda")
dev.predict_sweep_dev(X, n, m5, 0, 50, pred)
torch.cuda.synchronize()
print("profile_all done")
