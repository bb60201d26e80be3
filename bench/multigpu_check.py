"""torchrun script: row-sharded fit over WORLD_SIZE GPUs vs a single-GPU fit of the same global matrix
(rank 0), both through libjchemo_b200.  Prints the relative errors; exits non-zero above 1e-10."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import jchemo_b200 as jc
from jchemo_b200 import device as dev, sharded

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev.init(lr); dev.use_current_stream()
n, p, q, nlv = 400_000, 500, 10, 25
lo, hi = sharded.shard_rows(n, rank, world)
nl = hi - lo
X = dev.colmajor_empty(nl, p); Y = dev.colmajor_empty(nl, q)
dev.fill_uniform(X, nl, 1, lo, n); dev.fill_uniform(Y, nl, 2, lo, n)
w = torch.empty((1, dev.even_up(nl)), dtype=torch.float64, device="cuda")
dev.fill_uniform(w, nl, 3, lo, n); w = (w + 0.5).reshape(-1)
m = dev.DeviceModel(nl, p, q, nlv)
sharded.fit_sharded(X, Y, w, nl, m, scal=True)
torch.cuda.synchronize()
# gather scores for the check
ok = True
if rank == 0:
    Xg = dev.colmajor_empty(n, p); Yg = dev.colmajor_empty(n, q)
    dev.fill_uniform(Xg, n, 1); dev.fill_uniform(Yg, n, 2)
    wg = torch.empty((1, dev.even_up(n)), dtype=torch.float64, device="cuda")
    dev.fill_uniform(wg, n, 3); wg = (wg + 0.5).reshape(-1)
    m1 = dev.DeviceModel(n, p, q, nlv)
    dev.fit_dev(Xg, Yg, wg, n, m1, scal=True)
    torch.cuda.synchronize()
    def rel(a, b): return float((a - b).norm() / b.norm())
    s = torch.sign((m.W * m1.W).sum(1))
    errs = {"xmeans": rel(m.xmeans, m1.xmeans), "xscales": rel(m.xscales, m1.xscales),
            "R": rel(m.R * s[:, None], m1.R), "TT": rel(m.TT, m1.TT),
            "B": rel((m.R.T / m.xscales[:, None]) @ m.C, (m1.R.T / m1.xscales[:, None]) @ m1.C),
            "T_shard": rel(m.T[:nlv, :nl] * s[:, None], m1.T[:nlv, :nl]),
            "weights_shard": rel(m.weights[:nl], m1.weights[:nl])}
    print("sharded (%d GPUs) vs single GPU:" % world, errs)
    ok = all(v < 1e-10 for v in errs.values())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
