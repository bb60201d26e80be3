"""torchrun script: row-sharded fit over WORLD_SIZE GPUs vs a single-GPU fit of the same global matrix
(rank 0), both through libjchemo_b200; with --comm peer the exchange runs through the library's CUDA-IPC peer
windows (csrc/comm.cu) and is also compared, bit for bit across ranks, with the NCCL carrier's result.
Also covers a rank that holds no rows.  Prints the relative errors and PARITY_OK; exits non-zero above 1e-10."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import jchemo_b200 as jc
from jchemo_b200 import device as dev, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--comm", default="nccl", choices=["nccl", "peer"])
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev.init(lr); dev.use_current_stream()
comm = sharded.PeerComm(dev.packed_len(500, 10)) if a.comm == "peer" else None


def rel(x, y):
    return float((x - y).norm() / y.norm())


def case(n, p, q, nlv, weighted, scal, reps=1, fused=True):
    lo, hi = sharded.shard_rows(n, rank, world)
    nl = hi - lo
    X = dev.colmajor_empty(max(nl, 2), p); Y = dev.colmajor_empty(max(nl, 2), q)
    w = None
    if nl > 0:
        dev.fill_uniform(X, nl, 1, lo, n); dev.fill_uniform(Y, nl, 2, lo, n)
    if weighted:
        w = torch.empty((1, dev.even_up(max(nl, 2))), dtype=torch.float64, device="cuda")
        if nl > 0:
            dev.fill_uniform(w, nl, 3, lo, n)
        w = (w + 0.5).reshape(-1)
    m = dev.DeviceModel(max(nl, 2), p, q, nlv)
    for _ in range(reps):                      # several exchanges back to back: the double-buffered windows
        sharded.fit_sharded(X, Y, w, nl, m, scal=scal, comm=comm, fused=fused)
    torch.cuda.synchronize()
    errs = {}
    if comm is not None:
        m2 = dev.DeviceModel(max(nl, 2), p, q, nlv)
        sharded.fit_sharded(X, Y, w, nl, m2, scal=scal, comm=None)      # NCCL carrier
        torch.cuda.synchronize()
        s2 = torch.sign((m.W * m2.W).sum(1))
        errs["peer_vs_nccl_R"] = rel(m.R * s2[:, None], m2.R)
        errs["peer_vs_nccl_TT"] = rel(m.TT, m2.TT)
    # every rank holds the same model bits
    chk = torch.stack([m.R.sum(), m.C.sum(), m.TT.sum(), m.xmeans.sum()])
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    errs["ranks_differ"] = 0.0 if torch.equal(lo_, hi_) else 1.0
    if rank == 0:
        Xg = dev.colmajor_empty(n, p); Yg = dev.colmajor_empty(n, q)
        dev.fill_uniform(Xg, n, 1); dev.fill_uniform(Yg, n, 2)
        wg = None
        if weighted:
            wg = torch.empty((1, dev.even_up(n)), dtype=torch.float64, device="cuda")
            dev.fill_uniform(wg, n, 3); wg = (wg + 0.5).reshape(-1)
        m1 = dev.DeviceModel(n, p, q, nlv)
        dev.fit_dev(Xg, Yg, wg, n, m1, scal=scal)
        torch.cuda.synchronize()
        s = torch.sign((m.W * m1.W).sum(1))
        errs.update({"xmeans": rel(m.xmeans, m1.xmeans), "xscales": rel(m.xscales, m1.xscales),
                     "R": rel(m.R * s[:, None], m1.R), "TT": rel(m.TT, m1.TT),
                     "B": rel((m.R.T / m.xscales[:, None]) @ m.C, (m1.R.T / m1.xscales[:, None]) @ m1.C),
                     "T_shard": rel(m.T[:nlv, :nl] * s[:, None], m1.T[:nlv, :nl]),
                     "weights_shard": rel(m.weights[:nl], m1.weights[:nl])})
        print(f"sharded ({world} GPUs, {a.comm}, fused={fused}) n={n} p={p} q={q} vs single GPU:", errs, flush=True)
    ok = all(v < 1e-10 for v in errs.values())
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


ok = case(400_000, 500, 10, 25, True, True, reps=3)
ok &= case(30_001, 37, 3, 6, False, False, reps=4)
ok &= case(2, 3, 1, 1, False, False)            # rank 1 (and up) hold no rows
if comm is not None:                            # the unfused form (push + sum kernels) and a mix of both
    ok &= case(30_001, 37, 3, 6, True, True, reps=3, fused=False)
    ok &= case(100_000, 130, 2, 9, False, False, reps=2, fused=True)
    ok &= case(2, 3, 1, 1, False, False, fused=False)
if comm is not None:
    ok &= jc.lib().jcb200_comm_timeouts() == 0
    comm.close()
if rank == 0 and ok:
    print("PARITY_OK", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
