"""Row-sharded fit: one process per GPU, rows of X/Y/w partitioned in contiguous blocks.

The only exchange of the path is one sum all-reduce of the packed partial-Gram buffer
[Gxx | Gxy | gyy | sx | sy | sw] preceded by a (p+q)-double broadcast of the pivot.
After it every rank holds bit-identical inputs, runs K3/K4 redundantly and computes the scores of
its own rows; predict/transform shard by rows with no communication.

Two carriers for that exchange: `PeerComm` — the library's own kernels over CUDA-IPC peer windows
(NVLink, no collective call, no host synchronisation; the B200 path) — and torch.distributed
(`broadcast_pivot` / `reduce_packed`: NCCL on GPUs, gloo in the CPU tests of the host-side protocol).
"""
import torch
import torch.distributed as dist

from . import device as dev


def bind_host_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (Linux sysfs), so that page-locked
    buffers allocated afterwards are local to the GPU's PCIe root.  Returns the node, or None when the
    platform does not expose one (e.g. a VM: numa_node = -1); never raises."""
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shard_rows(n, rank, world):
    """Contiguous row block [lo, hi) of rank `rank`; boundaries are even so every shard of a
    column-major device buffer stays 16-byte aligned."""
    per = -(-n // world)
    per += per & 1
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def broadcast_pivot(pivot, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(pivot, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return pivot


def reduce_packed(packed, group=None):
    """Sum the packed partial-Gram buffers of all ranks, in place, result on every rank."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


class PeerComm:
    """The exchange of the sharded fit through peer HBM (jcb200_comm_*, csrc/comm.cu) instead of a collective
    library: every rank pushes its packed block into slot [rank] of every rank's CUDA-IPC window over NVLink and
    the next kernel on each rank adds the slots in rank order once the `world` flags are up.  No host
    synchronisation and no NCCL call inside the fit; torch.distributed only carries the 64-byte IPC handles
    once, here.  Every rank must construct it (collective) and issue the same sequence of fits."""

    def __init__(self, max_packed_len, group=None):
        import ctypes as C
        from . import _lib
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        handle = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        _lib.check(_lib.lib().jcb200_comm_create(self.rank, self.world, int(max_packed_len), handle), "comm_create")
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(handle), group=group)
            blob = b"".join(gathered)
            assert len(blob) == self.world * _lib.IPC_HANDLE_BYTES
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            _lib.check(_lib.lib().jcb200_comm_connect(buf), "comm_connect")
            dist.barrier(group=group)        # every window is mapped before the first push

    def pivot(self, X, Y, n_local, pivot):
        dev.comm_pivot_dev(X, Y, max(n_local, 1), pivot)

    def allreduce(self, packed):
        dev.comm_allreduce_dev(packed)

    def gram(self, X, Y, w, n_local, pivot):
        dev.comm_gram_dev(X, Y, w, n_local, pivot)

    def solve(self, pivot, model, scal=False):
        dev.comm_solve_dev(pivot, model, scal)

    def close(self):
        from . import _lib
        _lib.check(_lib.lib().jcb200_comm_destroy(), "comm_destroy")


def fit_sharded(X, Y, w, n_local, model, scal=False, group=None, pivot=None, packed=None, comm=None, fused=True):
    """X [p, ld], Y [q, ld], w [n_local] or None hold this rank's rows on its GPU.  The library's kernels and
    the collectives must share one stream: the fit is issued on torch's CURRENT stream (the library is switched
    to it here).  `comm`: a PeerComm (peer-HBM exchange) — else torch.distributed (NCCL / gloo) carries the
    pivot broadcast and the packed all-reduce.  With a PeerComm the exchange is FUSED by default: K1b stores the
    reduced block into every rank's window and K3 reads the sum of the slots (`fused=False`: separate push and
    sum kernels on `packed`).  A rank may hold no rows (n_local == 0): it contributes zeros and still takes part
    in the exchange."""
    dev.use_current_stream()
    p, q = X.shape[0], Y.shape[0]
    if pivot is None:
        pivot = torch.empty(p + q + 1, dtype=torch.float64, device=X.device)
    if packed is None:
        packed = torch.empty(dev.packed_len(p, q), dtype=torch.float64, device=X.device)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if comm is not None:
        comm.pivot(X, Y, n_local, pivot)            # rank 0 computes and publishes, the others fetch
    else:
        if rank == 0:
            dev.pivot_dev(X, Y, n_local, pivot)
        broadcast_pivot(pivot, group)
    if comm is not None and fused:
        comm.gram(X, Y, w, n_local, pivot)      # K1b writes into every rank's window: the exchange is the store
        comm.solve(pivot, model, scal)          # K3 reads the sum of the slots once the flags are up
    else:
        if n_local > 0:
            dev.gram_dev(X, Y, w, n_local, pivot, packed)
        else:
            packed.zero_()
        if comm is not None:
            comm.allreduce(packed)
        else:
            reduce_packed(packed, group)
        dev.solve_dev(packed, pivot, model, scal)
    if n_local > 0:
        if model.nlv > 0:
            dev.scores_dev(X, n_local, model, pivot=pivot)
        dev.weights_dev(w, n_local, model)
    return model


def chunk_bounds(n):
    """Row chunks of the streamed host fit: n/8 rows each, the last one cut again into n/16, n/32, n/32 (what is
    left to do when the last byte has arrived is K1 on 1/32 of the rows); boundaries are even."""
    if n < 400000:
        return [0, n]
    chunk = dev.even_up(-(-n // 8))
    b = [0]
    while n - b[-1] > chunk:
        b.append(b[-1] + chunk)
    rest = n - b[-1]
    half, quarter = dev.even_up(rest // 2), dev.even_up(rest // 4)
    if quarter >= 4096:
        b += [b[-1] + half, b[-1] + half + quarter]
    b.append(n)
    return b


_side = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=device)
    return _side[key]


def fit_sharded_from_host(hX, hY, hw, X, Y, w, n_local, model, scal=False, group=None, pivot=None, packed=None,
                          hT=None, comm=None):
    """fit_sharded with this rank's rows in page-locked HOST tensors hX [p, n_local], hY [q, n_local]
    (hw [n_local] or None); X, Y, w are the device buffers they are streamed into.  The rows go over in
    chunks on a side stream and K1 runs on chunk i while chunk i+1 is in flight, the partial Grams
    accumulating in `packed`; the pivot comes from rank 0's first chunk.  Issued on torch's CURRENT stream (the
    library is switched to it here).  The host paces the pipeline — the
    copy of chunk i+1 is issued after the kernel on chunk i — because a copy queued ahead of a kernel was
    observed to hold the kernel back.  With hT [nlv, n_local] (page-locked) the scores are
    copied back in four row blocks, each under the next block's K5."""
    dev.use_current_stream()
    p, q = X.shape[0], Y.shape[0]
    if pivot is None:
        pivot = torch.empty(p + q + 1, dtype=torch.float64, device=X.device)
    if packed is None:
        packed = torch.empty(dev.packed_len(p, q), dtype=torch.float64, device=X.device)
    if n_local <= 0:
        raise ValueError("fit_sharded_from_host: this rank holds no rows (use fit_sharded, which accepts empty shards)")
    main, side = torch.cuda.current_stream(X.device), _side_stream(X.device)
    b = chunk_bounds(n_local)
    side.wait_stream(main)

    def issue(ci):
        r0, r1 = b[ci], b[ci + 1]
        dev.copy_rows_async(X[:, r0:r1], hX[:, r0:r1], r1 - r0, side)
        dev.copy_rows_async(Y[:, r0:r1], hY[:, r0:r1], r1 - r0, side)
        with torch.cuda.stream(side):
            if hw is not None and ci == 0:
                w[:n_local].copy_(hw, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        return ev

    ev = issue(0)
    for ci in range(len(b) - 1):
        r0, r1 = b[ci], b[ci + 1]
        ev.synchronize()
        Xc, Yc = X[:, r0:r1], Y[:, r0:r1]
        if ci == 0:
            if comm is not None:
                comm.pivot(Xc, Yc, r1 - r0, pivot)
            else:
                dev.pivot_dev(Xc, Yc, r1 - r0, pivot)
                broadcast_pivot(pivot, group)
        dev.gram_dev(Xc, Yc, None if w is None or hw is None else w[r0:r1], r1 - r0, pivot, packed,
                     accumulate=ci > 0)
        if ci + 2 < len(b):
            ev = issue(ci + 1)
    if comm is not None:
        comm.allreduce(packed)
    else:
        reduce_packed(packed, group)
    dev.solve_dev(packed, pivot, model, scal)
    if model.nlv > 0:
        if hT is None or n_local < 400000:
            dev.scores_dev(X, n_local, model, pivot=pivot)
            if hT is not None:
                hT.copy_(model.T[:model.nlv, :n_local], non_blocking=True)
        else:
            blk = dev.even_up(-(-n_local // 4))
            for r0 in range(0, n_local, blk):
                r1 = min(n_local, r0 + blk)
                dev.scores_dev(X[:, r0:r1], r1 - r0, model, out=model.T[:, r0:r1], pivot=pivot)
                done = torch.cuda.Event()
                done.record(main)
                side.wait_event(done)
                dev.copy_rows_async(hT[:, r0:r1], model.T[:model.nlv, r0:r1], r1 - r0, side)
            main.wait_stream(side)
    dev.weights_dev(w if hw is not None else None, n_local, model)
    return model
