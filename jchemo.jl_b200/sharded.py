"""Row-sharded fit: one process per GPU, rows of X/Y/w partitioned in contiguous blocks.

The only exchange of the path is one sum all-reduce of the packed partial-Gram buffer
[Gxx | Gxy | gyy | sx | sy | sw] preceded by a (p+q)-double broadcast of the pivot.
After it every rank holds bit-identical inputs, runs K3/K4 redundantly and computes the scores of
its own rows; predict/transform shard by rows with no communication.

`reduce_packed` is the whole host-side protocol and is backend-agnostic (NCCL on GPUs, gloo in the
CPU tests).
"""
import torch
import torch.distributed as dist

from . import device as dev


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
        dist.broadcast(pivot, src=0, group=group)
    return pivot


def reduce_packed(packed, group=None):
    """Sum the packed partial-Gram buffers of all ranks, in place, result on every rank."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def fit_sharded(X, Y, w, n_local, model, scal=False, group=None, pivot=None, packed=None):
    """X [p, ld], Y [q, ld], w [n_local] or None hold this rank's rows on its GPU."""
    p, q = X.shape[0], Y.shape[0]
    if pivot is None:
        pivot = torch.empty(p + q + 1, dtype=torch.float64, device=X.device)
    if packed is None:
        packed = torch.empty(dev.packed_len(p, q), dtype=torch.float64, device=X.device)
    dev.pivot_dev(X, Y, n_local, pivot)
    broadcast_pivot(pivot, group)
    dev.gram_dev(X, Y, w, n_local, pivot, packed)
    reduce_packed(packed, group)
    dev.solve_dev(packed, pivot, model, scal)
    if model.nlv > 0:
        dev.scores_dev(X, n_local, model, pivot=pivot)
    dev.weights_dev(w, n_local, model)
    return model
