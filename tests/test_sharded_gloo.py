"""world_size-2 gloo tests (CPU) of the row-sharded fit's host-side protocol: shard boundaries, the
pivot broadcast and the single packed-Gram all-reduce (jchemo_b200.sharded).  The per-shard packed
buffers are built here with NumPy from their definition — what K1 + K1b produce on a GPU —
and the reduced buffer is finalised with a NumPy restatement of K3; the result must equal the oracle's
preamble (means, scales, X'DX, X'DY) on the unsharded data."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth


def packed_partial(X, Y, w, c):
    """[Gxx | Gxy | gyy | sx | sy | sw] of one row shard about the pivot c, A operand raw (K1)."""
    p, q = X.shape[1], Y.shape[1]
    Xb, Yb = (X - c[:p]) * w[:, None], (Y - c[p:p + q]) * w[:, None]
    Gxx = X.T @ Xb                      # acc_ij = sum_k x_ki w_k (x_kj - c_j)
    Gxy = X.T @ Yb
    gyy = np.einsum("ij,ij->j", Y, Yb)
    return np.concatenate([Gxx.ravel(order="F"), Gxy.ravel(order="F"), gyy, Xb.sum(0), Yb.sum(0),
                           [w.sum()]])


def finalize(packed, c, p, q, scal):
    """NumPy restatement of K3 (jchemo.jl_b200/csrc/k4_solve.cu: finalize_*_kernel)."""
    Gxx = packed[:p * p].reshape(p, p, order="F")
    Gxy = packed[p * p:p * p + p * q].reshape(p, q, order="F")
    o = p * p + p * q
    gyy, sx, sy, S = packed[o:o + q], packed[o + q:o + q + p], packed[o + q + p:o + 2 * q + p], packed[-1]
    cx, cy = c[:p], c[p:p + q]
    dx, dy = sx / S, sy / S
    XtX = (Gxx - np.outer(cx, sx)) / S - np.outer(dx, dx)
    XtY = (Gxy - np.outer(cx, sy)) / S - np.outer(dx, dy)
    vy = (gyy - cy * sy) / S - dy * dy
    xs = np.sqrt(np.diag(XtX)) if scal else np.ones(p)
    ys = np.sqrt(vy) if scal else np.ones(q)
    XtX = np.triu(XtX) + np.triu(XtX, 1).T       # K3 mirrors the upper triangle
    return cx + dx, cy + dy, xs, ys, XtX / np.outer(xs, xs), XtY / np.outer(xs, ys)


def _worker(rank, world, port, n, p, q, scal, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jchemo_b200 import sharded
    lo, hi = sharded.shard_rows(n, rank, world)
    X = synth.synth_matrix(1, hi - lo, p, row0=lo, n_global=n) * 3.0 + 10.0
    Y = synth.synth_matrix(2, hi - lo, q, row0=lo, n_global=n)
    w = synth.synth_weights(n, uniform=False)[lo:hi]
    pivot = torch.from_numpy(np.concatenate([X[:64].mean(0), Y[:64].mean(0), [1.0]]) if rank == 0
                             else np.zeros(p + q + 1))
    sharded.broadcast_pivot(pivot)
    c = pivot.numpy()
    packed = torch.from_numpy(packed_partial(X, Y, w, c))
    sharded.reduce_packed(packed)
    if rank == 0:
        np.save(out, np.concatenate([c, packed.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scal", [False, True])
def test_two_rank_reduce_matches_unsharded(tmp_path, scal):
    n, p, q, world = 1001, 17, 3, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "packed.npy")
    mp.spawn(_worker, args=(world, port, n, p, q, scal, out), nprocs=world, join=True)
    blob = np.load(out)
    c, packed = blob[:p + q + 1], blob[p + q + 1:]
    xm, ym, xs, ys, XtX, XtY = finalize(packed, c, p, q, scal)
    # unsharded truth from the oracle's own preamble (plskern.jl:117-132)
    X = synth.synth_matrix(1, n, p) * 3.0 + 10.0
    Y = synth.synth_matrix(2, n, q)
    w = synth.synth_weights(n, uniform=False)
    w = w / w.sum()
    xm0, ym0 = w @ X, w @ Y
    Xc, Yc = X - xm0, Y - ym0
    xs0 = np.sqrt(w @ Xc ** 2) if scal else np.ones(p)
    ys0 = np.sqrt(w @ Yc ** 2) if scal else np.ones(q)
    Xc, Yc = Xc / xs0, Yc / ys0
    np.testing.assert_allclose(xm, xm0, rtol=1e-13)
    np.testing.assert_allclose(ym, ym0, rtol=1e-13)
    np.testing.assert_allclose(xs, xs0, rtol=1e-12)
    np.testing.assert_allclose(ys, ys0, rtol=1e-12)
    np.testing.assert_allclose(XtX, Xc.T @ (w[:, None] * Xc), rtol=0, atol=1e-12)
    np.testing.assert_allclose(XtY, Xc.T @ (w[:, None] * Yc), rtol=0, atol=1e-12)


def test_shard_rows_partition():
    from jchemo_b200 import sharded
    for n, world in [(10, 2), (1001, 2), (1_000_000, 8), (7, 8), (1, 4)]:
        cuts = [sharded.shard_rows(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert b == c and a <= b
        assert all(lo % 2 == 0 for lo, hi in cuts if hi > lo)      # non-empty shards stay 16-byte aligned
