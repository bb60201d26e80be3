"""Device-resident entry points over torch CUDA tensors (torch is plumbing: memory, streams,
torch.distributed).  Used by bench.py and by the row-sharded multi-GPU fit; all compute is in
libjchemo_b200.so through the "_dev" C ABI.

Matrices are column-major: a logical [n, p] matrix is held as a torch tensor of shape [p, ld]
(row-major) with ld >= n even, i.e. element (i, j) at storage offset i + j*ld.
"""
import ctypes as C

import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def even_up(n):
    return (n + 1) & ~1


def colmajor_empty(n, p, device="cuda"):
    """[p, ld] float64 tensor holding a column-major n x p matrix (ld = n rounded up to even)."""
    return torch.empty((p, even_up(n)), dtype=torch.float64, device=device)


def colmajor_empty_xy(n, p, q, device="cuda"):
    """X [p, ld] and Y [q, ld] as views of ONE [(p + q), ld] tensor: Y's columns lie right behind X's, so the Gram
    kernel sees a single n x (p + q) matrix (joint column space: no separate X'Y blocks, csrc/k1_gram.cu)."""
    t = torch.empty((p + q, even_up(n)), dtype=torch.float64, device=device)
    return t[:p], t[p:]


def use_current_stream():
    """Launch the library's kernels on torch's current stream (so torch.cuda.Event sees them)."""
    _lib.check(_lib.lib().jcb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream), 1),
               "set_stream")


def use_own_stream():
    _lib.check(_lib.lib().jcb200_set_stream(None, 0), "set_stream")


def init(device_index):
    _lib.check(_lib.lib().jcb200_init(int(device_index)), "init")


def fill_uniform(t, n_rows, seed, row0=0, n_global=None):
    """t: [cols, ld] tensor; fills the first n_rows of every column (SURVEY 8d generator)."""
    n_global = n_rows if n_global is None else n_global
    _lib.check(_lib.lib().jcb200_fill_uniform_dev(_p(t), t.shape[1], n_rows, t.shape[0], seed, row0,
                                                  n_global), "fill_uniform")


class DeviceModel:
    """Device-resident Plsr fields (column-major)."""

    def __init__(self, n, p, q, nlv, device="cuda"):
        f = dict(dtype=torch.float64, device=device)
        self.n, self.p, self.q, self.nlv = n, p, q, nlv
        self.T = colmajor_empty(n, max(nlv, 1), device)
        self.P = torch.empty((max(nlv, 1), p), **f)
        self.R = torch.empty((max(nlv, 1), p), **f)
        self.W = torch.empty((max(nlv, 1), p), **f)
        self.C = torch.empty((max(nlv, 1), q), **f)
        self.TT = torch.empty(max(nlv, 1), **f)
        self.xmeans = torch.empty(p, **f)
        self.xscales = torch.empty(p, **f)
        self.ymeans = torch.empty(q, **f)
        self.yscales = torch.empty(q, **f)
        self.weights = torch.empty(even_up(n), **f)
        self.sumw = torch.empty(2, **f)


def fit_dev(X, Y, w, n, model, scal=False, writeback=False):
    """Whole single-GPU fit on device-resident inputs: X [p, ld], Y [q, ld], w [n] or None."""
    p, q = X.shape[0], Y.shape[0]
    rc = _lib.lib().jcb200_plskern_fit_dev(
        _p(X), X.shape[1], _p(Y), Y.shape[1], _p(w), n, p, q, model.nlv, int(scal), int(writeback),
        _p(model.T), model.T.shape[1], _p(model.P), _p(model.R), _p(model.W), _p(model.C),
        _p(model.TT), _p(model.xmeans), _p(model.xscales), _p(model.ymeans), _p(model.yscales),
        _p(model.weights))
    _lib.check(rc, "plskern_fit_dev")
    return model


def packed_len(p, q):
    return int(_lib.lib().jcb200_packed_len(p, q))


def _ld(t):
    """Leading dimension of a [cols, ld] tensor or of a row-block view t[:, r0:r1] of one."""
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def copy_rows_async(dst, src, rows, stream):
    """dst[:, :rows] <- src[:, :rows] between a page-locked host tensor and a device tensor (either direction;
    both [cols, ld] or row-block views): one strided DMA on `stream` (torch's own copy_ of a strided view
    goes through a contiguous temporary on the host)."""
    to_device = 1 if dst.is_cuda else 0
    _lib.check(_lib.lib().jcb200_copy_rows_async(_p(dst), _ld(dst), _p(src), _ld(src), rows, dst.shape[0],
                                                 to_device, C.c_void_p(stream.cuda_stream)), "copy_rows_async")


def pivot_dev(X, Y, n, pivot):
    _lib.check(_lib.lib().jcb200_pivot_dev(_p(X), _ld(X), _p(Y), _ld(Y), n, X.shape[0],
                                           Y.shape[0], _p(pivot)), "pivot_dev")


def gram_dev(X, Y, w, n, pivot, packed, accumulate=False):
    """Partial Gram of n rows (X, Y may be row-block views); accumulate=True adds to `packed`."""
    _lib.check(_lib.lib().jcb200_gram_dev(_p(X), _ld(X), _p(Y), _ld(Y), _p(w), n, X.shape[0],
                                          Y.shape[0], _p(pivot), _p(packed), int(accumulate)),
               "gram_dev")


def comm_pivot_dev(X, Y, n, pivot):
    """Pivot of rank 0's rows, published to every rank through the peer windows (jcb200_comm_pivot_dev)."""
    _lib.check(_lib.lib().jcb200_comm_pivot_dev(_p(X), _ld(X), _p(Y), _ld(Y), n, X.shape[0], Y.shape[0],
                                                _p(pivot)), "comm_pivot_dev")


def comm_allreduce_dev(packed):
    _lib.check(_lib.lib().jcb200_comm_allreduce_dev(_p(packed), packed.numel()), "comm_allreduce_dev")


def comm_gram_dev(X, Y, w, n, pivot):
    """K1 + K1b with the reduced block written straight into every rank's peer window (fused exchange)."""
    _lib.check(_lib.lib().jcb200_comm_gram_dev(_p(X), _ld(X), _p(Y), _ld(Y), _p(w), n, X.shape[0], Y.shape[0],
                                               _p(pivot)), "comm_gram_dev")


def comm_solve_dev(pivot, model, scal=False):
    """K3 on the sum of the window's slots (after the flags of this exchange) + K4."""
    _lib.check(_lib.lib().jcb200_comm_solve_dev(
        _p(pivot), model.p, model.q, model.nlv, int(scal), _p(model.P), _p(model.R), _p(model.W), _p(model.C),
        _p(model.TT), _p(model.xmeans), _p(model.xscales), _p(model.ymeans), _p(model.yscales), _p(model.sumw)),
        "comm_solve_dev")


def solve_dev(packed, pivot, model, scal=False):
    _lib.check(_lib.lib().jcb200_solve_dev(
        _p(packed), _p(pivot), model.p, model.q, model.nlv, int(scal), _p(model.P), _p(model.R),
        _p(model.W), _p(model.C), _p(model.TT), _p(model.xmeans), _p(model.xscales), _p(model.ymeans),
        _p(model.yscales), _p(model.sumw)), "solve_dev")


def scores_dev(X, n, model, out=None, pivot=None):
    """T = ((X - xmeans) ./ xscales) * R for the rows of X (fit scores or transform).  `pivot`: the pivot
    buffer of the fit these rows belong to (carries K1's centring decision), fit scores only."""
    out = model.T if out is None else out
    if pivot is not None:
        _lib.check(_lib.lib().jcb200_scores_dev(_p(X), _ld(X), n, model.p, model.q, _p(model.xmeans),
                                                _p(model.xscales), _p(model.R), model.nlv, _p(pivot),
                                                _p(out), _ld(out)), "scores_dev")
        return out
    _lib.check(_lib.lib().jcb200_xmul_dev(_p(X), X.shape[1], n, model.p, _p(model.xmeans),
                                          _p(model.xscales), _p(model.R), model.p, model.nlv, None,
                                          _p(out), out.shape[1]), "xmul_dev")
    return out


def weights_dev(w, n, model):
    _lib.check(_lib.lib().jcb200_weights_dev(_p(w), n, _p(model.sumw), _p(model.weights)),
               "weights_dev")


def predict_sweep_dev(X, m, model, k_lo, k_hi, out=None):
    """Predictions for every k in k_lo..k_hi in one pass; out: [(k_hi-k_lo+1), q, m] tensor."""
    nk = k_hi - k_lo + 1
    if out is None:
        out = torch.empty((nk, model.q, m), dtype=torch.float64, device=X.device)
    _lib.check(_lib.lib().jcb200_predict_sweep_dev(
        _p(X), X.shape[1], m, model.p, model.q, _p(model.R), _p(model.C), model.nlv, _p(model.xmeans),
        _p(model.xscales), _p(model.ymeans), _p(model.yscales), k_lo, k_hi, _p(out)),
        "predict_sweep_dev")
    return out


def center_scale_dev(X, n, mu, sigma):
    _lib.check(_lib.lib().jcb200_center_scale_dev(_p(X), X.shape[1], n, X.shape[0], _p(mu), _p(sigma)),
               "center_scale_dev")


def set_phase_timing(on):
    """Phase events on / off (the K1 event ring of gram_timings stays on): a caller that times whole fits itself
    takes the per-phase event records out of its timed region."""
    _lib.check(_lib.lib().jcb200_set_phase_timing(1 if on else 0), "set_phase_timing")


def sync_timings():
    _lib.check(_lib.lib().jcb200_sync_timings(), "sync_timings")
    return _lib.last_timings()
