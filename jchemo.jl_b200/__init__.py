"""jchemo.jl_b200 — B200-native kernel-PLS path of Jchemo.jl (plskern / transform / coef / predict).

Importable as `jchemo_b200` (see /jchemo_b200.py: the directory name carries a dot).
The compute lives in libjchemo_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/jchemo_b200.h); this package is the host-side mirror of the reference's function API.
"""
from ._lib import JchemoB200Error, NonFiniteError, lib, last_timings, LIB_PATH, SIGNATURES  # noqa: F401
from .plskern import (Plsr, plskern, plskern_bang, transform, coef, predict, summary,  # noqa: F401
                      xfit, xfit_bang, xresid, xresid_bang, ensure_mat, CoefResult, PredResult,
                      resident, resident_add, resident_drop, last_fit_info)

from .gridscore import gridscorelv, gridcvlv, locwlv, residual_sums  # noqa: F401

__all__ = ["gridscorelv", "gridcvlv", "locwlv", "Plsr", "plskern", "plskern_bang", "transform", "coef", "predict", "summary", "xfit", "xfit_bang",
           "xresid", "xresid_bang", "ensure_mat", "resident", "resident_add", "resident_drop", "last_fit_info",
           "JchemoB200Error", "NonFiniteError", "lib", "last_timings"]


def init_multi(device_ids):
    """Bind the library to several GPUs of one box (single process): the host-pointer fit then shards
    the rows over them (jcb200_init_multi).  Call before any other entry point."""
    import ctypes as C
    ids = (C.c_int * len(device_ids))(*device_ids)
    from ._lib import check
    check(lib().jcb200_init_multi(len(device_ids), ids), "init_multi")
