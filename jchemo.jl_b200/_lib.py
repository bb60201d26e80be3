"""ctypes binding of libjchemo_b200.so — the same C ABI a Julia `ccall` binds (include/jchemo_b200.h).

The library is the product; this module fails loudly when it is missing or cannot be loaded.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# JCB_LIB selects an alternative build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("JCB_LIB") or os.path.join(HERE, "libjchemo_b200.so")

c_dp = C.POINTER(C.c_double)
i64, i32 = C.c_int64, C.c_int32

# name -> (restype, argtypes); every symbol include/jchemo_b200.h declares
SIGNATURES = {
    "jcb200_version": (C.c_int, []),
    "jcb200_last_error": (C.c_char_p, []),
    "jcb200_init": (C.c_int, [C.c_int]),
    "jcb200_init_multi": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "jcb200_device_count": (C.c_int, []),
    "jcb200_shutdown": (None, []),
    "jcb200_set_stream": (C.c_int, [C.c_void_p, i32]),
    "jcb200_last_timings": (C.c_int, [c_dp, C.c_int]),
    "jcb200_sync_timings": (C.c_int, []),
    "jcb200_set_phase_timing": (C.c_int, [C.c_int]),
    "jcb200_gram_timings": (C.c_int, [c_dp, C.c_int]),
    "jcb200_launch_count": (i64, []),
    "jcb200_host_register": (C.c_int, [C.c_void_p, i64]),
    "jcb200_host_unregister": (C.c_int, [C.c_void_p]),
    "jcb200_host_alloc": (C.c_void_p, [i64]),
    "jcb200_host_free": (C.c_int, [C.c_void_p]),
    "jcb200_resident_add": (C.c_int, [C.c_void_p, i64, i64, i64]),
    "jcb200_resident_drop": (C.c_int, [C.c_void_p]),
    "jcb200_resident_count": (C.c_int, []),
    "jcb200_last_fit_info": (C.c_int, [C.POINTER(i32)]),
    "jcb200_comm_create": (C.c_int, [i32, i32, i64, C.c_void_p]),
    "jcb200_comm_connect": (C.c_int, [C.c_void_p]),
    "jcb200_comm_destroy": (C.c_int, []),
    "jcb200_comm_timeouts": (C.c_int, []),
    "jcb200_comm_pivot_dev": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, i64, i64, i64, C.c_void_p]),
    "jcb200_comm_allreduce_dev": (C.c_int, [C.c_void_p, i64]),
    "jcb200_comm_gram_dev": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64, i64, i64, C.c_void_p]),
    "jcb200_comm_solve_dev": (C.c_int, [C.c_void_p, i64, i64, i32, i32] + [C.c_void_p] * 10),
    "jcb200_plskern_fit": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64, i64, i64, i32,
                                     i32, i32, C.c_void_p, i64] + [C.c_void_p] * 10 +
                           [C.POINTER(i32)]),
    "jcb200_transform": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, i32,
                                   C.c_void_p, i64]),
    "jcb200_coef": (C.c_int, [C.c_void_p] * 6 + [i64, i64, i32, C.c_void_p, C.c_void_p]),
    "jcb200_predict_sweep": (C.c_int, [C.c_void_p, i64, i64, i64, i64, C.c_void_p, C.c_void_p, i32] +
                             [C.c_void_p] * 4 + [i32, i32, C.POINTER(C.c_void_p)]),
    "jcb200_gridscore": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, i64, i64, i64, C.c_void_p, C.c_void_p,
                                   i32] + [C.c_void_p] * 4 + [i32, i32] + [C.c_void_p] * 4),
    "jcb200_gridcv": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, i64, i64, i64, C.c_void_p, C.c_void_p, i32,
                                i32, i32, i32, i32] + [C.c_void_p] * 4),
    "jcb200_xfit": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32, i32,
                              C.c_void_p, i64]),
    "jcb200_locw_plskern": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, i64, i64, i64, C.c_void_p, i64, i64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, i32, C.c_void_p]),
    "jcb200_summary": (C.c_int, [C.c_void_p, i64, i64, i64] + [C.c_void_p] * 5 + [i32] + [C.c_void_p] * 3),
    "jcb200_packed_len": (i64, [i64, i64]),
    "jcb200_pivot_dev": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, i64, i64, i64, C.c_void_p]),
    "jcb200_gram_dev": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64, i64, i64,
                                  C.c_void_p, C.c_void_p, i32]),
    "jcb200_solve_dev": (C.c_int, [C.c_void_p, C.c_void_p, i64, i64, i32, i32] + [C.c_void_p] * 10),
    "jcb200_xmul_dev": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, i64,
                                  i32, C.c_void_p, C.c_void_p, i64]),
    "jcb200_copy_rows_async": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, i64, i64, i32, C.c_void_p]),
    "jcb200_scores_dev": (C.c_int, [C.c_void_p, i64, i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, i32,
                                    C.c_void_p, C.c_void_p, i64]),
    "jcb200_predict_sweep_dev": (C.c_int, [C.c_void_p, i64, i64, i64, i64, C.c_void_p, C.c_void_p, i32] +
                                 [C.c_void_p] * 4 + [i32, i32, C.c_void_p]),
    "jcb200_center_scale_dev": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p]),
    "jcb200_weights_dev": (C.c_int, [C.c_void_p, i64, C.c_void_p, C.c_void_p]),
    "jcb200_fill_uniform_dev": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_uint64, i64, i64]),
    "jcb200_plskern_fit_dev": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64, i64, i64,
                                         i32, i32, i32, C.c_void_p, i64] + [C.c_void_p] * 10),
}

PHASES = ["h2d", "pivot", "gram", "reduce", "finalize", "lvloop", "scores", "writeback", "d2h", "total"]

_lib = None


ENONFINITE = -5      # JCB200_ENONFINITE
IPC_HANDLE_BYTES = 64


class JchemoB200Error(RuntimeError):
    pass


class NonFiniteError(JchemoB200Error, ValueError):
    """X, Y or the weights contain NaN / Inf (the reference: ArgumentError from LAPACK's svd, plskern.jl:154)."""


def lib():
    """The loaded library; raises if libjchemo_b200.so is absent (build it: python __graft_entry__.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JchemoB200Error(
                f"{LIB_PATH} not found: build the CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().jcb200_last_error().decode("utf-8", "replace")
        cls = NonFiniteError if rc == ENONFINITE else JchemoB200Error
        raise cls(f"{what} failed (status {rc}): {msg}")


def gram_timings(k):
    """Durations (ms) of the last k K1 launches, most recent first."""
    buf = (C.c_double * k)()
    n = lib().jcb200_gram_timings(buf, k)
    return [buf[i] for i in range(max(n, 0))]


def last_timings():
    buf = (C.c_double * len(PHASES))()
    n = lib().jcb200_last_timings(buf, len(PHASES))
    return {PHASES[i]: buf[i] for i in range(n)}
