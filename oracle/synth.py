"""Counter-based synthetic inputs (SURVEY.md §8 d) — test infrastructure.

`u(s, k) = (mix64(s*0xD1342543DE82EF95 + (k+1)*0x9E3779B97F4A7C15) >> 11) * 2**-53`
with `mix64` the splitmix64 finaliser, all arithmetic mod 2**64.  Element
(i, j) of an n-row column-major matrix uses counter `k = i + j*n` (zero based),
so the host oracle and the device fill kernel (`jcb200_fill_uniform`) produce
identical bits, and a row shard [r0, r0+m) of the global matrix is generated
with `row0=r0, n_global=n`.

Mirrors the README's `rand(n, p)` U[0,1) data (`/root/reference/README.md:86-87`);
Julia's own RNG stream is not reproduced (the README sets no seed).
"""
import numpy as np

SEED_X, SEED_Y, SEED_W, SEED_N = 1, 2, 3, 4

_C1 = np.uint64(0xD1342543DE82EF95)
_C2 = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix64(z):
    z = z ^ (z >> np.uint64(30))
    z = z * _M1
    z = z ^ (z >> np.uint64(27))
    z = z * _M2
    z = z ^ (z >> np.uint64(31))
    return z


def u01(seed, k):
    """u(seed, k) for an array of uint64 counters k."""
    k = np.asarray(k, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * _C1 + (k + np.uint64(1)) * _C2
        z = _mix64(z)
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def synth_matrix(seed, n_rows, n_cols, row0=0, n_global=None, chunk_cols=64):
    """Column-major (Fortran) float64 [n_rows, n_cols] slab of the global matrix."""
    n_global = n_rows if n_global is None else n_global
    out = np.empty((n_rows, n_cols), dtype=np.float64, order="F")
    i = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    for j0 in range(0, n_cols, chunk_cols):
        j1 = min(n_cols, j0 + chunk_cols)
        j = np.arange(j0, j1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            k = i[:, None] + j[None, :] * np.uint64(n_global)
        out[:, j0:j1] = u01(seed, k)
    return out


def synth_weights(n_rows, row0=0, uniform=True):
    """ones(n) (C1, C2, C4, C5) or 0.5 + u(seed_W, i) (C3); unnormalised."""
    if uniform:
        return np.ones(n_rows, dtype=np.float64)
    i = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    return 0.5 + u01(SEED_W, i)
