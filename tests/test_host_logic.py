"""CPU checks of host-side logic and of the arithmetic the LV-loop kernel is built on (no GPU):
 * row-chunk boundaries of the streamed fits,
 * the dominant-eigenvector routine of csrc/eig.cuh (repeated squaring, power-of-two scaling, the
   tr(A^2)/tr(A)^2 stopping rule) restated in NumPy against LAPACK,
 * the distributed form of the LV loop (csrc/k4_solve.cu: lvdist_kernel) restated in NumPy — every
   p-vector sliced over 16 "CTAs", four all-to-all exchanges per LV with fixed-order sums, w normalised
   together with r — against the oracle's plskern on the same Gram inputs."""
import numpy as np
import pytest

import oracle
from oracle import synth


def test_chunk_bounds_cover_rows_with_even_boundaries():
    from jchemo_b200 import sharded
    for n in (1, 2, 1001, 399_999, 400_000, 400_001, 1_000_000, 999_999, 12_345_679):
        b = sharded.chunk_bounds(n)
        assert b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:]))
        assert all(x % 2 == 0 for x in b[:-1])            # every chunk starts 16-byte aligned
        if n >= 400_000:
            sizes = np.diff(b)
            assert len(sizes) == 10 and sizes[-1] <= n // 32 + 2 and sizes[:7].min() >= n // 8


def eig_dominant(M, maxit=64):
    """NumPy restatement of eig_dominant_warp (csrc/eig.cuh)."""
    tr = np.trace(M)
    if not tr > 0.0:
        v = np.zeros(M.shape[0])
        v[0] = 1.0
        return v, 0
    A = M / tr
    tau_prev = scl_prev = 1.0
    it = 0
    for it in range(maxit):
        tau = np.trace(A)
        scl = 2.0 ** (-2 * (np.frexp(tau)[1] - 1))      # exact power of two, tau = f * 2^e with f in [1, 2)
        A = scl * (A @ A)
        if it > 0 and tau >= (1.0 - 1e-4) * scl_prev * tau_prev * tau_prev:
            break
        tau_prev, scl_prev = tau, scl
    best = int(np.argmax(np.diag(A)))
    return A[:, best].copy(), it + 1


def test_eig_rule_matches_lapack():
    rng = np.random.default_rng(0)
    worst, most = 0.0, 0
    for _ in range(500):
        q = int(rng.integers(2, 17))
        B = rng.standard_normal((int(rng.integers(q, 60)), q)) * np.exp(rng.standard_normal(q) * rng.uniform(0, 3))
        M = B.T @ B
        v, its = eig_dominant(M)
        v /= np.linalg.norm(v)
        lam, V = np.linalg.eigh(M)
        err = min(np.linalg.norm(v - V[:, -1]), np.linalg.norm(v + V[:, -1]))
        worst = max(worst, err * (lam[-1] - lam[-2]) / lam[-1])     # error scaled by the relative gap
        most = max(most, its)
    assert worst < 1e-14 and most <= 20
    v, its = eig_dominant(np.zeros((3, 3)))
    assert its == 0 and v.tolist() == [1.0, 0.0, 0.0]


def lvloop_distributed(XtX, XtY, nlv, ncta=16):
    """NumPy restatement of lvdist_kernel: slices, exchanges A-D, fixed-order sums of the 16 partials."""
    p, q = XtY.shape
    per = -(-p // ncta)
    sl = [slice(min(p, c * per), min(p, (c + 1) * per)) for c in range(ncta)]
    xs = [XtY[s].copy() for s in sl]
    P, R, W = np.zeros((p, nlv)), np.zeros((p, nlv)), np.zeros((p, nlv))
    C, TT = np.zeros((q, nlv)), np.zeros(nlv)

    def allsum(parts):                       # slot c of every CTA's buffer, summed in the same fixed order
        tot = np.zeros_like(parts[0])
        for x in parts:
            tot = tot + x
        return tot
    for a in range(nlv):
        if q > 1:
            M = allsum([x.T @ x for x in xs])                               # exchange A
            v, _ = eig_dominant(M)
        else:
            v = np.ones(1)
        wt = [x @ v for x in xs]                                            # w~ slices (unnormalised)
        dn = allsum([np.concatenate([P[s, :a].T @ w_, [w_ @ w_]]) for s, w_ in zip(sl, wt)])   # exchange B
        d, nrm = dn[:a], np.sqrt(dn[a])
        r = np.concatenate([(w_ - R[s, :a] @ d) / nrm for s, w_ in zip(sl, wt)])              # exchange C (gather)
        u = allsum([x.T @ r[s] for s, x in zip(sl, xs)])
        zp = np.concatenate([XtX[s] @ r for s in sl])
        tt = allsum([np.array([r[s] @ zp[s]]) for s in sl])[0]                                 # exchange D
        c = u / tt
        for s, x in zip(sl, xs):
            x -= np.outer(zp[s], c)
        P[:, a], R[:, a], W[:, a] = zp / tt, r, np.concatenate(wt) / nrm
        C[:, a], TT[a] = c, tt
    return W, R, P, C, TT


@pytest.mark.parametrize("n,p,q,nlv,scal", [(400, 37, 3, 8, False), (300, 20, 1, 6, True), (500, 70, 10, 12, True),
                                            (200, 9, 2, 5, False)])
def test_distributed_lv_loop_matches_oracle(n, p, q, nlv, scal):
    X = synth.synth_matrix(1, n, p)
    Y = synth.synth_matrix(2, n, q) + X[:, :q]
    w = synth.synth_weights(n, uniform=False)
    ref = oracle.plskern(X, Y, w, nlv=nlv, scal=scal)
    wn = w / w.sum()
    Xc, Yc = (X - ref.xmeans) / ref.xscales, (Y - ref.ymeans) / ref.yscales
    XtX = Xc.T @ (wn[:, None] * Xc)
    XtY = Xc.T @ (wn[:, None] * Yc)
    W, R, P, C, TT = lvloop_distributed(XtX, XtY, nlv)
    s = np.sign(np.sum(W * ref.W, axis=0))
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)    # noqa: E731
    assert rel(W * s, ref.W) < 1e-10 and rel(R * s, ref.R) < 1e-10 and rel(P * s, ref.P) < 1e-10
    assert rel(C * s, ref.C) < 1e-10 and rel(TT, ref.TT) < 1e-10


def test_numa_binding_is_a_noop_without_a_gpu_or_numa_info():
    import os
    from jchemo_b200 import sharded
    before = os.sched_getaffinity(0)
    node = sharded.bind_host_to_gpu_numa(0)
    assert node is None or isinstance(node, int)
    if node is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): exactly one JSON line on stdout
    with the contract's keys; ranks other than 0 print nothing and exit 0."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JCB_BENCH_CPU_ROWS="4000")
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "plskern_fit_fp64_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    res1 = subprocess.run(cmd, env=dict(env, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=300)
    assert res1.returncode == 0 and res1.stdout.strip() == ""


def test_ensure_mat_coercions_mirror_the_reference():
    """`ensure_mat` (/root/reference/src/utility.jl:544-548): Matrix stays, Vector -> n x 1, DataFrame -> Matrix;
    the host mirror and the oracle agree, and a transposed view (Julia's Adjoint) is accepted."""
    import pandas as pd
    import jchemo_b200 as jc
    from oracle import plskern_ref
    v = np.arange(3.0)
    M = np.arange(6.0).reshape(3, 2)
    df = pd.DataFrame(M, columns=["a", "b"])
    for f in (jc.ensure_mat, plskern_ref.ensure_mat):
        assert f(v).shape == (3, 1) and f(M).shape == (3, 2) and f(M.T).shape == (2, 3)
        out = f(df)
        assert isinstance(out, np.ndarray) and out.shape == (3, 2) and np.array_equal(out, M)
