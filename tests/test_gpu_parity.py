"""GPU parity tests: the CUDA path, called through the C ABI (host-pointer entry points, the call a
Julia `ccall` would make), against the NumPy oracle and the committed golden fixtures.

Tolerance (north star): relative Frobenius error <= 1e-10 on B, T and predictions after per-LV sign
alignment.  For the q = 1 `rand` case (C3) the reference cannot reproduce its own late
LVs: columns where ||XtY_a|| / ||XtY_1|| < 1e-6 are graded on B / predictions only.
"""
import numpy as np
import pytest

import oracle
from oracle import make_golden, synth
from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def jc():
    import __graft_entry__ as ge
    ge.build()
    import jchemo_b200
    return jchemo_b200


@pytest.mark.parametrize("name", ["c1", "c1_wscal", "c2_cut", "c3_cut", "c5_cut", "edge_odd"])
def test_fit_matches_golden_and_oracle(jc, name):
    cfg, z = load_golden(name)
    X, Y, w, Xnew = make_golden.inputs(cfg)
    X0, Y0 = X.copy(), Y.copy()
    fm = jc.plskern(X, Y, None if cfg["uniform"] else w, nlv=cfg["nlv"], scal=cfg["scal"])
    assert np.array_equal(X, X0) and np.array_equal(Y, Y0)          # plskern leaves inputs untouched
    a = z["TT"].shape[0]
    assert fm.T.shape == (cfg["n"], a) and fm.P.shape == (cfg["p"], a) and fm.C.shape == (cfg["q"], a)
    assert fm.niter is None and fm.V is fm.P
    for f in ["xmeans", "xscales", "ymeans", "yscales"]:
        assert relerr(getattr(fm, f), z[f]) < 1e-13, f
    assert relerr(fm.weights[z["rows"]], z["weights_rows"]) < 1e-14
    s = np.sign(np.sum(z["W"] * fm.W, axis=0))
    ok = np.ones(a, dtype=bool)
    if cfg["q"] == 1:
        # ||XtY_a|| = |c_a| * tt_a / |r_a' XtY_a / ||XtY_a||| ~ |C[0,a]| TT[a] up to O(1): conditioning proxy
        k = np.abs(z["C"][0]) * z["TT"]
        ok = k / k[0] > 1e-6
    assert ok.sum() >= 3
    assert relerr(fm.TT[ok], z["TT"][ok]) < TOL
    assert relerr((fm.T * s)[z["rows"]][:, ok], z["T_rows"][:, ok]) < TOL
    assert relerr(np.linalg.norm(fm.T, axis=0)[ok], z["T_colnorm"][ok]) < TOL
    for f in ["P", "R", "W"]:
        assert relerr((getattr(fm, f) * s)[:, ok], z[f][:, ok]) < TOL, f
    assert relerr((fm.C * s)[:, ok], z["C"][:, ok]) < TOL
    # coefficients and predictions at every stored k: always graded (B is well-posed)
    for i, k in enumerate(z["ks"]):
        cf = jc.coef(fm, nlv=int(k))
        assert cf.B.shape == (cfg["p"], cfg["q"]) and cf.int.shape == (1, cfg["q"])
        if k > 0:
            assert relerr(cf.B, z["B_ks"][i]) < TOL, k
        else:
            assert np.all(cf.B == 0)
        assert relerr(cf.int, z["int_ks"][i]) < TOL, k
    pr = jc.predict(fm, Xnew, nlv=range(0, a + 1)).pred
    assert len(pr) == a + 1
    for k in range(a + 1):
        assert relerr(pr[k][z["mrows"]], z["pred_all_rows"][k]) < TOL, k
    Tn = jc.transform(fm, Xnew)
    assert relerr((Tn * s)[z["mrows"]][:, ok], z["Tnew_rows"][:, ok]) < TOL
    single = jc.predict(fm, Xnew).pred                              # nlv = nothing -> one matrix
    assert relerr(single[z["mrows"]], z["pred_all_rows"][a]) < TOL


def test_plskern_bang_writes_back(jc):
    cfg = dict(n=1000, p=37, q=3, nlv=4, m=10, uniform=False, scal=True)
    X, Y, w, _ = make_golden.inputs(cfg)
    Xr, Yr = X.copy(), Y.copy()
    ref = oracle.plskern_bang(Xr, Yr, w, nlv=4, scal=True)          # reference leaves Xr, Yr scaled
    fm = jc.plskern_bang(X, Y, w, nlv=4, scal=True)
    assert relerr(X, Xr) < 1e-12 and relerr(Y, Yr) < 1e-12          # plskern.jl:125-126 side effect
    s = oracle.sign_align(ref, fm)
    assert relerr(fm.T * s, ref.T) < TOL


def test_invariants_on_device_fit(jc):
    n, p, q, nlv = 5000, 120, 4, 12
    X = synth.synth_matrix(1, n, p)
    Y = synth.synth_matrix(2, n, q) + X[:, :q]
    w = synth.synth_weights(n, uniform=False)
    fm = jc.plskern(X, Y, w, nlv=nlv, scal=True)
    assert abs(fm.weights.sum() - 1) < 1e-13
    np.testing.assert_allclose(np.linalg.norm(fm.W, axis=0), 1, atol=1e-13)
    Xc = (X - fm.xmeans) / fm.xscales
    assert relerr(fm.T, Xc @ fm.R) < 1e-12
    G = fm.T.T @ (fm.weights[:, None] * fm.T)
    assert np.abs(G - np.diag(fm.TT)).max() / fm.TT.max() < 1e-12
    assert np.abs(fm.P.T @ fm.R - np.eye(nlv)).max() < 1e-11
    Yc = (Y - fm.ymeans) / fm.yscales
    Cexp = (Yc.T @ (fm.weights[:, None] * fm.T)) / fm.TT
    assert relerr(fm.C, Cexp) < 1e-11


def test_edge_cases(jc):
    X = synth.synth_matrix(1, 60, 8)
    Y = synth.synth_matrix(2, 60, 2)
    fm0 = jc.plskern(X, Y, nlv=0)                                   # nlv = 0 is a valid fit
    assert fm0.T.shape == (60, 0) and fm0.R.shape == (8, 0)
    np.testing.assert_allclose(jc.predict(fm0, X[:3]).pred, np.tile(fm0.ymeans, (3, 1)), atol=1e-15)
    assert jc.transform(fm0, X[:3]).shape == (3, 0)
    fm = jc.plskern(X, Y[:, 0], nlv=99)                             # vector Y, nlv clamped to min(n,p)
    assert fm.T.shape == (60, 8) and fm.C.shape == (1, 8)
    ref = oracle.plskern(X, Y[:, 0], nlv=99)
    assert relerr(jc.coef(fm).B, oracle.coef(ref)[0]) < 1e-8       # full rank: conditioning-limited
    pr = jc.predict(fm, X[:4], nlv=[2, 5]).pred                     # widened to 2:5
    assert len(pr) == 4
    pr = jc.predict(fm, X[:4], nlv=range(-3, 50)).pred              # clamped to 0:8
    assert len(pr) == 9
    r = oracle.predict(ref, X[:4], nlv=range(-3, 50))
    for k in range(6):
        assert relerr(pr[k], r[k]) < 1e-9
    # integer, C-ordered inputs and integer weights are coerced like ensure_mat + Float64.()
    Xi = (X * 100).astype(np.int64)
    fmi = jc.plskern(np.ascontiguousarray(Xi), Y, np.arange(1, 61), nlv=1)
    refi = oracle.plskern(Xi.astype(float), Y, np.arange(1, 61), nlv=1)
    assert relerr(jc.predict(fmi, Xi[:1]).pred, oracle.predict(refi, Xi[:1].astype(float))) < TOL


def test_offset_heavy_data(jc):
    """Spectra-like data with large column offsets: the pivot + exact correction must not lose digits."""
    n, p, q, nlv = 4000, 64, 2, 6
    X = synth.synth_matrix(1, n, p) * 1e-2 + 1e3 + np.arange(p)[None, :] * 10.0
    Y = synth.synth_matrix(2, n, q) + 5e2
    fm = jc.plskern(X, Y, nlv=nlv)
    ref = oracle.plskern(X, Y, nlv=nlv)
    assert relerr(jc.predict(fm, X[:50]).pred, oracle.predict(ref, X[:50])) < TOL
    s = oracle.sign_align(ref, fm)
    assert relerr(fm.T * s, ref.T) < 1e-8     # the oracle's own cancellation floor on this data


def test_full_size_properties(jc):
    """BASELINE C2 at full size (n=1e6, p=500, q=10, nlv=25) through size-independent properties:
    scores are D-orthogonal with T'DT = diag(TT), P'R = I, ||w|| = 1, T = Xc R on a row sample,
    and the fit agrees with itself under a row permutation of the inputs."""
    n, p, q, nlv = 1_000_000, 500, 10, 25
    X = synth.synth_matrix(1, n, p)
    Y = synth.synth_matrix(2, n, q)
    fm = jc.plskern(X, Y, nlv=nlv)
    np.testing.assert_allclose(np.linalg.norm(fm.W, axis=0), 1, atol=1e-12)
    G = fm.T.T @ fm.T / n
    assert np.abs(G - np.diag(fm.TT)).max() / fm.TT.max() < 1e-10
    assert np.abs(fm.P.T @ fm.R - np.eye(nlv)).max() < 1e-9
    rows = np.arange(0, n, 997)
    assert relerr(fm.T[rows], (X[rows] - fm.xmeans) @ fm.R) < 1e-11
    assert relerr(fm.xmeans, X.mean(axis=0)) < 1e-13
    perm = np.random.default_rng(0).permutation(n)
    fm2 = jc.plskern(X[perm], Y[perm], nlv=nlv)
    assert relerr(jc.coef(fm2).B, jc.coef(fm).B) < TOL
    s = np.sign(np.sum(fm.W * fm2.W, axis=0))
    assert relerr((fm2.T * s)[np.argsort(perm)][rows], fm.T[rows]) < 1e-9


@pytest.mark.parametrize("score", ["msep", "rmsep", "ssr", "bias", "sep", "r2", "rpd"])
def test_gridscorelv_fused_scores(jc, score):
    """Next row: gridscorelv for fun = plskern, every nlv scored on the device in one pass,
    against the oracle's fit + predict + score loop (gridscore.jl:167-221, scores.jl)."""
    n, m, p, q, nlv = 3000, 1201, 150, 3, 12
    Xtr = synth.synth_matrix(1, n, p)
    Ytr = synth.synth_matrix(2, n, q) + Xtr[:, :q] * 2.0 + Xtr[:, q:2 * q]
    Xte = synth.synth_matrix(4, m, p)
    Yte = synth.synth_matrix(5, m, q) + Xte[:, :q] * 2.0 + Xte[:, q:2 * q]
    got = jc.gridscorelv(Xtr, Ytr, Xte, Yte, score=score, nlv=range(0, nlv + 1))
    ref = oracle.gridscorelv(Xtr, Ytr, Xte, Yte, score=score, nlv=range(0, nlv + 1))
    assert list(got["nlv"]) == list(ref["nlv"]) == list(range(nlv + 1))
    for j in range(q):
        g, r = np.asarray(got[f"y{j + 1}"]), ref[f"y{j + 1}"]
        tol = 1e-9 if score == "bias" else TOL       # bias is a difference of means: absolute scale
        assert np.linalg.norm(g - r) <= tol * max(np.linalg.norm(r), 1e-3), (score, j)


def test_gridscorelv_ranges_and_vector_y(jc):
    X = synth.synth_matrix(1, 400, 20)
    y = synth.synth_matrix(2, 400, 1)[:, 0] + X[:, 0]
    got = jc.gridscorelv(X[:300], y[:300], X[300:], y[300:], score="rmsep", nlv=[2, 5])   # widened to 2:5
    ref = oracle.gridscorelv(X[:300], y[:300], X[300:], y[300:], score="rmsep", nlv=[2, 5])
    assert list(got["nlv"]) == [2, 3, 4, 5]
    assert relerr(np.asarray(got["y1"]), ref["y1"]) < TOL
    got0 = jc.gridscorelv(X[:300], y[:300], X[300:], y[300:], score="msep", nlv=0)         # nlv = 0 only
    ref0 = oracle.gridscorelv(X[:300], y[:300], X[300:], y[300:], score="msep", nlv=0)
    assert relerr(np.asarray(got0["y1"]), ref0["y1"]) < TOL


def test_c4_shape_cut(jc):
    """BASELINE C4's column shape (p=2000, q=10, nlv=50) on a row cut: more Gram groups than SMs, XtY too
    large for shared memory in the LV loop (global path, two cluster barriers per LV), 63 column chunks
    in the score kernel."""
    n, p, q, nlv, m = 6001, 2000, 10, 50, 333
    X = synth.synth_matrix(1, n, p)
    Y = synth.synth_matrix(2, n, q) + X[:, :q] - X[:, q:2 * q]
    Xnew = synth.synth_matrix(4, m, p)
    fm = jc.plskern(X, Y, nlv=nlv)
    ref = oracle.plskern(X, Y, nlv=nlv)
    s = oracle.sign_align(ref, fm)
    assert relerr(fm.T * s, ref.T) < TOL
    assert relerr(fm.R * s, ref.R) < TOL
    assert relerr(jc.coef(fm).B, oracle.coef(ref)[0]) < TOL
    pr = jc.predict(fm, Xnew, nlv=range(0, nlv + 1)).pred
    rr = oracle.predict(ref, Xnew, nlv=range(0, nlv + 1))
    assert max(relerr(a, b) for a, b in zip(pr, rr)) < TOL


def test_wide_y_and_many_lvs(jc):
    """q > 32 (several Y blocks, generic eigen path) and nlv > 64 (two passes of the score kernel)."""
    n, p, q, nlv = 2000, 150, 40, 70
    X = synth.synth_matrix(1, n, p)
    Y = synth.synth_matrix(2, n, q) + X[:, :q]
    fm = jc.plskern(X, Y, nlv=nlv, scal=True)
    ref = oracle.plskern(X, Y, nlv=nlv, scal=True)
    s = oracle.sign_align(ref, fm)
    ok = np.arange(nlv) < 40          # beyond rank(Y'X) = q the kernel is rounding noise in the reference too
    assert relerr((fm.T * s)[:, ok], ref.T[:, ok]) < TOL
    assert relerr(jc.coef(fm, nlv=40).B, oracle.coef(ref, nlv=40)[0]) < TOL
    assert relerr(jc.transform(fm, X[:100], nlv=40) * s[:40], oracle.transform(ref, X[:100], nlv=40)) < TOL


def test_device_api_rejects_misaligned(jc):
    import ctypes as C
    import torch
    from jchemo_b200 import device as dev, _lib
    dev.init(0)
    X = torch.zeros(7 * 101 + 8, dtype=torch.float64, device="cuda")
    Y = torch.zeros(101 + 8, dtype=torch.float64, device="cuda")
    piv = torch.zeros(9, dtype=torch.float64, device="cuda")
    pk = torch.zeros(int(_lib.lib().jcb200_packed_len(7, 1)), dtype=torch.float64, device="cuda")
    rc = _lib.lib().jcb200_gram_dev(C.c_void_p(X.data_ptr()), 101, C.c_void_p(Y.data_ptr()), 101, None,
                                    100, 7, 1, C.c_void_p(piv.data_ptr()), C.c_void_p(pk.data_ptr()), 0)
    assert rc == -4 and b"aligned" in _lib.lib().jcb200_last_error()          # JCB200_EALIGN: odd ld
    rc = _lib.lib().jcb200_gram_dev(None, 100, C.c_void_p(Y.data_ptr()), 100, None, 100, 7, 1,
                                    C.c_void_p(piv.data_ptr()), C.c_void_p(pk.data_ptr()), 0)
    assert rc == -1                                                            # JCB200_EINVAL


def test_concurrent_callers(jc):
    """The reference's callers invoke `fun` from Threads.@threads loops (src/locwlv.jl:18): entry points
    must be re-entrant.  Four host threads fit different problems at once; each result must equal the
    single-threaded one bit for bit (one mutex serialises the device work)."""
    import threading
    probs = []
    for t in range(4):
        n, p, q = 500 + 37 * t, 30 + t, 1 + (t % 3)
        X = synth.synth_matrix(10 + t, n, p)
        Y = synth.synth_matrix(20 + t, n, q) + X[:, :q]
        probs.append((X, Y, 3 + t))
    solo = [jc.plskern(X, Y, nlv=k) for X, Y, k in probs]
    out = [None] * 4

    def work(i):
        for _ in range(5):
            X, Y, k = probs[i]
            fm = jc.plskern(X, Y, nlv=k)
            out[i] = (fm, jc.predict(fm, X[:7]).pred)
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(4):
        assert np.array_equal(out[i][0].T, solo[i].T) and np.array_equal(out[i][0].R, solo[i].R)
        assert np.array_equal(out[i][1], jc.predict(solo[i], probs[i][0][:7]).pred)


def test_summary_explained_variance(jc):
    """Next row: Base.summary(::Plsr, X), plskern.jl:246-260."""
    n, p, q, nlv = 2500, 90, 2, 7
    X = synth.synth_matrix(1, n, p) * (1 + np.arange(p))[None, :]
    Y = synth.synth_matrix(2, n, q) + X[:, :q]
    w = synth.synth_weights(n, uniform=False)
    for scal in (False, True):
        fm = jc.plskern(X, Y, w, nlv=nlv, scal=scal)
        got = jc.summary(fm, X).explvarx
        ref = oracle.summary(oracle.plskern(X, Y, w, nlv=nlv, scal=scal), X)
        for col in ("var", "pvar", "cumpvar"):
            assert relerr(np.asarray(got[col]), ref[col]) < TOL, (scal, col)
        assert list(got["nlv"]) == list(range(1, nlv + 1))


@pytest.mark.parametrize("scal", [False, True])
def test_gridcvlv_gram_downdating(jc, scal):
    """Next row: gridcvlv for fun = plskern by Gram down-dating (one pass over X for all
    folds) against the oracle's K fits on K row-copies (gridcv.jl:187-228).  Two repetitions: a shuffled
    3-fold partition (segmkf-like, odd fold sizes) and a single test set (segmts-like)."""
    n, p, q, nlv = 1501, 60, 2, 8
    X = synth.synth_matrix(1, n, p) * (1.0 + np.arange(p) / p)[None, :]
    Y = synth.synth_matrix(2, n, q) + X[:, :q] * 1.5
    rng = np.random.default_rng(3)
    idx = rng.permutation(n)
    segm = [[np.sort(idx[0:500]), np.sort(idx[500:1001]), np.sort(idx[1001:])],
            [np.sort(rng.permutation(n)[:333])]]
    got = jc.gridcvlv(X, Y, segm=segm, score="rmsep", nlv=range(0, nlv + 1), scal=scal)
    ref_res, ref_rep = oracle.gridcvlv(X, Y, segm=segm, score="rmsep", nlv=range(0, nlv + 1), scal=scal)
    g_rep, g_res = got.res_rep, got.res
    assert list(g_rep["repl"]) == list(ref_rep["repl"]) and list(g_rep["segm"]) == list(ref_rep["segm"])
    assert list(g_rep["nlv"]) == list(ref_rep["nlv"])
    for c in ("y1", "y2"):
        assert relerr(np.asarray(g_rep[c]), ref_rep[c]) < TOL, c
        assert relerr(np.asarray(g_res[c]), ref_res[c]) < TOL, c


def test_empty_new_data(jc):
    """Zero rows of new data: empty results of the right shape, no device call needed."""
    X = synth.synth_matrix(1, 80, 6)
    Y = synth.synth_matrix(2, 80, 2)
    fm = jc.plskern(X, Y, nlv=3)
    E = np.empty((0, 6), order="F")
    assert jc.transform(fm, E).shape == (0, 3)
    assert jc.predict(fm, E).pred.shape == (0, 2)
    pr = jc.predict(fm, E, nlv=range(0, 4)).pred
    assert len(pr) == 4 and all(z.shape == (0, 2) for z in pr)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        jc.predict(fm, np.zeros((3, 5)))
    with pytest.raises(jc.JchemoB200Error):
        jc.plskern(np.empty((0, 6), order="F"), np.empty((0, 2), order="F"), nlv=1)     # n = 0 is rejected


def test_single_process_multi_gpu():
    """jcb200_init_multi: the host-pointer fit sharded over 2 GPUs inside the library (peer-memory Gram
    reduce) against the oracle.  Needs a fresh process (the library binds its devices once) and 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JCB_SKIP_E2E="1")
    res = subprocess.run([sys.executable, os.path.join(root, "bench", "multigpu_inproc.py"), "2"], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert '"parity_ok": true' in res.stdout


def test_constant_y_multiresponse(jc):
    """Constant Y with q > 1.  After centring XtY is rounding noise (or exactly zero): predictions must
    stay at ymeans.  For XtY == 0 exactly the reference's svd(XtY).U[:, 1] is e_1 (LAPACK returns U = I for
    a zero matrix, plskern.jl:154) and c = 0; the device path must do the same instead of 0/0."""
    X = synth.synth_matrix(1, 200, 12)
    Y = np.tile(np.array([[3.0, -1.5]]), (200, 1))
    fm = jc.plskern(X, Y, nlv=2)
    ref = oracle.plskern(X, Y, nlv=2)
    np.testing.assert_allclose(jc.predict(fm, X[:5]).pred, oracle.predict(ref, X[:5]), atol=1e-12)
    Y0 = np.zeros((200, 2))                       # exactly zero after centring: XtY == 0 exactly
    fm0 = jc.plskern(X, Y0, nlv=1)
    ref0 = oracle.plskern(X, Y0, nlv=1)
    assert np.all(np.isfinite(fm0.T)) and np.all(fm0.C == 0) and np.all(ref0.C == 0)
    assert abs(abs(fm0.W[0, 0]) - 1.0) < 1e-15 and np.all(fm0.W[1:, 0] == 0)        # w = e_1
    assert relerr(fm0.T * np.sign(fm0.W[0, 0] * ref0.W[0, 0]), ref0.T) < TOL
    np.testing.assert_allclose(jc.predict(fm0, X[:5]).pred, 0.0, atol=1e-300)


@pytest.mark.parametrize("q,scal,weighted", [(1, False, True), (3, True, True), (2, False, False)])
def test_locwlv_batched_tiny_fits(jc, q, scal, weighted):
    """Next row: locwlv for fun = plskern — one tiny weighted fit per query row, all in one
    kernel launch — against the oracle's loop of fits (locwlv.jl:9-48).  Ragged neighbourhoods, one (q = 1)
    whose neighbours share a single Y value; a neighbourhood smaller than nlv throws as the reference does."""
    rng = np.random.default_rng(7)
    ntr, p, m, nlv = 400, 37, 23, 6
    Xtr = synth.synth_matrix(1, ntr, p)
    Ytr = synth.synth_matrix(2, ntr, q) + Xtr[:, :q] * 2.0
    X = synth.synth_matrix(4, m, p)
    listnn, listw = [], []
    for i in range(m):
        k = [9, 60, 33, 100][i % 4]
        s = np.sort(rng.choice(ntr, size=k, replace=False))
        listnn.append(s)
        listw.append(0.2 + rng.random(k))
    if q == 1:
        Ytr[listnn[5], 0] = 1.25                      # all neighbours of row 5 share one value
    got = jc.locwlv(Xtr, Ytr, X, listnn=listnn, listw=listw if weighted else None, nlv=range(0, nlv + 1),
                    scal=scal).pred
    ref = oracle.locwlv(Xtr, Ytr, X, listnn=listnn, listw=listw if weighted else None, nlv=range(0, nlv + 1),
                        scal=scal)
    assert len(got) == len(ref) == nlv + 1
    for a in range(nlv + 1):
        assert relerr(got[a], ref[a]) < 1e-9, a       # k = 9 neighbourhoods are fitted close to full rank
    listnn[2] = listnn[2][:4]
    listw[2] = listw[2][:4]
    with pytest.raises(ValueError):
        jc.locwlv(Xtr, Ytr, X, listnn=listnn, listw=listw, nlv=nlv)
    with pytest.raises(ValueError):
        oracle.locwlv(Xtr, Ytr, X, listnn=listnn, listw=listw, nlv=nlv)


@pytest.mark.parametrize("scal", [False, True])
def test_xfit_xresid(jc, scal):
    """Next row: xfit / xresid and their bang forms (xfit.jl:33-99) for nlv = nothing, a
    partial nlv and nlv = 0; ragged m and p so that tile edges are exercised."""
    n, p, q, nlv, m = 700, 203, 3, 9, 333
    X, Y = synth.synth_matrix(1, n, p), synth.synth_matrix(2, n, q)
    w = synth.synth_weights(n, uniform=False)
    Xn = synth.synth_matrix(4, m, p) * 3.0 + 1.0
    fm = jc.plskern(X, Y, w, nlv=nlv, scal=scal)
    ref = oracle.plskern(X, Y, w, nlv=nlv, scal=scal)
    for k in (None, 4, 0, 50):
        want = oracle.xfit(ref, Xn, nlv=k)
        got = jc.xfit(fm, Xn, nlv=k)
        assert got.shape == (m, p) and relerr(got, want) < 1e-10, k
        e_want = oracle.xresid(ref, Xn, nlv=k)
        e_got = jc.xresid(fm, Xn, nlv=k)
        assert np.max(np.abs(e_got - e_want)) < 1e-10 * np.max(np.abs(Xn)), k
        Z = np.asfortranarray(Xn.copy())
        assert jc.xfit_bang(fm, Z, nlv=k) is Z and relerr(Z, want) < 1e-10
        Z = np.asfortranarray(Xn.copy())
        jc.xresid_bang(fm, Z, nlv=k)
        assert np.array_equal(Z, e_got)
    with pytest.raises(TypeError):
        jc.xfit_bang(fm, np.ascontiguousarray(Xn))           # xfit! takes X::Matrix only
    # full rank: the model reproduces its own training X
    fm_full = jc.plskern(X[:40, :30], Y[:40], nlv=30)
    assert np.max(np.abs(jc.xresid(fm_full, X[:40, :30]))) < 1e-9


def test_locwlv_neighbourhood_sizes(jc):
    """Neighbourhood sizes around the kernel's 32-row block edges (a missing barrier once showed only for
    33 <= k <= 64)."""
    rng = np.random.default_rng(11)
    ntr, p, m, nlv = 400, 37, 6, 3
    Xtr = synth.synth_matrix(1, ntr, p)
    Ytr = synth.synth_matrix(2, ntr, 1) + Xtr[:, :1] * 2.0
    X = synth.synth_matrix(4, m, p)
    for k in (5, 31, 32, 33, 34, 63, 64, 65, 129, 257):
        listnn = [np.sort(rng.choice(ntr, size=k, replace=False)) for _ in range(m)]
        got = jc.locwlv(Xtr, Ytr, X, listnn=listnn, nlv=range(0, nlv + 1)).pred
        ref = oracle.locwlv(Xtr, Ytr, X, listnn=listnn, nlv=range(0, nlv + 1))
        for a in range(nlv + 1):
            assert relerr(got[a], ref[a]) < 1e-11, (k, a)


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,q,nlv,scal", [
    (300, 5, 2, 5, False),       # p < 16: most CTAs of the cluster own an empty slice
    (400, 17, 16, 10, True),     # widest warp-level eigenproblem, odd p
    (400, 33, 17, 8, False),     # q > 16: portable 8-CTA form with the generic eigen path
    (3000, 700, 3, 40, True),    # XtX slice too large for shared memory: rows from L2
    (2500, 610, 1, 12, False),   # q = 1, shared-memory slice at its size limit
    (1500, 100, 4, 90, False),   # nlv > 64: several rounds of the dot exchange
])
@pytest.mark.parametrize("linear", [True, False])
def test_lvloop_forms(jc, n, p, q, nlv, scal, linear, monkeypatch):
    """Every form of the LV-loop kernel (the linear form that builds Rho / Zeta beside the eigenvector iteration;
    the four-exchange 16-CTA cluster form with the XtX slice in shared memory or in L2; the portable 8-CTA
    fallback) against the oracle, on data with a real X-Y relation so that every requested LV is well determined."""
    monkeypatch.setenv("JCB_LV_LINEAR", "1" if linear else "0")
    rng = np.random.default_rng(7)
    X = synth.synth_matrix(1, n, p)
    B = rng.standard_normal((p, q))
    Y = X @ B + 0.1 * synth.synth_matrix(2, n, q)
    fm = jc.plskern(X, Y, nlv=nlv, scal=scal)
    ref = oracle.plskern(X, Y, nlv=nlv, scal=scal)
    k = min(nlv, p)
    assert fm.T.shape == (n, k)
    # LVs far beyond the signal's rank are rounding noise in the reference too: grade the leading ones
    lead = min(k, max(q, 6))
    s = oracle.sign_align(ref, fm)
    assert relerr((fm.T * s)[:, :lead], ref.T[:, :lead]) < TOL
    assert relerr((fm.W * s)[:, :lead], ref.W[:, :lead]) < TOL
    assert relerr(fm.TT[:lead], ref.TT[:lead]) < TOL
    assert relerr(jc.coef(fm, nlv=lead).B, oracle.coef(ref, nlv=lead)[0]) < TOL
    assert relerr(jc.predict(fm, X[:64]).pred, oracle.predict(ref, X[:64])) < TOL
    np.testing.assert_allclose(np.linalg.norm(fm.W, axis=0), 1, atol=1e-12)
    assert np.abs(fm.P[:, :lead].T @ fm.R[:, :lead] - np.eye(lead)).max() < TOL


@pytest.mark.parametrize("weighted", [False, True])
def test_streamed_host_fit_matches_resident_fit(jc, weighted):
    """sharded.fit_sharded_from_host (rows streamed in chunks from page-locked host memory under K1, scores
    copied back in row blocks) against the same fit on device-resident inputs."""
    import torch
    from jchemo_b200 import device as dev, sharded
    n, p, q, nlv = 400_001, 48, 3, 7
    torch.cuda.set_device(0)
    dev.init(0)
    dev.use_current_stream()
    try:
        X = dev.colmajor_empty(n, p)
        Y = dev.colmajor_empty(n, q)
        dev.fill_uniform(X, n, 1)
        dev.fill_uniform(Y, n, 2)
        w = hw = w2 = None
        if weighted:
            w = torch.empty((1, dev.even_up(n)), dtype=torch.float64, device="cuda")
            dev.fill_uniform(w, n, 3)
            w = (w + 0.5).reshape(-1)
            hw = w[:n].cpu().pin_memory()
            w2 = torch.zeros_like(w)
        hX = torch.empty((p, n), dtype=torch.float64).pin_memory()
        hY = torch.empty((q, n), dtype=torch.float64).pin_memory()
        hX.copy_(X[:, :n])
        hY.copy_(Y[:, :n])
        m0 = dev.DeviceModel(n, p, q, nlv)
        sharded.fit_sharded(X, Y, w, n, m0, scal=weighted)
        X2 = torch.zeros_like(X)
        Y2 = torch.zeros_like(Y)
        m1 = dev.DeviceModel(n, p, q, nlv)
        hT = torch.empty((nlv, n), dtype=torch.float64).pin_memory()
        assert len(sharded.chunk_bounds(n)) - 1 == 10
        sharded.fit_sharded_from_host(hX, hY, hw, X2, Y2, w2, n, m1, scal=weighted, hT=hT)
        torch.cuda.synchronize()
        s = torch.sign((m0.W[:nlv] * m1.W[:nlv]).sum(dim=1))
        T0 = (m0.T[:nlv, :n] * s[:, None]).cpu().numpy()
        assert relerr(hT.numpy(), T0) < 1e-11
        assert relerr(m1.T[:nlv, :n].cpu().numpy(), T0) < 1e-11
        assert relerr(m1.xmeans.cpu().numpy(), m0.xmeans.cpu().numpy()) < 1e-13
        assert relerr(m1.xscales.cpu().numpy(), m0.xscales.cpu().numpy()) < 1e-13
        assert relerr(m1.weights[:n].cpu().numpy(), m0.weights[:n].cpu().numpy()) < 1e-14
        B0 = (m0.R[:nlv].T @ m0.C[:nlv]).cpu().numpy()
        B1 = (m1.R[:nlv].T @ m1.C[:nlv]).cpu().numpy()
        assert relerr(B1, B0) < 1e-11
    finally:
        dev.use_own_stream()


def test_chunked_host_paths_on_small_inputs(jc, monkeypatch):
    """The row-chunk machinery of the host paths (chunked H2D under K1 with the cut last chunk, row-block
    score copy-back, three-stream pipelines of predict / transform / xfit) normally starts at 400 000 rows;
    JCB_CHUNK_MIN_ROWS forces it on a small ragged input, with page-locked outputs, against the oracle."""
    monkeypatch.setenv("JCB_CHUNK_MIN_ROWS", "64")
    monkeypatch.setenv("JCB_PINNED_MIN_BYTES", "1")
    n, p, q, nlv, m = 1003, 37, 3, 6, 517
    X = synth.synth_matrix(1, n, p) + np.arange(p)[None, :]
    Y = synth.synth_matrix(2, n, q) + X[:, :q]
    w = synth.synth_weights(n, uniform=False)
    Xnew = synth.synth_matrix(4, m, p) + np.arange(p)[None, :]
    for scal in (False, True):
        fm = jc.plskern(X, Y, w, nlv=nlv, scal=scal)
        ref = oracle.plskern(X, Y, w, nlv=nlv, scal=scal)
        s = oracle.sign_align(ref, fm)
        assert relerr(fm.T * s, ref.T) < TOL
        assert relerr(fm.weights, ref.weights) < 1e-14
        assert relerr(jc.transform(fm, Xnew) * s, oracle.transform(ref, Xnew)) < TOL
        pr = jc.predict(fm, Xnew, nlv=range(0, nlv + 1)).pred
        rr = oracle.predict(ref, Xnew, nlv=range(0, nlv + 1))
        assert max(relerr(a, b) for a, b in zip(pr, rr)) < TOL
        assert relerr(jc.predict(fm, Xnew, nlv=3).pred, oracle.predict(ref, Xnew, nlv=3)) < TOL
        assert relerr(jc.xfit(fm, Xnew, nlv=4), oracle.xfit(ref, Xnew, nlv=4)) < TOL
        E = jc.xresid(fm, Xnew)
        assert np.abs(E - oracle.xresid(ref, Xnew)).max() < 1e-9 * np.abs(Xnew).max()
    # plskern!: the centred / scaled X and Y come back through the same chunked path
    Xb, Yb = np.asfortranarray(X.copy()), np.asfortranarray(Y.copy())
    jc.plskern_bang(Xb, Yb, w, nlv=nlv, scal=True)
    Xr, Yr = X.copy(order="F"), Y.copy(order="F")
    oracle.plskern_bang(Xr, Yr, w, nlv=nlv, scal=True)
    assert relerr(Xb, Xr) < 1e-12 and relerr(Yb, Yr) < 1e-12


def test_device_single_k_prediction_matches_sweep(jc):
    """Device API: a single-k prediction (coef + narrow GEMM, the reference's own arithmetic) against the last
    matrix of the sweep 0:k and against the oracle."""
    import torch
    from jchemo_b200 import device as dev, sharded
    n, p, q, nlv, m = 5000, 70, 3, 9, 777
    torch.cuda.set_device(0)
    dev.init(0)
    dev.use_current_stream()
    try:
        X, Y = synth.synth_matrix(1, n, p), synth.synth_matrix(2, n, q)
        Y = Y + X[:, :q]
        Xn = synth.synth_matrix(4, m, p)
        dX, dY, dXn = dev.colmajor_empty(n, p), dev.colmajor_empty(n, q), dev.colmajor_empty(m, p)
        dX[:, :n].copy_(torch.from_numpy(np.ascontiguousarray(X.T)))
        dY[:, :n].copy_(torch.from_numpy(np.ascontiguousarray(Y.T)))
        dXn[:, :m].copy_(torch.from_numpy(np.ascontiguousarray(Xn.T)))
        model = dev.DeviceModel(n, p, q, nlv)
        sharded.fit_sharded(dX, dY, None, n, model)
        for k in (1, 5, nlv):
            one = dev.predict_sweep_dev(dXn, m, model, k, k)
            swp = dev.predict_sweep_dev(dXn, m, model, 0, k)
            torch.cuda.synchronize()
            a, b = one[0].cpu().numpy().T, swp[k].cpu().numpy().T
            assert relerr(a, b) < 1e-12
            ref = oracle.plskern(X, Y, nlv=nlv)
            assert relerr(a, oracle.predict(ref, Xn, nlv=k)) < TOL
    finally:
        dev.use_own_stream()
