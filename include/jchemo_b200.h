/*
 * jchemo_b200.h — C ABI of libjchemo_b200.so (sm_100a), the B200-native replacement for the
 * kernel-PLS path of Jchemo.jl.
 *
 * The reference has no FFI of its own: the boundary it exposes is the exported Julia function API
 *   plskern / plskern!            /root/reference/src/plskern.jl:106-178   (export: src/Jchemo.jl:247)
 *   transform / coef / predict    /root/reference/src/plskern.jl:187-238   (export: src/Jchemo.jl:302)
 * Each entry point below names the reference lines it replaces.  A Julia (or ctypes) host binds
 * these with `ccall`; see INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *  - every matrix is IEEE Float64, column-major, with an explicit leading dimension (elements);
 *  - every array is caller-allocated and caller-owned; the library keeps no host pointer after return;
 *  - return value: 0 = success, < 0 = bad argument, > 0 = CUDA failure (value is the cudaError_t);
 *    the message is available, per calling thread, from jcb200_last_error();
 *  - there is NO CPU fallback: without a usable sm_100 device every compute call fails;
 *  - entry points are re-entrant: one process-wide mutex serialises device work (the GPU is one
 *    resource; the reference's callers invoke `fun` from Threads.@threads loops, src/locwlv.jl:18);
 *  - "_dev" entry points take DEVICE pointers (benchmarking without PCIe; multi-process sharding),
 *    the others take HOST pointers and include the host<->device copies.
 */
#ifndef JCHEMO_B200_H
#define JCHEMO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JCB200_VERSION 200 /* 0.2.0 */

/* status codes (< 0: argument errors) */
#define JCB200_OK 0
#define JCB200_EINVAL (-1)   /* n, p, q <= 0; nlv < 0; ld < rows; NULL where an array is required   */
#define JCB200_ENODEV (-2)   /* no CUDA device, or compute capability != 10.x                      */
#define JCB200_ENOMEM (-3)   /* device allocation failed                                           */
#define JCB200_EALIGN (-4)   /* "_dev" pointer not 16-byte aligned or odd leading dimension         */
#define JCB200_ENONFINITE (-5) /* X, Y or the weights contain NaN / Inf (the reference throws from LAPACK's
                                  svd at /root/reference/src/plskern.jl:154 for q > 1 and returns an all-NaN
                                  model for q == 1); outputs are unspecified, plskern! leaves X, Y untouched */

/* number of phases reported by jcb200_last_timings (ms each, CUDA events on the library stream) */
#define JCB200_NPHASE 10
enum jcb200_phase {
    JCB200_T_H2D = 0,      /* host -> device copies of X, Y, w                       */
    JCB200_T_PIVOT = 1,    /* strided-sample pivot                                    */
    JCB200_T_GRAM = 2,     /* K1: fused weight/centre + DMMA SYRK (dominant kernel)   */
    JCB200_T_REDUCE = 3,   /* K1b: ordered split-K reduce into the packed buffer      */
    JCB200_T_FINALIZE = 4, /* K3: means, scales, pivot correction, mirror             */
    JCB200_T_LVLOOP = 5,   /* K4: persistent latent-variable loop                     */
    JCB200_T_SCORES = 6,   /* K5/K6: (X - mu) * M streaming GEMM (scores / predictions) */
    JCB200_T_WRITEBACK = 7,/* K7: in-place centring/scaling of X, Y (plskern! only)   */
    JCB200_T_D2H = 8,      /* device -> host copies                                   */
    JCB200_T_TOTAL = 9     /* first event to last event of the call                   */
};

/* ---- library management ------------------------------------------------------------------- */
int jcb200_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* jcb200_last_error(void);
/* Bind the library to CUDA device `device` (default 0 on first use). Idempotent. */
int jcb200_init(int device);
/* Single-process multi-GPU: bind the library to `ngpu` (1..8) peer-accessible devices of one box.  The
 * host-pointer fit then shards X, Y, w by contiguous row blocks over them: every device streams its rows
 * over its own PCIe link and builds its partial Gram; the one exchange of the path is a sum
 * of the packed partial Grams that every device reads straight out of its peers' memory over NVLink (no
 * collective library, fixed order, bit-identical on all devices); K3/K4 run redundantly, K5 per shard.
 * transform / predict / gridscore keep using device_ids[0]. */
int jcb200_init_multi(int ngpu, const int* device_ids);
/* Number of devices the library is bound to (0 before initialisation). */
int jcb200_device_count(void);
void jcb200_shutdown(void);
/* external != 0: launch on the caller's stream `cuda_stream` (a cudaStream_t; 0 is the legacy default
 * stream, e.g. torch's current stream); external == 0: back to the library's own stream. */
int jcb200_set_stream(void* cuda_stream, int32_t external);
/* Per-phase times (ms) of the last successful call; returns the number of phases written. */
int jcb200_last_timings(double* ms, int cap);
/* Synchronise the library stream and collect the phase times of the preceding "_dev" calls. */
int jcb200_sync_timings(void);
/* Phase events on (default) / off.  Each record is a stream operation of its own; callers that time whole fits
   themselves switch them off.  The K1 event ring (jcb200_gram_timings) stays on. */
int jcb200_set_phase_timing(int on);
/* Durations (ms, most recent first) of the last K1 Gram-kernel launches, from CUDA events recorded
 * around each launch on the launching stream; synchronises that stream.  Returns the count written
 * (negative/positive status codes on failure are not used here: at most `cap`, at most 256). */
int jcb200_gram_timings(double* ms, int cap);
/* Number of kernels this library has launched since load (claim for bench.py's gpu_launches). */
int64_t jcb200_launch_count(void);
/* Page-lock / unlock a caller's host array so that the copies run at full PCIe speed. */
int jcb200_host_register(void* ptr, int64_t bytes);
int jcb200_host_unregister(void* ptr);
/* Page-locked host memory from a reuse pool, for large caller-owned OUTPUT arrays (the scores T): a copy
 * into it runs at PCIe speed instead of through the pageable staging path.  Freed blocks are kept (up to
 * 3 GB) for later calls.  jcb200_host_free may be called from a finalizer thread. */
void* jcb200_host_alloc(int64_t bytes);
int jcb200_host_free(void* ptr);

/* ---- resident matrices (device-resident data handle) ------------------------------------------
 * The reference's workflows reuse one X: gridscorelv fits then predicts (/root/reference/src/gridscore.jl:179-180),
 * summary(fm, X) takes the training X (/root/reference/src/plskern.jl:246-249), gridcvlv repeats over the same
 * X, Y.  jcb200_resident_add uploads a host matrix ONCE and keys the device copy by (pointer, leading dimension,
 * shape); every host-pointer entry point below that is handed exactly that matrix as X (or Y) then skips the
 * host-to-device transfer.  The caller must not modify a resident matrix on the host between calls; calls that
 * modify it themselves keep the copies consistent (plskern! centres both) or drop the entry (xfit!, xresid!).
 * Single-device mode only (after jcb200_init_multi the sharded fit ignores the registry).  At most 16 entries. */
int jcb200_resident_add(const double* A, int64_t lda, int64_t rows, int64_t cols);
int jcb200_resident_drop(const double* A);
int jcb200_resident_count(void);

/* Facts about the last successful jcb200_plskern_fit of the calling thread: the number of latent variables
 * that carry information (TT[a] > 0 and C[:, a] != 0).  A degenerate LV — XtY deflated to exactly zero, e.g. constant y or more
 * LVs than the data carry; the reference divides 0/0 at /root/reference/src/plskern.jl:152,166 — is returned
 * inert (w = e_1, c = 0, P = 0), so predictions stay finite and equal those of the last informative LV. */
int jcb200_last_fit_info(int32_t* nlv_effective);

/* ---- host-pointer entry points (the drop-in path) ------------------------------------------- */

/* plskern!/plskern — /root/reference/src/plskern.jl:106-178 with utility.jl:76-81,193-195,262-264,
 * 312-323,482-487,715-723 folded in.  nlv is clamped to min(n, p, nlv) (:116); *nlv_out receives it.
 * w == NULL means ones(n) (:106,:112).  writeback_xy != 0 reproduces plskern!'s side effect: X and Y
 * leave centred (and scaled) in the caller's arrays (:125-129); 0 leaves them untouched (plskern).
 * Outputs (all required, column-major): T n*nlv (ld = ldt), P,R,W p*nlv (ld = p), C q*nlv (ld = q),
 * TT nlv, xmeans p, xscales p, ymeans q, yscales q, w_out n (the normalised weights, :117). */
int jcb200_plskern_fit(double* X, int64_t ldx, double* Y, int64_t ldy, const double* w,
                       int64_t n, int64_t p, int64_t q, int32_t nlv, int32_t scal,
                       int32_t writeback_xy,
                       double* T, int64_t ldt, double* P, double* R, double* W, double* C,
                       double* TT, double* xmeans, double* xscales, double* ymeans,
                       double* yscales, double* w_out, int32_t* nlv_out);

/* transform(::Plsr, X; nlv) — plskern.jl:187-195: T_out (m*nlv, ld = ldt) =
 * ((X - xmeans) ./ xscales) * R[:, 1:nlv].  R has leading dimension p.  nlv == 0 is a no-op. */
int jcb200_transform(const double* X, int64_t ldx, int64_t m, int64_t p,
                     const double* xmeans, const double* xscales, const double* R,
                     int32_t nlv, double* T_out, int64_t ldt);

/* coef(::Plsr; nlv = k) — plskern.jl:207-217: B (p*q, ld = p) = D(1/xscales) R[:,1:k] C[:,1:k]' D(yscales),
 * intercept (q) = ymeans' - xmeans' B.  k == 0 gives zeros and ymeans.  Tiny; computed on the device
 * for parity with the predict path. */
int jcb200_coef(const double* R, const double* C, const double* xmeans, const double* xscales,
                const double* ymeans, const double* yscales, int64_t p, int64_t q, int32_t k,
                double* B, double* intercept);

/* predict(::Plsr, X; nlv = k_lo:k_hi) — plskern.jl:226-238: pred_out[i] (m*q, ld = m) =
 * int_k + X * B_k for k = k_lo + i; the range is contiguous by construction (:229).
 * One pass over X for the whole range. a = number of LVs of the model (columns of R and C). */
int jcb200_predict_sweep(const double* X, int64_t ldx, int64_t m, int64_t p, int64_t q,
                         const double* R, const double* C, int32_t a,
                         const double* xmeans, const double* xscales,
                         const double* ymeans, const double* yscales,
                         int32_t k_lo, int32_t k_hi, double* const* pred_out);

/* gridscorelv fused scoring (next row) — /root/reference/src/gridscore.jl:179-185 with
 * the scores of src/scores.jl (msep :155-158, rmsep :268, ssr :426-429, bias :25-28, sep :400, r2
 * :190-195, rpd :332-335): for every k in k_lo..k_hi the residual sums of the validation set (X, Y)
 *   ssr[i + j*nk]    = sum_rows (Y[:, j] - pred_k[:, j])^2 ,   sumres[i + j*nk] = sum_rows (Y[:, j] - pred_k[:, j])
 * (k = k_lo + i, nk = k_hi - k_lo + 1, column-major nk x q), plus ysum[j] = sum Y[:, j] and
 * ysumsq[j] = sum Y[:, j]^2, in one pass over X and without materialising any prediction. */
int jcb200_gridscore(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t m, int64_t p,
                     int64_t q, const double* R, const double* C, int32_t a, const double* xmeans,
                     const double* xscales, const double* ymeans, const double* yscales, int32_t k_lo,
                     int32_t k_hi, double* ssr, double* sumres, double* ysum, double* ysumsq);

/* gridcvlv for fun = plskern, ONE repetition (next row) — /root/reference/src/gridcv.jl:
 * 187-228: for every segment j the model is fitted on the rows NOT in segment j (uniform weights) and
 * scored on segment j for nlv = k_lo..k_hi.  perm (n, zero-based) lists the rows segment by segment,
 * segment j = perm[seg_start[j] : seg_start[j+1]]; rows after seg_start[nseg] are in every training set.
 * Gram down-dating: one pass over X for all K Grams, then K solves and K scoring passes over the
 * segments (one more pass in total) instead of K fits on K row-copies.  Outputs as jcb200_gridscore, one
 * block per segment: ssr, sumres nseg x (nk x q), ysum, ysumsq nseg x q.  k_hi must not exceed
 * min(p, smallest training set).  reuse_xy != 0 keeps the device copy of X, Y of the previous call with
 * the same pointers and shape (repetitions of one CV). */
int jcb200_gridcv(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t n, int64_t p,
                  int64_t q, const int64_t* perm, const int64_t* seg_start, int32_t nseg, int32_t k_lo,
                  int32_t k_hi, int32_t scal, int32_t reuse_xy, double* ssr, double* sumres, double* ysum,
                  double* ysumsq);

/* xfit / xresid (next row) — /root/reference/src/xfit.jl:33-56 and :88-99.
 * resid == 0: out = ((X - xmeans)/xscales) R[:, 1:nlv] P[:, 1:nlv]' * diag(xscales) + xmeans  (X_fit, original
 * scale; nlv == 0 gives every row = xmeans, xfit.jl:41-45).  resid != 0: out = X - X_fit.
 * X m x p (ldx), out m x p (ldo); out may alias X (the bang forms overwrite their argument). */
int jcb200_xfit(const double* X, int64_t ldx, int64_t m, int64_t p, const double* xmeans, const double* xscales,
                const double* R, const double* P, int32_t nlv, int32_t resid, double* out, int64_t ldo);

/* locwlv for fun = plskern (next row) — /root/reference/src/locwlv.jl:9-48: for every row i
 * of X (m x p) a weighted kernel-PLS fit on its neighbours Xtrain[s_i, :], Ytrain[s_i, :] (s_i =
 * nn_idx[nn_off[i] : nn_off[i+1]], zero based; weights nn_w likewise or NULL for ones) with
 * min(k_i, p, k_hi) LVs, then the prediction of row i for every nlv in k_lo..k_hi (a model with fewer LVs
 * answers with all it has; the reference throws there — locwlv.jl:37 — and the host mirrors raise before calling).  pred is m x q x nk column-major
 * (the reference's zpred).  One CTA per row runs the whole tiny fit: thousands of fits in one launch.
 * q == 1 with identical neighbour responses returns that value (locwlv.jl:24-28).  q <= 16. */
int jcb200_locw_plskern(const double* Xtrain, int64_t ldxt, const double* Ytrain, int64_t ldyt, int64_t ntr,
                        int64_t p, int64_t q, const double* X, int64_t ldx, int64_t m, const int64_t* nn_idx,
                        const int64_t* nn_off, const double* nn_w, int32_t k_lo, int32_t k_hi, int32_t scal,
                        double* pred);

/* Base.summary(::Plsr, X) (next row) — /root/reference/src/plskern.jl:246-260: explained
 * X-variance per LV.  One pass over X for sstot = sum(weights' * ((X - xmeans)./xscales).^2); then
 * tt_adj = colsum(P.^2) .* TT, pvar = tt_adj / sstot, cumpvar = cumsum(pvar), xvar = tt_adj / n
 * (a values each).  `weights` are the model's normalised weights (n). */
int jcb200_summary(const double* X, int64_t ldx, int64_t n, int64_t p, const double* xmeans,
                   const double* xscales, const double* weights, const double* P, const double* TT,
                   int32_t a, double* xvar, double* pvar, double* cumpvar);

/* ---- device-pointer entry points (staged fit; one process per GPU shards rows) ------------- */

/* Length (doubles) of the packed partial-Gram buffer [Gxx p*p | Gxy p*q | gyy q | sx p | sy q | sw 1]:
 * the single buffer a row-sharded fit all-reduces (sum) across GPUs. */
int64_t jcb200_packed_len(int64_t p, int64_t q);

/* Strided-sample pivot c (p+q+1 doubles, device): column means of up to 4096 evenly spaced rows of
 * the shard, or all zeros when every column has mean^2 <= 64 variance (centring then costs more
 * FP64-pipe cycles than it saves digits); element p+q is 1.0 when centring is on, 0.0 otherwise.
 * Multi-GPU: rank 0 computes it and broadcasts, so all partial Grams share one pivot. */
int jcb200_pivot_dev(const double* dX, int64_t ldx, const double* dY, int64_t ldy,
                     int64_t n, int64_t p, int64_t q, double* d_pivot);

/* K1 + K1b: packed (+)= [X-c Y-c]' D [X-c Y-c] upper blocks, weighted column sums and sum(w) of this
 * row shard; dw == NULL means unit weights.  accumulate != 0 adds to d_packed instead of overwriting
 * (row chunks streamed from the host). Replaces plskern.jl:117-132 and the per-LV GEMVs :162,:167. */
int jcb200_gram_dev(const double* dX, int64_t ldx, const double* dY, int64_t ldy, const double* dw,
                    int64_t n, int64_t p, int64_t q, const double* d_pivot, double* d_packed,
                    int32_t accumulate);

/* K3 + K4 on the (all-reduced) packed buffer: means, scales, X'DX, X'DY, then the LV loop
 * (plskern.jl:118-126,149-175 in Gram form).  Device outputs: P,R,W p*nlv (ld p), C q*nlv, TT nlv,
 * xmeans, xscales p, ymeans, yscales q, dsumw TWO doubles: [0] the global sum of weights, [1] 0.0 / 1.0 =
 * the input was finite / contained NaN or Inf (the LV loop is then skipped).  nlv already clamped. */
int jcb200_solve_dev(const double* d_packed, const double* d_pivot, int64_t p, int64_t q,
                     int32_t nlv, int32_t scal, double* dP, double* dR, double* dW, double* dC,
                     double* dTT, double* dxmeans, double* dxscales, double* dymeans,
                     double* dyscales, double* dsumw);

/* K5: dOut (m*ncol, ld = ldo) = bias' + ((X - mu) ./ sigma) * M, M p*ncol (ld = ldm); bias may be NULL.
 * Scores T (M = R), transform, and single-k predictions (M = R C' D(yscales), bias = ymeans). */
int jcb200_xmul_dev(const double* dX, int64_t ldx, int64_t m, int64_t p, const double* dmu,
                    const double* dsigma, const double* dM, int64_t ldm, int32_t ncol,
                    const double* dbias, double* dOut, int64_t ldo);

/* Plumbing for the streamed sharded fit: asynchronous copy of a rows x cols block of a column-major matrix
 * between page-locked host memory and the device (one strided DMA, cudaMemcpy2DAsync) on the caller's stream;
 * to_device = 1 host -> device, 0 device -> host. */
int jcb200_copy_rows_async(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols,
                           int32_t to_device, void* cuda_stream);

/* K5 on the rows a fit was built from (T = Xc R, plskern.jl:162,170): as jcb200_xmul_dev with M = R, plus the
 * pivot buffer of the same fit (p + q + 1 doubles, written by jcb200_pivot_dev; may be NULL).  Its last
 * element is K1's centring decision: when every column has mean^2 <= 64 variance the scores are formed as
 * X M - mu'M, without a subtraction per element. */
int jcb200_scores_dev(const double* dX, int64_t ldx, int64_t n, int64_t p, int64_t q, const double* dxmeans,
                      const double* dxscales, const double* dR, int32_t nlv, const double* d_pivot,
                      double* dT, int64_t ldt);

/* K6: predictions for every k in k_lo..k_hi in one pass: dPred holds (k_hi-k_lo+1) consecutive
 * m*q matrices (ld = m, stride m*q). */
int jcb200_predict_sweep_dev(const double* dX, int64_t ldx, int64_t m, int64_t p, int64_t q,
                             const double* dR, const double* dC, int32_t a,
                             const double* dxmeans, const double* dxscales,
                             const double* dymeans, const double* dyscales,
                             int32_t k_lo, int32_t k_hi, double* dPred);

/* K7: in place X[:, j] = (X[:, j] - mu[j]) / sigma[j]  (center! / cscale!, utility.jl:76-81,482-487). */
int jcb200_center_scale_dev(double* dX, int64_t ldx, int64_t n, int64_t p, const double* dmu,
                            const double* dsigma);

/* Normalised weights (mweight, utility.jl:715-723): dw_out[i] = (dw ? dw[i] : 1) / *dsumw. */
int jcb200_weights_dev(const double* dw, int64_t n, const double* dsumw, double* dw_out);

/* K8: counter-based U[0,1) fill: element (i, j) of the n_rows*n_cols shard starting at
 * global row row0 of an n_global-row matrix gets u(seed, (row0 + i) + j * n_global). */
int jcb200_fill_uniform_dev(double* d, int64_t ld, int64_t n_rows, int64_t n_cols, uint64_t seed,
                            int64_t row0, int64_t n_global);

/* ---- peer-memory exchange for the row-sharded fit, one process per GPU -------------
 * The only exchange of the path is the sum of the packed partial Grams (and, before it, the pivot of rank 0).
 * Instead of a collective library the ranks read and write each other's HBM over NVLink / NVSwitch through
 * CUDA IPC windows: K1b's packed block is pushed, coalesced, into slot [rank] of every peer's window and a
 * flag is raised there; the next kernel on each rank waits for its `world` flags and adds the slots in rank
 * order (bit-identical on every rank, deterministic).  No host synchronisation, no NCCL call in the fit.
 *   1. every rank: jcb200_comm_create(rank, world, max_packed_len, handle)   -> 64-byte IPC handle
 *   2. exchange the handles by any means (torch.distributed.all_gather_object, a file, MPI ...)
 *   3. every rank: jcb200_comm_connect(all_handles)                         (world * 64 bytes, rank order)
 *   4. per fit:    jcb200_comm_pivot_dev ... jcb200_gram_dev ... jcb200_comm_allreduce_dev ... jcb200_solve_dev
 * All ranks must issue the same sequence of comm calls. */
#define JCB200_IPC_HANDLE_BYTES 64
int jcb200_comm_create(int32_t rank, int32_t world, int64_t max_packed_len, void* handle_out);
int jcb200_comm_connect(const void* all_handles);
int jcb200_comm_destroy(void);
/* Waits that were given up since jcb200_comm_create (a peer never arrived within ~4 s: it died or issued a
 * different sequence of calls).  A waiter that gives up poisons its result with NaN, so the fit then fails with
 * JCB200_ENONFINITE instead of hanging the GPU.  Synchronises the library stream. */
int jcb200_comm_timeouts(void);
/* Rank 0 computes the pivot of ITS rows (jcb200_pivot_dev) and publishes it; the other ranks fetch it out of
 * rank 0's window.  d_pivot (p + q + 1 doubles) is valid on every rank afterwards (stream order). */
int jcb200_comm_pivot_dev(const double* dX, int64_t ldx, const double* dY, int64_t ldy, int64_t n,
                          int64_t p, int64_t q, double* d_pivot);
/* d_packed (len doubles) <- sum over ranks of d_packed, in place, same bits on every rank (push kernel + sum
 * kernel; used when the partial Gram was accumulated over row chunks). */
int jcb200_comm_allreduce_dev(double* d_packed, int64_t len);
/* The FUSED form for device-resident shards — the exchange has no kernel of its own:
 *   jcb200_comm_gram_dev   K1 on this rank's n rows (n == 0: the rank contributes zeros), then K1b writes the reduced
 *                          block straight into slot [rank] of every rank's window over NVLink and raises the flags;
 *   jcb200_comm_solve_dev  K3 waits for the `world` flags, reads the packed Gram as the sum of the slots in rank
 *                          order, then K4 — arguments as jcb200_solve_dev without d_packed. */
int jcb200_comm_gram_dev(const double* dX, int64_t ldx, const double* dY, int64_t ldy, const double* dw, int64_t n,
                         int64_t p, int64_t q, const double* d_pivot);
int jcb200_comm_solve_dev(const double* d_pivot, int64_t p, int64_t q, int32_t nlv, int32_t scal, double* dP,
                          double* dR, double* dW, double* dC, double* dTT, double* dxmeans, double* dxscales,
                          double* dymeans, double* dyscales, double* dsumw);

/* Whole single-GPU fit on device-resident inputs (pivot, gram, solve, scores [, write-back]);
 * dT n*nlv (ld = ldt), dw_out n.  Used by bench.py for the HBM-resident `value`. */
int jcb200_plskern_fit_dev(double* dX, int64_t ldx, double* dY, int64_t ldy, const double* dw,
                           int64_t n, int64_t p, int64_t q, int32_t nlv, int32_t scal,
                           int32_t writeback_xy,
                           double* dT, int64_t ldt, double* dP, double* dR, double* dW, double* dC,
                           double* dTT, double* dxmeans, double* dxscales, double* dymeans,
                           double* dyscales, double* dw_out);

#ifdef __cplusplus
}
#endif
#endif /* JCHEMO_B200_H */
