// K3 (finalise the all-reduced Gram) and K4 (persistent latent-variable loop).
//
// K3 replaces colmean / colstd / center! / cscale! (/root/reference/src/plskern.jl:118-129,
// utility.jl:195,264,314-323) in Gram form: with S = sum(w), delta = s/S (s = weighted column sums
// about the pivot c):  means = c + delta,  Xc'DXc = G/S - delta delta',  scales = sqrt(diag) when
// scal, and the scaled cross-products D^-1 (.) D^-1.  The upper triangle is mirrored.
//
// K4 replaces the LV loop (/root/reference/src/plskern.jl:149-175).  Gram-form recurrences:
//   w  = XtY[:,1]/||.||                      (q == 1, :151-152)
//      = XtY v / ||XtY v||, v = dominant eigenvector of XtY'XtY  (q > 1; equals svd(XtY).U[:,1] up
//        to sign, :154)
//   r  = w - sum_{j<a} (w'P_j) R_j           (:156-161, classical Gram-Schmidt on the original w)
//   zp = XtX r ; tt = r'zp                   (:162-164,167 without touching X)
//   c  = XtY' r / tt                          (:165-166, XtY before deflation)
//   XtY -= zp c' ; P_a = zp/tt                (:168-169)
// Two forms of one persistent launch.  lvdist_kernel (q <= 16, the usual case): a 16-CTA cluster with every
// p-vector and p-row matrix sliced over the CTAs in shared memory and four st.async / mbarrier exchanges
// per LV.  lvloop_kernel (portable fallback, any q): a cluster of 8 CTAs, the p x p matvec and the
// deflation row-sliced across the CTAs (one or two cluster barriers per LV), everything O(p q) recomputed
// redundantly — and bit-identically — by every CTA so that no broadcast is needed.
#include <cooperative_groups.h>

#include "jcb_internal.cuh"
#include "eig.cuh"

namespace cg = cooperative_groups;

namespace jcb {

constexpr int LV_THREADS = 512;
constexpr int LV_CLUSTER = 8;

// ----------------------------------------------------------------------------------------- K3
// Element i of the packed Gram: one buffer, or the sum over the slots of the peer window in rank order.
__device__ __forceinline__ double packed_at(const PackedSrc& s, int64_t i) {
    double v = __ldcg(s.base + i);
    for (int r = 1; r < s.n; ++r) v += __ldcg(s.base + (int64_t)r * s.stride + i);
    return v;
}

// stats: delta, means, scales (ONE block), and the non-finite check: a NaN or Inf anywhere in X, Y or the
// weights reaches the weighted column sums (or their total), so testing p + q + 1 numbers covers the input.
// sumw[0] = S, sumw[1] = 0 (finite) / 1 (non-finite input: the LV loop is skipped, the host entry points fail
// with JCB200_ENONFINITE — the reference throws from svd at plskern.jl:154).
// Row-sharded fit with the fused exchange: the kernel first waits (acquire, system scope) until every rank's
// flag of this exchange is up in the own window — the slots then hold all partial Grams — and also writes the
// SUMMED column sums and sum(w) to `sums` (p + q + 1 doubles) for the second kernel.
__global__ void finalize_stats_kernel(const PackedSrc src, const double* __restrict__ pivot, int p, int q,
                                      int scal, double* __restrict__ xmeans, double* __restrict__ xscales,
                                      double* __restrict__ ymeans, double* __restrict__ yscales,
                                      double* __restrict__ sumw, double* __restrict__ delta,
                                      double* __restrict__ sums) {
    const int64_t P = p, Q = q;
    int bad = 0;
    if (src.flags && threadIdx.x < src.n) {
        const unsigned long long* f = src.flags + threadIdx.x * src.flag_stride;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= src.seq) break;
            __nanosleep(40);
            if (clock64() - t0 > 8000000000ll) {       // a peer never arrived (~4 s): fail the fit, do not hang
                bad = 1;
                if (src.timeouts) atomicAdd(src.timeouts, 1u);
                break;
            }
        }
    }
    bad = __syncthreads_or(bad);
    const int64_t o_gyy = P * P + P * Q, o_sx = o_gyy + Q, o_sy = o_sx + P;
    const double S = packed_at(src, o_sy + Q);
    bad |= !isfinite(S);
    for (int j = threadIdx.x; j < p + q; j += blockDim.x) {
        if (j < p) {
            const double sxj = packed_at(src, o_sx + j);
            bad |= !isfinite(sxj);
            const double d = sxj / S;
            delta[j] = d;
            sums[j] = sxj;
            xmeans[j] = pivot[j] + d;
            // K1 leaves acc_jj = G_jj + c_j s_j (A operand raw): remove the rank-one term first
            xscales[j] = scal ? sqrt((packed_at(src, j + (int64_t)j * P) - pivot[j] * sxj) / S - d * d) : 1.0;
        } else {
            const int k = j - p;
            const double syk = packed_at(src, o_sy + k);
            bad |= !isfinite(syk);
            const double d = syk / S;
            delta[j] = d;
            sums[j] = syk;
            ymeans[k] = pivot[j] + d;
            yscales[k] = scal ? sqrt((packed_at(src, o_gyy + k) - pivot[j] * syk) / S - d * d) : 1.0;
        }
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        sums[p + q] = S;
        sumw[0] = S;
        sumw[1] = bad ? 1.0 : 0.0;
    }
}

// XtX (full, mirrored) and XtY, centred exactly and scaled
__global__ void finalize_gram_kernel(const PackedSrc src, const double* __restrict__ pivot,
                                     const double* __restrict__ delta, const double* __restrict__ sums,
                                     const double* __restrict__ xscales,
                                     const double* __restrict__ yscales, int p, int q,
                                     double* __restrict__ XtX, double* __restrict__ XtY) {
    const int64_t P = p;
    const double* sx = sums;
    const double* sy = sums + p;
    const double S = sums[p + q];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // row
    const int j = blockIdx.y;                              // column of [XtX | XtY]
    if (i >= p) return;
    if (j < p) {
        if (i > j) return;
        // K1: acc_ij = G_ij + c_i s_j (A operand raw, B operand centred about the pivot c)
        const double v = ((packed_at(src, i + (int64_t)j * P) - pivot[i] * sx[j]) / S - delta[i] * delta[j]) /
                         (xscales[i] * xscales[j]);
        XtX[i + (int64_t)j * P] = v;
        XtX[j + (int64_t)i * P] = v;
    } else {
        const int k = j - p;
        XtY[i + (int64_t)k * P] =
            ((packed_at(src, P * P + i + (int64_t)k * P) - pivot[i] * sy[k]) / S - delta[i] * delta[p + k]) /
            (xscales[i] * yscales[k]);
    }
}

// ----------------------------------------------------------------------------------------- K4

// Deterministic block-wide sum; every thread receives the result. `red` holds >= 32 doubles.
__device__ double block_sum(double v, double* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (lane < (LV_THREADS >> 5)) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

struct LvParams {
    const double* XtX;   // p x p, symmetric, read only
    double* XtY;         // p x q (K3 output); deflated in place only when it does not fit in smem
    double* zp;          // 2 * p, exchange buffer (double-buffered by LV parity)
    double* P;
    double* R;
    double* W;
    double* C;
    double* TT;
    int p, q, nlv;
    const double* status; // K3's non-finite flag (sumw[1]): non-zero = skip the loop (all CTAs alike)
    int xty_smem;        // XtY resident in shared memory (every CTA deflates its own full copy)
    double* Ppriv;       // LV_CLUSTER private copies of P and R (p x nlv each): with XtY in smem a CTA
    double* Rpriv;       // reads back only what it wrote itself, so one cluster barrier per LV suffices
#ifdef JCB_K1_TRACE
    long long* trace;    // debug builds: accumulated clock64 per phase (CTA 0, thread 0)
#endif
};

#ifdef JCB_K1_TRACE
#define LV_MARK(idx)                                                        \
    do {                                                                    \
        if (rank == 0 && tid == 0 && prm.trace) {                           \
            const long long _t = clock64();                                 \
            atomicAdd((unsigned long long*)&prm.trace[idx],                 \
                      (unsigned long long)(_t - tmark)); /* RED: no stall */ \
            tmark = _t;                                                     \
        }                                                                   \
    } while (0)
#else
#define LV_MARK(idx) do { } while (0)
#endif
// lvlin_kernel: thread 0 (the eigenvector warp) accumulates into trace[0..7], thread 32 (first of the fifteen
// builder warps) into trace[8..15]
#ifdef JCB_K1_TRACE
#define LVL_MARK(idx)                                                                          \
    do {                                                                                       \
        if (rank == 0 && (tid == 0 || tid == 32) && prm.trace) {                               \
            const long long _t = clock64();                                                    \
            atomicAdd((unsigned long long*)&prm.trace[(tid == 0 ? 0 : 8) + (idx)],             \
                      (unsigned long long)(_t - lmark));                                       \
            lmark = _t;                                                                        \
        }                                                                                      \
    } while (0)
#else
#define LVL_MARK(idx) do { } while (0)
#endif

// dot of a shared vector with a global column, lanes strided, 8 independent loads in flight.  The
// ragged last chunk is PREDICATED, not peeled: a peeled remainder loop splits the warp (p = 500: lanes
// 20..31 fell out of the vector path) and serialises up to eight L2 round trips behind each other.
__device__ __forceinline__ double warp_dot_gs(const double* __restrict__ gcol,
                                              const double* __restrict__ svec, int p, int lane) {
    double s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] = 0.0;
    for (int k0 = lane; k0 < p; k0 += 8 * 32) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (k0 + u * 32 < p) ? __ldcg(gcol + k0 + u * 32) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += v[u] * svec[min(k0 + u * 32, p - 1)];
    }
    return warp_sum(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
}

// two global columns against one shared vector: 16 loads in flight per lane (colB may be null)
__device__ __forceinline__ void warp_dot2_gs(const double* __restrict__ colA,
                                             const double* __restrict__ colB,
                                             const double* __restrict__ svec, int p, int lane,
                                             double& outA, double& outB) {
    double s[8], t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] = t[u] = 0.0;
    const bool hasA = colA != nullptr, hasB = colB != nullptr;
    for (int k0 = lane; k0 < p; k0 += 8 * 32) {
        double v[8], v2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool in = k0 + u * 32 < p;
            v[u] = (in && hasA) ? __ldcg(colA + k0 + u * 32) : 0.0;
            v2[u] = (in && hasB) ? __ldcg(colB + k0 + u * 32) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double x = svec[min(k0 + u * 32, p - 1)];
            s[u] += v[u] * x;
            t[u] += v2[u] * x;
        }
    }
    outA = warp_sum(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
    outB = warp_sum(((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7])));
}

__device__ __forceinline__ double warp_dot_ss(const double* __restrict__ a,
                                              const double* __restrict__ b, int p, int lane) {
    double s0 = 0.0, s1 = 0.0;
    int k = lane;
    for (; k + 32 < p; k += 64) {
        s0 += a[k] * b[k];
        s1 += a[k + 32] * b[k + 32];
    }
    if (k < p) s0 += a[k] * b[k];
    return warp_sum(s0 + s1);
}

__global__ void __cluster_dims__(LV_CLUSTER, 1, 1) __launch_bounds__(LV_THREADS, 1)
lvloop_kernel(const LvParams prm) {
    if (prm.status[0] != 0.0) return;          // non-finite input: every CTA leaves before the first barrier
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double sm[];
    const int p = prm.p, q = prm.q, nlv = prm.nlv;
    double* w_s = sm;             // p
    double* r_s = w_s + p;        // p
    double* zp_s = r_s + p;       // p
    double* A_s = zp_s + p;       // q*q
    double* B_s = A_s + q * q;    // q*q
    double* M_s = B_s + q * q;    // q*q
    double* N_s = M_s + q * q;    // q*q
    double* v_s = N_s + q * q;    // q
    double* c_s = v_s + q;        // q
    double* d_s = c_s + q;        // nlv (dots w'P_j)
    double* red = d_s + nlv;      // 64
    double* xty_s = red + 64;     // p*q when prm.xty_smem

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = LV_THREADS >> 5;
    const int rank = (int)cluster.block_rank();
    const int per = (p + LV_CLUSTER - 1) / LV_CLUSTER;
    const int lo = min(p, rank * per), hi = min(p, lo + per);
    const int64_t P64 = p;
    const bool xs = prm.xty_smem != 0;
    const double* Pr = xs ? prm.Ppriv + (int64_t)rank * P64 * nlv : prm.P;   // where this CTA reads P, R
    const double* Rr = xs ? prm.Rpriv + (int64_t)rank * P64 * nlv : prm.R;
    // column j of XtY: shared copy, or global (L2) when it does not fit
    if (xs) {
        for (int e = tid; e < p * q; e += LV_THREADS) xty_s[e] = prm.XtY[e];
        __syncthreads();
    }

#ifdef JCB_K1_TRACE
    long long tmark = clock64();
#endif
    for (int a = 0; a < nlv; ++a) {
        LV_MARK(9);
        // ---------------------------------------------------------------- (1) weight vector w
        if (q == 1) {
            double s = 0.0;
            for (int k = tid; k < p; k += LV_THREADS) {
                const double x = xs ? xty_s[k] : __ldcg(prm.XtY + k);
                w_s[k] = x;
                s += x * x;
            }
            // XtY == 0 (constant y, or more LVs than the data carry): the reference divides 0/0 here
            // (plskern.jl:152); like the q > 1 branch (svd of a zero matrix: U = I) take w = e_1, the LV is
            // then inert (tt = 0 -> c = 0, P = 0) and predictions stay finite
            const double nrm2 = block_sum(s, red);
            if (nrm2 > 0.0) {
                const double nrm = sqrt(nrm2);
                for (int k = tid; k < p; k += LV_THREADS) w_s[k] /= nrm;
            } else {
                for (int k = tid; k < p; k += LV_THREADS) w_s[k] = (k == 0) ? 1.0 : 0.0;
            }
        } else {
            // M = XtY' XtY (symmetric q x q): a warp takes up to 4 (i <= j) pairs per pass so that their
            // loads and reductions overlap
            const int npair = q * (q + 1) / 2;
            for (int pr0 = warp; pr0 < npair; pr0 += 4 * nwarp) {
                int pi[4], pj[4];
                double acc4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int pr = min(pr0 + u * nwarp, npair - 1);
                    int i = 0, rem = pr;
                    while (rem >= q - i) { rem -= q - i; ++i; }
                    pi[u] = i;
                    pj[u] = i + rem;
                    acc4[u] = 0.0;
                }
                if (xs) {
                    const double* c0 = xty_s + pi[0] * p; const double* d0 = xty_s + pj[0] * p;
                    const double* c1 = xty_s + pi[1] * p; const double* d1 = xty_s + pj[1] * p;
                    const double* c2 = xty_s + pi[2] * p; const double* d2 = xty_s + pj[2] * p;
                    const double* c3 = xty_s + pi[3] * p; const double* d3 = xty_s + pj[3] * p;
#pragma unroll 4
                    for (int k = lane; k < p; k += 32) {
                        acc4[0] += c0[k] * d0[k];
                        acc4[1] += c1[k] * d1[k];
                        acc4[2] += c2[k] * d2[k];
                        acc4[3] += c3[k] * d3[k];
                    }
                } else {
                    for (int k = lane; k < p; k += 32) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            acc4[u] += __ldcg(prm.XtY + k + (int64_t)pi[u] * P64) *
                                       __ldcg(prm.XtY + k + (int64_t)pj[u] * P64);
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc4[u] += __shfl_xor_sync(0xffffffffu, acc4[u], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (pr0 + u * nwarp < npair) {
                            M_s[pi[u] * q + pj[u]] = acc4[u];
                            M_s[pj[u] * q + pi[u]] = acc4[u];
                        }
                    }
                }
            }
            __syncthreads();
            LV_MARK(0);
            if (q <= 16) {
                if (warp == 0) eig_dominant_warp_q(q, M_s, A_s, B_s, v_s, lane);
                __syncthreads();
            } else {
                // generic q: one thread per entry, ping-pong buffers, one named barrier per squaring,
                // trace renormalisation + convergence check every third squaring (as above)
                const int qq = q * q;
                const int ew = min(nwarp, (qq + 31) / 32);          // warps that take part
                const int eth = ew * 32;
                if (warp < ew) {
                    double tr = 0.0;
                    for (int i = 0; i < q; ++i) tr += M_s[i * q + i];
                    double* cur = A_s;
                    double* nxt = B_s;
                    for (int e = tid; e < qq; e += eth) {
                        cur[e] = M_s[e] / tr;
                        N_s[e] = cur[e];            // last normalised state, for the convergence check
                    }
                    asm volatile("bar.sync 1, %0;" ::"r"(eth) : "memory");
                    for (int iter = 0; iter < 90; ++iter) {
                        for (int e = tid; e < qq; e += eth) {
                            const int i = e / q, j = e - i * q;
                            const double* ai = cur + i * q;
                            const double* aj = cur + j * q;
                            double s0 = 0.0, s1 = 0.0;
                            int k = 0;
                            for (; k + 1 < q; k += 2) {
                                s0 += ai[k] * aj[k];
                                s1 += ai[k + 1] * aj[k + 1];
                            }
                            if (k < q) s0 += ai[k] * aj[k];
                            nxt[e] = s0 + s1;
                        }
                        asm volatile("bar.sync 1, %0;" ::"r"(eth) : "memory");
                        double* t = cur; cur = nxt; nxt = t;
                        if (iter % 3 == 2) {
                            double t2 = 0.0;
                            for (int d = 0; d < q; ++d) t2 += cur[d * q + d];
                            double chg = 0.0;
                            for (int e = tid; e < qq; e += eth) {
                                const double nv = cur[e] / t2;
                                chg = fmax(chg, fabs(nv - N_s[e]));
                                nxt[e] = nv;
                            }
#pragma unroll
                            for (int o = 16; o; o >>= 1)
                                chg = fmax(chg, __shfl_xor_sync(0xffffffffu, chg, o));
                            if (lane == 0) red[32 + warp] = chg;
                            asm volatile("bar.sync 1, %0;" ::"r"(eth) : "memory");
                            t = cur; cur = nxt; nxt = t;
                            double cmax = 0.0;
                            for (int w2 = 0; w2 < ew; ++w2) cmax = fmax(cmax, red[32 + w2]);
                            if (cmax < 1e-4) break;
                            for (int e = tid; e < qq; e += eth) N_s[e] = cur[e];   // own entries only
                        }
                    }
                    // v = column with the largest diagonal entry of the converged matrix
                    if (warp == 0) {
                        asm volatile("bar.sync 1, %0;" ::"r"(eth) : "memory");
                        int best = 0;
                        for (int i = 1; i < q; ++i) if (cur[i * q + i] > cur[best * q + best]) best = i;
                        for (int i = lane; i < q; i += 32) v_s[i] = cur[i * q + best];
                    } else {
                        asm volatile("bar.sync 1, %0;" ::"r"(eth) : "memory");
                    }
                }
                __syncthreads();
            }
            LV_MARK(1);
            LV_MARK(2);
            // w = XtY v / ||.||
            double s = 0.0;
            for (int k = tid; k < p; k += LV_THREADS) {
                double t = 0.0;
                if (xs) {
                    for (int j = 0; j < q; ++j) t += xty_s[k + (int64_t)j * P64] * v_s[j];
                } else {
                    for (int j = 0; j < q; ++j) t += __ldcg(prm.XtY + k + (int64_t)j * P64) * v_s[j];
                }
                w_s[k] = t;
                s += t * t;
            }
            const double nrm = sqrt(block_sum(s, red));
            if (nrm > 0.0) {
                for (int k = tid; k < p; k += LV_THREADS) w_s[k] /= nrm;
            } else {
                // XtY == 0 (e.g. constant Y): LAPACK's svd of a zero matrix returns U = I, so the
                // reference takes w = e_1 (plskern.jl:154) and carries on with c = 0
                for (int k = tid; k < p; k += LV_THREADS) w_s[k] = (k == 0) ? 1.0 : 0.0;
            }
        }
        __syncthreads();
        LV_MARK(3);
        // ---------------------------------------------------------------- (2) r
        for (int j = warp; j < a; j += nwarp) {
            const double s = warp_dot_gs(Pr + (int64_t)j * P64, w_s, p, lane);
            if (lane == 0) d_s[j] = s;
        }
        __syncthreads();
        for (int k = tid; k < p; k += LV_THREADS) {
            double rv = w_s[k];
            for (int j = 0; j < a; j += 8) {              // ragged last batch predicated
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[u] = (j + u < a) ? __ldcg(Rr + k + (int64_t)(j + u) * P64) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) rv -= d_s[min(j + u, a - 1)] * v[u];
            }
            r_s[k] = rv;
        }
        __syncthreads();
        LV_MARK(4);
        // ---------------------------------------------------------------- (3) zp slice = XtX[lo:hi, :] r
        double* zp_g = prm.zp + (a & 1) * P64;
        for (int i = lo + warp; i < hi; i += 2 * nwarp) {
            // two rows per pass: 16 independent L2 loads in flight per lane
            const int i2 = i + nwarp;
            const bool two = i2 < hi;
            const double* row = prm.XtX + (int64_t)i * P64;   // symmetric: row i == column i
            const double* row2 = prm.XtX + (int64_t)(two ? i2 : i) * P64;
            double s[8], t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] = t[u] = 0.0;
            for (int k = lane; k < p; k += 8 * 32) {      // ragged last chunk predicated (see warp_dot_gs)
                double v[8], v2[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool in = k + u * 32 < p;
                    v[u] = in ? row[k + u * 32] : 0.0;
                    v2[u] = in ? row2[k + u * 32] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double rk = r_s[min(k + u * 32, p - 1)];
                    s[u] += v[u] * rk;
                    t[u] += v2[u] * rk;
                }
            }
            const double z1 = warp_sum(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
            const double z2 = warp_sum(((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7])));
            if (lane == 0) {
                __stcg(zp_g + i, z1);
                if (two) __stcg(zp_g + i2, z2);
            }
        }
        // u = XtY' r (pre-deflation XtY; c = u / tt once tt is known)
        for (int j = warp; j < q; j += nwarp) {
            double t;
            if (xs) t = warp_dot_ss(xty_s + (int64_t)j * P64, r_s, p, lane);
            else t = warp_dot_gs(prm.XtY + (int64_t)j * P64, r_s, p, lane);
            if (lane == 0) c_s[j] = t;
        }
        LV_MARK(5);
        // barrier 1: zp slices visible; every CTA has finished reading the pre-deflation XtY
        cluster.sync();
        LV_MARK(6);
        // ---------------------------------------------------------------- (4) tt, c
        double s = 0.0;
        for (int k = tid; k < p; k += LV_THREADS) {
            const double z = __ldcg(zp_g + k);
            zp_s[k] = z;
            s += r_s[k] * z;
        }
        const double tt = block_sum(s, red);
        __syncthreads();
        // tt == 0 (r = 0: XtY vanished, more LVs asked than the data carry): the reference divides 0/0;
        // here the LV is inert (c = 0, P = 0) so predictions stay finite
        for (int j = tid; j < q; j += LV_THREADS) c_s[j] = tt > 0.0 ? c_s[j] / tt : 0.0;
        __syncthreads();
        LV_MARK(7);
        // ---------------------------------------------------------------- (5) deflate, store
        if (xs) {
            // every CTA deflates its own full shared copy (bit-identical everywhere)
            for (int e = tid; e < p * q; e += LV_THREADS) {
                const int j = e / p, i = e - j * p;
                xty_s[e] -= zp_s[i] * c_s[j];
            }
        } else {
            const int nsl = hi - lo;
            for (int e = tid; e < nsl * q; e += LV_THREADS) {
                const int j = e / nsl, i = lo + (e - j * nsl);
                double* dst = prm.XtY + i + (int64_t)j * P64;
                __stcg(dst, __ldcg(dst) - zp_s[i] * c_s[j]);
            }
        }
        for (int i = lo + tid; i < hi; i += LV_THREADS) {
            __stcg(prm.P + i + (int64_t)a * P64, tt > 0.0 ? zp_s[i] / tt : 0.0);
            __stcg(prm.R + i + (int64_t)a * P64, r_s[i]);
            prm.W[i + (int64_t)a * P64] = w_s[i];
        }
        if (rank == 0) {
            for (int j = tid; j < q; j += LV_THREADS) prm.C[j + (int64_t)a * q] = c_s[j];
            if (tid == 0) prm.TT[a] = tt;
        }
        if (xs) {
            // private full columns a of P and R; only this CTA reads them back (next LVs)
            double* Pw = prm.Ppriv + (int64_t)rank * P64 * nlv + (int64_t)a * P64;
            double* Rw = prm.Rpriv + (int64_t)rank * P64 * nlv + (int64_t)a * P64;
            for (int i = tid; i < p; i += LV_THREADS) {
                __stcg(Pw + i, tt > 0.0 ? zp_s[i] / tt : 0.0);
                __stcg(Rw + i, r_s[i]);
            }
            __syncthreads();
        } else {
            // barrier 2: columns a of P, R and the deflated global XtY visible to the whole cluster
            cluster.sync();
        }
    }
}


// ----------------------------------------------------------------------------------------- K4, distributed form
// The same recurrences on ONE NON-PORTABLE CLUSTER OF 16 CTAs (every B200 GPC has >= 16 SMs), with every
// p-vector and p-row matrix SLICED over the CTAs (CTA c owns rows [c*per, (c+1)*per) of XtY, P, R, w, r,
// zp and of the symmetric XtX) and all of it resident in shared memory; used whenever q <= 16.
// Nothing is recomputed redundantly except the q x q eigenproblem, and nothing but the XtX slice of a
// large p is read from L2 inside the loop.  Four small all-to-all exchanges through DISTRIBUTED SHARED
// MEMORY per LV (each CTA stores its partial into slot [rank] of every CTA's buffer with st.async, which
// counts the bytes on the RECEIVER's mbarrier: no cluster barrier inside the loop, a CTA goes on as soon
// as its own 16 slots are full; every CTA sums the slots in the same fixed order, so all CTAs hold
// bit-identical values):
//   A  partial M = XtY_c' XtY_c (upper triangle)            -> M, then v = dominant eigenvector (every CTA)
//   B  partial d~ = P_c' w~_c and |w~_c|^2, w~ = XtY v       -> d~, nrm   (w is normalised together with r)
//   C  r_c = (w~_c - R_c d~)/nrm all-gathered, partial u = XtY_c' r_c
//   D  zp_c = XtX[c,:] r (XtX slice in shared memory when GS), per-CTA partials of tt = r'zp
// then c = u/tt, P_c = zp_c/tt, XtY_c -= zp_c c' locally.  Buffers and barriers need no double-buffering:
// a CTA can only send exchange k+1 after it has completed exchange k, which needs every peer's part of k,
// which each peer sends (behind a block barrier) after it has consumed its buffer of exchange k-1.
constexpr int LV16_CLUSTER = 16;
static_assert(LV_THREADS / 32 == LV16_CLUSTER, "warp w sends to CTA w: one warp per CTA of the cluster");

struct LvdLayout {
    int pe, sp, xs, Ps, Rs, w, r, zp, rfull, exA, exB, exU, exD, M, A, B, v, c, d, red, g, total;   // doubles
    int lenZ, zloc, exZ, Zs;     // ZF: this CTA's partial Z = P_c'XtY_c, the 16 slots it is exchanged through, the sum
};
__host__ __device__ inline LvdLayout lvd_layout(int p, int q, int nlv, int per, bool gs, bool zf = false) {
    LvdLayout L;
    auto ev = [](int x) { return (x + 1) & ~1; };
    const int nt = q * (q + 1) / 2;
    L.pe = ev(p);
    L.sp = per | 1;                       // odd pitch: columns of a slice fall into different banks
    int o = 0;
    L.rfull = o; o += L.pe + 2;           // first: 16-byte aligned for the double2 matvec
    L.xs = o; o += ev(q * L.sp);
    L.Ps = o; o += ev(nlv * L.sp);
    L.Rs = o; o += ev(nlv * L.sp);
    L.w = o; o += ev(L.sp);
    L.r = o; o += ev(L.sp);
    L.zp = o; o += ev(L.sp);
    L.exA = o; o += ev(LV16_CLUSTER * nt);
    L.exB = o; o += ev(LV16_CLUSTER * (nlv + 1));
    L.exU = o; o += ev(LV16_CLUSTER * q);
    L.exD = o; o += LV16_CLUSTER;
    L.M = o; o += ev(q * q);
    L.A = o; o += ev(q * q);
    L.B = o; o += ev(q * q);
    L.v = o; o += ev(q);
    L.c = o; o += ev(q);
    L.d = o; o += ev(nlv + 1);
    L.red = o; o += 32;
    L.lenZ = ev(nlv * q);
    L.zloc = o; o += zf ? L.lenZ : 0;
    L.exZ = o; o += zf ? LV16_CLUSTER * L.lenZ : 0;
    L.Zs = o; o += zf ? L.lenZ : 0;
    L.g = o; o += gs ? per * L.pe : 0;
    L.total = o;
    return L;
}

// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// 8-byte store into a peer CTA's shared memory that completes 8 tx-bytes on the PEER's mbarrier: the
// receiver learns that the data has landed by waiting on its own barrier (one-way latency, no cluster
// barrier, no release fence on the sender)
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr),
                 "l"(__double_as_longlong(v)), "r"(rbar)
                 : "memory");
}

// bulk copy of a contiguous block of this CTA's shared memory into a peer CTA's shared memory; completes `bytes`
// tx-bytes on the PEER's mbarrier.  ONE instruction per (block, destination) where st.async needs one per double:
// 13 700 st.async per LV made the first version of lvlin_kernel three times slower than lvdist_kernel.
// bytes: multiple of 16; both addresses 16-byte aligned; the source must have been fenced to the async proxy.
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}

// ZF (q > 1): the Gram-Schmidt dots d = P'w~ and |w~|^2 need no exchange of their own.  w~ = XtY v, so d = (P'XtY) v
// = Z v and |w~|^2 = v'M v: every CTA sends its partial Z_c = P_c'XtY_c (a x q doubles, one bulk DSMEM copy per
// peer) right beside its partial M, the copies and the ordered sum over the 16 slots run on warps 1-15 WHILE warp 0
// iterates the eigenvector, and exchange B (dots, send, wait, sum: 3.3 K of an LV's 20 K cycles) disappears.
template <bool GS, bool ZF>
__global__ void __launch_bounds__(LV_THREADS, 1) lvdist_kernel(const LvParams prm) {
    if (prm.status[0] != 0.0) return;          // non-finite input: every CTA leaves before the first barrier
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) uint64_t bars[6];          // A, B, C, D, B2 (the rare w = e_1 redo), Z
    const int p = prm.p, q = prm.q, nlv = prm.nlv;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = LV_THREADS >> 5;
    const int rank = (int)cluster.block_rank();
    constexpr int ncta = LV16_CLUSTER;
    const int per = (p + ncta - 1) / ncta;
    const int lo = min(p, rank * per), hi = min(p, lo + per), nsl = hi - lo;
    const LvdLayout L = lvd_layout(p, q, nlv, per, GS, ZF);
    const bool zf = ZF && q > 1;
    const int lenZ = L.lenZ;
    double* zloc = sm + L.zloc;
    double* exZ = sm + L.exZ;
    double* Zs = sm + L.Zs;
    const int pe = L.pe, sp = L.sp, nt = q * (q + 1) / 2, nd = nlv + 1;
    double* rfull = sm + L.rfull;
    double* xs = sm + L.xs;
    double* Ps = sm + L.Ps;
    double* Rs = sm + L.Rs;
    double* w_s = sm + L.w;
    double* r_s = sm + L.r;
    double* zp_s = sm + L.zp;
    double* exA = sm + L.exA;
    double* exB = sm + L.exB;
    double* exU = sm + L.exU;
    double* exD = sm + L.exD;
    double* M_s = sm + L.M;
    double* A_s = sm + L.A;
    double* B_s = sm + L.B;
    double* v_s = sm + L.v;
    double* c_s = sm + L.c;
    double* d_s = sm + L.d;
    double* red = sm + L.red;
    double* g_s = sm + L.g;
    const int64_t P64 = p;
    const uint32_t barA = smem_u32(&bars[0]), barB = smem_u32(&bars[1]), barC = smem_u32(&bars[2]),
                   barD = smem_u32(&bars[3]), barB2 = smem_u32(&bars[4]), barZ = smem_u32(&bars[5]);
    auto evn = [](int x) { return (x + 1) & ~1; };
    const uint32_t bytesA = (uint32_t)(ncta * nt * 8), bytesC = (uint32_t)((p + ncta * q) * 8),
                   bytesD = (uint32_t)(ncta * 8);

    if (tid == 0) {
        for (int b = 0; b < 6; ++b) mbar_init(&bars[b], 1);
        fence_barrier_init();
        // arm LV 0 (a peer's data may arrive before the arming of a phase: the pending arrival keeps it open)
        if (q > 1) mbar_arrive_expect_tx(&bars[0], bytesA);
        if (!zf) mbar_arrive_expect_tx(&bars[1], (uint32_t)(ncta * 8));
        if (zf && nlv > 1) mbar_arrive_expect_tx(&bars[5], (uint32_t)(ncta * evn(q) * 8));   // Z of LV 1
        mbar_arrive_expect_tx(&bars[2], bytesC);
        mbar_arrive_expect_tx(&bars[3], bytesD);
    }
    for (int e = tid; e < nsl * q; e += LV_THREADS) {
        const int j = e / nsl, i = e - j * nsl;
        xs[j * sp + i] = prm.XtY[lo + i + (int64_t)j * P64];
    }
    for (int e = p + tid; e < pe + 2; e += LV_THREADS) rfull[e] = 0.0;   // pad read by the 16-byte loop
    if (GS) {
        // row i of the symmetric XtX == column i: contiguous, coalesced
        for (int e = tid; e < nsl * pe; e += LV_THREADS) {
            const int rr = e / pe, k = e - rr * pe;
            g_s[e] = (k < p) ? prm.XtX[(int64_t)(lo + rr) * P64 + k] : 0.0;
        }
    }
    // triangle index -> (i <= j) for the M exchange
    int ti = 0, tj = 0;
    if (tid < nt) {
        int rem = tid;
        while (rem >= q - ti) { rem -= q - ti; ++ti; }
        tj = ti + rem;
    }
    int nB2 = 0;        // completed phases of the redo barrier
    __syncthreads();
    cluster.sync();     // every CTA is running and its barriers are initialised before the first remote store

    // partial dots of the w~ slice with the P slice columns j < a and with itself (j == a) -> d_s (local), then
    // sent to slot [rank] of every CTA's exB, signalling `bar` there.
    auto dots_and_send = [&](const int a, const uint32_t bar) {
        for (int j0 = warp; j0 <= a; j0 += 4 * nwarp) {
            double acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * nwarp;
                const double* col = (j < a) ? Ps + j * sp : w_s;       // j == a: the norm
                double s = 0.0;
                if (j <= a)
                    for (int i = lane; i < nsl; i += 32) s += col[i] * w_s[i];
                acc[u] = s;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j0 + u * nwarp <= a) d_s[j0 + u * nwarp] = acc[u];
            }
        }
        __syncthreads();
        if (tid <= a) {
            // thread j carries dot j to every CTA (and later replaces d_s[j] by the sum over the CTAs itself)
            const double val = d_s[tid];
            const uint32_t dst = smem_u32(exB + rank * nd + tid);
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) st_async_f64(mapa_u32(dst, cta), val, mapa_u32(bar, cta));
        }
    };

#ifdef JCB_K1_TRACE
    long long tmark = clock64();
#endif
    for (int a = 0; a < nlv; ++a) {
        const uint32_t par = a & 1;
        const bool more = a + 1 < nlv;
        LV_MARK(9);
        // ---------------------------------------------------------------- A: M = XtY'XtY, v
        if (q > 1) {
            if (tid < nt) {
                const double* ci = xs + ti * sp;
                const double* cj = xs + tj * sp;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int k = 0;
                for (; k + 3 < nsl; k += 4) {
                    s0 += ci[k] * cj[k];
                    s1 += ci[k + 1] * cj[k + 1];
                    s2 += ci[k + 2] * cj[k + 2];
                    s3 += ci[k + 3] * cj[k + 3];
                }
                for (; k < nsl; ++k) s0 += ci[k] * cj[k];
                const double mval = (s0 + s1) + (s2 + s3);
                // one destination CTA per st.async instruction (the lanes carry the entries of the triangle)
                const uint32_t dst = smem_u32(exA + rank * nt + tid);
#pragma unroll
                for (int cta = 0; cta < ncta; ++cta) st_async_f64(mapa_u32(dst, cta), mval, mapa_u32(barA, cta));
                mbar_wait(&bars[0], par);
                if (tid == 0 && more) mbar_arrive_expect_tx(&bars[0], bytesA);
                double v[ncta];
#pragma unroll
                for (int cta = 0; cta < ncta; ++cta) v[cta] = exA[cta * nt + tid];
                double s = 0.0;
#pragma unroll
                for (int cta = 0; cta < ncta; ++cta) s += v[cta];
                M_s[ti * q + tj] = s;
                M_s[tj * q + ti] = s;
            }
            __syncthreads();
            LV_MARK(0);
            if (warp == 0) {
                eig_dominant_warp_q(q, M_s, A_s, B_s, v_s, lane);
                if (zf) {
                    // |w~|^2 = v'M v
                    double sv = 0.0;
                    if (lane < q) {
                        double t0 = 0.0, t1 = 0.0;
                        int k = 0;
                        for (; k + 1 < q; k += 2) {
                            t0 += M_s[lane * q + k] * v_s[k];
                            t1 += M_s[lane * q + k + 1] * v_s[k + 1];
                        }
                        if (k < q) t0 += M_s[lane * q + k] * v_s[k];
                        sv = v_s[lane] * (t0 + t1);
                    }
                    sv = warp_sum(sv);
                    if (lane == 0) d_s[a] = sv;
                }
                LV_MARK(2);
            } else if (zf && a > 0 && (warp & 3) != 0) {
                // beside the eigenvector iteration, on the 12 warps of the three OTHER SMSPs (warp 0's own SMSP is
                // left to it): partial Z_c[j][k] = P_c[:, j]' XtY_c[:, k] for the a finished LVs, staged in zloc, sent
                // to slot [rank] of every CTA's exZ by 16 bulk copies; one warp waits for the 16 incoming slots (15
                // warps polling the barrier slowed the eigenvector warp down), then all 12 sum them in rank order
                constexpr int NWORK = (LV_THREADS / 32 - LV_THREADS / 128) * 32;      // 384
                const int wl = (warp - 1 - (warp >> 2)) * 32 + lane;
                const int nz = a * q;
                for (int e = wl; e < evn(nz); e += NWORK) {
                    double s0 = 0.0, s1 = 0.0;
                    if (e < nz) {
                        const int j = e / q, k = e - j * q;
                        const double* pc = Ps + j * sp;
                        const double* xc = xs + k * sp;
                        int i = 0;
                        for (; i + 1 < nsl; i += 2) {
                            s0 += pc[i] * xc[i];
                            s1 += pc[i + 1] * xc[i + 1];
                        }
                        if (i < nsl) s0 += pc[i] * xc[i];
                    }
                    zloc[e] = s0 + s1;
                }
                fence_proxy_async();
                asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");
                if (wl < ncta)
                    bulk_s2c(mapa_u32(smem_u32(exZ + rank * lenZ), wl), smem_u32(zloc), (uint32_t)(evn(nz) * 8),
                             mapa_u32(barZ, wl));
                if (warp == 1) {
                    mbar_wait(&bars[5], (uint32_t)((a - 1) & 1));
                    if (lane == 0 && more) mbar_arrive_expect_tx(&bars[5], (uint32_t)(ncta * evn(nz + q) * 8));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");
                for (int e = wl; e < nz; e += NWORK) {
                    double v[ncta];
#pragma unroll
                    for (int cta = 0; cta < ncta; ++cta) v[cta] = exZ[cta * lenZ + e];
                    double s = 0.0;
#pragma unroll
                    for (int cta = 0; cta < ncta; ++cta) s += v[cta];
                    Zs[e] = s;
                }
            }
            __syncthreads();
            LV_MARK(1);
        }
        // ---------------------------------------------------------------- B: w~ slice, dots d~ = P'w~, |w~|^2
        if (tid < nsl) {
            double t = xs[tid];
            if (q > 1) {
                t = 0.0;
                for (int j = 0; j < q; ++j) t += xs[j * sp + tid] * v_s[j];
            }
            w_s[tid] = t;
        }
        if (zf) {
            // d_j = Z[j, :] v (j < a) on warps 1.. (|w~|^2 = v'M v came with the eigenvector): no exchange
            if (tid >= 32 && tid - 32 < a) {
                const double* zr = Zs + (tid - 32) * q;
                double s0 = 0.0, s1 = 0.0;
                int k = 0;
                for (; k + 1 < q; k += 2) {
                    s0 += zr[k] * v_s[k];
                    s1 += zr[k + 1] * v_s[k + 1];
                }
                if (k < q) s0 += zr[k] * v_s[k];
                d_s[tid - 32] = s0 + s1;
            }
            __syncthreads();
            LV_MARK(10);
            LV_MARK(11);
            LV_MARK(12);
        } else {
            __syncthreads();
            LV_MARK(10);
            dots_and_send(a, barB);
            LV_MARK(11);
            if (tid <= a) {
                mbar_wait(&bars[1], par);
                LV_MARK(12);
                if (tid == 0 && more) mbar_arrive_expect_tx(&bars[1], (uint32_t)(ncta * (a + 2) * 8));
                double v[ncta];
#pragma unroll
                for (int cta = 0; cta < ncta; ++cta) v[cta] = exB[cta * nd + tid];
                double s = 0.0;
#pragma unroll
                for (int cta = 0; cta < ncta; ++cta) s += v[cta];
                d_s[tid] = s;
            }
            __syncthreads();
        }
        if (!(d_s[a] > 0.0)) {
            // XtY == 0 (e.g. constant Y): LAPACK's svd of a zero matrix returns U = I, so the reference
            // takes w = e_1 (plskern.jl:154) and carries on with c = 0.  Redo the dots on e_1 (all CTAs
            // take this branch together: d_s is bit-identical everywhere).
            cluster.sync();      // peers have finished reading the first pass out of their exB
            if (tid == 0) mbar_arrive_expect_tx(&bars[4], (uint32_t)(ncta * (a + 1) * 8));
            if (tid < nsl) w_s[tid] = (lo + tid == 0) ? 1.0 : 0.0;
            __syncthreads();
            dots_and_send(a, barB2);
            if (tid <= a) {
                mbar_wait(&bars[4], nB2 & 1);
                double s = 0.0;
                for (int cta = 0; cta < ncta; ++cta) s += exB[cta * nd + tid];
                d_s[tid] = s;
            }
            ++nB2;
            __syncthreads();
        }
        LV_MARK(3);
        // ---------------------------------------------------------------- C: r slice -> all CTAs, partial u
        {
            const double nrm = sqrt(d_s[a]);
            if (tid < nsl) {
                const double wv = w_s[tid];
                double r0 = wv, r1 = 0.0;
                int j = 0;
                for (; j + 1 < a; j += 2) {
                    r0 -= d_s[j] * Rs[j * sp + tid];
                    r1 -= d_s[j + 1] * Rs[(j + 1) * sp + tid];
                }
                if (j < a) r0 -= d_s[j] * Rs[j * sp + tid];
                w_s[tid] = wv / nrm;
                r_s[tid] = (r0 + r1) / nrm;
            }
        }
        __syncthreads();
        for (int e = tid; e < nsl * ncta; e += LV_THREADS) {
            const int cta = e / nsl, i = e - cta * nsl;
            st_async_f64(mapa_u32(smem_u32(rfull + lo + i), cta), r_s[i], mapa_u32(barC, cta));
        }
        for (int j = warp; j < q; j += nwarp) {
            double s = 0.0;
            for (int i = lane; i < nsl; i += 32) s += xs[j * sp + i] * r_s[i];
            s = warp_sum(s);
            if (lane == 0) c_s[j] = s;
        }
        __syncthreads();
        if (lane < q)
            st_async_f64(mapa_u32(smem_u32(exU + rank * q + lane), warp), c_s[lane], mapa_u32(barC, warp));
        mbar_wait(&bars[2], par);       // the whole r and the partial u's have landed
        LV_MARK(4);
        // ---------------------------------------------------------------- D: zp slice = XtX[lo:hi, :] r
        {
            double ttw = 0.0;
            for (int i = warp; i < nsl; i += 2 * nwarp) {
                const int i2 = i + nwarp;
                const bool two = i2 < nsl;
                double z1, z2;
                if (GS) {
                    const double2* g1 = reinterpret_cast<const double2*>(g_s + (int64_t)i * pe);
                    const double2* g2 = reinterpret_cast<const double2*>(g_s + (int64_t)(two ? i2 : i) * pe);
                    const double2* rr = reinterpret_cast<const double2*>(rfull);
                    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll 4
                    for (int k2 = lane; k2 < (pe >> 1); k2 += 32) {
                        const double2 x = rr[k2], u = g1[k2], v = g2[k2];
                        s0 += u.x * x.x;
                        s1 += u.y * x.y;
                        t0 += v.x * x.x;
                        t1 += v.y * x.y;
                    }
                    z1 = warp_sum(s0 + s1);
                    z2 = warp_sum(t0 + t1);
                } else {
                    const double* ra = prm.XtX + (int64_t)(lo + i) * P64;
                    const double* rb = prm.XtX + (int64_t)(lo + (two ? i2 : i)) * P64;
                    double s[8], t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) s[u] = t[u] = 0.0;
                    for (int k = lane; k < p; k += 8 * 32) {
                        double v[8], v2[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const bool in = k + u * 32 < p;
                            v[u] = in ? ra[k + u * 32] : 0.0;
                            v2[u] = in ? rb[k + u * 32] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const double rk = rfull[min(k + u * 32, p - 1)];
                            s[u] += v[u] * rk;
                            t[u] += v2[u] * rk;
                        }
                    }
                    z1 = warp_sum(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
                    z2 = warp_sum(((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7])));
                }
                ttw += r_s[i] * z1;
                if (two) ttw += r_s[i2] * z2;
                if (lane == 0) {
                    zp_s[i] = z1;
                    if (two) zp_s[i2] = z2;
                }
            }
            if (lane == 0) red[warp] = ttw;
        }
        __syncthreads();
        {
            const double ttc = warp_sum(lane < nwarp ? red[lane] : 0.0);     // every warp, same order
            if (lane == 0) st_async_f64(mapa_u32(smem_u32(exD + rank), warp), ttc, mapa_u32(barD, warp));
        }
        LV_MARK(5);
        mbar_wait(&bars[3], par);
        if (tid == 0 && more) {
            mbar_arrive_expect_tx(&bars[2], bytesC);     // C's phase completed above, D's just now
            mbar_arrive_expect_tx(&bars[3], bytesD);
        }
        LV_MARK(6);
        // ---------------------------------------------------------------- tt, c, deflate, store
        double tt;
        {
            // 16 per-CTA partials, summed in the same order by every warp of every CTA
            tt = warp_sum(lane < ncta ? exD[lane] : 0.0);
        }
        // tt == 0 (r = 0: XtY vanished, more LVs asked than the data carry): the reference divides 0/0;
        // here the LV is inert (c = 0, P = 0) so predictions stay finite
        if (tid < q) {
            double v[ncta];
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) v[cta] = exU[cta * q + tid];
            double u = 0.0;
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) u += v[cta];
            const double cv = tt > 0.0 ? u / tt : 0.0;
            c_s[tid] = cv;
            if (rank == 0) prm.C[tid + (int64_t)a * q] = cv;
        }
        if (rank == 0 && tid == 0) prm.TT[a] = tt;
        __syncthreads();
        LV_MARK(7);
        for (int e = tid; e < nsl * q; e += LV_THREADS) {
            const int j = e / nsl, i = e - j * nsl;
            xs[j * sp + i] -= zp_s[i] * c_s[j];
        }
        if (tid < nsl) {
            const double pv = tt > 0.0 ? zp_s[tid] / tt : 0.0;
            Ps[a * sp + tid] = pv;
            Rs[a * sp + tid] = r_s[tid];
            prm.P[lo + tid + (int64_t)a * P64] = pv;
            prm.R[lo + tid + (int64_t)a * P64] = r_s[tid];
            prm.W[lo + tid + (int64_t)a * P64] = w_s[tid];
        }
        __syncthreads();
    }
    cluster.sync();     // no CTA may exit while a peer can still store into its shared memory
}

// ----------------------------------------------------------------------------------------- K4, linear form
// lvlin_kernel — the same 16-CTA cluster, but everything downstream of the eigenvector is LINEAR in it:
//   w~ = XtY v,   r~ = w~ - R (P'w~) = Rho v   with  Rho  = XtY - R (P'XtY)         (p x q)
//   zp~ = XtX r~ = Zeta v                      with  Zeta = XtX Rho                  (p x q)
//   tt = r'zp = v'(Rho'Zeta) v / |w~|^2,  u = XtY'r = (XtY'Rho) v / |w~|,  |w~|^2 = v'M v,  M = XtY'XtY
// so Rho, Zeta and the two small q x q products do not need v: fifteen warps build them WHILE warp 0 runs the
// eigenvector iteration (half of an LV's critical path in lvdist_kernel), and three exchanges per LV
// (A: partial M and Z = XtY'P;  G: all-gather of the Rho slices;  D: partial Rho'Zeta and XtY'Rho) replace four
// that all sat behind v.  The q matvecs with XtX cost 10x the flops of the one they replace and are still off the
// critical path.  An extra column e_1 rides along with XtY's q columns: when XtY has deflated to zero (v'Mv = 0;
// LAPACK's svd of a zero matrix gives U = I, plskern.jl:154) the LV simply takes v = that column — the degenerate
// case needs no code path of its own.  Used when its buffers fit in shared memory (p x (q+1) doubles for the
// gathered Rho: p <= ~1800 at q = 10); lvdist_kernel otherwise.  Same determinism: every sum over CTAs runs in
// rank order in every CTA, so all CTAs hold the same bits.
constexpr int LVL_QAMAX = 18;            // q + 1 rounded up to even, q <= 16

struct LvlLayout {
    int pe, sp, qa, qp, nt, nta, lenA, lenD;
    int rfull, xs, Ps, Rs, rho, zeta, zp, exA, exD, M, A, B, Z, Al, Be, v, c, sc, total;   // doubles
};
__host__ __device__ inline LvlLayout lvl_layout(int p, int q, int nlv, int per) {
    LvlLayout L;
    auto ev = [](int x) { return (x + 1) & ~1; };
    L.pe = ev(p);
    L.sp = per | 1;
    L.qa = q + 1;
    L.qp = ev(q + 1);
    L.nt = q * (q + 1) / 2;
    L.nta = L.qa * (L.qa + 1) / 2;
    L.lenA = ev(L.nt + L.qa * nlv);       // even: slots stay 16-byte aligned for the bulk copies
    L.lenD = ev(L.nta + q * L.qa);
    int o = 0;
    L.rfull = o; o += L.pe * L.qp;          // first: 16-byte aligned rows for the double2 loads
    L.xs = o; o += ev(L.qa * L.sp);
    L.Ps = o; o += ev(nlv * L.sp);
    L.Rs = o; o += ev(nlv * L.sp);
    L.rho = o; o += ev(L.qa * L.sp);
    L.zeta = o; o += ev(L.qa * L.sp);
    L.zp = o; o += ev(L.sp);
    L.exA = o; o += ev(LV16_CLUSTER * L.lenA);
    L.exD = o; o += ev(LV16_CLUSTER * L.lenD);
    L.M = o; o += ev(q * q);
    L.A = o; o += ev(q * q);
    L.B = o; o += ev(q * q);
    L.Z = o; o += ev(L.qa * nlv);
    L.Al = o; o += ev(L.nta);
    L.Be = o; o += ev(q * L.qa);
    L.v = o; o += ev(L.qa);
    L.c = o; o += ev(q);
    L.sc = o; o += 8;
    L.total = o;
    return L;
}

__host__ __device__ inline int pe_qp_pad(const LvlLayout& L) { return L.pe * L.qp; }

// one row of the XtX slice (held in registers: xr[u] = XtX[row][lane + 32 u]) against the gathered Rho: all 2*NJ2
// columns at once, lanes over k, one butterfly per column at the end (every lane ends up with every sum)
template <int NJ2, int KU>
__device__ __forceinline__ void zeta_row(const double (&xr)[KU], const double* __restrict__ rfull, int qp, int p,
                                         int lane, double (&out)[2 * NJ2]) {
    double acc[2 * NJ2];
#pragma unroll
    for (int j = 0; j < 2 * NJ2; ++j) acc[j] = 0.0;
#pragma unroll
    for (int u = 0; u < KU; ++u) {
        const int k = lane + 32 * u;
        if (k < p) {
            const double2* rk = reinterpret_cast<const double2*>(rfull + (int64_t)k * qp);
#pragma unroll
            for (int j2 = 0; j2 < NJ2; ++j2) {
                const double2 r = rk[j2];
                acc[2 * j2] += xr[u] * r.x;
                acc[2 * j2 + 1] += xr[u] * r.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2 * NJ2; ++j) out[j] = warp_sum(acc[j]);
}

template <int NJ2>
__device__ __forceinline__ void zeta_rows(const double* __restrict__ XtX, int64_t P64, int lo, int nsl, int p,
                                          const double* __restrict__ rfull, int qp, int qa, int sp,
                                          double* __restrict__ zeta_s, uint64_t* barG, uint32_t par, int gwarp,
                                          int ngw, int lane, long long* tr) {
    constexpr int KU = 16;                // 32 * 16 = 512 columns per register pass
    // rows gwarp, gwarp + ngw, ... of the slice.  The first row's XtX values are requested BEFORE the wait for
    // the gathered Rho, so their L2 latency hides behind the exchange; later rows behind the previous row's FMAs.
    double xr[KU];
    int i = gwarp;
    const int npass = (p + 32 * KU - 1) / (32 * KU);
    auto load = [&](int row, int pass) {
        const double* g = XtX + (int64_t)(lo + row) * P64 + pass * 32 * KU;
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            const int k = pass * 32 * KU + lane + 32 * u;
            xr[u] = (row < nsl && k < p) ? __ldcg(g + lane + 32 * u) : 0.0;
        }
    };
    long long t0 = tr ? clock64() : 0;
    load(i, 0);
    mbar_wait(barG, par);
    if (tr) {      // debug builds: [16] request of the first row + wait for the gathered Rho, [17] the rows themselves
        const long long t1 = clock64();
        atomicAdd((unsigned long long*)&tr[16], (unsigned long long)(t1 - t0));
        t0 = t1;
    }
    for (; i < nsl; i += ngw) {
        double tot[2 * NJ2];
#pragma unroll
        for (int j = 0; j < 2 * NJ2; ++j) tot[j] = 0.0;
        for (int pass = 0; pass < npass; ++pass) {
            if (pass > 0) load(i, pass);
            double out[2 * NJ2];
            zeta_row<NJ2, KU>(xr, rfull + (int64_t)pass * 32 * KU * qp, qp, p - pass * 32 * KU, lane, out);
#pragma unroll
            for (int j = 0; j < 2 * NJ2; ++j) tot[j] += out[j];
            if (pass + 1 == npass && i + ngw < nsl) load(i + ngw, 0);      // next row while this one is stored
        }
#pragma unroll
        for (int j = 0; j < 2 * NJ2; ++j)
            if (j < qa && lane == (j & 31)) zeta_s[j * sp + i] = tot[j];
    }
    if (tr) atomicAdd((unsigned long long*)&tr[17], (unsigned long long)(clock64() - t0));
}

__global__ void __launch_bounds__(LV_THREADS, 1) lvlin_kernel(const LvParams prm) {
    if (prm.status[0] != 0.0) return;          // non-finite input: every CTA leaves before the first barrier
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) uint64_t bars[3];          // A (M, Z), G (Rho gather), D (Rho'Zeta, XtY'Rho)
    const int p = prm.p, q = prm.q, nlv = prm.nlv;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster.block_rank();
    constexpr int ncta = LV16_CLUSTER;
    const int per = (p + ncta - 1) / ncta;
    const int lo = min(p, rank * per), hi = min(p, lo + per), nsl = hi - lo;
    const LvlLayout L = lvl_layout(p, q, nlv, per);
    const int sp = L.sp, qa = L.qa, qp = L.qp, nt = L.nt, nta = L.nta, lenA = L.lenA, lenD = L.lenD;
    double* rfull = sm + L.rfull;
    double* xs = sm + L.xs;
    double* Ps = sm + L.Ps;
    double* Rs = sm + L.Rs;
    double* rho_s = sm + L.rho;
    double* zeta_s = sm + L.zeta;
    double* zp_s = sm + L.zp;
    double* exA = sm + L.exA;
    double* exD = sm + L.exD;
    double* M_s = sm + L.M;
    double* A_s = sm + L.A;
    double* B_s = sm + L.B;
    double* Z_s = sm + L.Z;
    double* Al_s = sm + L.Al;
    double* Be_s = sm + L.Be;
    double* v_s = sm + L.v;
    double* c_s = sm + L.c;
    double* sc_s = sm + L.sc;
    const int64_t P64 = p;
    const uint32_t barA = smem_u32(&bars[0]), barG = smem_u32(&bars[1]), barD = smem_u32(&bars[2]);
    // every exchange: a CTA writes its own part in place and bulk-copies it to the 15 peers
    auto evn = [](int x) { return (x + 1) & ~1; };
    const uint32_t bytesG = (uint32_t)((p - nsl) * qp * 8), bytesD = (uint32_t)((ncta - 1) * lenD * 8);

    if (tid == 0) {
        for (int b = 0; b < 3; ++b) mbar_init(&bars[b], 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bars[0], (uint32_t)((ncta - 1) * evn(nt) * 8));   // LV 0: M only (Z has no column yet)
        mbar_arrive_expect_tx(&bars[1], bytesG);
        mbar_arrive_expect_tx(&bars[2], bytesD);
    }
    for (int e = tid; e < pe_qp_pad(L); e += LV_THREADS) rfull[e] = 0.0;     // pad column of Rho stays zero
    for (int e = tid; e < nsl * q; e += LV_THREADS) {
        const int j = e / nsl, i = e - j * nsl;
        xs[j * sp + i] = prm.XtY[lo + i + (int64_t)j * P64];
    }
    for (int i = tid; i < nsl; i += LV_THREADS) xs[q * sp + i] = (lo + i == 0) ? 1.0 : 0.0;     // the e_1 column
    // triangle index -> (i <= j) for the M entries (nt <= 136 < LV_THREADS)
    int ti = 0, tj = 0;
    if (tid < nt) {
        int rem = tid;
        while (rem >= q - ti) { rem -= q - ti; ++ti; }
        tj = ti + rem;
    }
    __syncthreads();
    cluster.sync();     // every CTA is running and its barriers are initialised before the first remote store

    constexpr int NGT = LV_THREADS - 32;        // threads of warps 1..15
#ifdef JCB_K1_TRACE
    long long lmark = clock64();
#endif
    for (int a = 0; a < nlv; ++a) {
        const uint32_t par = a & 1;
        const bool more = a + 1 < nlv;
        const int nA = nt + qa * a;
        LVL_MARK(7);
        // ---------------------------------------------------------------- 1: partial M, Z -> slot [rank] everywhere
        for (int e = tid; e < nA; e += LV_THREADS) {
            const double* ci;
            const double* cj;
            if (e < nt) {
                ci = xs + ti * sp;
                cj = xs + tj * sp;
            } else {
                const int f = e - nt, j = f / qa, i = f - j * qa;
                ci = xs + i * sp;
                cj = Ps + j * sp;
            }
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = 0;
            for (; k + 3 < nsl; k += 4) {
                s0 += ci[k] * cj[k];
                s1 += ci[k + 1] * cj[k + 1];
                s2 += ci[k + 2] * cj[k + 2];
                s3 += ci[k + 3] * cj[k + 3];
            }
            for (; k < nsl; ++k) s0 += ci[k] * cj[k];
            exA[rank * lenA + e] = (s0 + s1) + (s2 + s3);          // own slot, in place
        }
        fence_proxy_async();
        __syncthreads();
        if (tid < ncta && tid != rank) {
            const uint32_t src = smem_u32(exA + rank * lenA);
            bulk_s2c(mapa_u32(src, tid), src, (uint32_t)(evn(nA) * 8), mapa_u32(barA, tid));
        }
        LVL_MARK(0);                     // partial M, Z + send
        // ---------------------------------------------------------------- 2: sums in rank order -> M, Z
        mbar_wait(&bars[0], par);
        LVL_MARK(1);                     // wait A
        if (tid == 0 && more) mbar_arrive_expect_tx(&bars[0], (uint32_t)((ncta - 1) * evn(nt + qa * (a + 1)) * 8));
        for (int e = tid; e < nA; e += LV_THREADS) {
            double v[ncta];
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) v[cta] = exA[cta * lenA + e];
            double s = 0.0;
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) s += v[cta];
            if (e < nt) {
                M_s[ti * q + tj] = s;
                M_s[tj * q + ti] = s;
            } else {
                Z_s[e - nt] = s;                    // Z_s[j * qa + i] = (column i of [XtY | e_1])' P_j
            }
        }
        __syncthreads();
        LVL_MARK(2);                     // sum A + sync
        if (warp == 0) {
            // ------------------------------------------------------------ 3a: eigenvector (one warp)
            if (q > 1) {
                eig_dominant_warp_q(q, M_s, A_s, B_s, v_s, lane);
            } else {
                if (lane == 0) v_s[0] = 1.0;
                __syncwarp();
            }
            double t = 0.0;
            if (lane < q) {
                double r = 0.0;
                for (int j = 0; j < q; ++j) r += M_s[lane * q + j] * v_s[j];
                t = r * v_s[lane];
            }
            t = warp_sum(t);                                    // |XtY v|^2 = v'M v
            const bool degen = !(t > 0.0);
            const double vl = (lane < q) ? v_s[lane] : 0.0;
            __syncwarp();
            if (lane <= q) v_s[lane] = degen ? (lane == q ? 1.0 : 0.0) : vl;      // degenerate: w = e_1
            if (lane == 0) sc_s[0] = degen ? 1.0 : t;
            __syncwarp();
            LVL_MARK(3);                 // thread 0: eigenvector
        } else {
            // ------------------------------------------------------------ 3b: Rho, Zeta, small products (15 warps)
            const int gt = tid - 32;
            for (int idx = gt; idx < qa * nsl; idx += NGT) {
                const int j = idx / nsl, i = idx - j * nsl;
                double r0 = xs[j * sp + i], r1 = 0.0;
                int l = 0;
                for (; l + 1 < a; l += 2) {
                    r0 -= Rs[l * sp + i] * Z_s[l * qa + j];
                    r1 -= Rs[(l + 1) * sp + i] * Z_s[(l + 1) * qa + j];
                }
                if (l < a) r0 -= Rs[l * sp + i] * Z_s[l * qa + j];
                const double val = r0 + r1;
                rho_s[j * sp + i] = val;
                rfull[(int64_t)(lo + i) * qp + j] = val;                // own rows of the gathered Rho, in place
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"r"(NGT) : "memory");
            if (gt < ncta && gt != rank && nsl > 0) {
                const uint32_t src = smem_u32(rfull + (int64_t)lo * qp);
                bulk_s2c(mapa_u32(src, gt), src, (uint32_t)(nsl * qp * 8), mapa_u32(barG, gt));
            }
            LVL_MARK(3);                 // thread 32: Rho + send
            // Zeta slice = XtX[lo:hi, :] Rho (waits for the gathered Rho inside, after requesting its first row)
#ifdef JCB_K1_TRACE
            long long* ztr = (rank == 0 && tid == 32) ? prm.trace : nullptr;
#else
            long long* ztr = nullptr;
#endif
            switch (qp / 2) {
#define JCB_ZR(N) case N: zeta_rows<N>(prm.XtX, P64, lo, nsl, p, rfull, qp, qa, sp, zeta_s, &bars[1], par, warp - 1, 15, lane, ztr); break;
                JCB_ZR(1) JCB_ZR(2) JCB_ZR(3) JCB_ZR(4) JCB_ZR(5) JCB_ZR(6) JCB_ZR(7) JCB_ZR(8) JCB_ZR(9)
#undef JCB_ZR
                default: break;
            }
            LVL_MARK(4);                 // thread 32: wait G + Zeta rows of warp 1
            if (gt == 0 && more) mbar_arrive_expect_tx(&bars[1], bytesG);
            asm volatile("bar.sync 1, %0;" ::"r"(NGT) : "memory");       // rho_s, zeta_s of every row are in place
            LVL_MARK(5);                 // thread 32: barrier after Zeta (slowest builder warp)
            for (int e = gt; e < nta + q * qa; e += NGT) {
                const double* ci;
                const double* cj;
                if (e < nta) {                      // upper triangle of Rho'Zeta
                    int i1 = 0, rem = e;
                    while (rem >= qa - i1) { rem -= qa - i1; ++i1; }
                    ci = rho_s + i1 * sp;
                    cj = zeta_s + (i1 + rem) * sp;
                } else {                            // XtY'Rho (q x qa)
                    const int f = e - nta, j1 = f / qa, j2 = f - j1 * qa;
                    ci = xs + j1 * sp;
                    cj = rho_s + j2 * sp;
                }
                double s0 = 0.0, s1 = 0.0;
                int k = 0;
                for (; k + 1 < nsl; k += 2) {
                    s0 += ci[k] * cj[k];
                    s1 += ci[k + 1] * cj[k + 1];
                }
                if (k < nsl) s0 += ci[k] * cj[k];
                exD[rank * lenD + e] = s0 + s1;                         // own slot, in place
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"r"(NGT) : "memory");
            if (gt < ncta && gt != rank) {
                const uint32_t src = smem_u32(exD + rank * lenD);
                bulk_s2c(mapa_u32(src, gt), src, (uint32_t)(lenD * 8), mapa_u32(barD, gt));
            }
            LVL_MARK(6);                 // thread 32: small products + send
        }
        // ---------------------------------------------------------------- 4: sums in rank order -> Rho'Zeta, XtY'Rho
        mbar_wait(&bars[2], par);
        LVL_MARK(4 + (tid == 0 ? 0 : 3));        // thread 0: [4] wait D (= the builders' branch); thread 32: [7'] wait D
        if (tid == 0 && more) mbar_arrive_expect_tx(&bars[2], bytesD);
        for (int e = tid; e < nta + q * qa; e += LV_THREADS) {
            double v[ncta];
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) v[cta] = exD[cta * lenD + e];
            double s = 0.0;
#pragma unroll
            for (int cta = 0; cta < ncta; ++cta) s += v[cta];
            if (e < nta) Al_s[e] = s;
            else Be_s[e - nta] = s;
        }
        __syncthreads();                 // also publishes v and |w~|^2 of warp 0
        LVL_MARK(5);                     // thread 0: sum D + sync
        // ---------------------------------------------------------------- 5: tt, c (one warp, same in every CTA)
        if (warp == 0) {
            const double nrm2 = sc_s[0], nrm = sqrt(nrm2);
            double t = 0.0;
            for (int e = lane; e < nta; e += 32) {
                int i1 = 0, rem = e;
                while (rem >= qa - i1) { rem -= qa - i1; ++i1; }
                const int i2 = i1 + rem;
                t += Al_s[e] * v_s[i1] * v_s[i2] * (i1 == i2 ? 1.0 : 2.0);
            }
            const double tt = warp_sum(t) / nrm2;
            // tt == 0 (r = 0: more LVs asked than the data carry): the reference divides 0/0; here the LV is inert
            if (lane < q) {
                double u = 0.0;
                for (int i = 0; i < qa; ++i) u += Be_s[lane * qa + i] * v_s[i];
                const double cv = tt > 0.0 ? (u / nrm) / tt : 0.0;
                c_s[lane] = cv;
                if (rank == 0) prm.C[lane + (int64_t)a * q] = cv;
            }
            if (lane == 0) {
                sc_s[1] = tt;
                sc_s[2] = nrm;
                if (rank == 0) prm.TT[a] = tt;
            }
        }
        __syncthreads();
        // ---------------------------------------------------------------- 6: this CTA's slices of w, r, zp; deflate
        {
            const double tt = sc_s[1], nrm = sc_s[2];
            if (tid < nsl) {
                double wv = 0.0, rv = 0.0, zv = 0.0;
                for (int j = 0; j < qa; ++j) {
                    const double vj = v_s[j];
                    wv += xs[j * sp + tid] * vj;
                    rv += rho_s[j * sp + tid] * vj;
                    zv += zeta_s[j * sp + tid] * vj;
                }
                wv /= nrm;
                rv /= nrm;
                zv /= nrm;
                const double pv = tt > 0.0 ? zv / tt : 0.0;
                zp_s[tid] = zv;
                Ps[a * sp + tid] = pv;
                Rs[a * sp + tid] = rv;
                prm.P[lo + tid + (int64_t)a * P64] = pv;
                prm.R[lo + tid + (int64_t)a * P64] = rv;
                prm.W[lo + tid + (int64_t)a * P64] = wv;
            }
        }
        __syncthreads();
        for (int e = tid; e < nsl * q; e += LV_THREADS) {
            const int j = e / nsl, i = e - j * nsl;
            xs[j * sp + i] -= zp_s[i] * c_s[j];
        }
        __syncthreads();
        LVL_MARK(6);                     // thread 0: tt, c, slices, deflate
    }
    cluster.sync();     // no CTA may exit while a peer can still store into its shared memory
}

#ifdef JCB_K1_TRACE
static long long* g_lv_trace = nullptr;
extern "C" int jcb200_debug_lv_trace(long long* host) {
    if (!g_lv_trace) return -1;
    cudaDeviceSynchronize();
    cudaMemcpy(host, g_lv_trace, 24 * sizeof(long long), cudaMemcpyDeviceToHost);
    return 24;
}
#endif

int launch_solve(Ctx* c, const double* d_packed, const double* d_pivot, int64_t p, int64_t q,
                 int nlv, int scal, double* dP, double* dR, double* dW, double* dC, double* dTT,
                 double* dxmeans, double* dxscales, double* dymeans, double* dyscales,
                 double* dsumw) {
    PackedSrc src;
    src.base = d_packed;
    src.n = 1;
    return launch_solve_src(c, src, d_pivot, p, q, nlv, scal, dP, dR, dW, dC, dTT, dxmeans, dxscales, dymeans,
                            dyscales, dsumw);
}

int launch_solve_src(Ctx* c, const PackedSrc& src, const double* d_pivot, int64_t p, int64_t q,
                     int nlv, int scal, double* dP, double* dR, double* dW, double* dC, double* dTT,
                     double* dxmeans, double* dxscales, double* dymeans, double* dyscales,
                     double* dsumw) {
    // workspace: XtX p*p | XtY p*q | delta p+q | sums p+q+2 | zp 2p | Ppriv, Rpriv LV_CLUSTER*p*nlv each (8-CTA form)
    const size_t need = (size_t)(p * p + p * q + 2 * (p + q) + 2 + 2 * p + 2 * (size_t)LV_CLUSTER * p * nlv) * 8;
    JCB_TRY(ensure(c->solve_ws, need));
    double* XtX = (double*)c->solve_ws.p;
    double* XtY = XtX + p * p;
    double* delta = XtY + p * q;
    double* sums = delta + (p + q);
    double* zp = sums + (p + q + 2);
    double* Ppriv = zp + 2 * p;
    double* Rpriv = Ppriv + (size_t)LV_CLUSTER * p * nlv;

    phase_begin(c, JCB200_T_FINALIZE);
    finalize_stats_kernel<<<1, 512, 0, c->stream>>>(
        src, d_pivot, (int)p, (int)q, scal, dxmeans, dxscales, dymeans, dyscales, dsumw, delta, sums);
    JCB_LAUNCH_CHECK();
    dim3 grid((unsigned)((p + 127) / 128), (unsigned)(p + q));
    finalize_gram_kernel<<<grid, 128, 0, c->stream>>>(src, d_pivot, delta, sums, dxscales, dyscales, (int)p,
                                                      (int)q, XtX, XtY);
    JCB_LAUNCH_CHECK();
    phase_end(c, JCB200_T_FINALIZE);

    if (nlv <= 0) return 0;
    LvParams prm;
    prm.XtX = XtX;
    prm.XtY = XtY;
    prm.zp = zp;
    prm.P = dP;
    prm.R = dR;
    prm.W = dW;
    prm.C = dC;
    prm.TT = dTT;
    prm.p = (int)p;
    prm.q = (int)q;
    prm.nlv = nlv;
    prm.status = dsumw + 1;
    prm.Ppriv = Ppriv;
    prm.Rpriv = Rpriv;
#ifdef JCB_K1_TRACE
    {
        static long long* tb = nullptr;
        if (!tb) cudaMalloc(&tb, 24 * sizeof(long long));
        cudaMemsetAsync(tb, 0, 24 * sizeof(long long), c->stream);
        prm.trace = tb;
        g_lv_trace = tb;
    }
#endif
    // distributed 16-CTA form (q <= 16), with the XtX slice in shared memory when it fits
    static int lv16_ok = -1;             // -1: not probed; 0: no 16-CTA cluster on this device
    const bool force8 = getenv("JCB_LV_CLUSTER8") != nullptr;
    const int per = (int)((p + LV16_CLUSTER - 1) / LV16_CLUSTER);
    // linear form (everything after the eigenvector is a q-term combination: Rho, Zeta built beside the
    // eigenvector iteration), OPT-IN with JCB_LV_LINEAR=1: correct on every test shape but measured SLOWER than the
    // four-exchange form at C2 (0.68 vs 0.27 ms; per-phase clock64 trace in profiles/lv_trace_r02.txt): the q + 1
    // matvecs read the gathered Rho from shared memory once per row (1.5 MB per LV at 128 B/clk), the eigenvector
    // warp runs 3.4x slower beside fifteen DFMA-heavy warps, and every exchange costs ~2 K cycles whatever it
    // carries.  Kept as the starting point for a k-split Zeta (lane = row, XtX read by symmetry, Rho broadcast).
    {
        const char* e = getenv("JCB_LV_LINEAR");
        const bool want = e && atoi(e) != 0;
        const size_t smem_lin = (size_t)lvl_layout((int)p, (int)q, nlv, per).total * 8;
        if (want && q <= 16 && per <= LV_THREADS && lv16_ok != 0 && !force8 && smem_lin <= 227 * 1024 - 64) {
            JCB_CUDA(cudaFuncSetAttribute(lvlin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lin));
            JCB_CUDA(cudaFuncSetAttribute(lvlin_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(LV16_CLUSTER);
            cfg.blockDim = dim3(LV_THREADS);
            cfg.dynamicSmemBytes = smem_lin;
            cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = LV16_CLUSTER;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            static int lin_ok = -1;
            if (lin_ok < 0) {
                int ncl = 0;
                const cudaError_t ce = cudaOccupancyMaxActiveClusters(&ncl, lvlin_kernel, &cfg);
                if (ce != cudaSuccess) (void)cudaGetLastError();
                lin_ok = (ce == cudaSuccess && ncl >= 1) ? 1 : 0;
            }
            if (lin_ok == 1) {
                prm.xty_smem = 1;
                phase_begin(c, JCB200_T_LVLOOP);
                JCB_CUDA(cudaLaunchKernelEx(&cfg, lvlin_kernel, prm));
                JCB_LAUNCH_CHECK();
                phase_end(c, JCB200_T_LVLOOP);
                return 0;
            }
        }
    }
    if (q <= 16 && per <= LV_THREADS && nlv <= LV_THREADS && lv16_ok != 0 && !force8) {
        const size_t max_smem = 227 * 1024 - 64;      // the kernel also holds five mbarriers statically
        const size_t sm_gs = (size_t)lvd_layout((int)p, (int)q, nlv, per, true).total * 8;
        const size_t sm_l2 = (size_t)lvd_layout((int)p, (int)q, nlv, per, false).total * 8;
        const bool gs = sm_gs <= max_smem && getenv("JCB_LV_NO_GS") == nullptr;
        // Z folded into exchange A (q > 1) when its 18 buffers of nlv * q doubles fit beside everything else
        // (measured, round 2: C2 0.271 -> 0.252 ms; slices of 63 rows -2 %; C4, slices of 125 rows, +8 %: the partial Z
        // costs a * q dot products over the slice per LV, beside a latency-bound eigenvector warp)
        static const int zf_per_max = getenv("JCB_LV_ZF_PER") ? atoi(getenv("JCB_LV_ZF_PER")) : 64;
        const bool zf = q > 1 && nlv <= LV_THREADS - 32 && per <= zf_per_max && getenv("JCB_LV_NO_ZF") == nullptr &&
                        (size_t)lvd_layout((int)p, (int)q, nlv, per, gs, true).total * 8 <= max_smem;
        const size_t smem16 = zf ? (size_t)lvd_layout((int)p, (int)q, nlv, per, gs, true).total * 8 : (gs ? sm_gs : sm_l2);
        if (smem16 <= max_smem) {
            auto kern = gs ? (zf ? lvdist_kernel<true, true> : lvdist_kernel<true, false>)
                           : (zf ? lvdist_kernel<false, true> : lvdist_kernel<false, false>);
            JCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            JCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(LV16_CLUSTER);
            cfg.blockDim = dim3(LV_THREADS);
            cfg.dynamicSmemBytes = smem16;
            cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = LV16_CLUSTER;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (lv16_ok < 0) {
                int ncl = 0;
                const cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
                if (e != cudaSuccess) (void)cudaGetLastError();
                lv16_ok = (e == cudaSuccess && ncl >= 1) ? 1 : 0;
            }
            if (lv16_ok == 1) {
                prm.xty_smem = 1;
                phase_begin(c, JCB200_T_LVLOOP);
                JCB_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
                JCB_LAUNCH_CHECK();
                phase_end(c, JCB200_T_LVLOOP);
                return 0;
            }
        }
    }
    // portable 8-CTA form (any q; XtY in L2 when it does not fit in shared memory)
    size_t smem = (size_t)(3 * p + 4 * q * q + 2 * q + nlv + 64) * 8;
    prm.xty_smem = (smem + (size_t)p * q * 8 <= 160 * 1024) ? 1 : 0;
    if (prm.xty_smem) smem += (size_t)p * q * 8;
    if (smem > 200 * 1024) {
        set_error("solve: p=%lld q=%lld need %zu bytes of shared memory (> 200 KB)", (long long)p,
                  (long long)q, smem);
        return JCB200_EINVAL;
    }
    JCB_CUDA(cudaFuncSetAttribute(lvloop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    phase_begin(c, JCB200_T_LVLOOP);
    lvloop_kernel<<<LV_CLUSTER, LV_THREADS, smem, c->stream>>>(prm);
    JCB_LAUNCH_CHECK();
    phase_end(c, JCB200_T_LVLOOP);
    return 0;
}

// coef (plskern.jl:207-217): B = D(1/xs) R[:,1:k] C[:,1:k]' D(ys); int = ymeans' - xmeans' B.
__global__ void coef_B_kernel(const double* __restrict__ R, const double* __restrict__ C,
                              const double* __restrict__ xs, const double* __restrict__ ys, int p,
                              int q, int k, double* __restrict__ B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= p) return;
    double s = 0.0;
    for (int a = 0; a < k; ++a) s += R[i + (int64_t)a * p] * C[j + (int64_t)a * q];
    B[i + (int64_t)j * p] = (1.0 / xs[i]) * s * ys[j];
}
__global__ void coef_int_kernel(const double* __restrict__ B, const double* __restrict__ xmeans,
                                const double* __restrict__ ymeans, int p, int q,
                                double* __restrict__ intercept) {
    const int j = blockIdx.x;
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) s += xmeans[i] * B[i + (int64_t)j * p];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) intercept[j] = ymeans[j] - t;
    }
}

int launch_coef(Ctx* c, const double* dR, const double* dC, const double* dxmeans,
                const double* dxscales, const double* dymeans, const double* dyscales, int64_t p,
                int64_t q, int k, double* dB, double* dint) {
    dim3 grid((unsigned)((p + 127) / 128), (unsigned)q);
    coef_B_kernel<<<grid, 128, 0, c->stream>>>(dR, dC, dxscales, dyscales, (int)p, (int)q, k, dB);
    JCB_LAUNCH_CHECK();
    coef_int_kernel<<<(unsigned)q, 256, 0, c->stream>>>(dB, dxmeans, dymeans, (int)p, (int)q, dint);
    JCB_LAUNCH_CHECK();
    return 0;
}

}  // namespace jcb
