// K9 — batched tiny kernel-PLS fits for the local (kNN) models: SURVEY 8f rank 4.
//
// Replaces the body of locwlv's thread loop (/root/reference/src/locwlv.jl:18-40) for fun = plskern:
// for every row i of X, `fun(Xtrain[s, :], Ytrain[s, :], listw[i]; nlv = max)` on its k neighbours
// s = listnn[i], then `predict(fm, X[i:i, :]; nlv = a)` for every a.  Called one by one through the big
// fit path these ~100-row problems would each pay a dozen launches and a host round trip, so they get
// their own kernel: ONE CTA per query row runs the whole fit + predictions.
//
// With k << p (e.g. 100 neighbours, 1000 wavelengths) the p x p Gram is the wrong tool; the CTA runs the
// reference's own recurrences (Dayal & MacGregor kernel #1, plskern.jl:149-175) on the gathered k x p
// slab: t = Xs r, zp = Xs'(D t), two passes per LV over a slab that sits in L2 (or L1).  The slab lives in
// a per-CTA global scratch, column-major with the k rows contiguous, so both passes are coalesced.
#include <algorithm>

#include "eig.cuh"
#include "jcb_internal.cuh"

namespace jcb {

constexpr int LW_THREADS = 1024;   // one CTA per SM: 32 warps hide the L2 latency of the two slab passes per LV
constexpr int LW_WARPS = LW_THREADS / 32;

struct LocwParams {
    const double* Xtr;
    int64_t ldxt;
    const double* Ytr;
    int64_t ldyt;
    const double* X;
    int64_t ldx;
    int64_t m;
    int p, q;
    const int64_t* nn_idx;   // concatenated neighbour rows (zero based)
    const int64_t* nn_off;   // m + 1 offsets
    const double* nn_w;      // concatenated weights or nullptr
    int k_lo, k_hi, scal, kmax, amax, tpsz;   // kmax: even upper bound of the slab pitch; tpsz: doubles in tp_s
    double* scratch;         // per CTA: Xs kmax*p | P p*amax | R p*amax
    int64_t scratch_stride;
    double* pred;            // m x q x nk (column-major)
};

__device__ __forceinline__ double lw_block_sum(double v, double* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (lane < LW_WARPS) ? red[lane] : 0.0;
    return warp_sum(t);
}

__global__ void __launch_bounds__(LW_THREADS, 1) locw_plskern_kernel(const LocwParams prm) {
    extern __shared__ __align__(16) double sm[];
    const int p = prm.p, q = prm.q, kmax = prm.kmax, amax = prm.amax;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // shared layout
    double* wts = sm;                    // kmax
    double* t_s = wts + kmax;            // kmax
    double* dt_s = t_s + kmax;           // kmax
    double* tp_s = dt_s + kmax;          // tpsz: partial t per column group
    double* Ys = tp_s + prm.tpsz;        // kmax * q
    double* xm = Ys + kmax * q;          // p
    double* xsc = xm + p;                // p
    double* w_s = xsc + p;               // p
    double* r_s = w_s + p;               // p
    double* zp_s = r_s + p;              // p
    double* xty = zp_s + p;              // p * q
    double* C_s = xty + p * q;           // q * amax
    double* M_s = C_s + q * amax;        // q*q
    double* A_s = M_s + q * q;           // q*q
    double* B_s = A_s + q * q;           // q*q
    double* v_s = B_s + q * q;           // q
    double* c_s = v_s + q;               // q
    double* ym = c_s + q;                // q
    double* ysc = ym + q;                // q
    double* d_s = ysc + q;               // amax
    double* tn_s = d_s + amax;           // amax
    double* red = tn_s + amax;           // 32
    __shared__ int same_s;

    for (int64_t i = blockIdx.x; i < prm.m; i += gridDim.x) {
        const int64_t o0 = prm.nn_off[i];
        const int k = (int)(prm.nn_off[i + 1] - o0);
        const int64_t* idx = prm.nn_idx + o0;
        const int kp = (k + 1) & ~1;                                             // even slab pitch: 16-byte rows pairs
        double* Xs = prm.scratch + (int64_t)blockIdx.x * prm.scratch_stride;     // k x p, ld = kp
        double* Pg = Xs + (int64_t)kmax * p;
        double* Rg = Pg + (int64_t)p * amax;
        const int a_fit = min(min(k, p), prm.k_hi);                              // plskern.jl:116
        const int nk = prm.k_hi - prm.k_lo + 1;
        __syncthreads();
        // ---- weights (mweight, utility.jl:715-723) and Y of the neighbours
        double s = 0.0;
        for (int r = tid; r < k; r += LW_THREADS) {
            const double wv = prm.nn_w ? prm.nn_w[o0 + r] : 1.0;
            wts[r] = wv;
            s += wv;
        }
        const double sw = lw_block_sum(s, red);
        for (int r = tid; r < k; r += LW_THREADS) wts[r] /= sw;
        for (int e = tid; e < k * q; e += LW_THREADS) {
            const int c = e / k, r = e - c * k;
            Ys[e] = prm.Ytr[idx[r] + (int64_t)c * prm.ldyt];
        }
        if (tid == 0) same_s = 1;
        __syncthreads();
        // locwlv.jl:24-28: q == 1 and all neighbours share one Y value -> that value for every nlv
        if (q == 1) {
            for (int r = tid; r < k; r += LW_THREADS)
                if (Ys[r] != Ys[0]) same_s = 0;
        }
        __syncthreads();
        if (q == 1 && same_s) {
            for (int e = tid; e < nk; e += LW_THREADS) prm.pred[i + (int64_t)e * prm.m * q] = Ys[0];
            continue;
        }
        // ---- Y: means, scales, centre/scale in shared memory
        for (int c = warp; c < q; c += LW_WARPS) {
            double a0 = 0.0;
            for (int r = lane; r < k; r += 32) a0 += wts[r] * Ys[c * k + r];
            a0 = warp_sum(a0);
            double v0 = 0.0;
            for (int r = lane; r < k; r += 32) {
                const double d = Ys[c * k + r] - a0;
                v0 += wts[r] * d * d;
            }
            v0 = warp_sum(v0);
            const double sc = prm.scal ? sqrt(v0) : 1.0;
            for (int r = lane; r < k; r += 32) Ys[c * k + r] = (Ys[c * k + r] - a0) / sc;
            if (lane == 0) {
                ym[c] = a0;
                ysc[c] = sc;
            }
        }
        __syncthreads();     // the XtY accumulation below reads the centred Ys
        // ---- X: gather the neighbours' rows column by column, means, scales, centre/scale, and XtY
        for (int j = warp; j < p; j += LW_WARPS) {
            const double* src = prm.Xtr + (int64_t)j * prm.ldxt;
            double* col = Xs + (int64_t)j * kp;
            if (lane == 0 && kp > k) col[k] = 0.0;
            double a0 = 0.0;
            for (int r = lane; r < k; r += 32) {
                const double x = src[idx[r]];
                col[r] = x;
                a0 += wts[r] * x;
            }
            a0 = warp_sum(a0);
            double sc = 1.0;
            if (prm.scal) {
                double v0 = 0.0;
                for (int r = lane; r < k; r += 32) {
                    const double d = col[r] - a0;
                    v0 += wts[r] * d * d;
                }
                sc = sqrt(warp_sum(v0));
            }
            double acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.0;
            for (int r = lane; r < k; r += 32) {
                const double x = (col[r] - a0) / sc;
                col[r] = x;
                const double xw = x * wts[r];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < q) acc[c] += xw * Ys[c * k + r];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < q) {
                    const double v = warp_sum(acc[c]);
                    if (lane == 0) xty[j + c * p] = v;
                }
            for (int c = 8; c < q; ++c) {      // q > 8: one more pass per extra response
                double a1 = 0.0;
                for (int r = lane; r < k; r += 32) a1 += col[r] * wts[r] * Ys[c * k + r];
                a1 = warp_sum(a1);
                if (lane == 0) xty[j + c * p] = a1;
            }
            if (lane == 0) {
                xm[j] = a0;
                xsc[j] = sc;
            }
        }
        __syncthreads();
        // ---- LV loop (plskern.jl:149-175, the reference's own recurrences)
        for (int a = 0; a < a_fit; ++a) {
            if (q == 1) {
                double s2 = 0.0;
                for (int j = tid; j < p; j += LW_THREADS) {
                    w_s[j] = xty[j];
                    s2 += xty[j] * xty[j];
                }
                const double nrm = sqrt(lw_block_sum(s2, red));
                for (int j = tid; j < p; j += LW_THREADS) w_s[j] /= nrm;
            } else {
                for (int e = warp; e < q * (q + 1) / 2; e += LW_WARPS) {
                    int ci = 0, rem = e;
                    while (rem >= q - ci) { rem -= q - ci; ++ci; }
                    const int cj = ci + rem;
                    double s2 = 0.0;
                    for (int j = lane; j < p; j += 32) s2 += xty[j + ci * p] * xty[j + cj * p];
                    s2 = warp_sum(s2);
                    if (lane == 0) { M_s[ci * q + cj] = s2; M_s[cj * q + ci] = s2; }
                }
                __syncthreads();
                if (warp == 0) eig_dominant_warp_q(q, M_s, A_s, B_s, v_s, lane);
                __syncthreads();
                double s2 = 0.0;
                for (int j = tid; j < p; j += LW_THREADS) {
                    double tv = 0.0;
                    for (int c = 0; c < q; ++c) tv += xty[j + c * p] * v_s[c];
                    w_s[j] = tv;
                    s2 += tv * tv;
                }
                const double nrm = sqrt(lw_block_sum(s2, red));
                if (nrm > 0.0) {
                    for (int j = tid; j < p; j += LW_THREADS) w_s[j] /= nrm;
                } else {
                    for (int j = tid; j < p; j += LW_THREADS) w_s[j] = (j == 0) ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            // r = w - sum_{l<a} (w'P_l) R_l
            for (int l = warp; l < a; l += LW_WARPS) {
                double s2 = 0.0;
                for (int j = lane; j < p; j += 32) s2 += w_s[j] * Pg[j + (int64_t)l * p];
                s2 = warp_sum(s2);
                if (lane == 0) d_s[l] = s2;
            }
            __syncthreads();
            for (int j = tid; j < p; j += LW_THREADS) {
                double rv = w_s[j];
                for (int l = 0; l < a; ++l) rv -= d_s[l] * Rg[j + (int64_t)l * p];
                r_s[j] = rv;
            }
            __syncthreads();
            // t = Xs r: the slab is cut into 64-row blocks (a lane holds two rows: 16-byte loads) and, when there
            // are fewer blocks than warps, into column groups; one partial vector per column group
            const int RB = (kp + 63) >> 6;
            const int NCG = RB >= LW_WARPS ? 1 : LW_WARPS / RB;
            {
                const int cg = RB >= LW_WARPS ? 0 : warp / RB;
                if (cg < NCG) {
                    for (int rb = RB >= LW_WARPS ? warp : warp % RB; rb < RB; rb += LW_WARPS) {
                        const int row = rb * 64 + 2 * lane;
                        if (row < kp) {
                            double ax = 0.0, ay = 0.0;
                            const double* base = Xs + row;
#pragma unroll 4
                            for (int j = cg; j < p; j += NCG) {
                                const double2 x = *reinterpret_cast<const double2*>(base + (int64_t)j * kp);
                                const double rj = r_s[j];
                                ax = fma(x.x, rj, ax);
                                ay = fma(x.y, rj, ay);
                            }
                            *reinterpret_cast<double2*>(tp_s + cg * (RB * 64) + row) = make_double2(ax, ay);
                        }
                    }
                }
            }
            __syncthreads();
            double stt = 0.0;
            for (int r = tid; r < k; r += LW_THREADS) {
                double tv = 0.0;
                for (int w8 = 0; w8 < NCG; ++w8) tv += tp_s[w8 * (RB * 64) + r];
                t_s[r] = tv;
                const double d = wts[r] * tv;
                dt_s[r] = d;
                stt += tv * d;
            }
            if (tid == 0 && kp > k) dt_s[k] = 0.0;
            const double tt = lw_block_sum(stt, red);
            // c = XtY' r / tt (XtY before deflation)
            for (int c = warp; c < q; c += LW_WARPS) {
                double s2 = 0.0;
                for (int j = lane; j < p; j += 32) s2 += xty[j + c * p] * r_s[j];
                s2 = warp_sum(s2);
                if (lane == 0) c_s[c] = tt > 0.0 ? s2 / tt : 0.0;
            }
            __syncthreads();
            // zp = Xs'(D t): columns from the last to the first — the t pass ran first to last, so the tail of
            // the slab is the part most likely still in L2.  A warp takes four columns at a time (16-byte loads,
            // two rows per lane) and folds the four lane-partials with one transposing butterfly
            for (int qd = ((p + 3) >> 2) - 1 - warp; qd >= 0; qd -= LW_WARPS) {
                const int j0 = qd * 4;
                const double* c0 = Xs + (int64_t)j0 * kp;
                const double* c1 = Xs + (int64_t)min(j0 + 1, p - 1) * kp;
                const double* c2 = Xs + (int64_t)min(j0 + 2, p - 1) * kp;
                const double* c3 = Xs + (int64_t)min(j0 + 3, p - 1) * kp;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int row = 2 * lane; row < kp; row += 64) {
                    const double2 d = *reinterpret_cast<const double2*>(dt_s + row);
                    const double2 x0 = *reinterpret_cast<const double2*>(c0 + row);
                    const double2 x1 = *reinterpret_cast<const double2*>(c1 + row);
                    const double2 x2 = *reinterpret_cast<const double2*>(c2 + row);
                    const double2 x3 = *reinterpret_cast<const double2*>(c3 + row);
                    a0 = fma(x0.y, d.y, fma(x0.x, d.x, a0));
                    a1 = fma(x1.y, d.y, fma(x1.x, d.x, a1));
                    a2 = fma(x2.y, d.y, fma(x2.x, d.x, a2));
                    a3 = fma(x3.y, d.y, fma(x3.x, d.x, a3));
                }
                const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
                double k0 = (h16 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, h16 ? a0 : a2, 16);
                double k1 = (h16 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, h16 ? a1 : a3, 16);
                double kk = (h8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
                kk += __shfl_xor_sync(0xffffffffu, kk, 4);
                kk += __shfl_xor_sync(0xffffffffu, kk, 2);
                kk += __shfl_xor_sync(0xffffffffu, kk, 1);
                // lanes 0-7 hold column j0, 8-15 j0 + 1, 16-23 j0 + 2, 24-31 j0 + 3
                if ((lane & 7) == 0 && j0 + (lane >> 3) < p) zp_s[j0 + (lane >> 3)] = kk;
            }
            __syncthreads();
            // XtY -= zp c'; P_a = zp / tt; R_a = r
            for (int j = tid; j < p; j += LW_THREADS) {
                const double z = zp_s[j];
                for (int c = 0; c < q; ++c) xty[j + c * p] -= z * c_s[c];
                Pg[j + (int64_t)a * p] = tt > 0.0 ? z / tt : 0.0;
                Rg[j + (int64_t)a * p] = r_s[j];
            }
            for (int c = tid; c < q; c += LW_THREADS) C_s[c + a * q] = c_s[c];
            __syncthreads();
        }
        // ---- predictions for the query row: t_new = ((x - xmeans)/xscales) R, pred_a = ymeans + ys .* C t_new
        for (int l = warp; l < a_fit; l += LW_WARPS) {
            double s2 = 0.0;
            for (int j = lane; j < p; j += 32)
                s2 += (prm.X[i + (int64_t)j * prm.ldx] - xm[j]) / xsc[j] * Rg[j + (int64_t)l * p];
            s2 = warp_sum(s2);
            if (lane == 0) tn_s[l] = s2;
        }
        __syncthreads();
        for (int e = tid; e < nk * q; e += LW_THREADS) {
            const int kk = e / q, c = e - kk * q;
            const int use = min(prm.k_lo + kk, a_fit);           // predict clamps nlv to the model's LVs (:229)
            double pv = 0.0;
            for (int l = 0; l < use; ++l) pv += tn_s[l] * C_s[c + l * q];
            prm.pred[i + (int64_t)c * prm.m + (int64_t)kk * prm.m * q] = ym[c] + ysc[c] * pv;
        }
    }
}

int launch_locw(Ctx* c, const double* dXtr, int64_t ldxt, const double* dYtr, int64_t ldyt, int64_t ntr,
                const double* dX, int64_t ldx, int64_t m, int64_t p, int64_t q, const int64_t* d_idx,
                const int64_t* d_off, const double* d_w, int kmax, int k_lo, int k_hi, int scal,
                double* d_pred) {
    (void)ntr;
    kmax = (kmax + 1) & ~1;                                   // even slab pitch and shared-memory strides
    const int tpsz = std::max(kmax, LW_WARPS * 64);
    const int amax = std::max(1, std::min<int>(std::min<int64_t>(kmax, p), k_hi));
    const size_t smem = (size_t)(3 * kmax + tpsz + kmax * q + 5 * p + p * q + q * amax + 3 * q * q +
                                 4 * q + 2 * amax + 32) * 8;
    if (q > 16 || smem > 220 * 1024) {
        set_error("locw: problem too large for the batched kernel (kmax=%d p=%lld q=%lld needs %zu bytes of "
                  "shared memory, q <= 16)", kmax, (long long)p, (long long)q, smem);
        return JCB200_EINVAL;
    }
    const int grid = (int)std::min<int64_t>(m, (int64_t)c->num_sms);
    const int64_t stride = ((int64_t)kmax * p + 2 * p * amax + 1) & ~(int64_t)1;
    JCB_TRY(ensure(c->locw_ws, (size_t)grid * stride * 8));
    LocwParams prm;
    prm.Xtr = dXtr; prm.ldxt = ldxt; prm.Ytr = dYtr; prm.ldyt = ldyt; prm.X = dX; prm.ldx = ldx; prm.m = m;
    prm.p = (int)p; prm.q = (int)q; prm.nn_idx = d_idx; prm.nn_off = d_off; prm.nn_w = d_w;
    prm.k_lo = k_lo; prm.k_hi = k_hi; prm.scal = scal; prm.kmax = kmax; prm.amax = amax; prm.tpsz = tpsz;
    prm.scratch = (double*)c->locw_ws.p; prm.scratch_stride = stride; prm.pred = d_pred;
    JCB_CUDA(cudaFuncSetAttribute(locw_plskern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    locw_plskern_kernel<<<grid, LW_THREADS, smem, c->stream>>>(prm);
    JCB_LAUNCH_CHECK();
    return 0;
}

}  // namespace jcb
