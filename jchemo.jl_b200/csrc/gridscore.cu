// gridscorelv fused scoring: validation scores for EVERY number of LVs k_lo..k_hi
// without materialising the predictions.  Replaces, for fun = plskern, the predict + score loop of
//   /root/reference/src/gridscore.jl:179-185  (pred = predict(fm, X; nlv = nlv).pred; score(pred[i], Y))
// with  msep/rmsep/ssr/bias/sep/r2/rpd  of /root/reference/src/scores.jl:25-28,155-158,190-195,268,
// 332-335,400,426-429, which are all functions of the residual sums  sum_i r_k  and  sum_i r_k^2.
//
// With T = ((X - xmeans)/xscales) R (scores of the validation rows, K5) and E = Y - pred_{k_hi},
//   r_k = Y - pred_k = E + sum_{l >= k} t_l cy_l' ,   cy_l = C[:, l] .* yscales,
// so  sum r_k^2  and  sum r_k  follow from the small matrices  T'T, T'E, diag(E'E), 1'T, 1'E, which one
// launch of the Gram kernel K1 on [T | E Y] delivers (pivot 0: scores and residuals are already
// centred).  Building from the residual at the LARGEST k keeps every term of the size of the answer
// (no cancellation for good fits).  One pass over X, ~1 GB of extra traffic, nothing written per k.
#include <algorithm>
#include <vector>

#include "jcb_internal.cuh"

namespace jcb {

// Yaug[:, 0:q] = Y - ymeans - sum_{l<k} T[:, l] cy[l, :]   (residual at k LVs),  Yaug[:, q:2q] = Y
__global__ void resid_kernel(const double* __restrict__ Y, int64_t ldy, const double* __restrict__ T,
                             int64_t ldt, const double* __restrict__ C, const double* __restrict__ ys,
                             const double* __restrict__ ymeans, int64_t m, int q, int k,
                             double* __restrict__ Yaug, int64_t lda) {
    extern __shared__ double cy[];          // [k][q]
    for (int e = threadIdx.x; e < k * q; e += blockDim.x) {
        const int l = e / q, j = e - l * q;
        cy[e] = C[j + (int64_t)l * q] * ys[j];
    }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m;
         i += (int64_t)gridDim.x * blockDim.x) {
        for (int j0 = 0; j0 < q; j0 += 8) {
            double r[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) r[u] = (j0 + u < q) ? Y[i + (int64_t)(j0 + u) * ldy] - ymeans[j0 + u] : 0.0;
            for (int l = 0; l < k; ++l) {
                const double t = T[i + (int64_t)l * ldt];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (j0 + u < q) r[u] -= t * cy[l * q + j0 + u];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (j0 + u < q) {
                    Yaug[i + (int64_t)(j0 + u) * lda] = r[u];
                    Yaug[i + (int64_t)(q + j0 + u) * lda] = Y[i + (int64_t)(j0 + u) * ldy];
                }
        }
    }
}

// device part: dT (m x ka, ld ldt) must hold the scores; writes the packed Gram of [T | E Y] to
// d_packed (jcb200_packed_len(ka, 2q) doubles).  ka = max(k_hi, 1).
int launch_gridscore_gram(Ctx* c, const double* dY, int64_t ldy, const double* dT, int64_t ldt,
                          const double* dC, const double* dys, const double* dymeans, int64_t m, int q,
                          int k_hi, int ka, double* dYaug, int64_t lda, double* d_pivot0,
                          double* d_packed) {
    const size_t smem = (size_t)std::max(1, k_hi * q) * 8;
    if (smem > 96 * 1024) {
        set_error("gridscore: k_hi * q = %d too large", k_hi * q);
        return JCB200_EINVAL;
    }
    if (smem > 48 * 1024)
        JCB_CUDA(cudaFuncSetAttribute(resid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((m + 255) / 256, 4 * 148);
    resid_kernel<<<grid, 256, smem, c->stream>>>(dY, ldy, dT, ldt, dC, dys, dymeans, m, q, k_hi, dYaug, lda);
    JCB_LAUNCH_CHECK();
    JCB_CUDA(cudaMemsetAsync(d_pivot0, 0, (size_t)(ka + 2 * q + 1) * 8, c->stream));   // pivot 0, centring off
    return launch_gram(c, dT, ldt, dYaug, lda, nullptr, m, ka, 2 * q, d_pivot0, d_packed, 0);
}

// host part: residual sums for k = k_lo..k_hi from the packed Gram (pivot 0: raw second moments)
void gridscore_from_packed(const double* pk, int ka, int q, int k_lo, int k_hi, const double* C,
                           const double* ys, double* ssr, double* sumres, double* ysum, double* ysumsq) {
    const int64_t P = ka, Q = 2 * q;
    const double* gxx = pk;
    const double* gxy = pk + P * P;
    const double* gyy = gxy + P * Q;
    const double* sx = gyy + Q;
    const double* sy = sx + P;
    auto tt = [&](int a, int b) { return a <= b ? gxx[a + (int64_t)b * P] : gxx[b + (int64_t)a * P]; };
    for (int j = 0; j < q; ++j) {
        ysum[j] = sy[q + j];
        ysumsq[j] = gyy[q + j];
        double s2 = gyy[j], s1 = sy[j];                 // k = k_hi: E itself
        std::vector<double> cy(std::max(k_hi, 1));
        for (int l = 0; l < k_hi; ++l) cy[l] = C[j + (int64_t)l * q] * ys[j];
        for (int k = k_hi; k >= k_lo; --k) {
            if (k < k_hi) {                              // add LV l = k to the residual
                const int l = k;
                double cross = 0.0;
                for (int l2 = l + 1; l2 < k_hi; ++l2) cross += cy[l2] * tt(l, l2);
                s2 += 2.0 * cy[l] * gxy[l + (int64_t)j * P] + cy[l] * cy[l] * tt(l, l) + 2.0 * cy[l] * cross;
                s1 += cy[l] * sx[l];
            }
            ssr[(k - k_lo) + (int64_t)j * (k_hi - k_lo + 1)] = s2;
            sumres[(k - k_lo) + (int64_t)j * (k_hi - k_lo + 1)] = s1;
        }
    }
}

}  // namespace jcb

// =============================================================================================
// gridcvlv by Gram down-dating — /root/reference/src/gridcv.jl:187-228 for
// fun = plskern: for every segment s of a repetition the reference fits on rmrow(X, s) and scores on
// X[s, :] — K fits and K row-copies of X.  Every quantity a fit needs is a sum over rows, so
//   Gram(training rows of segment j) = Gram(all rows) - Gram(rows of segment j):
// the rows are permuted once so that each segment is a contiguous slab, K1 runs once per slab (one pass
// over X in total, all slabs about the same pivot), and per segment only K3/K4 (solve) and one scoring
// pass over its own slab remain.
namespace jcb {

// dst row r of every column = src row src_row[r] (gap rows: -1 -> 0)
__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                   int64_t ldd, const int64_t* __restrict__ src_row, int64_t nrows) {
    const int64_t j = blockIdx.y;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src_row[r];
        dst[r + j * ldd] = s >= 0 ? src[s + j * lds] : 0.0;
    }
}

// out = a - b (packed buffers), and acc += b
__global__ void packed_sub_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                  double* __restrict__ out, int64_t len) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] - b[i];
}
__global__ void packed_add_kernel(double* __restrict__ acc, const double* __restrict__ b, int64_t len,
                                  int first) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x)
        acc[i] = (first ? 0.0 : acc[i]) + b[i];
}

int launch_gather_rows(Ctx* c, const double* src, int64_t lds, double* dst, int64_t ldd,
                       const int64_t* d_src_row, int64_t nrows, int64_t ncols) {
    dim3 grid((unsigned)std::min<int64_t>((nrows + 255) / 256, 256), (unsigned)ncols);
    gather_rows_kernel<<<grid, 256, 0, c->stream>>>(src, lds, dst, ldd, d_src_row, nrows);
    JCB_LAUNCH_CHECK();
    return 0;
}
int launch_packed_sub(Ctx* c, const double* a, const double* b, double* out, int64_t len) {
    packed_sub_kernel<<<(int)std::min<int64_t>((len + 255) / 256, 1184), 256, 0, c->stream>>>(a, b, out, len);
    JCB_LAUNCH_CHECK();
    return 0;
}
int launch_packed_add(Ctx* c, double* acc, const double* b, int64_t len, int first) {
    packed_add_kernel<<<(int)std::min<int64_t>((len + 255) / 256, 1184), 256, 0, c->stream>>>(acc, b, len, first);
    JCB_LAUNCH_CHECK();
    return 0;
}

}  // namespace jcb
