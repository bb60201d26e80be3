// Small device routines shared by the LV-loop kernel (k4_solve.cu) and the batched local fits (k9_locw.cu).
#pragma once
#include <stdint.h>

namespace jcb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Dominant eigenvector of the symmetric PSD Q x Q matrix M by repeated squaring, ONE warp, compile-time Q
// (all loops unrolled, no integer division, no block barrier).  L = 32/Q lanes share row i; a lane owns
// the entries (i, jl + u*L).  A <- A*A (A symmetric: column j = row j, both operands contiguous rows)
// converges quadratically to a multiple of v v'.  The matrix is renormalised by its trace every third
// squaring, where convergence is checked: if the normalised matrix moved by < 1e-4 over the last three
// squarings, the contamination three squarings ago was <~ 1e-4 and is now its 8th power.
// Result: column `best` (largest diagonal) of the converged matrix in v_out[0..Q).
template <int Q>
__device__ __forceinline__ void eig_dominant_warp(const double* __restrict__ M_s, double* __restrict__ bufA,
                                                  double* __restrict__ bufB, double* __restrict__ v_out,
                                                  const int lane) {
    constexpr int L = 32 / Q;
    constexpr int NJ = (Q + L - 1) / L;
    const int i = lane / L, jl = lane - i * L;
    const bool act = i < Q;
    double tr = 0.0;
#pragma unroll
    for (int d = 0; d < Q; ++d) tr += M_s[d * Q + d];
    const double itr = 1.0 / tr;
    double prev[NJ];
    double* cur = bufA;
    double* nxt = bufB;
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
        const int j = jl + u * L;
        prev[u] = 0.0;
        if (act && j < Q) {
            prev[u] = M_s[i * Q + j] * itr;
            cur[i * Q + j] = prev[u];
        }
    }
    __syncwarp();
    for (int iter = 0; iter < 90; ++iter) {
        double acc[NJ];
#pragma unroll
        for (int u = 0; u < NJ; ++u) acc[u] = 0.0;
        if (act) {
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                const double aik = cur[i * Q + k];
#pragma unroll
                for (int u = 0; u < NJ; ++u) {
                    const int j = jl + u * L;
                    if (j < Q) acc[u] += aik * cur[j * Q + k];
                }
            }
        }
        if (iter % 3 != 2) {
            if (act) {
#pragma unroll
                for (int u = 0; u < NJ; ++u) {
                    const int j = jl + u * L;
                    if (j < Q) nxt[i * Q + j] = acc[u];
                }
            }
            __syncwarp();
        } else {
            // trace of the new matrix = sum of its diagonal entries: collect them through a shuffle sum
            double dg = 0.0;
#pragma unroll
            for (int u = 0; u < NJ; ++u) {
                const int j = jl + u * L;
                if (act && j == i) dg = acc[u];
            }
            const double inv = 1.0 / warp_sum(dg);
            double chg = 0.0;
#pragma unroll
            for (int u = 0; u < NJ; ++u) {
                const int j = jl + u * L;
                if (act && j < Q) {
                    const double nv = acc[u] * inv;
                    chg = fmax(chg, fabs(nv - prev[u]));
                    prev[u] = nv;
                    nxt[i * Q + j] = nv;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) chg = fmax(chg, __shfl_xor_sync(0xffffffffu, chg, o));
            __syncwarp();
            if (chg < 1e-4) {
                cur = nxt;
                break;
            }
        }
        double* t = cur;
        cur = nxt;
        nxt = t;
    }
    __syncwarp();
    int best = 0;
#pragma unroll
    for (int d = 1; d < Q; ++d)
        if (cur[d * Q + d] > cur[best * Q + best]) best = d;
    if (lane < Q) v_out[lane] = cur[lane * Q + best];
}

// dispatch on the runtime q (2..16); executed by one full warp
__device__ __forceinline__ void eig_dominant_warp_q(const int q, const double* __restrict__ M_s,
                                                    double* __restrict__ bufA, double* __restrict__ bufB,
                                                    double* __restrict__ v_out, const int lane) {
    switch (q) {
#define JCB_EIG(Q) case Q: eig_dominant_warp<Q>(M_s, bufA, bufB, v_out, lane); break;
        JCB_EIG(2) JCB_EIG(3) JCB_EIG(4) JCB_EIG(5) JCB_EIG(6) JCB_EIG(7) JCB_EIG(8)
        JCB_EIG(9) JCB_EIG(10) JCB_EIG(11) JCB_EIG(12) JCB_EIG(13) JCB_EIG(14)
        JCB_EIG(15) JCB_EIG(16)
#undef JCB_EIG
        default: break;
    }
}

}  // namespace jcb
