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
// (all loops unrolled, no integer division, no block barrier, no shuffles).  A <- A*A (A symmetric:
// column j = row j, both operands contiguous rows) converges quadratically to a multiple of v v'.
//  * Only the upper triangle is computed: the NT = Q(Q+1)/2 entries are dealt round-robin to the lanes
//    (lane l owns entries l, l+32, ...), each entry is one row-by-row dot (two FMA chains, 16-byte loads
//    when Q is even and the buffers are aligned) and is stored to both (i,j) and (j,i).
//  * After every squaring every lane reads the Q diagonal entries itself and adds them in the same order,
//    so the trace tau is known to all lanes without a reduction.  The next squaring is scaled by the power
//    of two 2^(-2 e), e = exponent(tau): exact, no division, trace stays O(1).
//  * Convergence: tr(A^2)/tr(A)^2 = 1 - 2 rho + O(rho^2), rho = (lambda_2/lambda_1)^(2^k) the remaining
//    contamination.  When a squaring shows rho < 5e-5, its result carries rho^2 and ONE more squaring
//    leaves rho^4 < 1e-17: the trace of a matrix is taken inside the squaring that consumes it, and the
//    test on it ends the loop right after that squaring.
// Result: column `best` (largest diagonal) of the converged matrix in v_out[0..Q); M = 0 gives e_1.
template <int Q>
__device__ __forceinline__ void eig_dominant_warp(const double* __restrict__ M_s, double* __restrict__ bufA,
                                                  double* __restrict__ bufB, double* __restrict__ v_out,
                                                  const int lane) {
    constexpr int NT = Q * (Q + 1) / 2;
    constexpr int NE = (NT + 31) / 32;
    int ei[NE], ej[NE];
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        int rem = lane + 32 * u, i = 0;
        if (rem >= NT) rem = NT - 1;                  // idle slot: recompute the last entry (same value)
        while (rem >= Q - i) { rem -= Q - i; ++i; }
        ei[u] = i;
        ej[u] = i + rem;
    }
    double tr = 0.0;
#pragma unroll
    for (int d = 0; d < Q; ++d) tr += M_s[d * Q + d];
    if (!(tr > 0.0)) {
        if (lane < Q) v_out[lane] = (lane == 0) ? 1.0 : 0.0;
        __syncwarp();
        return;
    }
    // start from M scaled by the exact power of two that brings its trace into [1, 2) (no division)
    const double itr = __hiloint2double((2046 - ((__double2hiint(tr) >> 20) & 0x7ff)) << 20, 0);
    for (int e = lane; e < Q * Q; e += 32) bufA[e] = M_s[e] * itr;
    __syncwarp();
    double* cur = bufA;
    double* nxt = bufB;
    const bool vec = (Q % 2 == 0) &&
                     (((unsigned long long)(uintptr_t)bufA | (unsigned long long)(uintptr_t)bufB) & 15ull) == 0ull;
    // Warps issue in order: the trace of the CURRENT matrix is taken inside the squaring that consumes it
    // (diagonal loads issued with the row loads, a 4-level add tree that runs beside the FMA chains, the
    // scale applied as the last multiply), so neither the trace nor the test adds to the critical path.
    double tau_prev = 1.0, scl_prev = 1.0;
    for (int iter = 0; iter < 64; ++iter) {
        double dg[Q];
#pragma unroll
        for (int d = 0; d < Q; ++d) dg[d] = cur[d * Q + d];
        double acc[NE][2];
#pragma unroll
        for (int u = 0; u < NE; ++u) {
            const double* ri = cur + ei[u] * Q;
            const double* rj = cur + ej[u] * Q;
            double s0 = 0.0, s1 = 0.0;
            if (vec) {
                const double2* a2 = reinterpret_cast<const double2*>(ri);
                const double2* b2 = reinterpret_cast<const double2*>(rj);
#pragma unroll
                for (int k = 0; k < Q / 2; ++k) {
                    const double2 x = a2[k], y = b2[k];
                    s0 += x.x * y.x;
                    s1 += x.y * y.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k + 1 < Q; k += 2) {
                    s0 += ri[k] * rj[k];
                    s1 += ri[k + 1] * rj[k + 1];
                }
                if (Q & 1) s0 += ri[Q - 1] * rj[Q - 1];
            }
            acc[u][0] = s0;
            acc[u][1] = s1;
        }
        // tau = trace(cur): pairwise tree, same order in every lane
#pragma unroll
        for (int w2 = 1; w2 < Q; w2 <<= 1) {
#pragma unroll
            for (int d = 0; d + w2 < Q; d += 2 * w2) dg[d] += dg[d + w2];
        }
        const double tau = dg[0];
        // scl = 2^(-2 e) with tau = f * 2^e, f in [1, 2)
        const int ex = ((__double2hiint(tau) >> 20) & 0x7ff) - 1023;
        const double scl = __hiloint2double((1023 - 2 * ex) << 20, 0);
#pragma unroll
        for (int u = 0; u < NE; ++u) {
            const double val = (acc[u][0] + acc[u][1]) * scl;
            nxt[ei[u] * Q + ej[u]] = val;
            nxt[ej[u] * Q + ei[u]] = val;
        }
        __syncwarp();
        double* t = cur;
        cur = nxt;
        nxt = t;
        // tau / (scl_prev tau_prev^2) = tr(A_{k-1}^2)/tr(A_{k-1})^2 = 1 - 2 rho_{k-1}: rho_{k-1} < 5e-5 means the
        // matrix just squared carried rho^2 and the one just stored carries rho^4 < 1e-17
        if (iter > 0 && tau >= (1.0 - 1e-4) * scl_prev * tau_prev * tau_prev) break;
        tau_prev = tau;
        scl_prev = scl;
    }
    // column with the largest diagonal entry: all loads first, then a pairwise tournament (ties keep the lower index)
    double bd[Q];
    int bi[Q];
#pragma unroll
    for (int d = 0; d < Q; ++d) {
        bd[d] = cur[d * Q + d];
        bi[d] = d;
    }
#pragma unroll
    for (int w2 = 1; w2 < Q; w2 <<= 1) {
#pragma unroll
        for (int d = 0; d + w2 < Q; d += 2 * w2)
            if (bd[d + w2] > bd[d]) {
                bd[d] = bd[d + w2];
                bi[d] = bi[d + w2];
            }
    }
    const int best = bi[0];
    if (lane < Q) v_out[lane] = cur[lane * Q + best];
    __syncwarp();
}

// dispatch on the runtime q (2..16); executed by one full warp
__device__ __forceinline__ void eig_dominant_warp_q(const int q, const double* __restrict__ M_s,
                                                    double* __restrict__ bufA, double* __restrict__ bufB,
                                                    double* __restrict__ v_out, const int lane) {
    switch (q) {
#define JCB_EIG(Q) case Q: eig_dominant_warp<Q>(M_s, bufA, bufB, v_out, lane); break;
#ifdef JCB_EIG_ONLY_Q
        JCB_EIG(JCB_EIG_ONLY_Q)          // code-size experiment: a single instantiation
#else
        JCB_EIG(2) JCB_EIG(3) JCB_EIG(4) JCB_EIG(5) JCB_EIG(6) JCB_EIG(7) JCB_EIG(8)
        JCB_EIG(9) JCB_EIG(10) JCB_EIG(11) JCB_EIG(12) JCB_EIG(13) JCB_EIG(14)
        JCB_EIG(15) JCB_EIG(16)
#endif
#undef JCB_EIG
        default: break;
    }
}

}  // namespace jcb
