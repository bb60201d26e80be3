// K10 — xfit / xresid (next row): X_fit = (T P') * diag(xscales) + 1 xmeans' and E = X - X_fit,
// /root/reference/src/xfit.jl:33-56 and :88-99.  T = ((X - xmeans) / xscales) R comes from K5; this kernel is
// the second skinny product, fused with the return to the original scale (scale! by 1 ./ xscales then
// center! by -xmeans, xfit.jl:50-53) and, for xresid, with the subtraction from X.  The result may overwrite the
// X slab in HBM (as xfit! overwrites its argument); the host path writes chunk-local result buffers instead, so
// that the copy back to the host never reads the allocation the next chunk is being copied into.
//
// HBM-bound on the m x p write (plus the read of X for xresid): 8 m p (16 m p) bytes against 2 m p nlv flops.
#include "jcb_internal.cuh"

namespace jcb {

constexpr int XF_MT = 128;    // rows per CTA tile
constexpr int XF_NT = 64;     // columns per CTA tile
constexpr int XF_KC = 32;     // LVs per staged chunk

// thread (tx = tid & 15, ty = tid >> 4): row pairs 32 i + 2 tx (+1), i = 0..3, columns 4 ty .. 4 ty + 3 of the
// tile — a half-warp's 16-byte stores to one column cover 256 contiguous bytes
__global__ void __launch_bounds__(256) xfit_kernel(const double* X, int64_t ldx, double* Out, int64_t ldo,
                                                   int64_t m, int p,
                                                   const double* __restrict__ T, int64_t ldt,
                                                   const double* __restrict__ P, int64_t ldp, int nlv,
                                                   const double* __restrict__ xm, const double* __restrict__ xs,
                                                   int resid) {
    __shared__ __align__(16) double Ts[XF_KC][XF_MT];
    __shared__ __align__(16) double Ps[XF_KC][XF_NT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t r0 = (int64_t)blockIdx.x * XF_MT;
    const int c0 = blockIdx.y * XF_NT;
    double acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.0;
    for (int k0 = 0; k0 < nlv; k0 += XF_KC) {
        const int kc = min(XF_KC, nlv - k0);
        __syncthreads();
        for (int e = tid; e < kc * XF_MT; e += 256) {
            const int k = e / XF_MT, r = e - k * XF_MT;
            Ts[k][r] = (r0 + r < m) ? T[r0 + r + (int64_t)(k0 + k) * ldt] : 0.0;
        }
        for (int e = tid; e < kc * XF_NT; e += 256) {
            const int k = e / XF_NT, j = e - k * XF_NT;
            Ps[k][j] = (c0 + j < p) ? P[c0 + j + (int64_t)(k0 + k) * ldp] : 0.0;
        }
        __syncthreads();
        for (int k = 0; k < kc; ++k) {
            double tv[8], pv[4];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&Ts[k][16 * i + 2 * tx]);
                tv[i] = v.x;
                tv[i + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&Ps[k][4 * ty + j]);
                pv[j] = v.x;
                pv[j + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fma(tv[i], pv[j], acc[j][i]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = c0 + 4 * ty + j;
        if (col >= p) continue;
        const double mu = xm[col], sc = xs[col];
        const double* scol = X + (int64_t)col * ldx + r0 + 2 * tx;
        double* dcol = Out + (int64_t)col * ldo + r0 + 2 * tx;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const int64_t row = r0 + 16 * i + 2 * tx;
            if (row >= m) break;
            double* dst = dcol + 16 * i;
            const double* src = scol + 16 * i;
            // scale!(X, 1 ./ xscales); center!(X, -xmeans); xresid: X .- xfit(...)
            const double f0 = acc[j][i] * sc + mu, f1 = acc[j][i + 1] * sc + mu;
            if (row + 1 < m) {
                double2 v = make_double2(f0, f1);
                if (resid) {
                    const double2 x = *reinterpret_cast<const double2*>(src);
                    v.x = x.x - f0;
                    v.y = x.y - f1;
                }
                *reinterpret_cast<double2*>(dst) = v;
            } else {
                dst[0] = resid ? src[0] - f0 : f0;
            }
        }
    }
}

// dOut (ld ldo, even) may be the X slab itself (in place, as xfit! overwrites its argument) or a separate buffer
int launch_xfit(Ctx* c, const double* dX, int64_t ldx, double* dOut, int64_t ldo, int64_t m, int64_t p,
                const double* dT, int64_t ldt, const double* dP, int64_t ldp, int nlv, const double* dxm,
                const double* dxs, int resid) {
    dim3 grid((unsigned)((m + XF_MT - 1) / XF_MT), (unsigned)((p + XF_NT - 1) / XF_NT));
    xfit_kernel<<<grid, 256, 0, c->stream>>>(dX, ldx, dOut, ldo, m, (int)p, dT, ldt, dP, ldp, nlv, dxm, dxs,
                                             resid);
    JCB_LAUNCH_CHECK();
    return 0;
}

}  // namespace jcb
