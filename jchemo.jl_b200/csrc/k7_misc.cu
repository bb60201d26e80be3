// K7 (in-place centring / scaling write-back), normalised weights, K8 (counter-based fill).
#include <algorithm>

#include "jcb_internal.cuh"

namespace jcb {

// center! / cscale! (/root/reference/src/utility.jl:76-81,482-487): X[:, j] = (X[:, j] - mu[j]) / sigma[j],
// subtract then divide per element as the reference does.  HBM-bound: 128-bit loads/stores along rows.
__global__ void center_scale_kernel(double* __restrict__ X, int64_t ldx, int64_t n, int p,
                                    const double* __restrict__ mu, const double* __restrict__ sigma) {
    const int j = blockIdx.y;
    const double m = mu[j], s = sigma ? sigma[j] : 1.0;
    double* col = X + (int64_t)j * ldx;
    const bool vec = (((uintptr_t)col & 15) == 0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t n2 = n >> 1;
        double2* c2 = reinterpret_cast<double2*>(col);
        for (int64_t i = t0; i < n2; i += stride) {
            double2 v = c2[i];
            v.x = (v.x - m) / s;
            v.y = (v.y - m) / s;
            c2[i] = v;
        }
        if ((n & 1) && t0 == 0) col[n - 1] = (col[n - 1] - m) / s;
    } else {
        for (int64_t i = t0; i < n; i += stride) col[i] = (col[i] - m) / s;
    }
}

int launch_center_scale(Ctx* c, double* dX, int64_t ldx, int64_t n, int64_t p, const double* dmu,
                        const double* dsigma) {
    if (n <= 0 || p <= 0) return 0;
    int gx = (int)std::min<int64_t>((n / 2 + 255) / 256, 64);
    if (gx < 1) gx = 1;
    dim3 grid(gx, (unsigned)p);
    center_scale_kernel<<<grid, 256, 0, c->stream>>>(dX, ldx, n, (int)p, dmu, dsigma);
    JCB_LAUNCH_CHECK();
    return 0;
}

// mweight (/root/reference/src/utility.jl:715-723)
__global__ void weights_kernel(const double* __restrict__ w, int64_t n, const double* __restrict__ sumw,
                               double* __restrict__ out) {
    const double S = *sumw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (w ? w[i] : 1.0) / S;
}

int launch_weights(Ctx* c, const double* dw, int64_t n, const double* dsumw, double* dw_out) {
    int grid = (int)std::min<int64_t>((n + 255) / 256, 1184);
    weights_kernel<<<grid, 256, 0, c->stream>>>(dw, n, dsumw, dw_out);
    JCB_LAUNCH_CHECK();
    return 0;
}

// K8: u(s, k) = (mix64(s*0xD1342543DE82EF95 + (k+1)*0x9E3779B97F4A7C15) >> 11) * 2^-53
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

__global__ void fill_uniform_kernel(double* __restrict__ d, int64_t ld, int64_t n_rows, int64_t n_cols,
                                    uint64_t seed, int64_t row0, int64_t n_global) {
    const int64_t j = blockIdx.y;
    const uint64_t base = seed * 0xD1342543DE82EF95ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = (uint64_t)(row0 + i) + (uint64_t)j * (uint64_t)n_global;
        const uint64_t z = mix64(base + (k + 1) * 0x9E3779B97F4A7C15ull);
        d[i + j * ld] = (double)(z >> 11) * 0x1.0p-53;
    }
}

int launch_fill_uniform(Ctx* c, double* d, int64_t ld, int64_t n_rows, int64_t n_cols, uint64_t seed,
                        int64_t row0, int64_t n_global) {
    if (n_rows <= 0 || n_cols <= 0) return 0;
    int gx = (int)std::min<int64_t>((n_rows + 255) / 256, 128);
    dim3 grid(gx, (unsigned)n_cols);
    fill_uniform_kernel<<<grid, 256, 0, c->stream>>>(d, ld, n_rows, n_cols, seed, row0, n_global);
    JCB_LAUNCH_CHECK();
    return 0;
}

// summary(::Plsr, X) (/root/reference/src/plskern.jl:246-260): sstot = sum_ij w_i ((x_ij - mu_j)/sigma_j)^2,
// the weighted total sum of squares of the centred/scaled X.  One HBM-bound pass; per-block partial sums
// are written to `partial` and added in a fixed order by the last kernel (deterministic).
__global__ void __launch_bounds__(256)
sstot_partial_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, int p,
                     const double* __restrict__ mu, const double* __restrict__ sigma,
                     const double* __restrict__ w, double* __restrict__ partial) {
    const int j = blockIdx.y;
    const double m = mu[j], is = 1.0 / sigma[j];
    const double* col = X + (int64_t)j * ldx;
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double d = (col[i] - m) * is;
        s += w[i] * d * d;
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        partial[(int64_t)j * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void sstot_final_kernel(const double* __restrict__ partial, int64_t count, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) s += partial[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        *out = t;
    }
}

int launch_sstot(Ctx* c, const double* dX, int64_t ldx, int64_t n, int64_t p, const double* dmu,
                 const double* dsigma, const double* dw, double* dpartial, int gx, double* dout) {
    dim3 grid(gx, (unsigned)p);
    sstot_partial_kernel<<<grid, 256, 0, c->stream>>>(dX, ldx, n, (int)p, dmu, dsigma, dw, dpartial);
    JCB_LAUNCH_CHECK();
    sstot_final_kernel<<<1, 1024, 0, c->stream>>>(dpartial, (int64_t)gx * p, dout);
    JCB_LAUNCH_CHECK();
    return 0;
}

}  // namespace jcb
