// Dependent-issue latencies on B200 (sm_100a): DFMA, DADD, DMMA (same accumulator), LDS.64, SHFL.
// One warp, one CTA, clock64 around a chain of N dependent ops.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(long long* out, double* sink, double x, int n) {
    __shared__ double sm[64];
    sm[threadIdx.x] = threadIdx.x * 1.0;
    __syncthreads();
    double a = x, c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, 1.0000001, 1e-9);
    long long t1 = clock64();
    double b = x;
    for (int i = 0; i < n; ++i) b = b + 1e-9;
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) dmma884(c0, c1, x, 1e-3);
    long long t3 = clock64();
    int idx = threadIdx.x;
    for (int i = 0; i < n; ++i) idx = (int)sm[idx & 31] & 31;
    long long t4 = clock64();
    double s = x;
    for (int i = 0; i < n; ++i) s = __shfl_xor_sync(0xffffffffu, s, 1) + 0.0;
    long long t5 = clock64();
    // DADD feeding a DMMA feeding ... (alternating dependent chain)
    double e0 = 0, e1 = 0, f = x;
    for (int i = 0; i < n; ++i) { f = f + e0; dmma884(e0, e1, f, 1e-3); }
    long long t6 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5;
    }
    sink[threadIdx.x] = a + b + c0 + c1 + idx + s + e0 + e1 + f;
}
int main() {
    long long* out; double* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 32 * 8);
    const int n = 4096;
    k<<<1, 32>>>(out, sink, 1.0, n); k<<<1, 32>>>(out, sink, 1.0, n);
    long long h[6]; cudaMemcpy(h, out, 48, cudaMemcpyDeviceToHost);
    const char* nm[6] = {"DFMA dependent", "DADD dependent", "DMMA same accumulator", "LDS.64 -> cvt -> LDS chain",
                         "SHFL.64 + DADD chain", "DADD -> DMMA -> DADD chain (per pair)"};
    printf("{");
    for (int i = 0; i < 6; ++i) printf("\"%s\": %.1f%s", nm[i], (double)h[i] / n, i < 5 ? ", " : "}\n");
    return 0;
}
