// FP64 peak microbenchmarks for B200 (sm_100a): register-resident DMMA (mma.sync m8n8k4 f64),
// DFMA, and a mixed DMMA+DFMA loop (tells whether the two share a pipe).  The roofline denominator
// for the Gram kernel (K1) is measured with this, since MEASURED_PEAKS.json has no FP64 figure.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/fp64_peak bench/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NM DMMAs + NF DFMAs per iteration, independent chains.
template <int NM, int NF>
__global__ void __launch_bounds__(1024) k_mixed(double* out, int iters, double a0, double b0) {
    double acc[NM][2];
    double f[NF];
#pragma unroll
    for (int i = 0; i < NM; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            dmma884(acc[i][0], acc[i][1], a, b);
            if (i < NF) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f,\n", prop.name, sms, clk_khz / 1000.0);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    printf(" \"results\": [\n");
    int first = 1;
    auto report = [&](const char* name, int warps, int ctas_per_sm, double flops, float ms) {
        printf("%s  {\"kernel\": \"%s\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
               first ? "" : ",\n", name, warps, ctas_per_sm, ms, flops / ms * 1e-9);
        first = 0;
    };
    int warp_cfgs[] = {4, 8, 16, 32};
    for (int w : warp_cfgs) {
        int threads = w * 32;
        int grid = sms;
        {   // DMMA, 16 independent accumulators per warp
            float ms = time_ms([&] { k_dmma<16><<<grid, threads>>>(out, iters, 1.0, 1e-3); }, 5);
            double flops = (double)grid * w * iters * 16 * 512.0;
            report("dmma884_acc16", w, 1, flops, ms);
        }
        {
            float ms = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, iters, 1.0, 1e-3); }, 5);
            double flops = (double)grid * w * iters * 8 * 512.0;
            report("dmma884_acc8", w, 1, flops, ms);
        }
        {
            float ms = time_ms([&] { k_dfma<16><<<grid, threads>>>(out, iters, 1.0000001, 1e-3); }, 5);
            double flops = (double)grid * threads * iters * 16 * 2.0;
            report("dfma_acc16", w, 1, flops, ms);
        }
        {   // mixed: 16 DMMA + 4 DFMA per iteration (DFMA is 1/4 of DMMA instruction count, like the K1 loop)
            float ms = time_ms([&] { k_mixed<16, 4><<<grid, threads>>>(out, iters, 1.0000001, 1e-3); }, 5);
            double flops = (double)grid * w * iters * (16 * 512.0 + 4 * 64.0);
            report("mixed_16dmma_4dfma", w, 1, flops, ms);
        }
        {   // mixed: 16 DMMA + 16 DFMA per iteration
            float ms = time_ms([&] { k_mixed<16, 16><<<grid, threads>>>(out, iters, 1.0000001, 1e-3); }, 5);
            double flops = (double)grid * w * iters * (16 * 512.0 + 16 * 64.0);
            report("mixed_16dmma_16dfma", w, 1, flops, ms);
        }
    }
    // sustained: DMMA for ~3 s to see the power-capped clock
    {
        int w = 16, threads = w * 32, grid = sms;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 0;
        for (; launches < 60; ++launches) k_dmma<16><<<grid, threads>>>(out, iters * 4, 1.0, 1e-3);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)launches * grid * w * (iters * 4.0) * 16 * 512.0;
        report("dmma884_acc16_sustained", w, 1, flops, ms);
    }
    printf("\n ]}\n");
    CK(cudaFree(out));
    return 0;
}
