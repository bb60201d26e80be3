// What does a DFMA / DADD cost when it is issued into a stream of DMMAs on B200 (sm_100a)?
// Register-resident loops, no memory traffic: NM DMMA.8x8x4 per iteration over 8 accumulators plus NF scalar
// FP64 instructions per iteration, placed in different ways.  Reports FP64-pipe cycles per iteration per SMSP
// (1965 MHz assumed for the conversion is NOT used: cycles come from clock64 of one warp) and the cost per
// scalar instruction relative to the DMMA-only loop.
//   mode 0: DFMAs (independent chains) spread evenly between the DMMAs                  (K5 leftover columns)
//   mode 1: DFMAs batched at the end of the iteration
//   mode 2: DADD feeds the B operand of the next NM/NF DMMAs (dependent, issued right before) (K1 centring)
//   mode 3: all DADDs of the iteration first, then the DMMAs that use them
//   mode 4: DADDs produce the operands of the NEXT iteration, spread evenly (software pipelined)
//   mode 5: DFMAs in ONE dependent chain, spread evenly
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/fp64_mix bench/fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double vfma(double x, double a, double b) {
    double r;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(x), "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double vsub(double x, double c) {
    double r;
    asm volatile("sub.f64 %0, %1, %2;" : "=d"(r) : "d"(x), "d"(c));
    return r;
}
__device__ __forceinline__ void opaque(double& x) { asm volatile("" : "+d"(x)); }

template <int NM, int NF, int MODE>
__global__ void __launch_bounds__(512, 1) k_mix(double* out, long long* cyc, int iters, double a0, double b0, double c0) {
    constexpr int NA = 8;
    constexpr int NFF = NF > 0 ? NF : 1;
    constexpr int PER = NF > 0 ? NM / NFF : NM;
    double acc[NA][2];
    double f[NFF], bop[NFF], bnext[NFF];
#pragma unroll
    for (int i = 0; i < NA; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NFF; ++i) { f[i] = i + threadIdx.x; bop[i] = b0 + i; bnext[i] = b0 + i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < NM; ++i) {
                dmma884(acc[i % NA][0], acc[i % NA][1], a, b);
                if (NF > 0 && (i % PER) == PER - 1 && i / PER < NF) {
                    if (MODE == 0) f[i / PER] = vfma(f[i / PER], a, acc[i % NA][1]);   // after DMMA i
                    else f[0] = vfma(f[0], a, acc[i % NA][1]);
                }
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma884(acc[i % NA][0], acc[i % NA][1], a, b);
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = vfma(f[i], a, b);
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < NFF; ++j) {
                bop[j] = vsub(f[j], c0);
#pragma unroll
                for (int i = 0; i < PER; ++i) dmma884(acc[(j * PER + i) % NA][0], acc[(j * PER + i) % NA][1], a, bop[j]);
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < NFF; ++j) bop[j] = vsub(f[j], c0);
#pragma unroll
            for (int j = 0; j < NFF; ++j)
#pragma unroll
                for (int i = 0; i < PER; ++i) dmma884(acc[(j * PER + i) % NA][0], acc[(j * PER + i) % NA][1], a, bop[j]);
        } else if (MODE == 4) {
#pragma unroll
            for (int j = 0; j < NFF; ++j) {
#pragma unroll
                for (int i = 0; i < PER; ++i) dmma884(acc[(j * PER + i) % NA][0], acc[(j * PER + i) % NA][1], a, bop[j]);
                bnext[j] = vsub(f[j], c0);
            }
#pragma unroll
            for (int j = 0; j < NFF; ++j) bop[j] = bnext[j];
        }
        if (MODE >= 2 && MODE <= 4) {
            // new raw operands every iteration (stands in for the LDS): integer-side update, no FP64 instruction
#pragma unroll
            for (int j = 0; j < NFF; ++j) {
                opaque(f[j]);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < NFF; ++i) s += f[i] + bop[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

static double* g_out;
static long long* g_cyc;
static int g_first = 1;

template <int NM, int NF, int MODE>
static void run(int warps, int sms, double base_cyc_per_iter, double* cyc_out) {
    const int iters = 20000;
    k_mix<NM, NF, MODE><<<sms, warps * 32>>>(g_out, g_cyc, iters, 1.0000001, 1e-3, 0.5);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    long long cyc = 0;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        k_mix<NM, NF, MODE><<<sms, warps * 32>>>(g_out, g_cyc, iters, 1.0000001, 1e-3, 0.5);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) { best = ms; CK(cudaMemcpy(&cyc, g_cyc, 8, cudaMemcpyDeviceToHost)); }
    }
    // cycles of the SMSP's FP64 pipe per iteration of ONE warp = elapsed cycles / iters / warps per SMSP
    const double wps = warps / 4.0 < 1.0 ? 1.0 : warps / 4.0;
    const double per_iter = (double)cyc / iters / wps;
    const double ideal = NM * 16.0 + NF * 2.0;
    double per_scalar = -1.0;
    if (NF > 0 && base_cyc_per_iter > 0) per_scalar = (per_iter - base_cyc_per_iter) / (NF > 0 ? NF : 1);
    printf("%s  {\"nm\": %d, \"nf\": %d, \"mode\": %d, \"warps\": %d, \"ms\": %.4f, \"cyc_per_iter_per_warp_slot\": %.2f, "
           "\"ideal\": %.1f, \"pipe_eff\": %.4f, \"cyc_per_scalar\": %.2f}",
           g_first ? "" : ",\n", NM, NF, MODE, warps, best, per_iter, ideal, ideal / per_iter, per_scalar);
    g_first = 0;
    if (cyc_out) *cyc_out = per_iter;
}

template <int NM>
static void family(int warps, int sms) {
    double base = 0;
    run<NM, 0, 0>(warps, sms, 0, &base);
    run<NM, 1, 0>(warps, sms, base, nullptr);
    run<NM, 2, 0>(warps, sms, base, nullptr);
    run<NM, 4, 0>(warps, sms, base, nullptr);
    run<NM, 2, 1>(warps, sms, base, nullptr);
    run<NM, 4, 1>(warps, sms, base, nullptr);
    run<NM, 4, 5>(warps, sms, base, nullptr);
    run<NM, 4, 2>(warps, sms, base, nullptr);
    run<NM, 4, 3>(warps, sms, base, nullptr);
    run<NM, 4, 4>(warps, sms, base, nullptr);
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    CK(cudaMalloc(&g_out, sizeof(double) * sms * 512));
    CK(cudaMalloc(&g_cyc, 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", prop.name, sms);
    int warp_cfgs[] = {4, 8, 16};
    for (int w : warp_cfgs) {
        family<12>(w, sms);   // K5/K6 at nlv = 50: 12 DMMAs + 4 DFMAs per k4-step
        family<16>(w, sms);   // K1: 16 DMMAs + 4 DADDs per k4-step
    }
    // K5 at nlv = 25: 6 DMMAs + 2 DFMAs
    for (int w : warp_cfgs) {
        double base = 0;
        run<6, 0, 0>(w, sms, 0, &base);
        run<6, 1, 0>(w, sms, base, nullptr);
        run<6, 2, 0>(w, sms, base, nullptr);
        run<6, 2, 1>(w, sms, base, nullptr);
    }
    printf("\n]}\n");
    return 0;
}
