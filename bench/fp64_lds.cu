// Where do the ~30 idle FP64-pipe cycles per k4-step of the streaming kernels (K5, K1) come from?
// Shared-memory-resident model of the K5 inner loop: per k4-step one LDS.128 (A fragment: two row blocks) and NB
// LDS.64 (B fragments), then 2 x NB DMMA.8x8x4.  No global traffic, no barriers, so what is left is the
// LDS -> DMMA structure itself.  Reports FP64-pipe cycles per k4-step per warp slot against the ideal 32 x NB.
//   mode 0: as K5 (loads and DMMAs of the same k4-step in program order, the compiler schedules)
//   mode 1: explicit register double buffering (fragments of step k+1 are loaded before the DMMAs of step k)
//   mode 2: no LDS at all, operands rotate through registers
//   mode 3: A from shared memory, B constant        mode 4: B from shared memory, A constant
//   mode 5: as 0, but the 8 k4-steps of a chunk are NOT unrolled (#pragma unroll 1)
//   mode 6: as 1 with two steps of lookahead
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/fp64_lds bench/fp64_lds.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 lds128(const double* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ double lds64(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

constexpr int KC = 32, MPITCH = 36;

template <int NB, int MODE, int NCW>
__global__ void __launch_bounds__(NCW * 32, 1) k_lds(double* out, int chunks, int zero) {
    constexpr int MT = 16 * NCW, PITCH = MT + 4;
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                       // [KC][PITCH]
    double* ms = sm + KC * PITCH;          // [NB*8][MPITCH]
    for (int i = threadIdx.x; i < KC * PITCH + NB * 8 * MPITCH; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, kk = lane & 3, m0 = warp * 16;
    double acc[2][NB][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[h][nb][0] = acc[h][nb][1] = 0.0;
    const double* pa = xs + kk * PITCH + m0 + 2 * g;
    const double* pb = ms + g * MPITCH + kk;
    double2 ca = *reinterpret_cast<const double2*>(pa);
    double cb[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) cb[nb] = pb[nb * 8 * MPITCH];

    const double* pac = pa;
    const double* pbc = pb;
    // plain loads (the compiler schedules them as in K5); the address depends on the chunk through a runtime zero,
    // so nothing can be hoisted out of the chunk loop
    auto loadA = [&](int k4) { return *reinterpret_cast<const double2*>(pac + k4 * 4 * PITCH); };
    auto loadB = [&](int k4, int nb) { return pbc[nb * 8 * MPITCH + k4 * 4]; };

#pragma unroll 1
    for (int ch = 0; ch < chunks; ++ch) {
        pac = pa + 2 * (ch & zero);
        pbc = pb + (ch & zero);
        if (MODE == 0 || MODE == 3 || MODE == 4) {
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                double2 a = ca;
                if (MODE != 4) a = loadA(k4);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    double b = cb[nb];
                    if (MODE != 3) b = loadB(k4, nb);
                    dmma884(acc[0][nb][0], acc[0][nb][1], a.x, b);
                    dmma884(acc[1][nb][0], acc[1][nb][1], a.y, b);
                }
            }
        } else if (MODE == 5) {
#pragma unroll 1
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                const double2 a = loadA(k4);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const double b = loadB(k4, nb);
                    dmma884(acc[0][nb][0], acc[0][nb][1], a.x, b);
                    dmma884(acc[1][nb][0], acc[1][nb][1], a.y, b);
                }
            }
        } else if (MODE == 1 || MODE == 6) {
            constexpr int LA = MODE == 1 ? 1 : 2;
            double2 a[LA + 1];
            double b[LA + 1][NB];
#pragma unroll
            for (int l = 0; l < LA; ++l) {
                a[l] = loadA(l);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) b[l][nb] = loadB(l, nb);
            }
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                const int cur = k4 % (LA + 1), nxt = (k4 + LA) % (LA + 1);
                const int kn = (k4 + LA) % (KC / 4);
                a[nxt] = loadA(kn);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) b[nxt][nb] = loadB(kn, nb);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    dmma884(acc[0][nb][0], acc[0][nb][1], a[cur].x, b[cur][nb]);
                    dmma884(acc[1][nb][0], acc[1][nb][1], a[cur].y, b[cur][nb]);
                }
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    dmma884(acc[0][nb][0], acc[0][nb][1], (k4 & 1) ? ca.x : ca.y, cb[(nb + k4) % NB]);
                    dmma884(acc[1][nb][0], acc[1][nb][1], (k4 & 1) ? ca.y : ca.x, cb[(nb + k4) % NB]);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) s += acc[h][nb][0] + acc[h][nb][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double* g_out;
static int g_first = 1;

template <int NB, int MODE, int NCW>
static void run(int sms) {
    const int chunks = 4000;
    const int smem = (KC * (16 * NCW + 4) + NB * 8 * MPITCH) * 8;
    CK(cudaFuncSetAttribute(k_lds<NB, MODE, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_lds<NB, MODE, NCW><<<sms, NCW * 32, smem>>>(g_out, chunks, 0);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        k_lds<NB, MODE, NCW><<<sms, NCW * 32, smem>>>(g_out, chunks, 0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * 1.965e9 / (chunks * 8.0) / (NCW / 4.0);
    printf("%s  {\"nb\": %d, \"mode\": %d, \"warps\": %d, \"ms\": %.4f, \"cyc_per_k4_per_warp_slot\": %.2f, \"ideal\": %d, \"eff\": %.4f, \"extra\": %.2f}",
           g_first ? "" : ",\n", NB, MODE, NCW, best, cyc, 32 * NB, 32.0 * NB / cyc, cyc - 32.0 * NB);
    g_first = 0;
}

template <int NB, int NCW>
static void family(int sms) {
    run<NB, 0, NCW>(sms);
    run<NB, 1, NCW>(sms);
    run<NB, 6, NCW>(sms);
    run<NB, 2, NCW>(sms);
    run<NB, 3, NCW>(sms);
    run<NB, 4, NCW>(sms);
    run<NB, 5, NCW>(sms);
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    CK(cudaMalloc(&g_out, sizeof(double) * sms * 512));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", prop.name, sms);
    family<3, 16>(sms);
    family<3, 8>(sms);
    family<6, 8>(sms);
    family<6, 16>(sms);
    family<8, 8>(sms);
    printf("\n]}\n");
    return 0;
}
