// Internal declarations shared by the kernels and the C-ABI shim of libjchemo_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <string>

#include "../../include/jchemo_b200.h"

namespace jcb {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern int64_t g_launches;   // kernels launched by this library (guarded by the API mutex)

#define JCB_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            jcb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                      \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)

#define JCB_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        jcb::g_launches++;                                                                 \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            jcb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),     \
                           __FILE__, __LINE__);                                            \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)

#define JCB_TRY(call)                 \
    do {                              \
        int _r = (call);              \
        if (_r != 0) return _r;       \
    } while (0)

// ---------------------------------------------------------------- context
struct Buf {            // grow-only device scratch
    void* p = nullptr;
    size_t bytes = 0;
};

struct Ctx {
    bool ready = false;
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t out_stream = nullptr;   // device-to-host leg of the row-chunk pipelines (transform, predict)
    cudaEvent_t pipe_ev[4];              // [0..1] chunk computed, [2..3] chunk copied out
    cudaStream_t stream = nullptr;   // stream in use (own_stream or external)
    // K1 scratch
    Buf partials;        // split-K partial units
    void* sched_host = nullptr;  // pinned staging for a new schedule
    size_t sched_host_bytes = 0;
    // K1 schedules, cached by (columns seen by the kernel, stages): a streamed fit launches K1 on row chunks of up to
    // three different lengths per call, and a cache miss drains the stream
    struct Sched {
        int64_t p = -1, q = -1, nst = -1;
        int ngroups = 0, nsegs = 0;
        int64_t zone_len = 0;
        size_t off_segs = 0, off_cta = 0, off_gseg = 0;
        Buf dev;                 // device copy
        uint64_t stamp = 0;      // last use
    };
    static constexpr int NSCHED = 6;
    Sched sched[NSCHED];
    uint64_t sched_clock = 0;
    bool k1_attr_set = false;
    cudaEvent_t mg_ev[2];           // multi-GPU: [0] pivot / packed ready, [1] reduced
    // solve / xmul scratch
    Buf pivot_ws;
    void* pivot_ctr_zeroed = nullptr;   // where the pivot kernel's completion counter was last zeroed
    void* pivot_ctr_base = nullptr;     // ... and the workspace allocation it lived in
    Buf pivot_sample;               // device copy of the host row sample the streamed fit takes its pivot from
    void* pivot_host = nullptr;     // page-locked staging of that sample
    size_t pivot_host_bytes = 0;
    Buf locw_ws;
    Buf solve_ws;
    Buf xmul_ws;
    Buf coef_ws;         // B (p x q) and intercept of a single-k prediction on the device path
    const double* xmul_center_flag = nullptr;   // set around the fit's score pass: K1's centring decision (device)
    // general scratch for the host-pointer API
    Buf hX, hY, hW, hT, hSmall, hPred;
    Buf cvX, cvY, cvIdx, cvPk;      // gridcv: permuted copies, row map, per-segment packed buffers
    const void* cv_hostX = nullptr; // host pointers / shape of the copy resident in hX, hY (reuse_xy)
    const void* cv_hostY = nullptr;
    int64_t cv_n = 0, cv_p = 0, cv_q = 0;
    // per-launch K1 timing ring (bench.py's roofline: average K1 duration over the timed region)
    static constexpr int GRAM_RING = 256;
    cudaEvent_t gram_ev0[GRAM_RING], gram_ev1[GRAM_RING];
    int64_t gram_calls = 0;
    // pinned staging slots for pageable host arrays (hostcopy.cu)
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2];
    cudaEvent_t chunk_ev[3];
    cudaEvent_t blk_ev[8];               // host fit: row block b scored (and centred in place)
    // timing: a phase may occur several times in one call (row chunks, score blocks); every occurrence has
    // its own event pair and the reported time is their SUM (kernel time, not the span under the copies)
    static constexpr int PHASE_SLOTS = 24;
    cudaEvent_t ev_begin[JCB200_NPHASE][PHASE_SLOTS], ev_end[JCB200_NPHASE][PHASE_SLOTS];
    int ev_cnt[JCB200_NPHASE];        // completed occurrences
    bool ev_open[JCB200_NPHASE];      // begin recorded, end pending
    double last_ms[JCB200_NPHASE];
    // resident matrices (jcb200_resident_add): host arrays whose device copy is kept across calls
    struct Resident {
        const void* host;
        int64_t ld_host, rows, cols;
        double* dev;
        int64_t ld;
    };
    static constexpr int MAX_RESIDENT = 16;
    Resident resident[MAX_RESIDENT];
    int n_resident = 0;
};

Ctx* ctx();                               // the process-wide context (API mutex must be held)
int ensure(Buf& b, size_t bytes);         // grow-only cudaMalloc
void phase_begin(Ctx* c, int ph);
void phase_begin_on(Ctx* c, int ph, cudaStream_t st);
void phase_end_on(Ctx* c, int ph, cudaStream_t st);
int h2d_2d(Ctx* c, double* dDst, int64_t ldd, const double* hSrc, int64_t lds, int64_t rows, int64_t cols,
           cudaStream_t st);
int d2h_2d(Ctx* c, double* hDst, int64_t ldd, const double* dSrc, int64_t lds, int64_t rows, int64_t cols,
           cudaStream_t st);
void free_staging(Ctx* c);
bool is_pinned(const void* p);           // page-locked host memory (async copies run at PCIe speed)
void* pinned_alloc(size_t bytes);
int pinned_free(void* p);
void pinned_release_all();
void phase_end(Ctx* c, int ph);

// Where K1b writes the reduced Gram block: one buffer (the fit's own packed buffer), or — row-sharded fit, fused
// exchange — slot [rank] of EVERY rank's peer window, after which the last block raises flag [rank] in every window.
struct ReduceDst {
    int n = 1;                                 // destinations
    double* p[8] = {};
    unsigned int* done = nullptr;              // completion counter of the launch (own window); null: no signalling
    unsigned long long* flag[8] = {};          // flag [rank] of every window
    unsigned long long seq = 0;
};
// Where K3 reads the packed Gram: one buffer, or the sum over the `n` slots of the own peer window (rank order:
// the same bits on every rank), after waiting for the `n` flags of this exchange.
struct PackedSrc {
    const double* base = nullptr;
    int64_t stride = 0;
    int n = 1;
    const unsigned long long* flags = nullptr; // n flags, `flag_stride` apart; null: nothing to wait for
    int flag_stride = 16;
    unsigned long long seq = 0;
    unsigned int* timeouts = nullptr;          // counts waits that were given up
};

// ---------------------------------------------------------------- kernel launchers
int launch_pivot(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy, int64_t n,
                 int64_t p, int64_t q, double* d_pivot);
int launch_gram(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy,
                const double* dw, int64_t n, int64_t p, int64_t q, const double* d_pivot,
                double* d_packed, int accumulate);
int launch_gram_to(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy,
                   const double* dw, int64_t n, int64_t p, int64_t q, const double* d_pivot,
                   const ReduceDst& dst, int accumulate);
int launch_solve(Ctx* c, const double* d_packed, const double* d_pivot, int64_t p, int64_t q,
                 int nlv, int scal, double* dP, double* dR, double* dW, double* dC, double* dTT,
                 double* dxmeans, double* dxscales, double* dymeans, double* dyscales,
                 double* dsumw);
int launch_solve_src(Ctx* c, const PackedSrc& src, const double* d_pivot, int64_t p, int64_t q,
                     int nlv, int scal, double* dP, double* dR, double* dW, double* dC, double* dTT,
                     double* dxmeans, double* dxscales, double* dymeans, double* dyscales,
                     double* dsumw);
int launch_xmul(Ctx* c, const double* dX, int64_t ldx, int64_t m, int64_t p, const double* dmu,
                const double* dsigma, const double* dM, int64_t ldm, int ncol, const double* dbias,
                double* dOut, int64_t ldo);
int launch_predict_sweep(Ctx* c, const double* dX, int64_t ldx, int64_t m, int64_t p, int64_t q,
                         const double* dR, const double* dC, int a, const double* dxmeans,
                         const double* dxscales, const double* dymeans, const double* dyscales,
                         int k_lo, int k_hi, double* dPred);
int launch_center_scale(Ctx* c, double* dX, int64_t ldx, int64_t n, int64_t p, const double* dmu,
                        const double* dsigma);
int launch_weights(Ctx* c, const double* dw, int64_t n, const double* dsumw, double* dw_out);
int launch_fill_uniform(Ctx* c, double* d, int64_t ld, int64_t n_rows, int64_t n_cols,
                        uint64_t seed, int64_t row0, int64_t n_global);
int launch_sstot(Ctx* c, const double* dX, int64_t ldx, int64_t n, int64_t p, const double* dmu,
                 const double* dsigma, const double* dw, double* dpartial, int gx, double* dout);
int launch_locw(Ctx* c, const double* dXtr, int64_t ldxt, const double* dYtr, int64_t ldyt, int64_t ntr,
                const double* dX, int64_t ldx, int64_t m, int64_t p, int64_t q, const int64_t* d_idx,
                const int64_t* d_off, const double* d_w, int kmax, int k_lo, int k_hi, int scal, double* d_pred);
int launch_xfit(Ctx* c, const double* dX, int64_t ldx, double* dOut, int64_t ldo, int64_t m, int64_t p,
                const double* dT, int64_t ldt, const double* dP, int64_t ldp, int nlv, const double* dxm,
                const double* dxs, int resid);
int launch_coef(Ctx* c, const double* dR, const double* dC, const double* dxmeans,
                const double* dxscales, const double* dymeans, const double* dyscales, int64_t p,
                int64_t q, int k, double* dB, double* dint);

int launch_gridscore_gram(Ctx* c, const double* dY, int64_t ldy, const double* dT, int64_t ldt,
                          const double* dC, const double* dys, const double* dymeans, int64_t m, int q,
                          int k_hi, int ka, double* dYaug, int64_t lda, double* d_pivot0, double* d_packed);
int launch_gather_rows(Ctx* c, const double* src, int64_t lds, double* dst, int64_t ldd,
                       const int64_t* d_src_row, int64_t nrows, int64_t ncols);
int launch_packed_sub(Ctx* c, const double* a, const double* b, double* out, int64_t len);
int launch_packed_add(Ctx* c, double* acc, const double* b, int64_t len, int first);
void gridscore_from_packed(const double* pk, int ka, int q, int k_lo, int k_hi, const double* C,
                           const double* ys, double* ssr, double* sumres, double* ysum, double* ysumsq);

// peer-memory exchange (comm.cu); the API mutex must be held
int comm_create_locked(Ctx* c, int rank, int world, int64_t max_packed_len, void* handle_out);
int comm_connect_locked(const void* all_handles);
void comm_destroy_locked();
int comm_timeouts_locked();
int comm_pivot(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy, int64_t n, int64_t p,
               int64_t q, double* d_pivot);
int comm_allreduce(Ctx* c, double* d_packed, int64_t len);
int comm_gram(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy, const double* dw, int64_t n,
              int64_t p, int64_t q, const double* d_pivot);
int comm_solve(Ctx* c, const double* d_pivot, int64_t p, int64_t q, int nlv, int scal, double* dP, double* dR,
               double* dW, double* dC, double* dTT, double* dxmeans, double* dxscales, double* dymeans,
               double* dyscales, double* dsumw);

inline int64_t packed_len(int64_t p, int64_t q) { return p * p + p * q + q + p + q + 1; }
// offsets into the packed buffer
inline int64_t off_gxy(int64_t p, int64_t q) { return p * p; }
inline int64_t off_gyy(int64_t p, int64_t q) { return p * p + p * q; }
inline int64_t off_sx(int64_t p, int64_t q) { return p * p + p * q + q; }
inline int64_t off_sy(int64_t p, int64_t q) { return p * p + p * q + q + p; }
inline int64_t off_sw(int64_t p, int64_t q) { return p * p + p * q + q + p + q; }

}  // namespace jcb

// ---------------------------------------------------------------- device PTX helpers
#ifdef __CUDACC__
namespace jcb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {   // never blocks
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 2-D tiled TMA load: coordinates (c0 = innermost/row, c1 = column)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const CUtensorMap* map, int c0,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0)
        : "memory");
}
// 1-D bulk copy global -> shared (no tensor map): 16-byte aligned, size multiple of 16
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes,
                                          uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col); lowers to DMMA.8x8x4 on sm_100a.
// lane = 4*g + kk holds A[g][kk], B[kk][g], and C[g][2*kk], C[g][2*kk+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace jcb
#endif
