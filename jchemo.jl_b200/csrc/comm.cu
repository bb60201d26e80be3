// Peer-memory exchange of the row-sharded fit, one process per GPU.
//
// The path has ONE exchange: the sum of the packed partial Grams [Gxx | Gxy | gyy | sx | sy | sw] (2 MB at C2,
// 32 MB at C4), preceded by the pivot of rank 0 (p + q + 1 doubles).  Both go straight through peer HBM over
// NVLink / NVSwitch, mapped into every process with CUDA IPC — no collective library, no host round trip:
//
//   push   every rank copies its packed block, coalesced 16-byte stores, into slot [rank] of EVERY rank's window
//          (its own included); the last block to finish raises flag [rank] in every window (release, system
//          scope, after a system fence: the data has landed before the flag can be seen)
//   sum    the next kernel on the stream waits until all `world` flags of its own window carry this exchange's
//          sequence number (acquire, system scope) and adds the slots in rank order: local HBM reads only, the
//          same order on every rank, so every rank holds the same bits (K3/K4 then run redundantly and agree)
//
// Windows are double-buffered by the parity of the sequence number: a rank can only reach exchange k+2 after
// every peer has pushed k+1, which each peer does (stream order) after it has finished reading exchange k.
// The pivot uses the same scheme with rank 0 as the only writer.
#include <algorithm>
#include <cstddef>
#include <cstring>

#include "jcb_internal.cuh"

namespace jcb {

constexpr int COMM_MAX = 8;
constexpr int64_t COMM_PIVOT_CAP = 16384;           // doubles per pivot buffer (p + q + 1 <= 16384)
constexpr size_t COMM_HDR_BYTES = 4096;
constexpr int COMM_FLAG_STRIDE = 16;                // uint64 per flag: one 128-byte line each

struct CommHeader {                                  // lives at the start of every window
    unsigned long long flag_packed[COMM_MAX * COMM_FLAG_STRIDE];   // [r * 16]: rank r's block of exchange `seq` landed
    unsigned long long flag_pivot[COMM_FLAG_STRIDE];               // rank 0's pivot of fit `seq` landed
    unsigned int done;                                             // blocks of the running push that have finished
    unsigned int timeouts;                                         // waits given up (a peer never arrived)
};
// A wait that never ends would hang the GPU: after ~4 s (a peer has died or left the sequence) the waiter gives
// up, counts it in the header and poisons its result with NaN, which K3's non-finite check turns into an error.
constexpr long long COMM_WAIT_CYCLES = 8000000000ll;
static_assert(sizeof(CommHeader) <= COMM_HDR_BYTES, "header");

struct Comm {
    bool ready = false, connected = false;
    int rank = 0, world = 1;
    int64_t cap = 0;                 // doubles per slot
    unsigned char* win = nullptr;    // own window
    unsigned char* peer[COMM_MAX] = {};   // every rank's window in this process's address space
    unsigned long long seq_packed = 0, seq_pivot = 0;
    unsigned char** d_peer = nullptr;     // device copy of peer[]
};
static Comm g_comm;

static inline size_t comm_window_bytes(int world, int64_t cap) {
    return COMM_HDR_BYTES + 2 * (size_t)COMM_PIVOT_CAP * 8 + 2 * (size_t)world * (size_t)cap * 8;
}
__host__ __device__ static inline double* win_pivot(unsigned char* w, int parity) {
    return (double*)(w + COMM_HDR_BYTES) + (size_t)parity * COMM_PIVOT_CAP;
}
__host__ __device__ static inline double* win_slot(unsigned char* w, int parity, int world, int64_t cap, int r) {
    return (double*)(w + COMM_HDR_BYTES + 2 * (size_t)COMM_PIVOT_CAP * 8) + ((size_t)parity * world + r) * (size_t)cap;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// push: packed block -> slot [rank] of every window, then flag [rank] everywhere (last block)
__global__ void __launch_bounds__(256)
comm_push_kernel(const double* __restrict__ src, int64_t len, unsigned char* const* __restrict__ peers, int rank,
                 int world, int64_t cap, int parity, unsigned long long seq) {
    double* dst[COMM_MAX];
#pragma unroll
    for (int r = 0; r < COMM_MAX; ++r) dst[r] = r < world ? win_slot(peers[r], parity, world, cap, rank) : nullptr;
    const int64_t n2 = len >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = reinterpret_cast<const double2*>(src)[i];
#pragma unroll
        for (int r = 0; r < COMM_MAX; ++r)
            if (r < world) reinterpret_cast<double2*>(dst[r])[i] = v;
    }
    if ((len & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double v = src[len - 1];
        for (int r = 0; r < world; ++r) dst[r][len - 1] = v;
    }
    __threadfence_system();          // this thread's remote stores are performed before anything that follows
    __syncthreads();
    if (threadIdx.x == 0) {
        CommHeader* own = reinterpret_cast<CommHeader*>(peers[rank]);
        const unsigned int prev = atomicAdd(&own->done, 1u);
        if (prev == gridDim.x - 1) {          // every block has fenced its stores
            own->done = 0;
            __threadfence_system();
            for (int r = 0; r < world; ++r)
                st_release_sys(&reinterpret_cast<CommHeader*>(peers[r])->flag_packed[rank * COMM_FLAG_STRIDE], seq);
        }
    }
}

// sum: wait for the `world` flags of this exchange, then out = sum over ranks of slot[r] (rank order)
__global__ void __launch_bounds__(256)
comm_sum_kernel(double* __restrict__ out, int64_t len, unsigned char* win, int world, int64_t cap, int parity,
                unsigned long long seq) {
    CommHeader* hdr = reinterpret_cast<CommHeader*>(win);
    int late = 0;
    if (threadIdx.x < world) {
        const unsigned long long* f = &hdr->flag_packed[threadIdx.x * COMM_FLAG_STRIDE];
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            __nanosleep(40);
            if (clock64() - t0 > COMM_WAIT_CYCLES) {
                late = 1;
                break;
            }
        }
    }
    late = __syncthreads_or(late);
    if (late) {
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            atomicAdd(&hdr->timeouts, 1u);
            out[len - 1] = __longlong_as_double(0x7ff8000000000000ll);     // sum(w) = NaN: K3 flags the fit
        }
        return;
    }
    const double* s0 = win_slot(win, parity, world, cap, 0);
    const int64_t n2 = len >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v[COMM_MAX];
#pragma unroll
        for (int r = 0; r < COMM_MAX; ++r)
            if (r < world) v[r] = __ldcg(reinterpret_cast<const double2*>(s0 + (size_t)r * cap) + i);
        double2 s = v[0];
#pragma unroll
        for (int r = 1; r < COMM_MAX; ++r)
            if (r < world) {
                s.x += v[r].x;
                s.y += v[r].y;
            }
        reinterpret_cast<double2*>(out)[i] = s;
    }
    if ((len & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(s0 + (size_t)r * cap + len - 1);
        out[len - 1] = s;
    }
}

// rank 0: pivot -> every peer's pivot buffer, then the pivot flag there
__global__ void __launch_bounds__(256)
comm_pivot_publish_kernel(const double* __restrict__ pivot, int npv, unsigned char* const* __restrict__ peers,
                          int world, int parity, unsigned long long seq) {
    for (int r = 1; r < world; ++r) {
        double* dst = win_pivot(peers[r], parity);
        for (int i = threadIdx.x; i < npv; i += blockDim.x) dst[i] = pivot[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        for (int r = 1; r < world; ++r)
            st_release_sys(&reinterpret_cast<CommHeader*>(peers[r])->flag_pivot[0], seq);
}

// rank > 0: wait for rank 0's pivot of this fit, copy it out of the own window
__global__ void __launch_bounds__(256)
comm_pivot_fetch_kernel(double* __restrict__ pivot, int npv, unsigned char* win, int parity, unsigned long long seq) {
    CommHeader* hdr = reinterpret_cast<CommHeader*>(win);
    int late = 0;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(&hdr->flag_pivot[0]) < seq) {
            __nanosleep(40);
            if (clock64() - t0 > COMM_WAIT_CYCLES) {
                late = 1;
                atomicAdd(&hdr->timeouts, 1u);
                break;
            }
        }
    }
    late = __syncthreads_or(late);
    const double* src = win_pivot(win, parity);
    for (int i = threadIdx.x; i < npv; i += blockDim.x)
        pivot[i] = late ? __longlong_as_double(0x7ff8000000000000ll) : __ldcg(src + i);
}

void comm_destroy_locked() {
    Comm& m = g_comm;
    if (!m.ready) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < m.world; ++r)
        if (r != m.rank && m.peer[r]) cudaIpcCloseMemHandle(m.peer[r]);
    if (m.d_peer) cudaFree(m.d_peer);
    if (m.win) cudaFree(m.win);
    m = Comm();
}

int comm_create_locked(Ctx* c, int rank, int world, int64_t max_packed_len, void* handle_out) {
    Comm& m = g_comm;
    if (m.ready) {
        set_error("comm_create: a communicator exists; call jcb200_comm_destroy first");
        return JCB200_EINVAL;
    }
    if (world < 1 || world > COMM_MAX || rank < 0 || rank >= world || max_packed_len < 1 || !handle_out) {
        set_error("comm_create: need 1 <= world <= %d, 0 <= rank < world, max_packed_len > 0", COMM_MAX);
        return JCB200_EINVAL;
    }
    m.rank = rank;
    m.world = world;
    m.cap = (max_packed_len + 1) & ~(int64_t)1;
    const size_t bytes = comm_window_bytes(world, m.cap);
    cudaError_t e = cudaMalloc(&m.win, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("comm_create: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        m = Comm();
        return JCB200_ENOMEM;
    }
    JCB_CUDA(cudaMemset(m.win, 0, COMM_HDR_BYTES));
    JCB_CUDA(cudaMalloc(&m.d_peer, COMM_MAX * sizeof(unsigned char*)));
    JCB_CUDA(cudaDeviceSynchronize());
    memset(handle_out, 0, JCB200_IPC_HANDLE_BYTES);
    if (world > 1) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) <= JCB200_IPC_HANDLE_BYTES, "IPC handle size");
        JCB_CUDA(cudaIpcGetMemHandle(&h, m.win));
        memcpy(handle_out, &h, sizeof(h));
    }
    m.peer[rank] = m.win;
    m.ready = true;
    m.connected = world == 1;
    if (m.connected) JCB_CUDA(cudaMemcpy(m.d_peer, m.peer, sizeof(m.peer), cudaMemcpyHostToDevice));
    (void)c;
    return 0;
}

int comm_connect_locked(const void* all_handles) {
    Comm& m = g_comm;
    if (!m.ready || !all_handles) {
        set_error("comm_connect: call jcb200_comm_create first");
        return JCB200_EINVAL;
    }
    if (m.connected) return 0;
    for (int r = 0; r < m.world; ++r) {
        if (r == m.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)all_handles + (size_t)r * JCB200_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("comm_connect: cudaIpcOpenMemHandle of rank %d's window failed: %s", r, cudaGetErrorString(e));
            return (int)e;
        }
        m.peer[r] = (unsigned char*)ptr;
    }
    JCB_CUDA(cudaMemcpy(m.d_peer, m.peer, sizeof(m.peer), cudaMemcpyHostToDevice));
    m.connected = true;
    return 0;
}

int comm_timeouts_locked() {
    Comm& m = g_comm;
    if (!m.ready) return 0;
    unsigned int t = 0;
    cudaMemcpy(&t, m.win + offsetof(CommHeader, timeouts), sizeof(t), cudaMemcpyDeviceToHost);
    return (int)t;
}

int comm_pivot(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy, int64_t n, int64_t p,
               int64_t q, double* d_pivot) {
    Comm& m = g_comm;
    if (!m.ready || !m.connected) {
        set_error("comm_pivot: no connected communicator");
        return JCB200_EINVAL;
    }
    const int npv = (int)(p + q + 1);
    if (npv > COMM_PIVOT_CAP) {
        set_error("comm_pivot: p + q + 1 = %d exceeds %lld", npv, (long long)COMM_PIVOT_CAP);
        return JCB200_EINVAL;
    }
    const unsigned long long seq = ++m.seq_pivot;
    const int parity = (int)(seq & 1);
    if (m.rank == 0) {
        JCB_TRY(launch_pivot(c, dX, ldx, dY, ldy, n, p, q, d_pivot));
        if (m.world > 1) {
            comm_pivot_publish_kernel<<<1, 256, 0, c->stream>>>(d_pivot, npv, m.d_peer, m.world, parity, seq);
            JCB_LAUNCH_CHECK();
        }
    } else {
        comm_pivot_fetch_kernel<<<1, 256, 0, c->stream>>>(d_pivot, npv, m.win, parity, seq);
        JCB_LAUNCH_CHECK();
    }
    return 0;
}

// Fused exchange of the sharded fit: K1b writes this rank's reduced block straight into slot [rank] of every
// window (its own included) and raises the flags; comm_solve then runs K3 on the SUM of the slots.  No kernel of
// the exchange's own, no extra pass over the block.
int comm_gram(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy, const double* dw, int64_t n,
              int64_t p, int64_t q, const double* d_pivot) {
    Comm& m = g_comm;
    if (!m.ready || !m.connected) {
        set_error("comm_gram: no connected communicator");
        return JCB200_EINVAL;
    }
    const int64_t len = packed_len(p, q);
    if (len > m.cap) {
        set_error("comm_gram: packed length %lld exceeds the window capacity %lld", (long long)len, (long long)m.cap);
        return JCB200_EINVAL;
    }
    const unsigned long long seq = ++m.seq_packed;
    const int parity = (int)(seq & 1);
    if (n <= 0) {
        // a rank without rows contributes zeros: push them from a zeroed scratch block
        JCB_TRY(ensure(c->pivot_sample, (size_t)len * 8 + 16));
        double* z = (double*)c->pivot_sample.p;
        JCB_CUDA(cudaMemsetAsync(z, 0, (size_t)len * 8, c->stream));
        const int grid = (int)std::min<int64_t>(((len >> 1) + 255) / 256, 4 * (int64_t)c->num_sms);
        comm_push_kernel<<<grid > 0 ? grid : 1, 256, 0, c->stream>>>(z, len, m.d_peer, m.rank, m.world, m.cap, parity, seq);
        JCB_LAUNCH_CHECK();
        return 0;
    }
    ReduceDst dst;
    dst.n = m.world;
    for (int r = 0; r < m.world; ++r) {
        // own window first: K1b's accumulate-free stores go to dst.p[*] alike, order is irrelevant
        dst.p[r] = win_slot(m.peer[r], parity, m.world, m.cap, m.rank);
        dst.flag[r] = &reinterpret_cast<CommHeader*>(m.peer[r])->flag_packed[m.rank * COMM_FLAG_STRIDE];
    }
    dst.done = &reinterpret_cast<CommHeader*>(m.win)->done;
    dst.seq = seq;
    return launch_gram_to(c, dX, ldx, dY, ldy, dw, n, p, q, d_pivot, dst, 0);
}

int comm_solve(Ctx* c, const double* d_pivot, int64_t p, int64_t q, int nlv, int scal, double* dP, double* dR,
               double* dW, double* dC, double* dTT, double* dxmeans, double* dxscales, double* dymeans,
               double* dyscales, double* dsumw) {
    Comm& m = g_comm;
    if (!m.ready || !m.connected || m.seq_packed == 0) {
        set_error("comm_solve: no exchange in flight (call jcb200_comm_gram_dev first)");
        return JCB200_EINVAL;
    }
    const unsigned long long seq = m.seq_packed;
    const int parity = (int)(seq & 1);
    PackedSrc src;
    src.base = win_slot(m.win, parity, m.world, m.cap, 0);
    src.stride = m.cap;
    src.n = m.world;
    CommHeader* hdr = reinterpret_cast<CommHeader*>(m.win);
    src.flags = hdr->flag_packed;
    src.flag_stride = COMM_FLAG_STRIDE;
    src.seq = seq;
    src.timeouts = &hdr->timeouts;
    return launch_solve_src(c, src, d_pivot, p, q, nlv, scal, dP, dR, dW, dC, dTT, dxmeans, dxscales, dymeans,
                            dyscales, dsumw);
}

int comm_allreduce(Ctx* c, double* d_packed, int64_t len) {
    Comm& m = g_comm;
    if (!m.ready || !m.connected) {
        set_error("comm_allreduce: no connected communicator");
        return JCB200_EINVAL;
    }
    if (len < 1 || len > m.cap) {
        set_error("comm_allreduce: length %lld exceeds the window capacity %lld", (long long)len, (long long)m.cap);
        return JCB200_EINVAL;
    }
    if (((uintptr_t)d_packed & 15) != 0) {
        set_error("comm_allreduce: buffer must be 16-byte aligned");
        return JCB200_EALIGN;
    }
    if (m.world == 1) return 0;
    const unsigned long long seq = ++m.seq_packed;
    const int parity = (int)(seq & 1);
    // enough blocks to keep the NVLink write queues full, few enough that the completion counter is cheap
    const int grid = (int)std::min<int64_t>(((len >> 1) + 255) / 256, 4 * (int64_t)c->num_sms);
    phase_begin(c, JCB200_T_REDUCE);
    comm_push_kernel<<<grid > 0 ? grid : 1, 256, 0, c->stream>>>(d_packed, len, m.d_peer, m.rank, m.world, m.cap,
                                                                parity, seq);
    JCB_LAUNCH_CHECK();
    comm_sum_kernel<<<grid > 0 ? grid : 1, 256, 0, c->stream>>>(d_packed, len, m.win, m.world, m.cap, parity, seq);
    JCB_LAUNCH_CHECK();
    phase_end(c, JCB200_T_REDUCE);
    return 0;
}

}  // namespace jcb
