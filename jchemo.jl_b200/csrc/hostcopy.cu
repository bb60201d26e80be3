// Host <-> device transfers of column-major matrices for the host-pointer entry points.
//
// The reference's callers hand over ordinary (pageable) Julia arrays.  A cudaMemcpy from/to pageable
// memory is staged by the driver through a single thread (measured here: 208 MB of scores took 42 ms,
// ~5 GB/s, of which most is first-touch page faults of the destination).  These helpers
//   * copy directly (one cudaMemcpy2DAsync) when the host array is page-locked, and otherwise
//   * stage through two pinned 32 MB slots filled / drained by a small pool of worker threads, so the
//     memcpy + page faults run on several cores while the DMA of the neighbouring slot is in flight.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "jcb_internal.cuh"

namespace jcb {

// ------------------------------------------------------------------------------------------ thread pool
class Pool {
  public:
    static Pool& get() {
        static Pool p;
        return p;
    }
    // runs fn(0..njobs-1) on the workers and the calling thread; returns when all are done
    void run(int njobs, const std::function<void(int)>& fn) {
        if (njobs <= 0) return;
        if (njobs == 1 || workers_.empty()) {
            for (int i = 0; i < njobs; ++i) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            njobs_ = njobs;
            next_ = 0;
            pending_ = njobs;
            ++epoch_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    Pool() {
        // workers + the calling thread: 8 copy threads (JCB_STAGE_THREADS overrides; measured on a 16-vCPU box:
        // 16 and 24 threads are slower than 8 — the staging is bound by host memory traffic, not by cores)
        unsigned hc = std::thread::hardware_concurrency();
        int n = (int)std::min<unsigned>(hc > 1 ? hc - 1 : 0, 7);
        if (const char* e = getenv("JCB_STAGE_THREADS")) n = std::max(0, std::min(atoi(e) - 1, 63));
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void work() {
        for (;;) {
            const std::function<void(int)>* fn;
            int i;
            {
                // job acquisition is atomic with the (fn, counter) pair of the current run
                std::lock_guard<std::mutex> lk(m_);
                if (!fn_ || next_ >= njobs_) return;
                i = next_++;
                fn = fn_;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int njobs_ = 0, pending_ = 0;
    int next_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

constexpr size_t SLOT_BYTES = 32u << 20;

static int ensure_staging(Ctx* c) {
    if (c->stage[0]) return 0;
    for (int s = 0; s < 2; ++s) {
        JCB_CUDA(cudaHostAlloc(&c->stage[s], SLOT_BYTES, cudaHostAllocPortable));
        JCB_CUDA(cudaEventCreateWithFlags(&c->stage_ev[s], cudaEventDisableTiming));
    }
    return 0;
}

// Copy with non-temporal stores: the staged bytes are read next by the DMA engine (host -> device) or not at
// all by these threads (device -> host), so they need not displace the cache, and a streaming store skips the
// read-for-ownership of the destination line — a third of the memory traffic of a plain memcpy.
static const bool g_stream_copy = getenv("JCB_STAGE_MEMCPY") == nullptr;     // JCB_STAGE_MEMCPY=1: plain memcpy
static void stream_copy(double* dst, const double* src, size_t n) {
#if defined(__x86_64__) && defined(__SSE2__)
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) { dst[i] = src[i]; ++i; }          // 16-byte align the stores
    for (; i + 8 <= n; i += 8) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 2));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 4));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 6));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 2), b);
        _mm_stream_si128((__m128i*)(dst + i + 4), c);
        _mm_stream_si128((__m128i*)(dst + i + 6), d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
#else
    memcpy(dst, src, n * 8);
#endif
}

// memcpy of a (rows x cols) tile between a dense buffer (ld = rows) and a strided host matrix
static void tile_memcpy(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows,
                        int64_t cols) {
    // split every column into pieces of <= 1 MB so that the pool has enough jobs
    const int64_t piece = 131072;
    const int64_t ppc = (rows + piece - 1) / piece;
    const int njobs = (int)(ppc * cols);
    Pool::get().run(njobs, [&](int j) {
        const int64_t col = j / ppc, r0 = (j % ppc) * piece;
        const int64_t nr = std::min(piece, rows - r0);
        if (g_stream_copy) stream_copy(dst + col * ldd + r0, src + col * lds + r0, (size_t)nr);
        else memcpy(dst + col * ldd + r0, src + col * lds + r0, (size_t)nr * 8);
    });
}

// tiles of at most SLOT_BYTES: whole columns when a column fits, row pieces otherwise
struct Tiler {
    int64_t rows, cols, rchunk, cchunk;
    Tiler(int64_t r, int64_t c) : rows(r), cols(c) {
        const int64_t slot = (int64_t)(SLOT_BYTES / 8);
        rchunk = std::min(rows, slot);
        cchunk = std::max<int64_t>(1, slot / std::max<int64_t>(rchunk, 1));
    }
    int64_t ntiles() const { return ((rows + rchunk - 1) / rchunk) * ((cols + cchunk - 1) / cchunk); }
    void tile(int64_t t, int64_t& r0, int64_t& nr, int64_t& c0, int64_t& nc) const {
        const int64_t nrt = (rows + rchunk - 1) / rchunk;
        const int64_t ct = t / nrt, rt = t % nrt;
        r0 = rt * rchunk;
        nr = std::min(rchunk, rows - r0);
        c0 = ct * cchunk;
        nc = std::min(cchunk, cols - c0);
    }
};

int h2d_2d(Ctx* c, double* dDst, int64_t ldd, const double* hSrc, int64_t lds, int64_t rows,
           int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    if (is_pinned(hSrc) || (size_t)rows * cols * 8 < (1u << 20)) {
        JCB_CUDA(cudaMemcpy2DAsync(dDst, ldd * 8, hSrc, lds * 8, rows * 8, cols, cudaMemcpyHostToDevice, st));
        return 0;
    }
    JCB_TRY(ensure_staging(c));
    Tiler tl(rows, cols);
    const int64_t nt = tl.ntiles();
    for (int64_t t = 0; t < nt; ++t) {
        const int s = (int)(t & 1);
        int64_t r0, nr, c0, nc;
        tl.tile(t, r0, nr, c0, nc);
        if (t >= 2) JCB_CUDA(cudaEventSynchronize(c->stage_ev[s]));     // slot drained by the DMA
        tile_memcpy((double*)c->stage[s], nr, hSrc + r0 + c0 * lds, lds, nr, nc);
        JCB_CUDA(cudaMemcpy2DAsync(dDst + r0 + c0 * ldd, ldd * 8, c->stage[s], nr * 8, nr * 8, nc,
                                   cudaMemcpyHostToDevice, st));
        JCB_CUDA(cudaEventRecord(c->stage_ev[s], st));
    }
    // the slots are reused by the next transfer: make sure the last DMAs have read them
    JCB_CUDA(cudaEventSynchronize(c->stage_ev[0]));
    if (nt > 1) JCB_CUDA(cudaEventSynchronize(c->stage_ev[1]));
    return 0;
}

int d2h_2d(Ctx* c, double* hDst, int64_t ldd, const double* dSrc, int64_t lds, int64_t rows,
           int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    if (is_pinned(hDst) || (size_t)rows * cols * 8 < (1u << 20)) {
        JCB_CUDA(cudaMemcpy2DAsync(hDst, ldd * 8, dSrc, lds * 8, rows * 8, cols, cudaMemcpyDeviceToHost, st));
        return 0;
    }
    JCB_TRY(ensure_staging(c));
    Tiler tl(rows, cols);
    const int64_t nt = tl.ntiles();
    auto issue = [&](int64_t t) -> int {
        const int s = (int)(t & 1);
        int64_t r0, nr, c0, nc;
        tl.tile(t, r0, nr, c0, nc);
        JCB_CUDA(cudaMemcpy2DAsync(c->stage[s], nr * 8, dSrc + r0 + c0 * lds, lds * 8, nr * 8, nc,
                                   cudaMemcpyDeviceToHost, st));
        JCB_CUDA(cudaEventRecord(c->stage_ev[s], st));
        return 0;
    };
    JCB_TRY(issue(0));
    for (int64_t t = 0; t < nt; ++t) {
        const int s = (int)(t & 1);
        if (t + 1 < nt) JCB_TRY(issue(t + 1));      // the other slot was drained in the previous iteration
        JCB_CUDA(cudaEventSynchronize(c->stage_ev[s]));
        int64_t r0, nr, c0, nc;
        tl.tile(t, r0, nr, c0, nc);
        tile_memcpy(hDst + r0 + c0 * ldd, ldd, (const double*)c->stage[s], nr, nr, nc);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ pinned pool
// Page-locked host blocks for large outputs (scores T): a D2H copy into pinned memory runs at PCIe speed
// (208 MB in ~4 ms) instead of ~14 ms through the staging threads.  cudaHostAlloc is slow (~0.3 ms/MB),
// so freed blocks are kept and reused by later calls of similar size.
struct PinnedBlock {
    void* p;
    size_t bytes;
    bool in_use;
};
static std::mutex g_pool_mutex;
static std::vector<PinnedBlock> g_pool;
constexpr size_t POOL_KEEP_BYTES = (size_t)6 << 30;     // a C5 predict sweep returns 4.1 GB

void* pinned_alloc(size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (auto& b : g_pool)
        if (!b.in_use && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + 4096) {
            b.in_use = true;
            return b.p;
        }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    g_pool.push_back(PinnedBlock{p, bytes, true});
    return p;
}

int pinned_free(void* p) {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    size_t idle = 0;
    for (auto& b : g_pool)
        if (!b.in_use) idle += b.bytes;
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (g_pool[i].p == p) {
            if (idle + g_pool[i].bytes > POOL_KEEP_BYTES) {
                cudaFreeHost(p);
                g_pool.erase(g_pool.begin() + i);
            } else {
                g_pool[i].in_use = false;
            }
            return 0;
        }
    return -1;
}

void pinned_release_all() {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (size_t i = 0; i < g_pool.size();)
        if (!g_pool[i].in_use) {
            cudaFreeHost(g_pool[i].p);
            g_pool.erase(g_pool.begin() + i);
        } else {
            ++i;
        }
}

void free_staging(Ctx* c) {
    for (int s = 0; s < 2; ++s) {
        if (c->stage[s]) {
            cudaFreeHost(c->stage[s]);
            cudaEventDestroy(c->stage_ev[s]);
            c->stage[s] = nullptr;
        }
    }
}

}  // namespace jcb
