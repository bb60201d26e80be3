// K5 / K6 — streaming FP64 tensor-core GEMM over the rows of X:
//     Out = bias' + ((X - mu) ./ sigma) * M           (M is p x ncol, ncol small)
// K5: scores T = Xc R (fit, /root/reference/src/plskern.jl:162,170) and transform (:187-195);
//     single-k predictions with M = B_k, bias = ymeans (:233-234).
// K6: predict for a whole contiguous range k_lo:k_hi (:226-238) in ONE pass over X: the tile of
//     scores T = Xc R is formed once and pred_k = ymeans + sum_{j<k} t_j (c_j .* yscales)' is
//     accumulated over k in the epilogue, writing one m x q matrix per k.
//
// Design: X is column-major, so a tile of 128 rows x 32 columns is 32 contiguous 1 KB column
// segments; the producer warp brings each with one bulk async copy (cp.async.bulk, completes on an
// mbarrier) into a [column][132]-pitched shared tile — the 132 pitch makes the A-fragment loads
// (128-bit, two rows per lane) conflict free.  M is pre-packed (1/sigma folded in, zero padded) in the
// exact staged layout, one bulk copy per stage.  8 consumer warps own 16 rows each; centring happens
// on the fragments.  Results are staged through shared memory so global stores are 128-byte rows.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "jcb_internal.cuh"

namespace jcb {

// rows per tile = 16 per consumer warp; shared pitch of a column = rows + 4 (== 4 mod 16: conflict free)
#ifndef JCB_XM_KC
#define JCB_XM_KC 32
#endif
#ifndef JCB_XM_OCC
#define JCB_XM_OCC 1
#endif
constexpr int XM_KC = JCB_XM_KC;  // columns of X per stage (<= 32: one bulk copy per producer lane)
constexpr int XM_MPITCH = XM_KC + 4;   // shared pitch of a row of packed M^T (== 4 mod 16: conflict free)
constexpr int XM_OCC = JCB_XM_OCC;     // CTAs per SM the launch is sized for
// consumer warps per CTA: 8 (128-row tiles) or 16 (256-row tiles, 2 KB column segments, 4 warps per SMSP:
// the kernel is latency bound with 2) — template parameter NCW
constexpr int XM_MAXNB = 8;       // up to 64 output columns per pass

struct XmulParams {
    const double* X;
    int64_t ldx;
    int64_t m;
    int p;
    int nchunk;            // ceil(p / 32)
    const double* Mt;      // packed: [chunk][NP][36]
    const double* mu;      // padded to nchunk*32
    const double* zeros;   // >= 128 zero doubles (source for padding columns)
    const double* bias;    // NP (zero padded) or nullptr
    double* Out;
    int64_t ldo;
    int ncol;              // real output columns in this pass
    int aligned;           // X is 16-byte aligned with even ldx: bulk copies allowed
    int out_aligned;       // Out is 16-byte aligned with an even ldo: the direct epilogue may use 16-byte stores
    int direct;            // !SWEEP: results go from the accumulator fragments straight to global memory (no
                           // staging tile in shared memory: room for one more pipeline stage)
    const double* cflag;   // device flag of the fit (pivot[p+q]): != 1.0 = every column has mean^2 <= 64 var, so the
                           // scores may be formed as X M - mu'M (no DADD per element beside the DMMAs); else null
    // sweep epilogue
    const double* Cy;      // [a][q]: C[j,k] * yscales[j]
    const double* ymeans;
    double* Pred;          // (k_hi-k_lo+1) matrices m x q
    int q, k_lo, k_hi;
    int nstage;
};

// packs M (p x ncol, ld ldm) into Mt[chunk][n][36] with 1/sigma folded in; pads mu
__global__ void xmul_pack_kernel(const double* __restrict__ M, int64_t ldm, const double* __restrict__ sigma,
                                 const double* __restrict__ mu, int p, int ncol, int NP, int nchunk,
                                 double* __restrict__ Mt, double* __restrict__ mu_pad, double* __restrict__ zeros) {
    // the 128 zero doubles that stand in for padding columns: written here, by a kernel — a cudaMemsetAsync may
    // be scheduled on the copy engine that is busy with the host-to-device stream of the row-chunk pipelines,
    // and would hold the compute stream back until the whole transfer has finished
    if (blockIdx.x == 0 && threadIdx.x < 128) zeros[threadIdx.x] = 0.0;
    const int64_t total = (int64_t)nchunk * NP * XM_MPITCH;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % XM_MPITCH);
        const int n = (int)((e / XM_MPITCH) % NP);
        const int ch = (int)(e / ((int64_t)XM_MPITCH * NP));
        const int k = ch * XM_KC + kk;
        double v = 0.0;
        if (kk < XM_KC && k < p && n < ncol) v = M[k + (int64_t)n * ldm] / (sigma ? sigma[k] : 1.0);
        Mt[e] = v;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nchunk * XM_KC; k += gridDim.x * blockDim.x)
        mu_pad[k] = k < p ? mu[k] : 0.0;
}

// the sweep epilogue (51 prefix sums and 51 x q stores per row) runs on its own warps so that it overlaps the
// consumers' DMMA work on the next tile instead of serialising with it.  Its DFMAs queue behind the DMMAs on
// the shared FP64 pipe, so the work is spread over many warps: NCW - 1 of them (with the producer warp the CTA then
// has 2 NCW warps = 512 threads at NCW = 8, i.e. 128 registers per thread for the consumers; one warp more and
// ptxas caps everybody at 96), tasks of (32 rows, one response) dealt round-robin.
template <bool SWEEP, int NCW>
__host__ __device__ constexpr int xm_epw() { return SWEEP ? NCW - 1 : 0; }

template <int NPB, int NEX, bool SWEEP, int NCW, bool DIRECT>
__global__ void __launch_bounds__((NCW + 1 + xm_epw<SWEEP, NCW>()) * 32, XM_OCC)
xmul_kernel(const XmulParams prm) {
    constexpr int XM_NCW = NCW;
    constexpr int XM_MT = 16 * NCW;
    constexpr int XM_PITCH = XM_MT + 4;
    constexpr int XM_EPW = xm_epw<SWEEP, NCW>();
    constexpr int XM_THREADS = (NCW + 1 + XM_EPW) * 32;
    // NPB column blocks go through DMMA; NEX (<= 2) leftover columns are plain DFMA dot products on the
    // A fragments this lane already holds (a whole padded 8-column block for 1-2 columns would cost
    // 1/NPB more DMMA time: at nlv = 25 the score GEMM drops from 4 to 3 blocks)
    constexpr int NP = NPB * 8;
    constexpr int NPT = NP + NEX;
    constexpr int XBYTES = XM_KC * XM_PITCH * 8;        // 33792
    constexpr int MBYTES = NPT * XM_MPITCH * 8;
    constexpr int STAGE = XBYTES + MBYTES;
    extern __shared__ __align__(128) unsigned char smem[];
    const int nstage = prm.nstage;
    unsigned char* stage_base = smem;
    double* out_s = reinterpret_cast<double*>(smem + (size_t)nstage * STAGE);   // [NPT][132]
    // DIRECT is a template parameter: the epilogue that is not used must not be in the kernel (register allocation
    // and the placement of the leftover DFMAs in the main loop moved by 5 % with unrelated epilogue code present)
    constexpr bool direct = !SWEEP && DIRECT;
    double* mu_s = out_s + (direct ? 0 : NPT * XM_PITCH);                         // nchunk*32
    double* cy_s = mu_s + prm.nchunk * XM_KC;                                    // ncol*q (sweep only)
    double* cb_s = cy_s + (SWEEP ? ((prm.ncol * prm.q + 1) & ~1) : 0);          // NPT (+ pad): -mu'M, centre-free path
    uint64_t* full = reinterpret_cast<uint64_t*>(cb_s + 64 + 8);
    uint64_t* empty = full + nstage;
    uint64_t* tfull = empty + nstage;      // sweep: score tile in out_s complete (consumers -> epilogue warps)
    uint64_t* tempty = tfull + 1;          // sweep: out_s read out (epilogue warps -> consumers)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < prm.nchunk * XM_KC; k += XM_THREADS) mu_s[k] = prm.mu[k];
    if (SWEEP)
        for (int k = threadIdx.x; k < prm.ncol * prm.q; k += XM_THREADS) cy_s[k] = prm.Cy[k];
    // Centring puts two DADDs per k4-step between the fragment load and the six DMMAs that use it (no pipe time —
    // bench/fp64_mix.cu — but latency in this warp's chain: the centred kernel is 2-3 % slower, 0.885 vs 0.864 ms at
    // C2).  When the fit's pivot pass found every column well
    // scaled about zero (mean^2 <= 64 var: at most two digits to lose), T = X M - mu'M: raw fragments into
    // the DMMAs and one constant per output column in the epilogue.
    const bool center = SWEEP || !(prm.cflag != nullptr && *prm.cflag != 1.0);
    if (!SWEEP && !center) {
        __syncthreads();                                   // mu_s complete
        for (int col = warp; col < NPT; col += XM_THREADS / 32) {
            double sacc = 0.0;
            for (int k = lane; k < prm.nchunk * XM_KC; k += 32)
                sacc += mu_s[k] * prm.Mt[((int64_t)(k / XM_KC) * NPT + col) * XM_MPITCH + (k % XM_KC)];
#pragma unroll
            for (int o = 16; o; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) cb_s[col] = -sacc;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], XM_NCW);
        }
        if (SWEEP) {
            mbar_init(tfull, XM_NCW);
            mbar_init(tempty, XM_EPW);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t ntiles = (prm.m + XM_MT - 1) / XM_MT;
    const int nchunk = prm.nchunk;
    uint32_t it = 0;

    if (warp == XM_NCW) {
        // ------------------------------------------------------------------ producer warp
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int64_t row0 = t * XM_MT;
            const int rows = (int)min((int64_t)XM_MT, prm.m - row0);
            // aligned shards have an even leading dimension, so a ragged last tile may copy one padding
            // row (rows rounded up to even: bulk copies move multiples of 16 bytes) and stay in bounds
            const bool bulk = prm.aligned != 0;
            const int crow = (rows + 1) & ~1;
            for (int ch = 0; ch < nchunk; ++ch, ++it) {
                const int buf = it % nstage;
                const uint32_t ph = (it / nstage) & 1;
                if (lane == 0) mbar_wait(&empty[buf], ph ^ 1);
                __syncwarp();
                double* xs = reinterpret_cast<double*>(stage_base + (size_t)buf * STAGE);
                double* ms = xs + XM_KC * XM_PITCH;
                const double* msrc = prm.Mt + (int64_t)ch * NPT * XM_MPITCH;
                if (bulk) {
                    if (lane == 0) mbar_arrive_expect_tx(&full[buf], XM_KC * crow * 8 + MBYTES);
                    __syncwarp();
                    const int k = ch * XM_KC + lane;
                    const double* src = k < prm.p ? prm.X + row0 + (int64_t)k * prm.ldx : prm.zeros;
                    if (lane < XM_KC) bulk_load(xs + lane * XM_PITCH, src, crow * 8, &full[buf]);
                    if (lane == 0) bulk_load(ms, msrc, MBYTES, &full[buf]);
                } else {
                    // ragged / unaligned tile: plain loads through registers
                    for (int kc = 0; kc < XM_KC; ++kc) {
                        const int k = ch * XM_KC + kc;
                        for (int r = lane; r < XM_MT; r += 32)
                            xs[kc * XM_PITCH + r] =
                                (k < prm.p && r < rows) ? prm.X[row0 + r + (int64_t)k * prm.ldx] : 0.0;
                    }
                    for (int e = lane; e < NPT * XM_MPITCH; e += 32) ms[e] = msrc[e];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[buf]);   // release: generic-proxy writes above
                }
            }
        }
        return;
    }

    if (SWEEP && warp > XM_NCW) {
        // ------------------------------------------------------------------ sweep epilogue warps
        // pred_k[row, j] = ymeans[j] + sum_{l<k} T[row, l] * Cy[l][j].  A task = (group of 32 rows, response j):
        // one row per lane, k outer, up to 6 tasks of a warp carried in registers (independent DFMA chains), Cy
        // from shared memory; a warp's store covers 32 consecutive rows of one column (256 bytes)
        const int ew = warp - XM_NCW - 1;
        constexpr int NRG = XM_MT / 32, NT = 6;
        const int q = prm.q;
        const int ntask = NRG * q;
        const int64_t msz = prm.m * (int64_t)q;
        uint32_t tn = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tn) {
            const int64_t row0 = t * XM_MT;
            mbar_wait(tfull, tn & 1);
            for (int base = ew; base < ntask; base += XM_EPW * NT) {
                double pv[NT];
                const double* ts[NT];
                const double* cy[NT];
                double* dk[NT];
                bool ok[NT];
#pragma unroll
                for (int u = 0; u < NT; ++u) {
                    const int task = base + u * XM_EPW;
                    const bool valid = task < ntask;
                    const int rg = valid ? task / q : 0, j = valid ? task - rg * q : 0;
                    const int r = rg * 32 + lane;
                    pv[u] = prm.ymeans[j];
                    ts[u] = out_s + r;
                    cy[u] = cy_s + j;
                    dk[u] = prm.Pred + row0 + r + (int64_t)j * prm.m;
                    ok[u] = valid && row0 + r < prm.m;
                }
                for (int k = 0; k <= prm.k_hi; ++k) {
                    if (k >= prm.k_lo) {
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
                            if (ok[u]) __stcs(dk[u], pv[u]);
                            dk[u] += msz;
                        }
                    }
                    if (k < prm.k_hi) {
#pragma unroll
                        for (int u = 0; u < NT; ++u) pv[u] = fma(ts[u][k * XM_PITCH], cy[u][k * q], pv[u]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    uint32_t tn = 0;
    const int g = lane >> 2, kk = lane & 3;
    const int m0 = warp * 16;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * XM_MT;
        const int rows = (int)min((int64_t)XM_MT, prm.m - row0);
        double acc[2][NPB][2];
        double ex[2][NEX > 0 ? NEX : 1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int nb = 0; nb < NPB; ++nb) acc[h][nb][0] = acc[h][nb][1] = 0.0;
#pragma unroll
            for (int e = 0; e < NEX; ++e) ex[h][e] = 0.0;
        }
        auto chunk_loop = [&](auto cen_tag) {
            constexpr bool CEN = decltype(cen_tag)::value;
            for (int ch = 0; ch < nchunk; ++ch, ++it) {
                const int buf = it % nstage;
                const uint32_t ph = (it / nstage) & 1;
                mbar_wait(&full[buf], ph);
                const double* xs = reinterpret_cast<const double*>(stage_base + (size_t)buf * STAGE);
                const double* ms = xs + XM_KC * XM_PITCH;
                const double* mus = mu_s + ch * XM_KC;
#pragma unroll
                for (int k4 = 0; k4 < XM_KC / 4; ++k4) {
                    const int k = k4 * 4 + kk;
                    double2 a = *reinterpret_cast<const double2*>(xs + k * XM_PITCH + m0 + 2 * g);
                    if (CEN) {
                        const double mk = mus[k];
                        a.x -= mk;
                        a.y -= mk;
                    }
#pragma unroll
                    for (int nb = 0; nb < NPB; ++nb) {
                        const double b = ms[(nb * 8 + g) * XM_MPITCH + k];
                        dmma(acc[0][nb][0], acc[0][nb][1], a.x, b);
                        dmma(acc[1][nb][0], acc[1][nb][1], a.y, b);
                    }
#pragma unroll
                    for (int e = 0; e < NEX; ++e) {
                        const double b = ms[(NP + e) * XM_MPITCH + k];
                        ex[0][e] += a.x * b;
                        ex[1][e] += a.y * b;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[buf]);
            }
        };
        if (center) chunk_loop(std::true_type{});
        else chunk_loop(std::false_type{});
        if constexpr (direct) {
            // ---- epilogue, direct: lane (g, kk) holds rows m0 + 2g, m0 + 2g + 1 (h = 0, 1) of columns
            // nb*8 + 2kk (+1).  The two rows of a column go out as ONE 16-byte store, so a warp-wide store covers
            // four columns x 128 contiguous bytes (16 full sectors instead of 32 half-written ones: -5 % at
            // nlv = 16, neutral at 24+).  Unaligned outputs and the ragged last rows fall back to 8-byte stores.
            // What the result stores cost (round 2, JCB_XM_DBG_L2 builds, profiles/r02_xmul_decomposition.md):
            // 2.7 us of a 30 us tile at nlv = 24 although they are 5 % of the bytes — they queue in the LSU in front
            // of the next tile's fragment loads.  Sending them through the TMA engine instead (results written over
            // the warp's own rows of the tile's last stage, one 128-byte bulk copy per column) was built and
            // measured: -3.5 % at nlv = 24, +5 % at nlv = 25, +2 % at 16 — the bulk stores queue behind the
            // producer's loads and hold the stage back; not kept.
            const int row = m0 + 2 * g;
            double* o = prm.Out + row0 + row;
            const bool vec = prm.out_aligned && row + 1 < rows;
#pragma unroll
            for (int nb = 0; nb < NPB; ++nb) {
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                    const int col = nb * 8 + 2 * kk + e2;
                    if (col < prm.ncol) {
                        double v0 = acc[0][nb][e2], v1 = acc[1][nb][e2];
                        double add = 0.0;
                        if (!center) add += cb_s[col];
                        if (prm.bias) add += prm.bias[col];
                        v0 += add;
                        v1 += add;
                        double* oc = o + (int64_t)col * prm.ldo;
                        if (vec) {
                            *reinterpret_cast<double2*>(oc) = make_double2(v0, v1);
                        } else {
                            if (row < rows) oc[0] = v0;
                            if (row + 1 < rows) oc[1] = v1;
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int e = 0; e < NEX; ++e) {
                    double v = ex[h][e];                  // partial over this lane's k (kk); sum the 4 kk lanes
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int col = NP + e;
                    if (kk == 0 && row + h < rows && col < prm.ncol) {
                        if (!center) v += cb_s[col];
                        if (prm.bias) v += prm.bias[col];
                        o[h + (int64_t)col * prm.ldo] = v;
                    }
                }
            }
            continue;
        }
        // ---- epilogue: fragments -> shared (this warp's 16 rows) -> 128-byte global rows
        if (SWEEP && tn > 0) mbar_wait(tempty, (tn - 1) & 1);     // the previous tile has been read out
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nb = 0; nb < NPB; ++nb) {
                out_s[(nb * 8 + 2 * kk) * XM_PITCH + m0 + 2 * g + h] = acc[h][nb][0];
                out_s[(nb * 8 + 2 * kk + 1) * XM_PITCH + m0 + 2 * g + h] = acc[h][nb][1];
            }
#pragma unroll
        for (int e = 0; e < NEX; ++e)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double v = ex[h][e];                      // partial over this lane's k (kk); sum the 4 kk lanes
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (kk == 0) out_s[(NP + e) * XM_PITCH + m0 + 2 * g + h] = v;
            }
        __syncwarp();
        const int r = lane & 15, half = lane >> 4;
        const bool rok = (m0 + r) < rows;
        if (!SWEEP) {
            for (int col = half; col < prm.ncol; col += 2) {
                if (rok) {
                    double v = out_s[col * XM_PITCH + m0 + r];
                    if (!center) v += cb_s[col];
                    if (prm.bias) v += prm.bias[col];
                    prm.Out[row0 + m0 + r + (int64_t)col * prm.ldo] = v;
                }
            }
        } else {
            if (lane == 0) mbar_arrive(tfull);      // (after the __syncwarp above) hand the tile over
            ++tn;
        }
        __syncwarp();
    }
}

// Cy[k][j] = C[j,k] * yscales[j]
__global__ void sweep_cy_kernel(const double* __restrict__ C, const double* __restrict__ ys, int q, int a,
                                double* __restrict__ Cy) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a * q) return;
    const int k = e / q, j = e - k * q;
    Cy[e] = C[j + (int64_t)k * q] * ys[j];
}

template <int NPB, int NEX, bool SWEEP, int NCW>
static int launch_xmul_w(Ctx* c, XmulParams& prm) {
    constexpr int NP = NPB * 8 + NEX;
    constexpr int XM_MT = 16 * NCW;
    constexpr int XM_PITCH = XM_MT + 4;
    constexpr int XM_THREADS = (NCW + 1 + xm_epw<SWEEP, NCW>()) * 32;
    const int stage = XM_KC * XM_PITCH * 8 + NP * XM_MPITCH * 8;
    // Results either go through a staging tile in shared memory (128-byte global rows) or straight from the
    // accumulator fragments to global memory.  The direct form is chosen when dropping the staging tile turns a
    // 2-stage pipeline into a 3-stage one (C2: 0.96 -> 0.875 ms: HBM and the FP64 pipe are balanced there and
    // the deeper pipeline smooths their bursts); with 3+ stages anyway it gains nothing and its stores are
    // slightly less efficient.  JCB_XM_DIRECT=0/1 forces one form.
    static int direct_env = -2;
    if (direct_env == -2) {
        const char* e = getenv("JCB_XM_DIRECT");
        direct_env = e ? atoi(e) : -1;
    }
    const int budget = (XM_OCC == 2 ? 113 : 226) * 1024;
    const int fixed_base = prm.nchunk * XM_KC * 8 + 128 + (64 + 8) * 8 +
                           (SWEEP ? ((prm.ncol * prm.q + 1) & ~1) * 8 : 0);
    const int fixed_staged = fixed_base + NP * XM_PITCH * 8;
    const int ns_staged = std::min(4, (budget - fixed_staged) / stage), ns_direct = std::min(4, (budget - fixed_base) / stage);
    prm.direct = 0;
    if (!SWEEP) prm.direct = direct_env >= 0 ? (direct_env ? 1 : 0) : ((ns_staged < 3 && ns_direct > ns_staged) ? 1 : 0);
    const int fixed = prm.direct ? fixed_base : fixed_staged;
    int nstage = prm.direct ? ns_direct : ns_staged;
    if (nstage > 4) nstage = 4;
    if (nstage < 2) {
        if (NCW == 16) return -1000;                // wide tile does not fit: caller falls back to NCW = 8
        set_error("xmul: p=%d too large for the shared-memory budget", prm.p);
        return JCB200_EINVAL;
    }
    prm.nstage = nstage;
    const int smem = nstage * stage + fixed;
    const int64_t ntiles = (prm.m + XM_MT - 1) / XM_MT;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)XM_OCC * c->num_sms);
    if (!SWEEP && prm.direct) {
        if constexpr (!SWEEP) {
            JCB_CUDA(cudaFuncSetAttribute(xmul_kernel<NPB, NEX, SWEEP, NCW, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            xmul_kernel<NPB, NEX, SWEEP, NCW, true><<<grid, XM_THREADS, smem, c->stream>>>(prm);
        }
    } else {
        JCB_CUDA(cudaFuncSetAttribute(xmul_kernel<NPB, NEX, SWEEP, NCW, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        xmul_kernel<NPB, NEX, SWEEP, NCW, false><<<grid, XM_THREADS, smem, c->stream>>>(prm);
    }
    JCB_LAUNCH_CHECK();
    return 0;
}

// 256-row tiles (16 consumer warps) for narrow outputs without the sweep epilogue, else 128-row tiles
template <int NPB, int NEX, bool SWEEP>
static int launch_xmul_t(Ctx* c, XmulParams& prm) {
    if constexpr (!SWEEP && NPB <= 7) {
        static int wide = -1, wide_npb = 4;
        if (wide < 0) {
            const char* e = getenv("JCB_XM_WIDE");
            wide = e ? atoi(e) : 1;
            const char* e2 = getenv("JCB_XM_WIDE_NPB");      // widest output (8-column blocks) the 256-row tile takes
            wide_npb = e2 ? atoi(e2) : 4;
        }
        if (wide && NPB <= wide_npb && prm.m >= 256 * 148) {
            const int r = launch_xmul_w<NPB, NEX, SWEEP, 16>(c, prm);
            if (r != -1000) return r;
        }
    }
    return launch_xmul_w<NPB, NEX, SWEEP, 8>(c, prm);
}

template <bool SWEEP, int NEX>
static int dispatch_xmul_n(Ctx* c, XmulParams& prm, int npb) {
    switch (npb) {
        case 1: return launch_xmul_t<1, NEX, SWEEP>(c, prm);
        case 2: return launch_xmul_t<2, NEX, SWEEP>(c, prm);
        case 3: return launch_xmul_t<3, NEX, SWEEP>(c, prm);
        case 4: return launch_xmul_t<4, NEX, SWEEP>(c, prm);
        case 5: return launch_xmul_t<5, NEX, SWEEP>(c, prm);
        case 6: return launch_xmul_t<6, NEX, SWEEP>(c, prm);
        case 7: return launch_xmul_t<7, NEX, SWEEP>(c, prm);
        default: return launch_xmul_t<8, NEX, SWEEP>(c, prm);
    }
}
template <bool SWEEP>
static int dispatch_xmul(Ctx* c, XmulParams& prm, int npb, int nex) {
    if (nex == 1) return dispatch_xmul_n<SWEEP, 1>(c, prm, npb);
    if (nex == 2) return dispatch_xmul_n<SWEEP, 2>(c, prm, npb);
    return dispatch_xmul_n<SWEEP, 0>(c, prm, npb);
}

// workspace layout: zeros[128] | mu_pad | bias_pad[64] | Cy | Mt
static int xmul_common(Ctx* c, const double* dX, int64_t ldx, int64_t m, int64_t p, const double* dmu,
                       const double* dsigma, const double* dM, int64_t ldm, int ncol_total,
                       const double* dbias, double* dOut, int64_t ldo, bool sweep, const double* dC,
                       const double* dys, const double* dymeans, int q, int k_lo, int k_hi,
                       double* dPred) {
    const int nchunk = (int)((p + XM_KC - 1) / XM_KC);
    const int maxcol = XM_MAXNB * 8;
    if (sweep && ncol_total > maxcol) {
        set_error("predict sweep supports at most %d latent variables in one pass", maxcol);
        return JCB200_EINVAL;
    }
    const size_t ws_doubles = 128 + (size_t)nchunk * XM_KC + 64 + (size_t)maxcol * (q > 0 ? q : 1) +
                              (size_t)nchunk * (maxcol + 2) * XM_MPITCH;
    JCB_TRY(ensure(c->xmul_ws, ws_doubles * 8));
    double* zeros = (double*)c->xmul_ws.p;
    double* mu_pad = zeros + 128;
    double* bias_pad = mu_pad + (size_t)nchunk * XM_KC;
    double* Cy = bias_pad + 64;
    double* Mt = Cy + (size_t)maxcol * (q > 0 ? q : 1);
    const int aligned = (((uintptr_t)dX & 15) == 0 && (ldx & 1) == 0) ? 1 : 0;

    for (int c0 = 0; c0 < ncol_total; c0 += maxcol) {
        const int ncol = std::min(maxcol, ncol_total - c0);
        // full 8-column blocks by DMMA; 1-2 leftover columns by DFMA instead of a padded block
        static int nex_max = -1;
        if (nex_max < 0) {
            const char* e = getenv("JCB_XM_NEX_MAX");        // leftover columns taken by DFMA (else a padded DMMA block)
            nex_max = e ? atoi(e) : 2;
        }
        int npb = ncol / 8, nex = ncol % 8;
        if (npb == 0 || nex > nex_max) {
            npb = (ncol + 7) / 8;
            nex = 0;
        }
        const int NP = npb * 8 + nex;
        xmul_pack_kernel<<<64, 256, 0, c->stream>>>(dM + (int64_t)c0 * ldm, ldm, dsigma, dmu, (int)p, ncol,
                                                    NP, nchunk, Mt, mu_pad, zeros);
        JCB_LAUNCH_CHECK();
        XmulParams prm;
        memset(&prm, 0, sizeof(prm));
        prm.X = dX;
        prm.ldx = ldx;
        prm.m = m;
        prm.p = (int)p;
        prm.nchunk = nchunk;
        prm.Mt = Mt;
        prm.mu = mu_pad;
        prm.zeros = zeros;
        prm.bias = dbias ? dbias + c0 : nullptr;
        prm.Out = dOut ? dOut + (int64_t)c0 * ldo : nullptr;
        prm.ldo = ldo;
        prm.ncol = ncol;
        prm.aligned = aligned;
        prm.out_aligned = (prm.Out && (((uintptr_t)prm.Out & 15) == 0) && (ldo & 1) == 0) ? 1 : 0;
        prm.cflag = sweep ? nullptr : c->xmul_center_flag;
        if (sweep) {
            sweep_cy_kernel<<<(ncol * q + 255) / 256, 256, 0, c->stream>>>(dC, dys, q, ncol, Cy);
            JCB_LAUNCH_CHECK();
            prm.Cy = Cy;
            prm.ymeans = dymeans;
            prm.Pred = dPred;
            prm.q = q;
            prm.k_lo = k_lo;
            prm.k_hi = k_hi;
            JCB_TRY(dispatch_xmul<true>(c, prm, npb, nex));
        } else {
            JCB_TRY(dispatch_xmul<false>(c, prm, npb, nex));
        }
    }
    return 0;
}

int launch_xmul(Ctx* c, const double* dX, int64_t ldx, int64_t m, int64_t p, const double* dmu,
                const double* dsigma, const double* dM, int64_t ldm, int ncol, const double* dbias,
                double* dOut, int64_t ldo) {
    if (ncol <= 0 || m <= 0) return 0;
    phase_begin(c, JCB200_T_SCORES);
    int r = xmul_common(c, dX, ldx, m, p, dmu, dsigma, dM, ldm, ncol, dbias, dOut, ldo, false, nullptr,
                        nullptr, nullptr, 0, 0, 0, nullptr);
    phase_end(c, JCB200_T_SCORES);
    return r;
}

// pred_k for k < k_lo..k_hi all need only the first k_hi score columns
__global__ void fill_ymeans_kernel(const double* __restrict__ ymeans, int64_t m, int q, int nk,
                                   double* __restrict__ Pred) {
    const int64_t total = (int64_t)nk * m * q;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x)
        Pred[e] = ymeans[(e / m) % q];
}

int launch_predict_sweep(Ctx* c, const double* dX, int64_t ldx, int64_t m, int64_t p, int64_t q,
                         const double* dR, const double* dC, int a, const double* dxmeans,
                         const double* dxscales, const double* dymeans, const double* dyscales,
                         int k_lo, int k_hi, double* dPred) {
    if (m <= 0) return 0;
    phase_begin(c, JCB200_T_SCORES);
    int r = 0;
    if (k_hi == 0) {
        // nlv = 0 only: predictions are ymeans broadcast (plskern.jl:210-215,234)
        fill_ymeans_kernel<<<256, 256, 0, c->stream>>>(dymeans, m, (int)q, k_hi - k_lo + 1, dPred);
        g_launches++;
        if (cudaGetLastError() != cudaSuccess) r = 1;
    } else if (k_lo == k_hi) {
        // single k: the reference's own arithmetic, pred = int + X B (plskern.jl:233-234), as
        // ymeans + (X - xmeans) B with B = coef(k): q output columns instead of k score columns — HBM-bound
        r = ensure(c->coef_ws, (size_t)(p * q + q) * 8);
        double* dB = (double*)c->coef_ws.p;
        if (r == 0) r = launch_coef(c, dR, dC, dxmeans, dxscales, dymeans, dyscales, p, q, k_hi, dB, dB + p * q);
        if (r == 0)
            r = xmul_common(c, dX, ldx, m, p, dxmeans, nullptr, dB, p, (int)q, dymeans, dPred, m, false, nullptr,
                            nullptr, nullptr, 0, 0, 0, nullptr);
    } else {
        r = xmul_common(c, dX, ldx, m, p, dxmeans, dxscales, dR, p, k_hi, nullptr, nullptr, 0, true, dC,
                        dyscales, dymeans, (int)q, k_lo, k_hi, dPred);
    }
    phase_end(c, JCB200_T_SCORES);
    return r;
}

}  // namespace jcb
