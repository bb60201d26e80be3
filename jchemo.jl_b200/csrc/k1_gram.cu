// K1 — fused weight / centre + FP64 tensor-core SYRK/GEMM:  G = [X-c Y-c]' D [X-c Y-c]
// (upper 32x32 units only), weighted column sums and sum(w), for one row shard.
//
// Replaces, in Gram form, the reference's preamble and per-LV GEMVs:
//   mweight/colmean/center!/cscale! + X'(D Y)      /root/reference/src/plskern.jl:117-132
//   mul!(t, X, r), mul!(zp, X', dt) per LV          /root/reference/src/plskern.jl:162,167
//
// Design (B200, sm_100a):
//  * X, Y are column-major (rows contiguous) so the contraction dimension K = rows is contiguous for
//    both operands.  The augmented column space is cut into 32-column blocks (X blocks, then Y blocks);
//    a 32x32 output "unit" (block_a, block_b) is owned by one warp: 16 DMMA.8x8x4 accumulators.
//  * A CTA (16 warps; thread 0 doubles as the TMA producer and polls the ring between k8-steps) works on a
//    "group": two ranges of up to 4 consecutive column blocks staged per pipeline stage and up to 16 units on them
//    (an off-diagonal 128x128 super-tile, or a diagonal super-tile + the X'Y / Y'Y units).  Diagonal units skip the
//    8x8 blocks below the diagonal.
//  * Rows are streamed in stages of KT = 56 rows: two 2-D TMA loads per stage (FP64 tensor map, box 56 rows x 128
//    columns, zero fill out of bounds) into a [column][KT] shared-memory tile; KT = 8 (mod 16) makes the 128-bit
//    fragment loads bank-conflict free.  Full/empty mbarriers form a 2-stage ring (2 x 114 KB).
//  * Centring (x - c, c = strided-sample pivot) and weighting happen on the fragments in registers;
//    the exact correction G - delta delta' is applied in K3.  Zero-filled tail rows get weight 0.
//  * Weighted stream-K: the (group, stage) space is cut into equal-cost contiguous pieces, one per SM,
//    so every SM is busy for the same time; each piece writes its partial units to a workspace and
//    K1b sums them in a fixed order (deterministic, no atomics).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "jcb_internal.cuh"

namespace jcb {

#ifndef JCB_KT
#define JCB_KT 56
#endif
#ifndef JCB_K8_BAR
#define JCB_K8_BAR 0   /* per-SMSP pacing barrier at every k8-step: measured neutral (8.39 vs 8.40 ms), left off */
#endif
#ifndef JCB_NSTAGE
#define JCB_NSTAGE 2
#endif
#ifndef JCB_POLL_MOD
#define JCB_POLL_MOD 4
#define JCB_POLL_REM 1
#endif
constexpr int KT = JCB_KT;             // rows per pipeline stage (KT % 16 == 8: conflict-free LDS.128)
constexpr int CB = 32;                 // columns per block
constexpr int MAXSLOT = 8;             // column blocks staged per stage
constexpr int NSTAGE = JCB_NSTAGE;     // pipeline depth
constexpr int NCW = 16;                // consumer warps per CTA
constexpr int K1_THREADS = NCW * 32;    // thread 0 doubles as the TMA producer (polls between k8-steps)
constexpr int UNIT_STRIDE = 1088;      // doubles per partial unit: 1024 G + 32 column sums + sum(w) + pad
constexpr int SLOT_DOUBLES = CB * KT;
constexpr int STAGE_BYTES = MAXSLOT * SLOT_DOUBLES * 8 + 512;  // data + weight tile (KT doubles, padded)
constexpr int K1_SMEM = NSTAGE * STAGE_BYTES + 64;                     // + mbarriers
static_assert(KT % 16 == 8, "KT must be 8 mod 16 (bank-conflict-free fragment loads)");
static_assert(K1_SMEM <= 227 * 1024, "K1 shared memory budget");

struct UnitDesc {
    int8_t sa, sb;      // slots of the A (rows of G) and B (columns of G) blocks
    int8_t kind;        // 0 idle, 1 full, 2 diagonal (skip 8x8 blocks with mb > nb)
    int8_t mbc, nbc;    // active 8-blocks in A / B (edge blocks are narrower)
    int8_t sums;        // bit0: accumulate column sums of the B block, bit1: accumulate sum(w)
    int8_t pad[2];
};
struct GroupDesc {
    int32_t nslots;           // 4 * nranges: range r = blocks blk[4r .. 4r+3] (consecutive; -1 = unused)
    int32_t nranges;
    int32_t blk[MAXSLOT];
    UnitDesc unit[NCW];
};
// A segment = the share of one group that one CTA accumulates into one partial.  Rows are swept zone by
// zone (all CTAs are in the same zone at the same time, so the 4 groups that read a column block find
// it in L2): inside every zone the segment owns the stage offsets [x0, x1) (16.16 fixed point,
// dithered per zone so that fractional boundaries average out exactly).
struct SegDesc {
    int32_t group, pad;
    int64_t x0, x1;
};

struct StageIter {
    int64_t x0, x1, nst;
    int32_t L, nz, z, cur, end;
    __device__ __forceinline__ void init(const SegDesc& sd, int64_t nst_, int32_t L_, int32_t nz_) {
        x0 = sd.x0; x1 = sd.x1; nst = nst_; L = L_; nz = nz_; z = -1; cur = end = 0;
    }
    __device__ __forceinline__ bool next(int& st) {
        while (cur >= end) {
            if (++z >= nz) return false;
            const int64_t r = ((int64_t)z * 40503) & 0xFFFF;
            const int64_t base = (int64_t)z * L;
            const int64_t lz = min((int64_t)L, nst - base);
            const int64_t b0 = min((x0 + r) >> 16, lz), b1 = min((x1 + r) >> 16, lz);
            cur = (int32_t)(base + b0);
            end = (int32_t)(base + b1);
        }
        st = cur++;
        return true;
    }
};

struct GramParams {
    const GroupDesc* groups;
    const SegDesc* segs;
    const int32_t* cta_seg;   // [ncta + 1]
    const double* pivot;      // p + q
    double* partials;         // nsegs * NCW * UNIT_STRIDE
    int64_t n, nst;           // rows, stages
    int32_t p, q, nbx;
    int32_t weighted;
    int32_t zone_len, nzones; // stages per zone, number of zones
#ifdef JCB_K1_TRACE
    long long* trace;         // debug builds: [cta < 4][stage < 64][warp][3] clock64 stamps
#endif
};

// ------------------------------------------------------------------------------------------ kernel
// ---------------------------------------------------------------------------------------------
// Inner loop.  One k8-step of one 32x32 unit: 8 rows of the staged tile, two DMMA k4-halves.
//  * The A operand (rows of G) is used RAW; only the B operand is centred (and weighted):
//      acc_ij = sum_k x_ki * w_k (x_kj - c_j)  =  G_ij + c_i s_j ,  s_j = sum_k w_k (x_kj - c_j)
//    K3 subtracts the rank-one term c_i s_j exactly (s is accumulated by the diagonal units); with the
//    4K-row pivot s_j is tiny, so nothing is lost to cancellation.  This halves the FP64 adds in front of the
//    DMMAs (a DADD costs the pipe nothing beside a DMMA stream — bench/fp64_mix.cu — but it is one more link in
//    the issuing warp's LDS -> DADD -> DMMA chain).
//  * DIAG units skip the 8x8 blocks below the diagonal and accumulate the column sums; NBC < 4 (edge
//    and Y blocks) skips empty column blocks; MASKED is the zero-filled tail stage of the unweighted
//    kernel.  All three are COMPILE-TIME: a predicated-off DMMA/DMUL still occupies the FP64 pipe on
//    sm_100a (measured, profiles/k1_r01_notes.md), so skipped work must not be emitted at all.
// All k8-steps of one stage of one 32x32 unit.  Per k8-step: A fragments (raw), then the B fragments
// NG column blocks at a time (centred, weighted; diagonal units also accumulate the column sums).
// Within a group all first k-halves are issued before any second half, so the two dependent DMMAs of
// an accumulator are >= 8 (full) or >= 10 (diagonal) issue slots apart.
// (A register-level software pipeline of the B transform was tried and was slower: 128 registers.)
template <bool WEIGHTED, bool DIAG, int NBC, bool CENTER, bool MASKED, typename Poll>
__device__ __forceinline__ void stage_steps(const double* __restrict__ tA, const double* __restrict__ tB,
                                            const double* __restrict__ wt, const double (&pB)[4],
                                            double (&acc)[4][4][2], double (&bsum)[4], double& wsum,
                                            const bool sum_w, const int krow0, const int rows_valid,
                                            const int bar_id, const int bar_threads, Poll& poll) {
    constexpr int MBC = DIAG ? NBC : 4;
    constexpr int NG = DIAG ? NBC : (NBC >= 2 ? 2 : 1);
#pragma unroll 1
    for (int k8 = 0; k8 < KT / 8; ++k8) {
#if JCB_K8_BAR
        // The warp scheduler favours some warps of an SMSP; left alone they run a full stage ahead, block on
        // the ring, and leave too few warps to hide the LDS -> DADD -> DMMA latency of the others (measured:
        // 11 % of the FP64 pipe idle).  A named barrier among the SMSP's active warps at every k8-step keeps
        // them within one step of each other.
        if (bar_threads > 32) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
#endif
        // thread 0 looks for a released ring slot at k8-steps 1 and 5 of the 7.  The poll sits on warp 0's critical
        // path: every k8-step 0.871 of the FP64 peak at C2, every second 0.878, steps 1 and 5 0.879 (C3 0.878 ->
        // 0.889, C4 shard 0.886 -> 0.897), once per stage 0.845-0.863, never (blocking only) 0.79; twice per k8-step,
        // or issuing from the warp that completes the release, makes the 16 warps run in lockstep and lose more at
        // the stage boundaries than the earlier refill gains (0.835).
        if ((k8 % JCB_POLL_MOD) == JCB_POLL_REM) poll();
        double2 a[MBC];
#pragma unroll
        for (int mb = 0; mb < MBC; ++mb)
            a[mb] = *reinterpret_cast<const double2*>(tA + k8 * 8 + mb * 8 * KT);
        double2 w2 = make_double2(1.0, 1.0);
        if (WEIGHTED) {
            w2 = *reinterpret_cast<const double2*>(wt + k8 * 8);
        } else if (MASKED) {
            w2.x = (krow0 + k8 * 8 < rows_valid) ? 1.0 : 0.0;
            w2.y = (krow0 + k8 * 8 + 1 < rows_valid) ? 1.0 : 0.0;
        }
        if (DIAG && sum_w) wsum += w2.x + w2.y;
#pragma unroll
        for (int n0 = 0; n0 < NBC; n0 += NG) {
            double2 b[NG];
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const int nb = n0 + j;
                if (nb < NBC) {
                    b[j] = *reinterpret_cast<const double2*>(tB + k8 * 8 + nb * 8 * KT);
                    if (CENTER) {
                        b[j].x -= pB[nb];
                        b[j].y -= pB[nb];
                    }
                    if (WEIGHTED || MASKED) {
                        b[j].x *= w2.x;
                        b[j].y *= w2.y;
                    }
                    if (DIAG) {
                        // column sums ride in the accumulator blocks below the diagonal, which a
                        // diagonal unit never uses: two independent DADD chains per column block
                        // (a dependent DADD pair would stall this warp behind the other warps' DMMAs)
                        double(&sx)[2] = nb < 3 ? acc[3][nb] : acc[2][0];
                        sx[0] += b[j].x;
                        sx[1] += b[j].y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const int nb = n0 + j;
                if (nb < NBC) {
#pragma unroll
                    for (int mb = 0; mb < MBC; ++mb)
                        if (!DIAG || mb <= nb) dmma(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[j].x);
                }
            }
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const int nb = n0 + j;
                if (nb < NBC) {
#pragma unroll
                    for (int mb = 0; mb < MBC; ++mb)
                        if (!DIAG || mb <= nb) dmma(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[j].y);
                }
            }
        }
    }
}

// Request side (thread 0 only): the cursor walks this CTA's (segment, stage) list; a request needs its
// ring slot released by all 16 warps, which thread 0 polls for (never blocking) between k8-steps.
struct Producer {
    StageIter iter;
    int sg, seg_end;
    int next;         // stage of the pending request, -1 when exhausted
    uint32_t issued;  // requests issued so far (ring position)
    // cached description of the current segment's group: up to 2 ranges of <= 4 consecutive blocks,
    // range r always lands in ring slots [4r, 4r + 4) with ONE 128-column TMA box
    int nranges, col[2], isy[2];
};

__device__ __forceinline__ void producer_load_group(Producer& P, const GramParams& prm) {
    const GroupDesc* gd = &prm.groups[prm.segs[P.sg].group];
    P.nranges = gd->nranges;
    for (int r = 0; r < 2; ++r) {
        const int b = gd->blk[4 * r];
        P.isy[r] = b >= prm.nbx;
        P.col[r] = (P.isy[r] ? b - prm.nbx : b) * CB;
    }
}

__device__ __forceinline__ void producer_advance(Producer& P, const GramParams& prm) {
    P.next = -1;
    while (P.sg < P.seg_end) {
        int st;
        if (P.iter.next(st)) {
            P.next = st;
            return;
        }
        if (++P.sg < P.seg_end) {
            P.iter.init(prm.segs[P.sg], prm.nst, prm.zone_len, prm.nzones);
            producer_load_group(P, prm);
        }
    }
}

template <bool WEIGHTED>
__device__ __noinline__ void producer_issue(Producer& P, const GramParams& prm,
                                            const CUtensorMap* mapX, const CUtensorMap* mapY,
                                            const CUtensorMap* mapW, unsigned char* smem,
                                            uint64_t* full) {
    const int buf = P.issued % NSTAGE;
#ifdef JCB_K1_TRACE
    if (prm.trace && (blockIdx.x % 45) == 0 && P.issued < 64)
        prm.trace[4 * 64 * NCW * 3 + (blockIdx.x / 45) * 64 + P.issued] = clock64();
#endif
    const uint32_t bytes = P.nranges * 4 * SLOT_DOUBLES * 8 + (WEIGHTED ? KT * 8 : 0);
    mbar_arrive_expect_tx(&full[buf], bytes);
    unsigned char* base = smem + buf * STAGE_BYTES;
    const int row0 = P.next * KT;
    tma_load_2d(base, P.isy[0] ? mapY : mapX, row0, P.col[0], &full[buf]);
    if (P.nranges > 1)
        tma_load_2d(base + 4 * SLOT_DOUBLES * 8, P.isy[1] ? mapY : mapX, row0, P.col[1], &full[buf]);
    if (WEIGHTED) tma_load_1d(base + MAXSLOT * SLOT_DOUBLES * 8, mapW, row0, &full[buf]);
    ++P.issued;
    producer_advance(P, prm);
}

struct K1Shared {
    unsigned char* smem;
    uint64_t* full;
    uint64_t* empty;
};

// All stages of one segment for one warp-unit variant: a tight loop with the variant dispatch hoisted
// out (the per-stage bookkeeping is on the critical path: at a stage boundary the four warps of an
// SMSP leave the DMMA pipe idle together).
template <bool WEIGHTED, bool DIAG, int NBC, bool CENTER>
__device__ __forceinline__ void run_segment(const SegDesc& seg, const UnitDesc u, const GramParams& prm,
                                            const CUtensorMap* mapX, const CUtensorMap* mapY,
                                            const CUtensorMap* mapW, const K1Shared sh, Producer& P,
                                            const bool producer, uint32_t& it, const int g, const int kk,
                                            const double (&pB)[4], double (&acc)[4][4][2],
                                            double (&bsum)[4], double& wsum, const int bar_id,
                                            const int bar_threads) {
    const bool sum_w = (u.sums & 2) != 0;
    const int lane = threadIdx.x & 31;
    const int offA = u.sa * SLOT_DOUBLES + g * KT + 2 * kk;
    const int offB = u.sb * SLOT_DOUBLES + g * KT + 2 * kk;
    StageIter c_iter;
    c_iter.init(seg, prm.nst, prm.zone_len, prm.nzones);
    int st;
    for (; c_iter.next(st); ++it) {
        const int buf = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        if (producer) {   // the stage about to be consumed must have been requested
            while (P.issued <= it && P.next >= 0) {
                mbar_wait(&sh.empty[P.issued % NSTAGE], ((P.issued / NSTAGE) & 1) ^ 1);
                producer_issue<WEIGHTED>(P, prm, mapX, mapY, mapW, sh.smem, sh.full);
            }
        }
        __syncwarp();
#ifdef JCB_K1_TRACE
        const bool tr = prm.trace && (blockIdx.x % 45) == 0 && it < 64 && lane == 0;
        long long* trp = prm.trace + (((int64_t)(blockIdx.x / 45) * 64 + it) * NCW + (threadIdx.x >> 5)) * 3;
        if (tr) trp[0] = clock64();
#endif
        mbar_wait(&sh.full[buf], ph);
#ifdef JCB_K1_TRACE
        if (tr) trp[1] = clock64();
#endif
        const double* base = reinterpret_cast<const double*>(sh.smem + buf * STAGE_BYTES);
        const double* tA = base + offA;
        const double* tB = base + offB;
        const double* wt = base + MAXSLOT * SLOT_DOUBLES + 2 * kk;
        const int64_t rows_left = prm.n - (int64_t)st * KT;
        auto poll = [&]() {
            if (producer && P.next >= 0 && P.issued < it + NSTAGE &&
                mbar_test_wait(&sh.empty[P.issued % NSTAGE], ((P.issued / NSTAGE) & 1) ^ 1))
                producer_issue<WEIGHTED>(P, prm, mapX, mapY, mapW, sh.smem, sh.full);
        };
        // tail stage of the unweighted kernel: zero-filled rows must get weight 0 when they are centred
        // (0 - c != 0) or counted (sum of weights); without centring they contribute nothing anyway
        if (!WEIGHTED && rows_left < KT && (CENTER || (DIAG && sum_w)))
            stage_steps<WEIGHTED, DIAG, NBC, CENTER, true>(tA, tB, wt, pB, acc, bsum, wsum, sum_w, 2 * kk,
                                                           (int)rows_left, bar_id, bar_threads, poll);
        else
            stage_steps<WEIGHTED, DIAG, NBC, CENTER, false>(tA, tB, wt, pB, acc, bsum, wsum, sum_w,
                                                            2 * kk, KT, bar_id, bar_threads, poll);
        __syncwarp();
#ifdef JCB_K1_TRACE
        if (tr) trp[2] = clock64();
#endif
        if (lane == 0) mbar_arrive(&sh.empty[buf]);
    }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(K1_THREADS, 1)
gram_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
            const __grid_constant__ CUtensorMap mapW, const GramParams prm) {
    extern __shared__ __align__(1024) unsigned char smem[];
    K1Shared sh;
    sh.smem = smem;
    sh.full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    sh.empty = sh.full + NSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], NCW);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const int seg_begin = prm.cta_seg[blockIdx.x], seg_end = prm.cta_seg[blockIdx.x + 1];
    const bool producer = (threadIdx.x == 0);

    Producer P;
    P.sg = seg_begin;
    P.seg_end = seg_end;
    P.next = -1;
    P.issued = 0;
    if (producer && seg_begin < seg_end) {
        P.iter.init(prm.segs[P.sg], prm.nst, prm.zone_len, prm.nzones);
        producer_load_group(P, prm);
        producer_advance(P, prm);
        for (int s = 0; s < NSTAGE && P.next >= 0; ++s)
            producer_issue<WEIGHTED>(P, prm, &mapX, &mapY, &mapW, smem, sh.full);
    }

    const int g = lane >> 2, kk = lane & 3;
    // pivot[p + q] != 0: the columns are centred about the pivot inside the loop; == 0: the pivot is all
    // zeros (well-scaled columns, decided by the pivot kernel) and the loop is pure LDS + DMMA
    const bool center = prm.pivot[prm.p + prm.q] != 0.0;
    uint32_t it = 0;  // stages consumed by this CTA so far (ring position)
    for (int sg = seg_begin; sg < seg_end; ++sg) {
        const SegDesc seg = prm.segs[sg];
        const GroupDesc* gd = &prm.groups[seg.group];
        const UnitDesc u = gd->unit[warp];
#if JCB_K8_BAR
        __syncthreads();   // the pacing barriers of two segments (different thread counts) must not overlap
#endif
        // active warps on this warp's SMSP (warp % 4) in this group: they pace each other per k8-step
        int nact = 0;
#pragma unroll
        for (int j = 0; j < NCW / 4; ++j) nact += gd->unit[(warp & 3) + 4 * j].kind != 0;
        const int bar_id = 1 + (warp & 3), bar_threads = nact * 32;
        double acc[4][4][2];
        double bsum[4], pB[4];
        double wsum = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bsum[i] = 0.0;
            pB[i] = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        }
        if (u.kind) {
            // pivots of this lane's B columns (0 for padding columns: TMA zero-fills them)
            const int bb = gd->blk[u.sb];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (bb >= prm.nbx) {
                    const int col = (bb - prm.nbx) * CB + i * 8 + g;
                    if (col < prm.q) pB[i] = prm.pivot[prm.p + col];
                } else {
                    const int col = bb * CB + i * 8 + g;
                    if (col < prm.p) pB[i] = prm.pivot[col];
                }
            }
        }
#define JCB_SEG(D, N)                                                                                \
    do {                                                                                             \
        if (center)                                                                                  \
            run_segment<WEIGHTED, D, N, true>(seg, u, prm, &mapX, &mapY, &mapW, sh, P, producer, it, g, \
                                              kk, pB, acc, bsum, wsum, bar_id, bar_threads);         \
        else                                                                                         \
            run_segment<WEIGHTED, D, N, false>(seg, u, prm, &mapX, &mapY, &mapW, sh, P, producer, it, \
                                               g, kk, pB, acc, bsum, wsum, bar_id, bar_threads);     \
    } while (0)
        switch (u.kind ? ((u.kind == 2 ? 4 : 0) + (u.nbc - 1)) : -1) {
            case 0: JCB_SEG(false, 1); break;
            case 1: JCB_SEG(false, 2); break;
            case 2: JCB_SEG(false, 3); break;
            case 3: JCB_SEG(false, 4); break;
            case 4: JCB_SEG(true, 1); break;
            case 5: JCB_SEG(true, 2); break;
            case 6: JCB_SEG(true, 3); break;
            case 7: JCB_SEG(true, 4); break;
            default: {   // idle warp of a small group: keep the ring protocol going
                StageIter c_iter;
                c_iter.init(seg, prm.nst, prm.zone_len, prm.nzones);
                int st;
                for (; c_iter.next(st); ++it) {
                    const int buf = it % NSTAGE;
                    if (producer) {
                        while (P.issued <= it && P.next >= 0) {
                            mbar_wait(&sh.empty[P.issued % NSTAGE], ((P.issued / NSTAGE) & 1) ^ 1);
                            producer_issue<WEIGHTED>(P, prm, &mapX, &mapY, &mapW, smem, sh.full);
                        }
                    }
                    __syncwarp();
                    mbar_wait(&sh.full[buf], (it / NSTAGE) & 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.empty[buf]);
                }
            } break;
        }
#undef JCB_SEG
        // ---- write this warp's partial unit (fragment order; K1b knows the layout)
        if (u.kind) {
            double* out = prm.partials + ((int64_t)sg * NCW + warp) * UNIT_STRIDE;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb)
                    *reinterpret_cast<double2*>(out + ((mb * 4 + nb) * 32 + lane) * 2) =
                        make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            if (u.kind == 2) {
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    double v = nb < 3 ? acc[3][nb][0] + acc[3][nb][1] : acc[2][0][0] + acc[2][0][1];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (kk == 0) out[1024 + nb * 8 + g] = v;
                }
                double v = wsum;
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (lane == 0) out[1056] = v;
            }
        }
    }
}

// K1b — ordered split-K reduce of the partial units into the packed layout
// [Gxx p*p | Gxy p*q | gyy q | sx p | sy q | sw].  One block per (group, warp-unit): the 32x32 unit is summed over
// its segments in fragment order (fixed order: deterministic), transposed through shared memory and written as
// 32 columns of 32 consecutive doubles (256-byte runs) to EVERY destination — the fit's own packed buffer, or, in
// the row-sharded fit, slot [rank] of every rank's peer window over NVLink: the exchange needs no copy of its own.
// The last block to finish (system fence + completion counter) raises this rank's flag in every window.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const GroupDesc* __restrict__ groups, const int32_t* __restrict__ group_seg,
                   const double* __restrict__ partials, const ReduceDst dst, int p, int q,
                   int nbx, int accumulate, int joint) {
    __shared__ double tile[32 * 33];
    const int gi = blockIdx.x / NCW, wu = blockIdx.x % NCW;
    const GroupDesc* gd = &groups[gi];
    const UnitDesc u = gd->unit[wu];
    if (u.kind) {
        const int sbeg = group_seg[gi], send = group_seg[gi + 1];
        const int ba = gd->blk[u.sa], bb = gd->blk[u.sb];
        const bool ay = ba >= nbx, by = bb >= nbx;
        const int rowbase = (ay ? ba - nbx : ba) * CB, colbase = (by ? bb - nbx : bb) * CB;
        const int64_t P = p, Q = q;
        const int64_t o_gxy = P * P, o_gyy = o_gxy + P * Q, o_sx = o_gyy + Q, o_sy = o_sx + P, o_sw = o_sy + Q;
        const int nelem = 1024 + (u.kind == 2 ? 32 : 0);
        for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
            // fixed summation order (deterministic), but eight loads in flight: one dependent load per segment made
            // this kernel a chain of memory latencies (38 us for 22 MB)
            double s = 0.0;
            const double* src = partials + (int64_t)wu * UNIT_STRIDE + e;
            int sg = sbeg;
            for (; sg + 8 <= send; sg += 8) {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldcs(src + (int64_t)(sg + k) * NCW * UNIT_STRIDE);
#pragma unroll
                for (int k = 0; k < 8; ++k) s += v[k];
            }
            for (; sg < send; ++sg) s += __ldcs(src + (int64_t)sg * NCW * UNIT_STRIDE);
            if (e < 1024) {
                const int blk = e >> 6, ln = (e >> 1) & 31, half = e & 1;
                const int mb = blk >> 2, nb = blk & 3;
                tile[(nb * 8 + (ln & 3) * 2 + half) * 33 + mb * 8 + (ln >> 2)] = s;
            } else {
                const int col = colbase + (e - 1024);
                int64_t off = -1;
                if (joint) {                 // one column space [X | Y]: the block may straddle the two
                    if (col < p) off = o_sx + col;
                    else if (col < p + q) off = o_sy + (col - p);
                } else if (!by) {
                    if (col < p) off = o_sx + col;
                } else {
                    if (col < q) off = o_sy + col;
                }
                if (off >= 0) {
                    if (accumulate) s += dst.p[0][off];
#pragma unroll
                    for (int d = 0; d < 8; ++d)
                        if (d < dst.n) dst.p[d][off] = s;
                }
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 1024; idx += blockDim.x) {
            const int cl = idx >> 5, r = idx & 31;
            if (u.kind == 2 && (r >> 3) > (cl >> 3)) continue;      // below the diagonal: never computed
            const int row = rowbase + r, col = colbase + cl;
            int64_t off = -1;
            if (joint) {
                if (col < p) {
                    if (row < p) off = row + (int64_t)col * P;                               // X'X
                } else if (col < p + q) {
                    if (row < p) off = o_gxy + row + (int64_t)(col - p) * P;                 // X'Y
                    else if (row == col) off = o_gyy + (row - p);                            // diag Y'Y
                }
            } else if (!ay && !by) {
                if (row < p && col < p) off = row + (int64_t)col * P;
            } else if (!ay && by) {
                if (row < p && col < q) off = o_gxy + row + (int64_t)col * P;
            } else if (ay && by) {
                if (row == col && row < q) off = o_gyy + row;
            }
            if (off >= 0) {
                double v = tile[cl * 33 + r];
                if (accumulate) v += dst.p[0][off];
#pragma unroll
                for (int d = 0; d < 8; ++d)
                    if (d < dst.n) dst.p[d][off] = v;
            }
        }
        if ((u.sums & 2) && threadIdx.x == 0) {
            double s = 0.0;
            for (int sg = sbeg; sg < send; ++sg)
                s += partials[((int64_t)sg * NCW + wu) * UNIT_STRIDE + 1056];
            if (accumulate) s += dst.p[0][o_sw];
            for (int d = 0; d < dst.n; ++d) dst.p[d][o_sw] = s;
        }
    }
    if (dst.done) {
        __threadfence_system();          // this thread's (remote) stores are performed before what follows
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(dst.done, 1u);
            if (prev == gridDim.x - 1) {          // every block has fenced its stores
                *dst.done = 0;
                __threadfence_system();
                for (int d = 0; d < dst.n; ++d)
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst.flag[d]), "l"(dst.seq) : "memory");
            }
        }
    }
}

// Strided-sample pivot: mean of up to 4096 rows per column, taken as 16 evenly spaced chunks of 256 consecutive
// rows (coalesced 2 KB reads), all 16 loads of a thread in flight at once: the kernel is a chain of memory / TLB
// latencies, not of bytes (16 K rows in 8 dependent rounds took 15 us; one round takes ~4).  One block per column.
// The pivot only has to be CLOSE to the mean (K3 corrects exactly): with 4096 rows the correction terms c_i s_j
// and delta delta' stay ~sigma/64 relative, far below the rounding level even for offset-heavy data.
// ratio[col] = mean^2 / variance of the sample tells how much centring matters for this column.
//
// The LAST block to finish takes the data-dependent centring decision (it used to be a kernel of its own):
// centring puts a DADD between every fragment load and its DMMAs; when every column has mean^2 <= 64 variance, second
// moments about 0 lose at most ~2 digits to the mean (K3 removes it exactly: pivot = 0 is just another pivot),
// far inside the 1e-10 budget.  pivot[p + q] = 1 (the data need centring), 0 (they do not, and the pivot was
// zeroed: K1 runs without centring; only with JCB_AUTO_NOCENTER=1, because K1 itself gains nothing from it) or
// 2 (they do not, but K1 keeps its pivot: the default).  K1 centres whenever the flag is non-zero; the score
// pass K5, which is 2-3 % faster without the DADDs in front of its DMMAs, goes centre-free on 0 and 2 (k5_xmul.cu).
__global__ void __launch_bounds__(256)
pivot_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y, int64_t ldy,
             int64_t n, int p, int q, double* __restrict__ pivot, double* __restrict__ ratio,
             unsigned int* __restrict__ counter, int force_center) {
    const int col = blockIdx.x;
    const double* src = col < p ? X + (int64_t)col * ldx : Y + (int64_t)(col - p) * ldy;
    __shared__ double red[16];
    __shared__ int s_last, s_need;
    double s = 0.0, s2 = 0.0;
    int64_t cnt;
    if (n <= 4096) {
        cnt = n;
        for (int64_t i = threadIdx.x; i < n; i += 256) {
            const double v = src[i];
            s += v;
            s2 += v * v;
        }
    } else {
        cnt = 4096;
        const int64_t stride = n / 16;             // chunk c covers rows [c*stride, c*stride + 256)
        double v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = src[c * stride + threadIdx.x];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            s += v[c];
            s2 += v[c] * v[c];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = s;
        red[8 + (threadIdx.x >> 5)] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) {
            t += red[w];
            t2 += red[8 + w];
        }
        const double m = t / (double)cnt;
        const double var = t2 / (double)cnt - m * m;
        pivot[col] = m;
        // a (near-)constant column has no usable variance estimate: always centre
        ratio[col] = (var > 1e-12 * (m * m) && var > 0.0) ? (m * m) / var : 1e300;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
        s_need = 0;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last block: every column's ratio is visible (fence + counter)
    __threadfence();
    const int ncol = p + q;
    int need = 0;
    for (int c = threadIdx.x; c < ncol; c += 256)
        if (!(__ldcg(ratio + c) <= 64.0)) need = 1;
    if (need) s_need = 1;
    __syncthreads();
    need = s_need;
    if (!need && !force_center)
        for (int c = threadIdx.x; c < ncol; c += 256) pivot[c] = 0.0;
    if (threadIdx.x == 0) {
        pivot[ncol] = need ? 1.0 : (force_center ? 2.0 : 0.0);
        *counter = 0;                               // ready for the next launch
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int make_map_2d(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld,
                       int box_rows, int box_cols) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return JCB200_ENODEV;
    }
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(2d) failed with CUresult %d (rows=%lld cols=%lld ld=%lld)",
                  (int)r, (long long)rows, (long long)cols, (long long)ld);
        return JCB200_EINVAL;
    }
    return 0;
}

static int make_map_1d(CUtensorMap* map, const double* base, int64_t len, int box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return JCB200_ENODEV;
    }
    cuuint64_t dims[1] = {(cuuint64_t)len};
    cuuint64_t strides[1] = {0};
    cuuint32_t boxd[1] = {(cuuint32_t)box};
    cuuint32_t estr[1] = {1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, const_cast<double*>(base), dims,
                     strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(1d) failed with CUresult %d", (int)r);
        return JCB200_EINVAL;
    }
    return 0;
}

// Cost of a unit in DMMA pairs per k8-step, used to balance SMSPs and SMs.
// FP64-pipe cost of a unit per k8-step, in 32-cycle units (one DMMA pair = 1): DMMA pairs plus the
// B-fragment adds (2 DADD per column block; diagonal units add 2 more for the column sums).
static double unit_cost(const UnitDesc& u) {
    if (u.kind == 2) return u.nbc * (u.nbc + 1) / 2 + 0.25 * u.nbc;   // upper triangle of nbc x nbc blocks
    return 4 * u.nbc + 0.125 * u.nbc;                                 // 4 row blocks x nbc column blocks
}

struct Schedule {
    std::vector<GroupDesc> groups;
    std::vector<double> gcost;
    std::vector<SegDesc> segs;
    std::vector<int32_t> cta_seg;    // ncta + 1
    std::vector<int32_t> group_seg;  // ngroups + 1
};

// Fixed per-stage overhead of a group in the same units as unit_cost (fragment transforms, barrier).  Calibrated on
// B200 with the joint [X | Y] layout (bench/k1_sweep.py, profiles/k1_sweep_r02.txt): when the groups are few and the
// rows are swept in zones (C2: 10 groups, C3: 36) the optimum is 2-3 (C2 8.12 -> 8.06 ms), when there are about as
// many groups as SMs (C4: 136) it is 5 (153.8 ms; 156.8 at 3, 158.5 at 2).  JCB_STAGE_OVERHEAD overrides both.
static double stage_overhead(bool zoned) {
    static double v = -2;
    if (v < -1) {
        const char* e = getenv("JCB_STAGE_OVERHEAD");
        v = e ? atof(e) : -1.0;
    }
    return v >= 0 ? v : (zoned ? 3.0 : 5.0);
}

static void build_groups(int64_t p, int64_t q, Schedule& S) {
    const int nbx = (int)((p + CB - 1) / CB), nby = (int)((q + CB - 1) / CB);
    auto blk_cols = [&](int b) {  // real columns in block b
        if (b < nbx) return (int)std::min<int64_t>(CB, p - (int64_t)b * CB);
        return (int)std::min<int64_t>(CB, q - (int64_t)(b - nbx) * CB);
    };
    auto nb8 = [&](int b) { return (blk_cols(b) + 7) / 8; };
    const int ns = (nbx + 3) / 4;
    struct RawUnit { int ba, bb, kind, sums; };
    // rA, rB: up to 4 consecutive blocks each; range r occupies ring slots [4r, 4r + 4)
    auto emit = [&](const std::vector<int>& rA, const std::vector<int>& rB, std::vector<RawUnit>& units) {
        auto slot_of = [&](int blk) {
            for (size_t i = 0; i < rA.size(); ++i)
                if (rA[i] == blk) return (int)i;
            for (size_t i = 0; i < rB.size(); ++i)
                if (rB[i] == blk) return 4 + (int)i;
            return -1;
        };
        // split into groups of at most NCW units that share the slot list
        for (size_t u0 = 0; u0 < units.size(); u0 += NCW) {
            GroupDesc gd;
            memset(&gd, 0, sizeof(gd));
            gd.nranges = rB.empty() ? 1 : 2;
            gd.nslots = 4 * gd.nranges;
            for (int i = 0; i < MAXSLOT; ++i) gd.blk[i] = -1;
            for (size_t i = 0; i < rA.size(); ++i) gd.blk[i] = rA[i];
            for (size_t i = 0; i < rB.size(); ++i) gd.blk[4 + i] = rB[i];
            // unused slots of a range still receive the (zero-filled or neighbouring) columns of the
            // 128-column box; give them the block index that really lands there for bookkeeping
            for (int r = 0; r < gd.nranges; ++r)
                for (int i = 1; i < 4; ++i)
                    if (gd.blk[4 * r + i] < 0) gd.blk[4 * r + i] = gd.blk[4 * r] + i;
            std::vector<UnitDesc> us;
            for (size_t k = u0; k < std::min(units.size(), u0 + NCW); ++k) {
                UnitDesc u;
                memset(&u, 0, sizeof(u));
                u.sa = (int8_t)slot_of(units[k].ba);
                u.sb = (int8_t)slot_of(units[k].bb);
                u.kind = (int8_t)units[k].kind;
                u.mbc = (int8_t)nb8(units[k].ba);
                u.nbc = (int8_t)nb8(units[k].bb);
                u.sums = (int8_t)units[k].sums;
                us.push_back(u);
            }
            // LPT assignment of units to warps so that the 4 SMSPs (warp % 4) carry equal FP64-pipe load
            std::sort(us.begin(), us.end(),
                      [](const UnitDesc& a, const UnitDesc& b) { return unit_cost(a) > unit_cost(b); });
            double load[4] = {0, 0, 0, 0};
            int used[4] = {0, 0, 0, 0};
            for (const UnitDesc& u : us) {
                int best = -1;
                for (int s = 0; s < 4; ++s)
                    if (used[s] < NCW / 4 && (best < 0 || load[s] < load[best])) best = s;
                gd.unit[used[best] * 4 + best] = u;
                used[best]++;
                load[best] += unit_cost(u);
            }
            double mx = std::max(std::max(load[0], load[1]), std::max(load[2], load[3]));
            S.groups.push_back(gd);
            S.gcost.push_back(mx);              // + the per-stage overhead, added once the regime is known
        }
    };
    bool sw_assigned = false;
    for (int I = 0; I < ns; ++I) {
        std::vector<int> bi;
        for (int b = I * 4; b < std::min(nbx, I * 4 + 4); ++b) bi.push_back(b);
        // diagonal super-tile (+ X'Y and Y'Y units when they fit in the same group)
        {
            std::vector<int> ry;
            std::vector<RawUnit> units;
            for (size_t a = 0; a < bi.size(); ++a)
                for (size_t b = a; b < bi.size(); ++b) {
                    RawUnit u{bi[a], bi[b], a == b ? 2 : 1, a == b ? 1 : 0};
                    if (a == b && !sw_assigned) {
                        u.sums |= 2;
                        sw_assigned = true;
                    }
                    units.push_back(u);
                }
            const bool merge_y = nby <= 4 &&
                                 (int)units.size() + (int)bi.size() * nby + (I == 0 ? nby : 0) <= NCW;
            if (merge_y) {
                for (int y = 0; y < nby; ++y) ry.push_back(nbx + y);
                for (int b : bi)
                    for (int y = 0; y < nby; ++y) units.push_back(RawUnit{b, nbx + y, 1, 0});
                if (I == 0)
                    for (int y = 0; y < nby; ++y) units.push_back(RawUnit{nbx + y, nbx + y, 2, 1});
            }
            emit(bi, ry, units);
            if (!merge_y) {
                // separate X_I' Y groups, Y blocks taken 4 at a time
                for (int y0 = 0; y0 < nby; y0 += 4) {
                    std::vector<int> s2;
                    std::vector<RawUnit> u2;
                    for (int y = y0; y < std::min(nby, y0 + 4); ++y) s2.push_back(nbx + y);
                    for (int b : bi)
                        for (int y = y0; y < std::min(nby, y0 + 4); ++y)
                            u2.push_back(RawUnit{b, nbx + y, 1, 0});
                    emit(bi, s2, u2);
                }
                if (I == 0) {
                    for (int y0 = 0; y0 < nby; y0 += MAXSLOT) {
                        std::vector<int> s3, s4;
                        std::vector<RawUnit> u3;
                        for (int y = y0; y < std::min(nby, y0 + MAXSLOT); ++y) {
                            (y < y0 + 4 ? s3 : s4).push_back(nbx + y);
                            u3.push_back(RawUnit{nbx + y, nbx + y, 2, 1});
                        }
                        emit(s3, s4, u3);
                    }
                }
            }
        }
        for (int J = I + 1; J < ns; ++J) {
            std::vector<int> bj;
            for (int b = J * 4; b < std::min(nbx, J * 4 + 4); ++b) bj.push_back(b);
            std::vector<RawUnit> units;
            for (int a : bi)
                for (int b : bj) units.push_back(RawUnit{a, b, 1, 0});
            emit(bi, bj, units);
        }
    }
}

static void build_segments(int64_t nstages, int ncta, int64_t zone_len, Schedule& S) {
    const int ng = (int)S.groups.size();
    const double L = (double)zone_len;
    double total = 0;
    for (int g = 0; g < ng; ++g) total += S.gcost[g] * L;
    const double share = total / ncta;
    S.segs.clear();
    S.cta_seg.assign(ncta + 1, 0);
    S.group_seg.assign(ng + 1, 0);
    // walk groups in order; CTA c owns the flattened cost interval [c*share, (c+1)*share) of a zone
    std::vector<std::vector<SegDesc>> per_cta(ncta);
    const int64_t xmax = zone_len << 16;
    double pos = 0;  // flattened cost position of the start of the current group
    for (int g = 0; g < ng; ++g) {
        const double gc = S.gcost[g];
        const double end = pos + gc * L;
        auto bnd = [&](int cta) {  // stage offset (16.16) where CTA `cta` starts inside this group
            double b = ((double)cta * share - pos) / gc;
            int64_t xb = (int64_t)llround(b * 65536.0);
            return std::max<int64_t>(0, std::min<int64_t>(xmax, xb));
        };
        int c_lo = std::min(ncta - 1, (int)(pos / share + 1e-9));
        int c_hi = std::min(ncta - 1, (int)(end / share - 1e-9));
        for (int cta = c_lo; cta <= c_hi; ++cta) {
            const int64_t x0 = (cta == c_lo) ? 0 : bnd(cta);
            const int64_t x1 = (cta == c_hi) ? xmax : bnd(cta + 1);
            if (x1 > x0) per_cta[cta].push_back(SegDesc{g, 0, x0, x1});
        }
        pos = end;
    }
    // the walk emits CTA-contiguous (K1) and group-contiguous (K1b) order at once
    for (int c = 0; c < ncta; ++c) {
        S.cta_seg[c] = (int32_t)S.segs.size();
        for (const SegDesc& sd : per_cta[c]) S.segs.push_back(sd);
    }
    S.cta_seg[ncta] = (int32_t)S.segs.size();
    int cur = 0;
    for (int g = 0; g < ng; ++g) {
        S.group_seg[g] = cur;
        while (cur < (int)S.segs.size() && S.segs[cur].group == g) ++cur;
    }
    S.group_seg[ng] = cur;
    (void)nstages;
}

#ifdef JCB_K1_TRACE
static long long* g_trace = nullptr;
extern "C" int jcb200_debug_trace(long long* host, int n) {
    if (!g_trace) return -1;
    cudaDeviceSynchronize();
    cudaMemcpy(host, g_trace, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost);
    return 4 * 64 * NCW * 3 + 256;
}
#endif

int launch_pivot(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy, int64_t n,
                 int64_t p, int64_t q, double* d_pivot) {
    JCB_TRY(ensure(c->pivot_ws, (size_t)(p + q) * 8 + 16));
    // Measured on B200 (profiles/k1_r01_notes.md): the centring-free loop is NOT faster (8.65 vs 8.45 ms
    // at C2 — the loop is bound by LDS -> DMMA issue latency, not by the 8 DADDs per k8-step), so
    // centring stays on unless JCB_AUTO_NOCENTER=1 asks for the data-dependent decision.
    static int force = -1;
    if (force < 0) {
        const char* e = getenv("JCB_AUTO_NOCENTER");
        force = (e && atoi(e)) ? 0 : 1;
    }
    // the completion counter lives behind the ratios; it is zeroed when the workspace is (re)allocated and
    // resets itself at the end of every launch
    unsigned int* counter = (unsigned int*)((double*)c->pivot_ws.p + (p + q));
    if (c->pivot_ctr_zeroed != (void*)counter || c->pivot_ctr_base != c->pivot_ws.p) {
        JCB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), c->stream));
        c->pivot_ctr_zeroed = (void*)counter;
        c->pivot_ctr_base = c->pivot_ws.p;
    }
    pivot_kernel<<<(int)(p + q), 256, 0, c->stream>>>(dX, ldx, dY, ldy, n, (int)p, (int)q, d_pivot,
                                                      (double*)c->pivot_ws.p, counter, force);
    JCB_LAUNCH_CHECK();
    return 0;
}

int launch_gram(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy,
                const double* dw, int64_t n, int64_t p, int64_t q, const double* d_pivot,
                double* d_packed, int accumulate) {
    ReduceDst dst;
    dst.n = 1;
    dst.p[0] = d_packed;
    return launch_gram_to(c, dX, ldx, dY, ldy, dw, n, p, q, d_pivot, dst, accumulate);
}

int launch_gram_to(Ctx* c, const double* dX, int64_t ldx, const double* dY, int64_t ldy,
                   const double* dw, int64_t n, int64_t p, int64_t q, const double* d_pivot,
                   const ReduceDst& dst, int accumulate) {
    if (accumulate && dst.n != 1) {
        set_error("gram: accumulation needs a single destination");
        return JCB200_EINVAL;
    }
    if (((uintptr_t)dX & 15) || ((uintptr_t)dY & 15) || (dw && ((uintptr_t)dw & 15)) || (ldx & 1) ||
        (ldy & 1)) {
        set_error("gram: device pointers must be 16-byte aligned and leading dimensions even "
                  "(TMA requirement); repack the shard");
        return JCB200_EALIGN;
    }
    const int64_t nstages = (n + KT - 1) / KT;
    const int ncta = c->num_sms;
    // Joint column space: when Y lies right behind X in memory with the same leading dimension (the library's own
    // staging buffers; `device.colmajor_empty_xy`), [X | Y] IS one n x (p + q) matrix and the whole packed Gram is
    // the upper triangle of ITS Gram: no separate Y blocks — Y's q columns ride in the padding of X's last
    // 32-column block (p = 500: 20 + 10 of 32) instead of costing a 16-column-wide X'Y unit per X block (-5 % of the
    // kernel's FP64-pipe time at C2).  K1b maps the joint coordinates back to the packed layout.
    static int joint_ok = -1;
    if (joint_ok < 0) {
        const char* e = getenv("JCB_GRAM_JOINT");
        joint_ok = (e && atoi(e) == 0) ? 0 : 1;
    }
    const bool joint = joint_ok && dY == dX + p * ldx && ldy == ldx;
    const int64_t pj = joint ? p + q : p, qj = joint ? 0 : q;      // what the kernel sees
    const int nbx = (int)((pj + CB - 1) / CB);
    // ---- schedule (cached on (pj, qj, nstages); a few entries: the streamed fits launch K1 on chunks of up to three
    // lengths per call, and a miss drains the stream)
    Ctx::Sched* sc = nullptr;
    for (int i = 0; i < Ctx::NSCHED; ++i)
        if (c->sched[i].p == pj && c->sched[i].q == qj && c->sched[i].nst == nstages) sc = &c->sched[i];
    if (!sc) {
        sc = &c->sched[0];
        for (int i = 1; i < Ctx::NSCHED; ++i)
            if (c->sched[i].stamp < sc->stamp) sc = &c->sched[i];          // least recently used
        Schedule S;
        build_groups(pj, qj, S);
        // zone = the row range all CTAs sweep together; sized so that one zone of [X Y] (~25 MB) stays
        // in the 126 MB L2 while the groups that share its column blocks read it.  With more groups
        // than CTAs there is nothing to co-schedule: one zone.
        int64_t zone_len = nstages;
        const char* zenv = getenv("JCB_ZONE_MB");
        const double zone_mb = zenv ? atof(zenv) : 25.0;
        const bool zoned = (int)S.groups.size() * 2 <= ncta && zone_mb > 0;
        for (double& g : S.gcost) g += stage_overhead(zoned);
        if (zoned) {
            const int64_t rows = (int64_t)(zone_mb * 1e6 / ((double)(pj + qj) * 8.0));
            zone_len = std::max<int64_t>(16, rows / KT);
            if (zone_len > nstages) zone_len = nstages;
        }
        build_segments(nstages, ncta, zone_len, S);
        auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
        const size_t gb = up16(S.groups.size() * sizeof(GroupDesc)),
                     sb = up16(S.segs.size() * sizeof(SegDesc)), cb = up16(S.cta_seg.size() * 4),
                     qb = up16(S.group_seg.size() * 4);
        const size_t tot = gb + sb + cb + qb;
        // a previous launch may still read the entry that is being replaced, and the staging buffer may still be
        // on its way to the device: drain the stream before touching either
        JCB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->sched_host_bytes < tot) {
            if (c->sched_host) cudaFreeHost(c->sched_host);
            JCB_CUDA(cudaMallocHost(&c->sched_host, tot));
            c->sched_host_bytes = tot;
        }
        JCB_TRY(ensure(sc->dev, tot));
        unsigned char* h = (unsigned char*)c->sched_host;
        memset(h, 0, tot);
        memcpy(h, S.groups.data(), S.groups.size() * sizeof(GroupDesc));
        memcpy(h + gb, S.segs.data(), S.segs.size() * sizeof(SegDesc));
        memcpy(h + gb + sb, S.cta_seg.data(), S.cta_seg.size() * 4);
        memcpy(h + gb + sb + cb, S.group_seg.data(), S.group_seg.size() * 4);
        JCB_CUDA(cudaMemcpyAsync(sc->dev.p, h, tot, cudaMemcpyHostToDevice, c->stream));
        sc->off_segs = gb;
        sc->off_cta = gb + sb;
        sc->off_gseg = gb + sb + cb;
        sc->p = pj;
        sc->q = qj;
        sc->nst = nstages;
        sc->zone_len = zone_len;
        sc->ngroups = (int)S.groups.size();
        sc->nsegs = (int)S.segs.size();
    }
    sc->stamp = ++c->sched_clock;
    const int ng = sc->ngroups, nsegs = sc->nsegs;
    unsigned char* d = (unsigned char*)sc->dev.p;
    const GroupDesc* dgroups = (const GroupDesc*)d;
    const SegDesc* dsegs = (const SegDesc*)(d + sc->off_segs);
    const int32_t* dcta = (const int32_t*)(d + sc->off_cta);
    const int32_t* dgseg = (const int32_t*)(d + sc->off_gseg);
    JCB_TRY(ensure(c->partials, (size_t)nsegs * NCW * UNIT_STRIDE * 8));

    CUtensorMap mapX, mapY, mapW;
    JCB_TRY(make_map_2d(&mapX, dX, n, pj, ldx, KT, 4 * CB));
    if (joint) mapY = mapX;
    else JCB_TRY(make_map_2d(&mapY, dY, n, q, ldy, KT, 4 * CB));
    if (dw) {
        JCB_TRY(make_map_1d(&mapW, dw, n, KT));
    } else {
        mapW = mapX;
    }
    GramParams prm;
    prm.groups = dgroups;
    prm.segs = dsegs;
    prm.cta_seg = dcta;
    prm.pivot = d_pivot;
    prm.partials = (double*)c->partials.p;
    prm.n = n;
    prm.nst = nstages;
    prm.zone_len = (int32_t)sc->zone_len;
    prm.nzones = (int32_t)((nstages + sc->zone_len - 1) / sc->zone_len);
    prm.p = (int)pj;
    prm.q = (int)qj;
    prm.nbx = nbx;
    prm.weighted = dw ? 1 : 0;
#ifdef JCB_K1_TRACE
    {
        static long long* tracebuf = nullptr;
        if (!tracebuf) cudaMalloc(&tracebuf, (4 * 64 * NCW * 3 + 256) * sizeof(long long));
        cudaMemsetAsync(tracebuf, 0, (4 * 64 * NCW * 3 + 256) * sizeof(long long), c->stream);
        prm.trace = tracebuf;
        g_trace = tracebuf;
    }
#endif

    if (!c->k1_attr_set) {      // function attributes are per device
        JCB_CUDA(cudaFuncSetAttribute(gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      K1_SMEM));
        JCB_CUDA(cudaFuncSetAttribute(gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      K1_SMEM));
        c->k1_attr_set = true;
    }
    phase_begin(c, JCB200_T_GRAM);
    const int slot = (int)(c->gram_calls % Ctx::GRAM_RING);
    cudaEventRecord(c->gram_ev0[slot], c->stream);
    if (dw)
        gram_kernel<true><<<ncta, K1_THREADS, K1_SMEM, c->stream>>>(mapX, mapY, mapW, prm);
    else
        gram_kernel<false><<<ncta, K1_THREADS, K1_SMEM, c->stream>>>(mapX, mapY, mapW, prm);
    JCB_LAUNCH_CHECK();
    cudaEventRecord(c->gram_ev1[slot], c->stream);
    c->gram_calls++;
    phase_end(c, JCB200_T_GRAM);
    phase_begin(c, JCB200_T_REDUCE);
    gram_reduce_kernel<<<ng * NCW, 256, 0, c->stream>>>(dgroups, dgseg, (const double*)c->partials.p,
                                                        dst, (int)p, (int)q, nbx, accumulate, joint ? 1 : 0);
    JCB_LAUNCH_CHECK();
    phase_end(c, JCB200_T_REDUCE);
    return 0;
}

}  // namespace jcb
