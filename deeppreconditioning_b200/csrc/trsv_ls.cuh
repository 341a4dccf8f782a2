// trsv_ls.cuh — K4, second algorithm: the LEVEL-STREAM triangular solve. One CTA solves one system.
//
// The sync-free solve (sptrsv.cuh) pays an L2 store -> poll hop per level (0.5 us between most SM pairs of a B200,
// measured: tools/microbench/pingpong.cu). Matrices with NARROW levels (2-D 5-point factors: <= n rows per level,
// 2n-1 levels) are latency bound by that hop by a factor ~30. Here every dependency stays inside one SM:
//
//   * the factor is stored a second time in LEVEL ORDER (dp_sptrsv_permute): row r of the copy is row perm[r] of T,
//     columns renumbered to positions in that order, entries of a row in T's order (so the sum stays bit-identical
//     to plain substitution). The copy is still triangular, and a level is a contiguous run of rows AND of entries;
//   * the CTA streams the copy through the TMA tile pipeline (tilepipe.cuh) - 512 rows per tile, one thread per row,
//     the tile's row pointers and (vectors in level order) right-hand sides riding on the same stage;
//     the row moves to registers (window slots of its dependencies, coefficients, 1 / diagonal) and the stage is handed
//     back at once;
//   * dependencies are awaited on a SHARED-MEMORY WINDOW of the solution, indexed by position and pre-armed with the
//     kPending NaN pattern - the protocol of the sync-free solve (data is its own flag, one 8-byte store per row), with a
//     shared-memory round trip (~30 cycles) in place of the L2 hop. There is no level barrier: a row is solved as soon as
//     ITS dependencies are, warps of the CTA work on different levels (and tiles) at the same time.
//     Round 1 ordered the levels by a ring of mbarriers, one phase per level step, every warp arriving at every step
//     whether it had rows in the level or not. Both forms take ~0.45 us per level of a 316^2 factor (profiles/r2/README.md:
//     neither a gate word before the polls, nor four producer warps, nor a one-tile lookahead of the row fetch, nor
//     per-chunk mbarriers in place of the polls moved that number); this one needs no level array, takes its vectors by
//     position (bulk copies) and is the faster one inside PCG on level-ordered systems.
//   * a solved value also goes to global memory (original numbering through `perm`, or by position when the caller keeps
//     its vectors in level order: perm == nullptr, coalesced).
// The window. Tile t's slots are armed by the producer warp right before it issues tile t's bulk copies, i.e. after
// every warp has released tile t - kStages: those warps have finished every tile up to t - kStages - 1 and may still be
// polling for rows of tile t - kStages, whose dependencies reach back at most kTileRows positions (the eligibility
// limit), into tile t - kStages - 1. A window of kStages + 2 tiles therefore never re-arms a slot that can still be
// read, and a consumer that has been handed tile t (its `full` barrier) finds the slots of tile t armed: the arming
// stores are ordered before the copy's issue, whose completion the consumer acquires.
// A factor qualifies (dp_sptrsv_ls_limits) when a 512-row tile of the copy fits one pipeline stage, a row has at most
// kLsRowEntries entries (they live in registers) and no dependency is further back than kTileRows positions; stencil
// IC(0) factors do. Everything else is solved by the sync-free kernel (or, in batches, the tile-stream kernel).
#pragma once

#include "tilepipe.cuh"

namespace dp {

constexpr int kLsRowEntries = 4;           // diagonal + 3 dependencies: 5-/7-point factors
constexpr int kLsDepDistance = kTileRows;  // a dependency is at most this many positions back
static_assert(kTileRows == 512, "window slots are computed with q >> 9");

// Geometry of a solve with kStages pipeline stages: per stage the tile's values and column indices (kLsCap entries, the
// layout of PipeT), its kTileRows + 1 row pointers and - when the vectors are indexed by position - its kTileRows
// right-hand sides: everything a tile needs that is contiguous travels by bulk copy, nothing is fetched through registers
// on a warp's path from one tile to the next (an L2 round trip there was the floor of a level step: 1.6 levels per tile).
// The window follows the stages.
constexpr int kLsRowptrSlots = kTileRows + 8;  // kTileRows + 1 used; 16-byte granular copies
template <int kStages>
struct LsGeom {
    static constexpr int kWindowTiles = kStages + 2;
    static constexpr size_t kMatrixBytes = PipeGeom<kLsCap, kStages>::kBytes;  // values, then columns (PipeT)
    static constexpr size_t kRowptrAt = kMatrixBytes;
    static constexpr size_t kRhsAt = kRowptrAt + (size_t)kStages * kLsRowptrSlots * 4;
    static constexpr size_t kWindowAt = kRhsAt + (size_t)kStages * kTileRows * 8;
    static constexpr size_t kBytes = kWindowAt + (size_t)kWindowTiles * kTileRows * 8;
    static_assert(kMatrixBytes % 16 == 0 && kRhsAt % 16 == 0 && kWindowAt % 16 == 0, "bulk-copy alignment");
};
constexpr int kLsStagesFused = 3;  // inside the fused PCG kernel: laid over the SpMV pipeline's bytes (2 CTAs per SM)
constexpr int kLsStagesAlone = 7;  // the standalone batch kernel: one CTA per SM, deep enough for the DRAM latency
static_assert(LsGeom<kLsStagesFused>::kBytes <= kPipeRawBytes, "the fused kernel lends its SpMV stages to the solve");

// Shared memory of the solve besides its bytes: the pipeline's barriers and its item count, carried from one solve to
// the next (the mbarrier parities depend on it).
template <int kStages>
struct LsSharedT {
    PipeBarriers<kStages> bar;
    unsigned items;
    int pad;
    __device__ __forceinline__ void init() {  // thread 0, once per kernel, before a CTA barrier
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bar.full[s], 1u);
            mbar_init(&bar.empty[s], (unsigned)kWarpsPerBlock);
        }
        mbar_fence_init();
        items = 0u;
    }
};
using LsShared = LsSharedT<kLsStagesFused>;

struct LsFactor {
    const int* rowptr;  // nullptr: no level-ordered copy, use the sync-free solve
    const int* col;     // positions (level order)
    const double* val;  // the diagonal entry of each row holds 1 / T_ii
    const int* perm;    // position -> row of T; nullptr: rhs and x are indexed by position
    const int* lvl;     // level of the row at each position (not read by the solve; kept for the callers' bookkeeping)
    int n, nnz;
};

__device__ __forceinline__ unsigned long long lds_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}

#ifdef DPCG_LS_TRACE  // per-tile timeline of warp `DPCG_LS_TRACE_WARP` of CTA 0 (tools/trace_ls.py)
#ifndef DPCG_LS_TRACE_WARP
#define DPCG_LS_TRACE_WARP 0
#endif
__device__ long long g_ls_trace[8 * 256];
#define LS_TRACE(tile, slot)                                                                        \
    if (threadIdx.x == 32 * DPCG_LS_TRACE_WARP && blockIdx.x == 0 && (tile) < 256) g_ls_trace[(tile) * 8 + (slot)] = clock64()
#else
#define LS_TRACE(tile, slot)
#endif

constexpr unsigned kLsSpinBudget = 1u << 22;  // shared-memory polls before a row gives up (malformed copy): no hang

// Solve T x = rhs for one system with the whole CTA. `bytes`: LsGeom<kStages>::kBytes of shared memory (no other
// pipeline may have items in flight). CTAs of kBlock threads let their last warp double as the producer; CTAs of
// kBlock + 32 threads give it a warp of its own (the standalone kernel). F.perm == nullptr: rhs and x are indexed by
// position and rhs must be 16-byte aligned (its tiles travel by bulk copy); else they are in the original numbering and a
// row's right-hand side is gathered through registers two tiles ahead. F.rowptr must be 16-byte aligned.
// All threads must call; ends with a CTA barrier.
template <bool kUpper, int kStages>
__device__ __forceinline__ void trsv_level_stream(const LsFactor& F, const double* rhs, double* x, unsigned char* bytes,
                                                  LsSharedT<kStages>& ls) {
    using Geom = LsGeom<kStages>;
    using PipeLs = PipeT<kLsCap, kStages>;
    constexpr int kWt = Geom::kWindowTiles;
    constexpr int kDeps = kLsRowEntries - 1;
    PipeLs pipe;
    pipe.init(bytes, &ls.bar, false);
    pipe.resume(ls.items);
    int* rowptr_st = reinterpret_cast<int*>(bytes + Geom::kRowptrAt);
    double* rhs_st = reinterpret_cast<double*>(bytes + Geom::kRhsAt);
    unsigned long long* win = reinterpret_cast<unsigned long long*>(bytes + Geom::kWindowAt);
    const int n = F.n;
    const int ntiles = (n + kTileRows - 1) / kTileRows;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool positions = F.perm == nullptr;
    const bool dedicated = blockDim.x > kBlock;               // a producer warp of its own
    const bool producer = dedicated ? tid >= kBlock : tid >= kBlock - kWarp;
    __syncthreads();  // the previous solve's window is no longer read

    // ---- producer warp: tile t = wait for its stage, arm its window slots, issue its bulk copies --------------------
    unsigned issued = pipe.p_count;  // warp-uniform copy of the item count
    int bound_base = -1, bound_lo = 0, bound_hi = 0;  // lane k: entry range of tile bound_base + k
    auto produce = [&](int t) {
        if (t - bound_base >= 32 || bound_base < 0) {  // entry ranges of the next 32 tiles, one per lane
            bound_base = t;
            const int tt = t + lane;
            bound_lo = tt < ntiles ? __ldg(F.rowptr + min(tt * kTileRows, n)) : 0;
            bound_hi = tt < ntiles ? __ldg(F.rowptr + min((tt + 1) * kTileRows, n)) : 0;
        }
        const int cs = __shfl_sync(kFull, bound_lo, t - bound_base), ce = __shfl_sync(kFull, bound_hi, t - bound_base);
        const unsigned stage = issued % kStages, use = issued / kStages;
        if (lane == 0 && use > 0) {
            while (!mbar_test_wait(&pipe.empty[stage], (use - 1u) & 1u)) {
            }
        }
        __syncwarp();
        unsigned long long* slots = win + (size_t)(t % kWt) * kTileRows;
#pragma unroll
        for (int k = 0; k < kTileRows / kWarp; ++k) sts_volatile_u64(slots + k * kWarp + lane, kPending);
        __syncwarp();
        if (lane == 0) {  // (the four copies from four lanes at once were measured slower: 294 -> 353 us per 316^2 solve)
            const int as = cs & ~3, r0 = t * kTileRows;
            const unsigned nr = (unsigned)min(kTileRows, n - r0);
            const unsigned ncol = (unsigned)(((ce + 3) & ~3) - as), nval = (unsigned)(((ce + 1) & ~1) - as);
            const unsigned rp_bytes = ((nr + 1u) * 4u + 15u) & ~15u, b_bytes = positions ? (nr * 8u + 15u) & ~15u : 0u;
            unsigned long long* bar = &pipe.full[stage];
            const unsigned long long pol = l2_policy_stream();
            mbar_arrive_expect_tx(bar, ncol * 4u + nval * 8u + rp_bytes + b_bytes);
            bulk_g2s(rowptr_st + (size_t)stage * kLsRowptrSlots, F.rowptr + r0, rp_bytes, bar, pol);
            if (positions) bulk_g2s(rhs_st + (size_t)stage * kTileRows, rhs + r0, b_bytes, bar, pol);
            bulk_g2s(pipe.val0 + (size_t)stage * PipeLs::kSlots, F.val + as, nval * 8u, bar, pol);
            bulk_g2s(pipe.col0 + (size_t)stage * PipeLs::kSlots, F.col + as, ncol * 4u, bar, pol);
        }
        ++issued;
    };
    if (dedicated && producer) {  // free running: the stages' `empty` barriers pace it
        for (int t = 0; t < ntiles; ++t) produce(t);
    }
    constexpr int kAhead = kStages - 1;  // shared producer: tiles in flight beyond the one being solved
    if (!dedicated && producer)
        for (int t = 0; t < min(ntiles, kAhead); ++t) produce(t);

    // ---- consumers: one thread per row ---------------------------------------------------------------------------------
    if (!(dedicated && producer)) {
        // original numbering only: the row's index two tiles ahead, its right-hand side one tile ahead (registers)
        int orig_1 = -1, orig_2 = -1;
        double b_1 = 0.0;
        auto load_orig = [&](int tile) {
            const int r = tile * kTileRows + tid;
            return (tile < ntiles && r < n) ? ldg_here_s32(F.perm + r) : -1;
        };
        if (!positions) {
            orig_1 = load_orig(0), orig_2 = load_orig(1);
            if (orig_1 >= 0) b_1 = ldcg_here_f64(rhs + orig_1);
        }
        for (int t = 0; t < ntiles; ++t) {
            const int r = t * kTileRows + tid;
            const bool valid = r < n;
            int orig = r;
            double bi = 0.0;
            if (!positions) {
                orig = orig_1, bi = b_1;
                orig_1 = orig_2;
                if (orig_1 >= 0) b_1 = ldcg_here_f64(rhs + orig_1);
                orig_2 = load_orig(t + 2);
            }
            if (!dedicated && producer && t + kAhead < ntiles) produce(t + kAhead);
            LS_TRACE(t, 0);
            const unsigned stage = pipe.wait_item();
            LS_TRACE(t, 1);
            const double* __restrict__ sv = pipe.stage_val(stage);
            const int* __restrict__ sc = pipe.stage_col(stage);
            const int* __restrict__ srp = rowptr_st + (size_t)stage * kLsRowptrSlots;
            // the whole row moves to registers: window slots of the dependencies, their coefficients, 1 / diagonal
            const int as = srp[0] & ~3;
            const int rs = valid ? srp[tid] : 0, re = valid ? srp[tid + 1] : 0;
            if (positions && valid) bi = rhs_st[(size_t)stage * kTileRows + tid];
            const int e = (kUpper ? rs + 1 : rs) - as, ndep = valid ? (kUpper ? re : re - 1) - as - e : 0;
            const int tm = t % kWt, tp = (t + kWt - 1) % kWt;  // window tile of this tile / of the previous one
            double rcp = 0.0, v[kDeps];
            int w[kDeps];
#pragma unroll
            for (int u = 0; u < kDeps; ++u) {
                w[u] = 0, v[u] = 0.0;
                if (u < ndep) {
                    const int q = sc[e + u];  // position of the dependency: in this tile or the one before
                    w[u] = ((q >> 9) == t ? tm : tp) * kTileRows + (q & (kTileRows - 1));
                    v[u] = sv[e + u];
                }
            }
            if (valid) rcp = sv[(kUpper ? rs : re - 1) - as];  // the copy stores 1 / diagonal (dp_sptrsv_permute)
            pipe.release();  // the stage goes back before any waiting
            LS_TRACE(t, 2);
            unsigned long long* mine = win + (size_t)tm * kTileRows + tid;
            bool done = !valid;
            unsigned long long uu[kDeps];
#pragma unroll
            for (int u = 0; u < kDeps; ++u) uu[u] = kPending;
            for (unsigned spins = 0;; ++spins) {
                if (!done) {
                    bool all = true;
#pragma unroll
                    for (int u = 0; u < kDeps; ++u) {
                        if (u < ndep && uu[u] == kPending) {
                            uu[u] = lds_volatile_u64(win + w[u]);
                            all = all && uu[u] != kPending;
                        }
                    }
                    if (all || spins > kLsSpinBudget) {  // (budget: a malformed copy yields NaNs, not a hang)
                        double sum = 0.0;
#pragma unroll
                        for (int u = 0; u < kDeps; ++u)
                            if (u < ndep) sum = __dadd_rn(sum, __dmul_rn(v[u], as_double(uu[u])));
                        const double xsol = __dmul_rn(__dsub_rn(bi, sum), rcp);
                        sts_volatile_u64(mine, as_bits(xsol));  // rows of later levels (this warp's too) wait for it
                        x[orig] = xsol;
                        done = true;
                    }
                }
                if (__all_sync(kFull, done)) break;
            }
            LS_TRACE(t, 4);
        }
    }
    __syncthreads();
    if (tid == 0) ls.items = pipe.c_count;
    __syncthreads();
}

}  // namespace dp
