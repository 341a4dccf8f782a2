// trsv_ls.cuh — K4, second algorithm: the LEVEL-STREAM triangular solve. One CTA solves one system.
//
// The sync-free solve (sptrsv.cuh) pays an L2 store -> poll hop per level (0.5 us between most SM pairs of a B200,
// measured: tools/microbench/pingpong.cu). Matrices with NARROW levels (2-D 5-point factors: <= n rows per level,
// 2n-1 levels) are latency bound by that hop by a factor ~30. Here every dependency stays inside one SM:
//
//   * the factor is stored a second time in LEVEL ORDER (dp_sptrsv_permute): row r of the copy is row perm[r] of T,
//     columns renumbered to positions in that order, entries of a row in T's order (so the sum stays bit-identical
//     to plain substitution). The copy is still triangular, and a level is a contiguous run of rows AND of entries;
//   * the CTA streams the copy through the TMA tile pipeline (tilepipe.cuh) - 512 rows per tile, one thread per row;
//   * levels are ordered by a ring of mbarriers, one phase per level step: a warp arrives at a step as soon as its own
//     rows of that level are solved (at once if it has none) and waits on the previous step only right before it
//     solves rows itself. Warps with nothing to do in the remaining levels of a tile are already fetching the next
//     tile's rows into registers, so the hand-over between tiles is off the critical path;
//   * a solved value goes to global memory in the original numbering and into a shared-memory window indexed by
//     position (the last kLsWindow positions), where the next levels pick it up with shared-memory latency.
// A factor qualifies (dp_sptrsv_ls_limits) when a 512-row tile of the copy fits one pipeline stage, a row has at most
// kLsRowEntries entries (they live in registers) and no dependency is further back than the window reaches; stencil
// IC(0) factors do. Everything else is solved by the sync-free kernel.
// Per level the cost is an mbarrier hand-off plus a few shared-memory reads instead of an L2 hop; a batch runs one system per
// CTA, so 64+ systems keep the whole HBM busy. Wide levels (3-D factors) stay with the sync-free multi-SM solve.
#pragma once

#include "tilepipe.cuh"

namespace dp {

constexpr int kLsWindow = 1024;   // positions of the solution kept in shared memory (power of two)
constexpr int kLsRowEntries = 4;  // diagonal + 3 dependencies: 5-/7-point factors

#ifdef DPCG_LS_TRACE
__device__ long long g_ls_trace[8 * 256];
#define LS_TRACE(tile, slot) \
    if (threadIdx.x == 0 && blockIdx.x == 0 && (tile) < 256) g_ls_trace[(tile) * 8 + (slot)] = clock64()
#else
#define LS_TRACE(tile, slot)
#endif

#ifdef DPCG_LS_TRACE
__device__ __forceinline__ void g_ls_trace_levels(int tile, int nl) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && tile < 256) g_ls_trace[tile * 8 + 7] = nl;
}
#else
__device__ __forceinline__ void g_ls_trace_levels(int, int) {}
#endif

constexpr int kLsRing = 16;  // level steps a warp may be ahead of the slowest one (enforced every kLsRing / 2 steps)

using LsPipe = PipeT<kLsCap, kLsStages>;  // 5 stages of 1536 entries over the SpMV pipeline's bytes

// Shared memory of the solve besides the stage bytes and the tile table. `items` carries the pipeline's item count
// from one solve to the next (the mbarrier parities depend on it).
struct LsShared {
    PipeBarriers<kLsStages> bar;
    unsigned items;
    int pad;
    unsigned long long steps;  // level steps taken so far (carried like `items`: fixes the parities of the ring below)
    alignas(8) unsigned long long level_done[kLsRing];  // step s completes phase s / kLsRing of slot s % kLsRing
    double win[kLsWindow];
    __device__ __forceinline__ void init() {  // thread 0, once per kernel, before a CTA barrier
        for (int s = 0; s < kLsStages; ++s) {
            mbar_init(&bar.full[s], 1u);
            mbar_init(&bar.empty[s], (unsigned)kWarpsPerBlock);
        }
        for (int s = 0; s < kLsRing; ++s) mbar_init(&level_done[s], (unsigned)kWarpsPerBlock);
        mbar_fence_init();
        items = 0u, steps = 0ull;
    }
};

struct LsFactor {
    const int* rowptr;  // nullptr: no level-ordered copy, use the sync-free solve
    const int* col;     // positions (level order)
    const double* val;  // the diagonal entry of each row holds 1 / T_ii
    const int* perm;    // position -> row of T
    const int* lvl;     // level of the row at each position (non-decreasing)
    int n, nnz;
};

// Solve T x = rhs for one system with the whole CTA. `stage_bytes`: the pipeline bytes (no other pipeline may have items
// in flight), `tab`: >= `tabcap` tile descriptors of shared memory. rhs and x are in the original numbering.
// All kBlock threads must call; ends with a CTA barrier.
template <bool kUpper>
__device__ __forceinline__ void trsv_level_stream(const LsFactor& F, const double* rhs, double* x, unsigned char* stage_bytes,
                                                  LsShared& ls, TileDesc* tab, int tabcap) {
    LsPipe pipe;
    pipe.init(stage_bytes, &ls.bar, false);
    pipe.resume(ls.items);
    double* win = ls.win;
    const CsrView P{F.rowptr, F.col, F.val, F.n, F.nnz};
    const int n = F.n;
    const int ntiles = (n + kTileRows - 1) / kTileRows;
    const int tid = threadIdx.x;
    constexpr int kDeps = kLsRowEntries - 1;
    constexpr int kAhead = kLsStages - 2;      // tiles in flight; the stage issued into was released a whole tile ago
    constexpr int kProducer = kBlock - kWarp;  // lane 0 of the LAST warp: its rows come last in every tile
    const int lane = tid & 31;
    unsigned long long step0 = ls.steps;  // step of the first level of the current tile
    const unsigned long long solve_step0 = step0;  // the first level of a solve has no dependencies
    auto step_wait = [&](unsigned long long s) {  // all 16 warps have arrived at step s
        while (!mbar_try_wait(&ls.level_done[s % kLsRing], (unsigned)(s / kLsRing) & 1u)) {
        }
    };
    for (int ga = 0; ga < ntiles; ga += tabcap) {
        const int cnt = min(tabcap, ntiles - ga);
        __syncthreads();  // the previous round's table (and the previous system's window) are no longer read
        for (int i = tid; i < cnt; i += kBlock) {
            TileDesc d;
            tile_desc_fill(d, P, ga + i);
            const int base = (ga + i) * kTileRows;
            d.n = __ldg(F.lvl + base);                             // first level of the tile (field reused)
            d.nnz = __ldg(F.lvl + min(n, base + kTileRows) - 1);   // last level of the tile
            d.sys = 0;
            tab[i] = d;
        }
        __syncthreads();
        // every tile is one item: the producer lane keeps kAhead tiles in flight
        pipe.tab = tab, pipe.ntiles = cnt;
        pipe.p_count = pipe.c_count;  // every thread tracks the count; only the producer's copy is used to issue
        if (tid == kProducer)
            for (int t = 0; t < min(cnt, kAhead); ++t) pipe.issue_tile(t);
        // per-row metadata one tile ahead: original row, level, row extent
        int orig_n = -1, lvl_n = -1, rs_n = 0, re_n = 0;
        auto load_meta = [&](int tile) {
            const int r = tile * kTileRows + tid;
            orig_n = -1, lvl_n = -1, rs_n = 0, re_n = 0;
            if (r < n) {
                orig_n = ldg_here_s32(F.perm + r), lvl_n = ldg_here_s32(F.lvl + r);
                rs_n = ldg_here_s32(F.rowptr + r), re_n = ldg_here_s32(F.rowptr + r + 1);
            }
        };
        load_meta(ga);
        int pre = 0;  // level steps of the upcoming tile this warp has already arrived at
        auto step_arrive = [&](unsigned long long s) {
            // a slot of the ring is reused every kLsRing steps: stay less than a ring ahead of the slowest warp
            if (s % (kLsRing / 2) == 0ull && s >= (unsigned long long)(kLsRing / 2)) step_wait(s - kLsRing / 2);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ls.level_done[s % kLsRing]);
        };
        for (int i = 0; i < cnt; ++i) {
            const TileDesc& d = tab[i];
            LS_TRACE(ga + i, 0);
            const int orig = orig_n, lvl = lvl_n, rs = rs_n, re = re_n;
            const double bi = orig >= 0 ? ldcg_here_f64(rhs + orig) : 0.0;
            if (i + 1 < cnt) load_meta(ga + i + 1);
            LS_TRACE(ga + i, 1);
            if (tid == kProducer && i + kAhead < cnt) pipe.issue_tile(i + kAhead);
            LS_TRACE(ga + i, 2);
            const unsigned stage = pipe.wait_item();
            LS_TRACE(ga + i, 3);
            const double* __restrict__ sv = pipe.stage_val(stage);
            const int* __restrict__ sc = pipe.stage_col(stage);
            const int as = d.cs & ~3;
            const int e = (kUpper ? rs + 1 : rs) - as, ndep = (kUpper ? re : re - 1) - as - e;
            // the whole row moves to registers: window slots of the dependencies, their coefficients, 1 / diagonal
            double rcp = 0.0, v[kDeps];
            int w[kDeps];
#pragma unroll
            for (int u = 0; u < kDeps; ++u) {
                w[u] = 0, v[u] = 0.0;
                if (u < ndep) w[u] = sc[e + u] & (kLsWindow - 1), v[u] = sv[e + u];
            }
            if (orig >= 0) rcp = sv[(kUpper ? rs : re - 1) - as];  // the copy stores 1 / diagonal (dp_sptrsv_permute)
            const int slot = (d.ltile * kTileRows + tid) & (kLsWindow - 1);
            // levels of this warp's rows: [wl0, wl1] (levels ascend with the position; -1 = the warp has no rows)
            const int wl0 = __shfl_sync(kFull, lvl, 0);
            const int wl1 = __reduce_max_sync(kFull, lvl);
            const int lv0 = d.n, lv1 = d.nnz;
            double xsol = 0.0;
            LS_TRACE(ga + i, 4);
            g_ls_trace_levels(ga + i, lv1 - lv0 + 1);
#pragma unroll 1
            for (int l = lv0 + pre; l <= lv1; ++l) {
                const unsigned long long s = step0 + (unsigned long long)(l - lv0);
                if (wl0 >= 0 && l >= wl0 && l <= wl1) {
                    if (s != solve_step0) step_wait(s - 1ull);  // every row of the earlier levels is solved
                    if (lvl == l) {
                        double sum = 0.0;
#pragma unroll
                        for (int u = 0; u < kDeps; ++u)
                            if (u < ndep) sum = __dadd_rn(sum, __dmul_rn(v[u], win[w[u]]));
                        xsol = __dmul_rn(__dsub_rn(bi, sum), rcp);
                        win[slot] = xsol;
                    }
                }
                step_arrive(s);
            }
            step0 += (unsigned long long)(lv1 - lv0 + 1);
            // Before the hand-over (store, stage release, next rows into registers) the warp already arrives at the
            // steps of the next tile that come before its own first level there: nobody waits for its hand-over.
            pre = 0;
            if (i + 1 < cnt) {
                const int lv0n = tab[i + 1].n, lv1n = tab[i + 1].nnz;
                const int wl0n = __shfl_sync(kFull, lvl_n, 0);  // lvl_n: this thread's level in the next tile
                pre = wl0n < 0 ? lv1n - lv0n + 1 : max(0, wl0n - lv0n);
                for (int k = 0; k < pre; ++k) step_arrive(step0 + (unsigned long long)k);
            }
            // the global copy (original numbering)
            LS_TRACE(ga + i, 5);
            if (orig >= 0) x[orig] = xsol;
            pipe.release();
            LS_TRACE(ga + i, 6);
        }
    }
    __syncthreads();
    if (tid == 0) ls.items = pipe.c_count, ls.steps = step0;
    __syncthreads();
}

}  // namespace dp
