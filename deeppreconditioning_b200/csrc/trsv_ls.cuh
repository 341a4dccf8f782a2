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
//   * rows of a tile are released level by level with a CTA barrier; a solved value goes to global memory in the
//     original numbering and into a shared-memory window indexed by position (the last kLsWindow positions), where
//     the next levels pick it up with shared-memory latency.
// A factor qualifies (dp_sptrsv_ls_limits) when a 512-row tile of the copy fits one pipeline stage, a row has at most
// kLsRowEntries entries (they live in registers) and no dependency is further back than the window reaches; stencil
// IC(0) factors do. Everything else is solved by the sync-free kernel.
// Per level the cost is a CTA barrier plus a few shared-memory reads instead of an L2 hop; a batch runs one system per
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

using LsPipe = PipeT<kLsCap, kLsStages>;  // 5 stages of 1536 entries over the SpMV pipeline's bytes

// Shared memory of the solve besides the stage bytes and the tile table. `items` carries the pipeline's item count
// from one solve to the next (the mbarrier parities depend on it).
struct LsShared {
    PipeBarriers<kLsStages> bar;
    unsigned items;
    int pad;
    double win[kLsWindow];
    __device__ __forceinline__ void init() {  // thread 0, once per kernel, before a CTA barrier
        for (int s = 0; s < kLsStages; ++s) {
            mbar_init(&bar.full[s], 1u);
            mbar_init(&bar.empty[s], (unsigned)kWarpsPerBlock);
        }
        mbar_fence_init();
        items = 0u;
    }
};

struct LsFactor {
    const int* rowptr;  // nullptr: no level-ordered copy, use the sync-free solve
    const int* col;     // positions (level order)
    const double* val;
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
    constexpr int kAhead = kLsStages - 2;  // tiles in flight; the stage issued into was released a whole tile ago
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
        // every tile is one item: thread 0 keeps kAhead tiles in flight with the lean producer
        pipe.tab = tab, pipe.ntiles = cnt;
        if (tid == 0)
            for (int t = 0; t < min(cnt, kAhead); ++t) pipe.issue_tile(t);
        // per-row metadata one tile ahead: original row, level, row extent
        int orig_n = -1, lvl_n = -1, rs_n = 0, re_n = 0;
        auto load_meta = [&](int tile) {
            const int r = tile * kTileRows + tid;
            orig_n = -1, lvl_n = -1, rs_n = 0, re_n = 0;
            if (r < n) {
                orig_n = __ldg(F.perm + r), lvl_n = __ldg(F.lvl + r);
                rs_n = __ldg(F.rowptr + r), re_n = __ldg(F.rowptr + r + 1);
            }
        };
        load_meta(ga);
        for (int i = 0; i < cnt; ++i) {
            const TileDesc& d = tab[i];
            LS_TRACE(ga + i, 0);
            const int orig = orig_n, lvl = lvl_n, rs = rs_n, re = re_n;
            const double bi = orig >= 0 ? __ldcg(rhs + orig) : 0.0;
            if (i + 1 < cnt) load_meta(ga + i + 1);
            LS_TRACE(ga + i, 1);
            if (tid == 0 && i + kAhead < cnt) pipe.issue_tile(i + kAhead);
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
                w[u] = 0, v[u] = 0.0;  // an absent dependency reads slot 0 with coefficient 0 ... and is skipped below
                if (u < ndep) w[u] = sc[e + u] & (kLsWindow - 1), v[u] = sv[e + u];
            }
            if (orig >= 0) rcp = __ddiv_rn(1.0, sv[(kUpper ? rs : re - 1) - as]);
            const int slot = (d.ltile * kTileRows + tid) & (kLsWindow - 1);
            const int lv1 = d.nnz;
            double xsol = 0.0;
            LS_TRACE(ga + i, 4);
            g_ls_trace_levels(ga + i, lv1 - d.n + 1);
#pragma unroll 1
            for (int l = d.n; l <= lv1; ++l) {
                if (lvl == l) {
                    double sum = 0.0;
#pragma unroll
                    for (int u = 0; u < kDeps; ++u)
                        if (u < ndep) sum = __dadd_rn(sum, __dmul_rn(v[u], win[w[u]]));
                    xsol = __dmul_rn(__dsub_rn(bi, sum), rcp);
                    win[slot] = xsol;
                }
                __syncthreads();
            }
            // the global copy (original numbering) leaves after the level loop
            LS_TRACE(ga + i, 5);
            if (orig >= 0) x[orig] = xsol;
            pipe.release();
            LS_TRACE(ga + i, 6);
        }
    }
    __syncthreads();
    if (tid == 0) ls.items = pipe.c_count;
    __syncthreads();
}

}  // namespace dp
