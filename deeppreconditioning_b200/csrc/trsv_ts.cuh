// trsv_ts.cuh — K4, third algorithm: the TILE-STREAM triangular solve for BATCHES of factors with WIDE levels.
//
// No reference counterpart; same arithmetic as sptrsv.cuh / trsv_ls.cuh (plain substitution, sum in column
// order, products and sums rounded separately, multiplication by the stored reciprocal 1 / T_ii), so the solution is
// bit-identical to oracle_sptrsv_lower/upper.
//
// Why a third algorithm. The sync-free solve (sptrsv.cuh) reads the factor in its ORIGINAL order, one lane per row of
// a level: on a 3-D stencil a level is a diagonal plane i+j+k = l, its rows are n-1 apart, so every lane of a warp
// touches its own sectors of rowptr/col/val/b (each fetched again for the neighbouring rows of later levels) and the
// stream is a chain of dependent scattered loads: 0.13 of the HBM peak on 16 x 128^3. The level-stream solve
// (trsv_ls.cuh) streams a level-ordered copy but keeps a system inside ONE SM - right for 2-D factors whose levels
// are narrower than a tile, hopeless for levels of 10^4..10^5 rows. Here the two are combined:
//
//   * the factor is read from its LEVEL-ORDERED copy (dp_sptrsv_permute: row r = row perm[r] of T, columns renumbered
//     to positions, diagonal stored as reciprocal). A 512-row TILE of the copy is one contiguous span of col/val, moved
//     by the TMA engine through the tile pipeline (tilepipe.cuh), one thread per row - the matrix stream is as
//     coalesced as an SpMV's;
//   * the tiles of ALL systems of the batch form one global sequence (tile-major, systems interleaved) that the
//     persistent CTAs take round robin. All CTAs are co-resident (cooperative launch) and a dependency always lives at
//     a lower position of the same system, i.e. in an earlier item of the sequence: forward progress by induction;
//   * dependencies are awaited on the solution vector in POSITION space (xp, pre-armed with the kPending NaN pattern:
//     data is its own flag, one store per row, no fences - the protocol of sptrsv.cuh). Consecutive rows of a level
//     depend on nearly consecutive positions of the previous level, so a warp's polls are a few sectors wide;
//   * with levels wider than the window of tiles in flight (G / nsys tiles per system) a tile's dependencies were
//     solved long before it starts: no waiting at all, the solve streams at SpMV speed. Narrow levels (the corners
//     of the diagonal sweep) pay one L2 hop per level, shared by all systems of the batch.
// Everything a tile needs that is contiguous travels with its first pipeline item: the 513 row pointers and the 512
// right-hand sides. Nothing is prefetched through registers, so the only latency a warp sees per tile is the poll of
// its dependencies (measured before this: the row extents fetched one tile ahead through registers cost a DRAM
// latency per tile, 3 us per tile and CTA, 0.52 of peak).
// The kernel works in POSITION SPACE only: b and x are indexed by position, x doubles as the polled vector. A caller
// that keeps its vectors in level order (dp_trsv_ls_system_t.perm == NULL, precond.LevelOrdering) gets exactly that.
// For vectors in the original numbering the host side (sptrsv.cu) brackets the solve with a gather b_pos = b[perm]
// and a scatter x[perm] = x_pos, one system after the other: a row's b / x sector is shared by rows of 4 consecutive
// levels, which survives in L2 inside one system's pass but not inside the batch solve, where the level fronts of all
// systems compete for it (measured with the gather/scatter inside the solve: 8 x 256^3 0.30 of peak against 0.49).
// REVERSED position space (rev): position p is row n - 1 - p of b and x. This is the backward solve (L^T) of a system
// that is kept in the level order of its FORWARD solve: walking its rows from the last to the first is a valid order
// for L^T (every dependency of row i is a row j > i), and a row's dependencies sit one forward level further on, as
// far away as in the forward solve - so one ordering of the system serves both solves of a PCG iteration.
#pragma once

#include "sptrsv.cuh"
#include "tilepipe.cuh"
#include "trsv_ls.cuh"

#ifndef DPCG_TS_CAP
#define DPCG_TS_CAP 2048
#endif
#ifndef DPCG_TS_STAGES
#define DPCG_TS_STAGES 3
#endif
#ifndef DPCG_TS_SLEEP
#define DPCG_TS_SLEEP 32
#endif
#ifndef DPCG_TS_ROUND
#define DPCG_TS_ROUND 96
#endif

namespace dp {

// Timeline of the short-row kernel (experiments, -DDPCG_TS_TRACE): thread 0 of CTA 0 and of the middle CTA record
// clock64 at six points of each of their first kTsTraceTiles tiles (look: before / after the stage wait, loads issued;
// finish: coefficients read, first try done, tile done); read back with dp_debug_ts_trace.
#ifdef DPCG_TS_TRACE
constexpr int kTsTraceTiles = 1024;
__device__ long long g_ts_trace[2 * kTsTraceTiles * 8];
#define TS_TRACE_AT(item, slot)                                                                          \
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && (item) < kTsTraceTiles)  \
    g_ts_trace[((blockIdx.x ? 1 : 0) * kTsTraceTiles + (item)) * 8 + (slot)] = clock64()
#else
#define TS_TRACE_AT(item, slot)
#endif

constexpr int kTsCap = DPCG_TS_CAP;        // entries per stage: a 512-row tile of a 7-point factor (4 per row) is one item
constexpr int kTsStages = DPCG_TS_STAGES;
constexpr int kTsRound = DPCG_TS_ROUND;    // tile descriptors per table refill
constexpr int kTsInflight = 4;             // dependencies polled per lane and round
#ifndef DPCG_TS_FAST
#define DPCG_TS_FAST 3
#endif
constexpr int kTsFast = DPCG_TS_FAST;      // register path: rows with at most this many dependencies (5-/7-point factors: 2 / 3)
constexpr int kTsSlots = kTsCap + 8;       // up to 3 lead-in entries (16-byte alignment) + tail rounding
static_assert(kTsCap % 4 == 0 && kTsCap >= 64, "stage capacity");

struct TsSysDev {
    LsFactor F;       // level-ordered copy (perm / lvl unused here)
    const double* b;  // position space
    double* x;        // position space: solution AND polled vector, armed with kPending before the launch
    int upper, ntiles;
    int rev, pad;     // position p is row n - 1 - p of b / x (see above)
    const int* skip;  // optional device flag, read at launch: non-zero = leave this system out (PCG: it has converged)
};

// One 512-row tile of one system's level-ordered copy. rowptr == nullptr: a system with fewer tiles, nothing to do.
struct TsTile {
    const int* rowptr;
    const int* col;
    const double* val;
    const double* b;
    double* x;
    int n, cs, ce, ltile, upper, rev;
};

struct TsStage {
    double val[kTsSlots];
    int col[kTsSlots];
    double b[kTileRows + 2];      // (reversed slices start at an even row: one lead-in entry)
    int rowptr[kTileRows + 8];    // kTileRows + 1 used
};
static_assert(sizeof(TsStage) % 16 == 0 && (kTsSlots * 8) % 16 == 0 && (kTsSlots * 4) % 16 == 0, "bulk-copy alignment");

struct TsSmem {
    alignas(16) TsStage stage[kTsStages];
    alignas(8) unsigned long long full[kTsStages];   // producer -> consumers: bytes have landed
    alignas(8) unsigned long long empty[kTsStages];  // consumers -> producer: all 16 warps are done with the stage
    TsTile tab[kTsRound];
};

__device__ __forceinline__ int ts_blocks(const TsTile& d) { return d.rowptr ? (d.ce - d.cs + kTsCap - 1) / kTsCap : 0; }

// The pipeline of tilepipe.cuh with two changes: the tile's metadata rides on its first item, and the producer is a
// WARP OF ITS OWN (warp kWarpsPerBlock of a kTsThreads-thread CTA; lane 0 issues). Round 1 let the warp that released a
// stage last re-arm it; the per-tile timeline (tools/trace_ts.py, profiles/r2/ts_trace_r1_kernel.log) showed what that
// costs: under load the four bulk copies of an item block their issuing thread for ~2000 cycles (TMA back-pressure), the
// warp that paid them was late for its next tile, hence last again, and its serial path - bytes, stage -> registers, issue,
// poll - became the period of the whole CTA (4700 cycles per 512 rows). The producer warp takes the issue off every
// consumer's path. Items are numbered since kernel start: item i lives in stage i % kTsStages, its (i / kTsStages)-th use.
// ... and FOUR of them, one per array of an item: in steady state every mbarrier.arrive.expect_tx and every cp.async.bulk
// costs its issuing thread ~255 ns whatever the size of the copy (tools/microbench/tma_stream.cu,
// profiles/r2/tma_stream.log: an item of 1 / 2 / 4 copies takes 510 / 765 / 1280 ns from one thread, with 3 or 7 stages
// alike), so one producer could not issue a 4-array item faster than every 1.28 us - 24 KB per 1.28 us and CTA is below
// the HBM share of an SM. Issued from four warps the four copies overlap.
constexpr int kTsProducers = 4;
constexpr int kTsThreads = kBlock + kTsProducers * kWarp;  // 16 consumer warps (one thread per row of a tile) + producers

struct TsPipe {
    TsSmem* sm;
    int ntiles;
    unsigned c_count;  // consumers: items this warp has released; producer lane: items issued
    unsigned a_count;  // consumers: items this warp has acquired (the short-row kernel holds two at a time)

    __device__ __forceinline__ void init(TsSmem* s) {
        sm = s, ntiles = 0, c_count = 0u, a_count = 0u;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < kTsStages; ++k) {
                mbar_init(&sm->full[k], (unsigned)kTsProducers);
                mbar_init(&sm->empty[k], (unsigned)kWarpsPerBlock);
            }
            mbar_fence_init();
        }
        __syncthreads();
    }
    // (t, j) -> the following item of the round; false at its end. Start from (0, -1).
    __device__ __forceinline__ bool next_item(int& t, int& j) const {
        ++j;
        while (t < ntiles && j >= ts_blocks(sm->tab[t])) ++t, j = 0;
        return t < ntiles;
    }
    // Producer warp `role` (0..3 = values, columns, row pointers, right-hand sides; lane 0 acts): arm `stage` (free: every
    // consumer warp has released its previous item) with this role's array of block j of tile t. The stage's `full`
    // barrier counts kTsProducers arrivals: a role without a copy for this item (blocks j > 0 carry no metadata) just
    // arrives.
    __device__ __forceinline__ void issue(int role, unsigned stage, int t, int j) {
        const TsTile& d = sm->tab[t];
        TsStage& st = sm->stage[stage];
        const int bs = d.cs + j * kTsCap;
        const int be = min(d.ce, bs + kTsCap);
        const int as = bs & ~3;
        unsigned long long* bar = &sm->full[stage];
        const unsigned long long pol = l2_policy_stream();
        if (role == 0) {
            const unsigned nval = (unsigned)(((be + 1) & ~1) - as);
            mbar_arrive_expect_tx(bar, nval * 8u);
            bulk_g2s(st.val, d.val + as, nval * 8u, bar, pol);
        } else if (role == 1) {
            const unsigned ncol = (unsigned)(((be + 3) & ~3) - as);
            mbar_arrive_expect_tx(bar, ncol * 4u);
            bulk_g2s(st.col, d.col + as, ncol * 4u, bar, pol);
        } else if (j > 0) {
            mbar_arrive(bar);
        } else {
            const int r0 = d.ltile * kTileRows;
            const unsigned nr = (unsigned)min(kTileRows, d.n - r0);
            if (role == 2) {
                const unsigned rp_bytes = ((nr + 1u) * 4u + 15u) & ~15u;
                mbar_arrive_expect_tx(bar, rp_bytes);
                bulk_g2s(st.rowptr, d.rowptr + r0, rp_bytes, bar, pol);
            } else {
                // rows [r0, r0 + nr), or - reversed - rows [n - r0 - nr, n - r0) from the even row below
                const int b0 = d.rev ? (d.n - r0 - (int)nr) & ~1 : r0;
                const unsigned b_bytes = ((unsigned)((d.rev ? d.n - r0 : r0 + (int)nr) - b0) * 8u + 15u) & ~15u;
                mbar_arrive_expect_tx(bar, b_bytes);
                bulk_g2s(st.b, d.b + b0, b_bytes, bar, pol);
            }
        }
    }
    // Lane 0 of producer warp `role`, once per round (after the round's table is complete and visible): this role's part
    // of every item of the round, in order, each as soon as its stage has been released by all consumer warps.
    __device__ __forceinline__ void produce(int role, int count) {
        ntiles = count;
        int t = 0, j = -1;
        while (next_item(t, j)) {
            const unsigned stage = c_count % kTsStages, use = c_count / kTsStages;
            if (use > 0) {
                while (!mbar_try_wait_hint(&sm->empty[stage], (use - 1u) & 1u, 2000u)) {
                }
            }
            issue(role, stage, t, j);
            ++c_count;
        }
    }
    __device__ __forceinline__ void begin(int count) { ntiles = count; }  // consumers
    __device__ __forceinline__ unsigned acquire() {
        const unsigned stage = a_count % kTsStages;
        // suspend instead of spinning: warps that wait for bytes must not take issue slots from the warps that work
        while (!mbar_try_wait_hint(&sm->full[stage], (a_count / kTsStages) & 1u, 2000u)) {
        }
        ++a_count;
        return stage;
    }
    // Wait for the bytes of the item after the one just released, without taking it (acquire() then returns at once).
    __device__ __forceinline__ void await_next() {
        const unsigned stage = a_count % kTsStages;
        while (!mbar_try_wait_hint(&sm->full[stage], (a_count / kTsStages) & 1u, 2000u)) {
        }
    }
    // Hand back the stage of the oldest item this warp still holds.
    __device__ __forceinline__ void release(int, int) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&sm->empty[c_count % kTsStages]);
        ++c_count;
    }
};

// One tile whose rows have at most kTsFast dependencies each (so the tile is ONE pipeline item): the row moves to
// registers (dependency positions, coefficients, 1 / diagonal), the STAGE IS HANDED BACK BEFORE ANY POLL, and the row is
// finished from registers. The stage is held for a few shared-memory reads instead of an L2 round trip (or a whole
// level of waiting), so its next item is on its way while this tile still polls; where all dependencies are long
// solved (levels wider than the batch's window of tiles in flight) the code is straight-line. A kernel of its own
// (kShort): sharing registers with the general loop below spilled, and the spill traffic cost more than the early
// hand-back won (profiles/README.md).
// `more`: another tile follows in this round. Its bytes are awaited WHILE this tile's dependency loads are in flight
// (they are an L2 round trip of 1000-1600 cycles under load, tools/trace_ts.py): the two longest waits of a warp's path
// through a tile overlap, at no cost in registers.
// Measured and rejected in round 2 (profiles/r2/README.md): a full look/finish software pipeline over two tiles (the
// stage must then be held across the next tile's look, which leaves one of three stages for the stream: 0.67 -> 0.65
// on 8 x 256^3) and conflict-free rotated stage reads (the 4-way bank conflicts of 4-entry rows are real - 53 % of the
// wavefronts - but the phase is latency bound: no gain).
__device__ __forceinline__ void ts_tile_short(const TsTile& d, bool more, TsSmem& sm, TsPipe& pipe, const AbortCtl& ctl, bool& dead) {
    const int tid = threadIdx.x;
    double* xp = d.x;
    const int r = d.ltile * kTileRows + tid;
    const bool valid = r < d.n;
    const int xm = d.rev ? -1 : 1, xo = d.rev ? d.n - 1 : 0;
    const int as = d.cs & ~3;
    int m = 0;
    double bi = 0.0, rcp = 0.0;
    int c[kTsFast];
    double v[kTsFast];
#pragma unroll
    for (int k = 0; k < kTsFast; ++k) c[k] = 0, v[k] = 0.0;
    const unsigned trace_item = pipe.c_count;
    (void)trace_item;
    TS_TRACE_AT(trace_item, 0);
    const unsigned stage = pipe.acquire();
    TS_TRACE_AT(trace_item, 1);
    const TsStage& st = sm.stage[stage];
    bool done = !valid || dead;
    bool bad = false;  // the caller's DP_TRSV_SHORT_ROWS promise does not hold for this row
    if (valid) {
        const int rs = st.rowptr[tid], re = st.rowptr[tid + 1];
        const int q = (d.upper ? rs + 1 : rs) - as;  // dependencies: stage entries [q, q + m)
        m = re - rs - 1;
        bad = m > kTsFast || m < 0 || re - as > kTsSlots;
        if (bad) m = 0;
        else rcp = st.val[(d.upper ? rs : re - 1) - as];  // the diagonal's entry holds 1 / T_ii
        const int r0 = d.ltile * kTileRows;
        bi = d.rev ? st.b[(d.n - 1 - r) - ((d.n - r0 - min(kTileRows, d.n - r0)) & ~1)] : st.b[tid];
#pragma unroll
        for (int k = 0; k < kTsFast; ++k)
            if (k < m) c[k] = xo + xm * st.col[q + k], v[k] = st.val[q + k];
    }
    if (bad) ctl.raise(DP_ERR_STRUCTURE), dead = true;
    dead = __any_sync(kFull, dead);
    if (dead) done = true;
    TS_TRACE_AT(trace_item, 2);
    pipe.release(0, 0);
    TS_TRACE_AT(trace_item, 3);
    unsigned long long u[kTsFast];
#pragma unroll
    for (int k = 0; k < kTsFast; ++k) {
        u[k] = 0ull;
        if (!done && k < m) u[k] = ld_relaxed_u64(xp + c[k]);
    }
    if (more) pipe.await_next();  // the next tile's bytes land while the loads above are in flight
    // finish the row from one snapshot of its dependencies; publishes at once (rows of the same warp may wait for it)
    auto try_finish = [&](const unsigned long long (&w)[kTsFast]) {
        bool ready = true;
#pragma unroll
        for (int k = 0; k < kTsFast; ++k) ready = ready && (k >= m || w[k] != kPending);
        if (ready) {
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < kTsFast; ++k)
                if (k < m) sum = __dadd_rn(sum, __dmul_rn(v[k], as_double(w[k])));
            const double xv = __dmul_rn(__dsub_rn(bi, sum), rcp);
            st_relaxed_u64(xp + (xo + xm * r), as_bits(xv));
            done = true;
        }
    };
    auto repoll = [&](unsigned long long (&w)[kTsFast]) {
#pragma unroll
        for (int k = 0; k < kTsFast; ++k)
            if (k < m && w[k] == kPending) w[k] = ld_relaxed_u64(xp + c[k]);
    };
    if (!done) try_finish(u);
    TS_TRACE_AT(trace_item, 4);
    if (__all_sync(kFull, done)) {  // the streaming regime: every dependency was solved long ago
        TS_TRACE_AT(trace_item, 5);
        return;
    }
    // Waiting for a level: one poll per dependency in flight. Two polls half a round trip apart were measured SLOWER
    // (0.97 -> 1.05 us per level, profiles/r1/ts_two_polls_negative.log): the L2 traffic of the pollers is part of the hop.
    for (unsigned idle = 0;;) {
        if (!done) repoll(u);
        if (!done) try_finish(u);
        if (__all_sync(kFull, done)) break;
        ++idle;
        if (idle > kSpinBudget) ctl.raise(DP_ERR_TIMEOUT), dead = true;
        if ((idle & 255u) == 0 && ctl.aborted()) dead = true;
        dead = __any_sync(kFull, dead);
        if (dead) done = true;
#if DPCG_TS_SLEEP > 0
        else __nanosleep(DPCG_TS_SLEEP);
#endif
    }
    TS_TRACE_AT(trace_item, 5);
}

// One tile of any shape: rows of any length, tiles of several pipeline items, rows cut by an item boundary.
__device__ __forceinline__ void ts_tile_general(const TsTile& d, int i, TsSmem& sm, TsPipe& pipe, const AbortCtl& ctl, bool& dead) {
    const int tid = threadIdx.x;
    double* xp = d.x;
    const bool upper = d.upper != 0;
    const int cs = d.cs, ce = d.ce;
    const int r = d.ltile * kTileRows + tid;
    const bool valid = r < d.n;
    // where position p lives in xp: p, or n - 1 - p when the caller's vectors are in the REVERSE of this
    // factor's position order (the backward solve of a system kept in the forward solve's level order)
    const int xm = d.rev ? -1 : 1, xo = d.rev ? d.n - 1 : 0;
    int dpos = 0, q = 0, end = 0;
    double bi = 0.0;
    bool done = !valid || dead, have_rcp = false;
    double sum = 0.0, rcp = 0.0;
    unsigned long long u[kTsInflight];
#pragma unroll
    for (int k = 0; k < kTsInflight; ++k) u[k] = kPending;
    const int nb = (ce - cs + kTsCap - 1) / kTsCap;
    for (int j = 0; j < nb; ++j) {
        const int bs = cs + j * kTsCap;
        const int be = min(ce, bs + kTsCap);
        const int as = bs & ~3;
        const unsigned stage = pipe.acquire();
        const TsStage& st = sm.stage[stage];
        if (j == 0 && valid) {  // the tile's metadata rides on its first item
            const int rs = st.rowptr[tid], re = st.rowptr[tid + 1];
            dpos = upper ? rs : re - 1;   // the diagonal's entry (holds 1 / T_ii)
            q = upper ? rs + 1 : rs;      // dependencies: entries [q, end)
            end = upper ? re : re - 1;
            const int r0 = d.ltile * kTileRows;
            bi = d.rev ? st.b[(d.n - 1 - r) - ((d.n - r0 - min(kTileRows, d.n - r0)) & ~1)] : st.b[tid];
        }
        const double* __restrict__ sv = st.val;
        const int* __restrict__ sc = st.col;
        if (!done && !have_rcp && dpos >= bs && dpos < be) rcp = sv[dpos - as], have_rcp = true;
        const int qe = min(end, be);  // this row's dependencies inside the block: [q, qe) (empty if q >= be)
        unsigned idle = 0;
        for (;;) {
            if (!done && q == end && have_rcp) {  // publish at once: rows of the same warp may wait for it
                const double xv = __dmul_rn(__dsub_rn(bi, sum), rcp);
                st_relaxed_u64(xp + (xo + xm * r), as_bits(xv));
                done = true;
            }
            const bool pending = !done && q < qe;
            if (!__any_sync(kFull, pending)) break;
            bool progress = false;
            if (pending) {
                const int m = min(kTsInflight, qe - q);
                bool all = true;
#pragma unroll
                for (int k = 0; k < kTsInflight; ++k) {
                    if (k < m && u[k] == kPending) {
                        u[k] = ld_relaxed_u64(xp + (xo + xm * sc[q + k - as]));
                        all = all && u[k] != kPending;
                    }
                }
                if (all) {
#pragma unroll
                    for (int k = 0; k < kTsInflight; ++k) {
                        if (k < m) sum = __dadd_rn(sum, __dmul_rn(sv[q + k - as], as_double(u[k])));
                        u[k] = kPending;
                    }
                    q += m;
                    progress = true;
                }
            }
            if (!__any_sync(kFull, progress)) {
                ++idle;
                if (idle > kSpinBudget) ctl.raise(DP_ERR_TIMEOUT), dead = true;
                if ((idle & 255u) == 0 && ctl.aborted()) dead = true;
                dead = __any_sync(kFull, dead);
                if (dead) done = true;
#if DPCG_TS_SLEEP > 0
                else __nanosleep(DPCG_TS_SLEEP);
#endif
            }
        }
        pipe.release(i, j);
    }
}

// The whole batch. All kTsThreads threads of every CTA of a cooperative launch must call. kShort: every system of the
// batch carries DP_TRSV_SHORT_ROWS (no row with more than kTsFast dependencies; violations raise DP_ERR_STRUCTURE).
template <bool kShort>
__device__ __forceinline__ void trsv_tile_stream(const TsSysDev* __restrict__ sys, int nsys, int max_tiles, TsSmem& sm,
                                                 const AbortCtl& ctl) {
    TsPipe pipe;
    pipe.init(&sm);
    const int tid = threadIdx.x;
    const bool producer = tid >= kBlock;  // the last kTsProducers warps
    const int G = gridDim.x;
    const int items = max_tiles * nsys;  // < 2^31: checked by the host
    const int mine = items > (int)blockIdx.x ? (items - (int)blockIdx.x + G - 1) / G : 0;
    bool dead = false;  // the solve was aborted: keep the pipeline moving (every item is acquired and released), solve nothing
    for (int j0 = 0; j0 < mine; j0 += kTsRound) {
        const int cnt = min(kTsRound, mine - j0);
        __syncthreads();  // every warp has consumed every item of the previous round: its table is free
        for (int i = tid; i < cnt; i += kTsThreads) {
            const long long g = blockIdx.x + (long long)(j0 + i) * G;
            const int s = (int)(g % nsys), t = (int)(g / nsys);
            const TsSysDev S = sys[s];
            TsTile d;
            d.rowptr = nullptr, d.col = S.F.col, d.val = S.F.val, d.b = S.b, d.x = S.x;
            d.n = S.F.n, d.cs = 0, d.ce = 0, d.ltile = t, d.upper = S.upper, d.rev = S.rev;
            if (t < S.ntiles && !(S.skip && ld_relaxed_s32(S.skip) != 0)) {  // (the flag does not change during the launch)
                d.rowptr = S.F.rowptr;
                d.cs = __ldg(S.F.rowptr + min(t * kTileRows, S.F.n));
                d.ce = __ldg(S.F.rowptr + min((t + 1) * kTileRows, S.F.n));
                if (kShort && d.ce - d.cs > kTsCap) d.ce = d.cs + kTsCap;  // keeps the tile ONE item; reported by ts_tile_short
            }
            sm.tab[i] = d;
        }
        __syncthreads();
        if (producer) {
            if ((tid & 31) == 0) pipe.produce((tid - kBlock) >> 5, cnt);
            continue;
        }
        pipe.begin(cnt);
        if (kShort) {
            int nxt = 0;  // next tile of the round that has rows
            while (nxt < cnt && !sm.tab[nxt].rowptr) ++nxt;
            while (nxt < cnt) {
                const int i = nxt;
                for (++nxt; nxt < cnt && !sm.tab[nxt].rowptr; ++nxt) {
                }
                ts_tile_short(sm.tab[i], nxt < cnt, sm, pipe, ctl, dead);
            }
        } else {
            for (int i = 0; i < cnt; ++i) {
                const TsTile& d = sm.tab[i];
                if (!d.rowptr) continue;
                ts_tile_general(d, i, sm, pipe, ctl, dead);
            }
        }
    }
    __syncthreads();
}

// Host side (sptrsv.cu): arm the polled vectors of a device-resident descriptor array (`arm`; a caller whose own
// kernels leave them armed passes false) and launch the solve. `word` (8 bytes, zeroed by the caller once) carries the
// abort bit. No host synchronisation.
int ts_solve_launch(const TsSysDev* sys_dev, int nsys, int max_tiles, int nmax, bool short_rows, bool arm,
                    unsigned long long* word, int* flag, cudaStream_t s);

}  // namespace dp
