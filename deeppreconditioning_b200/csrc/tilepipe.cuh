// tilepipe.cuh — the CSR tile pipeline (K2 primitive): TMA-staged, lane-per-row SpMV over 512-row tiles.
//
// Replaces `A @ p` / `M @ r` (cg.py:60,61,75,81) for the standalone SpMV and for every SpMV-shaped PCG phase.
//
// A tile = 512 consecutive rows of one CSR matrix = ONE contiguous span of col[]/val[]. The span is cut into blocks
// of at most kPipeCap entries; each block is an ITEM of the pipeline: lane 0 of warp 0 arms the stage's `full`
// mbarrier with the byte count and issues two bulk async copies (cp.async.bulk, SASS UBLKCP: the TMA engine streams
// col[] and val[] into a shared-memory stage), up to kPipeStages items ahead, so the HBM stream never waits for the
// arithmetic. Every warp passes through every item on its own: wait on `full`, walk the part of ITS 32 rows that
// lies in the block, arrive on the stage's `empty` mbarrier (16 arrivals re-arm the stage). Warps are not coupled by
// CTA barriers while streaming, so different warps work on different blocks of a tile at the same time.
// Inside a stage every lane walks its own row:
//     sum = sum + val[q] * x[col[q]]        q ascending, product and sum rounded separately (no FMA)
// which is bit-identical to the sequential CPU loop (scipy csr_matvec / oracle_spmv_csr). Reading the stage at a
// stride of one row length is bank-conflict free for odd row lengths (5/7-point stencils, the CNN's 15/row), and -
// what matters most - the 32 gather addresses of one load instruction are 32 CONSECUTIVE rows' k-th neighbours, i.e.
// 2-3 cache lines for stencil-like matrices instead of ~10 for an entry-major sweep: the kernel stays HBM-bound
// instead of L1-wavefront-bound. No product staging, no shared-memory stores at all.
//
// Bulk copies need 16-byte aligned addresses and sizes: a block's copy starts at the 16-byte boundary at or below its
// first entry and ends at the one at or above its last. The few bytes read beyond the matrix's last entry stay inside
// the 16-byte unit that holds that entry (arrays are 16-byte aligned, dpcg.h), so they can never touch another page.
//
// Packed stream (kPacked): the same pipeline over the lossless 6-byte copy of a matrix (dp_csr_pack: fp32 values, 16-bit
// columns relative to the tile's smallest column). The lanes widen on load (exact) and run the identical fp64 recurrence:
// same bits, half the matrix bytes from HBM; a stage of the same bytes holds twice the entries (the CNN factor's tile is one
// item). When the CTA carries a warp beyond its kBlock row threads, that PRODUCER warp is the issuer: it walks the same
// tiles, reads nothing, and sends every item the moment its stage has been handed back (pcg.cu, packed engine).
// Experiment (DPCG_PACK_WINDOWS=1, off): half of the stage bytes hold gather WINDOWS - the columns [base, base + span) a
// banded tile touches (~1150 for 512 rows of a 316-wide 2-D factor), copied once per tile into shared memory by bulk copies,
// one tile ahead, so that the gathers become shared-memory loads. Same values, same arithmetic, same bits; the row loops
// halve, but the stages do too, and the time comes back as waiting for bytes (profiles/r2/pack_experiments.md).
#pragma once

#include <type_traits>

#include "common.cuh"
#include "spmv.cuh"

#ifndef DPCG_PIPE_CAP
#define DPCG_PIPE_CAP 3840
#endif
#ifndef DPCG_PIPE_STAGES
#define DPCG_PIPE_STAGES 2
#endif
#ifndef DPCG_PIPE_UNROLL
#define DPCG_PIPE_UNROLL 4
#endif
#ifndef DPCG_PIPE_EVICT_FIRST
#define DPCG_PIPE_EVICT_FIRST 1
#endif
#ifndef DPCG_PACK_WINDOWS
#define DPCG_PACK_WINDOWS 0  // experiment: gather windows in shared memory behind (half-size) packed stages
#endif
#ifndef DPCG_PACK_CAP
#if DPCG_PACK_WINDOWS
#define DPCG_PACK_CAP DPCG_PIPE_CAP        // entries per stage of the packed stream: half the bytes, the rest holds the windows
#else
#define DPCG_PACK_CAP (2 * DPCG_PIPE_CAP)  // ... the same bytes per stage as the fp64 stream: twice the entries
#endif
#endif
#ifndef DPCG_PACK_STAGES
#define DPCG_PACK_STAGES DPCG_PIPE_STAGES
#endif
#ifndef DPCG_PACK_PRODUCER_WARP
#define DPCG_PACK_PRODUCER_WARP 1  // the packed PCG engine runs a 17th warp that only issues the stages' copies
#endif
#ifndef DPCG_GHOST_PREFETCH
#define DPCG_GHOST_PREFETCH 1  // the producer warp prefetches the rows' own vector lines of the tiles it passes: 0 off, 1 L2, 2 L1
#endif
#ifndef DPCG_GHOST_PREFETCH_GATHER
#define DPCG_GHOST_PREFETCH_GATHER 1  // ... and the lines of the columns a banded tile gathers from (TileDesc base / span)
#endif
#ifndef DPCG_PIPE_WAIT_HINT
#define DPCG_PIPE_WAIT_HINT 1000  // ns a consumer warp may sleep in try_wait on a stage's `full` barrier (0: spin)
#endif

namespace dp {

// -DDPCG_PIPE_TRACE: timeline of two warps of CTA 0 through the tile pipeline (label << 48 | clock), read back with
// dp_debug_pipe_trace. Not compiled into the shipped library.
#ifdef DPCG_PIPE_TRACE
constexpr int kPipeTraceCap = 1 << 15;
__device__ unsigned long long g_pipe_trace[2][kPipeTraceCap];
#define DP_PIPE_MARK(label) tr_mark(label)
#else
#define DP_PIPE_MARK(label) ((void)0)
#endif

constexpr int kPipeCap = DPCG_PIPE_CAP;        // entries per stage (multiple of 4)
constexpr int kPipeStages = DPCG_PIPE_STAGES;
constexpr int kPipeUnroll = DPCG_PIPE_UNROLL;  // gathers in flight per thread
constexpr int kPackCap = DPCG_PACK_CAP;
constexpr int kPackStages = DPCG_PACK_STAGES;
constexpr bool kPackWindows = DPCG_PACK_WINDOWS != 0;
constexpr bool kPackProducerWarp = DPCG_PACK_PRODUCER_WARP != 0;
constexpr int kMaxStages = kPipeStages > kPackStages ? kPipeStages : kPackStages;

// ---- PTX wrappers (sm_90+/sm_100a) ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (ns): a warp that waits for bytes sleeps in hardware instead of taking issue slots
// from the warps that work.
__device__ __forceinline__ bool mbar_try_wait_hint(unsigned long long* bar, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, unsigned parity) {  // never suspends
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`. 16-byte aligned addresses and size.
// The matrix stream is read once per SpMV: mark its lines evict-first so that the vectors (gathered with L2
// latency instead of DRAM latency when they survive) keep the cache.
__device__ __forceinline__ unsigned long long l2_policy_stream() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// A working set that fits the L2 (one small system: 55 MB against 126 MB) should stay there instead: the next iteration
// streams the same matrices again.
__device__ __forceinline__ unsigned long long l2_policy_keep() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                         unsigned long long policy) {
#if DPCG_PIPE_EVICT_FIRST
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
#endif
}
// 16-byte asynchronous copy global -> shared (LDGSTS) and its completion hook: the mbarrier receives one arrival from this
// thread once all its earlier cp.async have landed (.noinc: the arrival is part of the barrier's initial count).
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The producer warp runs up to kStages tiles ahead of the row warps: it asks for the 32 lines that hold the 512 rows of a
// vector in tile `tile` (one line per lane), so that the row warps' loads at the head of that tile find them in cache instead of
// paying an HBM round trip (tools/trace_pipe.py: 1.1 us at the head of every APPLY1 tile).
__device__ __forceinline__ void prefetch_tile_rows(const double* v, int tile, int n) {
#if DPCG_GHOST_PREFETCH
    const int row = tile * kTileRows + (int)(threadIdx.x & 31) * 16;
    if (row < n) {
#if DPCG_GHOST_PREFETCH == 1
        asm volatile("prefetch.global.L2 [%0];" ::"l"(v + row));
#else
        asm volatile("prefetch.global.L1 [%0];" ::"l"(v + row));
#endif
    }
#else
    (void)v, (void)tile, (void)n;
#endif
}
// ... and for the columns [base, base + span) a banded tile gathers from (span <= kWinCap doubles = 88 lines: three rounds).
__device__ __forceinline__ void prefetch_columns(const double* v, int base, int span) {
#if DPCG_GHOST_PREFETCH && DPCG_GHOST_PREFETCH_GATHER
    for (int c = (int)(threadIdx.x & 31) * 16; c < span; c += 32 * 16) {
#if DPCG_GHOST_PREFETCH == 1
        asm volatile("prefetch.global.L2 [%0];" ::"l"(v + (base & ~15) + c));
#else
        asm volatile("prefetch.global.L1 [%0];" ::"l"(v + (base & ~15) + c));
#endif
    }
#else
    (void)v, (void)base, (void)span;
#endif
}
// ---- descriptors ------------------------------------------------------------------------------------------------
struct TileDesc {
    const int* rowptr;  // nullptr: this tile streams nothing in this phase
    const int* col;
    const double* val;
    int n, nnz;  // rows / stored entries of the matrix
    int cs, ce;  // entries of the tile
    int ltile;   // tile index inside its matrix
    int sys;     // system id (PCG), unused by the standalone kernel
    int base;    // packed stream: smallest column of the tile (col = base + col16); col / val then point at col16 / val32
    int span;    // packed stream: doubles of the tile's gather window [base & ~1, ...) if it fits kWinCap, else 0 (no window)
};
constexpr int kWinCap = 1408;  // doubles per gather window (a 512-row tile of a 316-wide 2-D factor spans 1146 columns)
constexpr int kWinBufs = 2;    // the tile being consumed and the next one
constexpr int kWinVecs = 2;    // vectors a gather functor may combine

// Shared memory of a pipeline of geometry (kCap entries per stage, kStages stages): the stages' bytes (values first,
// then column indices) and one full/empty mbarrier pair per stage. Several geometries may be laid over the same bytes
// (each with its own barriers) as long as only one of them has items in flight at a time.
template <int kCap, int kStages, bool kPacked = false>
struct PipeGeom {
    // lead-in entries down to the 16-byte boundary of the span's first entry (3, packed: 7) + tail rounding
    static constexpr int kSlots = kCap + (kPacked ? 16 : 8);
    static constexpr int kEntryBytes = kPacked ? 6 : 12;
    static constexpr size_t kStageBytes = (size_t)kStages * kSlots * kEntryBytes;
    static constexpr size_t kWinBytes = (kPacked && kPackWindows) ? (size_t)kWinBufs * kWinVecs * kWinCap * 8 : 0;  // behind the stages
    static constexpr size_t kBytes = kStageBytes + kWinBytes;
    static_assert(kStageBytes % 16 == 0, "windows start 16-byte aligned");
    static_assert(kCap % 8 == 0 && kCap >= 64, "stage capacity");
};
template <int kStages>
struct PipeBarriers {
    alignas(8) unsigned long long full[kStages];   // producer -> consumers: bytes have landed
    alignas(8) unsigned long long empty[kStages];  // consumers -> producer: all 16 warps are done with the stage
    alignas(8) unsigned long long wfull[kWinBufs];   // gather windows (packed stream): every thread's copies have landed
    alignas(8) unsigned long long wempty[kWinBufs];  // ... all 16 warps are done with the window
};
// The level-stream triangular solve (trsv_ls.cuh) lays a finer geometry (stages of kLsCap entries: a stencil factor has 3
// entries per row) plus its shared-memory solution window over the same bytes (checked there).
constexpr int kLsCap = 1536;
// (3 level-stream stages: matrix 3 * 18 528, row pointers 3 * 2 080, right-hand sides 3 * 4 096, a window of 5 tiles)
constexpr size_t kLsFusedBytes = 3 * ((size_t)(kLsCap + 8) * 12 + (kTileRows + 8) * 4 + kTileRows * 8) + 5 * kTileRows * 8;
constexpr size_t max3(size_t a, size_t b, size_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
constexpr size_t kPipeRawBytes = max3(PipeGeom<kPipeCap, kPipeStages>::kBytes, PipeGeom<kPackCap, kPackStages, true>::kBytes, kLsFusedBytes);
struct PipeShared {
    alignas(16) unsigned char bytes[kPipeRawBytes];
    PipeBarriers<kMaxStages> bar;
};

// Fill the stream part of a descriptor for tile `ltile` of matrix M (one dependent load pair; CTA-parallel in the
// table builders).
template <bool kPacked = false>
__device__ __forceinline__ void tile_desc_fill(TileDesc& d, const CsrView& M, int ltile) {
    d.rowptr = M.rowptr;
    d.col = kPacked ? reinterpret_cast<const int*>(M.col16) : M.col;
    d.val = kPacked ? reinterpret_cast<const double*>(M.val32) : M.val;
    d.n = M.n;
    d.nnz = M.nnz;
    d.ltile = ltile;
    d.base = 0, d.span = 0;
    if (kPacked && M.rowptr) {
        const int2 info = __ldg(reinterpret_cast<const int2*>(M.tbase) + ltile);  // (smallest column, columns spanned)
        d.base = info.x;
        const int wlen = ((info.x + info.y + 1) & ~1) - (info.x & ~1);
        d.span = (info.y > 0 && wlen <= kWinCap) ? wlen : 0;
    }
    if (M.rowptr) {
        d.cs = __ldg(M.rowptr + min(ltile * kTileRows, M.n));
        d.ce = __ldg(M.rowptr + min((ltile + 1) * kTileRows, M.n));
    } else {
        d.cs = d.ce = 0;
    }
}

// Register state of a pipeline; lives for the whole kernel (or is saved/restored through `counts()`/`resume()`).
// Items are numbered since kernel start: item i lives in stage i % kStages and is the (i / kStages)-th use of that
// stage, which fixes the mbarrier parities.
template <int kCap, int kStages, bool kPacked = false>
struct PipeT {
    static constexpr int kSlots = PipeGeom<kCap, kStages, kPacked>::kSlots;
    static constexpr bool kIsPacked = kPacked;
    using ValT = typename std::conditional<kPacked, float, double>::type;
    using ColT = typename std::conditional<kPacked, unsigned short, int>::type;
    static constexpr int kValUnit = 16 / (int)sizeof(ValT);  // entries per 16-byte unit of the bulk copies
    static constexpr int kColUnit = 16 / (int)sizeof(ColT);
    static constexpr int kAlign = kColUnit > kValUnit ? kColUnit : kValUnit;  // a block's copies start at this entry boundary
    ValT* val0;  // stage s: val0 + s * kSlots, col0 + s * kSlots
    ColT* col0;
    unsigned long long* full;
    unsigned long long* empty;
    const TileDesc* tab;
    int ntiles;
    unsigned c_count;   // items this warp has consumed
    bool ghost;         // this thread belongs to the producer warp (beyond the kBlock row threads): it owns no row, issues the
                        // items as early as their stages allow and neither waits for bytes nor hands stages back
    int role;           // 0: this thread issues the items, -1: it does not
    unsigned p_count;   // items issued (meaningful in the issuer threads only, like the cursor below)
    int p_tile, p_blk;  // next item of the round to issue
    unsigned t_count;   // streaming tiles this warp has reduced (ring index of tile_reduce_async)
    int flip;           // tile_reduce scratch buffer in use next
    int early;          // 1 + id of the table whose first items are already in flight (begin_early), else 0
    int keep_l2;        // the matrices fit the L2 together with the vectors: do not mark their lines evict-first
    // gather windows (packed stream): buffer b, vector v at win0 + (b * kWinVecs + v) * kWinCap. Counters are CTA-uniform
    // (every thread walks the same tiles): windows filled / consumed since kernel start, buffer = count % kWinBufs.
#ifdef DPCG_PIPE_TRACE
    int tr_pos = 0;
    unsigned long long tr_phase = 0;
    __device__ __forceinline__ void tr_mark(unsigned long long label) {
        label += 8 * tr_phase;
        if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 160) && tr_pos < kPipeTraceCap) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            g_pipe_trace[threadIdx.x ? 1 : 0][tr_pos++] = (label << 48) | (t & 0xFFFFFFFFFFFFull);
        }
    }
#endif
    double* win0;
    unsigned long long* wfull;
    unsigned long long* wempty;
    unsigned w_issued, w_count;

    static __device__ __forceinline__ int blocks(const TileDesc& d) { return (d.ce - d.cs + kCap - 1) / kCap; }
    __device__ __forceinline__ const ValT* stage_val(unsigned s) const { return val0 + (size_t)s * kSlots; }
    __device__ __forceinline__ const ColT* stage_col(unsigned s) const { return col0 + (size_t)s * kSlots; }
    // the 16-byte aligned copies that bring entries [as, be) of a tile's arrays into a stage (as = bs & ~(kAlign - 1))
    __device__ __forceinline__ void issue_copies(const TileDesc& d, unsigned stage, int as, int be, unsigned long long pol) {
        const unsigned ncol = (unsigned)(((be + kColUnit - 1) & ~(kColUnit - 1)) - as);
        const unsigned nval = (unsigned)(((be + kValUnit - 1) & ~(kValUnit - 1)) - as);
        unsigned long long* bar = &full[stage];
        mbar_arrive_expect_tx(bar, ncol * (unsigned)sizeof(ColT) + nval * (unsigned)sizeof(ValT));
        bulk_g2s(val0 + (size_t)stage * kSlots, reinterpret_cast<const ValT*>(d.val) + as, nval * (unsigned)sizeof(ValT), bar, pol);
        bulk_g2s(col0 + (size_t)stage * kSlots, reinterpret_cast<const ColT*>(d.col) + as, ncol * (unsigned)sizeof(ColT), bar, pol);
    }

    // Bind to shared memory. `fresh`: initialise the barriers (once per kernel and geometry); ends with a CTA barrier.
    template <class Bars>
    __device__ __forceinline__ void init(unsigned char* bytes, Bars* bar, bool fresh = true) {
        val0 = reinterpret_cast<ValT*>(bytes);
        col0 = reinterpret_cast<ColT*>(bytes + (size_t)kStages * kSlots * sizeof(ValT));
        full = bar->full, empty = bar->empty;
        tab = nullptr;
        ntiles = 0;
        c_count = 0u, p_count = 0u, p_tile = 0, p_blk = 0, t_count = 0u, flip = 0, early = 0, keep_l2 = 0;
        ghost = threadIdx.x >= (unsigned)kBlock;
        // the issuer: lane 0 of the producer warp when the CTA has one, else thread 0 (which then also does its rows)
        role = threadIdx.x == (blockDim.x > (unsigned)kBlock ? (unsigned)kBlock : 0u) ? 0 : -1;
        win0 = reinterpret_cast<double*>(bytes + PipeGeom<kCap, kStages, kPacked>::kStageBytes);
        wfull = bar->wfull, wempty = bar->wempty;
        w_issued = 0u, w_count = 0u;
        if (fresh) {
            if (threadIdx.x == 0) {
#pragma unroll
                for (int s = 0; s < kStages; ++s) {
                    mbar_init(&full[s], 1u);
                    mbar_init(&empty[s], (unsigned)kWarpsPerBlock);
                }
                if (kPacked) {
#pragma unroll
                    for (int b = 0; b < kWinBufs; ++b) {
                        mbar_init(&wfull[b], 1u);
                        mbar_init(&wempty[b], (unsigned)kWarpsPerBlock);
                    }
                }
                mbar_fence_init();
            }
            __syncthreads();
        }
    }
    // A pipeline that is only used now and then inside a long kernel keeps nothing but its item count between uses
    // (all items consumed, producer and consumers agree): resume(count) after init(..., false).
    __device__ __forceinline__ void resume(unsigned items) { c_count = p_count = items; }

    // Issuer threads: issue the next item of the round if its stage is free. `blocking`: wait for the stage
    // (only legal when this warp has itself consumed the stage's previous item). Returns true if an item went out.
    // An item is three asynchronous operations - arm the stage's `full` barrier with the byte count, bulk copy of the values,
    // bulk copy of the columns. Inside the busy kernel this one-lane chain costs its warp ~0.86 us per item
    // (tools/trace_pipe.py), which is why the packed PCG engine gives it to a producer warp of its own; splitting the three
    // operations over three warps or dealing the items round robin among the row warps was measured and does not help
    // (profiles/r2/pack_experiments.md).
    __device__ __forceinline__ bool issue_one(bool blocking) {
        while (p_tile < ntiles) {
            if (p_blk < blocks(tab[p_tile])) break;
            ++p_tile, p_blk = 0;
        }
        if (p_tile >= ntiles) return false;
        if (p_count - c_count >= (unsigned)kStages) return false;  // this warp still owns the stage's previous item
        const unsigned stage = p_count % kStages, use = p_count / kStages;
        if (use > 0) {
            const unsigned par = (use - 1u) & 1u;
            if (blocking) {
                while (!mbar_try_wait(&empty[stage], par)) {
                }
            } else if (!mbar_test_wait(&empty[stage], par)) {
                return false;
            }
        }
        const TileDesc& d = tab[p_tile];
        const int bs = d.cs + p_blk * kCap;
        const int be = min(d.ce, bs + kCap);
        const int as = bs & ~(kAlign - 1);
        issue_copies(d, stage, as, be, keep_l2 ? l2_policy_keep() : l2_policy_stream());
        DP_PIPE_MARK(7);
        ++p_blk, ++p_count;
        return true;
    }

    // All threads, after the table `t` (shared memory) is complete and visible. Every warp has consumed every item of
    // the previous round (the caller's CTA barrier), so all stages are free.
    __device__ __forceinline__ void begin(const TileDesc* t, int count) {
        tab = t;
        ntiles = count;
        p_tile = 0, p_blk = 0;
        if (role >= 0) {
            while (issue_one(true)) {
            }
        }
    }

    // Lean producer for streams whose tiles are exactly one item each (the level-stream solve): thread 0 issues tile t
    // of the round into the next stage, waiting for the stage if it has been used before. No table walk.
    __device__ __forceinline__ void issue_tile(int t) {
        const unsigned stage = p_count % kStages, use = p_count / kStages;
        if (use > 0) {  // spin without suspending: the caller issues into a stage that was released a tile ago
            while (!mbar_test_wait(&empty[stage], (use - 1u) & 1u)) {
            }
        }
        const TileDesc& d = tab[t];
        issue_copies(d, stage, d.cs & ~(kAlign - 1), d.ce, l2_policy_stream());
        ++p_count;
    }
    // ... and its consumer side: wait for the next item without any producer duty.
    __device__ __forceinline__ unsigned wait_item() {
        const unsigned stage = c_count % kStages;
        while (!mbar_try_wait(&full[stage], (c_count / kStages) & 1u)) {
        }
        return stage;
    }

    // Start the NEXT round before the current phase's grid barrier: the matrices are immutable, so the first items of
    // the next phase's table can be on their way while the CTAs wait for each other. Every thread calls it after its
    // own last tile; thread 0 waits for the stages (the other warps release them as they finish). The next round must
    // open with begin_resume(id) instead of begin(), or be drained with drain_early().
    __device__ __forceinline__ void begin_early(const TileDesc* t, int count, int id) {
        begin(t, count);
        early = id + 1;
    }
    __device__ __forceinline__ bool begin_resume(const TileDesc* t, int count, int id) {
        if (early != id + 1) return false;
        early = 0;
        tab = t, ntiles = count;  // (already set by begin_early; kept for symmetry)
        return true;
    }
    // Consume what begin_early put in flight without using it (the kernel is about to exit, or the prediction of the
    // next table failed). begin() issues min(kStages, items of the table) items.
    __device__ __forceinline__ void drain_early() {
        if (!early) return;
        early = 0;
        int issued = 0;
        for (int t = 0; t < ntiles && issued < kStages; ++t) issued += blocks(tab[t]);
        issued = min(issued, kStages);
        if (role >= 0) p_tile = ntiles;  // nothing more to issue from that table
        for (int j = 0; j < issued; ++j) {
            const unsigned stage = c_count % kStages;
            while (!ghost && !mbar_try_wait(&full[stage], (c_count / kStages) & 1u)) {
            }
            release();
        }
    }

    // Consumer side of one item: acquire() waits until the item's bytes have landed and returns its stage; release()
    // hands the stage back once this warp has read what it needs. Every warp acquires and releases every item of the
    // round, in order. Thread 0 doubles as the producer: the item about to be consumed must be out, then it tops up.
    __device__ __forceinline__ unsigned acquire() {
        if (role >= 0) {
            while (p_count <= c_count) issue_one(true);
            while (issue_one(false)) {
            }
        }
        DP_PIPE_MARK(6);
        const unsigned stage = c_count % kStages;
        const unsigned par = (c_count / kStages) & 1u;
        if (ghost) return stage;  // the producer warp reads nothing
#if DPCG_PIPE_WAIT_HINT
        // (ncu, round 2: the plain try_wait spin was 13.5 % of the fused kernel's executed instructions)
        while (!mbar_try_wait_hint(&full[stage], par, DPCG_PIPE_WAIT_HINT)) {
        }
#else
        while (!mbar_try_wait(&full[stage], par)) {
        }
#endif
        return stage;
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0 && !ghost) mbar_arrive(&empty[c_count % kStages]);
        ++c_count;
    }

    // ---- gather windows (packed stream) -----------------------------------------------------------------------------
    // Copy the window of tile `t` (columns [t.base & ~1, + t.span) of the functor's vectors) into the next buffer: two bulk
    // copies issued by ONE thread (lane 0 of warp kWinWarp - not thread 0, which feeds the matrix stages). The buffer's
    // previous tenant must have been released by all 16 warps first; the issuing warp has itself left that tile and thread 0
    // had issued all of its items by then, so nobody waits in a circle. Every thread counts the window (CTA-uniform).
    static constexpr int kWinWarp = kWarpsPerBlock / 2;
    template <class Gather>
    __device__ __forceinline__ void window_issue(const TileDesc& t, const Gather& x) {
        // (with a producer warp in the CTA its lane 0 issues the windows as well, tiles ahead of the rows)
        if (blockDim.x > (unsigned)kBlock ? threadIdx.x == (unsigned)kBlock : threadIdx.x == kWinWarp * 32) {
            const unsigned buf = w_issued % kWinBufs, fill = w_issued / kWinBufs;
            if (fill > 0) {
                while (!mbar_try_wait_hint(&wempty[buf], (fill - 1u) & 1u, 1000u)) {
                }
            }
            const int off = t.base & ~1;
            const unsigned bytes = (unsigned)t.span * 8u;  // span is even: a multiple of 16 bytes
            const unsigned long long pol = l2_policy_keep();
            // the vectors were written with ordinary stores (by other CTAs, before the last grid barrier): order those
            // generic-proxy writes before the async-proxy reads of the bulk copies
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_arrive_expect_tx(&wfull[buf], bytes * (unsigned)Gather::kVecs);
#pragma unroll
            for (int v = 0; v < Gather::kVecs; ++v)
                bulk_g2s(win0 + (size_t)(buf * kWinVecs + v) * kWinCap, x.vec(v) + off, bytes, &wfull[buf], pol);
        }
        ++w_issued;
    }
    // Start of a windowed tile: make sure its window is on its way, send the next tile's after it (same system: the
    // functor's vectors are that system's), wait for this one. Returns the window of vector 0 (vector v: + v * kWinCap).
    template <class Gather>
    __device__ __forceinline__ const double* window_acquire(const TileDesc& d, const Gather& x) {
        if (w_issued == w_count) window_issue(d, x);
        const int next = (int)(&d - tab) + 1;
        if (next < ntiles && w_issued == w_count + 1u) {
            const TileDesc& nx = tab[next];
            if (nx.span > 0 && nx.sys == d.sys) window_issue(nx, x);
        }
        const unsigned buf = w_count % kWinBufs;
        while (!ghost && !mbar_try_wait_hint(&wfull[buf], (w_count / kWinBufs) & 1u, 1000u)) {
        }
        return win0 + (size_t)buf * kWinVecs * kWinCap;
    }
    __device__ __forceinline__ void window_release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0 && !ghost) mbar_arrive(&wempty[w_count % kWinBufs]);
        ++w_count;
    }

    // The items of tile d: every warp acquires and releases each of them; `load(c)` yields the gathered value of column c.
    template <int kUnroll, class Load>
    __device__ __forceinline__ double consume_items(const TileDesc& d, int rs, int re, bool compute, const Load& load) {
        double sum = 0.0;
        const int nb = blocks(d);
        for (int j = 0; j < nb; ++j) {
            const int bs = d.cs + j * kCap;
            const int be = min(d.ce, bs + kCap);
            const int as = bs & ~(kAlign - 1);
            const unsigned stage = acquire();
            DP_PIPE_MARK(2);
            if (compute && !ghost) {
                const ValT* __restrict__ sv = stage_val(stage);
                const ColT* __restrict__ sc = stage_col(stage);
                const int qe = min(re, be) - as;
                for (int q = max(rs, bs) - as; q < qe; q += kUnroll) {
                    int c[kUnroll];
                    double v[kUnroll], xv[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const bool on = q + u < qe;
                        c[u] = on ? (int)sc[q + u] : 0;
                        v[u] = on ? (double)sv[q + u] : 0.0;  // packed: fp32 -> fp64 is exact
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) xv[u] = (q + u < qe) ? load(c[u]) : 0.0;
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u)
                        if (q + u < qe) sum = __dadd_rn(sum, __dmul_rn(v[u], xv[u]));
                }
            }
            DP_PIPE_MARK(3);
            release();
        }
        return sum;
    }

    // Row sum of this thread's row (entries [rs, re) of the matrix, empty for rows >= n) of tile d.
    // Every warp of the CTA must call it for every tile of the round, in order (`compute == false` only drains).
    template <int kUnroll = kPipeUnroll, class Gather>
    __device__ __forceinline__ double tile_spmv(const TileDesc& d, int rs, int re, const Gather& x, bool compute) {
        if constexpr (kPacked && kPackWindows && Gather::kVecs > 0) {
            if (d.span > 0) {  // CTA-uniform: the gathers of this tile are served from its shared-memory window
                DP_PIPE_MARK(0);
                const double* w = window_acquire(d, x);
                DP_PIPE_MARK(1);
                const int shift = d.base & 1;  // the window starts at the even column at or below base
                const double sum = consume_items<kUnroll>(d, rs, re, compute, [&](int c16) {
                    return x.combine(w[c16 + shift], Gather::kVecs > 1 ? w[kWinCap + c16 + shift] : 0.0);
                });
                window_release();
                DP_PIPE_MARK(4);
                return sum;
            }
        }
        const int base = kPacked ? d.base : 0;
        DP_PIPE_MARK(0);
        const double sum = consume_items<kUnroll>(d, rs, re, compute, [&](int c) { return x(base + c); });
        DP_PIPE_MARK(4);
        return sum;
    }

    // Drain the items of a tile whose result is not wanted (its system finished in this very iteration). A window that was
    // already sent for it (never the case today: windows are only sent ahead inside one system) is consumed as well.
    __device__ __forceinline__ void tile_skip(const TileDesc& d) {
        if (kPacked && kPackWindows && d.span > 0 && w_issued > w_count) {
            const unsigned buf = w_count % kWinBufs;
            while (!ghost && !mbar_try_wait(&wfull[buf], (w_count / kWinBufs) & 1u)) {
            }
            window_release();
        }
        consume_items<1>(d, 0, 0, false, [](int) { return 0.0; });
    }
};

using Pipe = PipeT<kPipeCap, kPipeStages>;  // the SpMV geometry: whole tiles per stage
// ... and the packed stream over the same bytes: stages of the same entry count (half the bytes) + the gather windows
using PipePacked = PipeT<kPackCap, kPackStages, true>;

// Row extent of this thread's row in tile d (coalesced; issue one tile ahead to hide the latency).
__device__ __forceinline__ void tile_row_extent(const TileDesc& d, int& rs, int& re) {
    rs = re = 0;
    if (d.rowptr) {
        const int row = d.ltile * kTileRows + (int)threadIdx.x;
        rs = __ldg(d.rowptr + min(row, d.n));
        re = __ldg(d.rowptr + min(row + 1, d.n));
    }
}

// CTA sum of kN <= 3 values per thread with ONE barrier: `scratch` is double buffered (pipe.flip alternates).
// Per value bit-identical to block_sum (lane butterfly -> 16 warp sums -> 16-wide butterfly).
constexpr int kReduceSlots = 3 * kWarpsPerBlock;

template <int kN, class P>
__device__ __forceinline__ void tile_reduce(double (&v)[kN], double (*scratch2x)[kReduceSlots], P& pipe) {
    static_assert(kN <= 3, "scratch holds 3 values per warp");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* scratch = scratch2x[pipe.flip];
    pipe.flip ^= 1;
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0 && warp < kWarpsPerBlock) {
#pragma unroll
        for (int i = 0; i < kN; ++i) scratch[i * kWarpsPerBlock + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = half_warp_sum(scratch[i * kWarpsPerBlock + (lane & (kWarpsPerBlock - 1))]);
}


// The same sum for a tile that went through the pipeline, WITHOUT a CTA barrier: every warp parks its warp sums in a
// ring slot and bumps the slot's counter; the warp that arrives last folds the 16 warp sums (same fixed order, same
// bits as tile_reduce) and hands them to `write` in its lane 0. Warps are at most kPipeStages streaming tiles apart
// (a stage is re-armed only after all 16 warps passed it), so a ring of kRedRing slots is never overrun.
constexpr int kRedRing = 8;
static_assert(kMaxStages < kRedRing, "reduction ring must outlast the warps' maximum distance");

struct TileRed {
    double slot[kRedRing][3][kWarpsPerBlock];
    int count[kRedRing];
};

template <int kN, class P, class Write>
__device__ __forceinline__ void tile_reduce_async(double (&v)[kN], TileRed& red, P& pipe, const Write& write) {
    static_assert(kN <= 3, "ring slots hold 3 values per warp");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned r = pipe.t_count % kRedRing;
    ++pipe.t_count;
    if (pipe.ghost) return;  // the producer warp owns no row
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = warp_sum(v[i]);
    int old = 0;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kN; ++i) red.slot[r][i][warp] = v[i];
        __threadfence_block();
        old = atomicAdd(&red.count[r], 1);
    }
    old = __shfl_sync(kFull, old, 0);
    if (old == kWarpsPerBlock - 1) {
        __threadfence_block();
#pragma unroll
        for (int i = 0; i < kN; ++i) v[i] = half_warp_sum(red.slot[r][i][lane & (kWarpsPerBlock - 1)]);
        if (lane == 0) {
            red.count[r] = 0;
            write(v);
        }
    }
}

}  // namespace dp
