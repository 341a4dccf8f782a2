// pcg.cu — K5: the whole preconditioned-CG loop of cg.py:50-90 on the device, for a batch of independent systems.
//
// Reference semantics kept exactly (cg.py line numbers):
//   58-62  x = x0; r = b - A x; z = M r; p = z
//   66     res_0 = <z0,z0>/<b,b>            (iteration 0 checks the PRECONDITIONED residual)
//   70-72  for at most max_iter bodies: stop as soon as res < rtol (res is the SQUARED relative residual)
//   75-83  Ap = A p; rz = <r,z>; a = rz/<Ap,p>; x += a p; r -= a Ap; z = M r; beta = <r,z>/rz; p = z + beta p
//   86     res = <r,r>/<b,b>
//   90     iterations = bodies executed
// Element-wise updates round like torch (product, then sum); SpMV/SpTRSV rows are sequential sums; dot products
// are fixed-order trees (lane butterfly -> 16 warp sums -> per-tile partial -> strided sum + tree over the tiles),
// so a solve is bitwise reproducible and independent of which CTA processed which tile or of the batch it ran in.
//
// Phase structure (one grid-wide barrier after each phase; k = body index, "old/new" = double buffers):
//   A       p_new = z + beta p_old fused into the gather of Ap = A p_new; partial <Ap,p>; convergence check
//   APPLY1  r_new = r_old - a Ap, x += a p fused into the first half of z = M r_new; partial <r,r> (and <r,z>)
//             IDENTITY/JACOBI: element-wise z        CSR: z = M r_new        MULTIPLY: t = L^T r_new
//   APPLY2  MULTIPLY: z = L t; partial <r,z>
//   FWD/BWD SOLVE: y = L^-1 r_new, z = L^-T y (sync-free, sptrsv.cuh), then DOTRZ: partial <r,z>
// The initial z0 = M r0 runs the same APPLY phases with kInit (no update, <z,z> into the <r,r> slot).
//
// Scheduling. A tile is 512 consecutive rows of one system, processed by one CTA (one row per thread); its matrix
// entries arrive through the TMA tile pipeline (tilepipe.cuh), warps of a CTA are not coupled by barriers while they
// stream, and the per-tile dot partials are folded by whichever warp finishes last.
//   fused engine  : one persistent cooperative launch. Unfinished systems live in a compacted ACTIVE list; its
//                   tiles are split into gridDim contiguous ranges, so a CTA streams through consecutive tiles of
//                   (mostly) one system and evaluates that system's scalars once per phase. When systems finish,
//                   CTA 0 rebuilds the list (double buffered) during the next APPLY1 phase.
//   stepped engine: one launch per phase over the static list of all systems (same contiguous ranges); the host polls
//                   the finished counter every check_every iterations. Same phase code, same bits.
#include <stdlib.h>

#include <vector>

#include "sptrsv.cuh"
#include "trsv_ts.cuh"
#include "tilepipe.cuh"
#include "trsv_ls.cuh"

namespace dp {

int trsv_lookahead();                                          // sptrsv.cu
int pipe_trace_read_spmv(uint64_t* out_host, int capacity);    // spmv.cu (trace builds)

struct SysDev {
    int n, precond, ntiles;
    int rearm_t;  // SOLVE: PH_DOTRZ re-arms y (= t) for the next forward solve (see dp_pcg_solve_f64)
    int fwd_look, bwd_look;  // SOLVE: chunks of the widest level (parking distance of the sync-free solves)
    CsrView A, M, Mt;
    const double* dinv;
    const int* fwd_plan;
    const int* bwd_plan;
    LsFactor fwd_ls, bwd_ls;  // SOLVE, optional: level-ordered copies (rowptr == nullptr: sync-free solve)
    const double* b;
    double* x;
    double* r[2];
    double* p[2];
    double* z[2];
    double* ap;
    double* t;  // MULTIPLY: t = L^T r;  SOLVE: y = L^-1 r
    double* part_rr;
    double* part_pap;
    double* part_bb;
    double* part_rz[2];
    double* scal;  // [0],[1]: published <r,z> by body parity; [2]: <b,b>
    int* iters_out;
    double* res_out;
    double* history;
    double* coef;
};

struct Ctx {
    const SysDev* sys;
    int* state;           // [nsys] 1 = finished (written in phase A only, read in the other phases)
    const int* tile_ofs;  // [nsys+1] static tile prefix (stepped engine)
    const int* fwd_ofs;   // [nsys+1] plan chunk prefixes (SOLVE)
    const int* bwd_ofs;
    int* act_sys[2];      // active list, double buffered: system ids ...
    int* act_ofs[2];      // ... and tile prefix (count+1 entries)
    int* act_meta;        // [2][2]: {count, total tiles}
    int nsys, total_tiles, total_fwd, total_bwd;
    int pw_fwd, pw_bwd;
    int has_multiply, has_solve, has_ls;
    int keep_l2;  // the batch's matrices and vectors fit the L2 (a single small system): stream them without evict-first
    double rtol;
    int max_iter;
    unsigned long long* word;  // grid barrier (+ finished count in the upper half)
    int* n_done;               // finished count for the host (stepped engine)
    int* flag;
    long long* trace;          // optional timeline of CTA 0 (DPCG_TRACE=1): (label, clock64) pairs
    int trace_cap;
};

// Tile tables: the stream descriptors of the tiles of this CTA's range, one table per SpMV-shaped phase. They
// survive across iterations (the range only changes when the active list is rebuilt), so the producer thread finds
// the next item's addresses in shared memory instead of chasing rowptr through L2.
#ifndef DPCG_ROUND_TILES
#define DPCG_ROUND_TILES 64
#endif
constexpr int kMaxRoundTiles = DPCG_ROUND_TILES;
#ifndef DPCG_APPLY2_UNROLL
#define DPCG_APPLY2_UNROLL kPipeUnroll
#endif
constexpr int kApply2Unroll = DPCG_APPLY2_UNROLL;  // single-gather phase: can afford more loads in flight
#ifndef DPCG_PHASEA_UNROLL
#define DPCG_PHASEA_UNROLL 3
#endif
#ifndef DPCG_APPLY1_UNROLL
#define DPCG_APPLY1_UNROLL 3
#endif
constexpr int kPhaseAUnroll = DPCG_PHASEA_UNROLL;  // double-gather phases: two loads per entry
constexpr int kApply1Unroll = DPCG_APPLY1_UNROLL;
enum Table { TAB_A = 0, TAB_P1 = 1, TAB_P2 = 2 };  // A | L^T (MULTIPLY) or M (CSR) | L (MULTIPLY)

struct Smem {
    PipeShared pipe;
    TileDesc tab[3][kMaxRoundTiles];
    double scratch[2][3 * kWarpsPerBlock];  // tile_reduce (double buffered)
    TileRed red;                            // tile_reduce_async (tiles that went through the pipeline)
    LsShared ls;                            // level-stream solves: barriers of their pipeline geometry
    double scratch2[3 * kWarpsPerBlock];    // block_sum* of the per-system scalar evaluation
    SysDev sys;  // descriptor of the system this CTA is working on (survives across phases)
    int sys_id;
    int tab_ver[3], tab_ga[3], tab_gb[3];  // what each table currently describes (ver 0 = nothing)
    int scan[kWarpsPerBlock + 1];
    int trace_pos;
    int bcast;
};

static_assert(sizeof(SysDev) % 8 == 0, "SysDev is copied as 8-byte words");
static_assert(sizeof(Smem) <= (233472 / 2) - 1024, "two CTAs per SM: 228 KB of shared memory, 1 KB reserved per CTA");

enum Phase { PH_INIT = 0, PH_A = 1, PH_APPLY1 = 2, PH_APPLY2 = 3, PH_FWD = 4, PH_BWD = 5, PH_DOTRZ = 6 };

__device__ __forceinline__ void trace(const Ctx& ctx, Smem& sm, int label) {
    if (ctx.trace_cap > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        const int i = sm.trace_pos;
        if (i < ctx.trace_cap) {
            ctx.trace[2 * i] = label;
            ctx.trace[2 * i + 1] = clock64();
            sm.trace_pos = i + 1;
        }
    }
}

// CTA-uniform: make sm.sys hold system s (one L2 round trip only when the system changes).
__device__ __forceinline__ const SysDev& load_sys(const Ctx& ctx, int s, Smem& sm) {
    if (sm.sys_id != s) {
        __syncthreads();
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(ctx.sys + s);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sm.sys);
        for (int i = threadIdx.x; i < (int)(sizeof(SysDev) / 8); i += kBlock) dst[i] = __ldg(src + i);
        if (threadIdx.x == 0) sm.sys_id = s;
        __syncthreads();
    }
    return sm.sys;
}

__device__ __forceinline__ int find_segment(const int* ofs, int count, int g) {
    int lo = 0, hi = count;  // largest i with ofs[i] <= g
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(ofs + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// Per-CTA cache of the scalars of the system it is currently working on (CTA-uniform registers).
struct Scal {
    int sys = -1;
    bool active = false;
    double v = 0.0;  // beta (phase A) or a (APPLY1)
};

// ---- PH_INIT: r0 = b - A x0 (cg.py:60), <b,b>, p = 0, arm the sync-free buffers -------------------------------
template <class P>
__device__ __forceinline__ void phase_init(const Ctx& ctx, const SysDev& S, const TileDesc& d, int rs, int re,
                                           Smem& sm, P& pipe) {
    const int tile = d.ltile;
    const int row = tile * kTileRows + threadIdx.x;
    const double ax = pipe.tile_spmv(d, rs, re, GatherPlain{S.x}, true);
    double bb[1] = {0.0};
    if (row < S.n && row_thread()) {
        const double bi = S.b[row];
        S.r[1][row] = __dsub_rn(bi, ax);
        S.p[0][row] = 0.0;
        bb[0] = __dmul_rn(bi, bi);
        if (S.precond == DP_PRECOND_SOLVE) {
            st_relaxed_u64(S.z[0] + row, kPending);
            st_relaxed_u64(S.t + row, kPending);
        }
    }
    if (threadIdx.x == 0 && tile == 0) ctx.state[d.sys] = 0;
    double* part_bb = S.part_bb;
    auto write = [part_bb, tile](const double (&v)[1]) { part_bb[tile] = v[0]; };
    if (d.ce > d.cs) {
        tile_reduce_async<1>(bb, sm.red, pipe, write);
    } else {
        tile_reduce<1>(bb, sm.scratch, pipe);
        if (threadIdx.x == 0) write(bb);
    }
}

// ---- PH_A ---------------------------------------------------------------------------------------------------
// kCheckState: the static schedule (stepped engine) still visits finished systems and must skip them; the flag is
// written by another CTA in this very phase, so it is read once per CTA and broadcast (a per-thread read could
// split the CTA around a barrier). The fused engine's active list never contains a finished system.
template <bool kCheckState, class P>
__device__ __forceinline__ void phase_a(const Ctx& ctx, const SysDev& S, const TileDesc& d, int rs, int re, int k,
                                        Smem& sm, Scal& sc, P& pipe) {
    const int s = d.sys, tile = d.ltile;
    const int row = tile * kTileRows + threadIdx.x;
    const bool valid = row < S.n && row_thread();
    const double* z = S.z[k & 1];
    const double* po = S.p[k & 1];
    // loads that do not depend on this phase's scalars go first
    double zr = 0.0, pr = 0.0;
    if (valid) zr = z[row], pr = po[row];
    if (pipe.ghost) {
        if (d.span > 0) prefetch_columns(z, d.base, d.span), prefetch_columns(po, d.base, d.span);  // (covers the rows' own lines)
        else prefetch_tile_rows(z, tile, S.n), prefetch_tile_rows(po, tile, S.n);
    }
    if (sc.sys != s) {
        sc.sys = s;
        sc.active = false;
        bool done = false;
        if (kCheckState) done = __syncthreads_or(threadIdx.x == 0 ? ld_relaxed_s32(ctx.state + s) : 0) != 0;
        if (!done) {
            double v[2] = {0.0, 0.0};
            const double* prr = S.part_rr;
            const double* prz = S.part_rz[k & 1];
            if (row_thread()) {
                for (int i = threadIdx.x; i < S.ntiles; i += kBlock) {
                    v[0] = __dadd_rn(v[0], __ldcg(prr + i));
                    v[1] = __dadd_rn(v[1], __ldcg(prz + i));
                }
            }
            const double bb = __ldcg(S.scal + 2);
            const double rz_prev = k > 0 ? __ldcg(S.scal + ((k - 1) & 1)) : 1.0;
            block_sum_n<2>(v, sm.scratch2);
            const double res = v[0] / bb;  // cg.py:17
            const bool finished = (res < ctx.rtol) || (k >= ctx.max_iter);  // cg.py:70-72
            sc.active = !finished;
            sc.v = k > 0 ? v[1] / rz_prev : 0.0;  // beta, cg.py:82 (p_0 = z_0, cg.py:62)
            if (tile == 0 && threadIdx.x == 0) {  // tile 0 of a system is visited by exactly one CTA per phase
                if (S.history) S.history[k] = res;
                if (finished) {
                    *S.iters_out = k;
                    *S.res_out = res;
                    ctx.state[s] = 1;
                    atomicAdd(ctx.n_done, 1);
                    atomicAdd(ctx.word, kDoneUnit);
                } else {
                    S.scal[k & 1] = v[1];
                    if (S.coef) S.coef[2 * k + 1] = sc.v;  // beta behind this body's p (cg.py:82,83)
                }
            }
        }
    }
    if (!sc.active) {
        pipe.tile_skip(d);
        return;
    }
    double* pn = S.p[(k + 1) & 1];
    const double ap = pipe.template tile_spmv<kPhaseAUnroll>(d, rs, re, GatherZBetaP{z, po, sc.v}, true);  // cg.py:75
    double pap[1] = {0.0};
    if (valid) {
        const double pi = __dadd_rn(zr, __dmul_rn(sc.v, pr));  // cg.py:83
        pn[row] = pi;
        S.ap[row] = ap;
        pap[0] = __dmul_rn(ap, pi);
        if (S.precond == DP_PRECOND_SOLVE) st_relaxed_u64(S.z[(k + 1) & 1] + row, kPending);  // re-arm next z
    }
    double* part_pap = S.part_pap;
    auto write = [part_pap, tile](const double (&v)[1]) { part_pap[tile] = v[0]; };
    if (d.ce > d.cs) {
        tile_reduce_async<1>(pap, sm.red, pipe, write);
    } else {
        tile_reduce<1>(pap, sm.scratch, pipe);
        if (threadIdx.x == 0) write(pap);
    }
}

// ---- PH_APPLY1 ----------------------------------------------------------------------------------------------
template <bool kInit, class P>
__device__ __forceinline__ void phase_apply1(const Ctx& ctx, const SysDev& S, const TileDesc& d, int rs, int re, int k,
                                             Smem& sm, Scal& sc, P& pipe) {
    const int s = d.sys, tile = d.ltile;
    const int row = tile * kTileRows + threadIdx.x;
    const bool valid = row < S.n && row_thread();
    const double* ro = S.r[k & 1];
    double* rnw = S.r[(k + 1) & 1];
    const double* pn = S.p[(k + 1) & 1];
    double* zn = S.z[(k + 1) & 1];
    const int precond = S.precond;
    // scalar-independent loads first
    double r_old = 0.0, ap_i = 0.0, x_i = 0.0, p_i = 0.0, dinv_i = 0.0;
    if (pipe.ghost) {
        if (d.span > 0) {  // the gather of L^T r / M r reaches beyond this tile's rows
            prefetch_columns(ro, d.base, d.span);
            if (!kInit) prefetch_columns(S.ap, d.base, d.span);
        }
        prefetch_tile_rows(ro, tile, S.n);
        if (!kInit) prefetch_tile_rows(S.ap, tile, S.n), prefetch_tile_rows(S.x, tile, S.n), prefetch_tile_rows(pn, tile, S.n);
    }
    if (valid) {
        r_old = ro[row];
        if (!kInit) ap_i = S.ap[row], x_i = S.x[row], p_i = pn[row];
        if (precond == DP_PRECOND_JACOBI) dinv_i = __ldg(S.dinv + row);
    }
    if (sc.sys != s) {
        sc.sys = s;
        sc.active = kInit || ld_relaxed_s32(ctx.state + s) == 0;  // written before the previous barrier: uniform
        sc.v = 0.0;
        if (!kInit && sc.active) {
            sc.v = __ldcg(S.scal + (k & 1)) / block_reduce_array(S.part_pap, S.ntiles, sm.scratch2);  // a, cg.py:78
            if (tile == 0 && threadIdx.x == 0 && S.coef) S.coef[2 * k] = sc.v;  // tile 0: one CTA per system and phase
        }
    }
    if (!sc.active) {
        pipe.tile_skip(d);
        return;
    }
    const double a = sc.v;

    // (Streaming the matrix first and updating r / x afterwards, so that this row's loads fly under the stream instead of
    // stalling the head of the tile, was measured and is slower: the row loop then waits for them on the same scoreboard -
    // APPLY1 head 1.0 -> 0.4 us, row loop 2.1 -> 3.2 us, profiles/r2/pack_experiments.md.)
    double rn = 0.0;
    if (valid) {
        rn = kInit ? r_old : __dsub_rn(r_old, __dmul_rn(a, ap_i));      // cg.py:80
        rnw[row] = rn;
        if (!kInit) S.x[row] = __dadd_rn(x_i, __dmul_rn(a, p_i));       // cg.py:79
    }
    double zi = 0.0;
    bool have_z = true;
    switch (precond) {
        case DP_PRECOND_IDENTITY:
            zi = rn;
            break;
        case DP_PRECOND_JACOBI:
            zi = __dmul_rn(dinv_i, rn);
            break;
        case DP_PRECOND_CSR:  // table P1 streams M
            zi = kInit ? pipe.tile_spmv(d, rs, re, GatherWork{ro}, true)
                       : pipe.template tile_spmv<kApply1Unroll>(d, rs, re, GatherRMinusAAp{ro, S.ap, a}, true);
            break;
        case DP_PRECOND_MULTIPLY: {  // table P1 streams L^T
            const double ti = kInit ? pipe.tile_spmv(d, rs, re, GatherWork{ro}, true)
                                    : pipe.template tile_spmv<kApply1Unroll>(d, rs, re, GatherRMinusAAp{ro, S.ap, a}, true);
            if (valid) S.t[row] = ti;
            have_z = false;
            break;
        }
        default:  // DP_PRECOND_SOLVE: z comes from FWD/BWD
            have_z = false;
            break;
    }
    if (have_z && valid) zn[row] = zi;
    // partial dot products of this tile: <r,r> (cg.py:86), <r,z> (cg.py:76,82), and <z,z> for iteration 0 (cg.py:66)
    double v[3] = {__dmul_rn(rn, rn), __dmul_rn(rn, zi), __dmul_rn(zi, zi)};
    double* part_rr = S.part_rr;
    double* part_rz = S.part_rz[(k + 1) & 1];
    auto write = [part_rr, part_rz, tile, have_z](const double (&w)[3]) {
        if (!kInit) part_rr[tile] = w[0];
        if (have_z) {
            part_rz[tile] = w[1];
            if (kInit) part_rr[tile] = w[2];
        }
    };
    if (d.ce > d.cs) {
        tile_reduce_async<3>(v, sm.red, pipe, write);
    } else {
        tile_reduce<3>(v, sm.scratch, pipe);
        if (threadIdx.x == 0) write(v);
    }
    if (kInit && tile == 0) {  // publish <b,b> once (all part_bb were written before the previous barrier)
        const double bb = block_reduce_array(S.part_bb, S.ntiles, sm.scratch2);
        if (threadIdx.x == 0) S.scal[2] = bb;
    }
}

// ---- PH_APPLY2 (MULTIPLY): z = L t ---------------------------------------------------------------------------
template <bool kInit, class P>
__device__ __forceinline__ void phase_apply2(const Ctx& ctx, const SysDev& S, const TileDesc& d, int rs, int re, int k,
                                             Smem& sm, Scal& sc, P& pipe) {
    const int s = d.sys, tile = d.ltile;
    if (sc.sys != s) {
        sc.sys = s;
        sc.active = S.precond == DP_PRECOND_MULTIPLY && (kInit || ld_relaxed_s32(ctx.state + s) == 0);
    }
    if (!sc.active) {
        pipe.tile_skip(d);
        return;
    }
    const int row = tile * kTileRows + threadIdx.x;
    const bool valid = row < S.n && row_thread();
    double rn = 0.0;
    if (valid) rn = S.r[(k + 1) & 1][row];
    if (pipe.ghost) {
        prefetch_tile_rows(S.r[(k + 1) & 1], tile, S.n);
        if (d.span > 0) prefetch_columns(S.t, d.base, d.span);
    }
    const double zi = pipe.template tile_spmv<kApply2Unroll>(d, rs, re, GatherWork{S.t}, true);
    if (valid) S.z[(k + 1) & 1][row] = zi;
    double v[2] = {__dmul_rn(rn, zi), __dmul_rn(zi, zi)};
    double* part_rr = S.part_rr;
    double* part_rz = S.part_rz[(k + 1) & 1];
    auto write = [part_rr, part_rz, tile](const double (&w)[2]) {
        part_rz[tile] = w[0];
        if (kInit) part_rr[tile] = w[1];
    };
    if (d.ce > d.cs) {
        tile_reduce_async<2>(v, sm.red, pipe, write);
    } else {
        tile_reduce<2>(v, sm.scratch, pipe);
        if (threadIdx.x == 0) write(v);
    }
}

// ---- PH_DOTRZ (SOLVE): <r,z> after the backward solve -----------------------------------------------------------
template <bool kInit, class P>
__device__ __forceinline__ void phase_dotrz(const Ctx& ctx, const SysDev& S, const TileDesc& d, int k, Smem& sm, Scal& sc,
                                            P& pipe) {
    const int s = d.sys, tile = d.ltile;
    if (sc.sys != s) {
        sc.sys = s;
        sc.active = S.precond == DP_PRECOND_SOLVE && (kInit || ld_relaxed_s32(ctx.state + s) == 0);
    }
    if (!sc.active) return;
    const int row = tile * kTileRows + threadIdx.x;
    double rn = 0.0, zi = 0.0;
    if (row < S.n && row_thread()) {
        rn = S.r[(k + 1) & 1][row];
        zi = S.z[(k + 1) & 1][row];
        // a sync-free forward solve polls y and needs it armed before every application; the sync-free backward solve
        // re-arms it on the way (RhsConsume), the level-stream and tile-stream ones read it plainly and cannot
        if (S.rearm_t) st_relaxed_u64(S.t + row, kPending);
    }
    double v[2] = {__dmul_rn(rn, zi), __dmul_rn(zi, zi)};
    tile_reduce<2>(v, sm.scratch, pipe);
    if (threadIdx.x == 0) {
        S.part_rz[(k + 1) & 1][tile] = v[0];
        if (kInit) S.part_rr[tile] = v[1];
    }
}

// ---- PH_FWD / PH_BWD (SOLVE): y = L^-1 r_new ; z_new = L^-T y ---------------------------------------------------
// The participating warps are dealt to the systems (system s gets warps s, s + nsys, ...), so the solves of a batch
// advance side by side: the critical path (levels x L2 round trip) is paid once per batch, not once per system.
// Out of line on purpose: the level-stream solve keeps two tiles of row metadata in registers; inlined into the
// persistent kernel it would push the SpMV phases into register spills.
template <bool kUpper>
__device__ __noinline__ void trsv_level_stream_outlined(const LsFactor* F, const double* rhs, double* x,
                                                        unsigned char* stage_bytes, LsShared* ls) {
    trsv_level_stream<kUpper, kLsStagesFused>(*F, rhs, x, stage_bytes, *ls);
}

// Systems that come with a level-ordered copy of the factor are solved by ONE CTA each (trsv_ls.cuh), on the bytes of
// the SpMV pipeline (idle between the phases).
template <bool kUpper, bool kInit>
__device__ __forceinline__ void phase_trsv_ls(const Ctx& ctx, int k, Smem& sm) {
    for (int s = blockIdx.x; s < ctx.nsys; s += gridDim.x) {
        const SysDev* S = ctx.sys + s;
        if (S->precond != DP_PRECOND_SOLVE) continue;
        const LsFactor F = kUpper ? S->bwd_ls : S->fwd_ls;
        if (!F.rowptr) continue;
        if (!kInit && ld_relaxed_s32(ctx.state + s) != 0) continue;
        const double* rhs = kUpper ? S->t : S->r[(k + 1) & 1];
        double* x = kUpper ? S->z[(k + 1) & 1] : S->t;
        trsv_level_stream_outlined<kUpper>(&F, rhs, x, sm.pipe.bytes, &sm.ls);
    }
}

template <bool kUpper, bool kInit>
__device__ __forceinline__ bool phase_trsv(const Ctx& ctx, int k, const Smem& sm) {
    const int pw = kUpper ? ctx.pw_bwd : ctx.pw_fwd;
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    if (gw >= pw) return true;
    const int* ofs = kUpper ? ctx.bwd_ofs : ctx.fwd_ofs;
    const AbortCtl ctl{ctx.word, ctx.flag};
    const int nsys = ctx.nsys;
    const int wps = pw >= nsys ? pw / nsys : 1;       // warps per system
    const int local = pw >= nsys ? gw / nsys : 0;     // this warp's index inside its system's team
    if (local >= wps) return true;
    for (int s = pw >= nsys ? gw % nsys : gw; s < nsys; s += pw) {
        const SysDev& S = (sm.sys_id == s) ? sm.sys : ctx.sys[s];
        if (S.precond != DP_PRECOND_SOLVE) continue;
        if ((kUpper ? S.bwd_ls.rowptr : S.fwd_ls.rowptr) != nullptr) continue;  // solved by phase_trsv_ls
        if (!kInit && ld_relaxed_s32(ctx.state + s) != 0) continue;
        const long long nchunks = __ldg(ofs + s + 1) - __ldg(ofs + s);
        bool ok;
        if (kUpper)
            ok = sptrsv_stream<true>(S.Mt, S.bwd_plan, local, wps, nchunks, S.bwd_look, RhsConsume{S.t}, S.z[(k + 1) & 1], ctl);
        else
            ok = sptrsv_stream<false>(S.M, S.fwd_plan, local, wps, nchunks, S.fwd_look, RhsPlain{S.r[(k + 1) & 1]}, S.t, ctl);
        if (!ok) return false;
    }
    return true;
}

// ---- tile schedule ---------------------------------------------------------------------------------------------
template <int kPhase, bool kInit, bool kCheckState, class P>
__device__ __forceinline__ void run_tile(const Ctx& ctx, const SysDev& S, const TileDesc& d, int rs, int re, int k,
                                         Smem& sm, Scal& sc, P& pipe) {
    if (kPhase == PH_INIT) phase_init(ctx, S, d, rs, re, sm, pipe);
    if (kPhase == PH_A) phase_a<kCheckState>(ctx, S, d, rs, re, k, sm, sc, pipe);
    if (kPhase == PH_APPLY1) phase_apply1<kInit>(ctx, S, d, rs, re, k, sm, sc, pipe);
    if (kPhase == PH_APPLY2) phase_apply2<kInit>(ctx, S, d, rs, re, k, sm, sc, pipe);
    if (kPhase == PH_DOTRZ) phase_dotrz<kInit>(ctx, S, d, k, sm, sc, pipe);
}

// Describe tiles [ga, gb) of list `cur` for table kTab: which system / local tile, and what the phase streams.
template <int kTab, bool kPacked>
__device__ __forceinline__ void build_table(const Ctx& ctx, int cur, int ga, int gb, Smem& sm) {
    const int count = __ldcg(ctx.act_meta + 2 * cur);
    const int* asys = ctx.act_sys[cur];
    const int* aofs = ctx.act_ofs[cur];
    for (int i = threadIdx.x; i < gb - ga; i += kBlock) {
        const int g = ga + i;
        const int seg = count == 1 ? 0 : find_segment(aofs, count, g);
        const int s = __ldcg(asys + seg);
        const int lt = g - __ldcg(aofs + seg);
        const SysDev* S = ctx.sys + s;
        const int precond = S->precond;
        CsrView M{nullptr, nullptr, nullptr, 0, 0};
        if (kTab == TAB_A) M = S->A;
        if (kTab == TAB_P1 && precond == DP_PRECOND_MULTIPLY) M = S->Mt;
        if (kTab == TAB_P1 && precond == DP_PRECOND_CSR) M = S->M;
        if (kTab == TAB_P2 && precond == DP_PRECOND_MULTIPLY) M = S->M;
        TileDesc d;
        tile_desc_fill<kPacked>(d, M, lt);
        d.sys = s;
        sm.tab[kTab][i] = d;
    }
}

template <int kTab, bool kPacked>
__device__ __forceinline__ void ensure_table(const Ctx& ctx, int cur, int ver, int ga, int gb, Smem& sm) {
    if (sm.tab_ver[kTab] == ver && sm.tab_ga[kTab] == ga && sm.tab_gb[kTab] == gb) return;  // CTA-uniform
    __syncthreads();  // every thread has compared the keys / left the old table
    build_table<kTab, kPacked>(ctx, cur, ga, gb, sm);
    if (threadIdx.x == 0) sm.tab_ver[kTab] = ver, sm.tab_ga[kTab] = ga, sm.tab_gb[kTab] = gb;
    __syncthreads();
}

// The tiles of the systems in list `cur`, one contiguous range per CTA, streamed through the tile pipeline.
// fused engine: `cur` = the active list (unfinished systems); stepped engine: list 0 = all systems, kCheckState.
// `next_tab` (fused engine only, -1 = none): the table the NEXT phase will stream; when it is still valid for this
// CTA's range its first items are issued before the grid barrier (Pipe::begin_early).
template <int kPhase, bool kInit, bool kCheckState, class P>
__device__ __forceinline__ void run_tiles(const Ctx& ctx, int k, int cur, int ver, int total, Smem& sm, P& pipe,
                                          int next_tab = -1) {
    constexpr int kTab = kPhase == PH_APPLY1 ? TAB_P1 : kPhase == PH_APPLY2 ? TAB_P2 : TAB_A;
    constexpr bool kStream = kPhase != PH_DOTRZ;
    // `total` = tiles of list `cur`: kept in a register by the caller (it only changes when the list is rebuilt), which
    // saves an L2 round trip at the head of every phase
    // contiguous ranges that differ by at most one tile: the first `rem` CTAs take one more
    const int quot = total / (int)gridDim.x, rem = total % (int)gridDim.x;
    const int g0 = (int)blockIdx.x * quot + min((int)blockIdx.x, rem), g1 = g0 + quot + ((int)blockIdx.x < rem ? 1 : 0);
    Scal sc;
#ifdef DPCG_PIPE_TRACE
    pipe.tr_phase = kPhase;
#endif
    for (int ga = g0; ga < g1; ga += kMaxRoundTiles) {
        const int gb = min(g1, ga + kMaxRoundTiles);
        ensure_table<kTab, P::kIsPacked>(ctx, cur, ver, ga, gb, sm);
        const TileDesc* tab = sm.tab[kTab];
        if (kStream) {
            if (!pipe.begin_resume(tab, gb - ga, kTab)) {
                pipe.drain_early();  // (never taken: the prediction below is exact; kept as the safe way out)
                pipe.begin(tab, gb - ga);
            }
        }
        for (int i = 0; i < gb - ga; ++i) {
            const TileDesc& d = tab[i];
            const SysDev& S = load_sys(ctx, d.sys, sm);
            // row extents: requested here, they land while the scalars are evaluated and the stage arrives (holding them
            // in registers a tile ahead only produced spills whose stores waited for the loads)
            int rs = 0, re = 0;
            if (kStream) tile_row_extent(d, rs, re);
            run_tile<kPhase, kInit, kCheckState>(ctx, S, d, rs, re, k, sm, sc, pipe);
#ifdef DPCG_PIPE_TRACE
            pipe.tr_mark(5);
#endif
        }
    }
    if (next_tab >= 0 && g1 > g0 && g1 - g0 <= kMaxRoundTiles && sm.tab_ver[next_tab] == ver &&
        sm.tab_ga[next_tab] == g0 && sm.tab_gb[next_tab] == g1)
        pipe.begin_early(sm.tab[next_tab], g1 - g0, next_tab);
}

// CTA 0: compact the active list `cur` into `cur ^ 1`, dropping the systems whose state flag is up.
__device__ __forceinline__ void rebuild_active(const Ctx& ctx, int cur, Smem& sm) {
    const int count = __ldcg(ctx.act_meta + 2 * cur);
    const int* asys = ctx.act_sys[cur];
    const int* aofs = ctx.act_ofs[cur];
    int* nsys_out = ctx.act_sys[cur ^ 1];
    int* nofs_out = ctx.act_ofs[cur ^ 1];
    int carry_n = 0, carry_t = 0;
    for (int i0 = 0; i0 < count; i0 += kBlock) {
        const int i = i0 + threadIdx.x;
        int s = -1, alive = 0, nt = 0;
        if (i < count && row_thread()) {
            s = __ldcg(asys + i);
            alive = ld_relaxed_s32(ctx.state + s) == 0;
            nt = alive ? __ldcg(aofs + i + 1) - __ldcg(aofs + i) : 0;
        }
        int tot_n, tot_t;
        const int pos = block_exclusive_scan_int(alive, &tot_n, sm.scan);
        const int tofs = block_exclusive_scan_int(nt, &tot_t, sm.scan);
        if (alive) {
            nsys_out[carry_n + pos] = s;
            nofs_out[carry_n + pos] = carry_t + tofs;
        }
        carry_n += tot_n;
        carry_t += tot_t;
    }
    if (threadIdx.x == 0) {
        nofs_out[carry_n] = carry_t;
        ctx.act_meta[2 * (cur ^ 1)] = carry_n;
        ctx.act_meta[2 * (cur ^ 1) + 1] = carry_t;
    }
}

// Returns false on abort. `done` = finished count of the last barrier.
// `list_stays`: the active list (hence every table) survives this iteration, so the last phase may start table A early.
template <bool kInit, bool kSolve, class P>
__device__ __forceinline__ bool apply_preconditioner(const Ctx& ctx, int k, int cur, int ver, int total, GridBarrier& bar,
                                                     Smem& sm, P& pipe, bool list_stays) {
    const int tab_a = list_stays ? (int)TAB_A : -1;
    run_tiles<PH_APPLY1, kInit, false>(ctx, k, cur, ver, total, sm, pipe,
                                       ctx.has_multiply ? (int)TAB_P2 : (kSolve ? -1 : tab_a));
    trace(ctx, sm, 8 * PH_APPLY1 + 1);
    if (bar.sync() < 0) return false;
    trace(ctx, sm, 8 * PH_APPLY1 + 2);
    if (ctx.has_multiply) {
        run_tiles<PH_APPLY2, kInit, false>(ctx, k, cur, ver, total, sm, pipe, kSolve ? -1 : tab_a);
        trace(ctx, sm, 8 * PH_APPLY2 + 1);
        if (bar.sync() < 0) return false;
        trace(ctx, sm, 8 * PH_APPLY2 + 2);
    }
    if (kSolve) {
        if (ctx.has_ls) phase_trsv_ls<false, kInit>(ctx, k, sm);
        const bool f = phase_trsv<false, kInit>(ctx, k, sm);
        if (bar.sync() < 0 || !f) return false;
        if (ctx.has_ls) phase_trsv_ls<true, kInit>(ctx, k, sm);
        const bool b = phase_trsv<true, kInit>(ctx, k, sm);
        if (bar.sync() < 0 || !b) return false;
        run_tiles<PH_DOTRZ, kInit, false>(ctx, k, cur, ver, total, sm, pipe, tab_a);
        if (bar.sync() < 0) return false;
    }
    return true;
}

template <class P>
__device__ __forceinline__ Smem& smem_init(unsigned char* raw, P& pipe) {
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    if (threadIdx.x == 0) {
        sm.sys_id = -1, sm.trace_pos = 0;
        sm.tab_ver[0] = sm.tab_ver[1] = sm.tab_ver[2] = 0;
        for (int i = 0; i < kRedRing; ++i) sm.red.count[i] = 0;
        sm.ls.init();
    }
    pipe.init(sm.pipe.bytes, &sm.pipe.bar);  // ends with a CTA barrier
    return sm;
}

// The whole solve in one persistent cooperative launch: device-side loop control, no host round trips.
// kSolve: the batch holds SOLVE-mode systems. Two instantiations, so that the triangular-solve code (two inlined sync-free
// streams, the level-stream call) does not weigh on the register allocation of the batches that never run it - the
// benchmarked multiply-mode batches among them.
// kPacked: every matrix the batch streams comes with its packed copy (dp_csr_pack): 6 instead of 12 bytes per entry through
// the same pipeline, same bits.
// The packed engine runs kBlock + 32 threads: warp 16 is the PRODUCER warp. It walks the same control flow as the row warps
// (every CTA and grid barrier), owns no row, contributes to no sum, and its lane 0 issues every item of the tile pipeline as
// soon as the item's stage has been handed back - up to kStages items ahead of the rows, instead of 0.86 us per item on the
// path of warp 0 (profiles/r2/pack_experiments.md). At 544 threads the kernel must fit 56 registers; this instantiation does.
constexpr int kFusedThreadsPacked = kPackProducerWarp ? kBlock + kWarp : kBlock;
template <bool kSolve, bool kPacked>
__global__ void __launch_bounds__(kPacked ? kFusedThreadsPacked : kBlock, 2) pcg_fused_kernel(Ctx ctx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename std::conditional<kPacked, PipePacked, Pipe>::type pipe;
    Smem& sm = smem_init(smem_raw, pipe);
    pipe.keep_l2 = ctx.keep_l2;
    GridBarrier bar{ctx.word, ctx.flag, &sm.bcast, 0u, gridDim.x};
    int cur = 0;          // active-list buffer in use
    int ver = 1;          // bumped whenever the list (hence every CTA's tile range) changes
    int done_built = 0;   // finished count the list `cur` reflects
    int total = ctx.total_tiles;  // tiles of the active list (all systems at first)
    run_tiles<PH_INIT, false, false>(ctx, -1, cur, ver, total, sm, pipe, TAB_P1);
    if (bar.sync() < 0) return;
    if (!apply_preconditioner<true, kSolve>(ctx, -1, cur, ver, total, bar, sm, pipe, true)) return;
    for (int k = 0; k <= ctx.max_iter; ++k) {
        trace(ctx, sm, 8 * PH_A + 0);
        run_tiles<PH_A, false, false>(ctx, k, cur, ver, total, sm, pipe, TAB_P1);
        trace(ctx, sm, 8 * PH_A + 1);
        const int done = bar.sync();
        trace(ctx, sm, 8 * PH_A + 2);
        if (done < 0) return;
        if (done >= ctx.nsys) {
            pipe.drain_early();  // do not exit with bulk copies in flight
            break;
        }
        const bool rebuild = done != done_built;  // same decision in every CTA
        if (rebuild && blockIdx.x == 0) rebuild_active(ctx, cur, sm);
        if (!apply_preconditioner<false, kSolve>(ctx, k, cur, ver, total, bar, sm, pipe, !rebuild)) return;  // >= 1 barrier: the new list is visible
        if (rebuild) {
            cur ^= 1, done_built = done, ++ver;
            total = __ldcg(ctx.act_meta + 2 * cur + 1);
        }
    }
}

// Stepped engine: the same phases, one launch each (the launch boundary is the barrier), over the static list of
// all systems (list 0 as uploaded by the host); finished systems are skipped by their state flag.
template <int kPhase, bool kInit>
__global__ void __launch_bounds__(kBlock, 2) pcg_phase_kernel(Ctx ctx, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pipe pipe;
    Smem& sm = smem_init(smem_raw, pipe);
    pipe.keep_l2 = ctx.keep_l2;
    if (kPhase == PH_FWD) {
        if (ctx.has_ls) phase_trsv_ls<false, kInit>(ctx, k, sm);
        phase_trsv<false, kInit>(ctx, k, sm);
    } else if (kPhase == PH_BWD) {
        if (ctx.has_ls) phase_trsv_ls<true, kInit>(ctx, k, sm);
        phase_trsv<true, kInit>(ctx, k, sm);
    } else {
        run_tiles<kPhase, kInit, true>(ctx, k, 0, 1, ctx.total_tiles, sm, pipe);
    }
}

// ---- lean element-wise kernels of the stepped engine for batches whose systems are all SOLVE-mode (tile-stream) --------
// In SOLVE mode PH_APPLY1 streams no matrix (r_new = r_old - a Ap, x += a p, <r,r>) and PH_DOTRZ is <r,z> plus re-arming y:
// pure vector passes. Through pcg_phase_kernel they run two 512-thread CTAs per SM (its shared-memory pipeline and 64
// registers), one tile per CTA in flight and a dependent load round per tile: 0.53 / 0.41 of the HBM peak on 8 x 256^3
// (tools/c5_timeline.py), a quarter of a config-5 iteration. These kernels do the same arithmetic in the same order -
// same per-tile partial sums (lane butterfly -> 16 warp sums -> half-warp butterfly), same scalars, same bits - with
// three CTAs per SM and nothing but registers.
__device__ __forceinline__ double tile_sum(double v, double (*scratch)[kWarpsPerBlock], int& flip) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sc = scratch[flip];
    flip ^= 1;
    v = warp_sum(v);
    if (lane == 0) sc[warp] = v;
    __syncthreads();
    return half_warp_sum(sc[lane & (kWarpsPerBlock - 1)]);
}

struct LeanCursor {  // walks the static list of all systems' tiles: contiguous range of this CTA
    int g, g1, s, lo, hi;
    __device__ __forceinline__ void init(const Ctx& ctx) {
        const int total = ctx.total_tiles;
        const int quot = total / (int)gridDim.x, rem = total % (int)gridDim.x;
        g = (int)blockIdx.x * quot + min((int)blockIdx.x, rem);
        g1 = g + quot + ((int)blockIdx.x < rem ? 1 : 0);
        s = g < g1 ? find_segment(ctx.tile_ofs, ctx.nsys, g) : 0;
        lo = __ldg(ctx.tile_ofs + s), hi = __ldg(ctx.tile_ofs + s + 1);
    }
    __device__ __forceinline__ bool advance(const Ctx& ctx) {  // true when the system changed
        bool changed = false;
        while (g >= hi) ++s, lo = hi, hi = __ldg(ctx.tile_ofs + s + 1), changed = true;
        return changed;
    }
};

template <bool kInit>
__global__ void __launch_bounds__(kBlock, 3) solve_apply1_kernel(Ctx ctx, int k) {
    __shared__ double scratch[2][kWarpsPerBlock];
    __shared__ double scratch2[kWarpsPerBlock];
    int flip = 0;
    LeanCursor cur;
    cur.init(ctx);
    bool fresh = true, active = false;
    double a = 0.0;
    const double *ro = nullptr, *ap = nullptr, *pn = nullptr;
    double *rnw = nullptr, *x = nullptr, *part_rr = nullptr;
    int n = 0;
    for (; cur.g < cur.g1; ++cur.g) {
        if (cur.advance(ctx) || fresh) {
            fresh = false;
            const SysDev* S = ctx.sys + cur.s;
            n = S->n;
            ro = S->r[k & 1], rnw = S->r[(k + 1) & 1], pn = S->p[(k + 1) & 1], ap = S->ap, x = S->x, part_rr = S->part_rr;
            active = kInit || ld_relaxed_s32(ctx.state + cur.s) == 0;  // written before the previous launch ended: uniform
            a = 0.0;
            if (!kInit && active) a = __ldcg(S->scal + (k & 1)) / block_reduce_array(S->part_pap, S->ntiles, scratch2);  // cg.py:78
        }
        const int tile = cur.g - cur.lo;
        if (!active) continue;
        const int row = tile * kTileRows + threadIdx.x;
        double rn = 0.0;
        if (row < n) {
            const double r_old = ro[row];
            if (kInit) {
                rn = r_old;
            } else {
                const double ap_i = ap[row], x_i = x[row], p_i = pn[row];
                rn = __dsub_rn(r_old, __dmul_rn(a, ap_i));   // cg.py:80
                x[row] = __dadd_rn(x_i, __dmul_rn(a, p_i));  // cg.py:79
            }
            rnw[row] = rn;
        }
        if (!kInit) {
            const double rr = tile_sum(__dmul_rn(rn, rn), scratch, flip);  // cg.py:86
            if (threadIdx.x == 0) part_rr[tile] = rr;
        }
        if (tile == 0) {  // one CTA per system and launch
            const SysDev* S = ctx.sys + cur.s;
            if (!kInit && threadIdx.x == 0 && S->coef) S->coef[2 * k] = a;
            if (kInit) {  // publish <b,b> once (all part_bb were written by the previous launch)
                const double bb = block_reduce_array(S->part_bb, S->ntiles, scratch2);
                if (threadIdx.x == 0) S->scal[2] = bb;
            }
        }
    }
}

template <bool kInit>
__global__ void __launch_bounds__(kBlock, 3) solve_dotrz_kernel(Ctx ctx, int k) {
    __shared__ double scratch[2][kWarpsPerBlock];
    int flip = 0;
    LeanCursor cur;
    cur.init(ctx);
    bool fresh = true, active = false;
    const double *rn = nullptr, *zn = nullptr;
    double *t = nullptr, *part_rz = nullptr, *part_rr = nullptr;
    int n = 0, rearm = 0;
    for (; cur.g < cur.g1; ++cur.g) {
        if (cur.advance(ctx) || fresh) {
            fresh = false;
            const SysDev* S = ctx.sys + cur.s;
            n = S->n, rearm = S->rearm_t;
            rn = S->r[(k + 1) & 1], zn = S->z[(k + 1) & 1], t = S->t, part_rz = S->part_rz[(k + 1) & 1], part_rr = S->part_rr;
            active = kInit || ld_relaxed_s32(ctx.state + cur.s) == 0;
        }
        const int tile = cur.g - cur.lo;
        if (!active) continue;
        const int row = tile * kTileRows + threadIdx.x;
        double ri = 0.0, zi = 0.0;
        if (row < n) {
            ri = rn[row], zi = zn[row];
            if (rearm) st_relaxed_u64(t + row, kPending);  // y for the next forward solve
        }
        const double rz = tile_sum(__dmul_rn(ri, zi), scratch, flip);  // cg.py:76,82
        if (threadIdx.x == 0) part_rz[tile] = rz;
        if (kInit) {  // iteration 0 checks <z0,z0> (cg.py:66)
            const double zz = tile_sum(__dmul_rn(zi, zi), scratch, flip);
            if (threadIdx.x == 0) part_rr[tile] = zz;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
static inline int64_t pad32(int64_t v) { return (v + 31) / 32 * 32; }
static inline int ntiles_of(int n) { return (n + kTileRows - 1) / kTileRows; }
constexpr int kTraceCap = 4096;

struct WsLayout {
    size_t word, n_done, meta, sys, state, tile_ofs, fwd_ofs, bwd_ofs, act_sys[2], act_ofs[2], trace, ts_word, ts_sys[4], total;
};
static WsLayout ws_layout(int nsys) {
    WsLayout w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += align_up(bytes, 256); return at; };
    const size_t ints = sizeof(int) * ((size_t)nsys + 1);
    w.word = take(8);
    w.n_done = take(4);
    w.meta = take(16);
    w.sys = take(sizeof(SysDev) * (size_t)nsys);
    w.state = take(ints);
    w.tile_ofs = take(ints);
    w.fwd_ofs = take(ints);
    w.bwd_ofs = take(ints);
    for (int i = 0; i < 2; ++i) w.act_sys[i] = take(ints), w.act_ofs[i] = take(ints);
    w.trace = take(sizeof(long long) * 2 * kTraceCap);
    w.ts_word = take(8);
    for (int i = 0; i < 4; ++i) w.ts_sys[i] = take(sizeof(TsSysDev) * (size_t)nsys);  // fwd/bwd x parity of (k + 1)
    w.total = off;
    return w;
}

// Every PCG kernel carries the tile pipeline's stages in dynamic shared memory (> 48 KB: opt in once per kernel).
template <class Kernel>
static int allow_smem(Kernel kernel) {
    return allow_dynamic_smem((const void*)kernel, sizeof(Smem));  // cached per (device, kernel)
}

template <int kPhase, bool kInit>
static int launch_phase(const Ctx& ctx, int k, int grid, bool cooperative, cudaStream_t s) {
    int st = allow_smem(pcg_phase_kernel<kPhase, kInit>);
    if (st != DP_OK) return st;
    if (cooperative) {
        Ctx c = ctx;
        void* args[] = {&c, &k};
        DP_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_phase_kernel<kPhase, kInit>, dim3(grid), dim3(kBlock), args,
                                            sizeof(Smem), s));
    } else {
        pcg_phase_kernel<kPhase, kInit><<<grid, kBlock, sizeof(Smem), s>>>(ctx, k);
        DP_LAUNCH_CHECK();
    }
    return DP_OK;
}

// Tile-stream solves of the stepped engine (DP_SOLVE_TILE_STREAM): device descriptor arrays of the SOLVE systems,
// forward (r_new -> y) and backward (y -> z_new), one pair per parity of k + 1 (r and z are double buffered).
struct TsPhases {
    const TsSysDev* fwd[2];
    const TsSysDev* bwd[2];
    int nsys, max_tiles, nmax;
    bool short_rows;  // every system promises DP_TRSV_SHORT_ROWS (solve_algorithm bit 1)
    bool all_solve;   // the batch holds nothing but tile-stream SOLVE systems: APPLY1 / DOTRZ are pure vector passes
    unsigned long long* word;
};

template <bool kInit>
static int launch_lean(bool dotrz, const Ctx& ctx, int k, cudaStream_t s) {
    const void* kernel = dotrz ? (const void*)solve_dotrz_kernel<kInit> : (const void*)solve_apply1_kernel<kInit>;
    int grid = coop_grid(kernel, kBlock, 0);  // resident CTAs (cached per device): one contiguous tile range each
    if (grid > ctx.total_tiles) grid = ctx.total_tiles;
    if (dotrz)
        solve_dotrz_kernel<kInit><<<grid, kBlock, 0, s>>>(ctx, k);
    else
        solve_apply1_kernel<kInit><<<grid, kBlock, 0, s>>>(ctx, k);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

template <bool kInit>
static int launch_apply(const Ctx& ctx, int k, int tile_grid, int coop, const TsPhases* ts, cudaStream_t s) {
    int st;
    const bool lean = ts != nullptr && ts->all_solve;  // every system is a tile-stream SOLVE system: pure vector passes
    if (lean) {
        if ((st = launch_lean<kInit>(false, ctx, k, s)) != DP_OK) return st;
    } else if ((st = launch_phase<PH_APPLY1, kInit>(ctx, k, tile_grid, false, s)) != DP_OK) return st;
    if (ctx.has_multiply && (st = launch_phase<PH_APPLY2, kInit>(ctx, k, tile_grid, false, s)) != DP_OK) return st;
    if (ctx.has_solve) {
        if (ts) {  // finished systems are solved along (their vectors are scratch by then): no host round trip
            const int par = (k + 1) & 1;
            // no arming launches: y is armed by PH_INIT and re-armed by PH_DOTRZ, z_new by APPLY1 (the lines the sync-free
            // solves rely on as well)
            if ((st = ts_solve_launch(ts->fwd[par], ts->nsys, ts->max_tiles, ts->nmax, ts->short_rows, false, ts->word, ctx.flag, s)) != DP_OK) return st;
            if ((st = ts_solve_launch(ts->bwd[par], ts->nsys, ts->max_tiles, ts->nmax, ts->short_rows, false, ts->word, ctx.flag, s)) != DP_OK) return st;
        } else {
            if ((st = launch_phase<PH_FWD, kInit>(ctx, k, coop, true, s)) != DP_OK) return st;
            if ((st = launch_phase<PH_BWD, kInit>(ctx, k, coop, true, s)) != DP_OK) return st;
        }
        if (lean) {
            if ((st = launch_lean<kInit>(true, ctx, k, s)) != DP_OK) return st;
        } else if ((st = launch_phase<PH_DOTRZ, kInit>(ctx, k, tile_grid, false, s)) != DP_OK) return st;
    }
    return DP_OK;
}

}  // namespace dp

using namespace dp;

extern "C" {

int64_t dp_pcg_work_doubles(int32_t n) {
    if (n < 0) return -1;
    return 8 * pad32(n) + 5 * pad32(ntiles_of(n)) + 32;
}

size_t dp_pcg_workspace_bytes(int32_t nsys) { return ws_layout(nsys < 0 ? 0 : nsys).total; }

int dp_device_info(int* sm_count_host, int* pcg_ctas_per_sm_host, int* l2_bytes_host) {
    int dev = 0, sms = 0, l2 = 0;
    DP_CUDA(cudaGetDevice(&dev));
    DP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DP_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    if (sm_count_host) *sm_count_host = sms;
    if (l2_bytes_host) *l2_bytes_host = l2;
    if (pcg_ctas_per_sm_host) {
        if (allow_smem(pcg_fused_kernel<false, false>) != DP_OK) return DP_ERR_CUDA;
        *pcg_ctas_per_sm_host = coop_grid((const void*)pcg_fused_kernel<false, false>, kBlock, sizeof(Smem)) / (sms > 0 ? sms : 1);
    }
    return DP_OK;
}

int dp_debug_pcg_trace(const void* workspace, int32_t nsys, int64_t* out_host, int32_t capacity) {
    if (!workspace || !out_host || capacity <= 0 || nsys <= 0) return DP_ERR_INVALID;
    const WsLayout lay = ws_layout(nsys);
    const int cap = capacity < kTraceCap ? capacity : kTraceCap;
    DP_CUDA(cudaMemcpy(out_host, static_cast<const char*>(workspace) + lay.trace, sizeof(long long) * 2 * (size_t)cap,
                       cudaMemcpyDeviceToHost));
    return DP_OK;
}

int dp_debug_pipe_trace(uint64_t* out_host, int32_t capacity, int32_t unit) {
#ifdef DPCG_PIPE_TRACE
    if (!out_host || capacity <= 0) return DP_ERR_INVALID;
    if (unit == 1) return dp::pipe_trace_read_spmv(out_host, capacity);  // (every translation unit has its own trace buffer)
    const int cap = capacity < kPipeTraceCap ? capacity : kPipeTraceCap;
    for (int w = 0; w < 2; ++w)
        DP_CUDA(cudaMemcpyFromSymbol(out_host + (size_t)w * capacity, g_pipe_trace, sizeof(unsigned long long) * (size_t)cap,
                                     sizeof(unsigned long long) * (size_t)w * kPipeTraceCap, cudaMemcpyDeviceToHost));
    return DP_OK;
#else
    (void)out_host, (void)capacity, (void)unit;
    return DP_ERR_INVALID;  // built without -DDPCG_PIPE_TRACE
#endif
}

int dp_pcg_solve_f64(const dp_pcg_system_t* systems_host, int32_t nsys, const dp_pcg_params_t* params_host,
                     int32_t* flag_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!systems_host || nsys <= 0 || !params_host || !flag_out || !workspace) return DP_ERR_INVALID;
    if (params_host->max_iter < 0) return DP_ERR_INVALID;
    if (!aligned16(workspace)) return DP_ERR_ALIGNMENT;
    const WsLayout lay = ws_layout(nsys);
    if (workspace_bytes < lay.total) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);

    std::vector<SysDev> sys((size_t)nsys);
    std::vector<int> tile_ofs((size_t)nsys + 1, 0), fwd_ofs((size_t)nsys + 1, 0), bwd_ofs((size_t)nsys + 1, 0);
    std::vector<int> ident((size_t)nsys + 1, 0);
    int has_multiply = 0, has_solve = 0, has_ls = 0, n_solve = 0, n_ts = 0;
    bool packed = true;  // every streamed matrix of every system comes with its packed copy
    long long working_set_bytes = 0, matrix_entries = 0;
    long long sum_fwd_lvl = 0, sum_bwd_lvl = 0;
    std::vector<TsSysDev> ts_sys[4];
    std::vector<int> ts_owner;  // system index of each tile-stream descriptor
    bool ts_short = true;
    for (int i = 0; i < nsys; ++i) {
        const dp_pcg_system_t& u = systems_host[i];
        if (u.n <= 0 || !u.a_rowptr || !u.a_col || !u.a_val || !u.b || !u.x || !u.work || !u.iters_out || !u.res_out)
            return DP_ERR_INVALID;
        if (!aligned16(u.a_col) || !aligned16(u.a_val) || !aligned16(u.work)) return DP_ERR_ALIGNMENT;
        SysDev d{};
        d.n = u.n;
        d.precond = u.precond;
        d.ntiles = ntiles_of(u.n);
        d.A = CsrView{u.a_rowptr, u.a_col, u.a_val, u.n, u.a_nnz, u.a_col16, u.a_val32, u.a_tile_base};
        d.M = CsrView{u.m_rowptr, u.m_col, u.m_val, u.n, u.m_nnz, u.m_col16, u.m_val32, u.m_tile_base};
        d.Mt = CsrView{u.mt_rowptr, u.mt_col, u.mt_val, u.n, u.mt_nnz, u.mt_col16, u.mt_val32, u.mt_tile_base};
        {
            auto has_pack = [](const CsrView& V) { return V.col16 && V.val32 && V.tbase && aligned16(V.col16) && aligned16(V.val32); };
            if (!has_pack(d.A)) packed = false;
            if ((u.precond == DP_PRECOND_MULTIPLY || u.precond == DP_PRECOND_CSR) && !has_pack(d.M)) packed = false;
            if (u.precond == DP_PRECOND_MULTIPLY && !has_pack(d.Mt)) packed = false;
            if (u.precond == DP_PRECOND_SOLVE) packed = false;
        }
        d.dinv = u.dinv;
        d.fwd_plan = u.fwd_plan;
        d.bwd_plan = u.bwd_plan;
        d.fwd_ls = LsFactor{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
        d.bwd_ls = d.fwd_ls;
        const bool tile_stream = u.precond == DP_PRECOND_SOLVE && (u.solve_algorithm & DP_SOLVE_TILE_STREAM) != 0;
        if (tile_stream && !(u.solve_algorithm & DP_SOLVE_SHORT_ROWS)) ts_short = false;
        if (u.precond == DP_PRECOND_SOLVE) ++n_solve;
        if (tile_stream) {
            if (!u.fwd_ls_rowptr || !u.fwd_ls_col || !u.fwd_ls_val || !u.bwd_ls_rowptr || !u.bwd_ls_col || !u.bwd_ls_val)
                return DP_ERR_INVALID;
            if (!aligned16(u.fwd_ls_rowptr) || !aligned16(u.fwd_ls_col) || !aligned16(u.fwd_ls_val) ||
                !aligned16(u.bwd_ls_rowptr) || !aligned16(u.bwd_ls_col) || !aligned16(u.bwd_ls_val))
                return DP_ERR_ALIGNMENT;
            ++n_ts;
        } else if (u.precond == DP_PRECOND_SOLVE) {
            if (u.fwd_ls_rowptr && u.fwd_ls_col && u.fwd_ls_val) {  // perm == NULL: the system is in this solve's level order
                if (!aligned16(u.fwd_ls_col) || !aligned16(u.fwd_ls_val) || !aligned16(u.fwd_ls_rowptr)) return DP_ERR_ALIGNMENT;
                d.fwd_ls = LsFactor{u.fwd_ls_rowptr, u.fwd_ls_col, u.fwd_ls_val, u.fwd_ls_perm, u.fwd_ls_level, u.n, u.m_nnz};
                has_ls = 1;
            }
            if (u.bwd_ls_rowptr && u.bwd_ls_col && u.bwd_ls_val) {
                if (!aligned16(u.bwd_ls_col) || !aligned16(u.bwd_ls_val) || !aligned16(u.bwd_ls_rowptr)) return DP_ERR_ALIGNMENT;
                d.bwd_ls = LsFactor{u.bwd_ls_rowptr, u.bwd_ls_col, u.bwd_ls_val, u.bwd_ls_perm, u.bwd_ls_level, u.n, u.mt_nnz};
                has_ls = 1;
            }
        }
        // y is polled by a forward solve that is sync-free (or tile-stream) and must be re-armed after every backward
        // solve that does not consume-and-re-arm it itself (level-stream, tile-stream)
        d.rearm_t = tile_stream || (u.precond == DP_PRECOND_SOLVE && !d.fwd_ls.rowptr && d.bwd_ls.rowptr) ? 1 : 0;
        d.fwd_look = u.fwd_max_level_chunks > 0 ? u.fwd_max_level_chunks : 1;
        d.bwd_look = u.bwd_max_level_chunks > 0 ? u.bwd_max_level_chunks : 1;
        int fwd_chunks = 0, bwd_chunks = 0;
        switch (u.precond) {
            case DP_PRECOND_IDENTITY: break;
            case DP_PRECOND_JACOBI:
                if (!u.dinv) return DP_ERR_INVALID;
                break;
            case DP_PRECOND_CSR:
                if (!u.m_rowptr || !u.m_col || !u.m_val) return DP_ERR_INVALID;
                if (!aligned16(u.m_col) || !aligned16(u.m_val)) return DP_ERR_ALIGNMENT;
                break;
            case DP_PRECOND_SOLVE:
                if (!u.fwd_plan || !u.bwd_plan || u.fwd_nchunks <= 0 || u.bwd_nchunks <= 0) return DP_ERR_INVALID;
                fwd_chunks = u.fwd_nchunks;
                bwd_chunks = u.bwd_nchunks;
                sum_fwd_lvl += u.fwd_max_level_chunks > 0 ? u.fwd_max_level_chunks : (1 << 20);
                sum_bwd_lvl += u.bwd_max_level_chunks > 0 ? u.bwd_max_level_chunks : (1 << 20);
                has_solve = 1;
                [[fallthrough]];  // needs L and L^T as well
            case DP_PRECOND_MULTIPLY:
                if (!u.m_rowptr || !u.m_col || !u.m_val || !u.mt_rowptr || !u.mt_col || !u.mt_val) return DP_ERR_INVALID;
                if (!aligned16(u.m_col) || !aligned16(u.m_val) || !aligned16(u.mt_col) || !aligned16(u.mt_val))
                    return DP_ERR_ALIGNMENT;
                if (u.precond == DP_PRECOND_MULTIPLY) has_multiply = 1;
                break;
            default: return DP_ERR_INVALID;
        }
        d.b = u.b;
        d.x = u.x;
        const long long entries = (long long)u.a_nnz + (u.precond >= DP_PRECOND_MULTIPLY ? u.m_nnz : 0) +
                                  (u.precond == DP_PRECOND_MULTIPLY || u.precond == DP_PRECOND_SOLVE ? u.mt_nnz : 0);
        matrix_entries += entries;
        working_set_bytes += 12ll * entries + 8ll * 10 * u.n;
        const int64_t np = pad32(u.n), tp = pad32(d.ntiles);
        double* w = u.work;
        d.r[0] = w; d.r[1] = w + np; d.p[0] = w + 2 * np; d.p[1] = w + 3 * np;
        d.z[0] = w + 4 * np;
        d.z[1] = u.precond == DP_PRECOND_SOLVE ? w + 5 * np : d.z[0];
        d.ap = w + 6 * np; d.t = w + 7 * np;
        double* q = w + 8 * np;
        d.part_rr = q; d.part_pap = q + tp; d.part_bb = q + 2 * tp; d.part_rz[0] = q + 3 * tp; d.part_rz[1] = q + 4 * tp;
        d.scal = q + 5 * tp;
        d.iters_out = u.iters_out;
        d.res_out = u.res_out;
        d.history = u.history;
        d.coef = u.coef;
        if (tile_stream) {  // y = t, z double buffered like r: the vectors are the iteration's own (position space)
            for (int par = 0; par < 2; ++par) {
                TsSysDev f{}, g{};
                f.F = LsFactor{u.fwd_ls_rowptr, u.fwd_ls_col, u.fwd_ls_val, nullptr, nullptr, u.n, u.m_nnz};
                f.b = d.r[par], f.x = d.t, f.upper = 0, f.ntiles = d.ntiles, f.rev = 0;
                f.skip = g.skip = reinterpret_cast<const int*>(ws + lay.state) + i;  // converged systems leave the solves at once
                g.F = LsFactor{u.bwd_ls_rowptr, u.bwd_ls_col, u.bwd_ls_val, nullptr, nullptr, u.n, u.mt_nnz};
                g.b = d.t, g.x = d.z[par], g.upper = 1, g.ntiles = d.ntiles, g.rev = 1;
                ts_sys[par].push_back(f);
                ts_sys[2 + par].push_back(g);
            }
            ts_owner.push_back(i);
        }
        sys[(size_t)i] = d;
        ident[(size_t)i] = i;
        tile_ofs[(size_t)i + 1] = tile_ofs[(size_t)i] + d.ntiles;
        fwd_ofs[(size_t)i + 1] = fwd_ofs[(size_t)i] + fwd_chunks;
        bwd_ofs[(size_t)i + 1] = bwd_ofs[(size_t)i] + bwd_chunks;
    }

    // the tile-stream solves replace the FWD/BWD phases of the stepped engine for ALL solve-mode systems of the batch
    if (n_ts && (n_ts != n_solve || params_host->engine != DP_ENGINE_STEPPED)) return DP_ERR_INVALID;
    {
        const char* e = getenv("DPCG_NO_PACK");  // experiments / A-B runs: ignore the packed copies
        if ((e && e[0] == '1') || params_host->engine != DP_ENGINE_FUSED) packed = false;
    }
    const void* fused = has_solve ? (const void*)pcg_fused_kernel<true, false>
                        : packed  ? (const void*)pcg_fused_kernel<false, true>
                                  : (const void*)pcg_fused_kernel<false, false>;
    if (allow_dynamic_smem(fused, sizeof(Smem)) != DP_OK || allow_smem(pcg_phase_kernel<PH_FWD, false>) != DP_OK) return DP_ERR_CUDA;
    const int fused_threads = (!has_solve && packed) ? kFusedThreadsPacked : kBlock;
    const int coop = coop_grid(fused, fused_threads, sizeof(Smem));
    const int coop_phase = coop_grid((const void*)pcg_phase_kernel<PH_FWD, false>, kBlock, sizeof(Smem));
    auto clamp_pw = [](long long lvl_chunks, int grid) {
        long long pw = (long long)trsv_lookahead() * lvl_chunks;
        if (pw < 32) pw = 32;
        const long long w = (long long)grid * kWarpsPerBlock;
        return (int)(pw > w ? w : pw);
    };

    Ctx ctx{};
    ctx.sys = reinterpret_cast<const SysDev*>(ws + lay.sys);
    ctx.state = reinterpret_cast<int*>(ws + lay.state);
    ctx.tile_ofs = reinterpret_cast<const int*>(ws + lay.tile_ofs);
    ctx.fwd_ofs = reinterpret_cast<const int*>(ws + lay.fwd_ofs);
    ctx.bwd_ofs = reinterpret_cast<const int*>(ws + lay.bwd_ofs);
    for (int i = 0; i < 2; ++i) {
        ctx.act_sys[i] = reinterpret_cast<int*>(ws + lay.act_sys[i]);
        ctx.act_ofs[i] = reinterpret_cast<int*>(ws + lay.act_ofs[i]);
    }
    ctx.act_meta = reinterpret_cast<int*>(ws + lay.meta);
    ctx.nsys = nsys;
    ctx.total_tiles = tile_ofs[(size_t)nsys];
    ctx.total_fwd = fwd_ofs[(size_t)nsys];
    ctx.total_bwd = bwd_ofs[(size_t)nsys];
    ctx.has_multiply = has_multiply;
    ctx.has_solve = has_solve;
    ctx.has_ls = has_ls;
    {   // bytes one iteration touches: matrices (12 B per entry) + the 8 work vectors, b and x
        int dev = 0, l2 = 0;
        DP_CUDA(cudaGetDevice(&dev));
        DP_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
        const char* e = getenv("DPCG_KEEP_L2");  // experiments: 0 / 1 forces the policy
        if (packed) working_set_bytes -= matrix_entries * 6;
        ctx.keep_l2 = e ? (e[0] == '1') : (working_set_bytes * 10 < (long long)l2 * 6);
    }
    ctx.rtol = params_host->rtol;
    ctx.max_iter = params_host->max_iter;
    ctx.word = reinterpret_cast<unsigned long long*>(ws + lay.word);
    ctx.n_done = reinterpret_cast<int*>(ws + lay.n_done);
    ctx.flag = flag_out;
    ctx.trace = reinterpret_cast<long long*>(ws + lay.trace);
    const char* tr = getenv("DPCG_TRACE");
    ctx.trace_cap = (tr && tr[0] == '1') ? kTraceCap : 0;

    const int meta[4] = {nsys, ctx.total_tiles, 0, 0};
    DP_CUDA(cudaMemsetAsync(ws + lay.word, 0, 8, s));
    DP_CUDA(cudaMemsetAsync(ws + lay.n_done, 0, 4, s));
    if (ctx.trace_cap) DP_CUDA(cudaMemsetAsync(ws + lay.trace, 0, sizeof(long long) * 2 * kTraceCap, s));
    DP_CUDA(cudaMemcpyAsync(ws + lay.meta, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaMemcpyAsync(ws + lay.sys, sys.data(), sizeof(SysDev) * (size_t)nsys, cudaMemcpyHostToDevice, s));
    const size_t ints = sizeof(int) * ((size_t)nsys + 1);
    DP_CUDA(cudaMemcpyAsync(ws + lay.tile_ofs, tile_ofs.data(), ints, cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaMemcpyAsync(ws + lay.fwd_ofs, fwd_ofs.data(), ints, cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaMemcpyAsync(ws + lay.bwd_ofs, bwd_ofs.data(), ints, cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaMemcpyAsync(ws + lay.act_sys[0], ident.data(), ints, cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaMemcpyAsync(ws + lay.act_ofs[0], tile_ofs.data(), ints, cudaMemcpyHostToDevice, s));

    if (params_host->engine == DP_ENGINE_FUSED) {
        // right-size the grid: a CTA per tile is enough for the tile phases; the SpTRSV phases spread their
        // participating warps one per CTA first. Fewer CTAs = cheaper grid barrier for small systems.
        int grid = ctx.total_tiles;
        if (has_solve) {
            const long long want = (long long)trsv_lookahead() * (sum_fwd_lvl > sum_bwd_lvl ? sum_fwd_lvl : sum_bwd_lvl);
            if (want > grid) grid = want > coop ? coop : (int)want;
        }
        if (grid > coop) grid = coop;
        if (grid < 1) grid = 1;
        ctx.pw_fwd = clamp_pw(sum_fwd_lvl, grid);
        ctx.pw_bwd = clamp_pw(sum_bwd_lvl, grid);
        void* args[] = {&ctx};
        DP_CUDA(cudaLaunchCooperativeKernel(fused, dim3(grid), dim3(fused_threads), args, sizeof(Smem), s));
        return DP_OK;
    }
    if (params_host->engine != DP_ENGINE_STEPPED) return DP_ERR_INVALID;

    ctx.pw_fwd = clamp_pw(sum_fwd_lvl, coop_phase);
    ctx.pw_bwd = clamp_pw(sum_bwd_lvl, coop_phase);
    TsPhases ts_phases{};
    const TsPhases* ts = nullptr;
    auto ts_upload = [&](const std::vector<int>& keep) -> int {  // descriptors of the systems still iterating
        std::vector<TsSysDev> act;
        ts_phases.nsys = (int)keep.size(), ts_phases.max_tiles = 0, ts_phases.nmax = 0;
        for (int q = 0; q < 4; ++q) {
            act.clear();
            for (int idx : keep) act.push_back(ts_sys[q][(size_t)idx]);
            // pageable source: the call returns once the bytes are staged
            DP_CUDA(cudaMemcpyAsync(ws + lay.ts_sys[q], act.data(), sizeof(TsSysDev) * act.size(), cudaMemcpyHostToDevice, s));
        }
        for (int idx : keep) {
            const TsSysDev& f = ts_sys[0][(size_t)idx];
            if (f.ntiles > ts_phases.max_tiles) ts_phases.max_tiles = f.ntiles;
            if (f.F.n > ts_phases.nmax) ts_phases.nmax = f.F.n;
        }
        return DP_OK;
    };
    std::vector<int> ts_keep;
    if (n_ts) {
        DP_CUDA(cudaMemsetAsync(ws + lay.ts_word, 0, 8, s));
        for (int q = 0; q < n_ts; ++q) ts_keep.push_back(q);
        if (ts_upload(ts_keep) != DP_OK) return DP_ERR_CUDA;
        for (int par = 0; par < 2; ++par) {
            ts_phases.fwd[par] = reinterpret_cast<const TsSysDev*>(ws + lay.ts_sys[par]);
            ts_phases.bwd[par] = reinterpret_cast<const TsSysDev*>(ws + lay.ts_sys[2 + par]);
        }
        ts_phases.short_rows = ts_short;
        ts_phases.all_solve = n_ts == nsys;
        ts_phases.word = reinterpret_cast<unsigned long long*>(ws + lay.ts_word);
        ts = &ts_phases;
    }
    const int tile_grid = ctx.total_tiles < coop * 4 ? ctx.total_tiles : coop * 4;
    const int every = params_host->check_every > 0 ? params_host->check_every : 1;
    int st;
    if ((st = launch_phase<PH_INIT, false>(ctx, -1, tile_grid, false, s)) != DP_OK) return st;
    if ((st = launch_apply<true>(ctx, -1, tile_grid, coop_phase, ts, s)) != DP_OK) return st;
    for (int k = 0; k <= ctx.max_iter; ++k) {
        if ((st = launch_phase<PH_A, false>(ctx, k, tile_grid, false, s)) != DP_OK) return st;
        if (k % every == 0 || k == ctx.max_iter) {  // the one host sync per convergence check
            int done = 0;
            DP_CUDA(cudaMemcpyAsync(&done, ctx.n_done, sizeof(int), cudaMemcpyDeviceToHost, s));
            DP_CUDA(cudaStreamSynchronize(s));
            if (done >= nsys) break;
            if (ts && done > nsys - (int)ts_keep.size()) {  // some systems finished: stop solving them along
                std::vector<int> state((size_t)nsys);
                DP_CUDA(cudaMemcpyAsync(state.data(), ctx.state, sizeof(int) * (size_t)nsys, cudaMemcpyDeviceToHost, s));
                DP_CUDA(cudaStreamSynchronize(s));
                std::vector<int> keep;
                for (int idx : ts_keep)
                    if (state[(size_t)ts_owner[(size_t)idx]] == 0) keep.push_back(idx);
                if (!keep.empty() && keep.size() < ts_keep.size()) {
                    ts_keep.swap(keep);
                    if (ts_upload(ts_keep) != DP_OK) return DP_ERR_CUDA;
                }
            }
        }
        if ((st = launch_apply<false>(ctx, k, tile_grid, coop_phase, ts, s)) != DP_OK) return st;
    }
    return DP_OK;
}

}  // extern "C"
