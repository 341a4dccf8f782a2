// sptrsv.cu — standalone K4 entry point (dp_sptrsv_solve_f64) and the sync-free IC(0) factorisation (dp_ic0_f64).
#include <stdlib.h>

#include <vector>

#include "sptrsv.cuh"
#include "trsv_ls.cuh"
#include "trsv_ts.cuh"

namespace dp {

__global__ void fill_u64_kernel(unsigned long long* __restrict__ p, long long count, unsigned long long v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// Participating warps: `pw` of them, spread one per CTA first (gw = warp * gridDim + block), chunk c -> warp c % pw.
__global__ void __launch_bounds__(kBlock, 2)
sptrsv_kernel(CsrView T, int upper, const int* __restrict__ plan, long long nchunks, int pw, int lookback,
              const double* __restrict__ b, double* x, unsigned long long* word, int* flag) {
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    if (gw >= pw) return;
    const AbortCtl ctl{word, flag};
    const RhsPlain rhs{b};
    if (upper)
        sptrsv_stream<true>(T, plan, gw, pw, nchunks, lookback, rhs, x, ctl);
    else
        sptrsv_stream<false>(T, plan, gw, pw, nchunks, lookback, rhs, x, ctl);
}

// Batch of independent solves in one launch: the resident warps are dealt to the systems (system s gets warps
// s, s + nsys, ...), so the solves advance side by side and the HBM stream of the batch hides each system's
// level-by-level critical path (BASELINE configs 3 and 5: many systems per GPU).
struct TrsvSysDev {
    CsrView T;
    const int* plan;
    const double* b;
    double* x;
    long long nchunks;
    int upper, lookback;
};

__global__ void fill_pending_batch_kernel(const TrsvSysDev* __restrict__ sys, int nsys) {
    for (int s = blockIdx.y; s < nsys; s += gridDim.y) {
        unsigned long long* x = reinterpret_cast<unsigned long long*>(sys[s].x);
        const int n = sys[s].T.n;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = kPending;
    }
}

__global__ void __launch_bounds__(kBlock, 2)
sptrsv_batch_kernel(const TrsvSysDev* __restrict__ sys, int nsys, int wps, unsigned long long* word, int* flag) {
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    const int total = gridDim.x * kWarpsPerBlock;
    const AbortCtl ctl{word, flag};
    const int local = total >= nsys ? gw / nsys : 0;
    if (local >= wps) return;
    for (int s = total >= nsys ? gw % nsys : gw; s < nsys; s += total) {
        const TrsvSysDev S = sys[s];
        const int team = min((long long)wps, S.nchunks) > 0 ? (int)min((long long)wps, S.nchunks) : 1;
        if (local >= team) continue;
        const RhsPlain rhs{S.b};
        bool ok;
        if (S.upper)
            ok = sptrsv_stream<true>(S.T, S.plan, local, team, S.nchunks, S.lookback, rhs, S.x, ctl);
        else
            ok = sptrsv_stream<false>(S.T, S.plan, local, team, S.nchunks, S.lookback, rhs, S.x, ctl);
        if (!ok) return;
    }
}

// ---- level-stream solve (trsv_ls.cuh): one CTA per system, systems dealt round robin -----------------------------
struct LsSysDev {
    LsFactor F;
    const double* b;
    double* x;
    int upper, pad;
};

struct LsSmem {
    alignas(16) unsigned char bytes[LsGeom<kLsStagesAlone>::kBytes];
    LsSharedT<kLsStagesAlone> ls;
};
constexpr int kLsThreads = kBlock + kWarp;  // 16 warps of rows + the producer warp

// One CTA per SM: a solve is latency bound and wants the whole SM's shared memory for a deep pipeline; bigger batches loop.
__global__ void __launch_bounds__(kLsThreads, 1) sptrsv_ls_batch_kernel(const LsSysDev* __restrict__ sys, int nsys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LsSmem& sm = *reinterpret_cast<LsSmem*>(smem_raw);
    if (threadIdx.x == 0) sm.ls.init();
    __syncthreads();
    for (int s = blockIdx.x; s < nsys; s += gridDim.x) {
        const LsSysDev S = sys[s];
        if (S.upper)
            trsv_level_stream<true, kLsStagesAlone>(S.F, S.b, S.x, sm.bytes, sm.ls);
        else
            trsv_level_stream<false, kLsStagesAlone>(S.F, S.b, S.x, sm.bytes, sm.ls);
    }
}

// ---- tile-stream solve (trsv_ts.cuh): the tiles of all systems in one sequence, dealt to persistent CTAs ---------
__global__ void arm_positions_kernel(const TsSysDev* __restrict__ sys, int nsys) {
    for (int s = blockIdx.y; s < nsys; s += gridDim.y) {
        unsigned long long* x = reinterpret_cast<unsigned long long*>(sys[s].x);
        const int n = sys[s].F.n;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = kPending;
    }
}

// Vectors in the ORIGINAL numbering: b_pos = b[perm] before the solve, x[perm] = x_pos after it, in passes of their own
// and one system after the other (blockIdx.y would interleave the systems: the point is that the 4 consecutive levels
// which share a sector of b / x stay in L2 between their visits - true for one system's fronts, not for a batch's).
struct TsPermDev {
    const int* perm;
    const double* b;  // original numbering
    double* x;        // original numbering
    double* b_pos;    // workspace
    double* x_pos;    // workspace
    int n, pad;
};
// (One row per thread and step: eight rows in flight per thread, a grid stride apart, were measured 40 % SLOWER -
// neighbouring positions of a level are neighbouring rows of a plane, and the unrolled form visits them far apart.)
__global__ void ts_gather_kernel(const TsPermDev* __restrict__ sys, int nsys) {
    for (int s = 0; s < nsys; ++s) {
        const TsPermDev S = sys[s];
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < S.n; p += gridDim.x * blockDim.x)
            S.b_pos[p] = __ldg(S.b + __ldg(S.perm + p));
    }
}
__global__ void ts_scatter_kernel(const TsPermDev* __restrict__ sys, int nsys) {
    for (int s = 0; s < nsys; ++s) {
        const TsPermDev S = sys[s];
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < S.n; p += gridDim.x * blockDim.x)
            S.x[__ldg(S.perm + p)] = S.x_pos[p];
    }
}

template <bool kShort>
__global__ void __launch_bounds__(kTsThreads, 2)
sptrsv_ts_batch_kernel(const TsSysDev* __restrict__ sys, int nsys, int max_tiles, unsigned long long* word, int* flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TsSmem& sm = *reinterpret_cast<TsSmem*>(smem_raw);
    const AbortCtl ctl{word, flag};
    trsv_tile_stream<kShort>(sys, nsys, max_tiles, sm, ctl);
}

// IC(0), up-looking, one lane per row, rows of a level per chunk (same plan as the forward solve of tril(A)):
//   L_ij = (A_ij - sum_{k<j} L_ik L_jk) / L_jj,   L_ii = sqrt(A_ii - sum_k L_ik^2)      (oracle_ic0)
// Row j is complete when its diagonal (last entry) leaves the kPending state; the producer fences between the
// off-diagonal stores and the diagonal store, the consumer between seeing the diagonal and reading the row.
__global__ void __launch_bounds__(kBlock, 2)
ic0_kernel(CsrView A, double* l, const int* __restrict__ plan, long long nchunks, int pw, unsigned long long* word,
           int* flag) {
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    if (gw >= pw) return;
    const int lane = threadIdx.x & 31;
    const AbortCtl ctl{word, flag};
    for (long long c = gw; c < nchunks; c += pw) {
        const int row = __ldg(plan + c * 32 + lane);
        const bool valid = row >= 0;
        int rs = 0, re = 0;
        if (valid) {
            rs = __ldg(A.rowptr + row);
            re = __ldg(A.rowptr + row + 1);
        }
        int p = rs;
        unsigned idle = 0;
        for (;;) {
            const bool pending = valid && p < re;
            if (!__any_sync(kFull, pending)) break;
            bool progress = false;
            if (pending) {
                const int j = __ldg(A.col + p);
                double s = __ldg(A.val + p);
                if (j < row) {
                    const int js = __ldg(A.rowptr + j), dj = __ldg(A.rowptr + j + 1) - 1;
                    const unsigned long long ujj = ld_relaxed_u64(l + dj);
                    if (ujj != kPending) {
                        __threadfence();
                        int pi = rs, pj = js;
                        while (pi < p && pj < dj) {
                            const int ci = __ldg(A.col + pi), cj = __ldg(A.col + pj);
                            if (ci == cj) {
                                s = __dsub_rn(s, __dmul_rn(l[pi], as_double(ld_relaxed_u64(l + pj))));
                                ++pi, ++pj;
                            } else if (ci < cj) ++pi; else ++pj;
                        }
                        l[p] = __ddiv_rn(s, as_double(ujj));
                        ++p, progress = true;
                    }
                } else {  // diagonal: last entry of the row
                    for (int pi = rs; pi < p; ++pi) s = __dsub_rn(s, __dmul_rn(l[pi], l[pi]));
                    if (!(s > 0.0)) atomicCAS(flag, 0, (int)DP_ERR_STRUCTURE);
                    __threadfence();
                    st_relaxed_u64(l + p, as_bits(__dsqrt_rn(s)));
                    ++p, progress = true;
                }
            }
            if (!__any_sync(kFull, progress)) {
                if (++idle > kSpinBudget) {
                    ctl.raise(DP_ERR_TIMEOUT);
                    return;
                }
                if ((idle & 1023u) == 0 && ctl.aborted()) return;
                __nanosleep(64);
            }
        }
    }
}

int trsv_lookahead() {  // levels' worth of warps that take part in a solve (DPCG_TRSV_LOOKAHEAD: experiments)
    static int cached = 0;
    if (!cached) {
        const char* e = getenv("DPCG_TRSV_LOOKAHEAD");
        cached = e && atoi(e) > 0 ? atoi(e) : 4;
    }
    return cached;
}

static int participating_warps(int max_level_chunks, int grid) {
    const int w = grid * kWarpsPerBlock;
    if (max_level_chunks <= 0) return w;
    long long pw = (long long)trsv_lookahead() * max_level_chunks;
    if (pw < 32) pw = 32;
    return pw > w ? w : (int)pw;
}

int ts_solve_launch(const TsSysDev* sys_dev, int nsys, int max_tiles, int nmax, bool short_rows, bool arm,
                    unsigned long long* word, int* flag, cudaStream_t s) {
    if (arm) {
        const int fill_x = (nmax + 255) / 256 < sm_count() * 4 ? (nmax + 255) / 256 : sm_count() * 4;
        arm_positions_kernel<<<dim3(fill_x, nsys < 1024 ? nsys : 1024), 256, 0, s>>>(sys_dev, nsys);
        DP_LAUNCH_CHECK();
    }
    const void* kernel = short_rows ? (const void*)sptrsv_ts_batch_kernel<true> : (const void*)sptrsv_ts_batch_kernel<false>;
    if (allow_dynamic_smem(kernel, sizeof(TsSmem)) != DP_OK) return DP_ERR_CUDA;
    int grid = coop_grid(kernel, kTsThreads, sizeof(TsSmem));  // cached per device
    if (const char* e = getenv("DPCG_TS_GRID")) {  // experiments: fewer CTAs = fewer pollers
        const int g = atoi(e);
        if (g > 0 && g < grid) grid = g;
    }
    const long long items = (long long)max_tiles * nsys;
    if (items >= (1ll << 31) - 1024) return DP_ERR_INVALID;  // the kernel counts items in 32 bits
    if (items < grid) grid = (int)items;
    void* args[] = {&sys_dev, &nsys, &max_tiles, &word, &flag};
    DP_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kTsThreads), args, sizeof(TsSmem), s));
    return DP_OK;
}

}  // namespace dp

using namespace dp;

extern "C" {

size_t dp_sptrsv_workspace_bytes(void) { return 256; }

int dp_sptrsv_solve_f64(int32_t n, const int32_t* rowptr, const int32_t* col, const double* val, int32_t upper,
                        const int32_t* plan, int64_t nchunks, int32_t max_level_chunks, const double* b, double* x,
                        int32_t* flag_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0 || nchunks < 0 || !flag_out || !workspace) return DP_ERR_INVALID;
    if (n == 0) return DP_OK;
    if (!rowptr || !col || !val || !plan || !b || !x) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_workspace_bytes()) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* word = static_cast<unsigned long long*>(workspace);
    DP_CUDA(cudaMemsetAsync(word, 0, sizeof(unsigned long long), s));
    fill_u64_kernel<<<sm_count() * 4, 256, 0, s>>>(reinterpret_cast<unsigned long long*>(x), n, kPending);
    DP_LAUNCH_CHECK();
    int grid = coop_grid((const void*)sptrsv_kernel, kBlock, 0);
    int pw = participating_warps(max_level_chunks, grid);
    if (pw < grid) grid = pw;  // one participating warp per CTA: do not launch idle CTAs
    CsrView T{rowptr, col, val, n, 0};
    int upper_i = upper ? 1 : 0;
    long long nch = nchunks;
    int lookback = max_level_chunks > 0 ? max_level_chunks : 1;
    void* args[] = {&T, &upper_i, (void*)&plan, &nch, &pw, &lookback, (void*)&b, &x, &word, &flag_out};
    DP_CUDA(cudaLaunchCooperativeKernel((const void*)sptrsv_kernel, dim3(grid), dim3(kBlock), args, 0, s));
    return DP_OK;
}

size_t dp_sptrsv_batch_workspace_bytes(int32_t nsys) {
    return 256 + sizeof(TrsvSysDev) * (size_t)(nsys > 0 ? nsys : 0);
}

int dp_sptrsv_solve_batch_f64(const dp_trsv_system_t* systems_host, int32_t nsys, int32_t* flag_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!systems_host || nsys <= 0 || !flag_out || !workspace) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_batch_workspace_bytes(nsys)) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<TrsvSysDev> dev((size_t)nsys);
    long long want = 0;
    int nmax = 0;
    for (int i = 0; i < nsys; ++i) {
        const dp_trsv_system_t& u = systems_host[i];
        if (u.n <= 0 || u.nchunks <= 0 || !u.rowptr || !u.col || !u.val || !u.plan || !u.b || !u.x) return DP_ERR_INVALID;
        TrsvSysDev d{};
        d.T = CsrView{u.rowptr, u.col, u.val, u.n, 0};
        d.plan = u.plan, d.b = u.b, d.x = u.x, d.nchunks = u.nchunks, d.upper = u.upper ? 1 : 0;
        d.lookback = u.max_level_chunks > 0 ? u.max_level_chunks : 1;
        dev[(size_t)i] = d;
        const long long w = (long long)trsv_lookahead() * d.lookback;
        if (w > want) want = w;
        if (u.n > nmax) nmax = u.n;
    }
    char* ws = static_cast<char*>(workspace);
    unsigned long long* word = reinterpret_cast<unsigned long long*>(ws);
    TrsvSysDev* sys = reinterpret_cast<TrsvSysDev*>(ws + 256);
    DP_CUDA(cudaMemsetAsync(word, 0, sizeof(unsigned long long), s));
    DP_CUDA(cudaMemcpyAsync(sys, dev.data(), sizeof(TrsvSysDev) * (size_t)nsys, cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaStreamSynchronize(s));  // `dev` is a stack-lifetime staging buffer
    const int fill_x = (nmax + 255) / 256 < sm_count() * 4 ? (nmax + 255) / 256 : sm_count() * 4;
    fill_pending_batch_kernel<<<dim3(fill_x, nsys < 1024 ? nsys : 1024), 256, 0, s>>>(sys, nsys);
    DP_LAUNCH_CHECK();
    int grid = coop_grid((const void*)sptrsv_batch_kernel, kBlock, 0);
    const long long total = (long long)grid * kWarpsPerBlock;
    long long wps = total >= nsys ? total / nsys : 1;  // warps per system
    if (want < 32) want = 32;
    if (wps > want) wps = want;
    const long long need_warps = wps * nsys;
    if (need_warps < total) {  // one participating warp per CTA first: do not launch idle CTAs
        const long long g = need_warps < grid ? need_warps : grid;
        grid = (int)(g > 0 ? g : 1);
    }
    int wps_i = (int)wps, nsys_i = nsys;
    void* args[] = {&sys, &nsys_i, &wps_i, &word, &flag_out};
    DP_CUDA(cudaLaunchCooperativeKernel((const void*)sptrsv_batch_kernel, dim3(grid), dim3(kBlock), args, 0, s));
    return DP_OK;
}

#ifdef DPCG_TS_TRACE
int dp_debug_ts_trace(long long* out_host) {
    DP_CUDA(cudaMemcpyFromSymbol(out_host, g_ts_trace, sizeof(long long) * 2 * kTsTraceTiles * 8));
    return DP_OK;
}
#endif

#ifdef DPCG_LS_TRACE
int dp_debug_ls_trace(long long* out_host) {
    DP_CUDA(cudaMemcpyFromSymbol(out_host, g_ls_trace, sizeof(long long) * 8 * 256));
    return DP_OK;
}
#endif

void dp_sptrsv_ls_limits(int32_t* limits_host) {
    limits_host[0] = kLsCap;                  // one pipeline item per 512-row tile
    limits_host[1] = kLsRowEntries;           // a row's entries live in registers
    limits_host[2] = kLsDepDistance;          // every dependency is still in the shared-memory window
    limits_host[3] = kTileRows;               // rows of the widest level: a window slot is never rewritten by a row of
                                              // the level that still reads it
}

size_t dp_sptrsv_ls_workspace_bytes(int32_t nsys) { return align_up(sizeof(LsSysDev) * (size_t)(nsys > 0 ? nsys : 0), 256); }

int dp_sptrsv_ls_prepare(const dp_trsv_ls_system_t* systems_host, int32_t nsys, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (!systems_host || nsys <= 0 || !workspace) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_ls_workspace_bytes(nsys)) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<LsSysDev> dev((size_t)nsys);
    for (int i = 0; i < nsys; ++i) {
        const dp_trsv_ls_system_t& u = systems_host[i];
        if (u.n <= 0 || !u.rowptr_p || !u.col_p || !u.val_p || !u.b || !u.x) return DP_ERR_INVALID;
        if (!u.perm && u.b == u.x) return DP_ERR_INVALID;
        if (!aligned16(u.col_p) || !aligned16(u.val_p) || !aligned16(u.rowptr_p) || (!u.perm && !aligned16(u.b)))
            return DP_ERR_ALIGNMENT;  // their tiles travel by 16-byte granular bulk copies
        LsSysDev d{};
        d.F = LsFactor{u.rowptr_p, u.col_p, u.val_p, u.perm, u.level_sorted, u.n, u.nnz};
        d.b = u.b, d.x = u.x, d.upper = u.upper ? 1 : 0;
        dev[(size_t)i] = d;
    }
    DP_CUDA(cudaMemcpyAsync(workspace, dev.data(), sizeof(LsSysDev) * (size_t)nsys, cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaStreamSynchronize(s));  // `dev` is a stack-lifetime staging buffer
    return DP_OK;
}

int dp_sptrsv_ls_launch(int32_t nsys, void* workspace, size_t workspace_bytes, void* stream) {
    if (nsys <= 0 || !workspace) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_ls_workspace_bytes(nsys)) return DP_ERR_WORKSPACE;
    if (allow_dynamic_smem((const void*)sptrsv_ls_batch_kernel, sizeof(LsSmem)) != DP_OK) return DP_ERR_CUDA;
    const int resident = sm_count();
    sptrsv_ls_batch_kernel<<<nsys < resident ? nsys : resident, kLsThreads, sizeof(LsSmem), (cudaStream_t)stream>>>(
        static_cast<const LsSysDev*>(workspace), nsys);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

int dp_sptrsv_ls_solve_batch_f64(const dp_trsv_ls_system_t* systems_host, int32_t nsys, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    const int st = dp_sptrsv_ls_prepare(systems_host, nsys, workspace, workspace_bytes, stream);
    return st != DP_OK ? st : dp_sptrsv_ls_launch(nsys, workspace, workspace_bytes, stream);
}

/* ---- tile-stream batch solve ---------------------------------------------------------------------------------- */
static size_t ts_header_bytes(int32_t nsys) { return 256 + align_up(sizeof(TsSysDev) * (size_t)(nsys > 0 ? nsys : 0), 256); }

void dp_sptrsv_ts_limits(int32_t* limits_host) {
    limits_host[0] = kTsCap;       // entries of a 512-row tile that fit one pipeline item
    limits_host[1] = kTsFast + 1;  // entries per row (diagonal included) of the register path: DP_TRSV_SHORT_ROWS
}

size_t dp_sptrsv_ts_workspace_bytes(const dp_trsv_ls_system_t* systems_host, int32_t nsys) {
    size_t bytes = ts_header_bytes(nsys) + align_up(sizeof(TsPermDev) * (size_t)(nsys > 0 ? nsys : 0), 256);
    for (int i = 0; systems_host && i < nsys; ++i)  // b_pos and x_pos of the systems that come in the original numbering
        if (systems_host[i].perm) bytes += 2 * align_up(sizeof(double) * (size_t)(systems_host[i].n > 0 ? systems_host[i].n : 0), 256);
    return bytes;
}

// What dp_sptrsv_ts_launch needs to know about a prepared batch: kept in the first 256 bytes of the workspace, behind
// the abort word.
struct TsPrepared {
    unsigned long long word;  // abort bit of the solve (zeroed by every launch)
    int nsys, max_tiles, nmax, short_rows, nperm, pad;
};
static_assert(sizeof(TsPrepared) <= 256, "workspace header");

int dp_sptrsv_ts_prepare(const dp_trsv_ls_system_t* systems_host, int32_t nsys, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (!systems_host || nsys <= 0 || !workspace) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_ts_workspace_bytes(systems_host, nsys)) return DP_ERR_WORKSPACE;
    if (!aligned16(workspace)) return DP_ERR_ALIGNMENT;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    TsSysDev* sys = reinterpret_cast<TsSysDev*>(ws + 256);
    size_t off = ts_header_bytes(nsys);
    TsPermDev* perm_sys = reinterpret_cast<TsPermDev*>(ws + off);
    off += align_up(sizeof(TsPermDev) * (size_t)nsys, 256);
    std::vector<TsSysDev> dev((size_t)nsys);
    std::vector<TsPermDev> perms;
    TsPrepared head{};
    head.nsys = nsys, head.short_rows = 1;
    for (int i = 0; i < nsys; ++i) {
        const dp_trsv_ls_system_t& u = systems_host[i];
        if (u.n <= 0 || !u.rowptr_p || !u.col_p || !u.val_p || !u.b || !u.x || u.b == u.x) return DP_ERR_INVALID;
        if (!(u.flags & DP_TRSV_SHORT_ROWS)) head.short_rows = 0;
        if (!aligned16(u.col_p) || !aligned16(u.val_p) || !aligned16(u.rowptr_p) || (!u.perm && !aligned16(u.b)))
            return DP_ERR_ALIGNMENT;  // spans of all four arrays are moved by 16-byte granular bulk copies
        TsSysDev d{};
        d.F = LsFactor{u.rowptr_p, u.col_p, u.val_p, nullptr, nullptr, u.n, u.nnz};
        d.b = u.b, d.x = u.x, d.upper = u.upper ? 1 : 0;
        d.rev = (u.flags & DP_TRSV_REVERSED) ? 1 : 0;
        if (d.rev && u.perm) return DP_ERR_INVALID;  // reversed positions are a form of position space
        if (u.perm) {  // original numbering: the solve runs on position-space copies in the workspace
            const size_t vec = align_up(sizeof(double) * (size_t)u.n, 256);
            TsPermDev g{u.perm, u.b, u.x, reinterpret_cast<double*>(ws + off), reinterpret_cast<double*>(ws + off + vec), u.n, 0};
            off += 2 * vec;
            d.b = g.b_pos, d.x = g.x_pos;
            perms.push_back(g);
        }
        d.ntiles = (u.n + kTileRows - 1) / kTileRows;
        if (d.ntiles > head.max_tiles) head.max_tiles = d.ntiles;
        if (u.n > head.nmax) head.nmax = u.n;
        dev[(size_t)i] = d;
    }
    head.nperm = (int)perms.size();
    if ((long long)head.max_tiles * nsys >= (1ll << 31) - 1024) return DP_ERR_INVALID;  // the kernel counts items in 32 bits
    DP_CUDA(cudaMemcpyAsync(ws, &head, sizeof(head), cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaMemcpyAsync(sys, dev.data(), sizeof(TsSysDev) * (size_t)nsys, cudaMemcpyHostToDevice, s));
    if (head.nperm)
        DP_CUDA(cudaMemcpyAsync(perm_sys, perms.data(), sizeof(TsPermDev) * perms.size(), cudaMemcpyHostToDevice, s));
    DP_CUDA(cudaStreamSynchronize(s));  // staging buffers are about to go out of scope
    return DP_OK;
}

// `prepared_host`: the five integers dp_sptrsv_ts_prepare derived (the caller keeps them; nothing is read back).
int dp_sptrsv_ts_launch(int32_t nsys, int32_t max_tiles, int32_t nmax, int32_t short_rows, int32_t nperm, int32_t* flag_out,
                        void* workspace, void* stream) {
    if (nsys <= 0 || max_tiles <= 0 || !flag_out || !workspace) return DP_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    unsigned long long* word = reinterpret_cast<unsigned long long*>(ws);
    TsSysDev* sys = reinterpret_cast<TsSysDev*>(ws + 256);
    TsPermDev* perm_sys = reinterpret_cast<TsPermDev*>(ws + ts_header_bytes(nsys));
    DP_CUDA(cudaMemsetAsync(word, 0, sizeof(unsigned long long), s));
    const int pass_grid = (nmax + 255) / 256 < sm_count() * 8 ? (nmax + 255) / 256 : sm_count() * 8;
    if (nperm) {
        ts_gather_kernel<<<pass_grid, 256, 0, s>>>(perm_sys, nperm);
        DP_LAUNCH_CHECK();
    }
    const int st = ts_solve_launch(sys, nsys, max_tiles, nmax, short_rows != 0, true, word, flag_out, s);
    if (st != DP_OK) return st;
    if (nperm) {
        ts_scatter_kernel<<<pass_grid, 256, 0, s>>>(perm_sys, nperm);
        DP_LAUNCH_CHECK();
    }
    return DP_OK;
}

int dp_sptrsv_ts_solve_batch_f64(const dp_trsv_ls_system_t* systems_host, int32_t nsys, int32_t* flag_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    if (!flag_out) return DP_ERR_INVALID;
    int st = dp_sptrsv_ts_prepare(systems_host, nsys, workspace, workspace_bytes, stream);
    if (st != DP_OK) return st;
    int max_tiles = 0, nmax = 0, short_rows = 1, nperm = 0;
    for (int i = 0; i < nsys; ++i) {
        const dp_trsv_ls_system_t& u = systems_host[i];
        const int nt = (u.n + kTileRows - 1) / kTileRows;
        if (nt > max_tiles) max_tiles = nt;
        if (u.n > nmax) nmax = u.n;
        if (!(u.flags & DP_TRSV_SHORT_ROWS)) short_rows = 0;
        if (u.perm) ++nperm;
    }
    return dp_sptrsv_ts_launch(nsys, max_tiles, nmax, short_rows, nperm, flag_out, workspace, stream);
}

int dp_ic0_f64(int32_t n, const int32_t* rowptr, const int32_t* col, const double* a_val, double* l_val,
               const int32_t* plan, int64_t nchunks, int32_t max_level_chunks, int32_t* flag_out, void* workspace,
               size_t workspace_bytes, void* stream) {
    if (n < 0 || nchunks < 0 || !flag_out || !workspace) return DP_ERR_INVALID;
    if (n == 0) return DP_OK;
    if (!rowptr || !col || !a_val || !l_val || !plan) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_workspace_bytes()) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* word = static_cast<unsigned long long*>(workspace);
    DP_CUDA(cudaMemsetAsync(word, 0, sizeof(unsigned long long), s));
    // nnz = rowptr[n] lives on the device: arm every entry of L (caller sized l_val like a_val)
    int nnz = 0;
    DP_CUDA(cudaMemcpyAsync(&nnz, rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    DP_CUDA(cudaStreamSynchronize(s));
    fill_u64_kernel<<<sm_count() * 4, 256, 0, s>>>(reinterpret_cast<unsigned long long*>(l_val), nnz, kPending);
    DP_LAUNCH_CHECK();
    int grid = coop_grid((const void*)ic0_kernel, kBlock, 0);
    int pw = participating_warps(max_level_chunks, grid);
    if (pw < grid) grid = pw;
    CsrView A{rowptr, col, a_val, n, nnz};
    long long nch = nchunks;
    void* args[] = {&A, &l_val, (void*)&plan, &nch, &pw, &word, &flag_out};
    DP_CUDA(cudaLaunchCooperativeKernel((const void*)ic0_kernel, dim3(grid), dim3(kBlock), args, 0, s));
    return DP_OK;
}

}  // extern "C"
