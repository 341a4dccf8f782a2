// levels.cu — K3: level sets of a triangular CSR pattern, level-sorted permutation, warp-padded execution plan.
// No reference counterpart; integer output is bit-exact against oracle_levels/oracle_level_perm.
//
//   level[i] = 1 + max(level[j] : j a dependency of row i), 0 without dependencies
//
// computed in ONE sync-free cooperative launch: a warp owns 32 consecutive positions (rows in solve order),
// dependencies outside the warp are awaited by spinning on level[] itself (-1 = not yet known), dependencies
// inside the warp are resolved by a 32-step shuffle sweep. perm is a stable LSD radix sort of the rows by level.
#include "common.cuh"
#include "scan.cuh"

namespace dp {

constexpr int kLevelThreads = 128;  // 4 warps per CTA, 1 CTA per SM: few pollers, the frontier is narrow anyway

__global__ void __launch_bounds__(kLevelThreads)
levels_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col, int upper, int* level,
              int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (kLevelThreads / 32);
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;  // consecutive chunks land on different SMs
    const int nchunks = (n + 31) / 32;
    for (int chunk = gw; chunk < nchunks; chunk += nwarps) {
        const int base = chunk * 32;          // in position space: pos = upper ? n-1-row : row
        const int pos = base + lane;
        const bool valid = pos < n;
        const int row = upper ? n - 1 - pos : pos;
        int e = 0, re = 0;
        if (valid) {
            e = __ldg(rowptr + row);
            re = __ldg(rowptr + row + 1);
            // contract: diagonal last (lower) / first (upper)
            const int dpos = upper ? e : re - 1;
            if (re <= e || __ldg(col + dpos) != row) atomicCAS(flag, 0, (int)DP_ERR_STRUCTURE);
        }
        int ext_max = -1;
        unsigned imask = 0;
        unsigned idle = 0;
        for (;;) {
            const bool pending = valid && e < re;
            if (!__any_sync(kFull, pending)) break;
            bool progress = false;
            if (pending) {
                const int j = __ldg(col + e);
                const int pj = upper ? n - 1 - j : j;
                if (j == row) {
                    ++e, progress = true;
                } else if ((unsigned)j >= (unsigned)n || pj > pos) {  // wrong triangle / out of range
                    atomicCAS(flag, 0, (int)DP_ERR_STRUCTURE);
                    ++e, progress = true;
                } else if (pj >= base) {  // produced by a lower lane of this warp
                    imask |= 1u << (pj - base);
                    ++e, progress = true;
                } else {
                    const int lv = ld_relaxed_s32(level + j);
                    if (lv >= 0) {
                        ext_max = max(ext_max, lv);
                        ++e, progress = true;
                    }
                }
            }
            if (!__any_sync(kFull, progress)) {
                if (++idle > kSpinBudget) {
                    atomicCAS(flag, 0, (int)DP_ERR_TIMEOUT);
                    return;
                }
                if (ld_relaxed_s32(flag) == (int)DP_ERR_TIMEOUT) return;
                __nanosleep(100);
            }
        }
        // in-warp sweep: lane k is final once lanes < k are
        const unsigned need = __reduce_or_sync(kFull, imask);
        int int_max = -1, lvl = 0;
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            if (lane == k) lvl = max(ext_max, int_max) + 1;
            if ((need >> k) & 1u) {
                const int v = __shfl_sync(kFull, lvl, k);
                if ((imask >> k) & 1u) int_max = max(int_max, v);
            }
        }
        if (valid) st_relaxed_s32(level + row, lvl);
    }
}

__global__ void level_hist_kernel(int n, const int* __restrict__ level, int* __restrict__ count, int* __restrict__ nlevels) {
    int local_max = -1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int l = level[i];
        atomicAdd(count + l, 1);
        local_max = max(local_max, l);
    }
    local_max = __reduce_max_sync(kFull, local_max);
    if ((threadIdx.x & 31) == 0 && local_max >= 0) atomicMax(nlevels, local_max + 1);
}

// ---- stable LSD radix sort of (key = level, value = row), 8 bits per pass -----------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortPerWarp = 256;
constexpr int kSortSpan = kSortWarps * kSortPerWarp;  // 2048 items per CTA

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const int* __restrict__ keys, int n, int shift, int nblocks, int* __restrict__ table) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortSpan;
    for (int k = threadIdx.x; k < kSortSpan; k += kSortThreads) {
        const int i = base + k;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255], 1);
    }
    __syncthreads();
    table[threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];  // digit-major: smaller digits first, then CTAs
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ vals_in /* null: identity */, int n,
                     int shift, int nblocks, const int* __restrict__ table, int* __restrict__ keys_out,
                     int* __restrict__ vals_out) {
    __shared__ int wh[kSortWarps][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = lane; d < 256; d += 32) wh[warp][d] = 0;
    __syncwarp();
    const int start = blockIdx.x * kSortSpan + warp * kSortPerWarp;
    const unsigned lt = (1u << lane) - 1u;
    for (int g = 0; g < kSortPerWarp / 32; ++g) {
        const int i = start + g * 32 + lane;
        const int d = i < n ? ((keys_in[i] >> shift) & 255) : 256 + lane;
        const unsigned peers = __match_any_sync(kFull, d);
        if (i < n && (peers & lt) == 0) wh[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: running start offset of digit d for every warp of this CTA
        const int d = threadIdx.x;
        int run = table[d * nblocks + blockIdx.x];
        for (int w = 0; w < kSortWarps; ++w) {
            const int t = wh[w][d];
            wh[w][d] = run;
            run += t;
        }
    }
    __syncthreads();
    for (int g = 0; g < kSortPerWarp / 32; ++g) {
        const int i = start + g * 32 + lane;
        const int key = i < n ? keys_in[i] : 0;
        const int d = i < n ? ((key >> shift) & 255) : 256 + lane;
        const unsigned peers = __match_any_sync(kFull, d);
        if (i < n) {
            const int pos = wh[warp][d] + __popc(peers & lt);
            keys_out[pos] = key;
            vals_out[pos] = vals_in ? vals_in[i] : i;
        }
        __syncwarp();
        if (i < n && (peers & lt) == 0) wh[warp][d] += __popc(peers);
        __syncwarp();
    }
}

__global__ void plan_build_kernel(int nlevels, const int* __restrict__ perm, const int* __restrict__ level_ptr,
                                  const int* __restrict__ chunk_ptr, int* __restrict__ plan, long long nchunks) {
    const int lane = threadIdx.x & 31;
    for (long long c = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; c < nchunks;
         c += ((long long)gridDim.x * blockDim.x) >> 5) {
        int lo = 0, hi = nlevels;  // last level l with chunk_ptr[l] <= c
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunk_ptr[mid] <= c) lo = mid; else hi = mid;
        }
        const int idx = level_ptr[lo] + (int)(c - chunk_ptr[lo]) * 32 + lane;
        plan[c * 32 + lane] = idx < level_ptr[lo + 1] ? perm[idx] : -1;
    }
}

// ---- level-ordered copy of a triangular factor (input of the level-stream solve, trsv_ls.cuh) -------------------
__global__ void inv_perm_kernel(int n, const int* __restrict__ perm, int* __restrict__ inv) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) inv[perm[r]] = r;
}

__global__ void perm_rowlen_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ perm,
                                   const int* __restrict__ level, int* __restrict__ len, int* __restrict__ level_sorted) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += gridDim.x * blockDim.x) {
        if (r < n) {
            const int row = perm[r];
            len[r] = rowptr[row + 1] - rowptr[row];
            level_sorted[r] = level[row];
        } else {
            len[r] = 0;
        }
    }
}

// One warp per row of the copy: entries keep their order inside the row, columns become positions, and the diagonal
// entry (last in a lower row, first in an upper one) is replaced by its reciprocal, rounded exactly like the solve
// kernels round it (__ddiv_rn(1, d)): the level-stream solve multiplies by it.
__global__ void perm_fill_kernel(int n, int upper, const int* __restrict__ rowptr, const int* __restrict__ col,
                                 const double* __restrict__ val, const int* __restrict__ perm, const int* __restrict__ inv,
                                 const int* __restrict__ rowptr_p, int* __restrict__ col_p, double* __restrict__ val_p) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const int row = perm[r];
        const int src = rowptr[row], cnt = rowptr[row + 1] - src, dst = rowptr_p[r];
        const int dk = upper ? 0 : cnt - 1;
        for (int k = lane; k < cnt; k += 32) {
            col_p[dst + k] = inv[col[src + k]];
            val_p[dst + k] = k == dk ? __ddiv_rn(1.0, val[src + k]) : val[src + k];
        }
    }
}

// What the level-stream solve needs to know about the copy: stats[0] = most entries in a 512-row tile, stats[1] = most
// entries in a row, stats[2] = largest distance (in positions) between a row and one of its dependencies.
__global__ void perm_stats_kernel(int n, const int* __restrict__ rowptr_p, const int* __restrict__ col_p, int* __restrict__ stats) {
    int tile_max = 0, row_max = 0, dist_max = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int rs = rowptr_p[r], re = rowptr_p[r + 1];
        row_max = max(row_max, re - rs);
        for (int q = rs; q < re; ++q) dist_max = max(dist_max, r - col_p[q]);
        if (r % kTileRows == 0) tile_max = max(tile_max, rowptr_p[min(n, r + kTileRows)] - rs);
    }
    tile_max = __reduce_max_sync(kFull, tile_max);
    row_max = __reduce_max_sync(kFull, row_max);
    dist_max = __reduce_max_sync(kFull, dist_max);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(stats + 0, tile_max);
        atomicMax(stats + 1, row_max);
        atomicMax(stats + 2, dist_max);
    }
}

// chunks[l] = ceil(size(l) / 32) for l < *nlevels, 0 beyond (the scan below runs over n + 1 entries: nlevels is only
// known on the device); summary[2] = widest level in chunks.
__global__ void level_chunks_kernel(int n, const int* __restrict__ nlevels, const int* __restrict__ level_ptr,
                                    int* __restrict__ chunks, int* __restrict__ summary) {
    const int nl = *nlevels;
    int widest = 0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= n; l += gridDim.x * blockDim.x) {
        const int c = l < nl ? (level_ptr[l + 1] - level_ptr[l] + 31) / 32 : 0;
        chunks[l] = c;
        widest = max(widest, c);
    }
    widest = __reduce_max_sync(kFull, widest);
    if ((threadIdx.x & 31) == 0 && widest > 0) atomicMax(summary + 2, widest);
}
__global__ void plan_summary_kernel(int n, const int* __restrict__ nlevels, const int* __restrict__ chunk_ptr,
                                    int* __restrict__ summary) {
    summary[0] = *nlevels;
    summary[1] = chunk_ptr[min(*nlevels, n)];  // exclusive scan: entry nlevels is the total
}

static int grid_for(long long items, int threads) {
    long long b = (items + threads - 1) / threads;
    long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

static int sort_blocks(int n) { return (n + kSortSpan - 1) / kSortSpan; }

struct AnalyseWs {
    int* keys[2];
    int* vals;
    int* table;
    void* scan_ws;
    size_t bytes;
};

static AnalyseWs carve_analyse_ws(void* ws, int n) {
    AnalyseWs w{};
    const size_t nb = (size_t)sort_blocks(n > 0 ? n : 1);
    const size_t table_items = 256 * nb;
    const size_t scan_items = table_items > (size_t)n + 1 ? table_items : (size_t)n + 1;
    size_t off = 0;
    char* p = static_cast<char*>(ws);
    auto take = [&](size_t bytes) { void* q = p ? p + off : nullptr; off += align_up(bytes, 256); return q; };
    w.keys[0] = static_cast<int*>(take(sizeof(int) * (size_t)n));
    w.keys[1] = static_cast<int*>(take(sizeof(int) * (size_t)n));
    w.vals = static_cast<int*>(take(sizeof(int) * (size_t)n));
    w.table = static_cast<int*>(take(sizeof(int) * table_items));
    w.scan_ws = take(scan_workspace_bytes((long long)scan_items));
    w.bytes = off;
    return w;
}

}  // namespace dp

using namespace dp;

extern "C" {

size_t dp_sptrsv_analyse_workspace_bytes(int32_t n) { return carve_analyse_ws(nullptr, n < 0 ? 0 : n).bytes; }

int dp_sptrsv_analyse(int32_t n, const int32_t* rowptr, const int32_t* col, int32_t upper, int32_t* level,
                      int32_t* perm, int32_t* level_ptr, int32_t* nlevels_out, int32_t* flag_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (n < 0 || !rowptr || !level || !perm || !level_ptr || !nlevels_out || !flag_out || !workspace)
        return DP_ERR_INVALID;
    if (!aligned16(workspace)) return DP_ERR_ALIGNMENT;
    AnalyseWs w = carve_analyse_ws(workspace, n);
    if (workspace_bytes < w.bytes) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    DP_CUDA(cudaMemsetAsync(nlevels_out, 0, sizeof(int), s));
    DP_CUDA(cudaMemsetAsync(level_ptr, 0, sizeof(int) * ((size_t)n + 1), s));
    if (n == 0) return DP_OK;
    if (!col) return DP_ERR_INVALID;

    // 1. levels (sync-free, cooperative: every warp must be resident while it spins)
    DP_CUDA(cudaMemsetAsync(level, 0xFF, sizeof(int) * (size_t)n, s));
    {
        int grid = sm_count();
        const int need = ((n + 31) / 32 + kLevelThreads / 32 - 1) / (kLevelThreads / 32);
        if (need < grid) grid = need;
        int upper_i = upper ? 1 : 0;
        void* args[] = {&n, (void*)&rowptr, (void*)&col, &upper_i, &level, &flag_out};
        DP_CUDA(cudaLaunchCooperativeKernel((const void*)levels_kernel, dim3(grid), dim3(kLevelThreads), args, 0, s));
    }
    // 2. level_ptr = exclusive scan of the level histogram; nlevels = max + 1
    level_hist_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, level, level_ptr, nlevels_out);
    DP_LAUNCH_CHECK();
    int st = exclusive_scan_i32(level_ptr, level_ptr, (long long)n + 1, w.scan_ws, s);
    if (st != DP_OK) return st;
    // 3. perm = stable sort of rows by level: LSD radix, 8 bits per pass, enough passes for any level < n
    int bits = 1;
    while ((1ll << bits) < (long long)n) ++bits;
    const int passes = (bits + 7) / 8;
    const int nb = sort_blocks(n);
    const int* kin = level;
    const int* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
        int* kout = w.keys[p & 1];
        int* vout = ((passes - 1 - p) & 1) == 0 ? perm : w.vals;
        radix_hist_kernel<<<nb, kSortThreads, 0, s>>>(kin, n, 8 * p, nb, w.table);
        DP_LAUNCH_CHECK();
        st = exclusive_scan_i32(w.table, w.table, 256ll * nb, w.scan_ws, s);
        if (st != DP_OK) return st;
        radix_scatter_kernel<<<nb, kSortThreads, 0, s>>>(kin, vin, n, 8 * p, nb, w.table, kout, vout);
        DP_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
    }
    return DP_OK;
}

size_t dp_sptrsv_permute_workspace_bytes(int32_t n) {
    const size_t m = (size_t)(n < 0 ? 0 : n);
    return align_up(sizeof(int) * m, 256) + align_up(scan_workspace_bytes((long long)m + 1), 256);
}

int dp_sptrsv_permute(int32_t n, int32_t upper, const int32_t* rowptr, const int32_t* col, const double* val, const int32_t* perm,
                      const int32_t* level, int32_t* rowptr_p, int32_t* col_p, double* val_p, int32_t* level_sorted,
                      int32_t* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0 || !rowptr_p || !stats_out || !workspace) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_permute_workspace_bytes(n)) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    DP_CUDA(cudaMemsetAsync(stats_out, 0, 3 * sizeof(int), s));
    DP_CUDA(cudaMemsetAsync(rowptr_p, 0, sizeof(int), s));
    if (n == 0) return DP_OK;
    if (!rowptr || !col || !val || !perm || !level || !col_p || !val_p || !level_sorted) return DP_ERR_INVALID;
    if (!aligned16(col_p) || !aligned16(val_p) || !aligned16(workspace)) return DP_ERR_ALIGNMENT;
    int* inv = static_cast<int*>(workspace);
    void* scan_ws = static_cast<char*>(workspace) + align_up(sizeof(int) * (size_t)n, 256);
    inv_perm_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, perm, inv);
    DP_LAUNCH_CHECK();
    perm_rowlen_kernel<<<grid_for((long long)n + 1, 256), 256, 0, s>>>(n, rowptr, perm, level, rowptr_p, level_sorted);
    DP_LAUNCH_CHECK();
    const int st = exclusive_scan_i32(rowptr_p, rowptr_p, (long long)n + 1, scan_ws, s);
    if (st != DP_OK) return st;
    perm_fill_kernel<<<grid_for(32ll * n, 256), 256, 0, s>>>(n, upper ? 1 : 0, rowptr, col, val, perm, inv, rowptr_p, col_p, val_p);
    DP_LAUNCH_CHECK();
    perm_stats_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, rowptr_p, col_p, stats_out);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

size_t dp_sptrsv_plan_sizes_workspace_bytes(int32_t n) {
    const long long m = (long long)(n < 0 ? 0 : n) + 1;
    return align_up(sizeof(int) * (size_t)m, 256) + scan_workspace_bytes(m);
}

int dp_sptrsv_plan_sizes(int32_t n, const int32_t* nlevels, const int32_t* level_ptr, int32_t* chunk_ptr,
                         int32_t* summary_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0 || !nlevels || !level_ptr || !chunk_ptr || !summary_out || !workspace) return DP_ERR_INVALID;
    if (workspace_bytes < dp_sptrsv_plan_sizes_workspace_bytes(n)) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int* chunks = static_cast<int*>(workspace);
    void* scan_ws = static_cast<char*>(workspace) + align_up(sizeof(int) * ((size_t)n + 1), 256);
    DP_CUDA(cudaMemsetAsync(summary_out, 0, 3 * sizeof(int), s));
    level_chunks_kernel<<<grid_for((long long)n + 1, 256), 256, 0, s>>>(n, nlevels, level_ptr, chunks, summary_out);
    DP_LAUNCH_CHECK();
    const int st = exclusive_scan_i32(chunks, chunk_ptr, (long long)n + 1, scan_ws, s);
    if (st != DP_OK) return st;
    plan_summary_kernel<<<1, 1, 0, s>>>(n, nlevels, chunk_ptr, summary_out);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

int64_t dp_sptrsv_plan_chunks(int32_t nlevels, const int32_t* level_ptr_host) {
    if (nlevels < 0 || (nlevels > 0 && !level_ptr_host)) return -1;
    int64_t c = 0;
    for (int l = 0; l < nlevels; ++l) c += (level_ptr_host[l + 1] - level_ptr_host[l] + 31) / 32;
    return c;
}

int dp_sptrsv_plan_build(int32_t n, int32_t nlevels, const int32_t* perm, const int32_t* level_ptr,
                         const int32_t* chunk_ptr, int32_t* plan, int64_t nchunks, void* stream) {
    if (n < 0 || nlevels < 0 || nchunks < 0) return DP_ERR_INVALID;
    if (nchunks == 0) return DP_OK;
    if (!perm || !level_ptr || !chunk_ptr || !plan) return DP_ERR_INVALID;
    plan_build_kernel<<<grid_for(nchunks * 32, 256), 256, 0, (cudaStream_t)stream>>>(nlevels, perm, level_ptr,
                                                                                    chunk_ptr, plan, nchunks);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

}  // extern "C"
