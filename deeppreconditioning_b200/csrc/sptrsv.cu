// sptrsv.cu — standalone K4 entry point (dp_sptrsv_solve_f64) and the sync-free IC(0) factorisation (dp_ic0_f64).
#include "sptrsv.cuh"

namespace dp {

__global__ void fill_u64_kernel(unsigned long long* __restrict__ p, long long count, unsigned long long v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// Participating warps: `pw` of them, spread one per CTA first (gw = warp * gridDim + block), chunk c -> warp c % pw.
__global__ void __launch_bounds__(kBlock, 2)
sptrsv_kernel(CsrView T, int upper, const int* __restrict__ plan, long long nchunks, int pw,
              const double* __restrict__ b, double* x, unsigned long long* word, int* flag) {
    const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    if (gw >= pw) return;
    const AbortCtl ctl{word, flag};
    const RhsPlain rhs{b};
    const bool light = pw <= 128;
    for (long long c = gw; c < nchunks; c += pw) {
        const bool ok = upper ? sptrsv_chunk<true>(T, plan + c * 32, rhs, x, ctl, light)
                              : sptrsv_chunk<false>(T, plan + c * 32, rhs, x, ctl, light);
        if (!ok) return;
    }
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

int coop_grid(const void* kernel, int threads, size_t smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * sm_count();
}

static int participating_warps(int max_level_chunks, int grid) {
    const int w = grid * kWarpsPerBlock;
    if (max_level_chunks <= 0) return w;
    long long pw = 4ll * max_level_chunks;
    if (pw < 32) pw = 32;
    return pw > w ? w : (int)pw;
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
    void* args[] = {&T, &upper_i, (void*)&plan, &nch, &pw, (void*)&b, &x, &word, &flag_out};
    DP_CUDA(cudaLaunchCooperativeKernel((const void*)sptrsv_kernel, dim3(grid), dim3(kBlock), args, 0, s));
    return DP_OK;
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
