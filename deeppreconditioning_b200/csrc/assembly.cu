// assembly.cu — K1: CSR assembly of L / L^T / A straight from the spconv COO layout, explicit CSR transpose,
// inverse diagonal. Replaces the dense N x N round trips of test.py:61-68 and test.py:100-105.
//
// Pipeline (all integer work exact, values only widened fp32 -> fp64):
//   count  : one thread per COO entry, filter (batch, row < n, col < n, row >= col, value != 0), atomicAdd per row
//   scan   : exclusive prefix sum -> rowptr
//   fill   : entries dropped into their row segment at an atomically claimed slot (order inside a row arbitrary)
//   sort   : every row sorted by column -> the output no longer depends on the atomic order: bit-exact
//            against torch's to_sparse_csr() / scipy; a repeated (row, col) raises DP_ERR_STRUCTURE.
#include "common.cuh"
#include "scan.cuh"

namespace dp {

struct CooEntry {
    int r, c;
    float v;
    bool keep;
};

__device__ __forceinline__ CooEntry coo_load(const int* __restrict__ ind, const float* __restrict__ feat, long long e,
                                             int batch, int n) {
    CooEntry t;
    const int b = ind[3 * e];
    t.r = ind[3 * e + 1];
    t.c = ind[3 * e + 2];
    t.v = feat[e];
    // `[batch, 0, :n, :n]` slice, lower triangle (model.py:53-54 zeroes the rest by value), exact zeros dropped
    // like Tensor.to_sparse_csr() (test.py:105). NaN != 0 is kept, -0.0 is dropped, as torch does.
    t.keep = b == batch && (unsigned)t.r < (unsigned)n && (unsigned)t.c < (unsigned)n && t.r >= t.c && t.v != 0.0f;
    return t;
}

__global__ void coo_count_kernel(const int* __restrict__ ind, const float* __restrict__ feat, long long nnz, int batch,
                                 int n, int mode, int* __restrict__ count) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nnz;
         e += (long long)gridDim.x * blockDim.x) {
        const CooEntry t = coo_load(ind, feat, e, batch, n);
        if (!t.keep) continue;
        if (mode == DP_ASSEMBLE_TRIL) {
            atomicAdd(count + t.r, 1);
        } else if (mode == DP_ASSEMBLE_TRIL_T) {
            atomicAdd(count + t.c, 1);
        } else {
            atomicAdd(count + t.r, 1);
            if (t.r > t.c) atomicAdd(count + t.c, 1);
        }
    }
}

__global__ void coo_fill_kernel(const int* __restrict__ ind, const float* __restrict__ feat, long long nnz, int batch,
                                int n, int mode, const int* __restrict__ rowptr, int* __restrict__ cursor,
                                int* __restrict__ col, double* __restrict__ val) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nnz;
         e += (long long)gridDim.x * blockDim.x) {
        const CooEntry t = coo_load(ind, feat, e, batch, n);
        if (!t.keep) continue;
        const double v = (double)t.v;  // exact widening (test.py:68,105: .to(torch.float64))
        if (mode != DP_ASSEMBLE_TRIL_T) {
            const int pos = rowptr[t.r] + atomicAdd(cursor + t.r, 1);
            col[pos] = t.c;
            val[pos] = v;
        }
        if (mode == DP_ASSEMBLE_TRIL_T || (mode == DP_ASSEMBLE_SYMMETRISE && t.r > t.c)) {
            const int pos = rowptr[t.c] + atomicAdd(cursor + t.c, 1);
            col[pos] = t.r;
            val[pos] = v;
        }
    }
}

// One thread per row: insertion sort by column (rows on this path hold 3..~25 entries), then duplicate check.
__global__ void csr_sort_rows_kernel(int n, const int* __restrict__ rowptr, int* __restrict__ col,
                                     double* __restrict__ val, int* __restrict__ flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int rs = rowptr[i], re = rowptr[i + 1];
        for (int p = rs + 1; p < re; ++p) {
            const int kc = col[p];
            const double kv = val[p];
            int q = p - 1;
            while (q >= rs && col[q] > kc) {
                col[q + 1] = col[q];
                val[q + 1] = val[q];
                --q;
            }
            col[q + 1] = kc;
            val[q + 1] = kv;
        }
        for (int p = rs + 1; p < re; ++p)
            if (col[p] == col[p - 1]) atomicCAS(flag, 0, (int)DP_ERR_STRUCTURE);
    }
}

__global__ void csr_count_cols_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                      int* __restrict__ count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) atomicAdd(count + col[p], 1);
}

__global__ void csr_transpose_fill_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                          const double* __restrict__ val, const int* __restrict__ rowptr_t,
                                          int* __restrict__ cursor, int* __restrict__ col_t,
                                          double* __restrict__ val_t) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const int j = col[p];
            const int pos = rowptr_t[j] + atomicAdd(cursor + j, 1);
            col_t[pos] = i;
            val_t[pos] = val[p];
        }
}

__global__ void copy_last_kernel(const int* __restrict__ rowptr, int n, int* __restrict__ nnz_out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *nnz_out = rowptr[n];
}

__global__ void inv_diagonal_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                    const double* __restrict__ val, double* __restrict__ dinv, int* __restrict__ flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double d = 0.0;
        bool found = false;
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p)
            if (col[p] == i) {
                d = val[p];
                found = true;
            }
        if (!found) atomicCAS(flag, 0, (int)DP_ERR_STRUCTURE);
        dinv[i] = found ? __ddiv_rn(1.0, d) : 0.0;  // `1 / matrix.diagonal()`, test.py:76
    }
}


// nnz(L L^T) without forming the product (test.py:107-109 reports the density of the explicit M = L L^T that
// test.py:104-105 stores): entry (i, j) is structurally non-zero iff rows i and j of L share a column, i.e.
//   row i of L L^T  =  union over the columns c of row i of { rows j with L_jc != 0 }  =  union of rows c of L^T.
// One warp per row: candidate j from the k-th list counts iff it is in none of the lists 0..k-1 (binary search, every
// row of L^T is sorted). Integer work, exact; a 64-bit total by atomics.
__global__ void __launch_bounds__(256) aat_nnz_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                      const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                                      unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long mine = 0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int rs = rowptr[i], re = rowptr[i + 1];
        for (int k = rs; k < re; ++k) {
            const int c = col[k];
            const int ts = rowptr_t[c], te = rowptr_t[c + 1];
            for (int q = ts + lane; q < te; q += 32) {
                const int j = col_t[q];
                bool seen = false;
                for (int k2 = rs; k2 < k && !seen; ++k2) {
                    const int c2 = col[k2];
                    int lo = rowptr_t[c2], hi = rowptr_t[c2 + 1];
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        const int v = col_t[mid];
                        if (v == j) { seen = true; break; }
                        if (v < j) lo = mid + 1; else hi = mid;
                    }
                }
                mine += seen ? 0ull : 1ull;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFull, mine, o);
    if (lane == 0 && mine) atomicAdd(total, mine);
}

static int grid_for(long long items, int threads) {
    long long b = (items + threads - 1) / threads;
    long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

// workspace layout: [cursor/count: (n+1) int][scan scratch]
static size_t assembly_ws_bytes(int n) { return align_up(sizeof(int) * ((size_t)n + 1), 256) + scan_workspace_bytes((long long)n + 1); }

}  // namespace dp

using namespace dp;

extern "C" {

size_t dp_csr_from_coo_workspace_bytes(int32_t n, int64_t nnz_in) {
    (void)nnz_in;
    return assembly_ws_bytes(n < 0 ? 0 : n);
}

int dp_csr_from_coo(const int32_t* indices, const float* features, int64_t nnz_in, int32_t batch, int32_t n,
                    int32_t mode, int32_t* rowptr, int32_t* col, double* val, int32_t* nnz_out, int32_t* flag_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0 || nnz_in < 0 || mode < 0 || mode > 2 || !rowptr || !flag_out || !workspace) return DP_ERR_INVALID;
    if (nnz_in > 0 && (!indices || !features || !col || !val)) return DP_ERR_INVALID;
    if (!aligned16(col) || !aligned16(val) || !aligned16(workspace)) return DP_ERR_ALIGNMENT;
    if (workspace_bytes < assembly_ws_bytes(n)) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int* cursor = static_cast<int*>(workspace);
    void* scan_ws = static_cast<char*>(workspace) + align_up(sizeof(int) * ((size_t)n + 1), 256);

    DP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)n + 1), s));
    if (nnz_in > 0) {
        coo_count_kernel<<<grid_for(nnz_in, 256), 256, 0, s>>>(indices, features, nnz_in, batch, n, mode, cursor);
        DP_LAUNCH_CHECK();
    }
    int st = exclusive_scan_i32(cursor, rowptr, (long long)n + 1, scan_ws, s);
    if (st != DP_OK) return st;
    DP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)n + 1), s));
    if (nnz_in > 0) {
        coo_fill_kernel<<<grid_for(nnz_in, 256), 256, 0, s>>>(indices, features, nnz_in, batch, n, mode, rowptr,
                                                              cursor, col, val);
        DP_LAUNCH_CHECK();
        csr_sort_rows_kernel<<<grid_for(n, 128), 128, 0, s>>>(n, rowptr, col, val, flag_out);
        DP_LAUNCH_CHECK();
    }
    if (nnz_out) {
        copy_last_kernel<<<1, 32, 0, s>>>(rowptr, n, nnz_out);
        DP_LAUNCH_CHECK();
    }
    return DP_OK;
}

size_t dp_csr_transpose_workspace_bytes(int32_t n, int32_t nnz) {
    (void)nnz;
    return assembly_ws_bytes(n < 0 ? 0 : n);
}

int dp_csr_transpose(int32_t n, int32_t nnz, const int32_t* rowptr, const int32_t* col, const double* val,
                     int32_t* rowptr_t, int32_t* col_t, double* val_t, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (n < 0 || nnz < 0 || !rowptr || !rowptr_t || !workspace) return DP_ERR_INVALID;
    if (nnz > 0 && (!col || !val || !col_t || !val_t)) return DP_ERR_INVALID;
    if (!aligned16(col_t) || !aligned16(val_t) || !aligned16(workspace)) return DP_ERR_ALIGNMENT;
    if (workspace_bytes < assembly_ws_bytes(n)) return DP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int* cursor = static_cast<int*>(workspace);
    void* scan_ws = static_cast<char*>(workspace) + align_up(sizeof(int) * ((size_t)n + 1), 256);

    DP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)n + 1), s));
    if (n > 0 && nnz > 0) {
        csr_count_cols_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, rowptr, col, cursor);
        DP_LAUNCH_CHECK();
    }
    int st = exclusive_scan_i32(cursor, rowptr_t, (long long)n + 1, scan_ws, s);
    if (st != DP_OK) return st;
    DP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)n + 1), s));
    if (n > 0 && nnz > 0) {
        csr_transpose_fill_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, rowptr, col, val, rowptr_t, cursor, col_t, val_t);
        DP_LAUNCH_CHECK();
        // flag: duplicates cannot appear in a transpose of a valid CSR; reuse cursor[n] as a sink
        csr_sort_rows_kernel<<<grid_for(n, 128), 128, 0, s>>>(n, rowptr_t, col_t, val_t, cursor + n);
        DP_LAUNCH_CHECK();
    }
    return DP_OK;
}

int dp_csr_inv_diagonal(int32_t n, const int32_t* rowptr, const int32_t* col, const double* val, double* dinv,
                        int32_t* flag_out, void* stream) {
    if (n < 0 || !rowptr || !dinv || !flag_out) return DP_ERR_INVALID;
    if (n == 0) return DP_OK;
    inv_diagonal_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, rowptr, col, val, dinv, flag_out);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

int dp_csr_aat_nnz(int32_t n, const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t, const int32_t* col_t,
                   int64_t* nnz_out, void* stream) {
    if (n < 0 || !nnz_out || (n > 0 && (!rowptr || !col || !rowptr_t || !col_t))) return DP_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    DP_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), s));
    if (n == 0) return DP_OK;
    aat_nnz_kernel<<<grid_for((long long)n * 32, 256), 256, 0, s>>>(n, rowptr, col, rowptr_t, col_t,
                                                                    reinterpret_cast<unsigned long long*>(nnz_out));
    DP_LAUNCH_CHECK();
    return DP_OK;
}

}  // extern "C"
