// spmv.cuh — warp-granular CSR "stream" SpMV primitive (K2), shared by the standalone kernel and the PCG phases.
//
// A warp owns 32 consecutive rows (a chunk). Its stored entries form ONE contiguous span of col[]/val[], which
// the warp streams with fully coalesced loads, multiplies by the gathered x[col] and parks the products in its
// private 2 KB slice of shared memory.
// Each lane then adds up the products of its own row, left to right. Products and sums are rounded separately
// (__dmul_rn/__dadd_rn, no FMA contraction) so a row sum is bit-identical to the sequential CPU loop
// `sum += val[p] * x[col[p]]` (scipy csr_matvec / oracle_spmv_csr) — replaces `A @ p`, `M @ r`, cg.py:60,61,75,81.
#pragma once

#include "common.cuh"

namespace dp {

struct CsrView {
    const int* __restrict__ rowptr;
    const int* __restrict__ col;
    const double* __restrict__ val;
    int n;
    int nnz;
};

// Gather functors: what "x[c]" means for a phase. Vectors written earlier in the same persistent kernel are read
// with plain (coherent after the grid barrier's fence) loads; matrix data goes through the read-only path.
struct GatherPlain {
    const double* x;
    __device__ __forceinline__ double operator()(int c) const { return x[c]; }
};
struct GatherReadOnly {  // standalone kernels: x is immutable for the whole launch
    const double* __restrict__ x;
    __device__ __forceinline__ double operator()(int c) const { return __ldg(x + c); }
};
// p_new[c] = z[c] + beta * p_old[c]   (cg.py:83), evaluated on the fly so the p-update needs no pass of its own.
struct GatherZBetaP {
    const double* z;
    const double* p;
    double beta;
    __device__ __forceinline__ double operator()(int c) const { return __dadd_rn(z[c], __dmul_rn(beta, p[c])); }
};
// r_new[c] = r_old[c] - a * Ap[c]     (cg.py:80), evaluated on the fly for the gather of L^T r / M r.
struct GatherRMinusAAp {
    const double* r;
    const double* ap;
    double a;
    __device__ __forceinline__ double operator()(int c) const { return __dsub_rn(r[c], __dmul_rn(a, ap[c])); }
};

// Row extent of this lane's row and of the whole 32-row chunk. Independent of any vector: can be issued early
// (before the scalars of a phase are known) to overlap its latency.
struct ChunkHead {
    int rs, re;  // entries of row (base + lane); empty for rows >= n
    int cs, ce;  // entries of the chunk
};

__device__ __forceinline__ ChunkHead spmv_head(const CsrView& A, int base) {
    const int row = base + (threadIdx.x & 31);
    ChunkHead h;
    h.rs = __ldg(A.rowptr + min(row, A.n));
    h.re = __ldg(A.rowptr + min(row + 1, A.n));
    h.cs = __shfl_sync(kFull, h.rs, 0);
    h.ce = __shfl_sync(kFull, h.re, 31);
    return h;
}

// Row sum of this lane's row. `stage` = this warp's kStageCap doubles of shared memory. All 32 lanes must call.
//
// The chunk's span is streamed in rounds of 32 consecutive entries per load instruction (lane l takes entry e0+l):
// col/val loads are perfectly coalesced (128 B / 256 B per instruction) and - what matters more - the 32 GATHER
// addresses of one instruction belong to ~6 consecutive rows, i.e. to a handful of cache lines (the stencil's
// diagonals), instead of 32 different rows: 4-5x fewer L1 wavefronts per gather than a quad-per-lane layout, and
// conflict-free shared-memory stores. kSpmvUnroll rounds are in flight before the first product is needed.
constexpr int kSpmvUnroll = 8;

template <class Gather>
__device__ __forceinline__ double spmv_body(const CsrView& A, const ChunkHead& h, const Gather& x, double* stage) {
    const int lane = threadIdx.x & 31;
    double sum = 0.0;
    for (int bs = h.cs; bs < h.ce; bs += kStageCap) {
        const int be = min(bs + kStageCap, h.ce);
        for (int e0 = bs + lane; e0 < be; e0 += kWarp * kSpmvUnroll) {
            int c[kSpmvUnroll];
            double v[kSpmvUnroll];
#pragma unroll
            for (int u = 0; u < kSpmvUnroll; ++u) {
                const int e = e0 + kWarp * u;
                c[u] = e < be ? __ldg(A.col + e) : -1;
            }
#pragma unroll
            for (int u = 0; u < kSpmvUnroll; ++u) {
                const int e = e0 + kWarp * u;
                v[u] = e < be ? __ldg(A.val + e) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kSpmvUnroll; ++u)
                if (c[u] >= 0) stage[e0 + kWarp * u - bs] = __dmul_rn(v[u], x(c[u]));
        }
        __syncwarp();
        const int lo = max(h.rs, bs), hi = min(h.re, be);
        for (int q = lo; q < hi; ++q) sum = __dadd_rn(sum, stage[q - bs]);
        __syncwarp();
    }
    return sum;
}

// Variant: every lane takes 4 consecutive entries per step (one int4 + two double2 loads). Fewer load
// instructions, but the 32 gather addresses of an instruction spread over ~25 rows. Same bits as spmv_body.
template <class Gather>
__device__ __forceinline__ double spmv_body_quad(const CsrView& A, const ChunkHead& h, const Gather& x, double* stage) {
    const int lane = threadIdx.x & 31;
    double sum = 0.0;
    for (int bs = h.cs & ~3; bs < h.ce; bs += kStageCap) {
        const int be = min(bs + kStageCap, h.ce);
#pragma unroll 2
        for (int e = bs + 4 * lane; e < be; e += 4 * kWarp) {
            int c0, c1, c2, c3;
            double v0, v1, v2, v3;
            if (e + 4 <= A.nnz) {
                const int4 cc = __ldg(reinterpret_cast<const int4*>(A.col + e));
                const double2 va = __ldg(reinterpret_cast<const double2*>(A.val + e));
                const double2 vb = __ldg(reinterpret_cast<const double2*>(A.val + e + 2));
                c0 = cc.x, c1 = cc.y, c2 = cc.z, c3 = cc.w;
                v0 = va.x, v1 = va.y, v2 = vb.x, v3 = vb.y;
            } else {  // last (partial) quad of the whole matrix
                c0 = (e + 0 < A.nnz) ? __ldg(A.col + e + 0) : 0;
                c1 = (e + 1 < A.nnz) ? __ldg(A.col + e + 1) : 0;
                c2 = (e + 2 < A.nnz) ? __ldg(A.col + e + 2) : 0;
                c3 = 0;
                v0 = (e + 0 < A.nnz) ? __ldg(A.val + e + 0) : 0.0;
                v1 = (e + 1 < A.nnz) ? __ldg(A.val + e + 1) : 0.0;
                v2 = (e + 2 < A.nnz) ? __ldg(A.val + e + 2) : 0.0;
                v3 = 0.0;
            }
            const double x0 = x(c0), x1 = x(c1), x2 = x(c2), x3 = x(c3);
            double2* dst = reinterpret_cast<double2*>(stage + (e - bs));
            dst[0] = make_double2(__dmul_rn(v0, x0), __dmul_rn(v1, x1));
            dst[1] = make_double2(__dmul_rn(v2, x2), __dmul_rn(v3, x3));
        }
        __syncwarp();
        const int lo = max(h.rs, bs), hi = min(h.re, be);
        for (int q = lo; q < hi; ++q) sum = __dadd_rn(sum, stage[q - bs]);
        __syncwarp();
    }
    return sum;
}

template <class Gather>
__device__ __forceinline__ double spmv_chunk(const CsrView& A, int base, const Gather& x, double* stage) {
    return spmv_body(A, spmv_head(A, base), x, stage);
}

}  // namespace dp
