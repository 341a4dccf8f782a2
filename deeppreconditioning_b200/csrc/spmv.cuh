// spmv.cuh — CSR view and the gather functors of the SpMV-shaped kernels (K2). The streaming engine itself is
// tilepipe.cuh; products and sums are rounded separately everywhere (__dmul_rn/__dadd_rn, no FMA contraction) so a
// row sum is bit-identical to the sequential CPU loop `sum += val[p] * x[col[p]]` (scipy csr_matvec /
// oracle_spmv_csr) — replaces `A @ p`, `M @ r`, cg.py:60,61,75,81.
#pragma once

#include "common.cuh"

namespace dp {

struct CsrView {
    const int* __restrict__ rowptr;
    const int* __restrict__ col;
    const double* __restrict__ val;
    int n;
    int nnz;
    // optional packed stream copy (dp_csr_pack): fp32 values, 16-bit columns relative to the tile's smallest column
    const unsigned short* __restrict__ col16;
    const float* __restrict__ val32;
    const int* __restrict__ tbase;
};

// Gather functors: what "x[c]" means for a phase. Vectors written earlier in the same persistent kernel are read
// with plain (coherent after the grid barrier's fence) loads; matrix data goes through the read-only path.
// kVecs > 0: the functor reads kVecs vectors of the solve's own work block (16-byte aligned, padded to 32 entries) and may
// be served from a shared-memory WINDOW of those vectors instead (tilepipe.cuh, packed stream): vec(i) names the vectors,
// combine() is the same arithmetic on the two loaded values, so both routes give the same bits.
struct GatherPlain {
    static constexpr int kVecs = 0;
    const double* x;
    __device__ __forceinline__ double operator()(int c) const { return x[c]; }
};
struct GatherReadOnly {  // standalone kernels: x is immutable for the whole launch
    static constexpr int kVecs = 0;
    const double* __restrict__ x;
    __device__ __forceinline__ double operator()(int c) const { return __ldg(x + c); }
};
struct GatherWork {  // a plain gather from one of the work vectors
    static constexpr int kVecs = 1;
    const double* x;
    __device__ __forceinline__ const double* vec(int) const { return x; }
    __device__ __forceinline__ double combine(double a, double) const { return a; }
    __device__ __forceinline__ double operator()(int c) const { return x[c]; }
};
// p_new[c] = z[c] + beta * p_old[c]   (cg.py:83), evaluated on the fly so the p-update needs no pass of its own.
struct GatherZBetaP {
    static constexpr int kVecs = 2;
    const double* z;
    const double* p;
    double beta;
    __device__ __forceinline__ const double* vec(int i) const { return i ? p : z; }
    __device__ __forceinline__ double combine(double zc, double pc) const { return __dadd_rn(zc, __dmul_rn(beta, pc)); }
    __device__ __forceinline__ double operator()(int c) const { return combine(z[c], p[c]); }
};
// r_new[c] = r_old[c] - a * Ap[c]     (cg.py:80), evaluated on the fly for the gather of L^T r / M r.
struct GatherRMinusAAp {
    static constexpr int kVecs = 2;
    const double* r;
    const double* ap;
    double a;
    __device__ __forceinline__ const double* vec(int i) const { return i ? ap : r; }
    __device__ __forceinline__ double combine(double rc, double apc) const { return __dsub_rn(rc, __dmul_rn(a, apc)); }
    __device__ __forceinline__ double operator()(int c) const { return combine(r[c], ap[c]); }
};

}  // namespace dp
