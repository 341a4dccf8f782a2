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

}  // namespace dp
