// sptrsv.cuh — K4 primitive: one warp resolves one plan chunk (<= 32 rows of ONE level) of a sparse triangular
// solve, sync-free. No reference counterpart; the arithmetic is plain substitution
//   x_i = (b_i - sum_j T_ij x_j) * (1 / T_ii),   sum sequential in column order, products/sums rounded separately
// so the result is bit-identical to oracle_sptrsv_lower/upper.
//
// Dependencies are awaited on the solution vector itself: x[] is pre-filled with the kPending NaN pattern and a
// consumer spins (ld.relaxed.gpu, L1-bypassing) until the 8-byte word changes — data is its own flag, one store
// per row, no fences. Rows of a chunk share a level, so they never depend on each other. Forward progress:
// chunks are handed to resident warps in plan order (round robin), every dependency lives in an earlier chunk.
#pragma once

#include "common.cuh"
#include "spmv.cuh"

namespace dp {

struct AbortCtl {
    unsigned long long* word;  // GridBarrier word (bit 63 = abort)
    int* flag;
    __device__ __forceinline__ bool aborted() const { return (ld_volatile_u64(word) & kAbortBit) != 0; }
    __device__ __forceinline__ void raise(int status) const {
        atomicOr(word, kAbortBit);
        atomicCAS(flag, 0, status);
    }
};

// Rhs functors: value of the right-hand side for `row`.
struct RhsPlain {
    const double* b;
    __device__ __forceinline__ double operator()(int row) const { return __ldcg(b + row); }
};
// Backward solve inside PCG: consume y[row] and immediately re-arm it for the next forward solve.
struct RhsConsume {
    double* y;
    __device__ __forceinline__ double operator()(int row) const {
        const double v = __ldcg(y + row);
        st_relaxed_u64(y + row, kPending);
        return v;
    }
};

// Returns false if the solve was aborted (spin budget exhausted somewhere). All 32 lanes must call.
// `light`: few warps take part in this solve (small levels), so every lane may poll all its dependencies right
// away; otherwise the warp first watches ONE word (one sector request per poll) until its chunk is about to be
// ready, which keeps the L2 request rate of thousands of waiting warps negligible.
template <bool kUpper, class Rhs>
__device__ __forceinline__ bool sptrsv_chunk(const CsrView& T, const int* __restrict__ plan32, const Rhs& rhs,
                                             double* x, const AbortCtl& ctl, bool light) {
    const int lane = threadIdx.x & 31;
    const int row = __ldg(plan32 + lane);
    const bool valid = row >= 0;
    int e = 0, end = 0;
    double rcp = 0.0, b = 0.0, sum = 0.0;
    int c[4] = {0, 0, 0, 0};
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    int probe = -1;
    if (valid) {
        const int rs = __ldg(T.rowptr + row), re = __ldg(T.rowptr + row + 1);
        e = kUpper ? rs + 1 : rs;        // off-diagonal entries [e, end)
        end = kUpper ? re : re - 1;
        // everything that does not depend on other rows is fetched before waiting
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e + k < end) c[k] = __ldg(T.col + e + k), v[k] = __ldg(T.val + e + k);
        rcp = __ddiv_rn(1.0, __ldg(T.val + (kUpper ? rs : re - 1)));
        b = rhs(row);
        // the off-diagonal entry nearest to the diagonal is (heuristically) the last to become available
        if (!light && lane == 0 && e < end) probe = __ldg(T.col + (kUpper ? e : end - 1));
    }
    if (!light) {
        probe = __shfl_sync(kFull, probe, 0);
        if (probe >= 0) {
            unsigned spins = 0;
            while (ld_relaxed_u64(x + probe) == kPending) {
                if (++spins > kSpinBudget) {
                    ctl.raise(DP_ERR_TIMEOUT);
                    return false;
                }
                if ((spins & 1023u) == 0 && ctl.aborted()) return false;
                __nanosleep(64);
            }
        }
    }
    unsigned idle = 0;
    for (;;) {
        const bool pending = valid && e < end;
        if (!__any_sync(kFull, pending)) break;
        bool progress = false;
        if (pending) {
            // up to 4 entries in flight; consumed strictly in column order
            const int m = min(4, end - e);
            unsigned long long u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = k < m ? ld_relaxed_u64(x + c[k]) : kPending;
            int used = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (used == k && u[k] != kPending) {
                    sum = __dadd_rn(sum, __dmul_rn(v[k], as_double(u[k])));
                    ++used;
                }
            }
            if (used) {
                e += used;
                progress = true;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (e + k < end) c[k] = __ldg(T.col + e + k), v[k] = __ldg(T.val + e + k);
            }
        }
        if (!__any_sync(kFull, progress)) {
            if (++idle > kSpinBudget) {
                ctl.raise(DP_ERR_TIMEOUT);
                return false;
            }
            if ((idle & 1023u) == 0 && ctl.aborted()) return false;
            __nanosleep(32);
        }
    }
    if (valid) st_relaxed_u64(x + row, as_bits(__dmul_rn(__dsub_rn(b, sum), rcp)));
    return true;
}

}  // namespace dp
