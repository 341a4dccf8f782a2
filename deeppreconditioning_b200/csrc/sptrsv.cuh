// sptrsv.cuh — K4 primitive: a warp resolves a SEQUENCE of plan chunks (<= 32 rows of ONE level each) of a sparse
// triangular solve, sync-free. No reference counterpart; the arithmetic is plain substitution
//   x_i = (b_i - sum_j T_ij x_j) * (1 / T_ii),   sum sequential in column order, products/sums rounded separately
// so the result is bit-identical to oracle_sptrsv_lower/upper.
//
// Dependencies are awaited on the solution vector itself: x[] is pre-filled with the kPending NaN pattern and a
// consumer spins (ld.relaxed.gpu, L1-bypassing) until the 8-byte word changes — data is its own flag, one store
// per row, no fences. Rows of a chunk share a level, so they never depend on each other. Forward progress:
// chunks are handed to resident warps in plan order (round robin), every dependency lives in an earlier chunk.
//
// The critical path of a solve is (levels) x (store -> L2 -> poll hit). Everything else is taken off that path:
//   * chunk metadata is software-pipelined (plan row two chunks ahead, row extent one chunk ahead, entries / diagonal /
//     rhs requested on arrival, before any waiting), so the dependent loads of a chunk never queue up behind its wait;
//   * a warp whose chunk is still levels away parks on ONE word (lane 0, x of a row `lookback` chunks earlier in the
//     plan, i.e. about one level back: one sector request per poll); once that is solved every lane polls its own dependencies
//     back to back, so the last store is seen one L2 round trip later and the L2 never sees thousands of idle pollers.
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

constexpr int kTrsvInflight = 4;  // dependencies polled per lane and round

// Everything a lane needs to solve its row except the dependencies' values.
struct TrsvRow {
    int row;       // -1: idle lane
    int e, end;    // off-diagonal entries [e, end)
    double rcp, b;
    int c[kTrsvInflight];
    double v[kTrsvInflight];
};

template <bool kUpper>
__device__ __forceinline__ void trsv_row_extent(const CsrView& T, int row, int& e, int& end, int& dpos) {
    e = end = dpos = 0;
    if (row >= 0) {
        const int rs = __ldg(T.rowptr + row), re = __ldg(T.rowptr + row + 1);
        e = kUpper ? rs + 1 : rs;
        end = kUpper ? re : re - 1;
        dpos = kUpper ? rs : re - 1;
    }
}

template <class Rhs>
__device__ __forceinline__ void trsv_row_entries(const CsrView& T, int row, int e, int end, int dpos, const Rhs& rhs,
                                                 TrsvRow& r) {
    r.row = row, r.e = e, r.end = end;
    r.rcp = 0.0, r.b = 0.0;
#pragma unroll
    for (int k = 0; k < kTrsvInflight; ++k) r.c[k] = 0, r.v[k] = 0.0;
    if (row >= 0) {
#pragma unroll
        for (int k = 0; k < kTrsvInflight; ++k)
            if (e + k < end) r.c[k] = __ldg(T.col + e + k), r.v[k] = __ldg(T.val + e + k);
        r.rcp = __ddiv_rn(1.0, __ldg(T.val + dpos));
        r.b = rhs(row);
    }
}

// Solve chunks c0, c0 + stride, ... < cend of the plan. `lookback` = chunks per (largest) level: the parking word of
// chunk c is the first row of chunk c - lookback. Returns false if the solve was aborted. All 32 lanes must call.
template <bool kUpper, class Rhs>
__device__ __forceinline__ bool sptrsv_stream(const CsrView& T, const int* __restrict__ plan, long long c0, long long stride,
                                              long long cend, int lookback, const Rhs& rhs, double* x, const AbortCtl& ctl) {
    const int lane = threadIdx.x & 31;
    auto plan_row = [&](long long c) { return c < cend ? __ldg(plan + c * 32 + lane) : -1; };
    auto park_row = [&](long long c) { return (c < cend && c >= lookback) ? __ldg(plan + (c - lookback) * 32) : -1; };

    // pipeline fill: chunk c0 up to its row extent, c0 + stride up to its row index
    int row0, e0, end0, dpos0, park0;
    int row1, park1;
    row0 = plan_row(c0), park0 = park_row(c0);
    row1 = plan_row(c0 + stride), park1 = park_row(c0 + stride);
    trsv_row_extent<kUpper>(T, row0, e0, end0, dpos0);
    for (long long c = c0; c < cend; c += stride) {
        // everything that does not depend on other rows is requested before waiting on anything: this chunk's
        // entries, the next chunk's row extent, the one after's row index
        TrsvRow cur;
        trsv_row_entries(T, row0, e0, end0, dpos0, rhs, cur);
        const int row2 = plan_row(c + 2 * stride), park2 = park_row(c + 2 * stride);
        int e1, end1, dpos1;
        trsv_row_extent<kUpper>(T, row1, e1, end1, dpos1);

        // park until the plan is about one level away from this chunk
        if (park0 >= 0) {
            unsigned spins = 0;
            while (ld_relaxed_u64(x + park0) == kPending) {
                if (++spins > kSpinBudget) {
                    ctl.raise(DP_ERR_TIMEOUT);
                    return false;
                }
                if ((spins & 1023u) == 0 && ctl.aborted()) return false;
            }
        }
        // resolve: every lane polls its own dependencies, a batch of kTrsvInflight at a time, and re-reads only the ones
        // that are still pending; a complete batch is summed in column order
        double sum = 0.0;
        unsigned idle = 0;
        unsigned long long u[kTrsvInflight];
#pragma unroll
        for (int k = 0; k < kTrsvInflight; ++k) u[k] = kPending;
        for (;;) {
            const bool pending = cur.row >= 0 && cur.e < cur.end;
            if (!__any_sync(kFull, pending)) break;
            bool progress = false;
            if (pending) {
                const int m = min(kTrsvInflight, cur.end - cur.e);
                bool all = true;
#pragma unroll
                for (int k = 0; k < kTrsvInflight; ++k) {
                    if (k < m && u[k] == kPending) {
                        u[k] = ld_relaxed_u64(x + cur.c[k]);
                        all = all && u[k] != kPending;
                    }
                }
                if (all) {
#pragma unroll
                    for (int k = 0; k < kTrsvInflight; ++k)
                        if (k < m) sum = __dadd_rn(sum, __dmul_rn(cur.v[k], as_double(u[k])));
                    cur.e += m;
                    progress = true;
#pragma unroll
                    for (int k = 0; k < kTrsvInflight; ++k) u[k] = kPending;
                    if (cur.e < cur.end) {  // rows with more than kTrsvInflight dependencies: fetch the next batch
#pragma unroll
                        for (int k = 0; k < kTrsvInflight; ++k)
                            if (cur.e + k < cur.end) cur.c[k] = __ldg(T.col + cur.e + k), cur.v[k] = __ldg(T.val + cur.e + k);
                    }
                }
            }
            if (!__any_sync(kFull, progress)) {
                if (++idle > kSpinBudget) {
                    ctl.raise(DP_ERR_TIMEOUT);
                    return false;
                }
                if ((idle & 1023u) == 0 && ctl.aborted()) return false;
            }
        }
        if (cur.row >= 0) st_relaxed_u64(x + cur.row, as_bits(__dmul_rn(__dsub_rn(cur.b, sum), cur.rcp)));
        // rotate the pipeline
        row0 = row1, e0 = e1, end0 = end1, dpos0 = dpos1, park0 = park1;
        row1 = row2, park1 = park2;
    }
    return true;
}

}  // namespace dp
