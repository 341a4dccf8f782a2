// common.cuh — shared device helpers for libdpcg (sm_100a). No torch, no libraries: CUDA runtime only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dpcg.h"

namespace dp {

constexpr int kWarp = 32;
constexpr int kBlock = 512;                  // threads per CTA of every row-parallel kernel
constexpr int kWarpsPerBlock = kBlock / kWarp;
constexpr int kTileRows = kBlock;            // one CTA pass covers 512 consecutive rows (16 chunks of 32)
constexpr unsigned kFull = 0xffffffffu;
// A kernel may carry warps beyond the kBlock row threads (the producer warp of the packed PCG engine, tilepipe.cuh): they
// walk the same control flow - every CTA barrier - and read the results of the CTA-wide reductions, but own no row and
// never contribute to a sum or a scan.
__device__ __forceinline__ bool row_thread() { return threadIdx.x < (unsigned)kBlock; }

// "Not yet produced" marker for sync-free dependency resolution: a quiet NaN with a payload that IEEE
// arithmetic on this GPU never generates (computed NaNs are canonical 0x7ff8000000000000).
constexpr unsigned long long kPending = 0xFFFBADC0FFEE0B20ull;

// Spin budgets (polls) before a kernel gives up and raises DP_ERR_TIMEOUT instead of hanging the GPU.
constexpr unsigned kSpinBudget = 1u << 24;

const char* set_cuda_error(cudaError_t e);
// Per-DEVICE host-side caches (spmv.cu): a process may solve on cuda:0 and then on cuda:1, and function attributes,
// occupancy and SM counts belong to the current device, not to the process or the host thread.
int sm_count();                                              // SMs of the current device
int allow_dynamic_smem(const void* kernel, size_t bytes);    // opt `kernel` into > 48 KB of dynamic shared memory (DP_OK / DP_ERR_CUDA)
int coop_grid(const void* kernel, int threads, size_t smem); // co-resident CTAs of `kernel` on the current device

#define DP_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) {                       \
            ::dp::set_cuda_error(e__);                  \
            return DP_ERR_CUDA;                         \
        }                                               \
    } while (0)

#define DP_LAUNCH_CHECK() DP_CUDA(cudaGetLastError())

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- memory-model helpers --------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const void* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(void* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const void* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_s32(void* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Read-only loads the compiler must issue WHERE THEY ARE WRITTEN: a plain __ldg is a pure expression to nvcc, which
// happily sinks a software-prefetch load down to its first use a whole tile later (measured: 700 cycles of exposed
// latency per tile in the level-stream solve). volatile asm keeps its place among the other volatile statements.
__device__ __forceinline__ int ldg_here_s32(const int* p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldcg_here_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double as_double(unsigned long long u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ unsigned long long as_bits(double d) { return (unsigned long long)__double_as_longlong(d); }

// ---- deterministic reductions ------------------------------------------------------------------------------
// Butterfly over the 32 lanes: every lane ends with the same bits; the order is fixed by the lane index only.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ double half_warp_sum(double v) {  // lanes 0..15 hold data
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Sum of one value per thread over the CTA (kBlock threads). `scratch` holds kWarpsPerBlock doubles.
// Result valid in every thread. Fixed order: lane butterfly, then a 16-wide butterfly over the warp sums.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // scratch may still be read by a previous call
    if (lane == 0 && warp < kWarpsPerBlock) scratch[warp] = v;
    __syncthreads();
    double w = scratch[lane & (kWarpsPerBlock - 1)];
    return half_warp_sum(w);
}

// Same reduction for kN values at once (one pair of CTA barriers instead of kN): per value bit-identical to block_sum.
// `scratch` holds kN * kWarpsPerBlock doubles.
template <int kN>
__device__ __forceinline__ void block_sum_n(double (&v)[kN], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0 && warp < kWarpsPerBlock) {
#pragma unroll
        for (int i = 0; i < kN; ++i) scratch[i * kWarpsPerBlock + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = half_warp_sum(scratch[i * kWarpsPerBlock + (lane & (kWarpsPerBlock - 1))]);
}

// Exclusive prefix sum of one int per thread over the CTA; *total = CTA sum. `sh` holds kWarpsPerBlock + 1 ints.
__device__ __forceinline__ int block_exclusive_scan_int(int v, int* total, int* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31 && warp < kWarpsPerBlock) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < kWarpsPerBlock ? sh[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kWarpsPerBlock) sh[lane] = winc - w;
        if (lane == 31) sh[kWarpsPerBlock] = winc;
    }
    __syncthreads();
    *total = sh[kWarpsPerBlock];
    return sh[warp & (kWarpsPerBlock - 1)] + inc - v;  // (warps beyond the row threads: value unused)
}

// Sum of `count` doubles stored at `p` (global, produced before the last grid-wide barrier): thread t adds
// p[t], p[t+kBlock], ... in that order, then block_sum. Same bits in every CTA.
__device__ __forceinline__ double block_reduce_array(const double* p, int count, double* scratch) {
    double v = 0.0;
    if (row_thread())
        for (int i = threadIdx.x; i < count; i += kBlock) v = __dadd_rn(v, __ldcg(p + i));
    return block_sum(v, scratch);
}

// ---- grid-wide barrier (cooperative launch guarantees co-residency) ---------------------------------------
// Monotonic 64-bit word: low 32 bits count arrivals, bit 63 is a sticky ABORT raised by any spin loop that ran
// out of budget. Same protocol as cooperative groups (fence - arrive - spin - fence by thread 0, bracketed by
// CTA barriers); sync() returns false in every thread of every CTA once ABORT is up, so kernels can unwind
// instead of hanging the GPU.
constexpr unsigned long long kAbortBit = 1ull << 63;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

constexpr unsigned long long kDoneUnit = 1ull << 32;  // bits 32..62: number of finished systems (PCG bookkeeping)

struct GridBarrier {
    unsigned long long* word;
    int* flag;   // device status word (DP_ERR_TIMEOUT)
    int* bcast;  // one int of shared memory
    unsigned epoch;
    unsigned nblocks;
    __device__ __forceinline__ void raise_abort(int status) {
        atomicOr(word, kAbortBit);
        atomicCAS(flag, 0, status);
    }
    __device__ __forceinline__ bool aborted() const { return (ld_volatile_u64(word) & kAbortBit) != 0; }
    // Returns -1 once ABORT is up, else the finished-systems count carried in the upper half of the word (the same
    // value in every CTA: increments happen strictly between two barriers).
    __device__ __forceinline__ int sync() {
        __syncthreads();
        epoch += nblocks;
        int info = 0;
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(word, 1ull);
            unsigned spins = 0;
            for (;;) {
                const unsigned long long v = ld_volatile_u64(word);
                if (v & kAbortBit) { info = -1; break; }
                if ((int)((unsigned)v - epoch) >= 0) { info = (int)((v >> 32) & 0x7fffffffu); break; }
                if (++spins > kSpinBudget) { raise_abort(DP_ERR_TIMEOUT); info = -1; break; }
            }
            __threadfence();
        }
        // broadcast thread 0's reading (note: __syncthreads_or reduces a predicate, not the integer)
        if (threadIdx.x == 0) *bcast = info;
        __syncthreads();
        return *bcast;
    }
};

}  // namespace dp
