// scan.cu — three-kernel exclusive scan (CTA sums -> scan of sums -> local scan + offset).
#include "scan.cuh"

namespace dp {

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* sh /* kScanThreads/32 + 1 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? sh[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kScanThreads / 32) sh[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) sh[kScanThreads / 32] = winc;       // CTA total
    }
    __syncthreads();
    *total = sh[kScanThreads / 32];
    return sh[warp] + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(const int* __restrict__ in, long long m, int* __restrict__ sums) {
    __shared__ int sh[kScanThreads / 32 + 1];
    const long long base = (long long)blockIdx.x * kScanSpan;
    int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + (long long)k * kScanThreads + threadIdx.x;
        if (i < m) v += in[i];
    }
    int total;
    block_exclusive_scan(v, &total, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// Single CTA: exclusive scan of sums[0..nb) in place, any nb.
__global__ void __launch_bounds__(kScanThreads) scan_of_sums_kernel(int* sums, int nb) {
    __shared__ int sh[kScanThreads / 32 + 1];
    int carry = 0;
    for (int start = 0; start < nb; start += kScanThreads) {
        const int i = start + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total, sh);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* in, int* out, long long m, const int* __restrict__ sums) {
    __shared__ int sh[kScanThreads / 32 + 1];
    const long long base = (long long)blockIdx.x * kScanSpan + (long long)threadIdx.x * kScanItems;
    int v[kScanItems];
    int local = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < m) ? in[base + k] : 0;
        local += v[k];
    }
    int total;
    int run = sums[blockIdx.x] + block_exclusive_scan(local, &total, sh);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < m) out[base + k] = run;
        run += v[k];
    }
}

size_t scan_workspace_bytes(long long m) {
    long long nb = (m + kScanSpan - 1) / kScanSpan;
    return align_up((size_t)(nb > 0 ? nb : 1) * sizeof(int), 256);
}

int exclusive_scan_i32(const int* in, int* out, long long m, void* ws, cudaStream_t stream) {
    if (m <= 0) return DP_OK;
    const int nb = (int)((m + kScanSpan - 1) / kScanSpan);
    int* sums = static_cast<int*>(ws);
    scan_sums_kernel<<<nb, kScanThreads, 0, stream>>>(in, m, sums);
    DP_LAUNCH_CHECK();
    scan_of_sums_kernel<<<1, kScanThreads, 0, stream>>>(sums, nb);
    DP_LAUNCH_CHECK();
    scan_apply_kernel<<<nb, kScanThreads, 0, stream>>>(in, out, m, sums);
    DP_LAUNCH_CHECK();
    return DP_OK;
}

}  // namespace dp
