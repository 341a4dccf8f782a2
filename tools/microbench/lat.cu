// lat.cu — dependent-chain latencies on B200 that bound the level-stream solve: fp64 add/mul, shared-memory load,
// CTA barrier with 16 warps. nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dchain(double a, double b, int n, long long* out, double* sink) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = __dadd_rn(x, b);
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) x = __dmul_rn(x, b);
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) x = __dadd_rn(__dmul_rn(x, b), a);
    long long t3 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t1, out[2] = t3 - t2;
    sink[threadIdx.x] = x;
}
__global__ void ldschain(int n, long long* out, int* sink) {
    __shared__ int ring[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) ring[i] = (i + 33) & 1023;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) idx = ring[idx];
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = idx;
}
__global__ void barchain(int n, int active_warps, long long* out, double* sink) {
    __shared__ double win[1024];
    win[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double x = 1.0;
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (warp < active_warps) {  // the level-stream inner step: 3 window reads, 3 mul+add, scale, window write
            double s = 0.0;
            s = __dadd_rn(s, __dmul_rn(1.0000001, win[(threadIdx.x + i) & 1023]));
            s = __dadd_rn(s, __dmul_rn(0.9999999, win[(threadIdx.x + i + 317) & 1023]));
            s = __dadd_rn(s, __dmul_rn(1.0000002, win[(threadIdx.x + i + 5) & 1023]));
            x = __dmul_rn(__dsub_rn(x, s), 0.5);
            win[(threadIdx.x + i + 158) & 1023] = x;
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// The same level step ordered by a ring of mbarriers (16 arrivals per step, one per warp): owners of step i are the
// warps [first, first + active) rotating with i; they wait for step i-1, compute, arrive; the others just arrive.
template <bool kTest>
__global__ void ringchain(int n, int active_warps, long long* out, double* sink) {
    __shared__ double win[1024];
    __shared__ __align__(8) unsigned long long ring[16];
    win[threadIdx.x] = threadIdx.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" ::"r"(smem_u32(&ring[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    double x = 1.0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        const int first = (i * 5) & 15;
        const bool owner = ((warp - first) & 15) < active_warps;
        if (owner) {
            if (i > 0) {
                unsigned long long* b = &ring[(i - 1) & 15];
                const unsigned par = ((i - 1) >> 4) & 1;
                while (!(kTest ? test_wait(b, par) : try_wait(b, par))) {
                }
            }
            double s = 0.0;
            s = __dadd_rn(s, __dmul_rn(1.0000001, win[(threadIdx.x + i) & 1023]));
            s = __dadd_rn(s, __dmul_rn(0.9999999, win[(threadIdx.x + i + 317) & 1023]));
            s = __dadd_rn(s, __dmul_rn(1.0000002, win[(threadIdx.x + i + 5) & 1023]));
            x = __dmul_rn(__dsub_rn(x, s), 0.5);
            win[(threadIdx.x + i + 158) & 1023] = x;
        }
        if ((i & 7) == 0 && i >= 8) {
            unsigned long long* b = &ring[(i - 8) & 15];
            while (!try_wait(b, ((i - 8) >> 4) & 1)) {
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ring[i & 15])) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x;
}
int main() {
    long long* out; double* sink; int* isink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 8192); cudaMalloc(&isink, 8192);
    long long h[3];
    const int n = 4096;
    dchain<<<1, 32>>>(1.0, 1.0000001, n, out, sink); dchain<<<1, 32>>>(1.0, 1.0000001, n, out, sink);
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("dependent DADD %.1f, DMUL %.1f, DMUL+DADD %.1f cycles\n", (double)h[0] / n, (double)h[1] / n, (double)h[2] / n);
    ldschain<<<1, 32>>>(n, out, isink); ldschain<<<1, 32>>>(n, out, isink);
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("dependent LDS %.1f cycles\n", (double)h[0] / n);
    for (int aw : {0, 1, 5, 16}) {
        barchain<<<1, 512>>>(n, aw, out, sink); barchain<<<1, 512>>>(n, aw, out, sink);
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("level step (16 warps, %2d active) %.1f cycles per step\n", aw, (double)h[0] / n);
    }
    for (int aw : {1, 5, 16}) {
        ringchain<false><<<1, 512>>>(n, aw, out, sink); ringchain<false><<<1, 512>>>(n, aw, out, sink);
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("level step, mbarrier ring + try_wait (16 warps, %2d owners) %.1f cycles per step\n", aw, (double)h[0] / n);
        ringchain<true><<<1, 512>>>(n, aw, out, sink); ringchain<true><<<1, 512>>>(n, aw, out, sink);
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("level step, mbarrier ring + test_wait spin (16 warps, %2d owners) %.1f cycles per step\n", aw, (double)h[0] / n);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
