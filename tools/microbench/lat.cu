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
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
