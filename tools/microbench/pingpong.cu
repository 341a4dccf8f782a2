// pingpong.cu — store -> poll latency between two SMs through global memory (the hop that bounds sync-free SpTRSV).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu ; ./pingpong
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_cg(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

// mode 0: relaxed.gpu ld/st; 1: volatile ld + relaxed st; 2: ld.cg + st; 3: whole warp polls (32 lanes same word)
// CTA a and CTA b bounce a counter: a writes odd values to w[0], b answers on w[16] (different lines).
__global__ void pingpong(unsigned long long* w, int a, int b, int iters, int mode, long long* out, unsigned* sm) {
    const int me = blockIdx.x == a ? 0 : blockIdx.x == b ? 1 : -1;
    if (me < 0) return;
    const bool warp_poll = mode == 3;
    if (!warp_poll && threadIdx.x != 0) return;
    if (threadIdx.x >= 32) return;
    unsigned long long* mine = w + (me == 0 ? 0 : 16);
    const unsigned long long* theirs = w + (me == 0 ? 16 : 0);
    if (threadIdx.x == 0) sm[me] = smid();
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            if (threadIdx.x == 0) st_relaxed(mine, (unsigned long long)i);
            unsigned long long v;
            do {
                v = mode == 1 ? ld_volatile(theirs) : mode == 2 ? ld_cg(theirs) : ld_relaxed(theirs);
            } while (v < (unsigned long long)i);
        } else {
            unsigned long long v;
            do {
                v = mode == 1 ? ld_volatile(theirs) : mode == 2 ? ld_cg(theirs) : ld_relaxed(theirs);
            } while (v < (unsigned long long)i);
            if (threadIdx.x == 0) st_relaxed(mine, (unsigned long long)i);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[me] = t1 - t0;
}

// plain L2 load latency by pointer chasing with ld.relaxed.gpu / ld.cg
__global__ void chase(const unsigned long long* p, int iters, int mode, long long* out) {
    unsigned long long idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = mode == 0 ? ld_relaxed(p + idx) : ld_cg(p + idx);
    long long t1 = clock64();
    out[0] = t1 - t0;
    out[1] = (long long)idx;
}

int main() {
    unsigned long long* w;
    long long* out;
    unsigned* sm;
    cudaMalloc(&w, 4096);
    cudaMalloc(&out, 64);
    cudaMalloc(&sm, 64);
    const int iters = 2000;
    const char* names[] = {"ld.relaxed.gpu", "ld.volatile", "ld.cg", "warp-wide ld.relaxed.gpu"};
    int pairs[][2] = {{0, 1}, {0, 2}, {0, 37}, {0, 74}, {0, 100}, {0, 147}, {10, 120}};
    for (int mode = 0; mode < 4; ++mode)
        for (auto& pr : pairs) {
            cudaMemset(w, 0, 4096);
            pingpong<<<148, 32>>>(w, pr[0], pr[1], iters, mode, out, sm);
            long long h[2];
            unsigned hs[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(hs, sm, 8, cudaMemcpyDeviceToHost);
            printf("%-26s cta %3d(sm %3u) <-> cta %3d(sm %3u): %.0f cycles per hop (round trip / 2)\n", names[mode], pr[0],
                   hs[0], pr[1], hs[1], (double)h[0] / iters / 2.0);
        }
    // pointer chase over a small L2-resident ring (stride 4 KB)
    const int n = 4096;
    unsigned long long* ring;
    cudaMalloc(&ring, (size_t)n * 512 * 8);
    unsigned long long* hring = new unsigned long long[(size_t)n * 512]();
    for (int i = 0; i < n; ++i) hring[(size_t)i * 512] = (unsigned long long)((i + 1) % n) * 512;
    cudaMemcpy(ring, hring, (size_t)n * 512 * 8, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        chase<<<1, 1>>>(ring, n, mode, out);
        chase<<<1, 1>>>(ring, n, mode, out);
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("pointer chase %s (L2-resident 16 MB ring): %.0f cycles per load\n", mode == 0 ? "ld.relaxed.gpu" : "ld.cg", (double)h[0] / n);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
