// tma_issue.cu — how long does the ISSUING thread spend in cp.async.bulk (1-D TMA copy), and when do the bytes land?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue tma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const char* src, int bytes, int copies, long long* out) {
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes * copies) : "memory");
        long long t1 = clock64();
        for (int c = 0; c < copies; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(buf + (size_t)c * bytes)),
                         "l"(src + (size_t)c * bytes), "r"(bytes), "r"(smem_u32(&bar))
                         : "memory");
        long long t2 = clock64();
        unsigned ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        long long t3 = clock64();
        out[0] = t1 - t0, out[1] = t2 - t1, out[2] = t3 - t2;
    }
}
int main() {
    char* src; long long* out;
    cudaMalloc(&src, 64 << 20); cudaMemset(src, 1, 64 << 20); cudaMalloc(&out, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int sizes[] = {1024, 4096, 16384, 49152, 98304};
    for (int copies : {1, 2, 4})
        for (int b : sizes) {
            if ((long long)b * copies > 196608) continue;
            long long h[3];
            for (int rep = 0; rep < 2; ++rep) {  // second run: source in L2
                probe<<<1, 32, b * copies>>>(src + (rep == 0 ? (size_t)(copies * 8 + b / 1024) << 20 >> 4 : (size_t)(copies * 8 + b / 1024) << 20 >> 4), b, copies, out);
                cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
                printf("%d x %6d B (%s): expect_tx %lld cyc, issue of the copies %lld cyc, then %lld cyc until landed\n", copies, b,
                       rep ? "L2" : "DRAM", h[0], h[1], h[2]);
            }
        }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
