// tma_stream.cu — sustained rate of a producer/consumer 1-D bulk-copy (TMA) pipeline per SM: how does it depend on the
// size of a copy, the number of copies per pipeline item, the depth of the pipeline and the number of CTAs streaming?
// (The level-stream and tile-stream triangular solves move 4 small arrays per 512-row tile; the SpMV pipeline 2 large ones.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
constexpr int kMaxStages = 16;
// item = `copies` bulk copies of `bytes` each (consecutive in memory); warp 0 lane 0 produces, warps 1..4 consume
// (wait full, arrive empty). Each CTA streams `items` items from its own region.
__global__ void stream(const char* src, size_t region, int bytes, int copies, int stages, int items, long long* cycles) {
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ __align__(8) unsigned long long full[kMaxStages], empty[kMaxStages];
    const int consumers = blockDim.x / 32 - 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(consumers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const char* mine = src + (size_t)blockIdx.x * region;
    const size_t item_bytes = (size_t)bytes * copies;
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int i = 0; i < items; ++i) {
            const int s = i % stages, use = i / stages;
            if (use > 0) while (!try_wait(&empty[s], (use - 1) & 1)) {}
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((unsigned)item_bytes) : "memory");
            for (int c = 0; c < copies; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(buf + (size_t)s * item_bytes + (size_t)c * bytes)),
                             "l"(mine + ((size_t)i * item_bytes + (size_t)c * bytes) % region), "r"(bytes), "r"(smem_u32(&full[s]))
                             : "memory");
        }
    } else if (threadIdx.x >= 32) {
        for (int i = 0; i < items; ++i) {
            const int s = i % stages, use = i / stages;
            while (!try_wait(&full[s], use & 1)) {}
            __syncwarp();
            if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
    const size_t total = 8ull << 30;  // 8 GB source: far beyond the 126 MB L2
    char* src; long long* out;
    cudaMalloc(&src, total); cudaMemset(src, 1, total); cudaMalloc(&out, 8 * 1024);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grids[] = {1, 148, 296};
    struct Cfg { int bytes, copies, stages; } cfgs[] = {
        {24576, 1, 3}, {12288, 2, 3}, {6144, 4, 3},      // one 24 KB item as 1 / 2 / 4 copies, 3 stages (tile-stream geometry)
        {24576, 1, 7}, {12288, 2, 7}, {6144, 4, 7},      // 7 stages (level-stream geometry)
        {46080, 1, 2}, {23040, 2, 2},                    // SpMV geometry: 2 stages of 46 KB
        {49152, 1, 3}, {8192, 1, 8}, {4096, 1, 16}, {16384, 1, 8}};
    for (int g : grids)
        for (auto c : cfgs) {
            const size_t item = (size_t)c.bytes * c.copies;
            const size_t smem = item * c.stages;
            if (smem > 200 * 1024) continue;
            if (g == 296 && smem > 100 * 1024) continue;
            const int items = 2000;
            const size_t region = total / g / item * item;
            stream<<<g, 160, smem>>>(src, region, c.bytes, c.copies, c.stages, 200, out);  // warm
            cudaEventRecord(e0);
            stream<<<g, 160, smem>>>(src, region, c.bytes, c.copies, c.stages, items, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double gbs = (double)g * items * item / ms / 1e6;
            printf("%3d CTAs, item %5zu B = %d x %5d B, %2d stages: %8.1f GB/s total, %6.1f GB/s per CTA, %6.0f ns per item\n", g, item,
                   c.copies, c.bytes, c.stages, gbs, gbs / g, 1e6 * ms / items);
        }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
