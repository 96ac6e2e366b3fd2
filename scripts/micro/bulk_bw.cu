// Micro-benchmark: per-SM and chip throughput of cp.async.bulk (global -> shared, mbarrier complete_tx) as a function of
// the copy size, the number of copies in flight and where the data lives (L2-resident shared block vs unique HBM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// each CTA: `iters` stages of `n_sub` copies of `sub` bytes, `depth` stages in flight
__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, size_t cta_stride, size_t region, int iters, int depth, int n_sub, int sub) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* ring = smem + 1024;
    const uint32_t b0 = smem_u32(bars);
    if (threadIdx.x == 0) { for (int i = 0; i < 2 * depth; ++i) mbar_init(b0 + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const size_t stage = (size_t)n_sub * sub;
    const uint8_t* base = src + blockIdx.x * cta_stride;
    if (threadIdx.x == 0) {
        size_t off = 0;
        for (int it = 0; it < iters; ++it) {
            const int s = it % depth;
            mbar_wait(b0 + 8 * (depth + s), ((it / depth) & 1) ^ 1);
            mbar_expect_tx(b0 + 8 * s, (uint32_t)stage);
            for (int j = 0; j < n_sub; ++j) bulk_g2s(smem_u32(ring + s * stage + (size_t)j * sub), base + off + (size_t)j * sub, sub, b0 + 8 * s);
            off += stage; if (off + stage > region) off = 0;
        }
    } else if (threadIdx.x == 32) {
        for (int it = 0; it < iters; ++it) {
            const int s = it % depth;
            mbar_wait(b0 + 8 * s, (it / depth) & 1);
            mbar_arrive(b0 + 8 * (depth + s));
        }
    }
}
int main() {
    const size_t total = (size_t)4 << 30;
    uint8_t* d; cudaMalloc(&d, total); cudaMemset(d, 1, total);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int subs[] = {1024, 4096, 16384, 49152};
    printf("%-8s %-6s %-6s %-7s %-6s %10s %10s\n", "where", "ctas", "sub", "n_sub", "depth", "GB/s", "GB/s/SM");
    for (int where = 0; where < 2; ++where)
        for (int ctas : {1, 128})
            for (int sub : subs)
                for (int stage_kb : {16, 64})
                    for (int depth : {2, 3, 6}) {
                        const int n_sub = stage_kb * 1024 / sub;
                        if (n_sub < 1 || (size_t)depth * stage_kb * 1024 > 220 * 1024) continue;
                        const size_t per_cta = (size_t)24 << 20;                 // bytes each CTA pulls
                        const int iters = (int)(per_cta / (stage_kb * 1024));
                        // where 0: every CTA streams the same 2 MB block (L2 hits); 1: its own 24 MB (HBM)
                        const size_t cta_stride = where ? per_cta : 0, region = where ? per_cta : ((size_t)2 << 20);
                        for (int rep = 0; rep < 2; ++rep) {
                            cudaEventRecord(e0);
                            k<<<ctas, 64, 1024 + depth * stage_kb * 1024>>>(d, cta_stride, region, iters, depth, n_sub, sub);
                            cudaEventRecord(e1); cudaEventSynchronize(e1);
                        }
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        const double gbs = (double)per_cta * ctas / (ms * 1e6);
                        printf("%-8s %-6d %-6d %-7d %-6d %10.1f %10.1f\n", where ? "hbm" : "l2", ctas, sub, n_sub, depth, gbs, gbs / ctas);
                    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
