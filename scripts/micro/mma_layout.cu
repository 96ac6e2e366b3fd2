// Micro-benchmark: tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) cycles per instruction as a function of the
// no-swizzle K-major operand layout (LBO = byte distance of the two K octets, SBO = byte distance of 8-row groups).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_layout mma_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
struct Cfg { int N, a_lbo, a_sbo, b_lbo, b_sbo, n_mma, a_step, b_step; };
__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        const long long t0 = clock64();
        for (int i = 0; i < c.n_mma; ++i) {
            const uint64_t ad = make_desc(a0 + (i & 7) * c.a_step, c.a_lbo, c.a_sbo), bd = make_desc(b0 + (i & 7) * c.b_step, c.b_lbo, c.b_sbo);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int n = 256;
    printf("%-5s %-7s %-7s %-7s %-7s %12s %12s\n", "N", "a_lbo", "a_sbo", "b_lbo", "b_sbo", "issue cyc/mma", "total cyc/mma");
    for (int N : {64, 128, 256}) {
        Cfg list[] = {
            {N, 2048, 128, N * 16, 128, n, 4096, 2 * N * 16},            // planes of 128 rows (heads kernels)
            {N, 128, 256, 128, 256, n, 4096, N * 32},                     // K-adjacent core matrices contiguous
            {N, 2048 + 128, 128, N * 16 + 128, 128, n, 4352, 2 * N * 16 + 256},   // padded plane pitch
            {N, 2048, 128, 128, 256, n, 4096, N * 32},                    // A planes, B canonical
            {N, 128, 256, N * 16, 128, n, 4096, 2 * N * 16},              // A canonical, B planes
            {N, 5136, 128, N * 16, 128, n, 10272, 2 * N * 16},            // odd plane pitch as in the conv kernels
        };
        for (const Cfg& c : list) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 200 * 1024>>>(c, d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); }
            printf("%-5d %-7d %-7d %-7d %-7d %12.1f %12.1f\n", c.N, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, (double)h[0] / n, (double)h[1] / n);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
