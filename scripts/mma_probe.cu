// Micro-probe: cycles per tcgen05.mma (M=128, K=16, fp16, no-swizzle K-major smem operands) as a function of
// N and of how many independent TMEM accumulators the stream of MMAs rotates over.
#include "../openwakeword_b200/csrc/tc_common.cuh"
#include <cuda_runtime.h>
#include <vector>

__global__ void __launch_bounds__(128, 1) probe(int N, int n_acc, int n_mma, int lbo_units, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        const long long t0 = clock64();
        for (int k = 0; k < n_mma; ++k) {
            const uint64_t ad = make_desc(a0 + (k % 8) * 16, (uint32_t)lbo_units * 16u, 128u);
            const uint64_t bd = make_desc(b0, (uint32_t)N * 16u, 128u);
            tc_mma_f16(tm + (uint32_t)(k % n_acc) * (512 / n_acc), ad, bd, idesc, k >= n_acc);
        }
        const long long t1 = clock64();
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512)); }
}

// variant: straight-line issue, descriptors hoisted, issued from a converged warp through elect.sync
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int NACC>
__global__ void __launch_bounds__(128, 1) probe2(int N, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t ad0 = make_desc(smem_u32(smem), 136u * 16u, 128u);
        const uint64_t bd0 = make_desc(smem_u32(smem) + 96 * 1024, (uint32_t)N * 16u, 128u);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    tc_mma_f16(tm + (uint32_t)(k % NACC) * (512 / NACC), ad0 + (uint64_t)(k & 7), bd0, idesc, (rep | (k >= NACC)) ? 1u : 0u);
            }
            t1 = clock64();
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512)); }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Ns[] = {32, 96, 256};
    const int accs[] = {1, 2, 4};
    const int lbos[] = {136};
    for (int lbo : lbos) for (int N : Ns) for (int na : accs) {
        if (N * na > 512) continue;
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<1, 128, 170 * 1024>>>(N, na, 64, lbo, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        printf("lbo %5d units  N=%3d  accumulators=%d : issue %6lld cyc, done %6lld cyc -> %.1f cyc/MMA (pipe floor %d)\n",
               lbo, N, na, h[0], h[1], h[1] / 64.0, 128 * N / 256);
    }
    cudaFuncSetAttribute(probe2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : Ns) for (int na = 1; na <= 2; ++na) {
        if (N * na > 512) continue;
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
            if (na == 1) probe2<1><<<1, 128, 170 * 1024>>>(N, d); else probe2<2><<<1, 128, 170 * 1024>>>(N, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        printf("straight-line elect  N=%3d  accumulators=%d : issue %6lld cyc, done %6lld cyc -> %.1f cyc/MMA (pipe floor %d)\n",
               N, na, h[0], h[1], h[1] / 64.0, 128 * N / 256);
    }
    return 0;
}
