// Micro-probes for the two open questions behind the fused kernel's heads phase:
//  (1) how fast can ONE SM stream an L2-resident matrix through a shared-memory ring with cp.async.bulk, as a function
//      of chunk size, ring depth and how many SMs do it at once (the first-layer matrix, 786 KB per CTA, is streamed
//      that way; the phase behaves as if the rate were ~20 B/clk/SM);
//  (2) what a warp-wide shared-memory BROADCAST load costs per instruction for 4 / 8 / 16-byte accesses and 1..16 warps
//      (the first layer reads every feature value as a broadcast).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/smem_probe scripts/smem_probe.cu
// Run on the GPU box: ./scripts/smem_probe      (prints one line per configuration; nothing here is on the product path)
#include "../openwakeword_b200/csrc/tc_common.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

// ---- (1) bulk-copy streaming: warp 0 lane 0 produces, warps 1..4 consume (wait + touch one word + release) ------------
__global__ void __launch_bounds__(160, 1) stream_probe(const uint8_t* src, size_t src_bytes, int chunk, int slots, int n_chunks,
                                                       long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[8], empty[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const uint8_t* my = src + ((size_t)blockIdx.x * 65536) % (src_bytes - (size_t)chunk * 64);   // stay inside the buffer
    const long long t0 = clock64();
    if (warp == 0) {
        if (lane == 0) {
            uint32_t par = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % slots;
                mbar_wait(smem_u32(&empty[s]), ((par >> s) & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(&full[s]), (uint32_t)chunk);
                bulk_g2s(smem_u32(smem + (size_t)s * chunk), my + (size_t)(c % 64) * chunk, (uint32_t)chunk, smem_u32(&full[s]));
                par ^= 1u << s;
            }
        }
    } else {
        uint32_t par = 0, acc = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % slots;
            mbar_wait(smem_u32(&full[s]), (par >> s) & 1u);
            par ^= 1u << s;
            acc += reinterpret_cast<const uint32_t*>(smem + (size_t)s * chunk)[threadIdx.x];
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
        }
        if (acc == 0x12345678u) out[2] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

// ---- (2) broadcast loads: every lane of every warp reads the SAME address, VEC floats wide, n_iter times ---------------
template <int VEC>
__global__ void __launch_bounds__(512, 1) bcast_probe(int n_iter, long long* out) {
    __shared__ __align__(16) float buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 1.0f;
    __syncthreads();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;             // independent chains: the loads, not the adds, set the pace
    const long long t0 = clock64();
    for (int it = 0; it < n_iter; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int o = ((it * 8 + u) * VEC) & 4095 & ~(VEC - 1);
            if (VEC == 1) { if (u & 1) a1 += buf[o]; else a0 += buf[o]; }
            if (VEC == 2) { const float2 v = *reinterpret_cast<const float2*>(buf + o); a0 += v.x; a1 += v.y; }
            if (VEC == 4) { const float4 v = *reinterpret_cast<const float4*>(buf + o); a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
        }
    }
    const long long t1 = clock64();
    if (a0 + a1 + a2 + a3 == 12345.f) out[3] = 1;
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 64);
    const size_t src_bytes = 64u << 20;                          // 64 MB: L2 resident on B200 after the first pass
    uint8_t* d_src; cudaMalloc(&d_src, src_bytes); cudaMemset(d_src, 1, src_bytes);
    cudaFuncSetAttribute(stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("# bulk-copy streaming through a smem ring (bytes per clock per SM)\n");
    for (int grid : {1, sms}) for (int chunk : {12288, 24576, 49152}) for (int slots : {1, 2, 4}) {
        if ((size_t)chunk * slots > 196608) continue;
        const int n_chunks = 786432 / chunk * 4;                 // four passes of a 768 KB matrix
        long long h[4] = {0, 0, 0, 0};
        for (int rep = 0; rep < 3; ++rep) {                      // first reps warm L2
            stream_probe<<<grid, 160, (size_t)chunk * slots>>>(d_src, src_bytes, chunk, slots, n_chunks, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("stream_probe failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("grid %3d chunk %5d B slots %d : %8lld clk for %d chunks = %6.1f B/clk/SM\n", grid, chunk, slots, h[0], n_chunks,
               (double)chunk * n_chunks / (double)h[0]);
    }
    printf("# shared-memory broadcast loads (clocks per warp-wide load instruction, all warps of the CTA loading)\n");
    for (int warps : {1, 4, 8, 16}) {
        long long h[4];
        const int n_iter = 4096;
        bcast_probe<1><<<1, warps * 32>>>(n_iter, d_out); cudaDeviceSynchronize(); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        const double c1 = (double)h[0] / (n_iter * 8);
        bcast_probe<2><<<1, warps * 32>>>(n_iter, d_out); cudaDeviceSynchronize(); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        const double c2 = (double)h[0] / (n_iter * 8);
        bcast_probe<4><<<1, warps * 32>>>(n_iter, d_out); cudaDeviceSynchronize(); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        const double c4 = (double)h[0] / (n_iter * 8);
        printf("warps %2d : LDS.32 %.2f  LDS.64 %.2f  LDS.128 %.2f clk per instruction per warp (x%d warps sharing the pipe)\n", warps, c1, c2, c4, warps);
    }
    return 0;
}
