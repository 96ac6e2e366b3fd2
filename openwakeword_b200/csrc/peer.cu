// Multi-GPU result gather without a collective call ("optionally fused"): the last kernel of a step
// (heads phase of tc_inc_kernel, or heads_kernel) stores its scores straight into a buffer that lives on the gathering
// rank's GPU - d_scores of oww_step is then a peer-mapped address and the stores travel over NVLink - and a one-thread
// kernel publishes a step counter behind them.  This file holds the plumbing for that: allocations that can be shared
// between the one-process-per-GPU ranks (CUDA IPC), and the signal / wait kernels.  No data-path NCCL call remains.
#include "oww_internal.h"
#include <algorithm>
#include <cstring>

namespace {

__global__ void peer_signal_kernel(unsigned long long* flag, unsigned long long value) {
    // stream order put this kernel after the step's last kernel: its stores are performed; make them visible system-wide
    // before the counter (release at system scope), so a peer that acquires the counter sees the scores
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

// block copy with the SMs: coalesced 4-byte words (16-byte vectors where both ends allow it); the destination may be a
// peer mapping - the stores then cross NVLink as full lines.  (cudaMemcpyAsync on an IPC mapping took 20 ms per call here.)
__global__ void __launch_bounds__(256) peer_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n_words) {
    const size_t stride = (size_t)gridDim.x * 256;
    if (((((size_t)dst) | ((size_t)src)) & 15) == 0) {
        const size_t n4 = n_words / 4;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride)
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
        for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n_words; i += stride) dst[i] = src[i];
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_words; i += stride) dst[i] = src[i];
    }
}

__global__ void peer_wait_kernel(const unsigned long long* flags, int n, int stride, unsigned long long value,
                                 long long timeout_cycles, int* timed_out) {
    // one lane per flag; every lane spins until its counter reaches `value`
    const int i = threadIdx.x;
    if (i >= n) return;
    const unsigned long long* f = flags + (size_t)i * stride;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= value) break;
        // a dead peer must not hang the stream - and must not poison the context either: raise the handle's flag
        // (oww_peer_status) and let the stream go on; the consumer's data for this step is then stale
        if (clock64() - t0 > timeout_cycles) { atomicExch(timed_out, 1); break; }
        __nanosleep(200);
    }
}

}  // namespace

extern "C" {

int oww_peer_alloc(oww_ctx* ctx, size_t bytes, void** d_ptr, unsigned char handle_out[64]) {
    if (!ctx || !d_ptr || !handle_out || bytes == 0) return oww_fail(ctx, OWW_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    OWW_CUDA(ctx, cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        return oww_fail(ctx, OWW_ECUDA, "peer allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, 64);
    *d_ptr = p;
    return OWW_OK;
}

int oww_peer_free(oww_ctx* ctx, void* d_ptr) {
    if (!ctx) return OWW_EINVAL;
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_ptr) OWW_CUDA(ctx, cudaFree(d_ptr));
    return OWW_OK;
}

int oww_peer_open(oww_ctx* ctx, const unsigned char handle[64], void** d_ptr) {
    if (!ctx || !handle || !d_ptr) return oww_fail(ctx, OWW_EINVAL, "null argument");
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    OWW_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return OWW_OK;
}

int oww_peer_close(oww_ctx* ctx, void* d_ptr) {
    if (!ctx) return OWW_EINVAL;
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_ptr) OWW_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return OWW_OK;
}

int oww_peer_copy(oww_ctx* ctx, void* d_dst, const void* d_src, size_t bytes, void* stream) {
    if (!ctx || !d_dst || !d_src) return oww_fail(ctx, OWW_EINVAL, "null argument");
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bytes % 4 || ((size_t)d_dst & 3) || ((size_t)d_src & 3)) return oww_fail(ctx, OWW_EINVAL, "oww_peer_copy moves 4-byte words");
    const size_t n_words = bytes / 4;
    if (n_words == 0) return OWW_OK;
    const unsigned grid = (unsigned)std::min<size_t>((n_words / 4 + 255) / 256 + 1, (size_t)ctx->sm_count * 4);
    peer_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<uint32_t*>(d_dst), static_cast<const uint32_t*>(d_src), n_words);
    OWW_LAUNCH_CHECK(ctx);
    return OWW_OK;
}

int oww_peer_signal(oww_ctx* ctx, uint64_t* d_flag, uint64_t value, void* stream) {
    if (!ctx || !d_flag) return oww_fail(ctx, OWW_EINVAL, "null argument");
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(d_flag), value);
    OWW_LAUNCH_CHECK(ctx);
    return OWW_OK;
}

int oww_peer_wait(oww_ctx* ctx, const uint64_t* d_flags, int n, int stride, uint64_t value, double timeout_s, void* stream) {
    if (!ctx || !d_flags) return oww_fail(ctx, OWW_EINVAL, "null argument");
    if (n < 1 || n > 1024 || stride < 1) return oww_fail(ctx, OWW_EINVAL, "n=%d stride=%d", n, stride);
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    if (khz <= 0) khz = 1900000;
    if (!(timeout_s > 0)) timeout_s = 10.0;
    const long long cycles = (long long)(timeout_s * 1e3 * khz);
    if (!ctx->d_peer_err) {
        OWW_CUDA(ctx, cudaMalloc(&ctx->d_peer_err, sizeof(int)));
        OWW_CUDA(ctx, cudaMemset(ctx->d_peer_err, 0, sizeof(int)));
    }
    peer_wait_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(d_flags),
                                                                            n, stride, value, cycles, ctx->d_peer_err);
    OWW_LAUNCH_CHECK(ctx);
    return OWW_OK;
}

int oww_peer_status(oww_ctx* ctx, int* timed_out) {
    if (!ctx || !timed_out) return oww_fail(ctx, OWW_EINVAL, "null argument");
    *timed_out = 0;
    if (!ctx->d_peer_err) return OWW_OK;
    OWW_CUDA(ctx, cudaSetDevice(ctx->device));
    OWW_CUDA(ctx, cudaMemcpy(timed_out, ctx->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost));   // synchronises
    if (*timed_out) OWW_CUDA(ctx, cudaMemset(ctx->d_peer_err, 0, sizeof(int)));
    return OWW_OK;
}

}  // extern "C"
