// K3 (tensor-core path): the wake-word heads' wide first layer as a tcgen05 GEMM, the small later layers on CUDA cores.
//
// Same graphs as heads.cu (reference: <head>.onnx sessions, /root/reference/openwakeword/model.py:137-138,153-159,287-302;
// family /root/reference/openwakeword/train.py:56-83,144-165).  heads.cu tiles 8 or 32 streams per CTA and streams the
// first-layer matrix W1 (n_in*96 x D1, 0.4-1.7 MB) through shared memory once per tile: at 8192 streams x 6 heads that is
// 0.78 ms per step, 7 % of its HBM roofline.  Here a CTA owns 128 streams of one head:
//   D[128 streams x D1] (fp32, TMEM) = sum over the n_in feature rows c of  X_c[128 x 96] * W1_c[96 x D1]
// X_c is gathered from the per-stream feature rings (fp32) by 16 converter warps, split into fp16 hi + lo parts
// (x*4 = hi + lo, |lo| <= ulp(hi)/2) and written to shared memory in the no-swizzle K-major UMMA core-matrix order
// ([k-octet][128 rows][16 B]; LBO = 2048 B, SBO = 128 B); W1 is pre-packed on the host as fp16 hi + lo of W*2^s in the
// same order and streamed by one cp.async.bulk per feature row.  Three MMA terms per K step (hi*hi + lo*hi + hi*lo,
// fp32 accumulate) reproduce the fp32 product to ~2^-21 relative, so the scores stay within 2e-5 of heads.cu / the
// oracle; n_terms = 1 or 2 trades that for speed.  After the mainloop the 128 accumulator rows are read back with
// tcgen05.ld (one row per thread), scaled by 2^-(s+2), and the rest of the head (bias, [LayerNorm], ReLU, the small
// Linear layers, sigmoid / softmax) runs on CUDA cores out of shared memory with the same arithmetic as heads.cu.
// Warp roles: 0-15 converters (then epilogue / later layers), 16 = W1 bulk-copy producer, 17 = TMEM owner + MMA issuer.
#include "oww_internal.h"
#include "tc_common.cuh"
#include <cmath>
#include <cstring>

namespace {

constexpr int kHtWorkers = 16;                       // converter / epilogue warps
constexpr int kHtThreads = (kHtWorkers + 2) * 32;    // 576
constexpr int kHtTile = 128;                         // streams per CTA = MMA M
constexpr int kHtMaxStages = 3;
constexpr int kHtAPlane = kHtTile * 16;              // bytes per k-octet plane of an A tile (LBO)
constexpr int kHtABytes = 12 * kHtAPlane;            // one 128 x 96 fp16 tile: 24 KB
constexpr float kHtXScale = 4.0f;                    // features are scaled by 2^2 before the hi/lo split (keeps lo normal)
constexpr int kHtHPitch = 129;                       // fp32 row pitch of the hidden-activation buffers (layers <= 128 wide; conflict-free per-row access)

__device__ __forceinline__ bool ht_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ht_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct HeadsTcArgs {
    HeadDev head[16];
    const uint8_t* w1[16];        // packed first layer: per feature row c: [term hi|lo][octet 12][NP][8] fp16
    int np[16];                   // D1 rounded up to a multiple of 16 (MMA N)
    float unscale[16];            // 2^-(s+2)
    FeatSrc src;
    int n; float* out; int out_stride; int combine_max;
    int n_terms;                  // 1: hi*hi   2: + lo*hi   3: + hi*lo (default)
    int stages, stage_bytes;
};

// Where sample s's window starts: row index r0 of feature row 0 (rows r0 + c, c = 0..n_in-1; negative = not available ->
// zeros), the row mask (ring) or -1 (linear), and the sample's base pointer.
struct HtRows { const float* base; int r0, mask; };
__device__ __forceinline__ HtRows ht_rows(const FeatSrc& src, int n_in, int s, int n) {
    HtRows w{nullptr, 0, -1};
    if (s >= n) return w;
    if (src.count) {                                       // per-stream ring
        w.base = src.base + (int64_t)s * src.stride;
        w.r0 = src.count[s] - src.back - n_in; w.mask = src.rows_mask;
    } else if (src.steps > 0) {                            // sliding windows over per-clip linear feature rows (bulk path)
        const int clip = s / src.steps, st = s - clip * src.steps;
        w.base = src.base + (int64_t)clip * src.stride;
        w.r0 = src.row0 + st + 1 - n_in;
    } else {                                               // linear [n][n_in][96]
        w.base = src.base + (int64_t)s * src.stride;
    }
    return w;
}

__global__ void __launch_bounds__(kHtThreads, 1) heads_tc_kernel(const __grid_constant__ HeadsTcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int hid = blockIdx.y;
    const HeadDev& H = a.head[hid];
    const int NP = a.np[hid];
    const int n_in = H.n_in;
    const int s0 = blockIdx.x * kHtTile;
    const int S = a.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    uint8_t* stage0 = smem + 1024;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto w_full = [&](int s) { return bar0 + 8u * (kHtMaxStages + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * kHtMaxStages + s); };
    const uint32_t acc_full = bar0 + 8u * (3 * kHtMaxStages);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kHtMaxStages; ++s) { mbar_init(a_full(s), kHtWorkers); mbar_init(w_full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kHtWorkers + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t w_term_bytes = 12u * (uint32_t)NP * 16u;

    if (warp == kHtWorkers) {
        // ===================== W1 producer: one bulk copy per feature row =====================
        if (lane == 0) {
            const uint32_t bytes = w_term_bytes * (a.n_terms >= 3 ? 2u : 1u);
            for (int c = 0; c < n_in; ++c) {
                const int s = c % S;
                mbar_wait(empty(s), (((uint32_t)(c / S)) & 1u) ^ 1u);
                mbar_expect_tx(w_full(s), bytes);
                bulk_g2s(smem_u32(stage0 + s * a.stage_bytes + 2 * kHtABytes), a.w1[hid] + (size_t)c * 2u * w_term_bytes, bytes, w_full(s));
            }
        }
    } else if (warp == kHtWorkers + 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t desc_hi = (uint32_t)(make_desc(0, 0u, 128u) >> 32);                 // SBO + version
        const uint32_t a_lo_fields = (uint32_t)make_desc(0, (uint32_t)kHtAPlane, 128u);    // LBO in the low word
        const uint32_t b_lo_fields = (uint32_t)make_desc(0, (uint32_t)NP * 16u, 128u);
        uint32_t accumulate = 0;
        for (int c = 0; c < n_in; ++c) {
            const int s = c % S;
            const uint32_t par = ((uint32_t)(c / S)) & 1u;
            mbar_wait(a_full(s), par);
            mbar_wait(w_full(s), par);
            tc_fence_after();
            const uint32_t st_addr = smem_u32(stage0 + s * a.stage_bytes);
            const uint32_t a_unit[2] = {st_addr >> 4, (st_addr + kHtABytes) >> 4};
            const uint32_t w_unit[2] = {(st_addr + 2 * kHtABytes) >> 4, (st_addr + 2 * kHtABytes + w_term_bytes) >> 4};
            if (ht_elect_one()) {
                for (int t = 0; t < a.n_terms; ++t) {
                    const uint32_t au = a_unit[t == 1 ? 1 : 0];       // term 1 = x_lo * w_hi
                    const uint32_t wu = w_unit[t == 2 ? 1 : 0];       // term 2 = x_hi * w_lo
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const uint32_t alo = a_lo_fields | ((au + (uint32_t)(2 * q) * (kHtAPlane >> 4)) & 0x3FFFu);
                        const uint32_t blo = b_lo_fields | ((wu + (uint32_t)(2 * q * NP)) & 0x3FFFu);
                        tc_mma_f16(tmem_base, ((uint64_t)desc_hi << 32) | alo, ((uint64_t)desc_hi << 32) | blo, idesc, accumulate);
                        accumulate = 1;
                    }
                }
                tc_commit(empty(s));                                  // stage free once these MMAs have read it
                if (c == n_in - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ===================== converters: fp32 ring rows -> fp16 hi/lo A tiles =====================
        // thread = (row 8*warp + lane%8, octet quad lane/8): three octets (32 B of fp32 each) per feature row
        const int row = 8 * warp + (lane & 7), jq = lane >> 3;
        const int s = s0 + row;
        float4 buf[2][6];
        const HtRows rows_of = ht_rows(a.src, n_in, s, a.n);
        auto load = [&](int c, float4* v) {
            const float* p = nullptr;
            if (c < n_in && rows_of.base) {
                const int r = rows_of.r0 + c;
                if (r >= 0) p = rows_of.base + (int64_t)(rows_of.mask >= 0 ? (r & rows_of.mask) : r) * 96;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (p) {
                    const float4* q = reinterpret_cast<const float4*>(p + (4 * i + jq) * 8);
                    v[2 * i] = __ldcg(q); v[2 * i + 1] = __ldcg(q + 1);
                } else {
                    v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f); v[2 * i + 1] = v[2 * i];
                }
            }
        };
        auto convert_store = [&](int c, const float4* v) {
            const int st = c % S;
            mbar_wait(empty(st), (((uint32_t)(c / S)) & 1u) ^ 1u);
            uint8_t* A = stage0 + st * a.stage_bytes;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float x[8] = {v[2 * i].x, v[2 * i].y, v[2 * i].z, v[2 * i].w, v[2 * i + 1].x, v[2 * i + 1].y, v[2 * i + 1].z, v[2 * i + 1].w};
                __half2 hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x0 = x[2 * e] * kHtXScale, x1 = x[2 * e + 1] * kHtXScale;
                    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                    hi[e] = __halves2half2(h0, h1);
                    lo[e] = __floats2half2_rn(x0 - __half2float(h0), x1 - __half2float(h1));
                }
                const int off = (4 * i + jq) * kHtAPlane + row * 16;
                *reinterpret_cast<uint4*>(A + off) = *reinterpret_cast<uint4*>(hi);
                if (a.n_terms >= 2) *reinterpret_cast<uint4*>(A + kHtABytes + off) = *reinterpret_cast<uint4*>(lo);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(st));
        };
        load(0, buf[0]);
        load(1, buf[1]);
        for (int c = 0; c < n_in; c += 2) {
            convert_store(c, buf[0]);
            load(c + 2, buf[0]);
            if (c + 1 < n_in) {
                convert_store(c + 1, buf[1]);
                load(c + 3, buf[1]);
            }
        }

        // ===================== epilogue: accumulators -> hidden activations, later layers on CUDA cores =====================
        // the stage ring is dead once acc_full has fired (every MMA has read its operands): reuse it
        float* hA = reinterpret_cast<float*>(stage0);                    // [128][kHtHPitch]
        float* hB = hA + kHtTile * kHtHPitch;
        float* wS = hB + kHtTile * kHtHPitch;                            // weight rows of a later layer, [rows][Dp]
        const int wS_floats = (int)((227 * 1024 - 1024 - 2 * kHtTile * kHtHPitch * 4) / 4);
        const int tid = threadIdx.x;                                      // 0..511
        const int D1 = H.dims[1];
        mbar_wait(acc_full, 0);
        tc_fence_after();
        if (warp < 4) {
            const int r = warp * 32 + lane;
            const float us = a.unscale[hid];
            const float* bias = H.blob + H.b_off[0];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < NP; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (c0 + e < D1) hA[r * kHtHPitch + c0 + e] = fmaf(__uint_as_float(v[e]), us, __ldg(bias + c0 + e));
            }
            tc_fence_before();
        }
        ht_bar_sync(1, kHtWorkers * 32);
        float* cur = hA; float* nxt = hB;
        for (int l = 0; l < H.n_layers; ++l) {
            const int D = H.dims[l + 1], K = H.dims[l];
            if (l > 0) {
                // out[r][d] = b[d] + sum_k cur[r][k] W[k][d], k ascending (as heads.cu).  thread = (row tid%128, column
                // quarter tid/128); W rows staged through shared memory in chunks, zero-padded to Dp columns.
                const int Dp = (D + 31) & ~31, dq = Dp / 4;
                const int r = tid & 127, cq = tid >> 7;
                const float* W = H.blob + H.w_off[l];
                int rows = wS_floats / Dp;
                if (rows > K) rows = K;
                const int n_ct = dq / 8;                               // register tiles of 8 columns per thread
                for (int ct = 0; ct < n_ct; ++ct) {
                    float acc[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
                    for (int k0 = 0; k0 < K; k0 += rows) {
                        const int kr = min(rows, K - k0);
                        if (ct == 0 || rows < K) {                     // (re)stage this chunk of weight rows
                            ht_bar_sync(1, kHtWorkers * 32);
                            for (int i = tid; i < kr * Dp; i += kHtWorkers * 32) {
                                const int kk = i / Dp, d = i - kk * Dp;
                                wS[i] = d < D ? __ldg(W + (int64_t)(k0 + kk) * D + d) : 0.f;
                            }
                            ht_bar_sync(1, kHtWorkers * 32);
                        }
                        const float* wr = wS + cq * dq + ct * 8;
                        const float* xr = cur + r * kHtHPitch + k0;
                        for (int kk = 0; kk < kr; ++kk) {
                            const float x = xr[kk];
                            const float4 w0 = *reinterpret_cast<const float4*>(wr + kk * Dp);
                            const float4 w1 = *reinterpret_cast<const float4*>(wr + kk * Dp + 4);
                            acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
                            acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
                            acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
                            acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
                        }
                    }
                    const float* bias = H.blob + H.b_off[l];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int d = cq * dq + ct * 8 + e;
                        if (d < D) nxt[r * kHtHPitch + d] = acc[e] + __ldg(bias + d);
                    }
                }
                ht_bar_sync(1, kHtWorkers * 32);
                float* t = cur; cur = nxt; nxt = t;
            }
            if (l < H.n_layers - 1) {
                // [LayerNorm] + ReLU: one thread per row (its row of `cur` is bank-conflict free at pitch 129)
                if (tid < kHtTile) {
                    float* rowp = cur + tid * kHtHPitch;
                    if (H.layernorm) {
                        float sum = 0.f;
                        for (int d = 0; d < D; ++d) sum += rowp[d];
                        const float mu = sum / (float)D;
                        float sq = 0.f;
                        for (int d = 0; d < D; ++d) { const float c = rowp[d] - mu; sq = fmaf(c, c, sq); }
                        const float rstd = 1.0f / sqrtf(sq / (float)D + 1e-5f);
                        const float* g = H.blob + H.g_off[l];
                        const float* hb = H.blob + H.h_off[l];
                        for (int d = 0; d < D; ++d) rowp[d] = fmaxf((rowp[d] - mu) * rstd * __ldg(g + d) + __ldg(hb + d), 0.f);
                    } else {
                        for (int d = 0; d < D; ++d) rowp[d] = fmaxf(rowp[d], 0.f);
                    }
                }
                ht_bar_sync(1, kHtWorkers * 32);
            }
        }
        // final activation + store: one thread per row
        const int n_out = H.dims[H.n_layers];
        if (tid < kHtTile && s0 + tid < a.n) {
            float* rowp = cur + tid * kHtHPitch;
            if (H.final_act == 4) {
                for (int d = 0; d < n_out; ++d) rowp[d] = fmaxf(rowp[d], 0.f);
            } else if (H.final_act == 1) {
                for (int d = 0; d < n_out; ++d) rowp[d] = 1.0f / (1.0f + expf(-rowp[d]));
            } else if (H.final_act == 2 || H.final_act == 3) {
                float m = -INFINITY;
                for (int d = 0; d < n_out; ++d) {
                    if (H.final_act == 3) rowp[d] = fmaxf(rowp[d], 0.f);
                    m = fmaxf(m, rowp[d]);
                }
                float sum = 0.f;
                for (int d = 0; d < n_out; ++d) { rowp[d] = expf(rowp[d] - m); sum += rowp[d]; }
                for (int d = 0; d < n_out; ++d) rowp[d] = rowp[d] / sum;
            }
            float* o = a.out + (int64_t)(s0 + tid) * a.out_stride + H.col0;
            for (int d = 0; d < n_out; ++d) o[d] = a.combine_max ? fmaxf(o[d], rowp[d]) : rowp[d];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kHtWorkers + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}

}  // namespace

// Host side: pack W1 of a head as fp16 hi/lo of W * 2^s in UMMA K-major core-matrix order, one block per feature row.
int oww_heads_tc_pack(oww_ctx* ctx, Head& h, const float* w1 /* [n_in*96][D1] */) {
    const int n_in = h.desc.n_in, D1 = h.desc.dims[1];
    h.tc_ok = false;
    if (D1 > 128) return OWW_OK;                                   // wider first layers stay on heads.cu
    for (int l = 1; l <= h.desc.n_layers; ++l) if (h.desc.dims[l] > 128) return OWW_OK;     // hidden buffers are 128 wide
    const int NP = (D1 + 15) & ~15;
    float amax = 0.f;
    for (size_t i = 0; i < (size_t)n_in * 96 * D1; ++i) amax = std::fmax(amax, std::fabs(w1[i]));
    int s = 0;
    if (amax > 0.f && std::isfinite(amax)) {
        int e; std::frexp(amax, &e);                               // amax = m * 2^e, m in [0.5, 1)
        s = 14 - e;                                                // amax * 2^s in [2^13, 2^14)
        if (s > 24) s = 24;
        if (s < -8) s = -8;
    }
    const float sc = std::ldexp(1.0f, s);
    const size_t term = (size_t)12 * NP * 8;                       // halves per (feature row, term)
    std::vector<__half> packed((size_t)n_in * 2 * term);
    for (int c = 0; c < n_in; ++c)
        for (int j = 0; j < 12; ++j)
            for (int n = 0; n < NP; ++n)
                for (int e = 0; e < 8; ++e) {
                    const int k = c * 96 + j * 8 + e;
                    const float v = n < D1 ? w1[(size_t)k * D1 + n] * sc : 0.f;
                    const __half hi = __float2half_rn(v);
                    const __half lo = __float2half_rn(v - __half2float(hi));
                    const size_t at = ((size_t)c * 2) * term + ((size_t)j * NP + n) * 8 + e;
                    packed[at] = hi;
                    packed[at + term] = lo;
                }
    OWW_CUDA(ctx, cudaMalloc(&h.d_w1_tc, packed.size() * sizeof(__half)));
    OWW_CUDA(ctx, cudaMemcpy(h.d_w1_tc, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
    h.tc_np = NP;
    h.tc_unscale = std::ldexp(1.0f, -s) / kHtXScale;
    h.tc_ok = true;
    return OWW_OK;
}

bool oww_heads_tc_supported(const oww_ctx* ctx, int head_id) {
    if (!ctx->tc_heads || ctx->cfg.cnn_mode == OWW_CNN_FP32_WINDOW) return false;     // mode 0 stays fp32 end to end
    return head_id >= 0 && head_id < (int)ctx->heads.size() && ctx->heads[head_id].tc_ok;
}

// Every head of the handle on the same samples: the tensor-core kernel for the heads it covers, heads.cu for the rest,
// then the conditional verifier pairs (applied to THIS call's scores, i.e. per chunk - model.py runs the whole gated
// graph per chunk and takes the max over chunks afterwards).
namespace {
__global__ void gate_kernel(float* scores, int n, int stride, const Gate* gates, int n_gates) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_gates) return;
    const int s = i / n_gates;
    const Gate g = gates[i - s * n_gates];
    float* o = scores + (int64_t)s * stride;
    if (o[g.main_col] > g.thr) o[g.main_col] = o[g.ver_col];
}
__global__ void max_combine_kernel(float* dst, const float* src, int n, int cols, int dst_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    const int s = i / cols, c = i - s * cols;
    float* d = dst + (int64_t)s * dst_stride + c;
    *d = fmaxf(*d, src[i]);
}
}  // namespace

int oww_heads_all(oww_ctx* ctx, const FeatSrc& src, int n, float* d_out, int out_stride, int combine_max, cudaStream_t s) {
    if (ctx->heads.empty() || n <= 0) return OWW_OK;
    uint32_t tc_mask = 0, cc_mask = 0;
    for (int i = 0; i < (int)ctx->heads.size(); ++i) {
        if (oww_heads_tc_supported(ctx, i)) tc_mask |= 1u << i; else cc_mask |= 1u << i;
    }
    float* out = d_out; int stride = out_stride; int comb = combine_max;
    const bool via_tmp = combine_max && !ctx->gates.empty();       // gate this chunk's raw scores before the max
    if (via_tmp) {
        const size_t need = (size_t)n * ctx->n_out_total;
        if (ctx->scores_tmp_floats < need) {
            cudaFree(ctx->d_scores_tmp); ctx->d_scores_tmp = nullptr; ctx->scores_tmp_floats = 0;
            OWW_CUDA(ctx, cudaMalloc(&ctx->d_scores_tmp, need * sizeof(float)));
            ctx->scores_tmp_floats = need;
        }
        out = ctx->d_scores_tmp; stride = ctx->n_out_total; comb = 0;
    }
    int rc;
    if (tc_mask && (rc = oww_heads_tc_launch(ctx, -1, src, n, out, stride, 0, comb, s, tc_mask))) return rc;
    if (cc_mask && (rc = oww_heads_launch(ctx, -1, src, n, out, stride, 0, comb, s, cc_mask))) return rc;
    if (!ctx->gates.empty()) {
        const int total = n * (int)ctx->gates.size();
        gate_kernel<<<(total + 255) / 256, 256, 0, s>>>(out, n, stride, ctx->d_gates, (int)ctx->gates.size());
        OWW_LAUNCH_CHECK(ctx);
    }
    if (via_tmp) {
        const int total = n * ctx->n_out_total;
        max_combine_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_out, ctx->d_scores_tmp, n, ctx->n_out_total, out_stride);
        OWW_LAUNCH_CHECK(ctx);
    }
    return OWW_OK;
}

// Same contract as oww_heads_launch (heads.cu).
int oww_heads_tc_launch(oww_ctx* ctx, int head_id, const FeatSrc& src, int n, float* d_out, int out_stride,
                        int out_col0, int combine_max, cudaStream_t s, uint32_t head_mask) {
    if (n <= 0) return OWW_OK;
    int sel[16], nh = 0;
    if (head_id >= 0) sel[nh++] = head_id;
    else {
        if (ctx->heads.size() > 16) return oww_fail(ctx, OWW_EUNSUPPORTED, "at most 16 heads per launch");
        for (int i = 0; i < (int)ctx->heads.size(); ++i) if (head_mask >> i & 1u) sel[nh++] = i;
    }
    if (nh == 0) return OWW_OK;
    HeadsTcArgs a;
    std::memset(&a, 0, sizeof(a));
    int np_max = 16;
    for (int i = 0; i < nh; ++i) {
        const Head& h = ctx->heads[sel[i]];
        if (!h.tc_ok) return oww_fail(ctx, OWW_EUNSUPPORTED, "head %d has no tensor-core packing", sel[i]);
        HeadDev& d = a.head[i];
        d.blob = h.d_blob;
        d.n_in = h.desc.n_in; d.n_layers = h.desc.n_layers; d.layernorm = h.desc.layernorm; d.final_act = h.desc.final_act;
        for (int l = 0; l <= h.desc.n_layers; ++l) d.dims[l] = h.desc.dims[l];
        for (int l = 0; l < h.desc.n_layers; ++l) {
            d.w_off[l] = (int)h.w_off[l]; d.b_off[l] = (int)h.b_off[l];
            d.g_off[l] = (int)h.g_off[l]; d.h_off[l] = (int)h.h_off[l];
        }
        d.col0 = (head_id < 0 ? h.col0 : 0) + out_col0;
        a.w1[i] = reinterpret_cast<const uint8_t*>(h.d_w1_tc);
        a.np[i] = h.tc_np;
        a.unscale[i] = h.tc_unscale;
        if (h.tc_np > np_max) np_max = h.tc_np;
    }
    a.src = src; a.n = n; a.out = d_out; a.out_stride = out_stride; a.combine_max = combine_max;
    a.n_terms = ctx->tc_heads_terms;
    a.stage_bytes = 2 * kHtABytes + 2 * 12 * np_max * 16;
    const int kSmem = 227 * 1024;
    a.stages = (kSmem - 1024) / a.stage_bytes;
    if (a.stages > kHtMaxStages) a.stages = kHtMaxStages;
    if (a.stages < 2) return oww_fail(ctx, OWW_EUNSUPPORTED, "tensor-core heads: stage of %d bytes does not fit twice", a.stage_bytes);
    if (!ctx->heads_tc_attr_set) {
        OWW_CUDA(ctx, cudaFuncSetAttribute(heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        ctx->heads_tc_attr_set = true;
    }
    dim3 grid((n + kHtTile - 1) / kHtTile, nh);
    heads_tc_kernel<<<grid, kHtThreads, kSmem, s>>>(a);
    OWW_LAUNCH_CHECK(ctx);
    return OWW_OK;
}
