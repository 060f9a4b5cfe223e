// Row-wise pieces of the relation encoder that are not GEMMs: LayerNorm (PreNorm, model_veto.py:125-132)
// and the 19-token multi-head attention core (Attention.forward, model_veto.py:86-96).  Both are
// HBM/L2-streaming kernels: one warp per row (LayerNorm) or per (sequence, head) (attention), warp
// shuffles for the reductions, vectorised coalesced loads and stores.
#include "common.cuh"

namespace veto {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- LayerNorm
// one warp per row of 576: 9 float2 per lane, two-pass mean / variance in registers (eps 1e-5)
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
                 int64_t rows, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    constexpr int PER = kDim / 64;  // 9
    float2 wv[PER], bv[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        wv[j] = __ldg((const float2*)w + lane + 32 * j);
        bv[j] = __ldg((const float2*)b + lane + 32 * j);
    }
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const float2* xr = (const float2*)(x + row * ldx);
        float2 v[PER];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = xr[lane + 32 * j];
            s += v[j].x + v[j].y;
        }
        const float mean = warp_sum(s) * (1.f / kDim);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const float dx = v[j].x - mean, dy = v[j].y - mean;
            q += dx * dx + dy * dy;
        }
        const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kDim) + 1e-5f);
        const size_t o = (size_t)row * kDim;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            float2 y;
            y.x = (v[j].x - mean) * rstd * wv[j].x + bv[j].x;
            y.y = (v[j].y - mean) * rstd * wv[j].y + bv[j].y;
            const size_t e = o + 2 * (lane + 32 * j);
            if (out_f32) *(float2*)(out_f32 + e) = y;
            if (out_hi) {
                __nv_bfloat16 h0, h1, l0, l1;
                split_bf16(y.x, h0, l0);
                split_bf16(y.y, h1, l1);
                *(uint32_t*)(out_hi + e) = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                if (out_lo)
                    *(uint32_t*)(out_lo + e) = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
        }
    }
}

// ---------------------------------------------------------------- attention
// One CTA (6 warps) per sequence, warp h = head h.  K and V of the head are staged in shared memory
// (broadcast reads), lane i < 19 owns query row i: its q row, its 19 scores and its 96-wide output
// live in registers, so the softmax needs no cross-lane traffic at all.
constexpr int ATT_THREADS = kHeads * 32;
constexpr int ATT_SMEM = 2 * kHeads * kTokens * kHeadDim * (int)sizeof(float);

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const float* __restrict__ qkv, int64_t n_seq, float* out_f32, __nv_bfloat16* out_hi,
                 __nv_bfloat16* out_lo) {
    extern __shared__ float4 att_smem[];  // K then V: [kHeads][kTokens][kHeadDim] fp32 each (87.5 KB)
    float (*sK)[kTokens][kHeadDim] = reinterpret_cast<float (*)[kTokens][kHeadDim]>(att_smem);
    float (*sV)[kTokens][kHeadDim] = sK + kHeads;
    const int h = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const float scale = 0.10206207261596575f;  // 96 ** -0.5 (model_veto.py:74), rounded to fp32 like the reference
    constexpr int LD = 3 * kDim;               // 1728
    constexpr int V4 = kHeadDim / 4;           // 24 float4 per row

    for (int64_t seq = blockIdx.x; seq < n_seq; seq += gridDim.x) {
        const float* base = qkv + (size_t)seq * kTokens * LD + h * kHeadDim;
        __syncwarp();
        for (int idx = lane; idx < kTokens * V4; idx += 32) {
            const int t = idx / V4, c = idx - t * V4;
            const float4 kk = __ldg((const float4*)(base + (size_t)t * LD + kDim) + c);
            const float4 vv = __ldg((const float4*)(base + (size_t)t * LD + 2 * kDim) + c);
            *((float4*)&sK[h][t][0] + c) = kk;
            *((float4*)&sV[h][t][0] + c) = vv;
        }
        __syncwarp();
        if (lane < kTokens) {
            float s[kTokens];
#pragma unroll
            for (int j = 0; j < kTokens; ++j) s[j] = 0.f;
            const float4* qrow = (const float4*)(base + (size_t)lane * LD);
#pragma unroll 4
            for (int c = 0; c < V4; ++c) {
                const float4 q = __ldg(qrow + c);
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    const float4 k4 = *((const float4*)&sK[h][j][0] + c);
                    s[j] = fmaf(q.x, k4.x, s[j]);
                    s[j] = fmaf(q.y, k4.y, s[j]);
                    s[j] = fmaf(q.z, k4.z, s[j]);
                    s[j] = fmaf(q.w, k4.w, s[j]);
                }
            }
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < kTokens; ++j) {
                s[j] *= scale;
                m = fmaxf(m, s[j]);
            }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < kTokens; ++j) {
                s[j] = expf(s[j] - m);
                sum += s[j];
            }
            const float inv = 1.f / sum;
#pragma unroll
            for (int j = 0; j < kTokens; ++j) s[j] *= inv;

            const size_t o = ((size_t)seq * kTokens + lane) * kDim + h * kHeadDim;
#pragma unroll 2
            for (int c = 0; c < V4; ++c) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    const float4 v4 = *((const float4*)&sV[h][j][0] + c);
                    a.x = fmaf(s[j], v4.x, a.x);
                    a.y = fmaf(s[j], v4.y, a.y);
                    a.z = fmaf(s[j], v4.z, a.z);
                    a.w = fmaf(s[j], v4.w, a.w);
                }
                if (out_f32) *(float4*)(out_f32 + o + 4 * c) = a;
                if (out_hi) {
                    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                    split_bf16(a.x, h0, l0); split_bf16(a.y, h1, l1);
                    split_bf16(a.z, h2, l2); split_bf16(a.w, h3, l3);
                    *(uint2*)(out_hi + o + 4 * c) = pack_bf16x4(h0, h1, h2, h3);
                    if (out_lo) *(uint2*)(out_lo + o + 4 * c) = pack_bf16x4(l0, l1, l2, l3);
                }
            }
        }
    }
}

}  // namespace

int layernorm_rows(const float* x, int64_t ldx, const float* w, const float* b, int64_t rows, const ActOut& out,
                   cudaStream_t s) {
    if (rows <= 0) return VETO_OK;
    VETO_REQUIRE((ldx & 1) == 0, VETO_ERR_ARG, "layernorm: row stride must be even");
    const int64_t blocks_needed = (rows + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    layernorm_kernel<<<grid, 256, 0, s>>>(x, ldx, w, b, rows, out.f32, out.hi, out.lo);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int attention_seq(const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s) {
    if (n_seq <= 0) return VETO_OK;
    static bool attr_set = false;
    if (!attr_set) {
        VETO_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    const int64_t cap = (int64_t)num_sms() * 2;
    const int grid = (int)(n_seq < cap ? n_seq : cap);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, s>>>(qkv, n_seq, out.f32, out.hi, out.lo);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
