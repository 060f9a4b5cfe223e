// Relation-token construction: the [R,19,576] encoder input of Transformer.forward (model_veto.py:52-64)
// from the box-level projections (box_stage.cu).  Per pair (s,o):
//   row 0      cls_token + pos_embedding
//   rows 1..16 cat(proj_d(depth patches), proj_v(rgb patches)) + pos_embedding, with
//              proj(cat(s,o)) = S[s] + O[o]                      (model_veto.py:105-113, call order
//                                                                 roi_relation_predictors.py:4124)
//   row 17     ReLU(location_projection(cat(pos[s],pos[o]))) + pos_embedding   (:4118-4119)
//   row 18     ReLU(class_projection(cat(emb[s],emb[o]))) + pos_embedding      (:4120-4121)
// pos_embedding is one [1,1,576] vector added to every token (model_veto.py:62).  Pure gather + add:
// 144 threads, thread t owns float4 column t of all 19 rows; loads and stores are 128-bit and coalesced.
#include "stages.cuh"

namespace veto {
namespace {

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 relu4(float4 a) {
    return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}

constexpr int kTokThreads = kDim / 4;  // 144
constexpr int kStatParts = kDim / 64;  // 9 partial (sum, sum of squares) per row, as the gemm_tc2 epilogues emit them

// OPS: the rows also go out in operand format (bf16 hi + lo or f16c8) with their LayerNorm statistics partials, so that
// layer 0's to_qkv runs LayerNorm-fused on them like every later Linear (api.cu) — the LayerNorm pass over the fresh
// tokens and, with the residual stream in operand format, their fp32 copy disappear.  x may then be NULL.
template <bool OPS>
__global__ void __launch_bounds__(kTokThreads)
tokens_kernel(TokenSources src, const int32_t* __restrict__ subj, const int32_t* __restrict__ obj, int64_t n_pairs,
              float* __restrict__ x, ActOut xo, float2* __restrict__ parts) {
    // per-thread (sum, sum of squares) of its four columns, reduced per 64 columns after the pair's rows are out: 4.5 ms per
    // step; 16-lane butterflies per row (no shared memory, no block barriers) measured 5.6 ms
    __shared__ float2 sq[OPS ? kTokens * kTokThreads : 1];
    const int t = threadIdx.x;
    const float4 pos = __ldg((const float4*)src.pos + t);
    const float4 clspos = __ldg((const float4*)src.clspos + t);
    const bool depth_part = t < kDimDepth / 4;
    const int64_t M = n_pairs * kTokens;
    for (int64_t r = blockIdx.x; r < n_pairs; r += gridDim.x) {
        const int s = subj[r], o = obj[r];
        float4* xr = x ? (float4*)(x + (size_t)r * kTokens * kDim) + t : nullptr;
        auto emit = [&](int row, float4 v) {
            if (xr) xr[(size_t)row * kTokThreads] = v;
            if constexpr (OPS) {
                const size_t off = ((size_t)r * kTokens + row) * kDim + 4 * t;
                if (xo.fmt == FMT_F16C8) {
                    store_act4_f16c8(xo.hi, xo.lo, off, v);
                } else {
                    uint2 hh, ll;
                    split_pair(v.x, v.y, hh.x, ll.x);
                    split_pair(v.z, v.w, hh.y, ll.y);
                    *(uint2*)(xo.hi + off) = hh;
                    *(uint2*)(xo.lo + off) = ll;
                }
                sq[row * kTokThreads + t] = make_float2((v.x + v.y) + (v.z + v.w), (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
            }
        };
        emit(0, clspos);
        if (depth_part) {
            const float4* ps = (const float4*)(src.so_d + (size_t)s * kPatches * 2 * kDimDepth) + t;
            const float4* po = (const float4*)(src.so_d + (size_t)o * kPatches * 2 * kDimDepth + kDimDepth) + t;
#pragma unroll 4
            for (int p = 0; p < kPatches; ++p)
                emit(1 + p, add4(add4(__ldg(ps + (size_t)p * (2 * kDimDepth / 4)), __ldg(po + (size_t)p * (2 * kDimDepth / 4))), pos));
        } else {
            const int tv = t - kDimDepth / 4;
            const float4* ps = (const float4*)(src.so_v + (size_t)s * kPatches * 2 * kDimRgb) + tv;
            const float4* po = (const float4*)(src.so_v + (size_t)o * kPatches * 2 * kDimRgb + kDimRgb) + tv;
#pragma unroll 4
            for (int p = 0; p < kPatches; ++p)
                emit(1 + p, add4(add4(__ldg(ps + (size_t)p * (2 * kDimRgb / 4)), __ldg(po + (size_t)p * (2 * kDimRgb / 4))), pos));
        }
        const float4 ls = __ldg((const float4*)(src.lso + (size_t)s * 2 * kDim) + t);
        const float4 lo = __ldg((const float4*)(src.lso + (size_t)o * 2 * kDim + kDim) + t);
        emit(17, add4(relu4(add4(ls, lo)), pos));
        const float4 cs = __ldg((const float4*)(src.cso + (size_t)s * 2 * kDim) + t);
        const float4 co = __ldg((const float4*)(src.cso + (size_t)o * 2 * kDim + kDim) + t);
        emit(18, add4(relu4(add4(cs, co)), pos));
        if constexpr (OPS) {
            // 19 rows x 9 partials, each the sum over 16 threads' (= 64 columns') contributions
            __syncthreads();
            for (int e = t; e < kTokens * kStatParts; e += kTokThreads) {
                const int row = e / kStatParts, part = e - row * kStatParts;
                const float2* q = sq + row * kTokThreads + part * 16;
                float sv = 0.f, qv = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    sv += q[i].x;
                    qv += q[i].y;
                }
                parts[(size_t)part * M + (size_t)r * kTokens + row] = make_float2(sv, qv);
            }
            __syncthreads();
        }
    }
}

// rel_dists += freq_bias.index_with_labels(stack(obj_pred[s], obj_pred[o])) (model_motifs.py:29-38;
// roi_relation_predictors.py:1134-1135 in the heads that use it) — optional, OFF for VETO parity.
__global__ void freq_bias_kernel(float* __restrict__ logits, int num_out, const float* __restrict__ table,
                                 const int64_t* __restrict__ labels, int num_obj, const int32_t* __restrict__ subj,
                                 const int32_t* __restrict__ obj, int64_t n_pairs) {
    const int64_t total = n_pairs * num_out;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / num_out;
        const int c = (int)(e - r * num_out);
        const int64_t row = labels[subj[r]] * num_obj + labels[obj[r]];
        logits[e] += __ldg(table + row * num_out + c);
    }
}

}  // namespace

int build_tokens(const TokenSources& src, const int32_t* subj, const int32_t* obj, int64_t n_pairs, float* x, ActOut xo,
                 float2* stats_partials, cudaStream_t s) {
    if (n_pairs <= 0) return VETO_OK;
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(n_pairs < cap ? n_pairs : cap);
    if (xo.hi) {
        VETO_REQUIRE(xo.lo && stats_partials, VETO_ERR_ARG, "build_tokens: operand-format rows need both arrays and the statistics buffer");
        tokens_kernel<true><<<grid, kTokThreads, 0, s>>>(src, subj, obj, n_pairs, x, xo, stats_partials);
    } else {
        VETO_REQUIRE(x != nullptr, VETO_ERR_ARG, "build_tokens: no output");
        tokens_kernel<false><<<grid, kTokThreads, 0, s>>>(src, subj, obj, n_pairs, x, xo, nullptr);
    }
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int add_freq_bias(float* logits, int num_out, const float* table, const int64_t* labels, int num_obj,
                  const int32_t* subj, const int32_t* obj, int64_t n_pairs, cudaStream_t s) {
    if (n_pairs <= 0) return VETO_OK;
    const int64_t blocks = (n_pairs * num_out + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    freq_bias_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(logits, num_out, table, labels, num_obj, subj, obj,
                                                                         n_pairs);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
