// Stage kernels of the relation head (internal).
#pragma once
#include "common.cuh"

namespace veto {

// ---- weight packing (pack.cu) ----
// dst[o + h*out, i] = src[o, h*in + i]: splits Linear(2*in -> out) over cat(subject, object) inputs into
// a [2*out, in] matrix whose first `out` rows act on the subject and last `out` rows on the object.
int pack_halves(const float* src, float* dst, int out, int in, cudaStream_t s);
// patch projection: dst[o + h*out, p*256 + c] = src[o, p*512 + h*256 + c] (model_veto.py:109-113 layout
// 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' with c over cat(subject 256, object 256))
int pack_patch(const float* src, float* dst, int out, cudaStream_t s);
int pack_bias2(const float* b, float* dst, int out, cudaStream_t s);          // [b, 0]
int pack_add(const float* a, const float* b, float* dst, int n, cudaStream_t s);
int pack_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t s);
// the same for up to kMaxSplitJobs arrays in ONE launch (the per-step re-split of the weights was 26 tiny launches)
constexpr int kMaxSplitJobs = 2 + 6 * VETO_MAX_LAYERS;   // + the LayerNorm-folded to_qkv / FF1 weights
struct SplitJob {
    const float* src;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    size_t n;
    int fmt = FMT_BF16;   // FMT_F16C8: the WEIGHT side of the f16c8 format (common.cuh): hi = fp16(2048 w), lo = (value, residual) e4m3 bytes
    const float* col_scale = nullptr;   // optional [row_len]: element (r, k) is multiplied by col_scale[k] first (LayerNorm weight
    int row_len = 0;                    //   folded into a Linear weight, gemm_tc2.cu EPI_*_LN)
};
// c1[n] = sum_k gamma[k] W[n,k],  c2[n] = sum_k beta[k] W[n,k] (+ bias[n]): the constants of a LayerNorm-fused Linear
int ln_fold_consts(const float* W, const float* gamma, const float* beta, const float* bias, int N, int K, float* c1, float* c2,
                   cudaStream_t s);
int pack_split_bf16_multi(const SplitJob* jobs, int count, cudaStream_t s);

// ---- box stage (box_stage.cu) ----
// pos_embed (BN1d eval -> Linear(4,128) -> ReLU, roi_relation_predictors.py:4042-4047,4097-4102) and the class
// embedding (hard lookup :4087 / soft softmax@W :4095) of every box.
int box_embed(const float* boxes, const int64_t* labels, const float* obj_logits, int num_obj, int n_boxes,
              const veto_weights& w, float* pos_out, float* emb_out, cudaStream_t s, const float* batch_stats = nullptr,
              const DropSpec& pos_drop = DropSpec());
// roi [N,256,8,8] -> patch rows [N*16, 1024] in (p1 p2 c) order
int patchify(const float* roi, int n_boxes, const ActOut& out, cudaStream_t s);

// ---- token build (tokens.cu) ----
struct TokenSources {
    const float* so_d;    // [N*16, 1024]: cols [0,512) subject part (+bias), [512,1024) object part
    const float* so_v;    // [N*16, 128]
    const float* lso;     // [N, 1152]: location projection subject(+bias) | object
    const float* cso;     // [N, 1152]: class projection
    const float* clspos;  // [576] cls_token + pos_embedding
    const float* pos;     // [576] pos_embedding
};
// x (fp32, may be NULL when xo is given) and / or xo: the rows in operand format + their LayerNorm statistics partials
// [kDim / 64][n_pairs * 19] for ln_stats_finalize
int build_tokens(const TokenSources& src, const int32_t* subj, const int32_t* obj, int64_t n_pairs, float* x, ActOut xo,
                 float2* stats_partials, cudaStream_t s);
int add_freq_bias(float* logits, int num_out, const float* table, const int64_t* labels, int num_obj,
                  const int32_t* subj, const int32_t* obj, int64_t n_pairs, cudaStream_t s);

}  // namespace veto
