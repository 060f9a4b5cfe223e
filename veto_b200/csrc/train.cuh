// Training-branch stage kernels (internal; train.cu).  See include/veto_b200.h veto_relation_train_step.
#pragma once
#include "common.cuh"

namespace veto {

// an activation array in any storage format (fp32, or bf16 hi [+ lo])
struct ActIn {
    const float* f32 = nullptr;
    const __nv_bfloat16* hi = nullptr;
    const __nv_bfloat16* lo = nullptr;
};

constexpr int kColsumMaxChunks = 256;

// nn.CrossEntropyLoss(weight) with mean reduction (roi_relation_predictors.py:4070,4134-4135): scalar loss and
// d loss / d logits, over columns [col0, col0 + C) of a matrix with row stride ld; rows whose label is negative are
// outside the loss (zero gradient).  row_scratch: >= 2*rows + 1 floats.
int ce_loss_grad(const float* logits, int ld, int col0, int C, const int64_t* labels, const float* weight, int64_t rows,
                 float* row_scratch, float* loss_out, float* dlogits, cudaStream_t s);
// out[c] (+)= sum_r src[r, c]; deterministic two-stage sum.  scratch: >= colsum_scratch_floats(cols)
size_t colsum_scratch_floats(int cols);
int colsum(const ActIn& src, int64_t ld, int64_t rows, int cols, float* scratch, float* out, bool accumulate, cudaStream_t s);
// t_out[c, r] = f(src[r, c]) with f = optional GELU then optional dropout (element index r * drop_ld + c); columns
// rows..rows_pad-1 of t_out are zero.  rm_out (optional formats) receives f(src) row-major with row stride ld_rm.
int transpose_f32(const float* src, int64_t ld_src, int64_t rows, int cols, bool gelu, const DropSpec& drop, int64_t drop_ld,
                  const ActOut& t_out, int64_t ld_dst, int64_t rows_pad, const ActOut& rm_out, int64_t ld_rm, cudaStream_t s);
// dst[c, r] = src[r, c] for up to kMaxTransposeJobs dense fp32 matrices in ONE launch (the transposed encoder weights of the
// input-gradient GEMMs: 4 per layer, re-derived every step)
constexpr int kMaxTransposeJobs = 4 * VETO_MAX_LAYERS;
struct TransposeJob {
    const float* src;  // [rows, cols] dense
    ActOut dst;        // [cols, rows] dense, any subset of f32 / hi / lo
    int rows, cols;
};
int transpose_f32_multi(const TransposeJob* jobs, int count, cudaStream_t s);
int transpose_bf16(const __nv_bfloat16* src, int64_t ld_src, int64_t rows, int cols, __nv_bfloat16* dst, int64_t ld_dst,
                   int64_t rows_pad, cudaStream_t s);
int splitk_reduce(const float* partial, int slices, size_t n, size_t stride, float* out, cudaStream_t s);
int dropout_inplace(float* x, size_t n, const DropSpec& drop, cudaStream_t s);
// fp32 [n] -> activation storage format, optionally times the dropout keep-scale of element index e (n % 4 == 0)
// fp32 -> activation storage format; act != ACT_NONE first applies the activation (re-computation of a saved pre-activation)
int convert_act(const float* src, size_t n, const DropSpec& drop, const ActOut& out, cudaStream_t s, int act = ACT_NONE);

// LayerNorm backward over rows of 576: dx = dres + dLN(dy) (dx may alias dres or dy), g_gamma / g_beta overwritten.
// Optionally also writes op_out = dropout(dx) in the operand format of the next backward GEMMs and its column sums
// g_op_colsum[576] (the bias gradient of the Linear below).  partial: >= ln_bwd_blocks(rows) * 1728 floats;
// colsum_scratch: >= colsum_scratch_floats(1728) + 1728 floats.
int ln_bwd_blocks(int64_t rows);
int layernorm_bwd(const float* x, int64_t ldx, const float* dy, const float* gamma, const float* dres, float* dx, int64_t rows,
                  float* partial, float* colsum_scratch, float* g_gamma, float* g_beta, cudaStream_t s,
                  const DropSpec& drop = DropSpec(), const ActOut& op_out = ActOut(), float* g_op_colsum = nullptr);
// attention core backward: qkv fp32 [n_seq*19, 1728] (saved), d_out fp32 [n_seq*19, 576] -> d_qkv [n_seq*19, 1728]
int attention_bwd(const float* qkv, const float* d_out, int64_t n_seq, const ActOut& d_qkv, cudaStream_t s);
// token gather backward: dx [R,19,576] -> d_so_d [N*16,1024], d_so_v [N*16,128], d_lso / d_cso [N,1152]
int tokens_bwd(const float* dx, const int32_t* subj, const int32_t* obj, const int32_t* rel_offsets, const int32_t* box_offsets,
               int n_images, int n_boxes, const float* lso, const float* cso, float* d_so_d, float* d_so_v, float* d_lso,
               float* d_cso, cudaStream_t s);
// BatchNorm1d(4) batch statistics of the box geometry: stats[0..3] mean, [4..7] biased variance; running stats updated
int bn_batch_stats(const float* boxes, int n_boxes, float momentum, float* stats, float* running_mean, float* running_var,
                   cudaStream_t s);
// scratch: >= pos_embed_bwd_scratch_floats(); colsum_scratch: >= colsum_scratch_floats(648)
constexpr int kPosBwdBlocks = 64;
constexpr int kPosBwdCols = 5 * kPosDim + 8;  // 648
size_t pos_embed_bwd_scratch_floats();
int pos_embed_bwd(const float* boxes, int n_boxes, const float* stats, const veto_weights& w, const float* pos_out,
                  const float* d_pos, float drop_scale, float* scratch, float* colsum_scratch, float* g_pos_w, float* g_pos_b,
                  float* g_bn_w, float* g_bn_b, cudaStream_t s);
int embed_bwd(const float* d_emb, const int64_t* labels, const float* obj_logits, int num_obj, int n_boxes, float* g_embed,
              cudaStream_t s);
int unpatchify(const float* d_patch, int n_boxes, float* d_roi, cudaStream_t s);
int unpack_halves(const float* g_packed, float* g_src, int out, int in, cudaStream_t s);
int unpack_patch(const float* g_packed, float* g_src, int out, cudaStream_t s);
int copy_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int cols, cudaStream_t s);

}  // namespace veto
