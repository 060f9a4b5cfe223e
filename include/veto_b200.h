/*
 * veto_b200.h — C ABI of libveto_b200.so: the B200 (sm_100a) implementation of the VETO
 * relation-prediction hot path of visinf/veto (reference tree paths are relative to the reference
 * repository root; see SURVEY.md §8 for the scope table this ABI covers).
 *
 * Conventions (SURVEY.md §8b "what a C-ABI replacement must export"):
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every *_dev pointer is device memory owned by the caller (PyTorch's caching allocator on the
 *     reference side); the library never frees or retains caller buffers beyond the call — packed
 *     weights and workspaces are caller-allocated too (sizes from the *_bytes functions);
 *   - every entry point enqueues on the given stream (a cudaStream_t passed as void*) and returns
 *     without synchronising, except where a host copy of a small table is documented;
 *   - return value 0 = ok, < 0 = error; veto_last_error() gives the message (thread-local).
 *     No C++ exception crosses the ABI (the reference raises RuntimeError from AT_ERROR /
 *     AT_ASSERTM, pysgg/csrc/ROIAlign.h:11-44; the Python host turns a negative code into one).
 */
#ifndef VETO_B200_H
#define VETO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VETO_ABI_VERSION 4   /* 2: veto_train_inputs grew the MEET group-head fields; 3: depth backbone entry points;
                              4: veto_meet_group_labels, precision modes F16C8 / F16 */

enum {
    VETO_OK = 0,
    VETO_ERR_ARG = -1,          /* bad argument / shape */
    VETO_ERR_CUDA = -2,         /* a CUDA runtime or driver call failed */
    VETO_ERR_UNSUPPORTED = -3,  /* shape outside what the kernels are built for */
    VETO_ERR_WORKSPACE = -4     /* workspace / packed buffer too small */
};

/* Arithmetic of the token-wise GEMMs of the encoder (model_veto.py:15-26).
 *   FP32   : fp32 SIMT FMA everywhere (the reference's own precision, DTYPE "float32").
 *   BF16X3 : tcgen05 tensor cores, every operand split into bf16 hi+lo, 3 MMAs per product,
 *            fp32 accumulate in TMEM — fp32-grade results (rel. error ~1e-5).
 *   BF16   : tcgen05 tensor cores, single bf16 pass, fp32 accumulate (stated tolerance 3e-2).
 *   F16C8  : tcgen05 tensor cores, ONE fp16 product (kind::f16) plus the two first-order rounding corrections as one
 *            e4m3 product over 2K (kind::f8f6f4, twice the rate): 2 bf16-MMA equivalents instead of BF16X3's 3,
 *            rel. error ~1e-4 (bar 1e-3).  Inference forward of the encoder; the per-box patch projections and the
 *            whole training step run as BF16X3.  Operand ranges |activation| < 3584, |weight| < 32 (saturating).
 *   F16    : single fp16 product (stated tolerance 5e-3); training runs as BF16. */
enum { VETO_PREC_FP32 = 0, VETO_PREC_BF16X3 = 1, VETO_PREC_BF16 = 2, VETO_PREC_F16C8 = 3, VETO_PREC_F16 = 4 };

typedef void* veto_stream_t; /* cudaStream_t */

int veto_abi_version(void);
const char* veto_last_error(void);
/* 0 if the current device can run the sm_100a kernels, VETO_ERR_UNSUPPORTED otherwise. */
int veto_device_check(void);

/* ------------------------------------------------------------------------------------------
 * a1. Candidate-pair enumeration.
 * Replaces RelationSampling.prepare_test_pairs
 * (pysgg/modeling/roi_heads/relation_head/sampling.py:31-52).
 *
 * For image b with n_b boxes: all ordered pairs (i,j), i != j, row-major (== torch.nonzero of
 * ones-eye); if require_overlap, only pairs with boxlist_iou > 0 (structures/boxlist_ops.py:54-87,
 * +1 convention); if more than max_pairs remain, the top max_pairs by scores[i]*scores[j] (fp32),
 * ordered by (product descending, row-major index ascending) — the reference's torch.sort is
 * unstable, this is the tie-break its CPU path shows; if none remain, the single placeholder
 * pair (0,0) (sampling.py:47-51).
 *
 * n_boxes_host [n_images] (host).  Image b's pairs are written at row out_offset(b) =
 * sum_{a<b} cap(a), cap(a) = max(1, min(n_a*(n_a-1), max_pairs)); counts_out_dev[b] = number of
 * valid rows.  With require_overlap == 0 counts are known on the host (== cap) and
 * counts_out_dev may be NULL.  scratch_dev: >= 3*(n_images+1) int32.
 * boxes_dev [N,4] fp32 xyxy (needed iff require_overlap); scores_dev [N] fp32 (needed iff some
 * image exceeds max_pairs).  The filtered / capped path holds one image's candidates in shared
 * memory: n_b*(n_b-1) <= 16384 (n_b <= 128) there, else VETO_ERR_UNSUPPORTED.
 * The small offset tables are copied host->device on `stream` before the kernel. */
int veto_pairs_capacity(const int32_t* n_boxes_host, int n_images, int max_pairs, int64_t* total_rows);
int veto_pairs_enumerate(const int32_t* n_boxes_host, int n_images,
                         const float* boxes_dev, const float* scores_dev,
                         int require_overlap, int max_pairs,
                         int64_t* pairs_out_dev, int32_t* counts_out_dev,
                         int32_t* scratch_dev, veto_stream_t stream);

/* Per-image local pair indices -> global box indices (roi_relation_predictors.py:4104-4115,
 * a host loop over CPU tensors in the reference).  pairs_dev int64 [R,2]; rel_offsets_dev /
 * box_offsets_dev int32 [n_images+1] prefix sums; subj/obj out int32 [R]. */
int veto_pairs_globalize(const int64_t* pairs_dev, int64_t n_pairs,
                         const int32_t* rel_offsets_dev, const int32_t* box_offsets_dev, int n_images,
                         int32_t* subj_out_dev, int32_t* obj_out_dev, veto_stream_t stream);

/* f2. Training-time relation sampling for ground-truth boxes: RelationSampling.gtbox_relsample
 * (pysgg/modeling/roi_heads/relation_head/sampling.py:54-107), whole batch, one launch.
 * rel_matrix_dev: the images' target "relation" matrices [n_b, n_b] int64 back to back, image b at cell offset
 * mat_offsets_dev[b] (int32 [n_images+1]); box_offsets_dev int32 [n_images+1]; n_boxes_host [n_images] (n_b <= 128).
 * Per image: foreground = the cells with relation > 0 in row-major order — a random num_pos_per_image-subset in
 * random order when there are more (:91-94); background = every other ordered pair i != j in random order, cut to
 * batch_size_per_image - (foreground kept) (:97-99).  The random order is a counter-based hash of (seed, image, cell)
 * — a different stream than the reference's torch.randperm, same distribution (every subset / order equally likely).
 * Outputs: image b's rows start at row b * batch_size_per_image: pairs_out_dev int64 [n_images*batch,2] (foreground
 * rows first), labels_out_dev int64 [n_images*batch] (relation label, 0 for background), counts_out_dev int32
 * [n_images,2] = (foreground rows, total rows); binary_out_dev int64, laid out like rel_matrix_dev: the symmetric
 * binary relatedness matrix rel_sym_binarys (:77-82). */
int veto_relsample_gtbox(const int64_t* rel_matrix_dev, const int32_t* mat_offsets_dev, const int32_t* box_offsets_dev,
                         const int32_t* n_boxes_host, int n_images, int batch_size_per_image, int num_pos_per_image,
                         uint64_t seed, int64_t* pairs_out_dev, int64_t* labels_out_dev, int32_t* counts_out_dev,
                         int64_t* binary_out_dev, veto_stream_t stream);

/* f2 (cont.). RelationSampling.detect_relsample + motif_rel_fg_bg_sampling (sampling.py:109-309): the training
 * sampler for DETECTED boxes (SGDet, SGCls), whole batch, one launch.  Per image: detections prp_* (boxes [P,4] xyxy,
 * labels [P], pred_scores [P]; rows prp_offsets[b] ..), ground truth tgt_* (boxes [T,4], labels [T], relation matrix
 * [T,T] at cell offset rel_offsets[b]); P, T <= 128 (host counts n_prp_host / n_tgt_host are checked).
 *   foreground: for every ground-truth relation (nonzero() order) the detection pairs (a, b), a != b, whose boxes match
 *     its head / tail (same label, IoU > fg_thres); more than num_sample_per_gt_rel (<= 4) of them: that many drawn
 *     without replacement with probability ~ iou_head * iou_tail; more than num_pos_per_image rows overall: a uniform
 *     subset.  background: the other pairs between foreground-labelled detections (require_overlap: only overlapping
 *     ones), the 2 * num_neg best by pred_scores[a] * pred_scores[b], of which num_neg = min(batch - #fg, #bg) are drawn
 *     uniformly.  Nothing at all: two (0,0,0) rows.  Random draws: counter-based hash of (seed, image, item).
 * Outputs: image b's rows start at b * batch_size_per_image: triplets_out_dev int64 [n_images*batch, 3] (subject,
 * object, label; foreground first), corrsp_out_dev int64 [n_images*batch] (index of the ground-truth relation of a
 * foreground row, -1 for background), counts_out_dev int32 [n_images, 2] = (foreground rows, total rows),
 * binary_out_dev int64 [P,P] per image at cell offset bin_offsets[b] (symmetric relatedness of detections, :216-227),
 * locating_out_dev fp32 [sum P] (1 where a detection overlaps any ground-truth box by more than fg_thres, :137-141). */
int veto_relsample_detect(const float* prp_boxes_dev, const int64_t* prp_labels_dev, const float* prp_scores_dev,
                          const int32_t* prp_offsets_dev, const float* tgt_boxes_dev, const int64_t* tgt_labels_dev,
                          const int32_t* tgt_offsets_dev, const int64_t* tgt_rel_dev, const int32_t* rel_offsets_dev,
                          const int32_t* bin_offsets_dev, const int32_t* n_prp_host, const int32_t* n_tgt_host,
                          int n_images, float fg_thres, int require_overlap, int num_sample_per_gt_rel,
                          int batch_size_per_image, int num_pos_per_image, uint64_t seed, int64_t* triplets_out_dev,
                          int64_t* corrsp_out_dev, int32_t* counts_out_dev, int64_t* binary_out_dev,
                          float* locating_out_dev, veto_stream_t stream);

/* f2 / a10. Training-time group sampling + relabelling of VETOPredictor_MEET — replaces the per-pair Python loops of
 * VETOPredictor_MEET.forward (roi_relation_predictors.py:3940-3969: cur_chosen_matrix, one .item() sync per pair) and
 * Ensemble.forward (:3812-3821: per-head relabelling).  rel_labels_dev int64 [n_pairs] (global predicate ids, 0 =
 * background); incre_idx_dev int32 [num_rel] = 1-based group of every predicate (incre_idx_list, extra_function_utils.py:
 * 39-52); rates_dev double [n_groups, num_rel] = generate_sample_rate_vector_sep2 (:185-257); local_label_dev int32
 * [n_groups, num_rel] = head k's local label of predicate p (0 for p = 0, 1-based position among the head's members,
 * n_k + 1 for any other foreground predicate).  zero_mode: 0 'rand_insert', 1 'rand_choose', 2 'all_include'
 * (GCL_SETTING.ZERO_LABEL_PADDING_MODE).  Output head_labels_out_dev int64 [n_groups, n_pairs]: the head-local label of
 * every pair the head trains on, -1 elsewhere (veto_train_inputs.head_labels).  Draws are counter-based in (seed, pair);
 * draws_dev (double [n_pairs], the u of every foreground / rand_choose pair) and bg_heads_dev (int32 [n_pairs], the head of
 * every rand_insert background pair) optionally inject the reference's own `random` stream (parity tests); NULL otherwise. */
int veto_meet_group_labels(const int64_t* rel_labels_dev, int64_t n_pairs, const int32_t* incre_idx_dev,
                           const double* rates_dev, const int32_t* local_label_dev, int n_groups, int num_rel,
                           int zero_mode, uint64_t seed, const double* draws_dev, const int32_t* bg_heads_dev,
                           int64_t* head_labels_out_dev, veto_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a3. ROIAlign.  veto_roi_align_forward is the one-for-one replacement of
 * _C.roi_align_forward(input, rois, spatial_scale, ph, pw, sampling_ratio)
 * (pysgg/csrc/vision.cpp:11, ROIAlign.h:11-24, cuda/ROIAlign_cuda.cu:65-122,257-299): legacy
 * (non-aligned) ROIAlign, input NCHW fp32 [B,C,H,W], rois [N,5] fp32 (batch idx, x1,y1,x2,y2),
 * out [N,C,ph,pw] fp32.  Bit-exact with the reference's CPU kernel (ROIAlign_cpu.cpp:114-219).
 * sampling_ratio must be > 0 and ph*pw*sampling_ratio^2 <= 1024.
 *
 * veto_roi_align_backward replaces _C.roi_align_backward (ROIAlign_cuda.cu:178-254): scatters
 * grad [N,C,ph,pw] into grad_input [B,C,H,W] (zero-initialised here) with fp32 atomics. */
int veto_roi_align_forward(const float* input_dev, int batch, int channels, int height, int width,
                           const float* rois_dev, int n_rois, float spatial_scale,
                           int pooled_h, int pooled_w, int sampling_ratio,
                           float* out_dev, veto_stream_t stream);
int veto_roi_align_backward(const float* grad_dev, const float* rois_dev, int n_rois, float spatial_scale,
                            int pooled_h, int pooled_w, int batch, int channels, int height, int width,
                            int sampling_ratio, float* grad_input_dev, veto_stream_t stream);

/* a2+a3 fused: Pooler.forward depth + RGB branch (pysgg/modeling/poolers.py:109-171) as used by
 * VETOFeatureExtractor.forward (box_head/roi_box_feature_extractors.py:116-135) in ONE launch:
 * rois from boxes + per-image box offsets (poolers.py:96-107), FPN level per box by LevelMapper
 * (poolers.py:32-43; area with the +1 convention, bounding_box.py:249-253), RGB pooled from its
 * level's map, depth always from `depth_dev` with `depth_scale` (poolers.py:144-153).
 * feats_dev: n_levels (<= 4) device pointers (host array), each [B,C,feat_h[l],feat_w[l]].
 * boxes_dev [N,4] fp32 xyxy; box_offsets_dev int32 [n_images+1].
 * Outputs out_rgb_dev / out_depth_dev [N,C,pool,pool] fp32; levels_out_dev int32 [N] or NULL (give it: the FPN level
 * of a box is then computed once by a separate launch and read by the gather kernels, instead of once per CTA). */
int veto_roi_gather_forward(const float* const* feats_dev, const int32_t* feat_h, const int32_t* feat_w,
                            const float* scales, int n_levels, int k_min, int k_max,
                            const float* depth_dev, int depth_h, int depth_w, float depth_scale,
                            int batch, int channels,
                            const float* boxes_dev, const int32_t* box_offsets_dev, int n_images, int n_boxes,
                            int pool, int sampling_ratio,
                            float* out_rgb_dev, float* out_depth_dev, int32_t* levels_out_dev,
                            veto_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a5-a10. The relation head proper: VETOPredictor.forward / Ensemble.forward
 * (roi_relation_predictors.py:4074-4139, 3752-3853) with VETOTransformer (model_veto.py).
 * The architecture constants of configs/VETO_final.yaml are compiled in (T_INPUT_DIM 576, 6 heads
 * of 96, MLP 1152, PATCH_SIZE 2 on 8x8 maps -> 19 tokens, 256 ROI channels, proj_d 512 + proj_v 64,
 * 128-d position and 200-d class embeddings); veto_config is validated against them. */
typedef struct {
    int32_t dim;        /* 576  MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER.T_INPUT_DIM */
    int32_t layers;     /* 6    ENC_LAYERS (1..16) */
    int32_t heads;      /* 6    NHEADS */
    int32_t mlp_dim;    /* 1152 (model_veto.py:35) */
    int32_t channels;   /* 256  ROI feature channels */
    int32_t pool;       /* 8    POOLER_RESOLUTION */
    int32_t patch;      /* 2    PATCH_SIZE */
    int32_t num_obj;    /* 151 (VG) / 201 (GQA) */
    int32_t num_out;    /* rel_out rows: 51 / 101, or for MEET the sum over heads of (n_k + 2) */
    int32_t precision;  /* VETO_PREC_* */
} veto_config;

#define VETO_MAX_LAYERS 16

/* Device pointers to the reference module's parameters, fp32, in state_dict layout
 * (SURVEY.md §8a "State-dict keys"). */
typedef struct {
    const float* obj_embed;                 /* obj_embed.weight            [num_obj,200] */
    const float* class_proj_w;              /* class_projection.0.weight   [576,400] */
    const float* class_proj_b;              /* class_projection.0.bias     [576] */
    const float* bn_weight;                 /* pos_embed.0.weight          [4] */
    const float* bn_bias;                   /* pos_embed.0.bias            [4] */
    const float* bn_mean;                   /* pos_embed.0.running_mean    [4] */
    const float* bn_var;                    /* pos_embed.0.running_var     [4] */
    const float* pos_w;                     /* pos_embed.1.weight          [128,4] */
    const float* pos_b;                     /* pos_embed.1.bias            [128] */
    const float* loc_proj_w;                /* location_projection.0.weight[576,256] */
    const float* loc_proj_b;                /* location_projection.0.bias  [576] */
    const float* cls_token;                 /* ...transformer.cls_token    [576] */
    const float* pos_embedding;             /* ...transformer.pos_embedding[576] */
    const float* proj_d_w;                  /* ...patch_embed.proj_d.weight[512,2048] */
    const float* proj_d_b;                  /* ...patch_embed.proj_d.bias  [512] */
    const float* proj_v_w;                  /* ...patch_embed.proj_v.weight[64,2048] */
    const float* proj_v_b;                  /* ...patch_embed.proj_v.bias  [64] */
    const float* ln1_w[VETO_MAX_LAYERS];    /* layers.i.0.norm.weight      [576] */
    const float* ln1_b[VETO_MAX_LAYERS];    /* layers.i.0.norm.bias        [576] */
    const float* qkv_w[VETO_MAX_LAYERS];    /* layers.i.0.fn.to_qkv.weight [1728,576] (no bias) */
    const float* out_w[VETO_MAX_LAYERS];    /* layers.i.0.fn.to_out.0.weight [576,576] */
    const float* out_b[VETO_MAX_LAYERS];    /* layers.i.0.fn.to_out.0.bias [576] */
    const float* ln2_w[VETO_MAX_LAYERS];    /* layers.i.1.norm.weight      [576] */
    const float* ln2_b[VETO_MAX_LAYERS];    /* layers.i.1.norm.bias        [576] */
    const float* ff1_w[VETO_MAX_LAYERS];    /* layers.i.1.fn.net.0.weight  [1152,576] */
    const float* ff1_b[VETO_MAX_LAYERS];    /* layers.i.1.fn.net.0.bias    [1152] */
    const float* ff2_w[VETO_MAX_LAYERS];    /* layers.i.1.fn.net.3.weight  [576,1152] */
    const float* ff2_b[VETO_MAX_LAYERS];    /* layers.i.1.fn.net.3.bias    [576] */
    const float* rel_out_w;                 /* rel_out.weight [num_out,576] (MEET: heads concatenated) */
    const float* rel_out_b;                 /* rel_out.bias   [num_out] */
} veto_weights;

/* Derived weights (subject/object-factored projections, bf16 hi/lo splits for the tensor-core
 * modes).  The caller allocates veto_packed_bytes(cfg) of device memory, veto_pack_weights fills
 * it; it must be re-packed whenever the source parameters change.  The source pointers in
 * veto_weights must stay valid for every later forward call (they are read directly too). */
size_t veto_packed_bytes(const veto_config* cfg);
int veto_pack_weights(const veto_config* cfg, const veto_weights* w, void* packed_dev, size_t packed_bytes,
                      veto_stream_t stream);

typedef struct {
    int32_t n_boxes;                /* N */
    int64_t n_pairs;                /* R */
    const float* boxes;             /* [N,4] fp32 xyxy pixels (BoxList.bbox) */
    const int64_t* labels;          /* [N] hard class ids: predcls labels (…:4087) or MEET obj_preds
                                       (…:3783); NULL => soft embedding from `obj_logits` */
    const float* obj_logits;        /* [N,num_obj] predict_logits (softmax @ obj_embed, …:4095) or NULL */
    const float* roi_rgb;           /* roi_features        [N,256,8,8] */
    const float* roi_depth;         /* roi_depth_features  [N,256,8,8] */
    const int32_t* subj;            /* [R] global subject box index */
    const int32_t* obj;             /* [R] global object box index */
    const float* freq_bias;         /* optional [num_obj*num_obj, num_out] table (model_motifs.py:29-38),
                                       added as row labels[s]*num_obj+labels[o]; NULL = off (the
                                       reference never applies it for VETO: SURVEY.md note A) */
} veto_inputs;

typedef struct {
    float* rel_logits;              /* [R,num_out] */
    float* rel_features;            /* optional [R,576]: encoder CLS output (x[:,0], model_veto.py:25) */
    float* tokens;                  /* optional [R,19,576]: encoder input (debug / stage parity) */
} veto_outputs;

/* Pairs are processed in chunks of `chunk_pairs` (0 = library default) so that the activations of
 * one chunk stay L2-resident; the workspace scales with N and the chunk, not with R. */
size_t veto_workspace_bytes(const veto_config* cfg, int32_t n_boxes, int64_t n_pairs, int32_t chunk_pairs);
int veto_relation_forward(const veto_config* cfg, const veto_weights* w, const void* packed_dev,
                          const veto_inputs* in, const veto_outputs* out,
                          void* workspace_dev, size_t workspace_bytes, int32_t chunk_pairs,
                          veto_stream_t stream);
/* ------------------------------------------------------------------------------------------
 * a11. Training step of the relation head: VETOPredictor.forward in train() mode
 * (roi_relation_predictors.py:4074-4136: same trunk with the Dropout layers active and BatchNorm1d batch
 * statistics, then add_losses['rel_loss'] = CrossEntropyLoss(weight)(rel_dists, cat(rel_labels))) AND the backward
 * pass autograd would run from that loss (tools/relation_train_net.py:451-452), in one call: forward with saved
 * activations, loss, then gradients of every parameter the loss depends on, and of the ROI features.
 *
 * Dropout (pos_embed Dropout(0.1), Transformer.pos_drop EMB_DROPOUT, Attention.to_out T_DROPOUT; model_veto.py:43,
 * 83-85) uses a counter-based mask, keyed by `seed` and the element index, regenerated in the backward pass.
 * BatchNorm1d(4, momentum) normalises with the statistics of all boxes of the step and updates the running
 * statistics in place (bn_running_mean / bn_running_var, NULL = leave them).
 *
 * veto_grads mirrors veto_weights (fp32, state_dict layout); every non-NULL gradient is OVERWRITTEN (not
 * accumulated).  bn_mean / bn_var have no gradient and must be NULL.  Parameters the reference never uses
 * (obj_embed2, bbox_embed: SURVEY.md §8a) have none either.  Every reduction has a fixed order: the step is
 * bitwise reproducible. */
typedef struct {
    const int64_t* rel_labels;      /* [R] predicate class per pair (cat(rel_labels), …:4134) */
    const float* class_weight;      /* [num_out] CrossEntropyLoss weight (criterion_loss_rel.weight) or NULL = ones */
    const int32_t* rel_offsets;     /* [n_images+1] device: the pairs of image b are rows rel_offsets[b] .. rel_offsets[b+1] */
    const int32_t* box_offsets;     /* [n_images+1] device: boxes of image b */
    int32_t n_images;
    float p_pos_dropout;            /* 0.1  pos_embed Dropout (…:4046) */
    float p_emb_dropout;            /* EMB_DROPOUT 0.35 (model_veto.py:43,63) */
    float p_attn_dropout;           /* T_DROPOUT 0.35  (model_veto.py:83-85) */
    uint64_t seed;
    float bn_momentum;              /* 0.001 (…:4043) */
    float* bn_running_mean;         /* [4] updated in place, or NULL */
    float* bn_running_var;          /* [4] updated in place, or NULL */
    /* MEET group heads (VETOPredictor_MEET / Ensemble in train() mode, roi_relation_predictors.py:3812-3846): with
     * n_heads > 1 the num_out logit columns are n_heads classifiers, head k owning columns head_offsets[k] ..
     * head_offsets[k+1]; the loss of head k is the plain mean CE over the pairs the group sampling chose for it,
     * against group-local labels: head_labels[k*R + r] = label of pair r in head k, or -1 = pair r is not in head k's
     * loss.  outputs.loss then holds n_heads values ('group_k_CE_loss') and the gradients are those of their SUM
     * (what the trainer back-propagates, tools/relation_train_net.py:451-452).  rel_labels / class_weight are unused.
     * n_heads <= 1: the single rel_out head above. */
    int32_t n_heads;
    const int32_t* head_offsets;    /* HOST [n_heads+1] */
    const int64_t* head_labels;     /* device [n_heads, R] */
} veto_train_inputs;

typedef struct {
    float* obj_embed; float* class_proj_w; float* class_proj_b;
    float* bn_weight; float* bn_bias; float* bn_mean; float* bn_var;
    float* pos_w; float* pos_b; float* loc_proj_w; float* loc_proj_b;
    float* cls_token; float* pos_embedding;
    float* proj_d_w; float* proj_d_b; float* proj_v_w; float* proj_v_b;
    float* ln1_w[VETO_MAX_LAYERS]; float* ln1_b[VETO_MAX_LAYERS];
    float* qkv_w[VETO_MAX_LAYERS];
    float* out_w[VETO_MAX_LAYERS]; float* out_b[VETO_MAX_LAYERS];
    float* ln2_w[VETO_MAX_LAYERS]; float* ln2_b[VETO_MAX_LAYERS];
    float* ff1_w[VETO_MAX_LAYERS]; float* ff1_b[VETO_MAX_LAYERS];
    float* ff2_w[VETO_MAX_LAYERS]; float* ff2_b[VETO_MAX_LAYERS];
    float* rel_out_w; float* rel_out_b;
} veto_grads;   /* same field order as veto_weights */

typedef struct {
    float* loss;                    /* [1] rel_loss ([n_heads] group losses for MEET) */
    float* rel_logits;              /* optional [R,num_out]: the training-mode logits */
    float* grad_roi_depth;          /* optional [N,256,8,8]: d loss / d roi_depth_features (flows on into the depth
                                       backbone through veto_roi_align_backward) */
    float* grad_roi_rgb;            /* optional [N,256,8,8] (the RGB backbone is frozen in the reference: normally NULL) */
} veto_train_outputs;

size_t veto_train_workspace_bytes(const veto_config* cfg, int32_t n_boxes, int64_t n_pairs);
int veto_relation_train_step(const veto_config* cfg, const veto_weights* w, const void* packed_dev,
                             const veto_inputs* in, const veto_train_inputs* tin, const veto_grads* grads,
                             const veto_train_outputs* out, void* workspace_dev, size_t workspace_bytes,
                             veto_stream_t stream);

/* Total number of kernel launches this thread has enqueued through the library so far (monotonic). */
int64_t veto_last_launch_count(void);

/* Per-stage device timing (used by bench.py for the roofline line; not a profiler): between begin and end every
 * kernel launch of this thread is bracketed by CUDA events on `stream`; end synchronises and adds, per stage tag,
 * the elapsed milliseconds and the launch count into the caller's host arrays of length VETO_PROFILE_TAGS. */
#define VETO_PROFILE_TAGS 24
int veto_profile_begin(veto_stream_t stream);
int veto_profile_end(double* ms_by_tag_host, int64_t* launches_by_tag_host);
const char* veto_profile_tag_name(int tag);

/* ------------------------------------------------------------------------------------------
 * a12 (row f1). PostProcessor.forward, vanilla branch (relation_head/inference.py:398-453):
 * softmax over rel logits, max over classes 1.., triple score rel*obj_s*obj_o, per-image sort
 * descending (ties: original row ascending).  obj_scores [N] fp32; rel_offsets int32 [n_images+1]
 * over rows of rel_logits; pairs int64 [R,2] local indices; box_offsets int32 [n_images+1].
 * Outputs (all [R] rows, image-segmented like the input): sorted pairs int64 [R,2], class
 * probabilities [R,num_rel], labels int64 [R], triple scores [R].  One image is sorted in
 * shared memory: R_i <= 16384.  */
int veto_postprocess(const float* rel_logits_dev, int num_rel, const int64_t* pairs_dev,
                     const float* obj_scores_dev, const int32_t* rel_offsets_dev,
                     const int32_t* box_offsets_dev, int n_images, int64_t n_pairs,
                     int64_t* pairs_out_dev, float* probs_out_dev, int64_t* labels_out_dev,
                     float* triple_out_dev, veto_stream_t stream);

/* f1. PostProcessor.forward, MEET 'ensemble' branch (ENSEMBLE_LEARNING.ENABLED, EXPERT_GROUP False;
 * relation_head/inference.py:284-397).  group_logits_dev [R,num_out]: the group heads' logits side by side (head k =
 * columns head_offsets[k] .. head_offsets[k+1], n_k + 2 of them: background, the n_k member predicates, out-of-group).
 * Per image and head: softmax over the head's columns, the out-of-group column dropped, score / head-local label = max
 * over the member columns, triple = score * obj_s * obj_o; the image's G*R_i candidate rows (group-major) are ranked by
 * triple score descending (ties: merged index ascending) and each row's probabilities are scattered into the global
 * predicate columns col_map_dev[num_out] (head column -> global predicate id; the dropped columns are never read).
 * Outputs are image-segmented with G*R_i rows per image at row offset G*rel_offsets[i]: pairs int64 [G*R,2]
 * (the reference stores them in a float32 tensor, :380), probabilities [G*R,num_rel], head-local labels int64 [G*R]
 * (:371,388), triple scores [G*R].  One image is sorted in shared memory: G*R_i <= 16384.
 * The reference consumes image 0 only (TEST.IMS_PER_BATCH 1); this entry point handles a batch. */
int veto_postprocess_meet(const float* group_logits_dev, int num_out, const int32_t* head_offsets_dev, int n_heads,
                          const int32_t* col_map_dev, int num_rel, const int64_t* pairs_dev,
                          const float* obj_scores_dev, const int32_t* rel_offsets_dev,
                          const int32_t* box_offsets_dev, int n_images, int64_t n_pairs, int64_t* pairs_out_dev,
                          float* probs_out_dev, int64_t* labels_out_dev, float* triple_out_dev, veto_stream_t stream);

/* f4. Triplet matching of the recall metrics: SGRecall.calculate_recall -> _compute_pred_matches
 * (pysgg/data/datasets/evaluation/vg/sgg_eval.py:44-117,138-186), non-phrdet modes, a batch of images in one launch.
 * A triplet row = (subject class, predicate, object class), a box row = (subject xyxy, object xyxy) fp32 [.,8];
 * image b's ground-truth rows are gt_offsets[b] .. gt_offsets[b+1], its predictions (in rank order) pred_offsets[b] ..
 * A prediction matches a ground-truth triplet when the three labels are equal and both boxes overlap by IoU >=
 * iou_thres (boxlist_iou, +1 convention, fp32).  first_match_dev int32 [G]: rank (within the image) of the first
 * matching prediction, INT32_MAX = none — recall@K of an image = #{g : first_match[g] < K} / G_b (:158-160);
 * pred_hits_dev int32 [P]: number of ground-truth triplets each prediction matches (len(pred_to_gt[p])). */
int veto_sgg_match(const int64_t* gt_triplets_dev, const float* gt_boxes_dev, const int32_t* gt_offsets_dev,
                   const int64_t* pred_triplets_dev, const float* pred_boxes_dev, const int32_t* pred_offsets_dev,
                   int n_images, float iou_thres, int32_t* first_match_dev, int32_t* pred_hits_dev, veto_stream_t stream);

/* f1 (cont.). PostProcessor.forward, MEET EXPERT_GROUP branch (relation_head/inference.py:93-283): three experts per
 * group; group_logits_dev [R,num_out] holds the 3*G heads expert-major (head e*G + j = expert e of group j, columns
 * head_offsets[e*G+j] ..; the three experts of a group have the same width n_j + 2), col_map_dev as above.  A candidate
 * (group, pair) survives when the experts' predicted classes agree: consensus = 0: all three ('U'; score and
 * probabilities = the mean over the experts), consensus = 1: at least two ('C'; mean over the agreeing expert pairs of
 * the pair means, with the reference's mean(p1, p1) for the pair (1,2), :191-193).  Survivors of an image are ranked by
 * score; outputs as veto_postprocess_meet (capacity G*R_i rows per image at row offset G*rel_offsets[i]) plus
 * counts_out_dev int32 [n_images] = survivors per image (rows beyond it are unspecified). */
int veto_postprocess_meet_vote(const float* group_logits_dev, int num_out, const int32_t* head_offsets_dev, int n_groups,
                               const int32_t* col_map_dev, int num_rel, int consensus, const int64_t* pairs_dev,
                               const float* obj_scores_dev, const int32_t* rel_offsets_dev,
                               const int32_t* box_offsets_dev, int n_images, int64_t n_pairs, int64_t* pairs_out_dev,
                               float* probs_out_dev, int64_t* labels_out_dev, float* triple_out_dev,
                               int32_t* counts_out_dev, veto_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a10. Ensemble.nms_per_cls (roi_relation_predictors.py:3855-3874) with nms_overlaps
 * (relation_head/utils_relation.py:56-79): MEET's greedy per-class label assignment at SGDet test time.
 * Per image, n rounds: take the arg max of the [n, num_obj] score tile (column 0 excluded; first index in
 * row-major order on ties, like numpy.argmax), give that box that class, zero the class's score of every box
 * whose class box overlaps it by IoU >= thresh (+1 box convention, fp32, the reference's operation order), and
 * retire the box.  scores_dev [N,num_obj] fp32 = softmax of the detector distribution (the caller computes it:
 * the reference feeds softmax(one_hot(pred_labels))); boxes_per_cls_dev [N,num_obj,4] fp32 xyxy;
 * box_offsets_dev int32 [n_images+1]; n_boxes_host [n_images]; labels_out_dev int64 [N].
 * late_nms = 1 selects obj_prediction_nms (relation_head/utils_relation.py:94-128), the late NMS of
 * PostProcessor.forward at SGDet test time (relation_head/inference.py:414-417): same rounds, but the background
 * column counts as score 0 (not -1) and a box keeps its first assignment; scores = softmax(predict_logits).
 * One CTA per image with the score tile in shared memory: n * num_obj <= 56 320 (80 boxes x 151 classes = 12 080). */
int veto_obj_nms_per_cls(const float* scores_dev, const float* boxes_per_cls_dev, const int32_t* box_offsets_dev,
                         const int32_t* n_boxes_host, int n_images, int num_obj, float thresh, int late_nms,
                         int64_t* labels_out_dev, veto_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f3. The depth backbone (registry "R-18-C4": pysgg/modeling/backbone/backbone.py:83-93 building
 * ResNetDepth, pysgg/modeling/backbone/resnet_depth.py:11-47 — torchvision's ResNet-18 with a one-channel first
 * convolution, truncated after layer3), the producer of `depth_features` (detector/generalized_rcnn.py:53-54) and the
 * one other module relation_train_net.py:166-170 trains: forward in eval() or train() mode (BatchNorm2d on batch
 * statistics, running statistics updated in place) and the backward pass from the gradient of its output.
 *
 * The 15 convolutions in module order, each followed by its BatchNorm2d:
 *   0 conv1 (1->64, 7x7 s2 p3) | layer1: 1,2 = block 0 conv1,conv2; 3,4 = block 1 | layer2: 5,6 = block 0 conv1 (s2),
 *   conv2; 7 = block 0 downsample (1x1 s2); 8,9 = block 1 | layer3: 10,11,12 and 13,14 likewise.
 * Convolutions run as an NHWC im2col into the operand format of the precision mode followed by the library's GEMMs
 * (forward: tcgen05 bf16 / bf16x3 or fp32 SIMT; weight gradients on the MN-major tcgen05 GEMM). */
#define VETO_DEPTH_CONVS 15
typedef struct {
    const float* conv_w[VETO_DEPTH_CONVS];   /* [Cout,Cin,k,k] as nn.Conv2d holds them */
    const float* bn_w[VETO_DEPTH_CONVS];     /* [Cout] */
    const float* bn_b[VETO_DEPTH_CONVS];
    float* bn_mean[VETO_DEPTH_CONVS];        /* running_mean / running_var: read in eval mode, updated in training */
    float* bn_var[VETO_DEPTH_CONVS];
} veto_depth_weights;
typedef struct {
    float* conv_w[VETO_DEPTH_CONVS];         /* overwritten (not accumulated) */
    float* bn_w[VETO_DEPTH_CONVS];
    float* bn_b[VETO_DEPTH_CONVS];
} veto_depth_grads;
/* output spatial size for an input of height x width (the stride-16 map) */
void veto_depth_backbone_out_size(int height, int width, int* out_h, int* out_w);
size_t veto_depth_backbone_workspace_bytes(int precision, int batch, int height, int width, int training);
/* depth_dev [B,1,H,W] fp32 -> out_dev [B,256,H/16,W/16] fp32 NCHW (what the Pooler reads).  training != 0: batch
 * statistics with `momentum` (nn.BatchNorm2d default 0.1), eps 1e-5, and the workspace keeps what the backward needs. */
int veto_depth_backbone_forward(int precision, const veto_depth_weights* w, const float* depth_dev, int batch, int height,
                                int width, int training, float momentum, float* out_dev, void* workspace_dev,
                                size_t workspace_bytes, veto_stream_t stream);
/* backward of the last training-mode forward that used this workspace: grad_out_dev [B,256,H/16,W/16] NCHW -> g */
int veto_depth_backbone_backward(int precision, const veto_depth_weights* w, const float* grad_out_dev, int batch,
                                 int height, int width, const veto_depth_grads* g, void* workspace_dev,
                                 size_t workspace_bytes, veto_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Test hooks (used by tests/ only): one GEMM  C[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual)
 * through the SIMT or the tcgen05 kernel, fp32 in / fp32 out. scratch_dev: >= 4*(M*K + N*K) bytes. */
int veto_test_gemm(const float* a_dev, const float* w_dev, const float* bias_dev, const float* residual_dev,
                   float* c_dev, int M, int N, int K, int act, int precision,
                   void* scratch_dev, size_t scratch_bytes, veto_stream_t stream);
int veto_test_layernorm(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev,
                        int64_t rows, veto_stream_t stream);
int veto_test_attention(const float* qkv_dev, float* out_dev, int64_t n_seq, veto_stream_t stream);
/* the tensor-core attention kernel: fp32 result in out_dev; scratch_dev >= 4 * n_seq*19*576 bytes (bf16 hi/lo outputs);
 * split = 1 for the bf16x3 scheme, 0 for single-pass bf16 */
int veto_test_attention_tc(const float* qkv_dev, float* out_dev, void* scratch_dev, int64_t n_seq, int split,
                           veto_stream_t stream);

/* the weight-gradient GEMM out[Nw,Kw] = y[rows,Nw]^T @ x[rows,Kw] (both operands read in place as MN-major tcgen05
 * operands); scratch_dev >= 4*(rows*Nw + rows*Kw) + 4*Nw*Kw*split_k bytes; geometry_host = {LBO, SBO, K advance} bytes of
 * the shared-memory descriptors or NULL for the production constants (the parameter exists for the bring-up sweep) */
int veto_test_gemm_tn(const float* y_dev, const float* x_dev, float* out_dev, int rows, int Nw, int Kw, int precision,
                      int split_k, const uint32_t* geometry_host, void* scratch_dev, size_t scratch_bytes,
                      veto_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VETO_B200_H */
