// C ABI of libveto_b200.so: error plumbing, weight packing, workspace planning and the orchestration of
// VETOPredictor.forward / Ensemble.forward (roi_relation_predictors.py:4074-4139, 3752-3853) over the
// stage kernels.  Declarations: include/veto_b200.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "api_internal.cuh"
#include "train.cuh"

namespace veto {

static thread_local char g_err[768] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("%s: %s (%s) at %s:%d", what, cudaGetErrorString(e), cudaGetErrorName(e), file, line);
    return VETO_ERR_CUDA;
}

// per-stage event timing (veto_profile_begin / _end)
struct ProfEvent { cudaEvent_t ev; int tag; };
static thread_local bool g_prof = false;
static thread_local cudaStream_t g_prof_stream = nullptr;
static thread_local std::vector<ProfEvent> g_prof_events;
static thread_local int g_tag = 0;
static const char* kTagNames[VETO_PROFILE_TAGS] = {"other", "pairs", "roi_gather", "box_stage", "tokens", "layernorm",
                                                   "gemm_qkv", "attention", "gemm_out", "gemm_ff1", "gemm_ff2",
                                                   "classifier", "postprocess", "pack", "bwd_dgrad", "bwd_other", "bwd_attention", "bwd_layernorm",
                                                   "bwd_box_stage", "bwd_wgrad", "loss", "", "", ""};

void set_tag(int tag) { g_tag = tag; }
void count_launch(int n) {
    g_launches += n;
    if (g_prof) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) {
            cudaEventRecord(e, g_prof_stream);
            g_prof_events.push_back({e, g_tag});
        }
    }
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

namespace {

struct WorkLayout {
    size_t pos, emb, lso, cso, pa_d, pa_v, so_d, so_v;  // box level
    size_t x, xn, qkv, h, q_cls, x_cls;                 // chunk level
    size_t xo, ln_parts, ln_stats;                      // LayerNorm fusion: x in operand format, its row statistics
    size_t total;
    int32_t chunk;
};

// 7976 pairs = 151 544 token rows = 592 CTA-pair tiles of 256 rows: with 74 SM pairs every encoder GEMM of a chunk is a
// whole number of waves (8 x 74 row tiles times N/192 = 9, 6 or 3 column tiles).  Measured on configs[2] (202 240 pairs,
// profiles/r2_chunk_sweep.jsonl): 499 -> 514 ms, 997 -> 450, 1994 -> 423, 3988 -> 412, 7976 -> 404 ms per step: the
// intermediates never fit L2 usefully, so larger chunks only amortise launch tails (2.4 GB of workspace at this size).
constexpr int32_t kDefaultChunk = 7976;

// VETO_ATTENTION_SPLIT=0: the fp32-qkv attention kernel in every mode (A/B measurements, diagnosis)
bool split_attention_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("VETO_ATTENTION_SPLIT");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

// The residual stream between the layers lives in operand format only (EPI_RESOP_*): no fp32 copy of x is written or read
// (2.3 KB / row less HBM traffic for each to_out / FF2 launch; logits 8.4e-5 -> 8.8e-5).  The first version read the
// residual with one 12-byte load per four values one chunk ahead and was SLOWER (to_out 51.0 -> 57.9 ms,
// profiles/r2_modes_residual_ab.jsonl); with the next tile's lines prefetched into L2 and the raw words two chunks ahead
// it is faster: to_out 52.2 -> 46.6 ms, FF2 63.4 -> 60.9 ms per step (profiles/r2_modes_residual_ab2.jsonl).
// VETO_RESIDUAL_OPERAND=0 keeps the fp32 residual stream (A/B measurements).
bool resop_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("VETO_RESIDUAL_OPERAND");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

// VETO_LN_STATS_EPILOGUE=1: (mean, rstd) of the LayerNorm-fused rows inside the consuming epilogues instead of by the
// ln_stats_finalize launches (the same formula; the results agree to ~1e-4 of the logit range, not bit for bit — the
// compiler contracts E[x^2] - mean^2 differently in the two places).  Measured SLOWER (profiles/
// r2_modes_stats_epilogue_ab.jsonl): the 11 launches per chunk cost 3.1 ms per step, but the 36 extra loads per thread and
// tile lengthen the to_qkv / FF1 epilogues by 7 and 10 ms — those epilogues are the critical path of their kernels.
bool stats_in_epilogue() {
    static int on = -1;
    if (on < 0) on = getenv("VETO_LN_STATS_EPILOGUE") ? 1 : 0;
    return on != 0;
}

// q, k, v between to_qkv and attention_split in the (sequence, head) item layout (common.cuh qkv_item_offset);
// VETO_QKV_ITEM_LAYOUT=0 keeps the row-major [rows, 1728] arrays (A/B measurements)
bool qkv_item_layout_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("VETO_QKV_ITEM_LAYOUT");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

// VETO_LN_FUSION=0 keeps the LayerNorm kernels everywhere (A/B measurements, diagnosis)
bool ln_fusion_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("VETO_LN_FUSION");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

WorkLayout work_layout(const veto_config& c, int32_t n_boxes, int64_t n_pairs, int32_t chunk_pairs) {
    WorkLayout W{};
    Carver k;
    int32_t chunk = chunk_pairs > 0 ? chunk_pairs : kDefaultChunk;
    if (n_pairs > 0 && chunk > n_pairs) chunk = (int32_t)n_pairs;
    if (chunk < 1) chunk = 1;
    W.chunk = chunk;
    const size_t N = (size_t)(n_boxes > 0 ? n_boxes : 1);
    W.pos = k.take(sizeof(float) * N * kPosDim);
    W.emb = k.take(sizeof(float) * N * kEmbDim);
    W.lso = k.take(sizeof(float) * N * 2 * kDim);
    W.cso = k.take(sizeof(float) * N * 2 * kDim);
    W.pa_d = k.take(act_bytes(prec_box(c.precision), N * kPatches * kPatchVec));
    W.pa_v = k.take(act_bytes(prec_box(c.precision), N * kPatches * kPatchVec));
    W.so_d = k.take(sizeof(float) * N * kPatches * 2 * kDimDepth);
    W.so_v = k.take(sizeof(float) * N * kPatches * 2 * kDimRgb);
    const size_t M = (size_t)chunk * kTokens;
    W.x = k.take(sizeof(float) * M * kDim);
    W.xn = k.take(act_bytes(c.precision, M * kDim));
    W.qkv = k.take(sizeof(float) * M * 3 * kDim);
    W.h = k.take(act_bytes(c.precision, M * kMlp));
    W.q_cls = k.take(sizeof(float) * (size_t)chunk * kDim);
    W.x_cls = k.take(sizeof(float) * (size_t)chunk * kDim);
    if (prec_two_arrays(c.precision)) {
        W.xo = k.take(act_bytes(c.precision, M * kDim));
        W.ln_parts = k.take(sizeof(float2) * (kDim / 64) * M);
        W.ln_stats = k.take(sizeof(float2) * M);
    }
    W.total = k.off;
    return W;
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_abi_version(void) { return VETO_ABI_VERSION; }
extern "C" const char* veto_last_error(void) { return g_err; }
extern "C" int64_t veto_last_launch_count(void) { return g_launches; }

extern "C" int veto_device_check(void) {
    int dev = 0, major = 0, minor = 0;
    VETO_CUDA(cudaGetDevice(&dev));
    VETO_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    VETO_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    VETO_REQUIRE(major == 10 && minor == 0, VETO_ERR_UNSUPPORTED,
                 "libveto_b200 is built for sm_100a only; the current device is sm_%d%d", major, minor);
    return VETO_OK;
}

extern "C" size_t veto_packed_bytes(const veto_config* cfg) {
    if (check_config(cfg)) return 0;
    return packed_layout(*cfg).total;
}

extern "C" int veto_pack_weights(const veto_config* cfg, const veto_weights* w, void* packed_dev, size_t packed_bytes,
                                 veto_stream_t stream) {
    int rc = check_config(cfg);
    if (rc) return rc;
    VETO_REQUIRE(w && packed_dev, VETO_ERR_ARG, "veto_pack_weights: NULL argument");
    const PackedLayout L = packed_layout(*cfg);
    VETO_REQUIRE(packed_bytes >= L.total, VETO_ERR_WORKSPACE, "veto_pack_weights: packed buffer %zu < %zu bytes", packed_bytes,
                 L.total);
    cudaStream_t s = (cudaStream_t)stream;
    char* P = (char*)packed_dev;
    set_tag(TAG_PACK);
    if ((rc = pack_halves(w->loc_proj_w, (float*)(P + L.w_loc2), kDim, kPosDim, s))) return rc;
    if ((rc = pack_bias2(w->loc_proj_b, (float*)(P + L.b_loc2), kDim, s))) return rc;
    if ((rc = pack_halves(w->class_proj_w, (float*)(P + L.w_cls2), kDim, kEmbDim, s))) return rc;
    if ((rc = pack_bias2(w->class_proj_b, (float*)(P + L.b_cls2), kDim, s))) return rc;
    if ((rc = pack_patch(w->proj_d_w, (float*)(P + L.w_d2), kDimDepth, s))) return rc;
    if ((rc = pack_bias2(w->proj_d_b, (float*)(P + L.b_d2), kDimDepth, s))) return rc;
    if ((rc = pack_patch(w->proj_v_w, (float*)(P + L.w_v2), kDimRgb, s))) return rc;
    if ((rc = pack_bias2(w->proj_v_b, (float*)(P + L.b_v2), kDimRgb, s))) return rc;
    if ((rc = pack_add(w->cls_token, w->pos_embedding, (float*)(P + L.clspos), kDim, s))) return rc;
    if (cfg->precision != VETO_PREC_FP32) {
        // one launch re-splits every GEMM weight into its bf16 hi (+ lo) operand arrays
        SplitJob jobs[kMaxSplitJobs];
        int nj = 0;
        auto split = [&](const float* src, size_t hi, size_t lo, size_t n, int fmt = FMT_BF16, const float* col_scale = nullptr,
                         int row_len = 0) {
            jobs[nj++] = SplitJob{src, (__nv_bfloat16*)(P + hi), lo ? (__nv_bfloat16*)(P + lo) : nullptr, n, fmt, col_scale, row_len};
        };
        const int efmt = prec_encoder_fmt(cfg->precision);   // encoder weights in the mode's operand format
        split((const float*)(P + L.w_d2), L.d2_hi, L.d2_lo, (size_t)2 * kDimDepth * kPatchVec);
        split((const float*)(P + L.w_v2), L.v2_hi, L.v2_lo, (size_t)2 * kDimRgb * kPatchVec);
        for (int l = 0; l < cfg->layers; ++l) {
            VETO_REQUIRE(w->qkv_w[l] && w->out_w[l] && w->ff1_w[l] && w->ff2_w[l], VETO_ERR_ARG, "layer %d weights missing", l);
            split(w->qkv_w[l], L.qkv_hi[l], L.qkv_lo[l], (size_t)3 * kDim * kDim, efmt);
            split(w->out_w[l], L.out_hi[l], L.out_lo[l], (size_t)kDim * kDim, efmt);
            split(w->ff1_w[l], L.ff1_hi[l], L.ff1_lo[l], (size_t)kMlp * kDim, efmt);
            split(w->ff2_w[l], L.ff2_hi[l], L.ff2_lo[l], (size_t)kDim * kMlp, efmt);
            // LayerNorm-fused copies: gamma folded into the columns, and the constants of the epilogue
            VETO_REQUIRE(w->ln1_w[l] && w->ln1_b[l] && w->ln2_w[l] && w->ln2_b[l] && w->ff1_b[l], VETO_ERR_ARG,
                         "layer %d LayerNorm parameters missing", l);
            split(w->qkv_w[l], L.qkvf_hi[l], L.qkvf_lo[l], (size_t)3 * kDim * kDim, efmt, w->ln1_w[l], kDim);
            split(w->ff1_w[l], L.ff1f_hi[l], L.ff1f_lo[l], (size_t)kMlp * kDim, efmt, w->ln2_w[l], kDim);
            float* cq = (float*)(P + L.c_qkv[l]);
            float* cf = (float*)(P + L.c_ff1[l]);
            if ((rc = ln_fold_consts(w->qkv_w[l], w->ln1_w[l], w->ln1_b[l], nullptr, 3 * kDim, kDim, cq, cq + 3 * kDim, s))) return rc;
            if ((rc = ln_fold_consts(w->ff1_w[l], w->ln2_w[l], w->ln2_b[l], w->ff1_b[l], kMlp, kDim, cf, cf + kMlp, s))) return rc;
        }
        if ((rc = pack_split_bf16_multi(jobs, nj, s))) return rc;
    }
    return VETO_OK;
}

extern "C" size_t veto_workspace_bytes(const veto_config* cfg, int32_t n_boxes, int64_t n_pairs, int32_t chunk_pairs) {
    if (check_config(cfg)) return 0;
    return work_layout(*cfg, n_boxes, n_pairs, chunk_pairs).total;
}

extern "C" int veto_relation_forward(const veto_config* cfg, const veto_weights* w, const void* packed_dev,
                                     const veto_inputs* in, const veto_outputs* out, void* workspace_dev,
                                     size_t workspace_bytes, int32_t chunk_pairs, veto_stream_t stream) {
    int rc = check_config(cfg);
    if (rc) return rc;
    VETO_REQUIRE(w && packed_dev && in && out && workspace_dev, VETO_ERR_ARG, "veto_relation_forward: NULL argument");
    VETO_REQUIRE(in->n_boxes >= 0 && in->n_pairs >= 0, VETO_ERR_ARG, "veto_relation_forward: negative sizes");
    if (in->n_pairs == 0 || in->n_boxes == 0) return VETO_OK;
    VETO_REQUIRE(in->boxes && in->roi_rgb && in->roi_depth && in->subj && in->obj && (in->labels || in->obj_logits),
                 VETO_ERR_ARG, "veto_relation_forward: missing input pointer");
    VETO_REQUIRE(out->rel_logits, VETO_ERR_ARG, "veto_relation_forward: rel_logits output missing");
    VETO_REQUIRE(!in->freq_bias || in->labels, VETO_ERR_ARG, "frequency bias needs hard labels");
    const int prec = cfg->precision;
    const PackedLayout L = packed_layout(*cfg);
    const WorkLayout W = work_layout(*cfg, in->n_boxes, in->n_pairs, chunk_pairs);
    VETO_REQUIRE(workspace_bytes >= W.total, VETO_ERR_WORKSPACE, "veto_relation_forward: workspace %zu < %zu bytes",
                 workspace_bytes, W.total);
    VETO_REQUIRE(W.chunk <= 65536, VETO_ERR_UNSUPPORTED, "chunk_pairs=%d > 65536", W.chunk);
    cudaStream_t s = (cudaStream_t)stream;
    const char* P = (const char*)packed_dev;
    char* B = (char*)workspace_dev;
    const int N = in->n_boxes;

    // ---------------- box stage: everything that depends on one box only ----------------
    float* pos = (float*)(B + W.pos);
    float* emb = (float*)(B + W.emb);
    float* lso = (float*)(B + W.lso);
    float* cso = (float*)(B + W.cso);
    float* so_d = (float*)(B + W.so_d);
    float* so_v = (float*)(B + W.so_v);
    set_tag(TAG_BOX);
    if ((rc = box_embed(in->boxes, in->labels, in->obj_logits, cfg->num_obj, N, *w, pos, emb, s))) return rc;
    {
        GemmEpilogue ep;
        ep.bias = (const float*)(P + L.b_loc2);
        ep.out.f32 = lso;
        ep.ldc = 2 * kDim;
        if ((rc = gemm_simt(pos, kPosDim, (const float*)(P + L.w_loc2), N, 2 * kDim, kPosDim, ep, s))) return rc;
        ep.bias = (const float*)(P + L.b_cls2);
        ep.out.f32 = cso;
        if ((rc = gemm_simt(emb, kEmbDim, (const float*)(P + L.w_cls2), N, 2 * kDim, kEmbDim, ep, s))) return rc;
    }
    {
        const size_t pe = (size_t)N * kPatches * kPatchVec;
        ActBuf pa_d = act_at(B, W.pa_d, prec_box(prec), pe), pa_v = act_at(B, W.pa_v, prec_box(prec), pe);
        if ((rc = patchify(in->roi_depth, N, pa_d.out(), s))) return rc;
        if ((rc = patchify(in->roi_rgb, N, pa_v.out(), s))) return rc;
        GemmEpilogue ep;
        ep.bias = (const float*)(P + L.b_d2);
        ep.out.f32 = so_d;
        ep.ldc = 2 * kDimDepth;
        WRef wd{(const float*)(P + L.w_d2), bf(P, L.d2_hi), bf(P, L.d2_lo)};
        if ((rc = linear(prec, pa_d, kPatchVec, wd, N * kPatches, 2 * kDimDepth, kPatchVec, ep, s))) return rc;
        ep.bias = (const float*)(P + L.b_v2);
        ep.out.f32 = so_v;
        ep.ldc = 2 * kDimRgb;
        WRef wv{(const float*)(P + L.w_v2), bf(P, L.v2_hi), bf(P, L.v2_lo)};
        if ((rc = linear(prec, pa_v, kPatchVec, wv, N * kPatches, 2 * kDimRgb, kPatchVec, ep, s))) return rc;
    }

    // ---------------- pair stage, chunk by chunk ----------------
    TokenSources ts{so_d, so_v, lso, cso, (const float*)(P + L.clspos), w->pos_embedding};
    float* x = (float*)(B + W.x);
    float* qkv = (float*)(B + W.qkv);
    for (int64_t r0 = 0; r0 < in->n_pairs; r0 += W.chunk) {
        const int64_t rc_pairs = (in->n_pairs - r0 < W.chunk) ? (in->n_pairs - r0) : W.chunk;
        const int M = (int)(rc_pairs * kTokens);
        ActBuf xn = enc_act_at(B, W.xn, prec, (size_t)M * kDim);
        ActBuf hb = enc_act_at(B, W.h, prec, (size_t)M * kMlp);
        const int R = (int)rc_pairs;
        // LayerNorm fusion (tensor-core modes): the epilogues that produce x (to_out, FF2) also write it in operand format
        // with its row statistics, and the next Linear (FF1, the next layer's to_qkv) runs on those raw rows with the
        // LayerNorm weight folded into its weight (gemm_tc2.cu EPI_*_LN) — the LayerNorm pass over x and its normalised
        // copy never exist.  The token kernel emits its rows the same way, so layer 0 is no exception.
        // Only in the modes that carry ~16 bits per operand (bf16x3, f16c8): on raw rows the products are rounded relative
        // to |x|, not |x - mean|, and the single-product modes (bf16, f16) have no headroom for that next to their tolerance.
        const bool fuse_ln = ln_fusion_enabled() && prec_two_arrays(prec);
        ActBuf xo = fuse_ln ? enc_act_at(B, W.xo, prec, (size_t)M * kDim) : ActBuf();
        float2* ln_parts = fuse_ln ? (float2*)(B + W.ln_parts) : nullptr;
        float2* ln_stats = fuse_ln ? (float2*)(B + W.ln_stats) : nullptr;
        bool x_ops_ready = false;     // xo / ln_stats hold the current x
        // an epilogue that produces x.  Un-fused: x = acc + bias + x in fp32.  Fused: the result goes out in operand
        // format with its statistics partials; once xo holds x (after layer 0's to_out) the residual is read from xo too and
        // the fp32 copy is not written any more — the residual stream lives in operand format only.
        auto emit_x = [&](GemmEpilogue& e) {
            if (!fuse_ln) return;
            e.out.hi = xo.hi;
            e.out.lo = xo.lo;
            e.out.fmt = xo.fmt;
            e.stats_partials = ln_parts;
            if (x_ops_ready && resop_enabled()) {
                e.residual = nullptr;
                e.out.f32 = nullptr;
                e.res_op = xo.out();
            }
        };
        auto set_ln_in = [&](GemmEpilogue& e) {
            if (stats_in_epilogue()) {
                e.ln_parts = ln_parts;
                e.ln_parts_rows = M;
            } else {
                e.ln_stats = ln_stats;
            }
        };
        auto finish_x = [&]() -> int {
            if (!fuse_ln) return VETO_OK;
            set_tag(TAG_LN);
            x_ops_ready = true;
            if (!stats_in_epilogue()) return ln_stats_finalize(ln_parts, kDim / 64, M, ln_stats, s);
            return VETO_OK;   // the consuming epilogues reduce the partials themselves
        };
        // the fp32 rows of the fresh tokens are only needed without the fusion, for the fp32 residual stream, or on request
        const bool tok_f32 = !fuse_ln || !resop_enabled() || out->tokens != nullptr;
        set_tag(TAG_TOKENS);
        if ((rc = build_tokens(ts, in->subj + r0, in->obj + r0, rc_pairs, tok_f32 ? x : nullptr, xo.out(), ln_parts, s))) return rc;
        if (out->tokens)
            VETO_CUDA(cudaMemcpyAsync(out->tokens + (size_t)r0 * kTokens * kDim, x, sizeof(float) * (size_t)M * kDim,
                                      cudaMemcpyDeviceToDevice, s));
        if ((rc = finish_x())) return rc;
        for (int l = 0; l + 1 < cfg->layers; ++l) {
            // x = to_out(softmax(q k^T * scale) v) + x      (PreNorm + Attention, model_veto.py:18-19,86-96)
            GemmEpilogue e1;
            e1.ldc = 3 * kDim;
            // q, k, v leave the GEMM as bf16 hi + lo arrays (the bytes of the fp32 buffer) for attention_split.cu in every
            // mode whose attention core runs the split products; the single bf16 product and fp32 keep the fp32 buffer
            const bool split_qkv = split_attention_enabled() && prec != VETO_PREC_FP32 && prec != VETO_PREC_BF16;
            __nv_bfloat16* qkv_hi = (__nv_bfloat16*)qkv;
            __nv_bfloat16* qkv_lo = qkv_hi + (size_t)M * 3 * kDim;
            const bool item_qkv = split_qkv && qkv_item_layout_enabled();
            if (split_qkv) {
                e1.out.hi = qkv_hi;
                e1.out.lo = qkv_lo;
                e1.out.fmt = FMT_BF16;
                e1.qkv_item_layout = item_qkv;
            } else {
                e1.out.f32 = qkv;
            }
            set_tag(TAG_QKV);
            if (x_ops_ready) {
                const float* cq = (const float*)(P + L.c_qkv[l]);
                set_ln_in(e1);
                e1.ln_c1 = cq;
                e1.bias = cq + 3 * kDim;
                WRef wq{nullptr, bf(P, L.qkvf_hi[l]), bf(P, L.qkvf_lo[l])};
                if ((rc = linear(prec, xo, kDim, wq, M, 3 * kDim, kDim, e1, s))) return rc;
            } else {
                set_tag(TAG_LN);
                if ((rc = layernorm_rows(x, kDim, w->ln1_w[l], w->ln1_b[l], M, xn.out(), s))) return rc;
                WRef wq{w->qkv_w[l], bf(P, L.qkv_hi[l]), bf(P, L.qkv_lo[l])};
                set_tag(TAG_QKV);
                if ((rc = linear(prec, xn, kDim, wq, M, 3 * kDim, kDim, e1, s))) return rc;
            }
            set_tag(TAG_ATT);
            if (split_qkv) rc = attention_seq_split(qkv_hi, qkv_lo, rc_pairs, xn.out(), item_qkv, s);
            else rc = attention_seq(qkv, rc_pairs, xn.out(), s);
            if (rc) return rc;
            GemmEpilogue e2;
            e2.bias = w->out_b[l];
            e2.residual = x;
            e2.out.f32 = x;
            e2.ldc = kDim;
            emit_x(e2);
            WRef wo{w->out_w[l], bf(P, L.out_hi[l]), bf(P, L.out_lo[l])};
            set_tag(TAG_OUT);
            if ((rc = linear(prec, xn, kDim, wo, M, kDim, kDim, e2, s))) return rc;
            if ((rc = finish_x())) return rc;
            // x = W2 gelu(W1 LN(x) + b1) + b2 + x           (PreNorm + FeedForward, model_veto.py:20,134-146)
            GemmEpilogue e3;
            e3.act = ACT_GELU;
            e3.out = hb.out();
            e3.ldc = kMlp;
            if (fuse_ln) {
                const float* cf = (const float*)(P + L.c_ff1[l]);
                set_ln_in(e3);
                e3.ln_c1 = cf;
                e3.bias = cf + kMlp;
                WRef w1{nullptr, bf(P, L.ff1f_hi[l]), bf(P, L.ff1f_lo[l])};
                set_tag(TAG_FF1);
                if ((rc = linear(prec, xo, kDim, w1, M, kMlp, kDim, e3, s))) return rc;
            } else {
                set_tag(TAG_LN);
                if ((rc = layernorm_rows(x, kDim, w->ln2_w[l], w->ln2_b[l], M, xn.out(), s))) return rc;
                e3.bias = w->ff1_b[l];
                WRef w1{w->ff1_w[l], bf(P, L.ff1_hi[l]), bf(P, L.ff1_lo[l])};
                set_tag(TAG_FF1);
                if ((rc = linear(prec, xn, kDim, w1, M, kMlp, kDim, e3, s))) return rc;
            }
            GemmEpilogue e4;
            e4.bias = w->ff2_b[l];
            e4.residual = x;
            e4.out.f32 = x;
            e4.ldc = kDim;
            emit_x(e4);
            WRef w2{w->ff2_w[l], bf(P, L.ff2_hi[l]), bf(P, L.ff2_lo[l])};
            set_tag(TAG_FF2);
            if ((rc = linear(prec, hb, kMlp, w2, M, kDim, kMlp, e4, s))) return rc;
            if ((rc = finish_x())) return rc;
        }
        {
            // Last layer: only x[:,0] leaves the encoder (model_veto.py:25), so only the CLS row needs a query, an
            // attention output, the out-projection and the feed-forward; K and V still need every token.  The CLS
            // rows are addressed in place with a row stride of 19*576 (TMA tensor map / lda), results stay compact.
            const int l = cfg->layers - 1;
            const size_t woff = (size_t)kDim * kDim;  // rows [576, 1728) of to_qkv.weight = K and V projections
            float* q_cls = (float*)(B + W.q_cls);
            float* x_cls = (float*)(B + W.x_cls);
            GemmEpilogue e1;
            e1.out.f32 = qkv + kDim;
            e1.ldc = 3 * kDim;
            GemmEpilogue eq;
            eq.out.f32 = q_cls;
            eq.ldc = kDim;
            if (x_ops_ready) {   // LayerNorm fused (see above): K, V of every row, the query of the CLS rows (row stride 19)
                const float* cq = (const float*)(P + L.c_qkv[l]);
                set_ln_in(e1);
                set_ln_in(eq);
                e1.ln_c1 = cq + kDim;
                e1.bias = cq + 3 * kDim + kDim;
                eq.ln_c1 = cq;
                eq.bias = cq + 3 * kDim;
                eq.ln_row_stride = kTokens;
                WRef wkv{nullptr, bf(P, L.qkvf_hi[l]) + woff, bf(P, L.qkvf_lo[l]) ? bf(P, L.qkvf_lo[l]) + woff : nullptr};
                WRef wq{nullptr, bf(P, L.qkvf_hi[l]), bf(P, L.qkvf_lo[l])};
                set_tag(TAG_QKV);
                if ((rc = linear(prec, xo, kDim, wkv, M, 2 * kDim, kDim, e1, s))) return rc;
                if ((rc = linear(prec, xo, kTokens * kDim, wq, R, kDim, kDim, eq, s))) return rc;
            } else {
                set_tag(TAG_LN);
                if ((rc = layernorm_rows(x, kDim, w->ln1_w[l], w->ln1_b[l], M, xn.out(), s))) return rc;
                WRef wkv{w->qkv_w[l] + woff, bf(P, L.qkv_hi[l]) ? bf(P, L.qkv_hi[l]) + woff : nullptr,
                         bf(P, L.qkv_lo[l]) ? bf(P, L.qkv_lo[l]) + woff : nullptr};
                set_tag(TAG_QKV);
                if ((rc = linear(prec, xn, kDim, wkv, M, 2 * kDim, kDim, e1, s))) return rc;
                WRef wq{w->qkv_w[l], bf(P, L.qkv_hi[l]), bf(P, L.qkv_lo[l])};
                if ((rc = linear(prec, xn, kTokens * kDim, wq, R, kDim, kDim, eq, s))) return rc;
            }
            ActBuf ao = enc_act_at(B, W.xn, prec, (size_t)R * kDim);  // xn is free once K, V and q exist
            set_tag(TAG_ATT);
            if ((rc = attention_cls(q_cls, qkv, rc_pairs, ao.out(), s))) return rc;
            GemmEpilogue e2;
            e2.bias = w->out_b[l];
            if (x_ops_ready && resop_enabled()) e2.res_op = xo.out();     // the residual stream is in operand format (see emit_x)
            else e2.residual = x;
            e2.ldr = kTokens * kDim;
            e2.out.f32 = x_cls;
            e2.ldc = kDim;
            WRef wo{w->out_w[l], bf(P, L.out_hi[l]), bf(P, L.out_lo[l])};
            set_tag(TAG_OUT);
            if ((rc = linear(prec, ao, kDim, wo, R, kDim, kDim, e2, s))) return rc;
            ActBuf xn_cls = enc_act_at(B, W.xn, prec, (size_t)R * kDim);
            set_tag(TAG_LN);
            if ((rc = layernorm_rows(x_cls, kDim, w->ln2_w[l], w->ln2_b[l], R, xn_cls.out(), s))) return rc;
            ActBuf h_cls = enc_act_at(B, W.h, prec, (size_t)R * kMlp);
            GemmEpilogue e3;
            e3.bias = w->ff1_b[l];
            e3.act = ACT_GELU;
            e3.out = h_cls.out();
            e3.ldc = kMlp;
            WRef w1{w->ff1_w[l], bf(P, L.ff1_hi[l]), bf(P, L.ff1_lo[l])};
            set_tag(TAG_FF1);
            if ((rc = linear(prec, xn_cls, kDim, w1, R, kMlp, kDim, e3, s))) return rc;
            GemmEpilogue e4;
            e4.bias = w->ff2_b[l];
            e4.residual = x_cls;
            e4.out.f32 = x_cls;
            e4.ldc = kDim;
            WRef w2{w->ff2_w[l], bf(P, L.ff2_hi[l]), bf(P, L.ff2_lo[l])};
            set_tag(TAG_FF2);
            if ((rc = linear(prec, h_cls, kMlp, w2, R, kDim, kMlp, e4, s))) return rc;
            // rel_out on x[:,0] (model_veto.py:25; roi_relation_predictors.py:4125): always fp32 FMA
            GemmEpilogue ec;
            ec.bias = w->rel_out_b;
            ec.out.f32 = out->rel_logits + (size_t)r0 * cfg->num_out;
            ec.ldc = cfg->num_out;
            set_tag(TAG_CLS);
            if ((rc = gemm_simt(x_cls, kDim, w->rel_out_w, R, cfg->num_out, kDim, ec, s))) return rc;
            if (out->rel_features)
                VETO_CUDA(cudaMemcpyAsync(out->rel_features + (size_t)r0 * kDim, x_cls, sizeof(float) * (size_t)R * kDim,
                                          cudaMemcpyDeviceToDevice, s));
        }
    }
    set_tag(TAG_OTHER);
    if (in->freq_bias)
        if ((rc = add_freq_bias(out->rel_logits, cfg->num_out, in->freq_bias, in->labels, cfg->num_obj, in->subj, in->obj,
                                in->n_pairs, s)))
            return rc;
    return VETO_OK;
}

extern "C" int veto_profile_begin(veto_stream_t stream) {
    VETO_REQUIRE(!g_prof, VETO_ERR_ARG, "veto_profile_begin: already profiling");
    g_prof_stream = (cudaStream_t)stream;
    g_prof_events.clear();
    cudaEvent_t e;
    VETO_CUDA(cudaEventCreate(&e));
    VETO_CUDA(cudaEventRecord(e, g_prof_stream));
    g_prof_events.push_back({e, -1});
    g_prof = true;
    return VETO_OK;
}

extern "C" int veto_profile_end(double* ms_by_tag_host, int64_t* launches_by_tag_host) {
    VETO_REQUIRE(g_prof && ms_by_tag_host && launches_by_tag_host, VETO_ERR_ARG, "veto_profile_end: not profiling / NULL");
    g_prof = false;
    cudaError_t err = cudaEventSynchronize(g_prof_events.back().ev);
    FILE* dump = nullptr;  // VETO_PROFILE_DUMP=<path>: one line per launch (index, stage tag, microseconds), for diagnosis
    if (const char* path = getenv("VETO_PROFILE_DUMP")) dump = fopen(path, "a");
    for (size_t i = 1; i < g_prof_events.size() && err == cudaSuccess; ++i) {
        float ms = 0.f;
        err = cudaEventElapsedTime(&ms, g_prof_events[i - 1].ev, g_prof_events[i].ev);
        const int t = g_prof_events[i].tag;
        if (dump) fprintf(dump, "%zu %s %.1f\n", i, (t >= 0 && t < VETO_PROFILE_TAGS) ? kTagNames[t] : "?", ms * 1e3f);
        if (t >= 0 && t < VETO_PROFILE_TAGS) {
            ms_by_tag_host[t] += ms;
            launches_by_tag_host[t] += 1;
        }
    }
    if (dump) fclose(dump);
    for (auto& pe : g_prof_events) cudaEventDestroy(pe.ev);
    g_prof_events.clear();
    VETO_CUDA(err);
    return VETO_OK;
}

extern "C" const char* veto_profile_tag_name(int tag) { return (tag >= 0 && tag < VETO_PROFILE_TAGS) ? kTagNames[tag] : ""; }

// ------------------------------------------------------------------------------------------ test hooks
extern "C" int veto_test_gemm(const float* a_dev, const float* w_dev, const float* bias_dev, const float* residual_dev,
                              float* c_dev, int M, int N, int K, int act, int precision, void* scratch_dev,
                              size_t scratch_bytes, veto_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    // act bits: low byte = activation; 0x100 = write bf16 hi/lo outputs (into the scratch) instead of fp32;
    // 0x200 = operands were already split by a previous call with the same scratch (timing runs)
    const bool split_out = (act & 0x100) != 0, reuse = (act & 0x200) != 0;
    GemmEpilogue ep;
    ep.bias = bias_dev;
    ep.residual = residual_dev;
    ep.act = act & 0xff;
    ep.out.f32 = c_dev;
    ep.ldc = N;
    if (precision == VETO_PREC_FP32) return gemm_simt(a_dev, K, w_dev, M, N, K, ep, s);
    const size_t ae = (size_t)M * K, we = (size_t)N * K, ce = (size_t)M * N;
    VETO_REQUIRE(scratch_dev && scratch_bytes >= 4 * (ae + we) + (split_out ? 4 * ce : 0), VETO_ERR_WORKSPACE,
                 "veto_test_gemm: scratch too small");
    __nv_bfloat16* a_hi = (__nv_bfloat16*)scratch_dev;
    __nv_bfloat16* a_lo = a_hi + ae;
    __nv_bfloat16* w_hi = a_lo + ae;
    __nv_bfloat16* w_lo = w_hi + we;
    int rc;
    const bool c8 = prec_encoder_fmt(precision) == FMT_F16C8;
    if (!reuse) {
        if (c8) {
            SplitJob jobs[2] = {SplitJob{a_dev, a_hi, a_lo, ae, FMT_F16C8_ACT}, SplitJob{w_dev, w_hi, w_lo, we, FMT_F16C8}};
            if ((rc = pack_split_bf16_multi(jobs, 2, s))) return rc;
        } else {
            if ((rc = pack_split_bf16(a_dev, a_hi, a_lo, ae, s))) return rc;
            if ((rc = pack_split_bf16(w_dev, w_hi, w_lo, we, s))) return rc;
        }
    }
    if (split_out) {
        ep.out.f32 = nullptr;
        ep.out.hi = w_lo + we;
        ep.out.lo = prec_two_arrays(precision) ? ep.out.hi + ce : nullptr;
        ep.out.fmt = prec_encoder_fmt(precision);
    }
    GemmOperand A, W;
    A.hi = a_hi; A.lo = a_lo;
    W.hi = w_hi; W.lo = w_lo;
    return gemm_tc_auto(A, W, M, N, K, prec_encoder_passes(precision), ep, s);
}

extern "C" int veto_test_layernorm(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev, int64_t rows,
                                   veto_stream_t stream) {
    ActOut o;
    o.f32 = y_dev;
    return layernorm_rows(x_dev, kDim, w_dev, b_dev, rows, o, (cudaStream_t)stream);
}

extern "C" int veto_test_attention(const float* qkv_dev, float* out_dev, int64_t n_seq, veto_stream_t stream) {
    ActOut o;
    o.f32 = out_dev;
    return attention_seq(qkv_dev, n_seq, o, (cudaStream_t)stream);
}

extern "C" int veto_test_attention_tc(const float* qkv_dev, float* out_dev, void* scratch_dev, int64_t n_seq, int split,
                                      veto_stream_t stream) {
    VETO_REQUIRE(qkv_dev && out_dev && scratch_dev, VETO_ERR_ARG, "veto_test_attention_tc: NULL argument");
    ActOut o;
    o.f32 = out_dev;
    o.hi = (__nv_bfloat16*)scratch_dev;
    o.lo = split ? o.hi + (size_t)n_seq * kTokens * kDim : nullptr;
    return attention_tc(qkv_dev, n_seq, o, (cudaStream_t)stream);
}

extern "C" int veto_test_gemm_tn(const float* y_dev, const float* x_dev, float* out_dev, int rows, int Nw, int Kw, int precision,
                                 int split_k, const uint32_t* geometry_host, void* scratch_dev, size_t scratch_bytes,
                                 veto_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    VETO_REQUIRE(y_dev && x_dev && out_dev && scratch_dev, VETO_ERR_ARG, "veto_test_gemm_tn: NULL argument");
    VETO_REQUIRE(precision == VETO_PREC_BF16X3 || precision == VETO_PREC_BF16, VETO_ERR_ARG, "veto_test_gemm_tn: tensor-core modes only");
    const size_t ye = (size_t)rows * Nw, xe = (size_t)rows * Kw, oe = (size_t)Nw * Kw;
    const int slices = gemm_tn2_slices(rows, split_k);
    VETO_REQUIRE(scratch_bytes >= 4 * (ye + xe) + (slices > 1 ? 4 * oe * slices : 0), VETO_ERR_WORKSPACE,
                 "veto_test_gemm_tn: scratch too small");
    __nv_bfloat16* y_hi = (__nv_bfloat16*)scratch_dev;
    __nv_bfloat16* y_lo = y_hi + ye;
    __nv_bfloat16* x_hi = y_lo + ye;
    __nv_bfloat16* x_lo = x_hi + xe;
    float* partial = (float*)(x_lo + xe);
    int rc;
    if ((rc = pack_split_bf16(y_dev, y_hi, y_lo, ye, s))) return rc;
    if ((rc = pack_split_bf16(x_dev, x_hi, x_lo, xe, s))) return rc;
    GemmOperand Y, X;
    Y.hi = y_hi; Y.lo = y_lo;
    X.hi = x_hi; X.lo = x_lo;
    const int passes = precision == VETO_PREC_BF16X3 ? 3 : 1;
    if ((rc = gemm_tn2(Y, X, Nw, Kw, rows, passes, slices > 1 ? partial : out_dev, Kw, slices, oe, s, geometry_host))) return rc;
    if (slices > 1) rc = splitk_reduce(partial, slices, oe, oe, out_dev, s);
    return rc;
}
