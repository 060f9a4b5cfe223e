// veto_relation_train_step: VETOPredictor.forward in train() mode plus the backward pass from rel_loss
// (roi_relation_predictors.py:4074-4136; tools/relation_train_net.py:451-452), orchestrated over the stage kernels.
//
// Forward keeps what the backward needs (per layer: the residual stream before each PreNorm, the normalised rows,
// qkv, the attention output, the FeedForward pre-activation and activation).  Backward runs every Linear's two
// gradients on the same tcgen05 GEMM kernel as the forward:
//   dX = dY @ W          ->  gemm(A = dY [M,N],    W' = W^T  [K,N])      (W^T packed once per step)
//   dW = dY^T @ X        ->  gemm_tn2(dY [M,N], X [M,K])                  (both read in place as MN-major tcgen05
//                                                                         operands, split-K over the M rows; fp32
//                                                                         mode: transposes + the SIMT GEMM)
// so the bf16x3 mode keeps its fp32-grade products in the gradients too.  Declarations: include/veto_b200.h.
#include <string.h>

#include "api_internal.cuh"
#include "train.cuh"

namespace veto {
namespace {

constexpr int kMaxSplit = 16;
inline int64_t pad64(int64_t v) { return (v + 63) / 64 * 64; }
inline size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

struct LayerSave {
    size_t xn1, qkv, ao, x_mid, xn2, h, h_pre;
};
struct LayerWT {
    size_t qkv, out, ff1, ff2;  // transposed weights (operand format)
};

struct TrainLayout {
    // box level (forward)
    size_t pos, emb, lso, cso, pa_d, pa_v, so_d, so_v, bn_stats;
    // saved activations
    size_t x_in[VETO_MAX_LAYERS + 1];
    LayerSave L[VETO_MAX_LAYERS];
    size_t logits;
    // backward temporaries
    size_t dlogits, ce_scratch, tc1, tc2, wcT, wc_partial, pos_partial, box_partial;
    size_t dx, tmp, a576, a1728, T1, T2, splitk, ln_partial, colsum_scratch;
    LayerWT WT[VETO_MAX_LAYERS];
    size_t d2T, v2T, loc2T, cls2T;
    size_t d_so_d, d_so_v, d_lso, d_cso, d_pos, d_emb, d_pa, a_box, tb1, tb2;
    size_t g_w_d2, g_w_v2, g_w_loc2, g_w_cls2;
    size_t total;
    int64_t M, Mp, Kb, Rp, Nb;
};

// VETO_TRAIN_RECOMPUTE=1: the LayerNorm outputs and the GELU output of a layer are not kept for the backward pass (one
// buffer each, shared by the layers) but re-computed from x_in / x_mid / the FF1 pre-activation right before the weight
// gradients that read them: 9216 of the 27 648 saved bytes per token row and layer less (16.0 -> 11.3 GB at the 4560-pair
// step of BASELINE configs[1]) for two LayerNorm passes and one GELU pass per layer more.  Off by default: the step
// is 1.8 ms (4 %) slower.  Read by the workspace-size query and the step alike.
bool train_recompute() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("VETO_TRAIN_RECOMPUTE");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on != 0;
}

TrainLayout train_layout(const veto_config& c, int32_t n_boxes, int64_t n_pairs) {
    TrainLayout T{};
    Carver k;
    const int prec = c.precision;
    const size_t N = (size_t)(n_boxes > 0 ? n_boxes : 1);
    const size_t R = (size_t)(n_pairs > 0 ? n_pairs : 1);
    const size_t M = R * kTokens;
    T.M = (int64_t)M;
    T.Mp = pad64((int64_t)M);
    T.Kb = pad64((int64_t)N * kPatches);
    T.Rp = pad64((int64_t)R);
    T.Nb = pad64((int64_t)N);
    const size_t f = sizeof(float);
    T.pos = k.take(f * N * kPosDim);
    T.emb = k.take(f * N * kEmbDim);
    T.lso = k.take(f * N * 2 * kDim);
    T.cso = k.take(f * N * 2 * kDim);
    T.pa_d = k.take(act_bytes(prec, N * kPatches * kPatchVec));
    T.pa_v = k.take(act_bytes(prec, N * kPatches * kPatchVec));
    T.so_d = k.take(f * N * kPatches * 2 * kDimDepth);
    T.so_v = k.take(f * N * kPatches * 2 * kDimRgb);
    T.bn_stats = k.take(f * 8);
    for (int l = 0; l <= c.layers; ++l) T.x_in[l] = k.take(f * M * kDim);
    const bool recompute = train_recompute();
    for (int l = 0; l < c.layers; ++l) {
        const bool own = !recompute || l == 0;   // with re-computation every layer uses layer 0's three buffers
        T.L[l].xn1 = own ? k.take(act_bytes(prec, M * kDim)) : T.L[0].xn1;
        T.L[l].qkv = k.take(f * M * 3 * kDim);
        T.L[l].ao = k.take(act_bytes(prec, M * kDim));
        T.L[l].x_mid = k.take(f * M * kDim);
        T.L[l].xn2 = own ? k.take(act_bytes(prec, M * kDim)) : T.L[0].xn2;
        T.L[l].h = own ? k.take(act_bytes(prec, M * kMlp)) : T.L[0].h;
        T.L[l].h_pre = k.take(f * M * kMlp);
    }
    T.logits = k.take(f * R * c.num_out);
    T.dlogits = k.take(f * R * c.num_out);
    T.ce_scratch = k.take(f * (2 * R + 16));
    T.tc1 = k.take(f * (size_t)c.num_out * T.Rp);
    T.tc2 = k.take(f * (size_t)kDim * T.Rp);
    T.wcT = k.take(f * (size_t)kDim * c.num_out);
    T.wc_partial = k.take(f * (size_t)kMaxSplit * kDim * c.num_out);
    T.pos_partial = k.take(f * pos_embed_bwd_scratch_floats());
    T.box_partial = k.take(f * (size_t)kMaxSplit * max_sz((size_t)T.Nb * kEmbDim, (size_t)2 * kDim * kEmbDim));
    T.dx = k.take(f * M * kDim);
    T.tmp = k.take(f * M * kDim);
    T.a576 = k.take(act_bytes(prec, M * kDim));
    T.a1728 = k.take(act_bytes(prec, M * 3 * kDim));
    // transposed operands: only the fp32 (SIMT) mode needs them, the tensor-core modes read dY and X in place
    const bool simt = prec == VETO_PREC_FP32;
    T.T1 = k.take(simt ? f * max_sz((size_t)3 * kDim * T.Mp, (size_t)2 * kDimDepth * T.Kb) : 0);
    T.T2 = k.take(simt ? f * max_sz((size_t)kMlp * T.Mp, (size_t)kPatchVec * T.Kb) : 0);
    T.splitk = k.take(simt ? 0 : f * (size_t)kMaxSplit * max_sz((size_t)3 * kDim * kDim, (size_t)2 * kDimDepth * kPatchVec));
    T.ln_partial = k.take(f * (size_t)ln_bwd_blocks((int64_t)M) * 3 * kDim);
    T.colsum_scratch = k.take(f * (colsum_scratch_floats(3 * kDim) + 3 * kDim));
    for (int l = 0; l < c.layers; ++l) {
        T.WT[l].qkv = k.take(act_bytes(prec, (size_t)3 * kDim * kDim));
        T.WT[l].out = k.take(act_bytes(prec, (size_t)kDim * kDim));
        T.WT[l].ff1 = k.take(act_bytes(prec, (size_t)kMlp * kDim));
        T.WT[l].ff2 = k.take(act_bytes(prec, (size_t)kMlp * kDim));
    }
    T.d2T = k.take(act_bytes(prec, (size_t)2 * kDimDepth * kPatchVec));
    T.v2T = k.take(act_bytes(prec, (size_t)2 * kDimRgb * kPatchVec));
    T.loc2T = k.take(f * (size_t)2 * kDim * kPosDim);
    T.cls2T = k.take(f * (size_t)2 * kDim * kEmbDim);
    T.d_so_d = k.take(f * N * kPatches * 2 * kDimDepth);
    T.d_so_v = k.take(f * N * kPatches * 2 * kDimRgb);
    T.d_lso = k.take(f * N * 2 * kDim);
    T.d_cso = k.take(f * N * 2 * kDim);
    T.d_pos = k.take(f * N * kPosDim);
    T.d_emb = k.take(f * N * kEmbDim);
    T.d_pa = k.take(f * N * kPatches * kPatchVec);
    T.a_box = k.take(act_bytes(prec, N * kPatches * 2 * kDimDepth));
    T.tb1 = k.take(f * (size_t)2 * kDim * T.Nb);
    T.tb2 = k.take(f * (size_t)kEmbDim * T.Nb);
    T.g_w_d2 = k.take(f * (size_t)2 * kDimDepth * kPatchVec);
    T.g_w_v2 = k.take(f * (size_t)2 * kDimRgb * kPatchVec);
    T.g_w_loc2 = k.take(f * (size_t)2 * kDim * kPosDim);
    T.g_w_cls2 = k.take(f * (size_t)2 * kDim * kEmbDim);
    T.total = k.off;
    return T;
}

struct Ctx {
    int prec;
    cudaStream_t s;
    char* B;
    const TrainLayout* T;
    float* colsum_scratch;
    float* splitk;

    ActBuf act(size_t off, size_t elems) const { return act_at(B, off, prec, elems); }
    float* f32(size_t off) const { return (float*)(B + off); }

    // C[M,N] = epilogue(A[M,K] @ W[N,K]^T), operands in the storage format of the precision mode
    int mm(const ActBuf& A, int lda, const ActBuf& W, int M, int N, int K, const GemmEpilogue& ep) const {
        return linear(prec, A, lda, WRef{W.f32, W.hi, W.lo}, M, N, K, ep, s);
    }
    // plain fp32 C[M,N] = A[M,K] @ W[N,K]^T for the small per-box GEMMs of the backward (few output tiles, long K): split-K
    // over up to kMaxSplit slices of at least 64 columns, fixed-order reduction
    int mm_small(const float* A, int lda, const float* W, int M, int N, int K, float* out, int ldc) const {
        GemmEpilogue ep;
        ep.ldc = ldc;
        int want = K / 64 < kMaxSplit ? K / 64 : kMaxSplit;
        if (N <= 128 && K >= 64 && M >= 256) want = 1;  // gemm_simt's skinny-output kernel already fills the machine
        const int slices = want > 1 ? gemm_simt_slices(K, want) : 1;
        if (slices <= 1) {
            ep.out.f32 = out;
            return gemm_simt(A, lda, W, M, N, K, ep, s);
        }
        float* partial = f32(T->box_partial);
        ep.out.f32 = partial;
        ep.split_k = want;
        ep.split_stride = (size_t)M * ldc;
        int rc = gemm_simt(A, lda, W, M, N, K, ep, s);
        if (rc) return rc;
        return splitk_reduce(partial, slices, (size_t)M * ldc, (size_t)M * ldc, out, s);
    }
    // gW[Nw,Kw] = dY[rows,Nw]^T @ X[rows,Kw]: the weight gradient of y = x W^T, operands row-major as the backward holds them
    int wgrad(const ActBuf& dY, int ldy, const ActBuf& X, int ldx, int64_t rows, int Nw, int Kw, float* gW) const {
        set_tag(TAG_BWD_WGRAD);
        const int rc = wgrad_impl(dY, ldy, X, ldx, rows, Nw, Kw, gW);
        set_tag(TAG_BWD_GEMM);
        return rc;
    }
    int wgrad_impl(const ActBuf& dY, int ldy, const ActBuf& X, int ldx, int64_t rows, int Nw, int Kw, float* gW) const {
        if (prec == VETO_PREC_FP32) {
            const int64_t Kp = pad64(rows);
            float* t1 = f32(T->T1);
            float* t2 = f32(T->T2);
            ActOut o1, o2;
            o1.f32 = t1; o2.f32 = t2;
            int rc = transpose_f32(dY.f32, ldy, rows, Nw, false, DropSpec(), 0, o1, Kp, Kp, ActOut(), 0, s);
            if (rc) return rc;
            if ((rc = transpose_f32(X.f32, ldx, rows, Kw, false, DropSpec(), 0, o2, Kp, Kp, ActOut(), 0, s))) return rc;
            GemmEpilogue ep;
            ep.ldc = Kw;
            ep.out.f32 = gW;
            return gemm_simt(t1, (int)Kp, t2, Nw, Kw, (int)Kp, ep, s);
        }
        // split-K: pick the slice count whose tile total fills whole waves of CTA pairs best (>= 8 K blocks per slice)
        const int pairs = num_sms() / 2;
        const int max_s = (int)((rows + 511) / 512) < kMaxSplit ? (int)((rows + 511) / 512) : kMaxSplit;
        int best = 1;
        double best_eff = 0.0;
        for (int want = 1; want <= (max_s > 1 ? max_s : 1); ++want) {
            const int sl = gemm_tn2_slices((int)rows, want);
            const double eff = gemm_tn2_efficiency(Nw, Kw, sl, pairs);
            if (eff > best_eff + 0.02) {
                best_eff = eff;
                best = sl;
            }
        }
        GemmOperand A, B;
        A.hi = dY.hi; A.lo = dY.lo; A.ld = ldy;
        B.hi = X.hi; B.lo = X.lo; B.ld = ldx;
        const size_t n = (size_t)Nw * Kw;
        int rc = gemm_tn2(A, B, Nw, Kw, (int)rows, prec == VETO_PREC_BF16X3 ? 3 : 1, best > 1 ? splitk : gW, Kw, best, n, s);
        if (rc) return rc;
        if (best > 1) rc = splitk_reduce(splitk, best, n, n, gW, s);
        return rc;
    }
    // fp32 -> the operand format of the precision mode (optionally through a dropout mask)
    int to_operand(const float* src, size_t n, const DropSpec& drop, const ActBuf& dst) const {
        return convert_act(src, n, drop, dst.out(), s);
    }
    int bias_grad(const ActIn& src, int64_t ld, int64_t rows, int cols, float* out) const {
        return colsum(src, ld, rows, cols, colsum_scratch, out, false, s);
    }
};

ActIn as_in(const ActBuf& b) {
    ActIn a;
    a.f32 = b.f32; a.hi = b.hi; a.lo = b.lo;
    return a;
}
ActIn f32_in(const float* p) {
    ActIn a;
    a.f32 = p;
    return a;
}
uint64_t sub_seed(uint64_t seed, uint64_t k) { return drop_hash(seed ^ 0xA5A5A5A5A5A5A5A5ull, k); }

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" size_t veto_train_workspace_bytes(const veto_config* cfg, int32_t n_boxes, int64_t n_pairs) {
    if (check_config(cfg)) return 0;
    return train_layout(*cfg, n_boxes, n_pairs).total;
}

#define RC(expr)                  \
    do {                          \
        if ((rc = (expr))) return rc; \
    } while (0)

extern "C" int veto_relation_train_step(const veto_config* cfg, const veto_weights* w, const void* packed_dev,
                                        const veto_inputs* in, const veto_train_inputs* tin, const veto_grads* g,
                                        const veto_train_outputs* out, void* workspace_dev, size_t workspace_bytes,
                                        veto_stream_t stream) {
    int rc = check_config(cfg);
    if (rc) return rc;
    VETO_REQUIRE(w && packed_dev && in && tin && g && out && workspace_dev, VETO_ERR_ARG, "veto_relation_train_step: NULL argument");
    VETO_REQUIRE(in->n_boxes > 0 && in->n_pairs > 0, VETO_ERR_ARG, "veto_relation_train_step: needs at least one box and one pair");
    VETO_REQUIRE(in->boxes && in->roi_rgb && in->roi_depth && in->subj && in->obj && (in->labels || in->obj_logits),
                 VETO_ERR_ARG, "veto_relation_train_step: missing input pointer");
    VETO_REQUIRE(!in->freq_bias, VETO_ERR_UNSUPPORTED, "the frequency-bias epilogue has no training branch (VETO never uses it)");
    if (tin->n_heads > 1) {
        VETO_REQUIRE(tin->head_offsets && tin->head_labels && !tin->class_weight, VETO_ERR_ARG,
                     "veto_relation_train_step: group heads need head_offsets and head_labels (and take no class weight)");
        VETO_REQUIRE(tin->head_offsets[0] == 0 && tin->head_offsets[tin->n_heads] == cfg->num_out, VETO_ERR_ARG,
                     "veto_relation_train_step: head_offsets must cover the num_out logit columns");
        for (int k = 0; k < tin->n_heads; ++k)
            VETO_REQUIRE(tin->head_offsets[k + 1] > tin->head_offsets[k], VETO_ERR_ARG, "veto_relation_train_step: empty head %d", k);
    }
    VETO_REQUIRE((tin->rel_labels || tin->n_heads > 1) && tin->rel_offsets && tin->box_offsets && tin->n_images > 0, VETO_ERR_ARG,
                 "veto_relation_train_step: rel_labels / rel_offsets / box_offsets missing");
    VETO_REQUIRE(out->loss, VETO_ERR_ARG, "veto_relation_train_step: loss output missing");
    VETO_REQUIRE(tin->p_pos_dropout >= 0.f && tin->p_pos_dropout < 1.f && tin->p_emb_dropout >= 0.f && tin->p_emb_dropout < 1.f &&
                     tin->p_attn_dropout >= 0.f && tin->p_attn_dropout < 1.f,
                 VETO_ERR_ARG, "dropout probabilities must be in [0, 1)");
    VETO_REQUIRE(!g->bn_mean && !g->bn_var, VETO_ERR_ARG, "running statistics have no gradient: grads.bn_mean / bn_var must be NULL");
    VETO_REQUIRE(g->obj_embed && g->class_proj_w && g->class_proj_b && g->bn_weight && g->bn_bias && g->pos_w && g->pos_b &&
                     g->loc_proj_w && g->loc_proj_b && g->cls_token && g->pos_embedding && g->proj_d_w && g->proj_d_b &&
                     g->proj_v_w && g->proj_v_b && g->rel_out_w && g->rel_out_b,
                 VETO_ERR_ARG, "veto_relation_train_step: a gradient pointer is NULL");
    for (int l = 0; l < cfg->layers; ++l)
        VETO_REQUIRE(g->ln1_w[l] && g->ln1_b[l] && g->qkv_w[l] && g->out_w[l] && g->out_b[l] && g->ln2_w[l] && g->ln2_b[l] &&
                         g->ff1_w[l] && g->ff1_b[l] && g->ff2_w[l] && g->ff2_b[l],
                     VETO_ERR_ARG, "veto_relation_train_step: a gradient pointer of layer %d is NULL", l);
    const int prec = cfg->precision;
    VETO_REQUIRE(prec <= VETO_PREC_BF16, VETO_ERR_UNSUPPORTED,
                 "veto_relation_train_step: f16c8 / f16 are inference modes; train with bf16x3 / bf16 / fp32 (the host layer maps them)");
    const PackedLayout L = packed_layout(*cfg);
    const TrainLayout T = train_layout(*cfg, in->n_boxes, in->n_pairs);
    VETO_REQUIRE(workspace_bytes >= T.total, VETO_ERR_WORKSPACE, "veto_relation_train_step: workspace %zu < %zu bytes",
                 workspace_bytes, T.total);
    VETO_REQUIRE(T.Mp < (1ll << 31) / 4, VETO_ERR_UNSUPPORTED, "too many pairs for one training step (%lld)", (long long)in->n_pairs);
    cudaStream_t s = (cudaStream_t)stream;
    const char* P = (const char*)packed_dev;
    char* B = (char*)workspace_dev;
    const int N = in->n_boxes;
    const int R = (int)in->n_pairs;
    const int M = (int)T.M;
    const int C = cfg->num_out;
    const int NL = cfg->layers;
    Ctx X{prec, s, B, &T, (float*)(B + T.colsum_scratch), (float*)(B + T.splitk)};
    const DropSpec drop_pos = make_drop(tin->p_pos_dropout, sub_seed(tin->seed, 1));
    const DropSpec drop_emb = make_drop(tin->p_emb_dropout, sub_seed(tin->seed, 2));

    // =====================================================================================  forward (training mode)
    float* pos = X.f32(T.pos);
    float* emb = X.f32(T.emb);
    float* lso = X.f32(T.lso);
    float* cso = X.f32(T.cso);
    float* so_d = X.f32(T.so_d);
    float* so_v = X.f32(T.so_v);
    float* bn_stats = X.f32(T.bn_stats);
    const size_t pe = (size_t)N * kPatches * kPatchVec;
    ActBuf pa_d = X.act(T.pa_d, pe), pa_v = X.act(T.pa_v, pe);
    set_tag(TAG_BOX);
    RC(bn_batch_stats(in->boxes, N, tin->bn_momentum, bn_stats, tin->bn_running_mean, tin->bn_running_var, s));
    RC(box_embed(in->boxes, in->labels, in->obj_logits, cfg->num_obj, N, *w, pos, emb, s, bn_stats, drop_pos));
    {
        GemmEpilogue ep;
        ep.bias = (const float*)(P + L.b_loc2);
        ep.out.f32 = lso;
        ep.ldc = 2 * kDim;
        RC(gemm_simt(pos, kPosDim, (const float*)(P + L.w_loc2), N, 2 * kDim, kPosDim, ep, s));
        ep.bias = (const float*)(P + L.b_cls2);
        ep.out.f32 = cso;
        RC(gemm_simt(emb, kEmbDim, (const float*)(P + L.w_cls2), N, 2 * kDim, kEmbDim, ep, s));
        RC(patchify(in->roi_depth, N, pa_d.out(), s));
        RC(patchify(in->roi_rgb, N, pa_v.out(), s));
        GemmEpilogue e2;
        e2.bias = (const float*)(P + L.b_d2);
        e2.out.f32 = so_d;
        e2.ldc = 2 * kDimDepth;
        WRef wd{(const float*)(P + L.w_d2), bf(P, L.d2_hi), bf(P, L.d2_lo)};
        RC(linear(prec, pa_d, kPatchVec, wd, N * kPatches, 2 * kDimDepth, kPatchVec, e2, s));
        e2.bias = (const float*)(P + L.b_v2);
        e2.out.f32 = so_v;
        e2.ldc = 2 * kDimRgb;
        WRef wv{(const float*)(P + L.w_v2), bf(P, L.v2_hi), bf(P, L.v2_lo)};
        RC(linear(prec, pa_v, kPatchVec, wv, N * kPatches, 2 * kDimRgb, kPatchVec, e2, s));
    }
    float* x_in[VETO_MAX_LAYERS + 1];
    for (int l = 0; l <= NL; ++l) x_in[l] = X.f32(T.x_in[l]);
    {
        TokenSources ts{so_d, so_v, lso, cso, (const float*)(P + L.clspos), w->pos_embedding};
        set_tag(TAG_TOKENS);
        RC(build_tokens(ts, in->subj, in->obj, R, x_in[0], ActOut(), nullptr, s));
        RC(dropout_inplace(x_in[0], (size_t)M * kDim, drop_emb, s));  // Transformer.pos_drop (model_veto.py:63)
    }
    for (int l = 0; l < NL; ++l) {
        const LayerSave& S = T.L[l];
        ActBuf xn1 = X.act(S.xn1, (size_t)M * kDim), ao = X.act(S.ao, (size_t)M * kDim), xn2 = X.act(S.xn2, (size_t)M * kDim);
        ActBuf hb = X.act(S.h, (size_t)M * kMlp);
        float* qkv = X.f32(S.qkv);
        float* x_mid = X.f32(S.x_mid);
        set_tag(TAG_LN);
        RC(layernorm_rows(x_in[l], kDim, w->ln1_w[l], w->ln1_b[l], M, xn1.out(), s));
        GemmEpilogue e1;
        e1.out.f32 = qkv;
        e1.ldc = 3 * kDim;
        WRef wq{w->qkv_w[l], bf(P, L.qkv_hi[l]), bf(P, L.qkv_lo[l])};
        set_tag(TAG_QKV);
        RC(linear(prec, xn1, kDim, wq, M, 3 * kDim, kDim, e1, s));
        set_tag(TAG_ATT);
        RC(attention_seq(qkv, R, ao.out(), s));
        GemmEpilogue e2;  // x_mid = Dropout(to_out(attn)) + x   (model_veto.py:19,83-85)
        e2.bias = w->out_b[l];
        e2.residual = x_in[l];
        e2.out.f32 = x_mid;
        e2.ldc = kDim;
        e2.drop = make_drop(tin->p_attn_dropout, sub_seed(tin->seed, 16 + l));
        WRef wo{w->out_w[l], bf(P, L.out_hi[l]), bf(P, L.out_lo[l])};
        set_tag(TAG_OUT);
        RC(linear(prec, ao, kDim, wo, M, kDim, kDim, e2, s));
        set_tag(TAG_LN);
        RC(layernorm_rows(x_mid, kDim, w->ln2_w[l], w->ln2_b[l], M, xn2.out(), s));
        GemmEpilogue e3;
        e3.bias = w->ff1_b[l];
        e3.act = ACT_GELU;
        e3.out = hb.out();
        e3.pre_f32 = X.f32(S.h_pre);
        e3.ldc = kMlp;
        WRef w1{w->ff1_w[l], bf(P, L.ff1_hi[l]), bf(P, L.ff1_lo[l])};
        set_tag(TAG_FF1);
        RC(linear(prec, xn2, kDim, w1, M, kMlp, kDim, e3, s));
        GemmEpilogue e4;
        e4.bias = w->ff2_b[l];
        e4.residual = x_mid;
        e4.out.f32 = x_in[l + 1];
        e4.ldc = kDim;
        WRef w2{w->ff2_w[l], bf(P, L.ff2_hi[l]), bf(P, L.ff2_lo[l])};
        set_tag(TAG_FF2);
        RC(linear(prec, hb, kMlp, w2, M, kDim, kMlp, e4, s));
    }
    // rel_out on x[:,0] (rows r*19 of the last residual stream), always fp32 FMA
    float* logits = out->rel_logits ? out->rel_logits : X.f32(T.logits);
    {
        GemmEpilogue ec;
        ec.bias = w->rel_out_b;
        ec.out.f32 = logits;
        ec.ldc = C;
        set_tag(TAG_CLS);
        RC(gemm_simt(x_in[NL], kTokens * kDim, w->rel_out_w, R, C, kDim, ec, s));
    }

    // =====================================================================================  loss and backward
    set_tag(TAG_LOSS);
    float* dlogits = X.f32(T.dlogits);
    if (tin->n_heads > 1) {
        // MEET group heads (roi_relation_predictors.py:3834-3846): one unweighted CE per head over its own columns and the
        // rows the group sampling chose for it; the step's gradient is that of the SUM of the head losses
        for (int k = 0; k < tin->n_heads; ++k) {
            const int c0 = tin->head_offsets[k], ck = tin->head_offsets[k + 1] - c0;
            RC(ce_loss_grad(logits, C, c0, ck, tin->head_labels + (size_t)k * R, nullptr, R, X.f32(T.ce_scratch), out->loss + k,
                            dlogits, s));
        }
    } else {
        RC(ce_loss_grad(logits, C, 0, C, tin->rel_labels, tin->class_weight, R, X.f32(T.ce_scratch), out->loss, dlogits, s));
    }
    float* dx = X.f32(T.dx);
    float* tmp = X.f32(T.tmp);
    set_tag(TAG_BWD_OTHER);
    {
        // rel_out: d b = colsum(dlogits); d W = dlogits^T x_cls; d x_cls = dlogits W (scattered into rows r*19 of dx)
        RC(X.bias_grad(f32_in(dlogits), C, R, C, g->rel_out_b));
        float* tc1 = X.f32(T.tc1);
        float* tc2 = X.f32(T.tc2);
        float* wcT = X.f32(T.wcT);
        ActOut o1, o2, o3;
        o1.f32 = tc1; o2.f32 = tc2; o3.f32 = wcT;
        RC(transpose_f32(dlogits, C, R, C, false, DropSpec(), 0, o1, T.Rp, T.Rp, ActOut(), 0, s));
        RC(transpose_f32(x_in[NL], (int64_t)kTokens * kDim, R, kDim, false, DropSpec(), 0, o2, T.Rp, T.Rp, ActOut(), 0, s));
        RC(transpose_f32(w->rel_out_w, kDim, C, kDim, false, DropSpec(), 0, o3, C, C, ActOut(), 0, s));
        // d W = dlogits^T x_cls: a [C, 576] output over R rows — split-K over the rows, fixed-order reduction
        GemmEpilogue ep;
        ep.ldc = kDim;
        const int wc_slices = gemm_simt_slices((int)T.Rp, kMaxSplit);
        ep.out.f32 = wc_slices > 1 ? X.f32(T.wc_partial) : g->rel_out_w;
        ep.split_k = wc_slices;
        ep.split_stride = (size_t)C * kDim;
        set_tag(TAG_BWD_GEMM);
        RC(gemm_simt(tc1, (int)T.Rp, tc2, C, kDim, (int)T.Rp, ep, s));
        if (wc_slices > 1) RC(splitk_reduce(X.f32(T.wc_partial), wc_slices, (size_t)C * kDim, (size_t)C * kDim, g->rel_out_w, s));
        VETO_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)M * kDim, s));
        GemmEpilogue ed;
        ed.out.f32 = dx;
        ed.ldc = kTokens * kDim;
        RC(gemm_simt(dlogits, C, wcT, R, kDim, C, ed, s));
    }
    // transposed encoder weights for the input-gradient GEMMs
    set_tag(TAG_PACK);
    ActBuf wT_qkv[VETO_MAX_LAYERS], wT_out[VETO_MAX_LAYERS], wT_ff1[VETO_MAX_LAYERS], wT_ff2[VETO_MAX_LAYERS];
    TransposeJob tjobs[kMaxTransposeJobs];
    for (int l = 0; l < NL; ++l) {
        wT_qkv[l] = X.act(T.WT[l].qkv, (size_t)3 * kDim * kDim);  // [576, 1728]
        wT_out[l] = X.act(T.WT[l].out, (size_t)kDim * kDim);      // [576, 576]
        wT_ff1[l] = X.act(T.WT[l].ff1, (size_t)kMlp * kDim);      // [576, 1152]
        wT_ff2[l] = X.act(T.WT[l].ff2, (size_t)kMlp * kDim);      // [1152, 576]
        tjobs[4 * l + 0] = TransposeJob{w->qkv_w[l], wT_qkv[l].out(), 3 * kDim, kDim};
        tjobs[4 * l + 1] = TransposeJob{w->out_w[l], wT_out[l].out(), kDim, kDim};
        tjobs[4 * l + 2] = TransposeJob{w->ff1_w[l], wT_ff1[l].out(), kMlp, kDim};
        tjobs[4 * l + 3] = TransposeJob{w->ff2_w[l], wT_ff2[l].out(), kDim, kMlp};
    }
    RC(transpose_f32_multi(tjobs, 4 * NL, s));  // one launch instead of 4 per layer

    ActBuf a576 = X.act(T.a576, (size_t)M * kDim);
    ActBuf a1728 = X.act(T.a1728, (size_t)M * 3 * kDim);
    float* ln_partial = X.f32(T.ln_partial);
    for (int l = NL - 1; l >= 0; --l) {
        const LayerSave& S = T.L[l];
        ActBuf xn1 = X.act(S.xn1, (size_t)M * kDim), ao = X.act(S.ao, (size_t)M * kDim), xn2 = X.act(S.xn2, (size_t)M * kDim);
        ActBuf hb = X.act(S.h, (size_t)M * kMlp);
        // ---- FeedForward second Linear: x_out = h W2^T + b2 + x_mid.  dx in operand format (a576) and its column
        // sums (ff2_b) come from the LayerNorm backward of the layer above; only the top layer converts here
        set_tag(TAG_BWD_OTHER);
        if (l == NL - 1) {
            RC(X.to_operand(dx, (size_t)M * kDim, DropSpec(), a576));
            RC(X.bias_grad(f32_in(dx), kDim, M, kDim, g->ff2_b[l]));
        }
        // with re-computation the shared buffers hold the TOP layer's values after the forward pass; every layer below
        // restores its own right before the weight gradient that reads it
        const bool redo = train_recompute() && l != NL - 1;
        if (redo) {
            set_tag(TAG_BWD_OTHER);
            RC(convert_act(X.f32(S.h_pre), (size_t)M * kMlp, DropSpec(), hb.out(), s, ACT_GELU));
        }
        set_tag(TAG_BWD_GEMM);
        RC(X.wgrad(a576, kDim, hb, kMlp, M, kDim, kMlp, g->ff2_w[l]));
        ActBuf dh = X.act(T.a1728, (size_t)M * kMlp);
        {
            GemmEpilogue ep;  // d h_pre = (dx W2) * gelu'(h_pre)
            ep.residual = X.f32(S.h_pre);
            ep.res_mode = RES_GELU_GRAD;
            ep.out = dh.out();
            ep.ldc = kMlp;
            RC(X.mm(a576, kDim, wT_ff2[l], M, kMlp, kDim, ep));
        }
        // ---- FeedForward first Linear: h_pre = LN2(x_mid) W1^T + b1
        set_tag(TAG_BWD_OTHER);
        RC(X.bias_grad(as_in(dh), kMlp, M, kMlp, g->ff1_b[l]));
        if (redo) {
            set_tag(TAG_BWD_LN);
            RC(layernorm_rows(X.f32(S.x_mid), kDim, w->ln2_w[l], w->ln2_b[l], M, xn2.out(), s));
        }
        set_tag(TAG_BWD_GEMM);
        RC(X.wgrad(dh, kMlp, xn2, kDim, M, kMlp, kDim, g->ff1_w[l]));
        {
            GemmEpilogue ep;
            ep.out.f32 = tmp;
            ep.ldc = kDim;
            RC(X.mm(dh, kMlp, wT_ff1[l], M, kDim, kMlp, ep));
        }
        set_tag(TAG_BWD_LN);
        // ---- attention output projection: x_mid = Dropout(ao Wo^T + bo) + x_in: the LayerNorm backward also emits
        // mask * dx in operand format (a576) and its column sums = the to_out bias gradient
        const DropSpec drop_l = make_drop(tin->p_attn_dropout, sub_seed(tin->seed, 16 + l));
        RC(layernorm_bwd(X.f32(S.x_mid), kDim, tmp, w->ln2_w[l], dx, dx, M, ln_partial, X.colsum_scratch, g->ln2_w[l], g->ln2_b[l], s,
                         drop_l, a576.out(), g->out_b[l]));
        set_tag(TAG_BWD_GEMM);
        RC(X.wgrad(a576, kDim, ao, kDim, M, kDim, kDim, g->out_w[l]));
        {
            GemmEpilogue ep;
            ep.out.f32 = tmp;
            ep.ldc = kDim;
            RC(X.mm(a576, kDim, wT_out[l], M, kDim, kDim, ep));
        }
        // ---- attention core, then to_qkv: qkv = LN1(x_in) Wqkv^T
        set_tag(TAG_BWD_ATT);
        ActBuf dqkv = X.act(T.a1728, (size_t)M * 3 * kDim);
        RC(attention_bwd(X.f32(S.qkv), tmp, R, dqkv.out(), s));
        if (redo) {
            set_tag(TAG_BWD_LN);
            RC(layernorm_rows(x_in[l], kDim, w->ln1_w[l], w->ln1_b[l], M, xn1.out(), s));
        }
        set_tag(TAG_BWD_GEMM);
        RC(X.wgrad(dqkv, 3 * kDim, xn1, kDim, M, 3 * kDim, kDim, g->qkv_w[l]));
        {
            GemmEpilogue ep;
            ep.out.f32 = tmp;
            ep.ldc = kDim;
            RC(X.mm(dqkv, 3 * kDim, wT_qkv[l], M, kDim, 3 * kDim, ep));
        }
        set_tag(TAG_BWD_LN);
        if (l > 0)  // dx feeds the FeedForward backward of layer l - 1: operand copy + ff2 bias gradient ride along
            RC(layernorm_bwd(x_in[l], kDim, tmp, w->ln1_w[l], dx, dx, M, ln_partial, X.colsum_scratch, g->ln1_w[l], g->ln1_b[l], s,
                             DropSpec(), a576.out(), g->ff2_b[l - 1]));
        else
            RC(layernorm_bwd(x_in[l], kDim, tmp, w->ln1_w[l], dx, dx, M, ln_partial, X.colsum_scratch, g->ln1_w[l], g->ln1_b[l], s));
    }

    // ---- encoder input: x = Dropout(cat(cls, patches, loc, cls) + pos_embedding)
    set_tag(TAG_BWD_BOX);
    RC(dropout_inplace(dx, (size_t)M * kDim, drop_emb, s));
    RC(X.bias_grad(f32_in(dx), kDim, M, kDim, g->pos_embedding));
    RC(X.bias_grad(f32_in(dx), (int64_t)kTokens * kDim, R, kDim, g->cls_token));
    float* d_so_d = X.f32(T.d_so_d);
    float* d_so_v = X.f32(T.d_so_v);
    float* d_lso = X.f32(T.d_lso);
    float* d_cso = X.f32(T.d_cso);
    RC(tokens_bwd(dx, in->subj, in->obj, tin->rel_offsets, tin->box_offsets, tin->n_images, N, lso, cso, d_so_d, d_so_v, d_lso,
                  d_cso, s));

    // ---- patch projections (per box): so = patches W2^T + b2, W2 the subject/object-factored proj_d / proj_v
    const int rowsB = N * kPatches;
    {
        ActBuf a_box = X.act(T.a_box, (size_t)rowsB * 2 * kDimDepth);
        float* g_w_d2 = X.f32(T.g_w_d2);
        RC(X.to_operand(d_so_d, (size_t)rowsB * 2 * kDimDepth, DropSpec(), a_box));
        RC(X.bias_grad(f32_in(d_so_d), 2 * kDimDepth, rowsB, kDimDepth, g->proj_d_b));  // the bias sits in the subject half
        set_tag(TAG_BWD_GEMM);
        RC(X.wgrad(a_box, 2 * kDimDepth, pa_d, kPatchVec, rowsB, 2 * kDimDepth, kPatchVec, g_w_d2));
        set_tag(TAG_BWD_BOX);
        RC(unpack_patch(g_w_d2, g->proj_d_w, kDimDepth, s));
        if (out->grad_roi_depth) {
            ActBuf d2T = X.act(T.d2T, (size_t)2 * kDimDepth * kPatchVec);  // [1024 in, 1024 out]
            RC(transpose_f32((const float*)(P + L.w_d2), kPatchVec, 2 * kDimDepth, kPatchVec, false, DropSpec(), 0, d2T.out(),
                             2 * kDimDepth, 2 * kDimDepth, ActOut(), 0, s));
            GemmEpilogue ep;
            ep.out.f32 = X.f32(T.d_pa);
            ep.ldc = kPatchVec;
            set_tag(TAG_BWD_GEMM);
            RC(X.mm(a_box, 2 * kDimDepth, d2T, rowsB, kPatchVec, 2 * kDimDepth, ep));
            set_tag(TAG_BWD_BOX);
            RC(unpatchify(X.f32(T.d_pa), N, out->grad_roi_depth, s));
        }
    }
    {
        ActBuf a_box = X.act(T.a_box, (size_t)rowsB * 2 * kDimRgb);
        float* g_w_v2 = X.f32(T.g_w_v2);
        RC(X.to_operand(d_so_v, (size_t)rowsB * 2 * kDimRgb, DropSpec(), a_box));
        RC(X.bias_grad(f32_in(d_so_v), 2 * kDimRgb, rowsB, kDimRgb, g->proj_v_b));
        set_tag(TAG_BWD_GEMM);
        RC(X.wgrad(a_box, 2 * kDimRgb, pa_v, kPatchVec, rowsB, 2 * kDimRgb, kPatchVec, g_w_v2));
        set_tag(TAG_BWD_BOX);
        RC(unpack_patch(g_w_v2, g->proj_v_w, kDimRgb, s));
        if (out->grad_roi_rgb) {
            ActBuf v2T = X.act(T.v2T, (size_t)2 * kDimRgb * kPatchVec);  // [1024 in, 128 out]
            RC(transpose_f32((const float*)(P + L.w_v2), kPatchVec, 2 * kDimRgb, kPatchVec, false, DropSpec(), 0, v2T.out(),
                             2 * kDimRgb, 2 * kDimRgb, ActOut(), 0, s));
            GemmEpilogue ep;
            ep.out.f32 = X.f32(T.d_pa);
            ep.ldc = kPatchVec;
            set_tag(TAG_BWD_GEMM);
            RC(X.mm(a_box, 2 * kDimRgb, v2T, rowsB, kPatchVec, 2 * kDimRgb, ep));
            set_tag(TAG_BWD_BOX);
            RC(unpatchify(X.f32(T.d_pa), N, out->grad_roi_rgb, s));
        }
    }

    // ---- location / class projections (per box, fp32 FMA like their forward) and the box embeddings
    {
        const int64_t Nb = T.Nb;
        float* tb1 = X.f32(T.tb1);
        float* tb2 = X.f32(T.tb2);
        ActOut o1, o2;
        o1.f32 = tb1; o2.f32 = tb2;
        // location_projection: lso = pos W_loc2^T + b
        RC(transpose_f32(d_lso, 2 * kDim, N, 2 * kDim, false, DropSpec(), 0, o1, Nb, Nb, ActOut(), 0, s));
        RC(transpose_f32(pos, kPosDim, N, kPosDim, false, DropSpec(), 0, o2, Nb, Nb, ActOut(), 0, s));
        set_tag(TAG_BWD_GEMM);
        RC(X.mm_small(tb1, (int)Nb, tb2, 2 * kDim, kPosDim, (int)Nb, X.f32(T.g_w_loc2), kPosDim));
        set_tag(TAG_BWD_BOX);
        RC(unpack_halves(X.f32(T.g_w_loc2), g->loc_proj_w, kDim, kPosDim, s));
        RC(X.bias_grad(f32_in(d_lso), 2 * kDim, N, kDim, g->loc_proj_b));
        ActOut ot;
        ot.f32 = X.f32(T.loc2T);  // [128, 1152]
        RC(transpose_f32((const float*)(P + L.w_loc2), kPosDim, 2 * kDim, kPosDim, false, DropSpec(), 0, ot, 2 * kDim, 2 * kDim,
                         ActOut(), 0, s));
        set_tag(TAG_BWD_GEMM);
        RC(X.mm_small(d_lso, 2 * kDim, X.f32(T.loc2T), N, kPosDim, 2 * kDim, X.f32(T.d_pos), kPosDim));
        set_tag(TAG_BWD_BOX);
        RC(pos_embed_bwd(in->boxes, N, bn_stats, *w, pos, X.f32(T.d_pos), drop_pos.scale, X.f32(T.pos_partial), X.colsum_scratch,
                         g->pos_w, g->pos_b, g->bn_weight, g->bn_bias, s));
        // class_projection: cso = emb W_cls2^T + b
        RC(transpose_f32(d_cso, 2 * kDim, N, 2 * kDim, false, DropSpec(), 0, o1, Nb, Nb, ActOut(), 0, s));
        RC(transpose_f32(emb, kEmbDim, N, kEmbDim, false, DropSpec(), 0, o2, Nb, Nb, ActOut(), 0, s));
        set_tag(TAG_BWD_GEMM);
        RC(X.mm_small(tb1, (int)Nb, tb2, 2 * kDim, kEmbDim, (int)Nb, X.f32(T.g_w_cls2), kEmbDim));
        set_tag(TAG_BWD_BOX);
        RC(unpack_halves(X.f32(T.g_w_cls2), g->class_proj_w, kDim, kEmbDim, s));
        RC(X.bias_grad(f32_in(d_cso), 2 * kDim, N, kDim, g->class_proj_b));
        ActOut oc;
        oc.f32 = X.f32(T.cls2T);  // [200, 1152]
        RC(transpose_f32((const float*)(P + L.w_cls2), kEmbDim, 2 * kDim, kEmbDim, false, DropSpec(), 0, oc, 2 * kDim, 2 * kDim,
                         ActOut(), 0, s));
        set_tag(TAG_BWD_GEMM);
        RC(X.mm_small(d_cso, 2 * kDim, X.f32(T.cls2T), N, kEmbDim, 2 * kDim, X.f32(T.d_emb), kEmbDim));
        set_tag(TAG_BWD_BOX);
        RC(embed_bwd(X.f32(T.d_emb), in->labels, in->obj_logits, cfg->num_obj, N, g->obj_embed, s));
    }
    set_tag(TAG_OTHER);
    return VETO_OK;
}
