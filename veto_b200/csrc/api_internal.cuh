// Internals shared by the forward (api.cu) and training (train_api.cu) orchestration: configuration check, the layout
// of the packed-weight buffer, activation buffers in the storage format of a precision mode, and the Linear dispatch.
#pragma once
#include "stages.cuh"

namespace veto {

constexpr size_t kAlign = 256;
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off += (bytes + kAlign - 1) / kAlign * kAlign;
        return o;
    }
};

static inline int check_config(const veto_config* c) {
    VETO_REQUIRE(c != nullptr, VETO_ERR_ARG, "veto_config is NULL");
    VETO_REQUIRE(c->dim == kDim && c->heads == kHeads && c->mlp_dim == kMlp && c->channels == kChannels &&
                     c->pool == kPool && c->patch == 2,
                 VETO_ERR_UNSUPPORTED,
                 "unsupported architecture (dim %d heads %d mlp %d channels %d pool %d patch %d): this library is built "
                 "for configs/VETO_final.yaml (576/6/1152/256/8/2)",
                 c->dim, c->heads, c->mlp_dim, c->channels, c->pool, c->patch);
    VETO_REQUIRE(c->layers >= 1 && c->layers <= VETO_MAX_LAYERS, VETO_ERR_UNSUPPORTED, "layers=%d outside 1..%d", c->layers,
                 VETO_MAX_LAYERS);
    VETO_REQUIRE(c->num_obj >= 2 && c->num_obj <= 512 && c->num_out >= 1, VETO_ERR_ARG, "bad num_obj=%d / num_out=%d",
                 c->num_obj, c->num_out);
    VETO_REQUIRE(c->precision >= VETO_PREC_FP32 && c->precision <= VETO_PREC_F16, VETO_ERR_ARG, "bad precision %d",
                 c->precision);
    return VETO_OK;
}

// Storage class of a precision mode: which arrays exist next to the fp32 ones.  f16c8 stores two 2-byte-per-element
// arrays like bf16x3, f16 one like bf16 — buffer sizes and offsets depend on this only, not on the element encoding.
static inline bool prec_two_arrays(int precision) { return precision == VETO_PREC_BF16X3 || precision == VETO_PREC_F16C8; }
// the per-box stage (patch projections: 0.1 % of the work) runs as bf16x3 in every tensor-core mode but the plain bf16 one
static inline int prec_box(int precision) {
    return (precision == VETO_PREC_F16C8 || precision == VETO_PREC_F16) ? VETO_PREC_BF16X3 : precision;
}
// operand format of the ENCODER GEMMs in a precision mode (the per-box patch projections stay bf16 / bf16x3)
static inline int prec_encoder_fmt(int precision) {
    return (precision == VETO_PREC_F16C8 || precision == VETO_PREC_F16) ? FMT_F16C8 : FMT_BF16;
}
static inline int prec_encoder_passes(int precision) {
    return precision == VETO_PREC_BF16X3 ? TC_BF16X3 : precision == VETO_PREC_F16C8 ? TC_F16C8 : precision == VETO_PREC_F16 ? TC_F16 : TC_BF16;
}

struct PackedLayout {
    size_t w_loc2, b_loc2, w_cls2, b_cls2, w_d2, b_d2, w_v2, b_v2, clspos;
    size_t d2_hi, d2_lo, v2_hi, v2_lo;
    size_t qkv_hi[VETO_MAX_LAYERS], qkv_lo[VETO_MAX_LAYERS], out_hi[VETO_MAX_LAYERS], out_lo[VETO_MAX_LAYERS];
    size_t ff1_hi[VETO_MAX_LAYERS], ff1_lo[VETO_MAX_LAYERS], ff2_hi[VETO_MAX_LAYERS], ff2_lo[VETO_MAX_LAYERS];
    // LayerNorm-fused inference (gemm_tc2.cu EPI_*_LN): to_qkv / FF1 weights with the LayerNorm weight folded in, and
    // their constants c1 | c2 (2 x N floats)
    size_t qkvf_hi[VETO_MAX_LAYERS], qkvf_lo[VETO_MAX_LAYERS], ff1f_hi[VETO_MAX_LAYERS], ff1f_lo[VETO_MAX_LAYERS];
    size_t c_qkv[VETO_MAX_LAYERS], c_ff1[VETO_MAX_LAYERS];
    size_t total;
};

static inline PackedLayout packed_layout(const veto_config& c) {
    PackedLayout L{};
    Carver k;
    L.w_loc2 = k.take(sizeof(float) * 2 * kDim * kPosDim);
    L.b_loc2 = k.take(sizeof(float) * 2 * kDim);
    L.w_cls2 = k.take(sizeof(float) * 2 * kDim * kEmbDim);
    L.b_cls2 = k.take(sizeof(float) * 2 * kDim);
    L.w_d2 = k.take(sizeof(float) * 2 * kDimDepth * kPatchVec);
    L.b_d2 = k.take(sizeof(float) * 2 * kDimDepth);
    L.w_v2 = k.take(sizeof(float) * 2 * kDimRgb * kPatchVec);
    L.b_v2 = k.take(sizeof(float) * 2 * kDimRgb);
    L.clspos = k.take(sizeof(float) * kDim);
    if (c.precision != VETO_PREC_FP32) {
        const bool lo = prec_two_arrays(c.precision);
        const bool box_lo = prec_two_arrays(prec_box(c.precision));
        const size_t e = sizeof(__nv_bfloat16);
        L.d2_hi = k.take(e * 2 * kDimDepth * kPatchVec);
        L.d2_lo = box_lo ? k.take(e * 2 * kDimDepth * kPatchVec) : 0;
        L.v2_hi = k.take(e * 2 * kDimRgb * kPatchVec);
        L.v2_lo = box_lo ? k.take(e * 2 * kDimRgb * kPatchVec) : 0;
        for (int l = 0; l < c.layers; ++l) {
            L.qkv_hi[l] = k.take(e * 3 * kDim * kDim);
            L.qkv_lo[l] = lo ? k.take(e * 3 * kDim * kDim) : 0;
            L.out_hi[l] = k.take(e * kDim * kDim);
            L.out_lo[l] = lo ? k.take(e * kDim * kDim) : 0;
            L.ff1_hi[l] = k.take(e * kMlp * kDim);
            L.ff1_lo[l] = lo ? k.take(e * kMlp * kDim) : 0;
            L.ff2_hi[l] = k.take(e * kDim * kMlp);
            L.ff2_lo[l] = lo ? k.take(e * kDim * kMlp) : 0;
            L.qkvf_hi[l] = k.take(e * 3 * kDim * kDim);
            L.qkvf_lo[l] = lo ? k.take(e * 3 * kDim * kDim) : 0;
            L.ff1f_hi[l] = k.take(e * kMlp * kDim);
            L.ff1f_lo[l] = lo ? k.take(e * kMlp * kDim) : 0;
            L.c_qkv[l] = k.take(sizeof(float) * 2 * 3 * kDim);
            L.c_ff1[l] = k.take(sizeof(float) * 2 * kMlp);
        }
    }
    L.total = k.off;
    return L;
}

// an activation buffer in the storage format of the precision mode
struct ActBuf {
    float* f32 = nullptr;
    __nv_bfloat16* hi = nullptr;
    __nv_bfloat16* lo = nullptr;
    int fmt = FMT_BF16;
    ActOut out() const { return ActOut{f32, hi, lo, fmt}; }
};

static inline size_t act_bytes(int precision, size_t elems) {
    if (precision == VETO_PREC_FP32) return elems * sizeof(float);
    if (prec_two_arrays(precision)) return elems * 2 * sizeof(__nv_bfloat16);
    return elems * sizeof(__nv_bfloat16);
}

static inline ActBuf act_at(void* base, size_t off, int precision, size_t elems) {
    ActBuf b;
    char* p = (char*)base + off;
    if (precision == VETO_PREC_FP32) b.f32 = (float*)p;
    else {
        b.hi = (__nv_bfloat16*)p;
        if (prec_two_arrays(precision)) b.lo = b.hi + elems;
    }
    return b;
}
// the same buffer as an operand of the encoder GEMMs (f16c8 / f16 modes store it in the f16c8 format)
static inline ActBuf enc_act_at(void* base, size_t off, int precision, size_t elems) {
    ActBuf b = act_at(base, off, precision, elems);
    b.fmt = prec_encoder_fmt(precision);
    return b;
}

struct WRef {  // a Linear weight in the forms the two GEMM paths need
    const float* f32;
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;
};

static inline int linear(int precision, const ActBuf& a, int lda, const WRef& w, int M, int N, int K, const GemmEpilogue& ep,
           cudaStream_t s) {
    if (precision == VETO_PREC_FP32) return gemm_simt(a.f32, lda, w.f32, M, N, K, ep, s);
    GemmOperand A, W;
    A.hi = a.hi; A.lo = a.lo;
    A.ld = lda;
    W.hi = w.hi; W.lo = w.lo;
    // a.fmt tells the box stage (bf16 operands in every tensor-core mode) from the encoder (format of the mode)
    const int passes = a.fmt == FMT_F16C8 ? prec_encoder_passes(precision) : (prec_two_arrays(prec_box(precision)) ? TC_BF16X3 : TC_BF16);
    return gemm_tc_auto(A, W, M, N, K, passes, ep, s);
}

static inline const __nv_bfloat16* bf(const void* base, size_t off) { return off ? (const __nv_bfloat16*)((const char*)base + off) : nullptr; }


}  // namespace veto
