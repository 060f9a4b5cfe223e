// Shared declarations of libveto_b200.so (internal; the public surface is include/veto_b200.h).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/veto_b200.h"

namespace veto {
// One-time per-DEVICE setup guard: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device only,
// and one process may drive several GPUs.  Usage: static DeviceOnce once; if (once.pending()) { ...set...; once.done(); }
struct DeviceOnce {
    unsigned long long mask[2] = {0ull, 0ull};   // devices 0..127; racing threads at worst repeat the (idempotent) setup
    static int device() {
        int d = 0;
        cudaGetDevice(&d);
        return d & 127;
    }
    bool pending() const {
        const int d = device();
        return ((mask[d >> 6] >> (d & 63)) & 1ull) == 0ull;
    }
    void done() {
        const int d = device();
        __atomic_fetch_or(&mask[d >> 6], 1ull << (d & 63), __ATOMIC_RELAXED);
    }
};
}  // namespace veto

namespace veto {

// ---- architecture constants of configs/VETO_final.yaml (validated against veto_config) ----
constexpr int kDim = 576;        // T_INPUT_DIM
constexpr int kHeads = 6;        // NHEADS
constexpr int kHeadDim = 96;     // 576 / 6 (model_veto.py:70)
constexpr int kMlp = 1152;       // mlp_dim = 2 * dim (model_veto.py:35)
constexpr int kTokens = 19;      // 1 cls + 16 patches + location + class (model_veto.py:52-61)
constexpr int kPatches = 16;     // (8/2)^2
constexpr int kChannels = 256;   // ROI feature channels
constexpr int kPool = 8;
constexpr int kPatchVec = 1024;  // p1*p2*c for ONE box (the reference's 2048 = subject + object halves)
constexpr int kDimDepth = 512;   // proj_d out (model_veto.py:105)
constexpr int kDimRgb = 64;      // proj_v out (model_veto.py:106)
constexpr int kPosDim = 128;     // pos_embed out (roi_relation_predictors.py:4042-4047)
constexpr int kEmbDim = 200;     // obj_embed dim

// ---- error plumbing ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);
// stage tags of the per-stage event timing (veto_profile_*); names in api.cu
enum { TAG_OTHER = 0, TAG_PAIRS, TAG_GATHER, TAG_BOX, TAG_TOKENS, TAG_LN, TAG_QKV, TAG_ATT, TAG_OUT, TAG_FF1, TAG_FF2,
       TAG_CLS, TAG_POST, TAG_PACK, TAG_BWD_GEMM, TAG_BWD_OTHER, TAG_BWD_ATT, TAG_BWD_LN, TAG_BWD_BOX, TAG_BWD_WGRAD, TAG_LOSS };
void set_tag(int tag);

#define VETO_CUDA(expr)                                                               \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) return ::veto::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define VETO_LAUNCH_CHECK()                                                                  \
    do {                                                                                     \
        ::veto::count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) return ::veto::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define VETO_REQUIRE(cond, code, ...)   \
    do {                                \
        if (!(cond)) {                  \
            ::veto::set_error(__VA_ARGS__); \
            return (code);              \
        }                               \
    } while (0)

// ---- activation storage formats between kernels ----
// F32     : one fp32 array
// BF16    : one bf16 array
// BF16X2  : bf16 hi array followed (at `lo`) by a bf16 lo array, value ~= hi + lo (16 mantissa bits)
// F16C8   : (fmt = FMT_F16C8) `hi` holds fp16 values, `lo` holds two e4m3 bytes per element (see the f16c8 block below);
//           same byte counts as BF16X2, so buffers and offsets do not depend on the format
enum { FMT_BF16 = 0, FMT_F16C8 = 1, FMT_F16C8_ACT = 2 };   // _ACT: pack jobs only (stages.cuh SplitJob)
struct ActOut {
    float* f32 = nullptr;
    __nv_bfloat16* hi = nullptr;
    __nv_bfloat16* lo = nullptr;
    int fmt = FMT_BF16;
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ uint2 pack_bf16x4(__nv_bfloat16 a, __nv_bfloat16 b, __nv_bfloat16 c, __nv_bfloat16 d) {
    uint2 r;
    r.x = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
    r.y = (uint32_t)__bfloat16_as_ushort(c) | ((uint32_t)__bfloat16_as_ushort(d) << 16);
    return r;
}

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));  // nn.GELU() exact (model_veto.py:139)
    return v;
}

// GELU for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7 + fp32 rounding, the same
// order as erff's own), one exp and one reciprocal, no branches — the exact-erff version made the FF1 epilogue
// ALU-bound (it costs about as many issue slots per tile as the MMAs take cycles).
__device__ __forceinline__ float gelu_fast(float v) {
    // z = |v| / sqrt2; t = 1 / (1 + 0.3275911 z); exp(-z^2) = 2^(-v^2 * log2(e) / 2): the two MUFU ops as plain .approx.ftz
    // instructions (no range fix-ups: t is in (0, 1], the exponent is <= 0) and the constants folded — 14 instructions
    float t, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(v), 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(v * v * -0.72134752044448170368f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = fmaf(-(p * t), ex, 1.f);               // erf(|v|/sqrt2)
    const float hv = 0.5f * v;
    return fmaf(hv, copysignf(e, v), hv);
}
__device__ __forceinline__ float apply_act_tc(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_GELU) return gelu_fast(v);
    return v;
}

// (x0, x1) -> packed bf16 hi pair and lo pair (lo = x - float(hi)): 6 instructions for two elements
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- f16c8 operand format (precision mode VETO_PREC_F16C8: one fp16 product + fp8 first-order corrections) ----
// A Linear y = a . w is evaluated on the tensor cores as
//     2^-11 * [ fp16(a) . fp16(2048 w)                        kind::f16,    K = 16 per instruction
//             + e4m3(256 r_a) . e4m3(8 w) + e4m3(a / 8) . e4m3(8 r_w) ]    kind::f8f6f4, K = 32 per instruction (2x rate)
// with r_a = a - fp16(a) and r_w = 2048 w - fp16(2048 w): the fp16 product carries 11 significant bits per operand and
// the two e4m3 products restore the first-order rounding residuals to ~4 more bits — 2 bf16-MMA equivalents instead of
// the 3 of the bf16x3 split, measured error 1e-4 of the logit range (tools/precision_study.py; bar 1e-3).  All scale
// factors are powers of two (exact).  Ranges: |a| < 3584 (e4m3(a/8) saturates at 448), |w| < 32 (fp16(2048 w)); both
// saturate instead of overflowing.
// Storage of an [rows, K] operand (K % 64 == 0): `hi` = fp16 [rows, K]; `lo` = bytes [rows, 2K]: for every 64-element
// K block, 64 bytes of the first e4m3 stream followed by 64 bytes of the second — activations (residual, value),
// weights (value, residual) — so that one 128-byte row of the K-major SWIZZLE_128B tile of A meets the matching row of W
// and the correction is ONE fp8 GEMM over 2K.  Byte-compatible with the bf16 hi / lo arrays ([rows, K] 2-byte elements).
constexpr float kC8ActRes = 256.f, kC8ActVal = 0.125f, kC8WScale = 2048.f, kC8WVal = 8.f, kC8WRes = 8.f;
constexpr float kC8AccScale = 1.f / 2048.f;

__device__ __forceinline__ uint32_t f16x2_sat(float x0, float x1) {   // x0 in the low half; saturating (one F2FP.SATFINITE)
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ float2 f16x2_to_float(uint32_t h) {
    return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
__device__ __forceinline__ uint32_t e4m3x2(float x0, float x1) {   // x0 in the low byte; saturating
    return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(x0, x1), __NV_SATFINITE, __NV_E4M3);
}
// byte offset inside the `lo` array of the first-stream byte of element `off` (flat index row * K + k, K % 64 == 0);
// the second-stream byte sits 64 bytes further
__device__ __forceinline__ size_t c8_byte(size_t off) { return ((off >> 6) << 7) + (off & 63); }

// two consecutive activations (off even) -> fp16 pair in hi, (residual, value) e4m3 pairs in lo (may be NULL: f16 mode)
__device__ __forceinline__ void store_act2_f16c8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, float x0, float x1) {
    const uint32_t h = f16x2_sat(x0, x1);
    *(uint32_t*)(hi + off) = h;
    if (lo) {
        const float2 hf = f16x2_to_float(h);
        uint8_t* p = (uint8_t*)lo + c8_byte(off);
        *(uint16_t*)p = (uint16_t)e4m3x2((x0 - hf.x) * kC8ActRes, (x1 - hf.y) * kC8ActRes);
        *(uint16_t*)(p + 64) = (uint16_t)e4m3x2(x0 * kC8ActVal, x1 * kC8ActVal);
    }
}
__device__ __forceinline__ void store_act4_f16c8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, float4 v) {
    uint2 h;
    h.x = f16x2_sat(v.x, v.y);
    h.y = f16x2_sat(v.z, v.w);
    *(uint2*)(hi + off) = h;
    if (lo) {
        const float2 a = f16x2_to_float(h.x), b = f16x2_to_float(h.y);
        uint8_t* p = (uint8_t*)lo + c8_byte(off);
        *(uint32_t*)p = e4m3x2((v.x - a.x) * kC8ActRes, (v.y - a.y) * kC8ActRes) |
                        (e4m3x2((v.z - b.x) * kC8ActRes, (v.w - b.y) * kC8ActRes) << 16);
        *(uint32_t*)(p + 64) = e4m3x2(v.x * kC8ActVal, v.y * kC8ActVal) | (e4m3x2(v.z * kC8ActVal, v.w * kC8ActVal) << 16);
    }
}
__device__ __forceinline__ void store_act1_f16c8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, float x) {
    const uint32_t hp = f16x2_sat(x, 0.f);
    const __half h = __ushort_as_half((unsigned short)(hp & 0xffffu));
    *((__half*)hi + off) = h;
    if (lo) {
        uint8_t* p = (uint8_t*)lo + c8_byte(off);
        p[0] = (uint8_t)(e4m3x2((x - __half2float(h)) * kC8ActRes, 0.f) & 0xffu);
        p[64] = (uint8_t)(e4m3x2(x * kC8ActVal, 0.f) & 0xffu);
    }
}

// ---- counter-based dropout mask (training branch) ----
// One splitmix64 hash per group of four consecutive elements; element e is KEPT iff the 16-bit field
// (hash(seed, e >> 2) >> (16 * (e & 3))) & 0xffff is >= thr16 = round(p * 65536).  Stateless, so the backward pass
// (and the numpy oracle, tests/) regenerates the mask from (seed, element index) instead of storing it.
__host__ __device__ __forceinline__ uint64_t drop_hash(uint64_t seed, uint64_t group) {
    uint64_t z = seed + (group + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct DropSpec {
    uint64_t seed = 0;
    uint32_t thr16 = 0;   // 0 = no dropout
    float scale = 1.f;    // 1 / (1 - p)
};
inline DropSpec make_drop(float p, uint64_t seed) {
    DropSpec d;
    if (p > 0.f) {
        d.seed = seed;
        d.thr16 = (uint32_t)(p * 65536.f + 0.5f);
        d.scale = 1.f / (1.f - p);
    }
    return d;
}
// keep-scale factors of the four elements 4*group .. 4*group+3
__device__ __forceinline__ float4 drop_scale4(const DropSpec& d, uint64_t group) {
    const uint64_t h = drop_hash(d.seed, group);
    return make_float4(((uint32_t)(h) & 0xffffu) >= d.thr16 ? d.scale : 0.f,
                       ((uint32_t)(h >> 16) & 0xffffu) >= d.thr16 ? d.scale : 0.f,
                       ((uint32_t)(h >> 32) & 0xffffu) >= d.thr16 ? d.scale : 0.f,
                       ((uint32_t)(h >> 48) & 0xffffu) >= d.thr16 ? d.scale : 0.f);
}
__device__ __forceinline__ float drop_scale1(const DropSpec& d, uint64_t e) {
    const uint64_t h = drop_hash(d.seed, e >> 2);
    return ((uint32_t)(h >> (16 * (e & 3))) & 0xffffu) >= d.thr16 ? d.scale : 0.f;
}

// d/dv of nn.GELU() (exact erf form): Phi(v) + v * phi(v)
__device__ __forceinline__ float gelu_grad(float v) {
    return 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v);
}
__device__ __forceinline__ float gelu_grad_fast(float v) {
    const float z = fabsf(v) * 0.70710678118654752440f;
    const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float ex = __expf(-z * z);                       // exp(-v^2 / 2)
    const float e = 1.f - p * t * ex;                      // erf(|v| / sqrt2)
    return 0.5f * (1.f + copysignf(e, v)) + v * 0.39894228040143267794f * ex;
}

enum { RES_ADD = 0, RES_GELU_GRAD = 1 };

struct GemmEpilogue {
    const float* bias = nullptr;      // [N]
    const float* residual = nullptr;  // [M, ldr] fp32 (may alias out.f32)
    int act = ACT_NONE;
    ActOut out;                       // any subset of f32 / hi / lo; row stride ldc
    int ldc = 0;
    int ldr = 0;                      // residual row stride (0 = ldc)
    // ---- training-branch extras (gemm_tc2 and gemm_simt only) ----
    float* pre_f32 = nullptr;         // optional [M, ldc]: acc + bias BEFORE the activation (saved for the backward pass)
    int res_mode = RES_ADD;           // RES_GELU_GRAD: out = (acc + bias) * gelu'(residual[row, col]) (FF1 backward)
    DropSpec drop;                    // dropout on act(acc + bias) before the residual add; element index row * ldc + col
    int split_k = 1;                  // > 1: partial products over K slices, slice s written at out.f32 + s * split_stride
    size_t split_stride = 0;          //      (no bias / act / residual; reduce with splitk_reduce)
    // ---- LayerNorm fusion (gemm_tc2, inference epilogues; gemm_tc2.cu EPI_*) ----
    const float2* ln_stats = nullptr; // (mean, rstd) per A row: the GEMM runs on the RAW rows with gamma folded into W and
    int ln_row_stride = 1;            //   the epilogue applies rstd * (acc - mean * ln_c1[n]) + bias[n] (bias = c2);
    const float* ln_c1 = nullptr;     //   row r's statistics sit at ln_stats[r * ln_row_stride]
    const float2* ln_parts = nullptr; // instead of ln_stats: the kDim / 64 partial (sum, sum of squares) planes the producing
    long long ln_parts_rows = 0;      //   epilogue wrote (plane stride in rows); the epilogue reduces them itself
    float2* stats_partials = nullptr; // [N / 64][M] partial (sum, sum of squares) of the OUTPUT rows (ln_stats_finalize)
    ActOut res_op;                    // residual given in OPERAND format (hi / lo / fmt; row stride ldr) instead of `residual`
    bool qkv_item_layout = false;     // to_qkv for attention_split.cu (N = 1728, hi / lo outputs, M = whole sequences): element
                                      //   (seq * 19 + tok, which * 576 + head * 96 + d) goes to qkv_item_offset(...) — every
                                      //   (sequence, head)'s q, k, v as three contiguous 19 x 96 blocks instead of 192-byte
                                      //   segments strided by a 3456-byte row
};
// offset of (row, col) of the [rows, 1728] to_qkv output in the item layout (in elements of the hi / lo arrays)
__host__ __device__ __forceinline__ size_t qkv_item_offset(int64_t seq, int tok, int which, int head, int d) {
    return (((size_t)seq * kHeads + head) * 3 + which) * (size_t)(kTokens * kHeadDim) + (size_t)tok * kHeadDim + d;
}
// (mean, rstd) of LayerNorm (eps 1e-5) over rows of kDim from the partial sums a gemm_tc2 epilogue wrote
int ln_stats_finalize(const float2* partials, int n_parts, int64_t rows, float2* stats, cudaStream_t s);

// `passes` of the tensor-core GEMMs: the operand format / product scheme
//   1 = bf16 single product, 3 = bf16x3 (hi*hi + lo*hi + hi*lo), 2 = f16c8 (fp16 product + one fp8 correction GEMM over
//   2K, see the f16c8 block above; gemm_tc2 only), 4 = fp16 single product (hi arrays of the f16c8 format; gemm_tc2 only)
enum { TC_BF16 = 1, TC_F16C8 = 2, TC_BF16X3 = 3, TC_F16 = 4 };

// A operand of a GEMM: fp32 (SIMT path) or bf16 hi[/lo] (tensor-core path); W likewise.
struct GemmOperand {
    const float* f32 = nullptr;
    const __nv_bfloat16* hi = nullptr;
    const __nv_bfloat16* lo = nullptr;
    int ld = 0;                       // row stride in elements (0 = K, dense)
};

// C[M,N] = epi(A[M,K] @ W[N,K]^T).  A row stride lda (elements), W row stride K.
int gemm_simt(const float* A, int lda, const float* W, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s);
int gemm_simt_slices(int K, int split_k);  // K slices an ep.split_k request really produces (partials at out.f32 + z * split_stride)
// passes = 1 (bf16) or 3 (bf16x3: hi*hi + lo*hi + hi*lo). Requires K % 64 == 0, N % 4 == 0, A.ld % 8 == 0.
int gemm_tc(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes,
            const GemmEpilogue& ep, cudaStream_t s);
int gemm_tc_init();  // resolves cuTensorMapEncodeTiled, sets smem attributes (idempotent)
// Implicit-GEMM ks x ks convolution on the same kernel: A = NHWC bf16 hi[/lo] activation [B,H,W,Cin] (the INPUT size; Cin % 64
// == 0), W = [N, ks*ks*Cin] with column (kh*ks + kw)*Cin + c; out[(b*Ho + h)*Wo + w, 0..N) fp32 with row stride ldc, Ho / Wo
// by the floor rule (gemm_tc.cu).  stride > 1 uses TMA element strides.
int conv_tc(const GemmOperand& A, const GemmOperand& W, int B, int H, int Wd, int Cin, int N, int ks, int pad, int stride, int passes,
            float* out, int ldc, cudaStream_t s);
// CTA-pair (cta_group::2) version, 256 x 192 tiles; needs N % 192 == 0 (gemm_tc2.cu)
bool gemm_tc2_supported(int N, int K);
int gemm_tc2_slices(int K, int split_k);  // K slices a GemmEpilogue::split_k request really produces
int gemm_tc2(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes, const GemmEpilogue& ep,
             cudaStream_t s);
// Weight-gradient GEMM G[Nw,Kw] = dY[rows,Nw]^T @ X[rows,Kw] with both operands read in place as MN-major tcgen05
// operands (gemm_tn2.cu).  split_k > 1: slice s of the rows is written at out + s * split_stride (gemm_tn2_slices tells
// how many slices a request really produces).  geometry = {LBO, SBO, K advance} bytes, NULL = the production constants.
bool gemm_tn2_supported(int Nw, int Kw, int ld_y, int ld_x);
int gemm_tn2_slices(int rows, int split_k);
int gemm_tn2_mn_tiles(int Nw, int Kw);                                  // output tiles per K slice (mixed 256 / 128-wide)
double gemm_tn2_efficiency(int Nw, int Kw, int slices, int pairs);         // busy fraction of the CTA pairs for a slice count
int gemm_tn2(const GemmOperand& dY, const GemmOperand& X, int Nw, int Kw, int rows, int passes, float* out, int ldc, int split_k,
             size_t split_stride, cudaStream_t s, const uint32_t* geometry = nullptr);
// Convolution form: G[Cout, ks*ks*Cin] = sum over output pixels of dY[b,h,w,:]^T x X[b,h*stride+kh-pad,w*stride+kw-pad,:], both
// NHWC bf16 hi[/lo] (dY [B,H,W,Cout], X [B,Hin,Win,Cin]), no im2col; gemm_tn2_conv_rows = the row count to size split-K with.
int gemm_tn2_conv_rows(int B, int H, int W);
int gemm_tn2_conv(const GemmOperand& dY, const GemmOperand& X, int B, int H, int W, int Hin, int Win, int Cout, int Cin, int ks, int pad,
                  int stride, int passes, float* out, int ldc, int split_k, size_t split_stride, cudaStream_t s);
// picks gemm_tc2 where it applies (env VETO_GEMM_2CTA=0 forces the single-CTA kernel), else gemm_tc
int gemm_tc_auto(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes, const GemmEpilogue& ep,
                 cudaStream_t s);

// LayerNorm over rows of kDim (eps 1e-5, model_veto.py:128) -> out format(s)
int layernorm_rows(const float* x, int64_t ldx, const float* w, const float* b, int64_t rows, const ActOut& out,
                   cudaStream_t s);
// softmax(q k^T * 96^-0.5) v per (sequence, head); qkv fp32 [n_seq*19, 1728] -> out [n_seq*19, 576]
int attention_seq(const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s);
// the same on q, k, v already split into bf16 hi + lo arrays [n_seq*19, 1728] (attention_split.cu; the to_qkv epilogue
// of gemm_tc2 writes them): the inference path of the tensor-core modes with 16-bit-mantissa operands
int attention_seq_split(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, int64_t n_seq, const ActOut& out, bool item_layout,
                        cudaStream_t s);
// tcgen05 version (attention_tc.cu): six sequences per 128-row tile, bf16 hi/lo split when out.lo is given
int attention_tc(const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s);
// the same for the CLS query row only (last encoder layer: only x[:,0] is consumed, model_veto.py:25):
// q_cls fp32 [n_seq, 576]; k, v from qkv [n_seq*19, 1728] cols 576..1727; out [n_seq, 576]
int attention_cls(const float* q_cls, const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s);

int num_sms();

}  // namespace veto
