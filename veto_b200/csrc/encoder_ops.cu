// Row-wise pieces of the relation encoder that are not GEMMs: LayerNorm (PreNorm, model_veto.py:125-132)
// and the 19-token multi-head attention core (Attention.forward, model_veto.py:86-96).  Both are
// HBM/L2-streaming kernels: one warp per row (LayerNorm) or per (sequence, head) (attention), warp
// shuffles for the reductions, vectorised coalesced loads and stores.
#include <stdlib.h>

#include "common.cuh"

namespace veto {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- LayerNorm
// One warp per row of 576 (9 float2 per lane), two-pass mean / variance in registers (eps 1e-5).  Two rows are in
// flight per warp and gamma / beta come from L1 instead of registers, so that 48 warps per SM keep enough bytes in
// flight for HBM (profiles/r1_rows_ncu.txt: the first version sat at 65 % of copy bandwidth with 24 warps per SM).
__global__ void __launch_bounds__(256, 4)
layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
                 int64_t rows, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int fmt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    constexpr int PER = kDim / 64;  // 9
    const float2* w2 = (const float2*)w;
    const float2* b2 = (const float2*)b;
    for (int64_t row0 = 2 * warp0; row0 < rows; row0 += 2 * nwarps) {
        const bool two = row0 + 1 < rows;
        const float2* xr0 = (const float2*)(x + row0 * ldx);
        const float2* xr1 = (const float2*)(x + (two ? row0 + 1 : row0) * ldx);
        float2 v0[PER], v1[PER];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v0[j] = xr0[lane + 32 * j];
            v1[j] = xr1[lane + 32 * j];
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            s0 += v0[j].x + v0[j].y;
            s1 += v1[j].x + v1[j].y;
        }
        const float mean0 = warp_sum(s0) * (1.f / kDim), mean1 = warp_sum(s1) * (1.f / kDim);
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const float ax = v0[j].x - mean0, ay = v0[j].y - mean0, bx = v1[j].x - mean1, by = v1[j].y - mean1;
            q0 += ax * ax + ay * ay;
            q1 += bx * bx + by * by;
        }
        const float rstd0 = 1.f / sqrtf(warp_sum(q0) * (1.f / kDim) + 1e-5f);
        const float rstd1 = 1.f / sqrtf(warp_sum(q1) * (1.f / kDim) + 1e-5f);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (r == 1 && !two) break;
            const float mean = r ? mean1 : mean0, rstd = r ? rstd1 : rstd0;
            const size_t o = (size_t)(row0 + r) * kDim;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const float2 v = r ? v1[j] : v0[j];
                const float2 wv = __ldg(w2 + lane + 32 * j), bv = __ldg(b2 + lane + 32 * j);
                float2 y;
                y.x = (v.x - mean) * rstd * wv.x + bv.x;
                y.y = (v.y - mean) * rstd * wv.y + bv.y;
                const size_t e = o + 2 * (lane + 32 * j);
                if (out_f32) *(float2*)(out_f32 + e) = y;
                if (out_hi) {
                    if (fmt == FMT_F16C8) {
                        store_act2_f16c8(out_hi, out_lo, e, y.x, y.y);
                    } else {
                        uint32_t hh, ll;
                        split_pair(y.x, y.y, hh, ll);
                        *(uint32_t*)(out_hi + e) = hh;
                        if (out_lo) *(uint32_t*)(out_lo + e) = ll;
                    }
                }
            }
        }
    }
}

// Row statistics for the LayerNorm-fused GEMMs: the epilogue that produced x wrote n_parts partial (sum, sum of squares)
// per row (one per 64 columns); mean and 1 / sqrt(var + eps) in a fixed summation order.
__global__ void ln_stats_finalize_kernel(const float2* __restrict__ partials, int n_parts, int64_t rows, float2* __restrict__ stats) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int p = 0; p < n_parts; ++p) {
            const float2 v = __ldg(partials + (size_t)p * rows + r);
            s += v.x;
            q += v.y;
        }
        const float mean = s * (1.f / kDim);
        const float var = fmaxf(q * (1.f / kDim) - mean * mean, 0.f);
        stats[r] = make_float2(mean, 1.f / sqrtf(var + 1e-5f));
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


// ---------------------------------------------------------------- attention on the tensor cores (mma.sync)
// One warp per (sequence, head).  S = Q K^T (19x19, padded to 32x24) and O = P V (19x96) run as m16n8k16 bf16
// warp MMAs with fp32 accumulation; with SPLIT every fp32 operand is split into bf16 hi + lo and each product
// takes three MMAs (hi*hi + lo*hi + hi*lo), which keeps ~16 mantissa bits — the same scheme as the GEMMs.
// Fragments are loaded straight from global memory in MMA layout (float2 / 32-byte row segments, every sector
// fully used); the softmax runs on the accumulator registers (a row lives in the 4 lanes of a quad) and P is
// re-used from registers as the A operand of the second product, flash-attention style.  No shared memory.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm(  // not volatile: a pure function of its operands, the scheduler may interleave independent accumulator chains
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// shared-memory staging of one (sequence, head): row strides chosen so that the fragment reads are conflict-free
// (Q/K: 64-bit reads at [g][2t] need stride = 8 mod 32 words; V: 32-bit reads at [2t][g] need 2*stride = 8 mod 32)
constexpr int ATT_QK_STRIDE = 104, ATT_V_STRIDE = 100;  // floats (416 B / 400 B: 16-byte aligned for cp.async)
constexpr int ATT_ITEM_FLOATS = 2 * kTokens * ATT_QK_STRIDE + kTokens * ATT_V_STRIDE;
constexpr int ATT_MMA_WARPS = 8;
constexpr int ATT_MMA_SMEM = ATT_MMA_WARPS * ATT_ITEM_FLOATS * (int)sizeof(float);

template <bool SPLIT>
__global__ void __launch_bounds__(ATT_MMA_WARPS * 32)
attention_mma_kernel(const float* __restrict__ qkv, int64_t n_seq, float* out_f32, __nv_bfloat16* out_hi,
                     __nv_bfloat16* out_lo) {
    extern __shared__ float4 att_smem[];
    constexpr int LD = 3 * kDim;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* sQ = reinterpret_cast<float*>(att_smem) + (threadIdx.x >> 5) * ATT_ITEM_FLOATS;
    float* sK = sQ + kTokens * ATT_QK_STRIDE;
    float* sV = sK + kTokens * ATT_QK_STRIDE;
    const float scale = 0.10206207261596575f;  // 96 ** -0.5 (model_veto.py:74)
    const int64_t items = n_seq * kHeads;
    for (int64_t item = (int64_t)blockIdx.x * ATT_MMA_WARPS + (threadIdx.x >> 5); item < items;
         item += (int64_t)gridDim.x * ATT_MMA_WARPS) {
        const int64_t seq = item / kHeads;
        const int h = (int)(item - seq * kHeads);
        const float* base = qkv + (size_t)seq * kTokens * LD + h * kHeadDim;
        // stage q, k, v of this (sequence, head): 57 rows of 384 B, 16-byte cp.async, fully coalesced
        __syncwarp();
        for (int idx = lane; idx < 3 * kTokens * (kHeadDim / 4); idx += 32) {
            const int m = idx / (kHeadDim / 4), c = idx - m * (kHeadDim / 4);
            const int which = m / kTokens, row = m - which * kTokens;
            const float* src = base + (size_t)row * LD + which * kDim + c * 4;
            float* dst = (which == 0 ? sQ + row * ATT_QK_STRIDE : which == 1 ? sK + row * ATT_QK_STRIDE : sV + row * ATT_V_STRIDE) + c * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        auto ld2 = [&](int row, int col, int which) -> float2 {  // which: 0 q, 1 k
            return (row < kTokens) ? *(const float2*)((which == 0 ? sQ : sK) + row * ATT_QK_STRIDE + col) : make_float2(0.f, 0.f);
        };

        // ---- S = Q K^T : 2 m-tiles (rows 16mt+g, +8) x 3 n-tiles (keys 8nt+g)
        float S[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) S[mt][nt][e] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < kHeadDim / 16; ++ks) {
            const int d0 = ks * 16 + 2 * t;
            uint32_t qh[2][4], ql[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float2 x0 = ld2(16 * mt + g, d0, 0), x1 = ld2(16 * mt + g + 8, d0, 0);
                const float2 x2 = ld2(16 * mt + g, d0 + 8, 0), x3 = ld2(16 * mt + g + 8, d0 + 8, 0);
                split_pair(x0.x, x0.y, qh[mt][0], ql[mt][0]);
                split_pair(x1.x, x1.y, qh[mt][1], ql[mt][1]);
                split_pair(x2.x, x2.y, qh[mt][2], ql[mt][2]);
                split_pair(x3.x, x3.y, qh[mt][3], ql[mt][3]);
            }
            uint32_t kh[3][2], kl[3][2];
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                const float2 y0 = ld2(8 * nt + g, d0, 1), y1 = ld2(8 * nt + g, d0 + 8, 1);
                split_pair(y0.x, y0.y, kh[nt][0], kl[nt][0]);
                split_pair(y1.x, y1.y, kh[nt][1], kl[nt][1]);
            }
            // split terms outermost: consecutive MMAs hit the 6 independent accumulators, dependent ones are 6 apart
#pragma unroll
            for (int term = 0; term < (SPLIT ? 3 : 1); ++term)
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        mma_bf16_16816(S[mt][nt], term == 1 ? ql[mt] : qh[mt], term == 2 ? kl[nt] : kh[nt]);
        }

        // ---- softmax over the 19 keys of each row (a row = the 4 lanes of a quad; e 0,1 -> row g, e 2,3 -> row g+8)
        uint32_t ph[2][2][4], pl[2][2][4];  // [mt][ks2][a-fragment]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                float m = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * nt + 2 * t + e;
                        float v = S[mt][nt][2 * hrow + e] * scale;
                        v = (col < kTokens) ? v : -INFINITY;
                        S[mt][nt][2 * hrow + e] = v;
                        m = fmaxf(m, v);
                    }
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float p = expf(S[mt][nt][2 * hrow + e] - m);  // exp(-inf) = 0 for the padding keys
                        S[mt][nt][2 * hrow + e] = p;
                        sum += p;
                    }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float inv = 1.f / sum;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) S[mt][nt][2 * hrow + e] *= inv;
            }
            // P as A fragments of the second product: k-step ks2 covers keys 16ks2..16ks2+15 = n-tiles 2ks2, 2ks2+1
#pragma unroll
            for (int ks2 = 0; ks2 < 2; ++ks2) {
                split_pair(S[mt][2 * ks2][0], S[mt][2 * ks2][1], ph[mt][ks2][0], pl[mt][ks2][0]);
                split_pair(S[mt][2 * ks2][2], S[mt][2 * ks2][3], ph[mt][ks2][1], pl[mt][ks2][1]);
                split_pair(S[mt][2 * ks2 + 1][0], S[mt][2 * ks2 + 1][1], ph[mt][ks2][2], pl[mt][ks2][2]);  // n-tile 3 is all zero
                split_pair(S[mt][2 * ks2 + 1][2], S[mt][2 * ks2 + 1][3], ph[mt][ks2][3], pl[mt][ks2][3]);
            }
        }

        // ---- O = P V : 12 n-tiles of 8 head dims, 2 k-steps of 16 keys; 4 n-tiles at a time to bound registers
        auto ldv = [&](int key, int d) -> float { return (key < kTokens) ? sV[key * ATT_V_STRIDE + d] : 0.f; };
#pragma unroll 1
        for (int nb = 0; nb < 3; ++nb) {
            float O[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) O[mt][j][e] = 0.f;
#pragma unroll
            for (int ks2 = 0; ks2 < 2; ++ks2) {
                const int k0 = 16 * ks2 + 2 * t;
                uint32_t vh[4][2], vl[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = 8 * (4 * nb + j) + g;
                    split_pair(ldv(k0, d), ldv(k0 + 1, d), vh[j][0], vl[j][0]);
                    split_pair(ldv(k0 + 8, d), ldv(k0 + 9, d), vh[j][1], vl[j][1]);
                }
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
                            mma_bf16_16816(O[mt][j], term == 1 ? pl[mt][ks2] : ph[mt][ks2], term == 2 ? vl[j] : vh[j]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int row = 16 * mt + g + 8 * hrow;
                    if (row < kTokens) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float v0 = O[mt][j][2 * hrow], v1 = O[mt][j][2 * hrow + 1];
                            const size_t o = ((size_t)seq * kTokens + row) * kDim + h * kHeadDim + 8 * (4 * nb + j) + 2 * t;
                            if (out_f32) *(float2*)(out_f32 + o) = make_float2(v0, v1);
                            if (out_hi) {
                                uint32_t hh, ll;
                                split_pair(v0, v1, hh, ll);
                                *(uint32_t*)(out_hi + o) = hh;
                                if (out_lo) *(uint32_t*)(out_lo + o) = ll;
                            }
                        }
                    }
                }
        }
    }
}


// ---------------------------------------------------------------- attention on warp MMAs, two warps per item
// Same arithmetic as attention_mma_kernel, but TWO warps share the staged Q, K, V of a (sequence, head): warp w takes
// k-steps 3w..3w+2 of S = Q K^T (the partial accumulators meet through a 19x20 tile each) and 6 of the 12 output
// n-tiles of O = P V; both run the softmax on the full S.  16 warps per SM instead of 8 for nearly the same shared
// memory: the one-warp kernel is staging-latency bound (cp.async -> wait -> compute -> store per item, 2 warps per
// scheduler), which the same change fixed in the backward kernel (train.cu: 696 -> 498 us).
constexpr int ATT2_PAIRS = 8;
constexpr int ATT2_TILE = kTokens * 20;
constexpr int ATT2_ITEM_FLOATS = ATT_ITEM_FLOATS + 2 * ATT2_TILE;
constexpr int ATT2_THREADS = ATT2_PAIRS * 64;
constexpr int ATT2_SMEM = ATT2_PAIRS * ATT2_ITEM_FLOATS * (int)sizeof(float);  // 213,952 B

template <bool SPLIT>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attention_mma2_kernel(const float* __restrict__ qkv, int64_t n_seq, float* out_f32, __nv_bfloat16* out_hi,
                      __nv_bfloat16* out_lo, int fmt) {
    extern __shared__ float4 att_smem[];
    constexpr int LD = 3 * kDim;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = threadIdx.x >> 6, w = (threadIdx.x >> 5) & 1, lane64 = threadIdx.x & 63;
    float* sQ = reinterpret_cast<float*>(att_smem) + pair * ATT2_ITEM_FLOATS;
    float* sK = sQ + kTokens * ATT_QK_STRIDE;
    float* sV = sK + kTokens * ATT_QK_STRIDE;
    float* myS = sV + kTokens * ATT_V_STRIDE + w * ATT2_TILE;
    const float* otherS = sV + kTokens * ATT_V_STRIDE + (w ^ 1) * ATT2_TILE;
    const float scale = 0.10206207261596575f;  // 96 ** -0.5 (model_veto.py:74)
    const int64_t items = n_seq * kHeads;
    auto pair_bar = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };
    for (int64_t item = (int64_t)blockIdx.x * ATT2_PAIRS + pair; item < items; item += (int64_t)gridDim.x * ATT2_PAIRS) {
        const int64_t seq = item / kHeads;
        const int h = (int)(item - seq * kHeads);
        const float* base = qkv + (size_t)seq * kTokens * LD + h * kHeadDim;
        for (int idx = lane64; idx < 3 * kTokens * (kHeadDim / 4); idx += 64) {
            const int m = idx / (kHeadDim / 4), c = idx - m * (kHeadDim / 4);
            const int which = m / kTokens, row = m - which * kTokens;
            const float* src = base + (size_t)row * LD + which * kDim + c * 4;
            float* dst = (which == 0 ? sQ + row * ATT_QK_STRIDE : which == 1 ? sK + row * ATT_QK_STRIDE : sV + row * ATT_V_STRIDE) + c * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        pair_bar();  // (1) staged operands visible to both warps
        auto ld2 = [&](int row, int col, int which) -> float2 {  // which: 0 q, 1 k
            return (row < kTokens) ? *(const float2*)((which == 0 ? sQ : sK) + row * ATT_QK_STRIDE + col) : make_float2(0.f, 0.f);
        };

        // ---- partial S = Q K^T over this warp's three k-steps: 2 m-tiles x 3 n-tiles
        float S[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) S[mt][nt][e] = 0.f;
#pragma unroll 1
        for (int ks = 3 * w; ks < 3 * w + 3; ++ks) {
            const int d0 = ks * 16 + 2 * t;
            uint32_t qh[2][4], ql[2][4], kh[3][2], kl[3][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float2 x0 = ld2(16 * mt + g, d0, 0), x1 = ld2(16 * mt + g + 8, d0, 0);
                const float2 x2 = ld2(16 * mt + g, d0 + 8, 0), x3 = ld2(16 * mt + g + 8, d0 + 8, 0);
                split_pair(x0.x, x0.y, qh[mt][0], ql[mt][0]);
                split_pair(x1.x, x1.y, qh[mt][1], ql[mt][1]);
                split_pair(x2.x, x2.y, qh[mt][2], ql[mt][2]);
                split_pair(x3.x, x3.y, qh[mt][3], ql[mt][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                const float2 y0 = ld2(8 * nt + g, d0, 1), y1 = ld2(8 * nt + g, d0 + 8, 1);
                split_pair(y0.x, y0.y, kh[nt][0], kl[nt][0]);
                split_pair(y1.x, y1.y, kh[nt][1], kl[nt][1]);
            }
#pragma unroll
            for (int term = 0; term < (SPLIT ? 3 : 1); ++term)
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        mma_bf16_16816(S[mt][nt], term == 1 ? ql[mt] : qh[mt], term == 2 ? kl[nt] : kh[nt]);
        }
        // ---- exchange the partial scores (e 0,1 -> row 16mt+g, e 2,3 -> row +8; cols 8nt+2t+e)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
                    if (row < kTokens && col < kTokens) myS[row * 20 + col] = S[mt][nt][e];
                }
        pair_bar();  // (2)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
                    if (row < kTokens && col < kTokens) S[mt][nt][e] += otherS[row * 20 + col];
                }

        // ---- softmax over the 19 keys of each row (a row = the 4 lanes of a quad)
        uint32_t ph[2][2][4], pl[2][2][4];  // [mt][ks2][a-fragment]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                float m = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * nt + 2 * t + e;
                        float v = S[mt][nt][2 * hrow + e] * scale;
                        v = (col < kTokens) ? v : -INFINITY;
                        S[mt][nt][2 * hrow + e] = v;
                        m = fmaxf(m, v);
                    }
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float p = expf(S[mt][nt][2 * hrow + e] - m);
                        S[mt][nt][2 * hrow + e] = p;
                        sum += p;
                    }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float inv = 1.f / sum;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) S[mt][nt][2 * hrow + e] *= inv;
            }
            split_pair(S[mt][0][0], S[mt][0][1], ph[mt][0][0], pl[mt][0][0]);
            split_pair(S[mt][0][2], S[mt][0][3], ph[mt][0][1], pl[mt][0][1]);
            split_pair(S[mt][1][0], S[mt][1][1], ph[mt][0][2], pl[mt][0][2]);
            split_pair(S[mt][1][2], S[mt][1][3], ph[mt][0][3], pl[mt][0][3]);
            split_pair(S[mt][2][0], S[mt][2][1], ph[mt][1][0], pl[mt][1][0]);
            split_pair(S[mt][2][2], S[mt][2][3], ph[mt][1][1], pl[mt][1][1]);
            ph[mt][1][2] = pl[mt][1][2] = ph[mt][1][3] = pl[mt][1][3] = 0u;  // keys 24..31 do not exist
        }

        // ---- O = P V : this warp's 6 n-tiles of 8 head dims (blocks 2w, 2w+1 of three), 2 k-steps of 16 keys
        auto ldv = [&](int key, int d) -> float { return (key < kTokens) ? sV[key * ATT_V_STRIDE + d] : 0.f; };
#pragma unroll 1
        for (int blk = 2 * w; blk < 2 * w + 2; ++blk) {
            float O[2][3][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) O[mt][j][e] = 0.f;
#pragma unroll
            for (int ks2 = 0; ks2 < 2; ++ks2) {
                const int k0 = 16 * ks2 + 2 * t;
                uint32_t vh[3][2], vl[3][2];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int d = 8 * (3 * blk + j) + g;
                    split_pair(ldv(k0, d), ldv(k0 + 1, d), vh[j][0], vl[j][0]);
                    split_pair(ldv(k0 + 8, d), ldv(k0 + 9, d), vh[j][1], vl[j][1]);
                }
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
                            mma_bf16_16816(O[mt][j], term == 1 ? pl[mt][ks2] : ph[mt][ks2], term == 2 ? vl[j] : vh[j]);
            }
            // the accumulator fragments own 2 columns of 8 rows each: through the Q tile (free since barrier 2) so that
            // the global stores below are whole 16-byte / 8-byte pieces of contiguous rows
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int row = 16 * mt + g + 8 * hrow;
                    if (row < kTokens) {
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            *(float2*)(sQ + row * ATT_QK_STRIDE + 8 * (3 * blk + j) + 2 * t) =
                                make_float2(O[mt][j][2 * hrow], O[mt][j][2 * hrow + 1]);
                    }
                }
        }
        pair_bar();  // (3) the 19 x 96 output tile is complete; V is no longer read
        for (int idx = lane64; idx < kTokens * (kHeadDim / 4); idx += 64) {
            const int row = idx / (kHeadDim / 4), c4 = idx - row * (kHeadDim / 4);
            const float4 v = *(const float4*)(sQ + row * ATT_QK_STRIDE + 4 * c4);
            const size_t o = ((size_t)seq * kTokens + row) * kDim + h * kHeadDim + 4 * c4;
            if (out_f32) *(float4*)(out_f32 + o) = v;
            if (out_hi) {
                if (fmt == FMT_F16C8) {
                    store_act4_f16c8(out_hi, out_lo, o, v);
                } else {
                    uint2 hh, ll;
                    split_pair(v.x, v.y, hh.x, ll.x);
                    split_pair(v.z, v.w, hh.y, ll.y);
                    *(uint2*)(out_hi + o) = hh;
                    if (out_lo) *(uint2*)(out_lo + o) = ll;
                }
            }
        }
        pair_bar();  // (4) the tile is stored before the next item's cp.async overwrites it
    }
}


// ---------------------------------------------------------------- CLS-only attention (last layer)
// One warp per (sequence, head): lane j < 19 scores key j against the single CLS query, the softmax is a pair of
// warp reductions, and lane L accumulates output dims L, L+32, L+64 with p_j broadcast by shuffle.
__global__ void __launch_bounds__(256)
attention_cls_kernel(const float* __restrict__ q_cls, const float* __restrict__ qkv, int64_t n_seq, float* out_f32,
                     __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int fmt) {
    constexpr int LD = 3 * kDim;
    const int lane = threadIdx.x & 31;
    const float scale = 0.10206207261596575f;
    const int64_t items = n_seq * kHeads;
    for (int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); item < items; item += (int64_t)gridDim.x * 8) {
        const int64_t seq = item / kHeads;
        const int h = (int)(item - seq * kHeads);
        const float4* q = (const float4*)(q_cls + (size_t)seq * kDim + h * kHeadDim);
        const float* kbase = qkv + (size_t)seq * kTokens * LD + kDim + h * kHeadDim;
        const float* vbase = kbase + kDim;
        float s = 0.f;
        if (lane < kTokens) {
            const float4* kr = (const float4*)(kbase + (size_t)lane * LD);
#pragma unroll 6
            for (int c = 0; c < kHeadDim / 4; ++c) {
                const float4 a = __ldg(q + c), b = __ldg(kr + c);
                s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
            }
        }
        s = (lane < kTokens) ? s * scale : -INFINITY;
        float m = s;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float p = expf(s - m);  // 0 for the padding lanes
        const float inv = 1.f / warp_sum(p);
        p *= inv;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kTokens; ++j) {
            const float pj = __shfl_sync(0xffffffffu, p, j);
            const float* vr = vbase + (size_t)j * LD + lane;
            acc[0] = fmaf(pj, __ldg(vr), acc[0]);
            acc[1] = fmaf(pj, __ldg(vr + 32), acc[1]);
            acc[2] = fmaf(pj, __ldg(vr + 64), acc[2]);
        }
        const size_t o = (size_t)seq * kDim + h * kHeadDim + lane;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (out_f32) out_f32[o + 32 * k] = acc[k];
            if (out_hi) {
                if (fmt == FMT_F16C8) {
                    store_act1_f16c8(out_hi, out_lo, o + 32 * k, acc[k]);
                } else {
                    __nv_bfloat16 hh, ll;
                    split_bf16(acc[k], hh, ll);
                    out_hi[o + 32 * k] = hh;
                    if (out_lo) out_lo[o + 32 * k] = ll;
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
    const int64_t blocks_needed = (rows + 15) / 16;  // 8 warps x 2 rows per block iteration
    const int64_t cap = (int64_t)num_sms() * 4;
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    layernorm_kernel<<<grid, 256, 0, s>>>(x, ldx, w, b, rows, out.f32, out.hi, out.lo, out.fmt);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int ln_stats_finalize(const float2* partials, int n_parts, int64_t rows, float2* stats, cudaStream_t s) {
    if (rows <= 0) return VETO_OK;
    const int64_t blocks = (rows + 255) / 256;
    const int grid = (int)(blocks < (int64_t)num_sms() * 8 ? blocks : (int64_t)num_sms() * 8);
    ln_stats_finalize_kernel<<<grid, 256, 0, s>>>(partials, n_parts, rows, stats);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int attention_seq(const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s) {
    if (n_seq <= 0) return VETO_OK;
    if (out.hi) {  // tensor-core modes: bf16x3 (hi + lo outputs) or single-pass bf16 (hi only)
        static int use_tc = -1;
        if (use_tc < 0) {
            // "tc" selects the tcgen05 kernel (attention_tc.cu): correct, but its load/convert phase is not yet
            // overlapped with the MMAs and it is slower (260 us vs 115 us per 1994 sequences), so warp-MMA is the default
            const char* e = getenv("VETO_ATTENTION");
            use_tc = (e && e[0] == 't') ? 1 : 0;
        }
        if (use_tc) return attention_tc(qkv, n_seq, out, s);
        static int one_warp = -1;
        if (one_warp < 0) one_warp = getenv("VETO_ATTENTION_ONE_WARP") ? 1 : 0;  // diagnosis: the one-warp-per-item kernel
        if (!one_warp) {
            static DeviceOnce attr2_set;
            if (attr2_set.pending()) {
                VETO_CUDA(cudaFuncSetAttribute(attention_mma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
                VETO_CUDA(cudaFuncSetAttribute(attention_mma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
                attr2_set.done();
            }
            const int64_t blocks2 = (n_seq * kHeads + ATT2_PAIRS - 1) / ATT2_PAIRS;
            const int grid2 = (int)(blocks2 < (int64_t)num_sms() ? blocks2 : (int64_t)num_sms());
            // the f16 mode (fp16 output, no residual bytes) keeps the split products INSIDE the attention core
            if (out.lo || out.fmt == FMT_F16C8) attention_mma2_kernel<true><<<grid2, ATT2_THREADS, ATT2_SMEM, s>>>(qkv, n_seq, out.f32, out.hi, out.lo, out.fmt);
            else attention_mma2_kernel<false><<<grid2, ATT2_THREADS, ATT2_SMEM, s>>>(qkv, n_seq, out.f32, out.hi, out.lo, out.fmt);
            VETO_LAUNCH_CHECK();
            return VETO_OK;
        }
        static DeviceOnce mma_attr_set;
        if (mma_attr_set.pending()) {
            VETO_CUDA(cudaFuncSetAttribute(attention_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MMA_SMEM));
            VETO_CUDA(cudaFuncSetAttribute(attention_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MMA_SMEM));
            mma_attr_set.done();
        }
        const int64_t blocks = (n_seq * kHeads + ATT_MMA_WARPS - 1) / ATT_MMA_WARPS;
        const int64_t capb = (int64_t)num_sms();  // one 187 KB CTA per SM
        const int gridb = (int)(blocks < capb ? blocks : capb);
        if (out.lo) attention_mma_kernel<true><<<gridb, ATT_MMA_WARPS * 32, ATT_MMA_SMEM, s>>>(qkv, n_seq, out.f32, out.hi, out.lo);
        else attention_mma_kernel<false><<<gridb, ATT_MMA_WARPS * 32, ATT_MMA_SMEM, s>>>(qkv, n_seq, out.f32, out.hi, out.lo);
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set.done();
    }
    const int64_t cap = (int64_t)num_sms() * 2;
    const int grid = (int)(n_seq < cap ? n_seq : cap);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, s>>>(qkv, n_seq, out.f32, out.hi, out.lo);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int attention_cls(const float* q_cls, const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s) {
    if (n_seq <= 0) return VETO_OK;
    const int64_t blocks = (n_seq * kHeads + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    attention_cls_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(q_cls, qkv, n_seq, out.f32, out.hi, out.lo, out.fmt);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
