// Training-branch kernels of the relation head that are not GEMMs: the cross-entropy loss of
// VETOPredictor.forward (roi_relation_predictors.py:4134-4135, nn.CrossEntropyLoss(weight)), and the backward twins
// of the row-wise forward kernels — LayerNorm (model_veto.py:125-132), the 19-token attention core (:86-96), the
// relation-token gather (tokens.cu; model_veto.py:52-64), the per-box embeddings (box_stage.cu;
// roi_relation_predictors.py:4042-4047, 4086-4102) — plus the data-movement helpers the weight-gradient GEMMs need
// (transposes into K-major bf16 hi/lo operands, split-K reduction, deterministic column sums).
//
// Every reduction here has a fixed order (no floating-point atomics): a training step is bitwise reproducible.
#include "stages.cuh"
#include <stdlib.h>

#include "train.cuh"

namespace veto {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

int grid_cap(size_t blocks, int per_sm) {
    const size_t cap = (size_t)num_sms() * per_sm;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

__device__ __forceinline__ float load_act(const float* f32, const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t e) {
    if (f32) return f32[e];
    float v = __bfloat162float(hi[e]);
    if (lo) v += __bfloat162float(lo[e]);
    return v;
}
__device__ __forceinline__ void store_act(float* f32, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t e, float v) {
    if (f32) f32[e] = v;
    if (hi) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        hi[e] = h;
        if (lo) lo[e] = l;
    }
}

// ---------------------------------------------------------------- cross entropy
// loss = sum_i w[y_i] * (logsumexp(z_i) - z_i[y_i]) / sum_i w[y_i]   (nn.CrossEntropyLoss(weight), mean reduction)
// over the C columns [col0, col0 + C) of a logits matrix with row stride ld (one head of the MEET group classifier, or
// the whole row for the single rel_out head); rows with a negative label are not part of this head's loss (the rows
// MEET's group sampling left out, roi_relation_predictors.py:3842-3846): weight 0, zero gradient.
// warp per row: row_loss[i] = w * nll, row_w[i] = w
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ logits, int ld, int col0, int C, const int64_t* __restrict__ labels,
               const float* __restrict__ weight, int64_t rows, float* __restrict__ row_loss, float* __restrict__ row_w) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int64_t y = labels[r];
        if (y < 0) {
            if (lane == 0) { row_loss[r] = 0.f; row_w[r] = 0.f; }
            continue;
        }
        const float* z = logits + r * ld + col0;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
        m = warp_max(m);
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += expf(z[c] - m);
        s = warp_sum(s);
        if (lane == 0) {
            const float w = weight ? weight[y] : 1.f;
            row_loss[r] = w * (logf(s) + m - z[y]);
            row_w[r] = w;
        }
    }
}
// single block: fixed-order tree sums of row_loss and row_w -> out[0] = loss, out[1] = 1 / sum(w)
// (no row in the head: loss = NaN like torch's mean over an empty selection, and a zero gradient scale)
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_w, int64_t rows, float* __restrict__ loss_out,
                   float* __restrict__ inv_w_out) {
    __shared__ double sl[1024], sw[1024];
    double a = 0.0, b = 0.0;
    for (int64_t r = threadIdx.x; r < rows; r += 1024) {
        a += row_loss[r];
        b += row_w[r];
    }
    sl[threadIdx.x] = a;
    sw[threadIdx.x] = b;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sl[threadIdx.x] += sl[threadIdx.x + o];
            sw[threadIdx.x] += sw[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *loss_out = (float)(sl[0] / sw[0]);
        *inv_w_out = sw[0] != 0.0 ? (float)(1.0 / sw[0]) : 0.f;
    }
}
// dlogits[i, col0 + c] = w[y_i] * (softmax(z_i)[c] - [c == y_i]) / sum(w)
__global__ void __launch_bounds__(256)
ce_grad_kernel(const float* __restrict__ logits, int ld, int col0, int C, const int64_t* __restrict__ labels,
               const float* __restrict__ weight, const float* __restrict__ inv_w, int64_t rows, float* __restrict__ dlogits) {
    const int lane = threadIdx.x & 31;
    const float iw = *inv_w;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int64_t y = labels[r];
        float* d = dlogits + r * ld + col0;
        if (y < 0) {
            for (int c = lane; c < C; c += 32) d[c] = 0.f;
            continue;
        }
        const float* z = logits + r * ld + col0;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
        m = warp_max(m);
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += expf(z[c] - m);
        s = warp_sum(s);
        const float w = (weight ? weight[y] : 1.f) * iw;
        const float inv = 1.f / s;
        for (int c = lane; c < C; c += 32) d[c] = w * (expf(z[c] - m) * inv - (c == y ? 1.f : 0.f));
    }
}

// ---------------------------------------------------------------- column sums (bias gradients)
// stage 1: block (128 columns) x row chunk -> partial[chunk, cols]; stage 2: fixed-order sum over chunks
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const float* f32, const __nv_bfloat16* hi, const __nv_bfloat16* lo, int64_t ld, int64_t rows, int cols,
                      int rows_per_chunk, float* __restrict__ partial) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    float a = 0.f;
    for (int64_t r = r0; r < r1; ++r) a += load_act(f32, hi, lo, (size_t)(r * ld + c));
    partial[(size_t)blockIdx.y * cols + c] = a;
}
__global__ void __launch_bounds__(128)
colsum_final_kernel(const float* __restrict__ partial, int chunks, int cols, float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= cols) return;
    float a = 0.f;
    for (int k = 0; k < chunks; ++k) a += partial[(size_t)k * cols + c];
    out[c] = accumulate ? out[c] + a : a;
}

// ---------------------------------------------------------------- transposes
// dst[c, r] = op(src[r, c]) (* dropout), r < rows; columns rows..rows_pad of dst are zero-filled (K padding of the
// weight-gradient GEMMs).  Optionally also writes the (dropped) values row-major into `rm` (same shape as src, ld cols).
constexpr int TR_OP_NONE = 0, TR_OP_GELU = 1;
__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int cols, int op, DropSpec drop,
                     int64_t drop_ld, float* t_f32, __nv_bfloat16* t_hi, __nv_bfloat16* t_lo, int64_t ld_dst, int64_t rows_pad,
                     float* rm_f32, __nv_bfloat16* rm_hi, __nv_bfloat16* rm_lo, int64_t ld_rm) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = r0 + ty + 8 * k;
        const int c = c0 + tx;
        float v = 0.f;
        if (r < rows && c < cols) {
            v = src[r * ld_src + c];
            if (op == TR_OP_GELU) v = apply_act(v, ACT_GELU);
            if (drop.thr16) v *= drop_scale1(drop, (uint64_t)(r * drop_ld + c));
            store_act(rm_f32, rm_hi, rm_lo, (size_t)(r * ld_rm + c), v);
        }
        tile[ty + 8 * k][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k;
        const int64_t r = r0 + tx;
        if (c < cols && r < rows_pad) store_act(t_f32, t_hi, t_lo, (size_t)((int64_t)c * ld_dst + r), tile[tx][ty + 8 * k]);
    }
}
struct TransposeJobs {
    TransposeJob j[kMaxTransposeJobs];
};
__global__ void __launch_bounds__(256) transpose_f32_multi_kernel(const __grid_constant__ TransposeJobs jobs) {
    __shared__ float tile[32][33];
    const TransposeJob& jb = jobs.j[blockIdx.z];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (r0 >= jb.rows || c0 >= jb.cols) return;  // the grid covers the largest matrix
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < jb.rows && c < jb.cols) ? __ldg(jb.src + (size_t)r * jb.cols + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (c < jb.cols && r < jb.rows) store_act(jb.dst.f32, jb.dst.hi, jb.dst.lo, (size_t)c * jb.rows + r, tile[tx][ty + 8 * k]);
    }
}
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld_src, int64_t rows, int cols,
                      __nv_bfloat16* __restrict__ dst, int64_t ld_dst, int64_t rows_pad) {
    __shared__ __nv_bfloat16 tile[64][66];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * 64;
    const int c0 = blockIdx.x * 64;
    // read: thread (ty, tx) moves the bf16 pair (row ty + 8k, cols 2tx, 2tx+1)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t r = r0 + ty + 8 * k;
        const int c = c0 + 2 * tx;
        __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
        if (r < rows) {
            if (c + 1 < cols) v = *(const __nv_bfloat162*)(src + r * ld_src + c);
            else if (c < cols) v.x = src[r * ld_src + c];
        }
        tile[ty + 8 * k][2 * tx] = v.x;
        tile[ty + 8 * k][2 * tx + 1] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + ty + 8 * k;
        const int64_t r = r0 + 2 * tx;
        if (c < cols && r < rows_pad) {
            __nv_bfloat162 v;
            v.x = tile[2 * tx][ty + 8 * k];
            v.y = tile[2 * tx + 1][ty + 8 * k];
            *(__nv_bfloat162*)(dst + (int64_t)c * ld_dst + r) = v;  // ld_dst and rows_pad are even
        }
    }
}

// out[e] = sum_s partial[s, e] (fixed order), optionally into the un-packed layout of a factored weight
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int slices, size_t n, size_t stride, float* __restrict__ out) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int s = 0; s < slices; ++s) a += partial[(size_t)s * stride + e];
        out[e] = a;
    }
}

// x[e] *= keep-scale(e)
__global__ void __launch_bounds__(256)
dropout_kernel(float* __restrict__ x, size_t n4, DropSpec drop) {
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
        float4 v = ((float4*)x)[g];
        const float4 d = drop_scale4(drop, g);
        v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
        ((float4*)x)[g] = v;
    }
}

// fp32 -> activation storage format (no transpose), optionally through the dropout mask of element index e; 4 per thread
__global__ void __launch_bounds__(256)
convert_kernel(const float* __restrict__ src, size_t n4, DropSpec drop, float* f32, __nv_bfloat16* hi, __nv_bfloat16* lo, int act) {
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg((const float4*)src + g);
        if (act != ACT_NONE) {   // the activation of the forward epilogue that produced the saved pre-activation: the
            if (hi) {            // tensor-core GEMMs apply apply_act_tc, the fp32 SIMT GEMM apply_act (bit-identical re-computation)
                v.x = apply_act_tc(v.x, act); v.y = apply_act_tc(v.y, act); v.z = apply_act_tc(v.z, act); v.w = apply_act_tc(v.w, act);
            } else {
                v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
            }
        }
        if (drop.thr16) {
            const float4 d = drop_scale4(drop, g);
            v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
        }
        if (f32) ((float4*)f32)[g] = v;
        if (hi) {
            uint2 hh, ll;
            split_pair(v.x, v.y, hh.x, ll.x);
            split_pair(v.z, v.w, hh.y, ll.y);
            ((uint2*)hi)[g] = hh;
            if (lo) ((uint2*)lo)[g] = ll;
        }
    }
}

// ---------------------------------------------------------------- LayerNorm backward
// warp per row.  xhat = (x - mean) * rstd, g = dy * gamma:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres);  dgamma += dy * xhat;  dbeta += dy
// Per-block partial dgamma / dbeta go to partial[block, 3, 576]; a column sum over blocks finishes them.
// The gradient that leaves (dx) is what the next backward GEMMs consume, so the kernel also writes it in their operand
// format (op_*: fp32 or bf16 hi/lo), through the dropout mask of the Linear below when there is one, and accumulates
// its column sums (partial[block, 2, :]) = that Linear's bias gradient — no separate convert / column-sum passes.
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, const float* __restrict__ gamma,
              const float* dres, float* dx, int64_t rows, float* __restrict__ partial, DropSpec drop,
              float* __restrict__ op_f32, __nv_bfloat16* __restrict__ op_hi, __nv_bfloat16* __restrict__ op_lo) {
    constexpr int PER = kDim / 64;  // 9 float2 per lane
    __shared__ float red[8][kDim];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool want_op = op_f32 || op_hi;
    // gamma is re-read through L1 where it is used (2.3 KB, always resident) instead of living in 18 registers: the
    // registers hold the row's THREE input streams, so that all of a row's global loads are in flight together (one
    // DRAM round trip per row instead of two — the residual used to be fetched after the reductions).  dres may alias
    // dx (in-place accumulation into the residual gradient): a row's residual is fully loaded before its first store.
    const float2* gam = (const float2*)gamma;
    float2 dg[PER], db[PER], ds[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        dg[j] = make_float2(0.f, 0.f);
        db[j] = make_float2(0.f, 0.f);
        ds[j] = make_float2(0.f, 0.f);
    }
    for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < rows; row += (int64_t)gridDim.x * 8) {
        const float2* xr = (const float2*)(x + row * ldx);
        const float2* dyr = (const float2*)(dy + row * kDim);
        const float2* rr = (const float2*)(dres ? dres + row * kDim : nullptr);
        float2 v[PER], d[PER], r[PER];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = __ldcs(xr + lane + 32 * j);   // streamed: every input element is read exactly once
            d[j] = __ldcs(dyr + lane + 32 * j);
            r[j] = rr ? __ldcs(rr + lane + 32 * j) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) s += v[j].x + v[j].y;
        const float mean = warp_sum(s) * (1.f / kDim);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const float ax = v[j].x - mean, ay = v[j].y - mean;
            q += ax * ax + ay * ay;
        }
        const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kDim) + 1e-5f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const float2 gm = __ldg(gam + lane + 32 * j);
            v[j].x = (v[j].x - mean) * rstd;  // xhat
            v[j].y = (v[j].y - mean) * rstd;
            dg[j].x += d[j].x * v[j].x;
            dg[j].y += d[j].y * v[j].y;
            db[j].x += d[j].x;
            db[j].y += d[j].y;
            d[j].x *= gm.x;  // g
            d[j].y *= gm.y;
            s1 += d[j].x + d[j].y;
            s2 += d[j].x * v[j].x + d[j].y * v[j].y;
        }
        s1 = warp_sum(s1) * (1.f / kDim);
        s2 = warp_sum(s2) * (1.f / kDim);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            float2 o;
            o.x = rstd * (d[j].x - s1 - v[j].x * s2) + r[j].x;
            o.y = rstd * (d[j].y - s1 - v[j].y * s2) + r[j].y;
            const size_t e = (size_t)row * kDim + 2 * (lane + 32 * j);
            *(float2*)(dx + e) = o;
            if (want_op) {
                if (drop.thr16) {  // elements e, e + 1 share a 4-group (e is even)
                    const uint64_t h = drop_hash(drop.seed, e >> 2);
                    const int sh = 16 * (int)(e & 3);
                    o.x = ((uint32_t)(h >> sh) & 0xffffu) >= drop.thr16 ? o.x * drop.scale : 0.f;
                    o.y = ((uint32_t)(h >> (sh + 16)) & 0xffffu) >= drop.thr16 ? o.y * drop.scale : 0.f;
                }
                ds[j].x += o.x;
                ds[j].y += o.y;
                if (op_f32) *(float2*)(op_f32 + e) = o;
                if (op_hi) {
                    uint32_t hh, ll;
                    split_pair(o.x, o.y, hh, ll);
                    *(uint32_t*)(op_hi + e) = hh;
                    if (op_lo) *(uint32_t*)(op_lo + e) = ll;
                }
            }
        }
    }
    for (int v = 0; v < 3; ++v) {  // dgamma, dbeta, colsum(op): cross-warp sums through one 18 KB buffer
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const float2 t = v == 0 ? dg[j] : v == 1 ? db[j] : ds[j];
            *(float2*)&red[wid][2 * (lane + 32 * j)] = t;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < kDim; c += 256) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += red[w][c];
            partial[(size_t)blockIdx.x * 3 * kDim + v * kDim + c] = a;
        }
    }
}

// ---------------------------------------------------------------- attention backward
// TWO warps per (sequence, head): each owns 48 of the 96 head dims of Q, K, V, dO in shared memory, so everything
// except the 19 x 19 score tile is private to a warp.  Lane i owns query row i:
//   partial S = Q K^T and dP = dO V^T over the warp's dims  -> the odd warp hands its partials to the even one through
//   the P / dS tiles (named barrier), which finishes P = softmax(S * scale), dS = P * (dP - rowsum(P * dP)) * scale and
//   leaves both tiles in shared memory (second barrier);
//   dQ = dS K (row i);  then lane j owns key row j: dK = dS^T Q, dV = P^T dO, each for the warp's own dims.
// Results overwrite the staging buffers that are no longer needed (dQ -> V, dK -> K, dV -> Q) and leave with coalesced
// 128-bit stores.  The first version (one warp per item, six warps per SM, synchronous staging loop) ran at 1.18 ms
// per launch, latency-bound; 14 warps per SM with cp.async staging hide the shared-memory latency.
// Rows are 96 floats with the float4 column index XOR-swizzled by the row (c ^ (row & 7) within groups of 8), so the
// per-lane row reads are bank-conflict free without padding: 7 items fit in 227 KB.
// Column ownership: warp h owns the float4 columns c = 8g + 4h + i (g < 3, i < 4).  The swizzled position of (row, c)
// is 8g + 4 (h ^ row bit 2) + (i ^ (row & 3)), so in the inner loops — which walk the 19 rows of a K / V / Q / dO tile
// with the row index unrolled — every shared-memory address is one of two per-warp base registers (h, 1 - h) plus an
// immediate.  That removed the integer work per access but not the time (888 us per launch either way): the kernel is
// bound by the shared-memory RETURN path — a lane-per-row dot product needs one LDS.128 per four FMAs, 2340 LDS.128 per
// (sequence, head) at 4 clk each.  It remains the fp32-mode kernel; the tensor-core modes use attention_bwd_mma_kernel.
constexpr int AB_ROW = kHeadDim;                 // floats per staged row
constexpr int AB_PS = 20;                        // row stride of the P / dS tiles
constexpr int AB_ITEM = 4 * kTokens * AB_ROW + 2 * kTokens * AB_PS;  // floats per item (32,224 B)
constexpr int AB_ITEMS = 7;
constexpr int AB_THREADS = AB_ITEMS * 64;
constexpr int AB_SMEM = AB_ITEMS * AB_ITEM * (int)sizeof(float);  // 225,568 B
constexpr int AB_BUF = kTokens * AB_ROW;         // floats per staged tile

__device__ __forceinline__ int ab_col(int row, int c) { return (c & ~7) | ((c ^ row) & 7); }  // swizzled float4 column
__device__ __forceinline__ void ab_bar(int slot) { asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory"); }
// float offset of (row, column i of the warp's group) relative to the per-group base `pa` (row bit 2 clear) / `pb` (set)
#define AB_AT(pa, pb, row, i) ((((row) >> 2) & 1 ? (pb) : (pa)) + (row) * AB_ROW + 4 * ((i) ^ ((row) & 3)))

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out, int64_t n_seq, float* g_f32,
                     __nv_bfloat16* g_hi, __nv_bfloat16* g_lo) {
    extern __shared__ float4 ab_smem[];
    constexpr int LD = 3 * kDim, HV4 = kHeadDim / 8;  // 12 float4 columns per row per warp
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int slot = wid >> 1, half = wid & 1;
    float* sQ = reinterpret_cast<float*>(ab_smem) + slot * AB_ITEM;
    float* sK = sQ + AB_BUF;
    float* sV = sK + AB_BUF;
    float* sDO = sV + AB_BUF;
    float* sP = sDO + AB_BUF;
    float* sDS = sP + kTokens * AB_PS;
    auto cell = [&](float* buf, int row, int c) { return (float4*)(buf + row * AB_ROW) + ab_col(row, c); };
    auto own_col = [&](int t) { return 8 * (t >> 2) + 4 * half + (t & 3); };  // t-th owned column
    // this lane's own row (query row i / key row j = lane): float offset of group 0, and the XOR of the in-group index
    const int lrow = lane * AB_ROW + 16 * (half ^ ((lane >> 2) & 1));
    const int lx = lane & 3;
    const float scale = 0.10206207261596575f;
    const int64_t items = n_seq * kHeads;
    // every warp of a slot runs the same number of iterations (the barriers are per slot)
    for (int64_t item = (int64_t)blockIdx.x * AB_ITEMS + slot; item < items; item += (int64_t)gridDim.x * AB_ITEMS) {
        const int64_t seq = item / kHeads;
        const int h = (int)(item - seq * kHeads);
        const float* base = qkv + (size_t)seq * kTokens * LD + h * kHeadDim;
        const float* dob = d_out + (size_t)seq * kTokens * kDim + h * kHeadDim;
        __syncwarp();
        // stage this warp's 12 columns of the 76 rows (q, k, v, dO): 912 16-byte cp.async, all in flight at once
        for (int idx = lane; idx < 4 * kTokens * HV4; idx += 32) {
            const int m = idx / HV4, c = own_col(idx - m * HV4);
            const int which = m / kTokens, row = m - which * kTokens;
            const float* src = which < 3 ? base + (size_t)row * LD + which * kDim + 4 * c : dob + (size_t)row * kDim + 4 * c;
            float4* dst = cell(sQ + which * AB_BUF, row, c);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        float ds[kTokens];
        {
            float p[kTokens], dp[kTokens];
#pragma unroll
            for (int j = 0; j < kTokens; ++j) { p[j] = 0.f; dp[j] = 0.f; }
            if (lane < kTokens) {
#pragma unroll 1
                for (int g = 0; g < 3; ++g) {
                    const float* ka = sK + 32 * g + 16 * half;          // K tile; the V tile sits AB_BUF floats behind it
                    const float* kb = sK + 32 * g + 16 * (half ^ 1);
                    const float* ql = sQ + lrow + 32 * g;               // own row of Q; dO sits 3 * AB_BUF behind it
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 q = *(const float4*)(ql + 4 * (i ^ lx));
                        const float4 go = *(const float4*)(ql + 3 * AB_BUF + 4 * (i ^ lx));
#pragma unroll
                        for (int j = 0; j < kTokens; ++j) {
                            const float4 k4 = *(const float4*)AB_AT(ka, kb, j, i);
                            const float4 v4 = *(const float4*)(AB_AT(ka, kb, j, i) + AB_BUF);
                            p[j] = fmaf(q.x, k4.x, p[j]); p[j] = fmaf(q.y, k4.y, p[j]);
                            p[j] = fmaf(q.z, k4.z, p[j]); p[j] = fmaf(q.w, k4.w, p[j]);
                            dp[j] = fmaf(go.x, v4.x, dp[j]); dp[j] = fmaf(go.y, v4.y, dp[j]);
                            dp[j] = fmaf(go.z, v4.z, dp[j]); dp[j] = fmaf(go.w, v4.w, dp[j]);
                        }
                    }
                }
            }
            if (half == 1 && lane < kTokens) {
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    sP[lane * AB_PS + j] = p[j];
                    sDS[lane * AB_PS + j] = dp[j];
                }
            }
            ab_bar(slot);  // the odd warp's partial scores are in the tiles
            if (half == 0 && lane < kTokens) {
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    p[j] = (p[j] + sP[lane * AB_PS + j]) * scale;
                    dp[j] += sDS[lane * AB_PS + j];
                    m = fmaxf(m, p[j]);
                }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    p[j] = expf(p[j] - m);
                    sum += p[j];
                }
                const float inv = 1.f / sum;
                float dsum = 0.f;
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    p[j] *= inv;
                    dsum = fmaf(p[j], dp[j], dsum);
                }
#pragma unroll
                for (int j = 0; j < kTokens; ++j) {
                    sP[lane * AB_PS + j] = p[j];
                    sDS[lane * AB_PS + j] = p[j] * (dp[j] - dsum) * scale;
                }
            }
            ab_bar(slot);  // P and dS are final
        }
        if (lane < kTokens) {
#pragma unroll
            for (int j = 0; j < kTokens; ++j) ds[j] = sDS[lane * AB_PS + j];
            // dQ (row = lane) over the warp's columns -> V buffer (V is dead: every lane of this warp is past dP,
            // and the other warp never touches these columns)
#pragma unroll 1
            for (int g = 0; g < 3; ++g) {
                const float* ka = sK + 32 * g + 16 * half;
                const float* kb = sK + 32 * g + 16 * (half ^ 1);
                float* ol = sV + lrow + 32 * g;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < kTokens; ++j) {
                        const float4 k4 = *(const float4*)AB_AT(ka, kb, j, i);
                        a.x = fmaf(ds[j], k4.x, a.x); a.y = fmaf(ds[j], k4.y, a.y);
                        a.z = fmaf(ds[j], k4.z, a.z); a.w = fmaf(ds[j], k4.w, a.w);
                    }
                    *(float4*)(ol + 4 * (i ^ lx)) = a;
                }
            }
        }
        __syncwarp();  // this warp's K columns are free now: dK overwrites them; lane j = key row j, column j of dS / P
        if (lane < kTokens) {
#pragma unroll
            for (int i = 0; i < kTokens; ++i) ds[i] = sDS[i * AB_PS + lane];
#pragma unroll 1
            for (int g = 0; g < 3; ++g) {
                const float* qa = sQ + 32 * g + 16 * half;
                const float* qb = sQ + 32 * g + 16 * (half ^ 1);
                float* ol = sK + lrow + 32 * g;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < kTokens; ++i) {
                        const float4 q4 = *(const float4*)AB_AT(qa, qb, i, c);
                        a.x = fmaf(ds[i], q4.x, a.x); a.y = fmaf(ds[i], q4.y, a.y);
                        a.z = fmaf(ds[i], q4.z, a.z); a.w = fmaf(ds[i], q4.w, a.w);
                    }
                    *(float4*)(ol + 4 * (c ^ lx)) = a;
                }
            }
        }
        __syncwarp();  // Q columns are free now: dV overwrites them
        if (lane < kTokens) {
#pragma unroll
            for (int i = 0; i < kTokens; ++i) ds[i] = sP[i * AB_PS + lane];
#pragma unroll 1
            for (int g = 0; g < 3; ++g) {
                const float* ga = sDO + 32 * g + 16 * half;
                const float* gb = sDO + 32 * g + 16 * (half ^ 1);
                float* ol = sQ + lrow + 32 * g;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < kTokens; ++i) {
                        const float4 g4 = *(const float4*)AB_AT(ga, gb, i, c);
                        a.x = fmaf(ds[i], g4.x, a.x); a.y = fmaf(ds[i], g4.y, a.y);
                        a.z = fmaf(ds[i], g4.z, a.z); a.w = fmaf(ds[i], g4.w, a.w);
                    }
                    *(float4*)(ol + 4 * (c ^ lx)) = a;
                }
            }
        }
        __syncwarp();
        // dq (in sV) -> cols [0,576), dk (sK) -> [576,1152), dv (sQ) -> [1152,1728): this warp's 48 dims of the head
        for (int idx = lane; idx < 3 * kTokens * HV4; idx += 32) {
            const int m = idx / HV4, c = own_col(idx - m * HV4);
            const int which = m / kTokens, row = m - which * kTokens;
            const float4 v = *cell(which == 0 ? sV : which == 1 ? sK : sQ, row, c);
            const size_t o = ((size_t)seq * kTokens + row) * LD + which * kDim + h * kHeadDim + 4 * c;
            if (g_f32) *(float4*)(g_f32 + o) = v;
            if (g_hi) {
                uint2 hh, ll;
                split_pair(v.x, v.y, hh.x, ll.x);
                split_pair(v.z, v.w, hh.y, ll.y);
                *(uint2*)(g_hi + o) = hh;
                if (g_lo) *(uint2*)(g_lo + o) = ll;
            }
        }
        // the P / dS tiles are rewritten by the odd warp in the next iteration: both warps must be past their reads
        ab_bar(slot);
    }
}
#undef AB_AT

// ---------------------------------------------------------------- attention backward on the tensor cores (mma.sync)
// The SIMT kernel above is bound by the shared-memory return path: a lane-per-row dot product needs one LDS.128 per
// four FMAs (2340 LDS.128 per (sequence, head) x 4 clk = the 888 us ncu measures).  Warp MMAs read every operand once
// per 16x8x16 tile instead.  Same bf16 hi/lo split as the forward kernel (attention_mma_kernel, encoder_ops.cu) and
// the GEMMs: each product = hi*hi + lo*hi + hi*lo, fp32 accumulate.
//   S  = Q K^T, dP = dO V^T       A = rows of Q / dO, B = rows of K / V        (64-bit fragment reads, 6 k-steps)
//   P  = softmax(S * scale), dS = P o (dP - rowsum(P o dP)) * scale            on the accumulator registers (quad shuffles)
//   dQ = dS K                      A = dS straight from the accumulators (flash-attention style), B = K by key pairs
//   dK = dS^T Q, dV = P^T dO       A = the transposed tiles, through a 19x20 fp32 tile in shared memory
// TWO warps per (sequence, head) share one staged copy of Q, K, V, dO: warp w takes k-steps 3w..3w+2 of S / dP (the
// partial accumulators meet through two small tiles) and half of the 12 output n-tiles of dQ, dK, dV — 12 warps per
// SM instead of 6 for the same shared memory (the first, one-warp version: 696 us, latency-bound at 1.5 warps per
// scheduler).  Independent accumulators are interleaved between the three split terms so that dependent MMAs are
// 6-8 instructions apart.  Row stride 104 floats makes the 64-bit row reads conflict-free; the key-pair (column)
// reads are 2-way conflicted.
constexpr int ABM_STRIDE = 104;
constexpr int ABM_BUF = kTokens * ABM_STRIDE;
constexpr int ABM_TILE = kTokens * AB_PS;
constexpr int ABM_ITEM = 4 * ABM_BUF + 4 * ABM_TILE;  // floats per warp pair (37,696 B)
constexpr int ABM_PAIRS = 6;
constexpr int ABM_THREADS = ABM_PAIRS * 64;
constexpr int ABM_SMEM = ABM_PAIRS * ABM_ITEM * (int)sizeof(float);  // 226,176 B

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool SPLIT>
__global__ void __launch_bounds__(ABM_THREADS, 1)
attention_bwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out, int64_t n_seq,
                         __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo) {
    extern __shared__ float4 ab_smem[];
    constexpr int LD = 3 * kDim;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = threadIdx.x >> 6, w = (threadIdx.x >> 5) & 1, lane64 = threadIdx.x & 63;
    float* sQ = reinterpret_cast<float*>(ab_smem) + pair * ABM_ITEM;
    float* sK = sQ + ABM_BUF;
    float* sV = sK + ABM_BUF;
    float* sDO = sV + ABM_BUF;
    float* tiles = sDO + ABM_BUF;                 // [warp][2][19][20]: partial S / dP of each warp
    float* myS = tiles + w * 2 * ABM_TILE;
    float* myDP = myS + ABM_TILE;
    const float* otherS = tiles + (w ^ 1) * 2 * ABM_TILE;
    const float* otherDP = otherS + ABM_TILE;
    float* sP = tiles;                            // final P / dS: warp 0's tiles, rewritten after both warps read them
    float* sDS = tiles + ABM_TILE;
    const float scale = 0.10206207261596575f;
    const int64_t items = n_seq * kHeads;
    auto pair_bar = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };
    // A / B fragments of "X rows" (row-major [19][96], k = head dim): 64-bit reads, rows past 18 are zero
    auto row2 = [&](const float* buf, int row, int col) -> float2 {
        return row < kTokens ? *(const float2*)(buf + row * ABM_STRIDE + col) : make_float2(0.f, 0.f);
    };
    // B fragments with k = token index (pairs of rows of one column), tokens past 18 are zero
    auto col1 = [&](const float* buf, int row, int col) -> float { return row < kTokens ? buf[row * ABM_STRIDE + col] : 0.f; };
    // transposed 19x19 tile element [k][m] (k = contraction index = tile row), zero outside
    auto tile = [&](const float* tl, int k, int m) -> float { return (k < kTokens && m < kTokens) ? tl[k * AB_PS + m] : 0.f; };

    // both warps of a pair run the same number of iterations (the barriers are per pair)
    for (int64_t item = (int64_t)blockIdx.x * ABM_PAIRS + pair; item < items; item += (int64_t)gridDim.x * ABM_PAIRS) {
        const int64_t seq = item / kHeads;
        const int h = (int)(item - seq * kHeads);
        const float* base = qkv + (size_t)seq * kTokens * LD + h * kHeadDim;
        const float* dob = d_out + (size_t)seq * kTokens * kDim + h * kHeadDim;
        for (int idx = lane64; idx < 4 * kTokens * (kHeadDim / 4); idx += 64) {
            const int m = idx / (kHeadDim / 4), c = idx - m * (kHeadDim / 4);
            const int which = m / kTokens, row = m - which * kTokens;
            const float* src = which < 3 ? base + (size_t)row * LD + which * kDim + 4 * c : dob + (size_t)row * kDim + 4 * c;
            float* dst = sQ + which * ABM_BUF + row * ABM_STRIDE + 4 * c;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        pair_bar();  // (1) the staged operands of both warps are visible

        // ---- partial S = Q K^T and dP = dO V^T over this warp's 3 k-steps of 16 head dims:
        // 2 m-tiles (rows 16mt+g, +8) x 3 n-tiles (keys 8nt+g)
        float S[2][3][4], DP[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) { S[mt][nt][e] = 0.f; DP[mt][nt][e] = 0.f; }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {          // pass 0: (Q, K) -> S; pass 1: (dO, V) -> dP
            const float* A = pass == 0 ? sQ : sDO;
            const float* B = pass == 0 ? sK : sV;
#pragma unroll 1
            for (int ks = 3 * w; ks < 3 * w + 3; ++ks) {
                const int d0 = ks * 16 + 2 * t;
                uint32_t ah[2][4], al[2][4], bh[3][2], bl[3][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float2 x0 = row2(A, 16 * mt + g, d0), x1 = row2(A, 16 * mt + g + 8, d0);
                    const float2 x2 = row2(A, 16 * mt + g, d0 + 8), x3 = row2(A, 16 * mt + g + 8, d0 + 8);
                    split_pair(x0.x, x0.y, ah[mt][0], al[mt][0]);
                    split_pair(x1.x, x1.y, ah[mt][1], al[mt][1]);
                    split_pair(x2.x, x2.y, ah[mt][2], al[mt][2]);
                    split_pair(x3.x, x3.y, ah[mt][3], al[mt][3]);
                }
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    const float2 y0 = row2(B, 8 * nt + g, d0), y1 = row2(B, 8 * nt + g, d0 + 8);
                    split_pair(y0.x, y0.y, bh[nt][0], bl[nt][0]);
                    split_pair(y1.x, y1.y, bh[nt][1], bl[nt][1]);
                }
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term)
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            if (pass == 0) mma_16816(S[mt][nt], term == 1 ? al[mt] : ah[mt], term == 2 ? bl[nt] : bh[nt]);
                            else mma_16816(DP[mt][nt], term == 1 ? al[mt] : ah[mt], term == 2 ? bl[nt] : bh[nt]);
                        }
            }
        }
        // ---- the two warps exchange their partial accumulators (e 0,1 -> row 16mt+g, e 2,3 -> row +8; cols 8nt+2t+e)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
                    if (row < kTokens && col < kTokens) {
                        myS[row * AB_PS + col] = S[mt][nt][e];
                        myDP[row * AB_PS + col] = DP[mt][nt][e];
                    }
                }
        pair_bar();  // (2) partials written
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
                    if (row < kTokens && col < kTokens) {
                        S[mt][nt][e] += otherS[row * AB_PS + col];
                        DP[mt][nt][e] += otherDP[row * AB_PS + col];
                    }
                }
        pair_bar();  // (3) partials consumed: warp 0's tiles may be overwritten with the final P / dS

        // ---- P = softmax(S * scale), dS = P o (dP - rowsum(P o dP)) * scale; a row = the 4 lanes of a quad.
        // S becomes P, DP becomes dS (both warps hold the full tiles; warp 0 publishes them).
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
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
                float dsum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float p = S[mt][nt][2 * hrow + e] * inv;
                        S[mt][nt][2 * hrow + e] = p;
                        dsum = fmaf(p, DP[mt][nt][2 * hrow + e], dsum);   // p = 0 on the padding keys
                    }
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                const int row = 16 * mt + g + 8 * hrow;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float p = S[mt][nt][2 * hrow + e];
                        const float dsv = p * (DP[mt][nt][2 * hrow + e] - dsum) * scale;
                        DP[mt][nt][2 * hrow + e] = dsv;
                        const int col = 8 * nt + 2 * t + e;
                        if (w == 0 && row < kTokens && col < kTokens) {
                            sP[row * AB_PS + col] = p;
                            sDS[row * AB_PS + col] = dsv;
                        }
                    }
            }
        pair_bar();  // (4) final P / dS tiles visible

        // output rows of this item: (seq * 19 + row) * 1728 + which * 576 + h * 96 + d
        auto store2 = [&](int which, int row, int d, float v0, float v1) {
            if (row < kTokens) {
                const size_t o = ((size_t)seq * kTokens + row) * LD + which * kDim + h * kHeadDim + d;
                uint32_t hh, ll;
                split_pair(v0, v1, hh, ll);
                *(uint32_t*)(g_hi + o) = hh;
                if (g_lo) *(uint32_t*)(g_lo + o) = ll;
            }
        };
        // C[19 x 96] = A[19 x 19] B[19 x 96] with A fragments in registers ([mt][ks2][4], hi / lo), B = rows of `buf`
        // taken by token pairs.  12 n-tiles of 8 head dims in 4 blocks of 3; this warp takes blocks 2w, 2w+1.
        auto product = [&](const uint32_t (&ah)[2][2][4], const uint32_t (&al)[2][2][4], const float* buf, int which) {
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
                    uint32_t bh[3][2], bl[3][2];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int d = 8 * (3 * blk + j) + g;
                        split_pair(col1(buf, k0, d), col1(buf, k0 + 1, d), bh[j][0], bl[j][0]);
                        split_pair(col1(buf, k0 + 8, d), col1(buf, k0 + 9, d), bh[j][1], bl[j][1]);
                    }
#pragma unroll
                    for (int term = 0; term < (SPLIT ? 3 : 1); ++term)
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt)
                                mma_16816(O[mt][j], term == 1 ? al[mt][ks2] : ah[mt][ks2], term == 2 ? bl[j] : bh[j]);
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int d = 8 * (3 * blk + j) + 2 * t;
                        store2(which, 16 * mt + g, d, O[mt][j][0], O[mt][j][1]);
                        store2(which, 16 * mt + g + 8, d, O[mt][j][2], O[mt][j][3]);
                    }
            }
        };

        uint32_t fh[2][2][4], fl[2][2][4];
        // ---- dQ = dS K : dS as A fragments straight from the accumulators; k-step ks2 covers keys 16ks2 .. 16ks2+15 =
        // n-tiles 2ks2, 2ks2+1 (n-tile 3 does not exist: keys 24..31 are padding)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            split_pair(DP[mt][0][0], DP[mt][0][1], fh[mt][0][0], fl[mt][0][0]);
            split_pair(DP[mt][0][2], DP[mt][0][3], fh[mt][0][1], fl[mt][0][1]);
            split_pair(DP[mt][1][0], DP[mt][1][1], fh[mt][0][2], fl[mt][0][2]);
            split_pair(DP[mt][1][2], DP[mt][1][3], fh[mt][0][3], fl[mt][0][3]);
            split_pair(DP[mt][2][0], DP[mt][2][1], fh[mt][1][0], fl[mt][1][0]);
            split_pair(DP[mt][2][2], DP[mt][2][3], fh[mt][1][1], fl[mt][1][1]);
            fh[mt][1][2] = fl[mt][1][2] = fh[mt][1][3] = fl[mt][1][3] = 0u;
        }
        product(fh, fl, sK, 0);
        // ---- dK = dS^T Q and dV = P^T dO : A[m = key][k = query] = tile[query][key]
        auto transposed = [&](const float* tl) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int ks2 = 0; ks2 < 2; ++ks2) {
                    const int k0 = 16 * ks2 + 2 * t, m0 = 16 * mt + g;
                    split_pair(tile(tl, k0, m0), tile(tl, k0 + 1, m0), fh[mt][ks2][0], fl[mt][ks2][0]);
                    split_pair(tile(tl, k0, m0 + 8), tile(tl, k0 + 1, m0 + 8), fh[mt][ks2][1], fl[mt][ks2][1]);
                    split_pair(tile(tl, k0 + 8, m0), tile(tl, k0 + 9, m0), fh[mt][ks2][2], fl[mt][ks2][2]);
                    split_pair(tile(tl, k0 + 8, m0 + 8), tile(tl, k0 + 9, m0 + 8), fh[mt][ks2][3], fl[mt][ks2][3]);
                }
        };
        transposed(sDS);
        product(fh, fl, sQ, 1);
        transposed(sP);
        product(fh, fl, sDO, 2);
        pair_bar();  // (5) both warps are done with the staged operands and the tiles
    }
}

// ---------------------------------------------------------------- relation-token backward
// grid (N boxes, 18 slots): slot p < 16 = patch token 1+p, 16 = location token, 17 = class token.  The block scans the
// pairs of the box's image in order and adds the token-gradient rows of the pairs the box is subject / object of
// (thread t owns float4 column t, like tokens_kernel): a deterministic gather instead of an atomic scatter.
__device__ __forceinline__ void acc4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

__global__ void __launch_bounds__(kDim / 4)
tokens_bwd_kernel(const float* __restrict__ dx, const int32_t* __restrict__ subj, const int32_t* __restrict__ obj,
                  const int32_t* __restrict__ rel_offsets, const int32_t* __restrict__ box_offsets, int n_images,
                  const float* __restrict__ lso, const float* __restrict__ cso, float* __restrict__ d_so_d,
                  float* __restrict__ d_so_v, float* __restrict__ d_lso, float* __restrict__ d_cso) {
    __shared__ int s_img;
    const int b = blockIdx.x, slot = blockIdx.y, t = threadIdx.x;
    if (t == 0) {
        int lo = 0, hi = n_images - 1;  // last image whose first box is <= b
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (box_offsets[mid] <= b) lo = mid; else hi = mid - 1;
        }
        // skip empty images that share the offset
        while (lo + 1 < n_images && box_offsets[lo + 1] <= b) ++lo;
        s_img = lo;
    }
    __syncthreads();
    const int r0 = rel_offsets[s_img], r1 = rel_offsets[s_img + 1];
    const int tok = slot < kPatches ? 1 + slot : slot + 1;  // 16 -> 17, 17 -> 18
    float4 as = make_float4(0.f, 0.f, 0.f, 0.f), ao = as;
    if (slot < kPatches) {
        for (int r = r0; r < r1; ++r) {
            const int s = subj[r], o = obj[r];
            if (s == b || o == b) {
                const float4 g = __ldg((const float4*)(dx + ((size_t)r * kTokens + tok) * kDim) + t);
                if (s == b) acc4(as, g);
                if (o == b) acc4(ao, g);
            }
        }
        if (t < kDimDepth / 4) {
            float4* dst = (float4*)(d_so_d + ((size_t)b * kPatches + slot) * 2 * kDimDepth);
            dst[t] = as;
            dst[kDimDepth / 4 + t] = ao;
        } else {
            const int tv = t - kDimDepth / 4;
            float4* dst = (float4*)(d_so_v + ((size_t)b * kPatches + slot) * 2 * kDimRgb);
            dst[tv] = as;
            dst[kDimRgb / 4 + tv] = ao;
        }
    } else {
        const float* pre = slot == kPatches ? lso : cso;
        float* dst = slot == kPatches ? d_lso : d_cso;
        const float4 own_s = __ldg((const float4*)(pre + (size_t)b * 2 * kDim) + t);
        const float4 own_o = __ldg((const float4*)(pre + (size_t)b * 2 * kDim + kDim) + t);
        for (int r = r0; r < r1; ++r) {
            const int s = subj[r], o = obj[r];
            if (s == b || o == b) {
                const float4 g = __ldg((const float4*)(dx + ((size_t)r * kTokens + tok) * kDim) + t);
                if (s == b) {  // ReLU(pre_s[b] + pre_o[o]) (tokens.cu)
                    const float4 other = __ldg((const float4*)(pre + (size_t)o * 2 * kDim + kDim) + t);
                    as.x += (own_s.x + other.x > 0.f) ? g.x : 0.f; as.y += (own_s.y + other.y > 0.f) ? g.y : 0.f;
                    as.z += (own_s.z + other.z > 0.f) ? g.z : 0.f; as.w += (own_s.w + other.w > 0.f) ? g.w : 0.f;
                }
                if (o == b) {
                    const float4 other = __ldg((const float4*)(pre + (size_t)s * 2 * kDim) + t);
                    ao.x += (other.x + own_o.x > 0.f) ? g.x : 0.f; ao.y += (other.y + own_o.y > 0.f) ? g.y : 0.f;
                    ao.z += (other.z + own_o.z > 0.f) ? g.z : 0.f; ao.w += (other.w + own_o.w > 0.f) ? g.w : 0.f;
                }
            }
        }
        ((float4*)(dst + (size_t)b * 2 * kDim))[t] = as;
        ((float4*)(dst + (size_t)b * 2 * kDim + kDim))[t] = ao;
    }
}

// ---------------------------------------------------------------- box stage, training mode
// BatchNorm1d(4) batch statistics over all boxes of the step (roi_relation_predictors.py:4042-4047 in train()):
// stats[0..3] = mean, stats[4..7] = biased variance; running stats updated with `momentum` (unbiased variance).
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ boxes, int n, float momentum, float* __restrict__ stats, float* running_mean,
                float* running_var) {
    __shared__ double red[256][4];
    auto feat = [&](int i, float (&f)[4]) {
        const float4 bx = __ldg((const float4*)boxes + i);
        const float w = bx.z - bx.x + 1.f, h = bx.w - bx.y + 1.f;
        f[0] = bx.x + 0.5f * w; f[1] = bx.y + 0.5f * h; f[2] = w; f[3] = h;
    };
    auto block_sum4 = [&](double (&a)[4]) {
        for (int k = 0; k < 4; ++k) red[threadIdx.x][k] = a[k];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o)
                for (int k = 0; k < 4; ++k) red[threadIdx.x][k] += red[threadIdx.x + o][k];
            __syncthreads();
        }
        for (int k = 0; k < 4; ++k) a[k] = red[0][k];
        __syncthreads();
    };
    double a[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += 256) {
        float f[4];
        feat(i, f);
        for (int k = 0; k < 4; ++k) a[k] += f[k];
    }
    block_sum4(a);
    double mean[4];
    for (int k = 0; k < 4; ++k) mean[k] = a[k] / n;
    double q[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += 256) {
        float f[4];
        feat(i, f);
        for (int k = 0; k < 4; ++k) q[k] += (f[k] - mean[k]) * (f[k] - mean[k]);
    }
    block_sum4(q);
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        stats[k] = (float)mean[k];
        stats[4 + k] = (float)(q[k] / n);
        if (running_mean) running_mean[k] = (1.f - momentum) * running_mean[k] + momentum * (float)mean[k];
        if (running_var) running_var[k] = (1.f - momentum) * running_var[k] + momentum * (float)(q[k] / (n > 1 ? n - 1 : 1));
    }
}

// pos = dropout(ReLU(Linear(4,128)(BN(f)))) backward, blocks of 128 threads (thread t = pos feature t), block b takes
// boxes b, b + gridDim.x, ...: partial[b] = { d pos_w [128,4] | d pos_b [128] | d bn_weight [4] | d bn_bias [4] } (648
// floats), summed over blocks by a column sum.  No gradient flows into the boxes.
__global__ void __launch_bounds__(128)
pos_embed_bwd_kernel(const float* __restrict__ boxes, int n, const float* __restrict__ stats, const float* __restrict__ bn_w,
                     const float* __restrict__ bn_b, const float* __restrict__ pos_w, const float* __restrict__ pos_out,
                     const float* __restrict__ d_pos, float drop_scale, float* __restrict__ partial) {
    float* g_pos_w = partial + (size_t)blockIdx.x * kPosBwdCols;
    float* g_pos_b = g_pos_w + 4 * kPosDim;
    float* g_bn_w = g_pos_b + kPosDim;
    float* g_bn_b = g_bn_w + 4;
    __shared__ float red[4][4];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    float w[4], gw[4] = {0.f, 0.f, 0.f, 0.f}, gb = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = pos_w[t * 4 + k];
    float mean[4], rstd[4], gam[4], bet[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mean[k] = stats[k];
        rstd[k] = 1.f / sqrtf(stats[4 + k] + 1e-5f);
        gam[k] = bn_w[k];
        bet[k] = bn_b[k];
    }
    float g_gamma = 0.f, g_beta = 0.f;  // thread k < 4 accumulates feature k
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const float4 bx = __ldg((const float4*)boxes + i);
        const float bw = bx.z - bx.x + 1.f, bh = bx.w - bx.y + 1.f;
        const float f[4] = {bx.x + 0.5f * bw, bx.y + 0.5f * bh, bw, bh};
        float xh[4], bn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xh[k] = (f[k] - mean[k]) * rstd[k];
            bn[k] = xh[k] * gam[k] + bet[k];
        }
        const float g = pos_out[(size_t)i * kPosDim + t] > 0.f ? d_pos[(size_t)i * kPosDim + t] * drop_scale : 0.f;
        gb += g;
        float part[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gw[k] = fmaf(g, bn[k], gw[k]);
            part[k] = warp_sum(g * w[k]);
        }
        __syncthreads();
        if (lane == 0)
            for (int k = 0; k < 4; ++k) red[wid][k] = part[k];
        __syncthreads();
        if (t < 4) {
            const float dbn = red[0][t] + red[1][t] + red[2][t] + red[3][t];
            g_gamma = fmaf(dbn, xh[t], g_gamma);
            g_beta += dbn;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) g_pos_w[t * 4 + k] = gw[k];
    g_pos_b[t] = gb;
    if (t < 4) {
        g_bn_w[t] = g_gamma;
        g_bn_b[t] = g_beta;
    }
}

// d obj_embed[c, :] = sum over boxes with label c of d_emb[n, :]   (block per class, fixed box order)
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ d_emb, const int64_t* __restrict__ labels, int n, float* __restrict__ g_embed) {
    const int c = blockIdx.x, d = threadIdx.x;
    if (d >= kEmbDim) return;
    float a = 0.f;
    for (int i = 0; i < n; ++i)
        if (labels[i] == c) a += d_emb[(size_t)i * kEmbDim + d];
    g_embed[(size_t)c * kEmbDim + d] = a;
}

// soft class embedding emb = softmax(z) @ E (roi_relation_predictors.py:4095): d E[c, :] = sum_n softmax(z_n)[c] d_emb[n, :]
// (obj_logits are detached in the reference, so no gradient flows into them).  Block per class.
__global__ void __launch_bounds__(256)
embed_soft_bwd_kernel(const float* __restrict__ d_emb, const float* __restrict__ obj_logits, int num_obj, int n,
                      float* __restrict__ g_embed) {
    __shared__ float s_p;
    const int c = blockIdx.x, d = threadIdx.x;
    float a = 0.f;
    for (int i = 0; i < n; ++i) {
        if (d < 32) {  // warp 0: softmax probability of class c for box i
            const float* z = obj_logits + (size_t)i * num_obj;
            float m = -INFINITY;
            for (int k = d; k < num_obj; k += 32) m = fmaxf(m, z[k]);
            m = warp_max(m);
            float s = 0.f;
            for (int k = d; k < num_obj; k += 32) s += expf(z[k] - m);
            s = warp_sum(s);
            if (d == 0) s_p = expf(z[c] - m) / s;
        }
        __syncthreads();
        if (d < kEmbDim) a = fmaf(s_p, d_emb[(size_t)i * kEmbDim + d], a);
        __syncthreads();
    }
    if (d < kEmbDim) g_embed[(size_t)c * kEmbDim + d] = a;
}

// inverse of patchify_kernel: d_patch rows [N*16, 1024] in (p1 p2 c) order -> d_roi [N,256,8,8]
__global__ void __launch_bounds__(256)
unpatchify_kernel(const float* __restrict__ d_patch, float* __restrict__ d_roi) {
    __shared__ float tile[32][65];
    const int n = blockIdx.x, c0 = blockIdx.y * 32, t = threadIdx.x;
    for (int e = t; e < 64 * 32; e += 256) {
        const int c = e & 31, q = e >> 5;
        const int patch = q >> 2, pp = q & 3;
        const int ph = patch >> 2, pw = patch & 3, p1 = pp >> 1, p2 = pp & 1;
        tile[c][(2 * ph + p1) * 8 + 2 * pw + p2] = d_patch[((size_t)n * kPatches + patch) * kPatchVec + pp * kChannels + c0 + c];
    }
    __syncthreads();
    float* dst = d_roi + ((size_t)n * kChannels + c0) * 64;
    for (int e = t; e < 32 * 64; e += 256) dst[e] = tile[e >> 6][e & 63];
}

// gradient of a factored weight back in the reference's layout (inverse index maps of pack_halves / pack_patch)
__global__ void unpack_halves_kernel(const float* __restrict__ g_packed, float* __restrict__ g_src, int out, int in) {
    const int total = 2 * out * in;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e % in, r = e / in;
        const int h = r / out, o = r - h * out;
        g_src[(size_t)o * 2 * in + h * in + i] = g_packed[e];
    }
}
__global__ void unpack_patch_kernel(const float* __restrict__ g_packed, float* __restrict__ g_src, int out) {
    const int total = 2 * out * kPatchVec;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int col = e % kPatchVec, r = e / kPatchVec;
        const int h = r / out, o = r - h * out;
        const int p = col / kChannels, c = col - p * kChannels;
        g_src[(size_t)o * 2 * kPatchVec + p * 2 * kChannels + h * kChannels + c] = g_packed[e];
    }
}

// dst[r * ld_dst + c] = src[r * ld_src + c] for a [rows, cols] block (strided row copy, fp32)
__global__ void __launch_bounds__(256)
copy_rows_kernel(const float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst, int64_t rows, int cols) {
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols;
        const int c = (int)(e - r * cols);
        dst[r * ld_dst + c] = src[r * ld_src + c];
    }
}

}  // namespace

// ================================================================ host wrappers
int ce_loss_grad(const float* logits, int ld, int col0, int C, const int64_t* labels, const float* weight, int64_t rows,
                 float* row_scratch, float* loss_out, float* dlogits, cudaStream_t s) {
    if (rows <= 0) return VETO_OK;
    float* row_loss = row_scratch;
    float* row_w = row_scratch + rows;
    float* inv_w = row_scratch + 2 * rows;
    const int grid = grid_cap((size_t)(rows + 7) / 8, 8);
    ce_rows_kernel<<<grid, 256, 0, s>>>(logits, ld, col0, C, labels, weight, rows, row_loss, row_w);
    VETO_LAUNCH_CHECK();
    ce_finalize_kernel<<<1, 1024, 0, s>>>(row_loss, row_w, rows, loss_out, inv_w);
    VETO_LAUNCH_CHECK();
    ce_grad_kernel<<<grid, 256, 0, s>>>(logits, ld, col0, C, labels, weight, inv_w, rows, dlogits);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

size_t colsum_scratch_floats(int cols) { return (size_t)kColsumMaxChunks * cols; }

int colsum(const ActIn& src, int64_t ld, int64_t rows, int cols, float* scratch, float* out, bool accumulate, cudaStream_t s) {
    if (cols <= 0) return VETO_OK;
    int chunks = (int)((rows + 127) / 128);
    if (chunks > kColsumMaxChunks) chunks = kColsumMaxChunks;
    if (chunks < 1) chunks = 1;
    const int per = (int)((rows + chunks - 1) / chunks);
    dim3 grid((cols + 127) / 128, chunks);
    colsum_partial_kernel<<<grid, 128, 0, s>>>(src.f32, src.hi, src.lo, ld, rows, cols, per > 0 ? per : 1, scratch);
    VETO_LAUNCH_CHECK();
    colsum_final_kernel<<<(cols + 127) / 128, 128, 0, s>>>(scratch, chunks, cols, out, accumulate ? 1 : 0);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int transpose_f32(const float* src, int64_t ld_src, int64_t rows, int cols, bool gelu, const DropSpec& drop, int64_t drop_ld,
                  const ActOut& t_out, int64_t ld_dst, int64_t rows_pad, const ActOut& rm_out, int64_t ld_rm, cudaStream_t s) {
    if (cols <= 0 || rows_pad <= 0) return VETO_OK;
    dim3 grid((cols + 31) / 32, (unsigned)((rows_pad + 31) / 32));
    VETO_REQUIRE(grid.y <= 65535, VETO_ERR_UNSUPPORTED, "transpose: %lld rows exceed one launch", (long long)rows_pad);
    transpose_f32_kernel<<<grid, 256, 0, s>>>(src, ld_src, rows, cols, gelu ? TR_OP_GELU : TR_OP_NONE, drop, drop_ld, t_out.f32,
                                              t_out.hi, t_out.lo, ld_dst, rows_pad, rm_out.f32, rm_out.hi, rm_out.lo, ld_rm);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int transpose_f32_multi(const TransposeJob* jobs, int count, cudaStream_t s) {
    if (count <= 0) return VETO_OK;
    VETO_REQUIRE(count <= kMaxTransposeJobs, VETO_ERR_ARG, "transpose_f32_multi: %d matrices > %d", count, kMaxTransposeJobs);
    TransposeJobs J{};
    int rows = 0, cols = 0;
    for (int i = 0; i < count; ++i) {
        J.j[i] = jobs[i];
        rows = jobs[i].rows > rows ? jobs[i].rows : rows;
        cols = jobs[i].cols > cols ? jobs[i].cols : cols;
    }
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, count);
    transpose_f32_multi_kernel<<<grid, 256, 0, s>>>(J);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int transpose_bf16(const __nv_bfloat16* src, int64_t ld_src, int64_t rows, int cols, __nv_bfloat16* dst, int64_t ld_dst,
                   int64_t rows_pad, cudaStream_t s) {
    if (cols <= 0 || rows_pad <= 0) return VETO_OK;
    VETO_REQUIRE((ld_dst & 1) == 0 && (rows_pad & 1) == 0 && (ld_src & 1) == 0, VETO_ERR_ARG, "transpose_bf16: odd strides");
    dim3 grid((cols + 63) / 64, (unsigned)((rows_pad + 63) / 64));
    VETO_REQUIRE(grid.y <= 65535, VETO_ERR_UNSUPPORTED, "transpose: %lld rows exceed one launch", (long long)rows_pad);
    transpose_bf16_kernel<<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst, rows_pad);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int splitk_reduce(const float* partial, int slices, size_t n, size_t stride, float* out, cudaStream_t s) {
    splitk_reduce_kernel<<<grid_cap((n + 255) / 256, 8), 256, 0, s>>>(partial, slices, n, stride, out);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int dropout_inplace(float* x, size_t n, const DropSpec& drop, cudaStream_t s) {
    if (!drop.thr16 || n == 0) return VETO_OK;
    VETO_REQUIRE(n % 4 == 0, VETO_ERR_ARG, "dropout: element count must be a multiple of 4");
    dropout_kernel<<<grid_cap((n / 4 + 255) / 256, 8), 256, 0, s>>>(x, n / 4, drop);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int convert_act(const float* src, size_t n, const DropSpec& drop, const ActOut& out, cudaStream_t s, int act) {
    if (n == 0) return VETO_OK;
    VETO_REQUIRE(n % 4 == 0, VETO_ERR_ARG, "convert_act: element count must be a multiple of 4");
    convert_kernel<<<grid_cap((n / 4 + 255) / 256, 8), 256, 0, s>>>(src, n / 4, drop, out.f32, out.hi, out.lo, act);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int ln_bwd_blocks(int64_t rows) { return grid_cap((size_t)(rows + 7) / 8, 4); }

int layernorm_bwd(const float* x, int64_t ldx, const float* dy, const float* gamma, const float* dres, float* dx, int64_t rows,
                  float* partial, float* colsum_scratch, float* g_gamma, float* g_beta, cudaStream_t s, const DropSpec& drop,
                  const ActOut& op_out, float* g_op_colsum) {
    if (rows <= 0) return VETO_OK;
    const int grid = ln_bwd_blocks(rows);
    ln_bwd_kernel<<<grid, 256, 0, s>>>(x, ldx, dy, gamma, dres, dx, rows, partial, drop, op_out.f32, op_out.hi, op_out.lo);
    VETO_LAUNCH_CHECK();
    // one two-stage column sum over the per-block partials finishes dgamma | dbeta | colsum(op) together
    const bool want_op = (op_out.f32 || op_out.hi) && g_op_colsum;
    const int cols = want_op ? 3 * kDim : 2 * kDim;
    ActIn p;
    p.f32 = partial;
    float* packed = colsum_scratch + colsum_scratch_floats(3 * kDim);
    int rc = colsum(p, 3 * kDim, grid, cols, colsum_scratch, packed, false, s);
    if (rc) return rc;
    VETO_CUDA(cudaMemcpyAsync(g_gamma, packed, sizeof(float) * kDim, cudaMemcpyDeviceToDevice, s));
    VETO_CUDA(cudaMemcpyAsync(g_beta, packed + kDim, sizeof(float) * kDim, cudaMemcpyDeviceToDevice, s));
    if (want_op) VETO_CUDA(cudaMemcpyAsync(g_op_colsum, packed + 2 * kDim, sizeof(float) * kDim, cudaMemcpyDeviceToDevice, s));
    return VETO_OK;
}

int attention_bwd(const float* qkv, const float* d_out, int64_t n_seq, const ActOut& d_qkv, cudaStream_t s) {
    if (n_seq <= 0) return VETO_OK;
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        VETO_CUDA(cudaFuncSetAttribute(attention_bwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ABM_SMEM));
        VETO_CUDA(cudaFuncSetAttribute(attention_bwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ABM_SMEM));
        attr_set.done();
    }
    static const bool simt_only = getenv("VETO_ATTN_BWD_SIMT") != nullptr;  // diagnosis: the fp32 SIMT kernel in every mode
    if (d_qkv.hi && !d_qkv.f32 && !simt_only) {  // tensor-core modes: bf16 hi (+ lo) operands for the qkv weight / input gradients
        const int64_t blocks = (n_seq * kHeads + ABM_PAIRS - 1) / ABM_PAIRS;
        const int grid = (int)(blocks < num_sms() ? blocks : num_sms());
        if (d_qkv.lo)
            attention_bwd_mma_kernel<true><<<grid, ABM_THREADS, ABM_SMEM, s>>>(qkv, d_out, n_seq, d_qkv.hi, d_qkv.lo);
        else
            attention_bwd_mma_kernel<false><<<grid, ABM_THREADS, ABM_SMEM, s>>>(qkv, d_out, n_seq, d_qkv.hi, nullptr);
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
    const int64_t blocks = (n_seq * kHeads + AB_ITEMS - 1) / AB_ITEMS;
    const int grid = (int)(blocks < num_sms() ? blocks : num_sms());
    attention_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, s>>>(qkv, d_out, n_seq, d_qkv.f32, d_qkv.hi, d_qkv.lo);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int tokens_bwd(const float* dx, const int32_t* subj, const int32_t* obj, const int32_t* rel_offsets, const int32_t* box_offsets,
               int n_images, int n_boxes, const float* lso, const float* cso, float* d_so_d, float* d_so_v, float* d_lso,
               float* d_cso, cudaStream_t s) {
    if (n_boxes <= 0) return VETO_OK;
    dim3 grid(n_boxes, kPatches + 2);
    tokens_bwd_kernel<<<grid, kDim / 4, 0, s>>>(dx, subj, obj, rel_offsets, box_offsets, n_images, lso, cso, d_so_d, d_so_v,
                                                d_lso, d_cso);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int bn_batch_stats(const float* boxes, int n_boxes, float momentum, float* stats, float* running_mean, float* running_var,
                   cudaStream_t s) {
    bn_stats_kernel<<<1, 256, 0, s>>>(boxes, n_boxes, momentum, stats, running_mean, running_var);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

size_t pos_embed_bwd_scratch_floats() { return (size_t)(kPosBwdBlocks + 1) * kPosBwdCols; }

int pos_embed_bwd(const float* boxes, int n_boxes, const float* stats, const veto_weights& w, const float* pos_out,
                  const float* d_pos, float drop_scale, float* scratch, float* colsum_scratch, float* g_pos_w, float* g_pos_b,
                  float* g_bn_w, float* g_bn_b, cudaStream_t s) {
    const int grid = n_boxes < kPosBwdBlocks ? (n_boxes > 0 ? n_boxes : 1) : kPosBwdBlocks;
    pos_embed_bwd_kernel<<<grid, 128, 0, s>>>(boxes, n_boxes, stats, w.bn_weight, w.bn_bias, w.pos_w, pos_out, d_pos, drop_scale,
                                              scratch);
    VETO_LAUNCH_CHECK();
    float* packed = scratch + (size_t)kPosBwdBlocks * kPosBwdCols;
    ActIn p;
    p.f32 = scratch;
    int rc = colsum(p, kPosBwdCols, grid, kPosBwdCols, colsum_scratch, packed, false, s);
    if (rc) return rc;
    VETO_CUDA(cudaMemcpyAsync(g_pos_w, packed, sizeof(float) * 4 * kPosDim, cudaMemcpyDeviceToDevice, s));
    VETO_CUDA(cudaMemcpyAsync(g_pos_b, packed + 4 * kPosDim, sizeof(float) * kPosDim, cudaMemcpyDeviceToDevice, s));
    VETO_CUDA(cudaMemcpyAsync(g_bn_w, packed + 5 * kPosDim, sizeof(float) * 4, cudaMemcpyDeviceToDevice, s));
    VETO_CUDA(cudaMemcpyAsync(g_bn_b, packed + 5 * kPosDim + 4, sizeof(float) * 4, cudaMemcpyDeviceToDevice, s));
    return VETO_OK;
}

int embed_bwd(const float* d_emb, const int64_t* labels, const float* obj_logits, int num_obj, int n_boxes, float* g_embed,
              cudaStream_t s) {
    if (labels) embed_bwd_kernel<<<num_obj, 256, 0, s>>>(d_emb, labels, n_boxes, g_embed);
    else embed_soft_bwd_kernel<<<num_obj, 256, 0, s>>>(d_emb, obj_logits, num_obj, n_boxes, g_embed);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int unpatchify(const float* d_patch, int n_boxes, float* d_roi, cudaStream_t s) {
    if (n_boxes <= 0) return VETO_OK;
    dim3 grid(n_boxes, kChannels / 32);
    unpatchify_kernel<<<grid, 256, 0, s>>>(d_patch, d_roi);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int unpack_halves(const float* g_packed, float* g_src, int out, int in, cudaStream_t s) {
    unpack_halves_kernel<<<grid_cap(((size_t)2 * out * in + 255) / 256, 8), 256, 0, s>>>(g_packed, g_src, out, in);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
int unpack_patch(const float* g_packed, float* g_src, int out, cudaStream_t s) {
    unpack_patch_kernel<<<grid_cap(((size_t)2 * out * kPatchVec + 255) / 256, 8), 256, 0, s>>>(g_packed, g_src, out);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int copy_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return VETO_OK;
    copy_rows_kernel<<<grid_cap((size_t)((rows * cols + 255) / 256), 8), 256, 0, s>>>(src, ld_src, dst, ld_dst, rows, cols);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
