// Box-level stage of the relation head and the weight re-packing that makes it possible.
//
// The reference builds, per PAIR, cat(subject, object) vectors and pushes them through Linear layers
// (roi_relation_predictors.py:4118-4123, model_veto.py:99-115): 2 x 131 KB of gathered ROI features and
// 18.9 M MACs of patch projection per pair.  Every one of those Linears is linear in the concatenation, so
// W . cat(s, o) = W_s . s + W_o . o: here each projection is evaluated once per BOX (N boxes instead of
// N(N-1) pairs) and the pair token becomes a gather + add (tokens.cu).  SURVEY.md §7 step 4.
#include "stages.cuh"

namespace veto {
namespace {

// ---------------------------------------------------------------- packing kernels
__global__ void pack_halves_kernel(const float* __restrict__ src, float* __restrict__ dst, int out, int in) {
    const int total = 2 * out * in;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e % in, r = e / in;
        const int h = r / out, o = r - h * out;
        dst[e] = src[(size_t)o * 2 * in + h * in + i];
    }
}

__global__ void pack_patch_kernel(const float* __restrict__ src, float* __restrict__ dst, int out) {
    const int total = 2 * out * kPatchVec;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int col = e % kPatchVec, r = e / kPatchVec;
        const int h = r / out, o = r - h * out;
        const int p = col / kChannels, c = col - p * kChannels;
        dst[e] = src[(size_t)o * 2 * kPatchVec + p * 2 * kChannels + h * kChannels + c];
    }
}

__global__ void pack_bias2_kernel(const float* __restrict__ b, float* __restrict__ dst, int out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * out; e += gridDim.x * blockDim.x) dst[e] = e < out ? b[e] : 0.f;
}

__global__ void pack_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst, int n) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) dst[e] = a[e] + b[e];
}

__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 h, l;
        split_bf16(src[e], h, l);
        hi[e] = h;
        if (lo) lo[e] = l;
    }
}

struct SplitJobs {
    SplitJob j[kMaxSplitJobs];
};
// blockIdx.y = array, blockIdx.x strides over its elements four at a time (every n is a multiple of 4)
__global__ void split_bf16_multi_kernel(const __grid_constant__ SplitJobs jobs) {
    const SplitJob& jb = jobs.j[blockIdx.y];
    const size_t n4 = jb.n >> 2;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(jb.src) + e);
        if (jb.col_scale) {   // LayerNorm weight folded into the Linear weight (row_len % 4 == 0)
            const float4 g = __ldg(reinterpret_cast<const float4*>(jb.col_scale + (4 * e) % (size_t)jb.row_len));
            v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
        }
        uint2 hi, lo;
        if (jb.fmt == FMT_F16C8) {
            // weight side of the f16c8 format: fp16(2048 w); bytes (e4m3(8 w), e4m3(8 r_w)) with r_w = 2048 w - fp16(2048 w)
            const float4 ws = make_float4(v.x * kC8WScale, v.y * kC8WScale, v.z * kC8WScale, v.w * kC8WScale);
            hi.x = f16x2_sat(ws.x, ws.y);
            hi.y = f16x2_sat(ws.z, ws.w);
            reinterpret_cast<uint2*>(jb.hi)[e] = hi;
            if (jb.lo) {
                const float2 a = f16x2_to_float(hi.x), b = f16x2_to_float(hi.y);
                uint8_t* p8 = (uint8_t*)jb.lo + c8_byte(4 * e);
                *(uint32_t*)p8 = e4m3x2(v.x * kC8WVal, v.y * kC8WVal) | (e4m3x2(v.z * kC8WVal, v.w * kC8WVal) << 16);
                *(uint32_t*)(p8 + 64) = e4m3x2((ws.x - a.x) * kC8WRes, (ws.y - a.y) * kC8WRes) |
                                        (e4m3x2((ws.z - b.x) * kC8WRes, (ws.w - b.y) * kC8WRes) << 16);
            }
            continue;
        }
        if (jb.fmt == FMT_F16C8_ACT) {   // activation side (test hook: production activations come out of the kernels' epilogues)
            store_act4_f16c8(jb.hi, jb.lo, 4 * e, v);
            continue;
        }
        split_pair(v.x, v.y, hi.x, lo.x);
        split_pair(v.z, v.w, hi.y, lo.y);
        reinterpret_cast<uint2*>(jb.hi)[e] = hi;
        if (jb.lo) reinterpret_cast<uint2*>(jb.lo)[e] = lo;
    }
}

// one warp per output n: c1 = sum_k gamma_k W[n,k], c2 = sum_k beta_k W[n,k] + bias[n]
__global__ void ln_fold_consts_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ bias, int N, int K, float* __restrict__ c1, float* __restrict__ c2) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float a = 0.f, b = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = W[(size_t)n * K + k];
        a = fmaf(gamma[k], w, a);
        b = fmaf(beta[k], w, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        c1[n] = a;
        c2[n] = b + (bias ? bias[n] : 0.f);
    }
}

int grid_for(size_t n) {
    const size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

// ---------------------------------------------------------------- box embeddings
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    float r = sh[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
    return r;
}

// one CTA (128 threads) per box
__global__ void __launch_bounds__(128)
box_embed_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ labels, const float* __restrict__ obj_logits,
                 int num_obj, const float* __restrict__ obj_embed, const float* __restrict__ bn_w,
                 const float* __restrict__ bn_b, const float* __restrict__ bn_mean, const float* __restrict__ bn_var,
                 const float* __restrict__ pos_w, const float* __restrict__ pos_b, float* __restrict__ pos_out,
                 float* __restrict__ emb_out, DropSpec pos_drop) {
    __shared__ float prob[512];
    __shared__ float red[4];
    const int n = blockIdx.x, t = threadIdx.x;
    const float4 bx = __ldg((const float4*)boxes + n);
    // BoxList.convert('xywh') (+1, bounding_box.py:72-75) then center_xywh (model_mpv2.py:342-345)
    const float w = bx.z - bx.x + 1.f, h = bx.w - bx.y + 1.f;
    const float in[4] = {bx.x + 0.5f * w, bx.y + 0.5f * h, w, h};
    float bn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) bn[k] = (in[k] - bn_mean[k]) / sqrtf(bn_var[k] + 1e-5f) * bn_w[k] + bn_b[k];
    if (t < kPosDim) {
        float a = pos_b[t];
#pragma unroll
        for (int k = 0; k < 4; ++k) a = fmaf(pos_w[t * 4 + k], bn[k], a);
        a = fmaxf(a, 0.f);
        if (pos_drop.thr16) a *= drop_scale1(pos_drop, (uint64_t)n * kPosDim + t);  // nn.Dropout(0.1), training only
        pos_out[(size_t)n * kPosDim + t] = a;
    }
    if (labels) {
        const int64_t lab = labels[n];
        for (int d = t; d < kEmbDim; d += blockDim.x) emb_out[(size_t)n * kEmbDim + d] = obj_embed[lab * kEmbDim + d];
    } else {
        // F.softmax(obj_logits, 1) @ obj_embed.weight (roi_relation_predictors.py:4095)
        const float* lg = obj_logits + (size_t)n * num_obj;
        float m = -INFINITY;
        for (int c = t; c < num_obj; c += blockDim.x) m = fmaxf(m, lg[c]);
        m = block_reduce(m, true, red);
        float sum = 0.f;
        for (int c = t; c < num_obj; c += blockDim.x) {
            const float e = expf(lg[c] - m);
            prob[c] = e;
            sum += e;
        }
        sum = block_reduce(sum, false, red);
        __syncthreads();
        for (int d = t; d < kEmbDim; d += blockDim.x) {
            float a = 0.f;
            for (int c = 0; c < num_obj; ++c) a = fmaf(prob[c] / sum, obj_embed[(size_t)c * kEmbDim + d], a);
            emb_out[(size_t)n * kEmbDim + d] = a;
        }
    }
}

// ---------------------------------------------------------------- patchify
// grid (N, 256/32): 32 channels x 64 pixels staged through shared memory so that both the NCHW read and
// the (p1 p2 c)-ordered write are coalesced.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ roi, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
    __shared__ float tile[32][65];
    const int n = blockIdx.x, c0 = blockIdx.y * 32, t = threadIdx.x;
    const float* src = roi + ((size_t)n * kChannels + c0) * 64;
    for (int e = t; e < 32 * 64; e += 256) tile[e >> 6][e & 63] = src[e];
    __syncthreads();
    // 64 (patch, pp) combinations x 32 channels
    for (int e = t; e < 64 * 32; e += 256) {
        const int c = e & 31, q = e >> 5;
        const int patch = q >> 2, pp = q & 3;
        const int ph = patch >> 2, pw = patch & 3, p1 = pp >> 1, p2 = pp & 1;
        const float v = tile[c][(2 * ph + p1) * 8 + 2 * pw + p2];
        const size_t o = ((size_t)n * kPatches + patch) * kPatchVec + pp * kChannels + c0 + c;
        if (out_f32) out_f32[o] = v;
        if (out_hi) {
            __nv_bfloat16 hh, ll;
            split_bf16(v, hh, ll);
            out_hi[o] = hh;
            if (out_lo) out_lo[o] = ll;
        }
    }
}

}  // namespace

int pack_halves(const float* src, float* dst, int out, int in, cudaStream_t s) {
    pack_halves_kernel<<<grid_for((size_t)2 * out * in), 256, 0, s>>>(src, dst, out, in);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
int pack_patch(const float* src, float* dst, int out, cudaStream_t s) {
    pack_patch_kernel<<<grid_for((size_t)2 * out * kPatchVec), 256, 0, s>>>(src, dst, out);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
int pack_bias2(const float* b, float* dst, int out, cudaStream_t s) {
    pack_bias2_kernel<<<grid_for((size_t)2 * out), 256, 0, s>>>(b, dst, out);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
int pack_add(const float* a, const float* b, float* dst, int n, cudaStream_t s) {
    pack_add_kernel<<<grid_for((size_t)n), 256, 0, s>>>(a, b, dst, n);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
int pack_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t s) {
    split_bf16_kernel<<<grid_for(n), 256, 0, s>>>(src, hi, lo, n);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int ln_fold_consts(const float* W, const float* gamma, const float* beta, const float* bias, int N, int K, float* c1, float* c2,
                   cudaStream_t s) {
    ln_fold_consts_kernel<<<(N + 7) / 8, 256, 0, s>>>(W, gamma, beta, bias, N, K, c1, c2);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int pack_split_bf16_multi(const SplitJob* jobs, int count, cudaStream_t s) {
    if (count <= 0) return VETO_OK;
    VETO_REQUIRE(count <= kMaxSplitJobs, VETO_ERR_ARG, "pack_split_bf16_multi: %d arrays > %d", count, kMaxSplitJobs);
    SplitJobs J{};
    size_t n_max = 0;
    for (int i = 0; i < count; ++i) {
        VETO_REQUIRE(jobs[i].src && jobs[i].hi && jobs[i].n % 4 == 0 && (!jobs[i].col_scale || (jobs[i].row_len > 0 && jobs[i].row_len % 4 == 0)),
                     VETO_ERR_ARG, "pack_split_bf16_multi: bad array %d", i);
        J.j[i] = jobs[i];
        n_max = jobs[i].n > n_max ? jobs[i].n : n_max;
    }
    const size_t bx = (n_max / 4 + 255) / 256;
    dim3 grid((unsigned)(bx < 64 ? (bx ? bx : 1) : 64), (unsigned)count);
    split_bf16_multi_kernel<<<grid, 256, 0, s>>>(J);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int box_embed(const float* boxes, const int64_t* labels, const float* obj_logits, int num_obj, int n_boxes,
              const veto_weights& w, float* pos_out, float* emb_out, cudaStream_t s, const float* batch_stats,
              const DropSpec& pos_drop) {
    if (n_boxes <= 0) return VETO_OK;
    VETO_REQUIRE(labels || obj_logits, VETO_ERR_ARG, "box_embed: need labels or obj_logits");
    VETO_REQUIRE(num_obj <= 512, VETO_ERR_UNSUPPORTED, "box_embed: num_obj=%d > 512", num_obj);
    // training mode normalises with the batch statistics (train.cu bn_batch_stats: mean[4], biased var[4])
    const float* mean = batch_stats ? batch_stats : w.bn_mean;
    const float* var = batch_stats ? batch_stats + 4 : w.bn_var;
    box_embed_kernel<<<n_boxes, 128, 0, s>>>(boxes, labels, obj_logits, num_obj, w.obj_embed, w.bn_weight, w.bn_bias, mean, var,
                                             w.pos_w, w.pos_b, pos_out, emb_out, pos_drop);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int patchify(const float* roi, int n_boxes, const ActOut& out, cudaStream_t s) {
    if (n_boxes <= 0) return VETO_OK;
    dim3 grid(n_boxes, kChannels / 32);
    patchify_kernel<<<grid, 256, 0, s>>>(roi, out.f32, out.hi, out.lo);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
