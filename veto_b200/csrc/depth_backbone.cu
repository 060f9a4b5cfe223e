// f3: the depth backbone — ResNetDepth (pysgg/modeling/backbone/resnet_depth.py:11-47: torchvision ResNet-18 with a
// one-channel conv1, truncated after layer3; built by backbone.py:83-93) forward in eval()/train() mode and backward.
//
// Activations live in HBM as NHWC fp32 ([B*H*W, C] row-major = the GEMM's row-major M x N) plus, in the tensor-core
// modes, a bf16 hi/lo copy that is the A operand of the next convolution.
//   * every convolution but conv1: implicit GEMM (conv_tc, gemm_tc.cu) — an output tile is an 8 x 16 pixel patch, the A
//     tile of a (tap, 64-channel block) one 4-D TMA box of the activation at the tap's offset (element strides for the
//     stride-2 ones), zero fill = padding.  Input gradient of the stride-1 3x3 ones: the same kernel over d(output) with
//     the flipped, transposed weights; of the strided ones: dcol = dY @ Wp gathered back to NHWC by col2im (no atomics:
//     every input pixel sums its taps).  Weight gradient: gemm_tn2_conv, pixel-patch K blocks, tap-shifted boxes.
//   * conv1 (one input channel) and everything in fp32 mode: im2col into the operand format -> the library GEMM against
//     the weights repacked as [Cout, k*k*Cin]; weight gradient dWp = dY^T @ col on the MN-major tcgen05 GEMM
// BatchNorm2d on batch statistics is a two-stage column reduction over the [M, C] rows (fp32 partials per block, fp64
// finalise in a fixed order: deterministic) and a fused normalise + residual + ReLU pass; its backward is the same
// shape (two column sums, one elementwise pass that emits d(conv output) directly in the GEMM operand format).
// The last block writes NCHW as well, the layout the Pooler (veto_roi_gather_forward) reads.
#include "api_internal.cuh"
#include "train.cuh"

namespace veto {
namespace {

constexpr int kConvs = VETO_DEPTH_CONVS;
constexpr int kMaxSplit = 64;    // weight-gradient K slices: Cout = 64 gives few output tiles, the rows are many
constexpr float kBnEps = 1e-5f;
constexpr int kBnMaxBlocks = 592;   // 4 per SM

struct ConvSpec {
    int cin, cout, k, stride, pad;
};
// module order (include/veto_b200.h): conv1 | layer1 | layer2 (5 = stride 2, 7 = downsample) | layer3 (10, 12 likewise)
const ConvSpec kSpec[kConvs] = {
    {1, 64, 7, 2, 3},
    {64, 64, 3, 1, 1}, {64, 64, 3, 1, 1}, {64, 64, 3, 1, 1}, {64, 64, 3, 1, 1},
    {64, 128, 3, 2, 1}, {128, 128, 3, 1, 1}, {64, 128, 1, 2, 0}, {128, 128, 3, 1, 1}, {128, 128, 3, 1, 1},
    {128, 256, 3, 2, 1}, {256, 256, 3, 1, 1}, {128, 256, 1, 2, 0}, {256, 256, 3, 1, 1}, {256, 256, 3, 1, 1},
};
struct BlockSpec {
    int c1, c2, ds;
};
const BlockSpec kBlocks[6] = {{1, 2, -1}, {3, 4, -1}, {5, 6, 7}, {8, 9, -1}, {10, 11, 12}, {13, 14, -1}};

inline int out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }
inline int64_t pad64(int64_t v) { return (v + 63) / 64 * 64; }
inline size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

struct ConvDims {
    int hin, win, hout, wout, K, Kp;
    int64_t M;
};

struct DepthLayout {
    ConvDims d[kConvs];
    int hp, wp;  // after the max-pool
    size_t col[kConvs], raw[kConvs], y[kConvs], yop[kConvs], stat[kConvs];  // stat: mean[C], rstd[C]; yop: y as GEMM operand
    size_t wp_[kConvs], wpT[kConvs];  // wpT: [Kp,Cout] transpose, or for the implicit convolutions the flipped [Cin, 9*Cout]
    bool has_yop[kConvs];
    int src[kConvs];                  // which tensor feeds conv i: -2 the depth image, -1 the max-pool output, else y[src]
    size_t pool, pool_op, pool_arg, bn_partial, bn_sums;
    size_t dcol, gA, gB, gC, dyop, dwp, splitk, T1, T2;
    size_t total;
};

DepthLayout depth_layout(int prec, int B, int H, int W, bool training) {
    DepthLayout L{};
    Carver k;
    const size_t f = sizeof(float);
    int h = H, w = W;
    auto set = [&](int i, int hin, int win) {
        const ConvSpec& c = kSpec[i];
        ConvDims& d = L.d[i];
        d.hin = hin; d.win = win;
        d.hout = out_dim(hin, c.k, c.stride, c.pad);
        d.wout = out_dim(win, c.k, c.stride, c.pad);
        d.K = c.k * c.k * c.cin;
        d.Kp = (int)pad64(d.K);
        d.M = (int64_t)B * d.hout * d.wout;
    };
    set(0, h, w);
    L.src[0] = -2;
    L.hp = out_dim(L.d[0].hout, 3, 2, 1);
    L.wp = out_dim(L.d[0].wout, 3, 2, 1);
    h = L.hp; w = L.wp;
    int prev = -1;
    for (const BlockSpec& b : kBlocks) {
        set(b.c1, h, w);
        L.src[b.c1] = prev;
        if (b.ds >= 0) {
            set(b.ds, h, w);
            L.src[b.ds] = prev;
        }
        set(b.c2, L.d[b.c1].hout, L.d[b.c1].wout);
        L.src[b.c2] = b.c1;
        prev = b.c2;
        h = L.d[b.c2].hout; w = L.d[b.c2].wout;
    }
    size_t col_max = 0, act_max = 0, yop_max = 0, w_max = 0;
    for (int i = 0; i < kConvs; ++i) {
        const ConvDims& d = L.d[i];
        const size_t M = (size_t)(d.M > 0 ? d.M : 1);
        col_max = max_sz(col_max, M * d.Kp);
        act_max = max_sz(act_max, M * kSpec[i].cout);
        yop_max = max_sz(yop_max, M * kSpec[i].cout);
        w_max = max_sz(w_max, (size_t)kSpec[i].cout * d.Kp);
    }
    // One im2col buffer serves every convolution that needs one (forward: the strided / first / fp32-mode ones; backward:
    // re-derived from the saved input right before each weight gradient).  Only conv1 keeps its own: its input, the
    // depth image, is not available to the backward call.
    const size_t shared_col = k.take(act_bytes(prec, col_max));
    const bool ops = prec != VETO_PREC_FP32;
    // the bf16 operand copy of y[i] is written only where an implicit (stride-1 3x3) convolution reads it
    bool feeds_implicit[kConvs] = {};
    for (int j = 0; j < kConvs; ++j)
        if (L.src[j] >= 0 && kSpec[j].cin % 64 == 0) feeds_implicit[L.src[j]] = true;
    for (int i = 0; i < kConvs; ++i) {
        const ConvDims& d = L.d[i];
        const size_t M = (size_t)(d.M > 0 ? d.M : 1);
        const int C = kSpec[i].cout;
        L.col[i] = (training && i == 0) ? k.take(act_bytes(prec, M * d.Kp)) : shared_col;
        L.raw[i] = k.take(f * M * C);
        L.y[i] = k.take(f * M * C);
        L.has_yop[i] = ops && feeds_implicit[i];
        L.yop[i] = L.has_yop[i] ? k.take(act_bytes(prec, M * C)) : 0;
        L.stat[i] = k.take(f * 2 * C);
        L.wp_[i] = k.take(act_bytes(prec, (size_t)C * d.Kp));
        L.wpT[i] = training ? k.take(act_bytes(prec, (size_t)C * d.Kp)) : 0;
    }
    L.pool = k.take(f * (size_t)B * L.hp * L.wp * 64);
    L.pool_op = ops ? k.take(act_bytes(prec, (size_t)B * L.hp * L.wp * 64)) : 0;
    L.pool_arg = k.take(training ? (size_t)B * L.hp * L.wp * 64 : 0);
    L.bn_partial = k.take(f * (size_t)kBnMaxBlocks * 2 * 256);
    L.bn_sums = k.take(f * 2 * 256);
    if (training) {
        L.dcol = k.take(f * col_max);
        L.gA = k.take(f * act_max);
        L.gB = k.take(f * act_max);
        L.gC = k.take(f * act_max);
        L.dyop = k.take(act_bytes(prec, yop_max));
        L.dwp = k.take(f * w_max);
        const bool simt = prec == VETO_PREC_FP32;
        L.splitk = k.take(simt ? 0 : f * kMaxSplit * w_max);
        size_t t1 = 0, t2 = 0;
        if (simt)
            for (int i = 0; i < kConvs; ++i) {
                t1 = max_sz(t1, (size_t)kSpec[i].cout * pad64(L.d[i].M));
                t2 = max_sz(t2, (size_t)L.d[i].Kp * pad64(L.d[i].M));
            }
        L.T1 = k.take(f * t1);
        L.T2 = k.take(f * t2);
    }
    L.total = k.off;
    return L;
}

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
// idx -> (pixel, channel group) for a power-of-two number of channel groups (C / 4 with C = 64, 128 or 256), and pixel ->
// (image, row, column) in 32-bit arithmetic (pixel counts are < 2^31, check_args): the 64-bit divisions of the plain
// form cost more issue slots than the 16-byte accesses these kernels exist for
__device__ __forceinline__ void split_groups(int64_t idx, int groups, int64_t& pixel, int& g) {
    const int shift = __ffs(groups) - 1;
    pixel = idx >> shift;
    g = (int)(idx & (groups - 1));
}
__device__ __forceinline__ void split_pixel(int64_t pixel, int height, int width, int& b, int& h, int& w) {
    const unsigned p = (unsigned)pixel;
    const unsigned t = p / (unsigned)width;
    w = (int)(p - t * (unsigned)width);
    b = (int)(t / (unsigned)height);
    h = (int)(t - (unsigned)b * (unsigned)height);
}

__device__ __forceinline__ void store8(const ActOut& o, size_t at, const float (&v)[8]) {
    if (o.f32) {
        *reinterpret_cast<float4*>(o.f32 + at) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o.f32 + at + 4) = make_float4(v[4], v[5], v[6], v[7]);
        return;
    }
    uint4 hi, lo;
    split_pair(v[0], v[1], hi.x, lo.x);
    split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z);
    split_pair(v[6], v[7], hi.w, lo.w);
    *reinterpret_cast<uint4*>(o.hi + at) = hi;
    if (o.lo) *reinterpret_cast<uint4*>(o.lo + at) = lo;
}
__device__ __forceinline__ void store4(const ActOut& o, size_t at, float4 v) {
    if (o.f32) {
        *reinterpret_cast<float4*>(o.f32 + at) = v;
        return;
    }
    uint2 hi, lo;
    split_pair(v.x, v.y, hi.x, lo.x);
    split_pair(v.z, v.w, hi.y, lo.y);
    *reinterpret_cast<uint2*>(o.hi + at) = hi;
    if (o.lo) *reinterpret_cast<uint2*>(o.lo + at) = lo;
}

// col[m, (kh*k + kw)*Cin + c] = x[b, ho*s - p + kh, wo*s - p + kw, c] (zero outside the image and for k >= K).
// One thread per 8 consecutive columns; VEC: Cin % 8 == 0, the 8 columns are one tap's consecutive channels.
// (index type: 32-bit when M * Kp / 8 fits — the 64-bit divisions of the decode cost more than the copy itself)
// KS1 > 0: one input channel and a KS1 x KS1 window known at compile time (conv1: the tap decode is divisions by constants)
template <bool VEC, typename idx_t, int KS1 = 0>
__global__ void im2col_kernel(const float* __restrict__ x, int hin, int win, int cin, int ks, int stride, int pad, int hout,
                              int wout, int K, int Kp, int64_t M, ActOut col) {
    const int groups = Kp >> 3;
    const idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((int64_t)idx >= M * groups) return;
    const idx_t m = idx / (idx_t)groups;
    const int k0 = (int)(idx - m * (idx_t)groups) << 3;
    const int wo = (int)(m % (idx_t)wout);
    const idx_t t = m / (idx_t)wout;
    const int ho = (int)(t % (idx_t)hout);
    const int64_t b = (int64_t)(t / (idx_t)hout);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (VEC) {
        if (k0 < K) {
            const int tap = k0 / cin, c0 = k0 - tap * cin;
            const int kh = tap / ks, kw = tap - kh * ks;
            const int hi = ho * stride - pad + kh, wi = wo * stride - pad + kw;
            if (hi >= 0 && hi < hin && wi >= 0 && wi < win) {
                const float4* src = reinterpret_cast<const float4*>(x + ((b * hin + hi) * win + wi) * cin + c0);
                const float4 a = __ldg(src), c = __ldg(src + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = k0 + j;
            if (kk < K) {
                const int tap = KS1 ? kk : kk / cin, c = KS1 ? 0 : kk - tap * cin;
                const int kh = KS1 ? tap / KS1 : tap / ks, kw = tap - kh * (KS1 ? KS1 : ks);
                const int hi = ho * stride - pad + kh, wi = wo * stride - pad + kw;
                if (hi >= 0 && hi < hin && wi >= 0 && wi < win) v[j] = __ldg(x + ((b * hin + hi) * win + wi) * cin + c);
            }
        }
    }
    store8(col, (size_t)m * Kp + k0, v);
}

// dx[b,h,w,c] (+)= sum over the taps that read this pixel of dcol[m(b,ho,wo), tap*Cin + c]  (+ add[..] where mask > 0)
__global__ void col2im_kernel(const float* __restrict__ dcol, int hin, int win, int cin, int ks, int stride, int pad, int hout,
                              int wout, int Kp, int64_t pixels, const float* __restrict__ add, const float* __restrict__ mask,
                              bool accumulate, float* dx) {
    const int groups = cin >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pixels * groups) return;
    int64_t p;
    int cgi, bi, h, w;
    split_groups(idx, groups, p, cgi);
    split_pixel(p, hin, win, bi, h, w);
    const int c0 = cgi << 2;
    const int64_t b = bi;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kh = 0; kh < ks; ++kh) {
        const int hs = h + pad - kh;
        if (hs < 0 || hs % stride) continue;
        const int ho = hs / stride;
        if (ho >= hout) continue;
        for (int kw = 0; kw < ks; ++kw) {
            const int ws = w + pad - kw;
            if (ws < 0 || ws % stride) continue;
            const int wo = ws / stride;
            if (wo >= wout) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(dcol + ((b * hout + ho) * wout + wo) * Kp +
                                                                    (kh * ks + kw) * cin + c0));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const size_t at = (size_t)p * cin + c0;
    if (add) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(add + at));
        const float4 mk = __ldg(reinterpret_cast<const float4*>(mask + at));
        acc.x += mk.x > 0.f ? a.x : 0.f;
        acc.y += mk.y > 0.f ? a.y : 0.f;
        acc.z += mk.z > 0.f ? a.z : 0.f;
        acc.w += mk.w > 0.f ? a.w : 0.f;
    }
    if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dx + at);
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    *reinterpret_cast<float4*>(dx + at) = acc;
}

// Column sums over [M, C] rows, two quantities per column:
//   MODE 0 (forward statistics): sum x, sum x^2
//   MODE 1 (backward):           sum dz, sum dz * xhat   with dz = dy * (mask > 0 or no mask), xhat = (x - mean) * rstd
// 256 threads: C/4 column groups x 1024/C row lanes; partial[block][2][C]
template <int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ mask, const float* __restrict__ stat, int C,
                                                        int64_t M, int64_t rows_per_block, float* __restrict__ partial) {
    __shared__ float4 sh[2][256];
    const int groups = C >> 2, lanes = 256 / groups;
    const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, mean = s0, rstd = s0;
    if (MODE == 1) {
        mean = __ldg(reinterpret_cast<const float4*>(stat) + g);
        rstd = __ldg(reinterpret_cast<const float4*>(stat + C) + g);
    }
    int64_t r = r0 + lane;
    if (MODE == 0) {
        // four rows in flight per thread: one 16-byte load per iteration leaves the statistics pass latency-bound
        for (; r + 3 * lanes < r1; r += 4 * lanes) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(x + (size_t)(r + j * lanes) * C + 4 * g));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s0.x += v[j].x; s0.y += v[j].y; s0.z += v[j].z; s0.w += v[j].w;
                s1.x = fmaf(v[j].x, v[j].x, s1.x); s1.y = fmaf(v[j].y, v[j].y, s1.y);
                s1.z = fmaf(v[j].z, v[j].z, s1.z); s1.w = fmaf(v[j].w, v[j].w, s1.w);
            }
        }
    }
    for (; r < r1; r += lanes) {
        const size_t at = (size_t)r * C + 4 * g;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + at));
        if (MODE == 0) {
            s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
            s1.x = fmaf(v.x, v.x, s1.x); s1.y = fmaf(v.y, v.y, s1.y); s1.z = fmaf(v.z, v.z, s1.z); s1.w = fmaf(v.w, v.w, s1.w);
        } else {
            float4 d = __ldg(reinterpret_cast<const float4*>(dy + at));
            if (mask) {
                const float4 mk = __ldg(reinterpret_cast<const float4*>(mask + at));
                d.x = mk.x > 0.f ? d.x : 0.f; d.y = mk.y > 0.f ? d.y : 0.f;
                d.z = mk.z > 0.f ? d.z : 0.f; d.w = mk.w > 0.f ? d.w : 0.f;
            }
            s0.x += d.x; s0.y += d.y; s0.z += d.z; s0.w += d.w;
            s1.x = fmaf(d.x, (v.x - mean.x) * rstd.x, s1.x); s1.y = fmaf(d.y, (v.y - mean.y) * rstd.y, s1.y);
            s1.z = fmaf(d.z, (v.z - mean.z) * rstd.z, s1.z); s1.w = fmaf(d.w, (v.w - mean.w) * rstd.w, s1.w);
        }
    }
    sh[0][threadIdx.x] = s0;
    sh[1][threadIdx.x] = s1;
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 a = sh[0][l * groups + g], b = sh[1][l * groups + g];
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
        }
        float* out = partial + (size_t)blockIdx.x * 2 * C;
        *reinterpret_cast<float4*>(out + 4 * g) = s0;
        *reinterpret_cast<float4*>(out + C + 4 * g) = s1;
    }
}

// Sum of the per-block partials of one column in fp64, in a fixed order: 1024 threads = 32 columns x 32 lanes, lane l
// takes blocks l, l + 32, ...; the 32 lane sums are combined in lane order.  Returns the totals to the lane-0 threads.
constexpr int kFinLanes = 32;
__device__ __forceinline__ bool reduce_partials(const float* __restrict__ partial, int blocks, int C, int c, int lane,
                                                double& s, double& q) {
    __shared__ double sh[2][kFinLanes][32];
    s = 0.0;
    q = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int b = lane; b < blocks; b += kFinLanes) {
            s += (double)__ldg(partial + (size_t)b * 2 * C + c);
            q += (double)__ldg(partial + (size_t)b * 2 * C + C + c);
        }
    }
    sh[0][lane][threadIdx.x & 31] = s;
    sh[1][lane][threadIdx.x & 31] = q;
    __syncthreads();
    if (lane != 0 || c >= C) return false;
    for (int l = 1; l < kFinLanes; ++l) {
        s += sh[0][l][threadIdx.x & 31];
        q += sh[1][l][threadIdx.x & 31];
    }
    return true;
}

// forward: mean / rstd from the partial sums (fp64, fixed order), running statistics as nn.BatchNorm2d updates them
__global__ void __launch_bounds__(1024) bn_stats_finalize_kernel(const float* __restrict__ partial, int blocks, int C, int64_t M,
                                                                float momentum, float* stat, float* running_mean,
                                                                float* running_var) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    double s, q;
    if (!reduce_partials(partial, blocks, C, c, threadIdx.x >> 5, s, q)) return;
    const double mean = s / (double)M;
    double var = q / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[c] = (float)mean;
    stat[C + c] = (float)(1.0 / sqrt(var + (double)kBnEps));
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
}
// eval mode: the same mean / rstd pair from the running statistics
__global__ void bn_stats_eval_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int C,
                                     float* stat) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    stat[c] = running_mean[c];
    stat[C + c] = 1.f / sqrtf(running_var[c] + kBnEps);
}
// backward: sums[0..C) = sum dz = g_beta, sums[C..2C) = sum dz * xhat = g_gamma
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int blocks, int C, float* sums,
                                                              float* g_gamma, float* g_beta) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    double s, q;
    if (!reduce_partials(partial, blocks, C, c, threadIdx.x >> 5, s, q)) return;
    sums[c] = (float)s;
    sums[C + c] = (float)q;
    g_beta[c] = (float)s;
    g_gamma[c] = (float)q;
}

// y = [relu]((x - mean) * rstd * gamma + beta [+ residual]); optionally also out_nchw[b, c, h*w] and yop = y as bf16 hi[/lo]
// (the A operand of the implicit-GEMM convolution that consumes it)
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stat, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ residual, int relu, int C, int64_t M,
                                int64_t hw, float* __restrict__ y, float* __restrict__ out_nchw, ActOut yop) {
    const int groups = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * groups) return;
    int64_t m;
    int g;
    split_groups(idx, groups, m, g);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + idx);
    const float4 mean = __ldg(reinterpret_cast<const float4*>(stat) + g);
    const float4 rstd = __ldg(reinterpret_cast<const float4*>(stat + C) + g);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + g);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + g);
    float4 o;
    o.x = (v.x - mean.x) * rstd.x * ga.x + be.x;
    o.y = (v.y - mean.y) * rstd.y * ga.y + be.y;
    o.z = (v.z - mean.z) * rstd.z * ga.z + be.z;
    o.w = (v.w - mean.w) * rstd.w * ga.w + be.w;
    if (residual) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(residual) + idx);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    reinterpret_cast<float4*>(y)[idx] = o;
    if (yop.hi) store4(yop, (size_t)idx * 4, o);
    if (out_nchw) {
        const int64_t b = m / hw, pix = m - b * hw;
        float* dst = out_nchw + ((size_t)b * C + 4 * g) * hw + pix;
        dst[0] = o.x; dst[hw] = o.y; dst[2 * hw] = o.z; dst[3 * hw] = o.w;
    }
}

// d(conv output) = gamma * rstd * (dz - sum(dz)/M - xhat * sum(dz*xhat)/M), in the GEMM operand format
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mask,
                                    const float* __restrict__ stat, const float* __restrict__ gamma,
                                    const float* __restrict__ sums, int C, int64_t M, ActOut out) {
    const int groups = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * groups) return;
    const int g = (int)(idx & (groups - 1));   // C / 4 is a power of two
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + idx);
    float4 d = __ldg(reinterpret_cast<const float4*>(dy) + idx);
    if (mask) {
        const float4 mk = __ldg(reinterpret_cast<const float4*>(mask) + idx);
        d.x = mk.x > 0.f ? d.x : 0.f; d.y = mk.y > 0.f ? d.y : 0.f;
        d.z = mk.z > 0.f ? d.z : 0.f; d.w = mk.w > 0.f ? d.w : 0.f;
    }
    const float4 mean = __ldg(reinterpret_cast<const float4*>(stat) + g);
    const float4 rstd = __ldg(reinterpret_cast<const float4*>(stat + C) + g);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + g);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sums) + g);
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(sums + C) + g);
    const float inv = 1.f / (float)M;
    float4 o;
    o.x = ga.x * rstd.x * (d.x - s0.x * inv - (v.x - mean.x) * rstd.x * s1.x * inv);
    o.y = ga.y * rstd.y * (d.y - s0.y * inv - (v.y - mean.y) * rstd.y * s1.y * inv);
    o.z = ga.z * rstd.z * (d.z - s0.z * inv - (v.z - mean.z) * rstd.z * s1.z * inv);
    o.w = ga.w * rstd.w * (d.w - s0.w * inv - (v.w - mean.w) * rstd.w * s1.w * inv);
    store4(out, (size_t)idx * 4, o);
}

// nn.MaxPool2d(3, 2, 1) on NHWC, 4 channels per thread.  arg (training) records which tap (kh * 3 + kw) holds the FIRST
// maximum in scan order — torch's `val > maxval` update rule; after a ReLU whole windows tie at zero.
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int hin, int win, int C, int hout, int wout, int64_t out_pixels,
                                   float* __restrict__ y, uchar4* __restrict__ arg, ActOut yop) {
    const int groups = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_pixels * groups) return;
    int64_t p;
    int g, bi, ho, wo;
    split_groups(idx, groups, p, g);
    split_pixel(p, hout, wout, bi, ho, wo);
    const int64_t b = bi;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 a = make_uchar4(255, 255, 255, 255);
    for (int kh = 0; kh < 3; ++kh) {
        const int h = ho * 2 - 1 + kh;
        if (h < 0 || h >= hin) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int w = wo * 2 - 1 + kw;
            if (w < 0 || w >= win) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((b * hin + h) * win + w) * C) + g);
            const unsigned char tap = (unsigned char)(kh * 3 + kw);
            if (v.x > m.x || a.x == 255) { m.x = v.x; a.x = tap; }
            if (v.y > m.y || a.y == 255) { m.y = v.y; a.y = tap; }
            if (v.z > m.z || a.z == 255) { m.z = v.z; a.z = tap; }
            if (v.w > m.w || a.w == 255) { m.w = v.w; a.w = tap; }
        }
    }
    reinterpret_cast<float4*>(y)[idx] = m;
    if (yop.hi) store4(yop, (size_t)idx * 4, m);
    if (arg) arg[idx] = a;
}

// backward, gather form: an input pixel sums the gradients of the <= 4 windows whose recorded arg-max tap is this pixel
__global__ void maxpool_bwd_kernel(const uchar4* __restrict__ arg, const float* __restrict__ dy, int hin, int win, int C,
                                   int hout, int wout, int64_t in_pixels, float* __restrict__ dx) {
    const int groups = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= in_pixels * groups) return;
    int64_t p;
    int g, bi, h, w;
    split_groups(idx, groups, p, g);
    split_pixel(p, hin, win, bi, h, w);
    const int64_t b = bi;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ho = h / 2; ho <= (h + 1) / 2 && ho < hout; ++ho) {
        const int kh = h - (ho * 2 - 1);
        for (int wo = w / 2; wo <= (w + 1) / 2 && wo < wout; ++wo) {
            const unsigned char tap = (unsigned char)(kh * 3 + (w - (wo * 2 - 1)));
            const size_t at = (((size_t)b * hout + ho) * wout + wo) * groups + g;
            const uchar4 a = __ldg(arg + at);
            const float4 d = __ldg(reinterpret_cast<const float4*>(dy) + at);
            acc.x += a.x == tap ? d.x : 0.f;
            acc.y += a.y == tap ? d.y : 0.f;
            acc.z += a.z == tap ? d.z : 0.f;
            acc.w += a.w == tap ? d.w : 0.f;
        }
    }
    reinterpret_cast<float4*>(dx)[idx] = acc;
}

// conv weight [Cout,Cin,k,k] -> Wp [Cout,Kp] (column = tap*Cin + c) and its transpose WpT [Kp,Cout], operand format
__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int ks, int K, int Kp, ActOut wp, ActOut wpT) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cout * Kp) return;
    const int o = idx / Kp, kk = idx - o * Kp;
    float v = 0.f;
    if (kk < K) {
        const int tap = kk / cin, c = kk - tap * cin;
        v = w[((size_t)o * cin + c) * ks * ks + tap];
    }
    auto put = [&](const ActOut& dst, size_t at) {
        if (dst.f32) dst.f32[at] = v;
        else if (dst.hi) {
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            dst.hi[at] = hi;
            if (dst.lo) dst.lo[at] = lo;
        }
    };
    put(wp, (size_t)idx);
    put(wpT, (size_t)kk * cout + o);
}
// conv weight [Cout,Cin,k,k] -> Wf [Cin, k*k*Cout] with column (kh'*k + kw')*Cout + o holding w[o, c, k-1-kh', k-1-kw']:
// the weights of the stride-1 convolution over d(output) that gives d(input)
__global__ void pack_conv_flipped_kernel(const float* __restrict__ w, int cout, int cin, int ks, ActOut wf) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int kk2 = ks * ks, Kf = kk2 * cout;
    if (idx >= cin * Kf) return;
    const int c = idx / Kf, r = idx - c * Kf;
    const int tap = r / cout, o = r - tap * cout;
    const float v = w[((size_t)o * cin + c) * kk2 + (kk2 - 1 - tap)];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    wf.hi[idx] = hi;
    if (wf.lo) wf.lo[idx] = lo;
}
// dx += add where mask > 0 (the identity branch of a BasicBlock through the block's ReLU)
__global__ void add_masked_kernel(float* __restrict__ dx, const float* __restrict__ add, const float* __restrict__ mask, int64_t n4) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n4) return;
    float4 v = reinterpret_cast<float4*>(dx)[idx];
    const float4 a = __ldg(reinterpret_cast<const float4*>(add) + idx);
    const float4 mk = __ldg(reinterpret_cast<const float4*>(mask) + idx);
    v.x += mk.x > 0.f ? a.x : 0.f;
    v.y += mk.y > 0.f ? a.y : 0.f;
    v.z += mk.z > 0.f ? a.z : 0.f;
    v.w += mk.w > 0.f ? a.w : 0.f;
    reinterpret_cast<float4*>(dx)[idx] = v;
}
__global__ void unpack_conv_grad_kernel(const float* __restrict__ gp, int cout, int cin, int ks, int Kp, float* __restrict__ g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int kk2 = ks * ks;
    if (idx >= cout * cin * kk2) return;
    const int tap = idx % kk2;
    const int c = (idx / kk2) % cin;
    const int o = idx / (kk2 * cin);
    g[idx] = gp[(size_t)o * Kp + tap * cin + c];
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int64_t hw, int64_t n, float* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int c = (int)(idx % C);
    const int64_t m = idx / C;
    const int64_t b = m / hw, pix = m - b * hw;
    dst[idx] = __ldg(src + ((size_t)b * C + c) * hw + pix);
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------------------------
struct Run {
    int prec;
    cudaStream_t s;
    char* B;
    const DepthLayout* L;
    const veto_depth_weights* w;
    int batch;
    bool training;
    float momentum;

    float* f32(size_t off) const { return (float*)(B + off); }
    ActBuf act(size_t off, size_t elems) const { return act_at(B, off, prec, elems); }
    ActBuf col(int i) const { return act(L->col[i], (size_t)L->d[i].M * L->d[i].Kp); }
    ActBuf wp(int i) const { return act(L->wp_[i], (size_t)kSpec[i].cout * L->d[i].Kp); }
    ActBuf wpT(int i) const { return act(L->wpT[i], (size_t)kSpec[i].cout * L->d[i].Kp); }

    // convolutions on >= 64 input channels run as implicit GEMMs (conv_tc: 4-D TMA boxes of the NHWC activation, nothing
    // materialised) in the tensor-core modes; conv1 goes through im2col.  VETO_DEPTH_IM2COL=1 forces im2col everywhere
    // (debugging).
    bool implicit(int i) const {
        static const bool off = [] { const char* e = getenv("VETO_DEPTH_IM2COL"); return e && e[0] == '1'; }();
        const ConvSpec& c = kSpec[i];
        return !off && prec != VETO_PREC_FP32 && c.cin % 64 == 0;
    }
    // the input gradient as a convolution of d(output) with the flipped weights exists for the stride-1 3x3 ones only
    bool implicit_dgrad(int i) const { return implicit(i) && kSpec[i].stride == 1 && kSpec[i].k == 3; }
    int passes() const { return prec == VETO_PREC_BF16X3 ? 3 : 1; }
    const float* src_f32(int i, const float* depth) const {
        const int src = L->src[i];
        return src == -2 ? depth : src == -1 ? f32(L->pool) : f32(L->y[src]);
    }
    ActBuf src_op(int i) const {
        const int src = L->src[i];
        if (src == -1) return act(L->pool_op, (size_t)batch * L->hp * L->wp * 64);
        return act(L->yop[src], (size_t)L->d[src].M * kSpec[src].cout);
    }
    ActOut yop_out(int i) const { return L->has_yop[i] ? act(L->yop[i], (size_t)L->d[i].M * kSpec[i].cout).out() : ActOut(); }

    int pack(int i) const {
        const ConvSpec& c = kSpec[i];
        const ConvDims& d = L->d[i];
        const bool flipped = training && implicit_dgrad(i);
        ActOut t = (training && !flipped) ? wpT(i).out() : ActOut();
        pack_conv_kernel<<<blocks_for((int64_t)c.cout * d.Kp, 256), 256, 0, s>>>(w->conv_w[i], c.cout, c.cin, c.k, d.K, d.Kp,
                                                                                wp(i).out(), t);
        VETO_LAUNCH_CHECK();
        if (flipped) {
            pack_conv_flipped_kernel<<<blocks_for((int64_t)c.cout * d.Kp, 256), 256, 0, s>>>(w->conv_w[i], c.cout, c.cin, c.k,
                                                                                            wpT(i).out());
            VETO_LAUNCH_CHECK();
        }
        return VETO_OK;
    }
    int im2col(int i, const float* x) const {
        const ConvSpec& c = kSpec[i];
        const ConvDims& d = L->d[i];
        const int64_t n = d.M * (d.Kp / 8);
        const bool small = n + 256 < ((int64_t)1 << 32);
        const unsigned grid = blocks_for(n, 256);
#define VETO_IM2COL(VEC, T) \
    im2col_kernel<VEC, T><<<grid, 256, 0, s>>>(x, d.hin, d.win, c.cin, c.k, c.stride, c.pad, d.hout, d.wout, d.K, d.Kp, d.M, col(i).out())
        if (c.cin % 8 == 0) {
            if (small) VETO_IM2COL(true, uint32_t);
            else VETO_IM2COL(true, int64_t);
        } else if (c.cin == 1 && c.k == 7 && small) {
            im2col_kernel<false, uint32_t, 7><<<grid, 256, 0, s>>>(x, d.hin, d.win, c.cin, c.k, c.stride, c.pad, d.hout, d.wout, d.K,
                                                                   d.Kp, d.M, col(i).out());
        } else {
            if (small) VETO_IM2COL(false, uint32_t);
            else VETO_IM2COL(false, int64_t);
        }
#undef VETO_IM2COL
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
    int conv(int i, const float* depth) const {
        const ConvSpec& c = kSpec[i];
        const ConvDims& d = L->d[i];
        int rc = pack(i);
        if (rc) return rc;
        const ActBuf W = wp(i);
        if (implicit(i)) {
            const ActBuf a = src_op(i);
            GemmOperand A, Wo;
            A.hi = a.hi; A.lo = a.lo;
            Wo.hi = W.hi; Wo.lo = W.lo;
            return conv_tc(A, Wo, batch, d.hin, d.win, c.cin, c.cout, c.k, c.pad, c.stride, passes(), f32(L->raw[i]), c.cout, s);
        }
        if ((rc = im2col(i, src_f32(i, depth)))) return rc;
        GemmEpilogue ep;
        ep.out.f32 = f32(L->raw[i]);
        ep.ldc = c.cout;
        return linear(prec, col(i), d.Kp, WRef{W.f32, W.hi, W.lo}, (int)d.M, c.cout, d.Kp, ep, s);
    }
    int reduce_blocks(int64_t M, int64_t* rows_per_block) const {
        int64_t rpb = (M + kBnMaxBlocks - 1) / kBnMaxBlocks;
        if (rpb < 64) rpb = 64;
        *rows_per_block = rpb;
        return (int)((M + rpb - 1) / rpb);
    }
    int bn(int i, const float* residual, bool relu, float* out_nchw) const {
        const int C = kSpec[i].cout;
        const int64_t M = L->d[i].M;
        float* stat = f32(L->stat[i]);
        if (training) {
            int64_t rpb;
            const int blocks = reduce_blocks(M, &rpb);
            bn_reduce_kernel<0><<<blocks, 256, 0, s>>>(f32(L->raw[i]), nullptr, nullptr, nullptr, C, M, rpb, f32(L->bn_partial));
            VETO_LAUNCH_CHECK();
            bn_stats_finalize_kernel<<<(C + 31) / 32, 1024, 0, s>>>(f32(L->bn_partial), blocks, C, M, momentum, stat,
                                                                    w->bn_mean[i], w->bn_var[i]);
            VETO_LAUNCH_CHECK();
        } else {
            bn_stats_eval_kernel<<<(C + 127) / 128, 128, 0, s>>>(w->bn_mean[i], w->bn_var[i], C, stat);
            VETO_LAUNCH_CHECK();
        }
        bn_apply_kernel<<<blocks_for(M * (C / 4), 256), 256, 0, s>>>(f32(L->raw[i]), stat, w->bn_w[i], w->bn_b[i], residual,
                                                                    relu ? 1 : 0, C, M, (int64_t)L->d[i].hout * L->d[i].wout,
                                                                    f32(L->y[i]), out_nchw, yop_out(i));
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }

    // ---- backward pieces ----
    // d(conv i output) from the gradient dy of the BatchNorm output (through the ReLU mask `mask > 0` when given)
    int bn_bwd(int i, const float* dy, const float* mask, const veto_depth_grads* g) const {
        const int C = kSpec[i].cout;
        const int64_t M = L->d[i].M;
        int64_t rpb;
        const int blocks = reduce_blocks(M, &rpb);
        bn_reduce_kernel<1><<<blocks, 256, 0, s>>>(f32(L->raw[i]), dy, mask, f32(L->stat[i]), C, M, rpb, f32(L->bn_partial));
        VETO_LAUNCH_CHECK();
        bn_bwd_finalize_kernel<<<(C + 31) / 32, 1024, 0, s>>>(f32(L->bn_partial), blocks, C, f32(L->bn_sums), g->bn_w[i],
                                                              g->bn_b[i]);
        VETO_LAUNCH_CHECK();
        bn_bwd_apply_kernel<<<blocks_for(M * (C / 4), 256), 256, 0, s>>>(f32(L->raw[i]), dy, mask, f32(L->stat[i]), w->bn_w[i],
                                                                        f32(L->bn_sums), C, M,
                                                                        act(L->dyop, (size_t)M * C).out());
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
    // weight gradient of conv i from d(conv output) in dyop and the saved im2col
    int wgrad(int i, const veto_depth_grads* g) const {
        const ConvSpec& c = kSpec[i];
        const ConvDims& d = L->d[i];
        const ActBuf dY = act(L->dyop, (size_t)d.M * c.cout);
        const ActBuf X = col(i);
        float* gp = f32(L->dwp);
        int rc;
        const bool conv_form = implicit(i);  // reduction over pixel patches of the NHWC operands, no im2col
        if (i != 0 && !conv_form && (rc = im2col(i, src_f32(i, nullptr)))) return rc;  // conv1 kept its own (its input is not passed here)
        if (prec == VETO_PREC_FP32) {
            const int64_t Mp = pad64(d.M);
            ActOut o1, o2;
            o1.f32 = f32(L->T1);
            o2.f32 = f32(L->T2);
            if ((rc = transpose_f32(dY.f32, c.cout, d.M, c.cout, false, DropSpec(), 0, o1, Mp, Mp, ActOut(), 0, s))) return rc;
            if ((rc = transpose_f32(X.f32, d.Kp, d.M, d.Kp, false, DropSpec(), 0, o2, Mp, Mp, ActOut(), 0, s))) return rc;
            GemmEpilogue ep;
            ep.ldc = d.Kp;
            ep.out.f32 = gp;
            if ((rc = gemm_simt(o1.f32, (int)Mp, o2.f32, c.cout, d.Kp, (int)Mp, ep, s))) return rc;
        } else {
            const int rows = conv_form ? gemm_tn2_conv_rows(batch, d.hout, d.wout) : (int)d.M;
            const int pairs = num_sms() / 2;
            const int max_s = (rows + 511) / 512 < kMaxSplit ? (rows + 511) / 512 : kMaxSplit;
            int best = 1;
            double best_eff = 0.0;
            for (int want = 1; want <= (max_s > 1 ? max_s : 1); ++want) {
                const int sl = gemm_tn2_slices(rows, want);
                const double eff = gemm_tn2_efficiency(c.cout, d.Kp, sl, pairs);
                if (eff > best_eff + 0.02) {
                    best_eff = eff;
                    best = sl;
                }
            }
            GemmOperand A, Bo;
            A.hi = dY.hi; A.lo = dY.lo; A.ld = c.cout;
            const size_t n = (size_t)c.cout * d.Kp;
            float* sk = f32(L->splitk);
            if (conv_form) {
                const ActBuf xin = src_op(i);
                Bo.hi = xin.hi; Bo.lo = xin.lo;
                rc = gemm_tn2_conv(A, Bo, batch, d.hout, d.wout, d.hin, d.win, c.cout, c.cin, c.k, c.pad, c.stride, passes(),
                                   best > 1 ? sk : gp, d.Kp, best, n, s);
            } else {
                Bo.hi = X.hi; Bo.lo = X.lo; Bo.ld = d.Kp;
                rc = gemm_tn2(A, Bo, c.cout, d.Kp, (int)d.M, passes(), best > 1 ? sk : gp, d.Kp, best, n, s);
            }
            if (rc) return rc;
            if (best > 1 && (rc = splitk_reduce(sk, best, n, n, gp, s))) return rc;
        }
        unpack_conv_grad_kernel<<<blocks_for((int64_t)c.cout * c.cin * c.k * c.k, 256), 256, 0, s>>>(gp, c.cout, c.cin, c.k, d.Kp,
                                                                                                  g->conv_w[i]);
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
    // input gradient of conv i: dx (NHWC [B,hin,win,cin]) (+)= col2im(dyop @ Wp)  (+ add where mask > 0)
    int dgrad(int i, float* dx, bool accumulate, const float* add, const float* mask) const {
        const ConvSpec& c = kSpec[i];
        const ConvDims& d = L->d[i];
        const ActBuf WT = wpT(i);
        if (implicit_dgrad(i)) {
            // d(input) = the stride-1 convolution of d(output) with the flipped, transposed weights
            VETO_REQUIRE(!accumulate, VETO_ERR_ARG, "depth backbone: implicit input gradient cannot accumulate");
            const ActBuf dy = act(L->dyop, (size_t)d.M * c.cout);
            GemmOperand A, Wo;
            A.hi = dy.hi; A.lo = dy.lo;
            Wo.hi = WT.hi; Wo.lo = WT.lo;
            int rc = conv_tc(A, Wo, batch, d.hout, d.wout, c.cout, c.cin, c.k, c.k - 1 - c.pad, 1, passes(), dx, c.cin, s);
            if (rc) return rc;
            if (add) {
                const int64_t n4 = (int64_t)batch * d.hin * d.win * (c.cin / 4);
                add_masked_kernel<<<blocks_for(n4, 256), 256, 0, s>>>(dx, add, mask, n4);
                VETO_LAUNCH_CHECK();
            }
            return VETO_OK;
        }
        GemmEpilogue ep;
        ep.out.f32 = f32(L->dcol);
        ep.ldc = d.Kp;
        int rc = linear(prec, act(L->dyop, (size_t)d.M * c.cout), c.cout, WRef{WT.f32, WT.hi, WT.lo}, (int)d.M, d.Kp, c.cout, ep, s);
        if (rc) return rc;
        const int64_t pixels = (int64_t)batch * d.hin * d.win;
        col2im_kernel<<<blocks_for(pixels * (c.cin / 4), 256), 256, 0, s>>>(f32(L->dcol), d.hin, d.win, c.cin, c.k, c.stride, c.pad,
                                                                           d.hout, d.wout, d.Kp, pixels, add, mask, accumulate, dx);
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
};

int check_args(int precision, int batch, int height, int width) {
    VETO_REQUIRE(precision >= VETO_PREC_FP32 && precision <= VETO_PREC_BF16, VETO_ERR_ARG, "bad precision %d", precision);
    VETO_REQUIRE(batch >= 1 && height >= 16 && width >= 16, VETO_ERR_ARG,
                 "depth backbone: needs batch >= 1 and an image of at least 16 x 16 (got %d x %d x %d)", batch, height, width);
    VETO_REQUIRE((int64_t)batch * ((height + 1) / 2) * ((width + 1) / 2) < (int64_t)1 << 31, VETO_ERR_UNSUPPORTED,
                 "depth backbone: %d x %d x %d exceeds 2^31 output pixels of conv1", batch, height, width);
    return VETO_OK;
}

}  // namespace
}  // namespace veto

using namespace veto;

#define RC(expr)                      \
    do {                              \
        if ((rc = (expr))) return rc; \
    } while (0)

extern "C" void veto_depth_backbone_out_size(int height, int width, int* out_h, int* out_w) {
    int h = height, w = width;
    h = out_dim(h, 7, 2, 3); w = out_dim(w, 7, 2, 3);
    for (int i = 0; i < 3; ++i) {  // max-pool, layer2, layer3
        h = out_dim(h, 3, 2, 1);
        w = out_dim(w, 3, 2, 1);
    }
    if (out_h) *out_h = h;
    if (out_w) *out_w = w;
}

extern "C" size_t veto_depth_backbone_workspace_bytes(int precision, int batch, int height, int width, int training) {
    if (check_args(precision, batch, height, width)) return 0;
    return depth_layout(precision, batch, height, width, training != 0).total;
}

extern "C" int veto_depth_backbone_forward(int precision, const veto_depth_weights* w, const float* depth_dev, int batch,
                                           int height, int width, int training, float momentum, float* out_dev,
                                           void* workspace_dev, size_t workspace_bytes, veto_stream_t stream) {
    int rc = check_args(precision, batch, height, width);
    if (rc) return rc;
    VETO_REQUIRE(w && depth_dev && out_dev && workspace_dev, VETO_ERR_ARG, "veto_depth_backbone_forward: NULL argument");
    const DepthLayout L = depth_layout(precision, batch, height, width, training != 0);
    VETO_REQUIRE(workspace_bytes >= L.total, VETO_ERR_ARG, "veto_depth_backbone_forward: workspace %zu < %zu bytes",
                 workspace_bytes, L.total);
    if (precision != VETO_PREC_FP32) RC(gemm_tc_init());
    Run R{precision, (cudaStream_t)stream, (char*)workspace_dev, &L, w, batch, training != 0, momentum};
    set_tag(TAG_OTHER);
    RC(R.conv(0, depth_dev));  // conv1 reads the depth image; every other convolution knows its source (DepthLayout::src)
    RC(R.bn(0, nullptr, true, nullptr));
    {
        const int64_t n = (int64_t)batch * L.hp * L.wp * (64 / 4);
        maxpool_fwd_kernel<<<blocks_for(n, 256), 256, 0, R.s>>>(R.f32(L.y[0]), L.d[0].hout, L.d[0].wout, 64, L.hp, L.wp,
                                                               (int64_t)batch * L.hp * L.wp, R.f32(L.pool),
                                                               training ? (uchar4*)(R.B + L.pool_arg) : nullptr,
                                                               precision != VETO_PREC_FP32
                                                                   ? R.act(L.pool_op, (size_t)batch * L.hp * L.wp * 64).out()
                                                                   : ActOut());
        VETO_LAUNCH_CHECK();
    }
    const float* x = R.f32(L.pool);
    for (int b = 0; b < 6; ++b) {
        const BlockSpec& bs = kBlocks[b];
        RC(R.conv(bs.c1, nullptr));
        RC(R.bn(bs.c1, nullptr, true, nullptr));
        const float* identity = x;
        if (bs.ds >= 0) {
            RC(R.conv(bs.ds, nullptr));
            RC(R.bn(bs.ds, nullptr, false, nullptr));
            identity = R.f32(L.y[bs.ds]);
        }
        RC(R.conv(bs.c2, nullptr));
        RC(R.bn(bs.c2, identity, true, b == 5 ? out_dev : nullptr));
        x = R.f32(L.y[bs.c2]);
    }
    return VETO_OK;
}

extern "C" int veto_depth_backbone_backward(int precision, const veto_depth_weights* w, const float* grad_out_dev, int batch,
                                            int height, int width, const veto_depth_grads* g, void* workspace_dev,
                                            size_t workspace_bytes, veto_stream_t stream) {
    int rc = check_args(precision, batch, height, width);
    if (rc) return rc;
    VETO_REQUIRE(w && grad_out_dev && g && workspace_dev, VETO_ERR_ARG, "veto_depth_backbone_backward: NULL argument");
    const DepthLayout L = depth_layout(precision, batch, height, width, true);
    VETO_REQUIRE(workspace_bytes >= L.total, VETO_ERR_ARG, "veto_depth_backbone_backward: workspace %zu < %zu bytes",
                 workspace_bytes, L.total);
    if (precision != VETO_PREC_FP32) RC(gemm_tc_init());
    Run R{precision, (cudaStream_t)stream, (char*)workspace_dev, &L, w, batch, true, 0.f};
    set_tag(TAG_OTHER);
    float* dX = R.f32(L.gA);   // gradient of the current block's output
    float* gB = R.f32(L.gB);
    float* gC = R.f32(L.gC);
    {
        const ConvDims& d = L.d[14];
        const int64_t n = d.M * 256;
        nchw_to_nhwc_kernel<<<blocks_for(n, 256), 256, 0, R.s>>>(grad_out_dev, 256, (int64_t)d.hout * d.wout, n, dX);
        VETO_LAUNCH_CHECK();
    }
    for (int b = 5; b >= 0; --b) {
        const BlockSpec& bs = kBlocks[b];
        const float* y_out = R.f32(L.y[bs.c2]);
        // main branch: bn2 <- relu mask of the block output
        RC(R.bn_bwd(bs.c2, dX, y_out, g));
        RC(R.wgrad(bs.c2, g));
        RC(R.dgrad(bs.c2, gB, false, nullptr, nullptr));
        RC(R.bn_bwd(bs.c1, gB, R.f32(L.y[bs.c1]), g));
        RC(R.wgrad(bs.c1, g));
        if (bs.ds < 0) {
            RC(R.dgrad(bs.c1, gC, false, dX, y_out));          // + the identity branch: dX through the block's ReLU
        } else {
            RC(R.dgrad(bs.c1, gC, false, nullptr, nullptr));
            RC(R.bn_bwd(bs.ds, dX, y_out, g));
            RC(R.wgrad(bs.ds, g));
            RC(R.dgrad(bs.ds, gC, true, nullptr, nullptr));
        }
        float* t = dX;
        dX = gC;
        gC = t;
    }
    // stem: max-pool, bn1 (ReLU mask), conv1 weight gradient (the depth image itself needs no gradient)
    {
        const ConvDims& d = L.d[0];
        const int64_t n = d.M * (64 / 4);
        maxpool_bwd_kernel<<<blocks_for(n, 256), 256, 0, R.s>>>((const uchar4*)(R.B + L.pool_arg), dX, d.hout, d.wout, 64, L.hp,
                                                               L.wp, d.M, gB);
        VETO_LAUNCH_CHECK();
    }
    RC(R.bn_bwd(0, gB, R.f32(L.y[0]), g));
    RC(R.wgrad(0, g));
    return VETO_OK;
}
