// ROIAlign gather for the relation head: replaces Pooler.forward's 5 ROIAlign launches + 4 nonzero syncs
// (pysgg/modeling/poolers.py:109-171) and _C.roi_align_forward / _backward
// (pysgg/csrc/cuda/ROIAlign_cuda.cu:65-122, 178-254) on this path.
//
// Compiled with --fmad=false: results are bit-identical to the reference CPU kernel
// (csrc/cpu/ROIAlign_cpu.cpp:17-219), which accumulates w1*v1 + w2*v2 + w3*v3 + w4*v4 sample by sample
// without fused multiply-adds and divides once by the sample count.
//
// Two kernels, both built on the same idea — stage feature-map elements in shared memory TRANSPOSED to
// [position][channel] so that lanes own CHANNELS while pooling: the tap positions / weights of a sample are then
// uniform across the lanes of a bin (broadcast 16-byte reads) and the four tap values are conflict-free rows, instead of
// 16 scattered 4-byte global loads per output element (the round-1 kernel: 1.9 % of HBM bandwidth, profiles/
// r2_hbm_kernels_before.txt):
//   * small maps (the depth map, P4, P5 of a 592x800 image: <= 2400 elements per channel): one CTA per (image, map,
//     8 channels) stages the WHOLE map once and pools every box of the image that reads this map from it — the map is
//     read from HBM/L2 exactly once per image instead of once per box (80 boxes per image share the depth map);
//   * large maps (P2, P3): one CTA per (roi, 32 channels) stages the bounding window of the roi's taps (<= ~900
//     elements for boxes below 224 px) with 4-byte cp.async copies, lanes along x (coalesced rows of the NCHW plane).
// Output tiles go back through shared memory so that global stores are contiguous rows of the [N,256,8,8] result.
// The arithmetic of a bin is unchanged (sample by sample, w1*v1 + w2*v2 + w3*v3 + w4*v4, one division): bit-identical.
#include <cuda_pipeline_primitives.h>

#include "common.cuh"

namespace veto {
namespace {

constexpr int kMaxTaps = 1024;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChanBlock = 32;                 // channels per CTA of the window kernel
constexpr int kMapChan = 8;                    // channels per CTA of the map-resident kernel
constexpr int kSmemBytes = 110 * 1024;         // two CTAs per SM (2 x (110 + 1) KB of the 228 KB)

struct RoiGeom {
    const float* src;   // [C, H, W] planes of the roi's image
    int H, W;
    float scale;
};

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Separable taps.  The sample grid of a roi is a tensor product: sample (bin row bh, iy) fixes y, (bin column bw, ix)
// fixes x (ROIAlign_cpu.cpp:37-52), validity and clamping act per axis (:54-90), and the four weights of a sample are
// products of one y factor and one x factor (w1 = hy*hx, w2 = hy*lx, w3 = ly*hx, w4 = ly*lx, :92-98) — the SAME fp32
// products whether they are formed once per sample or by every lane that uses them.  So a roi needs ph*sr axis entries
// for y and pw*sr for x (16 + 16 for the 8x8, sampling-ratio-2 pooling) instead of ph*pw*sr*sr = 256 four-tap records.
// An entry: (low, high) coordinate and the (l, h) interpolation factors; a coordinate outside the map gets l = h = 0
// (every weight it takes part in is an exact 0, as for the reference's empty samples) and low = high = -1.
struct AxisTap {
    int lo, hi;
    float l, h;
};
__device__ __forceinline__ AxisTap axis_tap(int i, float start, float bin, int sr, int extent) {
    const int b = i / sr, s = i - b * sr;
    float v = start + b * bin + (float)(s + .5f) * bin / (float)sr;
    AxisTap t{-1, -1, 0.f, 0.f};
    if (!(v < -1.0f || v > (float)extent)) {
        if (v <= 0.f) v = 0.f;
        int lo = (int)v, hi;
        if (lo >= extent - 1) { hi = lo = extent - 1; v = (float)lo; } else hi = lo + 1;
        t.lo = lo;
        t.hi = hi;
        t.l = v - (float)lo;
        t.h = 1.f - t.l;
    }
    return t;
}
// the axis entries of one roi, computed by ONE warp (lane i < ny + nx): ty[ny], tx[nx] as int4 (lo, hi, l bits, h bits).
// Returns through `range` (lane-uniform) the bounding coordinates (ymin, ymax, xmin, xmax) of the valid entries, -1 maxima
// when an axis has none.
__device__ __forceinline__ void roi_axis_taps(int4* ty, int4* tx, float4 box, float scale, int height, int width, int ph,
                                              int pw, int sr, int4& range) {
    const int lane = threadIdx.x & 31;
    const float roi_start_w = box.x * scale, roi_start_h = box.y * scale;
    const float roi_end_w = box.z * scale, roi_end_h = box.w * scale;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
    const float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
    const float bin_h = roi_height / (float)ph, bin_w = roi_width / (float)pw;
    const int ny = ph * sr, nx = pw * sr;
    int ymin = 0x7fffffff, ymax = -1, xmin = 0x7fffffff, xmax = -1;
    for (int i = lane; i < ny + nx; i += 32) {
        const bool is_y = i < ny;
        const AxisTap t = is_y ? axis_tap(i, roi_start_h, bin_h, sr, height) : axis_tap(i - ny, roi_start_w, bin_w, sr, width);
        (is_y ? ty[i] : tx[i - ny]) = make_int4(t.lo, t.hi, __float_as_int(t.l), __float_as_int(t.h));
        if (t.lo >= 0) {
            if (is_y) { ymin = min(ymin, t.lo); ymax = max(ymax, t.hi); }
            else { xmin = min(xmin, t.lo); xmax = max(xmax, t.hi); }
        }
    }
    range = make_int4(warp_min(ymin), warp_max(ymax), warp_min(xmin), warp_max(xmax));
}
// turn the coordinates of the entries into staging-buffer offsets: y -> (y - y0) * row_pitch, x -> (x - x0) * stride;
// entries outside the map point at offset 0 (their factors are 0)
__device__ __forceinline__ void axis_taps_to_offsets(int4* ty, int4* tx, int ny, int nx, int y0, int x0, int row_pitch, int stride) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < ny + nx; i += 32) {
        int4& t = i < ny ? ty[i] : tx[i - ny];
        if (t.x < 0) { t.x = 0; t.y = 0; }
        else if (i < ny) { t.x = (t.x - y0) * row_pitch; t.y = (t.y - y0) * row_pitch; }
        else { t.x = (t.x - x0) * stride; t.y = (t.y - x0) * stride; }
    }
}
// One bin of one channel from a [position][stride] staging buffer (`win` already offset by the lane's channel): samples in
// the reference's order (iy outer, ix inner), acc += w1*v1 + w2*v2 + w3*v3 + w4*v4, one division by the sample count.
template <int SR>
__device__ __forceinline__ float pool_bin(const int4* ty, const int4* tx, const float* win, int bh, int bw, int sr_rt, float count) {
    const int sr = SR > 0 ? SR : sr_rt;          // compile-time sampling ratio: both loops unroll
    float acc = 0.f;
#pragma unroll
    for (int iy = 0; iy < sr; ++iy) {
        const int4 ye = ty[bh * sr + iy];
        const float ly = __int_as_float(ye.z), hy = __int_as_float(ye.w);
#pragma unroll
        for (int ix = 0; ix < sr; ++ix) {
            const int4 xe = tx[bw * sr + ix];
            const float lx = __int_as_float(xe.z), hx = __int_as_float(xe.w);
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            const float v1 = win[ye.x + xe.x], v2 = win[ye.x + xe.y], v3 = win[ye.y + xe.x], v4 = win[ye.y + xe.y];
            acc += w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
        }
    }
    // one division by the sample count, as the reference; for the sampling ratios 1 / 2 / 4 the count is a power of two and
    // the multiplication by its reciprocal is the same correctly rounded result (an IEEE division costs ~10 instructions)
    if (SR == 1) return acc;
    if (SR == 2) return acc * 0.25f;
    if (SR == 4) return acc * 0.0625f;
    return acc / count;
}

// one sample of a roi in the reference's operation order (ROIAlign_cpu.cpp:17-112): corner coordinates (y_low, x_low,
// y_high, x_high) and the four weights; a sample outside the map gets y_low = -1 and zero weights
__device__ __forceinline__ void sample_taps(int k, float x1, float y1, float x2, float y2, float scale, int height, int width,
                                            int ph, int pw, int sr, int4& p, float4& w) {
    const float roi_start_w = x1 * scale, roi_start_h = y1 * scale;
    const float roi_end_w = x2 * scale, roi_end_h = y2 * scale;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
    const float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
    const float bin_h = roi_height / (float)ph, bin_w = roi_width / (float)pw;
    const int spp = sr * sr;
    const int bin = k / spp, s = k - bin * spp;
    const int bh = bin / pw, bw = bin - bh * pw;
    const int iy = s / sr, ix = s - iy * sr;
    float y = roi_start_h + bh * bin_h + (float)(iy + .5f) * bin_h / (float)sr;
    float x = roi_start_w + bw * bin_w + (float)(ix + .5f) * bin_w / (float)sr;
    p = make_int4(-1, 0, 0, 0);
    w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(y < -1.0f || y > (float)height || x < -1.0f || x > (float)width)) {
        if (y <= 0.f) y = 0.f;
        if (x <= 0.f) x = 0.f;
        int y_low = (int)y, x_low = (int)x, y_high, x_high;
        if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else y_high = y_low + 1;
        if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else x_high = x_low + 1;
        const float ly = y - (float)y_low, lx = x - (float)x_low;
        const float hy = 1.f - ly, hx = 1.f - lx;
        p = make_int4(y_low, x_low, y_high, x_high);
        w = make_float4(hy * hx, hy * lx, ly * hx, ly * lx);
    }
}

// taps of all bins of one roi into shared memory + the bounding window of the valid ones (bounds: ymin, ymax, xmin, xmax;
// initialised by the caller to (INT_MAX, -1, INT_MAX, -1); one shared-memory atomic per warp and bound)
__device__ __forceinline__ void compute_taps(int4* tp, float4* tw, int* bounds, float x1, float y1, float x2, float y2,
                                             float scale, int height, int width, int ph, int pw, int sr) {
    const int ntap = ph * pw * sr * sr;
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < ntap; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int4 p = make_int4(-1, 0, 0, 0);
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < ntap) {
            sample_taps(k, x1, y1, x2, y2, scale, height, width, ph, pw, sr, p, w);
            tp[k] = p;
            tw[k] = w;
        }
        if (bounds) {
            const bool ok = p.x >= 0;
            const int ymin = warp_min(ok ? p.x : 0x7fffffff), ymax = warp_max(ok ? p.z : -1);
            const int xmin = warp_min(ok ? p.y : 0x7fffffff), xmax = warp_max(ok ? p.w : -1);
            if (lane == 0 && ymax >= 0) {
                atomicMin(&bounds[0], ymin);
                atomicMax(&bounds[1], ymax);
                atomicMin(&bounds[2], xmin);
                atomicMax(&bounds[3], xmax);
            }
        }
    }
}

// One (roi, channel block) tile of the window kernel: channels [c0, c0 + nch) of the roi pooled into dst[(c0 + c) * bins + bin].
template <int POOL, int SR>
__device__ __forceinline__ void gather_tile(const RoiGeom& g, float4 box, int ph_rt, int pw_rt, int sr_rt, int c0, int nch,
                                            float* __restrict__ dst, uint8_t* smem) {
    const int ph = POOL > 0 ? POOL : ph_rt, pw = POOL > 0 ? POOL : pw_rt, sr = SR > 0 ? SR : sr_rt;   // 8 x 8, ratio 2: compile time
    const int bins = ph * pw, ny = ph * sr, nx = pw * sr;
    int4* ty = reinterpret_cast<int4*>(smem);
    int4* tx = ty + ny;
    float* tile = reinterpret_cast<float*>(tx + nx);                         // [kChanBlock][bins + 1]
    float* win = tile + kChanBlock * (bins + 1);
    const int win_floats = (kSmemBytes - (int)((uint8_t*)win - smem)) / (int)sizeof(float);
    __shared__ int4 s_range;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        int4 range;
        roi_axis_taps(ty, tx, box, g.scale, g.H, g.W, ph, pw, sr, range);
        if (lane == 0) s_range = range;
    }
    __syncthreads();
    const int4 range = s_range;
    const float count = (float)(sr * sr);
    const size_t plane = (size_t)g.H * g.W;
    if (range.y < 0 || range.w < 0) {          // every sample outside the map: zeros (0 / count)
        for (int e = tid; e < nch * bins; e += kThreads) dst[(size_t)c0 * bins + e] = 0.f;
        return;
    }
    const int y0 = range.x, x0 = range.z;
    const int wh = range.y - y0 + 1, ww = range.w - x0 + 1;
    const int nw = wh * ww;
    int cb = kChanBlock;
    while (cb > 1 && nw * (cb + 1) > win_floats) cb >>= 1;
    if (nw * (cb + 1) > win_floats) {
        // window larger than shared memory even for one channel (maps far beyond the 592x800 geometry): taps from global
        if (warp == 0) axis_taps_to_offsets(ty, tx, ny, nx, 0, 0, g.W, 1);
        __syncthreads();
        for (int e = tid; e < nch * bins; e += kThreads) {
            const int cl = e / bins, bin = e - cl * bins;
            dst[(size_t)(c0 + cl) * bins + bin] = pool_bin<0>(ty, tx, g.src + (size_t)(c0 + cl) * plane, bin / pw, bin % pw, sr, count);
        }
        return;
    }
    const int stride = cb + 1;
    if (warp == 0) axis_taps_to_offsets(ty, tx, ny, nx, y0, x0, ww * stride, stride);
    const int bins_per_iter = 32 / cb;           // a warp covers cb channels x (32 / cb) bins per step
    const int cl = lane % cb, bsub = lane / cb;
    for (int sub = 0; sub * cb < nch; ++sub) {
        const int nc = min(cb, nch - sub * cb);
        const float* src0 = g.src + (size_t)(c0 + sub * cb) * plane + (size_t)y0 * g.W + x0;
        // ---- stage the window: a thread owns window positions (consecutive lanes = consecutive x of a row: coalesced)
        //      and copies them for every channel of the block — two adds per 4-byte cp.async
        for (int pos = tid; pos < nw; pos += kThreads) {
            const int yy = pos / ww, xx = pos - yy * ww;
            const float* sp = src0 + (size_t)yy * g.W + xx;
            float* dp = win + (size_t)pos * stride;
            for (int c = 0; c < nc; ++c) __pipeline_memcpy_async(dp + c, sp + (size_t)c * plane, sizeof(float));
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncthreads();
        // ---- pool: lane = (bin slot, channel); the axis entries of a bin are uniform across the lanes of a bin slot
        for (int b0 = warp * bins_per_iter; b0 < bins; b0 += kWarps * bins_per_iter) {
            const int bin = b0 + bsub;
            if (bin < bins && cl < nc)
                tile[(sub * cb + cl) * (bins + 1) + bin] = pool_bin<SR>(ty, tx, win + cl, bin / pw, bin % pw, sr, count);
        }
        __syncthreads();
    }
    // ---- contiguous stores of the [nch][bins] tile
    for (int e = tid; e < nch * bins; e += kThreads) {
        const int c = e / bins, b = e - c * bins;
        dst[(size_t)c0 * bins + e] = tile[c * (bins + 1) + b];
    }
}

// grid (n_rois, channel blocks)
__global__ void __launch_bounds__(kThreads, 2)
roi_align_fwd_kernel(const float* __restrict__ input, int channels, int height, int width,
                     const float* __restrict__ rois, float scale, int ph, int pw, int sr, float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int n = blockIdx.x;
    const float* r = rois + (size_t)n * 5;
    const int b = (int)r[0];
    RoiGeom g{input + (size_t)b * channels * height * width, height, width, scale};
    const int c0 = blockIdx.y * kChanBlock;
    gather_tile<0, 0>(g, make_float4(r[1], r[2], r[3], r[4]), ph, pw, sr, c0, min(kChanBlock, channels - c0),
                      out + (size_t)n * channels * ph * pw, smem_raw);
}

struct GatherLevels {
    const float* feat[4];
    int h[4], w[4];
    float scale[4];
    int n_levels, k_min, k_max;
    int resident[4];   // 1: the level's map is pooled by the map-resident kernel, 0: by the window kernel
};

// LevelMapper (poolers.py:32-43) in fp32: floor(4 + log2(sqrt(area)/224 + 1e-6)) clamped to [k_min,k_max].
// log2 is evaluated in double and rounded once, i.e. the correctly rounded fp32 log2 of the fp32 argument.
__device__ __forceinline__ int map_level(float x1, float y1, float x2, float y2, int k_min, int k_max) {
    const float area = (x2 - x1 + 1.f) * (y2 - y1 + 1.f);
    const float s = sqrtf(area);
    const float arg = s / 224.f + 1e-6f;
    const float lg = (float)log2((double)arg);
    float lvl = floorf(4.f + lg);
    lvl = fminf(fmaxf(lvl, (float)k_min), (float)k_max);
    return (int)lvl - k_min;
}

// FPN level of every box, once (the double-precision log2 of map_level is ~300 instructions on this chip)
__global__ void roi_levels_kernel(const float* __restrict__ boxes, int n_boxes, int k_min, int k_max, int32_t* __restrict__ levels) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_boxes) return;
    const float4 bx = __ldg((const float4*)boxes + n);
    levels[n] = map_level(bx.x, bx.y, bx.z, bx.w, k_min, k_max);
}

// Window kernel, grid (N, channel blocks): the RGB features of the boxes whose FPN level is not map-resident
// (and, when the depth map is too large for the resident kernel, grid.z = 2: z = 1 pools the depth features).
template <int POOL, int SR>
__global__ void __launch_bounds__(kThreads, 2)
roi_gather_kernel(GatherLevels lv, const float* __restrict__ depth, int depth_h, int depth_w, float depth_scale,
                  int channels, const float* __restrict__ boxes, const int32_t* __restrict__ box_off, int n_images,
                  int pool, int sr, float* __restrict__ out_rgb, float* __restrict__ out_depth,
                  const int32_t* __restrict__ levels_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int n = blockIdx.x;
    const float4 bx = __ldg((const float4*)boxes + n);
    RoiGeom g;
    float* dst;
    int l = -1;
    if (blockIdx.z == 0) {
        // the level comes from roi_levels_kernel when the caller gave a levels buffer (one double-precision log2 per box
        // instead of one per thread of every CTA of the box)
        l = levels_out ? levels_out[n] : map_level(bx.x, bx.y, bx.z, bx.w, lv.k_min, lv.k_max);
        if (lv.resident[l]) return;
    }
    // image of box n (poolers.py:96-107: roi batch index)
    int lo = 0, hi = n_images - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (box_off[mid] <= n) lo = mid; else hi = mid - 1;
    }
    const int b = lo;
    if (blockIdx.z == 0) {
        g = RoiGeom{lv.feat[l] + (size_t)b * channels * lv.h[l] * lv.w[l], lv.h[l], lv.w[l], lv.scale[l]};
        dst = out_rgb + (size_t)n * channels * pool * pool;
    } else {
        g = RoiGeom{depth + (size_t)b * channels * depth_h * depth_w, depth_h, depth_w, depth_scale};
        dst = out_depth + (size_t)n * channels * pool * pool;
    }
    const int c0 = blockIdx.y * kChanBlock;
    gather_tile<POOL, SR>(g, bx, pool, pool, sr, c0, min(kChanBlock, channels - c0), dst, smem_raw);
}

// Map-resident kernel, grid (n_images, maps, channels / 8): map 0 = the depth map, map 1 + l = FPN level l (resident ones).
// The CTA stages its 8 channels of the whole map as [position][9]; after that every WARP works on its own: it takes the
// image's boxes round-robin, computes the 16 + 16 axis entries of a box (one per lane), pools its 8 channels x 64 bins
// (lane = (bin slot, channel)) into a private tile and stores the tile as contiguous rows — no block barrier per box.
// 16 warps per CTA and compact per-warp areas (16 + 16 axis entries = 512 B, an [8][65] tile = 2080 B): with the 68 KB
// map of a 38 x 50 level two CTAs = 32 warps share an SM.  The first version (8 warps, 2 CTAs) was latency-bound: 62 %
// issue-active at 23 % occupancy (profiles/r2_gather_kernels.txt).
constexpr int kResThreads = 512;
constexpr int kResWarps = kResThreads / 32;
template <int POOL, int SR>
__global__ void __launch_bounds__(kResThreads, 2)
roi_gather_resident_kernel(GatherLevels lv, const float* __restrict__ depth, int depth_h, int depth_w, float depth_scale,
                           int depth_resident, int channels, const float* __restrict__ boxes,
                           const int32_t* __restrict__ box_off, int pool_rt, int sr_rt, float* __restrict__ out_rgb,
                           float* __restrict__ out_depth, const int32_t* __restrict__ levels) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int pool = POOL > 0 ? POOL : pool_rt, sr = SR > 0 ? SR : sr_rt;
    const int img = blockIdx.x, m = blockIdx.y;
    const int l = m - 1;
    if (m == 0 ? !depth_resident : !lv.resident[l]) return;
    const int bins = pool * pool, ny = pool * sr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_tap_bytes = 2 * ny * (int)sizeof(int4);
    int4* ty = reinterpret_cast<int4*>(smem + (size_t)warp * warp_tap_bytes);
    int4* tx = ty + ny;
    float* tile = reinterpret_cast<float*>(smem + (size_t)kResWarps * warp_tap_bytes) + (size_t)warp * kMapChan * (bins + 1);
    float* map = reinterpret_cast<float*>(smem + (size_t)kResWarps * warp_tap_bytes) + (size_t)kResWarps * kMapChan * (bins + 1);
    const int H = m == 0 ? depth_h : lv.h[l], W = m == 0 ? depth_w : lv.w[l];
    const float scale = m == 0 ? depth_scale : lv.scale[l];
    const int c0 = blockIdx.z * kMapChan, nc = min(kMapChan, channels - c0);
    const int n0 = box_off[img], n1 = box_off[img + 1];
    // does any box of the image read this map?  (depth: all of them)
    if (m > 0) {
        int any = 0;
        for (int n = n0 + tid; n < n1; n += kResThreads) {
            if (levels) {
                any |= levels[n] == l;
            } else {
                const float4 bx = __ldg((const float4*)boxes + n);
                any |= map_level(bx.x, bx.y, bx.z, bx.w, lv.k_min, lv.k_max) == l;
            }
        }
        if (!__syncthreads_or(any)) return;
    } else if (n1 <= n0) {
        return;
    }
    const float* src = (m == 0 ? depth : lv.feat[l]) + ((size_t)img * channels + c0) * H * W;
    constexpr int stride = kMapChan + 1;
    const int hw = H * W;
    // the nc planes are one contiguous range of nc * hw floats: consecutive threads copy consecutive elements
    for (int e = tid; e < nc * hw; e += kResThreads) {
        const int c = e / hw, pos = e - c * hw;
        __pipeline_memcpy_async(map + (size_t)pos * stride + c, src + e, sizeof(float));
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    float* out = m == 0 ? out_depth : out_rgb;
    const float count = (float)(sr * sr);
    constexpr int kBinsPerIter = 32 / kMapChan;
    const int cl = lane % kMapChan, bsub = lane / kMapChan;
    int slot = 0;                                          // boxes of this map are dealt to the warps round-robin
    for (int n = n0; n < n1; ++n) {
        if (m > 0 && levels && levels[n] != l) continue;
        const float4 bx = __ldg((const float4*)boxes + n);
        if (m > 0 && !levels && map_level(bx.x, bx.y, bx.z, bx.w, lv.k_min, lv.k_max) != l) continue;
        if ((slot++ % kResWarps) != warp) continue;
        int4 range;
        roi_axis_taps(ty, tx, bx, scale, H, W, pool, pool, sr, range);
        __syncwarp();
        axis_taps_to_offsets(ty, tx, ny, ny, 0, 0, W * stride, stride);
        __syncwarp();
        for (int b0 = 0; b0 < bins; b0 += kBinsPerIter) {
            const int bin = b0 + bsub;
            if (cl < nc) tile[cl * (bins + 1) + bin] = pool_bin<SR>(ty, tx, map + cl, bin / pool, bin % pool, sr, count);
        }
        __syncwarp();
        float* dst = out + ((size_t)n * channels + c0) * bins;
        for (int e = lane; e < nc * bins; e += 32) {
            const int c = e / bins, b = e - c * bins;
            dst[e] = tile[c * (bins + 1) + b];
        }
        __syncwarp();                                      // the tile and the entries are reused for the warp's next box
    }
}

// backward (replaces RoIAlignBackwardFeature, ROIAlign_cuda.cu:178-254: one thread per output element, 16 global fp32
// atomics each): the same (roi, 32-channel block) tiling as the forward.  The roi's gradient tile is read with contiguous
// loads, scattered into a shared-memory copy of the window (lanes own channels, so the lanes of a warp never collide;
// warps working on neighbouring bins meet through shared-memory atomics), and the window is then added to the map once
// per element with row-contiguous global atomics — rois overlap, so the map itself still needs them.
__global__ void __launch_bounds__(kThreads, 2)
roi_align_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ rois, float scale, int channels, int height,
                     int width, int ph, int pw, int sr, float* __restrict__ grad_in) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = blockIdx.x;
    const float* r = rois + (size_t)n * 5;
    const int b = (int)r[0];
    const int c0 = blockIdx.y * kChanBlock;
    const int nch = min(kChanBlock, channels - c0);
    const int bins = ph * pw, spp = sr * sr, ntap = bins * spp;
    int4* tp = reinterpret_cast<int4*>(smem);
    float4* tw = reinterpret_cast<float4*>(smem + (size_t)ntap * sizeof(int4));
    float* tile = reinterpret_cast<float*>(smem + (size_t)ntap * (sizeof(int4) + sizeof(float4)));
    float* win = tile + kChanBlock * (bins + 1);
    const int win_floats = (kSmemBytes - (int)((uint8_t*)win - smem)) / (int)sizeof(float);
    __shared__ int bounds[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        bounds[0] = 0x7fffffff; bounds[1] = -1; bounds[2] = 0x7fffffff; bounds[3] = -1;
    }
    __syncthreads();
    compute_taps(tp, tw, bounds, r[1], r[2], r[3], r[4], scale, height, width, ph, pw, sr);
    for (int e = tid; e < nch * bins; e += kThreads) {
        const int c = e / bins, bb = e - c * bins;
        tile[c * (bins + 1) + bb] = grad[((size_t)n * channels + c0) * bins + e];
    }
    __syncthreads();
    if (bounds[1] < 0) return;                  // no sample inside the map: no gradient
    const int y0 = bounds[0], x0 = bounds[2];
    const int wh = bounds[1] - y0 + 1, ww = bounds[3] - x0 + 1;
    const int nw = wh * ww;
    const float count = (float)spp;
    const size_t plane = (size_t)height * width;
    float* dst_img = grad_in + ((size_t)b * channels + c0) * plane;
    int cb = kChanBlock;
    while (cb > 1 && nw * (cb + 1) > win_floats) cb >>= 1;
    if (nw * (cb + 1) > win_floats) {           // window beyond shared memory: the element-wise scatter
        for (int e = tid; e < nch * bins; e += kThreads) {
            const int c = e / bins, bin = e - c * bins;
            float* dst = dst_img + (size_t)c * plane;
            const float go = tile[c * (bins + 1) + bin];
            for (int s = 0; s < spp; ++s) {
                const int4 p = tp[bin * spp + s];
                const float4 w = tw[bin * spp + s];
                if (p.x < 0) continue;
                atomicAdd(dst + p.x * width + p.y, go * w.x / count);
                atomicAdd(dst + p.x * width + p.w, go * w.y / count);
                atomicAdd(dst + p.z * width + p.y, go * w.z / count);
                atomicAdd(dst + p.z * width + p.w, go * w.w / count);
            }
        }
        return;
    }
    const int stride = cb + 1;
    for (int k = tid; k < ntap; k += kThreads) {
        const int4 p = tp[k];
        tp[k] = p.x < 0 ? make_int4(-1, 0, 0, 0)
                        : make_int4(((p.x - y0) * ww + (p.y - x0)) * stride, ((p.x - y0) * ww + (p.w - x0)) * stride,
                                    ((p.z - y0) * ww + (p.y - x0)) * stride, ((p.z - y0) * ww + (p.w - x0)) * stride);
    }
    const int bins_per_iter = 32 / cb;
    const int cl = lane % cb, bsub = lane / cb;
    for (int sub = 0; sub * cb < nch; ++sub) {
        const int nc = min(cb, nch - sub * cb);
        for (int e = tid; e < nw * stride; e += kThreads) win[e] = 0.f;
        __syncthreads();
        for (int b0 = warp * bins_per_iter; b0 < bins; b0 += kWarps * bins_per_iter) {
            const int bin = b0 + bsub;
            if (bin < bins && cl < nc) {
                const float go = tile[(sub * cb + cl) * (bins + 1) + bin];
                for (int s = 0; s < spp; ++s) {
                    const int4 p = tp[bin * spp + s];
                    if (p.x < 0) continue;
                    const float4 w = tw[bin * spp + s];
                    atomicAdd(&win[p.x + cl], go * w.x / count);
                    atomicAdd(&win[p.y + cl], go * w.y / count);
                    atomicAdd(&win[p.z + cl], go * w.z / count);
                    atomicAdd(&win[p.w + cl], go * w.w / count);
                }
            }
        }
        __syncthreads();
        for (int rr = warp; rr < nc * wh; rr += kWarps) {
            const int c = rr / wh, yy = rr - c * wh;
            float* drow = dst_img + (size_t)(sub * cb + c) * plane + (size_t)(y0 + yy) * width + x0;
            const float* srow = win + (size_t)yy * ww * stride + c;
            for (int x = lane; x < ww; x += 32) {
                const float v = srow[x * stride];
                if (v != 0.f) atomicAdd(drow + x, v);
            }
        }
        __syncthreads();
    }
}

// Map-resident backward, grid (batch, channels / 8): the CTA owns 8 channels of one image's gradient map in shared memory
// ([position][9], zero-initialised), scatters the gradient tiles of the image's rois into it (roi index ascending; lanes
// own channels, warps working on neighbouring bins meet through shared-memory atomics) and writes the map out ONCE with
// plain coalesced stores: no global atomics and no separate memset — every element of grad_in is written exactly once.
__global__ void __launch_bounds__(kThreads, 2)
roi_align_bwd_resident_kernel(const float* __restrict__ grad, const float* __restrict__ rois, int n_rois, float scale,
                              int channels, int height, int width, int ph, int pw, int sr, float* __restrict__ grad_in) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int img = blockIdx.x;
    const int c0 = blockIdx.y * kMapChan, nc = min(kMapChan, channels - c0);
    const int bins = ph * pw, spp = sr * sr, ntap = bins * spp;
    int4* tp = reinterpret_cast<int4*>(smem);
    float4* tw = reinterpret_cast<float4*>(smem + (size_t)ntap * sizeof(int4));
    float* tile = reinterpret_cast<float*>(smem + (size_t)ntap * (sizeof(int4) + sizeof(float4)));   // [kMapChan][bins + 1]
    float* map = tile + kMapChan * (bins + 1);
    constexpr int stride = kMapChan + 1;
    const int hw = height * width;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < hw * stride; e += kThreads) map[e] = 0.f;
    const float count = (float)spp;
    const int bins_per_iter = 32 / kMapChan;
    const int cl = lane % kMapChan, bsub = lane / kMapChan;
    for (int n = 0; n < n_rois; ++n) {
        const float* r = rois + (size_t)n * 5;
        if ((int)r[0] != img) continue;                    // uniform across the CTA
        __syncthreads();                                   // the previous roi's taps / tile are consumed (first: map zeroed)
        for (int k = tid; k < ntap; k += kThreads) {
            int4 p;
            float4 w;
            sample_taps(k, r[1], r[2], r[3], r[4], scale, height, width, ph, pw, sr, p, w);
            tp[k] = p.x < 0 ? make_int4(-1, 0, 0, 0)
                            : make_int4((p.x * width + p.y) * stride, (p.x * width + p.w) * stride,
                                        (p.z * width + p.y) * stride, (p.z * width + p.w) * stride);
            tw[k] = w;
        }
        const float* gsrc = grad + ((size_t)n * channels + c0) * bins;
        for (int e = tid; e < nc * bins; e += kThreads) {
            const int c = e / bins, b = e - c * bins;
            tile[c * (bins + 1) + b] = gsrc[e];
        }
        __syncthreads();
        for (int b0 = warp * bins_per_iter; b0 < bins; b0 += kWarps * bins_per_iter) {
            const int bin = b0 + bsub;
            if (bin < bins && cl < nc) {
                const float go = tile[cl * (bins + 1) + bin];
                for (int s = 0; s < spp; ++s) {
                    const int4 p = tp[bin * spp + s];
                    if (p.x < 0) continue;
                    const float4 w = tw[bin * spp + s];
                    atomicAdd(&map[p.x + cl], go * w.x / count);
                    atomicAdd(&map[p.y + cl], go * w.y / count);
                    atomicAdd(&map[p.z + cl], go * w.z / count);
                    atomicAdd(&map[p.w + cl], go * w.w / count);
                }
            }
        }
    }
    __syncthreads();
    float* dst = grad_in + ((size_t)img * channels + c0) * hw;
    for (int e = tid; e < nc * hw; e += kThreads) {
        const int c = e / hw, pos = e - c * hw;
        dst[e] = map[(size_t)pos * stride + c];
    }
}

// floats the resident kernels need next to the map: taps + one [8][bins + 1] tile
__host__ inline bool map_fits_resident(int h, int w, int pool, int sr) {
    // forward: per-warp axis entries + per-warp [8][bins + 1] tiles; backward: the 2-D taps + one tile — the larger of the two
    const size_t fwd = (size_t)kResWarps * 2 * pool * sr * sizeof(int4) + (size_t)kResWarps * kMapChan * (pool * pool + 1) * sizeof(float);
    const size_t bwd = (size_t)pool * pool * sr * sr * (sizeof(int4) + sizeof(float4)) + (size_t)kMapChan * (pool * pool + 1) * sizeof(float);
    const size_t fixed = fwd > bwd ? fwd : bwd;
    return pool * sr <= 64 && fixed + (size_t)h * w * (kMapChan + 1) * sizeof(float) <= (size_t)kSmemBytes;
}

DeviceOnce g_attr_set;
int ensure_attrs() {
    if (!g_attr_set.pending()) return VETO_OK;
    VETO_CUDA(cudaFuncSetAttribute(roi_align_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(roi_gather_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(roi_gather_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(roi_align_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(roi_gather_resident_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(roi_gather_resident_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(roi_align_bwd_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    g_attr_set.done();
    return VETO_OK;
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_roi_align_forward(const float* input_dev, int batch, int channels, int height, int width,
                                      const float* rois_dev, int n_rois, float spatial_scale, int pooled_h, int pooled_w,
                                      int sampling_ratio, float* out_dev, veto_stream_t stream) {
    if (n_rois == 0) return VETO_OK;
    VETO_REQUIRE(input_dev && rois_dev && out_dev && batch > 0 && channels > 0 && height > 0 && width > 0 && n_rois > 0,
                 VETO_ERR_ARG, "veto_roi_align_forward: bad argument");
    VETO_REQUIRE(sampling_ratio > 0 && pooled_h > 0 && pooled_w > 0 &&
                     pooled_h * pooled_w * sampling_ratio * sampling_ratio <= kMaxTaps,
                 VETO_ERR_UNSUPPORTED, "veto_roi_align_forward: needs sampling_ratio > 0 and ph*pw*sr^2 <= %d", kMaxTaps);
    int rc = ensure_attrs();
    if (rc) return rc;
    set_tag(TAG_GATHER);
    dim3 grid(n_rois, (channels + kChanBlock - 1) / kChanBlock);
    roi_align_fwd_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        input_dev, channels, height, width, rois_dev, spatial_scale, pooled_h, pooled_w, sampling_ratio, out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

extern "C" int veto_roi_align_backward(const float* grad_dev, const float* rois_dev, int n_rois, float spatial_scale,
                                       int pooled_h, int pooled_w, int batch, int channels, int height, int width,
                                       int sampling_ratio, float* grad_input_dev, veto_stream_t stream) {
    VETO_REQUIRE(grad_input_dev && batch > 0 && channels > 0 && height > 0 && width > 0 && sampling_ratio > 0, VETO_ERR_ARG,
                 "veto_roi_align_backward: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    VETO_REQUIRE(n_rois == 0 || (grad_dev && rois_dev), VETO_ERR_ARG, "veto_roi_align_backward: bad argument");
    VETO_REQUIRE(pooled_h > 0 && pooled_w > 0 && pooled_h * pooled_w * sampling_ratio * sampling_ratio <= kMaxTaps,
                 VETO_ERR_UNSUPPORTED, "veto_roi_align_backward: needs ph*pw*sr^2 <= %d", kMaxTaps);
    int rc = ensure_attrs();
    if (rc) return rc;
    set_tag(TAG_GATHER);
    if (pooled_h == pooled_w && map_fits_resident(height, width, pooled_h, sampling_ratio)) {
        // the map of one image (8 channels) fits shared memory: no global atomics, no memset
        dim3 rgrid(batch, (channels + kMapChan - 1) / kMapChan);
        roi_align_bwd_resident_kernel<<<rgrid, kThreads, kSmemBytes, s>>>(grad_dev, rois_dev, n_rois, spatial_scale, channels,
                                                                         height, width, pooled_h, pooled_w, sampling_ratio,
                                                                         grad_input_dev);
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
    VETO_CUDA(cudaMemsetAsync(grad_input_dev, 0, (size_t)batch * channels * height * width * sizeof(float), s));
    if (n_rois == 0) return VETO_OK;
    dim3 grid(n_rois, (channels + kChanBlock - 1) / kChanBlock);
    roi_align_bwd_kernel<<<grid, kThreads, kSmemBytes, s>>>(grad_dev, rois_dev, spatial_scale, channels, height, width, pooled_h,
                                                           pooled_w, sampling_ratio, grad_input_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

extern "C" int veto_roi_gather_forward(const float* const* feats_dev, const int32_t* feat_h, const int32_t* feat_w,
                                       const float* scales, int n_levels, int k_min, int k_max, const float* depth_dev,
                                       int depth_h, int depth_w, float depth_scale, int batch, int channels,
                                       const float* boxes_dev, const int32_t* box_offsets_dev, int n_images, int n_boxes,
                                       int pool, int sampling_ratio, float* out_rgb_dev, float* out_depth_dev,
                                       int32_t* levels_out_dev, veto_stream_t stream) {
    if (n_boxes == 0) return VETO_OK;
    VETO_REQUIRE(feats_dev && feat_h && feat_w && scales && depth_dev && boxes_dev && box_offsets_dev && out_rgb_dev &&
                     out_depth_dev && n_images > 0 && n_images <= batch && channels > 0,
                 VETO_ERR_ARG, "veto_roi_gather_forward: bad argument");
    VETO_REQUIRE(n_levels >= 1 && n_levels <= 4 && k_max - k_min + 1 == n_levels, VETO_ERR_UNSUPPORTED,
                 "veto_roi_gather_forward: 1..4 FPN levels with k_max-k_min+1 == n_levels");
    VETO_REQUIRE(sampling_ratio > 0 && pool > 0 && pool * pool * sampling_ratio * sampling_ratio <= kMaxTaps,
                 VETO_ERR_UNSUPPORTED, "veto_roi_gather_forward: needs sampling_ratio > 0 and pool^2*sr^2 <= %d", kMaxTaps);
    int rc = ensure_attrs();
    if (rc) return rc;
    set_tag(TAG_GATHER);
    GatherLevels lv{};
    for (int l = 0; l < n_levels; ++l) {
        VETO_REQUIRE(feats_dev[l] && feat_h[l] > 0 && feat_w[l] > 0, VETO_ERR_ARG, "veto_roi_gather_forward: bad level %d", l);
        lv.feat[l] = feats_dev[l];
        lv.h[l] = feat_h[l];
        lv.w[l] = feat_w[l];
        lv.scale[l] = scales[l];
    }
    lv.n_levels = n_levels;
    lv.k_min = k_min;
    lv.k_max = k_max;
    // small maps are pooled by the map-resident kernel (one staging per image), the others per roi by the window kernel
    bool any_resident = false, all_resident = true;
    for (int l = 0; l < n_levels; ++l) {
        lv.resident[l] = map_fits_resident(feat_h[l], feat_w[l], pool, sampling_ratio) ? 1 : 0;
        any_resident |= lv.resident[l] != 0;
        all_resident &= lv.resident[l] != 0;
    }
    const int depth_resident = map_fits_resident(depth_h, depth_w, pool, sampling_ratio) ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (levels_out_dev) {
        roi_levels_kernel<<<(n_boxes + 127) / 128, 128, 0, s>>>(boxes_dev, n_boxes, k_min, k_max, levels_out_dev);
        VETO_LAUNCH_CHECK();
    }
    if (any_resident || depth_resident) {
        dim3 rgrid(n_images, 1 + n_levels, (channels + kMapChan - 1) / kMapChan);
        // the reference's pooling (8 x 8 bins, sampling ratio 2) has its loop bounds at compile time
        auto* kern = (pool == 8 && sampling_ratio == 2) ? roi_gather_resident_kernel<8, 2> : roi_gather_resident_kernel<0, 0>;
        kern<<<rgrid, kResThreads, kSmemBytes, s>>>(lv, depth_dev, depth_h, depth_w, depth_scale, depth_resident, channels, boxes_dev,
                                                box_offsets_dev, pool, sampling_ratio, out_rgb_dev, out_depth_dev, levels_out_dev);
        VETO_LAUNCH_CHECK();
    }
    if (!all_resident || !depth_resident) {
        dim3 grid(n_boxes, (channels + kChanBlock - 1) / kChanBlock, depth_resident ? 1 : 2);
        auto* kern = (pool == 8 && sampling_ratio == 2) ? roi_gather_kernel<8, 2> : roi_gather_kernel<0, 0>;
        kern<<<grid, kThreads, kSmemBytes, s>>>(lv, depth_dev, depth_h, depth_w, depth_scale, channels, boxes_dev, box_offsets_dev,
                                               n_images, pool, sampling_ratio, out_rgb_dev, out_depth_dev, levels_out_dev);
        VETO_LAUNCH_CHECK();
    }
    return VETO_OK;
}
