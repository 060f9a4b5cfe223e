// ROIAlign gather for the relation head: replaces Pooler.forward's 5 ROIAlign launches + 4 nonzero syncs
// (pysgg/modeling/poolers.py:109-171) and _C.roi_align_forward / _backward
// (pysgg/csrc/cuda/ROIAlign_cuda.cu:65-122, 178-254) on this path.
//
// Compiled with --fmad=false: results are bit-identical to the reference CPU kernel
// (csrc/cpu/ROIAlign_cpu.cpp:17-219), which accumulates w1*v1 + w2*v2 + w3*v3 + w4*v4 sample by sample
// without fused multiply-adds and divides once by the sample count.
//
// One CTA handles one (roi, map, channel slice).  The ph*pw*sr*sr bilinear taps of the roi (the same for
// every channel) are computed once into shared memory; the channel loop then streams the feature planes:
// consecutive threads own consecutive output bins of one channel, so stores are fully coalesced and the
// 16 taps of neighbouring bins hit the same L1 lines of the NCHW plane.
#include "common.cuh"

namespace veto {
namespace {

constexpr int kMaxTaps = 1024;
constexpr int kThreads = 256;

struct Taps {
    int p[4][kMaxTaps];
    float w[4][kMaxTaps];
};

// taps of all bins of one roi; reference operation order (ROIAlign_cpu.cpp:17-112)
__device__ __forceinline__ void compute_taps(Taps& t, float x1, float y1, float x2, float y2, float scale, int height,
                                             int width, int ph, int pw, int sr) {
    const float roi_start_w = x1 * scale, roi_start_h = y1 * scale;
    const float roi_end_w = x2 * scale, roi_end_h = y2 * scale;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
    const float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
    const float bin_h = roi_height / (float)ph, bin_w = roi_width / (float)pw;
    const int spp = sr * sr, ntap = ph * pw * spp;
    for (int k = threadIdx.x; k < ntap; k += blockDim.x) {
        const int bin = k / spp, s = k - bin * spp;
        const int bh = bin / pw, bw = bin - bh * pw;
        const int iy = s / sr, ix = s - iy * sr;
        float y = roi_start_h + bh * bin_h + (float)(iy + .5f) * bin_h / (float)sr;
        float x = roi_start_w + bw * bin_w + (float)(ix + .5f) * bin_w / (float)sr;
        int p1 = 0, p2 = 0, p3 = 0, p4 = 0;
        float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
        if (!(y < -1.0f || y > (float)height || x < -1.0f || x > (float)width)) {
            if (y <= 0.f) y = 0.f;
            if (x <= 0.f) x = 0.f;
            int y_low = (int)y, x_low = (int)x, y_high, x_high;
            if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else y_high = y_low + 1;
            if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else x_high = x_low + 1;
            const float ly = y - (float)y_low, lx = x - (float)x_low;
            const float hy = 1.f - ly, hx = 1.f - lx;
            p1 = y_low * width + x_low;  p2 = y_low * width + x_high;
            p3 = y_high * width + x_low; p4 = y_high * width + x_high;
            w1 = hy * hx; w2 = hy * lx; w3 = ly * hx; w4 = ly * lx;
        }
        t.p[0][k] = p1; t.p[1][k] = p2; t.p[2][k] = p3; t.p[3][k] = p4;
        t.w[0][k] = w1; t.w[1][k] = w2; t.w[2][k] = w3; t.w[3][k] = w4;
    }
}

__device__ __forceinline__ void pool_channels(const Taps& t, const float* __restrict__ plane0, int plane_elems,
                                              float* __restrict__ out0, int c_begin, int c_end, int bins, int spp) {
    const float count = (float)spp;
    const int total = (c_end - c_begin) * bins;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int cl = e / bins, bin = e - cl * bins;
        const int c = c_begin + cl;
        const float* src = plane0 + (size_t)c * plane_elems;
        float acc = 0.f;
        const int k0 = bin * spp;
        for (int s = 0; s < spp; ++s) {
            const int k = k0 + s;
            const float v1 = __ldg(src + t.p[0][k]), v2 = __ldg(src + t.p[1][k]);
            const float v3 = __ldg(src + t.p[2][k]), v4 = __ldg(src + t.p[3][k]);
            acc += t.w[0][k] * v1 + t.w[1][k] * v2 + t.w[2][k] * v3 + t.w[3][k] * v4;
        }
        out0[(size_t)c * bins + bin] = acc / count;
    }
}

__global__ void __launch_bounds__(kThreads)
roi_align_fwd_kernel(const float* __restrict__ input, int channels, int height, int width,
                     const float* __restrict__ rois, float scale, int ph, int pw, int sr, int c_per_cta,
                     float* __restrict__ out) {
    extern __shared__ float4 smem_raw4[];
    Taps& t = *reinterpret_cast<Taps*>(smem_raw4);
    const int n = blockIdx.x;
    const float* r = rois + (size_t)n * 5;
    const int b = (int)r[0];
    compute_taps(t, r[1], r[2], r[3], r[4], scale, height, width, ph, pw, sr);
    __syncthreads();
    const int c_begin = blockIdx.y * c_per_cta;
    const int c_end = min(channels, c_begin + c_per_cta);
    pool_channels(t, input + (size_t)b * channels * height * width, height * width,
                  out + (size_t)n * channels * ph * pw, c_begin, c_end, ph * pw, sr * sr);
}

struct GatherLevels {
    const float* feat[4];
    int h[4], w[4];
    float scale[4];
    int n_levels, k_min, k_max;
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

// grid (N, 2, channel slices): y = 0 RGB from the mapped FPN level, y = 1 depth
__global__ void __launch_bounds__(kThreads)
roi_gather_kernel(GatherLevels lv, const float* __restrict__ depth, int depth_h, int depth_w, float depth_scale,
                  int channels, const float* __restrict__ boxes, const int32_t* __restrict__ box_off, int n_images,
                  int pool, int sr, int c_per_cta, float* __restrict__ out_rgb, float* __restrict__ out_depth,
                  int32_t* __restrict__ levels_out) {
    extern __shared__ float4 smem_raw4[];
    Taps& t = *reinterpret_cast<Taps*>(smem_raw4);
    const int n = blockIdx.x;
    const float4 bx = __ldg((const float4*)boxes + n);
    // image of box n (poolers.py:96-107: roi batch index)
    int lo = 0, hi = n_images - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (box_off[mid] <= n) lo = mid; else hi = mid - 1;
    }
    const int b = lo;
    const float* src;
    float* dst;
    int H, W;
    float scale;
    if (blockIdx.y == 0) {
        const int l = map_level(bx.x, bx.y, bx.z, bx.w, lv.k_min, lv.k_max);
        if (levels_out && blockIdx.z == 0 && threadIdx.x == 0) levels_out[n] = l;
        H = lv.h[l]; W = lv.w[l]; scale = lv.scale[l];
        src = lv.feat[l] + (size_t)b * channels * H * W;
        dst = out_rgb + (size_t)n * channels * pool * pool;
    } else {
        H = depth_h; W = depth_w; scale = depth_scale;
        src = depth + (size_t)b * channels * H * W;
        dst = out_depth + (size_t)n * channels * pool * pool;
    }
    compute_taps(t, bx.x, bx.y, bx.z, bx.w, scale, H, W, pool, pool, sr);
    __syncthreads();
    const int c_begin = blockIdx.z * c_per_cta;
    const int c_end = min(channels, c_begin + c_per_cta);
    pool_channels(t, src, H * W, dst, c_begin, c_end, pool * pool, sr * sr);
}

// backward: scatter with fp32 atomics (ROIAlign_cuda.cu:178-254); one thread per (roi, c, bin)
__global__ void roi_align_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ rois, int64_t total,
                                     float scale, int channels, int height, int width, int ph, int pw, int sr,
                                     float* __restrict__ grad_in) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int bw = (int)(idx % pw);
        const int bh = (int)((idx / pw) % ph);
        const int c = (int)((idx / ((int64_t)pw * ph)) % channels);
        const int n = (int)(idx / ((int64_t)pw * ph * channels));
        const float* r = rois + (size_t)n * 5;
        const int b = (int)r[0];
        const float roi_start_w = r[1] * scale, roi_start_h = r[2] * scale;
        const float roi_end_w = r[3] * scale, roi_end_h = r[4] * scale;
        const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
        const float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
        const float bin_h = roi_height / (float)ph, bin_w = roi_width / (float)pw;
        float* dst = grad_in + ((size_t)b * channels + c) * height * width;
        const float go = grad[idx];
        const float count = (float)(sr * sr);
        for (int iy = 0; iy < sr; ++iy) {
            float y = roi_start_h + bh * bin_h + (float)(iy + .5f) * bin_h / (float)sr;
            for (int ix = 0; ix < sr; ++ix) {
                float x = roi_start_w + bw * bin_w + (float)(ix + .5f) * bin_w / (float)sr;
                float yy = y;
                if (yy < -1.0f || yy > (float)height || x < -1.0f || x > (float)width) continue;
                if (yy <= 0.f) yy = 0.f;
                if (x <= 0.f) x = 0.f;
                int y_low = (int)yy, x_low = (int)x, y_high, x_high;
                if (y_low >= height - 1) { y_high = y_low = height - 1; yy = (float)y_low; } else y_high = y_low + 1;
                if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else x_high = x_low + 1;
                const float ly = yy - (float)y_low, lx = x - (float)x_low;
                const float hy = 1.f - ly, hx = 1.f - lx;
                atomicAdd(dst + y_low * width + x_low, go * (hy * hx) / count);
                atomicAdd(dst + y_low * width + x_high, go * (hy * lx) / count);
                atomicAdd(dst + y_high * width + x_low, go * (ly * hx) / count);
                atomicAdd(dst + y_high * width + x_high, go * (ly * lx) / count);
            }
        }
    }
}

int channel_slices(int n_items, int channels) {
    // enough CTAs for >= 2 waves of 148 SMs x 2 resident CTAs, without slicing below 16 channels
    int slices = 1;
    while ((int64_t)n_items * slices < (int64_t)num_sms() * 4 && channels / (slices * 2) >= 16) slices *= 2;
    return slices;
}

DeviceOnce g_attr_set;
int ensure_attrs() {
    if (!g_attr_set.pending()) return VETO_OK;
    VETO_CUDA(cudaFuncSetAttribute(roi_align_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Taps)));
    VETO_CUDA(cudaFuncSetAttribute(roi_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Taps)));
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
    const int slices = channel_slices(n_rois, channels);
    const int c_per = (channels + slices - 1) / slices;
    dim3 grid(n_rois, slices);
    roi_align_fwd_kernel<<<grid, kThreads, sizeof(Taps), (cudaStream_t)stream>>>(
        input_dev, channels, height, width, rois_dev, spatial_scale, pooled_h, pooled_w, sampling_ratio, c_per, out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

extern "C" int veto_roi_align_backward(const float* grad_dev, const float* rois_dev, int n_rois, float spatial_scale,
                                       int pooled_h, int pooled_w, int batch, int channels, int height, int width,
                                       int sampling_ratio, float* grad_input_dev, veto_stream_t stream) {
    VETO_REQUIRE(grad_input_dev && batch > 0 && channels > 0 && height > 0 && width > 0 && sampling_ratio > 0, VETO_ERR_ARG,
                 "veto_roi_align_backward: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    VETO_CUDA(cudaMemsetAsync(grad_input_dev, 0, (size_t)batch * channels * height * width * sizeof(float), s));
    if (n_rois == 0) return VETO_OK;
    VETO_REQUIRE(grad_dev && rois_dev, VETO_ERR_ARG, "veto_roi_align_backward: bad argument");
    const int64_t total = (int64_t)n_rois * channels * pooled_h * pooled_w;
    const int64_t blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (int64_t)num_sms() * 32 ? blocks : (int64_t)num_sms() * 32);
    roi_align_bwd_kernel<<<grid, 256, 0, s>>>(grad_dev, rois_dev, total, spatial_scale, channels, height, width, pooled_h,
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
    const int slices = channel_slices(2 * n_boxes, channels);
    const int c_per = (channels + slices - 1) / slices;
    dim3 grid(n_boxes, 2, slices);
    roi_gather_kernel<<<grid, kThreads, sizeof(Taps), (cudaStream_t)stream>>>(
        lv, depth_dev, depth_h, depth_w, depth_scale, channels, boxes_dev, box_offsets_dev, n_images, pool, sampling_ratio,
        c_per, out_rgb_dev, out_depth_dev, levels_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
