/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's legacy (non-aligned) ROIAlign forward:
 *   /root/reference/pysgg/csrc/cpu/ROIAlign_cpu.cpp:114-219  (ROIAlignForward_cpu_kernel)
 *   /root/reference/pysgg/csrc/cpu/ROIAlign_cpu.cpp:17-112   (bilinear tap pre-computation)
 * and of the scatter backward the CUDA twin performs
 *   /root/reference/pysgg/csrc/cuda/ROIAlign_cuda.cu:124-254 (bilinear_interpolate_gradient,
 *   RoIAlignBackwardFeature) with the atomics replaced by a serial accumulation.
 *
 * Semantics kept on purpose: no half-pixel offset, roi_w = max(x2*s - x1*s, 1), samples outside
 * [-1,H]x[-1,W] contribute 0, coordinates clamped to the border, average of sr*sr bilinear samples,
 * fp32 arithmetic in the reference's operation order (w1*v1 + w2*v2 + w3*v3 + w4*v4 accumulated
 * sample by sample, then one division by the sample count).
 *
 * Pinned (tests/test_oracle.py) bit-exactly against torchvision.ops.roi_align(aligned=False) and
 * against oracle/_ref (the unmodified reference .cpp compiled from /root/reference).
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/liboracle_roialign.so oracle/roialign_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int p1, p2, p3, p4; float w1, w2, w3, w4; } tap_t;

static void taps_for_sample(int height, int width, float y, float x, tap_t* t) {
    if (y < -1.0 || y > height || x < -1.0 || x > width) { memset(t, 0, sizeof(*t)); return; }
    if (y <= 0) y = 0;
    if (x <= 0) x = 0;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else y_high = y_low + 1;
    if (x_low >= width - 1)  { x_high = x_low = width - 1;  x = (float)x_low; } else x_high = x_low + 1;
    float ly = y - y_low, lx = x - x_low;
    /* the reference writes `1. - ly` with a double literal: the subtraction happens in double and is
     * rounded back to float (ROIAlign_cpu.cpp:91) */
    float hy = (float)(1. - ly), hx = (float)(1. - lx);
    t->p1 = y_low * width + x_low;  t->p2 = y_low * width + x_high;
    t->p3 = y_high * width + x_low; t->p4 = y_high * width + x_high;
    t->w1 = hy * hx; t->w2 = hy * lx; t->w3 = ly * hx; t->w4 = ly * lx;
}

/* rois: [n_rois,5] = (batch_idx, x1, y1, x2, y2); input NCHW; out [n_rois,C,ph,pw] */
int oracle_roi_align_forward(const float* input, int channels, int height, int width,
                             const float* rois, int n_rois, float spatial_scale,
                             int pooled_h, int pooled_w, int sampling_ratio, float* out) {
    for (int n = 0; n < n_rois; n++) {
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        float roi_start_w = r[1] * spatial_scale, roi_start_h = r[2] * spatial_scale;
        float roi_end_w = r[3] * spatial_scale, roi_end_h = r[4] * spatial_scale;
        float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
        float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
        float bin_h = roi_height / (float)pooled_h, bin_w = roi_width / (float)pooled_w;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / pooled_h);
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / pooled_w);
        const float count = (float)(gh * gw);
        size_t ntap = (size_t)gh * gw * pooled_h * pooled_w;
        tap_t* taps = (tap_t*)malloc(ntap * sizeof(tap_t));
        if (!taps) return -1;
        size_t k = 0;
        for (int ph = 0; ph < pooled_h; ph++)
            for (int pw = 0; pw < pooled_w; pw++)
                for (int iy = 0; iy < gh; iy++) {
                    float yy = roi_start_h + ph * bin_h + (float)(iy + .5f) * bin_h / (float)gh;
                    for (int ix = 0; ix < gw; ix++) {
                        float xx = roi_start_w + pw * bin_w + (float)(ix + .5f) * bin_w / (float)gw;
                        taps_for_sample(height, width, yy, xx, &taps[k++]);
                    }
                }
        for (int c = 0; c < channels; c++) {
            const float* src = input + ((size_t)b * channels + c) * height * width;
            float* dst = out + (((size_t)n * channels + c) * pooled_h) * pooled_w;
            k = 0;
            for (int i = 0; i < pooled_h * pooled_w; i++) {
                float acc = 0.f;
                for (int s = 0; s < gh * gw; s++) {
                    tap_t t = taps[k++];
                    acc += t.w1 * src[t.p1] + t.w2 * src[t.p2] + t.w3 * src[t.p3] + t.w4 * src[t.p4];
                }
                dst[i] = acc / count;
            }
        }
        free(taps);
    }
    return 0;
}

/* grad_in [B,C,H,W] must be zero-initialised by the caller. */
int oracle_roi_align_backward(const float* grad_out, int channels, int height, int width,
                              const float* rois, int n_rois, float spatial_scale,
                              int pooled_h, int pooled_w, int sampling_ratio, float* grad_in) {
    for (int n = 0; n < n_rois; n++) {
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        float roi_start_w = r[1] * spatial_scale, roi_start_h = r[2] * spatial_scale;
        float roi_end_w = r[3] * spatial_scale, roi_end_h = r[4] * spatial_scale;
        float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
        float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
        float bin_h = roi_height / (float)pooled_h, bin_w = roi_width / (float)pooled_w;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / pooled_h);
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / pooled_w);
        const float count = (float)(gh * gw);
        for (int c = 0; c < channels; c++) {
            float* dst = grad_in + ((size_t)b * channels + c) * height * width;
            const float* g = grad_out + (((size_t)n * channels + c) * pooled_h) * pooled_w;
            for (int ph = 0; ph < pooled_h; ph++)
                for (int pw = 0; pw < pooled_w; pw++) {
                    float go = g[ph * pooled_w + pw];
                    for (int iy = 0; iy < gh; iy++) {
                        float yy = roi_start_h + ph * bin_h + (float)(iy + .5f) * bin_h / (float)gh;
                        for (int ix = 0; ix < gw; ix++) {
                            float xx = roi_start_w + pw * bin_w + (float)(ix + .5f) * bin_w / (float)gw;
                            tap_t t;
                            taps_for_sample(height, width, yy, xx, &t);
                            if (t.w1 == 0.f && t.w2 == 0.f && t.w3 == 0.f && t.w4 == 0.f) continue;
                            dst[t.p1] += go * t.w1 / count; dst[t.p2] += go * t.w2 / count;
                            dst[t.p3] += go * t.w3 / count; dst[t.p4] += go * t.w4 / count;
                        }
                    }
                }
        }
    }
    return 0;
}
