// Ensemble.nms_per_cls (roi_relation_predictors.py:3855-3874) — the per-class greedy label assignment MEET runs on
// the detector's boxes at SGDet test time — with nms_overlaps (relation_head/utils_relation.py:56-79) folded in.
// The reference builds the [n, n, C] overlap tensor per image, moves it and the scores to the host and loops n times
// over a numpy argmax.  Here: one CTA per image, the [n, C] score tile lives in shared memory, and each of the n
// rounds is  (1) block-wide arg max (first index in row-major order on ties, like numpy.argmax),
//            (2) label[box] = cls,
//            (3) scores[j, cls] = 0 for every box j whose class-`cls` box overlaps box `box`'s by IoU >= thresh —
//                the one [n] column of the overlap tensor that round needs, computed on the fly,
//            (4) scores[box, :] = -1.
// The same kernel serves obj_prediction_nms (relation_head/utils_relation.py:94-128), the late NMS of the vanilla
// post-processor at SGDet test time (inference.py:414-417): background column 0 instead of -1, and a box keeps its
// first assignment.
// The IoU uses the reference's +1 convention and operation order with explicit round-to-nearest intrinsics (no FMA
// contraction), so the >= thresh decision is bit-identical to the fp32 torch expression.
#include "stages.cuh"

namespace veto {
namespace {

constexpr int NMS_THREADS = 256;

__device__ __forceinline__ float box_area(const float4 b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

__global__ void __launch_bounds__(NMS_THREADS)
obj_nms_kernel(const float* __restrict__ scores, const float* __restrict__ boxes_per_cls, const int32_t* __restrict__ box_off,
               int num_obj, float thresh, int late_nms, int64_t* __restrict__ labels_out) {
    extern __shared__ float s_scores[];  // [n, num_obj]
    __shared__ float red_v[NMS_THREADS / 32];
    __shared__ int red_i[NMS_THREADS / 32];
    __shared__ int s_pick;
    const int b0 = box_off[blockIdx.x], n = box_off[blockIdx.x + 1] - b0;
    if (n <= 0) return;
    const int total = n * num_obj;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int e = tid; e < total; e += NMS_THREADS) {
        const int c = e % num_obj;
        // background column: out_dists_sampled[:, 0] = -1 (nms_per_cls) / prob_sampled[:, 0] = 0 (obj_prediction_nms)
        s_scores[e] = c == 0 ? (late_nms ? 0.f : -1.f) : scores[(size_t)b0 * num_obj + e];
    }
    for (int j = tid; j < n; j += NMS_THREADS) labels_out[b0 + j] = 0;
    __syncthreads();
    for (int round = 0; round < n; ++round) {
        // (1) arg max, smallest flat index among equal values
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int e = tid; e < total; e += NMS_THREADS) {
            const float v = s_scores[e];
            if (v > best) { best = v; besti = e; }   // e ascends per thread: the first maximum is kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { red_v[wid] = best; red_i[wid] = besti; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NMS_THREADS / 32; ++w)
                if (red_v[w] > best || (red_v[w] == best && red_i[w] < besti)) { best = red_v[w]; besti = red_i[w]; }
            s_pick = besti;
        }
        __syncthreads();
        const int box = s_pick / num_obj, cls = s_pick - box * num_obj;
        // (2) + (3): column `cls` of is_overlap[box, :, :]
        const float4 a = __ldg((const float4*)boxes_per_cls + (size_t)(b0 + box) * num_obj + cls);
        const float area_a = box_area(a);
        for (int j = tid; j < n; j += NMS_THREADS) {
            const float4 q = __ldg((const float4*)boxes_per_cls + (size_t)(b0 + j) * num_obj + cls);
            const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, q.z), fmaxf(a.x, q.x)), 1.f), 0.f);
            const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, q.w), fmaxf(a.y, q.y)), 1.f), 0.f);
            const float inter = __fmul_rn(w, h);
            // union = -inters + areas[None] + areas[:, None] at [box, j]: (-inter + area_j) + area_box
            const float uni = __fadd_rn(__fadd_rn(-inter, box_area(q)), area_a);
            if (__fdiv_rn(inter, uni) >= thresh) s_scores[j * num_obj + cls] = 0.f;
        }
        // obj_prediction_nms keeps the first (highest-probability) assignment of a box (utils_relation.py:119-123)
        if (tid == 0 && !(late_nms && labels_out[b0 + box] > 0)) labels_out[b0 + box] = cls;
        __syncthreads();
        // (4) the picked box leaves the pool
        for (int c = tid; c < num_obj; c += NMS_THREADS) s_scores[box * num_obj + c] = -1.f;
        __syncthreads();
    }
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_obj_nms_per_cls(const float* scores_dev, const float* boxes_per_cls_dev, const int32_t* box_offsets_dev,
                                    const int32_t* n_boxes_host, int n_images, int num_obj, float thresh, int late_nms,
                                    int64_t* labels_out_dev, veto_stream_t stream) {
    VETO_REQUIRE(scores_dev && boxes_per_cls_dev && box_offsets_dev && n_boxes_host && labels_out_dev && n_images >= 0 && num_obj > 1,
                 VETO_ERR_ARG, "veto_obj_nms_per_cls: bad arguments");
    if (n_images == 0) return VETO_OK;
    int n_max = 0;
    for (int b = 0; b < n_images; ++b) n_max = n_boxes_host[b] > n_max ? n_boxes_host[b] : n_max;
    if (n_max == 0) return VETO_OK;
    const size_t smem = (size_t)n_max * num_obj * sizeof(float);
    VETO_REQUIRE(smem <= 220 * 1024, VETO_ERR_UNSUPPORTED,
                 "veto_obj_nms_per_cls: %d boxes x %d classes does not fit one CTA's shared memory", n_max, num_obj);
    cudaStream_t s = (cudaStream_t)stream;
    VETO_CUDA(cudaFuncSetAttribute(obj_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set_tag(TAG_PAIRS);
    obj_nms_kernel<<<n_images, NMS_THREADS, smem, s>>>(scores_dev, boxes_per_cls_dev, box_offsets_dev, num_obj, thresh, late_nms, labels_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
