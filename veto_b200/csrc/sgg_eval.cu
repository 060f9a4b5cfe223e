// Triplet matching of the scene-graph recall metrics: SGRecall.calculate_recall -> _compute_pred_matches
// (pysgg/data/datasets/evaluation/vg/sgg_eval.py:44-117,138-186; intersect_2d, bbox_overlaps of
// pysgg/utils/miscellaneous.py:47-83; boxlist_iou of structures/boxlist_ops.py:54-87).
//
// The reference matches on the host, image by image: a [G, P, 3] broadcast equality of (subject class, predicate,
// object class), then per matched ground-truth triplet two IoU rows against the candidate predictions, and finally
// recall@K = |union of pred_to_gt[:K]| / G.  Everything the recall family needs is, per ground-truth triplet g, the
// RANK of the first prediction that matches it:  g is in the union of the first K predictions  <=>  first_match[g] < K.
// One CTA per image computes that: threads sweep the G x P pairs, a pair matches when the three labels agree and both
// boxes overlap by IoU >= iou_thres (+1 box convention, fp32, the reference's operation order), and an atomicMin in
// shared memory keeps the smallest matching rank.  A second output counts the matches of every prediction
// (len(pred_to_gt[p])), which the no-graph-constraint / accuracy variants read.
#include "stages.cuh"

namespace veto {
namespace {

__device__ __forceinline__ float iou_plus1(const float* a, const float* b) {
    const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a[2], a[0]), 1.f), __fadd_rn(__fsub_rn(a[3], a[1]), 1.f));
    const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.f));
    const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a[2], b[2]), fmaxf(a[0], b[0])), 1.f), 0.f);
    const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a[3], b[3]), fmaxf(a[1], b[1])), 1.f), 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

__global__ void __launch_bounds__(256)
sgg_match_kernel(const int64_t* __restrict__ gt_trip, const float* __restrict__ gt_box, const int32_t* __restrict__ gt_off,
                 const int64_t* __restrict__ pr_trip, const float* __restrict__ pr_box, const int32_t* __restrict__ pr_off,
                 float iou_thres, int32_t* __restrict__ first_match, int32_t* __restrict__ pred_hits) {
    const int b = blockIdx.x;
    const int g0 = gt_off[b], G = gt_off[b + 1] - g0;
    const int p0 = pr_off[b], P = pr_off[b + 1] - p0;
    for (int g = threadIdx.x; g < G; g += blockDim.x) first_match[g0 + g] = 0x7fffffff;
    for (int p = threadIdx.x; p < P; p += blockDim.x) pred_hits[p0 + p] = 0;
    __syncthreads();
    const int64_t total = (int64_t)G * P;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int g = (int)(e / P), p = (int)(e - (int64_t)g * P);   // consecutive threads walk the predictions of one g
        const int64_t* gt = gt_trip + 3 * (size_t)(g0 + g);
        const int64_t* pr = pr_trip + 3 * (size_t)(p0 + p);
        if (gt[0] != pr[0] || gt[1] != pr[1] || gt[2] != pr[2]) continue;      // intersect_2d
        const float* gb = gt_box + 8 * (size_t)(g0 + g);
        const float* pb = pr_box + 8 * (size_t)(p0 + p);
        if (iou_plus1(gb, pb) >= iou_thres && iou_plus1(gb + 4, pb + 4) >= iou_thres) {
            atomicMin(first_match + g0 + g, p);
            atomicAdd(pred_hits + p0 + p, 1);
        }
    }
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_sgg_match(const int64_t* gt_triplets_dev, const float* gt_boxes_dev, const int32_t* gt_offsets_dev,
                              const int64_t* pred_triplets_dev, const float* pred_boxes_dev, const int32_t* pred_offsets_dev,
                              int n_images, float iou_thres, int32_t* first_match_dev, int32_t* pred_hits_dev,
                              veto_stream_t stream) {
    if (n_images <= 0) return VETO_OK;
    VETO_REQUIRE(gt_triplets_dev && gt_boxes_dev && gt_offsets_dev && pred_triplets_dev && pred_boxes_dev && pred_offsets_dev &&
                     first_match_dev && pred_hits_dev,
                 VETO_ERR_ARG, "veto_sgg_match: NULL argument");
    set_tag(TAG_POST);
    sgg_match_kernel<<<n_images, 256, 0, (cudaStream_t)stream>>>(gt_triplets_dev, gt_boxes_dev, gt_offsets_dev, pred_triplets_dev,
                                                                pred_boxes_dev, pred_offsets_dev, iou_thres, first_match_dev,
                                                                pred_hits_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
