// PostProcessor.forward, vanilla branch (relation_head/inference.py:398-453), one CTA per image:
//   rel_class_prob = softmax(rel_logit);  rel_scores, rel_class = rel_class_prob[:,1:].max(1) (+1)
//   triple = rel_scores * obj_scores[s] * obj_scores[o];  sort descending; reorder pairs / probs / labels.
// The reference sort is unstable; ties are broken here by the original row (ascending).  Warp per row for
// the softmax (lanes over predicate classes, shuffle reductions), bitonic sort of 64-bit keys in shared
// memory, then a second warp-per-row pass that writes the probability rows in ranked order.
#include "common.cuh"

namespace veto {
namespace {

constexpr int kMaxRows = 16384;

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sort storage: images of up to kMaxRows rows sort their 64-bit keys in shared memory.  Larger images (MAX_PROPOSAL_PAIR
// 4096+ with the MEET heads merged) sort in place in global memory: the image's own pairs_out region (16 B per row)
// holds the npow <= 2 * rows keys, the ranked order is then parked in labels_out (one int64 per row) before the
// output pass overwrites both.  Labels and triple scores are recomputed in the output pass instead of being stored.
__device__ __forceinline__ void bitonic_sort_keys(unsigned long long* keys, int npow) {
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = threadIdx.x; q < npow; q += blockDim.x) {
                const int p = q ^ j;
                if (p > q) {
                    const unsigned long long a = keys[q], c = keys[p];
                    const bool up = ((q & k) == 0);
                    if ((a > c) == up) { keys[q] = c; keys[p] = a; }
                }
            }
            __syncthreads();
        }
    }
}
__device__ __forceinline__ unsigned long long rank_key(float t, int q) {
    unsigned int u = __float_as_uint(t);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)(~u) << 32) | (unsigned long long)(unsigned int)q;
}
// park the ranked order of an oversized image in labels_out (see above); returns after a CTA barrier
__device__ __forceinline__ void park_order(const unsigned long long* keys, int n, int64_t* labels_out_img, bool big) {
    if (!big) return;
    for (int r = threadIdx.x; r < n; r += blockDim.x) labels_out_img[r] = (int64_t)(keys[r] & 0xffffffffull);
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
postprocess_kernel(const float* __restrict__ logits, int num_rel, const int64_t* __restrict__ pairs,
                   const float* __restrict__ obj_scores, const int32_t* __restrict__ rel_off,
                   const int32_t* __restrict__ box_off, int64_t* __restrict__ pairs_out, float* __restrict__ probs_out,
                   int64_t* __restrict__ labels_out, float* __restrict__ triple_out) {
    extern __shared__ unsigned long long smem_keys[];  // [kMaxRows]
    const int b = blockIdx.x;
    const int r0 = rel_off[b], rows = rel_off[b + 1] - r0;
    if (rows <= 0) return;
    const bool big = rows > kMaxRows;
    unsigned long long* keys = big ? (unsigned long long*)(pairs_out + 2 * (size_t)r0) : smem_keys;
    const int boff = box_off[b];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int npow = 1;
    while (npow < rows) npow <<= 1;

    // softmax statistics, arg max over classes 1.. (first index on ties) and triple score of row q (whole warp)
    auto score_row = [&](int q, float& m, float& sum, int& besti, float& t) {
        const float* lg = logits + (size_t)(r0 + q) * num_rel;
        m = -INFINITY;
        for (int c = lane; c < num_rel; c += 32) m = fmaxf(m, lg[c]);
        m = wmax(m);
        sum = 0.f;
        float best = -INFINITY;
        besti = 0x7fffffff;
        for (int c = lane; c < num_rel; c += 32) {
            const float e = expf(lg[c] - m);
            sum += e;
            if (c >= 1 && e > best) { best = e; besti = c; }
        }
        sum = wsum(sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        const float score = best / sum;
        const longlong2 pr = *(const longlong2*)(pairs + 2 * (size_t)(r0 + q));
        t = score * obj_scores[boff + pr.x] * obj_scores[boff + pr.y];
    };

    for (int q = wid; q < rows; q += nw) {
        float m, sum, t;
        int besti;
        score_row(q, m, sum, besti, t);
        if (lane == 0) keys[q] = rank_key(t, q);
    }
    for (int q = rows + threadIdx.x; q < npow; q += blockDim.x) keys[q] = ~0ull;
    __syncthreads();
    bitonic_sort_keys(keys, npow);
    park_order(keys, rows, labels_out + r0, big);
    for (int rank = wid; rank < rows; rank += nw) {
        const int q = big ? (int)labels_out[r0 + rank] : (int)(keys[rank] & 0xffffffffull);
        float m, sum, t;
        int besti;
        score_row(q, m, sum, besti, t);
        const float* lg = logits + (size_t)(r0 + q) * num_rel;
        float* po = probs_out + (size_t)(r0 + rank) * num_rel;
        for (int c = lane; c < num_rel; c += 32) po[c] = expf(lg[c] - m) / sum;
        if (lane == 0) {
            *(longlong2*)(pairs_out + 2 * (size_t)(r0 + rank)) = *(const longlong2*)(pairs + 2 * (size_t)(r0 + q));
            labels_out[r0 + rank] = besti;
            triple_out[r0 + rank] = t;
        }
    }
}

// PostProcessor.forward, MEET 'ensemble' branch (EXPERT_GROUP False, relation_head/inference.py:284-397), one CTA per
// image.  Every group head k scores every pair: softmax over the head's n_k + 2 columns, the last (out-of-group) column
// is dropped (:345-346), rel_score / rel_class = max over the head's member columns 1..n_k; the reference's "chosen"
// filter (:356-361) keeps every row (rel_class is by construction one of the members), so the image's merged list has
// G * R rows, group-major.  It is ranked by triple score (ties: merged index ascending) and each row's probabilities
// are scattered from head-local into global predicate columns (chosen_labels_incr, :387), zeros elsewhere; the label
// stays head-local like the reference's (:371,388).
__global__ void __launch_bounds__(1024)
meet_postprocess_kernel(const float* __restrict__ logits, int ld, const int32_t* __restrict__ head_off, int n_heads,
                        const int32_t* __restrict__ col_map, int num_rel, const int64_t* __restrict__ pairs,
                        const float* __restrict__ obj_scores, const int32_t* __restrict__ rel_off,
                        const int32_t* __restrict__ box_off, int64_t* __restrict__ pairs_out, float* __restrict__ probs_out,
                        int64_t* __restrict__ labels_out, float* __restrict__ triple_out) {
    extern __shared__ unsigned long long smem_keys[];  // [kMaxRows]
    const int b = blockIdx.x;
    const int r0 = rel_off[b], rows = rel_off[b + 1] - r0;
    const int merged = rows * n_heads;
    if (merged <= 0) return;
    const bool big = merged > kMaxRows;
    const int boff = box_off[b];
    const size_t out0 = (size_t)r0 * n_heads;  // merged rows of the images before this one
    unsigned long long* keys = big ? (unsigned long long*)(pairs_out + 2 * out0) : smem_keys;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int npow = 1;
    while (npow < merged) npow <<= 1;

    // head k's softmax statistics over its n_k + 2 columns, arg max over the member columns 1..n_k and triple score of
    // merged row q = k * rows + r (whole warp)
    auto score_row = [&](int q, float& m, float& sum, int& besti, float& t) {
        const int k = q / rows, r = q - k * rows;
        const int c0 = head_off[k], nc = head_off[k + 1] - c0;
        const float* lg = logits + (size_t)(r0 + r) * ld + c0;
        m = -INFINITY;
        for (int c = lane; c < nc; c += 32) m = fmaxf(m, lg[c]);
        m = wmax(m);
        sum = 0.f;
        float best = -INFINITY;
        besti = 0x7fffffff;
        for (int c = lane; c < nc; c += 32) {
            const float e = expf(lg[c] - m);
            sum += e;
            if (c >= 1 && c < nc - 1 && e > best) { best = e; besti = c; }
        }
        sum = wsum(sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        const float score = best / sum;
        const longlong2 pr = *(const longlong2*)(pairs + 2 * (size_t)(r0 + r));
        t = score * obj_scores[boff + pr.x] * obj_scores[boff + pr.y];
    };

    for (int q = wid; q < merged; q += nw) {
        float m, sum, t;
        int besti;
        score_row(q, m, sum, besti, t);
        if (lane == 0) keys[q] = rank_key(t, q);
    }
    for (int q = merged + threadIdx.x; q < npow; q += blockDim.x) keys[q] = ~0ull;
    __syncthreads();
    bitonic_sort_keys(keys, npow);
    park_order(keys, merged, labels_out + out0, big);
    for (int rank = wid; rank < merged; rank += nw) {
        const int q = big ? (int)labels_out[out0 + rank] : (int)(keys[rank] & 0xffffffffull);
        const int k = q / rows, r = q - k * rows;
        const int c0 = head_off[k], nc = head_off[k + 1] - c0;
        const float* lg = logits + (size_t)(r0 + r) * ld + c0;
        float m, sum, t;
        int besti;
        score_row(q, m, sum, besti, t);
        float* po = probs_out + (out0 + rank) * num_rel;
        for (int c = lane; c < num_rel; c += 32) po[c] = 0.f;
        __syncwarp();
        for (int c = lane; c < nc - 1; c += 32) po[col_map[c0 + c]] = expf(lg[c] - m) / sum;  // distinct global columns
        if (lane == 0) {
            *(longlong2*)(pairs_out + 2 * (out0 + rank)) = *(const longlong2*)(pairs + 2 * (size_t)(r0 + r));
            labels_out[out0 + rank] = besti;
            triple_out[out0 + rank] = t;
        }
    }
}

// PostProcessor.forward, MEET EXPERT_GROUP branch (relation_head/inference.py:93-283): three experts per group, a
// candidate (group j, pair r) survives the vote when the experts' predicted classes agree — all three ('U', unanimous)
// or at least two ('C', consensus).  One warp handles one candidate: three softmaxes over the experts' heads (out-of-
// group column dropped), the per-expert class / triple score, then
//   unanimous: score = mean of the three triple scores, probabilities = mean of the three rows;
//   consensus: over the agreeing expert PAIRS (0,1), (1,2), (0,2): score = mean of the pair means, probabilities = mean
//              of the pair means — where the reference's (1,2) probability "mean" is mean(p1, p1) = p1 (:191-193);
//              reproduced as is.
// Survivors are ranked by score (ties: merged index) and scattered into global predicate columns like the 'ensemble'
// branch; the number of survivors per image goes to counts_out.  Heads are laid out expert-major: head e * G + j.
struct ExpertRow {
    float t[3];   // triple scores of the three experts
    int c[3];     // predicted head-local classes
};

__device__ __forceinline__ void expert_softmax(const float* lg, int nc, int lane, float& m, float& sum, float& best, int& besti) {
    m = -INFINITY;
    for (int c = lane; c < nc; c += 32) m = fmaxf(m, lg[c]);
    m = wmax(m);
    sum = 0.f;
    best = -INFINITY;
    besti = 0x7fffffff;
    for (int c = lane; c < nc; c += 32) {
        const float e = expf(lg[c] - m);
        sum += e;
        if (c >= 1 && c < nc - 1 && e > best) { best = e; besti = c; }
    }
    sum = wsum(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
}

// vote of one candidate: returns whether it survives; score / class of the survivor; w[e] = weight of expert e's
// probability row in the survivor's probability row (sums to 1)
__device__ __forceinline__ bool expert_vote(const ExpertRow& x, int consensus, float& score, int& cls, float (&w)[3]) {
    const bool a01 = x.c[0] == x.c[1], a12 = x.c[1] == x.c[2], a02 = x.c[0] == x.c[2];
    if (!consensus) {
        score = ((x.t[0] + x.t[1]) + x.t[2]) / 3.f;
        cls = x.c[2];
        w[0] = w[1] = w[2] = 1.f / 3.f;
        return a01 && a12 && a02;
    }
    const int count = (int)a01 + (int)a12 + (int)a02;
    if (count == 0) return false;
    float ssum = 0.f;
    w[0] = w[1] = w[2] = 0.f;
    if (a01) { ssum += (x.t[0] + x.t[1]) * 0.5f; w[0] += 0.5f; w[1] += 0.5f; cls = x.c[0]; }
    if (a12) { ssum += (x.t[1] + x.t[2]) * 0.5f; w[1] += 1.0f; cls = x.c[1]; }              // mean(p1, p1): :191-193
    if (a02) { ssum += (x.t[0] + x.t[2]) * 0.5f; w[0] += 0.5f; w[2] += 0.5f; cls = x.c[2]; }
    score = ssum / (float)count;
    const float inv = 1.f / (float)count;
    w[0] *= inv; w[1] *= inv; w[2] *= inv;
    return true;
}

__global__ void __launch_bounds__(1024)
meet_vote_kernel(const float* __restrict__ logits, int ld, const int32_t* __restrict__ head_off, int n_groups,
                 const int32_t* __restrict__ col_map, int num_rel, int consensus, const int64_t* __restrict__ pairs,
                 const float* __restrict__ obj_scores, const int32_t* __restrict__ rel_off, const int32_t* __restrict__ box_off,
                 int64_t* __restrict__ pairs_out, float* __restrict__ probs_out, int64_t* __restrict__ labels_out,
                 float* __restrict__ triple_out, int32_t* __restrict__ counts_out) {
    extern __shared__ unsigned long long smem_keys[];  // [kMaxRows]
    __shared__ int s_kept;
    const int b = blockIdx.x;
    const int r0 = rel_off[b], rows = rel_off[b + 1] - r0;
    const int merged = rows * n_groups;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    if (merged <= 0) {
        if (threadIdx.x == 0) counts_out[b] = 0;
        return;
    }
    const bool big = merged > kMaxRows;
    const int boff = box_off[b];
    const size_t out0 = (size_t)r0 * n_groups;
    unsigned long long* keys = big ? (unsigned long long*)(pairs_out + 2 * out0) : smem_keys;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int npow = 1;
    while (npow < merged) npow <<= 1;

    auto experts = [&](int q, ExpertRow& x, float (&mm)[3], float (&ss)[3]) {
        const int j = q / rows, r = q - j * rows;
        const longlong2 pr = *(const longlong2*)(pairs + 2 * (size_t)(r0 + r));
        const float so = obj_scores[boff + pr.x] * obj_scores[boff + pr.y];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int c0 = head_off[e * n_groups + j], nc = head_off[e * n_groups + j + 1] - c0;
            float best;
            int besti;
            expert_softmax(logits + (size_t)(r0 + r) * ld + c0, nc, lane, mm[e], ss[e], best, besti);
            x.c[e] = besti;
            x.t[e] = (best / ss[e]) * so;
        }
    };

    int my_kept = 0;
    for (int q = wid; q < merged; q += nw) {
        ExpertRow x;
        float mm[3], ss[3], w[3], score = 0.f;
        int cls = 0;
        experts(q, x, mm, ss);
        const bool keep = expert_vote(x, consensus, score, cls, w);
        if (lane == 0) {
            if (keep) {
                keys[q] = rank_key(score, q);
                ++my_kept;
            } else {
                keys[q] = ~0ull;
            }
        }
    }
    if (lane == 0 && my_kept) atomicAdd(&s_kept, my_kept);
    for (int q = merged + threadIdx.x; q < npow; q += blockDim.x) keys[q] = ~0ull;
    __syncthreads();
    bitonic_sort_keys(keys, npow);
    const int kept = s_kept;
    if (threadIdx.x == 0) counts_out[b] = kept;
    park_order(keys, kept, labels_out + out0, big);
    if (big) {
        // the key storage was this image's pairs_out region: rows past the survivors go back to the zeros the caller put there
        for (size_t i = 2 * (size_t)kept + threadIdx.x; i < (size_t)npow && i < 2 * (size_t)merged; i += blockDim.x)
            pairs_out[2 * out0 + i] = 0;
    }
    for (int rank = wid; rank < kept; rank += nw) {
        const int q = big ? (int)labels_out[out0 + rank] : (int)(keys[rank] & 0xffffffffull);
        const int j = q / rows, r = q - j * rows;
        ExpertRow x;
        float mm[3], ss[3], w[3], score = 0.f;
        int cls = 0;
        experts(q, x, mm, ss);
        expert_vote(x, consensus, score, cls, w);
        float* po = probs_out + (out0 + rank) * num_rel;
        for (int c = lane; c < num_rel; c += 32) po[c] = 0.f;
        __syncwarp();
        const int c0 = head_off[j], nc = head_off[j + 1] - c0;   // every expert of group j has the same width / column map
        for (int c = lane; c < nc - 1; c += 32) {
            float v = 0.f;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const float* lg = logits + (size_t)(r0 + r) * ld + head_off[e * n_groups + j];
                v += w[e] * (expf(lg[c] - mm[e]) / ss[e]);
            }
            po[col_map[c0 + c]] = v;
        }
        if (lane == 0) {
            *(longlong2*)(pairs_out + 2 * (out0 + rank)) = *(const longlong2*)(pairs + 2 * (size_t)(r0 + r));
            labels_out[out0 + rank] = cls;
            triple_out[out0 + rank] = score;
        }
    }
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_postprocess(const float* rel_logits_dev, int num_rel, const int64_t* pairs_dev,
                                const float* obj_scores_dev, const int32_t* rel_offsets_dev, const int32_t* box_offsets_dev,
                                int n_images, int64_t n_pairs, int64_t* pairs_out_dev, float* probs_out_dev,
                                int64_t* labels_out_dev, float* triple_out_dev, veto_stream_t stream) {
    if (n_images <= 0 || n_pairs <= 0) return VETO_OK;
    VETO_REQUIRE(rel_logits_dev && pairs_dev && obj_scores_dev && rel_offsets_dev && box_offsets_dev && pairs_out_dev &&
                     probs_out_dev && labels_out_dev && triple_out_dev && num_rel >= 2 && num_rel < 65536,
                 VETO_ERR_ARG, "veto_postprocess: bad argument");
    set_tag(TAG_POST);
    static DeviceOnce attr_set;
    const int smem = kMaxRows * (int)sizeof(unsigned long long);
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(postprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.done();
    }
    postprocess_kernel<<<n_images, 1024, smem, (cudaStream_t)stream>>>(rel_logits_dev, num_rel, pairs_dev, obj_scores_dev,
                                                                      rel_offsets_dev, box_offsets_dev, pairs_out_dev,
                                                                      probs_out_dev, labels_out_dev, triple_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

extern "C" int veto_postprocess_meet(const float* group_logits_dev, int num_out, const int32_t* head_offsets_dev, int n_heads,
                                     const int32_t* col_map_dev, int num_rel, const int64_t* pairs_dev,
                                     const float* obj_scores_dev, const int32_t* rel_offsets_dev,
                                     const int32_t* box_offsets_dev, int n_images, int64_t n_pairs, int64_t* pairs_out_dev,
                                     float* probs_out_dev, int64_t* labels_out_dev, float* triple_out_dev,
                                     veto_stream_t stream) {
    if (n_images <= 0 || n_pairs <= 0) return VETO_OK;
    VETO_REQUIRE(group_logits_dev && head_offsets_dev && col_map_dev && pairs_dev && obj_scores_dev && rel_offsets_dev &&
                     box_offsets_dev && pairs_out_dev && probs_out_dev && labels_out_dev && triple_out_dev && n_heads >= 1 &&
                     num_out >= 3 * n_heads && num_rel >= 2 && num_rel < 65536,
                 VETO_ERR_ARG, "veto_postprocess_meet: bad argument");
    set_tag(TAG_POST);
    static DeviceOnce attr_set;
    const int smem = kMaxRows * (int)sizeof(unsigned long long);
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(meet_postprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.done();
    }
    meet_postprocess_kernel<<<n_images, 1024, smem, (cudaStream_t)stream>>>(
        group_logits_dev, num_out, head_offsets_dev, n_heads, col_map_dev, num_rel, pairs_dev, obj_scores_dev, rel_offsets_dev,
        box_offsets_dev, pairs_out_dev, probs_out_dev, labels_out_dev, triple_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

extern "C" int veto_postprocess_meet_vote(const float* group_logits_dev, int num_out, const int32_t* head_offsets_dev, int n_groups,
                                          const int32_t* col_map_dev, int num_rel, int consensus, const int64_t* pairs_dev,
                                          const float* obj_scores_dev, const int32_t* rel_offsets_dev,
                                          const int32_t* box_offsets_dev, int n_images, int64_t n_pairs, int64_t* pairs_out_dev,
                                          float* probs_out_dev, int64_t* labels_out_dev, float* triple_out_dev,
                                          int32_t* counts_out_dev, veto_stream_t stream) {
    if (n_images <= 0) return VETO_OK;
    VETO_REQUIRE(group_logits_dev && head_offsets_dev && col_map_dev && pairs_dev && obj_scores_dev && rel_offsets_dev &&
                     box_offsets_dev && pairs_out_dev && probs_out_dev && labels_out_dev && triple_out_dev && counts_out_dev &&
                     n_groups >= 1 && num_out >= 9 * n_groups && num_rel >= 2 && num_rel < 65536 && n_pairs >= 0,
                 VETO_ERR_ARG, "veto_postprocess_meet_vote: bad argument");
    set_tag(TAG_POST);
    static DeviceOnce attr_set;
    const int smem = kMaxRows * (int)sizeof(unsigned long long);
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(meet_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.done();
    }
    meet_vote_kernel<<<n_images, 1024, smem, (cudaStream_t)stream>>>(
        group_logits_dev, num_out, head_offsets_dev, n_groups, col_map_dev, num_rel, consensus, pairs_dev, obj_scores_dev,
        rel_offsets_dev, box_offsets_dev, pairs_out_dev, probs_out_dev, labels_out_dev, triple_out_dev, counts_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
