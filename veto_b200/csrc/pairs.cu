// Candidate-pair enumeration (RelationSampling.prepare_test_pairs, sampling.py:31-52) and the
// local->global index step of the predictor (roi_relation_predictors.py:4104-4115).
//
// Two kernels:
//   pairs_dense_kernel   — no IoU filter, no image over the cap: closed form r -> (i, j) of the
//                          row-major nonzero(ones - eye) order; pure streaming write, 16 B per pair,
//                          grid-stride over all images at once.
//   pairs_filtered_kernel— IoU filter and / or top-max_pairs cap: one CTA per image; candidates are
//                          compacted in row-major order with a block scan, and over the cap sorted in
//                          shared memory by (score product desc, row-major index asc) with a bitonic
//                          network on 64-bit keys.
#include "common.cuh"

namespace veto {
namespace {

constexpr int kMaxCand = 16384;  // candidates of one image held in shared memory (n <= 128)

__global__ void pairs_dense_kernel(const int32_t* __restrict__ n_boxes, const int32_t* __restrict__ out_off,
                                   int n_images, int64_t total, int64_t* __restrict__ pairs) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
        // image of row r: last b with out_off[b] <= r
        int lo = 0, hi = n_images - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (out_off[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const int n = n_boxes[lo];
        const int q = (int)(r - out_off[lo]);
        longlong2 v;
        if (n < 2) {
            v.x = 0; v.y = 0;  // placeholder [[0,0]] (sampling.py:47-51)
        } else {
            const int i = q / (n - 1);
            int j = q - i * (n - 1);
            j += (j >= i);
            v.x = i; v.y = j;
        }
        *(longlong2*)(pairs + 2 * r) = v;
    }
}

// boxlist_iou(p, p) > 0 in fp32, operation by operation (structures/boxlist_ops.py:54-87, TO_REMOVE = 1)
__device__ __forceinline__ bool iou_positive(const float4 a, const float4 b) {
    const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
    const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
    const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
    const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
    const float inter = __fmul_rn(w, h);
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return iou > 0.f;
}

__global__ void __launch_bounds__(1024)
pairs_filtered_kernel(const int32_t* __restrict__ n_boxes, const int32_t* __restrict__ box_off,
                      const int32_t* __restrict__ out_off, const float* __restrict__ boxes,
                      const float* __restrict__ scores, int require_overlap, int max_pairs,
                      int64_t* __restrict__ pairs, int32_t* __restrict__ counts) {
    extern __shared__ unsigned long long keys[];  // [kMaxCand]
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const int b = blockIdx.x;
    const int n = n_boxes[b];
    const int boff = box_off[b];
    int64_t* out = pairs + 2 * (int64_t)out_off[b];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cand = n * (n - 1);  // host guarantees cand <= kMaxCand
    if (tid == 0) s_base = 0;
    __syncthreads();

    // pass 1: compact the surviving candidates, row-major, into keys[] as (i<<16 | j)
    for (int start = 0; start < cand; start += blockDim.x) {
        const int q = start + tid;
        bool keep = false;
        int i = 0, j = 0;
        if (q < cand) {
            i = q / (n - 1);
            j = q - i * (n - 1);
            j += (j >= i);
            keep = true;
            if (require_overlap) {
                const float4 bi = __ldg((const float4*)boxes + boff + i);
                const float4 bj = __ldg((const float4*)boxes + boff + j);
                keep = iou_positive(bi, bj);
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        const int before = __popc(ball & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(ball);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const int t = warp_tot[w];
            if (w < wid) woff += t;
            tot += t;
        }
        const int base = s_base;
        if (keep) keys[base + woff + before] = ((unsigned long long)i << 16) | (unsigned long long)j;
        __syncthreads();
        if (tid == 0) s_base = base + tot;
        __syncthreads();
    }
    const int kept = s_base;

    if (kept == 0) {
        if (tid == 0) {
            out[0] = 0; out[1] = 0;
            if (counts) counts[b] = 1;
        }
        return;
    }
    if (kept <= max_pairs) {
        for (int q = tid; q < kept; q += blockDim.x) {
            const unsigned long long k = keys[q];
            longlong2 v; v.x = (long long)(k >> 16); v.y = (long long)(k & 0xffff);
            *(longlong2*)(out + 2 * q) = v;
        }
        if (tid == 0 && counts) counts[b] = kept;
        return;
    }

    // over the cap: key = (~orderable(score product) << 32) | (row-major rank) ; ascending sort gives
    // product descending, rank ascending.  The (i,j) payload is recovered from a second array.
    int npow = 1;
    while (npow < kept) npow <<= 1;
    unsigned int* payload = (unsigned int*)(keys + kMaxCand);  // [kMaxCand] (i<<16 | j) by rank
    for (int q = tid; q < npow; q += blockDim.x) {
        unsigned long long key = ~0ull;
        if (q < kept) {
            const unsigned long long k = keys[q];
            const int i = (int)(k >> 16), j = (int)(k & 0xffff);
            payload[q] = (unsigned int)k;
            const float prod = __fmul_rn(__ldg(scores + boff + i), __ldg(scores + boff + j));
            unsigned int u = __float_as_uint(prod);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone float -> uint
            key = ((unsigned long long)(~u) << 32) | (unsigned long long)q;
        }
        keys[q] = key;  // slot q is read (above) and rewritten by this thread only
    }
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = tid; q < npow; q += blockDim.x) {
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
    for (int q = tid; q < max_pairs; q += blockDim.x) {
        const unsigned int pl = payload[(unsigned int)(keys[q] & 0xffffffffull)];
        longlong2 v; v.x = (long long)(pl >> 16); v.y = (long long)(pl & 0xffff);
        *(longlong2*)(out + 2 * q) = v;
    }
    if (tid == 0 && counts) counts[b] = max_pairs;
}

__global__ void pairs_globalize_kernel(const int64_t* __restrict__ pairs, int64_t n_pairs,
                                       const int32_t* __restrict__ rel_off, const int32_t* __restrict__ box_off,
                                       int n_images, int32_t* __restrict__ subj, int32_t* __restrict__ obj) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_pairs; r += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_images - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (rel_off[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const longlong2 v = *(const longlong2*)(pairs + 2 * r);
        subj[r] = (int32_t)v.x + box_off[lo];
        obj[r] = (int32_t)v.y + box_off[lo];
    }
}

inline int64_t cap_of(int n, int max_pairs) {
    int64_t c = (int64_t)n * (n - 1);
    if (c > max_pairs) c = max_pairs;
    return c < 1 ? 1 : c;
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_pairs_capacity(const int32_t* n_boxes_host, int n_images, int max_pairs, int64_t* total_rows) {
    VETO_REQUIRE(n_boxes_host && total_rows && n_images >= 0 && max_pairs > 0, VETO_ERR_ARG, "veto_pairs_capacity: bad argument");
    int64_t t = 0;
    for (int b = 0; b < n_images; ++b) {
        VETO_REQUIRE(n_boxes_host[b] >= 0, VETO_ERR_ARG, "veto_pairs_capacity: negative box count");
        t += cap_of(n_boxes_host[b], max_pairs);
    }
    *total_rows = t;
    return VETO_OK;
}

extern "C" int veto_pairs_enumerate(const int32_t* n_boxes_host, int n_images, const float* boxes_dev,
                                    const float* scores_dev, int require_overlap, int max_pairs,
                                    int64_t* pairs_out_dev, int32_t* counts_out_dev, int32_t* scratch_dev,
                                    veto_stream_t stream) {
    set_tag(TAG_PAIRS);
    cudaStream_t s = (cudaStream_t)stream;
    VETO_REQUIRE(n_boxes_host && pairs_out_dev && scratch_dev && n_images >= 0 && max_pairs > 0, VETO_ERR_ARG,
                 "veto_pairs_enumerate: bad argument");
    if (n_images == 0) return VETO_OK;
    // offset tables: [n_boxes | box_off | out_off], each n_images+1 ints
    const int stride = n_images + 1;
    int32_t* host = new int32_t[3 * (size_t)stride];
    bool over_cap = false, too_big = false;
    int64_t boff = 0, ooff = 0;
    for (int b = 0; b < n_images; ++b) {
        const int n = n_boxes_host[b];
        host[b] = n;
        host[stride + b] = (int32_t)boff;
        host[2 * stride + b] = (int32_t)ooff;
        const int64_t cand = (int64_t)n * (n - 1);
        if (cand > max_pairs) over_cap = true;
        if (cand > kMaxCand || n > 65535) too_big = true;
        boff += n;
        ooff += cap_of(n, max_pairs);
    }
    host[n_images] = 0;
    host[stride + n_images] = (int32_t)boff;
    host[2 * stride + n_images] = (int32_t)ooff;
    const bool filtered = require_overlap || over_cap;
    if (ooff > 0x7fffffffLL || (filtered && too_big) || (require_overlap && !boxes_dev) || (over_cap && !scores_dev) ||
        (require_overlap && !counts_out_dev)) {
        delete[] host;
        if (filtered && too_big)
            VETO_REQUIRE(false, VETO_ERR_UNSUPPORTED,
                         "veto_pairs_enumerate: the filtered/capped path holds one image in shared memory (n <= 128)");
        VETO_REQUIRE(false, VETO_ERR_ARG, "veto_pairs_enumerate: boxes/scores/counts pointer missing or too many pairs");
    }
    cudaError_t e = cudaMemcpyAsync(scratch_dev, host, 3 * (size_t)stride * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    delete[] host;  // pageable source: the copy is staged before cudaMemcpyAsync returns
    VETO_CUDA(e);
    const int32_t* d_n = scratch_dev;
    const int32_t* d_boff = scratch_dev + stride;
    const int32_t* d_ooff = scratch_dev + 2 * stride;
    if (!filtered) {
        const int64_t blocks = (ooff + 255) / 256;
        const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
        pairs_dense_kernel<<<grid, 256, 0, s>>>(d_n, d_ooff, n_images, ooff, pairs_out_dev);
        VETO_LAUNCH_CHECK();
        if (counts_out_dev) {
            // counts == capacities, known on the host
            int32_t* hc = new int32_t[n_images];
            for (int b = 0; b < n_images; ++b) hc[b] = (int32_t)cap_of(n_boxes_host[b], max_pairs);
            e = cudaMemcpyAsync(counts_out_dev, hc, n_images * sizeof(int32_t), cudaMemcpyHostToDevice, s);
            delete[] hc;
            VETO_CUDA(e);
        }
        return VETO_OK;
    }
    static DeviceOnce attr_set;
    const int smem = kMaxCand * (int)(sizeof(unsigned long long) + sizeof(unsigned int));
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(pairs_filtered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.done();
    }
    pairs_filtered_kernel<<<n_images, 1024, smem, s>>>(d_n, d_boff, d_ooff, boxes_dev, scores_dev, require_overlap,
                                                      max_pairs, pairs_out_dev, counts_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

extern "C" int veto_pairs_globalize(const int64_t* pairs_dev, int64_t n_pairs, const int32_t* rel_offsets_dev,
                                    const int32_t* box_offsets_dev, int n_images, int32_t* subj_out_dev,
                                    int32_t* obj_out_dev, veto_stream_t stream) {
    if (n_pairs <= 0) return VETO_OK;
    set_tag(TAG_PAIRS);
    VETO_REQUIRE(pairs_dev && rel_offsets_dev && box_offsets_dev && subj_out_dev && obj_out_dev && n_images > 0,
                 VETO_ERR_ARG, "veto_pairs_globalize: bad argument");
    const int64_t blocks = (n_pairs + 255) / 256;
    const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
    pairs_globalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pairs_dev, n_pairs, rel_offsets_dev, box_offsets_dev,
                                                                  n_images, subj_out_dev, obj_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
