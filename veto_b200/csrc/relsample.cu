// RelationSampling.gtbox_relsample (pysgg/modeling/roi_heads/relation_head/sampling.py:54-107) — the training-time
// relation sampler of the PredCls / SGCls settings — for a whole batch in one launch, one CTA per image.
//
// The reference loops over the images and, per image, runs about fifteen small torch ops with three `nonzero` host
// syncs and two `randperm`s: foreground pairs = nonzero(relation > 0) in row-major order (a random subset of
// num_pos_per_img of them when there are more), background pairs = every other ordered pair (i != j), randomly
// permuted and cut to batch_size_per_image - num_fg, plus the symmetric binary relatedness matrix.
//
// Here every cell (i, j) of the image's n x n relation matrix gets a 64-bit sort key
//     [ group : 2 | random : 32 | cell index : 16 ]        group 0 = foreground, 1 = background, 3 = diagonal / padding
// with random = a counter-based hash of (seed, image, cell) — zero for the foreground cells when all of them are kept,
// so that they stay in the reference's row-major order — and one bitonic sort in shared memory yields both selections:
// the first fg_kept keys, and the first bg_kept keys after the n_fg foreground ones.  n <= 128 (16384 cells).
#include "stages.cuh"

namespace veto {
namespace {

constexpr int RS_MAX_CELLS = 16384;
constexpr int RS_THREADS = 1024;

__device__ __forceinline__ uint32_t rs_hash(uint64_t seed, uint32_t image, uint32_t cell) {
    uint64_t z = seed + ((uint64_t)image << 32 | cell) * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

__global__ void __launch_bounds__(RS_THREADS)
relsample_gtbox_kernel(const int64_t* __restrict__ rel, const int32_t* __restrict__ mat_off, const int32_t* __restrict__ box_off,
                       int batch_size, int num_pos, uint64_t seed, int64_t* __restrict__ pairs_out,
                       int64_t* __restrict__ labels_out, int32_t* __restrict__ counts_out, int64_t* __restrict__ binary_out) {
    extern __shared__ unsigned long long keys[];  // [npow]
    __shared__ int s_nfg, s_nbg;
    const int b = blockIdx.x;
    const int n = box_off[b + 1] - box_off[b];
    const int cells = n * n;
    const int64_t* m = rel + mat_off[b];
    int64_t* bin = binary_out + mat_off[b];
    if (threadIdx.x == 0) { s_nfg = 0; s_nbg = 0; }
    __syncthreads();
    int npow = 1;
    while (npow < cells) npow <<= 1;
    int my_fg = 0, my_bg = 0;
    for (int e = threadIdx.x; e < cells; e += RS_THREADS) {
        const int i = e / n, j = e - i * n;
        const bool fg = m[e] > 0;
        my_fg += fg;
        my_bg += (!fg && i != j);
        bin[e] = (fg || m[j * n + i] > 0) ? 1 : 0;   // binary_rel[head, tail] = binary_rel[tail, head] = 1 (:79-81)
    }
    atomicAdd(&s_nfg, my_fg);
    atomicAdd(&s_nbg, my_bg);
    __syncthreads();
    const int n_fg = s_nfg, n_bg = s_nbg;
    const bool shuffle_fg = n_fg > num_pos;            // :91-94
    for (int e = threadIdx.x; e < npow; e += RS_THREADS) {
        unsigned long long key = ~0ull;
        if (e < cells) {
            const int i = e / n, j = e - i * n;
            const bool fg = m[e] > 0;
            if (fg) key = ((unsigned long long)(shuffle_fg ? rs_hash(seed, 2 * b, e) : 0u) << 16) | (unsigned)e;
            else if (i != j) key = (1ull << 62) | ((unsigned long long)rs_hash(seed, 2 * b + 1, e) << 16) | (unsigned)e;
        }
        keys[e] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = threadIdx.x; q < npow; q += RS_THREADS) {
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
    const int fg_kept = n_fg < num_pos ? n_fg : num_pos;                    // :95
    const int bg_room = batch_size - fg_kept;                               // :97
    const int bg_kept = n_bg < bg_room ? n_bg : (bg_room > 0 ? bg_room : 0);
    const size_t row0 = (size_t)b * batch_size;
    for (int r = threadIdx.x; r < fg_kept + bg_kept; r += RS_THREADS) {
        const int e = (int)(keys[r < fg_kept ? r : n_fg + (r - fg_kept)] & 0xffffull);
        const int i = e / n, j = e - i * n;
        pairs_out[2 * (row0 + r)] = i;
        pairs_out[2 * (row0 + r) + 1] = j;
        labels_out[row0 + r] = r < fg_kept ? m[e] : 0;
    }
    if (threadIdx.x == 0) {
        counts_out[2 * b] = fg_kept;
        counts_out[2 * b + 1] = fg_kept + bg_kept;
    }
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_relsample_gtbox(const int64_t* rel_matrix_dev, const int32_t* mat_offsets_dev, const int32_t* box_offsets_dev,
                                    const int32_t* n_boxes_host, int n_images, int batch_size_per_image, int num_pos_per_image,
                                    uint64_t seed, int64_t* pairs_out_dev, int64_t* labels_out_dev, int32_t* counts_out_dev,
                                    int64_t* binary_out_dev, veto_stream_t stream) {
    VETO_REQUIRE(n_images >= 0 && batch_size_per_image > 0 && num_pos_per_image >= 0 && num_pos_per_image <= batch_size_per_image,
                 VETO_ERR_ARG, "veto_relsample_gtbox: bad sizes");
    if (n_images == 0) return VETO_OK;
    VETO_REQUIRE(rel_matrix_dev && mat_offsets_dev && box_offsets_dev && n_boxes_host && pairs_out_dev && labels_out_dev &&
                     counts_out_dev && binary_out_dev,
                 VETO_ERR_ARG, "veto_relsample_gtbox: NULL argument");
    int n_max = 0;
    for (int b = 0; b < n_images; ++b) n_max = n_boxes_host[b] > n_max ? n_boxes_host[b] : n_max;
    VETO_REQUIRE(n_max * n_max <= RS_MAX_CELLS, VETO_ERR_UNSUPPORTED,
                 "veto_relsample_gtbox: %d boxes in one image; the n x n relation matrix is sorted in shared memory (n <= 128)", n_max);
    int npow = 1;
    while (npow < n_max * n_max) npow <<= 1;
    const int smem = npow * (int)sizeof(unsigned long long);
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(relsample_gtbox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       RS_MAX_CELLS * (int)sizeof(unsigned long long)));
        attr_set.done();
    }
    set_tag(TAG_PAIRS);
    relsample_gtbox_kernel<<<n_images, RS_THREADS, smem, (cudaStream_t)stream>>>(
        rel_matrix_dev, mat_offsets_dev, box_offsets_dev, batch_size_per_image, num_pos_per_image, seed, pairs_out_dev,
        labels_out_dev, counts_out_dev, binary_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

namespace veto {
namespace {

// =====================================================================================================================
// RelationSampling.detect_relsample + motif_rel_fg_bg_sampling (sampling.py:109-309): the SGDet / SGCls-with-detections
// training sampler, one CTA per image.
//   1. ious[T, P] between ground-truth and detected boxes (+1 convention); is_match = same label and IoU > fg_thres;
//      locating_match[p] = any ground-truth box overlaps detection p by more than fg_thres (:132-141).
//   2. per ground-truth relation i = (head h, tail t, label l), in nonzero() order: the candidates are the detection
//      pairs (a, b), a != b, with a matching h and b matching t (head-major order); all of them leave the background
//      pool and mark the symmetric binary matrix (:216-248); if there are more than num_sample_per_gt_rel, that many
//      are drawn WITHOUT replacement with probability proportional to iou[h, a] * iou[t, b] (:256-261) — here by
//      Efraimidis-Spirakis keys log(u) / w with counter-hash uniforms, which has the same distribution as numpy's
//      sequential weighted draw.
//   3. more than num_pos foreground rows: a uniformly random subset (:271-273).
//   4. background = remaining pairs between foreground-labelled detections (optionally only overlapping ones, :144-153),
//      the 2 * num_neg best by pred_scores[a] * pred_scores[b] (ties: pair index ascending), of which num_neg are drawn
//      uniformly, num_neg = min(batch_size - #fg, #bg) (:283-293).
//   5. nothing at all: two (0, 0, 0) rows (:298-304).
// One warp handles one ground-truth relation; the two selections over the P x P cells reuse the bitonic sort in shared
// memory.  P, T <= 128.
constexpr int RD_THREADS = 1024;
constexpr int RD_MAX_GT_REL = 2048;

__device__ __forceinline__ float rd_iou(const float4 a, const float4 b) {
    const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
    const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
    const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
    const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}
__device__ __forceinline__ unsigned rd_orderable(float v) {  // larger float -> larger unsigned
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void rd_sort(unsigned long long* keys, int npow) {
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

struct DetectParams {
    float fg_thres;
    int require_overlap, per_gt, batch_size, num_pos;
    uint64_t seed;
};

__global__ void __launch_bounds__(RD_THREADS)
relsample_detect_kernel(const float* __restrict__ prp_box, const int64_t* __restrict__ prp_lab, const float* __restrict__ prp_score,
                        const int32_t* __restrict__ prp_off, const float* __restrict__ tgt_box, const int64_t* __restrict__ tgt_lab,
                        const int32_t* __restrict__ tgt_off, const int64_t* __restrict__ tgt_rel, const int32_t* __restrict__ rel_off,
                        const int32_t* __restrict__ bin_off, DetectParams prm, int64_t* __restrict__ triplets,
                        int64_t* __restrict__ corrsp, int32_t* __restrict__ counts, int64_t* __restrict__ binary,
                        float* __restrict__ locating) {
    extern __shared__ unsigned long long keys[];                      // [16384] sort keys
    unsigned char* s_match = (unsigned char*)(keys + RS_MAX_CELLS);    // [T * P]  is_match
    unsigned* s_taken = (unsigned*)(s_match + RS_MAX_CELLS);           // [P * P / 32] cells removed from the background pool
    int* s_gt = (int*)(s_taken + RS_MAX_CELLS / 32);                   // [RD_MAX_GT_REL] cell index of every gt relation
    int* s_kept = s_gt + RD_MAX_GT_REL;                                // [RD_MAX_GT_REL + 1] rows kept per gt relation -> offsets
    unsigned long long* s_fg = (unsigned long long*)(s_kept + RD_MAX_GT_REL + 2);   // [4 * RD_MAX_GT_REL] packed fg rows
    __shared__ int s_n_gt, s_n_bg;
    const int b = blockIdx.x;
    const int p0 = prp_off[b], P = prp_off[b + 1] - p0;
    const int t0 = tgt_off[b], T = tgt_off[b + 1] - t0;
    const int64_t* rel = tgt_rel + rel_off[b];
    int64_t* bin = binary + bin_off[b];
    const float4* pb = (const float4*)prp_box + p0;
    const float4* tb = (const float4*)tgt_box + t0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t row0 = (size_t)b * prm.batch_size;

    // ---- 1. matches, locating_match, zeroed binary matrix
    for (int e = threadIdx.x; e < T * P; e += blockDim.x) {
        const int t = e / P, p = e - t * P;
        s_match[e] = (tgt_lab[t0 + t] == prp_lab[p0 + p] && rd_iou(tb[t], pb[p]) > prm.fg_thres) ? 1 : 0;
    }
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        bool any = false;
        for (int t = 0; t < T; ++t) any |= rd_iou(tb[t], pb[p]) > prm.fg_thres;
        locating[p0 + p] = any ? 1.f : 0.f;
    }
    for (int e = threadIdx.x; e < P * P; e += blockDim.x) bin[e] = 0;
    for (int e = threadIdx.x; e < (P * P + 31) / 32; e += blockDim.x) s_taken[e] = 0u;
    if (threadIdx.x == 0) {  // ground-truth relations in nonzero() (row-major) order
        int n = 0;
        for (int e = 0; e < T * T && n < RD_MAX_GT_REL; ++e)
            if (rel[e] != 0) s_gt[n++] = e;
        s_n_gt = n;
        s_n_bg = 0;
    }
    __syncthreads();
    const int n_gt = s_n_gt;

    // ---- 2. one warp per ground-truth relation: candidates, pool removal, binary marks, weighted sample
    for (int i = wid; i < n_gt; i += nw) {
        const int h = s_gt[i] / T, t = s_gt[i] - h * T;
        const unsigned char* mh = s_match + h * P;
        const unsigned char* mt = s_match + t * P;
        int count = 0;
        for (int e = lane; e < P * P; e += 32) {
            const int a = e / P, c = e - a * P;
            if (mh[a] && mt[c]) {
                bin[a * P + c] = 1;                      // :224-227 (symmetric, includes a == c)
                bin[c * P + a] = 1;
                if (a != c) {
                    atomicOr(&s_taken[e >> 5], 1u << (e & 31));   // :248 rel_possibility[head, tail] = 0
                    ++count;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        const int keep = count < prm.per_gt ? count : prm.per_gt;
        if (lane == 0) s_kept[i] = keep;
        // selection: all candidates in head-major order, or the `per_gt` largest Efraimidis-Spirakis keys
        unsigned long long chosen[4] = {~0ull, ~0ull, ~0ull, ~0ull};   // per_gt <= 4 rounds of warp arg-max
        const bool sample = count > prm.per_gt;
        int written = 0;
        if (!sample) {
            // enumerate in order: rank of a candidate = number of candidates before it
            int base = 0;
            for (int e0 = 0; e0 < P * P; e0 += 32) {
                const int e = e0 + lane;
                bool ok = false;
                if (e < P * P) {
                    const int a = e / P, c = e - a * P;
                    ok = mh[a] && mt[c] && a != c;
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int r = base + __popc(m & ((1u << lane) - 1));
                    s_fg[(size_t)i * 4 + r] = ((unsigned long long)i << 32) | (unsigned)e;
                }
                base += __popc(m);
            }
        } else {
            for (int round = 0; round < prm.per_gt && round < 4; ++round) {
                float best = -INFINITY;
                int beste = 0x7fffffff;
                for (int e = lane; e < P * P; e += 32) {
                    const int a = e / P, c = e - a * P;
                    if (!(mh[a] && mt[c] && a != c)) continue;
                    bool used = false;
                    for (int r = 0; r < round; ++r) used |= (unsigned)(chosen[r] & 0xffffffffull) == (unsigned)e;
                    if (used) continue;
                    const float w = __fmul_rn(rd_iou(tb[h], pb[a]), rd_iou(tb[t], pb[c]));       // :257
                    const float u = ((float)rs_hash(prm.seed, 4 * b + 2, (unsigned)(i * 16384 + e)) + 0.5f) * 2.3283064e-10f;
                    const float key = __fdividef(__logf(u), w);
                    if (key > best || (key == best && e < beste)) { best = key; beste = e; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oe = __shfl_xor_sync(0xffffffffu, beste, o);
                    if (ob > best || (ob == best && oe < beste)) { best = ob; beste = oe; }
                }
                chosen[round] = ((unsigned long long)i << 32) | (unsigned)beste;
                if (lane == 0) s_fg[(size_t)i * 4 + round] = chosen[round];
            }
        }
        (void)written;
    }
    __syncthreads();

    // ---- 3. foreground rows: compact (gt-relation order), cap at num_pos by a random subset
    if (threadIdx.x == 0) {
        int off = 0;
        for (int i = 0; i < n_gt; ++i) {
            const int k = s_kept[i];
            s_kept[i] = off;
            off += k;
        }
        s_kept[n_gt] = off;
    }
    __syncthreads();
    const int n_fg_all = s_kept[n_gt];
    const int n_fg = n_fg_all < prm.num_pos ? n_fg_all : prm.num_pos;
    if (n_fg_all > prm.num_pos) {   // keys = (hash, compact index); n_fg_all <= 4 * RD_MAX_GT_REL <= 8192
        int npow = 1;
        while (npow < n_fg_all) npow <<= 1;
        for (int q = threadIdx.x; q < npow; q += blockDim.x) keys[q] = ~0ull;
        __syncthreads();
        for (int i = threadIdx.x; i < n_gt; i += blockDim.x)
            for (int r = 0; r < s_kept[i + 1] - s_kept[i]; ++r) {
                const int ci = s_kept[i] + r;
                keys[ci] = ((unsigned long long)rs_hash(prm.seed, 4 * b + 3, (unsigned)ci) << 32) | (unsigned)(i * 4 + r);
            }
        __syncthreads();
        rd_sort(keys, npow);
        for (int r = threadIdx.x; r < n_fg; r += blockDim.x) {
            const unsigned long long f = s_fg[keys[r] & 0xffffffffull];
            const int i = (int)(f >> 32), e = (int)(f & 0xffffffffull);
            triplets[3 * (row0 + r)] = e / P;
            triplets[3 * (row0 + r) + 1] = e % P;
            triplets[3 * (row0 + r) + 2] = rel[s_gt[i]];
            corrsp[row0 + r] = i;
        }
    } else {
        for (int i = threadIdx.x; i < n_gt; i += blockDim.x)
            for (int r = 0; r < s_kept[i + 1] - s_kept[i]; ++r) {
                const unsigned long long f = s_fg[(size_t)i * 4 + r];
                const int e = (int)(f & 0xffffffffull);
                const size_t row = row0 + s_kept[i] + r;
                triplets[3 * row] = e / P;
                triplets[3 * row + 1] = e % P;
                triplets[3 * row + 2] = rel[s_gt[i]];
                corrsp[row] = i;
            }
    }
    __syncthreads();

    // ---- 4. background: the best 2 * num_neg by detection quality, then a uniform draw of num_neg
    int npow = 1;
    while (npow < P * P) npow <<= 1;
    int my_bg = 0;
    for (int e = threadIdx.x; e < npow; e += blockDim.x) {
        unsigned long long key = ~0ull;
        if (e < P * P) {
            const int a = e / P, c = e - a * P;
            bool ok = a != c && prp_lab[p0 + a] != 0 && prp_lab[p0 + c] != 0 && !((s_taken[e >> 5] >> (e & 31)) & 1u);
            if (ok && prm.require_overlap) {
                const float v = rd_iou(pb[a], pb[c]);
                ok = v > 0.f && v < 1.f;                                   // :145-146
            }
            if (ok) {
                const float q = __fmul_rn(prp_score[p0 + a], prp_score[p0 + c]);   // :287
                key = ((unsigned long long)(~rd_orderable(q)) << 32) | (unsigned)e;
                ++my_bg;
            }
        }
        keys[e] = key;
    }
    atomicAdd(&s_n_bg, my_bg);
    __syncthreads();
    const int n_bg = s_n_bg;
    int num_neg = prm.batch_size - n_fg;
    num_neg = num_neg < n_bg ? num_neg : n_bg;
    if (num_neg < 0) num_neg = 0;
    if (n_bg > 0) {
        rd_sort(keys, npow);
        int pool = 2 * num_neg < n_bg ? 2 * num_neg : n_bg;                 // :289 [: int(num_neg * 2.0)]
        int npow2 = 1;
        while (npow2 < pool) npow2 <<= 1;
        // re-key the pool by a hash: a uniform draw of num_neg of them (:290-291)
        unsigned long long mine[16];
        int cnt = 0;
        for (int q = threadIdx.x; q < npow2; q += blockDim.x, ++cnt) {
            unsigned long long k = ~0ull;
            if (q < pool) {
                const unsigned e = (unsigned)(keys[q] & 0xffffffffull);
                k = ((unsigned long long)rs_hash(prm.seed, 4 * b + 1, e) << 32) | e;
            }
            mine[cnt] = k;
        }
        __syncthreads();
        cnt = 0;
        for (int q = threadIdx.x; q < npow2; q += blockDim.x, ++cnt) keys[q] = mine[cnt];
        __syncthreads();
        rd_sort(keys, npow2);
        for (int r = threadIdx.x; r < num_neg; r += blockDim.x) {
            const int e = (int)(keys[r] & 0xffffffffull);
            const size_t row = row0 + n_fg + r;
            triplets[3 * row] = e / P;
            triplets[3 * row + 1] = e % P;
            triplets[3 * row + 2] = 0;
            corrsp[row] = -1;
        }
    }
    // ---- 5. nothing sampled at all: two placeholder rows (:298-304)
    int total = n_fg + num_neg;
    if (total == 0) {
        if (threadIdx.x < 2) {
            const size_t row = row0 + threadIdx.x;
            triplets[3 * row] = 0; triplets[3 * row + 1] = 0; triplets[3 * row + 2] = 0;
            corrsp[row] = -1;
        }
        total = 2;
    }
    if (threadIdx.x == 0) {
        counts[2 * b] = n_fg;
        counts[2 * b + 1] = total;
    }
}

}  // namespace
}  // namespace veto

extern "C" int veto_relsample_detect(const float* prp_boxes_dev, const int64_t* prp_labels_dev, const float* prp_scores_dev,
                                     const int32_t* prp_offsets_dev, const float* tgt_boxes_dev, const int64_t* tgt_labels_dev,
                                     const int32_t* tgt_offsets_dev, const int64_t* tgt_rel_dev, const int32_t* rel_offsets_dev,
                                     const int32_t* bin_offsets_dev, const int32_t* n_prp_host, const int32_t* n_tgt_host,
                                     int n_images, float fg_thres, int require_overlap, int num_sample_per_gt_rel,
                                     int batch_size_per_image, int num_pos_per_image, uint64_t seed, int64_t* triplets_out_dev,
                                     int64_t* corrsp_out_dev, int32_t* counts_out_dev, int64_t* binary_out_dev,
                                     float* locating_out_dev, veto_stream_t stream) {
    using namespace veto;
    VETO_REQUIRE(n_images >= 0 && batch_size_per_image >= 2 && num_pos_per_image >= 0 && num_pos_per_image <= batch_size_per_image &&
                     num_sample_per_gt_rel >= 1 && num_sample_per_gt_rel <= 4,
                 VETO_ERR_ARG, "veto_relsample_detect: bad sizes (num_sample_per_gt_rel must be 1..4)");
    if (n_images == 0) return VETO_OK;
    VETO_REQUIRE(prp_boxes_dev && prp_labels_dev && prp_scores_dev && prp_offsets_dev && tgt_boxes_dev && tgt_labels_dev &&
                     tgt_offsets_dev && tgt_rel_dev && rel_offsets_dev && bin_offsets_dev && n_prp_host && n_tgt_host &&
                     triplets_out_dev && corrsp_out_dev && counts_out_dev && binary_out_dev && locating_out_dev,
                 VETO_ERR_ARG, "veto_relsample_detect: NULL argument");
    for (int b = 0; b < n_images; ++b)
        VETO_REQUIRE(n_prp_host[b] <= 128 && n_tgt_host[b] <= 128, VETO_ERR_UNSUPPORTED,
                     "veto_relsample_detect: image %d has %d detections / %d ground-truth boxes (at most 128 each)", b, n_prp_host[b],
                     n_tgt_host[b]);
    const int smem = RS_MAX_CELLS * 8 + RS_MAX_CELLS + RS_MAX_CELLS / 8 + RD_MAX_GT_REL * 4 + (RD_MAX_GT_REL + 2) * 4 +
                     4 * RD_MAX_GT_REL * 8 + 64;
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(relsample_detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.done();
    }
    DetectParams prm{fg_thres, require_overlap, num_sample_per_gt_rel, batch_size_per_image, num_pos_per_image, seed};
    set_tag(TAG_PAIRS);
    relsample_detect_kernel<<<n_images, RD_THREADS, smem, (cudaStream_t)stream>>>(
        prp_boxes_dev, prp_labels_dev, prp_scores_dev, prp_offsets_dev, tgt_boxes_dev, tgt_labels_dev, tgt_offsets_dev, tgt_rel_dev,
        rel_offsets_dev, bin_offsets_dev, prm, triplets_out_dev, corrsp_out_dev, counts_out_dev, binary_out_dev, locating_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
