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
    static bool attr_set = false;
    if (!attr_set) {
        VETO_CUDA(cudaFuncSetAttribute(relsample_gtbox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       RS_MAX_CELLS * (int)sizeof(unsigned long long)));
        attr_set = true;
    }
    set_tag(TAG_PAIRS);
    relsample_gtbox_kernel<<<n_images, RS_THREADS, smem, (cudaStream_t)stream>>>(
        rel_matrix_dev, mat_offsets_dev, box_offsets_dev, batch_size_per_image, num_pos_per_image, seed, pairs_out_dev,
        labels_out_dev, counts_out_dev, binary_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
