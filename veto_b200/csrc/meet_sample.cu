// Training-time group sampling + relabelling of VETOPredictor_MEET on the device.
//
// Reference: VETOPredictor_MEET.forward walks the relation labels pair by pair in Python with one `.item()` device sync
// each (roi_relation_predictors.py:3940-3969) to decide which group heads a pair trains (cur_chosen_matrix), and
// Ensemble.forward relabels the chosen pairs per head with another per-element loop (:3812-3821).  Here one thread
// handles one pair and writes its column of the dense [n_groups, R] table of head-local labels (-1 = the pair is not in
// that head's loss) that veto_relation_train_step consumes; nothing crosses to the host.
//
//   background pair (label 0):  'rand_insert' : one uniformly drawn head          (random.randint(0, G - 1))
//                               'rand_choose' : all heads with probability 0.6    (random.random() >= 0.4)
//                               'all_include' : all heads
//   foreground pair of predicate p (group g_p = incre_idx[p], 1-based), one uniform u: walk a = G .. 1 and stop at the
//   first a with u <= rates[a-1][p] or a < g_p; the pair joins heads 0 .. a-1 (no head if the walk never stops).
//   head-local label: 0 stays 0, a member predicate of the head's group becomes its 1-based position among the members,
//   any other foreground predicate the head's out-of-group class (local_label table, built once on the host).
//
// Draws: counter-based (splitmix64 of seed and pair index, 53-bit uniform like random.random()), or — for parity with a
// seeded run of the reference — injected per pair (`draws`: the u of a foreground / rand_choose pair, `bg_heads`: the
// randint of a rand_insert background pair), taken from Python's `random` stream in the reference's order.
#include "common.cuh"

namespace veto {
namespace {

enum { ZERO_RAND_INSERT = 0, ZERO_RAND_CHOOSE = 1, ZERO_ALL_INCLUDE = 2 };
constexpr int kMaxGroups = 16;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
meet_group_labels_kernel(const int64_t* __restrict__ rel_labels, int64_t n_pairs, const int32_t* __restrict__ incre_idx,
                         const double* __restrict__ rates, const int32_t* __restrict__ local_label, int n_groups, int num_rel,
                         int zero_mode, uint64_t seed, const double* __restrict__ draws, const int32_t* __restrict__ bg_heads,
                         int64_t* __restrict__ out) {
    extern __shared__ double s_rates[];                       // [n_groups, num_rel]
    int32_t* s_local = (int32_t*)(s_rates + n_groups * num_rel);   // [n_groups, num_rel]
    int32_t* s_incre = s_local + n_groups * num_rel;               // [num_rel]
    for (int e = threadIdx.x; e < n_groups * num_rel; e += blockDim.x) {
        s_rates[e] = rates[e];
        s_local[e] = local_label[e];
    }
    for (int e = threadIdx.x; e < num_rel; e += blockDim.x) s_incre[e] = incre_idx[e];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p64 = rel_labels[i];
        const int p = (p64 < 0 || p64 >= num_rel) ? 0 : (int)p64;
        const uint64_t h = mix64(seed + ((uint64_t)i + 1) * 0x9E3779B97F4A7C15ull);
        const double u = draws ? draws[i] : (double)(h >> 11) * (1.0 / 9007199254740992.0);   // [0, 1), 53 bits
        int first = 0, count = 0;                              // the pair joins heads [first, first + count)
        if (p == 0) {
            if (zero_mode == ZERO_RAND_INSERT) {
                // high 32 bits of an independent hash, scaled: uniform over the heads up to 2^-32
                first = bg_heads ? bg_heads[i] : (int)(((mix64(h) >> 32) * (uint64_t)n_groups) >> 32);
                count = 1;
            } else if (zero_mode == ZERO_ALL_INCLUDE || u >= 0.4) {
                count = n_groups;
            }
        } else {
            const int g_p = s_incre[p];
            for (int a = n_groups; a >= 1; --a) {
                if (u <= s_rates[(a - 1) * num_rel + p] || a < g_p) {
                    count = a;
                    break;
                }
            }
        }
        for (int k = 0; k < n_groups; ++k)
            out[(size_t)k * n_pairs + i] = (k >= first && k < first + count) ? (int64_t)s_local[k * num_rel + p] : -1;
    }
}

}  // namespace
}  // namespace veto

using namespace veto;

extern "C" int veto_meet_group_labels(const int64_t* rel_labels_dev, int64_t n_pairs, const int32_t* incre_idx_dev,
                                      const double* rates_dev, const int32_t* local_label_dev, int n_groups, int num_rel,
                                      int zero_mode, uint64_t seed, const double* draws_dev, const int32_t* bg_heads_dev,
                                      int64_t* head_labels_out_dev, veto_stream_t stream) {
    if (n_pairs <= 0) return VETO_OK;
    VETO_REQUIRE(rel_labels_dev && incre_idx_dev && rates_dev && local_label_dev && head_labels_out_dev, VETO_ERR_ARG,
                 "veto_meet_group_labels: NULL argument");
    VETO_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && num_rel >= 2 && num_rel <= 1024, VETO_ERR_UNSUPPORTED,
                 "veto_meet_group_labels: n_groups=%d (1..%d), num_rel=%d (2..1024)", n_groups, kMaxGroups, num_rel);
    VETO_REQUIRE(zero_mode >= ZERO_RAND_INSERT && zero_mode <= ZERO_ALL_INCLUDE, VETO_ERR_ARG,
                 "veto_meet_group_labels: zero_mode %d (0 rand_insert, 1 rand_choose, 2 all_include)", zero_mode);
    set_tag(TAG_PAIRS);
    const int smem = n_groups * num_rel * (int)(sizeof(double) + sizeof(int32_t)) + num_rel * (int)sizeof(int32_t);
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(meet_group_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kMaxGroups * 1024 * 12 + 1024 * 4));
        attr_set.done();
    }
    const int64_t blocks = (n_pairs + 255) / 256;
    const int grid = (int)(blocks < (int64_t)num_sms() * 4 ? blocks : (int64_t)num_sms() * 4);
    meet_group_labels_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(rel_labels_dev, n_pairs, incre_idx_dev, rates_dev,
                                                                        local_label_dev, n_groups, num_rel, zero_mode, seed,
                                                                        draws_dev, bg_heads_dev, head_labels_out_dev);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}
