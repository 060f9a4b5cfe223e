// Attention core of the relation encoder (Attention.forward, model_veto.py:86-96) on the 5th-generation tensor cores.
//
// The 19-token sequences are far too small for a 128-row tcgen05 tile one at a time, so a work unit is SIX consecutive
// sequences (114 token rows, padded to 128) of one head:
//     S = Q K^T   (128 x 128, K = 96)   — block-diagonal: only the 19 x 19 blocks on the diagonal are meaningful
//     P = softmax(scale * S) restricted to each row's own sequence (everything else is written as exact zeros)
//     O = P V     (128 x 96,  K = 128)
// Five sixths of the S / P products are wasted, and it is still ~5x cheaper than mma.sync (profiles/r1_rows_ncu.txt: the
// warp-MMA version ran at 19 % tensor-pipe activity and 120 us per 2048 sequences): tcgen05 does 4096 MACs/clk/SM.
//
// Data path per unit: all 8 warps read q, k, v (fp32, coalesced) from the QKV buffer, split every value into bf16
// hi + lo and store the tiles into shared memory in the K-major SWIZZLE_128B layout UMMA expects (V is transposed on the
// way: the B operand of P V needs the key index contiguous); one thread issues the MMAs (3 per product in split mode:
// hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM); warps 0-3 own one S row per thread for the softmax (tcgen05.ld of the
// 64-column window that contains the row's sequence) and write P hi / lo over the Q tiles; all warps drain O.
#include <cuda.h>

#include "common.cuh"
#define VETO_TC_KERNEL "attention_tc"
#include "tcgen05.cuh"

namespace veto {
namespace {
using namespace tc;

constexpr int SEQ_PER_UNIT = 6;
constexpr int UNIT_ROWS = SEQ_PER_UNIT * kTokens;  // 114
constexpr int THREADS = 256;
constexpr int ATOM_QK = 128 * 128;                 // one 64-wide K block of a 128-row tile: 16 KB
constexpr int ATOM_VT = kHeadDim * 128;            // 96 rows (head dims) x 64 keys: 12 KB
constexpr int OFF_Q_HI = 0, OFF_Q_LO = 2 * ATOM_QK, OFF_K_HI = 4 * ATOM_QK, OFF_K_LO = 6 * ATOM_QK;
constexpr int OFF_VT_HI = 8 * ATOM_QK, OFF_VT_LO = OFF_VT_HI + 2 * ATOM_VT;
constexpr int OFF_BARS = OFF_VT_LO + 2 * ATOM_VT;  // 180224
constexpr int SMEM_BYTES = OFF_BARS + 64 + 1024;
constexpr int TMEM_COLS = 256;                     // S: columns [0,128), O: [128,224)
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads smem through it)

// K-major SWIZZLE_128B smem matrix descriptor (see gemm_tc.cu): SBO = 1024 B, version 1, layout type 2
// byte offset of 16-byte chunk `c` (0..7) of row `row` inside a K-major SWIZZLE_128B tile (128-byte rows, the
// chunk index XOR-ed with the row's position in its 8-row / 1024-byte group — the pattern TMA writes)
__device__ __forceinline__ uint32_t sw128(int row, int c) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    split_pair(v[0], v[1], hi.x, lo.x);
    split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z);
    split_pair(v[6], v[7], hi.w, lo.w);
}

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const float* __restrict__ qkv, int64_t n_seq, float* out_f32, __nv_bfloat16* out_hi,
                    __nv_bfloat16* out_lo) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared array itself: a round trip through uintptr_t makes every
    // access through the result a GENERIC load / store (LD.E / ST.E instead of LDS / STS — the epilogue staging paid for it)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar_s = (uint64_t*)(smem + OFF_BARS);
    uint64_t* bar_p = bar_s + 1;
    uint64_t* bar_o = bar_s + 2;
    uint32_t* tmem_slot = (uint32_t*)(bar_s + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int LD = 3 * kDim;
    const float scale = 0.10206207261596575f;  // 96 ** -0.5 (model_veto.py:74)
    const int64_t total_rows = n_seq * kTokens;
    const int64_t n_groups = (n_seq + SEQ_PER_UNIT - 1) / SEQ_PER_UNIT;
    const int64_t n_units = n_groups * kHeads;

    if (tid == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 128);
        mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    // V^T keys 120..127 (chunk 7 of key block 1) are never written per unit: zero them once (0 * garbage must be 0)
    for (int d = tid; d < kHeadDim; d += THREADS) {
        *(uint4*)(smem + OFF_VT_HI + ATOM_VT + sw128(d, 7)) = make_uint4(0, 0, 0, 0);
        *(uint4*)(smem + OFF_VT_LO + ATOM_VT + sw128(d, 7)) = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    uint32_t it = 0;
    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int64_t grp = unit / kHeads;
        const int h = (int)(unit - grp * kHeads);
        const int64_t row0 = grp * UNIT_ROWS;
        const uint32_t parity = it & 1;
        const float* base = qkv + (size_t)row0 * LD + h * kHeadDim;

        // ---------------- phase 1: q, k -> K-major tiles; v -> transposed K-major tile (bf16 hi / lo) ----------------
        constexpr int QK_TASKS = 2 * UNIT_ROWS * (kHeadDim / 8);  // 2736: (q|k, token, 8-wide chunk of head dims)
#pragma unroll 2
        for (int task = tid; task < QK_TASKS; task += THREADS) {
            const int which = task / (UNIT_ROWS * 12);
            const int rem = task - which * (UNIT_ROWS * 12);
            const int t = rem / 12, c12 = rem - t * 12;
            float v[8];
            if (row0 + t < total_rows) {
                const float4* src = (const float4*)(base + (size_t)t * LD + which * kDim + c12 * 8);
                const float4 a = __ldg(src), b = __ldg(src + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            uint4 hi, lo;
            split8(v, hi, lo);
            const uint32_t off = (uint32_t)((c12 >> 3) * ATOM_QK) + sw128(t, c12 & 7);
            *(uint4*)(smem + (which ? OFF_K_HI : OFF_Q_HI) + off) = hi;
            if (SPLIT) *(uint4*)(smem + (which ? OFF_K_LO : OFF_Q_LO) + off) = lo;
        }
        constexpr int KEY_BLOCKS = (UNIT_ROWS + 7) / 8;  // 15
        constexpr int VT_TASKS = kHeadDim * KEY_BLOCKS;  // 1440: (8 consecutive keys, head dim)
#pragma unroll 2
        for (int task = tid; task < VT_TASKS; task += THREADS) {
            const int kb = task / kHeadDim, d = task - kb * kHeadDim;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = kb * 8 + j;
                v[j] = (key < UNIT_ROWS && row0 + key < total_rows) ? __ldg(base + (size_t)key * LD + 2 * kDim + d) : 0.f;
            }
            uint4 hi, lo;
            split8(v, hi, lo);
            const uint32_t off = (uint32_t)((kb >> 3) * ATOM_VT) + sw128(d, kb & 7);
            *(uint4*)(smem + OFF_VT_HI + off) = hi;
            if (SPLIT) *(uint4*)(smem + OFF_VT_LO + off) = lo;
        }
        fence_async_smem();
        __syncthreads();

        // ---------------- phase 2: S = Q K^T ----------------
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc_s = make_idesc(128, 128);
            constexpr int PASSES = SPLIT ? 3 : 1;
#pragma unroll
            for (int pass = 0; pass < PASSES; ++pass) {
                const uint32_t a0 = smem_base + (pass == 1 ? OFF_Q_LO : OFF_Q_HI);
                const uint32_t b0 = smem_base + (pass == 2 ? OFF_K_LO : OFF_K_HI);
#pragma unroll
                for (int ks = 0; ks < kHeadDim / 16; ++ks) {
                    const uint32_t o = (uint32_t)((ks >> 2) * ATOM_QK + (ks & 3) * 32);
                    umma_bf16(tmem_base, make_smem_desc(a0 + o), make_smem_desc(b0 + o), idesc_s, (pass | ks) != 0 ? 1u : 0u);
                }
            }
            umma_commit(bar_s);
        }

        // ---------------- phase 3: row softmax, P (bf16 hi / lo) over the Q tiles ----------------
        if (warp < 4) {
            mbar_wait(bar_s, parity, 1);
            __syncwarp();  // lane 0 of warp 0 arrives late (it issued the MMAs): tcgen05.ld is warp-collective
            tc_fence_after();
            const int r = tid;                                   // S row == TMEM lane
            const int c0 = warp == 0 ? 0 : warp == 1 ? 16 : warp == 2 ? 56 : 64;  // 64-column window of this warp's rows
            uint32_t raw[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
            tmem_ld32(taddr, raw);
            tmem_ld32(taddr + 32, raw + 32);
            tmem_ld_wait();
            const bool row_ok = r < UNIT_ROWS;
            const int lo_col = (r / kTokens) * kTokens;          // first key of this row's sequence
            float p[64];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int col = c0 + j;
                const bool ok = row_ok && col >= lo_col && col < lo_col + kTokens;
                p[j] = ok ? __uint_as_float(raw[j]) * scale : -INFINITY;
                m = fmaxf(m, p[j]);
            }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                p[j] = (p[j] == -INFINITY) ? 0.f : __expf(p[j] - m);
                sum += p[j];
            }
            const float inv = row_ok ? 1.f / sum : 0.f;
            const int wc = c0 >> 3;                              // first 8-key chunk of the window
#pragma unroll
            for (int jw = 0; jw < 8; ++jw) {                     // the window: this row's probabilities (zeros elsewhere)
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = p[jw * 8 + e] * inv;
                uint4 hi, lo;
                split8(v, hi, lo);
                const int ck = wc + jw;
                const uint32_t off = (uint32_t)((ck >> 3) * ATOM_QK) + sw128(r, ck & 7);
                *(uint4*)(smem + OFF_Q_HI + off) = hi;
                if (SPLIT) *(uint4*)(smem + OFF_Q_LO + off) = lo;
            }
#pragma unroll
            for (int z = 0; z < 8; ++z) {                        // the other 8 chunks of the 128-key row: exact zeros
                const int ck = z < wc ? z : z + 8;
                const uint32_t off = (uint32_t)((ck >> 3) * ATOM_QK) + sw128(r, ck & 7);
                *(uint4*)(smem + OFF_Q_HI + off) = make_uint4(0, 0, 0, 0);
                if (SPLIT) *(uint4*)(smem + OFF_Q_LO + off) = make_uint4(0, 0, 0, 0);
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p);
        }

        // ---------------- phase 4: O = P V ----------------
        if (tid == 0) {
            mbar_wait(bar_p, parity, 2);
            tc_fence_after();
            constexpr uint32_t idesc_o = make_idesc(128, kHeadDim);
            constexpr int PASSES = SPLIT ? 3 : 1;
#pragma unroll
            for (int pass = 0; pass < PASSES; ++pass) {
                const uint32_t a0 = smem_base + (pass == 1 ? OFF_Q_LO : OFF_Q_HI);
                const uint32_t b0 = smem_base + (pass == 2 ? OFF_VT_LO : OFF_VT_HI);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t oa = (uint32_t)((ks >> 2) * ATOM_QK + (ks & 3) * 32);
                    const uint32_t ob = (uint32_t)((ks >> 2) * ATOM_VT + (ks & 3) * 32);
                    umma_bf16(tmem_base + 128, make_smem_desc(a0 + oa), make_smem_desc(b0 + ob), idesc_o, (pass | ks) != 0 ? 1u : 0u);
                }
            }
            umma_commit(bar_o);
        }

        // ---------------- phase 5: drain O (warps 0-3: head dims 0..47, warps 4-7: 48..95) ----------------
        mbar_wait(bar_o, parity, 3);
        __syncwarp();
        tc_fence_after();
        {
            const int q = warp & 3, half = warp >> 2;
            const int r = q * 32 + lane;
            const bool ok = r < UNIT_ROWS && row0 + r < total_rows;
            const size_t o0 = (size_t)(row0 + r) * kDim + h * kHeadDim + half * 48;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + 128 + half * 48 + i * 16, raw);
                tmem_ld_wait();
                if (ok) {
                    if (out_f32) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *(float4*)(out_f32 + o0 + i * 16 + j) = make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]),
                                                                               __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]));
                    }
                    if (out_hi) {
#pragma unroll
                        for (int j = 0; j < 16; j += 8) {
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(raw[j + e]);
                            uint4 hi, lo;
                            split8(v, hi, lo);
                            *(uint4*)(out_hi + o0 + i * 16 + j) = hi;
                            if (out_lo) *(uint4*)(out_lo + o0 + i * 16 + j) = lo;
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM and the operand tiles are free for the next unit
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

DeviceOnce g_attr_set;

}  // namespace

int attention_tc(const float* qkv, int64_t n_seq, const ActOut& out, cudaStream_t s) {
    if (n_seq <= 0) return VETO_OK;
    VETO_REQUIRE(out.hi || out.f32, VETO_ERR_ARG, "attention_tc: no output");
    if (g_attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        VETO_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        g_attr_set.done();
    }
    const int64_t units = (n_seq + SEQ_PER_UNIT - 1) / SEQ_PER_UNIT * kHeads;
    const int grid = (int)(units < num_sms() ? units : num_sms());
    if (out.lo) attention_tc_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(qkv, n_seq, out.f32, out.hi, out.lo);
    else attention_tc_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(qkv, n_seq, out.f32, out.hi, out.lo);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
