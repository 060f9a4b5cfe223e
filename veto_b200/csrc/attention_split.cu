// Attention core of the relation encoder (Attention.forward, model_veto.py:86-96) for the inference path, on operands the
// to_qkv GEMM already split: q, k, v arrive as bf16 hi + lo arrays [rows, 1728] (the epilogue of gemm_tc2 writes them in
// that form instead of fp32 — the same bytes), so the kernel stages them with 16-byte cp.async copies and feeds the warp
// MMAs through ldmatrix, with no fp32 -> bf16 conversion of its own.  The round-1 kernel (encoder_ops.cu
// attention_mma2_kernel, still used by the training step and the single-product modes) spent 40 % of its issue slots
// splitting q / k / v into hi + lo pairs after scalar shared-memory reads (profiles/r2_row_kernels_before.txt:
// 49-55 % issue-active at 24 % occupancy, 0.41 of HBM bandwidth).
//
// Work item = one (sequence, head): S = Q K^T (19 x 19, padded to 32 x 24), softmax over the 19 keys, O = P V (19 x 96);
// every product as hi*hi + lo*hi + hi*lo with fp32 accumulation (m16n8k16 bf16); the softmax exponentials are ex2.approx
// of log2(e)-scaled scores (expf was 13 % of the kernel's instructions), the rows 24..31 of the padded tile are skipped.
// Two warps per item: each takes half the k-steps of S (partials meet through a 19 x 20 tile) and half the output column
// blocks of P V.  Rows / keys beyond 19 are not staged: ldmatrix row pointers of the padding rows are clamped onto row 18
// (their scores are masked, their outputs never stored), the padding keys of V point at a row of zeros.
#include "common.cuh"

namespace veto {
namespace {

constexpr int AS_PAIRS = 8;
constexpr int AS_THREADS = AS_PAIRS * 64;
constexpr int AS_PITCH = 104;                        // bf16 elements per staged row: 208 B (ldmatrix rows conflict-free)
constexpr int AS_ARR = kTokens * AS_PITCH;           // elements per staged array
constexpr int AS_TILE = kTokens * 20;                // floats of one partial-score tile
constexpr int AS_ITEM_BYTES = 6 * AS_ARR * 2 + 2 * AS_TILE * 4;
constexpr int AS_SMEM = AS_PAIRS * AS_ITEM_BYTES + 256;   // + one row of zeros
constexpr int AS_OUT_PITCH = 104;                    // floats: the 19 x 96 output tile reuses the Q hi + lo arrays (7904 B)
static_assert(kTokens * AS_OUT_PITCH * 4 <= 2 * AS_ARR * 2, "output tile must fit the Q arrays");

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t* b) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(AS_THREADS, 1)
attention_split_kernel(const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo, int64_t n_seq,
                       float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int fmt, int item_layout) {
    extern __shared__ __align__(16) uint8_t as_smem[];
    constexpr int LD = 3 * kDim;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = threadIdx.x >> 6, w = (threadIdx.x >> 5) & 1, lane64 = threadIdx.x & 63;
    uint8_t* item_base = as_smem + (size_t)pair * AS_ITEM_BYTES;
    __nv_bfloat16* arr = reinterpret_cast<__nv_bfloat16*>(item_base);           // q_hi, q_lo, k_hi, k_lo, v_hi, v_lo
    const __nv_bfloat16* sQh = arr, *sQl = arr + AS_ARR, *sKh = arr + 2 * AS_ARR, *sKl = arr + 3 * AS_ARR;
    const __nv_bfloat16* sVh = arr + 4 * AS_ARR, *sVl = arr + 5 * AS_ARR;
    float* myS = reinterpret_cast<float*>(item_base + 6 * AS_ARR * 2) + w * AS_TILE;
    const float* otherS = reinterpret_cast<float*>(item_base + 6 * AS_ARR * 2) + (w ^ 1) * AS_TILE;
    float* sO = reinterpret_cast<float*>(item_base);                            // over q_hi + q_lo once the scores exist
    const __nv_bfloat16* zero_row = reinterpret_cast<const __nv_bfloat16*>(as_smem + (size_t)AS_PAIRS * AS_ITEM_BYTES);
    if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(as_smem + (size_t)AS_PAIRS * AS_ITEM_BYTES)[threadIdx.x] = 0u;
    __syncthreads();
    const float scale = 0.10206207261596575f * 1.4426950408889634f;  // 96 ** -0.5 (model_veto.py:74) times log2(e)
    const int64_t items = n_seq * kHeads;
    auto pair_bar = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };
    // ldmatrix row / column of this lane inside a 16 x 16 A tile and inside a pair of 8-row B tiles
    const int ld_slot = lane64 / 12, ld_chunk = lane64 - ld_slot * 12;      // staging: row slot 0..4 (5 = idle), 16-byte chunk 0..11
    const int st_slot = lane64 / 24, st_c4 = lane64 - st_slot * 24;         // output: row slot 0..1 (2 = idle), 4-column group 0..23
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;
    // staging of arrays [a0, a1) of an item (0, 1 = q hi / lo; 2, 3 = k; 4, 5 = v): 6 arrays x 19 rows of 12 16-byte chunks,
    // 60 of the pair's 64 lanes own (row slot, chunk) once and for all, so that every copy is two additions away from its
    // addresses (the flat index -> (array, row, chunk) decode of the first version cost a quarter of the kernel's instructions)
    auto stage = [&](int64_t it, int a0, int a1) {
        // item layout (common.cuh qkv_item_offset): the item's q, k, v are three contiguous 19 x 96 blocks — 3.6 KB runs
        // per array instead of 192-byte segments 3456 bytes apart; row-major layout: [rows, 1728]
        const int64_t sq = it / kHeads;
        const size_t gb = item_layout ? (size_t)it * 3 * (kTokens * kHeadDim)
                                      : (size_t)sq * kTokens * LD + (int)(it - sq * kHeads) * kHeadDim;
        const size_t arr_stride = item_layout ? (size_t)(kTokens * kHeadDim) : (size_t)kDim;
        const size_t row_stride = item_layout ? (size_t)kHeadDim : (size_t)LD;
        if (ld_slot < 5) {
            for (int a = a0; a < a1; ++a) {
                const __nv_bfloat16* src = ((a & 1) ? qkv_lo : qkv_hi) + gb + (a >> 1) * arr_stride + ld_chunk * 8;
                __nv_bfloat16* dst = arr + a * AS_ARR + ld_chunk * 8;
#pragma unroll
                for (int r0 = 0; r0 < 20; r0 += 5) {
                    const int row = r0 + ld_slot;
                    if (row < kTokens)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + row * AS_PITCH)),
                                     "l"(src + (size_t)row * row_stride) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t item_stride = (int64_t)gridDim.x * AS_PAIRS;
    const int64_t first = (int64_t)blockIdx.x * AS_PAIRS + pair;
    if (first < items) stage(first, 0, 6);
    for (int64_t item = first; item < items; item += item_stride) {
        const int64_t seq = item / kHeads;
        const int h = (int)(item - seq * kHeads);
        const int64_t next = item + item_stride;
        // the NEXT item's K is fetched as soon as the scores exist, its V after P V, its Q once the output tile (which
        // lives in the Q arrays) is stored: two thirds of an item's bytes arrive under the previous item's arithmetic
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        pair_bar();  // (1) staged operands visible to both warps

        // ---- partial S = Q K^T over this warp's three k-steps: 2 m-tiles x 3 n-tiles
        float S[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) S[mt][nt][e] = 0.f;
#pragma unroll 1
        for (int ks = 3 * w; ks < 3 * w + 3; ++ks) {
            const int k0 = ks * 16;
            uint32_t qh[2][4], ql[2][4], kh[3][2], kl[3][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int row = min(16 * mt + a_row, kTokens - 1);
                ldsm_x4(qh[mt], sQh + row * AS_PITCH + k0 + a_col);
                ldsm_x4(ql[mt], sQl + row * AS_PITCH + k0 + a_col);
            }
            {
                uint32_t t4[4];
                ldsm_x4(t4, sKh + b_row * AS_PITCH + k0 + b_col);
                kh[0][0] = t4[0]; kh[0][1] = t4[1]; kh[1][0] = t4[2]; kh[1][1] = t4[3];
                ldsm_x4(t4, sKl + b_row * AS_PITCH + k0 + b_col);
                kl[0][0] = t4[0]; kl[0][1] = t4[1]; kl[1][0] = t4[2]; kl[1][1] = t4[3];
                const int row2 = min(16 + (lane & 7), kTokens - 1);
                ldsm_x2(kh[2], sKh + row2 * AS_PITCH + k0 + b_col);
                ldsm_x2(kl[2], sKl + row2 * AS_PITCH + k0 + b_col);
            }
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        mma16816(S[mt][nt], term == 1 ? ql[mt] : qh[mt], term == 2 ? kl[nt] : kh[nt]);
        }
        // ---- exchange the partial scores (e 0,1 -> row 16mt+g, e 2,3 -> row +8; cols 8nt+2t+e)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
                    if (row < kTokens && col < kTokens) myS[row * 20 + col] = S[mt][nt][e];
                }
        pair_bar();  // (2) partial scores exchanged; Q and K are no longer read
        if (next < items) stage(next, 2, 4);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
                    if (row < kTokens && col < kTokens) S[mt][nt][e] += otherS[row * 20 + col];
                }

        // ---- softmax over the 19 keys of each row (a row = the 4 lanes of a quad)
        uint32_t ph[2][2][4], pl[2][2][4];  // [mt][ks2][a-fragment]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                if (mt == 1 && hrow == 1) {   // rows 24..31 do not exist: no softmax, zero probabilities
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt) S[mt][nt][2] = S[mt][nt][3] = 0.f;
                    continue;
                }
                float m = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * nt + 2 * t + e;
                        float v = S[mt][nt][2 * hrow + e] * scale;   // scale carries log2(e): the exponentials below are 2^x
                        v = (col < kTokens) ? v : -INFINITY;
                        S[mt][nt][2 * hrow + e] = v;
                        m = fmaxf(m, v);
                    }
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float p;   // ex2.approx: relative error 2^-22, 2^(-inf) = 0 for the padding keys; expf costs 8 instructions
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(S[mt][nt][2 * hrow + e] - m));
                        S[mt][nt][2 * hrow + e] = p;
                        sum += p;
                    }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float inv = 1.f / sum;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) S[mt][nt][2 * hrow + e] *= inv;
            }
            split_pair(S[mt][0][0], S[mt][0][1], ph[mt][0][0], pl[mt][0][0]);
            split_pair(S[mt][0][2], S[mt][0][3], ph[mt][0][1], pl[mt][0][1]);
            split_pair(S[mt][1][0], S[mt][1][1], ph[mt][0][2], pl[mt][0][2]);
            split_pair(S[mt][1][2], S[mt][1][3], ph[mt][0][3], pl[mt][0][3]);
            split_pair(S[mt][2][0], S[mt][2][1], ph[mt][1][0], pl[mt][1][0]);
            split_pair(S[mt][2][2], S[mt][2][3], ph[mt][1][1], pl[mt][1][1]);
            ph[mt][1][2] = pl[mt][1][2] = ph[mt][1][3] = pl[mt][1][3] = 0u;  // keys 24..31 do not exist
        }

        // ---- O = P V : this warp's 6 n-tiles of 8 head dims (blocks 2w, 2w+1 of three), 2 k-steps of 16 keys
#pragma unroll 1
        for (int blk = 2 * w; blk < 2 * w + 2; ++blk) {
            float O[2][3][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) O[mt][j][e] = 0.f;
#pragma unroll
            for (int ks2 = 0; ks2 < 2; ++ks2) {
                // B fragments of V[key][d] through ldmatrix.trans: rows = keys 16 ks2 + (0..15), columns = 8 head dims
                const int key = 16 * ks2 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int d0 = 8 * 3 * blk;
                uint32_t vh[3][2], vl[3][2];
                {
                    uint32_t t4[4];
                    const int dcol = d0 + (lane >> 4) * 8;     // matrices 2, 3: the next 8 head dims
                    const __nv_bfloat16* ph_ = key < kTokens ? sVh + key * AS_PITCH + dcol : zero_row;
                    const __nv_bfloat16* pl_ = key < kTokens ? sVl + key * AS_PITCH + dcol : zero_row;
                    ldsm_x4_trans(t4, ph_);
                    vh[0][0] = t4[0]; vh[0][1] = t4[1]; vh[1][0] = t4[2]; vh[1][1] = t4[3];
                    ldsm_x4_trans(t4, pl_);
                    vl[0][0] = t4[0]; vl[0][1] = t4[1]; vl[1][0] = t4[2]; vl[1][1] = t4[3];
                    const __nv_bfloat16* ph2 = key < kTokens ? sVh + key * AS_PITCH + d0 + 16 : zero_row;
                    const __nv_bfloat16* pl2 = key < kTokens ? sVl + key * AS_PITCH + d0 + 16 : zero_row;
                    ldsm_x2_trans(vh[2], ph2);
                    ldsm_x2_trans(vl[2], pl2);
                }
#pragma unroll
                for (int term = 0; term < 3; ++term)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
                            mma16816(O[mt][j], term == 1 ? pl[mt][ks2] : ph[mt][ks2], term == 2 ? vl[j] : vh[j]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int row = 16 * mt + g + 8 * hrow;
                    if (row < kTokens) {
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            *(float2*)(sO + row * AS_OUT_PITCH + 8 * (3 * blk + j) + 2 * t) =
                                make_float2(O[mt][j][2 * hrow], O[mt][j][2 * hrow + 1]);
                    }
                }
        }
        pair_bar();  // (3) the 19 x 96 output tile is complete; V is no longer read
        if (next < items) stage(next, 4, 6);
        for (int row = st_slot; row < kTokens && st_slot < 2; row += 2) {
            const int c4 = st_c4;
            const float4 v = *(const float4*)(sO + row * AS_OUT_PITCH + 4 * c4);
            const size_t o = ((size_t)seq * kTokens + row) * kDim + h * kHeadDim + 4 * c4;
            if (out_f32) *(float4*)(out_f32 + o) = v;
            if (out_hi) {
                if (fmt == FMT_F16C8) {
                    store_act4_f16c8(out_hi, out_lo, o, v);
                } else {
                    uint2 hh, ll;
                    split_pair(v.x, v.y, hh.x, ll.x);
                    split_pair(v.z, v.w, hh.y, ll.y);
                    *(uint2*)(out_hi + o) = hh;
                    if (out_lo) *(uint2*)(out_lo + o) = ll;
                }
            }
        }
        pair_bar();  // (4) the tile is stored before the next item's Q overwrites it
        if (next < items) stage(next, 0, 2);
    }
}

}  // namespace

int attention_seq_split(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, int64_t n_seq, const ActOut& out, bool item_layout,
                        cudaStream_t s) {
    if (n_seq <= 0) return VETO_OK;
    VETO_REQUIRE(qkv_hi && qkv_lo && (out.hi || out.f32), VETO_ERR_ARG, "attention_seq_split: missing argument");
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AS_SMEM));
        attr_set.done();
    }
    const int64_t blocks = (n_seq * kHeads + AS_PAIRS - 1) / AS_PAIRS;
    const int grid = (int)(blocks < (int64_t)num_sms() ? blocks : (int64_t)num_sms());
    attention_split_kernel<<<grid, AS_THREADS, AS_SMEM, s>>>(qkv_hi, qkv_lo, n_seq, out.f32, out.hi, out.lo, out.fmt, item_layout ? 1 : 0);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
