// Token-wise GEMM of the relation encoder, CTA-pair version (tcgen05 cta_group::2): the kernel the encoder layers
// run on.  Same contract as gemm_tc.cu (C[M,N] = epilogue(A[M,K] @ W[N,K]^T), fp32 TMEM accumulate; passes = TC_BF16 /
// TC_BF16X3 / TC_F16C8 / TC_F16 of common.cuh: one bf16 product, the three-product bf16 split, the fp16 product + e4m3
// corrections, one fp16 product), different mapping:
//
//   * a cluster of two CTAs (one SM pair) owns a 256 x 192 (or 256 x 256) output tile; each CTA holds its own 128 rows of
//     A and HALF of the W tile in shared memory, and one tcgen05.mma.cta_group::2 (M=256, N=192 / 256, K=16), issued
//     by the leader CTA, feeds both tensor cores.  The single-CTA kernel is shared-memory-bandwidth bound (every
//     operand byte is written once by TMA and read once per MMA: 2 x 40 KB per 384 tensor cycles > 128 B/clk,
//     profiles/r1_v2_gemm_ncu.txt); the pair halves the W traffic per SM;
//   * in 3-pass mode one pipeline stage holds A_hi, A_lo, W_hi, W_lo of a 64-wide K block and serves all three
//     products (hi*hi, lo*hi, hi*lo): each operand tile is written once instead of twice.
//
// Barrier protocol (all mbarriers live at the same offset in both CTAs):
//   full[s]   (leader)  : TMA of BOTH CTAs completes its bytes on the leader's barrier (peer bit of the address
//                         cleared, as SM100_TMA_2SM_LOAD does); the leader's producer arms it with 2 x stage bytes.
//   empty[s]  (each CTA): tcgen05.commit.cta_group::2 ... multicast 0b11 — both producers see the slot freed.
//   tfull[a]  (each CTA): the same multicast commit after the last K block: both epilogues may drain.
//   tempty[a] (leader)  : 2 x EW epilogue warps arrive (mapa + plain remote mbarrier.arrive, tcgen05.cuh).
// With CL = 4 (opt-in) two pairs share a cluster and the W tile: empty[s] then counts one commit per pair, multicast to all
// four CTAs.  What bounds the kernels and what the template parameters <EPI, BN, CL, ST, EW> buy: DESIGN.md 4b.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include <type_traits>

#include "common.cuh"
#define VETO_TC_KERNEL "gemm_tc2"
#include "tcgen05.cuh"

namespace veto {
namespace {
using namespace tc;

constexpr int BLOCK_M = 128;      // rows per CTA (256 per pair)
// Columns per pair tile: template parameter BN = 192 or 256 (each CTA stages BN / 2 rows of W).  192 divides every width of
// the encoder (576, 1152, 1728); 256 (+ one narrower last tile: 1728 = 6 x 256 + 192, 1152 = 4 x 256 + 128) re-reads the A
// tile 7 / 5 times instead of 9 / 6 and, with 16 epilogue warps, gives every warp 4 chunks instead of 6 / 5 / 5.  W travels in
// 32-row boxes so that one tensor map serves all widths.
constexpr int W_BOX_ROWS = 32;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
// Epilogue warps: three or four per TMEM lane quarter, interleaved over the 16-column chunks of the tile.  The epilogues are
// the critical path of the inference kernels (profiles/r2_modes_epilogue_diag.jsonl: with the epilogue switched off the
// mainloops run at the tensor peak), and each warp walks its chunks as one dependent chain — so the epilogues without a
// residual or GELU (64 - 80 registers: to_qkv) run 16 warps (16 chunks of a 256-wide tile: 4 per warp instead of 6 / 5 / 5,
// to_qkv 90 -> 86 ms per step); the residual ones need up to 128 registers per thread and stay at 12 warps (512 threads).
constexpr int MAX_EPI_WARPS = 16;
__host__ __device__ constexpr int epi_warps(int epi);   // defined after the EPI_* enum
__host__ __device__ constexpr int num_threads(int ew) { return (4 + ew) * 32; }
constexpr int EPI_COLS = 16;
constexpr int EPI_STAGE_BYTES = 32 * EPI_COLS * 4;
constexpr int BYTES_A = BLOCK_M * BLOCK_K * 2;   // 16 KB
__host__ __device__ constexpr int epi_bytes(int ew) { return ew * EPI_STAGE_BYTES; }
constexpr int MAX_STAGES = 6;
constexpr int TMEM_COLS = 512;
template <int BN>
struct TileN {
    static constexpr int kHalf = BN / 2;                         // W rows per CTA of a full-width tile
    static constexpr int kBytesB = kHalf * BLOCK_K * 2;          // 12 KB / 16 KB
    // 3 stages of two-array operands (56 KB / 64 KB each) or 6 stages of single-array operands
    static constexpr int kStageBytes = 2 * (BYTES_A + kBytesB);
    __host__ __device__ static constexpr int pipe(int stages) { return stages * kStageBytes; }
    static constexpr int kPipe = pipe(3);
    // dynamic shared memory of an instance with ew epilogue warps (no more than needed: what is left of the 256 KB is L1,
    // and the residual epilogues re-read their 128-byte lines from it chunk after chunk)
    __host__ __device__ static constexpr int smem(int ew, int stages) { return pipe(stages) + epi_bytes(ew) + 1024 + 256; }
    static_assert(kPipe + epi_bytes(MAX_EPI_WARPS) + 1024 + 256 <= 227 * 1024, "shared memory budget");
    static_assert(2 * BN <= TMEM_COLS, "two accumulator buffers");
};

// arrive on the barrier at the same offset in CTA `rank` of the cluster
// TMA load whose completion bytes go to the LEADER CTA's barrier (peer bit of the shared::cluster address cleared)
// arrive (once all MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair

// K-major SWIZZLE_128B descriptor (see gemm_tc.cu)

struct EpiParams {
    const float* bias;
    const float* residual;
    float* out_f32;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    int act;
    int ldc;
    int ldr;
    // training-branch extras (common.cuh GemmEpilogue)
    float* pre_f32;
    int res_mode;
    DropSpec drop;
    const __nv_bfloat16* res_hi;   // residual in OPERAND format (hi / lo arrays, res_fmt) instead of fp32: the residual stream of
    const __nv_bfloat16* res_lo;   //   the LayerNorm-fused inference path lives in operand format only
    int res_fmt;
    const float2* ln_stats;  // fused LayerNorm on the A operand: (mean, rstd) of row r at ln_stats[r * ln_row_stride]
    const float* ln_c1;      //   and c1[n] = sum_k gamma_k W[n,k]; `bias` then holds c2[n] = sum_k beta_k W[n,k] (+ bias)
    int ln_row_stride;
    const float2* ln_parts;  // alternatively the 9 partial (sum, sum of squares) planes of the rows, reduced here (no finalize launch)
    long long ln_parts_rows;
    float2* stats_partials;  // row statistics of the OUTPUT: partial (sum, sum of squares) [N / 64][M] for ln_stats_finalize
    int out_fmt;             // FMT_BF16 / FMT_F16C8: storage format of out_hi / out_lo
    float acc_scale;         // accumulator scale (f16c8 / f16: 2^-11, the weights are stored times 2048)
    int ksplit;              // K slices (wgrad: tiny output, huge K); tiles enumerate (slice, m, n)
    int kb_per;              // K blocks per slice
    long long split_stride;  // elements between the partial outputs of consecutive slices
    int diag;                // measurement builds only (-DVETO_TC2_DIAG, then VETO_GEMM_DIAG=n at run time; results invalid): 1 = one
                             //   chunk per warp and tile, 2 = no global stores, 3 = TMEM loads only, 4 = no epilogue work
    int w_box;               // rows per TMA box of W (32; 16 for the 4-CTA clusters: a quarter of a 192-wide tile is 48 rows)
};

// Epilogue variants.  EPI_GENERIC evaluates every option of EpiParams at run time (training branch, tests); the others
// are the inference epilogues of the encoder with their options fixed at compile time (the FF1 epilogue — GELU plus the
// operand-format store — was issue-bound: 58 instructions per element in the generic form, profiles/r2_gemm_f16c8_ncu.txt):
//   EPI_F32        C = acc                                               (to_qkv; no bias)
//   EPI_F32_LN     C = rstd * (acc - mean * c1) + c2                     (to_qkv on the raw residual stream: LayerNorm fused)
//   EPI_RES        C = acc + bias + residual                             (to_out, FF2)
//   EPI_RES_OPS    the same, and the result also in operand format + its row statistics (feeds a LayerNorm-fused GEMM)
//   EPI_GELU_OP    operand(gelu(acc + bias))                             (FF1)
//   EPI_GELU_OP_LN operand(gelu(rstd * (acc - mean * c1) + c2))          (FF1 on the raw residual stream)
// LayerNorm fusion: LN(x) W^T = rstd * (x (gamma . W)^T - mean * c1) + c2 — the GEMM runs on the raw rows of x with the
// LayerNorm weight folded into W (veto_pack_weights), the row statistics come from the epilogue that produced x.
//   EPI_OP / EPI_OP_LN   operand(acc) / operand(rstd * (acc - mean * c1) + c2)  (to_qkv for attention_split.cu: q, k, v as
//                        bf16 hi + lo arrays instead of fp32)
//   EPI_RESOP_OPS  acc + bias + residual with the residual READ from operand format and the result written in operand
//                  format only (+ row statistics): x never exists in fp32 between the layers (4.6 KB / row / layer less HBM
//                  traffic for the HBM-bound to_out / FF2 launches; measured effect on the logits: 8.4e-5 -> 8.8e-5)
//   EPI_RESOP_F32  the same residual source, fp32 result (the CLS rows of the last layer)
//   EPI_QKV / EPI_QKV_LN  EPI_OP / EPI_OP_LN with the outputs in the (sequence, head) item layout of common.cuh
//                  qkv_item_offset: q, k, v of an item as three contiguous 19 x 96 blocks for attention_split.cu
enum { EPI_GENERIC = 0, EPI_F32, EPI_F32_LN, EPI_RES, EPI_RES_OPS, EPI_GELU_OP, EPI_GELU_OP_LN, EPI_OP, EPI_OP_LN, EPI_RESOP_OPS,
       EPI_RESOP_F32, EPI_QKV, EPI_QKV_LN, EPI_COUNT };
__host__ __device__ constexpr int epi_warps(int epi) {
    // to_qkv (256-wide tiles: 225 KB of shared memory either way).  The GELU epilogues run 16 warps on their 256-wide
    // instances only (explicit EW argument at the launch): on 192-wide tiles 16 warps cross from 193 KB to 201 KB, i.e. into
    // the next shared-memory configuration of the SM (60 -> 28 KB of L1), which cost FF1 more than the fourth warp per lane
    // quarter gained (70.5 -> 72.3 ms per step).
    return (epi == EPI_F32 || epi == EPI_F32_LN || epi == EPI_OP || epi == EPI_OP_LN || epi == EPI_QKV || epi == EPI_QKV_LN) ? 16 : 12;
}

// four consecutive residual values from the operand format (flat element offset off, off % 4 == 0)
__device__ __forceinline__ float4 load_res_operand(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int fmt, size_t off) {
    const uint2 h = *reinterpret_cast<const uint2*>(hi + off);
    if (fmt == FMT_F16C8) {
        const float2 a = f16x2_to_float(h.x), b = f16x2_to_float(h.y);
        float4 v = make_float4(a.x, a.y, b.x, b.y);
        if (lo) {   // + e4m3(256 r) / 256: the first byte stream of the 64-element block
            const uint32_t rb = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(lo) + c8_byte(off));
            uint32_t r01, r23;
            asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r01) : "h"((unsigned short)(rb & 0xffffu)));
            asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r23) : "h"((unsigned short)(rb >> 16)));
            const float2 ra = f16x2_to_float(r01), rb2 = f16x2_to_float(r23);
            constexpr float inv = 1.f / kC8ActRes;
            v.x = fmaf(ra.x, inv, v.x); v.y = fmaf(ra.y, inv, v.y); v.z = fmaf(rb2.x, inv, v.z); v.w = fmaf(rb2.y, inv, v.w);
        }
        return v;
    }
    float4 v = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xffff0000u), __uint_as_float(h.y << 16),
                           __uint_as_float(h.y & 0xffff0000u));
    if (lo) {
        const uint2 l = *reinterpret_cast<const uint2*>(lo + off);
        v.x += __uint_as_float(l.x << 16); v.y += __uint_as_float(l.x & 0xffff0000u);
        v.z += __uint_as_float(l.y << 16); v.w += __uint_as_float(l.y & 0xffff0000u);
    }
    return v;
}

// the same in two steps: raw words now (12 bytes per four values: small enough to keep two chunks in flight), decode at use
struct RawRes {
    uint2 h;
    uint32_t l;
};
__device__ __forceinline__ RawRes load_res_raw(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int fmt, size_t off) {
    RawRes r;
    r.h = *reinterpret_cast<const uint2*>(hi + off);
    r.l = 0u;
    if (fmt == FMT_F16C8 && lo) r.l = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(lo) + c8_byte(off));
    return r;
}
__device__ __forceinline__ float4 decode_res_f16c8(const RawRes& r) {
    const float2 a = f16x2_to_float(r.h.x), b = f16x2_to_float(r.h.y);
    uint32_t r01, r23;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r01) : "h"((unsigned short)(r.l & 0xffffu)));
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r23) : "h"((unsigned short)(r.l >> 16)));
    const float2 ra = f16x2_to_float(r01), rb = f16x2_to_float(r23);
    constexpr float inv = 1.f / kC8ActRes;
    return make_float4(fmaf(ra.x, inv, a.x), fmaf(ra.y, inv, a.y), fmaf(rb.x, inv, b.x), fmaf(rb.y, inv, b.y));
}

// CL = CTAs per cluster.  2: one CTA pair.  4: two pairs on consecutive row tiles of the same column tile share the W
// tile — every CTA fetches half of its pair-half of W and multicasts it to the CTA of the same rank parity in the other
// pair, so the W bytes cross L2 -> SM once per two row tiles (measured: no gain, the kernels are epilogue-bound; opt-in).
// ST = pipeline stages of two-array operands: 3; 2 for the residual epilogues on K = 576 (to_out) — their kernels are short
// on L1 (what the 256 KB of an SM do not hold as shared memory serves the residual reads): to_out 43.6 -> 40.9 ms per step
// with 137 KB instead of 193 KB of shared memory, while FF2 (K = 1152) needs the third stage (60 -> 74 ms without it).
template <int EPI, int BN, int CL, int ST = 3, int EW = epi_warps(EPI)>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(num_threads(EW), 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                int M, int N, int K, int passes, EpiParams ep) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared array itself: a round trip through uintptr_t makes every
    // access through the result a GENERIC load / store (LD.E / ST.E instead of LDS / STS — the epilogue staging paid for it)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    using TN = TileN<BN>;
    constexpr int NUM_EPI_WARPS = EW;
    constexpr int BYTES_B = TN::kBytesB;
    constexpr int kPipeBytes = TN::pipe(ST);
    uint8_t* smem_epi = smem + kPipeBytes;
    uint64_t* bars = (uint64_t*)(smem_epi + epi_bytes(NUM_EPI_WARPS));
    uint64_t* full_bar = bars;                   // [MAX_STAGES]
    uint64_t* empty_bar = bars + MAX_STAGES;     // [MAX_STAGES]
    uint64_t* tmem_full = bars + 2 * MAX_STAGES; // [2]
    uint64_t* tmem_empty = tmem_full + 2;        // [2] (used in the leader)
    uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    static_assert(CL == 2 || CL == 4, "one or two CTA pairs per cluster");
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1u;        // CTA inside its pair
    const uint32_t prank = crank >> 1;       // pair inside the cluster
    const bool leader = rank == 0;

    const bool split = passes == TC_BF16X3 || passes == TC_F16C8;   // a stage holds hi and lo tiles of both operands
    const int stage_bytes = split ? 2 * (BYTES_A + BYTES_B) : (BYTES_A + BYTES_B);
    const int num_stages = kPipeBytes / stage_bytes;  // 3 (2) stages of two-array operands or 6 (4) of single-array ones
    const int num_m = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const int num_n = (N + BN - 1) / BN;
    const int mn_tiles = num_m * num_n;
    // CL == 4: a cluster walks (pair of row tiles, column tile); no split-K
    const int num_tiles = CL == 4 ? ((num_m + 1) >> 1) * num_n : mn_tiles * ep.ksplit;
    const int num_kb = K / BLOCK_K;
    // tile -> (K slice, m tile, n tile) and the K-block range of the slice
    auto tile_mn = [&](int tile, int& tm, int& tn) {
        if constexpr (CL == 4) {
            const int tm2 = tile / num_n;
            tn = tile - tm2 * num_n;
            tm = 2 * tm2 + (int)prank;   // may be one past the last row tile: TMA zero-fills, nothing is stored
            return 0;
        }
        const int t2 = tile % mn_tiles;
        tm = t2 / num_n;
        tn = t2 - tm * num_n;
        return tile / mn_tiles;
    };
    auto kb_range = [&](int ks, int& kb0, int& kb1) {
        kb0 = ks * ep.kb_per;
        kb1 = kb0 + ep.kb_per < num_kb ? kb0 + ep.kb_per : num_kb;
    };
    const int pair = blockIdx.x / CL, num_pairs = gridDim.x / CL;   // scheduling unit: the cluster

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_w_hi);
        if (split) {
            tma_prefetch_desc(&tm_a_lo);
            tma_prefetch_desc(&tm_w_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], CL / 2);   // one commit per pair that reads the stage
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc2(tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    // stage layout: [A_hi][A_lo][W_hi][W_lo] (3-pass) or [A][W] (1-pass)
    auto stage_ptr = [&](int s) { return smem + s * stage_bytes; };
    const int off_a_lo = BYTES_A;
    const int off_w_hi = split ? 2 * BYTES_A : BYTES_A;
    const int off_w_lo = off_w_hi + BYTES_B;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int tm, tn, kb0, kb1;
                kb_range(tile_mn(tile, tm, tn), kb0, kb1);
                const int m0 = tm * (2 * BLOCK_M) + rank * BLOCK_M;
                const int halfw = min(BN, N - tn * BN) >> 1;      // W rows of this CTA (the last column tile may be narrower)
                const int n0 = tn * BN + rank * halfw;
                const int tx_bytes = (split ? 2 : 1) * (BYTES_A + halfw * BLOCK_K * 2);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int k0 = kb * BLOCK_K;
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * tx_bytes);
                    uint8_t* sp = stage_ptr(stage);
                    tma_load_2d_pair(sp, &tm_a_hi, &full_bar[stage], k0, m0);
                    if (split) tma_load_2d_pair(sp + off_a_lo, &tm_a_lo, &full_bar[stage], k0, m0);
                    if constexpr (CL == 4) {
                        // this CTA's quarter of the column tile, to both CTAs of this rank parity
                        const int qw = halfw >> 1, r0 = (int)prank * qw;
                        const uint16_t mask = (uint16_t)(0x5u << rank);
                        for (int r = r0; r < r0 + qw; r += ep.w_box) {
                            tma_load_2d_pair_mc(sp + off_w_hi + r * (BLOCK_K * 2), &tm_w_hi, &full_bar[stage], k0, n0 + r, mask);
                            if (split)
                                tma_load_2d_pair_mc(sp + off_w_lo + r * (BLOCK_K * 2), &tm_w_lo, &full_bar[stage], k0, n0 + r, mask);
                        }
                    } else {
                        for (int r = 0; r < halfw; r += ep.w_box) {
                            tma_load_2d_pair(sp + off_w_hi + r * (BLOCK_K * 2), &tm_w_hi, &full_bar[stage], k0, n0 + r);
                            if (split) tma_load_2d_pair(sp + off_w_lo + r * (BLOCK_K * 2), &tm_w_lo, &full_bar[stage], k0, n0 + r);
                        }
                    }
                    if (++stage == num_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const bool fp16_ops = passes == TC_F16C8 || passes == TC_F16;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                int tm, tn, kb0, kb1;
                kb_range(tile_mn(tile, tm, tn), kb0, kb1);
                const int width = min(BN, N - tn * BN);
                const uint32_t idesc_fmt0 = make_idesc_fmt0(2 * BLOCK_M, width);   // fp16 (kind::f16) / e4m3 (kind::f8f6f4)
                const uint32_t idesc = fp16_ops ? idesc_fmt0 : make_idesc(2 * BLOCK_M, width);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t sp = smem_u32(stage_ptr(stage));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint32_t ko = k * UMMA_K * 2;
                        const uint64_t a_hi = make_smem_desc(sp + ko);
                        const uint64_t w_hi = make_smem_desc(sp + off_w_hi + ko);
                        umma2_bf16(tmem_d, a_hi, w_hi, idesc, (kb != kb0 || k != 0) ? 1u : 0u);   // kind::f16: bf16 or fp16 by idesc
                        if (passes == TC_BF16X3) {
                            umma2_bf16(tmem_d, make_smem_desc(sp + off_a_lo + ko), w_hi, idesc, 1u);
                            umma2_bf16(tmem_d, a_hi, make_smem_desc(sp + off_w_lo + ko), idesc, 1u);
                        } else if (passes == TC_F16C8) {
                            // 32 bytes of the 128-byte e4m3 rows = 32 of the 2 x 64 correction terms of this K block
                            umma2_f8(tmem_d, make_smem_desc(sp + off_a_lo + ko), make_smem_desc(sp + off_w_lo + ko), idesc_fmt0, 1u);
                        }
                    }
                    umma2_commit_mask(&empty_bar[stage], CL == 4 ? 0xFu : 0x3u);   // every CTA that writes into this stage
                    if (++stage == num_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma2_commit_mask(&tmem_full[acc], 0x3u << (2 * prank));   // the epilogues of this pair
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 rows) — see gemm_tc.cu for the data path ============
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;          // 0..2: which of the three warps of this lane quarter
        constexpr int kStride = NUM_EPI_WARPS / 4;  // chunks c = half, half + 3, ...
        float4* stage4 = reinterpret_cast<float4*>(smem_epi + (warp - 4) * EPI_STAGE_BYTES);
        const int rsub = lane >> 2, cg = lane & 3;
        int it = 0;
        if constexpr (EPI != EPI_GENERIC) {
            constexpr bool kItems = EPI == EPI_QKV || EPI == EPI_QKV_LN;
            constexpr bool kLnIn = EPI == EPI_F32_LN || EPI == EPI_GELU_OP_LN || EPI == EPI_OP_LN || EPI == EPI_QKV_LN;
            constexpr bool kResOp = EPI == EPI_RESOP_OPS || EPI == EPI_RESOP_F32;       // residual read from operand format
            constexpr bool kResid = EPI == EPI_RES || EPI == EPI_RES_OPS || kResOp;
            constexpr bool kGelu = EPI == EPI_GELU_OP || EPI == EPI_GELU_OP_LN;
            constexpr bool kOutF32 = EPI == EPI_F32 || EPI == EPI_F32_LN || EPI == EPI_RES || EPI == EPI_RES_OPS || EPI == EPI_RESOP_F32;
            constexpr bool kOutOp = kGelu || EPI == EPI_RES_OPS || EPI == EPI_OP || EPI == EPI_OP_LN || EPI == EPI_RESOP_OPS || kItems;
            constexpr bool kStats = EPI == EPI_RES_OPS || EPI == EPI_RESOP_OPS;
            static_assert(!kStats || BN == 192, "the row-statistics partials are laid out for 192-wide tiles");
            constexpr bool kBias = kResid || kGelu || kLnIn;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int tm, tn;
                tile_mn(tile, tm, tn);
                const int m0 = tm * (2 * BLOCK_M) + rank * BLOCK_M + q * 32;
                const int n0 = tn * BN;
                const int kChunks = min(BN, N - n0) / EPI_COLS;
                float4 res[4], res_next[4];
                auto load_res = [&](int c, float4 (&dst)[4]) {
                    const int col = n0 + c * EPI_COLS + cg * 4;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = m0 + rr * 8 + rsub;
                        if constexpr (kResOp)
                            dst[rr] = row < M ? load_res_operand(ep.res_hi, ep.res_lo, ep.res_fmt, (size_t)row * ep.ldr + col)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                        else
                            dst[rr] = row < M ? *(const float4*)(ep.residual + (size_t)row * ep.ldr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                float2 st[4];
                float s_acc[4] = {0.f, 0.f, 0.f, 0.f}, q_acc[4] = {0.f, 0.f, 0.f, 0.f};
                if constexpr (kLnIn) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = m0 + rr * 8 + rsub;
                        st[rr] = make_float2(0.f, 0.f);
                        if (row < M) {
                            if (ep.ln_parts) {   // ln_stats_finalize_kernel's arithmetic, in its order
                                const float2* pp = ep.ln_parts + (size_t)row * ep.ln_row_stride;
                                float sv = 0.f, qv = 0.f;
#pragma unroll
                                for (int part = 0; part < kDim / 64; ++part) {
                                    const float2 v = __ldg(pp + (size_t)part * ep.ln_parts_rows);
                                    sv += v.x;
                                    qv += v.y;
                                }
                                const float mean = sv * (1.f / kDim);
                                const float var = fmaxf(qv * (1.f / kDim) - mean * mean, 0.f);
                                st[rr] = make_float2(mean, 1.f / sqrtf(var + 1e-5f));
                            } else {
                                st[rr] = __ldg(ep.ln_stats + (size_t)row * ep.ln_row_stride);
                            }
                        }
                    }
                }
                RawRes raw_a[4], raw_b[4];
                auto load_raw = [&](int c, RawRes (&dst)[4]) {
                    const int col = n0 + c * EPI_COLS + cg * 4;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = min(m0 + rr * 8 + rsub, M - 1);   // rows beyond M are never stored
                        dst[rr] = load_res_raw(ep.res_hi, ep.res_lo, ep.res_fmt, (size_t)row * ep.ldr + col);
                    }
                };
                if constexpr (kResid) {
                    if constexpr (kResOp && BN == 192) {
                        if (ep.res_fmt == FMT_F16C8 && ep.res_lo) {
                            load_raw(half, raw_a);
                            load_raw(half + kStride, raw_b);
                        }
                    } else {
                        load_res(half, res_next);
                    }
                    if (tile + num_pairs < num_tiles) {   // the NEXT tile's residual lines of this warp into L2
                        int ntm, ntn;
                        tile_mn(tile + num_pairs, ntm, ntn);
                        const int pr = ntm * (2 * BLOCK_M) + rank * BLOCK_M + q * 32 + lane;
                        const int pc = ntn * BN;
                        if (pr < M) {
                            if constexpr (kResOp) {
                                // operand format: 128 B of hi and 128 B of lo per 64 columns; this warp third takes block `half`
                                const size_t off = (size_t)pr * ep.ldr + pc + half * 64;
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.res_hi + off));
                                if (ep.res_lo) {
                                    const size_t lo_b = ep.res_fmt == FMT_F16C8 ? c8_byte(off) : off * 2;
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(ep.res_lo) + lo_b));
                                }
                            } else {
                                for (int c = half; c < kChunks; c += kStride)
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.residual + (size_t)pr * ep.ldr + pc + c * EPI_COLS));
                            }
                        }
                    }
                }
                mbar_wait(&tmem_full[acc], acc_phase, 4);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
                const bool rows_full = m0 + 32 <= M, fmt_c8 = ep.out_fmt == FMT_F16C8;
                size_t item_row[kItems ? 4 : 1];   // item layout: the row's part of the offset (its sequence and token)
                if constexpr (kItems) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = m0 + rr * 8 + rsub;
                        const int sq = row / kTokens;
                        item_row[rr] = qkv_item_offset(sq, row - sq * kTokens, 0, 0, 0);
                    }
                }
                auto chunk_rows = [&](int c, const float4& bias4, const float4& c1, auto checked, auto c8fmt) {
                    constexpr bool kCheck = decltype(checked)::value, kC8 = decltype(c8fmt)::value;
                    const int col = n0 + c * EPI_COLS + cg * 4;
                    size_t item_col = 0;   // item layout: the column's part (q / k / v block of its head, head dimension)
                    if constexpr (kItems) {
                        const int which = col / kDim, rem = col - which * kDim, hd = rem / kHeadDim;
                        item_col = qkv_item_offset(0, 0, which, hd, rem - hd * kHeadDim);
                    }
                    float4 vv[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int lr = rr * 8 + rsub;
                        vv[rr] = stage4[lr * 4 + (cg ^ ((lr >> 1) & 3))];
                    }
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = m0 + rr * 8 + rsub;
                        float4 v = vv[rr];
                        if (!kCheck || row < M) {
                            if constexpr (kLnIn) {
                                const float mu = st[rr].x, rs = st[rr].y;
                                v.x = fmaf(rs, fmaf(v.x, ep.acc_scale, -mu * c1.x), bias4.x);
                                v.y = fmaf(rs, fmaf(v.y, ep.acc_scale, -mu * c1.y), bias4.y);
                                v.z = fmaf(rs, fmaf(v.z, ep.acc_scale, -mu * c1.z), bias4.z);
                                v.w = fmaf(rs, fmaf(v.w, ep.acc_scale, -mu * c1.w), bias4.w);
                            } else {
                                v.x = fmaf(v.x, ep.acc_scale, bias4.x); v.y = fmaf(v.y, ep.acc_scale, bias4.y);
                                v.z = fmaf(v.z, ep.acc_scale, bias4.z); v.w = fmaf(v.w, ep.acc_scale, bias4.w);
                            }
                            if constexpr (kGelu) {
                                v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w);
                            }
                            if constexpr (kResid) {
                                v.x += res[rr].x; v.y += res[rr].y; v.z += res[rr].z; v.w += res[rr].w;
                            }
                            size_t off;
                            if constexpr (kItems) off = item_row[rr] + item_col;
                            else off = (size_t)row * ep.ldc + col;
#ifdef VETO_TC2_DIAG
                            const bool st_ok = ep.diag != 2 || v.x == 1.2345678e-30f;
#else
                            constexpr bool st_ok = true;
#endif
                            if constexpr (kOutF32) if (st_ok) *(float4*)(ep.out_f32 + off) = v;
                            if constexpr (kOutOp) if (st_ok) {
                                if constexpr (kC8) {
                                    store_act4_f16c8(ep.out_hi, ep.out_lo, off, v);
                                } else {
                                    uint2 hh, ll;
                                    split_pair(v.x, v.y, hh.x, ll.x);
                                    split_pair(v.z, v.w, hh.y, ll.y);
                                    *(uint2*)(ep.out_hi + off) = hh;
                                    if (ep.out_lo) *(uint2*)(ep.out_lo + off) = ll;
                                }
                            }
                            if constexpr (kStats) {
                                s_acc[rr] += (v.x + v.y) + (v.z + v.w);
                                q_acc[rr] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                            }
                        }
                    }
                };
                auto chunk_body = [&](int c) {
#ifdef VETO_TC2_DIAG
                    if (ep.diag == 4) return;
#endif
                    // the per-column constants first: their global-load latency then hides under the TMEM load and the staging
                    // (issued where they are used they were the largest stall site of the to_qkv epilogue, profiles/
                    // r2_gemm_stall_sites.txt)
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    {
                        const int col = n0 + c * EPI_COLS + cg * 4;
                        if constexpr (kBias) bias4 = __ldg((const float4*)(ep.bias + col));
                        if constexpr (kLnIn) c1 = __ldg((const float4*)(ep.ln_c1 + col));
                    }
                    uint32_t r[16];
                    tmem_ld16(taddr + c * EPI_COLS, r);
                    tmem_ld_wait();
#ifdef VETO_TC2_DIAG
                    if (ep.diag == 3 && r[0] != 0x12345678u) return;
#endif
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        stage4[lane * 4 + (j ^ sw)] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    __syncwarp();
                    // the four row groups of the chunk as straight-line code when the warp's 32 rows all exist and with the
                    // output format fixed at compile time: a per-row `row < M` branch makes every row its own basic block
                    // (shared-memory load -> arithmetic -> conversion -> stores as one exposed chain, four times per chunk).
                    // to_qkv 89.4 -> 86.8 ms per step; NOT for the GELU epilogue (four interleaved GELUs cost FF1 registers
                    // and 3 ms: 70.9 -> 74.0) nor the residual ones (f16c8: no change; bf16x3: to_out 54 -> 64 ms), which keep
                    // the row-by-row form.
                    if (rows_full && !kGelu && !kResid) {
                        if (fmt_c8) chunk_rows(c, bias4, c1, std::false_type(), std::true_type());
                        else chunk_rows(c, bias4, c1, std::false_type(), std::false_type());
                    } else {
                        if (fmt_c8) chunk_rows(c, bias4, c1, std::true_type(), std::true_type());
                        else chunk_rows(c, bias4, c1, std::true_type(), std::false_type());
                    }
                    __syncwarp();
                };
                if constexpr (kResOp && BN == 192) {
                    // f16c8 residual stream: the four chunks of this warp unrolled, raw words two chunks ahead
                    if (ep.res_fmt == FMT_F16C8 && ep.res_lo) {
#pragma unroll
                        for (int i = 0; i < 12 / kStride; ++i) {
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) res[rr] = decode_res_f16c8((i & 1) ? raw_b[rr] : raw_a[rr]);
                            if (i + 2 < 12 / kStride) load_raw(half + (i + 2) * kStride, (i & 1) ? raw_b : raw_a);
                            chunk_body(half + i * kStride);
                        }
                    } else {
#pragma unroll 1
                        for (int c = half; c < kChunks; c += kStride) {
                            load_res(c, res);
                            chunk_body(c);
                        }
                    }
                } else {
#pragma unroll 1
                    for (int c = half; c < kChunks; c += kStride) {
                        if constexpr (kResid) {
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) res[rr] = res_next[rr];
                            if (c + kStride < kChunks) load_res(c + kStride, res_next);
                        }
                        chunk_body(c);
#ifdef VETO_TC2_DIAG
                        if (ep.diag == 1) break;
#endif
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], crank & ~1u);
                if constexpr (kStats) {
                    // this warp covered 64 of the row's columns (4 chunks x 16): one partial per (column tile, warp third)
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        float sv = s_acc[rr], qv = q_acc[rr];
                        sv += __shfl_xor_sync(0xffffffffu, sv, 1); qv += __shfl_xor_sync(0xffffffffu, qv, 1);
                        sv += __shfl_xor_sync(0xffffffffu, sv, 2); qv += __shfl_xor_sync(0xffffffffu, qv, 2);
                        const int row = m0 + rr * 8 + rsub;
                        if (cg == 0 && row < M) ep.stats_partials[(size_t)(tn * kStride + half) * M + row] = make_float2(sv, qv);
                    }
                }
            }
        } else
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            int tm, tn;
            const int ks = tile_mn(tile, tm, tn);
            const int m0 = tm * (2 * BLOCK_M) + rank * BLOCK_M + q * 32;
            const int n0 = tn * BN;
            const int kChunks = min(BN, N - n0) / EPI_COLS;
            const size_t split_off = (size_t)ks * (size_t)ep.split_stride;
            float4 res[4], res_next[4];
            auto load_res = [&](int c, float4 (&dst)[4]) {
                const int col = n0 + c * EPI_COLS + cg * 4;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int row = m0 + rr * 8 + rsub;
                    dst[rr] = (ep.residual && row < M && col < N)
                                  ? *(const float4*)(ep.residual + (size_t)row * ep.ldr + col)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            load_res(half, res_next);
            if (ep.residual && tile + num_pairs < num_tiles) {
                // pull the NEXT tile's residual lines of this warp into L2 while this tile is processed
                int ntm, ntn;
                tile_mn(tile + num_pairs, ntm, ntn);
                const int pr = ntm * (2 * BLOCK_M) + rank * BLOCK_M + q * 32 + lane;
                const int pc = ntn * BN;
                if (pr < M) {
                    for (int c = half; c < kChunks; c += kStride)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.residual + (size_t)pr * ep.ldr + pc + c * EPI_COLS));
                }
            }
            mbar_wait(&tmem_full[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = half; c < kChunks; c += kStride) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) res[rr] = res_next[rr];
                if (c + kStride < kChunks) load_res(c + kStride, res_next);
                uint32_t r[16];
                tmem_ld16(taddr + c * EPI_COLS, r);
                tmem_ld_wait();
                const int sw = (lane >> 1) & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    stage4[lane * 4 + (j ^ sw)] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                __syncwarp();
                const int col = n0 + c * EPI_COLS + cg * 4;
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ep.bias && col < N) bias4 = __ldg((const float4*)(ep.bias + col));
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int lr = rr * 8 + rsub;
                    float4 v = stage4[lr * 4 + (cg ^ ((lr >> 1) & 3))];
                    const int row = m0 + lr;
                    if (row < M && col < N) {
                        v.x = fmaf(v.x, ep.acc_scale, bias4.x); v.y = fmaf(v.y, ep.acc_scale, bias4.y);
                        v.z = fmaf(v.z, ep.acc_scale, bias4.z); v.w = fmaf(v.w, ep.acc_scale, bias4.w);
                        const size_t off = (size_t)row * ep.ldc + col + split_off;
                        if (ep.pre_f32) *(float4*)(ep.pre_f32 + off) = v;
                        if (ep.res_mode == RES_GELU_GRAD) {
                            v.x *= gelu_grad_fast(res[rr].x); v.y *= gelu_grad_fast(res[rr].y);
                            v.z *= gelu_grad_fast(res[rr].z); v.w *= gelu_grad_fast(res[rr].w);
                        } else {
                            v.x = apply_act_tc(v.x, ep.act); v.y = apply_act_tc(v.y, ep.act);
                            v.z = apply_act_tc(v.z, ep.act); v.w = apply_act_tc(v.w, ep.act);
                            if (ep.drop.thr16) {
                                const float4 ds = drop_scale4(ep.drop, off >> 2);
                                v.x *= ds.x; v.y *= ds.y; v.z *= ds.z; v.w *= ds.w;
                            }
                            v.x += res[rr].x; v.y += res[rr].y; v.z += res[rr].z; v.w += res[rr].w;
                        }
                        if (ep.out_f32) *(float4*)(ep.out_f32 + off) = v;
                        if (ep.out_hi) {
                            if (ep.out_fmt == FMT_F16C8) {
                                store_act4_f16c8(ep.out_hi, ep.out_lo, off, v);
                            } else {
                                uint2 hh, ll;
                                split_pair(v.x, v.y, hh.x, ll.x);
                                split_pair(v.z, v.w, hh.y, ll.y);
                                *(uint2*)(ep.out_hi + off) = hh;
                                if (ep.out_lo) *(uint2*)(ep.out_lo + off) = ll;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], crank & ~1u);  // the leader's MMA thread waits for both CTAs
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;
DeviceOnce g_inited;

struct MapKey {
    const void* p;
    uint64_t rows, cols, ld;
    uint32_t box_rows;
    bool operator==(const MapKey& o) const {
        return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        return std::hash<const void*>()(k.p) ^ (k.rows * 0x9E3779B97F4A7C15ull) ^ (k.cols << 20) ^ (k.ld << 7) ^ k.box_rows;
    }
};
// per host thread: no lock on the launch path (descriptors are pure functions of their key)
thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

int get_map(const __nv_bfloat16* p, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out) {
    MapKey key{p, rows, cols, ld, box_rows};
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
        *out = it->second;
        return VETO_OK;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)p, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%llu,%llu] ld %llu box %u at %p", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols, (unsigned long long)ld, box_rows, (const void*)p);
        return VETO_ERR_CUDA;
    }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return VETO_OK;
}

// co-resident 4-CTA clusters of a kernel (a GPC whose SM count is no multiple of 4 leaves SMs out): the persistent grid
int g_clusters4[EPI_COUNT][2] = {};
int max_clusters4(const void* fn, int smem, int threads) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() / 4 * 4);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 4;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int init2() {
    if (!g_inited.pending()) return VETO_OK;
    std::lock_guard<std::mutex> lk(g_mu);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VETO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VETO_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VETO_ERR_CUDA,
                 "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
#define VETO_TC2_ATTR(E, B) \
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<E, B, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileN<B>::smem(epi_warps(E), 3)))
    VETO_TC2_ATTR(EPI_GENERIC, 192); VETO_TC2_ATTR(EPI_F32, 192); VETO_TC2_ATTR(EPI_F32_LN, 192); VETO_TC2_ATTR(EPI_RES, 192);
    VETO_TC2_ATTR(EPI_RES_OPS, 192); VETO_TC2_ATTR(EPI_GELU_OP, 192); VETO_TC2_ATTR(EPI_GELU_OP_LN, 192); VETO_TC2_ATTR(EPI_OP, 192);
    VETO_TC2_ATTR(EPI_OP_LN, 192); VETO_TC2_ATTR(EPI_RESOP_OPS, 192); VETO_TC2_ATTR(EPI_RESOP_F32, 192);
    VETO_TC2_ATTR(EPI_QKV, 192); VETO_TC2_ATTR(EPI_QKV_LN, 192); VETO_TC2_ATTR(EPI_QKV, 256); VETO_TC2_ATTR(EPI_QKV_LN, 256);
    VETO_TC2_ATTR(EPI_GENERIC, 256); VETO_TC2_ATTR(EPI_F32, 256); VETO_TC2_ATTR(EPI_F32_LN, 256);
    VETO_TC2_ATTR(EPI_OP, 256); VETO_TC2_ATTR(EPI_OP_LN, 256);
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_GELU_OP, 256, 2, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileN<256>::smem(16, 3)));
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_GELU_OP_LN, 256, 2, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileN<256>::smem(16, 3)));
#undef VETO_TC2_ATTR
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_RES_OPS, 192, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TileN<192>::smem(epi_warps(EPI_RES_OPS), 2)));
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_RESOP_OPS, 192, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TileN<192>::smem(epi_warps(EPI_RESOP_OPS), 2)));
    // the 4-CTA-cluster instances: the full-sequence launches of the inference encoder
#define VETO_TC2_ATTR4(E, B)                                                                                              \
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<E, B, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileN<B>::smem(epi_warps(E), 3))); \
    g_clusters4[E][B == 256] = max_clusters4((const void*)gemm_tc2_kernel<E, B, 4>, TileN<B>::smem(epi_warps(E), 3), num_threads(epi_warps(E)))
    VETO_TC2_ATTR4(EPI_OP, 256); VETO_TC2_ATTR4(EPI_OP_LN, 256); VETO_TC2_ATTR4(EPI_GELU_OP, 192); VETO_TC2_ATTR4(EPI_GELU_OP_LN, 192);
    VETO_TC2_ATTR4(EPI_RES_OPS, 192); VETO_TC2_ATTR4(EPI_RESOP_OPS, 192);
#undef VETO_TC2_ATTR4
    if (getenv("VETO_GEMM_DEBUG"))
        fprintf(stderr, "veto gemm_tc2: co-resident 4-CTA clusters: to_qkv %d, ff1 %d, to_out / ff2 %d (of %d SMs)\n",
                g_clusters4[EPI_OP_LN][1], g_clusters4[EPI_GELU_OP_LN][0], g_clusters4[EPI_RESOP_OPS][0], num_sms());
    g_inited.done();
    return VETO_OK;
}

}  // namespace

bool gemm_tc2_supported(int N, int K) { return N % 192 == 0 && K % BLOCK_K == 0; }

// 256-wide column tiles (+ one narrower last tile): to_qkv, 1728 = 6 x 256 + 192 (7 tiles instead of 9: 92.3 -> 87.1 ms per
// inference step, profiles/r2_modes_bn256_ab.jsonl) and, with 16 epilogue warps, the inference FF1 (1152 = 4 x 256 + 128:
// 72.3 -> 69.4 ms, profiles/r2_modes_ff1_bn256_ab.jsonl; with 12 warps — 6 / 5 / 5 chunks per warp — it had gained nothing).
// 576 stays 3 x 192 (256 + 256 + 64 would be three tiles as well).
static int tile_width(int N, int ksplit, bool gelu_op_epilogue) {
    static int allow = -1;
    if (allow < 0) {
        const char* e = getenv("VETO_GEMM_BN256");
        allow = (e && e[0] == '0') ? 0 : 1;
    }
    const int rem = N % 256;
    return (allow && ksplit == 1 && N >= (gelu_op_epilogue ? 1152 : 1728) && (rem == 0 || (rem >= 128 && rem % 64 == 0))) ? 256 : 192;
}

static int diag_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VETO_GEMM_DIAG");
        v = e ? atoi(e) : 0;
    }
    return v;
}

static bool shallow_residual(int K, int passes) { return passes == TC_F16C8 && K <= 576; }

// number of K slices a split-K request really produces (every slice non-empty)
int gemm_tc2_slices(int K, int split_k) {
    const int num_kb = K / BLOCK_K;
    int ksplit = split_k > 1 ? split_k : 1;
    if (ksplit > num_kb) ksplit = num_kb;
    if (ksplit < 1) ksplit = 1;
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    return (num_kb + kb_per - 1) / kb_per;
}

int gemm_tc2(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes, const GemmEpilogue& ep,
             cudaStream_t s) {
    if (M <= 0 || N <= 0) return VETO_OK;
    VETO_REQUIRE(passes >= TC_BF16 && passes <= TC_F16, VETO_ERR_ARG, "gemm_tc2: passes must be 1 .. 4 (common.cuh TC_*)");
    VETO_REQUIRE(gemm_tc2_supported(N, K) && K > 0, VETO_ERR_UNSUPPORTED, "gemm_tc2: N=%d must be a multiple of %d, K=%d of %d",
                 N, 192, K, BLOCK_K);
    VETO_REQUIRE(ep.ldc % 4 == 0 && ep.ldr % 4 == 0 && A.ld % 8 == 0 && W.ld % 8 == 0, VETO_ERR_UNSUPPORTED,
                 "gemm_tc2: unaligned strides");
    const bool two_arrays = passes == TC_BF16X3 || passes == TC_F16C8;
    VETO_REQUIRE(A.hi && W.hi && (!two_arrays || (A.lo && W.lo)), VETO_ERR_ARG, "gemm_tc2: missing operand array");
    VETO_REQUIRE(ep.out.fmt == FMT_BF16 || (ep.ldc % 64 == 0 && ep.split_k <= 1), VETO_ERR_ARG,
                 "gemm_tc2: f16c8 outputs need ldc %% 64 == 0");
    int rc = init2();
    if (rc) return rc;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    const uint64_t lda = A.ld ? A.ld : K, ldw = W.ld ? W.ld : K;
    if ((rc = get_map(A.hi, M, K, lda, BLOCK_M, &ta_hi))) return rc;
    if ((rc = get_map(W.hi, N, K, ldw, W_BOX_ROWS, &tw_hi))) return rc;
    ta_lo = ta_hi;
    tw_lo = tw_hi;
    if (two_arrays) {   // the e4m3 byte pairs of f16c8 are addressed as 2-byte elements: the same tensor maps
        if ((rc = get_map(A.lo, M, K, lda, BLOCK_M, &ta_lo))) return rc;
        if ((rc = get_map(W.lo, N, K, ldw, W_BOX_ROWS, &tw_lo))) return rc;
    }
    const int num_kb = K / BLOCK_K;
    const int ksplit = gemm_tc2_slices(K, ep.split_k);
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    VETO_REQUIRE(ksplit == 1 || (!ep.bias && !ep.residual && ep.act == ACT_NONE && !ep.pre_f32 && !ep.drop.thr16 &&
                                 ep.out.f32 && !ep.out.hi),
                 VETO_ERR_ARG, "gemm_tc2: split-K writes plain fp32 partial products only");
    VETO_REQUIRE(ep.res_mode == RES_ADD || ep.residual, VETO_ERR_ARG, "gemm_tc2: RES_GELU_GRAD needs the pre-activation");
    VETO_REQUIRE(!ep.drop.thr16 || ep.ldc % 4 == 0, VETO_ERR_ARG, "gemm_tc2: dropout needs ldc % 4 == 0");
    // the residual epilogues (to_out / FF2: N = 576) are built for 192-wide tiles only
    const bool gelu_op = !ep.pre_f32 && ep.res_mode == RES_ADD && !ep.drop.thr16 && ksplit == 1 && ep.act == ACT_GELU &&
                         !ep.residual && ep.bias && ep.out.hi && !ep.out.f32 && !ep.stats_partials;   // EPI_GELU_OP[_LN] below
    const int bn = (ep.residual || ep.res_op.hi || ep.stats_partials) ? 192 : tile_width(N, ksplit, gelu_op);
    const int num_m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), num_n_tiles = (N + bn - 1) / bn;
    const int tiles = num_m_tiles * num_n_tiles * ksplit;
    static int pair_cap = -1;   // VETO_GEMM_PAIRS=n: diagnosis, run the persistent loop on n CTA pairs only
    if (pair_cap < 0) {
        const char* e = getenv("VETO_GEMM_PAIRS");
        pair_cap = e ? atoi(e) : 0;
    }
    int pairs_avail = num_sms() / 2;
    if (pair_cap > 0 && pair_cap < pairs_avail) pairs_avail = pair_cap;
    int grid = 2 * (tiles < pairs_avail ? tiles : pairs_avail);
    const float acc_scale = (passes == TC_F16C8 || passes == TC_F16) ? kC8AccScale : 1.f;
    EpiParams p{ep.bias, ep.residual, ep.out.f32, ep.out.hi, ep.out.lo, ep.act, ep.ldc, ep.ldr ? ep.ldr : ep.ldc,
                ep.pre_f32, ep.res_mode, ep.drop, ep.res_op.hi, ep.res_op.lo, ep.res_op.fmt, ep.ln_stats, ep.ln_c1,
                ep.ln_row_stride > 0 ? ep.ln_row_stride : 1, ep.ln_parts, ep.ln_parts_rows,
                ep.stats_partials, ep.out.fmt, acc_scale, ksplit, kb_per, (long long)ep.split_stride, diag_mode(), W_BOX_ROWS};
    // the compile-time epilogues of the inference encoder; anything else (training options, tests) is EPI_GENERIC
    const bool plain = !ep.pre_f32 && ep.res_mode == RES_ADD && !ep.drop.thr16 && ksplit == 1;
    const bool ln_in = ep.ln_stats != nullptr || ep.ln_parts != nullptr;
    VETO_REQUIRE(!ep.ln_parts || (K == kDim && ep.ln_parts_rows > 0), VETO_ERR_ARG, "gemm_tc2: statistics partials are per 64 of 576 columns");
    VETO_REQUIRE(!ln_in || (ep.ln_c1 && ep.bias), VETO_ERR_ARG, "gemm_tc2: fused LayerNorm needs ln_c1 and the c2 vector as bias");
    int epi = EPI_GENERIC;
    if (plain && ep.act == ACT_NONE && ep.res_op.hi && !ep.residual && ep.bias && !ln_in) {
        VETO_REQUIRE((ep.ldr ? ep.ldr : ep.ldc) % 64 == 0, VETO_ERR_ARG, "gemm_tc2: operand-format residual needs ldr %% 64 == 0");
        if (ep.out.hi && !ep.out.f32 && ep.stats_partials) epi = EPI_RESOP_OPS;
        else if (ep.out.f32 && !ep.out.hi && !ep.stats_partials) epi = EPI_RESOP_F32;
    } else if (plain && ep.act == ACT_NONE && !ep.residual && ep.out.f32 && !ep.out.hi && !ep.stats_partials) {
        if (ln_in) epi = EPI_F32_LN;
        else if (!ep.bias) epi = EPI_F32;
    } else if (plain && ep.act == ACT_NONE && ep.residual && ep.bias && ep.out.f32 && !ln_in && !ep.res_op.hi) {
        if (ep.out.hi && ep.stats_partials) epi = EPI_RES_OPS;
        else if (!ep.out.hi && !ep.stats_partials) epi = EPI_RES;
    } else if (plain && ep.act == ACT_GELU && !ep.residual && ep.bias && ep.out.hi && !ep.out.f32 && !ep.stats_partials) {
        epi = ln_in ? EPI_GELU_OP_LN : EPI_GELU_OP;
    } else if (plain && ep.act == ACT_NONE && !ep.residual && ep.out.hi && !ep.out.f32 && !ep.stats_partials) {
        if (ln_in) epi = ep.qkv_item_layout ? EPI_QKV_LN : EPI_OP_LN;
        else if (!ep.bias) epi = ep.qkv_item_layout ? EPI_QKV : EPI_OP;
    }
    VETO_REQUIRE(!ep.qkv_item_layout || ((epi == EPI_QKV || epi == EPI_QKV_LN) && N == 3 * kDim && M % kTokens == 0 && ep.out.fmt == FMT_BF16),
                 VETO_ERR_ARG, "gemm_tc2: the item layout is for to_qkv's bf16 hi / lo outputs over whole sequences");
    VETO_REQUIRE(epi != EPI_GENERIC || (!ln_in && !ep.stats_partials && !ep.res_op.hi), VETO_ERR_UNSUPPORTED,
                 "gemm_tc2: LayerNorm fusion / row statistics exist for the inference epilogues only");
    static int force_generic = -1;   // VETO_GEMM_GENERIC_EPI=1: diagnosis, every launch through the run-time epilogue
    if (force_generic < 0) force_generic = getenv("VETO_GEMM_GENERIC_EPI") ? 1 : 0;
    if (force_generic && !ln_in && !ep.stats_partials && !ep.res_op.hi && !ep.qkv_item_layout) epi = EPI_GENERIC;
    // 4-CTA clusters (W multicast between two row tiles) for the big two-array launches of the inference encoder.
    // Measured (profiles/r2_modes_cluster4_ab.jsonl): only 33 clusters of 4 are co-resident on the 148 SMs (132 SMs), the
    // step is 1.5 - 2 % SLOWER (to_qkv 89.7 -> 93.2 ms) although every SM does 7 - 8 % more work per unit time: a quarter
    // less L2 -> SM traffic buys little, the kernels are bound inside the SM (shared-memory pipe: TMA fill + UMMA operand
    // reads + epilogue staging, 58 % busy in profiles/r2_gemm_f16c8_ncu.txt), not by the L2.  Opt-in:
    // VETO_GEMM_CLUSTER4=1 (=2: also for small M, which is how the parity tests exercise the path).
    static int allow4 = -1;
    if (allow4 < 0) {
        const char* e = getenv("VETO_GEMM_CLUSTER4");
        allow4 = !e ? 0 : (e[0] == '2' ? 2 : (e[0] == '1' ? 1 : 0));
    }
    const bool has4 = (bn == 256 && (epi == EPI_OP || epi == EPI_OP_LN)) ||
                      (bn == 192 && (epi == EPI_GELU_OP || epi == EPI_GELU_OP_LN || epi == EPI_RES_OPS || epi == EPI_RESOP_OPS));
    const int clusters4 = has4 ? g_clusters4[epi][bn == 256] : 0;
    if (allow4 && two_arrays && has4 && pair_cap <= 0 && clusters4 >= 16 && (allow4 == 2 || num_m_tiles >= 4 * clusters4)) {
        p.w_box = 16;
        if ((rc = get_map(W.hi, N, K, ldw, 16, &tw_hi))) return rc;
        if ((rc = get_map(W.lo, N, K, ldw, 16, &tw_lo))) return rc;
        const int units = ((num_m_tiles + 1) / 2) * num_n_tiles;
        grid = 4 * (units < clusters4 ? units : clusters4);
#define VETO_TC2_LAUNCH4(E, B) \
    gemm_tc2_kernel<E, B, 4><<<grid, num_threads(epi_warps(E)), TileN<B>::smem(epi_warps(E), 3), s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p)
        switch (epi) {
            case EPI_OP: VETO_TC2_LAUNCH4(EPI_OP, 256); break;
            case EPI_OP_LN: VETO_TC2_LAUNCH4(EPI_OP_LN, 256); break;
            case EPI_GELU_OP: VETO_TC2_LAUNCH4(EPI_GELU_OP, 192); break;
            case EPI_GELU_OP_LN: VETO_TC2_LAUNCH4(EPI_GELU_OP_LN, 192); break;
            case EPI_RES_OPS: VETO_TC2_LAUNCH4(EPI_RES_OPS, 192); break;
            default: VETO_TC2_LAUNCH4(EPI_RESOP_OPS, 192); break;
        }
#undef VETO_TC2_LAUNCH4
        VETO_LAUNCH_CHECK();
        return VETO_OK;
    }
#define VETO_TC2_LAUNCH(E, B) \
    gemm_tc2_kernel<E, B, 2><<<grid, num_threads(epi_warps(E)), TileN<B>::smem(epi_warps(E), 3), s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p)
    // residual epilogues on K <= 576 (to_out) in f16c8 mode: two pipeline stages, the rest of the SM's memory as L1 (16
    // epilogue warps on top of that gained nothing; the bf16x3 products need the third stage)
    const bool shallow = shallow_residual(K, passes);
    if (bn == 256) {
        switch (epi) {
            case EPI_F32: VETO_TC2_LAUNCH(EPI_F32, 256); break;
            case EPI_F32_LN: VETO_TC2_LAUNCH(EPI_F32_LN, 256); break;
            case EPI_GELU_OP: gemm_tc2_kernel<EPI_GELU_OP, 256, 2, 3, 16><<<grid, num_threads(16), TileN<256>::smem(16, 3), s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p); break;
            case EPI_GELU_OP_LN: gemm_tc2_kernel<EPI_GELU_OP_LN, 256, 2, 3, 16><<<grid, num_threads(16), TileN<256>::smem(16, 3), s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p); break;
            case EPI_OP: VETO_TC2_LAUNCH(EPI_OP, 256); break;
            case EPI_OP_LN: VETO_TC2_LAUNCH(EPI_OP_LN, 256); break;
            case EPI_QKV: VETO_TC2_LAUNCH(EPI_QKV, 256); break;
            case EPI_QKV_LN: VETO_TC2_LAUNCH(EPI_QKV_LN, 256); break;
            default: VETO_TC2_LAUNCH(EPI_GENERIC, 256); break;
        }
    } else {
        switch (epi) {
            case EPI_F32: VETO_TC2_LAUNCH(EPI_F32, 192); break;
            case EPI_F32_LN: VETO_TC2_LAUNCH(EPI_F32_LN, 192); break;
            case EPI_RES: VETO_TC2_LAUNCH(EPI_RES, 192); break;
            case EPI_RES_OPS:
                if (shallow) gemm_tc2_kernel<EPI_RES_OPS, 192, 2, 2><<<grid, num_threads(epi_warps(EPI_RES_OPS)), TileN<192>::smem(epi_warps(EPI_RES_OPS), 2), s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p);
                else VETO_TC2_LAUNCH(EPI_RES_OPS, 192);
                break;
            case EPI_GELU_OP: VETO_TC2_LAUNCH(EPI_GELU_OP, 192); break;
            case EPI_GELU_OP_LN: VETO_TC2_LAUNCH(EPI_GELU_OP_LN, 192); break;
            case EPI_OP: VETO_TC2_LAUNCH(EPI_OP, 192); break;
            case EPI_OP_LN: VETO_TC2_LAUNCH(EPI_OP_LN, 192); break;
            case EPI_QKV: VETO_TC2_LAUNCH(EPI_QKV, 192); break;
            case EPI_QKV_LN: VETO_TC2_LAUNCH(EPI_QKV_LN, 192); break;
            case EPI_RESOP_OPS:
                if (shallow) gemm_tc2_kernel<EPI_RESOP_OPS, 192, 2, 2><<<grid, num_threads(epi_warps(EPI_RESOP_OPS)), TileN<192>::smem(epi_warps(EPI_RESOP_OPS), 2), s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p);
                else VETO_TC2_LAUNCH(EPI_RESOP_OPS, 192);
                break;
            case EPI_RESOP_F32: VETO_TC2_LAUNCH(EPI_RESOP_F32, 192); break;
            default: VETO_TC2_LAUNCH(EPI_GENERIC, 192); break;
        }
    }
#undef VETO_TC2_LAUNCH
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

int gemm_tc_auto(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes, const GemmEpilogue& ep,
                 cudaStream_t s) {
    static int use2 = -1;
    if (use2 < 0) {
        const char* e = getenv("VETO_GEMM_2CTA");
        use2 = (e && e[0] == '0') ? 0 : 1;
    }
    // what only the pair kernel implements (the fused inference epilogues, the fp16-based product schemes)
    const bool pair_only = ep.ln_stats || ep.ln_parts || ep.stats_partials || ep.res_op.hi || ep.qkv_item_layout ||
                           ep.out.fmt != FMT_BF16 || passes == TC_F16C8 || passes == TC_F16;
    if ((use2 || pair_only) && gemm_tc2_supported(N, K)) return gemm_tc2(A, W, M, N, K, passes, ep, s);  // any M: one K order
    return gemm_tc(A, W, M, N, K, passes, ep, s);
}

}  // namespace veto
