// Token-wise GEMM of the relation encoder on the 5th-generation tensor cores (sm_100a).
//
//   C[M,N] = epilogue( A[M,K] @ W[N,K]^T )       A, W bf16 (K-major), accumulate fp32 in TMEM
//
// This is the B200 replacement of the cuBLAS SGEMMs behind every nn.Linear of the reference
// encoder (model_veto.py:70-96 to_qkv / to_out, :134-146 FeedForward, :105-106 patch projections).
//
// Layout: persistent CTAs (one per SM) walk 128 x BLOCK_N output tiles; warp 0 is the TMA producer
// (cp.async.bulk.tensor, 128-byte swizzle, a 5-6 stage mbarrier ring), warp 1 issues tcgen05.mma
// (M=128, N=BLOCK_N, K=16 per instruction) into one of two TMEM accumulator stages, warps 4-7 drain
// the other stage with tcgen05.ld and apply bias / GELU / residual before storing.  With passes == 3
// the K loop runs three times over (A_hi,W_hi), (A_lo,W_hi), (A_hi,W_lo): the bf16x3 split that gives
// fp32-grade products on the bf16 pipe (every fp32 operand = hi + lo with 16 mantissa bits kept).
//
// CONV = true is the implicit-GEMM form of a stride-1 k x k convolution on NHWC activations (the depth backbone,
// depth_backbone.cu): an output tile is an 8 x 16 patch of pixels, and the A tile of (tap, 64-channel block) is ONE 4-D
// TMA box {64 channels, 16, 8, 1} of the activation tensor at the tap's offset — the padding is TMA's out-of-bounds
// zero fill, the shared-memory image is the same 128 rows x 128 B swizzled tile a 2-D box gives, so the MMA side does
// not change.  Nothing is materialised: the 9 taps of a pixel are re-read from L2.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#define VETO_TC_KERNEL "gemm_tc"
#include "tcgen05.cuh"

namespace veto {
namespace {
using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B: one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;                       // two per TMEM lane quarter, alternating column chunks
constexpr int NUM_THREADS = (4 + NUM_EPI_WARPS) * 32;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue
constexpr int EPI_COLS = 16;                           // columns per epilogue chunk (one tcgen05.ld x16)
constexpr int EPI_STAGE_BYTES = 32 * EPI_COLS * 4;     // per-warp transpose buffer: 32 rows x 16 fp32

template <int BLOCK_N>
struct TileCfg {
    static constexpr int kBytesA = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBytesB = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kBytesA + kBytesB;
    static constexpr int kStages = (BLOCK_N > 128) ? 5 : 6;
    static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kEpiBytes = NUM_EPI_WARPS * EPI_STAGE_BYTES;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    // CONV: one stage holds A_hi, A_lo, W_hi, W_lo of a (tap, channel block) and serves the three bf16x3 products, so
    // every operand tile crosses L2 -> shared memory once instead of once per product
    static constexpr int kConvStageBytes = 2 * kStageBytes;
    static constexpr int kConvStages = (192 * 1024) / kConvStageBytes;
    static constexpr int kConvSmemBytes = kConvStages * kConvStageBytes + kEpiBytes + 1024 + 256;
    static_assert(kSmemBytes <= 227 * 1024 && kConvSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert((BLOCK_N > 128 || kConvStages >= 3) && kConvStages <= kStages, "conv pipeline depth (barrier arrays are sized by kStages)");
    static_assert((BLOCK_N / EPI_COLS) % 2 == 0, "column chunks must split evenly over the two warps of a quarter");
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------

// Bounded wait: a protocol bug traps (visible as a launch failure) instead of hanging the GPU.



// arrives on `bar` once every tcgen05.mma issued so far by this thread has completed


// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows
// of 128 B) >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B (2) in [61,64).

// kind::f16 instruction descriptor: D fp32 (1<<4), A bf16 (1<<7), B bf16 (1<<10), both K-major,
// N>>3 in [17,23), M>>4 in [24,29).

constexpr int CONV_TH = 8, CONV_TW = 16;  // output patch of one tile (CONV_TH * CONV_TW == BLOCK_M)
struct ConvGeom {
    int H, W, Cin, ks, pad, n_th, n_tw, stride;  // H, W: OUTPUT size; input pixel = output pixel * stride + tap - pad
};

struct EpiParams {
    const float* bias;
    const float* residual;
    float* out_f32;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    int act;
    int ldc;
    int ldr;
};

template <int BLOCK_N, bool CONV>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               int M, int N, int K, int passes, EpiParams ep, ConvGeom cg_) {
    using C = TileCfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    // 1024-byte alignment by pointer arithmetic on the shared array itself: a round trip through uintptr_t makes every
    // access through the result a GENERIC load / store (LD.E / ST.E instead of LDS / STS — the epilogue staging paid for it)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::kStages * C::kBytesA;
    // per-epilogue-warp transpose buffers behind the pipeline (CONV: kConvStages stages of [A_hi][A_lo][W_hi][W_lo])
    uint8_t* smem_epi = smem + (CONV ? C::kConvStages * C::kConvStageBytes : C::kStages * C::kStageBytes);
    uint64_t* bars = (uint64_t*)(smem_epi + C::kEpiBytes);
    uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + C::kStages;      // [kStages]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * C::kStages;  // [2]        MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;         // [2]        epilogue -> MMA
    uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // CONV: M = batch * n_th * n_tw patches of 128 pixels
    const int num_m = CONV ? M : (M + BLOCK_M - 1) / BLOCK_M;
    const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int kb_per_pass = K / BLOCK_K;
    const int num_kb = CONV ? kb_per_pass : kb_per_pass * passes;
    constexpr int kPipe = CONV ? C::kConvStages : C::kStages;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_w_hi);
        if (passes > 1) {
            tma_prefetch_desc(&tm_a_lo);
            tma_prefetch_desc(&tm_w_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_base_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / num_n) * BLOCK_M;
                const int n0 = (tile % num_n) * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int pass = kb / kb_per_pass;
                    const int k0 = (kb - pass * kb_per_pass) * BLOCK_K;
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    mbar_arrive_expect_tx(&full_bar[stage], (CONV && passes == 3) ? C::kConvStageBytes : C::kStageBytes);
                    if (CONV) {
                        const int patch = tile / num_n;
                        const int tw = patch % cg_.n_tw, th = (patch / cg_.n_tw) % cg_.n_th, b = patch / (cg_.n_tw * cg_.n_th);
                        const int tap = k0 / cg_.Cin, c0 = k0 - tap * cg_.Cin;
                        const int kh = tap / cg_.ks, kw = tap - kh * cg_.ks;
                        const int x0 = tw * CONV_TW * cg_.stride + kw - cg_.pad, y0 = th * CONV_TH * cg_.stride + kh - cg_.pad;
                        uint8_t* sp = smem + stage * C::kConvStageBytes;
                        tma_load_4d(sp, &tm_a_hi, &full_bar[stage], c0, x0, y0, b);
                        tma_load_2d(sp + 2 * C::kBytesA, &tm_w_hi, &full_bar[stage], k0, n0);
                        if (passes == 3) {
                            tma_load_4d(sp + C::kBytesA, &tm_a_lo, &full_bar[stage], c0, x0, y0, b);
                            tma_load_2d(sp + 2 * C::kBytesA + C::kBytesB, &tm_w_lo, &full_bar[stage], k0, n0);
                        }
                        if (++stage == kPipe) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    tma_load_2d(smem_a + stage * C::kBytesA, (pass == 1) ? &tm_a_lo : &tm_a_hi, &full_bar[stage], k0, m0);
                    tma_load_2d(smem_b + stage * C::kBytesB, (pass == 2) ? &tm_w_lo : &tm_w_hi, &full_bar[stage], k0, n0);
                    if (++stage == kPipe) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(CONV ? smem + stage * C::kConvStageBytes : smem_a + stage * C::kBytesA);
                    const uint32_t b_addr = CONV ? a_addr + 2 * C::kBytesA : smem_u32(smem_b + stage * C::kBytesB);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t adesc = make_smem_desc(a_addr + k * UMMA_K * 2);
                        const uint64_t bdesc = make_smem_desc(b_addr + k * UMMA_K * 2);
                        umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (CONV && passes == 3) {  // hi*hi + lo*hi + hi*lo from the one stage
                            umma_bf16(tmem_d, make_smem_desc(a_addr + C::kBytesA + k * UMMA_K * 2), bdesc, idesc, 1u);
                            umma_bf16(tmem_d, adesc, make_smem_desc(b_addr + C::kBytesB + k * UMMA_K * 2), idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
                    if (++stage == kPipe) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // Warp w drains TMEM lanes [32q, 32q+32), q = w % 4 (the lane quarter a warp may touch), and the column
        // chunks c = half, half+2, ... of the tile (half = (w-4)/4).  A chunk of 16 accumulator columns arrives with
        // one row per thread; it is transposed through a swizzled 2 KB shared buffer so that the global accesses
        // (residual read, fp32 / bf16 stores) are row-contiguous: lane L handles rows 8*rr + L/4 (rr = 0..3) and the
        // float4 column group L % 4, i.e. every instruction covers 8 rows x 64 contiguous bytes.  The residual of the
        // next chunk is prefetched while the current one is processed.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        float4* stage4 = reinterpret_cast<float4*>(smem_epi + (warp - 4) * EPI_STAGE_BYTES);
        const int rsub = lane >> 2, cg = lane & 3;
        constexpr int kChunks = BLOCK_N / EPI_COLS;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (tile / num_n) * BLOCK_M + q * 32;
            const int n0 = (tile % num_n) * BLOCK_N;
            // CONV: tile row r = pixel (th*8 + r/16, tw*16 + r%16) of image b; rows outside the image are not stored
            int conv_b = 0, conv_h0 = 0, conv_w0 = 0;
            if (CONV) {
                const int patch = tile / num_n;
                conv_w0 = (patch % cg_.n_tw) * CONV_TW;
                conv_h0 = ((patch / cg_.n_tw) % cg_.n_th) * CONV_TH;
                conv_b = patch / (cg_.n_tw * cg_.n_th);
            }
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
            load_res(half, res_next);  // does not depend on the accumulator: overlaps the MMA of this tile
            mbar_wait(&tmem_full[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
            for (int c = half; c < kChunks; c += 2) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) res[rr] = res_next[rr];
                if (c + 2 < kChunks) load_res(c + 2, res_next);
                uint32_t r[16];
                tmem_ld16(taddr + c * EPI_COLS, r);
                tmem_ld_wait();
                // row `lane` -> 4 float4, XOR-swizzled by (row >> 1) & 3: conflict-free both ways
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
                    long long row = m0 + lr;
                    bool row_ok = row < M;
                    if (CONV) {
                        const int r = q * 32 + lr;
                        const int h = conv_h0 + r / CONV_TW, w = conv_w0 + r % CONV_TW;
                        row_ok = h < cg_.H && w < cg_.W;
                        row = ((long long)conv_b * cg_.H + h) * cg_.W + w;
                    }
                    if (row_ok && col < N) {  // N % 4 == 0 is enforced by the host
                        v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                        v.x = apply_act(v.x, ep.act); v.y = apply_act(v.y, ep.act);
                        v.z = apply_act(v.z, ep.act); v.w = apply_act(v.w, ep.act);
                        v.x += res[rr].x; v.y += res[rr].y; v.z += res[rr].z; v.w += res[rr].w;
                        const size_t off = (size_t)row * ep.ldc + col;
                        if (ep.out_f32) *(float4*)(ep.out_f32 + off) = v;
                        if (ep.out_hi) {
                            __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                            split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1);
                            split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
                            *(uint2*)(ep.out_hi + off) = pack_bf16x4(h0, h1, h2, h3);
                            if (ep.out_lo) *(uint2*)(ep.out_lo + off) = pack_bf16x4(l0, l1, l2, l3);
                        }
                    }
                }
                __syncwarp();  // the transpose buffer is reused by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
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

// bf16 row-major [rows, cols] with row stride ld -> tiled map with a {64, box_rows} box and 128-byte swizzle
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
        set_error("cuTensorMapEncodeTiled failed (%d) for [%llu,%llu] box %u at %p", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols, box_rows, (const void*)p);
        return VETO_ERR_CUDA;
    }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return VETO_OK;
}

template <int BLOCK_N>
int launch(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes, const GemmEpilogue& ep,
           cudaStream_t s) {
    using C = TileCfg<BLOCK_N>;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    int rc;
    const uint64_t lda = A.ld ? A.ld : K, ldw = W.ld ? W.ld : K;
    if ((rc = get_map(A.hi, M, K, lda, BLOCK_M, &ta_hi))) return rc;
    if ((rc = get_map(W.hi, N, K, ldw, BLOCK_N, &tw_hi))) return rc;
    ta_lo = ta_hi;
    tw_lo = tw_hi;
    if (passes == 3) {
        if ((rc = get_map(A.lo, M, K, lda, BLOCK_M, &ta_lo))) return rc;
        if ((rc = get_map(W.lo, N, K, ldw, BLOCK_N, &tw_lo))) return rc;
    }
    const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    EpiParams p{ep.bias, ep.residual, ep.out.f32, ep.out.hi, ep.out.lo, ep.act, ep.ldc, ep.ldr ? ep.ldr : ep.ldc};
    gemm_tc_kernel<BLOCK_N, false><<<grid, NUM_THREADS, C::kSmemBytes, s>>>(ta_hi, ta_lo, tw_hi, tw_lo, M, N, K, passes, p,
                                                                           ConvGeom{});
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

// NHWC bf16 activation [B,H,W,C] -> 4-D tiled map, box {64 channels, CONV_TW, CONV_TH, 1} pixels taken every `stride`-th
// (TMA element strides: the box spans CONV_T* x stride elements and loads every stride-th), 128-byte swizzle, zero fill
int get_map_nhwc(const __nv_bfloat16* p, int B, int H, int W, int C, int stride, CUtensorMap* out) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)(CONV_TW * stride), (cuuint32_t)(CONV_TH * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for NHWC [%d,%d,%d,%d] at %p", (int)r, B, H, W, C, (const void*)p);
        return VETO_ERR_CUDA;
    }
    return VETO_OK;
}

template <int BLOCK_N>
int launch_conv(const GemmOperand& A, const GemmOperand& W, int B, int Hin, int Win, int Cin, int N, int ks, int pad, int stride,
                int passes, float* out, int ldc, cudaStream_t s) {
    const int H = (Hin + 2 * pad - ks) / stride + 1, Wd = (Win + 2 * pad - ks) / stride + 1;  // output size
    using C = TileCfg<BLOCK_N>;
    const int K = ks * ks * Cin;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    int rc;
    if ((rc = get_map_nhwc(A.hi, B, Hin, Win, Cin, stride, &ta_hi))) return rc;
    if ((rc = get_map(W.hi, N, K, K, BLOCK_N, &tw_hi))) return rc;
    ta_lo = ta_hi;
    tw_lo = tw_hi;
    if (passes == 3) {
        if ((rc = get_map_nhwc(A.lo, B, Hin, Win, Cin, stride, &ta_lo))) return rc;
        if ((rc = get_map(W.lo, N, K, K, BLOCK_N, &tw_lo))) return rc;
    }
    ConvGeom g{H, Wd, Cin, ks, pad, (H + CONV_TH - 1) / CONV_TH, (Wd + CONV_TW - 1) / CONV_TW, stride};
    const int patches = B * g.n_th * g.n_tw;
    const int tiles = patches * ((N + BLOCK_N - 1) / BLOCK_N);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    EpiParams p{nullptr, nullptr, out, nullptr, nullptr, ACT_NONE, ldc, ldc};
    gemm_tc_kernel<BLOCK_N, true><<<grid, NUM_THREADS, C::kConvSmemBytes, s>>>(ta_hi, ta_lo, tw_hi, tw_lo, patches, N, K, passes, p, g);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace

int gemm_tc_init() {
    if (!g_inited.pending()) return VETO_OK;
    std::lock_guard<std::mutex> lk(g_mu);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VETO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VETO_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VETO_ERR_CUDA,
                 "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<192>::kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<128>::kSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<128>::kConvSmemBytes));
    VETO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<64>::kConvSmemBytes));
    g_inited.done();
    return VETO_OK;
}

int gemm_tc(const GemmOperand& A, const GemmOperand& W, int M, int N, int K, int passes, const GemmEpilogue& ep,
            cudaStream_t s) {
    if (M <= 0 || N <= 0) return VETO_OK;
    VETO_REQUIRE(passes == 1 || passes == 3, VETO_ERR_ARG, "gemm_tc: passes must be 1 or 3");
    VETO_REQUIRE(!ep.pre_f32 && ep.res_mode == RES_ADD && !ep.drop.thr16 && ep.split_k <= 1, VETO_ERR_UNSUPPORTED,
                 "gemm_tc: the training-branch epilogue options exist in gemm_tc2 / gemm_simt only (N=%d)", N);
    VETO_REQUIRE(K % BLOCK_K == 0 && K > 0, VETO_ERR_UNSUPPORTED, "gemm_tc: K=%d must be a positive multiple of %d", K, BLOCK_K);
    VETO_REQUIRE(N % 4 == 0 && ep.ldc % 4 == 0 && ep.ldr % 4 == 0, VETO_ERR_UNSUPPORTED,
                 "gemm_tc: N=%d, ldc=%d and ldr=%d must be multiples of 4", N, ep.ldc, ep.ldr);
    VETO_REQUIRE(A.ld % 8 == 0 && W.ld % 8 == 0, VETO_ERR_UNSUPPORTED, "gemm_tc: operand row strides must be multiples of 8");
    VETO_REQUIRE(A.hi && W.hi && (passes == 1 || (A.lo && W.lo)), VETO_ERR_ARG, "gemm_tc: missing bf16 operand");
    int rc = gemm_tc_init();
    if (rc) return rc;
    if (N % 192 == 0) return launch<192>(A, W, M, N, K, passes, ep, s);
    return launch<128>(A, W, M, N, K, passes, ep, s);
}

int conv_tc(const GemmOperand& A, const GemmOperand& W, int B, int H, int Wd, int Cin, int N, int ks, int pad, int stride, int passes,
            float* out, int ldc, cudaStream_t s) {
    VETO_REQUIRE(passes == 1 || passes == 3, VETO_ERR_ARG, "conv_tc: passes must be 1 or 3");
    VETO_REQUIRE(Cin % BLOCK_K == 0 && N % 4 == 0 && ldc % 4 == 0 && ks >= 1 && pad >= 0 && pad < ks && stride >= 1 && stride <= 4,
                 VETO_ERR_UNSUPPORTED,
                 "conv_tc: Cin=%d must be a multiple of %d, N=%d and ldc=%d of 4", Cin, BLOCK_K, N, ldc);
    VETO_REQUIRE(A.hi && W.hi && (passes == 1 || (A.lo && W.lo)) && out, VETO_ERR_ARG, "conv_tc: missing operand");
    int rc = gemm_tc_init();
    if (rc) return rc;
    if (N <= 64) return launch_conv<64>(A, W, B, H, Wd, Cin, N, ks, pad, stride, passes, out, ldc, s);
    return launch_conv<128>(A, W, B, H, Wd, Cin, N, ks, pad, stride, passes, out, ldc, s);
}

}  // namespace veto
