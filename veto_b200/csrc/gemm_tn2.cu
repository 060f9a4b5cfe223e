// Weight-gradient GEMM of the training step on the tensor cores, with BOTH operands read in place:
//
//   G[Nw, Kw] = dY[rows, Nw]^T @ X[rows, Kw]          (dW = dY^T X of a Linear y = x W^T; reduction over the token rows)
//
// dY and X are the row-major bf16 hi/lo activation arrays the backward pass already holds, so the reduction dimension
// (rows) is the SLOW dimension of both: for tcgen05 these are "MN-major" operands (instruction-descriptor bits 15 / 16),
// staged by TMA as [64 rows] x [64 contiguous elements = 128 B] boxes with the 128-byte swizzle — the canonical layout
// Swizzle<3,4,3> o ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) elements of cute's make_umma_desc<Major::MN>: a K row is 128 B,
// eight K rows form a 1024-byte swizzle atom (SBO), the next 64 MN elements live LBO bytes further (the next TMA box).
// No transposed copies of the activations are ever written (the first version of the backward spent 10 ms of a 51 ms
// step on them, profiles/r1_train_launches_v1.txt).
//
// Everything else follows gemm_tc2.cu: a cluster of two CTAs owns an output tile (tcgen05 cta_group::2, M = 256,
// K = 16), warp-specialised TMA / MMA / 12 epilogue warps, TMEM double buffering, one pipeline stage serving the three
// products of the bf16x3 split, and split-K over the rows (tiny output, huge reduction): partial tiles are written as
// plain fp32 and summed in a fixed order by splitk_reduce (train.cu).  Rows beyond the end of the arrays are
// zero-filled by TMA, so the row count needs no padding.
//
// Convolution form (gemm_tn2_conv, the depth backbone's weight gradients): dY and X are NHWC activations and the
// reduction runs over pixels.  A K block is a 4 x 16 pixel patch; the dY box is the patch itself, the X box of an output
// column block (tap, 64 channels) the same patch shifted by the tap — 4-D TMA boxes whose shared-memory image equals the
// 64-row 2-D box, with out-of-bounds zero fill standing in for the padding (and for patches hanging over the image
// edge).  Only the producer differs; no im2col matrix exists.
//
// Tile width: the first version used 256 x 128 tiles and ran at 62-68 % tensor-active (profiles/r1_train_ncu_full.txt):
// per pipeline stage each SM writes 48 KB (TMA) and reads 72 KB (3 products x (A 16 KB + its half of B 8 KB)) for 816
// tensor cycles = 147 B/clk, above the 128 B/clk shared-memory port.  A 256-wide tile halves the A traffic per MMA
// cycle (160 KB per 1632 cycles = 98 B/clk), but the 576-wide outputs are 2.25 such tiles.  So the column range of an
// output is cut into 256-wide tiles plus, for a remainder of at most 128 columns, one 128-wide tile (576 = 256 + 256 +
// 64 of 128; 1152 = 4 x 256 + 128): the tile width is a per-tile runtime value (instruction descriptor, number of B
// boxes, bytes expected by the stage barrier, epilogue chunk count), the stage layout is that of the wide tile.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#define VETO_TC_KERNEL "gemm_tn2"
#include "tcgen05.cuh"

namespace veto {
namespace {
using namespace tc;

constexpr int BLOCK_M = 128;      // output rows (Nw) per CTA, 256 per pair
constexpr int WIDE_N = 256;       // output columns (Kw) per wide pair tile; each CTA stages half of them
constexpr int NARROW_N = 128;     // ... per narrow tile (the remainder of a column range)
constexpr int BLOCK_K = 64;       // reduction rows per stage
constexpr int UMMA_K = 16;
constexpr int MN_CHUNK = 64;      // contiguous elements per TMA box row (128 B)
constexpr int NUM_EPI_WARPS = 12;
constexpr int NUM_THREADS = (4 + NUM_EPI_WARPS) * 32;
constexpr int EPI_COLS = 16;
constexpr int EPI_STAGE_BYTES = 32 * EPI_COLS * 4;
constexpr int CHUNK_BYTES = BLOCK_K * MN_CHUNK * 2;        // 8 KB: one TMA box
constexpr int BYTES_A = (BLOCK_M / MN_CHUNK) * CHUNK_BYTES;  // 16 KB
constexpr int BYTES_B = (WIDE_N / 2 / MN_CHUNK) * CHUNK_BYTES;  // 16 KB reserved (a narrow tile fills the first 8 KB)
constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_STAGE_BYTES;
constexpr int MAX_STAGES = 8;
constexpr int PIPE_BYTES = 192 * 1024;                      // 3 stages of 64 KB (3-pass) or 6 stages of 32 KB
constexpr int TMEM_COLS = 512;                              // two accumulators of up to 256 columns
constexpr int SMEM_BYTES = PIPE_BYTES + EPI_BYTES + 1024 + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// arrive on the barrier at the same offset in CTA `rank` of the cluster
// TMA load whose completion bytes go to the LEADER CTA's barrier (peer bit of the shared::cluster address cleared)
// arrive (once all MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair


// MN-major SWIZZLE_128B shared-memory descriptor: leading byte offset = distance between 64-element MN chunks,
// stride byte offset = distance between groups of eight K rows
// fp32 accumulate, bf16 A and B, both MN-major (bits 15, 16)

struct TnParams {
    float* out;              // [ksplit][Nw, ldc] fp32 partial products
    int ldc;
    int ksplit, kb_per;
    long long split_stride;
    uint32_t lbo, sbo, kadv;  // descriptor geometry (constants in production; parameters for the bring-up test)
    // convolution form: K block kb = pixel patch (b, th, tw) of CONV_PH x CONV_PW pixels; X column = (kh*ks + kw)*Cin + c
    int conv, Cin, ks, pad, n_th, n_tw, stride;
};
constexpr int CONV_PH = 4, CONV_PW = 16;  // CONV_PH * CONV_PW == BLOCK_K

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tn2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                int Nw, int Kw, int rows, int passes, TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared array itself: a round trip through uintptr_t makes every
    // access through the result a GENERIC load / store (LD.E / ST.E instead of LDS / STS — the epilogue staging paid for it)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_epi = smem + PIPE_BYTES;
    uint64_t* bars = (uint64_t*)(smem_epi + EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const bool split = passes == 3;
    const int stage_bytes = split ? 2 * (BYTES_A + BYTES_B) : (BYTES_A + BYTES_B);
    const int num_stages = PIPE_BYTES / stage_bytes;  // 3 or 6
    const int num_m = (Nw + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    // column tiles: n_wide tiles of 256, then at most one narrow tile of 128 for a remainder of <= 128 columns
    const int rem = Kw % WIDE_N;
    const int n_wide = Kw / WIDE_N + (rem > NARROW_N ? 1 : 0);
    const int num_n = n_wide + ((rem > 0 && rem <= NARROW_N) ? 1 : 0);
    const int mn_tiles = num_m * num_n;
    const int num_tiles = mn_tiles * p.ksplit;
    const int num_kb = (rows + BLOCK_K - 1) / BLOCK_K;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    auto tile_mn = [&](int tile, int& tm, int& tn) {
        const int t2 = tile % mn_tiles;
        tm = t2 / num_n;
        tn = t2 - tm * num_n;
        return tile / mn_tiles;
    };
    auto tile_width = [&](int tn) { return tn < n_wide ? WIDE_N : NARROW_N; };   // first column of tile tn = tn * WIDE_N
    auto kb_range = [&](int ks, int& kb0, int& kb1) {
        kb0 = ks * p.kb_per;
        kb1 = kb0 + p.kb_per < num_kb ? kb0 + p.kb_per : num_kb;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_b_hi);
        if (split) {
            tma_prefetch_desc(&tm_a_lo);
            tma_prefetch_desc(&tm_b_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
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

    // stage layout: [A_hi][A_lo][B_hi][B_lo] (3-pass) or [A][B] (1-pass); A = two 8 KB boxes, B = two (wide) or one
    auto stage_ptr = [&](int s) { return smem + s * stage_bytes; };
    const int off_a_lo = BYTES_A;
    const int off_b_hi = split ? 2 * BYTES_A : BYTES_A;
    const int off_b_lo = off_b_hi + BYTES_B;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int tm, tn, kb0, kb1;
                kb_range(tile_mn(tile, tm, tn), kb0, kb1);
                const int width = tile_width(tn);
                const int b_boxes = width / 2 / MN_CHUNK;              // 64-column boxes of X this CTA stages: 2 or 1
                const int m0 = tm * (2 * BLOCK_M) + rank * BLOCK_M;   // first dY column (output row) of this CTA
                const int n0 = tn * WIDE_N + rank * (width / 2);       // first X column (output column) of this CTA
                const int tx_bytes = 2 * (split ? 2 : 1) * (BYTES_A + b_boxes * CHUNK_BYTES);   // both CTAs of the pair
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int r0 = kb * BLOCK_K;
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                    uint8_t* sp = stage_ptr(stage);
                    if (p.conv) {
                        const int w0 = (kb % p.n_tw) * CONV_PW, h0 = ((kb / p.n_tw) % p.n_th) * CONV_PH, b = kb / (p.n_tw * p.n_th);
                        for (int j = 0; j < BLOCK_M / MN_CHUNK; ++j) {
                            tma_load_4d_pair(sp + j * CHUNK_BYTES, &tm_a_hi, &full_bar[stage], m0 + j * MN_CHUNK, w0, h0, b);
                            if (split)
                                tma_load_4d_pair(sp + off_a_lo + j * CHUNK_BYTES, &tm_a_lo, &full_bar[stage], m0 + j * MN_CHUNK, w0, h0, b);
                        }
                        for (int j = 0; j < b_boxes; ++j) {
                            const int col = n0 + j * MN_CHUNK;
                            const int tap = col / p.Cin;
                            // columns beyond the last tap: a box outside the channel range (zero fill)
                            const int c0 = tap < p.ks * p.ks ? col - tap * p.Cin : p.Cin;
                            const int kh = tap / p.ks, kw = tap - kh * p.ks;
                            const int xw = w0 * p.stride + kw - p.pad, xh = h0 * p.stride + kh - p.pad;
                            tma_load_4d_pair(sp + off_b_hi + j * CHUNK_BYTES, &tm_b_hi, &full_bar[stage], c0, xw, xh, b);
                            if (split) tma_load_4d_pair(sp + off_b_lo + j * CHUNK_BYTES, &tm_b_lo, &full_bar[stage], c0, xw, xh, b);
                        }
                        if (++stage == num_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < BLOCK_M / MN_CHUNK; ++j) {
                        tma_load_2d_pair(sp + j * CHUNK_BYTES, &tm_a_hi, &full_bar[stage], m0 + j * MN_CHUNK, r0);
                        if (split) tma_load_2d_pair(sp + off_a_lo + j * CHUNK_BYTES, &tm_a_lo, &full_bar[stage], m0 + j * MN_CHUNK, r0);
                    }
                    for (int j = 0; j < b_boxes; ++j) {
                        tma_load_2d_pair(sp + off_b_hi + j * CHUNK_BYTES, &tm_b_hi, &full_bar[stage], n0 + j * MN_CHUNK, r0);
                        if (split) tma_load_2d_pair(sp + off_b_lo + j * CHUNK_BYTES, &tm_b_lo, &full_bar[stage], n0 + j * MN_CHUNK, r0);
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
            constexpr uint32_t idesc_wide = make_idesc_mn(2 * BLOCK_M, WIDE_N), idesc_narrow = make_idesc_mn(2 * BLOCK_M, NARROW_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * WIDE_N;
                int tm, tn, kb0, kb1;
                kb_range(tile_mn(tile, tm, tn), kb0, kb1);
                const uint32_t idesc = tile_width(tn) == WIDE_N ? idesc_wide : idesc_narrow;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t sp = smem_u32(stage_ptr(stage));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint32_t ko = k * p.kadv;
                        const uint64_t a_hi = make_desc_mn(sp + ko, p.lbo, p.sbo);
                        const uint64_t b_hi = make_desc_mn(sp + off_b_hi + ko, p.lbo, p.sbo);
                        umma2_bf16(tmem_d, a_hi, b_hi, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        if (split) {
                            umma2_bf16(tmem_d, make_desc_mn(sp + off_a_lo + ko, p.lbo, p.sbo), b_hi, idesc, 1u);
                            umma2_bf16(tmem_d, a_hi, make_desc_mn(sp + off_b_lo + ko, p.lbo, p.sbo), idesc, 1u);
                        }
                    }
                    umma2_commit_both(&empty_bar[stage]);
                    if (++stage == num_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma2_commit_both(&tmem_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 output rows): plain fp32 partial tiles =====================
        const int q = warp & 3;
        const int third = (warp - 4) >> 2;
        constexpr int kStride = NUM_EPI_WARPS / 4;
        float4* stage4 = reinterpret_cast<float4*>(smem_epi + (warp - 4) * EPI_STAGE_BYTES);
        const int rsub = lane >> 2, cg = lane & 3;
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            int tm, tn;
            const int ks = tile_mn(tile, tm, tn);
            const int m0 = tm * (2 * BLOCK_M) + rank * BLOCK_M + q * 32;
            const int n0 = tn * WIDE_N;
            const int kChunks = tile_width(tn) / EPI_COLS;
            float* outp = p.out + (size_t)ks * (size_t)p.split_stride;
            mbar_wait(&tmem_full[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * WIDE_N;
#pragma unroll 1
            for (int c = third; c < kChunks; c += kStride) {
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
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int lr = rr * 8 + rsub;
                    const float4 v = stage4[lr * 4 + (cg ^ ((lr >> 1) & 3))];
                    const int row = m0 + lr;
                    if (row < Nw && col < Kw) *(float4*)(outp + (size_t)row * p.ldc + col) = v;
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
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
    bool operator==(const MapKey& o) const { return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        return std::hash<const void*>()(k.p) ^ (k.rows * 0x9E3779B97F4A7C15ull) ^ (k.cols << 20) ^ (k.ld << 7);
    }
};
// per host thread: no lock on the launch path (descriptors are pure functions of their key)
thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// [rows, cols] row-major bf16 with row stride ld: boxes of 64 rows x 64 contiguous elements, 128-byte swizzle
int get_map(const __nv_bfloat16* p, uint64_t rows, uint64_t cols, uint64_t ld, CUtensorMap* out) {
    MapKey key{p, rows, cols, ld};
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
        *out = it->second;
        return VETO_OK;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {MN_CHUNK, BLOCK_K};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)p, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%llu,%llu] ld %llu at %p (gemm_tn2)", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols, (unsigned long long)ld, (const void*)p);
        return VETO_ERR_CUDA;
    }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return VETO_OK;
}

int init_tn() {
    if (!g_inited.pending()) return VETO_OK;
    std::lock_guard<std::mutex> lk(g_mu);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VETO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VETO_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VETO_ERR_CUDA,
                 "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
    VETO_CUDA(cudaFuncSetAttribute(gemm_tn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_inited.done();
    return VETO_OK;
}

}  // namespace

bool gemm_tn2_supported(int Nw, int Kw, int ld_y, int ld_x) {
    return Nw % 8 == 0 && Kw % 8 == 0 && ld_y % 8 == 0 && ld_x % 8 == 0;
}

// output tiles of one K slice: 256-row tiles x (256-wide column tiles + at most one 128-wide tile for the remainder)
int gemm_tn2_mn_tiles(int Nw, int Kw) {
    const int rem = Kw % WIDE_N;
    const int n_wide = Kw / WIDE_N + (rem > NARROW_N ? 1 : 0);
    return ((Nw + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (n_wide + ((rem > 0 && rem <= NARROW_N) ? 1 : 0));
}

// fraction of the machine a launch with `slices` K slices keeps busy: tiles go round-robin to the CTA pairs, a wide
// tile costs two narrow ones
double gemm_tn2_efficiency(int Nw, int Kw, int slices, int pairs) {
    const int rem = Kw % WIDE_N;
    const int n_wide = Kw / WIDE_N + (rem > NARROW_N ? 1 : 0);
    const int num_n = n_wide + ((rem > 0 && rem <= NARROW_N) ? 1 : 0);
    const int mn = gemm_tn2_mn_tiles(Nw, Kw);
    std::vector<double> load((size_t)pairs, 0.0);
    double total = 0.0;
    for (int t = 0; t < mn * slices; ++t) {
        const double c = ((t % mn) % num_n) < n_wide ? 2.0 : 1.2;
        load[(size_t)(t % pairs)] += c;
        total += c;
    }
    double mx = 0.0;
    for (double l : load) mx = l > mx ? l : mx;
    return mx > 0.0 ? total / (mx * pairs) : 0.0;
}

int gemm_tn2_slices(int rows, int split_k) { return gemm_tc2_slices((rows + BLOCK_K - 1) / BLOCK_K * BLOCK_K, split_k); }

int gemm_tn2(const GemmOperand& dY, const GemmOperand& X, int Nw, int Kw, int rows, int passes, float* out, int ldc, int split_k,
             size_t split_stride, cudaStream_t s, const uint32_t* geometry) {
    if (Nw <= 0 || Kw <= 0 || rows <= 0) return VETO_OK;
    VETO_REQUIRE(passes == 1 || passes == 3, VETO_ERR_ARG, "gemm_tn2: passes must be 1 or 3");
    VETO_REQUIRE(dY.hi && X.hi && (passes == 1 || (dY.lo && X.lo)), VETO_ERR_ARG, "gemm_tn2: missing bf16 operand");
    const int ldy = dY.ld ? dY.ld : Nw, ldx = X.ld ? X.ld : Kw;
    VETO_REQUIRE(gemm_tn2_supported(Nw, Kw, ldy, ldx) && ldc % 4 == 0 && Kw % 4 == 0, VETO_ERR_UNSUPPORTED,
                 "gemm_tn2: Nw=%d Kw=%d ld %d/%d ldc %d must be multiples of 8 (ldc of 4)", Nw, Kw, ldy, ldx, ldc);
    int rc = init_tn();
    if (rc) return rc;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if ((rc = get_map(dY.hi, rows, Nw, ldy, &ta_hi))) return rc;
    if ((rc = get_map(X.hi, rows, Kw, ldx, &tb_hi))) return rc;
    ta_lo = ta_hi;
    tb_lo = tb_hi;
    if (passes == 3) {
        if ((rc = get_map(dY.lo, rows, Nw, ldy, &ta_lo))) return rc;
        if ((rc = get_map(X.lo, rows, Kw, ldx, &tb_lo))) return rc;
    }
    const int num_kb = (rows + BLOCK_K - 1) / BLOCK_K;
    const int ksplit = gemm_tn2_slices(rows, split_k);
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    const int tiles = gemm_tn2_mn_tiles(Nw, Kw) * ksplit;
    const int pairs_avail = num_sms() / 2;
    const int grid = 2 * (tiles < pairs_avail ? tiles : pairs_avail);
    TnParams p{out, ldc, ksplit, kb_per, (long long)split_stride,
               geometry ? geometry[0] : (uint32_t)CHUNK_BYTES, geometry ? geometry[1] : 1024u, geometry ? geometry[2] : 2048u,
               0, 0, 0, 0, 0, 0, 1};
    gemm_tn2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, Nw, Kw, rows, passes, p);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

namespace {
// NHWC bf16 [B,H,W,C]: boxes of a CONV_PH x CONV_PW pixel patch x 64 channels, 128-byte swizzle, zero fill
int get_map_nhwc(const __nv_bfloat16* p, int B, int H, int W, int C, int stride, CUtensorMap* out) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {MN_CHUNK, (cuuint32_t)(CONV_PW * stride), (cuuint32_t)(CONV_PH * stride), 1};  // every stride-th pixel
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for NHWC [%d,%d,%d,%d] at %p (gemm_tn2_conv)", (int)r, B, H, W, C, (const void*)p);
        return VETO_ERR_CUDA;
    }
    return VETO_OK;
}
}  // namespace

int gemm_tn2_conv_rows(int B, int H, int W) { return B * ((H + CONV_PH - 1) / CONV_PH) * ((W + CONV_PW - 1) / CONV_PW) * BLOCK_K; }

int gemm_tn2_conv(const GemmOperand& dY, const GemmOperand& X, int B, int H, int W, int Hin, int Win, int Cout, int Cin, int ks, int pad,
                  int stride, int passes, float* out, int ldc, int split_k, size_t split_stride, cudaStream_t s) {
    VETO_REQUIRE(passes == 1 || passes == 3, VETO_ERR_ARG, "gemm_tn2_conv: passes must be 1 or 3");
    VETO_REQUIRE(dY.hi && X.hi && (passes == 1 || (dY.lo && X.lo)) && out, VETO_ERR_ARG, "gemm_tn2_conv: missing operand");
    VETO_REQUIRE(Cout % 8 == 0 && Cin % MN_CHUNK == 0 && ldc % 4 == 0, VETO_ERR_UNSUPPORTED,
                 "gemm_tn2_conv: Cout=%d must be a multiple of 8, Cin=%d of %d, ldc=%d of 4", Cout, Cin, MN_CHUNK, ldc);
    int rc = init_tn();
    if (rc) return rc;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if ((rc = get_map_nhwc(dY.hi, B, H, W, Cout, 1, &ta_hi))) return rc;
    if ((rc = get_map_nhwc(X.hi, B, Hin, Win, Cin, stride, &tb_hi))) return rc;
    ta_lo = ta_hi;
    tb_lo = tb_hi;
    if (passes == 3) {
        if ((rc = get_map_nhwc(dY.lo, B, H, W, Cout, 1, &ta_lo))) return rc;
        if ((rc = get_map_nhwc(X.lo, B, Hin, Win, Cin, stride, &tb_lo))) return rc;
    }
    const int Kw = ks * ks * Cin;
    const int n_th = (H + CONV_PH - 1) / CONV_PH, n_tw = (W + CONV_PW - 1) / CONV_PW;
    const int num_kb = B * n_th * n_tw;
    const int rows = num_kb * BLOCK_K;
    const int ksplit = gemm_tn2_slices(rows, split_k);
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    const int tiles = gemm_tn2_mn_tiles(Cout, Kw) * ksplit;
    const int pairs_avail = num_sms() / 2;
    const int grid = 2 * (tiles < pairs_avail ? tiles : pairs_avail);
    TnParams p{out, ldc, ksplit, kb_per, (long long)split_stride, (uint32_t)CHUNK_BYTES, 1024u, 2048u, 1, Cin, ks, pad, n_th, n_tw, stride};
    gemm_tn2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, Cout, Kw, rows, passes, p);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
