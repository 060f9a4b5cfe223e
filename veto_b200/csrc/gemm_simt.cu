// fp32 SIMT GEMM:  C[M,N] = epilogue( A[M,K] @ W[N,K]^T ), every product an fp32 FMA.
//
// This is the reference-precision path (the reference runs DTYPE "float32", VETO_final.yaml:1) and the
// GEMM used for the small box-level projections and the predicate classifier in every precision mode,
// so that logits / argmax do not depend on tensor-core rounding.  128x128x16 tiles, 256 threads, 8x8
// register micro-tiles, register-prefetched double buffering.
#include "common.cuh"

namespace veto {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int PAD = 4;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int M, int N, int K,
                 const float* __restrict__ bias, const float* residual, int act, float* out_f32,
                 __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int ldc, int ldr, float* pre_f32, int res_mode,
                 DropSpec drop, int k_per, size_t split_stride) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K (tall-skinny weight gradients): slice blockIdx.z covers K columns [k_lo, k_hi), partial tile at z * split_stride
    const int k_lo = blockIdx.z * k_per;
    const int k_hi = (k_lo + k_per < K) ? k_lo + k_per : K;
    if (out_f32) out_f32 += (size_t)blockIdx.z * split_stride;

    // global -> register staging: each thread moves two float4 of A and two of W per K step
    const int lrow = tid >> 2;        // 0..63 (+64)
    const int lk = (tid & 3) * 4;     // 0,4,8,12
    const bool vec_ok = ((lda & 3) == 0) && ((K & 3) == 0) && ((((uintptr_t)A) & 15) == 0) && ((((uintptr_t)W) & 15) == 0);

    float4 ra[2], rb[2];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = m0 + lrow + h * 64;
            const int k = k0 + lk;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < M) {
                const float* p = A + (size_t)r * lda + k;
                if (vec_ok && k + 3 < k_hi) v = __ldg((const float4*)p);
                else {
                    if (k < k_hi) v.x = __ldg(p);
                    if (k + 1 < k_hi) v.y = __ldg(p + 1);
                    if (k + 2 < k_hi) v.z = __ldg(p + 2);
                    if (k + 3 < k_hi) v.w = __ldg(p + 3);
                }
            }
            ra[h] = v;
            const int c = n0 + lrow + h * 64;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < N) {
                const float* p = W + (size_t)c * K + k;
                if (vec_ok && k + 3 < k_hi) w = __ldg((const float4*)p);
                else {
                    if (k < k_hi) w.x = __ldg(p);
                    if (k + 1 < k_hi) w.y = __ldg(p + 1);
                    if (k + 2 < k_hi) w.z = __ldg(p + 2);
                    if (k + 3 < k_hi) w.w = __ldg(p + 3);
                }
            }
            rb[h] = w;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + h * 64;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
            As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
            Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (k_hi - k_lo + BK - 1) / BK;
    load_tile(k_lo);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile(k_lo + (kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *(const float4*)&As[buf][k][ty * 4];
            const float4 a1 = *(const float4*)&As[buf][k][64 + ty * 4];
            const float4 b0 = *(const float4*)&Bs[buf][k][tx * 4];
            const float4 b1 = *(const float4*)&Bs[buf][k][64 + tx * 4];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    const bool vec_out = ((N & 3) == 0) && ((ldc & 3) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = n0 + h * 64 + tx * 4;
            if (c >= N) continue;
            float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
            const size_t off = (size_t)r * ldc + c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c + j < N) {
                    if (bias) v[j] += __ldg(bias + c + j);
                    if (pre_f32) pre_f32[off + j] = v[j];
                    if (res_mode == RES_GELU_GRAD) {
                        v[j] *= gelu_grad(residual[(size_t)r * ldr + c + j]);
                    } else {
                        v[j] = apply_act(v[j], act);
                        if (drop.thr16) v[j] *= drop_scale1(drop, off + j);
                        if (residual) v[j] += residual[(size_t)r * ldr + c + j];
                    }
                }
            }
            if (vec_out && c + 3 < N) {
                if (out_f32) *(float4*)(out_f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
                if (out_hi) {
                    __nv_bfloat16 hh[4], ll[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_bf16(v[j], hh[j], ll[j]);
                    *(uint2*)(out_hi + off) = pack_bf16x4(hh[0], hh[1], hh[2], hh[3]);
                    if (out_lo) *(uint2*)(out_lo + off) = pack_bf16x4(ll[0], ll[1], ll[2], ll[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c + j < N) {
                        if (out_f32) out_f32[off + j] = v[j];
                        if (out_hi) {
                            __nv_bfloat16 hh, ll;
                            split_bf16(v[j], hh, ll);
                            out_hi[off + j] = hh;
                            if (out_lo) out_lo[off + j] = ll;
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- skinny outputs (the predicate classifier)
// C[M, N <= 128] = A[M,K] W[N,K]^T + bias: with N = 51 (VG) / 60 / 108 (MEET heads) the 128x128-tile kernel launches
// ceil(M/128) CTAs — 16 of 148 SMs for a 1994-pair inference chunk (124 us per chunk, 12.5 ms of the 435 ms inference
// step).  Here W^T lives in shared memory ([k][n], row stride = 1 mod 32: conflict-free both for the transposing fill
// and for the lane-per-column reads), a warp owns a row of A at a time (staged in shared memory, read back as
// broadcast float4), and lane j accumulates output columns j, j + 32, ...: every SM works, rows are independent and
// the k order is fixed, so results do not depend on how the rows are chunked.
constexpr int SK_THREADS = 256, SK_WARPS = SK_THREADS / 32, SK_ROWS = 2;
constexpr int SK_SMEM_FLOATS = 40960;  // 160 KB for the W^T tile

template <int NCH>
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int M, int N, int K, int KT,
                   const float* __restrict__ bias, float* __restrict__ out, int ldc) {
    extern __shared__ float sk_smem[];
    constexpr int NP = NCH * 32 + 1;
    float* sW = sk_smem;                         // [KT][NP]
    float* sX = sk_smem + (size_t)KT * NP;       // [SK_WARPS][SK_ROWS][KT]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* xs = sX + wid * SK_ROWS * KT;
    const int rows_per_pass = gridDim.x * SK_WARPS * SK_ROWS;
    const int passes = (M + rows_per_pass - 1) / rows_per_pass;
    for (int pass = 0; pass < passes; ++pass) {
        // a warp owns SK_ROWS consecutive rows at a time: every W^T value read from shared memory feeds SK_ROWS FMAs (the
        // one-row version was bound by its one shared-memory load per FMA)
        const int row0 = pass * rows_per_pass + (blockIdx.x * SK_WARPS + wid) * SK_ROWS;
        float acc[SK_ROWS][NCH];
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc[r][c] = 0.f;
        for (int k0 = 0; k0 < K; k0 += KT) {
            const int kt = (K - k0 < KT) ? K - k0 : KT;
            if (pass == 0 || K > KT) {           // the W^T tile stays resident when it holds all of K
                __syncthreads();
                for (int e = threadIdx.x; e < N * kt; e += SK_THREADS) {
                    const int n = e / kt, k = e - n * kt;  // consecutive threads walk k: coalesced in W, banks k + n
                    sW[k * NP + n] = __ldg(W + (size_t)n * K + k0 + k);
                }
                __syncthreads();
            }
            if (row0 < M) {
#pragma unroll
                for (int r = 0; r < SK_ROWS; ++r) {
                    const int row = min(row0 + r, M - 1);   // a row past the end repeats the last one; it is never stored
                    for (int k = lane; k < kt; k += 32) xs[r * KT + k] = __ldg(A + (size_t)row * lda + k0 + k);
                }
                __syncwarp();
                int k = 0;
                for (; k + 3 < kt; k += 4) {
                    float4 x[SK_ROWS];
#pragma unroll
                    for (int r = 0; r < SK_ROWS; ++r) x[r] = *(const float4*)(xs + r * KT + k);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const float* wp = sW + k * NP + lane + 32 * c;   // columns past N hold stale values: never stored
                        const float w0 = wp[0], w1 = wp[NP], w2 = wp[2 * NP], w3 = wp[3 * NP];
#pragma unroll
                        for (int r = 0; r < SK_ROWS; ++r) {
                            acc[r][c] = fmaf(x[r].x, w0, acc[r][c]);
                            acc[r][c] = fmaf(x[r].y, w1, acc[r][c]);
                            acc[r][c] = fmaf(x[r].z, w2, acc[r][c]);
                            acc[r][c] = fmaf(x[r].w, w3, acc[r][c]);
                        }
                    }
                }
                for (; k < kt; ++k) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const float wv = sW[k * NP + lane + 32 * c];
#pragma unroll
                        for (int r = 0; r < SK_ROWS; ++r) acc[r][c] = fmaf(xs[r * KT + k], wv, acc[r][c]);
                    }
                }
                __syncwarp();
            }
        }
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r) {
            const int row = row0 + r;
            if (row < M) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int n = lane + 32 * c;
                    if (n < N) out[(size_t)row * ldc + n] = acc[r][c] + (bias ? bias[n] : 0.f);
                }
            }
        }
    }
}

template <int NCH>
int launch_skinny(const float* A, int lda, const float* W, int M, int N, int K, const float* bias, float* out, int ldc,
                  cudaStream_t s) {
    constexpr int NP = NCH * 32 + 1;
    int KT = (SK_SMEM_FLOATS / NP) & ~3;
    if (KT > K) KT = (K + 3) & ~3;
    const int smem = (KT * NP + SK_WARPS * SK_ROWS * KT) * (int)sizeof(float);
    static DeviceOnce attr_set;
    if (attr_set.pending()) {
        VETO_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set.done();
    }
    int grid = (M + SK_WARPS * SK_ROWS - 1) / (SK_WARPS * SK_ROWS);
    if (grid > num_sms()) grid = num_sms();
    gemm_skinny_kernel<NCH><<<grid, SK_THREADS, smem, s>>>(A, lda, W, M, N, K, KT, bias, out, ldc);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace

int gemm_simt_slices(int K, int split_k) {
    if (split_k <= 1) return 1;
    const int k_per = ((K + split_k - 1) / split_k + BK - 1) / BK * BK;
    return (K + k_per - 1) / k_per;
}

int gemm_simt(const float* A, int lda, const float* W, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s) {
    if (M <= 0 || N <= 0) return VETO_OK;
    VETO_REQUIRE(K > 0 && A && W, VETO_ERR_ARG, "gemm_simt: bad operands");
    // skinny outputs with a plain (bias-only) epilogue: the predicate classifier
    if (N <= 128 && K >= 64 && M >= 256 && ep.split_k <= 1 && !ep.residual && ep.act == ACT_NONE && !ep.pre_f32 && !ep.drop.thr16 &&
        ep.out.f32 && !ep.out.hi && (lda % 4) == 0 && (((uintptr_t)A) & 15) == 0) {
        switch ((N + 31) / 32) {
            case 1: return launch_skinny<1>(A, lda, W, M, N, K, ep.bias, ep.out.f32, ep.ldc, s);
            case 2: return launch_skinny<2>(A, lda, W, M, N, K, ep.bias, ep.out.f32, ep.ldc, s);
            case 3: return launch_skinny<3>(A, lda, W, M, N, K, ep.bias, ep.out.f32, ep.ldc, s);
            default: return launch_skinny<4>(A, lda, W, M, N, K, ep.bias, ep.out.f32, ep.ldc, s);
        }
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    VETO_REQUIRE(grid.y <= 65535, VETO_ERR_UNSUPPORTED, "gemm_simt: M=%d too large for one launch", M);
    int slices = ep.split_k > 1 ? ep.split_k : 1;
    int k_per = K;
    if (slices > 1) {
        VETO_REQUIRE(!ep.bias && !ep.residual && ep.act == ACT_NONE && !ep.pre_f32 && !ep.drop.thr16 && ep.out.f32 && !ep.out.hi,
                     VETO_ERR_ARG, "gemm_simt: split-K writes plain fp32 partial products only");
        k_per = ((K + slices - 1) / slices + BK - 1) / BK * BK;  // whole K tiles per slice (keeps float4 alignment)
        slices = (K + k_per - 1) / k_per;
        VETO_REQUIRE(slices == gemm_simt_slices(K, ep.split_k), VETO_ERR_ARG, "gemm_simt: slice count mismatch");
        grid.z = slices;
    }
    VETO_REQUIRE(ep.res_mode == RES_ADD || ep.residual, VETO_ERR_ARG, "gemm_simt: RES_GELU_GRAD needs the pre-activation");
    gemm_simt_kernel<<<grid, 256, 0, s>>>(A, lda, W, M, N, K, ep.bias, ep.residual, ep.act, ep.out.f32, ep.out.hi,
                                          ep.out.lo, ep.ldc, ep.ldr ? ep.ldr : ep.ldc, ep.pre_f32, ep.res_mode, ep.drop, k_per,
                                          ep.split_stride);
    VETO_LAUNCH_CHECK();
    return VETO_OK;
}

}  // namespace veto
