// K2 -- all-timestep input projections  C[M,N] = A[M,K] * B[K,N] (+ bias[N]).
//
// Replaces the `x . kernel + bias` half of every Keras LSTMCell step of encoder
// layers > 0 (hoisted over all timesteps; basecaller.py:19-32 + SURVEY A.1) and the
// Luong memory layer `keys = values . memory_layer` (basecaller.py:131-134, A.3).
//
// Two implementations live here:
//   * gemm_tc   : tcgen05 / TMEM / TMA kernel (see proj_gemm_tc.cuh) -- the product path;
//   * gemm_simt : register-tiled FFMA kernel.  It exists to validate gemm_tc on the
//     device (tests call both through rvb_project) and as the exact-fp32 reference
//     when debugging; the model path never selects it implicitly.
#include "kernels.cuh"
#define RVB_HAVE_TC 1
#include <cuda_fp16.h>
#include "proj_gemm_tc.cuh"

namespace rvb {
namespace gemm {

constexpr int TM = 128, TN = 128, TK = 16, THREADS = 256;

// 128x128x16 tiles, 8x8 outputs per thread (two 4-wide strips per dimension).
__global__ void __launch_bounds__(THREADS) gemm_simt_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                            const float *__restrict__ bias, float *__restrict__ C,
                                                            long long M, int N, int K) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int a_row = tid >> 2, a_k = (tid & 3) * 4;       // + 64 rows for the second load
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;        // + 8 k for the second load
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            long long row = m0 + a_row + 64 * h;
            float4 v = make_float4(0, 0, 0, 0);
            if (row < M) v = __ldg(reinterpret_cast<const float4 *>(A + row * K + k0 + a_k));
            As[a_k + 0][a_row + 64 * h] = v.x; As[a_k + 1][a_row + 64 * h] = v.y;
            As[a_k + 2][a_row + 64 * h] = v.z; As[a_k + 3][a_row + 64 * h] = v.w;
            float4 w = __ldg(reinterpret_cast<const float4 *>(Bm + (size_t)(k0 + b_k + 8 * h) * N + n0 + b_n));
            *reinterpret_cast<float4 *>(&Bs[b_k + 8 * h][b_n]) = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4 *>(&Bs[k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int col = n0 + h * 64 + tx * 4;
            float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            if (bias != nullptr) {
                float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + col));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            *reinterpret_cast<float4 *>(C + row * N + col) = v;
        }
    }
}

int run_simt(const float *A, const float *Bm, const float *bias, float *C, long long M, int N, int K,
             cudaStream_t stream) {
    if (M <= 0) return RVB_OK;
    if (N % TN != 0 || K % TK != 0) return fail(RVB_ERR_ARG, "gemm_simt: N %% 128 and K %% 16 must be 0 (N=%d K=%d)", N, K);
    dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)(N / TN));
    { ProfScope ps(KK_GEMM, stream);
      gemm_simt_kernel<<<grid, THREADS, 0, stream>>>(A, Bm, bias, C, M, N, K); }
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

bool tc_available() { return true; }

// W[K,N] row-major -> tf32-exact hi part and remainder, both transposed to [N,K] (K-major B operand).
__global__ void split_transpose_kernel(const float *__restrict__ W, float *__restrict__ hiT, float *__restrict__ loT, int K, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const int n = i / K, k = i % K;
    const float v = W[(size_t)k * N + n];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hiT[i] = h;
    loT[i] = v - h;
}

int prepare_weights(const float *W, float *hiT, float *loT, int K, int N, cudaStream_t stream) {
    const int n = K * N;
    split_transpose_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, hiT, loT, K, N);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

int run_tc(const float *A, const float *WhiT, const float *WloT, const float *bias, float *C, long long M, int N, int K,
           int precision, int *abort_flag, cudaStream_t stream, long long lda, const CellEpilogue *cell) {
    if (M <= 0) return RVB_OK;
    if (N % 128 != 0 || K % tc::BK != 0) return fail(RVB_ERR_ARG, "gemm_tc: N %% 128 and K %% 32 must be 0 (N=%d K=%d)", N, K);
    const bool three = (precision == RVB_PREC_FP32);
    if (cell != nullptr) {
        if (N != GATES || !three) return fail(RVB_ERR_ARG, "gemm_tc: the fused cell epilogue needs N = 512 and the fp32-parity mode");
        return tc::launch_persistent<3, false>(A, nullptr, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda, false, cell);
    }
    static const bool legacy = getenv("RVB_GEMM_NONPERSISTENT") != nullptr;      // A/B switch for profiling
    if (N % 256 == 0 && !legacy) return three ? tc::launch_persistent<3, false>(A, nullptr, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda)
                                              : tc::launch_persistent<1, false>(A, nullptr, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda);
    if (N % 256 == 0) return three ? tc::launch<256, 3>(A, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda)
                                   : tc::launch<256, 1>(A, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda);
    return three ? tc::launch<128, 3>(A, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda)
                 : tc::launch<128, 1>(A, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda);
}

// ---- fp16-plane path: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) ------------------------------------
__global__ void split_f16_kernel(const float *__restrict__ X, __half *__restrict__ hi, __half *__restrict__ lo, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = X[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
}
// W[K,N] fp32 row-major -> fp16 hi / lo planes transposed to [N,K]
__global__ void split_transpose_f16_kernel(const float *__restrict__ W, __half *__restrict__ hiT, __half *__restrict__ loT, int K, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const int n = i / K, k = i % K;
    const float v = W[(size_t)k * N + n];
    const __half h = __float2half_rn(v);
    hiT[i] = h;
    loT[i] = __float2half_rn(v - __half2float(h));
}
int split_planes_f16(const float *X, void *hi, void *lo, long long n, cudaStream_t stream) {
    if (n <= 0) return RVB_OK;
    split_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(X, reinterpret_cast<__half *>(hi), reinterpret_cast<__half *>(lo), n);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}
int prepare_weights_f16(const float *W, void *hiT, void *loT, int K, int N, cudaStream_t stream) {
    const int n = K * N;
    split_transpose_f16_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, reinterpret_cast<__half *>(hiT), reinterpret_cast<__half *>(loT), K, N);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}
// blocked_out: C is written as [M / 128][N / 4][128 rows][4] (M % 128 == 0) with plain coalesced stores -- the layout
// the recurrent kernel reads lane-contiguously -- instead of row-major through the staged TMA store.
int run_tc_f16(const void *Ahi, const void *Alo, const void *WhiT, const void *WloT, const float *bias, float *C, long long M,
               int N, int K, int precision, int *abort_flag, cudaStream_t stream, bool blocked_out, const CellEpilogue *cell,
               long long lda, int n_out, bool blocked_half) {
    if (M <= 0) return RVB_OK;
    if (blocked_half && (!blocked_out || precision == RVB_PREC_FP32)) return fail(RVB_ERR_ARG, "gemm_tc_f16: fp16 output only for the blocked layout in reduced-precision mode");
    if (N % 256 != 0 || K % 64 != 0) return fail(RVB_ERR_ARG, "gemm_tc_f16: N %% 256 and K %% 64 must be 0 (N=%d K=%d)", N, K);
    if (blocked_out && M % 128 != 0) return fail(RVB_ERR_ARG, "gemm_tc_f16: blocked output needs M %% 128 == 0 (M=%lld)", M);
    if (cell != nullptr && (N != GATES || precision != RVB_PREC_FP32)) return fail(RVB_ERR_ARG, "gemm_tc_f16: the fused cell epilogue needs N = 512 and the fp32-parity mode");
    if (precision == RVB_PREC_FP32) return tc::launch_persistent<3, true>(Ahi, Alo, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda, blocked_out, cell, n_out);
    return tc::launch_persistent<1, true>(Ahi, Alo, WhiT, WloT, bias, C, M, N, K, abort_flag, stream, lda, blocked_out, nullptr, n_out, blocked_half);
}

}  // namespace gemm
}  // namespace rvb
