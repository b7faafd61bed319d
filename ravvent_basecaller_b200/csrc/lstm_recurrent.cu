// K3 -- persistent recurrent LSTM (one direction of one BiLSTM layer), fp32.
//
// Replaces the Keras RNN(LSTMCell) while-loop the reference builds in
// Encoder.__init__/call (basecaller.py:19-32, 48-59; LSTMCell math: SURVEY A.1).
//
// Work unit: (tile of 64 snippets, direction) = one 2-CTA cluster.  CTA `rank`
// owns hidden units [64*rank, 64*rank+64): its 256 gate columns of the recurrent
// kernel U stay resident in shared memory for all T steps (128 KB fp32); the
// full h vector of the tile is double-buffered in both CTAs' shared memory and
// each CTA pushes its half of the new h to the peer over DSMEM, followed by one
// cluster barrier per step.  Gate nonlinearities and the cell update are fused
// after the register-tiled FMA loop; c never leaves registers.
//   layer 0 : x_t (F = 1 raw / 5 event features) and the bias enter as extra
//             K rows (F weights rows + a row of ones), so no GEMM is needed;
//   layer>0 : the accumulators start from the pre-projected gates
//             G[b,t,dir,:] = y_{l-1}[b,t,:] W + b written by K2.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace rvb {
namespace rec {

constexpr int BT = 64;          // snippets per cluster
constexpr int HU = 64;          // hidden units per CTA
constexpr int THREADS = 256;


__device__ __forceinline__ float fsig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 2.0f * fsig(2.0f * x) - 1.0f; }

template <int F, bool PRE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) lstm_rec_kernel(Params p) {
    constexpr int KX = PRE ? UNITS : UNITS + F + 1;
    extern __shared__ __align__(16) float smem[];
    float *Us = smem;                    // [KX][4][64]
    float *hs = smem + KX * 256;         // [2][KX][64]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x >> 1;
    const int dir = cid & 1;
    const int b0 = (cid >> 1) * BT;
    const int tid = threadIdx.x;
    const int ug = tid & 15, rg = tid >> 4;
    const int T = p.T, B = p.B;

    {   // resident weights
        const float4 *src = reinterpret_cast<const float4 *>(p.wpack + (size_t)(dir * 2 + rank) * KX * 256);
        float4 *dst = reinterpret_cast<float4 *>(Us);
        for (int i = tid; i < KX * 64; i += THREADS) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < BT * UNITS; i += THREADS) {          // h0 -> buffer 0
        int row = i >> 7, k = i & 127, b = b0 + row;
        float v = 0.0f;
        if (p.state_in != nullptr && b < B) v = p.state_in[(((size_t)b * 2 + dir) * 2 + 0) * UNITS + k];
        hs[k * BT + row] = v;
    }
    if (!PRE) {
        const int t0 = dir ? T - 1 : 0;
        for (int i = tid; i < F * BT; i += THREADS) {
            int f = i / BT, row = i % BT, b = b0 + row;
            hs[(UNITS + f) * BT + row] = (b < B) ? p.x[((size_t)b * T + t0) * F + f] : 0.0f;
        }
        for (int i = tid; i < BT; i += THREADS) { hs[(UNITS + F) * BT + i] = 1.0f; hs[KX * BT + (UNITS + F) * BT + i] = 1.0f; }
    }
    float c[4][4], h[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int b = b0 + 4 * rg + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c[i][j] = 0.0f; h[i][j] = 0.0f;
            if (p.state_in != nullptr && b < B)
                c[i][j] = p.state_in[(((size_t)b * 2 + dir) * 2 + 1) * UNITS + HU * rank + 4 * ug + j];
        }
    }
    float *peer_hs = cluster.map_shared_rank(hs, rank ^ 1);
    cluster.sync();

    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        const float *hc = hs + (s & 1) * KX * BT;
        const int nxt = ((s & 1) ^ 1) * KX * BT;
        float acc[4][4][4];     // [row][gate][unit]
        if (PRE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int b = b0 + 4 * rg + i;
                const float *g = p.G + (((size_t)b * T + t) * 2 + dir) * GATES + HU * rank + 4 * ug;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 v = (b < B) ? __ldg(reinterpret_cast<const float4 *>(g + q * UNITS)) : make_float4(0, 0, 0, 0);
                    acc[i][q][0] = v.x; acc[i][q][1] = v.y; acc[i][q][2] = v.z; acc[i][q][3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][q][j] = 0.0f;
        }
        float xn[(F * BT + THREADS - 1) / THREADS];
        if (!PRE && s + 1 < T) {
            const int tn = dir ? T - 2 - s : s + 1;
#pragma unroll
            for (int e = 0; e < (F * BT + THREADS - 1) / THREADS; ++e) {
                int i = tid + e * THREADS;
                int f = i / BT, row = i % BT, b = b0 + row;
                xn[e] = (i < F * BT && b < B) ? __ldg(p.x + ((size_t)b * T + tn) * F + f) : 0.0f;
            }
        }
#pragma unroll 4
        for (int k = 0; k < KX; ++k) {
            const float4 hv = *reinterpret_cast<const float4 *>(hc + k * BT + 4 * rg);
            const float hr[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 uv = *reinterpret_cast<const float4 *>(Us + k * 256 + q * HU + 4 * ug);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][q][0] = fmaf(hr[i], uv.x, acc[i][q][0]);
                    acc[i][q][1] = fmaf(hr[i], uv.y, acc[i][q][1]);
                    acc[i][q][2] = fmaf(hr[i], uv.z, acc[i][q][2]);
                    acc[i][q][3] = fmaf(hr[i], uv.w, acc[i][q][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float ig = fsig(acc[i][0][j]), fg = fsig(acc[i][1][j]);
                float gg = ftanh(acc[i][2][j]), og = fsig(acc[i][3][j]);
                c[i][j] = fg * c[i][j] + ig * gg;
                h[i][j] = og * ftanh(c[i][j]);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 v = make_float4(h[0][j], h[1][j], h[2][j], h[3][j]);
            int off = nxt + (HU * rank + 4 * ug + j) * BT + 4 * rg;
            *reinterpret_cast<float4 *>(hs + off) = v;
            *reinterpret_cast<float4 *>(peer_hs + off) = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int b = b0 + 4 * rg + i;
            if (b < B)
                *reinterpret_cast<float4 *>(p.y + (size_t)b * p.y_bstride + (size_t)t * ENC_OUT + dir * UNITS + HU * rank + 4 * ug) =
                    make_float4(h[i][0], h[i][1], h[i][2], h[i][3]);
        }
        if (!PRE && s + 1 < T) {
#pragma unroll
            for (int e = 0; e < (F * BT + THREADS - 1) / THREADS; ++e) {
                int i = tid + e * THREADS;
                if (i < F * BT) hs[nxt + (UNITS + i / BT) * BT + (i % BT)] = xn[e];
            }
        }
        cluster.sync();
    }
    if (p.state_out != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int b = b0 + 4 * rg + i;
            if (b < B) {
                float *so = p.state_out + (((size_t)b * 2 + dir) * 2) * UNITS + HU * rank + 4 * ug;
                *reinterpret_cast<float4 *>(so) = make_float4(h[i][0], h[i][1], h[i][2], h[i][3]);
                *reinterpret_cast<float4 *>(so + UNITS) = make_float4(c[i][0], c[i][1], c[i][2], c[i][3]);
            }
        }
    }
}

template <int F, bool PRE>
static int launch(const Params &p, cudaStream_t stream) {
    constexpr int KX = PRE ? UNITS : UNITS + F + 1;
    constexpr size_t smem = sizeof(float) * (KX * 256 + 2 * KX * BT);
    RVB_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<F, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int tiles = (p.B + BT - 1) / BT;
    dim3 grid((unsigned)(tiles * 2 * 2));
    { ProfScope ps(KK_REC, stream);
      lstm_rec_kernel<F, PRE><<<grid, THREADS, smem, stream>>>(p); }
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

// feat: 1 / 5 for layer 0 (x given), 0 for layers fed by pre-projected gates.
int run(int feat, const Params &p, cudaStream_t stream) {
    if (p.B <= 0 || p.T <= 0) return RVB_OK;
    if (feat == 1) return launch<1, false>(p, stream);
    if (feat == 5) return launch<5, false>(p, stream);
    if (feat == 0) return launch<1, true>(p, stream);
    return fail(RVB_ERR_ARG, "lstm_rec: unsupported feature count %d", feat);
}

int kx_rows(int feat) { return feat == 0 ? UNITS : UNITS + feat + 1; }

}  // namespace rec
}  // namespace rvb
