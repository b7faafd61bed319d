// K4 + K5 -- persistent attention decoder with fused greedy / beam search.
//
// Replaces, per decode step, tfa AttentionWrapper.call over the StackedRNNCells LSTM,
// LuongAttention (score, masked softmax, context), the attention Dense layer and the
// `fc` output layer (basecaller.py:83-94, 117-134), and the search
// drivers BasicDecoder+GreedyEmbeddingSampler (basecaller.py:317-330; A.4) and
// BeamSearchDecoder step + gather_tree finalize (basecaller.py:296-315; A.5).
//
// One CTA owns up to RMAX = 32 decoder rows (snippets x beams) for ALL S steps:
// snippets are independent, so there is no grid-level dependency and the whole
// decode is a single launch.  All beams of a snippet live in the same CTA and share
// one read of that snippet's keys/values per step (tfa tile_batch would replicate
// them beam_width times).  Recurrent state (h, c, attention) never leaves shared
// memory; the beam reorder is a shared-memory gather by parent index.
#include "kernels.cuh"

namespace rvb {
namespace dec {

constexpr int RMAX = 32;
constexpr int THREADS = 256;
constexpr int TMAX = 256;          // memory length cap (230 for joint input)
constexpr int WMAX = 9;            // W * VOCAB <= 64 candidates (two per lane)
constexpr float F32_MIN = -3.4028234663852886e38f;   // tf.float32.min, tfa _mask_probs


__device__ __forceinline__ float fsig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 2.0f * fsig(2.0f * x) - 1.0f; }

// (value desc, index asc) warp arg-max; dead candidates carry idx = INT_MAX, val = -inf.
__device__ __forceinline__ void warp_argmax(float &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// tfa _beam_search_step for one batch entry, executed by one warp (A.5).
// total[] = log_probs + masked step log-probs for the W*V candidates (two per lane).
__device__ __forceinline__ void warp_topk(float c0, float c1, int n_cand, int W, int lane,
                                          float *out_score, int *out_idx) {
    float v0 = (lane < n_cand) ? c0 : -INFINITY;        int i0 = (lane < n_cand) ? lane : 0x7fffffff;
    float v1 = (lane + 32 < n_cand) ? c1 : -INFINITY;   int i1 = (lane + 32 < n_cand) ? lane + 32 : 0x7fffffff;
    for (int k = 0; k < W; ++k) {
        float v; int i;
        if (v0 > v1 || (v0 == v1 && i0 < i1)) { v = v0; i = i0; } else { v = v1; i = i1; }
        warp_argmax(v, i);
        if (lane == 0) { out_score[k] = v; out_idx[k] = i; }
        if (i0 == i) { v0 = -INFINITY; i0 = 0x7fffffff; }
        if (i1 == i) { v1 = -INFINITY; i1 = 0x7fffffff; }
    }
}

__global__ void __launch_bounds__(THREADS, 1) decoder_kernel(Params p) {
    extern __shared__ __align__(16) float smem[];
    float *buf = smem;                       // [640][32]: 0..127 prev attention | 128..255 h | 256..511 context
    float *cs = buf + 640 * RMAX;            // [128][32]
    float *attn = cs + UNITS * RMAX;         // [128][32]
    float *sc = attn + UNITS * RMAX;         // [32][TMAX]
    float *wfc_s = sc + RMAX * TMAX;         // [128*7]
    float *logit_s = wfc_s + UNITS * VOCAB;  // [32][8]
    float *lp_s = logit_s + RMAX * 8;        // [32] beam log-probs
    float *nsc_s = lp_s + RMAX;              // [32] new scores
    int *tok_s = reinterpret_cast<int *>(nsc_s + RMAX);   // [32]
    int *fin_s = tok_s + RMAX;               // [32]
    int *len_s = fin_s + RMAX;               // [32]
    int *srow_s = len_s + RMAX;              // [32] source row for the reorder
    int *nidx_s = srow_s + RMAX;             // [32] flat top-k index
    int *first_s = nidx_s + RMAX;            // [32] greedy: first END step ; beam: per-snippet all-finished step

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int W = p.W, S = p.S, Tm = p.Tm;
    const int SN = RMAX / W;
    const int s0 = blockIdx.x * SN;
    const int ns = min(SN, p.B - s0);
    const int R = ns * W;

    for (int i = tid; i < 640 * RMAX; i += THREADS) buf[i] = 0.0f;
    for (int i = tid; i < UNITS * RMAX; i += THREADS) { cs[i] = 0.0f; attn[i] = 0.0f; }
    for (int i = tid; i < UNITS * VOCAB; i += THREADS) wfc_s[i] = p.wfc[i];
    if (tid < RMAX) {
        tok_s[tid] = TOKEN_START;
        int k = tid % W;
        lp_s[tid] = (k == 0) ? 0.0f : -INFINITY;
        fin_s[tid] = (p.beam && k != 0) ? 1 : 0;
        len_s[tid] = 0;
        first_s[tid] = S;          // "never"
        srow_s[tid] = tid;
    }
    __syncthreads();

    const int u = tid & 127, half = tid >> 7;
    for (int t = 0; t < S; ++t) {
        // ---------------- phase 1: LSTM cell on [one_hot(token) | prev attention] ----------------
        {
            float acc[16][4];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                float4 w = __ldg(reinterpret_cast<const float4 *>(p.wtok + ((size_t)tok_s[half * 16 + r] * UNITS + u) * 4));
                acc[r][0] = w.x; acc[r][1] = w.y; acc[r][2] = w.z; acc[r][3] = w.w;
            }
#pragma unroll 4
            for (int k = 0; k < 2 * UNITS; ++k) {
                const float4 w = __ldg(reinterpret_cast<const float4 *>(p.wg + ((size_t)k * UNITS + u) * 4));
                const float4 *xr = reinterpret_cast<const float4 *>(buf + k * RMAX + half * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 x = xr[q];
                    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc[q * 4 + e][0] = fmaf(xv[e], w.x, acc[q * 4 + e][0]);
                        acc[q * 4 + e][1] = fmaf(xv[e], w.y, acc[q * 4 + e][1]);
                        acc[q * 4 + e][2] = fmaf(xv[e], w.z, acc[q * 4 + e][2]);
                        acc[q * 4 + e][3] = fmaf(xv[e], w.w, acc[q * 4 + e][3]);
                    }
                }
            }
            __syncthreads();                 // every read of the old h rows is done
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                int row = half * 16 + r;
                float c = cs[u * RMAX + row];
                float ig = fsig(acc[r][0]), fg = fsig(acc[r][1]), gg = ftanh(acc[r][2]), og = fsig(acc[r][3]);
                c = fg * c + ig * gg;
                cs[u * RMAX + row] = c;
                buf[(UNITS + u) * RMAX + row] = og * ftanh(c);
            }
        }
        __syncthreads();

        // ---------------- phase 2: Luong score -> masked softmax -> context, one warp per snippet ----
        for (int s = wid; s < ns; s += THREADS / 32) {
            const size_t bm = (size_t)(s0 + s) * Tm;
            float hq[WMAX][4];
#pragma unroll
            for (int w = 0; w < WMAX; ++w)
                if (w < W) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) hq[w][e] = buf[(UNITS + 4 * lane + e) * RMAX + s * W + w];
                }
            for (int tt = 0; tt < Tm; ++tt) {
                const float4 kv = __ldg(reinterpret_cast<const float4 *>(p.keys + (bm + tt) * UNITS + 4 * lane));
                const bool valid = p.mask[bm + tt] != 0;
#pragma unroll
                for (int w = 0; w < WMAX; ++w)
                    if (w < W) {
                        float d = kv.x * hq[w][0] + kv.y * hq[w][1] + kv.z * hq[w][2] + kv.w * hq[w][3];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                        if (lane == 0) sc[(s * W + w) * TMAX + tt] = valid ? d : -INFINITY;
                    }
            }
            __syncwarp();
            for (int w = 0; w < W; ++w) {
                float *row = sc + (s * W + w) * TMAX;
                float m = -INFINITY;
                for (int tt = lane; tt < Tm; tt += 32) m = fmaxf(m, row[tt]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float sum = 0.0f;
                for (int tt = lane; tt < Tm; tt += 32) { float e = __expf(row[tt] - m); row[tt] = e; sum += e; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float inv = 1.0f / sum;
                for (int tt = lane; tt < Tm; tt += 32) row[tt] *= inv;
            }
            __syncwarp();
            float ctx[WMAX][8];
#pragma unroll
            for (int w = 0; w < WMAX; ++w)
#pragma unroll
                for (int e = 0; e < 8; ++e) ctx[w][e] = 0.0f;
            for (int tt = 0; tt < Tm; ++tt) {
                const float *vp = p.values + (bm + tt) * ENC_OUT + 4 * lane;
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(vp));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(vp + UNITS));
#pragma unroll
                for (int w = 0; w < WMAX; ++w)
                    if (w < W) {
                        const float a = sc[(s * W + w) * TMAX + tt];
                        ctx[w][0] = fmaf(a, v0.x, ctx[w][0]); ctx[w][1] = fmaf(a, v0.y, ctx[w][1]);
                        ctx[w][2] = fmaf(a, v0.z, ctx[w][2]); ctx[w][3] = fmaf(a, v0.w, ctx[w][3]);
                        ctx[w][4] = fmaf(a, v1.x, ctx[w][4]); ctx[w][5] = fmaf(a, v1.y, ctx[w][5]);
                        ctx[w][6] = fmaf(a, v1.z, ctx[w][6]); ctx[w][7] = fmaf(a, v1.w, ctx[w][7]);
                    }
            }
#pragma unroll
            for (int w = 0; w < WMAX; ++w)
                if (w < W) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        buf[(2 * UNITS + 4 * lane + e) * RMAX + s * W + w] = ctx[w][e];
                        buf[(3 * UNITS + 4 * lane + e) * RMAX + s * W + w] = ctx[w][4 + e];
                    }
                }
        }
        __syncthreads();

        // ---------------- phase 3: attention = [h | context] . W_att ----------------------------------
        {
            float acc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) acc[r] = 0.0f;
#pragma unroll 4
            for (int k = 0; k < 3 * UNITS; ++k) {
                const float w = __ldg(p.watt + (size_t)k * UNITS + u);
                const float4 *xr = reinterpret_cast<const float4 *>(buf + (UNITS + k) * RMAX + half * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 x = xr[q];
                    acc[q * 4 + 0] = fmaf(x.x, w, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(x.y, w, acc[q * 4 + 1]);
                    acc[q * 4 + 2] = fmaf(x.z, w, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(x.w, w, acc[q * 4 + 3]);
                }
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) attn[u * RMAX + half * 16 + r] = acc[r];
        }
        __syncthreads();

        // ---------------- phase 4: logits = attention . fc + b -----------------------------------------
        if (tid < R * VOCAB) {
            const int r = tid / VOCAB, v = tid % VOCAB;
            float a = p.bfc[v];
#pragma unroll 8
            for (int k = 0; k < UNITS; ++k) a = fmaf(attn[k * RMAX + r], wfc_s[k * VOCAB + v], a);
            logit_s[r * 8 + v] = a;
        }
        __syncthreads();

        // ---------------- phase 5: search step ----------------------------------------------------------
        if (!p.beam) {
            if (tid < R) {
                const float *lg = logit_s + tid * 8;
                int best = 0; float bv = lg[0];
#pragma unroll
                for (int v = 1; v < VOCAB; ++v) if (lg[v] > bv) { bv = lg[v]; best = v; }   // ties -> lowest id
                const size_t b = (size_t)(s0 + tid);
                p.ids[b * S + t] = best;
#pragma unroll
                for (int v = 0; v < VOCAB; ++v) p.logits[(b * S + t) * VOCAB + v] = lg[v];
                tok_s[tid] = best;
                if (best == TOKEN_END && first_s[tid] == S) first_s[tid] = t;
            }
        } else {
            for (int s = wid; s < ns; s += THREADS / 32) {
                const int n_cand = W * VOCAB;
                float cand[2];
#pragma unroll
                for (int hcand = 0; hcand < 2; ++hcand) {
                    const int i = lane + 32 * hcand;
                    float tot = -INFINITY;
                    if (i < n_cand) {
                        const int k = i / VOCAB, v = i % VOCAB, row = s * W + k;
                        float slp;
                        if (fin_s[row]) slp = (v == TOKEN_END) ? 0.0f : F32_MIN;
                        else {
                            const float *lg = logit_s + row * 8;
                            float m = lg[0];
#pragma unroll
                            for (int q = 1; q < VOCAB; ++q) m = fmaxf(m, lg[q]);
                            float se = 0.0f;
#pragma unroll
                            for (int q = 0; q < VOCAB; ++q) se += expf(lg[q] - m);
                            slp = (lg[v] - m) - logf(se);
                        }
                        tot = lp_s[row] + slp;
                    }
                    cand[hcand] = tot;
                }
                __syncwarp();
                warp_topk(cand[0], cand[1], n_cand, W, lane, nsc_s + s * W, nidx_s + s * W);
                __syncwarp();
                int nfin = 0, nlen = 0, word = 0, parent = 0; float nlp = 0.0f;
                if (lane < W) {
                    const int idx = nidx_s[s * W + lane];
                    word = idx % VOCAB; parent = idx / VOCAB;
                    const int prow = s * W + parent;
                    const int pf = fin_s[prow];
                    nfin = pf | (word == TOKEN_END);
                    nlen = len_s[prow] + (pf ? 0 : 1);
                    nlp = nsc_s[s * W + lane];
                }
                __syncwarp();
                if (lane < W) {
                    const int row = s * W + lane;
                    fin_s[row] = nfin; len_s[row] = nlen; lp_s[row] = nlp; tok_s[row] = word;
                    srow_s[row] = s * W + parent;
                    const size_t o = ((size_t)(s0 + s) * S + t) * W + lane;
                    p.scores[o] = nlp; p.step_ids[o] = word; p.parent_ids[o] = parent;
                }
                const unsigned allfin = __ballot_sync(0xffffffffu, lane >= W || nfin);
                if (lane == 0 && allfin == 0xffffffffu && first_s[s] == S) first_s[s] = t;
            }
        }
        __syncthreads();

        // ---------------- state hand-over to the next step (beam: gather by parent) ---------------------
        if (!p.beam) {
            for (int i = tid; i < UNITS * RMAX; i += THREADS) buf[i] = attn[i];
        } else {
            float hv[16], cv[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int i = tid + e * THREADS, d = i >> 5, r = i & 31, sr = srow_s[r];
                hv[e] = buf[(UNITS + d) * RMAX + sr];
                cv[e] = cs[d * RMAX + sr];
                buf[d * RMAX + r] = attn[d * RMAX + sr];
            }
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int i = tid + e * THREADS, d = i >> 5, r = i & 31;
                buf[(UNITS + d) * RMAX + r] = hv[e];
                cs[d * RMAX + r] = cv[e];
            }
        }
        __syncthreads();
    }

    // ---------------- T = steps tfa's dynamic_decode would have executed ------------------------------
    if (tid == 0) {
        int tmax = 0;
        const int n = p.beam ? ns : R;
        for (int i = 0; i < n; ++i) tmax = max(tmax, min(first_s[i] + 1, S));
        atomicMax(p.steps, tmax);
    }
    // ---------------- finalize: gather_tree (A.5) ------------------------------------------------------
    if (p.beam) {
        __threadfence_block();
        __syncthreads();
        if (tid < R) {
            const int s = tid / W, k = tid % W;
            int maxlen = 0;
            for (int q = 0; q < W; ++q) maxlen = max(maxlen, len_s[s * W + q]);
            const int L = min(S, maxlen);
            const size_t o = (size_t)(s0 + s) * S * W;
            for (int tt = L; tt < S; ++tt) p.ids[o + (size_t)tt * W + k] = TOKEN_END;
            int parent = k;
            for (int level = L - 1; level >= 0; --level) {
                p.ids[o + (size_t)level * W + k] = p.step_ids[o + (size_t)level * W + parent];
                parent = p.parent_ids[o + (size_t)level * W + parent];
            }
            bool done = false;
            for (int tt = 0; tt < L; ++tt) {
                if (done) p.ids[o + (size_t)tt * W + k] = TOKEN_END;
                else if (p.ids[o + (size_t)tt * W + k] == TOKEN_END) done = true;
            }
        }
    }
}

constexpr size_t SMEM_FLOATS = 640 * RMAX + 2 * UNITS * RMAX + RMAX * TMAX + UNITS * VOCAB + RMAX * 8 + 2 * RMAX + 6 * RMAX;

int run(const Params &p, cudaStream_t stream) {
    if (p.B <= 0 || p.S <= 0) return RVB_OK;
    if (p.Tm > TMAX) return fail(RVB_ERR_ARG, "decoder: memory length %d > %d", p.Tm, TMAX);
    if (p.W < 1 || p.W > WMAX) return fail(RVB_ERR_ARG, "decoder: beam width must be in [1,%d]", WMAX);
    const size_t smem = SMEM_FLOATS * sizeof(float);
    RVB_CUDA(cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int SN = RMAX / p.W;
    dim3 grid((unsigned)((p.B + SN - 1) / SN));
    decoder_kernel<<<grid, THREADS, smem, stream>>>(p);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K5 standalone: one beam step on log-softmaxed rows / gather_tree (parity tests call these directly)
// ---------------------------------------------------------------------------------------------------
__global__ void beam_step_kernel(const float *slp, const float *lp, const uint8_t *fin, const long long *len,
                                 long long B, int W, int V, int end_token, float *scores, int32_t *word,
                                 int32_t *parent, uint8_t *nfin, long long *nlen) {
    __shared__ float sc_s[8][WMAX];
    __shared__ int idx_s[8][WMAX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.x * 8 + wid;
    if (b >= B) return;
    const int n_cand = W * V;
    float cand[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        float tot = -INFINITY;
        if (i < n_cand) {
            const int k = i / V, v = i % V;
            float s = fin[b * W + k] ? ((v == end_token) ? 0.0f : F32_MIN) : slp[(b * W + k) * V + v];
            tot = lp[b * W + k] + s;
        }
        cand[h] = tot;
    }
    warp_topk(cand[0], cand[1], n_cand, W, lane, sc_s[wid], idx_s[wid]);
    __syncwarp();
    if (lane < W) {
        const int idx = idx_s[wid][lane];
        const int wd = idx % V, pr = idx / V;
        const bool pf = fin[b * W + pr] != 0;
        scores[b * W + lane] = sc_s[wid][lane];
        word[b * W + lane] = wd;
        parent[b * W + lane] = pr;
        nfin[b * W + lane] = (pf || wd == end_token) ? 1 : 0;
        nlen[b * W + lane] = len[b * W + pr] + (pf ? 0 : 1);
    }
}

// step_ids / parent_ids / out: time-major [T,B,W] as tfa.seq2seq.gather_tree.
__global__ void gather_tree_kernel(const int32_t *step_ids, const int32_t *parent_ids, const int32_t *max_len,
                                   int T, long long B, int W, int end_token, int32_t *out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * W) return;
    const long long b = g / W; const int k = (int)(g % W);
    const int L = max(0, min(T, max_len[b]));
    const size_t stride = (size_t)B * W;
    for (int t = L; t < T; ++t) out[t * stride + b * W + k] = end_token;
    int parent = k;
    for (int level = L - 1; level >= 0; --level) {
        out[level * stride + b * W + k] = step_ids[level * stride + b * W + parent];
        parent = parent_ids[level * stride + b * W + parent];
    }
    bool done = false;
    for (int t = 0; t < L; ++t) {
        int32_t *o = out + t * stride + b * W + k;
        if (done) *o = end_token;
        else if (*o == end_token) done = true;
    }
}

}  // namespace dec
}  // namespace rvb

using namespace rvb;

extern "C" int rvb_beam_step(const float *d_slp, const float *d_lp, const uint8_t *d_fin, const int64_t *d_len,
                             int64_t batch, int W, int V, int end_token, float *d_scores, int32_t *d_word,
                             int32_t *d_parent, uint8_t *d_nfin, int64_t *d_nlen, void *stream) {
    if (batch < 0 || W < 1 || W > dec::WMAX || V < 1 || W * V > 64) return fail(RVB_ERR_ARG, "beam_step: need 1 <= W*V <= 64");
    if (batch == 0) return RVB_OK;
    dec::beam_step_kernel<<<(unsigned)((batch + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        d_slp, d_lp, d_fin, reinterpret_cast<const long long *>(d_len), batch, W, V, end_token, d_scores, d_word,
        d_parent, d_nfin, reinterpret_cast<long long *>(d_nlen));
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

extern "C" int rvb_gather_tree(const int32_t *d_step_ids, const int32_t *d_parent_ids, const int32_t *d_max_len,
                               int steps, int64_t batch, int W, int end_token, int32_t *d_out, void *stream) {
    if (batch < 0 || W < 1 || steps < 0) return fail(RVB_ERR_ARG, "gather_tree: bad shape");
    if (batch == 0 || steps == 0) return RVB_OK;
    const long long n = batch * W;
    dec::gather_tree_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_step_ids, d_parent_ids, d_max_len, steps, batch, W, end_token, d_out);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}
