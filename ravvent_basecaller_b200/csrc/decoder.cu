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
#include <cuda_fp16.h>

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

// One memory row as seen by a lane: 8 of the 256 columns.
//   fp32 memory: columns 4*lane..+3 and 128+4*lane..+3 (two 128-bit loads, each coalesced across the warp)
//   fp16 memory (reduced-precision mode): columns 8*lane..+7 (one 128-bit load)
template <bool VH> struct MemRow;
template <> struct MemRow<false> {
    float4 a, b;
    __device__ __forceinline__ void zero() { a = b = make_float4(0, 0, 0, 0); }
    __device__ __forceinline__ void load(const void *base, size_t row, int lane) {
        const float *vp = reinterpret_cast<const float *>(base) + row * ENC_OUT + 4 * lane;
        a = __ldg(reinterpret_cast<const float4 *>(vp));
        b = __ldg(reinterpret_cast<const float4 *>(vp + UNITS));
    }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ static __forceinline__ int col(int lane, int e) { return e < 4 ? 4 * lane + e : UNITS + 4 * lane + (e - 4); }
};
template <> struct MemRow<true> {
    uint4 r;
    __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ void load(const void *base, size_t row, int lane) {
        r = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(base) + row * ENC_OUT + 8 * lane));
    }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ static __forceinline__ int col(int lane, int e) { return 8 * lane + e; }
};

template <int WT, int RM, int MINB, bool VH>
__global__ void __launch_bounds__(THREADS, MINB) decoder_kernel(Params p) {
    extern __shared__ __align__(16) float smem[];
    float *buf = smem;                       // [640][32]: 0..127 prev attention | 128..255 h | 256..511 context
    float *cs = buf + 640 * RM;            // [cell][128][RM] cell states of the (up to 2) stacked LSTM cells
    float *attn = cs + 2 * UNITS * RM;     // [128][RM]
    float *qs = attn + UNITS * RM;         // [256][32] folded query q' = W_mem . h
    float *wfc_s = qs + ENC_OUT * RM;      // [128*7]
    float *logit_s = wfc_s + UNITS * VOCAB;  // [32][8]
    float *lp_s = logit_s + 64 * 8;        // [32] beam log-probs
    float *nsc_s = lp_s + 64;              // [32] new scores
    int *tok_s = reinterpret_cast<int *>(nsc_s + 64);   // [32]
    int *fin_s = tok_s + 64;               // [32]
    int *len_s = fin_s + 64;               // [32]
    int *srow_s = len_s + 64;              // [32] source row for the reorder
    int *nidx_s = srow_s + 64;             // [32] flat top-k index
    int *first_s = nidx_s + 64;            // [32] greedy: first END step ; beam: per-snippet all-finished step

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int W = p.W, S = p.S, Tm = p.Tm;
    const int SN = RM / W;
    const int s0 = blockIdx.x * SN;
    const int ns = min(SN, p.B - s0);
    const int R = ns * W;

    for (int i = tid; i < 640 * RM; i += THREADS) buf[i] = 0.0f;
    for (int i = tid; i < 2 * UNITS * RM; i += THREADS) cs[i] = 0.0f;
    for (int i = tid; i < UNITS * RM; i += THREADS) attn[i] = 0.0f;
    for (int i = tid; i < UNITS * VOCAB; i += THREADS) wfc_s[i] = p.wfc[i];
    if (tid < 64) {
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
    // buf rows: [0,128) previous attention | [128*(j+1), 128*(j+2)) h of stacked cell j | then 256 rows of context.
    // QROW = first row of the top cell's h (the attention query); [QROW, QROW+384) = [query | context] feeds phase 3.
    const int depth = p.depth;
    const int QROW = UNITS * depth;
    for (int t = 0; t < S; ++t) {
        // ---------------- phase 1: stacked LSTM cells (StackedRNNCells, basecaller.py:85-91) -------------
        // cell 0 sees [one_hot(token) | prev attention], cell j > 0 sees the new h of cell j-1.
        for (int cell = 0; cell < depth; ++cell) {
            constexpr int RH = RM / 2;
            const float *wcell = cell == 0 ? p.wg : p.wg1;
            const float *xin = buf + (size_t)cell * UNITS * RM;          // rows [128*cell, 128*cell + 256)
            float *ccell = cs + (size_t)cell * UNITS * RM;
            float acc[RH][4];
#pragma unroll
            for (int r = 0; r < RH; ++r) {
                const float4 w = cell == 0
                    ? __ldg(reinterpret_cast<const float4 *>(p.wtok + ((size_t)tok_s[half * RH + r] * UNITS + u) * 4))
                    : __ldg(reinterpret_cast<const float4 *>(p.b1 + (size_t)u * 4));
                acc[r][0] = w.x; acc[r][1] = w.y; acc[r][2] = w.z; acc[r][3] = w.w;
            }
#pragma unroll 8
            for (int k = 0; k < 2 * UNITS; ++k) {
                const float4 w = __ldg(reinterpret_cast<const float4 *>(wcell + ((size_t)k * UNITS + u) * 4));
                const float4 *xr = reinterpret_cast<const float4 *>(xin + k * RM + half * RH);
#pragma unroll
                for (int q = 0; q < RH / 4; ++q) {
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
            __syncthreads();                 // every read of this cell's old h rows is done
#pragma unroll
            for (int r = 0; r < RH; ++r) {
                int row = half * RH + r;
                float c = ccell[u * RM + row];
                float ig = fsig(acc[r][0]), fg = fsig(acc[r][1]), gg = ftanh(acc[r][2]), og = fsig(acc[r][3]);
                c = fg * c + ig * gg;
                ccell[u * RM + row] = c;
                buf[((cell + 1) * UNITS + u) * RM + row] = og * ftanh(c);
            }
            __syncthreads();
        }

        // ---------------- phase 2a: q' = W_mem . h  (Luong score = keys.h = values.(W_mem.h)) ---------
        // Folding the memory layer into the query means only `values` is streamed per step (A.3).
        {
            float acc[RM];
#pragma unroll
            for (int r = 0; r < RM; ++r) acc[r] = 0.0f;
#pragma unroll 8
            for (int k = 0; k < UNITS; ++k) {
                const float w = __ldg(p.wmemT + (size_t)k * ENC_OUT + tid);
                const float4 *xr = reinterpret_cast<const float4 *>(buf + (QROW + k) * RM);
#pragma unroll
                for (int q = 0; q < RM / 4; ++q) {
                    const float4 x = xr[q];
                    acc[q * 4 + 0] = fmaf(x.x, w, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(x.y, w, acc[q * 4 + 1]);
                    acc[q * 4 + 2] = fmaf(x.z, w, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(x.w, w, acc[q * 4 + 3]);
                }
            }
#pragma unroll
            for (int q = 0; q < RM / 4; ++q)
                *reinterpret_cast<float4 *>(qs + tid * RM + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        }
        __syncthreads();

        // ---------------- phase 2b: masked softmax(values.q') and context in ONE pass over values -----
        // One warp per snippet, online softmax (running max / sum), all beams of the snippet share the
        // stream.  A lane owns 8 of the 256 columns (MemRow); rows are software-pipelined in groups of 4
        // with the next group prefetched.
        for (int s = wid; s < ns; s += THREADS / 32) {
            const size_t bm = (size_t)(s0 + s) * Tm;
            const void *vmem = VH ? reinterpret_cast<const void *>(p.values16) : reinterpret_cast<const void *>(p.values);
            unsigned mbits = 0;                                   // validity of rows 8*lane .. 8*lane+7
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int tt = 8 * lane + j;
                if (tt < Tm && p.mask[bm + tt] != 0) mbits |= 1u << j;
            }
            float q[WT][8], acc[WT][8], mx[WT], den[WT];
#pragma unroll
            for (int w = 0; w < WT; ++w) {
                mx[w] = -INFINITY; den[w] = 0.0f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    acc[w][e] = 0.0f;
                    q[w][e] = (w < W) ? qs[MemRow<VH>::col(lane, e) * RM + s * W + w] : 0.0f;
                }
            }
            MemRow<VH> cur[4], nxt[4];
            unsigned vb_cur = __shfl_sync(0xffffffffu, mbits, 0) & 0xFu, vb_nxt = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cur[j].zero();
                if ((vb_cur >> j) & 1u) cur[j].load(vmem, bm + j, lane);
            }
            for (int t0 = 0; t0 < Tm; t0 += 4) {
                const int t1 = t0 + 4;
                vb_nxt = 0;
                if (t1 < Tm) vb_nxt = (__shfl_sync(0xffffffffu, mbits, t1 >> 3) >> (t1 & 7)) & 0xFu;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    nxt[j].zero();
                    if ((vb_nxt >> j) & 1u) nxt[j].load(vmem, bm + t1 + j, lane);
                }
                if (vb_cur != 0) {
                    float vv[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) cur[j].unpack(vv[j]);
#pragma unroll
                    for (int w = 0; w < WT; ++w)
                        if (w < W) {
                            float sj[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float d = vv[j][0] * q[w][0];
#pragma unroll
                                for (int e = 1; e < 8; ++e) d = fmaf(vv[j][e], q[w][e], d);
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                                sj[j] = ((vb_cur >> j) & 1u) ? d : -INFINITY;
                            }
                            const float mn = fmaxf(fmaxf(mx[w], fmaxf(sj[0], sj[1])), fmaxf(sj[2], sj[3]));
                            const float scale = __expf(mx[w] - mn);          // mx == -inf -> 0
                            float pj[4], ps = 0.0f;
#pragma unroll
                            for (int j = 0; j < 4; ++j) { pj[j] = __expf(sj[j] - mn); ps += pj[j]; }
                            den[w] = den[w] * scale + ps;
                            mx[w] = mn;
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[w][e] *= scale;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[w][e] = fmaf(pj[j], vv[j][e], acc[w][e]);
                        }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
                vb_cur = vb_nxt;
            }
#pragma unroll
            for (int w = 0; w < WT; ++w)
                if (w < W) {
                    // every position masked: tfa's softmax over all -inf yields NaN; keep that contract
                    const float inv = (den[w] > 0.0f) ? 1.0f / den[w] : __int_as_float(0x7fc00000);
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        buf[(QROW + UNITS + MemRow<VH>::col(lane, e)) * RM + s * W + w] = acc[w][e] * inv;
                }
        }
        __syncthreads();

        // ---------------- phase 3: attention = [h | context] . W_att ----------------------------------
        {
            constexpr int RH = RM / 2;
            float acc[RH];
#pragma unroll
            for (int r = 0; r < RH; ++r) acc[r] = 0.0f;
#pragma unroll 8
            for (int k = 0; k < 3 * UNITS; ++k) {
                const float w = __ldg(p.watt + (size_t)k * UNITS + u);
                const float4 *xr = reinterpret_cast<const float4 *>(buf + (QROW + k) * RM + half * RH);
#pragma unroll
                for (int q = 0; q < RH / 4; ++q) {
                    const float4 x = xr[q];
                    acc[q * 4 + 0] = fmaf(x.x, w, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(x.y, w, acc[q * 4 + 1]);
                    acc[q * 4 + 2] = fmaf(x.z, w, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(x.w, w, acc[q * 4 + 3]);
                }
            }
#pragma unroll
            for (int r = 0; r < RH; ++r) attn[u * RM + half * RH + r] = acc[r];
        }
        __syncthreads();

        // ---------------- phase 4: logits = attention . fc + b -----------------------------------------
        for (int o = tid; o < R * VOCAB; o += THREADS) {
            const int r = o / VOCAB, v = o % VOCAB;
            float a = p.bfc[v];
#pragma unroll 8
            for (int k = 0; k < UNITS; ++k) a = fmaf(attn[k * RM + r], wfc_s[k * VOCAB + v], a);
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

        // ---------------- early exit (beam search): every beam of every snippet of this CTA has finished -----
        // A finished beam can only continue with the end token at unchanged score, and top_k keeps the
        // (already descending) order, so the remaining steps are known: scores = current log-probs,
        // predicted id = end token, parent = identity (A.5).  Fill them and stop.
        if (p.beam) {
            bool all_done = true;
            for (int s = 0; s < ns; ++s) all_done = all_done && (first_s[s] != S);
            if (all_done) {
                for (int o = tid; o < R * (S - 1 - t); o += THREADS) {
                    const int row = o % R, tt = t + 1 + o / R;
                    const int s = row / W, k = row % W;
                    const size_t g = ((size_t)(s0 + s) * S + tt) * W + k;
                    p.scores[g] = lp_s[row]; p.step_ids[g] = TOKEN_END; p.parent_ids[g] = k;
                }
                break;
            }
        }

        // ---------------- state hand-over to the next step (beam: gather by parent) ---------------------
        if (!p.beam) {
            for (int i = tid; i < UNITS * RM; i += THREADS) buf[i] = attn[i];
        } else {
            constexpr int NE = UNITS * RM / THREADS;
            for (int cell = 0; cell < depth; ++cell) {
                float *hcell = buf + (size_t)(cell + 1) * UNITS * RM;
                float *ccell = cs + (size_t)cell * UNITS * RM;
                float hv[NE], cv[NE];
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int i = tid + e * THREADS, d = i / RM, r = i % RM, sr = srow_s[r];
                    hv[e] = hcell[d * RM + sr];
                    cv[e] = ccell[d * RM + sr];
                    if (cell == 0) buf[d * RM + r] = attn[d * RM + sr];
                }
                __syncthreads();
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int i = tid + e * THREADS, d = i / RM, r = i % RM;
                    hcell[d * RM + r] = hv[e];
                    ccell[d * RM + r] = cv[e];
                }
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

template <int RM>
constexpr size_t smem_floats() { return (size_t)640 * RM + 3 * UNITS * RM + ENC_OUT * RM + UNITS * VOCAB + 64 * 8 + 2 * 64 + 6 * 64; }

template <int WT, int RM, int MINB, bool VH>
static int launch_v(const Params &p, cudaStream_t stream) {
    const size_t smem = smem_floats<RM>() * sizeof(float);
    RVB_CUDA(cudaFuncSetAttribute(decoder_kernel<WT, RM, MINB, VH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int SN = RM / p.W;
    dim3 grid((unsigned)((p.B + SN - 1) / SN));
    { ProfScope ps(KK_DECODER, stream);
      decoder_kernel<WT, RM, MINB, VH><<<grid, THREADS, smem, stream>>>(p); }
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

// values16 != nullptr selects the fp16 attention memory (reduced-precision mode)
template <int WT, int RM, int MINB>
static int launch(const Params &p, cudaStream_t stream) {
    return p.values16 != nullptr ? launch_v<WT, RM, MINB, true>(p, stream) : launch_v<WT, RM, MINB, false>(p, stream);
}

int run(const Params &p, cudaStream_t stream) {
    if (p.B <= 0 || p.S <= 0) return RVB_OK;
    if (p.Tm > TMAX) return fail(RVB_ERR_ARG, "decoder: memory length %d > %d", p.Tm, TMAX);
    if (p.W < 1 || p.W > WMAX) return fail(RVB_ERR_ARG, "decoder: beam width must be in [1,%d]", WMAX);
    if (p.depth < 1 || p.depth > 2) return fail(RVB_ERR_ARG, "decoder: depth must be 1 or 2");
    // beam 1 / greedy: 16 rows per CTA and two CTAs per SM, so that one CTA's dense phases overlap the
    // other's HBM streaming; wider beams: 32 rows (all beams of a snippet stay in one CTA)
    if (p.W == 1) return launch<1, 16, 2>(p, stream);
    if (p.W == 5) return launch<5, 40, 1>(p, stream);      // 8 snippets x 5 beams: every warp owns a snippet in the attention pass
    if (p.W < 5) return launch<5, 32, 1>(p, stream);
    return launch<9, 32, 1>(p, stream);
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
