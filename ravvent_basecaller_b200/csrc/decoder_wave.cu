// K4 + K5, wave-level decoder: beam search at every width and greedy search, one or two stacked LSTM / GRU cells.
//
// tfa AttentionWrapper + LuongAttention + BeamSearchDecoder / BasicDecoder semantics (reference
// basecaller.py:296-315) but organised per decode step over ALL rows of a wave
// (rows = snippets x beams, ~47 000 for a 9 472-snippet wave at beam 5), so that the dense parts run on the
// tcgen05 projection kernel (K2, 3xTF32 = fp32 accuracy) instead of per-CTA FFMA loops that stall on L2 weight
// streams.  Per step:
//   (X[r] = [attention_prev[src(r)] | h_prev[src(r)]], src(r) = row of the parent beam, is written by the previous step's fc_search)
//   GEMM + cell     h, c   = LSTM(X . [W_att_in ; U] + W_token[token] + b, c_prev[src])   (K 256, N 512; the cell update is
//                            the GEMM's epilogue, gate columns in [unit][gate] order)     h -> XA[:, 0:128]
//   GEMM            Q'     = h . W_mem^T                                     (K 128, N 256; folded Luong query)
//   attention       ctx    = softmax_mask(values . q') . values              one warp per snippet, beams share the stream
//   GEMM            A      = [h | ctx] . W_attention_layer                   (K 384, N 128)
//   fc_search       logits = A . fc + b; tfa _beam_search_step (warp top-k); per-step outputs, next token / parent, next X
// and after the last step gather_tree.  The beam reorder never moves state: consumers read rows through src(r).
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace rvb {
namespace decw {

constexpr int WMAX = 9;
constexpr float F32_MIN = -3.4028234663852886e38f;


// ---- packed fp32 pairs: sm_100a executes two fp32 FMAs per issue slot (fma.rn.f32x2 -> FFMA2); nvcc does not pair them
//      on its own, and the attention kernel below is issue bound on exactly these FMAs -------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// A memory row as this lane sees it: columns 4*lane..+3 and 128 + 4*lane..+3 as four packed pairs.
struct Row8 { f32x2 p[4]; };
__device__ __forceinline__ Row8 load_row8(const float *src) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src)), c = __ldg(reinterpret_cast<const uint4 *>(src + UNITS));
    Row8 r;
    r.p[0] = (f32x2)a.x | ((f32x2)a.y << 32); r.p[1] = (f32x2)a.z | ((f32x2)a.w << 32);
    r.p[2] = (f32x2)c.x | ((f32x2)c.y << 32); r.p[3] = (f32x2)c.z | ((f32x2)c.w << 32);
    return r;
}

// fp16 memory rows (reduced-precision mode: half the decode-time HBM bytes), widened to fp32 in registers
__device__ __forceinline__ Row8 load_row8(const __half *src) {
    const uint2 a = __ldg(reinterpret_cast<const uint2 *>(src)), c = __ldg(reinterpret_cast<const uint2 *>(src + UNITS));
    auto widen = [](uint32_t v) -> f32x2 {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v));
        return pack2(f.x, f.y);
    };
    Row8 r;
    r.p[0] = widen(a.x); r.p[1] = widen(a.y); r.p[2] = widen(c.x); r.p[3] = widen(c.y);
    return r;
}

// ---- masked softmax(values . q') . values, one warp per snippet, single-pass online softmax ----------
template <int WT, typename VT>
__global__ void __launch_bounds__(128) attention_kernel(const VT *__restrict__ values, const uint8_t *__restrict__ mask,
                                                        const float *__restrict__ Q, float *__restrict__ xa, int B, int Tm, int W,
                                                        const int32_t *__restrict__ skip, uint16_t *__restrict__ xp_hi, uint16_t *__restrict__ xp_lo) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    if (skip[b]) return;          // every beam of this snippet has finished: fc_search emits the known continuation, nothing reads ctx
    const size_t bm = (size_t)b * Tm;
    unsigned mbits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int tt = 8 * lane + j;
        if (tt < Tm && mask[bm + tt] != 0) mbits |= 1u << j;
    }
    f32x2 q[WT][4], acc[WT][4];
    float mx[WT], den[WT];
#pragma unroll
    for (int w = 0; w < WT; ++w) {
        mx[w] = -INFINITY; den[w] = 0.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[w][e] = 0ull; q[w][e] = 0ull; }
        if (w < W) {
            const Row8 r = load_row8(Q + ((size_t)b * W + w) * ENC_OUT + 4 * lane);
#pragma unroll
            for (int e = 0; e < 4; ++e) q[w][e] = r.p[e];
        }
    }
    const VT *vbase = values + bm * ENC_OUT + 4 * lane;
    Row8 cur[4], nxt[4];
    unsigned vb_cur = __shfl_sync(0xffffffffu, mbits, 0) & 0xFu, vb_nxt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cur[j].p[0] = cur[j].p[1] = cur[j].p[2] = cur[j].p[3] = 0ull;
        if ((vb_cur >> j) & 1u) cur[j] = load_row8(vbase + (size_t)j * ENC_OUT);
    }
    for (int t0 = 0; t0 < Tm; t0 += 4) {
        const int t1 = t0 + 4;
        vb_nxt = 0;
        if (t1 < Tm) vb_nxt = (__shfl_sync(0xffffffffu, mbits, t1 >> 3) >> (t1 & 7)) & 0xFu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            nxt[j].p[0] = nxt[j].p[1] = nxt[j].p[2] = nxt[j].p[3] = 0ull;
            if ((vb_nxt >> j) & 1u) nxt[j] = load_row8(vbase + (size_t)(t1 + j) * ENC_OUT);
        }
        if (vb_cur != 0) {
            // all WT x 4 dot products first, then ONE interleaved butterfly over them: the reductions of different beams
            // are independent, and issuing them together hides the shuffle latency that a per-beam sequence exposes
            float d[WT][4];
#pragma unroll
            for (int w = 0; w < WT; ++w)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    f32x2 d2 = mul2(cur[j].p[0], q[w][0]);
                    d2 = fma2(cur[j].p[1], q[w][1], d2);
                    d2 = fma2(cur[j].p[2], q[w][2], d2);
                    d2 = fma2(cur[j].p[3], q[w][3], d2);
                    float dl, dh;
                    unpack2(d2, dl, dh);
                    d[w][j] = dl + dh;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int w = 0; w < WT; ++w)
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[w][j] += __shfl_xor_sync(0xffffffffu, d[w][j], o);
#pragma unroll
            for (int w = 0; w < WT; ++w)
                if (w < W) {
                    float sj[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) sj[j] = ((vb_cur >> j) & 1u) ? d[w][j] : -INFINITY;
                    const float mn = fmaxf(fmaxf(mx[w], fmaxf(sj[0], sj[1])), fmaxf(sj[2], sj[3]));
                    const float scale = __expf(mx[w] - mn);
                    float pj[4], ps = 0.0f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { pj[j] = __expf(sj[j] - mn); ps += pj[j]; }
                    den[w] = den[w] * scale + ps;
                    mx[w] = mn;
                    const f32x2 sc2 = pack2(scale, scale);
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[w][e] = mul2(acc[w][e], sc2);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const f32x2 p2 = pack2(pj[j], pj[j]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[w][e] = fma2(p2, cur[j].p[e], acc[w][e]);
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
        vb_cur = vb_nxt;
    }
#pragma unroll
    for (int w = 0; w < WT; ++w)
        if (w < W) {
            const float inv = (den[w] > 0.0f) ? 1.0f / den[w] : __int_as_float(0x7fc00000);   // all masked -> NaN like tfa
            float *o = xa + ((size_t)b * W + w) * (3 * UNITS) + UNITS;
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) unpack2(acc[w][e], v[2 * e], v[2 * e + 1]);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= inv;
            if (xp_hi != nullptr) {      // [h | ctx] as fp16 hi / lo planes: A operand of the attention-layer GEMM (the fp32 copy is not read then)
                const size_t po = ((size_t)b * W + w) * (3 * UNITS) + UNITS + 4 * lane;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const float *q = v + 4 * half;
                    const __half2 h0 = __floats2half2_rn(q[0], q[1]), h1 = __floats2half2_rn(q[2], q[3]);
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    const __half2 l0 = __floats2half2_rn(q[0] - f0.x, q[1] - f0.y), l1 = __floats2half2_rn(q[2] - f1.x, q[3] - f1.y);
                    *reinterpret_cast<uint2 *>(xp_hi + po + UNITS * half) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                    *reinterpret_cast<uint2 *>(xp_lo + po + UNITS * half) = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
                }
            } else {
                *reinterpret_cast<float4 *>(o + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(o + UNITS + 4 * lane) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
}

// (value desc, index asc) warp arg-max; dead candidates carry idx = INT_MAX, val = -inf.
__device__ __forceinline__ void warp_argmax(float &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// tf.math.top_k over the (at most 64) candidates of a snippet, two per lane: lane k receives the k-th best (value desc,
// index asc), k < W.  Used by fc_search and by the standalone beam step the bit-exactness test calls.
__device__ __forceinline__ void warp_topk(float v0, int i0, float v1, int i1, int W, int lane, float &sel_v, int &sel_i) {
    for (int k = 0; k < W; ++k) {
        float v; int i;
        if (v0 > v1 || (v0 == v1 && i0 < i1)) { v = v0; i = i0; } else { v = v1; i = i1; }
        warp_argmax(v, i);
        if (lane == k) { sel_v = v; sel_i = i; }
        if (i0 == i) { v0 = -INFINITY; i0 = 0x7fffffff; }
        if (i1 == i) { v1 = -INFINITY; i1 = 0x7fffffff; }
    }
}

// ---- logits = A . fc + b ; tfa _beam_search_step ; one warp per snippet ------------------------------------------
// Sums over the 32 lanes of N <= 32 per-lane values at once (recursive halving: 31 shuffles instead of 5 N): afterwards
// lane i holds the total of value i (lanes >= N hold padding).  v is used as scratch.
template <int N>
__device__ __forceinline__ float warp_sum_many(float (&v)[32], int lane) {
#pragma unroll
    for (int i = N; i < 32; ++i) v[i] = 0.0f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            if (i < N) {                                 // pairs entirely inside the padding stay zero
                const float keep = upper ? v[i + off] : v[i], send = upper ? v[i] : v[i + off];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return v[0];
}

// WT: beam bucket (1, 5 or 9), the loops over beams are unrolled to it
template <int WT>
__global__ void __launch_bounds__(128) fc_search_kernel(const float *__restrict__ att, const float *__restrict__ wfc, const float *__restrict__ bfc,
                                                        float *lp, int32_t *fin, int32_t *len, int32_t *tok, int32_t *parent,
                                                        int32_t *first_done, float *scores, int32_t *step_ids, int32_t *parent_ids,
                                                        int B, int W, int S, int t, const float *__restrict__ xa, float *__restrict__ X,
                                                        uint16_t *__restrict__ x_hi, uint16_t *__restrict__ x_lo,
                                                        const float *__restrict__ h0, uint16_t *__restrict__ x1_hi, uint16_t *__restrict__ x1_lo,
                                                        int32_t *__restrict__ skip, float *__restrict__ greedy_logits, int32_t *__restrict__ greedy_ids) {
    __shared__ float lg_s[4][WMAX * 8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x * 4 + wid;
    const bool active = b < B;
    const bool skipped = active && skip[b] != 0;
    if (!active) return;
    if (skipped) {
        // All W beams finished with finite scores: every candidate other than (beam k, end token) costs dtype.min, so tfa's
        // step selects exactly (k, end) for k = 0..W-1 in slot order (scores are already sorted): end token, identity
        // parents, unchanged scores, lengths and flags.  The state rows of this snippet are never read again.
        if (lane < W) {
            const size_t o = ((size_t)b * S + t) * W + lane;
            scores[o] = lp[(size_t)b * W + lane]; step_ids[o] = TOKEN_END; parent_ids[o] = lane;
        }
        return;
    }
    {
        // output layer: lane l owns features 4 l .. 4 l + 3 of every beam's attention vector (one coalesced float4 per row) and
        // the matching 4 x 7 block of the kernel; the W x 7 partial dot products are then summed over the lanes together
        float wr[4][VOCAB];
        {
            float wflat[4 * VOCAB];
#pragma unroll
            for (int j = 0; j < VOCAB; ++j) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(wfc + 4 * VOCAB * lane) + j);
                wflat[4 * j] = q.x; wflat[4 * j + 1] = q.y; wflat[4 * j + 2] = q.z; wflat[4 * j + 3] = q.w;
            }
#pragma unroll
            for (int d = 0; d < 4; ++d)
#pragma unroll
                for (int v = 0; v < VOCAB; ++v) wr[d][v] = wflat[d * VOCAB + v];
        }
        constexpr int NV = WT * VOCAB;                   // 7, 35 or 63 values
        float pa[32], pb[32];
#pragma unroll
        for (int k = 0; k < WT; ++k) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < W) a = __ldg(reinterpret_cast<const float4 *>(att + ((size_t)b * W + k) * UNITS) + lane);
#pragma unroll
            for (int v = 0; v < VOCAB; ++v) {
                const float t4 = fmaf(a.w, wr[3][v], fmaf(a.z, wr[2][v], fmaf(a.y, wr[1][v], a.x * wr[0][v])));
                const int i = k * VOCAB + v;
                if (i < 32) pa[i] = t4; else pb[i - 32] = t4;
            }
        }
        const float sa = warp_sum_many<(NV < 32 ? NV : 32)>(pa, lane);
        if (lane < NV && lane < W * VOCAB) lg_s[wid][(lane / VOCAB) * 8 + lane % VOCAB] = sa + __ldg(bfc + lane % VOCAB);
        if (NV > 32) {
            const float sb = warp_sum_many<(NV > 32 ? NV - 32 : 1)>(pb, lane);
            const int i = lane + 32;
            if (i < NV && i < W * VOCAB) lg_s[wid][(i / VOCAB) * 8 + i % VOCAB] = sb + __ldg(bfc + i % VOCAB);
        }
    }
    __syncwarp();
    const int n_cand = W * VOCAB;
    const size_t r0 = (size_t)b * W;
    if (greedy_logits != nullptr) {
        // BasicDecoder + GreedyEmbeddingSampler, W == 1: sample = argmax (ties -> lowest index), no masking of
        // finished rows (impute_finished = False: they keep decoding); T = first step at which every row has emitted the end token
        float v = (lane < VOCAB) ? lg_s[wid][lane] : -INFINITY;
        int i = (lane < VOCAB) ? lane : 0x7fffffff;
        if (lane < VOCAB) greedy_logits[((size_t)b * S + t) * VOCAB + lane] = v;
        warp_argmax(v, i);
        if (lane == 0) {
            greedy_ids[(size_t)b * S + t] = i;
            tok[r0] = i; parent[r0] = 0;
            if (i == TOKEN_END && first_done[b] == S) first_done[b] = t;
        }
    }
    int nfin = 1, nlen = 0, word = 0, par = 0;
    float sel_v = 0.0f;
    if (greedy_logits == nullptr) {
    float v0 = -INFINITY, v1 = -INFINITY; int i0 = 0x7fffffff, i1 = 0x7fffffff;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < n_cand) {
            const int k = i / VOCAB, v = i % VOCAB;
            float slp;
            if (fin[r0 + k]) slp = (v == TOKEN_END) ? 0.0f : F32_MIN;
            else {
                const float *lg = lg_s[wid] + k * 8;
                float m = lg[0];
#pragma unroll
                for (int q = 1; q < VOCAB; ++q) m = fmaxf(m, lg[q]);
                float se = 0.0f;
#pragma unroll
                for (int q = 0; q < VOCAB; ++q) se += expf(lg[q] - m);
                slp = (lg[v] - m) - logf(se);
            }
            const float tot = lp[r0 + k] + slp;
            if (h == 0) { v0 = tot; i0 = i; } else { v1 = tot; i1 = i; }
        }
    }
    int sel_i = 0;
    warp_topk(v0, i0, v1, i1, W, lane, sel_v, sel_i);
    if (lane < W) {
        word = sel_i % VOCAB; par = sel_i / VOCAB;
        const int pf = fin[r0 + par];
        nfin = pf | (word == TOKEN_END);
        nlen = len[r0 + par] + (pf ? 0 : 1);
    }
    __syncwarp();
    if (lane < W) {
        fin[r0 + lane] = nfin; len[r0 + lane] = nlen; lp[r0 + lane] = sel_v; tok[r0 + lane] = word; parent[r0 + lane] = par;
        const size_t o = ((size_t)b * S + t) * W + lane;
        scores[o] = sel_v; step_ids[o] = word; parent_ids[o] = par;
    }
    const unsigned allfin = __ballot_sync(0xffffffffu, lane >= W || nfin);
    if (lane == 0 && allfin == 0xffffffffu && first_done[b] == S) first_done[b] = t;
    // from the next step on this snippet is skipped -- only when the continuation is exactly known (finite scores; a
    // -inf slot, possible for widths > 7, would let dtype.min candidates of other beams in)
    const unsigned allfinite = __ballot_sync(0xffffffffu, lane >= W || sel_v > F32_MIN);
    if (lane == 0 && allfin == 0xffffffffu && allfinite == 0xffffffffu) skip[b] = 1;
    }   // beam search
    // input of the next step's cell GEMM, gathered through the parents chosen just now: X[r] = [attention[src] | h[src]]
    // (two stacked cells: h of cell 0 comes from h0, and the top cell's h goes into the second half of X1[r] = [h0_new[r] | h1[src]])
    for (int k = 0; k < W; ++k) {
        const size_t src = r0 + __shfl_sync(0xffffffffu, par, k);
        const float4 va = __ldg(reinterpret_cast<const float4 *>(att + src * UNITS) + lane);
        const float4 vtop = __ldg(reinterpret_cast<const float4 *>(xa + src * (3 * UNITS)) + lane);
        const float4 vh = (h0 != nullptr) ? __ldg(reinterpret_cast<const float4 *>(h0 + src * UNITS) + lane) : vtop;
        if (x_hi != nullptr) {                 // fp16 hi / lo planes (the cell GEMM runs on the fp16 pipe)
            auto put = [&](uint16_t *hi, uint16_t *lo, const float4 v, size_t o) {
                const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
                *reinterpret_cast<uint2 *>(hi + o) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                *reinterpret_cast<uint2 *>(lo + o) = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
            };
            put(x_hi, x_lo, va, (r0 + k) * (2 * UNITS) + 4 * lane);
            put(x_hi, x_lo, vh, (r0 + k) * (2 * UNITS) + UNITS + 4 * lane);
            if (x1_hi != nullptr) put(x1_hi, x1_lo, vtop, (r0 + k) * (2 * UNITS) + UNITS + 4 * lane);
        } else {
            float4 *dst = reinterpret_cast<float4 *>(X + (r0 + k) * (2 * UNITS));
            dst[lane] = va;
            dst[32 + lane] = vh;
        }
    }
}

__global__ void init_state_kernel(float *lp, int32_t *fin, int32_t *len, int32_t *tok, int32_t *parent, int32_t *first_done,
                                  int32_t *skip, long long rows, int W, int S) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int k = (int)(r % W);
    lp[r] = (k == 0) ? 0.0f : -INFINITY;
    fin[r] = (k == 0) ? 0 : 1;
    len[r] = 0; tok[r] = TOKEN_START; parent[r] = k;
    if (k == 0) { first_done[r / W] = S; skip[r / W] = 0; }
}

// tfa.seq2seq.gather_tree for beam slot k of one snippet; element (level, slot) of the snippet is at [level * ts + slot].
__device__ __forceinline__ void gather_tree_slot(const int32_t *step_ids, const int32_t *parent_ids, int32_t *out, size_t ts,
                                                 int L, int T, int k) {
    for (int tt = L; tt < T; ++tt) out[(size_t)tt * ts + k] = TOKEN_END;
    int par = k;
    for (int level = L - 1; level >= 0; --level) {
        out[(size_t)level * ts + k] = step_ids[(size_t)level * ts + par];
        par = parent_ids[(size_t)level * ts + par];
    }
    bool done = false;
    for (int tt = 0; tt < L; ++tt) {
        int32_t *o = out + (size_t)tt * ts + k;
        if (done) *o = TOKEN_END;
        else if (*o == TOKEN_END) done = true;
    }
}

// K5 standalone (parity tests): one tfa _beam_search_step on log-softmaxed rows, through the same top-k as fc_search
__global__ void beam_step_kernel(const float *slp, const float *lp, const uint8_t *fin, const long long *len, long long B, int W, int V,
                                 int end_token, float *scores, int32_t *word, int32_t *parent, uint8_t *nfin, long long *nlen) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n_cand = W * V;
    float v[2] = {-INFINITY, -INFINITY}; int ix[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < n_cand) {
            const int k = i / V, t = i % V;
            const float sl = fin[b * W + k] ? ((t == end_token) ? 0.0f : F32_MIN) : slp[(b * W + k) * V + t];
            v[h] = lp[b * W + k] + sl; ix[h] = i;
        }
    }
    float sel_v = 0.0f; int sel_i = 0;
    warp_topk(v[0], ix[0], v[1], ix[1], W, lane, sel_v, sel_i);
    if (lane < W) {
        const int wd = sel_i % V, pr = sel_i / V;
        const bool pf = fin[b * W + pr] != 0;
        scores[b * W + lane] = sel_v; word[b * W + lane] = wd; parent[b * W + lane] = pr;
        nfin[b * W + lane] = (pf || wd == end_token) ? 1 : 0;
        nlen[b * W + lane] = len[b * W + pr] + (pf ? 0 : 1);
    }
}

// K5 standalone: gather_tree on time-major [T,B,W] arrays as tfa.seq2seq.gather_tree takes them
__global__ void gather_tree_kernel(const int32_t *step_ids, const int32_t *parent_ids, const int32_t *max_len, int T, long long B, int W,
                                   int32_t *out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * W) return;
    const long long b = g / W;
    const int L = max(0, min(T, max_len[b]));
    gather_tree_slot(step_ids + b * W, parent_ids + b * W, out + b * W, (size_t)B * W, L, T, (int)(g % W));
}

// gather_tree on [B,S,W] arrays + T = max over snippets of (first all-finished step + 1)
__global__ void finalize_kernel(const int32_t *step_ids, const int32_t *parent_ids, const int32_t *len, const int32_t *first_done,
                                int32_t *ids, int32_t *steps, int B, int W, int S, int greedy) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * W) return;
    const int b = g / W, k = g % W;
    if (greedy) { atomicMax(steps, min(first_done[b] + 1, S)); return; }      // sample ids were written step by step
    int maxlen = 0;
    for (int q = 0; q < W; ++q) maxlen = max(maxlen, len[b * W + q]);
    const int L = min(S, maxlen);
    const size_t o = (size_t)b * S * W;
    gather_tree_slot(step_ids + o, parent_ids + o, ids + o, (size_t)W, L, S, k);
    if (k == 0) atomicMax(steps, min(first_done[b] + 1, S));
}

size_t workspace_floats(long long rows, int depth) {
    // X 256 | Z 512 | XA 384 | Q 256 | ATT 128 | c x2 256 | lp 1  + ints: fin len tok parent 4 + first_done + skip
    // depth 2 adds: X1 planes 256 | H0 128 | c of cell 1 x2 256
    return (size_t)rows * (256 + 512 + 384 + 256 + 128 + 256 + 1 + 4 + 2 + (depth == 2 ? 256 + 128 + 256 : 0)) + 64 + 64;
}

int run(const Params &p, cudaStream_t s) {
    if (p.B <= 0 || p.S <= 0) return RVB_OK;
    if (p.W < 1 || p.W > WMAX) return fail(RVB_ERR_ARG, "decoder_wave: beam width must be in [1,%d]", WMAX);
    if (p.greedy && (p.W != 1 || p.logits == nullptr)) return fail(RVB_ERR_ARG, "decoder_wave: greedy search needs width 1 and a logits buffer");
    const long long rows = (long long)p.B * p.W;
    float *X = p.ws, *Z = X + rows * 256, *XA = Z + rows * 512, *Q = XA + rows * 384, *ATT = Q + rows * 256;
    float *c0 = ATT + rows * 128, *c1 = c0 + rows * 128, *lp = c1 + rows * 128;
    int32_t *fin = reinterpret_cast<int32_t *>(lp + rows), *len = fin + rows, *tok = len + rows, *parent = tok + rows;
    int32_t *first_done = parent + rows;
    int32_t *skip = first_done + rows;        // per snippet: all beams finished, continuation known (see fc_search)
    const bool two = p.depth == 2;
    // depth 2 only; re-aligned to 256 bytes (the scalar arrays before it leave any 4-byte offset): TMA source + vector stores
    float *X1 = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(skip + rows) + 255) & ~uintptr_t(255));
    float *H0 = X1 + rows * 256, *d0 = H0 + rows * 128, *d1 = d0 + rows * 128;
    uint16_t *x1_hi = reinterpret_cast<uint16_t *>(X1), *x1_lo = x1_hi + rows * 256;
    // fp16-plane operands share the X / Z regions: X = [hi plane | lo plane] of [rows][256] halves, Z holds the h planes
    static const bool f16_off = getenv("RVB_DECODER_GEMM") && strcmp(getenv("RVB_DECODER_GEMM"), "tf32") == 0;
    const bool f16 = p.wg16_hi != nullptr && p.wm16_hi != nullptr && !f16_off;
    if (two && (!f16 || p.wg1_16_hi == nullptr || p.b1 == nullptr))
        return fail(RVB_ERR_STATE, "decoder_wave: decoder_depth 2 needs the fp16-plane weights");
    uint16_t *x_hi = reinterpret_cast<uint16_t *>(X), *x_lo = x_hi + rows * 256;
    // [h | ctx] as fp16 hi / lo planes, [rows][384] halves each, in the Z region (the fused cell epilogue never writes Z itself):
    // h from the cell epilogue (A operand of the query GEMM, K = 128, row pitch 384), ctx from the attention kernel, and the
    // whole row is the A operand of the attention-layer GEMM.  Without wa16 planes only the h part is used.
    static const bool xap_off = getenv("RVB_ATT_LAYER") && strcmp(getenv("RVB_ATT_LAYER"), "tf32") == 0;      // A/B switch
    const bool xap = f16 && p.wa16_hi != nullptr && !xap_off;
    const int h_ld = xap ? 3 * UNITS : UNITS;
    uint16_t *h_hi = reinterpret_cast<uint16_t *>(Z), *h_lo = h_hi + rows * h_ld;
    uint16_t *xp_hi = xap ? h_hi : nullptr, *xp_lo = xap ? h_lo : nullptr;
    RVB_CUDA(cudaMemsetAsync(X, 0, sizeof(float) * rows * 256, s));          // step 0: attention = h = 0
    RVB_CUDA(cudaMemsetAsync(XA, 0, sizeof(float) * rows * 384, s));
    RVB_CUDA(cudaMemsetAsync(ATT, 0, sizeof(float) * rows * 128, s));
    RVB_CUDA(cudaMemsetAsync(c0, 0, sizeof(float) * rows * 128, s));
    if (two) {
        RVB_CUDA(cudaMemsetAsync(X1, 0, sizeof(float) * rows * 256, s));         // h of both cells = 0
        RVB_CUDA(cudaMemsetAsync(H0, 0, sizeof(float) * rows * 128, s));
        RVB_CUDA(cudaMemsetAsync(d0, 0, sizeof(float) * rows * 128, s));
    }
    // profiling (bench.py): the attention kernel is timed per launch as its own kind, everything else as "decoder"
    {
        ProfScope ps(KK_DECODER, s);
        init_state_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(lp, fin, len, tok, parent, first_done, skip, rows, p.W, p.S);
        RVB_LAUNCH_CHECK();
    }
    int nl = 1;
    for (int t = 0; t < p.S; ++t) {
        float *cin = (t & 1) ? c1 : c0, *cout = (t & 1) ? c0 : c1;
        const unsigned ab = (unsigned)((p.B + 3) / 4);
        {
            ProfScope ps(KK_DECODER, s);
            // cell update fused into the GEMM epilogue; with fp16 weight planes both GEMMs run on the fp16 pipe (3 split passes)
            const gemm::CellEpilogue ce{p.wtok, tok, parent, cin, cout, XA, p.W, f16 ? h_hi : nullptr, f16 ? h_lo : nullptr, 0, h_ld, p.gru};
            if (two) {
                // cell 0: h0 (fp32, for the next step's gather) -> H0, and as fp16 planes into the first half of X1;
                // cell 1: X1 = [h0 | h1_prev[src]] . [W1 ; U1] + b1 with its own c ping-pong; its h is the query / attention-layer input
                float *din = (t & 1) ? d1 : d0, *dout = (t & 1) ? d0 : d1;
                const gemm::CellEpilogue ce0{p.wtok, tok, parent, cin, cout, H0, p.W, x1_hi, x1_lo, UNITS, 2 * UNITS, p.gru};
                const gemm::CellEpilogue ce1{p.b1, nullptr, parent, din, dout, XA, p.W, h_hi, h_lo, 0, h_ld, p.gru};
                RVB_CHECK(gemm::run_tc_f16(x_hi, x_lo, p.wg16_hi, p.wg16_lo, nullptr, Z, rows, GATES, 2 * UNITS, RVB_PREC_FP32, p.abort_flag, s, false, &ce0));
                RVB_CHECK(gemm::run_tc_f16(x1_hi, x1_lo, p.wg1_16_hi, p.wg1_16_lo, nullptr, Z, rows, GATES, 2 * UNITS, RVB_PREC_FP32, p.abort_flag, s, false, &ce1));
                RVB_CHECK(gemm::run_tc_f16(h_hi, h_lo, p.wm16_hi, p.wm16_lo, nullptr, Q, rows, ENC_OUT, UNITS, RVB_PREC_FP32, p.abort_flag, s, false, nullptr, h_ld));
                ++nl;
            } else if (f16) {
                RVB_CHECK(gemm::run_tc_f16(x_hi, x_lo, p.wg16_hi, p.wg16_lo, nullptr, Z, rows, GATES, 2 * UNITS, RVB_PREC_FP32, p.abort_flag, s, false, &ce));
                RVB_CHECK(gemm::run_tc_f16(h_hi, h_lo, p.wm16_hi, p.wm16_lo, nullptr, Q, rows, ENC_OUT, UNITS, RVB_PREC_FP32, p.abort_flag, s, false, nullptr, h_ld));
            } else {
                RVB_CHECK(gemm::run_tc(X, p.wg_hiT, p.wg_loT, nullptr, Z, rows, GATES, 2 * UNITS, RVB_PREC_FP32, p.abort_flag, s, 0, &ce));
                RVB_CHECK(gemm::run_tc(XA, p.wm_hiT, p.wm_loT, nullptr, Q, rows, ENC_OUT, UNITS, RVB_PREC_FP32, p.abort_flag, s, 3 * UNITS));
            }
        }
        {
            ProfScope ps(KK_ATTENTION, s);
            if (p.v_hi != nullptr && p.W >= 2) {   // both contractions on tcgen05 (attention_tc.cu)
                RVB_CHECK(atc::run(p.v_hi, p.v_lo, p.mask, Q, XA, skip, p.B, p.Tm, p.W, p.abort_flag, s, xp_hi, xp_lo));
            } else if (p.values16 != nullptr && p.att16_tc) {   // reduced-precision mode: one fp16 plane, same tcgen05 kernel
                RVB_CHECK(atc::run(p.values16, nullptr, p.mask, Q, XA, skip, p.B, p.Tm, p.W, p.abort_flag, s, xp_hi, xp_lo));
            } else if (p.values16 != nullptr) {    // reduced-precision mode: fp16 copy of the memory
                const __half *v16 = reinterpret_cast<const __half *>(p.values16);
                if (p.W == 1) attention_kernel<1, __half><<<ab, 128, 0, s>>>(v16, p.mask, Q, XA, p.B, p.Tm, p.W, skip, xp_hi, xp_lo);
                else if (p.W <= 5) attention_kernel<5, __half><<<ab, 128, 0, s>>>(v16, p.mask, Q, XA, p.B, p.Tm, p.W, skip, xp_hi, xp_lo);
                else attention_kernel<9, __half><<<ab, 128, 0, s>>>(v16, p.mask, Q, XA, p.B, p.Tm, p.W, skip, xp_hi, xp_lo);
            } else if (p.W == 1) attention_kernel<1, float><<<ab, 128, 0, s>>>(p.values, p.mask, Q, XA, p.B, p.Tm, p.W, skip, xp_hi, xp_lo);
            else if (p.W <= 5) attention_kernel<5, float><<<ab, 128, 0, s>>>(p.values, p.mask, Q, XA, p.B, p.Tm, p.W, skip, xp_hi, xp_lo);
            else attention_kernel<9, float><<<ab, 128, 0, s>>>(p.values, p.mask, Q, XA, p.B, p.Tm, p.W, skip, xp_hi, xp_lo);
        }
        {
            ProfScope ps(KK_DECODER, s);
            if (xap)        // [h | ctx] planes x the attention layer padded to one 256-column tile; 128 columns come out
                RVB_CHECK(gemm::run_tc_f16(xp_hi, xp_lo, p.wa16_hi, p.wa16_lo, nullptr, ATT, rows, 2 * UNITS, 3 * UNITS, RVB_PREC_FP32, p.abort_flag, s,
                                           false, nullptr, 0, UNITS));
            else RVB_CHECK(gemm::run_tc(XA, p.wa_hiT, p.wa_loT, nullptr, ATT, rows, UNITS, 3 * UNITS, RVB_PREC_FP32, p.abort_flag, s));
            auto fcs = (p.W == 1) ? fc_search_kernel<1> : (p.W <= 5) ? fc_search_kernel<5> : fc_search_kernel<9>;
            fcs<<<ab, 128, 0, s>>>(ATT, p.wfc, p.bfc, lp, fin, len, tok, parent, first_done, p.scores, p.step_ids,
                                   p.parent_ids, p.B, p.W, p.S, t, XA, X, f16 ? x_hi : nullptr, f16 ? x_lo : nullptr,
                                   two ? H0 : nullptr, two ? x1_hi : nullptr, two ? x1_lo : nullptr, skip,
                                   p.greedy ? p.logits : nullptr, p.greedy ? p.ids : nullptr);
            RVB_LAUNCH_CHECK();
        }
        nl += 2;
    }
    {
        ProfScope ps(KK_DECODER, s);
        finalize_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, s>>>(p.step_ids, p.parent_ids, len, first_done, p.ids, p.steps, p.B, p.W, p.S, p.greedy);
        RVB_LAUNCH_CHECK();
    }
    count_launch(nl + 1);
    return RVB_OK;
}

}  // namespace decw
}  // namespace rvb

using namespace rvb;

extern "C" int rvb_beam_step(const float *d_slp, const float *d_lp, const uint8_t *d_fin, const int64_t *d_len,
                             int64_t batch, int W, int V, int end_token, float *d_scores, int32_t *d_word,
                             int32_t *d_parent, uint8_t *d_nfin, int64_t *d_nlen, void *stream) {
    if (batch < 0 || W < 1 || W > decw::WMAX || V < 1 || W * V > 64) return fail(RVB_ERR_ARG, "beam_step: need 1 <= W*V <= 64");
    if (batch == 0) return RVB_OK;
    decw::beam_step_kernel<<<(unsigned)((batch + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        d_slp, d_lp, d_fin, reinterpret_cast<const long long *>(d_len), batch, W, V, end_token, d_scores, d_word,
        d_parent, d_nfin, reinterpret_cast<long long *>(d_nlen));
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

extern "C" int rvb_gather_tree(const int32_t *d_step_ids, const int32_t *d_parent_ids, const int32_t *d_max_len,
                               int steps, int64_t batch, int W, int end_token, int32_t *d_out, void *stream) {
    if (batch < 0 || W < 1 || steps < 0) return fail(RVB_ERR_ARG, "gather_tree: bad shape");
    if (end_token != TOKEN_END) return fail(RVB_ERR_ARG, "gather_tree: end_token must be %d", TOKEN_END);
    if (batch == 0 || steps == 0) return RVB_OK;
    const long long n = batch * W;
    decw::gather_tree_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_step_ids, d_parent_ids, d_max_len, steps, batch, W, d_out);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}
