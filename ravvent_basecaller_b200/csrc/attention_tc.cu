// K4 attention on the 5th-generation tensor cores: beam widths >= 2 in parity mode (decoder_wave.cu keeps the FFMA kernel for
// width 1, which runs at the HBM rate there), every width in reduced-precision mode (one fp16 plane, PLANES = 1 below).
//
// Replaces, per decode step, tfa LuongAttention + the context reduction of AttentionWrapper (reference basecaller.py:117-134,
// SURVEY A.3 / A.3b) for all beams of a snippet at once:
//     S[t, w]  = V[t, :] . q'[w, :]                 (q' = W_mem h: the memory layer is folded into the query)
//     P[t, w]  = softmax over the unmasked rows t
//     ctx[w,:] = sum_t P[t, w] V[t, :]
// The FFMA form of this (one warp per snippet, columns split over the lanes) needs a 5-stage shuffle butterfly per
// (row, beam) score and is instruction bound at beam 5 (47 % of the HBM rate).  Here both contractions run on tcgen05:
//   * the memory V is kept by K3 as fp16 hi + lo planes ([B, Tm, 256] each, the same bytes as fp32) and TMA-loaded in
//     128-row tiles, cut into "units" of 64 KB = 128 rows x 128 columns x (hi, lo) -- a ring of three units;
//   * scores: D_s[128 rows, .] = V_unit (A operand, K-major: K = columns) . Q^T (B operand: beams padded to 16 rows).  The
//     three split products V_hi.Q_hi, V_hi.Q_lo, V_lo.Q_hi that give fp32-level accuracy on the fp16 pipe take TWO reads of
//     the A operand, not three: Q_hi and Q_lo are stacked into one 32-row B operand (columns 0-15 | 16-31 of the
//     accumulator) and V_lo meets the first 16 rows of the same tile (columns 32-47); the softmax warps add the three;
//   * context: D_c[128 columns, .] = V_unit^T (the SAME shared-memory bytes read as an MN-major A operand: M = columns,
//     K = rows) . P (B operand written by the softmax warps as fp16 hi / lo, stacked the same way);
//   * an M = 128, K = 16 MMA costs ~40 cycles for any N <= 32 (tools/mma_small_n_bench.cu: it is bound by the 4 KB
//     shared-memory read of its A operand), which is why the number of A reads is what matters;
//   * four softmax warps own the 128 TMEM lanes: online softmax over the (at most two) row tiles.  Each tile has its own
//     context accumulator, combined with exp(m_tile - m_final) when the snippet is written out, so no accumulator is
//     rescaled in TMEM and tile 1's probabilities never wait for tile 0's context MMAs.
// One persistent CTA per SM walks its snippets; warp roles: w0 TMA producer, w1 TMEM allocation + MMA issue, w2..w5 softmax.
// Every mbarrier wait is bounded (abort flag), as in the projection kernel.
#include "proj_gemm_tc.cuh"

namespace rvb {
namespace atc {

using gemm::tc::make_desc;
using gemm::tc::make_idesc_f16;
using gemm::tc::mbar_arrive;
using gemm::tc::mbar_expect_tx;
using gemm::tc::mbar_init;
using gemm::tc::mbar_wait;
using gemm::tc::smem_u32;
using gemm::tc::umma_commit;
using gemm::tc::umma_f16;

constexpr int ROWS = 128;                       // memory rows per tile = TMEM lanes
constexpr int NB = 16;                          // beams padded to the smallest MMA N at M = 128
constexpr int BOX_BYTES = ROWS * 128;           // [128 rows][64 fp16], 128-byte swizzle
constexpr int RING_BYTES = 12 * BOX_BYTES;      // 192 KB: three units of both planes, or six of one
constexpr int MAX_RING = 6;
constexpr int QBOX = 2 * NB * 128;              // [hi beams 0-15 | lo beams 0-15][64 fp16]: one K-box of a stacked B operand
constexpr int LO_ROWS = NB * 128;               // byte offset of the lo rows inside a box (2 swizzle atoms of 8 rows)
constexpr int Q_BYTES = 4 * QBOX;               // 4 K-boxes of 64 columns
constexpr int P_TILE = 2 * QBOX;                // 2 K-boxes of 64 rows
constexpr int P_BYTES = 2 * P_TILE;             // one per row tile: tile 1 is written while tile 0's context MMAs read theirs
constexpr int THREADS = 192;
constexpr size_t SMEM = 1024 + (size_t)RING_BYTES + Q_BYTES + P_BYTES + 1024;
constexpr int ACC = 3 * NB;                     // accumulator columns per product group: hi.hi | hi.lo | lo.hi
constexpr int TMEM_COLS = 512;                  // scores 2 tiles x 48 | context (2 tiles x 2 column halves) x 48  = 288 -> 512
constexpr int WMAX_ = 9;                        // widest beam (decoder_wave.cu WMAX)

__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// MN-major, SWIZZLE_128B matrix descriptor: 64 M-elements (128 bytes) are contiguous, the next 64 are LBO bytes further,
// K advances by 128 bytes per element inside an 8-row swizzle atom and by SBO bytes per atom.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// tcgen05.mma with the descriptors given as (low word, high word): the issuer keeps base words and adds immediates
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        :: "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ __forceinline__ void split_f16(float v, uint16_t &hi, uint16_t &lo) {
    const __half h = __float2half_rn(v);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
}
// byte offset of element (row n, k) inside a K-major SWIZZLE_128B box of [rows][64 fp16]
__device__ __forceinline__ uint32_t kmajor_off(int n, int k) {
    return (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2);
}

// WT: beam bucket, loops over beams are unrolled to WT (1, 5 or 9).
// PLANES: 2 = fp16 hi + lo planes of the memory (fp32-level products), 1 = one fp16 plane (the reduced-precision mode's memory:
// half the bytes, and a whole snippet -- four 32 KB units -- fits the ring of six with room to run into the next one).
template <int WT, int PLANES>
__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const uint8_t *__restrict__ mask, const float *__restrict__ Q, float *__restrict__ xa,
                    const int32_t *__restrict__ skip, int B, int Tm, int W, int *abort_flag,
                    uint16_t *__restrict__ xp_hi, uint16_t *__restrict__ xp_lo) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char *ring = smem;
    constexpr int UNIT_BYTES = 2 * PLANES * BOX_BYTES;  // hi cols a | hi cols b [| lo cols a | lo cols b]: 128 columns of every plane
    constexpr int RING = RING_BYTES / UNIT_BYTES;
    unsigned char *qbuf = smem + (size_t)RING_BYTES;
    unsigned char *pbuf = qbuf + Q_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(pbuf + P_BYTES);
    uint64_t *full = bars, *empty = bars + MAX_RING;
    uint64_t *q_ready = bars + 2 * MAX_RING, *p_ready = q_ready + 1, *c_done = q_ready + 2, *s_ready = q_ready + 3;   // s_ready[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_ready + 2);
    float *red = reinterpret_cast<float *>(tmem_slot + 2);          // [2 buffers][4 warps][16]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (Tm + ROWS - 1) / ROWS;                     // 1 or 2 (Tm <= 256)

    if (threadIdx.x == 0) {
        for (int u = 0; u < RING; ++u) { mbar_init(&full[u], 1); mbar_init(&empty[u], 1); }
        mbar_init(q_ready, 128); mbar_init(p_ready, 128); mbar_init(c_done, 1);
        mbar_init(&s_ready[0], 1); mbar_init(&s_ready[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // beams W..15 of the B operands are zero for the whole kernel
    for (int i = threadIdx.x; i < (Q_BYTES + P_BYTES) / 16; i += THREADS) reinterpret_cast<uint4 *>(qbuf)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: units (tile, column half) in order =====================
        if (lane == 0) {
            uint32_t n = 0;                                         // units issued so far
            for (int b = blockIdx.x; b < B; b += gridDim.x) {
                if (skip[b]) continue;
                for (int t = 0; t < n_tiles; ++t)
                    for (int h = 0; h < 2; ++h, ++n) {
                        const uint32_t u = n % RING, round = n / RING;
                        if (round > 0 && !mbar_wait(&empty[u], (round - 1) & 1, abort_flag)) return;
                        unsigned char *dst = ring + (size_t)u * UNIT_BYTES;
                        mbar_expect_tx(&full[u], UNIT_BYTES);
                        tma_load_3d(&map_hi, &full[u], dst, 128 * h, ROWS * t, b);
                        tma_load_3d(&map_hi, &full[u], dst + BOX_BYTES, 128 * h + 64, ROWS * t, b);
                        if (PLANES == 2) {
                            tma_load_3d(&map_lo, &full[u], dst + 2 * BOX_BYTES, 128 * h, ROWS * t, b);
                            tma_load_3d(&map_lo, &full[u], dst + 3 * BOX_BYTES, 128 * h + 64, ROWS * t, b);
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s2 = make_idesc_f16(ROWS, 2 * NB), idesc_s1 = make_idesc_f16(ROWS, NB);   // A, B K-major
            constexpr uint32_t idesc_c2 = idesc_s2 | (1u << 15), idesc_c1 = idesc_s1 | (1u << 15);             // A MN-major (V^T)
            const uint32_t q0 = smem_u32(qbuf), p0 = smem_u32(pbuf), r0 = smem_u32(ring);
            uint32_t n = 0, nq = 0, np = 0;                           // units consumed, q_ready / p_ready phases seen
            // descriptor words: low = (address >> 4) | LBO field, high = SBO | version | swizzle; one thread issues ~100 tiny MMAs
            // per tile, so every instruction of descriptor arithmetic counts
            const uint64_t dk = make_desc(0), dm = make_desc_mn(0, BOX_BYTES, 1024);
            const uint32_t k_lo = (uint32_t)dk, k_hi = (uint32_t)(dk >> 32), m_lo = (uint32_t)dm, m_hi = (uint32_t)(dm >> 32);
            const uint32_t qw = k_lo + (q0 >> 4), pw = k_lo + (p0 >> 4);
            // scores of unit (t, h): 128 columns of K = 2 boxes x 4 k-steps; V_hi . [Q_hi | Q_lo] (N = 32), V_lo . Q_hi (N = 16)
            auto scores_unit = [&](uint32_t ubase, int t, int h) {
                const uint32_t aw = k_lo + (ubase >> 4), d = tmem_base + (uint32_t)(ACC * t);
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t bq = qw + (uint32_t)(((2 * h + jj) * QBOX + ks * 32) >> 4);
                        const uint32_t acc = (h | jj | ks) ? 1u : 0u;
                        umma_f16_w(d, aw + (uint32_t)((jj * BOX_BYTES + ks * 32) >> 4), k_hi, bq, k_hi, idesc_s2, acc);
                        if (PLANES == 2)
                            umma_f16_w(d + 2 * NB, aw + (uint32_t)(((2 + jj) * BOX_BYTES + ks * 32) >> 4), k_hi, bq, k_hi, idesc_s1, acc);
                    }
            };
            // context of unit (t, h): columns 128h..128h+127 (M), this tile's 128 rows (K = 8 k-steps of 16), into the
            // accumulator of (t, h): V_hi^T . [P_hi | P_lo], V_lo^T . P_hi
            auto ctx_unit = [&](uint32_t ubase, int t, int h) {
                const uint32_t aw = m_lo + (ubase >> 4), d = tmem_base + (uint32_t)(2 * ACC + ACC * (2 * t + h));
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t bp = pw + (uint32_t)((t * P_TILE + (ks >> 2) * QBOX + (ks & 3) * 32) >> 4);
                    umma_f16_w(d, aw + (uint32_t)((ks * 2048) >> 4), m_hi, bp, k_hi, idesc_c2, ks ? 1u : 0u);
                    if (PLANES == 2)
                        umma_f16_w(d + 2 * NB, aw + (uint32_t)((2 * BOX_BYTES + ks * 2048) >> 4), m_hi, bp, k_hi, idesc_c1, ks ? 1u : 0u);
                }
            };
            auto wait_unit = [&](uint32_t k) -> bool { return mbar_wait(&full[k % RING], (k / RING) & 1, abort_flag); };
            for (int b = blockIdx.x; b < B; b += gridDim.x) {
                if (skip[b]) continue;
                if (!mbar_wait(q_ready, nq++ & 1, abort_flag)) return;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t nb0 = n;                              // first unit of this snippet
                // scores of tile 0 (both column halves), then as much of tile 1 as the ring holds
                for (int h = 0; h < 2; ++h) {
                    if (!wait_unit(nb0 + h)) return;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    scores_unit(r0 + ((nb0 + h) % RING) * UNIT_BYTES, 0, h);
                }
                umma_commit(&s_ready[0]);
                if (n_tiles == 2) {
                    if (!wait_unit(nb0 + 2)) return;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    scores_unit(r0 + ((nb0 + 2) % RING) * UNIT_BYTES, 1, 0);
                    if (PLANES == 1) {                               // the whole snippet is in the ring
                        if (!wait_unit(nb0 + 3)) return;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        scores_unit(r0 + ((nb0 + 3) % RING) * UNIT_BYTES, 1, 1);
                        umma_commit(&s_ready[1]);
                    }
                }
                for (int t = 0; t < n_tiles; ++t) {
                    if (!mbar_wait(p_ready, np++ & 1, abort_flag)) return;       // P of tile t written
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t k = nb0 + 2 * t + h;
                        ctx_unit(r0 + (k % RING) * UNIT_BYTES, t, h);
                        umma_commit(&empty[k % RING]);                           // unit free when these MMAs retire
                    }
                    if (t == n_tiles - 1) umma_commit(c_done);                   // every context MMA of the snippet
                    if (PLANES == 2 && t == 0 && n_tiles == 2) {
                        // second column half of tile 1 lands in a unit freed just now
                        if (!wait_unit(nb0 + 3)) return;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        scores_unit(r0 + ((nb0 + 3) % RING) * UNIT_BYTES, 1, 1);
                        umma_commit(&s_ready[1]);
                    }
                }
                n += 2 * n_tiles;
            }
        }
    } else {
        // ===================== softmax warps: thread e <-> TMEM lane e =====================
        const int qd = warp & 3;
        const int e = 32 * qd + lane;                                // row inside a tile (scores) / column inside a half (context)
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * qd) << 16);
        const uint32_t q0 = smem_u32(qbuf), p0 = smem_u32(pbuf);
        uint32_t ns0 = 0, ns1 = 0, nc = 0, nred = 0;
        auto reduce16 = [&](float (&v)[WT], bool is_max) {        // all-reduce over the 128 softmax threads, per beam
#pragma unroll
            for (int w = 0; w < WT; ++w)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float other = __shfl_xor_sync(0xffffffffu, v[w], o);
                    v[w] = is_max ? fmaxf(v[w], other) : v[w] + other;
                }
            float *buf = red + (nred++ & 1) * 64;
            if (lane == 0) {
#pragma unroll
                for (int w = 0; w < WT; ++w) buf[(warp - 2) * 16 + w] = v[w];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int w = 0; w < WT; ++w) {
                const float a = buf[w], b2 = buf[16 + w], c2 = buf[32 + w], d2 = buf[48 + w];
                v[w] = is_max ? fmaxf(fmaxf(a, b2), fmaxf(c2, d2)) : (a + b2) + (c2 + d2);
            }
        };
        // queries of a snippet -> fp16 hi / lo B tiles
        auto write_q = [&](int b) {
            float qv[WT][2];
#pragma unroll
            for (int w = 0; w < WT; ++w)            // all loads first: one round trip to L2 instead of 2 W dependent ones
#pragma unroll
                for (int half = 0; half < 2; ++half)
                    qv[w][half] = (w < W) ? __ldg(Q + ((size_t)b * W + w) * ENC_OUT + e + 128 * half) : 0.0f;
#pragma unroll
            for (int w = 0; w < WT; ++w)
                if (w < W) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int k = e + 128 * half;
                        uint16_t hi, lo;
                        split_f16(qv[w][half], hi, lo);
                        const uint32_t off = (uint32_t)(k >> 6) * QBOX + kmajor_off(w, k & 63);
                        sts16(q0 + off, hi);
                        sts16(q0 + LO_ROWS + off, lo);
                    }
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(q_ready);
        };
        auto next_live = [&](int b) { b += gridDim.x; while (b < B && skip[b]) b += gridDim.x; return b; };
        int b = (int)blockIdx.x;
        while (b < B && skip[b]) b += gridDim.x;
        // mask bits of this thread's row in the two tiles of a snippet (bit t), fetched one snippet ahead: the load is an L2
        // round trip that would otherwise sit in front of every snippet's first softmax
        auto load_mask = [&](int b) {
            uint32_t m = 0;
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (ROWS * t + e < Tm && __ldg(mask + (size_t)b * Tm + ROWS * t + e) != 0) m |= 1u << t;
            return m;
        };
        uint32_t mk = 0;
        if (b < B) { write_q(b); mk = load_mask(b); }
        while (b < B) {
            const int bn = next_live(b);
            float mx[WT], lsum[WT], m0[WT];                     // running max, running sum, the max tile 0 was exponentiated against
#pragma unroll
            for (int w = 0; w < WT; ++w) { mx[w] = -INFINITY; lsum[w] = 0.0f; m0[w] = -INFINITY; }
            for (int t = 0; t < n_tiles; ++t) {
                const bool valid = (mk >> t) & 1u;
                if (!mbar_wait(&s_ready[t], (t == 0 ? ns0++ : ns1++) & 1, abort_flag)) return;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t r0[16], r1[16], r2[16];                 // V_hi.Q_hi | V_hi.Q_lo | V_lo.Q_hi
                tmem_ld16(tlane + (uint32_t)(ACC * t), r0);
                tmem_ld16(tlane + (uint32_t)(ACC * t + NB), r1);
                if (PLANES == 2) tmem_ld16(tlane + (uint32_t)(ACC * t + 2 * NB), r2);
                tmem_wait_ld16(r0);
                tmem_wait_ld16(r1);
                if (PLANES == 2) tmem_wait_ld16(r2);
                float s[WT], mnew[WT];
#pragma unroll
                for (int w = 0; w < WT; ++w) {
                    const float sv = (PLANES == 2) ? (__uint_as_float(r2[w]) + __uint_as_float(r1[w])) + __uint_as_float(r0[w])
                                                   : __uint_as_float(r1[w]) + __uint_as_float(r0[w]);
                    s[w] = valid ? sv : -INFINITY;
                    mnew[w] = s[w];
                }
                reduce16(mnew, true);
#pragma unroll
                for (int w = 0; w < WT; ++w) {
                    mnew[w] = fmaxf(mx[w], mnew[w]);
                    const float scale = (mnew[w] == -INFINITY) ? 1.0f : __expf(mx[w] - mnew[w]);   // mx = -inf, finite new max -> 0
                    const float pw = (s[w] == -INFINITY) ? 0.0f : __expf(s[w] - mnew[w]);
                    lsum[w] = lsum[w] * scale + pw;
                    mx[w] = mnew[w];
                    if (t == 0) m0[w] = mnew[w];
                    s[w] = pw;
                }
                // probabilities of this tile -> B operand of the context MMAs: P[beam w][row e], fp16 hi / lo
#pragma unroll
                for (int w = 0; w < WT; ++w)
                    if (w < W) {
                        uint16_t hi, lo;
                        split_f16(s[w], hi, lo);
                        const uint32_t off = (uint32_t)(t * P_TILE) + (uint32_t)(e >> 6) * QBOX + kmajor_off(w, e & 63);
                        sts16(p0 + off, hi);
                        sts16(p0 + LO_ROWS + off, lo);
                    }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(p_ready);
            }
            // The score MMAs of this snippet have retired (s_ready of its last tile was seen): the next snippet's queries can
            // go in now, so that its score MMAs are issued while this snippet's last context MMAs and output are in flight.
            uint32_t mk_next = 0;
            if (bn < B) { write_q(bn); mk_next = load_mask(bn); }
            // ---- normalise and write ctx[w, :] into the attention-layer input [h | ctx]
            reduce16(lsum, false);
            float inv[WT], sc0[WT];                              // 1 / sum, and tile 0's exp(m_0 - m_final) (tile 1 was taken against m_final)
#pragma unroll
            for (int w = 0; w < WT; ++w) {
                inv[w] = (lsum[w] > 0.0f) ? 1.0f / lsum[w] : __int_as_float(0x7fc00000);   // all masked -> NaN like tfa
                sc0[w] = (n_tiles == 1 || m0[w] == -INFINITY) ? ((n_tiles == 1) ? 1.0f : 0.0f) : __expf(m0[w] - mx[w]);
            }
            if (!mbar_wait(c_done, nc++ & 1, abort_flag)) return;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float acc[WT];
#pragma unroll
                for (int w = 0; w < WT; ++w) acc[w] = 0.0f;
                for (int t = 0; t < n_tiles; ++t) {
                    const uint32_t col = tlane + (uint32_t)(2 * ACC + ACC * (2 * t + h));
                    uint32_t c0[16], c1[16], c2[16];
                    tmem_ld16(col, c0);
                    tmem_ld16(col + NB, c1);
                    if (PLANES == 2) tmem_ld16(col + 2 * NB, c2);
                    tmem_wait_ld16(c0);
                    tmem_wait_ld16(c1);
                    if (PLANES == 2) tmem_wait_ld16(c2);
#pragma unroll
                    for (int w = 0; w < WT; ++w) {
                        const float v = (PLANES == 2) ? (__uint_as_float(c2[w]) + __uint_as_float(c1[w])) + __uint_as_float(c0[w])
                                                      : __uint_as_float(c1[w]) + __uint_as_float(c0[w]);
                        acc[w] = (t == 0) ? v * sc0[w] : acc[w] + v;
                    }
                }
#pragma unroll
                for (int w = 0; w < WT; ++w)
                    if (w < W) {
                        const size_t o = ((size_t)b * W + w) * (3 * UNITS) + UNITS + 128 * h + e;
                        const float v = acc[w] * inv[w];
                        if (xp_hi != nullptr) {      // fp16 hi / lo planes of [h | ctx]: A operand of the attention-layer GEMM
                            uint16_t hi, lo;
                            split_f16(v, hi, lo);
                            xp_hi[o] = hi; xp_lo[o] = lo;
                        } else xa[o] = v;
                    }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            b = bn;
            mk = mk_next;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

static int make_map3(CUtensorMap *map, const void *base, long long B, int Tm) {
    gemm::tc::EncodeTiledFn fn = gemm::tc::encode_fn();
    if (!fn) return fail(RVB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)ENC_OUT, (cuuint64_t)Tm, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ENC_OUT * 2, (cuuint64_t)Tm * ENC_OUT * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RVB_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return RVB_OK;
}

template <int WT, int PLANES>
static int launch(const CUtensorMap &mh, const CUtensorMap &ml, const uint8_t *mask, const float *Q, float *xa, const int32_t *skip,
                  int B, int Tm, int W, int *abort_flag, unsigned grid, cudaStream_t s, uint16_t *xp_hi, uint16_t *xp_lo) {
    // per launch, not once per process: the attribute is per device, and ShardedBasecaller drives every GPU from one process
    RVB_CUDA(cudaFuncSetAttribute(attention_tc_kernel<WT, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attention_tc_kernel<WT, PLANES><<<grid, THREADS, SMEM, s>>>(mh, ml, mask, Q, xa, skip, B, Tm, W, abort_flag, xp_hi, xp_lo);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}

// v_lo == nullptr: the memory is a single fp16 plane (reduced-precision mode)
int run(const uint16_t *v_hi, const uint16_t *v_lo, const uint8_t *mask, const float *Q, float *xa, const int32_t *skip,
        int B, int Tm, int W, int *abort_flag, cudaStream_t s, uint16_t *xp_hi, uint16_t *xp_lo) {
    if (B <= 0) return RVB_OK;
    if (W < 1 || W > WMAX_ || Tm < 1 || Tm > 2 * ROWS) return fail(RVB_ERR_ARG, "attention_tc: unsupported width %d / memory length %d", W, Tm);
    CUtensorMap mh, ml;
    RVB_CHECK(make_map3(&mh, v_hi, B, Tm));
    if (v_lo != nullptr) RVB_CHECK(make_map3(&ml, v_lo, B, Tm));
    else ml = mh;
    int dev = 0, sms = 0;
    RVB_CUDA(cudaGetDevice(&dev));
    RVB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(B < sms ? B : sms);
    if (v_lo != nullptr)
        return (W <= 5) ? launch<5, 2>(mh, ml, mask, Q, xa, skip, B, Tm, W, abort_flag, grid, s, xp_hi, xp_lo)
                        : launch<9, 2>(mh, ml, mask, Q, xa, skip, B, Tm, W, abort_flag, grid, s, xp_hi, xp_lo);
    if (W == 1) return launch<1, 1>(mh, ml, mask, Q, xa, skip, B, Tm, W, abort_flag, grid, s, xp_hi, xp_lo);
    return (W <= 5) ? launch<5, 1>(mh, ml, mask, Q, xa, skip, B, Tm, W, abort_flag, grid, s, xp_hi, xp_lo)
                    : launch<9, 1>(mh, ml, mask, Q, xa, skip, B, Tm, W, abort_flag, grid, s, xp_hi, xp_lo);
}

}  // namespace atc
}  // namespace rvb
