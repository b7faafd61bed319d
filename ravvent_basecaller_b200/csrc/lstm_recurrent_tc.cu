// K3 on tensor cores -- persistent recurrent LSTM with tcgen05.mma cta_group::2, fp32-level accuracy.
//
// Replaces the Keras RNN(LSTMCell) while-loop of Encoder.call (reference basecaller.py:19-32, 48-59).
//
// A thread-block cluster of two CTAs handles (256 snippets, one direction): each CTA owns 128 batch rows
// and their full 512 gate columns as fp32 accumulators in its own TMEM (128 lanes x 512 columns = all of
// it).  The recurrent kernel U is the B operand of the pair: each CTA keeps HALF of the gate columns
// resident in shared memory for all T steps, as fp16 hi + fp16 lo (U = U_hi + U_lo, 64 KB + 64 KB), and the
// 2-SM MMA reads both halves -- no per-step weight traffic and no DSMEM exchange of h.
// Split precision: h = h_hi + h_lo (fp16 each); per step the leader CTA issues
//     D  = h_lo.U_hi + h_hi.U_lo + h_hi.U_hi          (fp16 operands, fp32 accumulate)
// which drops only the lo.lo term (~2^-22 relative): the recurrence keeps fp32-parity accuracy on the
// fp16 pipe.  The epilogue warps (TMEM lane quarter = warp % 4) read the accumulators with tcgen05.ld, add
// the input contribution (layer 0: x_t.W + b from shared memory; layers > 0: the pre-projected gates
// written by K2), apply the fused gate nonlinearities / cell update (c stays in registers), write y to HBM
// and store the next h as fp16 hi/lo straight into the 128-byte-swizzled K-major A-operand tiles.
// MMAs and epilogue of consecutive steps overlap through a wavefront schedule over four column quarters
// (see the kernel comment): per step four cluster-scope mbarriers (quarter drained + K-block written ->
// MMA) and four multicast tcgen05.commit (quarter accumulated -> both CTAs' epilogues).  Waits are bounded.
#include <cuda_fp16.h>
#include <type_traits>

#include "kernels.cuh"
#include "cell_math.cuh"

namespace rvb {
namespace rectc {

#ifndef RVB_REC_EPI_WARPS
#define RVB_REC_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = RVB_REC_EPI_WARPS;      // 8 or 16: TMEM lane quarter = warp % 4, column group = (warp - 4) / 4
constexpr int CGROUPS = EPI_WARPS / 4;           // column groups per quarter (2 or 4)
constexpr int NQ = 4;                            // quarters of the gate columns = K-blocks of the next h (wavefront schedule)
constexpr int UQ = UNITS / NQ;                   // 32 units per quarter
constexpr int QCOLS = 4 * UQ;                    // 128 accumulator columns per quarter
constexpr int UPQ = UQ / CGROUPS;                // units per epilogue thread and quarter (16 or 8)
constexpr int CPQ = UPQ / 8;                     // 8-unit chunks per epilogue thread and quarter
constexpr int UPT = NQ * UPQ;                    // units (cell-state registers) per epilogue thread (64 or 32)
constexpr int NCH = NQ * CPQ;                    // 8-unit chunks per epilogue thread and step
constexpr int THREADS = 128 + 32 * EPI_WARPS;    // warpgroup 0: w1 = TMEM alloc + MMA issue (w0, w2, w3 idle); then the epilogue warpgroups
// Register budget: the launch allocates REG_LAUNCH per thread for all warps; warpgroup 0 then gives most of its share back
// (setmaxnreg.dec) and the epilogue warpgroups take it (setmaxnreg.inc), so the cell update keeps c, the pre-gate prefetch
// and a whole 8-unit chunk in registers without spilling.
#ifndef RVB_REC_PF_CHUNKS
#define RVB_REC_PF_CHUNKS 2
#endif
constexpr int PF_CHUNKS = RVB_REC_PF_CHUNKS;     // pre-gates: L2 prefetch this many chunks beyond the two register buffers (0 = off)
constexpr int REG_IDLE = 72;
constexpr int REG_EPI = EPI_WARPS == 8 ? 216 : 120;
constexpr int ROWS = 128;             // batch rows per CTA
constexpr int TILE_BYTES = 16384;     // [128 rows][64 fp16] K-major SW128 tile
constexpr int A_BYTES = 4 * TILE_BYTES;     // (hi|lo) x (kb 0|1)
constexpr int BQ_TILE_BYTES = 8192;         // [64 N rows of this CTA][64 fp16] K-major SW128 tile of one quarter
constexpr int B_BYTES = 16 * BQ_TILE_BYTES; // (quarter) x (hi|lo) x (kb)
constexpr int W0_FLOATS = 6 * GATES;        // layer 0: up to 5 feature rows + bias, [unit][gate] order
constexpr int W0_BYTES = W0_FLOATS * 4;
constexpr size_t SMEM = 1024 + A_BYTES + B_BYTES + W0_BYTES + 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 2.0f * fsig(2.0f * x) - 1.0f; }

using namespace cellmath;      // packed cell-update math, cell_math.cuh
__device__ __forceinline__ void lstm_pointwise(float zi, float zf, float zg, float zo, float c, float &cn, float &hn) {
    f32x2 c2, h2;
    lstm_pointwise2(pk(zi, zi), pk(zf, zf), pk(zg, zg), pk(zo, zo), pk(c, c), c2, h2);
    float t;
    upk(c2, cn, t); upk(h2, hn, t);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait with the default (CTA-scope) acquire.  NOTE: never poll with `.acquire.cluster`: every such
// try_wait makes ptxas emit CCTL.IVALL (L1 invalidate) -- it was 24 % of this kernel's stall samples.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *abort_flag) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
        if ((it & 1023u) == 1023u && *reinterpret_cast<volatile int *>(abort_flag) != 0) return false;
    }
    atomicExch(abort_flag, 1);
    return false;
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {          // K-major SW128, SBO = 1024 B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, A = B = fp16 (format 0), D = fp32, K-major both, M = 256 (pair), N = 128 (one quarter; 64 N rows per CTA)
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | (((uint32_t)QCOLS >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
// same, with the descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_f16_2sm_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {         // one lane of a converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// wait for this thread's TMEM loads; the registers are in/out operands so that no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}
// explicit shared-space accesses (the carved-up dynamic buffer is a generic pointer to the compiler: LD / ST instead of LDS / STS)
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Store 8 consecutive units of h (k = 64*kb + 8*chunk .. +7) of `row` as fp16 hi and lo into the A tiles (shared addresses).
__device__ __forceinline__ void store_h8(uint32_t hi_tile, uint32_t lo_tile, int row, int chunk, const float (&h)[8],
                                         uint4 &hi_out, uint4 &lo_out) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // packed conversions (F2FP.PACK_AB) keep the split off the XU pipe, which the gate nonlinearities saturate
        const __half2 hh = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(h[2 * i] - hf.x, h[2 * i + 1] - hf.y);
        hi[i] = *reinterpret_cast<const uint32_t *>(&hh);
        lo[i] = *reinterpret_cast<const uint32_t *>(&ll);
    }
    const uint32_t off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
    hi_out = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    lo_out = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    sts128(hi_tile + off, hi_out);
    sts128(lo_tile + off, lo_out);
}

// NPASS = 3: fp32-parity split precision (h_lo.U_hi + h_hi.U_lo + h_hi.U_hi); NPASS = 1: reduced-precision mode,
// a single h_hi.U_hi pass (fp16 operands, fp32 accumulate and cell state).
//
// Wavefront schedule.  The 512 gate columns (= [unit][gate], 4 per unit) are cut into NQ = 4 quarters of 32 units; quarter
// n of the accumulators and K-block n of the next h are the same 32 units.  Block (n, k) of a step = the MMAs that add
// h[:, K-block k] . U[K-block k, quarter n] into accumulator quarter n (NPASS x 2 instructions of M256 N128 K16).  Block
// (n, k) of step s+1 needs only (a) quarter n drained by the step-s epilogue and (b) K-block k written by it.  The epilogue
// walks the quarters in order and signals both events per quarter (`drained[j]` after its last TMEM load of quarter j,
// `written[j]` after the h stores), so the issuer runs blocks (j, i < j) after drained[j] and (i < j, j), (j, j) after
// written[j]: all the tensor work of step s+1 except the four blocks (i, 3) runs UNDER the epilogue of step s, and the
// epilogue of step s+1 starts one block after the one of step s ends.  One h buffer is enough: the accumulator barrier of
// quarter k is committed after block (3, k), the last reader of K-block k of the old h.
template <int F, bool PRE, int NPASS, bool GRU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) lstm_rec_tc_kernel(Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char *a_tiles = smem;                                   // [hi kb0][hi kb1][lo kb0][lo kb1], kb = K-block of 64
    unsigned char *b_tiles = smem + A_BYTES;                         // [(quarter*2 + part)*2 + kb] tiles of [64 N rows][64 K]
    float *w0s = reinterpret_cast<float *>(smem + A_BYTES + B_BYTES);   // [(F + 1)][512], layer 0 only
    uint64_t *drained = reinterpret_cast<uint64_t *>(smem + A_BYTES + B_BYTES + W0_BYTES);   // [NQ] (leader CTA's copies are used)
    uint64_t *written = drained + NQ;                 // [NQ]
    uint64_t *acc_ready = written + NQ;               // [NQ]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_ready + NQ);

    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cid = blockIdx.x >> 1;
    const int dir = cid & 1;
    const int b0 = (cid >> 1) * (2 * ROWS) + (int)rank * ROWS;
    const int T = p.T, B = p.B;

    // ---- one-time setup: resident B operand (pre-swizzled image), layer-0 rows, barriers, TMEM ----------
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.bimg + (size_t)(dir * 2 + rank) * (B_BYTES / 2));   // bimg counts uint16
        uint4 *dst = reinterpret_cast<uint4 *>(b_tiles);
        for (int i = threadIdx.x; i < B_BYTES / 16; i += THREADS) dst[i] = __ldg(src + i);
        if (!PRE)
            for (int i = threadIdx.x; i < (F + 1) * GATES; i += THREADS) w0s[i] = __ldg(p.w0 + (size_t)dir * W0_FLOATS + i);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int n = 0; n < NQ; ++n) {
            mbar_init(&drained[n], 2 * EPI_WARPS);    // epilogue warps x 2 CTAs
            mbar_init(&written[n], 2 * EPI_WARPS);
            mbar_init(&acc_ready[n], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_IDLE));
        // ================= MMA issuer: one thread of the leader CTA drives both SMs =================
        // The whole warp runs the schedule (warp-uniform control flow and descriptor arithmetic -> uniform datapath); only the
        // tcgen05 instructions themselves are predicated on one elected lane.
        if (warp == 1 && rank == 0) {
            // descriptors differ only in their low word (start address >> 4): keep the two bases and add immediates per
            // instruction -- hoisting all 192 descriptors out of the step loop costs more registers than this warp owns
            const uint64_t ad0 = make_desc(smem_u32(a_tiles)), bd0 = make_desc(smem_u32(b_tiles));
            const uint32_t dhi = (uint32_t)(ad0 >> 32);
            uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
            // block (n, k): quarter n of the accumulators += K-block k (32 units = 2 instructions of K 16) of h
            auto block = [&](int n, int k) {
#pragma unroll
                for (int combo = (NPASS == 3 ? 0 : 2); combo < 3; ++combo) {
                    const int pa = (combo == 0) ? 1 : 0;      // h_lo.U_hi, h_hi.U_lo, h_hi.U_hi
                    const int pb = (combo == 1) ? 1 : 0;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint32_t koff = (uint32_t)(((k & 1) * 2 + kk) * 32);
                        const uint32_t al = alo + (uint32_t)(((pa * 2 + (k >> 1)) * TILE_BYTES + koff) >> 4);
                        const uint32_t bl = blo + (uint32_t)(((((n * 2 + pb) * 2 + (k >> 1)) * BQ_TILE_BYTES) + koff) >> 4);
                        const bool first = (k == 0) && (combo == (NPASS == 3 ? 0 : 2)) && (kk == 0);
                        umma_f16_2sm_lo(tmem_base + (uint32_t)(n * QCOLS), al, bl, dhi, first ? 0u : 1u);
                    }
                }
            };
            bool ok = true;
            for (int s = 0; s < T && ok; ++s) {
                asm volatile("" : "+r"(alo), "+r"(blo));                      // opaque per step: no hoisting of the descriptor words
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    // Each CTA's h stores are made visible to ITS OWN async proxy (fence.proxy.async) before its warps arrive,
                    // and the tensor cores read them only after this thread has seen every arrival.  No cluster-scope fence
                    // here or on the arriving side: either one lowers to MEMBAR.ALL.GPU on the per-quarter critical path.
                    if (j > 0) {
                        if (!mbar_wait(&drained[j], s & 1, p.abort_flag)) { ok = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
#pragma unroll
                            for (int i = 0; i < j; ++i) block(j, i);          // (j, 0) opens quarter j (overwrite)
                        }
                        __syncwarp();
                    }
                    if (!mbar_wait(&written[j], s & 1, p.abort_flag)) { ok = false; break; }
                    if (j == 0 && !mbar_wait(&drained[0], s & 1, p.abort_flag)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
#pragma unroll
                        for (int i = 0; i < j; ++i) {
                            block(i, j);
                            if (j == NQ - 1) umma_commit_2sm(&acc_ready[i]);  // quarter i complete, K-block i of the old h free
                        }
                        block(j, j);
                        if (j == NQ - 1) umma_commit_2sm(&acc_ready[j]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_EPI));
        // ================= epilogue warps: gates, cell update, next h ================================
        const int q = warp & 3, cg = (warp - 4) >> 2;
        const int row = 32 * q + lane;
        const int b = b0 + row;
        const bool live = b < B;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const uint32_t a_sh = smem_u32(a_tiles), w0_sh = smem_u32(w0s);
        // this thread's units: quarter n, chunk ch  ->  u0(n, ch) = UQ*n + UPQ*cg + 8*ch  (8 consecutive units)
        float c[UPT];
        {
            float h0[UPT];
#pragma unroll
            for (int i = 0; i < UPT; ++i) { c[i] = 0.0f; h0[i] = 0.0f; }
            if (p.state_in != nullptr && live) {
                const float *si = p.state_in + (((size_t)b * 2 + dir) * 2) * UNITS;
#pragma unroll
                for (int n = 0; n < NQ; ++n)
#pragma unroll
                    for (int i = 0; i < UPQ; i += 4) {
                        const int u = UQ * n + UPQ * cg + i;
                        const float4 hv = *reinterpret_cast<const float4 *>(si + u);
                        const float4 cv = *reinterpret_cast<const float4 *>(si + UNITS + u);
                        const int ci = n * UPQ + i;
                        h0[ci] = hv.x; h0[ci + 1] = hv.y; h0[ci + 2] = hv.z; h0[ci + 3] = hv.w;
                        if (GRU) { c[ci] = hv.x; c[ci + 1] = hv.y; c[ci + 2] = hv.z; c[ci + 3] = hv.w; }      // GRU: the carried state is h
                        else { c[ci] = cv.x; c[ci + 1] = cv.y; c[ci + 2] = cv.z; c[ci + 3] = cv.w; }
                    }
            }
#pragma unroll
            for (int n = 0; n < NQ; ++n)
#pragma unroll
                for (int ch = 0; ch < CPQ; ++ch) {
                    const int u0 = UQ * n + UPQ * cg + 8 * ch;
                    float h8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) h8[u] = h0[n * UPQ + 8 * ch + u];
                    uint4 dh, dl;
                    store_h8(a_sh + (u0 >> 6) * TILE_BYTES, a_sh + (2 + (u0 >> 6)) * TILE_BYTES, row, (u0 & 63) >> 3, h8, dh, dl);
                }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int n = 0; n < NQ; ++n) { mbar_arrive_remote(&drained[n], 0); mbar_arrive_remote(&written[n], 0); }
        }

        bool ok = true;
        float *so = (p.state_out != nullptr && live) ? p.state_out + (((size_t)b * 2 + dir) * 2) * UNITS : nullptr;
        // pre-gates of (t, this row): float4 = one unit's four gates; consecutive units are gq float4 apart.  Blocked layout
        // (what K2 writes for the encoders): [row tile of 128][unit][row][4], so the 32 lanes of a warp read 512 contiguous bytes.
        // Rows b >= B load from a clamped (valid) row instead of being predicated off: a conditional refill makes the compiler
        // merge the loaded registers right after the load, i.e. wait for it, which defeats the prefetch.
        const long long gq = p.g_blocked ? ROWS : 1;
        const int bg = p.g_blocked ? (b < (int)p.g_rows_per_t ? b : 0) : (live ? b : 0);
        // Reduced-precision mode (NPASS == 1): K2 writes the pre-gates as fp16, 8 bytes per unit in the same blocked order --
        // half the K2-write / K3-read traffic of the encoder's largest tensor.
        constexpr bool G16 = PRE && NPASS == 1;
        using GV = typename std::conditional<G16, uint2, float4>::type;
        auto g_at = [&](int tt) -> const GV * {
            if (G16 || p.g_blocked) {
                const size_t rg = (size_t)tt * p.g_rows_per_t + bg;
                return reinterpret_cast<const GV *>(p.G) + ((rg >> 7) * (2 * GATES / 4) + dir * (GATES / 4)) * ROWS + (rg & 127);
            }
            return reinterpret_cast<const GV *>(p.G + (size_t)bg * p.g_bs + (size_t)tt * p.g_ts + dir * GATES);
        };
        auto g_quad = [](const GV &v, float &a, float &b2, float &c2, float &d2) {
            if constexpr (G16) {
                const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&v.x)), hi = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
                a = lo.x; b2 = lo.y; c2 = hi.x; d2 = hi.y;
            } else { a = v.x; b2 = v.y; c2 = v.z; d2 = v.w; }
        };
        // chunk i of a step (i = n*CPQ + ch) -> first unit
        auto unit_of = [&](int i) -> int { return UQ * (i / CPQ) + UPQ * cg + 8 * (i % CPQ); };
        // Software pipeline: the pre-gates of chunks i+1 and i+2 are in flight while chunk i is computed (two register
        // buffers by chunk parity; NCH is even, so the parity carries over the step boundary), and the accumulators of chunk
        // i+1 are loaded from TMEM as soon as those of chunk i have been consumed.
        GV gbuf[2][8];
        float xnext[F];
        const int bx = live ? b : 0;
        {
            const int t0 = dir ? T - 1 : 0;
            if (PRE) {
                const GV *g0 = g_at(t0);
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int i = 0; i < 8; ++i) gbuf[k][i] = __ldg(g0 + (size_t)(unit_of(k) + i) * gq);
            } else {
#pragma unroll
                for (int f = 0; f < F; ++f) xnext[f] = __ldg(p.x + ((size_t)bx * T + t0) * F + f);
            }
        }
        for (int s = 0; s < T && ok; ++s) {
            const int t = dir ? T - 1 - s : s;
            const int tn = dir ? T - 2 - s : s + 1;                     // next step's timestep (valid while s + 1 < T)
            float xin[F];
            if (!PRE) {
#pragma unroll
                for (int f = 0; f < F; ++f) xin[f] = xnext[f];
                const int tx = (s + 1 < T) ? tn : t;                    // last step: a harmless reload
#pragma unroll
                for (int f = 0; f < F; ++f) xnext[f] = __ldg(p.x + ((size_t)bx * T + tx) * F + f);
            }
            const GV *grow = PRE ? g_at(t) : nullptr;
            const GV *grow_next = PRE ? g_at(s + 1 < T ? tn : t) : nullptr;          // last step: harmless reloads of this step's rows
            float *yrow = (p.y16_hi == nullptr) ? p.y + (size_t)b * p.y_bs + (size_t)t * p.y_ts + dir * UNITS : nullptr;
            uint32_t r[32];
            ok = mbar_wait(&acc_ready[0], s & 1, p.abort_flag);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld32(tlane + (uint32_t)(4 * unit_of(0)), r);
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                // one 32-column chunk = 8 units at a time: the eight cell updates form one unrolled block, so eight
                // independent EX2 -> RCP -> EX2 -> RCP chains are in flight
                const int n = i / CPQ, ch = i % CPQ;
                const int u0 = unit_of(i);
                float z[32];
                if (PRE) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) g_quad(gbuf[i & 1][k], z[4 * k], z[4 * k + 1], z[4 * k + 2], z[4 * k + 3]);
                } else {
                    const uint32_t wr = w0_sh + 16 * u0;
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        float4 a = lds128(wr + (F * GATES + k) * 4);          // bias row
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float4 w = lds128(wr + (f * GATES + k) * 4);
                            a.x = fmaf(xin[f], w.x, a.x); a.y = fmaf(xin[f], w.y, a.y);
                            a.z = fmaf(xin[f], w.z, a.z); a.w = fmaf(xin[f], w.w, a.w);
                        }
                        z[k] = a.x; z[k + 1] = a.y; z[k + 2] = a.z; z[k + 3] = a.w;
                    }
                }
                tmem_wait_ld(r);
#pragma unroll
                for (int k = 0; k < 32; ++k) z[k] += __uint_as_float(r[k]);
                if (ch == CPQ - 1) {
                    // this warp has read all of quarter n: blocks (n, i < n) of the next step may overwrite it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && s + 1 < T) mbar_arrive_remote(&drained[n], 0);
                }
                // refill this parity's buffer with chunk i+2 (same step, or the next step's chunk i+2-NCH).  The proxy fence
                // at the end of a quarter (MEMBAR.ALL.CTA) waits for this thread's outstanding loads, so the last chunk of a
                // quarter issues its refill AFTER the fence and the others as early as possible.
                auto refill = [&]() {
                    const GV *nsrc = (i + 2 < NCH) ? grow : grow_next;
                    const int un = unit_of((i + 2) % NCH);
#pragma unroll
                    for (int k = 0; k < 8; ++k) gbuf[i & 1][k] = __ldg(nsrc + (size_t)(un + k) * gq);
                    if (PF_CHUNKS > 0) {
                        // registers hold two chunks (64 KB in flight per SM is not enough for the DRAM latency): keep PF_CHUNKS more on
                        // their way into L2.  Measured: 13.2 -> 11.6 us per step; a bulk prefetch of whole steps ADDS DRAM traffic.
                        const GV *psrc = (i + 2 + PF_CHUNKS < NCH) ? grow : grow_next;
                        const int up = unit_of((i + 2 + PF_CHUNKS) % NCH);
#pragma unroll
                        for (int k = 0; k < 8; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(psrc + (size_t)(up + k) * gq));
                    }
                };
                if (PRE && ch != CPQ - 1) refill();
                if (i + 1 < NCH) {
                    if ((i + 1) % CPQ == 0) {
                        ok = mbar_wait(&acc_ready[n + 1], s & 1, p.abort_flag);
                        if (!ok) break;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    tmem_ld32(tlane + (uint32_t)(4 * unit_of(i + 1)), r);
                }
                float h8[8];
#pragma unroll
                for (int u = 0; u < 8; u += 2) {                 // units u, u+1 as one packed pair
                    const int ci = n * UPQ + 8 * ch + u;
                    f32x2 cn, hn;
                    if (GRU) {
                        gru_pointwise2(pk(z[4 * u + 0], z[4 * u + 4]), pk(z[4 * u + 1], z[4 * u + 5]), pk(z[4 * u + 2], z[4 * u + 6]),
                                       pk(z[4 * u + 3], z[4 * u + 7]), pk(c[ci], c[ci + 1]), hn);
                        cn = hn;
                    } else
                        lstm_pointwise2(pk(z[4 * u + 0], z[4 * u + 4]), pk(z[4 * u + 1], z[4 * u + 5]), pk(z[4 * u + 2], z[4 * u + 6]),
                                        pk(z[4 * u + 3], z[4 * u + 7]), pk(c[ci], c[ci + 1]), cn, hn);
                    upk(cn, c[ci], c[ci + 1]);
                    upk(hn, h8[u], h8[u + 1]);
                }
                uint4 ph, pl;
                store_h8(a_sh + (u0 >> 6) * TILE_BYTES, a_sh + (2 + (u0 >> 6)) * TILE_BYTES, row, (u0 & 63) >> 3, h8, ph, pl);
                if (live) {
                    if (p.y16_hi != nullptr) {          // intermediate layer: fp16 hi/lo planes for the next projection
                        const size_t o = (size_t)b * p.y16_bs + (size_t)t * p.y16_ts + dir * UNITS + u0;
                        *reinterpret_cast<uint4 *>(p.y16_hi + o) = ph;
                        *reinterpret_cast<uint4 *>(p.y16_lo + o) = pl;
                    } else {
                        *reinterpret_cast<float4 *>(yrow + u0) = make_float4(h8[0], h8[1], h8[2], h8[3]);
                        *reinterpret_cast<float4 *>(yrow + u0 + 4) = make_float4(h8[4], h8[5], h8[6], h8[7]);
                        if (p.yv16 != nullptr)      // fp16 copy of the attention memory for the reduced-precision decoder
                            *reinterpret_cast<uint4 *>(p.yv16 + (size_t)b * p.y_bs + (size_t)t * p.y_ts + dir * UNITS + u0) = ph;
                    }
                    if (so != nullptr && s == T - 1) {
                        *reinterpret_cast<float4 *>(so + u0) = make_float4(h8[0], h8[1], h8[2], h8[3]);
                        *reinterpret_cast<float4 *>(so + u0 + 4) = make_float4(h8[4], h8[5], h8[6], h8[7]);
                    }
                }
                if (ch == CPQ - 1) {
                    // K-block n of the next h is written: release the blocks (i < n, n) and (n, n)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && s + 1 < T) mbar_arrive_remote(&written[n], 0);
                    if (PRE) refill();
                }
            }
        }
        if (so != nullptr && ok) {
#pragma unroll
            for (int n = 0; n < NQ; ++n)
#pragma unroll
                for (int i = 0; i < UPQ; i += 4) {
                    const int ci = n * UPQ + i;
                    *reinterpret_cast<float4 *>(so + UNITS + UQ * n + UPQ * cg + i) = make_float4(c[ci], c[ci + 1], c[ci + 2], c[ci + 3]);
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int F, bool PRE, int NPASS, bool GRU>
static int launch(const Params &p, cudaStream_t stream) {
    RVB_CUDA(cudaFuncSetAttribute(lstm_rec_tc_kernel<F, PRE, NPASS, GRU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    const int clusters = (p.B + 2 * ROWS - 1) / (2 * ROWS) * 2;          // x 2 directions
    {
        ProfScope ps(KK_REC, stream);
        lstm_rec_tc_kernel<F, PRE, NPASS, GRU><<<dim3((unsigned)(clusters * 2)), THREADS, SMEM, stream>>>(p);
    }
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

template <bool GRU>
static int run_cell(int feat, const Params &p, cudaStream_t stream) {
    const bool full = (p.precision == RVB_PREC_FP32);
    if (feat == 1) return full ? launch<1, false, 3, GRU>(p, stream) : launch<1, false, 1, GRU>(p, stream);
    if (feat == 5) return full ? launch<5, false, 3, GRU>(p, stream) : launch<5, false, 1, GRU>(p, stream);
    if (feat == 0) return full ? launch<1, true, 3, GRU>(p, stream) : launch<1, true, 1, GRU>(p, stream);
    return fail(RVB_ERR_ARG, "lstm_rec_tc: unsupported feature count %d", feat);
}

int run(int feat, const Params &p, cudaStream_t stream) {
    if (p.B <= 0 || p.T <= 0) return RVB_OK;
    if (feat == 0 && p.precision != RVB_PREC_FP32 && !(p.g16 && p.g_blocked))
        return fail(RVB_ERR_ARG, "lstm_rec_tc: the reduced-precision pre-gate launch takes blocked fp16 pre-gates (Params::g16)");
    if (p.g16 && (feat != 0 || p.precision == RVB_PREC_FP32)) return fail(RVB_ERR_ARG, "lstm_rec_tc: fp16 pre-gates only in reduced-precision mode");
    return p.gru ? run_cell<true>(feat, p, stream) : run_cell<false>(feat, p, stream);
}

// ---- host-side packing ---------------------------------------------------------------------------------
// B-operand image for (dir, rank): tiles [(quarter*2 + part)*2 + kb] of [64 N-rows][64 K] fp16, K-major SW128.
// N row n of quarter qn in CTA `rank` is gate column  n_global = qn*128 + rank*64 + n = unit*4 + gate.
void pack_b_image(const float *U /*[128][512] Keras order*/, int rank, uint16_t *img /*B_BYTES/2*/) {
    for (int qn = 0; qn < NQ; ++qn)
        for (int part = 0; part < 2; ++part)
            for (int kb = 0; kb < 2; ++kb)
                for (int n = 0; n < 64; ++n)
                    for (int k = 0; k < 64; ++k) {
                        const int ng = qn * QCOLS + rank * 64 + n;
                        const int unit = ng >> 2, gate = ng & 3;
                        const float v = U[(size_t)(kb * 64 + k) * GATES + gate * UNITS + unit];
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        const size_t off = (size_t)((qn * 2 + part) * 2 + kb) * BQ_TILE_BYTES + (n >> 3) * 1024 + (n & 7) * 128 +
                                           (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
                        img[off / 2] = __half_as_ushort(part == 0 ? hi : lo);
                    }
}

}  // namespace rectc
}  // namespace rvb
