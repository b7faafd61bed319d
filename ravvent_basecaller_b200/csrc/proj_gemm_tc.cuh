// K2 on the 5th-generation tensor cores:  C[M,N] = A[M,K] * W[K,N] (+ bias[N]),  fp32 in / fp32 out.
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a ring of shared-memory
//     stages, accumulators in TMEM, MMAs issued by one elected thread (tcgen05.mma kind::tf32),
//     accumulators read back with tcgen05.ld for the bias epilogue;
//   * fp32 parity ("3xTF32"): A = A_hi + A_lo and W = W_hi + W_lo with *_hi exactly representable in
//     tf32; D = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi recovers ~2^-21 relative accuracy on the tf32
//     pipe.  W is split once on the host; A is split per stage by four "split" warps between the TMA
//     landing and the MMA issue (in place for hi, a sibling buffer for lo; the swizzled placement is
//     irrelevant to an element-wise pass);
//   * NPASS = 1 skips the split (single tf32 pass) -- used by the bf16-tolerance mode.
//
// Tile: 128 rows x BN columns per CTA, BK = 32 fp32 (= one 128-byte swizzle row) per stage.
// Warp roles (256 threads): w0 = TMA producer, w1 = TMEM allocator + MMA issuer, w4..w7 = A split
// and epilogue (TMEM lane quarter = warp % 4).  Every mbarrier wait is bounded: on a timeout the
// kernel raises a flag and drains instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "kernels.cuh"
#include "cell_math.cuh"

namespace rvb {
namespace gemm {
namespace tc {

constexpr int BM = 128, BK = 32, THREADS = 256;
constexpr int UMMA_K = 8;                   // tf32: 32 bytes of K per instruction

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait; returns false (and the caller drains) if the phase never completes.
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

__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major; 1) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, K-major A and B.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 (fp16 operands, fp32 accumulate), K-major A and B
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

template <int BN, int NPASS>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 4;                  // 16 KB
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = (NPASS == 3) ? (2 * A_BYTES + 2 * B_BYTES) : (A_BYTES + B_BYTES);
    static constexpr int STAGES = (NPASS == 3) ? ((BN == 256) ? 2 : 3) : 4;
    static constexpr int TX_BYTES = A_BYTES + ((NPASS == 3) ? 2 : 1) * B_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int NPASS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
               const __grid_constant__ CUtensorMap map_blo, const float *__restrict__ bias, float *__restrict__ C,
               long long M, int N, int K, int *abort_flag) {
    using cfg = Cfg<BN, NPASS>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)cfg::STAGES * cfg::STAGE_BYTES);
    uint64_t *full = bars;                         // TMA landed
    uint64_t *ready = bars + cfg::STAGES;          // A split done (NPASS == 3)
    uint64_t *empty = bars + 2 * cfg::STAGES;      // MMAs of the stage retired
    uint64_t *acc_full = bars + 3 * cfg::STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * cfg::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = N / BN;
    const int n_tile = (int)(blockIdx.x % n_tiles), m_tile = (int)(blockIdx.x / n_tiles);
    const int num_kb = K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    auto stage_a = [&](int s) { return smem + (size_t)s * cfg::STAGE_BYTES; };
    auto stage_bhi = [&](int s) { return stage_a(s) + ((NPASS == 3) ? 2 : 1) * cfg::A_BYTES; };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
                if (round > 0 && !mbar_wait(&empty[s], (round - 1) & 1, abort_flag)) break;
                mbar_expect_tx(&full[s], cfg::TX_BYTES);
                tma_load_2d(&map_a, &full[s], stage_a(s), kb * BK, m_tile * BM);
                tma_load_2d(&map_bhi, &full[s], stage_bhi(s), kb * BK, n_tile * BN);
                if (NPASS == 3) tma_load_2d(&map_blo, &full[s], stage_bhi(s) + cfg::B_BYTES, kb * BK, n_tile * BN);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            bool ok = true;
            for (int kb = 0; kb < num_kb && ok; ++kb) {
                const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
                ok = mbar_wait((NPASS == 3) ? &ready[s] : &full[s], round & 1, abort_flag);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(stage_a(s)), b_hi = smem_u32(stage_bhi(s));
#pragma unroll
                for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                    const uint32_t koff = ks * UMMA_K * 4;           // bytes along K inside the swizzle row
                    const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                    if (NPASS == 3) {
                        umma_tf32(tmem_base, make_desc(a_hi + cfg::A_BYTES + koff), make_desc(b_hi + koff), idesc, first);
                        umma_tf32(tmem_base, make_desc(a_hi + koff), make_desc(b_hi + cfg::B_BYTES + koff), idesc, 1u);
                        umma_tf32(tmem_base, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
                    }
                }
                umma_commit(&empty[s]);            // frees the stage when these MMAs retire
            }
            umma_commit(acc_full);                 // arrives when every MMA above has retired
        }
    } else if (warp >= 4) {
        const int et = threadIdx.x - 128;          // 0..127
        bool ok = true;
        if (NPASS == 3) {
            // ===================== A split: hi in place, lo to the sibling buffer =====================
            for (int kb = 0; kb < num_kb && ok; ++kb) {
                const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
                ok = mbar_wait(&full[s], round & 1, abort_flag);
                if (!ok) break;
                float4 *hi = reinterpret_cast<float4 *>(stage_a(s));
                float4 *lo = reinterpret_cast<float4 *>(stage_a(s) + cfg::A_BYTES);
#pragma unroll
                for (int i = 0; i < cfg::A_BYTES / 16 / 128; ++i) {
                    float4 v = hi[et + i * 128];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[et + i * 128] = h;
                    lo[et + i * 128] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
                mbar_arrive(&ready[s]);
            }
        }
        // ===================== epilogue: TMEM -> registers -> (+bias) -> global =====================
        if (ok) ok = mbar_wait(acc_full, 0, abort_flag);
        if (ok) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int q = warp & 3;                                  // TMEM lane quarter of this warp
            const long long row = (long long)m_tile * BM + q * 32 + lane;
            float *crow = C + row * N + (size_t)n_tile * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        if (bias != nullptr) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + (size_t)n_tile * BN + c0 + j));
                            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        }
                        *reinterpret_cast<float4 *>(crow + c0 + j) = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Persistent variant (BN = 256): one CTA per SM loops over output tiles.  Two TMEM accumulator stages
// (2 x 256 columns) let the epilogue of tile i overlap the mainloop of tile i+1; the A split and the
// epilogue run on separate warp groups; C leaves through a 128-byte-swizzled staging buffer and TMA
// stores (cp.async.bulk.tensor ... global.shared::cta), which are fully coalesced and clip the ragged
// last row tile in hardware.
//   w0 TMA producer | w1 MMA issuer + TMEM alloc | w4..w7 A split (NPASS == 3) | w8..w11 epilogue
// ---------------------------------------------------------------------------------------------------
constexpr int PTHREADS = 384;
constexpr int PBN = 256;
constexpr int CSTAGE_BYTES = BM * 32 * 4;       // 128 rows x 32 fp32 columns
constexpr int CELL_TOKENS = 7;                  // rows of CellEpilogue::wtok when tok != nullptr (the decoder's vocabulary)

// PAIR: 0 = independent CTAs; 1 = cluster of two, W halves TMA-multicast to both; 2 = cluster of two driving one
// cta_group::2 MMA (M 256 = both CTAs' row tiles, each CTA holds only its half of the W stage -> 3 stages fit).
template <int NPASS, bool F16IN = false, int PAIR = 0>
struct PCfg {
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = (PAIR == 2 ? PBN / 2 : PBN) * BK * 4;
    static constexpr int STAGE_BYTES = (NPASS == 3) ? (2 * A_BYTES + 2 * B_BYTES) : (A_BYTES + B_BYTES);
    static constexpr int STAGES = (NPASS == 3) ? (PAIR == 2 ? 3 : 2) : 4;
    // fp16-plane inputs: A_lo also arrives by TMA; a stage then covers 64 K-elements (128 bytes of fp16)
    static constexpr int TX_BYTES = ((NPASS == 3 && F16IN) ? 2 : 1) * A_BYTES + ((NPASS == 3) ? 2 : 1) * B_BYTES;
    static constexpr int K_PER_STAGE = F16IN ? 64 : 32;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * CSTAGE_BYTES + 1024 + 256;
};

// Multicast load: the box lands at the same CTA-relative offset in every CTA of `mask`, and each of them gets the
// complete_tx on the mbarrier at the same CTA-relative offset.
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// cta_group::2 form: executed by both CTAs of the pair, the transaction bytes are credited to the LEADER's copy of the barrier
// (bit 24 of a shared-memory address selects the odd CTA of a pair; CUTLASS: SM100_TMA_2SM_LOAD_2D), so the MMA issuer
// waits on one barrier for both halves of a stage and no thread has to relay the peer's arrival.
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) {      // arrive on CTA `rank`'s copy of bar
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    // plain arrive, as K3 does for its `drained` barriers: the .release.cluster form lowers to MEMBAR.ALL.GPU (~1 us per warp and
    // tile here); the TMEM reads it orders are already complete (tcgen05.wait::ld + tcgen05.fence::before_thread_sync)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cluster_sync_pair() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}


// MC (fp16-plane encoder GEMM): CTAs are launched as clusters of two that walk the SAME column tile on adjacent row
// tiles in lockstep; each loads half of the W stage and multicasts it to both, so a CTA pulls 64 KB instead of 96 KB per
// stage out of L2 (the operand fill, not the MMA pipe, bounded the single-CTA form).  The MMAs stay cta_group::1; a
// stage is released by both CTAs' commits (empty barriers count 2, commits multicast to the pair).
template <int NPASS, bool F16IN, int PAIR = 0>
__global__ void __launch_bounds__(PTHREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_alo,
                          const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                          const __grid_constant__ CUtensorMap map_c,
                          const float *__restrict__ bias, long long M, int N, int K, int *abort_flag, float *__restrict__ c_blocked,
                          const CellEpilogue cell, int n_out, int blocked_half) {
    using cfg = PCfg<NPASS, F16IN, PAIR>;
    constexpr bool MC = PAIR == 1, SM2 = PAIR == 2;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char *cstage = smem + (size_t)cfg::STAGES * cfg::STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(cstage + 2 * CSTAGE_BYTES);
    uint64_t *full = bars, *ready = bars + cfg::STAGES, *empty = bars + 2 * cfg::STAGES;
    uint64_t *acc_full = bars + 3 * cfg::STAGES, *acc_empty = acc_full + 2;
    uint64_t *peer_full = acc_empty + 2;             // SM2, leader CTA: the peer's stage has landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(peer_full + cfg::STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = N / PBN;
    const long long m_tiles = (M + BM - 1) / BM;
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // work items: output tiles, or (MC) pairs of row-adjacent tiles of one column tile, one per CTA of the cluster
    const long long total = (long long)n_tiles * (PAIR ? (m_tiles + 1) / 2 : m_tiles);
    const long long w_first = PAIR ? (blockIdx.x >> 1) : blockIdx.x, w_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
    auto m_of = [&](long long w) -> long long { return PAIR ? 2 * (w / n_tiles) + rank : w / n_tiles; };
    const int num_kb = K / cfg::K_PER_STAGE;

    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], MC ? 2 : 1); }
        // SM2: the leader's acc_empty collects one arrival per epilogue warp of both CTAs
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], SM2 ? 16 : (F16IN ? 256 : 128)); }
        for (int s = 0; s < cfg::STAGES; ++s) mbar_init(&peer_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (SM2) { __syncthreads(); cluster_sync_pair(); }
    if (warp == 1) {
        if (SM2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (PAIR) cluster_sync_pair();                   // the peer's barriers / TMEM exist before anything is sent to them
    // Fused cell epilogue: the token rows of the input kernel (7 x N floats, or the one bias row of a stacked cell) go into the
    // staging area the plain epilogue would use, so that the per-chunk token add is a shared-memory read instead of a
    // dependent global load (the epilogue, not the MMA pipe, bounds this kernel).  Row pitch N + 4: rows land in different banks.
    float *tk_s = reinterpret_cast<float *>(cstage);
    const int tk_ld = N + 4;
    const bool tk_smem = cell.xa != nullptr && (size_t)CELL_TOKENS * (size_t)tk_ld * 4 <= (size_t)2 * CSTAGE_BYTES;
    if (tk_smem) {
        const int rows = cell.tok != nullptr ? CELL_TOKENS : 1;
        for (int i = threadIdx.x; i < rows * (N >> 2); i += PTHREADS) {
            const int r = i / (N >> 2), c4 = i - r * (N >> 2);
            *reinterpret_cast<float4 *>(tk_s + (size_t)r * tk_ld + 4 * c4) = __ldg(reinterpret_cast<const float4 *>(cell.wtok + (size_t)r * N) + c4);
        }
        __syncthreads();
    }

    auto stage_a = [&](int s) { return smem + (size_t)s * cfg::STAGE_BYTES; };
    auto stage_bhi = [&](int s) { return stage_a(s) + ((NPASS == 3) ? 2 : 1) * cfg::A_BYTES; };

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer =====
            long long g = 0; bool ok = true;
            for (long long tile = w_first; tile < total && ok; tile += w_step) {
                const int n_tile = (int)(tile % n_tiles); const long long m_tile = m_of(tile);
                for (int kb = 0; kb < num_kb; ++kb, ++g) {
                    const int s = (int)(g % cfg::STAGES); const long long round = g / cfg::STAGES;
                    if (round > 0 && !mbar_wait(&empty[s], (uint32_t)((round - 1) & 1), abort_flag)) { ok = false; break; }
                    const int kc = kb * cfg::K_PER_STAGE;
                    if (SM2) {
                        // both CTAs load their own A rows and their half of the W rows (the B operand of the pair's MMA is split
                        // across the CTAs); every byte is credited to the leader's barrier, which expects both CTAs' stages
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * cfg::TX_BYTES);
                        const int nrow = n_tile * PBN + (int)rank * (PBN / 2);
                        tma_load_2d_2sm(&map_a, &full[s], stage_a(s), kc, (int)(m_tile * BM));
                        if (NPASS == 3 && F16IN) tma_load_2d_2sm(&map_alo, &full[s], stage_a(s) + cfg::A_BYTES, kc, (int)(m_tile * BM));
                        tma_load_2d_2sm(&map_bhi, &full[s], stage_bhi(s), kc, nrow);
                        if (NPASS == 3) tma_load_2d_2sm(&map_blo, &full[s], stage_bhi(s) + cfg::B_BYTES, kc, nrow);
                        continue;
                    }
                    mbar_expect_tx(&full[s], cfg::TX_BYTES);
                    tma_load_2d(&map_a, &full[s], stage_a(s), kc, (int)(m_tile * BM));
                    if (NPASS == 3 && F16IN) tma_load_2d(&map_alo, &full[s], stage_a(s) + cfg::A_BYTES, kc, (int)(m_tile * BM));
                    if (MC) {   // this CTA's half of the W rows (box = PBN / 2 rows), delivered to both CTAs
                        const int half = (int)rank * (cfg::B_BYTES / 2), nrow = n_tile * PBN + (int)rank * (PBN / 2);
                        tma_load_2d_mc(&map_bhi, &full[s], stage_bhi(s) + half, kc, nrow, (uint16_t)3);
                        if (NPASS == 3) tma_load_2d_mc(&map_blo, &full[s], stage_bhi(s) + cfg::B_BYTES + half, kc, nrow, (uint16_t)3);
                    } else {
                        tma_load_2d(&map_bhi, &full[s], stage_bhi(s), kc, n_tile * PBN);
                        if (NPASS == 3) tma_load_2d(&map_blo, &full[s], stage_bhi(s) + cfg::B_BYTES, kc, n_tile * PBN);
                    }
                }
            }
        }
    } else if (warp == 1 && SM2 && rank == 1) {
        // the peer CTA issues no MMAs: the leader's cta_group::2 instructions read both CTAs' stages
    } else if (warp == 1) {
        if (lane == 0) {                                   // ===== MMA issuer =====
            constexpr uint32_t idesc = F16IN ? make_idesc_f16(SM2 ? 2 * BM : BM, PBN) : make_idesc_tf32(BM, PBN);
            long long g = 0, it = 0; bool ok = true;
            for (long long tile = w_first; tile < total && ok; tile += w_step, ++it) {
                const int as = (int)(it & 1); const long long ar = it >> 1;
                if (ar > 0 && !mbar_wait(&acc_empty[as], (uint32_t)((ar - 1) & 1), abort_flag)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(as * PBN);
                for (int kb = 0; kb < num_kb; ++kb, ++g) {
                    const int s = (int)(g % cfg::STAGES); const long long round = g / cfg::STAGES;
                    if (!mbar_wait((NPASS == 3 && !F16IN) ? &ready[s] : &full[s], (uint32_t)(round & 1), abort_flag)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(stage_a(s)), b_hi = smem_u32(stage_bhi(s));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {            // 4 x 32 bytes of K per 128-byte swizzle row
                        const uint32_t koff = ks * 32;
                        const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                        if (SM2) {
                            if (NPASS == 3) {
                                umma_f16_2sm(d, make_desc(a_hi + cfg::A_BYTES + koff), make_desc(b_hi + koff), idesc, first);
                                umma_f16_2sm(d, make_desc(a_hi + koff), make_desc(b_hi + cfg::B_BYTES + koff), idesc, 1u);
                                umma_f16_2sm(d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, 1u);
                            } else {
                                umma_f16_2sm(d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
                            }
                        } else if (F16IN) {
                            if (NPASS == 3) {
                                umma_f16(d, make_desc(a_hi + cfg::A_BYTES + koff), make_desc(b_hi + koff), idesc, first);
                                umma_f16(d, make_desc(a_hi + koff), make_desc(b_hi + cfg::B_BYTES + koff), idesc, 1u);
                                umma_f16(d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, 1u);
                            } else {
                                umma_f16(d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
                            }
                        } else if (NPASS == 3) {
                            umma_tf32(d, make_desc(a_hi + cfg::A_BYTES + koff), make_desc(b_hi + koff), idesc, first);
                            umma_tf32(d, make_desc(a_hi + koff), make_desc(b_hi + cfg::B_BYTES + koff), idesc, 1u);
                            umma_tf32(d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, 1u);
                        } else {
                            umma_tf32(d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
                        }
                    }
                    if (SM2) umma_commit_2sm(&empty[s]); else if (MC) umma_commit_mc(&empty[s], (uint16_t)3); else umma_commit(&empty[s]);
                }
                if (ok) { if (SM2) umma_commit_2sm(&acc_full[as]); else umma_commit(&acc_full[as]); }
            }
        }
    } else if (warp >= 4 && warp < 8 && !F16IN) {
        if (NPASS == 3) {                                  // ===== A split: hi in place, lo to the sibling buffer =====
            const int et = threadIdx.x - 128;
            long long g = 0; bool ok = true;
            for (long long tile = w_first; tile < total && ok; tile += w_step) {
                for (int kb = 0; kb < num_kb; ++kb, ++g) {
                    const int s = (int)(g % cfg::STAGES); const long long round = g / cfg::STAGES;
                    if (!mbar_wait(&full[s], (uint32_t)(round & 1), abort_flag)) { ok = false; break; }
                    float4 *hi = reinterpret_cast<float4 *>(stage_a(s));
                    float4 *lo = reinterpret_cast<float4 *>(stage_a(s) + cfg::A_BYTES);
#pragma unroll
                    for (int i = 0; i < cfg::A_BYTES / 16 / 128; ++i) {
                        float4 v = hi[et + i * 128];
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                        hi[et + i * 128] = h;
                        lo[et + i * 128] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&ready[s]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers (+bias) -> swizzled smem staging -> TMA store =====
        // fp32 inputs: one group (warps 8..11) with two staging buffers.  fp16-plane inputs need no split warps, so
        // warps 4..7 form a second group: the groups take alternate 32-column chunks, one staging buffer and one
        // named barrier each -- the epilogue, not the MMA pipe, bounded this kernel.
        constexpr int NGRP = F16IN ? 2 : 1;
        const int grp = (warp >= 8) ? 0 : 1;
        const int q = warp & 3;
        const int row = q * 32 + lane;                     // row of the tile == TMEM lane
        const int et = threadIdx.x - (grp == 0 ? 256 : 128);
        long long it = 0; int chunk_ctr = 0; bool ok = true;
        for (long long tile = w_first; tile < total && ok; tile += w_step, ++it) {
            const int n_tile = (int)(tile % n_tiles); const long long m_tile = m_of(tile);
            const int as = (int)(it & 1); const long long ar = it >> 1;
            // fused cell: what depends only on the row -- its token, the beam slot its state comes from, and the first chunk of
            // that state -- is fetched while the accumulator is still being produced
            const long long R = m_tile * BM + row;
            const bool rvalid = cell.xa != nullptr && R < M;
            int tokv = 0; long long src = 0;
            float4 ca = make_float4(0.f, 0.f, 0.f, 0.f), cb4 = ca;
            if (rvalid) {
                tokv = cell.tok != nullptr ? __ldg(cell.tok + R) : 0;
                const int Ri = (int)R;
                src = (long long)((Ri / cell.W) * cell.W + __ldg(cell.parent + R));
                const int u0 = (n_tile * PBN + 32 * grp) >> 2;
                ca = __ldg(reinterpret_cast<const float4 *>(cell.c_in + src * 128 + u0));
                cb4 = __ldg(reinterpret_cast<const float4 *>(cell.c_in + src * 128 + u0 + 4));
            }
            if (!mbar_wait(&acc_full[as], (uint32_t)(ar & 1), abort_flag)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 32 * grp; c0 < PBN; c0 += 32 * NGRP, ++chunk_ctr) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PBN + c0), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 32 * NGRP >= PBN) {               // this group's last read of the accumulator stage: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (SM2) { __syncwarp(); if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0); }
                    else mbar_arrive(&acc_empty[as]);
                }
                if (n_tile * PBN + c0 >= n_out) continue;  // columns beyond the output (a weight padded up to the 256-column tile)
                if (m_tile >= m_tiles) continue;           // the odd CTA of the last pair when the row tiles do not pair up
                if (c_blocked != nullptr) {                // blocked layout: lane = row, 16 bytes per lane, 512 contiguous bytes per warp store
                    const size_t q0i = ((size_t)m_tile * (N >> 2) + ((n_tile * PBN + c0) >> 2)) * BM + row;
                    float4 *dst = reinterpret_cast<float4 *>(c_blocked) + q0i;
                    uint2 *dst16 = reinterpret_cast<uint2 *>(c_blocked) + q0i;      // blocked_half: the same order with fp16 quads
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                        if (bias != nullptr) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + (size_t)n_tile * PBN + c0 + 4 * j));
                            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        }
                        if (blocked_half) {
                            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
                            dst16[(size_t)j * BM] = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                        } else dst[(size_t)j * BM] = v;
                    }
                    continue;
                }
                if (cell.xa != nullptr) {                  // fused LSTM cell of the wave-level decoder: 32 columns = 8 units x (i, f, g, o)
                    if (rvalid) {
                        const int col0 = n_tile * PBN + c0, u0 = col0 >> 2;
                        const float4 *tk = tk_smem ? reinterpret_cast<const float4 *>(tk_s + (size_t)tokv * tk_ld + col0)
                                                   : reinterpret_cast<const float4 *>(cell.wtok + (size_t)tokv * N + col0);
                        const float cin[8] = {ca.x, ca.y, ca.z, ca.w, cb4.x, cb4.y, cb4.z, cb4.w};
                        if (c0 + 32 * NGRP < PBN) {        // the next chunk's state, in flight under this chunk's math
                            ca = __ldg(reinterpret_cast<const float4 *>(cell.c_in + src * 128 + u0 + 8 * NGRP));
                            cb4 = __ldg(reinterpret_cast<const float4 *>(cell.c_in + src * 128 + u0 + 8 * NGRP + 4));
                        }
                        float cn[8], hn[8];
                        // two units per call: K3's packed f32x2 cell math (cell_math.cuh), 7 MUFU per LSTM unit instead of 10
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            using namespace cellmath;
                            const float4 ta = tk_smem ? tk[u] : __ldg(tk + u), tb = tk_smem ? tk[u + 1] : __ldg(tk + u + 1);
                            const f32x2 z0 = add2(pk(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 4])), pk(ta.x, tb.x));
                            const f32x2 z1 = add2(pk(__uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 5])), pk(ta.y, tb.y));
                            const f32x2 z2 = add2(pk(__uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 6])), pk(ta.z, tb.z));
                            const f32x2 z3 = add2(pk(__uint_as_float(r[4 * u + 3]), __uint_as_float(r[4 * u + 7])), pk(ta.w, tb.w));
                            f32x2 c2 = pk(cin[u], cin[u + 1]), h2;
                            if (cell.gru) {          // columns: z gate, r gate, candidate input part, candidate recurrent part
                                gru_pointwise2(z0, z1, z2, z3, c2, h2);
                                c2 = h2;
                            } else {
                                f32x2 cnew;
                                lstm_pointwise2(z0, z1, z2, z3, c2, cnew, h2);
                                c2 = cnew;
                            }
                            upk(c2, cn[u], cn[u + 1]);
                            upk(h2, hn[u], hn[u + 1]);
                        }
                        float4 *co = reinterpret_cast<float4 *>(cell.c_out + R * 128 + u0);
                        co[0] = make_float4(cn[0], cn[1], cn[2], cn[3]); co[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
                        float4 *ho = reinterpret_cast<float4 *>(cell.xa + R * (cell.xa_ld ? cell.xa_ld : 384) + u0);
                        ho[0] = make_float4(hn[0], hn[1], hn[2], hn[3]); ho[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                        if (cell.h_hi != nullptr) {          // fp16 hi / lo planes of h for the query GEMM
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __half2 hh = __floats2half2_rn(hn[2 * i], hn[2 * i + 1]);
                                const float2 hf = __half22float2(hh);
                                const __half2 ll = __floats2half2_rn(hn[2 * i] - hf.x, hn[2 * i + 1] - hf.y);
                                hi[i] = *reinterpret_cast<const uint32_t *>(&hh);
                                lo[i] = *reinterpret_cast<const uint32_t *>(&ll);
                            }
                            const long long hl = cell.h_ld ? cell.h_ld : 128;
                            *reinterpret_cast<uint4 *>(cell.h_hi + R * hl + u0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4 *>(cell.h_lo + R * hl + u0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    continue;
                }
                unsigned char *cb = cstage + (NGRP == 2 ? grp : (chunk_ctr & 1)) * CSTAGE_BYTES;
                if (et == 0) {                             // the store that last used cb has finished reading it
                    if (NGRP == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    if (bias != nullptr) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + (size_t)n_tile * PBN + c0 + 4 * j));
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                    }
                    *reinterpret_cast<float4 *>(cb + row * 128 + ((j ^ (row & 7)) << 4)) = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                if (et == 0) {
                    tma_store_2d(&map_c, cb, n_tile * PBN + c0, (int)(m_tile * BM));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_pair();                   // nothing may still be on its way to a CTA that exits
    if (warp == 1) {
        if (SM2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = BK columns x box_rows rows, 128-byte swizzle.
inline int make_map(CUtensorMap *map, const void *base, long long rows, int cols, int box_rows, int box_cols = BK, bool f16 = false,
                    long long ld = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(RVB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : cols) * (f16 ? 2 : 4)};     // row pitch in bytes
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RVB_OK;
}

template <int BN, int NPASS>
int launch(const float *A, const float *WhiT, const float *WloT, const float *bias, float *C, long long M, int N, int K,
           int *abort_flag, cudaStream_t stream, long long lda = 0) {
    using cfg = Cfg<BN, NPASS>;
    CUtensorMap ma, mh, ml;
    RVB_CHECK(make_map(&ma, A, M, K, BM, BK, false, lda));
    RVB_CHECK(make_map(&mh, WhiT, N, K, BN));
    RVB_CHECK(make_map(&ml, NPASS == 3 ? WloT : WhiT, N, K, BN));
    RVB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg::SMEM));
    const long long tiles = (long long)(N / BN) * ((M + BM - 1) / BM);
    if (tiles > 0x7fffffffLL) return fail(RVB_ERR_ARG, "projection too large for one launch");
    dim3 grid((unsigned)tiles);
    { ProfScope ps(KK_GEMM, stream);
      gemm_tc_kernel<BN, NPASS><<<grid, THREADS, cfg::SMEM, stream>>>(ma, mh, ml, bias, C, M, N, K, abort_flag); }
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

// F16IN: A (and W) are given as fp16 hi / lo planes ([M,K] / [N,K] row-major), K-elements per stage = 64.
template <int NPASS, bool F16IN>
int launch_persistent(const void *A, const void *Alo, const void *WhiT, const void *WloT, const float *bias, float *C,
                      long long M, int N, int K, int *abort_flag, cudaStream_t stream, long long lda = 0, bool blocked_out = false,
                      const CellEpilogue *cell_epi = nullptr, int n_out = 0, bool blocked_half = false) {
    if (n_out <= 0) n_out = N;                   // C has n_out <= N columns: W may be zero-padded up to a multiple of the tile
    float *c_blocked = blocked_out ? C : nullptr;
    CellEpilogue cell{};
    if (cell_epi != nullptr) cell = *cell_epi;
    // cluster form used for the big encoder GEMM; single-pass (reduced precision): measured 1.28 vs 1.17 ms, not used there
    constexpr int PAIRED = (F16IN && NPASS == 3) ? 2 : 0;
    using cfg = PCfg<NPASS, F16IN, 0>;
    using cfgp = PCfg<NPASS, F16IN, PAIRED>;
    CUtensorMap ma, mal, mh, ml, mc_map;
    const int bk = cfg::K_PER_STAGE;
    RVB_CHECK(make_map(&ma, A, M, K, BM, bk, F16IN, lda));
    RVB_CHECK(make_map(&mal, (NPASS == 3 && F16IN) ? Alo : A, M, K, BM, bk, F16IN, lda));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long m_tiles = (M + BM - 1) / BM;
    const long long tiles = (long long)(N / PBN) * m_tiles;
    // Cluster-of-two forms of the fp16-plane GEMM.  2 (the default for the big encoder projection, i.e. the blocked output):
    // one cta_group::2 MMA per pair -- each SM reads its A rows and HALF of W from shared memory, 8 KB per 64 cycles of
    // math, where the one-SM M128 N256 K16 MMA needs 12 KB = 96 cycles (tensor pipe <= 67 % busy, measured 62 %).
    // 1: independent MMAs, W halves TMA-multicast to both CTAs.  RVB_GEMM_PAIR = 0 / 1 / 2 overrides for A/B runs.
    static const int pair_env = getenv("RVB_GEMM_PAIR") ? atoi(getenv("RVB_GEMM_PAIR")) : -1;
    const int pair_mode = pair_env >= 0 ? pair_env : (blocked_out ? 2 : 0);
    const bool mc = PAIRED != 0 && pair_mode != 0 && tiles >= 2LL * sms;
    RVB_CHECK(make_map(&mh, WhiT, N, K, mc ? PBN / 2 : PBN, bk, F16IN));
    RVB_CHECK(make_map(&ml, NPASS == 3 ? WloT : WhiT, N, K, mc ? PBN / 2 : PBN, bk, F16IN));
    RVB_CHECK(make_map(&mc_map, C, M, n_out, BM, 32));
    if (mc) {
        auto kern = (pair_mode == 1) ? gemm_tc_persistent_kernel<NPASS, F16IN, PAIRED ? 1 : 0> : gemm_tc_persistent_kernel<NPASS, F16IN, PAIRED>;
        const size_t smem_bytes = (pair_mode == 1) ? cfg::SMEM : cfgp::SMEM;
        RVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        const long long pairs = (long long)(N / PBN) * ((m_tiles + 1) / 2);
        const unsigned clusters = (unsigned)(pairs < sms / 2 ? pairs : sms / 2);
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(2 * clusters); lc.blockDim = dim3(PTHREADS); lc.dynamicSmemBytes = smem_bytes; lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        ProfScope ps(KK_GEMM, stream);
        RVB_CUDA(cudaLaunchKernelEx(&lc, kern, ma, mal, mh, ml, mc_map, bias, M, N, K, abort_flag, c_blocked, cell, n_out, blocked_half ? 1 : 0));
    } else {
        RVB_CUDA(cudaFuncSetAttribute(gemm_tc_persistent_kernel<NPASS, F16IN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg::SMEM));
        const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
        ProfScope ps(KK_GEMM, stream);
        gemm_tc_persistent_kernel<NPASS, F16IN, 0><<<grid, PTHREADS, cfg::SMEM, stream>>>(ma, mal, mh, ml, mc_map, bias, M, N, K, abort_flag, c_blocked, cell, n_out, blocked_half ? 1 : 0);
    }
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

}  // namespace tc
}  // namespace gemm
}  // namespace rvb
