// Micro-benchmark: how long does a chain of small-N tcgen05.mma (M = 128, K = 16, kind::f16) take on one SM, as a function
// of N, of the A-operand layout (K-major vs MN-major) and of how many independent TMEM accumulators the chain alternates over?
// Used to size the attention_tc.cu MMA schedule.  Build: see tools/README or
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ravvent_basecaller_b200/csrc -o /tmp/mma_bench tools/mma_small_n_bench.cu
#include <cstdio>
#include <cstdlib>
#include "proj_gemm_tc.cuh"
namespace rvb { std::atomic<long long> g_launches{0}; }
using namespace rvb::gemm::tc;

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

constexpr int COUNT = 96;
// ELECT: the whole warp runs the loop and one elected lane issues (uniform datapath); otherwise `if (lane == 0)` (divergent)
template <int N, int MN, int ACC, int ELECT>
__global__ void __launch_bounds__(64, 1) bench(long long *out, int *abort_flag) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (128 * 1024) / 16; i += 64) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x < 32 && (ELECT || threadIdx.x == 0)) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        constexpr uint32_t idesc = make_idesc_f16(128, N) | (MN ? (1u << 15) : 0u);
        const uint64_t dk = make_desc(0), dm = desc_mn(0, 16384, 1024);
        const uint32_t k_lo = (uint32_t)dk, k_hi = (uint32_t)(dk >> 32), m_lo = (uint32_t)dm, m_hi = (uint32_t)(dm >> 32);
        uint32_t aw = (MN ? m_lo : k_lo) + (a0 >> 4), bw = k_lo + (b0 >> 4);
        const uint32_t ah = MN ? m_hi : k_hi;
        for (int rep = 0; rep < 3; ++rep) {
            asm volatile("" : "+r"(aw), "+r"(bw));
            const long long t0 = clock64();
            if (!ELECT || elect_one()) {
#pragma unroll
                for (int i = 0; i < COUNT; ++i) {
                    // walk the A operand the way attention_tc.cu does: 4 KB of fresh shared memory per MMA, 64 KB in all
                    const uint32_t aoff = MN ? (uint32_t)((i & 7) * 2048 + ((i >> 3) & 1) * 32768) : (uint32_t)((i & 3) * 32 + ((i >> 2) & 3) * 16384);
                    const uint32_t boff = (uint32_t)((i & 3) * 32 + ((i >> 2) & 3) * N * 128);
                    mma_w(tm + (uint32_t)((i % ACC) * N), aw + (aoff >> 4), ah, bw + (boff >> 4), k_hi, idesc, i >= ACC ? 1u : 0u);
                }
            }
            if (ELECT) __syncwarp();
            const long long t1 = clock64();
            if (!ELECT || elect_one()) umma_commit(&bar);
            if (ELECT) __syncwarp();
            mbar_wait(&bar, rep & 1, abort_flag);
            const long long t2 = clock64();
            if (blockIdx.x == 0 && threadIdx.x == 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

template <int N, int MN, int ACC, int ELECT>
static void run_one(long long *d, int *flag, int grid) {
    long long h[6];
    const size_t smem = 129 * 1024 + 1024;
    cudaFuncSetAttribute(bench<N, MN, ACC, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<N, MN, ACC, ELECT><<<grid, 64, smem>>>(d, flag);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("grid %3d  %s  A %s  N %3d  accumulators %d : issue %6.1f  total %6.1f\n", grid, ELECT ? "elect.sync" : "lane == 0 ", MN ? "MN-major" : "K-major ", N, ACC,
           (double)h[4] / COUNT, (double)h[5] / COUNT);
}
template <int N, int MN>
static void run_n(long long *d, int *flag, int grid) {
    run_one<N, MN, 1, 0>(d, flag, grid);
    run_one<N, MN, 1, 1>(d, flag, grid);
    run_one<N, MN, 2, 1>(d, flag, grid);
    run_one<N, MN, 3, 1>(d, flag, grid);
}

int main() {
    long long *d;
    int *flag;
    cudaMalloc(&d, 6 * sizeof(long long));
    cudaMalloc(&flag, 4);
    cudaMemset(flag, 0, 4);
    printf("%d MMAs M=128 K=16 kind::f16 per run; cycles per MMA (issue loop / until the commit is observed), third repetition\n", COUNT);
    for (int grid : {1, 148}) {
        run_n<16, 0>(d, flag, grid);
        run_n<16, 1>(d, flag, grid);
        run_n<32, 0>(d, flag, grid);
        run_n<64, 0>(d, flag, grid);
        run_n<64, 1>(d, flag, grid);
        run_n<128, 0>(d, flag, grid);
        run_n<128, 1>(d, flag, grid);
    }
    return 0;
}
