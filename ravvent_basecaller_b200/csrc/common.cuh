// Shared helpers for libravvent_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/ravvent_b200.h"

namespace rvb {

// ---- thread-local error string + launch counter ---------------------------
inline char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events on the launching stream ----
enum KernelKind { KK_EVENT = 0, KK_GEMM = 1, KK_REC = 2, KK_DECODER = 3, KK_OTHER = 4, KK_ATTENTION = 5, KK_COUNT = 6 };
void prof_record(int kind, cudaStream_t stream, bool begin);
extern std::atomic<int> g_prof_on;
inline int &prof_nest() { static thread_local int n = 0; return n; }
// Only the outermost scope records: a multi-kernel stage (decoder_wave) owns the time of the GEMMs it launches.
struct ProfScope {
    int kind; cudaStream_t stream; bool on;
    ProfScope(int k, cudaStream_t s) : kind(k), stream(s), on(g_prof_on.load(std::memory_order_relaxed) != 0 && prof_nest() == 0) {
        ++prof_nest();
        if (on) prof_record(kind, stream, true);
    }
    ~ProfScope() { --prof_nest(); if (on) prof_record(kind, stream, false); }
};

#define RVB_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return rvb::fail(RVB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
                             cudaGetErrorString(_e));                                        \
    } while (0)

#define RVB_CHECK(expr)                    \
    do {                                   \
        int _s = (expr);                   \
        if (_s != RVB_OK) return _s;       \
    } while (0)

#define RVB_LAUNCH_CHECK()                                                                   \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess)                                                               \
            return rvb::fail(RVB_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,  \
                             cudaGetErrorString(_e));                                        \
    } while (0)

// ---- model constants (the kernels are specialised for the reference's shipped
//      configuration: enc_units = dec_units = 128, vocab 7) -------------------
constexpr int UNITS = 128;      // enc_units == dec_units
constexpr int GATES = 4 * UNITS;
constexpr int VOCAB = 7;
constexpr int ENC_OUT = 2 * UNITS;
constexpr int TOKEN_END = 1;    // '^'  data_loader.py:21
constexpr int TOKEN_START = 2;  // '$'

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh with full fp32 accuracy (the fp32-parity path must not use tanh.approx)
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace rvb
