// K1 -- two-window t-statistic event detection for a ragged batch of reads.
//
// Replaces EventDetector.run/_add_sample/_compute_tstat/_detect_peak/_create_event
// (reference event_detection/event_detector.py:75-210).  Compile with -fmad=false:
// every float64 operation of the reference is a separately rounded IEEE op.
//
// Parallel decomposition (the reference is a strictly sequential per-sample loop):
//   * a read is cut into chunks of CH samples (one CTA each), a chunk into
//     sub-segments of SUB samples (one thread each);
//   * t-statistics are data-parallel given window sums (exact integers), and are
//     staged once per sample in shared memory (this is the float64-heavy part);
//   * the two coupled peak detectors are data-dependent state machines.  Every
//     sub-segment runs them speculatively from the reset state, `warmup` samples
//     early; the detectors re-synchronise within a few events;
//   * thread 0 then walks the sub-segments in order with the TRUE incoming state
//     (for chunk > 0: published by the previous chunk of the read through a
//     ticket-ordered flag chain, decoupled-look-back style) and accepts a
//     speculative result only if its start state equals the true state
//     bit-for-bit, otherwise re-runs that sub-segment from the true state.  The
//     result is therefore exactly the sequential one, for any `warmup`.
//   * accepted boundaries -> events; per-event sums are exact int64 range sums.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace rvb {
namespace ed {

constexpr int CH = 2048;
constexpr int SUB = 16;        // measured: 8 (all 8 warps busy, 56 instead of 64 steps) is slower, 1.32 vs 0.98 ms
constexpr int NSUB = CH / SUB;
constexpr int HMAX = 256;
constexpr int WMAX = 32;
constexpr int THREADS = 256;
constexpr int RAW_CAP = CH + HMAX + 2 * WMAX;
constexpr int TS_CAP = (CH + HMAX) + (CH + HMAX) / SUB + 2;
// The speculative walks read the staged t-statistics with a stride of SUB doubles (= 128 bytes) between the threads of a
// warp: unpadded, all 32 lanes hit the same bank pair (ncu: 118 M shared-memory bank conflicts per launch, the whole
// kernel's bottleneck).  One padding double per SUB entries spreads a warp's reads over 16 bank pairs.
__device__ __forceinline__ int tsi(int j) { return j + (j >> 4); }
static_assert(SUB == 16, "tsi() pads one double per 16 entries");
constexpr double FLT_MIN_D = 1.17549435e-38;   // event_detector.py:10
constexpr double FLT_MAX_D = 3.40282347e+38;   // event_detector.py:11

struct Det {
    double val;
    uint32_t masked_to;
    int32_t pos;
    int32_t valid;
    int32_t pad_;
};
struct Pair {
    Det s, l;
};
struct Chain {
    Pair st;
    long long m_st;      // absolute prefix index behind the reference's evt_st_sum (-1: unwritten slot == 0.0)
    uint32_t ev_st;      // reference evt_st
    int32_t ev_count;    // events of this read emitted so far
    int32_t flag;
    int32_t pad_;
};

struct Params {
    const void *signal;
    const long long *read_off;
    const long long *ev_off;
    const int32_t *chunk_read;
    const int32_t *chunk_idx;
    const int32_t *chunk_prev;   // ticket of the previous chunk of the same read (-1 for the first)
    int n_chunks;
    int w1, w2;
    double thr1, thr2, ph;
    int H;
    int32_t *ev_start, *ev_length;
    double *ev_mean, *ev_stdv;
    int32_t *ev_count;
    Chain *chain;
    int *ticket;
    int *status;
};

__device__ __forceinline__ void det_reset(Det &d) {
    d.val = FLT_MAX_D; d.masked_to = 0u; d.pos = -1; d.valid = 0;
}

// One update of a detector (event_detector.py:149-187).  `mask_target` is the long
// detector when d is "the short one" (the reference tests window-length equality).
__device__ __forceinline__ bool det_step(Det &d, Det *mask_target, double v, uint32_t mid, int w,
                                         double thr, double ph) {
    if (d.masked_to >= mid) return false;
    if (d.pos == -1) {
        if (v < d.val) d.val = v;
        else if (__dsub_rn(v, d.val) > ph) { d.val = v; d.pos = (int32_t)mid; }
        return false;
    }
    if (v > d.val) { d.val = v; d.pos = (int32_t)mid; }
    if (mask_target != nullptr && d.val > thr) {
        mask_target->masked_to = (uint32_t)((long long)d.pos + (long long)w);
        mask_target->pos = -1; mask_target->val = FLT_MAX_D; mask_target->valid = 0;
    }
    if (__dsub_rn(d.val, v) > ph && d.val > thr) d.valid = 1;
    if (d.valid && (double)((long long)mid - (long long)d.pos) > (double)w * 0.5) {
        d.pos = -1; d.val = v; d.valid = 0;
        return true;
    }
    return false;
}

__device__ __forceinline__ bool pair_step(Pair &p, double t1, double t2, uint32_t mid, const Params &q) {
    bool f1 = det_step(p.s, &p.l, t1, mid, q.w1, q.thr1, q.ph);
    bool f2 = det_step(p.l, (q.w1 == q.w2) ? &p.l : nullptr, t2, mid, q.w2, q.thr2, q.ph);
    return f1 || f2;
}

// A long-detector mask that lies behind every future buf_mid can never matter again.
__device__ __forceinline__ void pair_normalise(Pair &p, long long n, int w2) {
    uint32_t next_mid = (uint32_t)(n + 1 - w2);
    if (p.l.masked_to < next_mid) p.l.masked_to = 0u;
    if (p.s.masked_to < next_mid) p.s.masked_to = 0u;
}

__device__ __forceinline__ bool det_equal(const Det &a, const Det &b) {
    return __double_as_longlong(a.val) == __double_as_longlong(b.val) && a.masked_to == b.masked_to &&
           a.pos == b.pos && a.valid == b.valid;
}
__device__ __forceinline__ bool pair_equal(const Pair &a, const Pair &b) {
    return det_equal(a.s, b.s) && det_equal(a.l, b.l);
}

__device__ __forceinline__ double tstat_from_sums(long long s1, long long q1, long long s2, long long q2, int w) {
    double wf = (double)w;
    double m1 = __ddiv_rn(__ll2double_rn(s1), wf), m2 = __ddiv_rn(__ll2double_rn(s2), wf);
    double var = __dsub_rn(__ddiv_rn(__ll2double_rn(q1), wf), __dmul_rn(m1, m1));
    var = __dadd_rn(var, __ddiv_rn(__ll2double_rn(q2), wf));
    var = __dsub_rn(var, __dmul_rn(m2, m2));
    if (FLT_MIN_D > var) var = FLT_MIN_D;
    return __ddiv_rn(fabs(__dsub_rn(m2, m1)), __dsqrt_rn(__ddiv_rn(var, wf)));
}

// Index whose prefix sum the reference's ring holds in `slot` after n samples (-1: never written).
__device__ __forceinline__ long long ring_latest(long long n, long long slot, int buf) {
    long long r = (n - slot) % buf;
    if (r < 0) r += buf;
    return n - r;
}

template <typename T>
__device__ void range_sums(const T *raw, long long lo, long long hi, long long &s, long long &q) {
    s = 0; q = 0;
    for (long long j = lo; j < hi; ++j) {
        long long v = (long long)raw[j];
        s += v; q += v * v;
    }
}

// value(ma) - value(mb) for prefix sums, value(-1) == 0 (unwritten ring slot).
template <typename T>
__device__ void prefix_diff(const T *raw, long long ma, long long mb, long long &s, long long &q) {
    long long a = ma < 0 ? 0 : ma, b = mb < 0 ? 0 : mb;   // prefix(0) == 0 too
    if (a >= b) range_sums(raw, b, a, s, q);
    else { range_sums(raw, a, b, s, q); s = -s; q = -q; }
}

// Slow exact path: emulates the ring / u32 index arithmetic (warm-up of a read, or w > w2).
template <typename T>
__device__ double tstat_general(const T *raw, long long n, int w, int w2) {
    if ((unsigned long long)(n + 1) <= (unsigned long long)(2 * w) || w < 2) return 0.0;
    int buf = 1 + 2 * w2;
    uint32_t mid = (uint32_t)(n - w2);
    long long mi = ring_latest(n, mid % (uint32_t)buf, buf);
    long long ms = ring_latest(n, (uint32_t)(mid - (uint32_t)w) % (uint32_t)buf, buf);
    long long me = ring_latest(n, (uint32_t)(mid + (uint32_t)w) % (uint32_t)buf, buf);
    long long s1, q1, s2, q2;
    prefix_diff(raw, mi, ms, s1, q1);
    prefix_diff(raw, me, mi, s2, q2);
    return tstat_from_sums(s1, q1, s2, q2, w);
}

template <typename T>
__global__ void __launch_bounds__(THREADS) event_detect_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ts1 = reinterpret_cast<double *>(smem_raw);
    double *ts2 = ts1 + TS_CAP;
    Pair *sub_start = reinterpret_cast<Pair *>(ts2 + TS_CAP);
    Pair *sub_end = sub_start + NSUB;
    int *raw_s = reinterpret_cast<int *>(sub_end + NSUB);
    int *fire_n = raw_s + RAW_CAP;            // accepted fires: n relative to chunk start a (1..CH)
    uint32_t *sub_mask = reinterpret_cast<uint32_t *>(fire_n + CH);
    __shared__ int s_chunk, s_nfire, s_count0, s_abort, s_inner_bad, s_parallel_emit;
    __shared__ int sub_pre[NSUB];
    __shared__ uint32_t s_evst0;
    __shared__ long long s_mst0;

    const int tid = threadIdx.x;
    if (tid == 0) { s_chunk = atomicAdd(p.ticket, 1); s_abort = 0; }
    __syncthreads();
    const int c = s_chunk;
    if (c >= p.n_chunks) return;
    const int r = p.chunk_read[c];
    const int ci = p.chunk_idx[c];
    const long long base = p.read_off[r];
    const long long N = p.read_off[r + 1] - base;
    const T *raw = reinterpret_cast<const T *>(p.signal) + base;
    const int w1 = p.w1, w2 = p.w2, buf = 1 + 2 * w2;
    const long long a = (long long)ci * CH;                  // this chunk consumes n in (a, b]
    const long long b = min(N, a + (long long)CH);
    const int Hh = (ci == 0) ? 0 : p.H;                      // a >= CH >= H for ci > 0
    const long long ts_n0 = a - Hh;                          // ts index j <-> n = ts_n0 + 1 + j
    const int n_ts = (int)(b - ts_n0);
    const long long raw_lo = max(0LL, ts_n0 + 1 - 2 * w2);
    const int n_raw = (int)(b - raw_lo);

    if (sizeof(T) == 4) {
        // 128-bit coalesced loads: walk 16-byte aligned quads of the global array that cover [raw_lo, raw_lo + n_raw)
        const int *g = reinterpret_cast<const int *>(p.signal);
        const long long g_lo = base + raw_lo, g_hi = g_lo + n_raw;
        const long long q_lo = g_lo & ~3LL;
        for (long long qd = q_lo + 4LL * tid; qd < g_hi; qd += 4LL * THREADS) {
            if (qd >= g_lo && qd + 4 <= g_hi) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(g + qd));
                int *d = raw_s + (qd - g_lo);
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            } else {
                for (int k = 0; k < 4; ++k)
                    if (qd + k >= g_lo && qd + k < g_hi) raw_s[qd + k - g_lo] = __ldg(g + qd + k);
            }
        }
    } else {
        for (int j = tid; j < n_raw; j += THREADS) raw_s[j] = (int)raw[raw_lo + j];
    }
    __syncthreads();

    // ---- phase T: t-statistics, one per (sample, window) ------------------------------------
    const bool fast_ok = (w1 <= w2);
    for (int j = tid; j < n_ts; j += THREADS) {
        long long n = ts_n0 + 1 + j;
        double t1, t2;
        if (fast_ok && n >= 2 * w2) {
            // no index wrap: sum1 over [mid-w, mid), sum2 over [mid, mid+w), mid = n - w2
            int m0 = (int)(n - w2 - raw_lo);
            long long s1 = 0, q1 = 0, s2 = 0, q2 = 0;
            for (int k = 1; k <= w2; ++k) { long long v = raw_s[m0 - k]; s1 += v; q1 += v * v; }
            for (int k = 0; k < w2; ++k) { long long v = raw_s[m0 + k]; s2 += v; q2 += v * v; }
            t2 = (w2 < 2) ? 0.0 : tstat_from_sums(s1, q1, s2, q2, w2);
            if (w1 == w2) t1 = t2;
            else if (w1 < 2) t1 = 0.0;
            else {
                s1 = q1 = s2 = q2 = 0;
                for (int k = 1; k <= w1; ++k) { long long v = raw_s[m0 - k]; s1 += v; q1 += v * v; }
                for (int k = 0; k < w1; ++k) { long long v = raw_s[m0 + k]; s2 += v; q2 += v * v; }
                t1 = tstat_from_sums(s1, q1, s2, q2, w1);
            }
        } else {
            t1 = tstat_general(raw, n, w1, w2);
            t2 = tstat_general(raw, n, w2, w2);
        }
        ts1[tsi(j)] = t1; ts2[tsi(j)] = t2;
    }
    __syncthreads();

    // ---- phase S: speculative sub-segments ---------------------------------------------------
    const int nsub = (int)((b - a + SUB - 1) / SUB);
    if (tid < nsub) {
        long long seg0 = a + (long long)tid * SUB;           // consumes n in (seg0, seg1]
        long long seg1 = min(b, seg0 + SUB);
        long long hq = min((long long)p.H, seg0 - ts_n0);    // warm-up available in the staged range
        if (ci == 0) hq = min((long long)p.H, seg0);
        Pair st; det_reset(st.s); det_reset(st.l);
        for (long long n = seg0 - hq + 1; n <= seg0; ++n) {
            int j = (int)(n - ts_n0 - 1);
            pair_step(st, ts1[tsi(j)], ts2[tsi(j)], (uint32_t)(n - w2), p);
        }
        pair_normalise(st, seg0, w2);
        sub_start[tid] = st;
        uint32_t mask = 0;
        for (long long n = seg0 + 1; n <= seg1; ++n) {
            int j = (int)(n - ts_n0 - 1);
            if (pair_step(st, ts1[tsi(j)], ts2[tsi(j)], (uint32_t)(n - w2), p)) mask |= 1u << (int)(n - seg0 - 1);
        }
        pair_normalise(st, seg1, w2);
        sub_end[tid] = st;
        sub_mask[tid] = mask;
    }
    __syncthreads();

    // ---- phase C: exact verification against the true incoming state, chain to next chunk ----
    // Chain-independent part first, in parallel: sub-segment q is consistent with q-1 if the speculative
    // end state of q-1 equals the speculative start state of q.  If that holds for every q >= 1, the whole
    // chunk is accepted as soon as the TRUE incoming state equals sub_start[0]: the chain-critical path is
    // one compare + publish.  Anything else (first chunk of a read with its warm-up quirks, a mismatch,
    // w1 > w2) takes the exact sequential walk below.
    if (tid == 0) s_inner_bad = 0;
    __syncthreads();
    if (tid < nsub) {
        if (tid >= 1 && !pair_equal(sub_end[tid - 1], sub_start[tid])) atomicOr(&s_inner_bad, 1);
    }
    {
        // exclusive prefix count of the speculative fires: warp-shuffle scan + one pass over the per-warp totals
        __shared__ int warp_tot[THREADS / 32];
        const int lane = tid & 31, wid = tid >> 5;
        const int mine = (tid < nsub) ? __popc(sub_mask[tid]) : 0;
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        int off = 0;
        for (int w = 0; w < wid; ++w) off += warp_tot[w];
        if (tid < nsub) sub_pre[tid] = off + inc - mine;
    }
    __syncthreads();
    if (tid == 0) {
        Pair cur; uint32_t ev_st = 0; long long m_st = 0; int count = 0;
        if (ci == 0) { det_reset(cur.s); det_reset(cur.l); pair_normalise(cur, 0, w2); }
        else {
            const int cp = p.chunk_prev[c];
            volatile Chain *prev = p.chain + cp;
            unsigned spins = 0; bool ok = true;
            while (true) {
                int f;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(&prev->flag));
                if (f != 0) break;
                if (++spins > (1u << 24)) { ok = false; break; }
                __nanosleep(40);
            }
            if (!ok) { atomicExch(p.status, RVB_ERR_INTERNAL); s_abort = 1; }
            const Chain *pc = p.chain + cp;
            cur = pc->st; ev_st = pc->ev_st; m_st = pc->m_st; count = pc->ev_count;
        }
        s_evst0 = ev_st; s_mst0 = m_st; s_count0 = count;
        int nfire = 0;
        bool fast = (ci > 0) && fast_ok && !s_inner_bad && a >= (long long)buf && pair_equal(cur, sub_start[0]);
        if (fast) {
            const int total = sub_pre[nsub - 1] + __popc(sub_mask[nsub - 1]);
            if (total > 0) {
                int qf = 0; while (sub_mask[qf] == 0) ++qf;
                int ql = nsub - 1; while (sub_mask[ql] == 0) --ql;
                const long long n_first = a + (long long)qf * SUB + __ffs(sub_mask[qf]);
                const long long n_last = a + (long long)ql * SUB + (32 - __clz(sub_mask[ql]));
                const uint32_t en_first = (uint32_t)(n_first - w2) - (uint32_t)w1 + 1u;
                if ((long long)en_first - (long long)ev_st < 1) fast = false;      // would be rejected (:194-195)
                else { ev_st = (uint32_t)(n_last - w2) - (uint32_t)w1 + 1u; m_st = (long long)ev_st; }
            }
            if (fast) { cur = sub_end[nsub - 1]; nfire = total; }
        }
        s_parallel_emit = fast ? 1 : 0;
        if (!fast) {
            for (int q = 0; q < nsub; ++q) {
                long long seg0 = a + (long long)q * SUB, seg1 = min(b, seg0 + SUB);
                uint32_t mask;
                if (pair_equal(cur, sub_start[q])) { mask = sub_mask[q]; cur = sub_end[q]; }
                else {
                    mask = 0;
                    for (long long n = seg0 + 1; n <= seg1; ++n) {
                        int j = (int)(n - ts_n0 - 1);
                        if (pair_step(cur, ts1[tsi(j)], ts2[tsi(j)], (uint32_t)(n - w2), p)) mask |= 1u << (int)(n - seg0 - 1);
                    }
                    pair_normalise(cur, seg1, w2);
                }
                while (mask) {
                    int bit = __ffs(mask) - 1; mask &= mask - 1;
                    long long n = seg0 + 1 + bit;
                    uint32_t en = (uint32_t)(n - w2) - (uint32_t)w1 + 1u;       // event_detector.py:102-104
                    long long len = (long long)en - (long long)ev_st;
                    if (len < 1) continue;                                      // :194-195
                    fire_n[nfire++] = (int)(n - a);
                    ev_st = en; m_st = ring_latest(n, en % (uint32_t)buf, buf);
                    if (m_st < 0) m_st = -1;
                }
            }
        }
        s_nfire = nfire;
        Chain *me = p.chain + c;
        me->st = cur; me->ev_st = ev_st; me->m_st = m_st; me->ev_count = count + nfire;
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(&me->flag), "r"(1));
        if (b == N) p.ev_count[r] = count + nfire;
    }
    __syncthreads();
    if (s_abort) return;
    if (s_parallel_emit && tid < nsub) {           // accepted speculative fires -> ordered list, in parallel
        uint32_t mask = sub_mask[tid];
        int o = sub_pre[tid];
        while (mask) {
            int bit = __ffs(mask) - 1; mask &= mask - 1;
            fire_n[o++] = tid * SUB + 1 + bit;
        }
    }
    __syncthreads();

    // ---- phase E: event table rows (event_detector.py:189-210) --------------------------------
    const int nfire = s_nfire;
    const long long out0 = p.ev_off[r] + s_count0;
    const long long cap_end = p.ev_off[r + 1];
    if (tid == 0 && out0 + nfire > cap_end) atomicExch(p.status, RVB_ERR_OVERFLOW);
    for (int k = tid; k < nfire; k += THREADS) {
        long long n = a + fire_n[k];
        uint32_t en = (uint32_t)(n - w2) - (uint32_t)w1 + 1u;
        const bool steady = fast_ok && a >= (long long)buf;     // slot en%buf still holds S[en]: no 64-bit modulo
        long long m_en = steady ? (long long)en : ring_latest(n, en % (uint32_t)buf, buf);
        uint32_t st; long long m_st;
        if (k == 0) { st = s_evst0; m_st = s_mst0; }
        else {
            long long np = a + fire_n[k - 1];
            st = (uint32_t)(np - w2) - (uint32_t)w1 + 1u;
            m_st = steady ? (long long)st : ring_latest(np, st % (uint32_t)buf, buf);
        }
        long long len = (long long)en - (long long)st;
        long long s, q;
        prefix_diff(raw, m_en, m_st, s, q);
        double lf = (double)len;
        double mean = __ddiv_rn(__ll2double_rn(s), lf);
        double var = __dsub_rn(__ddiv_rn(__ll2double_rn(q), lf), __dmul_rn(mean, mean));
        if (FLT_MIN_D > var) var = FLT_MIN_D;
        long long o = out0 + k;
        if (o < cap_end) {
            p.ev_start[o] = (int32_t)st;
            p.ev_length[o] = (int32_t)(uint32_t)len;
            p.ev_mean[o] = mean;
            p.ev_stdv[o] = __dsqrt_rn(var);
        }
    }
}

constexpr size_t SMEM_BYTES = sizeof(double) * 2 * TS_CAP + sizeof(Pair) * 2 * NSUB + sizeof(int) * RAW_CAP +
                              sizeof(int) * CH + sizeof(uint32_t) * NSUB;

struct Layout {
    size_t ticket, chain, chunk_read, chunk_idx, chunk_prev, read_off, ev_off, total;
    long long n_chunks;
};

static Layout make_layout(const int64_t *h_read_off, int32_t n_reads) {
    Layout L{};
    long long nc = 0;
    for (int r = 0; r < n_reads; ++r) {
        long long n = h_read_off[r + 1] - h_read_off[r];
        nc += (n + CH - 1) / CH;
    }
    L.n_chunks = nc;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    L.ticket = off; off = al(off + 2 * sizeof(int));
    L.chain = off; off = al(off + sizeof(Chain) * (size_t)nc);
    L.chunk_read = off; off = al(off + sizeof(int32_t) * (size_t)nc);
    L.chunk_idx = off; off = al(off + sizeof(int32_t) * (size_t)nc);
    L.chunk_prev = off; off = al(off + sizeof(int32_t) * (size_t)nc);
    L.read_off = off; off = al(off + sizeof(int64_t) * (size_t)(n_reads + 1));
    L.ev_off = off; off = al(off + sizeof(int64_t) * (size_t)(n_reads + 1));
    L.total = off;
    return L;
}

}  // namespace ed
}  // namespace rvb

using namespace rvb;

extern "C" int rvb_event_detect_workspace_bytes(const int64_t *h_read_offsets, int32_t n_reads, size_t *bytes) {
    if (!h_read_offsets || !bytes || n_reads < 0) return fail(RVB_ERR_ARG, "event_detect_workspace_bytes: bad argument");
    *bytes = ed::make_layout(h_read_offsets, n_reads).total + 256;
    return RVB_OK;
}

extern "C" int rvb_event_detect(const void *d_signal, int sample_bytes, const int64_t *h_read_offsets,
                                int32_t n_reads, int w1, int w2, double thr1, double thr2, double peak_height,
                                const int64_t *h_event_offsets, int32_t *d_ev_start, int32_t *d_ev_length,
                                double *d_ev_mean, double *d_ev_stdv, int32_t *d_ev_count, void *d_workspace,
                                size_t workspace_bytes, int warmup, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (sample_bytes != 4 && sample_bytes != 2) return fail(RVB_ERR_ARG, "sample_bytes must be 2 or 4");
    if (w1 < 1 || w2 < 1 || w1 > ed::WMAX || w2 > ed::WMAX)
        return fail(RVB_ERR_ARG, "window lengths must be in [1,%d]", ed::WMAX);
    if (n_reads < 0 || !h_read_offsets || !h_event_offsets) return fail(RVB_ERR_ARG, "bad read/event offsets");
    // default speculative warm-up: the detectors re-synchronised within 32 samples in every case measured (0 of 1 592
    // sub-segments needed the exact re-run); shorter only costs time (the verification falls back), never exactness
    if (warmup < 0) warmup = 32;
    if (warmup > ed::HMAX) return fail(RVB_ERR_ARG, "warmup must be <= %d", ed::HMAX);
    if (n_reads == 0) return RVB_OK;
    for (int r = 0; r < n_reads; ++r)
        if (h_read_offsets[r + 1] < h_read_offsets[r] || h_event_offsets[r + 1] < h_event_offsets[r])
            return fail(RVB_ERR_ARG, "offsets must be non-decreasing");
    ed::Layout L = ed::make_layout(h_read_offsets, n_reads);
    if (workspace_bytes < L.total) return fail(RVB_ERR_ARG, "workspace too small: %zu < %zu", workspace_bytes, L.total);
    if (L.n_chunks > 0x7fffffffLL) return fail(RVB_ERR_ARG, "too many chunks");
    char *ws = reinterpret_cast<char *>(d_workspace);
    RVB_CUDA(cudaMemsetAsync(d_ev_count, 0, sizeof(int32_t) * n_reads, stream));
    if (L.n_chunks == 0) { RVB_CUDA(cudaStreamSynchronize(stream)); return RVB_OK; }
    // host-side chunk table (tiny): ticket -> (read, chunk index in read).  Tickets are handed out in table
    // order, so a chunk only ever waits on a lower ticket (forward progress).  The table is ordered by LAYER
    // (chunk 0 of every read, then chunk 1 of every read, ...): with many reads in the batch the predecessor of a
    // chunk finished a whole round of CTAs earlier, so the read-internal chain never ripples.  `chain` is indexed
    // by ticket; chunk (r,i) finds its predecessor (r,i-1) through chunk_prev.
    int32_t *h_tab = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)L.n_chunks);
    if (!h_tab) return fail(RVB_ERR_INTERNAL, "out of host memory");
    {
        long long max_nc = 0;
        for (int r = 0; r < n_reads; ++r) max_nc = std::max(max_nc, (long long)((h_read_offsets[r + 1] - h_read_offsets[r] + ed::CH - 1) / ed::CH));
        std::vector<int32_t> last_ticket((size_t)n_reads, -1);
        long long k = 0;
        for (long long i = 0; i < max_nc; ++i)
            for (int r = 0; r < n_reads; ++r) {
                const long long nc = (h_read_offsets[r + 1] - h_read_offsets[r] + ed::CH - 1) / ed::CH;
                if (i >= nc) continue;
                h_tab[k] = r; h_tab[L.n_chunks + k] = (int32_t)i; h_tab[2 * L.n_chunks + k] = last_ticket[r];
                last_ticket[r] = (int32_t)k;
                ++k;
            }
    }
    cudaError_t e = cudaMemsetAsync(ws + L.ticket, 0, L.chunk_read - L.ticket, stream);   // ticket, status, chain flags
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws + L.chunk_read, h_tab, sizeof(int32_t) * L.n_chunks, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws + L.chunk_idx, h_tab + L.n_chunks, sizeof(int32_t) * L.n_chunks, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws + L.chunk_prev, h_tab + 2 * L.n_chunks, sizeof(int32_t) * L.n_chunks, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws + L.read_off, h_read_offsets, sizeof(int64_t) * (n_reads + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws + L.ev_off, h_event_offsets, sizeof(int64_t) * (n_reads + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);     // pageable staging buffers are ours to free after this
    free(h_tab);
    if (e != cudaSuccess) return fail(RVB_ERR_CUDA, "event_detect setup: %s", cudaGetErrorString(e));

    ed::Params p{};
    p.signal = d_signal;
    p.read_off = reinterpret_cast<const long long *>(ws + L.read_off);
    p.ev_off = reinterpret_cast<const long long *>(ws + L.ev_off);
    p.chunk_read = reinterpret_cast<const int32_t *>(ws + L.chunk_read);
    p.chunk_idx = reinterpret_cast<const int32_t *>(ws + L.chunk_idx);
    p.chunk_prev = reinterpret_cast<const int32_t *>(ws + L.chunk_prev);
    p.n_chunks = (int)L.n_chunks;
    p.w1 = w1; p.w2 = w2; p.thr1 = thr1; p.thr2 = thr2; p.ph = peak_height; p.H = warmup;
    p.ev_start = d_ev_start; p.ev_length = d_ev_length; p.ev_mean = d_ev_mean; p.ev_stdv = d_ev_stdv;
    p.ev_count = d_ev_count;
    p.chain = reinterpret_cast<ed::Chain *>(ws + L.chain);
    p.ticket = reinterpret_cast<int *>(ws + L.ticket);
    p.status = p.ticket + 1;
    { ProfScope ps(KK_EVENT, stream);
    if (sample_bytes == 4) {
        RVB_CUDA(cudaFuncSetAttribute(ed::event_detect_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ed::SMEM_BYTES));
        ed::event_detect_kernel<int32_t><<<(unsigned)L.n_chunks, ed::THREADS, ed::SMEM_BYTES, stream>>>(p);
    } else {
        RVB_CUDA(cudaFuncSetAttribute(ed::event_detect_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ed::SMEM_BYTES));
        ed::event_detect_kernel<int16_t><<<(unsigned)L.n_chunks, ed::THREADS, ed::SMEM_BYTES, stream>>>(p);
    }
    }
    RVB_LAUNCH_CHECK();
    count_launch();
    int h_status = 0;
    RVB_CUDA(cudaMemcpyAsync(&h_status, p.status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    RVB_CUDA(cudaStreamSynchronize(stream));
    if (h_status == RVB_ERR_OVERFLOW) return fail(RVB_ERR_OVERFLOW, "event capacity of a read exceeded");
    if (h_status != 0) return fail(RVB_ERR_INTERNAL, "event_detect: chunk chain timed out");
    return RVB_OK;
}
