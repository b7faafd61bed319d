// Snippet builder: detected events + raw read -> padded model inputs, on the device.
//
// Replaces data_loader.prepare_snippets and its helpers for the inference outputs
// (reference data_loader.py: event table + scaler fit :74-79, trimming to the labelled range :82-87,
// raw standardisation :89-90, compute_fitting_event_ranges :29-46, convert_events_ranges_to_raw_ranges
// :48-51, slicing :96-99, pad_input_snippets :110-111).  Index logic is exact; values are computed in
// float64 and stored as float32 like the reference (sklearn accumulates the scaler moments in a
// different order, so stored values may differ by 1 float32 ulp).
// Built with -fmad=false.
#include "kernels.cuh"

namespace rvb {
namespace snip {

constexpr int MAX_RAW = 200, MAX_EV = 30, NF = 5, THREADS = 1024;

struct Scratch {
    double raw_mu, raw_sd;
    double ev_mu[NF], ev_sd[NF];
    long long first_start;      // start of the first kept event after re-anchoring (== label_start)
    long long last_len;         // modified length of the last kept event
    long long first_len;        // modified length of the first kept event
    int k0, k1;                 // kept events [k0, k1)
    int n_windows;
};

__device__ __forceinline__ double block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) t += sh[i];
    return t;
}
__device__ __forceinline__ long long block_sum_ll(long long v, long long *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    long long t = 0;
    for (int i = 0; i < THREADS / 32; ++i) t += sh[i];
    return t;
}

// feature f of event i of the UNTRIMMED table: length, mean, stdv, mean^2, delta-mean (data_loader.py:74-76)
__device__ __forceinline__ double feature(int f, int i, const int32_t *len, const double *mean, const double *stdv) {
    switch (f) {
        case 0: return (double)(uint32_t)len[i];
        case 1: return mean[i];
        case 2: return stdv[i];
        case 3: return mean[i] * mean[i];
        default: return i == 0 ? 0.0 : mean[i] - mean[i - 1];
    }
}

template <typename T>
__device__ __forceinline__ void stats_body(const T *raw, long long n, const int32_t *ev_start, const int32_t *ev_len,
                                           const double *ev_mean, const double *ev_stdv, int ne,
                                           long long lab0, long long lab1, int *P, Scratch *sc) {
    __shared__ double shd[THREADS / 32];
    __shared__ long long shl[THREADS / 32];
    __shared__ int s_k0, s_k1;
    const int tid = threadIdx.x;
    // raw moments from exact integer sums
    long long s = 0, q = 0;
    for (long long i = tid; i < n; i += THREADS) { long long v = (long long)raw[i]; s += v; q += v * v; }
    s = block_sum_ll(s, shl);
    q = block_sum_ll(q, shl);
    if (tid == 0) {
        double mu = (double)s / (double)n;
        double var = ((double)q - (double)s * mu) / (double)n;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        sc->raw_mu = mu; sc->raw_sd = (sd < 10 * 2.220446049250313e-16) ? 1.0 : sd;
    }
    // event feature moments over ALL events (two-pass)
    for (int f = 0; f < NF; ++f) {
        double a = 0.0;
        for (int i = tid; i < ne; i += THREADS) a += feature(f, i, ev_len, ev_mean, ev_stdv);
        const double mu = block_sum(a, shd) / (double)ne;
        double v = 0.0;
        for (int i = tid; i < ne; i += THREADS) { double d = feature(f, i, ev_len, ev_mean, ev_stdv) - mu; v += d * d; }
        const double var = block_sum(v, shd) / (double)ne;
        if (tid == 0) {
            double sd = sqrt(var);
            sc->ev_mu[f] = mu; sc->ev_sd[f] = (sd < 10 * 2.220446049250313e-16) ? 1.0 : sd;
        }
    }
    // kept range: start >= lab0 and end <= lab1 (events are contiguous and ordered)
    if (tid == 0) { s_k0 = ne; s_k1 = 0; }
    __syncthreads();
    for (int i = tid; i < ne; i += THREADS) {
        long long st = (long long)(uint32_t)ev_start[i], en = st + (long long)(uint32_t)ev_len[i];
        if (st >= lab0 && en <= lab1) { atomicMin(&s_k0, i); atomicMax(&s_k1, i + 1); }
    }
    __syncthreads();
    const int k0 = s_k0, k1 = s_k1;
    const int nk = max(0, k1 - k0);
    // modified lengths of the first / last kept event (data_loader.py:84-87), then inclusive cumsum P
    long long first_len = 0, last_len = 0;
    if (nk > 0) {
        first_len = (long long)(uint32_t)ev_len[k0] + ((long long)(uint32_t)ev_start[k0] - lab0);
        long long last_start = (nk == 1) ? lab0 : (long long)(uint32_t)ev_start[k1 - 1];
        last_len = lab1 - last_start;
    }
    auto mlen = [&](int j) -> long long {       // j relative to k0
        if (j == nk - 1) return last_len;
        if (j == 0) return first_len;
        return (long long)(uint32_t)ev_len[k0 + j];
    };
    __shared__ long long chunk_sum[THREADS];
    const int per = (nk + THREADS - 1) / THREADS;
    long long loc = 0;
    for (int j = tid * per; j < min(nk, (tid + 1) * per); ++j) loc += mlen(j);
    chunk_sum[tid] = loc;
    __syncthreads();
    if (tid == 0) {
        long long run = 0;
        for (int i = 0; i < THREADS; ++i) { long long t = chunk_sum[i]; chunk_sum[i] = run; run += t; }
        sc->k0 = k0; sc->k1 = k1; sc->first_len = first_len; sc->last_len = last_len; sc->first_start = lab0;
        sc->n_windows = 0x7fffffff;
    }
    __syncthreads();
    long long run = chunk_sum[tid];
    for (int j = tid * per; j < min(nk, (tid + 1) * per); ++j) { run += mlen(j); P[j] = (int)run; }
}

template <typename T>
__global__ void __launch_bounds__(THREADS) stats_kernel(const T *raw, long long n, const int32_t *ev_start, const int32_t *ev_len,
                                                        const double *ev_mean, const double *ev_stdv, int ne,
                                                        long long lab0, long long lab1, int *P, Scratch *sc) {
    stats_body<T>(raw, n, ev_start, ev_len, ev_mean, ev_stdv, ne, lab0, lab1, P, sc);
}
// Batched form: block r = read r (whole read labelled).  Reads and event tables are concatenated as rvb_event_detect takes
// and leaves them: samples of read r at read_off[r], its events at ev_off[r] (capacity offsets), count[r] of them valid.
// P and win_end use the event offsets too (a read has fewer windows than events).
template <typename T>
__global__ void __launch_bounds__(THREADS) stats_batch_kernel(const T *raw, const long long *read_off, const long long *ev_off,
                                                              const int32_t *count, const int32_t *ev_start, const int32_t *ev_len,
                                                              const double *ev_mean, const double *ev_stdv, int *P, Scratch *sc) {
    const int r = blockIdx.x;
    const long long n = read_off[r + 1] - read_off[r], e0 = ev_off[r];
    const int ne = count[r];
    if (n <= 0 || ne <= 0) {        // nothing to cut: no windows
        if (threadIdx.x == 0) { sc[r].k0 = 0; sc[r].k1 = 0; sc[r].n_windows = 0; }
        return;
    }
    stats_body<T>(raw + read_off[r], n, ev_start + e0, ev_len + e0, ev_mean + e0, ev_stdv + e0, ne, 0, n, P + e0, sc + r);
}

// compute_fitting_event_ranges (data_loader.py:29-46): window w starts at event w*stride and ends at the
// first event whose cumulative length (relative to the window start) exceeds MAX_RAW.
__device__ __forceinline__ void window_body(const int *P, Scratch *sc, int stride, int max_w, int *win_end, int w) {
    if (w >= max_w) return;
    const int nk = max(0, sc->k1 - sc->k0);
    const int first = w * stride;
    bool valid = first < nk;
    int end = 0;
    if (valid) {
        const long long base = (w == 0) ? 0 : (long long)P[first - 1];
        int lo = 0, hi = nk;                     // first j with P[j] - base > MAX_RAW
        while (lo < hi) { int mid = (lo + hi) >> 1; if ((long long)P[mid] - base > MAX_RAW) hi = mid; else lo = mid + 1; }
        end = lo;
        if (end >= nk || end == 0) valid = false;     // argmax of an all-False mask is 0 -> break (:40-41)
    }
    // the loop also stops after window w-1 when (w-1)*stride + stride - 1 >= nk (:43-44)
    if (valid && w > 0 && (long long)(w - 1) * stride + stride - 1 >= nk) valid = false;
    if (valid) win_end[w] = end;
    else atomicMin(&sc->n_windows, w);
}
__global__ void window_kernel(const int *P, Scratch *sc, int stride, int max_w, int *win_end) {
    window_body(P, sc, stride, max_w, win_end, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void window_batch_kernel(const int *P, Scratch *sc, const long long *ev_off, const int32_t *count, int stride, int *win_end) {
    const int r = blockIdx.y;
    const int ne = count[r];
    if (ne <= 0) return;
    const int max_w = (ne + stride - 1) / stride;
    window_body(P + ev_off[r], sc + r, stride, max_w, win_end + ev_off[r], blockIdx.x * blockDim.x + threadIdx.x);
}
// first snippet of every read = exclusive prefix sum of the reads' window counts (one block; n_reads is a few thousand at most)
__global__ void __launch_bounds__(THREADS) offsets_kernel(Scratch *sc, const int32_t *count, int stride, int n_reads, long long *snip_off) {
    __shared__ long long part[THREADS];
    const int tid = threadIdx.x;
    const int per = (n_reads + THREADS - 1) / THREADS;
    auto nw_of = [&](int r) -> long long {
        const int ne = count[r];
        const int max_w = ne > 0 ? (ne + stride - 1) / stride : 0;
        const int nw = min(sc[r].n_windows, max_w);      // n_windows stays INT_MAX when no window was invalid
        if (ne > 0) sc[r].n_windows = nw;
        return nw;
    };
    long long loc = 0;
    for (int r = tid * per; r < min(n_reads, (tid + 1) * per); ++r) loc += nw_of(r);
    part[tid] = loc;
    __syncthreads();
    if (tid == 0) {
        long long run = 0;
        for (int i = 0; i < THREADS; ++i) { const long long t = part[i]; part[i] = run; run += t; }
        snip_off[n_reads] = run;
    }
    __syncthreads();
    long long run = part[tid];
    for (int r = tid * per; r < min(n_reads, (tid + 1) * per); ++r) { snip_off[r] = run; run += sc[r].n_windows; }
}

template <typename T>
__device__ __forceinline__ void fill_body(const T *raw, long long n, const int32_t *ev_start, const int32_t *ev_len, const double *ev_mean,
                                          const double *ev_stdv, const int *win_end, const Scratch *sc, int stride, int max_w,
                                          float *raw_out, float *ev_out, int32_t *ranges_out, int w, long long ow) {
    // w: window of this read; ow: output slot (== w for one read, the global snippet index in a batch)
    const int nw = min(sc->n_windows, max_w);
    if (w >= nw) return;
    const int k0 = sc->k0, nk = sc->k1 - sc->k0;
    const int first = w * stride, end = win_end[w];
    auto mstart = [&](int j) -> long long { return j == 0 ? sc->first_start : (long long)(uint32_t)ev_start[k0 + j]; };
    const long long r0 = mstart(first), r1 = mstart(end - 1);      // raw span excludes the last event (:48-51)
    if (ranges_out != nullptr && threadIdx.x == 0) { ranges_out[2 * ow] = (int32_t)r0; ranges_out[2 * ow + 1] = (int32_t)r1; }
    if (raw_out != nullptr)
        for (int i = threadIdx.x; i < MAX_RAW; i += blockDim.x) {
            const long long idx = r0 + i;
            float v = 0.0f;
            if (idx < r1 && idx >= 0 && idx < n) v = (float)(((double)raw[idx] - sc->raw_mu) / sc->raw_sd);
            raw_out[(size_t)ow * MAX_RAW + i] = v;
        }
    for (int i = threadIdx.x; i < MAX_EV * NF; i += blockDim.x) {
        const int e = i / NF, f = i % NF, j = first + e;
        float v = 0.0f;
        if (j < end) {
            double x;
            if (f == 0) x = (j == nk - 1) ? (double)sc->last_len : (j == 0 ? (double)sc->first_len : (double)(uint32_t)ev_len[k0 + j]);
            else x = feature(f, k0 + j, ev_len, ev_mean, ev_stdv);
            v = (float)((x - sc->ev_mu[f]) / sc->ev_sd[f]);
        }
        ev_out[(size_t)ow * MAX_EV * NF + i] = v;
    }
}
template <typename T>
__global__ void fill_kernel(const T *raw, long long n, const int32_t *ev_start, const int32_t *ev_len, const double *ev_mean,
                            const double *ev_stdv, const int *win_end, const Scratch *sc, int stride, int max_w,
                            float *raw_out, float *ev_out, int32_t *ranges_out) {
    fill_body<T>(raw, n, ev_start, ev_len, ev_mean, ev_stdv, win_end, sc, stride, max_w, raw_out, ev_out, ranges_out, blockIdx.x, blockIdx.x);
}
// block (w, r): window w of read r, written at snippet snip_off[r] + w; snippets at or beyond `cap` are not written
template <typename T>
__global__ void fill_batch_kernel(const T *raw, const long long *read_off, const long long *ev_off, const int32_t *count,
                                  const int32_t *ev_start, const int32_t *ev_len, const double *ev_mean, const double *ev_stdv,
                                  const int *win_end, const Scratch *sc, const long long *snip_off, int stride, long long cap,
                                  float *raw_out, float *ev_out, int32_t *ranges_out) {
    const int r = blockIdx.y, w = blockIdx.x;
    const int ne = count[r];
    if (ne <= 0 || w >= sc[r].n_windows) return;
    const long long o = snip_off[r] + w;
    if (o >= cap) return;
    const long long e0 = ev_off[r];
    fill_body<T>(raw + read_off[r], read_off[r + 1] - read_off[r], ev_start + e0, ev_len + e0, ev_mean + e0, ev_stdv + e0, win_end + e0,
                 sc + r, stride, (ne + stride - 1) / stride, raw_out, ev_out, ranges_out, w, o);
}

}  // namespace snip
}  // namespace rvb

using namespace rvb;

extern "C" int rvb_build_snippets(const void *d_signal, int sample_bytes, int64_t n_samples, const int32_t *d_ev_start,
                                  const int32_t *d_ev_length, const double *d_ev_mean, const double *d_ev_stdv,
                                  int32_t n_events, int64_t label_start, int64_t label_end, int32_t stride,
                                  float *d_raw_snips, float *d_event_snips, int32_t max_snippets, int32_t *h_n_snippets,
                                  int32_t *d_raw_ranges, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h_n_snippets) return fail(RVB_ERR_ARG, "build_snippets: null count output");
    *h_n_snippets = 0;
    if (sample_bytes != 2 && sample_bytes != 4) return fail(RVB_ERR_ARG, "sample_bytes must be 2 or 4");
    if (stride < 1 || n_samples < 0 || n_events < 0 || max_snippets < 0) return fail(RVB_ERR_ARG, "build_snippets: bad argument");
    if (n_events == 0 || n_samples == 0) return RVB_OK;
    const int max_w = (n_events + stride - 1) / stride;
    char *scratch = nullptr;
    const size_t bytes = sizeof(snip::Scratch) + 256 + sizeof(int) * ((size_t)n_events + (size_t)max_w + 8);
    RVB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), bytes, stream));
    snip::Scratch *sc = reinterpret_cast<snip::Scratch *>(scratch);
    int *P = reinterpret_cast<int *>(scratch + ((sizeof(snip::Scratch) + 255) & ~size_t(255)));
    int *win_end = P + n_events;
    int st = RVB_OK;
    {
        ProfScope ps(KK_OTHER, stream);
        if (sample_bytes == 4)
            snip::stats_kernel<int32_t><<<1, snip::THREADS, 0, stream>>>(reinterpret_cast<const int32_t *>(d_signal), n_samples, d_ev_start,
                                                                         d_ev_length, d_ev_mean, d_ev_stdv, n_events, label_start, label_end, P, sc);
        else
            snip::stats_kernel<int16_t><<<1, snip::THREADS, 0, stream>>>(reinterpret_cast<const int16_t *>(d_signal), n_samples, d_ev_start,
                                                                         d_ev_length, d_ev_mean, d_ev_stdv, n_events, label_start, label_end, P, sc);
        snip::window_kernel<<<(max_w + 255) / 256, 256, 0, stream>>>(P, sc, stride, max_w, win_end);
        const int fill_w = max_w < max_snippets ? max_w : max_snippets;
        if (fill_w > 0) {
            if (sample_bytes == 4)
                snip::fill_kernel<int32_t><<<fill_w, 128, 0, stream>>>(reinterpret_cast<const int32_t *>(d_signal), n_samples, d_ev_start, d_ev_length,
                                                                       d_ev_mean, d_ev_stdv, win_end, sc, stride, fill_w, d_raw_snips, d_event_snips, d_raw_ranges);
            else
                snip::fill_kernel<int16_t><<<fill_w, 128, 0, stream>>>(reinterpret_cast<const int16_t *>(d_signal), n_samples, d_ev_start, d_ev_length,
                                                                       d_ev_mean, d_ev_stdv, win_end, sc, stride, fill_w, d_raw_snips, d_event_snips, d_raw_ranges);
        }
        count_launch(3);
    }
    cudaError_t e = cudaGetLastError();
    int nw = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&nw, &sc->n_windows, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFreeAsync(scratch, stream);
    if (e != cudaSuccess) st = fail(RVB_ERR_CUDA, "build_snippets: %s", cudaGetErrorString(e));
    if (st != RVB_OK) return st;
    if (nw > max_w) nw = max_w;                       // no window was invalid
    if (nw > max_snippets) return fail(RVB_ERR_OVERFLOW, "build_snippets: %d snippets > capacity %d", nw, max_snippets);
    *h_n_snippets = nw;
    return RVB_OK;
}

extern "C" int rvb_build_snippets_batch(const void *d_signal, int sample_bytes, const int64_t *h_read_offsets, int32_t n_reads,
                                        const int64_t *h_event_offsets, const int32_t *d_ev_start, const int32_t *d_ev_length,
                                        const double *d_ev_mean, const double *d_ev_stdv, const int32_t *d_counts,
                                        int32_t stride, float *d_raw_snips, float *d_event_snips, int64_t max_snippets,
                                        int64_t *d_snippet_offsets, int32_t *d_raw_ranges, int64_t *h_n_snippets, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h_n_snippets) return fail(RVB_ERR_ARG, "build_snippets_batch: null count output");
    *h_n_snippets = 0;
    if (sample_bytes != 2 && sample_bytes != 4) return fail(RVB_ERR_ARG, "sample_bytes must be 2 or 4");
    if (stride < 1 || n_reads < 0 || max_snippets < 0 || !h_read_offsets || !h_event_offsets || !d_event_snips)
        return fail(RVB_ERR_ARG, "build_snippets_batch: bad argument");
    if (n_reads == 0) return RVB_OK;
    if (n_reads > 65535) return fail(RVB_ERR_ARG, "build_snippets_batch: at most 65535 reads per call (%d)", n_reads);
    long long max_events = 0;
    for (int r = 0; r < n_reads; ++r) {
        if (h_read_offsets[r + 1] < h_read_offsets[r] || h_event_offsets[r + 1] < h_event_offsets[r])
            return fail(RVB_ERR_ARG, "build_snippets_batch: offsets must not decrease");
        max_events = std::max<long long>(max_events, h_event_offsets[r + 1] - h_event_offsets[r]);
    }
    const long long ev_cap = h_event_offsets[n_reads];
    if (ev_cap == 0) return RVB_OK;
    // scratch: per-read statistics | read offsets | event offsets | snippet offsets | P | win_end
    const size_t off_bytes = sizeof(long long) * (size_t)(n_reads + 1);
    const size_t sc_bytes = (sizeof(snip::Scratch) * (size_t)n_reads + 255) & ~size_t(255);
    const size_t bytes = sc_bytes + 3 * ((off_bytes + 255) & ~size_t(255)) + 2 * sizeof(int) * (size_t)ev_cap + 256;
    char *scratch = nullptr;
    RVB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), bytes, stream));
    snip::Scratch *sc = reinterpret_cast<snip::Scratch *>(scratch);
    long long *d_roff = reinterpret_cast<long long *>(scratch + sc_bytes);
    long long *d_eoff = reinterpret_cast<long long *>(reinterpret_cast<char *>(d_roff) + ((off_bytes + 255) & ~size_t(255)));
    long long *d_soff = reinterpret_cast<long long *>(reinterpret_cast<char *>(d_eoff) + ((off_bytes + 255) & ~size_t(255)));
    int *P = reinterpret_cast<int *>(reinterpret_cast<char *>(d_soff) + ((off_bytes + 255) & ~size_t(255)));
    int *win_end = P + ev_cap;
    cudaError_t e = cudaMemcpyAsync(d_roff, h_read_offsets, off_bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_eoff, h_event_offsets, off_bytes, cudaMemcpyHostToDevice, stream);
    long long total = 0;
    if (e == cudaSuccess) {
        ProfScope ps(KK_OTHER, stream);
        const int max_w = (int)((max_events + stride - 1) / stride);      // events per read are bounded by the capacity
        const dim3 wgrid((unsigned)((max_w + 255) / 256), (unsigned)n_reads), fgrid((unsigned)max_w, (unsigned)n_reads);
        if (sample_bytes == 4) {
            const int32_t *sig = reinterpret_cast<const int32_t *>(d_signal);
            snip::stats_batch_kernel<int32_t><<<n_reads, snip::THREADS, 0, stream>>>(sig, d_roff, d_eoff, d_counts, d_ev_start, d_ev_length, d_ev_mean, d_ev_stdv, P, sc);
            snip::window_batch_kernel<<<wgrid, 256, 0, stream>>>(P, sc, d_eoff, d_counts, stride, win_end);
            snip::offsets_kernel<<<1, snip::THREADS, 0, stream>>>(sc, d_counts, stride, n_reads, d_soff);
            snip::fill_batch_kernel<int32_t><<<fgrid, 128, 0, stream>>>(sig, d_roff, d_eoff, d_counts, d_ev_start, d_ev_length, d_ev_mean, d_ev_stdv, win_end,
                                                                        sc, d_soff, stride, max_snippets, d_raw_snips, d_event_snips, d_raw_ranges);
        } else {
            const int16_t *sig = reinterpret_cast<const int16_t *>(d_signal);
            snip::stats_batch_kernel<int16_t><<<n_reads, snip::THREADS, 0, stream>>>(sig, d_roff, d_eoff, d_counts, d_ev_start, d_ev_length, d_ev_mean, d_ev_stdv, P, sc);
            snip::window_batch_kernel<<<wgrid, 256, 0, stream>>>(P, sc, d_eoff, d_counts, stride, win_end);
            snip::offsets_kernel<<<1, snip::THREADS, 0, stream>>>(sc, d_counts, stride, n_reads, d_soff);
            snip::fill_batch_kernel<int16_t><<<fgrid, 128, 0, stream>>>(sig, d_roff, d_eoff, d_counts, d_ev_start, d_ev_length, d_ev_mean, d_ev_stdv, win_end,
                                                                        sc, d_soff, stride, max_snippets, d_raw_snips, d_event_snips, d_raw_ranges);
        }
        count_launch(4);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && d_snippet_offsets != nullptr)
        e = cudaMemcpyAsync(d_snippet_offsets, d_soff, off_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_soff + n_reads, sizeof(long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);          // the ONE round trip of the batch
    cudaFreeAsync(scratch, stream);
    if (e != cudaSuccess) return fail(RVB_ERR_CUDA, "build_snippets_batch: %s", cudaGetErrorString(e));
    if (total > max_snippets) return fail(RVB_ERR_OVERFLOW, "build_snippets_batch: %lld snippets > capacity %lld", total, (long long)max_snippets);
    *h_n_snippets = total;
    return RVB_OK;
}
