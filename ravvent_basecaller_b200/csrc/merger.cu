// Snippet -> read stitching on the GPU.
//
// Replaces Merger.merge of the reference (merger.py:146-248) together with the routine it calls,
// Biopython's pairwise2.align.localms / localds (affine-gap local alignment, first returned alignment), and the
// per-snippet preparation of ravvent_performance_evaluator.py:66-70 (tokens -> bases, beam scores -> probabilities).
//
// A read is a strictly sequential chain (snippet i is aligned against the last 25 bases of everything merged so
// far), reads are independent: ONE WARP PER READ.  Per snippet:
//   1. the 26 x 26 affine score matrices (best / gap-in-A / gap-in-B) are filled by anti-diagonals, lane = row (float64,
//      the same operations in the same order as the restated pairwise2 "fast" recurrence; this file is built with
//      -fmad=false so no a*b+c is contracted) -- 49 dependent steps instead of 625; pairwise2's trace matrix (seven rounded
//      comparisons per cell) is computed in a second pass in which the 625 cells are independent of each other, so it
//      runs 32 cells wide with full instruction-level parallelism instead of sitting in the recurrence's dependency chain;
//   2. lanes mark the admissible start cells in parallel (within 5e-4 of the best score, positive, ending on a
//      match, not a zero-score extension of another start);
//   3. lane 0 runs pairwise2's iterative back-trace with an explicit stack in shared memory (last start first; per
//      cell open-gap-in-A, match, open-gap-in-B, extend-A, extend-B; a gap in A may not follow a gap in B);
//   4. the gapped overlap is merged column by column (SingleMergerByLogits) and the read's tail is rewritten.
// Semantics, including the reference's early return when a snippet does not align after merging has begun, are
// spelled out in DESIGN.md (read stitching), which also states what is restated from where.
#include "common.cuh"

namespace rvb {
namespace mrg {

constexpr int OVL = 25;                 // merger.py:150
constexpr int DIM = OVL + 1;
constexpr int MAXCOL = 64;              // gapped alignment columns (<= 2 * OVL)
constexpr int STACK = 160;
constexpr int WARPS = 2;

struct ScoreSet { double match[4][4]; double open, extend; };

// gap_dir != 0: the entry was pushed from inside a gap walk that started at column gap_from of the alignment buffers;
// entries pushed later may start at a shorter prefix and overwrite those columns, so they are rebuilt on pop.
struct StackEntry { uint8_t n, row, col, col_gap, trace, gap_from, gap_dir, pad; };

struct WarpSmem {
    double score[DIM][DIM];
    double rowsc[DIM][DIM], colsc[DIM][DIM];   // running gap scores after each cell (inputs of the trace pass)
    uint8_t trace[DIM][DIM];
    uint32_t valid[DIM];                // bit c of valid[r]: admissible start cell
    StackEntry stack[STACK];
    uint8_t ali_a[MAXCOL], ali_b[MAXCOL];
    uint8_t seq_a[OVL], seq_b[OVL];
    float log_a[OVL], log_b[OVL];
    uint8_t out_seq[MAXCOL];
    float out_log[MAXCOL];
    int n_cols;                         // -1: no alignment
    int error;
};

constexpr uint8_t GAPC = 4;             // base codes 0..3 = A C G T

__device__ __forceinline__ int rint1000(double x) { return (int)(x * 1000.0 + 0.5); }

__device__ __forceinline__ double affine_penalty(int length, double open, double extend) {
    if (length <= 0) return 0.0;
    double p = open + extend * (double)length;
    p -= extend;
    return p;
}

// ---- 1. score matrices, lane = row - 1 ---------------------------------------------------------------------------------
__device__ double fill_matrices(WarpSmem &w, const ScoreSet &ss, int len_a, int len_b, int lane) {
    const double open = ss.open, extend = ss.extend;
    const double first_gap = affine_penalty(1, open, extend);
    const int row = lane + 1;
    const bool live = row <= len_a;
    const int a = live ? w.seq_a[lane] : 0;
    double row_score = affine_penalty(row, 2.0 * open, extend);
    double s_left = 0.0;                // score[row][col - 1]
    double s_cur = 0.0;                 // score[row][col] of this lane's previous step (what the lane below needs as "up")
    double cs_cur = 0.0;                // col_score[col] after this lane's previous step
    double up_prev = 0.0;               // score[row - 1][col - 1]
    double local_max = 0.0;
    for (int d = 0; d < len_a + len_b - 1; ++d) {
        const int col = d - lane + 1;
        double up = __shfl_up_sync(0xffffffffu, s_cur, 1);
        double cs_up = __shfl_up_sync(0xffffffffu, cs_cur, 1);
        const bool act = live && col >= 1 && col <= len_b;
        if (lane == 0) { up = 0.0; cs_up = affine_penalty(col, 2.0 * open, extend); }
        if (act) {
            const double nogap = up_prev + ss.match[a][w.seq_b[col - 1]];
            double row_open, row_extend, col_open, col_extend;
            if (row == len_a) { row_open = s_left; row_extend = row_score; }
            else { row_open = s_left + first_gap; row_extend = row_score + extend; }
            row_score = fmax(row_open, row_extend);
            if (col == len_b) { col_open = up; col_extend = cs_up; }
            else { col_open = up + first_gap; col_extend = cs_up + extend; }
            const double col_score = fmax(col_open, col_extend);
            const double best = fmax(fmax(nogap, col_score), row_score);
            local_max = fmax(local_max, best);
            const double sc = best < 0.0 ? 0.0 : best;
            w.score[row][col] = sc;
            w.rowsc[row][col] = row_score;
            w.colsc[row][col] = col_score;
            s_left = sc;
            s_cur = sc;
            cs_cur = col_score;
        }
        up_prev = (col >= 0 && col <= len_b) ? up : 0.0;        // score[row - 1][col] becomes the diagonal of the next step
        if (col == 0) up_prev = 0.0;                            // column 0 of the score matrix is 0
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmax(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    __syncwarp();
    return local_max;
}

// pairwise2's trace bits of one cell (1 open gap in A, 2 match, 4 open gap in B, 8 extend A, 16 extend B; 0 = None),
// recomputed from the three score matrices with the recurrence's own operations.
__device__ int trace_bits(const WarpSmem &w, const ScoreSet &ss, int len_a, int len_b, int row, int col) {
    if (row < 1 || col < 1) return 0;
    const double open = ss.open, extend = ss.extend;
    const double first_gap = affine_penalty(1, open, extend);
    const double nogap = w.score[row - 1][col - 1] + ss.match[w.seq_a[row - 1]][w.seq_b[col - 1]];
    const double s_left = w.score[row][col - 1], up = w.score[row - 1][col];
    const double rs_prev = (col == 1) ? affine_penalty(row, 2.0 * open, extend) : w.rowsc[row][col - 1];
    const double cs_prev = (row == 1) ? affine_penalty(col, 2.0 * open, extend) : w.colsc[row - 1][col];
    double row_open, row_extend, col_open, col_extend;
    if (row == len_a) { row_open = s_left; row_extend = rs_prev; }
    else { row_open = s_left + first_gap; row_extend = rs_prev + extend; }
    const double row_score = fmax(row_open, row_extend);
    if (col == len_b) { col_open = up; col_extend = cs_prev; }
    else { col_open = up + first_gap; col_extend = cs_prev + extend; }
    const double col_score = fmax(col_open, col_extend);
    const double best = fmax(fmax(nogap, col_score), row_score);
    if (best <= 0.0) return 0;
    const int rs = rint1000(row_score), cs = rint1000(col_score), bs = rint1000(best);
    const int row_trace = (rint1000(row_open) == rs ? 1 : 0) + (rint1000(row_extend) == rs ? 8 : 0);
    const int col_trace = (rint1000(col_open) == cs ? 4 : 0) + (rint1000(col_extend) == cs ? 16 : 0);
    int t = rint1000(nogap) == bs ? 2 : 0;
    if (rs == bs) t += row_trace;
    if (cs == bs) t += col_trace;
    return t;
}
__device__ void trace_pass(WarpSmem &w, const ScoreSet &ss, int len_a, int len_b, int lane) {
    const int n = len_a * len_b;
    for (int i = lane; i < n; i += 32) {
        const int row = i / len_b + 1, col = i % len_b + 1;
        w.trace[row][col] = (uint8_t)trace_bits(w, ss, len_a, len_b, row, col);
    }
    __syncwarp();
}

// ---- 2. admissible start cells ---------------------------------------------------------------------------------------
__device__ void mark_starts(WarpSmem &w, const ScoreSet &ss, int len_a, int len_b, double best, int lane) {
    const int row = lane + 1;
    uint32_t bits = 0;
    if (row <= len_a) {
        for (int col = 1; col <= len_b; ++col) {
            const double sc = w.score[row][col];
            if (rint1000(fabs(sc - best)) > 0) continue;
            const double sd = w.score[row - 1][col - 1];
            if (rint1000(fabs(sd - best)) <= 0 && sd == sc) continue;       // zero-extension of another start
            if (sc <= 0.0) continue;
            const int t = w.trace[row][col];
            if (((t - (t % 2)) % 4) != 2) continue;                          // must end on a match
            bits |= 1u << col;
        }
    }
    __syncwarp();                       // every lane has read the traces it needs before any is overwritten
    if (row < DIM) {
        w.valid[row] = bits;
        for (int col = 1; col <= len_b; ++col)
            if ((bits >> col) & 1u) w.trace[row][col] = 2;       // pairwise2 does the same to every admissible start
    }
    __syncwarp();
}

// ---- 3. pairwise2's iterative back-trace (lane 0) -------------------------------------------------------------------
__device__ void backtrace(WarpSmem &w, const ScoreSet &ss, int len_a, int len_b, double best) {
    w.n_cols = -1;
    for (int r0 = len_a; r0 >= 1; --r0)
        for (int c0 = len_b; c0 >= 1; --c0) {
            if (!((w.valid[r0] >> c0) & 1u)) continue;
            int n = 0;
            {   // unaligned right flanks, reversed, gap-padded to the same length
                const int col_d = len_b - c0, row_d = len_a - r0;
                const int L = max(col_d, row_d);
                for (int i = 0; i < L; ++i) {
                    const int ga = col_d - row_d, gb = row_d - col_d;
                    w.ali_a[i] = (i < ga) ? GAPC : w.seq_a[len_a - 1 - (i - max(ga, 0))];
                    w.ali_b[i] = (i < gb) ? GAPC : w.seq_b[len_b - 1 - (i - max(gb, 0))];
                }
                n = L;
            }
            int sp = 0;
            int row = r0, col = c0, t = 2;
            bool col_gap = false;
            while (true) {
                bool dead_end = false;
                while ((row > 0 || col > 0) && !dead_end) {
                    const int c_n = n, c_row = row, c_col = col;
                    const bool c_gap = col_gap;
                    bool finished = false;
                    if (t == 0) {
                        if (col && col_gap) dead_end = true;
                        else {          // _finish_backtrace: left flanks, right-aligned against each other
                            const int L = max(row, col);
                            if (n + L > MAXCOL) { w.error = 1; return; }
                            for (int i = 0; i < L; ++i) {
                                w.ali_a[n + i] = (i < row) ? w.seq_a[row - 1 - i] : GAPC;
                                w.ali_b[n + i] = (i < col) ? w.seq_b[col - 1 - i] : GAPC;
                            }
                            n += L;
                        }
                        finished = true;
                    } else if (t % 2 == 1) {                    // open gap in A
                        t -= 1;
                        if (col_gap) dead_end = true;
                        else { col -= 1; w.ali_a[n] = GAPC; w.ali_b[n] = w.seq_b[col]; ++n; col_gap = false; }
                    } else if (t % 4 == 2) {                    // match / mismatch
                        t -= 2; row -= 1; col -= 1;
                        w.ali_a[n] = w.seq_a[row]; w.ali_b[n] = w.seq_b[col]; ++n; col_gap = false;
                    } else if (t % 8 == 4) {                    // open gap in B
                        t -= 4; row -= 1;
                        w.ali_a[n] = w.seq_a[row]; w.ali_b[n] = GAPC; ++n; col_gap = true;
                    } else if (t == 8 || t == 24 || t == 16) {  // extend a gap: walk back to where it was opened
                        const bool in_a = (t != 16);
                        t -= in_a ? 8 : 16;
                        if (in_a && col_gap) dead_end = true;
                        else {
                            col_gap = !in_a;
                            const int target = in_a ? col : row;
                            const double target_score = w.score[row][col];
                            for (int k = 0; k < target; ++k) {
                                if (in_a) { col -= 1; w.ali_a[n] = GAPC; w.ali_b[n] = w.seq_b[col]; }
                                else { row -= 1; w.ali_a[n] = w.seq_a[row]; w.ali_b[n] = GAPC; }
                                ++n;
                                const double actual = w.score[row][col] + affine_penalty(k + 1, ss.open, ss.extend);
                                if (w.score[row][col] == best) { dead_end = true; break; }
                                const int tt = w.trace[row][col];
                                if (rint1000(actual) == rint1000(target_score) && k > 0) {
                                    if (tt == 0) break;
                                    if (sp >= STACK) { w.error = 2; return; }
                                    w.stack[sp++] = StackEntry{(uint8_t)n, (uint8_t)row, (uint8_t)col, (uint8_t)col_gap, (uint8_t)tt, (uint8_t)c_n, (uint8_t)(in_a ? 1 : 2), 0};
                                }
                                if (tt == 0) dead_end = true;
                            }
                        }
                    } else { w.error = 3; return; }              // a trace value the recurrence cannot produce
                    if (finished) break;
                    if (n >= MAXCOL - 1) { w.error = 1; return; }
                    if (t) {
                        if (sp >= STACK) { w.error = 2; return; }
                        w.stack[sp++] = StackEntry{(uint8_t)c_n, (uint8_t)c_row, (uint8_t)c_col, (uint8_t)c_gap, (uint8_t)t, 0, 0, 0};
                    }
                    t = w.trace[row][col];
                    if (w.score[row][col] == best) dead_end = true;          // went through a zero-score extension
                    else if (w.score[row][col] <= 0.0) t = 0;                // start of the local alignment
                }
                if (!dead_end) { w.n_cols = n; return; }
                if (sp == 0) break;                                          // next start cell
                const StackEntry e = w.stack[--sp];
                n = e.n; row = e.row; col = e.col; col_gap = e.col_gap != 0; t = e.trace;
                for (int i = e.gap_from; e.gap_dir != 0 && i < n; ++i) {     // columns of the gap walk that led here
                    const int back = n - 1 - i;                             // moves still to undo at column i
                    if (e.gap_dir == 1) { w.ali_a[i] = GAPC; w.ali_b[i] = w.seq_b[col + back]; }
                    else { w.ali_a[i] = w.seq_a[row + back]; w.ali_b[i] = GAPC; }
                }
            }
        }
}

// ---- per-snippet preparation: ids -> base codes + length (evaluator :67, basecaller.py:289-294) ------------------------
__global__ void prepare_kernel(const int32_t *__restrict__ ids, int n, int steps, uint8_t *__restrict__ bases, int32_t *__restrict__ lens) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int m = 0;
    for (int t = 0; t < steps; ++t) {
        const int tok = ids[(size_t)i * steps + t];
        if (tok >= 3 && tok <= 6) bases[(size_t)i * steps + m++] = (uint8_t)(tok - 3);
    }
    lens[i] = m;
}

// utils.calc_prob_logits_beam_search_scores (utils.py:123-128): exp(score_t - score_{t-1}), score_{-1} = 0
__global__ void probs_kernel(const float *__restrict__ scores, long long n, int steps, float *__restrict__ probs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = (int)(i % steps);
    const float prev = t ? scores[i - 1] : 0.0f;
    probs[i] = expf(scores[i] - prev);
}

__global__ void __launch_bounds__(32 * WARPS) merge_kernel(const uint8_t *__restrict__ bases, const int32_t *__restrict__ lens,
                                                           const float *__restrict__ probs, int steps,
                                                           const int32_t *__restrict__ read_off, int n_reads, ScoreSet ss,
                                                           uint8_t *__restrict__ seq_out, float *__restrict__ log_out,
                                                           int32_t *__restrict__ len_out, int *__restrict__ err_flag) {
    __shared__ WarpSmem smem[WARPS];
    const int lane = threadIdx.x & 31;
    const int read = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (read >= n_reads) return;
    WarpSmem &w = smem[threadIdx.x >> 5];
    if (lane == 0) w.error = 0;
    for (int i = lane; i < DIM; i += 32) { w.score[0][i] = 0.0; w.score[i][0] = 0.0; w.trace[0][i] = 0; w.trace[i][0] = 0; }    // never written afterwards
    __syncwarp();
    const int s0 = read_off[read], s1 = read_off[read + 1];
    uint8_t *mseq = seq_out + (size_t)s0 * steps;
    float *mlog = log_out + (size_t)s0 * steps;
    int mlen = 0;
    bool merge_flag = false;
    for (int s = s0; s < s1; ++s) {
        const uint8_t *aseq = bases + (size_t)s * steps;
        const float *alog = probs + (size_t)s * steps;
        const int alen = lens[s];
        if (s == s0) {                                           // seq_merged = snippets[0]
            for (int i = lane; i < alen; i += 32) { mseq[i] = aseq[i]; mlog[i] = alog[i]; }
            mlen = alen;
            __syncwarp();
            continue;
        }
        const int len_a = min(mlen, OVL), len_b = min(alen, OVL);
        int n_cols = -1;
        if (len_a > 0 && len_b > 0) {
            if (lane < len_a) { w.seq_a[lane] = mseq[mlen - len_a + lane]; w.log_a[lane] = mlog[mlen - len_a + lane]; }
            if (lane < len_b) { w.seq_b[lane] = aseq[lane]; w.log_b[lane] = alog[lane]; }
            __syncwarp();
            const double best = fill_matrices(w, ss, len_a, len_b, lane);
            trace_pass(w, ss, len_a, len_b, lane);
            mark_starts(w, ss, len_a, len_b, best, lane);
            if (lane == 0) {
                backtrace(w, ss, len_a, len_b, best);
                const int n = w.n_cols;
                if (n > 0) {                                     // SingleMergerByLogits over the gapped strings (merger.py:86-119)
                    int ia = 0, ib = 0;
                    for (int i = 0; i < n; ++i) {
                        const uint8_t ca = w.ali_a[n - 1 - i], cb = w.ali_b[n - 1 - i];
                        const float la = (ca == GAPC) ? -1.0f : w.log_a[ia];
                        const float lb = (cb == GAPC) ? -1.0f : w.log_b[ib];
                        ia += ca != GAPC; ib += cb != GAPC;
                        const bool take_b = (ca == GAPC) || (cb != GAPC && lb > la);
                        w.out_seq[i] = take_b ? cb : ca;
                        w.out_log[i] = take_b ? lb : la;
                    }
                }
            }
            __syncwarp();
            n_cols = w.n_cols;
            if (w.error) { if (lane == 0) atomicExch(err_flag, w.error); break; }
        }
        if (n_cols <= 0) {                                       // no alignment (merger.py:172-190)
            if (!merge_flag) {
                for (int i = lane; i < alen; i += 32) { mseq[i] = aseq[i]; mlog[i] = alog[i]; }
                mlen = alen;
                __syncwarp();
                continue;
            }
            break;                                               // "merged seq already found, so that is returned"
        }
        merge_flag = true;
        const int keep = mlen - len_a;                           // seq_merged[:-25]
        for (int i = lane; i < n_cols; i += 32) { mseq[keep + i] = w.out_seq[i]; mlog[keep + i] = w.out_log[i]; }
        const int rest = alen - len_b;                           // seq_appended[25:]
        for (int i = lane; i < rest; i += 32) { mseq[keep + n_cols + i] = aseq[len_b + i]; mlog[keep + n_cols + i] = alog[len_b + i]; }
        mlen = keep + n_cols + rest;
        __syncwarp();
    }
    if (lane == 0) len_out[read] = mlen;
}

static int score_set(int id, ScoreSet *ss) {
    static const double m2[4][4] = {{10., -3., -1., -4.}, {-3., 9., -5., 0.}, {-1., -5., 7., -3.}, {-4., 0., -3., 8.}};   // merger.py:136-141
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            ss->match[i][j] = (id == 0) ? (i == j ? 1.0 : -1.0) : (id == 1) ? (i == j ? 5.0 : -4.0) : m2[i][j];
    if (id == 0) { ss->open = -1.0; ss->extend = -0.2; }
    else if (id == 1) { ss->open = -3.0; ss->extend = -0.1; }
    else if (id == 2) { ss->open = -9.0; ss->extend = -2.0; }
    else return fail(RVB_ERR_ARG, "merge_reads: scores_id must be 0, 1 or 2 (merger.py:124-147)");
    return RVB_OK;
}

}  // namespace mrg
}  // namespace rvb

using namespace rvb;

extern "C" int rvb_beam_scores_to_probs(const float *d_scores, int64_t n_snippets, int steps, float *d_probs, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (n_snippets < 0 || steps < 0) return fail(RVB_ERR_ARG, "beam_scores_to_probs: bad argument");
    const long long n = (long long)n_snippets * steps;
    if (n == 0) return RVB_OK;
    if (!d_scores || !d_probs) return fail(RVB_ERR_ARG, "beam_scores_to_probs: null pointer");
    ProfScope ps(KK_OTHER, s);
    mrg::probs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_scores, n, steps, d_probs);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

extern "C" int rvb_merge_reads(const int32_t *d_ids, const float *d_probs, int64_t n_snippets, int steps,
                               const int32_t *d_read_offsets, int n_reads, int scores_id,
                               uint8_t *d_seq_out, float *d_prob_out, int32_t *d_len_out, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (n_snippets < 0 || steps < 0 || steps > 255 || n_reads < 0) return fail(RVB_ERR_ARG, "merge_reads: bad argument");
    mrg::ScoreSet ss;
    RVB_CHECK(mrg::score_set(scores_id, &ss));
    if (n_reads == 0) return RVB_OK;
    if (!d_read_offsets || !d_len_out) return fail(RVB_ERR_ARG, "merge_reads: null pointer");
    if (n_snippets > 0 && (!d_ids || !d_probs || !d_seq_out || !d_prob_out)) return fail(RVB_ERR_ARG, "merge_reads: null pointer");
    uint8_t *scratch = nullptr;
    const size_t base_bytes = ((size_t)n_snippets * steps + 15) & ~size_t(15);
    const size_t bytes = base_bytes + sizeof(int32_t) * (size_t)n_snippets + 16;
    RVB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), bytes, s));
    uint8_t *bases = scratch;
    int32_t *lens = reinterpret_cast<int32_t *>(scratch + base_bytes);
    int *err = reinterpret_cast<int *>(lens + n_snippets);
    int st = RVB_OK;
    do {
        if (cudaMemsetAsync(err, 0, sizeof(int), s) != cudaSuccess) { st = fail(RVB_ERR_CUDA, "merge_reads: memset"); break; }
        ProfScope ps(KK_OTHER, s);
        if (n_snippets > 0) mrg::prepare_kernel<<<(unsigned)((n_snippets + 127) / 128), 128, 0, s>>>(d_ids, (int)n_snippets, steps, bases, lens);
        mrg::merge_kernel<<<(unsigned)((n_reads + mrg::WARPS - 1) / mrg::WARPS), 32 * mrg::WARPS, 0, s>>>(
            bases, lens, d_probs, steps, d_read_offsets, n_reads, ss, d_seq_out, d_prob_out, d_len_out, err);
        if (cudaGetLastError() != cudaSuccess) { st = fail(RVB_ERR_CUDA, "merge_reads: kernel launch failed"); break; }
        count_launch(2);
        int h_err = 0;
        if (cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) { st = fail(RVB_ERR_CUDA, "merge_reads: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (h_err) st = fail(RVB_ERR_INTERNAL, "merge_reads: back-trace workspace exhausted (code %d)", h_err);
    } while (0);
    cudaFreeAsync(scratch, s);
    return st;
}
