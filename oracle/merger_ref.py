"""CPU oracle for snippet -> read stitching.  TEST INFRASTRUCTURE ONLY.

Restates, in plain Python:

  * the reference's merger.py: SeqLogitsPair.align_logits (:10-23), SingleMergerByLogits.merge
    (:86-119), Merger.merge (:146-248; 25-base overlap, score sets 0 / 1 / 2), and
    utils.calc_prob_logits_beam_search_scores (utils.py:123-128) feeding it as in
    ravvent_performance_evaluator.py:66-75;
  * the THIRD-PARTY routine Merger.merge calls, Biopython `Bio.pairwise2.align.localms / localds`
    (affine-gap local alignment, default keywords, first returned alignment).  Biopython is not
    vendored in /root/reference, not pinned by it (no requirements file) and not installed here, so its
    published algorithm is restated: the "fast" affine score / trace matrices (trace bits 1 = open gap in A,
    2 = match, 4 = open gap in B, 8 = extend gap in A, 16 = extend gap in B; comparisons on
    rint(x) = int(1000 x + 0.5); no end-gap penalty in the last row / column), the list of start cells
    within tolerance 0 of the best score, and the iterative stack back-trace (last start first; per cell the
    order open-A, match, open-B, extend-A, extend-B; a gap in A may not follow a gap in B; zero-score
    extensions are discarded; unaligned flanks are appended and padded with gaps).

PARITY UNPINNED: with neither Biopython nor TensorFlow available there is no reference output to pin this
restatement against; which of several co-optimal alignments comes first follows the restated back-trace
order and may differ from a given Biopython release.  Properties that do not depend on that order are
tested (optimal score vs brute force, exact reconstruction of error-free overlapping snippets).
"""
from __future__ import annotations

import numpy as np

GAP = "-"
OVERLAP = 25                                   # merger.py:150
MAX_ALIGNMENTS = 1000                          # pairwise2.MAX_ALIGNMENTS

SCORE_SETS = {                                 # merger.py:124-147
    0: dict(match=1.0, mismatch=-1.0, gap_open=-1.0, gap_extend=-0.2),
    1: dict(match=5.0, mismatch=-4.0, gap_open=-3.0, gap_extend=-0.1),
    2: dict(matrix={("A", "A"): 10., ("A", "C"): -3., ("A", "G"): -1., ("A", "T"): -4.,
                    ("C", "A"): -3., ("C", "C"): 9., ("C", "G"): -5., ("C", "T"): 0.,
                    ("G", "A"): -1., ("G", "C"): -5., ("G", "G"): 7., ("G", "T"): -3.,
                    ("T", "A"): -4., ("T", "C"): 0., ("T", "G"): -3., ("T", "T"): 8.},
            gap_open=-9.0, gap_extend=-2.0),
}


def rint(x):
    return int(x * 1000 + 0.5)


def affine_penalty(length, open_, extend):
    """pairwise2.calc_affine_penalty with penalize_extend_when_opening = 0."""
    if length <= 0:
        return 0.0
    penalty = open_ + extend * length
    penalty -= extend
    return penalty


def match_function(score_set):
    s = SCORE_SETS[score_set]
    if "matrix" in s:
        m = s["matrix"]
        return lambda a, b: m[(a, b)] if (a, b) in m else m[(b, a)]
    match, mismatch = s["match"], s["mismatch"]
    return lambda a, b: match if a == b else mismatch


def score_matrices(seq_a, seq_b, match_fn, open_, extend):
    """Local alignment, same affine penalty for both sequences, penalize_end_gaps = (False, False)."""
    len_a, len_b = len(seq_a), len(seq_b)
    first_gap = affine_penalty(1, open_, extend)
    score = [[0.0] * (len_b + 1) for _ in range(len_a + 1)]
    trace = [[None] * (len_b + 1) for _ in range(len_a + 1)]
    col_score = [0.0] + [affine_penalty(i, 2 * open_, extend) for i in range(1, len_b + 1)]
    local_max = 0.0
    for row in range(1, len_a + 1):
        row_score = affine_penalty(row, 2 * open_, extend)
        for col in range(1, len_b + 1):
            nogap = score[row - 1][col - 1] + match_fn(seq_a[row - 1], seq_b[col - 1])
            if row == len_a:
                row_open, row_extend = score[row][col - 1], row_score
            else:
                row_open, row_extend = score[row][col - 1] + first_gap, row_score + extend
            row_score = max(row_open, row_extend)
            if col == len_b:
                col_open, col_extend = score[row - 1][col], col_score[col]
            else:
                col_open, col_extend = score[row - 1][col] + first_gap, col_score[col] + extend
            col_score[col] = max(col_open, col_extend)
            best = max(nogap, col_score[col], row_score)
            local_max = max(local_max, best)
            score[row][col] = 0.0 if best < 0 else best
            rs, cs, bs = rint(row_score), rint(col_score[col]), rint(best)
            row_trace = (1 if rint(row_open) == rs else 0) + (8 if rint(row_extend) == rs else 0)
            col_trace = (4 if rint(col_open) == cs else 0) + (16 if rint(col_extend) == cs else 0)
            t = 2 if rint(nogap) == bs else 0
            if rs == bs:
                t += row_trace
            if cs == bs:
                t += col_trace
            trace[row][col] = None if best <= 0 else t
    return score, trace, local_max


def _finish_backtrace(seq_a, seq_b, ali_a, ali_b, row, col):
    if row:
        ali_a += seq_a[row - 1::-1]
    if col:
        ali_b += seq_b[col - 1::-1]
    if row > col:
        ali_b += GAP * (len(ali_a) - len(ali_b))
    elif col > row:
        ali_a += GAP * (len(ali_b) - len(ali_a))
    return ali_a, ali_b


def _find_gap_open(seq_a, seq_b, ali_a, ali_b, end, row, col, col_gap, score, trace, in_process, open_, extend,
                   target, direction, best_score):
    dead_end = False
    target_score = score[row][col]
    for n in range(target):
        if direction == "col":
            col -= 1
            ali_a += GAP
            ali_b += seq_b[col:col + 1]
        else:
            row -= 1
            ali_a += seq_a[row:row + 1]
            ali_b += GAP
        actual = score[row][col] + affine_penalty(n + 1, open_, extend)
        if score[row][col] == best_score:
            dead_end = True
            break
        if rint(actual) == rint(target_score) and n > 0:
            if not trace[row][col]:
                break
            in_process.append((ali_a, ali_b, end, row, col, col_gap, trace[row][col]))
        if not trace[row][col]:
            dead_end = True
    return ali_a, ali_b, row, col, dead_end


def local_align(seq_a, seq_b, score_set=0, first_only=True):
    """-> list of (gapped_a, gapped_b, score, begin, end) in pairwise2's order ([] when nothing aligns).
    first_only stops after the first completed back-trace (all that Merger.merge reads)."""
    if not seq_a or not seq_b:
        return []
    s = SCORE_SETS[score_set]
    open_, extend = s["gap_open"], s["gap_extend"]
    score, trace, best_score = score_matrices(seq_a, seq_b, match_function(score_set), open_, extend)
    len_a, len_b = len(seq_a), len(seq_b)
    starts = [(score[r][c], (r, c)) for r in range(len_a + 1) for c in range(len_b + 1)
              if rint(abs(score[r][c] - best_score)) <= 0]
    start_set = set(starts)
    in_process = []
    for sc, (row, col) in starts:
        if (sc, (row - 1, col - 1)) in start_set:        # zero-extension of another start
            continue
        if sc <= 0:
            continue
        t = trace[row][col]
        if t is None or (t - t % 2) % 4 != 2:            # must end on a match, not a gap
            continue
        trace[row][col] = 2
        end = -max(len_a - row, len_b - col) or None
        col_d, row_d = len_b - col, len_a - row
        ali_a = (col_d - row_d) * GAP + seq_a[len_a - 1:row - 1:-1]
        ali_b = (row_d - col_d) * GAP + seq_b[len_b - 1:col - 1:-1]
        in_process.append((ali_a, ali_b, end, row, col, False, 2))
    out, begin = [], 0
    while in_process and len(out) < MAX_ALIGNMENTS:
        dead_end = False
        ali_a, ali_b, end, row, col, col_gap, t = in_process.pop()
        while (row > 0 or col > 0) and not dead_end:
            cache = (ali_a, ali_b, end, row, col, col_gap)
            if not t:
                if col and col_gap:
                    dead_end = True
                else:
                    ali_a, ali_b = _finish_backtrace(seq_a, seq_b, ali_a, ali_b, row, col)
                break
            elif t % 2 == 1:                             # open gap in A
                t -= 1
                if col_gap:
                    dead_end = True
                else:
                    col -= 1
                    ali_a += GAP
                    ali_b += seq_b[col:col + 1]
                    col_gap = False
            elif t % 4 == 2:                             # match / mismatch
                t -= 2
                row -= 1
                col -= 1
                ali_a += seq_a[row:row + 1]
                ali_b += seq_b[col:col + 1]
                col_gap = False
            elif t % 8 == 4:                             # open gap in B
                t -= 4
                row -= 1
                ali_a += seq_a[row:row + 1]
                ali_b += GAP
                col_gap = True
            elif t in (8, 24):                           # extend gap in A
                t -= 8
                if col_gap:
                    dead_end = True
                else:
                    col_gap = False
                    ali_a, ali_b, row, col, dead_end = _find_gap_open(
                        seq_a, seq_b, ali_a, ali_b, end, row, col, col_gap, score, trace, in_process, open_, extend,
                        col, "col", best_score)
            elif t == 16:                                # extend gap in B
                t -= 16
                col_gap = True
                ali_a, ali_b, row, col, dead_end = _find_gap_open(
                    seq_a, seq_b, ali_a, ali_b, end, row, col, col_gap, score, trace, in_process, open_, extend,
                    row, "row", best_score)
            if t:
                in_process.append(cache + (t,))
            t = trace[row][col]
            if score[row][col] == best_score:            # went through a zero-score extension
                dead_end = True
            elif score[row][col] <= 0:                   # start of the local alignment
                begin = max(row, col)
                t = 0
        if not dead_end:
            a, b = ali_a[::-1], ali_b[::-1]
            e = len(a) if end is None else end + len(a)
            if begin < e and (a, b, best_score, begin, e) not in out:
                out.append((a, b, best_score, begin, e))
                if first_only:
                    break
    return out


# ---------------------------------------------------------------- merger.py
def align_logits(seq_gapped, logits):
    out, i = [], 0
    for ch in seq_gapped:
        if ch == GAP:
            out.append(-1.0)
        else:
            out.append(logits[i])
            i += 1
    return out


def merge_by_logits(seq1, logits1, seq2, logits2):
    """SingleMergerByLogits.merge (merger.py:86-119)."""
    assert len(seq1) == len(seq2)
    seq, logits = [], []
    for n1, n2, l1, l2 in zip(seq1, seq2, logits1, logits2):
        if n1 == GAP:
            seq.append(n2); logits.append(l2)
        elif n2 == GAP:
            seq.append(n1); logits.append(l1)
        elif l2 > l1:
            seq.append(n2); logits.append(l2)
        else:
            seq.append(n1); logits.append(l1)
    return "".join(seq), logits


def merge_read(snippets, score_set=0):
    """Merger.merge (merger.py:146-248).  snippets: list of (seq str, logits list) -> (seq, logits)."""
    seq_m, log_m = snippets[0][0], list(snippets[0][1])
    merge_flag = False
    for seq_app, log_app in snippets[1:]:
        log_app = list(log_app)
        algns = local_align(seq_m[-OVERLAP:], seq_app[:OVERLAP], score_set)
        if not algns:
            if not merge_flag:
                seq_m, log_m = seq_app, log_app
                continue
            return seq_m, log_m
        merge_flag = True
        a, b = algns[0][0], algns[0][1]
        l1 = align_logits(a, log_m[-OVERLAP:])
        l2 = align_logits(b, log_app[:OVERLAP])
        seq_o, log_o = merge_by_logits(a, l1, b, l2)
        seq_m = seq_m[:-OVERLAP] + seq_o + seq_app[OVERLAP:]
        log_m = log_m[:-OVERLAP] + log_o + log_app[OVERLAP:]
    return seq_m, log_m


def beam_scores_to_probs(beam_scores):
    """utils.calc_prob_logits_beam_search_scores (utils.py:123-128): exp(score_t - score_{t-1}), score_{-1} = 0."""
    s = np.asarray(beam_scores, dtype=np.float32)
    prev = np.zeros_like(s)
    prev[..., 1:] = s[..., :-1]
    return np.exp(s - prev)


def snippets_from_predictions(ids, scores):
    """ravvent_performance_evaluator.py:66-70: token rows -> (base string, probabilities[:len(seq)])."""
    probs = beam_scores_to_probs(scores)
    table = {3: "A", 4: "C", 5: "G", 6: "T"}
    out = []
    for row, p in zip(np.asarray(ids), probs):
        seq = "".join(table.get(int(t), "") for t in row)
        out.append((seq, [float(x) for x in p[:len(seq)]]))
    return out


# ---------------------------------------------------------------- independent checks used by the tests
def brute_force_local_score(seq_a, seq_b, score_set=0):
    """Best local affine-gap score by the textbook three-matrix Gotoh recurrence (no Biopython quirks)."""
    s = SCORE_SETS[score_set]
    mf, o, e = match_function(score_set), s["gap_open"], s["gap_extend"]
    NEG = float("-inf")
    n, m = len(seq_a), len(seq_b)
    M = [[0.0] * (m + 1) for _ in range(n + 1)]
    X = [[NEG] * (m + 1) for _ in range(n + 1)]
    Y = [[NEG] * (m + 1) for _ in range(n + 1)]
    best = 0.0
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            # a gap may open straight after a gap in the other sequence (pairwise2 allows "A-" / "-B" columns)
            X[i][j] = max(max(M[i - 1][j], Y[i - 1][j]) + o, X[i - 1][j] + e)
            Y[i][j] = max(max(M[i][j - 1], X[i][j - 1]) + o, Y[i][j - 1] + e)
            d = max(M[i - 1][j - 1], X[i - 1][j - 1], Y[i - 1][j - 1]) + mf(seq_a[i - 1], seq_b[j - 1])
            M[i][j] = max(0.0, d)
            best = max(best, M[i][j])
    return best
