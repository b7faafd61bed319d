"""Oracle for snippet -> read stitching (oracle/merger_ref.py).  Biopython is absent ("parity unpinned"):
the tests pin what does not depend on pairwise2's order of co-optimal alignments."""
import numpy as np
import pytest

from oracle import merger_ref as m


def _rand_seq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


def _region_score(a, b, begin, end, score_set):
    s = m.SCORE_SETS[score_set]
    mf = m.match_function(score_set)
    total, gap_a, gap_b = 0.0, False, False
    for x, y in zip(a[begin:end], b[begin:end]):
        if x == "-":
            total += s["gap_extend"] if gap_a else s["gap_open"]
            gap_a, gap_b = True, False
        elif y == "-":
            total += s["gap_extend"] if gap_b else s["gap_open"]
            gap_a, gap_b = False, True
        else:
            total += mf(x, y)
            gap_a = gap_b = False
    return total


@pytest.mark.parametrize("score_set", [0, 1, 2])
def test_local_alignment_is_optimal_and_consistent(score_set):
    rng = np.random.default_rng(100 + score_set)
    for _ in range(250):
        a, b = _rand_seq(rng, rng.integers(1, 26)), _rand_seq(rng, rng.integers(1, 26))
        al = m.local_align(a, b, score_set, first_only=False)
        best = m.brute_force_local_score(a, b, score_set)
        if not al:
            assert best <= 0
            continue
        for ga, gb, score, begin, end in al:
            assert len(ga) == len(gb) and ga.replace("-", "") == a and gb.replace("-", "") == b
            assert not any(x == "-" and y == "-" for x, y in zip(ga, gb))
            assert score == pytest.approx(best, abs=1e-9)
            assert _region_score(ga, gb, begin, end, score_set) == pytest.approx(score, abs=1e-9)
            assert ga[begin] != "-" and gb[begin] != "-" and ga[end - 1] != "-" and gb[end - 1] != "-"
        assert m.local_align(a, b, score_set)[0] == al[0]


def test_reference_main_example():
    # merger.py:251-257 (`python merger.py`): the 10-base shared core must survive the merge
    s1, s2 = "AGTTCAGCGATCGGATCCGCGTGC", "GAGATTTTATCCGCGTGCTGTTTACG"
    seq, logits = m.merge_read([(s1, [0.5] * len(s1)), (s2, [0.7] * len(s2))])
    assert "ATCCGCGTGC" in seq and seq.endswith("TGTTTACG") and len(seq) == len(logits)
    assert set(logits) <= {0.5, 0.7}


def test_error_free_overlapping_snippets_rebuild_the_read():
    rng = np.random.default_rng(3)
    for score_set in (0, 1, 2):
        read = _rand_seq(rng, 400)
        snips, pos = [], 0
        while pos + 30 <= len(read):
            snips.append((read[pos:pos + 30], [0.9] * 30))
            pos += 5                                               # consecutive snippets share exactly 25 bases
        seq, logits = m.merge_read(snips, score_set)
        assert seq == read[:pos + 25] and len(logits) == len(seq)


def test_merge_quirks():
    # leading snippets that do not align are dropped until one does (merger.py:176-183) ...
    seq, _ = m.merge_read([("AAAAAA", [.5] * 6), ("CCCCCC", [.5] * 6), ("CCCCGG", [.6] * 6)])
    assert seq.startswith("CCCC") and seq.endswith("GG") and "A" not in seq
    # ... but once merging has begun, a snippet that does not align ends the read (:184-190)
    two = m.merge_read([("AACCAACC", [.5] * 8), ("AACCAACC", [.5] * 8)])
    assert m.merge_read([("AACCAACC", [.5] * 8), ("AACCAACC", [.5] * 8), ("GGTTGGTT", [.5] * 8), ("AACCAACC", [.5] * 8)]) == two
    # empty snippets never align
    assert m.merge_read([("", []), ("ACGT", [.5] * 4)])[0] == "ACGT"
    assert m.merge_read([("ACGT", [.5] * 4), ("ACGT", [.4] * 4), ("", []), ("ACGT", [.5] * 4)])[0] == "ACGT"
    # higher probability wins on a mismatch inside the aligned overlap, ties keep the left base
    seq3, log3 = m.merge_read([("ACGTAACGT", [.5] * 9), ("ACGTCACGT", [.5, .5, .5, .5, .9, .5, .5, .5, .5])])
    assert seq3 == "ACGTCACGT" and max(log3) == .9
    assert m.merge_read([("ACGTAACGT", [.5] * 9), ("ACGTCACGT", [.5] * 9)])[0] == "ACGTAACGT"


def test_beam_scores_to_probs_and_snippets():
    scores = np.log(np.array([[0.5, 0.25, 0.125], [0.9, 0.81, 0.0081]], dtype=np.float32))
    p = m.beam_scores_to_probs(scores)
    assert np.allclose(p, [[0.5, 0.5, 0.5], [0.9, 0.9, 0.01]], rtol=1e-5)
    ids = np.array([[3, 6, 1], [2, 4, 4]])
    sn = m.snippets_from_predictions(ids, scores)
    assert sn[0][0] == "AT" and len(sn[0][1]) == 2 and sn[1][0] == "CC"
    assert sn[1][1] == pytest.approx([0.9, 0.9], rel=1e-5)          # the FIRST len(seq) probabilities, as the evaluator slices them
