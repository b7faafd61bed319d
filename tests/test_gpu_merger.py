"""GPU read stitching (csrc/merger.cu through rvb_merge_reads) vs oracle/merger_ref.py: bit-exact sequences and
probabilities on identical inputs -- SURVEY §8 f-1."""
import numpy as np
import pytest

from oracle import merger_ref as m

pytestmark = pytest.mark.gpu

CODE = {"A": 3, "C": 4, "G": 5, "T": 6}


def _rand_seq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


def _noisy_snippets(rng, read, width, stride, err):
    """Windows of `read` with substitutions / insertions / deletions at rate err, random probabilities."""
    out, pos = [], 0
    while pos + width <= len(read):
        s = []
        for ch in read[pos:pos + width]:
            u = rng.random()
            if u < err / 3:
                continue
            if u < 2 * err / 3:
                s.append(rng.choice(list("ACGT")))
            s.append(ch if u > err else rng.choice(list("ACGT")))
        s = "".join(s)[:33]
        out.append((s, [float(np.float32(x)) for x in rng.uniform(0.2, 1.0, len(s))]))
        pos += stride
    return out


def _pack(reads, S=33):
    n = sum(len(r) for r in reads)
    ids = np.zeros((n, S), np.int32)
    probs = np.zeros((n, S), np.float32)
    off, i = [0], 0
    for r in reads:
        for seq, lg in r:
            # bases interleaved with non-base tokens, as decoded rows are: start/end/pad tokens are skipped
            row = [CODE[c] for c in seq] + [1] + [0] * (S - len(seq) - 1)
            ids[i, :S] = row[:S]
            probs[i, :len(seq)] = lg
            i += 1
        off.append(i)
    return ids, probs, off


@pytest.mark.parametrize("score_set", [0, 1, 2])
def test_merge_reads_matches_oracle(score_set):
    import ravvent_basecaller_b200 as rb
    rng = np.random.default_rng(50 + score_set)
    reads = []
    for k in range(12):
        read = _rand_seq(rng, int(rng.integers(60, 500)))
        reads.append(_noisy_snippets(rng, read, 30, 5, err=[0.0, 0.03, 0.08, 0.2][k % 4]))
    reads.append([(_rand_seq(rng, 12), [0.5] * 12)])                       # a single-snippet read
    reads.append([("", []), ("ACGT", [.5] * 4), ("", []), ("ACGTT", [.6] * 5)])
    reads.append([("AACCAACC", [.5] * 8), ("AACCAACC", [.5] * 8), ("GGTTGGTT", [.5] * 8), ("AACCAACC", [.5] * 8)])
    reads.append([("AAAAAA", [.5] * 6), ("CCCCCC", [.5] * 6), ("CCCCGG", [.6] * 6)])
    ids, probs, off = _pack(reads)
    got = rb.Merger(score_set).merge_predictions(ids, None, off, probs=probs)
    assert len(got) == len(reads)
    n_gapped = 0
    for r, g in zip(reads, got):
        want_seq, want_log = m.merge_read(r, score_set)
        assert g.seq == want_seq
        assert np.array_equal(np.asarray(g.logits, np.float32), np.asarray(want_log, np.float32))
        n_gapped += len(want_seq) != sum(len(s) for s, _ in r)
    assert n_gapped > 0


def test_random_overlaps_exhaustive_small():
    """Many independent 2-snippet reads: every alignment shape (gaps, flanks, ties) against the oracle."""
    import ravvent_basecaller_b200 as rb
    rng = np.random.default_rng(77)
    reads = []
    for _ in range(1500):
        a = _rand_seq(rng, int(rng.integers(1, 34)))
        if rng.random() < 0.7:                                            # related second snippet: shifted copy with edits
            k = int(rng.integers(0, len(a)))
            b = list(a[k:] + _rand_seq(rng, int(rng.integers(0, 12))))
            for _e in range(int(rng.integers(0, 4))):
                if b:
                    j = int(rng.integers(0, len(b)))
                    op = rng.integers(0, 3)
                    if op == 0: b[j] = rng.choice(list("ACGT"))
                    elif op == 1: b.insert(j, rng.choice(list("ACGT")))
                    else: del b[j]
            b = "".join(b)[:33]
        else:
            b = _rand_seq(rng, int(rng.integers(0, 34)))
        reads.append([(a, [float(np.float32(x)) for x in rng.uniform(0, 1, len(a))]),
                      (b, [float(np.float32(x)) for x in rng.uniform(0, 1, len(b))])])
    ids, probs, off = _pack(reads)
    for score_set in (0, 1, 2):
        got = rb.Merger(score_set).merge_predictions(ids, None, off, probs=probs)
        bad = [i for i, (r, g) in enumerate(zip(reads, got))
               if (g.seq, [np.float32(x) for x in g.logits]) != (m.merge_read(r, score_set)[0], [np.float32(x) for x in m.merge_read(r, score_set)[1]])]
        assert not bad, (score_set, bad[:5], reads[bad[0]], got[bad[0]].seq, m.merge_read(reads[bad[0]], score_set)[0])


def test_reference_surface_and_scores_path():
    import torch
    import ravvent_basecaller_b200 as rb
    mg = rb.Merger()
    # merger.py:251-257
    s1, s2 = "AGTTCAGCGATCGGATCCGCGTGC", "GAGATTTTATCCGCGTGCTGTTTACG"
    out = mg.merge([rb.SeqLogitsPair(s1, [0.5] * len(s1)), rb.SeqLogitsPair(s2, [0.7] * len(s2))])
    want = m.merge_read([(s1, [0.5] * len(s1)), (s2, [0.7] * len(s2))])
    assert out.seq == want[0] and np.allclose(out.logits, want[1])
    assert rb.SeqLogitsPair.align_logits("A-C", [.1, .2]) == [.1, -1., .2]
    # beam scores -> probabilities on the device (utils.py:123-128), then the same merge from ids + scores
    rng = np.random.default_rng(4)
    scores = np.cumsum(np.log(rng.uniform(0.3, 1.0, (40, 33))), axis=1).astype(np.float32)
    ids = rng.integers(3, 7, (40, 33)).astype(np.int32)
    ids[:, 30:] = 1
    got = mg.merge_predictions(torch.from_numpy(ids).cuda(), torch.from_numpy(scores).cuda(), [0, 25, 40])
    probs = m.beam_scores_to_probs(scores)
    sn = m.snippets_from_predictions(ids, scores)
    for r, (a, b) in enumerate([(0, 25), (25, 40)]):
        want_seq, want_log = m.merge_read(sn[a:b])
        assert got[r].seq == want_seq
        assert np.allclose(got[r].logits, want_log, rtol=2e-6)
    assert np.isfinite(probs).all()
    with pytest.raises(ValueError):
        mg.merge_predictions(ids, scores, [0, 10])
    with pytest.raises(ValueError):
        rb.Merger(3)
    assert mg.merge_predictions(ids[:0], scores[:0], [0]) == []


def test_full_size_reads_rebuild_exactly():
    """Size-independent property at bench scale: error-free stride-5 windows of 100 reads x 1000 snippets."""
    import torch
    import ravvent_basecaller_b200 as rb
    rng = np.random.default_rng(8)
    R, n_snip, W, stride = 100, 1000, 30, 5
    L = (n_snip - 1) * stride + W
    reads = rng.integers(0, 4, (R, L)).astype(np.int32)
    idx = (np.arange(n_snip)[:, None] * stride + np.arange(W)[None, :])
    ids = np.ones((R * n_snip, 33), np.int32)
    ids[:, :W] = (reads[:, idx] + 3).reshape(R * n_snip, W)
    probs = np.full((R * n_snip, 33), 0.9, np.float32)
    mg = rb.Merger()
    off = np.arange(R + 1) * n_snip
    d_ids, d_probs = torch.from_numpy(ids).cuda(), torch.from_numpy(probs).cuda()
    mg.merge_predictions(d_ids, None, off, probs=d_probs)                        # warm-up
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    got = mg.merge_predictions(d_ids, None, off, probs=d_probs)
    t1.record()
    torch.cuda.synchronize()
    print(f"merge of {R} reads x {n_snip} snippets: {t0.elapsed_time(t1):.1f} ms incl. host unpack")
    for r in range(R):
        assert got[r].seq == "".join("ACGT"[c] for c in reads[r])
