"""Row-by-row comparison of a beam search against the CPU oracle.  TEST INFRASTRUCTURE ONLY
(tests/, __graft_entry__.smoke()).

A row may differ from the oracle only if, at the FIRST decode step where its raw per-step outputs
(step ids or parent ids of any beam slot) differ, the oracle's own selection was a near tie: the
smallest gap among its W+1 best candidates at that step (model_ref.topk_margin) is below `tie_eps`.
Everything before that step must match, rows without such a step must be identical in every output,
and the number of near-tie rows is bounded.  This is the beam-search analogue of the greedy rule
"identical except for documented near-tie argmax flips" (BASELINE.json north_star)."""
from __future__ import annotations

import numpy as np

from . import model_ref as mr


def _close(a, b, rtol, atol):
    with np.errstate(invalid="ignore"):
        inf = np.isneginf(a) & np.isneginf(b)
    np.testing.assert_allclose(np.where(inf, 0.0, a), np.where(inf, 0.0, b), rtol=rtol, atol=atol)


def check_beam(got, w, enc, mask, W, L, decoder_depth=1, tie_eps=1e-3, max_tie_frac=0.02, rtol=1e-3, atol=2e-4,
               full_length=False, label="", coinflip_eps=2e-5):
    """got = (pred [B,T,W], scores [B,T,W], step_ids [B,T,W], parent_ids [B,T,W]) from the implementation under test.
    Returns the number of near-tie rows (after asserting everything else)."""
    pred, sc, sid, par = (np.asarray(a) for a in got)
    rp, rs, rsid, rpar, margin = mr.beam_search(w, enc, mask, W, L, decoder_depth=decoder_depth, return_all=True,
                                                return_margins=True, full_length=full_length)
    n = pred.shape[0]
    assert rp.shape[0] == n and pred.shape[2] == W
    T = min(pred.shape[1], rp.shape[1])
    ties = []
    for r in range(n):
        d = np.flatnonzero((sid[r, :T] != rsid[r, :T]).any(axis=1) | (par[r, :T] != rpar[r, :T]).any(axis=1))
        if d.size == 0:
            assert np.array_equal(pred[r, :T], rp[r, :T]), f"{label} row {r}: same search trace, different gather_tree output"
            _close(sc[r, :T], rs[r, :T], rtol, atol)
            continue
        t = int(d[0])
        assert margin[r, t] < tie_eps, (f"{label} row {r}: diverges from the oracle at step {t} where the oracle's top-{W + 1} "
                                        f"margin is {margin[r, t]:.3g} (not a near tie)")
        _close(sc[r, :t], rs[r, :t], rtol, atol)
        ties.append((r, t, float(margin[r, t])))
    if not ties:
        assert pred.shape[1] == rp.shape[1], f"{label}: executed {pred.shape[1]} decode steps, the oracle {rp.shape[1]}"
    # Bound on their number: `max_tie_frac` of the rows, plus the rows the oracle itself cannot decide -- some step of
    # theirs has a margin below `coinflip_eps`, the size of the fp32 rounding differences between two summation orders
    # of the same score (with random-init weights every token is almost equally likely, so after ~30 steps a large share
    # of the rows holds two hypotheses that close; the numpy and torch oracles split the same rows, see
    # tests/test_oracle_second_opinion.py).
    coinflips = int((margin[:, :T].min(axis=1) < coinflip_eps).sum())
    limit = max(1, int(np.ceil(max_tie_frac * n))) + coinflips
    assert len(ties) <= limit, f"{label}: {len(ties)} near-tie rows of {n} (limit {limit}, {coinflips} undecidable): {ties[:5]}"
    if ties:
        print(f"[parity]{label} {len(ties)} of {n} rows diverge at near ties, {coinflips} rows undecidable at {coinflip_eps:g} "
              f"(row, step, margin): {ties[:8]}")
    return len(ties)
