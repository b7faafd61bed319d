"""Snippet-builder oracle vs the reference's prepare_snippets (golden fixtures)."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import event_ref, snippet_ref


def _ulp_close(a, b, ulps=1):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("k", [0, 1, 2])
def test_build_snippets_matches_reference(k):
    g = np.load(GOLDEN / "snippets_golden.npz")
    raw = g[f"raw_{k}"].astype(np.int32)
    lab0, lab1, stride = (int(v) for v in g[f"par_{k}"])
    ev = event_ref.detect_events(raw, 6, 9)
    out = snippet_ref.build_snippets(raw, ev.start, ev.length, ev.mean, ev.stdv, lab0, lab1, stride)
    assert out["raw"].shape == g[f"raw_snips_{k}"].shape
    assert out["event"].shape == g[f"event_snips_{k}"].shape
    # integer windowing logic: exact
    assert np.array_equal(out["raw_ranges"][:, 1] - out["raw_ranges"][:, 0], g[f"raw_lens_{k}"])
    assert np.array_equal(out["event_ranges"][:, 1] - out["event_ranges"][:, 0], g[f"event_lens_{k}"])
    # padding pattern identical, values within 1 float32 ulp
    assert np.array_equal(out["raw"] == 0, g[f"raw_snips_{k}"] == 0)
    assert np.array_equal(out["event"] == 0, g[f"event_snips_{k}"] == 0)
    assert _ulp_close(out["raw"], g[f"raw_snips_{k}"])
    assert _ulp_close(out["event"], g[f"event_snips_{k}"])


def test_fitting_ranges_properties():
    rng = np.random.default_rng(3)
    lens = rng.integers(2, 30, size=400)
    r = snippet_ref.fitting_event_ranges(lens, 6, 200)
    assert (r[:, 0] == np.arange(r.shape[0]) * 6).all()
    cum = np.concatenate(([0], np.cumsum(lens)))
    for a, b in r:
        assert cum[b] - cum[a] <= 200 < cum[b + 1] - cum[a]
