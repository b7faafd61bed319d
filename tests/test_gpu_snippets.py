"""GPU snippet builder (K1 -> rvb_build_snippets) against the reference's prepare_snippets output
(golden fixtures) and the oracle.  Window counts / padding pattern exact; values within 1 float32 ulp
(sklearn accumulates the scaler moments in a different order)."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import event_ref, snippet_ref

pytestmark = pytest.mark.gpu


def _ulp_close(a, b, ulps=1):
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)))


@pytest.mark.parametrize("k", [0, 1, 2])
def test_matches_reference_golden(k):
    from ravvent_basecaller_b200 import data_loader as dl
    g = np.load(GOLDEN / "snippets_golden.npz")
    raw = g[f"raw_{k}"].astype(np.int32)
    lab0, lab1, stride = (int(v) for v in g[f"par_{k}"])
    rs, es = dl.load_data_from_signal(raw, lab0, lab1, stride)
    rs, es = rs.cpu().numpy(), es.cpu().numpy()
    assert rs.shape == g[f"raw_snips_{k}"].shape and es.shape == g[f"event_snips_{k}"].shape
    assert np.array_equal(rs == 0, g[f"raw_snips_{k}"] == 0)
    assert np.array_equal(es == 0, g[f"event_snips_{k}"] == 0)
    assert _ulp_close(rs, g[f"raw_snips_{k}"]) and _ulp_close(es, g[f"event_snips_{k}"])


def test_against_oracle_on_a_full_read_and_feeds_the_event_model():
    """BASELINE configs[1]: GPU event detection feeding the event encoder, beam 1."""
    import ravvent_basecaller_b200 as rb
    from ravvent_basecaller_b200 import data_loader as dl
    from oracle import model_ref as mr
    raw = event_ref.synth_read(np.random.default_rng(42), 60000)
    rs, es = dl.load_data_from_signal(raw, stride=6)
    ev = event_ref.detect_events(raw, 6, 9)
    ref = snippet_ref.build_snippets(raw, ev.start, ev.length, ev.mean, ev.stdv, 0, raw.size, 6)
    assert rs.shape == ref["raw"].shape and es.shape == ref["event"].shape and rs.shape[0] > 900
    assert _ulp_close(rs.cpu().numpy(), ref["raw"]) and _ulp_close(es.cpu().numpy(), ref["event"])
    w = mr.init_weights(22)
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, 'event', 0.)
    bc.load_weights(w)
    from oracle.parity import check_beam
    got = bc.beam_search_prediction(es[:64], 1, 12, return_all_beams=True)
    # the GPU snippets may differ from the oracle's by 1 ulp (see _ulp_close): feed the oracle the GPU's own snippets
    enc, mask = mr.encode_input(w, es[:64].cpu().numpy(), "event")
    check_beam([g.cpu().numpy() for g in got], w, enc, mask, 1, 12, label=" event model on GPU-built snippets")


def test_short_and_empty_reads():
    from ravvent_basecaller_b200 import data_loader as dl
    for n in (0, 10, 150):
        raw = event_ref.synth_read(np.random.default_rng(n), n) if n else np.zeros(0, np.int32)
        rs, es = dl.load_data_from_signal(raw)
        assert rs.shape[0] == 0 and es.shape[0] == 0


@pytest.mark.parametrize("k", [0, 1, 2])
def test_signal_label_files_match_reference_loader(k, tmp_path):
    """§8f-3: the on-disk entry point (Chiron .signal / .label) against the reference's own
    load_data_from_single_signal_label, including the target token rows."""
    from ravvent_basecaller_b200 import data_loader as dl
    g = np.load(GOLDEN / "snippets_golden.npz")
    raw = g[f"raw_{k}"].astype(np.int64)
    stride = int(g[f"par_{k}"][2])
    sp, lp = tmp_path / "r.signal", tmp_path / "r.label"
    np.savetxt(sp, raw.reshape(1, -1), fmt="%d")
    with open(lp, "w") as f:
        for (a, b), c in zip(g[f"label_ranges_{k}"], g[f"label_syms_{k}"]):
            f.write(f"{a} {b} {chr(c)}\n")
    rs, es, tk = dl.load_data_from_single_signal_label(str(sp), str(lp), stride)
    assert rs.shape == g[f"raw_snips_{k}"].shape and es.shape == g[f"event_snips_{k}"].shape
    assert _ulp_close(rs, g[f"raw_snips_{k}"]) and _ulp_close(es, g[f"event_snips_{k}"])
    assert tk.dtype == np.int64 and np.array_equal(tk, g[f"tokens_{k}"])


def test_batched_builder_equals_the_per_read_one():
    """rvb_build_snippets_batch (one launch set, one host round trip for all reads) against load_data_from_signal per read:
    bit-identical snippets, ranges and per-read counts -- including a read too short to yield a window and an empty one."""
    import torch
    from ravvent_basecaller_b200 import data_loader as dl
    rng = np.random.default_rng(11)
    lens = [6000, 0, 300, 12345, 47, 9000]
    reads = []
    for n in lens:
        steps = rng.integers(0, 2, size=n).cumsum() % 7
        reads.append((500 + 40 * steps + rng.integers(-6, 7, size=n)).astype(np.int32))
    sig = np.concatenate(reads) if reads else np.zeros(0, np.int32)
    offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    raw_b, ev_b, soff, rng_b = dl.load_data_from_signals(sig, offs, stride=6, return_ranges=True)
    soff = soff.cpu().numpy()
    assert soff[0] == 0 and soff[-1] == ev_b.shape[0] == raw_b.shape[0]
    n_nonempty = 0
    for r, read in enumerate(reads):
        if read.size == 0:
            assert soff[r + 1] == soff[r]
            continue
        raw_1, ev_1, rng_1 = dl.load_data_from_signal(read, stride=6, return_ranges=True)
        assert soff[r + 1] - soff[r] == ev_1.shape[0], (r, soff[r + 1] - soff[r], ev_1.shape[0])
        sl = slice(int(soff[r]), int(soff[r + 1]))
        assert torch.equal(ev_b[sl], ev_1) and torch.equal(raw_b[sl], raw_1) and torch.equal(rng_b[sl], rng_1)
        n_nonempty += ev_1.shape[0] > 0
    assert n_nonempty >= 3
    # event-only form: no raw output
    none_raw, ev_only, soff2 = dl.load_data_from_signals(sig, offs, stride=6, with_raw=False)
    assert none_raw is None and torch.equal(ev_only, ev_b) and np.array_equal(soff2.cpu().numpy(), soff)
