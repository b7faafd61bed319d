"""rnn_type variants of the reference constructor (basecaller.py:25-46, 86-89, 195): GRU cells and unidirectional
encoders, against the CPU oracle (itself cross-checked against torch.nn.GRU / GRUCell in tests/test_oracle_second_opinion.py)."""
import numpy as np
import pytest

from oracle import model_ref as mr
from oracle.parity import check_beam

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 2e-5


def make(kind, rnn_type, w, depth=2, dec_depth=1):
    import ravvent_basecaller_b200 as rb
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, kind, 0., encoder_depth=depth, decoder_depth=dec_depth, rnn_type=rnn_type)
    bc.load_weights(w)
    return bc


@pytest.mark.parametrize("rnn_type", ["bigru", "gru", "lstm"])
@pytest.mark.parametrize("kind", ["joint", "raw"])
def test_rnn_type_matches_oracle(rnn_type, kind):
    w = mr.init_weights(13, rnn_type=rnn_type, random_bias=True)
    n, L = 72, 16
    raw, ev = mr.synth_chunks(np.random.default_rng(5), n)
    x = (raw, ev) if kind == "joint" else raw
    bc = make(kind, rnn_type, w)
    enc, mask = bc._encode_input(x)
    renc, rmask = mr.encode_input(w, x, kind)
    assert enc.shape == renc.shape and np.array_equal(mask, rmask)
    np.testing.assert_allclose(enc, renc, rtol=RTOL, atol=ATOL)
    for W in (1, 5):
        got = bc.beam_search_prediction(x, W, L, return_all_beams=True)
        check_beam(got, w, renc, rmask, W, L, atol=1e-4, label=f" {rnn_type} {kind} W={W}")
    ids, logits = bc.greedy_search_prediction(x, L)
    rid, rlog = mr.greedy_search(w, renc, rmask, L)
    assert ids.shape == rid.shape
    flips = 0
    for b in range(n):
        d = np.flatnonzero(ids[b] != rid[b])
        upto = ids.shape[1] if d.size == 0 else d[0] + 1
        np.testing.assert_allclose(logits[b, :upto], rlog[b, :upto], rtol=RTOL, atol=1e-4)
        if d.size:
            top2 = np.sort(rlog[b, d[0]])[-2:]
            assert top2[1] - top2[0] < 1e-3
            flips += 1
    assert flips <= max(1, n // 50)


def test_gru_depths_and_random_init():
    """(3, 2) with GRU cells, and the seeded Keras-default initialiser of the product equals the oracle's."""
    from ravvent_basecaller_b200 import weights
    for rt in ("bigru", "gru", "lstm"):
        a, b = weights.random_weights(22, rnn_type=rt), mr.init_weights(22, rnn_type=rt)
        assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    w = mr.init_weights(17, encoder_depth=3, decoder_depth=2, rnn_type="bigru", random_bias=True)
    x = mr.synth_chunks(np.random.default_rng(7), 40)
    bc = make("joint", "bigru", w, depth=3, dec_depth=2)
    renc, rmask = mr.encode_input(w, x, "joint", encoder_depth=3)
    got = bc.beam_search_prediction(x, 5, 14, return_all_beams=True)
    check_beam(got, w, renc, rmask, 5, 14, decoder_depth=2, label=" bigru (3,2)")
