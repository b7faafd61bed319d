"""Two independently written CPU restatements of the decoder path (numpy: oracle/model_ref.py, torch.nn:
oracle/model_ref_torch.py) must agree: integer outputs bit for bit in float64 (no near ties at that resolution),
and in float32 everywhere except at steps the numpy oracle itself flags as near ties.

The NN path is PARITY UNPINNED against TensorFlow (not installable here); this is the second opinion on the
AttentionWrapper step, the dynamic_decode stop rule, _beam_search_step and gather_tree."""
import numpy as np
import pytest
import torch

from oracle import model_ref as mr
from oracle import model_ref_torch as mt

N = 1000


@pytest.fixture(scope="module")
def encoded():
    w = mr.init_weights(22, random_bias=True)
    x = mr.synth_chunks(np.random.default_rng(77), N)
    enc, mask = mr.encode_input(w, x, "joint")
    return w, enc, mask


@pytest.mark.parametrize("W", [1, 5])
def test_beam_search_oracles_agree_float64(encoded, W):
    w, enc, mask = encoded
    L = 34
    a = mr.beam_search(w, enc, mask, W, L, dtype=np.float64, return_all=True)
    b = mt.beam_search(w, enc, mask, W, L, dtype=torch.float64, return_all=True)
    assert a[0].shape == b[0].shape, "dynamic_decode executed a different number of steps"
    assert np.array_equal(a[2], b[2]), "step ids differ"
    assert np.array_equal(a[3], b[3]), "parent ids differ"
    assert np.array_equal(a[0], b[0]), "gather_tree outputs differ"
    with np.errstate(invalid="ignore"):
        both_inf = np.isneginf(a[1]) & np.isneginf(b[1])
    np.testing.assert_allclose(np.where(both_inf, 0, a[1]), np.where(both_inf, 0, b[1]), rtol=1e-9, atol=1e-9)


def test_greedy_oracles_agree_float64(encoded):
    w, enc, mask = encoded
    ai, al = mr.greedy_search(w, enc, mask, 34, dtype=np.float64)
    bi, bl = mt.greedy_search(w, enc, mask, 34, dtype=torch.float64)
    assert np.array_equal(ai, bi)
    np.testing.assert_allclose(al, bl, rtol=1e-9, atol=1e-9)


def test_early_stop_rule_agrees(encoded):
    """End-token bias: decoding stops early, at the same step, with the same lengths-driven gather_tree."""
    w, enc, mask = encoded
    w = dict(w)
    b = w["decoder/fc/bias"].copy(); b[mr.TOKEN_END] += 2.5
    w["decoder/fc/bias"] = b
    a = mr.beam_search(w, enc[:200], mask[:200], 5, 34, dtype=np.float64, return_all=True)
    t = mt.beam_search(w, enc[:200], mask[:200], 5, 34, dtype=torch.float64, return_all=True)
    assert a[0].shape[1] < 33 and a[0].shape == t[0].shape
    assert np.array_equal(a[0], t[0]) and np.array_equal(a[2], t[2]) and np.array_equal(a[3], t[3])
    gi, _ = mr.greedy_search(w, enc[:200], mask[:200], 34, dtype=np.float64)
    ti, _ = mt.greedy_search(w, enc[:200], mask[:200], 34, dtype=torch.float64)
    assert gi.shape[1] < 33 and np.array_equal(gi, ti)


def test_decoder_depth_2_agrees():
    w = mr.init_weights(31, encoder_depth=1, decoder_depth=2, random_bias=True)
    x = mr.synth_chunks(np.random.default_rng(5), 64)
    enc, mask = mr.encode_input(w, x, "joint", encoder_depth=1)
    a = mr.beam_search(w, enc, mask, 5, 20, decoder_depth=2, dtype=np.float64, return_all=True)
    t = mt.beam_search(w, enc, mask, 5, 20, decoder_depth=2, dtype=torch.float64, return_all=True)
    assert np.array_equal(a[0], t[0]) and np.array_equal(a[2], t[2]) and np.array_equal(a[3], t[3])


def test_float32_differences_are_near_ties(encoded):
    """In float32 the two summation orders may split near ties: every row that differs must do so first at a step
    whose top-(W+1) margin (numpy oracle) is below 1e-3, and there are few of them."""
    w, enc, mask = encoded
    n, W, L = 400, 5, 34
    pa, sa, ia, para, margin = mr.beam_search(w, enc[:n], mask[:n], W, L, return_all=True, return_margins=True)
    pb, sb, ib, parb = mt.beam_search(w, enc[:n], mask[:n], W, L, return_all=True)
    T = min(ia.shape[1], ib.shape[1])
    differing = 0
    for r in range(n):
        d = np.flatnonzero((ia[r, :T] != ib[r, :T]).any(axis=1) | (para[r, :T] != parb[r, :T]).any(axis=1))
        if d.size:
            differing += 1
            assert margin[r, d[0]] < 1e-3, f"row {r} diverges at step {d[0]} with margin {margin[r, d[0]]}"
        else:
            assert np.array_equal(pa[r, :T], pb[r, :T])
    assert differing <= n // 50, differing


@pytest.mark.parametrize("rnn_type", ["gru", "bigru", "lstm"])
def test_rnn_type_variants_agree(rnn_type):
    """GRU cells (Keras reset_after=True vs torch.nn.GRUCell) and unidirectional encoders (basecaller.py:25-46, 86-89)."""
    w = mr.init_weights(9, rnn_type=rnn_type, random_bias=True)
    x = mr.synth_chunks(np.random.default_rng(6), 48)
    enc, mask = mr.encode_input(w, x, "joint")
    assert enc.shape == (48, 230, 256 if "bi" in rnn_type else 128)
    a = mr.beam_search(w, enc, mask, 5, 20, dtype=np.float64, return_all=True)
    t = mt.beam_search(w, enc, mask, 5, 20, dtype=torch.float64, return_all=True)
    assert np.array_equal(a[0], t[0]) and np.array_equal(a[2], t[2]) and np.array_equal(a[3], t[3])
    gi, gl = mr.greedy_search(w, enc, mask, 20, dtype=np.float64)
    ti, tl = mt.greedy_search(w, enc, mask, 20, dtype=torch.float64)
    assert np.array_equal(gi, ti)
    np.testing.assert_allclose(gl, tl, rtol=1e-9, atol=1e-9)


def test_gru_encoder_matches_torch_gru():
    """Encoder GRU layers against torch.nn.GRU (bidirectional, state hand-off as Encoder.call does it)."""
    u = 32
    w = mr.init_weights(5, enc_units=u, dec_units=u, rnn_type="bigru", random_bias=True)
    x = np.random.default_rng(0).normal(size=(6, 37, 1)).astype(np.float32)
    out, states = mr.encoder(x, w, "encoder_raw", 2, u)
    y = torch.from_numpy(x)
    h0 = torch.zeros(2, 6, u)
    perm = np.concatenate([np.arange(u, 2 * u), np.arange(0, u), np.arange(2 * u, 3 * u)])
    for l in range(2):
        gru = torch.nn.GRU(y.shape[-1], u, batch_first=True, bidirectional=True)
        with torch.no_grad():
            for d, suf in (("forward", ""), ("backward", "_reverse")):
                getattr(gru, "weight_ih_l0" + suf).copy_(torch.from_numpy(w[f"encoder_raw/layer{l}/{d}/kernel"].T[perm].copy()))
                getattr(gru, "weight_hh_l0" + suf).copy_(torch.from_numpy(w[f"encoder_raw/layer{l}/{d}/recurrent_kernel"].T[perm].copy()))
                getattr(gru, "bias_ih_l0" + suf).copy_(torch.from_numpy(w[f"encoder_raw/layer{l}/{d}/bias"][0][perm].copy()))
                getattr(gru, "bias_hh_l0" + suf).copy_(torch.from_numpy(w[f"encoder_raw/layer{l}/{d}/bias"][1][perm].copy()))
            y, h0 = gru(y, h0)
    np.testing.assert_allclose(out, y.numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(states[0], h0[0].numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(states[2], h0[1].numpy(), rtol=1e-4, atol=2e-6)
