"""NN oracle self-consistency.  There are no reference vectors for this path
(tensorflow / tensorflow_addons are not installable: parity unpinned), so the
restatement is cross-checked against independent implementations / properties."""
import numpy as np
import pytest
import torch

from oracle import model_ref as mr


def _torch_bilstm_layer(x, w, prefix, l, h0c0):
    """Independent second opinion: torch.nn.LSTM has the same i,f,g,o gate order."""
    n_in = x.shape[-1]
    u = w[f"{prefix}/layer{l}/forward/recurrent_kernel"].shape[0]
    lstm = torch.nn.LSTM(n_in, u, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for d, suf in (("forward", ""), ("backward", "_reverse")):
            getattr(lstm, "weight_ih_l0" + suf).copy_(torch.from_numpy(w[f"{prefix}/layer{l}/{d}/kernel"].T.copy()))
            getattr(lstm, "weight_hh_l0" + suf).copy_(torch.from_numpy(w[f"{prefix}/layer{l}/{d}/recurrent_kernel"].T.copy()))
            getattr(lstm, "bias_ih_l0" + suf).copy_(torch.from_numpy(w[f"{prefix}/layer{l}/{d}/bias"]))
            getattr(lstm, "bias_hh_l0" + suf).zero_()
        y, (hn, cn) = lstm(torch.from_numpy(x), h0c0)
    return y.numpy(), (hn, cn)


@pytest.mark.parametrize("prefix,feat,T", [("encoder_raw", 1, 37), ("encoder_event", 5, 30)])
def test_encoder_matches_torch_lstm(prefix, feat, T):
    w = mr.init_weights(seed=5, enc_units=32, dec_units=32, random_bias=True)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(6, T, feat)).astype(np.float32)
    out, states = mr.encoder(x, w, prefix, 2, 32)
    h0c0 = (torch.zeros(2, 6, 32), torch.zeros(2, 6, 32))
    y = x
    for l in range(2):
        y, h0c0 = _torch_bilstm_layer(y, w, prefix, l, h0c0)   # state hand-off (basecaller.py:51-57)
    np.testing.assert_allclose(out, y, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(states[0], h0c0[0][0].numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(states[3], h0c0[1][1].numpy(), rtol=1e-4, atol=2e-6)


def test_state_handoff_changes_outputs():
    """Guard against dropping the layer-to-layer state hand-off."""
    w = mr.init_weights(seed=5, enc_units=16, dec_units=16)
    x = np.random.default_rng(1).normal(size=(3, 20, 1)).astype(np.float32)
    out, _ = mr.encoder(x, w, "encoder_raw", 2, 16)
    out1, st1 = mr.encoder(x, w, "encoder_raw", 1, 16)
    w2 = {k.replace("layer1", "layer0"): v for k, v in w.items() if "encoder_raw/layer1" in k}
    out_nohand, _ = mr.encoder(out1, w2, "encoder_raw", 1, 16)
    assert np.abs(out - out_nohand).max() > 1e-3


def test_fp32_vs_fp64_encoder_close():
    w = mr.init_weights(seed=22)
    raw, ev = mr.synth_chunks(np.random.default_rng(2), 4)
    o32, m32 = mr.encode_input(w, (raw, ev), "joint", dtype=np.float32)
    o64, m64 = mr.encode_input(w, (raw, ev), "joint", dtype=np.float64)
    assert o32.shape == (4, 230, 256) and m32.shape == (4, 230)
    assert np.array_equal(m32, m64)
    np.testing.assert_allclose(o32, o64, rtol=1e-3, atol=1e-5)


def _brute_topk(total, W):
    order = sorted(range(total.size), key=lambda i: (-total[i], i))
    return order[:W]


def test_beam_step_against_bruteforce_with_ties():
    rng = np.random.default_rng(3)
    B, W, V = 16, 5, 7
    slp = mr.log_softmax(rng.integers(-2, 3, size=(B, W, V)).astype(np.float32))   # many exact ties
    lp = np.sort(rng.normal(-3, 1, size=(B, W)).astype(np.float32))[:, ::-1].copy()
    lp[:, 3] = lp[:, 2]
    fin = rng.random((B, W)) < 0.3
    lens = rng.integers(0, 9, size=(B, W)).astype(np.int64)
    sc, word, par, nlp, nfin, nlen = mr.beam_step(slp, lp, fin, lens)
    for b in range(B):
        rows = slp[b].copy()
        for k in range(W):
            if fin[b, k]:
                rows[k] = np.finfo(np.float32).min
                rows[k, mr.TOKEN_END] = 0.0
        total = (lp[b][:, None] + rows).reshape(-1)
        idx = _brute_topk(total, W)
        assert [i % V for i in idx] == word[b].tolist()
        assert [i // V for i in idx] == par[b].tolist()
        assert np.array_equal(total[idx], sc[b])
        for k, i in enumerate(idx):
            p = i // V
            assert nfin[b, k] == (fin[b, p] or (i % V) == mr.TOKEN_END)
            assert nlen[b, k] == lens[b, p] + (0 if fin[b, p] else 1)


def test_gather_tree_known_answer():
    # T=4, B=1, W=2 ; hand-traced
    ids = np.array([[[3, 4]], [[5, 6]], [[1, 3]], [[4, 1]]], dtype=np.int32)
    par = np.array([[[0, 0]], [[1, 0]], [[0, 1]], [[1, 0]]], dtype=np.int32)
    out = mr.gather_tree(ids, par, np.array([4]))
    # beam0: level3 id 4 parent 1 -> level2 id 3 (slot1) parent 1 -> level1 id 6 (slot1) parent 0 -> level0 id 3
    assert out[:, 0, 0].tolist() == [3, 6, 3, 4]
    # beam1: level3 id 1 parent 0 -> level2 id 1 (slot0) parent 0 -> level1 id 5 parent 1 -> level0 id 4 ; then truncate after first end
    assert out[:, 0, 1].tolist() == [4, 5, 1, 1]
    out3 = mr.gather_tree(ids, par, np.array([3]))
    assert out3[3, 0].tolist() == [1, 1]


def test_greedy_equals_beam1_until_end_token():
    w = mr.init_weights(seed=22)
    raw, ev = mr.synth_chunks(np.random.default_rng(4), 6)
    enc, mask = mr.encode_input(w, (raw, ev), "joint")
    gid, glog = mr.greedy_search(w, enc, mask, 12, full_length=True)
    bid, bsc = mr.beam_search(w, enc, mask, 1, 12, full_length=True)
    for b in range(6):
        end = np.flatnonzero(gid[b] == mr.TOKEN_END)
        L = end[0] + 1 if end.size else gid.shape[1]
        assert np.array_equal(gid[b, :L], bid[b, :L])
        assert (bid[b, L:] == mr.TOKEN_END).all()
    # scores are cumulative log-probs of the chosen tokens
    lsm = mr.log_softmax(glog)
    chosen = np.take_along_axis(lsm, gid[..., None].astype(np.int64), axis=-1)[..., 0]
    for b in range(6):
        end = np.flatnonzero(gid[b] == mr.TOKEN_END)
        L = end[0] + 1 if end.size else gid.shape[1]
        np.testing.assert_allclose(np.cumsum(chosen[b, :L]), bsc[b, :L], rtol=1e-4, atol=1e-5)


def test_dynamic_length_is_prefix_of_full_length():
    w = mr.init_weights(seed=22)
    raw = mr.synth_chunks(np.random.default_rng(5), 5, with_event=False)
    enc, mask = mr.encode_input(w, raw, "raw")
    a_ids, a_sc = mr.beam_search(w, enc, mask, 5, 10)
    f_ids, f_sc = mr.beam_search(w, enc, mask, 5, 10, full_length=True)
    T = a_ids.shape[1]
    assert np.array_equal(a_ids, f_ids[:, :T]) and np.array_equal(a_sc, f_sc[:, :T])
    g_ids, g_log = mr.greedy_search(w, enc, mask, 10)
    gf_ids, gf_log = mr.greedy_search(w, enc, mask, 10, full_length=True)
    assert np.array_equal(g_ids, gf_ids[:, :g_ids.shape[1]])


def test_tokens_to_nuc_sequences():
    assert mr.tokens_to_nuc_sequences(np.array([[3, 4, 5, 6, 1, 1], [2, 6, 0, 3, 1, 4]])) == ["ACGT", "TAC"]
    assert mr.called_bases(np.array([[3, 4, 5, 6, 1, 1], [2, 6, 0, 3, 1, 4]])).tolist() == [4, 2]
