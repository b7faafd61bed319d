"""CPU oracle: Ravvent encoders, Luong-attention decoder, greedy / beam search.
TEST INFRASTRUCTURE ONLY (also the timed CPU baseline of bench.py).

PARITY UNPINNED.  The reference builds this path from tf.keras and
tensorflow_addons.seq2seq layers (basecaller.py:20-30, 86-94, 117-122, 131-134,
300-313, 322-329); neither package is vendored under /root/reference nor
installable offline ("tensorflow >= 2.7", README.md:20; tensorflow_addons is
imported at basecaller.py:3 but listed nowhere, unpinned).  The reference ships
no tests, weights or recorded outputs for this path.  What follows restates the
published algorithms of those layers, one small function
per semantic, and is cross-checked against torch.nn.LSTM for the recurrent
cells (tests/test_oracle_model.py).

Reference wiring restated here (file:line under /root/reference):
  Encoder.call, state hand-off between layers          basecaller.py:48-59
  Decoder cell graph                                   basecaller.py:83-94, 117-134
  _prepare_input_mask / utils.input_mask               basecaller.py:384-393, utils.py:26-32
  _encode_input (raw / event / joint concat on time)   basecaller.py:395-416
  greedy_search_prediction                             basecaller.py:317-330
  beam_search_prediction (returns beam slot 0)         basecaller.py:296-315
  tokens_to_nuc_sequences                              basecaller.py:289-294
  vocabulary {'':0,'^':1,'$':2,a:3,c:4,g:5,t:6}        data_loader.py:20-26
"""
from __future__ import annotations

import numpy as np

VOCAB = {'': 0, '^': 1, '$': 2, 'a': 3, 'c': 4, 'g': 5, 't': 6}
INDEX_WORD = {v: k for k, v in VOCAB.items()}
TOKEN_PAD, TOKEN_END, TOKEN_START = 0, 1, 2
MAX_RAW_LEN, MAX_EVENT_LEN = 200, 30


# --------------------------------------------------------------------------
# weights (Appendix A.1 / A.6): Keras default initialisers, seeded numpy RNG
# --------------------------------------------------------------------------
def _glorot(rng, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def _orthogonal(rng, rows, cols):
    a = rng.normal(size=(max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return np.ascontiguousarray(q[:rows, :cols]).astype(np.float32)


def _lstm_weights(rng, n_in, units):
    b = np.zeros(4 * units, dtype=np.float32)
    b[units:2 * units] = 1.0                      # unit_forget_bias
    return {"kernel": _glorot(rng, n_in, 4 * units),
            "recurrent_kernel": _orthogonal(rng, units, 4 * units),
            "bias": b}


def _gru_weights(rng, n_in, units):
    """Keras GRUCell defaults (reset_after=True): kernel [in,3u], recurrent_kernel [u,3u], bias [2,3u] = (input, recurrent),
    gate blocks z, r, h."""
    return {"kernel": _glorot(rng, n_in, 3 * units),
            "recurrent_kernel": _orthogonal(rng, units, 3 * units),
            "bias": np.zeros((2, 3 * units), dtype=np.float32)}


def init_weights(seed=22, enc_units=128, dec_units=128, encoder_depth=2, decoder_depth=1,
                 vocab_size=7, random_bias=False, rnn_type="bilstm"):
    """Flat dict name -> float32 array (the .npz interchange layout of
    Basecaller.load_weights).  ``random_bias`` perturbs biases so tests see a
    non-trivial bias path.  rnn_type in {'bilstm', 'lstm', 'bigru', 'gru'} as in basecaller.py:25-46, 86-89:
    'bi' -> Bidirectional encoders (forward + backward weights, 2*enc_units outputs), otherwise forward only;
    'lstm' / 'gru' picks the cell of the encoders AND of the decoder (basecaller.py:195 strips the 'bi')."""
    rng = np.random.default_rng(seed)
    w = {}
    bi = "bi" in rnn_type
    cell_w = _lstm_weights if "lstm" in rnn_type else _gru_weights
    enc_out = (2 if bi else 1) * enc_units
    for enc, feat in (("encoder_raw", 1), ("encoder_event", 5)):
        for l in range(encoder_depth):
            n_in = feat if l == 0 else enc_out
            for d in (("forward", "backward") if bi else ("forward",)):
                for k, v in cell_w(rng, n_in, enc_units).items():
                    w[f"{enc}/layer{l}/{d}/{k}"] = v
    for j in range(decoder_depth):
        n_in = vocab_size + dec_units if j == 0 else dec_units
        for k, v in cell_w(rng, n_in, dec_units).items():
            w[f"decoder/cell{j}/{k}"] = v
    w["decoder/memory_layer/kernel"] = _glorot(rng, enc_out, dec_units)
    w["decoder/attention_layer/kernel"] = _glorot(rng, dec_units + enc_out, dec_units)
    w["decoder/fc/kernel"] = _glorot(rng, dec_units, vocab_size)
    w["decoder/fc/bias"] = np.zeros(vocab_size, dtype=np.float32)
    if random_bias:
        for k in w:
            if k.endswith("bias"):
                w[k] = (w[k] + rng.normal(0, 0.1, size=w[k].shape)).astype(np.float32)
    return w


# --------------------------------------------------------------------------
# cells and layers
# --------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, h, c, kernel, recurrent_kernel, bias):
    """Keras LSTMCell step (A.1): gate blocks i, f, g, o; state order [h, c]."""
    u = h.shape[-1]
    z = x @ kernel + h @ recurrent_kernel + bias
    i, f, g, o = z[:, :u], z[:, u:2 * u], z[:, 2 * u:3 * u], z[:, 3 * u:]
    c2 = sigmoid(f) * c + sigmoid(i) * np.tanh(g)
    h2 = sigmoid(o) * np.tanh(c2)
    return h2, c2


def gru_cell(x, h, kernel, recurrent_kernel, bias):
    """Keras GRUCell step, reset_after=True (the TF2 default; [Keras-recall], same formula as torch.nn.GRUCell):
    gate blocks z, r, h; the reset gate multiplies the recurrent part INCLUDING its bias."""
    u = h.shape[-1]
    mx = x @ kernel + bias[0]
    mh = h @ recurrent_kernel + bias[1]
    z = sigmoid(mx[:, :u] + mh[:, :u])
    r = sigmoid(mx[:, u:2 * u] + mh[:, u:2 * u])
    hh = np.tanh(mx[:, 2 * u:] + r * mh[:, 2 * u:])
    return z * h + (1.0 - z) * hh


def rnn_cell(x, h, c, kernel, recurrent_kernel, bias):
    """LSTMCell or GRUCell by the shape of the bias ([4u] vs [2,3u]); a GRU's state is h alone (c mirrors it)."""
    if bias.ndim == 2:
        h2 = gru_cell(x, h, kernel, recurrent_kernel, bias)
        return h2, h2
    return lstm_cell(x, h, c, kernel, recurrent_kernel, bias)


def rnn_direction(x, cell_w, h0, c0, reverse):
    """One Keras RNN(cell, return_sequences, return_state), optionally
    go_backwards with the output re-reversed as Bidirectional does (A.2)."""
    B, T, _ = x.shape
    h, c = h0, c0
    ys = np.empty((B, T, h0.shape[-1]), dtype=x.dtype)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h, c = rnn_cell(x[:, t], h, c, cell_w["kernel"], cell_w["recurrent_kernel"], cell_w["bias"])
        ys[:, t] = h
    return ys, h, c


def cast_weights(w, dtype):
    return {k: v.astype(dtype) for k, v in w.items()}


def encoder(x, w, prefix, depth, units):
    """Encoder.call (basecaller.py:48-59): depth x BiLSTM, layer i's final
    [h_f, c_f, h_b, c_b] seed layer i+1; no mask is applied."""
    B = x.shape[0]
    zeros = np.zeros((B, units), dtype=x.dtype)
    states = [zeros, zeros, zeros, zeros]
    out = x
    for l in range(depth):
        fw = {k: w[f"{prefix}/layer{l}/forward/{k}"] for k in ("kernel", "recurrent_kernel", "bias")}
        yf, hf, cf = rnn_direction(out, fw, states[0], states[1], reverse=False)
        if f"{prefix}/layer{l}/backward/kernel" not in w:          # unidirectional: RNN(cell) without Bidirectional
            out, states = yf, [hf, cf, states[2], states[3]]
            continue
        bw = {k: w[f"{prefix}/layer{l}/backward/{k}"] for k in ("kernel", "recurrent_kernel", "bias")}
        yb, hb, cb = rnn_direction(out, bw, states[2], states[3], reverse=True)
        out = np.concatenate([yf, yb], axis=-1)
        states = [hf, cf, hb, cb]
    return out, states


def input_mask(x, padding_value=0.0):
    """utils.input_mask: True where no feature equals the padding value."""
    return np.all(x != padding_value, axis=-1)


def encode_input(w, input_data, input_data_type, enc_units=128, encoder_depth=2, padding_value=0.0,
                 dtype=np.float32):
    """Basecaller._encode_input -> (enc_output [B,Tm,2u], mask [B,Tm])."""
    wc = cast_weights(w, dtype)
    if input_data_type == "joint":
        raw, event = input_data
        raw, event = np.asarray(raw, dtype), np.asarray(event, dtype)
        er, _ = encoder(raw, wc, "encoder_raw", encoder_depth, enc_units)
        ee, _ = encoder(event, wc, "encoder_event", encoder_depth, enc_units)
        return (np.concatenate([er, ee], axis=1),
                np.concatenate([input_mask(raw, padding_value), input_mask(event, padding_value)], axis=-1))
    x = np.asarray(input_data, dtype)
    prefix = "encoder_raw" if input_data_type == "raw" else "encoder_event"
    out, _ = encoder(x, wc, prefix, encoder_depth, enc_units)
    return out, input_mask(x, padding_value)


# --------------------------------------------------------------------------
# attention decoder (A.3, A.3b)
# --------------------------------------------------------------------------
class DecoderState:
    """AttentionWrapperState restricted to what inference uses."""
    def __init__(self, cells, attention):
        self.cells = cells            # list of (h, c) per stacked LSTM cell
        self.attention = attention    # [R, dec_units]

    def gather(self, idx):
        return DecoderState([(h[idx], c[idx]) for h, c in self.cells], self.attention[idx])


def setup_memory(w, enc_output, mask, dtype):
    """LuongAttention.setup_memory: values = memory*mask, keys = values @ memory_layer."""
    values = enc_output.astype(dtype) * mask[..., None].astype(dtype)
    keys = values @ w["decoder/memory_layer/kernel"].astype(dtype)
    return keys, values


def zero_state(rows, dec_units, decoder_depth, dtype):
    z = lambda: np.zeros((rows, dec_units), dtype=dtype)
    return DecoderState([(z(), z()) for _ in range(decoder_depth)], z())


def softmax(x):
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=-1, keepdims=True)


def log_softmax(x):
    m = np.max(x, axis=-1, keepdims=True)
    s = x - m
    return s - np.log(np.sum(np.exp(s), axis=-1, keepdims=True))


def decoder_step(w, tokens, state, keys, values, mask, decoder_depth, vocab_size=7):
    """AttentionWrapper.call + fc: -> (logits [R,V], next state, alignments)."""
    dtype = keys.dtype
    x = np.concatenate([np.eye(vocab_size, dtype=dtype)[tokens], state.attention], axis=-1)
    new_cells = []
    for j in range(decoder_depth):
        h, c = state.cells[j]
        h2, c2 = rnn_cell(x, h, c, w[f"decoder/cell{j}/kernel"], w[f"decoder/cell{j}/recurrent_kernel"],
                          w[f"decoder/cell{j}/bias"])
        new_cells.append((h2, c2))
        x = h2
    query = x
    score = np.einsum("rd,rtd->rt", query, keys)
    with np.errstate(invalid="ignore"):
        score = np.where(mask, score, -np.inf).astype(dtype)
        align = softmax(score)
    context = np.einsum("rt,rtd->rd", align, values)
    attention = np.concatenate([query, context], axis=-1) @ w["decoder/attention_layer/kernel"]
    logits = attention @ w["decoder/fc/kernel"] + w["decoder/fc/bias"]
    return logits, DecoderState(new_cells, attention), align


# --------------------------------------------------------------------------
# greedy search (A.4)
# --------------------------------------------------------------------------
def greedy_search(w, enc_output, mask, max_output_len, dec_units=128, decoder_depth=1, dtype=np.float32,
                  full_length=False):
    """BasicDecoder + GreedyEmbeddingSampler under dynamic_decode.
    -> (sample_id [B,T] int32, logits [B,T,V]).  T = executed steps unless
    ``full_length`` (then all max_output_len-1 steps are run and returned)."""
    wc = cast_weights(w, dtype)
    B = enc_output.shape[0]
    S = int(max_output_len) - 1
    keys, values = setup_memory(wc, enc_output, mask, dtype)
    state = zero_state(B, dec_units, decoder_depth, dtype)
    tokens = np.full(B, TOKEN_START, dtype=np.int32)
    finished = np.zeros(B, dtype=bool) | (0 >= S)
    ids, logs = [], []
    t = 0
    while (full_length and t < S) or (not full_length and not finished.all()):
        logits, state, _ = decoder_step(wc, tokens, state, keys, values, mask, decoder_depth)
        tokens = np.argmax(logits, axis=-1).astype(np.int32)
        finished = finished | (tokens == TOKEN_END) | (t + 1 >= S)
        ids.append(tokens)
        logs.append(logits)
        t += 1
    if not ids:
        return np.zeros((B, 0), np.int32), np.zeros((B, 0, 7), dtype)
    return np.stack(ids, axis=1), np.stack(logs, axis=1)


# --------------------------------------------------------------------------
# beam search (A.5)
# --------------------------------------------------------------------------
def beam_step(step_log_probs, log_probs, finished, lengths, end_token=TOKEN_END, with_margin=False):
    """_beam_search_step on already log-softmaxed rows.
    step_log_probs [B,W,V] ; log_probs [B,W] ; finished [B,W] bool ; lengths [B,W] int64.
    -> scores, word, parent, next_log_probs, next_finished, next_lengths."""
    B, W, V = step_log_probs.shape
    dt = step_log_probs.dtype
    fin_row = np.full(V, np.finfo(dt).min, dtype=dt)
    fin_row[end_token] = 0.0
    slp = np.where(finished[..., None], fin_row[None, None, :], step_log_probs)
    total = (log_probs[..., None] + slp).reshape(B, W * V)
    idx = np.argsort(-total, axis=-1, kind="stable")[:, :W]      # top_k: descending, ties -> lower index
    scores = np.take_along_axis(total, idx, axis=-1)
    word = (idx % V).astype(np.int32)
    parent = (idx // V).astype(np.int32)
    prev_fin = np.take_along_axis(finished, parent, axis=-1)
    next_finished = prev_fin | (word == end_token)
    next_lengths = np.take_along_axis(lengths, parent, axis=-1) + (~prev_fin).astype(np.int64)
    if with_margin:
        return scores, word, parent, scores.copy(), next_finished, next_lengths, topk_margin(total, W)
    return scores, word, parent, scores.copy(), next_finished, next_lengths


def topk_margin(total, W):
    """Smallest gap between neighbours among the W+1 best candidates of each row: how close the step was to a
    different selection OR a different slot order.  Pairs of exactly equal values are resolved by index in every
    implementation (top_k is stable) and pairs at the -inf / dtype.min level are not real candidates: both are
    skipped, except finite exact ties, which count as margin 0 (another summation order may split them)."""
    srt = -np.sort(-total.astype(np.float64), axis=-1)[:, :W + 1]
    a, b = srt[:, :-1], srt[:, 1:]
    with np.errstate(invalid="ignore"):
        gap = a - b
    real = np.isfinite(a) & np.isfinite(b) & (b > -1e30)
    gap = np.where(real, gap, np.inf)
    return gap.min(axis=-1)


def gather_tree(step_ids, parent_ids, max_sequence_lengths, end_token=TOKEN_END):
    """tfa.seq2seq.gather_tree on time-major [T,B,W] int32 arrays."""
    T, B, W = step_ids.shape
    out = np.full_like(step_ids, end_token)
    for b in range(B):
        L = min(T, int(max_sequence_lengths[b]))
        if L <= 0:
            continue
        for k in range(W):
            parent = k
            for level in range(L - 1, -1, -1):
                out[level, b, k] = step_ids[level, b, parent]
                parent = parent_ids[level, b, parent]
            done = False
            for t in range(L):
                if done:
                    out[t, b, k] = end_token
                elif out[t, b, k] == end_token:
                    done = True
    return out


def beam_search(w, enc_output, mask, beam_width, max_output_len, dec_units=128, decoder_depth=1,
                dtype=np.float32, full_length=False, return_all=False, return_margins=False):
    """BeamSearchDecoder under dynamic_decode + finalize.
    -> (predicted_ids[:, :, 0] [B,T] int32, scores[:, :, 0] [B,T]).
    return_all: all beams + the raw per-step ids / parents; return_margins adds margins [B,T] (topk_margin per step)."""
    wc = cast_weights(w, dtype)
    B, Tm, _ = enc_output.shape
    W = int(beam_width)
    S = int(max_output_len) - 1
    keys, values = setup_memory(wc, enc_output, mask, dtype)
    rep = np.repeat(np.arange(B), W)                       # tile_batch: row b*W+k
    keys, values, maskt = keys[rep], values[rep], mask[rep]
    state = zero_state(B * W, dec_units, decoder_depth, dtype)
    tokens = np.full(B * W, TOKEN_START, dtype=np.int32)
    log_probs = np.full((B, W), -np.inf, dtype=dtype)
    log_probs[:, 0] = 0.0
    finished = np.ones((B, W), dtype=bool)
    finished[:, 0] = False
    lengths = np.zeros((B, W), dtype=np.int64)
    all_done = bool(0 >= S)
    s_scores, s_ids, s_par, s_margin = [], [], [], []
    t = 0
    while (full_length and t < S) or (not full_length and not all_done):
        logits, state, _ = decoder_step(wc, tokens, state, keys, values, maskt, decoder_depth)
        slp = log_softmax(logits.reshape(B, W, -1))
        scores, word, parent, log_probs, finished, lengths, margin = beam_step(slp, log_probs, finished, lengths, with_margin=True)
        s_margin.append(margin)
        flat_parent = (np.arange(B)[:, None] * W + parent).reshape(-1)
        state = state.gather(flat_parent)
        tokens = word.reshape(-1)
        s_scores.append(scores); s_ids.append(word); s_par.append(parent)
        t += 1
        all_done = bool(finished.all()) or (t >= S)
    if not s_ids:
        return np.zeros((B, 0), np.int32), np.zeros((B, 0), dtype)
    step_ids, par_ids = np.stack(s_ids), np.stack(s_par)                 # [T,B,W]
    pred = gather_tree(step_ids, par_ids, lengths.max(axis=1).astype(np.int32))
    pred = np.transpose(pred, (1, 0, 2))
    sc = np.transpose(np.stack(s_scores), (1, 0, 2))
    extra = (np.stack(s_margin, axis=1),) if return_margins else ()
    if return_all:
        return (pred, sc, np.transpose(step_ids, (1, 0, 2)), np.transpose(par_ids, (1, 0, 2))) + extra
    return (pred[:, :, 0], sc[:, :, 0]) + extra


def tokens_to_nuc_sequences(tokens):
    """basecaller.py:289-294: ids -> text, drop ' ', '^', '$', upper-case."""
    out = []
    for row in np.asarray(tokens):
        txt = " ".join(INDEX_WORD[int(t)] for t in row if int(t) in INDEX_WORD)
        out.append(txt.replace(" ", "").replace("^", "").replace("$", "").upper())
    return out


def beam_scores_to_probs(beam_scores):
    """utils.calc_prob_logits_beam_search_scores (utils.py:123-128)."""
    s = np.asarray(beam_scores)
    prev = np.zeros_like(s)
    prev[..., 1:] = s[..., :-1]
    return np.exp(s - prev)


def masked_accuracy(y_true, y_pred, omit_vals):
    """utils.masked_accuracy (utils.py:15-24): matches / positions whose target is not in omit_vals."""
    total = count = 0
    for t, p in zip(np.asarray(y_true).ravel().tolist(), np.asarray(y_pred).ravel().tolist()):
        if t in [int(v) for v in omit_vals]:
            continue
        total += 1
        count += int(t == p)
    return count / total


def val_step(w, enc_output, mask, target_tokens, decoder_depth=1):
    """Basecaller._val_step (basecaller.py:267-279): greedy decode to the target length, zero-pad what
    dynamic_decode did not produce, masked mean cross-entropy (padding target positions excluded,
    basecaller.py:209-218) and masked accuracy (start / end targets excluded -- padding is NOT, as in the reference)."""
    target_tokens = np.asarray(target_tokens)
    L = target_tokens.shape[1]
    ids, logits = greedy_search(w, enc_output, mask, L, decoder_depth=decoder_depth)
    pad = L - 1 - ids.shape[1]
    logits = np.pad(logits.astype(np.float64), [(0, 0), (0, pad), (0, 0)])
    ids = np.pad(ids, [(0, 0), (0, pad)])
    real = target_tokens[:, 1:]
    num = den = 0.0
    for b in range(real.shape[0]):
        for t in range(real.shape[1]):
            if real[b, t] == TOKEN_PAD:
                continue
            z = logits[b, t]
            num += np.log(np.sum(np.exp(z - z.max()))) + z.max() - z[real[b, t]]
            den += 1.0
    return {"loss": num / den, "acc": masked_accuracy(real, ids, [TOKEN_START, TOKEN_END])}


def called_bases(ids):
    """Bases (tokens 3..6) before the first end token, per row."""
    ids = np.asarray(ids)
    is_end = ids == TOKEN_END
    first_end = np.where(is_end.any(axis=1), is_end.argmax(axis=1), ids.shape[1])
    before = np.arange(ids.shape[1])[None, :] < first_end[:, None]
    return ((ids >= 3) & (ids <= 6) & before).sum(axis=1)


def synth_chunks(rng, n, with_event=True):
    """Direct chunk generator: N(0,1) values, random
    valid length, zero tail, no exact zeros inside the valid part."""
    raw = rng.normal(size=(n, MAX_RAW_LEN, 1)).astype(np.float32)
    raw[raw == 0] = 1e-3
    rl = rng.integers(159, 196, size=n)
    raw[np.arange(MAX_RAW_LEN)[None, :] >= rl[:, None]] = 0.0
    if not with_event:
        return raw
    ev = rng.normal(size=(n, MAX_EVENT_LEN, 5)).astype(np.float32)
    ev[ev == 0] = 1e-3
    el = rng.integers(16, 28, size=n)
    ev[np.arange(MAX_EVENT_LEN)[None, :] >= el[:, None]] = 0.0
    return raw, ev
