"""Second-opinion CPU oracle for the decoder path, written against torch.nn modules.
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (same reason as model_ref.py: no TensorFlow /
tensorflow_addons here).

Purpose: `model_ref.py` restates tfa.seq2seq's AttentionWrapper step, `dynamic_decode` stop rule,
`BeamSearchDecoder` and `gather_tree` in numpy.  This file restates the SAME published semantics
 a second time along a different path -- torch.nn.LSTMCell / Linear
modules, batched matrix products, a stable descending sort for top_k, a level-synchronous vectorised
gather_tree -- so that an error of transcription in one of them shows up as a disagreement
(tests/test_oracle_second_opinion.py).  Nothing is shared with model_ref.py except the weight
dictionary layout and the token constants.

Reference call sites restated (file:line under /root/reference):
  Decoder.__init__ / build_rnn_cell / build_attention_mechanism   basecaller.py:63-139
  greedy_search_prediction                                        basecaller.py:317-330
  beam_search_prediction                                          basecaller.py:296-315
"""
from __future__ import annotations

import numpy as np
import torch

END, START = 1, 2


class TorchDecoder(torch.nn.Module):
    """tf.keras StackedRNNCells([LSTMCell]) inside tfa AttentionWrapper(LuongAttention) + Dense(vocab)."""

    def __init__(self, w, decoder_depth=1, dtype=torch.float32, vocab=7):
        super().__init__()
        u = w["decoder/cell0/recurrent_kernel"].shape[0]
        self.u, self.vocab, self.dt = u, vocab, dtype
        self.cells = torch.nn.ModuleList()
        for j in range(decoder_depth):
            k, rk, bias = (w[f"decoder/cell{j}/{n}"] for n in ("kernel", "recurrent_kernel", "bias"))
            if bias.ndim == 2:
                # Keras GRUCell(reset_after=True) == torch.nn.GRUCell with the gate blocks reordered: Keras z, r, h -> torch r, z, n
                cell = torch.nn.GRUCell(k.shape[0], u, dtype=dtype)
                perm = np.concatenate([np.arange(u, 2 * u), np.arange(0, u), np.arange(2 * u, 3 * u)])
                with torch.no_grad():
                    cell.weight_ih.copy_(torch.as_tensor(k.T[perm].copy(), dtype=dtype))
                    cell.weight_hh.copy_(torch.as_tensor(rk.T[perm].copy(), dtype=dtype))
                    cell.bias_ih.copy_(torch.as_tensor(bias[0][perm].copy(), dtype=dtype))
                    cell.bias_hh.copy_(torch.as_tensor(bias[1][perm].copy(), dtype=dtype))
            else:
                cell = torch.nn.LSTMCell(k.shape[0], u, dtype=dtype)      # gate order i, f, g, o == Keras i, f, c, o
                with torch.no_grad():
                    cell.weight_ih.copy_(torch.as_tensor(k.T.copy(), dtype=dtype))
                    cell.weight_hh.copy_(torch.as_tensor(rk.T.copy(), dtype=dtype))
                    cell.bias_ih.copy_(torch.as_tensor(bias, dtype=dtype))
                    cell.bias_hh.zero_()
            self.cells.append(cell)

        def dense(name, bias=None):
            k = w[name]
            lin = torch.nn.Linear(k.shape[0], k.shape[1], bias=bias is not None, dtype=dtype)
            with torch.no_grad():
                lin.weight.copy_(torch.as_tensor(k.T.copy(), dtype=dtype))
                if bias is not None:
                    lin.bias.copy_(torch.as_tensor(w[bias], dtype=dtype))
            return lin

        self.memory_layer = dense("decoder/memory_layer/kernel")          # LuongAttention: Dense(units, use_bias=False)
        self.attention_layer = dense("decoder/attention_layer/kernel")    # AttentionWrapper attention_layer_size=dec_units
        self.fc = dense("decoder/fc/kernel", "decoder/fc/bias")

    @torch.no_grad()
    def setup_memory(self, memory, mask, repeat=1):
        """attention_mechanism.setup_memory(memory, memory_mask); `repeat` = tile_batch multiplier."""
        memory = torch.as_tensor(np.asarray(memory), dtype=self.dt)
        mask = torch.as_tensor(np.asarray(mask), dtype=torch.bool)
        self.values = (memory * mask.unsqueeze(-1).to(self.dt)).repeat_interleave(repeat, dim=0)
        self.keys = self.memory_layer(self.values)
        self.mask = mask.repeat_interleave(repeat, dim=0)

    def zero_state(self, rows):
        z = lambda: torch.zeros(rows, self.u, dtype=self.dt)
        return {"cells": [(z(), z()) for _ in self.cells], "attention": z()}

    @torch.no_grad()
    def step(self, tokens, state):
        """One AttentionWrapper call followed by the output layer -> (logits, next state)."""
        x = torch.cat([torch.nn.functional.one_hot(tokens.long(), self.vocab).to(self.dt), state["attention"]], dim=1)
        cells = []
        for cell, (h, c) in zip(self.cells, state["cells"]):
            if isinstance(cell, torch.nn.GRUCell):
                h = cell(x, h)
                c = h
            else:
                h, c = cell(x, (h, c))
            cells.append((h, c))
            x = h
        score = torch.bmm(self.keys, x.unsqueeze(2)).squeeze(2)
        score = score.masked_fill(~self.mask, float("-inf"))
        align = torch.softmax(score, dim=1)
        context = torch.bmm(align.unsqueeze(1), self.values).squeeze(1)
        attention = self.attention_layer(torch.cat([x, context], dim=1))
        return self.fc(attention), {"cells": cells, "attention": attention}


def greedy_search(w, enc_output, mask, max_output_len, decoder_depth=1, dtype=torch.float32):
    """BasicDecoder(GreedyEmbeddingSampler) under dynamic_decode(maximum_iterations=max_output_len-1, impute_finished=False)."""
    dec = TorchDecoder(w, decoder_depth, dtype)
    B = len(enc_output)
    dec.setup_memory(enc_output, mask)
    state = dec.zero_state(B)
    tokens = torch.full((B,), START, dtype=torch.int64)
    max_iter = int(max_output_len) - 1
    finished = torch.zeros(B, dtype=torch.bool) | (max_iter <= 0)
    out_ids, out_logits, time = [], [], 0
    while not bool(finished.all()):
        logits, state = dec.step(tokens, state)
        sample = torch.argmax(logits, dim=1)                 # ties -> lowest index
        finished = finished | (sample == END) | (time + 1 >= max_iter)
        tokens = sample
        out_ids.append(sample)
        out_logits.append(logits)
        time += 1
    if not out_ids:
        return np.zeros((B, 0), np.int32), np.zeros((B, 0, dec.vocab))
    return torch.stack(out_ids, 1).to(torch.int32).numpy(), torch.stack(out_logits, 1).numpy()


def gather_tree(step_ids, parent_ids, max_len, end_token=END):
    """tfa.seq2seq.gather_tree, level-synchronous over all (batch, beam) pairs.  [T,B,W] int64 tensors."""
    T, B, W = step_ids.shape
    L = torch.clamp(torch.as_tensor(max_len, dtype=torch.int64), max=T)
    out = torch.full_like(step_ids, end_token)
    parent = torch.arange(W).expand(B, W).clone()
    for level in range(T - 1, -1, -1):
        live = (level < L).unsqueeze(1)                      # rows whose back-trace has started
        out[level] = torch.where(live, torch.gather(step_ids[level], 1, parent), out[level])
        parent = torch.where(live, torch.gather(parent_ids[level], 1, parent), parent)
    # everything after the first end token becomes the end token
    seen = torch.zeros(B, W, dtype=torch.bool)
    for t in range(T):
        out[t] = torch.where(seen, torch.full_like(out[t], end_token), out[t])
        seen = seen | (out[t] == end_token)
    return out


def beam_search(w, enc_output, mask, beam_width, max_output_len, decoder_depth=1, dtype=torch.float32, return_all=False):
    """BeamSearchDecoder(length_penalty_weight=0) under dynamic_decode, then finalize (gather_tree); slot 0 is returned
    as the reference does (basecaller.py:315)."""
    dec = TorchDecoder(w, decoder_depth, dtype)
    B, W, V = len(enc_output), int(beam_width), 7
    dec.setup_memory(enc_output, mask, repeat=W)             # tile_batch: row b*W + k
    state = dec.zero_state(B * W)
    tokens = torch.full((B * W,), START, dtype=torch.int64)
    neg_inf, fmin = float("-inf"), torch.finfo(dtype).min
    log_probs = torch.full((B, W), neg_inf, dtype=dtype); log_probs[:, 0] = 0.0
    finished = torch.ones(B, W, dtype=torch.bool); finished[:, 0] = False
    lengths = torch.zeros(B, W, dtype=torch.int64)
    max_iter = int(max_output_len) - 1
    done = max_iter <= 0
    base = (torch.arange(B) * W).unsqueeze(1)
    steps_sc, steps_id, steps_par, time = [], [], [], 0
    while not done:
        logits, state = dec.step(tokens, state)
        step_lp = torch.log_softmax(logits.view(B, W, V), dim=2)
        # _mask_probs: a finished beam can only continue with the end token, at no cost
        fin_row = torch.full((V,), fmin, dtype=dtype); fin_row[END] = 0.0
        step_lp = torch.where(finished.unsqueeze(2), fin_row.view(1, 1, V), step_lp)
        total = (log_probs.unsqueeze(2) + step_lp).view(B, W * V)
        order = torch.sort(total, dim=1, descending=True, stable=True).indices[:, :W]     # top_k: equal values keep index order
        scores = torch.gather(total, 1, order)
        word, beam = order % V, order // V
        was_finished = torch.gather(finished, 1, beam)
        lengths = torch.gather(lengths, 1, beam) + (~was_finished).long()
        finished = was_finished | (word == END)
        log_probs = scores
        flat = (base + beam).view(-1)
        state = {"cells": [(h[flat], c[flat]) for h, c in state["cells"]], "attention": state["attention"][flat]}
        tokens = word.reshape(-1)
        steps_sc.append(scores); steps_id.append(word); steps_par.append(beam)
        time += 1
        done = bool(finished.all()) or time >= max_iter
    if not steps_id:
        return np.zeros((B, 0), np.int32), np.zeros((B, 0))
    ids, par = torch.stack(steps_id), torch.stack(steps_par)                 # [T,B,W]
    pred = gather_tree(ids, par, lengths.max(dim=1).values).permute(1, 0, 2)
    sc = torch.stack(steps_sc).permute(1, 0, 2)
    if return_all:
        return pred.to(torch.int32).numpy(), sc.numpy(), ids.permute(1, 0, 2).to(torch.int32).numpy(), par.permute(1, 0, 2).to(torch.int32).numpy()
    return pred[:, :, 0].to(torch.int32).numpy(), sc[:, :, 0].numpy()
