"""Weight containers for Basecaller.load_weights.

The reference restores a Keras TF-format checkpoint
(ravvent_performance_evaluator.py:107).  Here the interchange is a flat
name -> float32 array mapping (an .npz file or a dict):

    encoder_{raw,event}/layer{l}/{forward,backward}/{kernel,recurrent_kernel,bias}
    decoder/cell0/{kernel,recurrent_kernel,bias}
    decoder/memory_layer/kernel, decoder/attention_layer/kernel, decoder/fc/{kernel,bias}

with Keras shapes (kernel [in,4u], recurrent_kernel [u,4u], bias [4u], gate order
i,f,g,o).  ``random_weights(seed)`` draws the Keras default initialisers
(glorot_uniform kernels, orthogonal recurrent kernels, zero biases with
unit_forget_bias; basecaller.py:22-23, 86) from a seeded numpy Generator."""
from __future__ import annotations

import numpy as np


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


def _lstm(rng, n_in, units):
    bias = np.zeros(4 * units, dtype=np.float32)
    bias[units:2 * units] = 1.0
    return {"kernel": _glorot(rng, n_in, 4 * units), "recurrent_kernel": _orthogonal(rng, units, 4 * units),
            "bias": bias}


def _gru(rng, n_in, units):
    """Keras GRUCell defaults (reset_after=True): bias [2, 3u] = (input, recurrent), gate blocks z, r, h."""
    return {"kernel": _glorot(rng, n_in, 3 * units), "recurrent_kernel": _orthogonal(rng, units, 3 * units),
            "bias": np.zeros((2, 3 * units), dtype=np.float32)}


def random_weights(seed=22, enc_units=128, dec_units=128, encoder_depth=2, decoder_depth=1, vocab_size=7, rnn_type="bilstm"):
    """rnn_type as in the reference constructor (basecaller.py:25-46, 86-89, 195): 'bi*' -> forward + backward encoder
    weights and a 2*enc_units memory; '*lstm' / '*gru' -> the cell of the encoders and the decoder."""
    rng = np.random.default_rng(seed)
    w = {}
    bi = "bi" in rnn_type
    cell = _lstm if "lstm" in rnn_type else _gru
    enc_out = (2 if bi else 1) * enc_units
    for enc, feat in (("encoder_raw", 1), ("encoder_event", 5)):
        for l in range(encoder_depth):
            n_in = feat if l == 0 else enc_out
            for d in (("forward", "backward") if bi else ("forward",)):
                for k, v in cell(rng, n_in, enc_units).items():
                    w[f"{enc}/layer{l}/{d}/{k}"] = v
    for j in range(decoder_depth):
        n_in = vocab_size + dec_units if j == 0 else dec_units
        for k, v in cell(rng, n_in, dec_units).items():
            w[f"decoder/cell{j}/{k}"] = v
    w["decoder/memory_layer/kernel"] = _glorot(rng, enc_out, dec_units)
    w["decoder/attention_layer/kernel"] = _glorot(rng, dec_units + enc_out, dec_units)
    w["decoder/fc/kernel"] = _glorot(rng, dec_units, vocab_size)
    w["decoder/fc/bias"] = np.zeros(vocab_size, dtype=np.float32)
    return w


def load_npz(path):
    with np.load(path) as z:
        return {k: np.asarray(z[k], dtype=np.float32) for k in z.files}
