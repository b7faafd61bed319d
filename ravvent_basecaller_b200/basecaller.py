"""Basecaller -- the reference's construct / load_weights / predict surface
(basecaller.py:156-207, 289-330, 384-416) over libravvent_b200's CUDA kernels.

No TensorFlow: host code only moves buffers and calls the C ABI.  Inputs may be
numpy arrays (host) or torch tensors (host or device); outputs mirror the input
kind (numpy in -> numpy out, torch in -> torch tensors on the model's device).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from . import tf_checkpoint as _tfckpt
from . import weights as _weights


class Basecaller:
    def __init__(self, enc_units: int, dec_units: int, batch_sz: int, tokenizer, input_data_type: str,
                 input_padding_value, encoder_depth: int = 2, decoder_depth: int = 1, rnn_type: str = 'bilstm',
                 teacher_forcing=True, attention_type: str = 'luong', beam_width: int = 5, *,
                 device: int | None = None, precision: str = 'fp32', wave_snippets: int = 0):
        if input_data_type not in ('raw', 'event', 'joint'):
            raise ValueError("input_data_type must be 'raw', 'event' or 'joint'")
        if rnn_type not in ('bilstm', 'lstm', 'bigru', 'gru'):
            raise NotImplementedError("rnn_type must be one of 'gru', 'lstm', 'bigru', 'bilstm' (basecaller.py:167)")
        if float(input_padding_value) != 0.0:
            raise NotImplementedError("input_padding_value must be 0.0 (data_loader.INPUT_PADDING)")
        self.batch_sz = batch_sz
        self.rnn_type = rnn_type
        self.tokenizer = tokenizer
        self.enc_units, self.dec_units = int(enc_units), int(dec_units)
        self.encoder_depth, self.decoder_depth = int(encoder_depth), int(decoder_depth)
        self.max_input_len = {'raw': 200, 'event': 30, 'joint': 230}[input_data_type]   # basecaller.py:180-185
        self.teacher_forcing = teacher_forcing
        self.input_data_type = input_data_type
        self.input_padding_value = input_padding_value
        self.attention_type = attention_type          # the reference hard-codes 'luong' in the decoder (:194)
        self.beam_width = beam_width
        self.vocab_size = len(tokenizer.word_index)
        self.output_start_token = np.int32(tokenizer.word_index['$'])
        self.output_end_token = np.int32(tokenizer.word_index['^'])
        self.output_padding_token = np.int32(tokenizer.word_index[''])
        if (int(self.output_start_token), int(self.output_end_token), int(self.output_padding_token)) != (2, 1, 0):
            raise NotImplementedError("token ids must be $=2, ^=1, ''=0 (data_loader.nuc_tk)")
        # descriptive stand-ins for the Keras sub-models the reference exposes as attributes
        self.encoder_raw = SimpleNamespace(enc_units=enc_units, layer_depth=encoder_depth, inputs_features_num=1, rnn_type=rnn_type)
        self.encoder_event = SimpleNamespace(enc_units=enc_units, layer_depth=encoder_depth, inputs_features_num=5, rnn_type=rnn_type)
        self.decoder = SimpleNamespace(dec_units=dec_units, layer_depth=decoder_depth, vocab_size=self.vocab_size,
                                       attention_type='luong', max_input_len=self.max_input_len)
        if _lib.device_count() == 0:
            raise _lib.RavventError(_lib.RVB_ERR_CUDA, "no CUDA device: Basecaller has no CPU fallback")
        dev = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", dev)
        self.precision = precision
        self._h = C.c_void_p()
        _lib.check(_lib.lib.rvb_model_create(C.byref(self._h), dev, self.enc_units, self.dec_units, self.encoder_depth,
                                             self.decoder_depth, self.vocab_size, _lib.INPUT_KIND[input_data_type],
                                             _lib.PRECISION[precision], int(wave_snippets)))
        if rnn_type != 'bilstm':      # 'bi' -> Bidirectional encoders; the cell kind also applies to the decoder (basecaller.py:195)
            _lib.check(_lib.lib.rvb_model_set_rnn(self._h, 1 if 'bi' in rnn_type else 0,
                                                  _lib.CELL_KIND['lstm' if 'lstm' in rnn_type else 'gru']))
        self._weights_loaded = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None and getattr(_lib, "lib", None) is not None:      # module may be gone at interpreter exit
            _lib.lib.rvb_model_destroy(h)
            self._h = None

    # -- Keras surface the evaluators touch ---------------------------------
    def compile(self, *args, **kwargs):
        """Keras no-op for inference (ravvent_performance_evaluator.py:104-106)."""
        return None

    def load_weights(self, source=None, *, seed=None):
        """source: a Keras TF-format checkpoint prefix (as in ravvent_performance_evaluator.py:107; read by
        tf_checkpoint.py without TensorFlow), a path to an .npz, or a dict name -> array (see weights.py);
        or seed= for Keras-default random initialisation."""
        if source is None:
            w = _weights.random_weights(22 if seed is None else seed, self.enc_units, self.dec_units,
                                        self.encoder_depth, self.decoder_depth, self.vocab_size, self.rnn_type)
        elif isinstance(source, dict):
            w = source
        else:
            path = str(source)
            if path.endswith(".npz"):
                w = _weights.load_npz(path)
            elif _tfckpt.is_checkpoint_prefix(path):        # Keras TF-format prefix, as the reference passes it
                w = _tfckpt.load_keras_checkpoint(path)
            else:
                raise FileNotFoundError(f"{path}: neither an .npz file nor a TF-format checkpoint prefix "
                                        "(<prefix>.index / <prefix>.data-00000-of-00001)")
        for name, arr in w.items():
            if self.input_data_type == 'raw' and name.startswith('encoder_event'):
                continue
            if self.input_data_type == 'event' and name.startswith('encoder_raw'):
                continue
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(_lib.lib.rvb_model_set_weight(self._h, name.encode(), a.ctypes.data, shape, a.ndim))
        _lib.check(_lib.lib.rvb_model_finalize(self._h))
        self._weights_loaded = True
        return self

    # -- helpers -------------------------------------------------------------
    def _split(self, input_data):
        if self.input_data_type == 'joint':
            raw, event = input_data
        elif self.input_data_type == 'raw':
            raw, event = input_data, None
        else:
            raw, event = None, input_data
        return raw, event

    @staticmethod
    def _same_batch(raw, event):
        if raw is not None and event is not None and raw.shape[0] != event.shape[0]:
            raise ValueError(f"raw and event batches differ: {raw.shape[0]} vs {event.shape[0]}")

    def _dev(self, x, feat):
        if x is None:
            return None, 0
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if t.dim() != 3 or t.shape[-1] != feat:
            raise ValueError(f"expected input of shape [batch, time, {feat}], got {tuple(t.shape)}")
        t = t.to(self.device, torch.float32).contiguous()
        return t, int(t.shape[1])

    @staticmethod
    def _is_host(input_data):
        xs = input_data if isinstance(input_data, (tuple, list)) else (input_data,)
        return all(not isinstance(x, torch.Tensor) for x in xs)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _ptr(t):
        return t.data_ptr() if t is not None else None

    # -- inference -------------------------------------------------------------
    def _encode_input(self, input_data, training=False):
        """-> (enc_output [B,Tm,2*enc_units] f32, input_mask [B,Tm] bool)  (basecaller.py:395-416)"""
        host = self._is_host(input_data)
        raw, event = self._split(input_data)
        raw, t_raw = self._dev(raw, 1)
        event, t_ev = self._dev(event, 5)
        self._same_batch(raw, event)
        B = int((raw if raw is not None else event).shape[0])
        Tm = t_raw + t_ev
        with torch.cuda.device(self.device):
            enc = torch.empty((B, Tm, 2 * self.enc_units), dtype=torch.float32, device=self.device)
            mask = torch.empty((B, Tm), dtype=torch.uint8, device=self.device)
            _lib.check(_lib.lib.rvb_encode(self._h, self._ptr(raw), t_raw, self._ptr(event), t_ev, B,
                                           enc.data_ptr(), mask.data_ptr(), self._stream()))
        _lib.check(_lib.lib.rvb_model_check(self._h))
        mask = mask.bool()
        if 'bi' not in self.rnn_type:        # unidirectional encoders: the memory is enc_units wide (the internal backward half is all zeros)
            enc = enc[:, :, :self.enc_units].contiguous()
        if host:
            return enc.cpu().numpy(), mask.cpu().numpy()
        return enc, mask

    def greedy_search_prediction(self, input_data, max_output_len):
        """-> (sample_id [B,T] int32, rnn_output [B,T,V] f32)  (basecaller.py:317-330)"""
        host = self._is_host(input_data)
        raw, event = self._split(input_data)
        raw, t_raw = self._dev(raw, 1)
        event, t_ev = self._dev(event, 5)
        self._same_batch(raw, event)
        B = int((raw if raw is not None else event).shape[0])
        S = max(int(max_output_len) - 1, 0)
        with torch.cuda.device(self.device):
            ids = torch.empty((B, S), dtype=torch.int32, device=self.device)
            logits = torch.empty((B, S, self.vocab_size), dtype=torch.float32, device=self.device)
            steps = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib.rvb_greedy(self._h, self._ptr(raw), t_raw, self._ptr(event), t_ev, B, int(max_output_len),
                                           ids.data_ptr(), logits.data_ptr(), steps.data_ptr(), self._stream()))
            T = int(steps.item())
            _lib.check(_lib.lib.rvb_model_check(self._h))
        ids, logits = ids[:, :T], logits[:, :T]
        if host:
            return ids.cpu().numpy(), logits.cpu().numpy()
        return ids, logits

    def beam_search_prediction(self, input_data, beam_width, max_output_len, return_all_beams=False):
        """-> (predicted_ids[:, :, 0] [B,T] int32, scores[:, :, 0] [B,T] f32)  (basecaller.py:296-315)"""
        W = int(beam_width)
        S = max(int(max_output_len) - 1, 0)
        if self._is_host(input_data) and not return_all_beams:
            raw, event = self._split(input_data)
            raw = None if raw is None else np.ascontiguousarray(raw, dtype=np.float32)
            event = None if event is None else np.ascontiguousarray(event, dtype=np.float32)
            for x, feat in ((raw, 1), (event, 5)):          # the C side trusts these shapes: same checks as _dev()
                if x is not None and (x.ndim != 3 or x.shape[-1] != feat):
                    raise ValueError(f"expected input of shape [batch, time, {feat}], got {tuple(x.shape)}")
            if raw is not None and event is not None and raw.shape[0] != event.shape[0]:
                raise ValueError(f"raw and event batches differ: {raw.shape[0]} vs {event.shape[0]}")
            B = int((raw if raw is not None else event).shape[0])
            ids = np.empty((B, S), dtype=np.int32)
            scores = np.empty((B, S), dtype=np.float32)
            steps = C.c_int32(0)
            _lib.check(_lib.lib.rvb_beam_host(
                self._h, None if raw is None else raw.ctypes.data, 0 if raw is None else raw.shape[1],
                None if event is None else event.ctypes.data, 0 if event is None else event.shape[1],
                B, W, int(max_output_len), ids.ctypes.data, scores.ctypes.data, C.byref(steps)))
            return ids[:, :steps.value], scores[:, :steps.value]
        host = self._is_host(input_data)
        raw, event = self._split(input_data)
        raw, t_raw = self._dev(raw, 1)
        event, t_ev = self._dev(event, 5)
        self._same_batch(raw, event)
        B = int((raw if raw is not None else event).shape[0])
        with torch.cuda.device(self.device):
            ids = torch.empty((B, S, W), dtype=torch.int32, device=self.device)
            scores = torch.empty((B, S, W), dtype=torch.float32, device=self.device)
            step_ids = torch.empty((B, S, W), dtype=torch.int32, device=self.device)
            parents = torch.empty((B, S, W), dtype=torch.int32, device=self.device)
            steps = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib.rvb_beam(self._h, self._ptr(raw), t_raw, self._ptr(event), t_ev, B, W, int(max_output_len),
                                         ids.data_ptr(), scores.data_ptr(), step_ids.data_ptr(), parents.data_ptr(),
                                         steps.data_ptr(), self._stream()))
            T = int(steps.item())
            _lib.check(_lib.lib.rvb_model_check(self._h))
        if return_all_beams:
            out = (ids[:, :T], scores[:, :T], step_ids[:, :T], parents[:, :T])
            return tuple(o.cpu().numpy() for o in out) if host else out
        ids, scores = ids[:, :T, 0], scores[:, :T, 0]
        if host:
            return ids.cpu().numpy(), scores.cpu().numpy()
        return ids, scores

    # -- validation (basecaller.py:209-218, 264-279) -----------------------------
    def loss_function(self, real, pred):
        """Mean sparse categorical cross-entropy from logits over the non-padding target positions."""
        real = np.asarray(real).astype(np.int64)
        pred = np.asarray(pred, dtype=np.float32)
        m = pred.max(axis=-1, keepdims=True)
        lse = (m + np.log(np.exp(pred - m).sum(axis=-1, keepdims=True, dtype=np.float32)))[..., 0]
        loss = lse - np.take_along_axis(pred, real[..., None], axis=-1)[..., 0]
        mask = (real != self.output_padding_token).astype(np.float32)
        return np.float32((mask * loss).sum(dtype=np.float32) / mask.sum(dtype=np.float32))

    def _val_step(self, data):
        from .data_loader import masked_accuracy, unpack_data_to_input_target
        input_data, target_tokens = unpack_data_to_input_target(data, self.input_data_type)
        target_tokens = target_tokens.cpu().numpy() if isinstance(target_tokens, torch.Tensor) else np.asarray(target_tokens)
        max_output_len = int(target_tokens.shape[1])
        ids, logits = self.greedy_search_prediction(input_data, max_output_len=max_output_len)
        if isinstance(ids, torch.Tensor):
            ids, logits = ids.cpu().numpy(), logits.cpu().numpy()
        pad = max_output_len - 1 - ids.shape[1]                 # dynamic_decode may have stopped early: zero-pad as the reference does
        logits = np.pad(logits, [(0, 0), (0, pad), (0, 0)])
        ids = np.pad(ids, [(0, 0), (0, pad)])
        real = target_tokens[:, 1:]
        return {'loss': self.loss_function(real, logits),
                'acc': masked_accuracy(real, ids.astype(np.int64), [self.output_start_token, self.output_end_token])}

    def test_step(self, inputs):
        return self._val_step(inputs)

    def tokens_to_nuc_sequences(self, result_tokens):
        """ids -> text; strips ' ', '^', '$' and upper-cases (basecaller.py:289-294)."""
        if isinstance(result_tokens, torch.Tensor):
            result_tokens = result_tokens.cpu().numpy()
        elif hasattr(result_tokens, "numpy"):
            result_tokens = result_tokens.numpy()
        result_text = self.tokenizer.sequences_to_texts(np.asarray(result_tokens))
        return [rt.replace(' ', '').replace('^', '').replace('$', '').upper() for rt in result_text]
