"""Inference-side constants and helpers of the reference's data_loader.py
(constants :12-26; prepare_snippets / pad_input_snippets :70-111), with the
heavy lifting on the GPU (EventDetector = K1, snippet builder kernels)."""
from __future__ import annotations

import numpy as np

ED_WINDOW_LENGTH_1 = 6      # data_loader.py:12
ED_WINDOW_LENGTH_2 = 9      # data_loader.py:13
INPUT_PADDING = 0.          # data_loader.py:14
MAX_RAW_LEN = 200           # data_loader.py:16
MAX_EVENT_LEN = 30          # data_loader.py:17


class CharTokenizer:
    """Stand-in for the Keras Tokenizer the reference configures by hand
    (data_loader.py:20-22): only word_index / index_word / sequences_to_texts /
    texts_to_sequences are used on the inference path."""

    def __init__(self, word_index):
        self.word_index = dict(word_index)
        self.index_word = {v: k for k, v in self.word_index.items()}

    def sequences_to_texts(self, sequences):
        # Keras joins the mapped words with ' ' and skips ids without a mapping
        return [" ".join(self.index_word[int(i)] for i in seq if int(i) in self.index_word) for seq in sequences]

    def texts_to_sequences(self, texts):
        return [[self.word_index[ch] for ch in t.lower() if ch in self.word_index] for t in texts]


nuc_tk = CharTokenizer({'': 0, '^': 1, '$': 2, 'a': 3, 'c': 4, 'g': 5, 't': 6})
NUC_TOKEN_END = nuc_tk.word_index['^']
NUC_TOKEN_START = nuc_tk.word_index['$']
NUC_TOKEN_PAD = nuc_tk.word_index['']


def unpack_data_to_input_target(data, input_data_type):
    """utils.unpack_data_to_input_target (utils.py:34-43)."""
    raw_sequence, events_sequence, target_sequence = data
    if input_data_type == 'raw':
        return raw_sequence, target_sequence
    if input_data_type == 'event':
        return events_sequence, target_sequence
    if input_data_type == 'joint':
        return (raw_sequence, events_sequence), target_sequence
    raise ValueError(input_data_type)


def masked_accuracy(y_true, y_pred, omit_vals):
    """Fraction of positions with y_true == y_pred among those whose y_true is not in omit_vals (utils.py:15-24)."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    mask = np.ones(y_true.shape, dtype=np.int64)
    for ov in omit_vals:
        mask *= (y_true != ov).astype(np.int64)
    return np.float64((mask * (y_true == y_pred)).sum()) / np.float64(mask.sum())


def calc_prob_logits_beam_search_scores(beam_scores):
    """utils.calc_prob_logits_beam_search_scores (utils.py:123-128): per-base
    probability exp(score[t] - score[t-1]) from cumulative beam scores."""
    s = np.asarray(beam_scores.cpu() if hasattr(beam_scores, "cpu") else beam_scores)
    prev = np.zeros_like(s)
    prev[..., 1:] = s[..., :-1]
    return np.exp(s - prev)


def load_data_from_signal(raw, label_start=None, label_end=None, stride=6, device=None, detector=None,
                          return_ranges=False):
    """Inference half of ``load_data_from_single_signal_label`` (data_loader.py:113-126) for one read that
    is already in memory: GPU event detection (K1), then the GPU snippet builder.

    raw: 1-D integer samples.  ``label_start`` / ``label_end`` are the first start / last end of the
    labelled sample range (``nuc_raw_ranges[0,0]`` / ``nuc_raw_ranges[-1,1]``); default: the whole read.
    -> (raw_snippets [Ns,200,1] f32, event_snippets [Ns,30,5] f32) torch tensors on the device."""
    import ctypes as C
    import torch
    from . import _lib
    from .event_detector import EventDetector
    arr = np.asarray(raw)
    n = int(arr.size)
    det = detector or EventDetector(ED_WINDOW_LENGTH_1, ED_WINDOW_LENGTH_2, device=device)
    dev = det.device
    if arr.dtype != np.int16:
        arr = arr.astype(np.int32)
    sig = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    ev = det.detect_batch(sig, [0, n])
    n_ev = int(ev["count"][0].item()) if n else 0
    lab0 = 0 if label_start is None else int(label_start)
    lab1 = n if label_end is None else int(label_end)
    cap = max(1, (n_ev + int(stride) - 1) // int(stride))
    with torch.cuda.device(dev):
        raw_s = torch.zeros((cap, MAX_RAW_LEN, 1), dtype=torch.float32, device=dev)
        ev_s = torch.zeros((cap, MAX_EVENT_LEN, 5), dtype=torch.float32, device=dev)
        cnt = C.c_int32(0)
        ranges = torch.zeros((cap, 2), dtype=torch.int32, device=dev) if return_ranges else None
        _lib.check(_lib.lib.rvb_build_snippets(
            sig.data_ptr(), sig.element_size(), n, ev["start"].data_ptr(), ev["length"].data_ptr(),
            ev["mean"].data_ptr(), ev["stdv"].data_ptr(), n_ev, lab0, lab1, int(stride),
            raw_s.data_ptr(), ev_s.data_ptr(), cap, C.byref(cnt), None if ranges is None else ranges.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream))
    if return_ranges:
        return raw_s[:cnt.value], ev_s[:cnt.value], ranges[:cnt.value]
    return raw_s[:cnt.value], ev_s[:cnt.value]


def load_data_from_signals(signal, read_offsets, stride=6, device=None, detector=None, with_raw=True, return_ranges=False):
    """``load_data_from_signal`` for a batch of whole reads in ONE pass: one event-detection launch for all reads (K1) and one
    call of the batched snippet builder (``rvb_build_snippets_batch``) -- two device-to-host round trips for the batch (the
    event counts, to size the outputs, and the snippet total) instead of one per read.  The reference loops over the reads of a directory on the host (ravvent_performance_evaluator.py:60-66).

    signal: 1-D integer samples of all reads concatenated (array or device tensor); read_offsets: n_reads + 1.
    -> (raw_snippets [Ns,200,1] f32 or None if not with_raw, event_snippets [Ns,30,5] f32, snippet_offsets [n_reads+1] int64)
    on the device; the snippets of read r are rows snippet_offsets[r] .. snippet_offsets[r+1]."""
    import ctypes as C
    import torch
    from . import _lib
    from .event_detector import EventDetector
    det = detector or EventDetector(ED_WINDOW_LENGTH_1, ED_WINDOW_LENGTH_2, device=device)
    dev = det.device
    offs = np.ascontiguousarray(np.asarray(read_offsets, dtype=np.int64))
    n_reads = offs.size - 1
    if isinstance(signal, torch.Tensor):
        sig = signal.to(dev)
        if sig.dtype not in (torch.int16, torch.int32):
            sig = sig.to(torch.int32)
    else:
        arr = np.asarray(signal)
        if arr.dtype != np.int16:
            arr = arr.astype(np.int32)
        sig = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    sig = sig.contiguous()
    ev = det.detect_batch(sig, offs)
    ev_offs = np.ascontiguousarray(ev["event_offsets"], dtype=np.int64)
    # capacity: a read of E events has at most ceil(E / stride) windows.  The detected counts come back first (one small
    # copy: the event arrays' own capacity, samples / 2, would over-allocate the outputs ~5x); every returned row is
    # written in full by the builder, so the buffers need no clearing.
    counts = ev["count"].cpu().numpy().astype(np.int64) if n_reads else np.zeros(0, np.int64)
    cap = max(int(np.sum((counts + int(stride) - 1) // int(stride))), 1)
    with torch.cuda.device(dev):
        raw_s = torch.empty((cap, MAX_RAW_LEN, 1), dtype=torch.float32, device=dev) if with_raw else None
        ev_s = torch.empty((cap, MAX_EVENT_LEN, 5), dtype=torch.float32, device=dev)
        soff = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
        ranges = torch.zeros((cap, 2), dtype=torch.int32, device=dev) if return_ranges else None
        total = C.c_int64(0)
        _lib.check(_lib.lib.rvb_build_snippets_batch(
            sig.data_ptr(), sig.element_size(), offs.ctypes.data, n_reads, ev_offs.ctypes.data,
            ev["start"].data_ptr(), ev["length"].data_ptr(), ev["mean"].data_ptr(), ev["stdv"].data_ptr(), ev["count"].data_ptr(),
            int(stride), None if raw_s is None else raw_s.data_ptr(), ev_s.data_ptr(), cap, soff.data_ptr(),
            None if ranges is None else ranges.data_ptr(), C.byref(total), torch.cuda.current_stream(dev).cuda_stream))
    n = total.value
    out = (None if raw_s is None else raw_s[:n], ev_s[:n], soff)
    return out + (ranges[:n],) if return_ranges else out


def _label_tokens(raw_ranges, nuc_raw_ranges, nuc_reference_symbols):
    """Target token rows of the snippets (data_loader.py:53-61, 101-108, 123-124): the labelled bases whose
    sample ranges intersect each snippet's raw range, wrapped in '$' ... '^', tokenised char-level and
    post-padded with the padding token to the longest row (int64)."""
    nuc_raw_ranges = np.asarray(nuc_raw_ranges).astype(np.int64)
    lens = nuc_raw_ranges[:, 1] - nuc_raw_ranges[:, 0]
    edges = nuc_raw_ranges[0, 0] + np.concatenate(([0], np.cumsum(lens)))      # id k covers [edges[k], edges[k+1])
    total = int(edges[-1])
    rows = []
    for r0, r1 in np.asarray(raw_ranges, dtype=np.int64):
        r1 = min(int(r1), total)
        if r1 <= r0:
            ids = np.zeros(0, dtype=np.int64)
        else:
            first = int(np.searchsorted(edges, r0, side="right") - 1)
            last = int(np.searchsorted(edges, r1 - 1, side="right") - 1)
            ids = np.arange(first, last + 1)                                   # first == -1 reproduces the reference's quirk
        text = '$' + ''.join(np.asarray(nuc_reference_symbols, dtype=object)[ids]) + '^'
        rows.append(nuc_tk.texts_to_sequences([text])[0])
    width = max((len(r) for r in rows), default=0)
    out = np.full((len(rows), width), NUC_TOKEN_PAD, dtype=np.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def load_data_from_single_signal_label(signal_path, label_path, stride, as_numpy=True, device=None):
    """Drop-in for data_loader.load_data_from_single_signal_label (data_loader.py:113-126): Chiron-style
    `.signal` (whitespace-separated integers) + `.label` (`start end base` rows) ->
    (raw_snippets [Ns,200,1] f32, event_snippets [Ns,30,5] f32, nuc_tk_snippets [Ns,L] int64).
    Event detection and snippet building run on the GPU; the label side is host numpy like the reference."""
    raw = np.loadtxt(signal_path, dtype=int)
    label = np.loadtxt(label_path, dtype=object, ndmin=2)
    nuc_raw_ranges = label[:, :2].astype(int)
    nuc_reference_symbols = label[:, 2]
    rs, es, ranges = load_data_from_signal(raw, int(nuc_raw_ranges[0, 0]), int(nuc_raw_ranges[-1, 1]), stride,
                                           device=device, return_ranges=True)
    tokens = _label_tokens(ranges.cpu().numpy(), nuc_raw_ranges, nuc_reference_symbols)
    if as_numpy:
        return rs.cpu().numpy(), es.cpu().numpy(), tokens
    return rs, es, tokens


def create_files_info(files_dir, stride=6, verbose=True, device=None):
    """data_loader.create_files_info (data_loader.py:129-156): index every `.signal` / `.label` pair of a directory with
    its snippet count into `files_info.snippets.stride_<stride>.json` (the list the evaluators iterate over)."""
    import json
    from pathlib import Path
    d = Path(files_dir)
    files_info_path = d / f'files_info.snippets.stride_{stride}.json'
    signals = sorted(p for p in d.iterdir() if p.suffix == '.signal')
    labels = sorted(p for p in d.iterdir() if p.suffix == '.label')
    files_info = []
    for signal_path, label_path in zip(signals, labels):
        raw_snippets, _, _ = load_data_from_single_signal_label(signal_path, label_path, stride, device=device)
        if verbose:
            print('{}'.format(signal_path.stem))
        files_info.append({'signal_path': signal_path.as_posix(), 'label_path': label_path.as_posix(),
                           'snippets_num': int(raw_snippets.shape[0])})
        with open(files_info_path, 'wt') as fi:
            json.dump(files_info, fi, indent=2)
    with open(files_info_path, 'wt') as fi:
        json.dump(files_info, fi, indent=2)
    return files_info_path
