"""CPU oracle: events -> model inputs ("snippets").  TEST INFRASTRUCTURE ONLY.

Restates the inference-relevant half of ``data_loader.prepare_snippets`` and
its helpers (file:line under /root/reference):
  * 7-column event table + feature scaler fit      data_loader.py:74-79
  * trimming to the labelled sample range           data_loader.py:82-87
  * whole-read raw standardisation                  data_loader.py:89-90
  * ``compute_fitting_event_ranges``                data_loader.py:29-46
  * ``convert_events_ranges_to_raw_ranges``         data_loader.py:48-51
  * slicing + ``pad_input_snippets``                data_loader.py:96-99, 110-111
Third-party semantics restated: sklearn ``StandardScaler``
= (x - mean) / sqrt(population variance) with zero variance -> scale 1; Keras
``pad_sequences(maxlen, 'float32', padding='post', truncating='post', value=0.)``.

Pinned: compared with the reference function itself (TF/keras imports stubbed)
by tools/make_golden.py -> tests/golden/snippets_*.npz.  Index outputs are
exact; float32 outputs agree to 1 ulp (sklearn accumulates the scaler moments
in a different order than ``np.mean`` / ``np.var``).
"""
from __future__ import annotations

import numpy as np

MAX_RAW_LEN = 200    # data_loader.py:16
MAX_EVENT_LEN = 30   # data_loader.py:17
N_EVENT_FEATURES = 5


def event_feature_table(start, length, mean, stdv) -> np.ndarray:
    """[n,7] float64: start, end, length, mean, stdv, mean**2, delta-mean."""
    start = np.asarray(start, dtype=np.float64)
    length = np.asarray(length, dtype=np.float64)
    mean = np.asarray(mean, dtype=np.float64)
    n = start.size
    tab = np.empty((n, 7), dtype=np.float64)
    tab[:, 0] = start
    tab[:, 1] = start + length
    tab[:, 2] = length
    tab[:, 3] = mean
    tab[:, 4] = np.asarray(stdv, dtype=np.float64)
    tab[:, 5] = [float(m) ** 2 for m in mean]       # Python float pow, as the reference
    tab[1:, 6] = mean[1:] - mean[:-1]
    if n:
        tab[0, 6] = 0.0
    return tab


def scaler_fit(x: np.ndarray):
    """StandardScaler.fit: column mean and scale (population std, 0 -> 1)."""
    mu = x.mean(axis=0)
    var = x.var(axis=0)
    scale = np.sqrt(var)
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
    return mu, scale


def fitting_event_ranges(lengths, stride: int, raw_max_len: int = MAX_RAW_LEN) -> np.ndarray:
    """Windows [first, end) of events, one every ``stride`` events, each ending
    at the first event whose cumulative length exceeds ``raw_max_len``."""
    cum = np.cumsum(lengths, axis=0, dtype=np.int32)
    n = len(lengths)
    out = []
    for first in range(0, n, stride):
        over = np.flatnonzero(cum > raw_max_len)
        if over.size == 0 or over[0] == 0:
            break
        out.append((first, int(over[0])))
        nxt = first + stride - 1
        if nxt >= n:
            break
        cum = cum - cum[nxt]
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


def pad_post(seqs, maxlen: int, width: int) -> np.ndarray:
    out = np.zeros((len(seqs), maxlen, width), dtype=np.float32)
    for i, s in enumerate(seqs):
        s = np.asarray(s)[:maxlen]
        out[i, :s.shape[0], :] = s.reshape(s.shape[0], width).astype(np.float32)
    return out


def build_snippets(raw, start, length, mean, stdv, label_start: int, label_end: int,
                   stride: int = 6):
    """-> dict(raw [Ns,200,1] f32, event [Ns,30,5] f32, event_ranges [Ns,2],
    raw_ranges [Ns,2], ev_mu/ev_scale [5], raw_mu/raw_scale)."""
    raw = np.asarray(raw)
    tab = event_feature_table(start, length, mean, stdv)
    ev_mu, ev_scale = scaler_fit(tab[:, 2:])
    keep = np.logical_and(tab[:, 0] >= label_start, tab[:, 1] <= label_end)
    tab = tab[keep]
    tab[0, 2] += tab[0, 0] - label_start
    tab[0, 0] = label_start
    tab[-1, 2] = label_end - tab[-1, 0]
    rawf = raw.astype(np.float64).reshape(-1, 1)
    raw_mu, raw_scale = scaler_fit(rawf)
    raw_sc = (rawf - raw_mu) / raw_scale
    ev_ranges = fitting_event_ranges(tab[:, 2], stride, MAX_RAW_LEN)
    if ev_ranges.shape[0] == 0:
        raw_ranges = np.zeros((0, 2), dtype=np.int64)
    else:
        raw_ranges = np.column_stack((tab[ev_ranges[:, 0], 0].astype(np.int32),
                                      tab[ev_ranges[:, 1] - 1, 0].astype(np.int32))).astype(np.int64)
    ev_sc = (tab[:, 2:] - ev_mu) / ev_scale
    raw_snips = [raw_sc[a:b] for a, b in raw_ranges]
    ev_snips = [ev_sc[a:b] for a, b in ev_ranges]
    return {
        "raw": pad_post(raw_snips, MAX_RAW_LEN, 1),
        "event": pad_post(ev_snips, MAX_EVENT_LEN, N_EVENT_FEATURES),
        "event_ranges": ev_ranges, "raw_ranges": raw_ranges,
        "ev_mu": ev_mu, "ev_scale": ev_scale,
        "raw_mu": raw_mu[0], "raw_scale": raw_scale[0],
    }
