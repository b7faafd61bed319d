"""CPU oracle: two-window t-statistic event detection (TEST INFRASTRUCTURE ONLY).

Restates the streaming detector of the reference
(``event_detection/event_detector.py``) as a function over absolute-index
prefix sums, which is the formulation the CUDA scan kernel uses.  Parity is
PINNED: ``tools/make_golden.py`` runs the reference module itself (imported
from /root/reference in the build container) on seeded signals and this file
is asserted identical on them (tests/test_oracle_event.py).

Mapping to the reference (file:line under /root/reference):
  * ring of prefix sums, ``_add_sample``      event_detector.py:85-107
  * ``get_buf_mid`` / u32 wrap                 event_detector.py:72-73, 281-287
  * ``_compute_tstat``                         event_detector.py:109-147
  * ``_detect_peak``                           event_detector.py:149-187
  * ``_create_event``                          event_detector.py:189-210

Semantics that must be kept (all reproduced below):
  * The reference keeps S[n] = sum(raw[:n]) and Q[n] = sum(raw[:n]**2) in a
    ring of ``BUF = 1 + 2*w2`` float64 slots; slot k therefore holds the value
    for the *latest* n' <= n with n' = k (mod BUF).  A slot that has not been
    written yet holds 0.0 (fresh detector, as ``data_loader.prepare_snippets``
    constructs one per read, data_loader.py:71).
  * ``buf_mid = (n - w2) mod 2**32`` with n = samples consumed so far; window
    edges ``buf_mid -/+ w`` are wrapped to u32 *before* the ``% BUF``; during
    warm-up this reads "wrong" slots and produces the spurious first event.
  * all arithmetic is IEEE float64, evaluated left to right, no FMA.
  * prefix sums of integer samples are exact in float64 (< 2**53), so any
    summation order gives the same bits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

FLT_MIN = 1.17549435e-38  # event_detector.py:10
FLT_MAX = 3.40282347e+38  # event_detector.py:11
U32 = 0xFFFFFFFF


@dataclass
class EventTable:
    """Structure-of-arrays result; one row per detected event."""
    start: np.ndarray   # int64
    length: np.ndarray  # int64
    mean: np.ndarray    # float64
    stdv: np.ndarray    # float64

    def __len__(self) -> int:
        return int(self.start.shape[0])


def _i32(v: int) -> int:
    v &= U32
    return v - (1 << 32) if v & 0x80000000 else v


class _Peak:
    """State of one peak detector (event_detector.py:46-68)."""
    __slots__ = ("w", "thr", "masked_to", "pos", "val", "valid")

    def __init__(self, w: int, thr: float):
        self.w, self.thr = w, thr
        self.masked_to, self.pos, self.val, self.valid = 0, -1, FLT_MAX, False

    def clear(self) -> None:
        self.pos, self.val, self.valid = -1, FLT_MAX, False


def prefix_sums(raw: np.ndarray):
    """S[n], Q[n] for n = 0..N as float64 (exact for integer samples)."""
    r = np.asarray(raw).astype(np.int64)
    S = np.zeros(r.size + 1, dtype=np.float64)
    Q = np.zeros(r.size + 1, dtype=np.float64)
    S[1:] = np.cumsum(r).astype(np.float64)
    Q[1:] = np.cumsum(r * r).astype(np.float64)
    return S, Q


def ring_lookup(P: np.ndarray, n: int, slot: int, buf: int) -> float:
    """Value the reference's ring holds in ``slot`` after n samples."""
    m = n - ((n - slot) % buf)
    return float(P[m]) if m >= 0 else 0.0


def tstat(S, Q, n: int, w: int, w2: int) -> float:
    """t-statistic the reference computes right after consuming n samples
    (event_detector.py:109-147; ``self.t == n + 1`` at that point)."""
    if (n + 1) <= 2 * w or w < 2:
        return 0.0
    buf = 1 + 2 * w2
    mid = (n - w2) & U32
    i = mid % buf
    st = ((mid - w) & U32) % buf
    en = ((mid + w) & U32) % buf
    wf = float(w)
    s_i, s_st, s_en = (ring_lookup(S, n, k, buf) for k in (i, st, en))
    q_i, q_st, q_en = (ring_lookup(Q, n, k, buf) for k in (i, st, en))
    sum1, sumsq1 = s_i - s_st, q_i - q_st
    sum2, sumsq2 = s_en - s_i, q_en - q_i
    mean1, mean2 = sum1 / wf, sum2 / wf
    var = sumsq1 / wf - mean1 * mean1 + sumsq2 / wf - mean2 * mean2
    var = max(var, FLT_MIN)
    return math.fabs(mean2 - mean1) / math.sqrt(var / wf)


def _step_peak(d: _Peak, other_long: _Peak | None, value: float, mid: int,
               peak_height: float) -> bool:
    """One update of a detector (event_detector.py:149-187).  ``other_long`` is
    the long detector when ``d`` is the short one (the short detector masks
    and clears the long one), else None."""
    if d.masked_to >= mid:
        return False
    if d.pos == -1:
        if value < d.val:
            d.val = value
        elif value - d.val > peak_height:
            d.val, d.pos = value, _i32(mid)
        return False
    if value > d.val:
        d.val, d.pos = value, _i32(mid)
    if other_long is not None and d.val > d.thr:
        other_long.masked_to = (d.pos + d.w) & U32
        other_long.clear()
    if d.val - value > peak_height and d.val > d.thr:
        d.valid = True
    if d.valid and (mid - d.pos) > d.w / 2:
        d.pos, d.val, d.valid = -1, value, False
        return True
    return False


def detect_events(raw, window_length1: int = 3, window_length2: int = 6,
                  threshold1: float = 1.4, threshold2: float = 9.0,
                  peak_height: float = 0.2) -> EventTable:
    """Events of one read; equals ``EventDetector(...).run(raw)`` of the
    reference on a freshly constructed detector (event_detector.py:75-83)."""
    w1, w2 = int(window_length1), int(window_length2)
    assert w1 > 0 and w2 > 0
    buf = 1 + 2 * w2
    S, Q = prefix_sums(raw)
    N = S.size - 1
    short, long_ = _Peak(w1, threshold1), _Peak(w2, threshold2)
    same_w = (w1 == w2)  # reference identifies "the short detector" by window length (:169)
    ev_st, ev_sum, ev_sq = 0, 0.0, 0.0
    starts, lengths, means, stdvs = [], [], [], []
    for n in range(1, N + 1):          # n = samples consumed
        mid = (n - w2) & U32
        t1 = tstat(S, Q, n, w1, w2)
        t2 = tstat(S, Q, n, w2, w2)
        p1 = _step_peak(short, long_, t1, mid, peak_height)
        p2 = _step_peak(long_, long_ if same_w else None, t2, mid, peak_height)
        if not (p1 or p2):
            continue
        en = (mid - w1 + 1) & U32
        length = float(en - ev_st)
        if length < FLT_MIN:
            continue
        slot = en % buf
        s_en, q_en = ring_lookup(S, n, slot, buf), ring_lookup(Q, n, slot, buf)
        mean = float(s_en - ev_sum) / length
        var = (q_en - ev_sq) / length - mean ** 2   # libm pow, as the reference (:201)
        starts.append(ev_st)
        lengths.append(int(length))
        means.append(mean)
        stdvs.append(math.sqrt(max(var, FLT_MIN)))
        ev_st, ev_sum, ev_sq = en, s_en, q_en
    return EventTable(np.asarray(starts, dtype=np.int64), np.asarray(lengths, dtype=np.int64),
                      np.asarray(means, dtype=np.float64), np.asarray(stdvs, dtype=np.float64))


def synth_read(rng: np.random.Generator, n_samples: int) -> np.ndarray:
    """Synthetic nanopore-like read: piecewise-constant levels
    U(250,550), dwell 2+Geometric(1/7), N(0,8^2) noise, rounded to int32."""
    n_lvl = n_samples // 3 + 8
    dwell = 2 + rng.geometric(1.0 / 7.0, size=n_lvl)
    level = rng.uniform(250.0, 550.0, size=n_lvl)
    sig = np.repeat(level, dwell)[:n_samples]
    assert sig.size == n_samples
    sig = sig + rng.normal(0.0, 8.0, size=n_samples)
    return np.rint(sig).astype(np.int32)
