"""Oracle (Python + C restatements) vs golden vectors produced by the reference's
own EventDetector (tools/make_golden.py).  Bit-exact on every field."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import event_ref


def _cases():
    g = np.load(GOLDEN / "event_golden.npz")
    for k in range(int(g["n_cases"])):
        w1, w2, t1, t2, ph = g[f"par_{k}"]
        yield k, g[f"raw_{k}"].astype(np.int32), (int(w1), int(w2), float(t1), float(t2), float(ph)), \
            g[f"start_{k}"], g[f"length_{k}"], g[f"mean_{k}"], g[f"stdv_{k}"]


CASES = list(_cases())


@pytest.mark.parametrize("case", CASES, ids=[f"case{c[0]}" for c in CASES])
def test_python_oracle_matches_reference(case):
    _, raw, par, start, length, mean, stdv = case
    ev = event_ref.detect_events(raw, *par)
    assert np.array_equal(ev.start, start)
    assert np.array_equal(ev.length, length)
    assert np.array_equal(ev.mean, mean)          # bitwise float64
    assert np.array_equal(ev.stdv, stdv)


def load_c_oracle():
    so = ROOT / "oracle" / "libravvent_oracle.so"
    if not so.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle")])
    lib = ctypes.CDLL(str(so))
    lib.rvo_detect_events.restype = ctypes.c_int64
    lib.rvo_detect_events.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int64]
    return lib


def c_detect(lib, raw, w1, w2, t1, t2, ph):
    raw = np.ascontiguousarray(raw, dtype=np.int32)
    cap = raw.size // 2 + 4
    st = np.empty(cap, np.int32); ln = np.empty(cap, np.int32)
    mu = np.empty(cap, np.float64); sd = np.empty(cap, np.float64)
    n = lib.rvo_detect_events(raw.ctypes.data, raw.size, w1, w2, t1, t2, ph,
                              st.ctypes.data, ln.ctypes.data, mu.ctypes.data, sd.ctypes.data, cap)
    assert 0 <= n <= cap
    return st[:n].astype(np.int64) & 0xFFFFFFFF, ln[:n].astype(np.int64) & 0xFFFFFFFF, mu[:n], sd[:n]


@pytest.mark.parametrize("case", CASES, ids=[f"case{c[0]}" for c in CASES])
def test_c_oracle_matches_reference(case):
    _, raw, par, start, length, mean, stdv = case
    st, ln, mu, sd = c_detect(load_c_oracle(), raw, *par)
    assert np.array_equal(st, start)
    assert np.array_equal(ln, length)
    assert np.array_equal(mu, mean)
    assert np.array_equal(sd, stdv)


def test_c_and_python_oracles_agree_on_fresh_signals():
    lib = load_c_oracle()
    rng = np.random.default_rng(7)
    for n in (0, 1, 17, 19, 20, 257, 5000):
        raw = event_ref.synth_read(rng, n) if n else np.zeros(0, np.int32)
        ev = event_ref.detect_events(raw, 6, 9)
        st, ln, mu, sd = c_detect(lib, raw, 6, 9, 1.4, 9.0, 0.2)
        assert np.array_equal(st, ev.start) and np.array_equal(ln, ev.length)
        assert np.array_equal(mu, ev.mean) and np.array_equal(sd, ev.stdv)


def test_events_tile_the_read():
    """Domain property: events are contiguous, first start 0."""
    rng = np.random.default_rng(11)
    raw = event_ref.synth_read(rng, 20000)
    ev = event_ref.detect_events(raw, 6, 9)
    assert ev.start[0] == 0
    assert np.array_equal(ev.start[1:], ev.start[:-1] + ev.length[:-1])
    assert (ev.length > 0).all()
