#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own modules.

Runs only in the build container (needs /root/reference, which does not exist
on the GPU box).  The fixtures it writes are committed; tests never read
/root/reference.

  * event_golden.npz    : EventDetector(w1,w2,...).run(raw) of
                          /root/reference/event_detection/event_detector.py
  * snippets_golden.npz : data_loader.prepare_snippets + pad_input_snippets of
                          /root/reference/data_loader.py with `tensorflow` /
                          `keras` stubbed (only pad_sequences is emulated, as
                          documented Keras behaviour; sklearn is the real one).

The NN path (basecaller.py) needs tensorflow + tensorflow_addons, which are
not installable here: no golden vectors exist for it ("parity unpinned").
"""
import sys
import types
from pathlib import Path

import numpy as np

REF = "/root/reference"
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.event_ref import synth_read  # noqa: E402  (signal generator only)


def _stub_tf_keras():
    def pad_sequences(seqs, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
        seqs = [np.asarray(s) for s in seqs]
        if maxlen is None:
            maxlen = max(len(s) for s in seqs)
        tail = seqs[0].shape[1:] if seqs else ()
        out = np.full((len(seqs), maxlen) + tuple(tail), value, dtype=dtype)
        for i, s in enumerate(seqs):
            assert padding == "post" and (truncating == "post" or len(s) <= maxlen)
            s = s[:maxlen]
            out[i, :len(s)] = s
        return out

    class Tokenizer:                       # char_level=True, lower=True, filters='' as configured at data_loader.py:20
        def __init__(self, **kw):
            self.word_index, self.index_word = {}, {}

        def texts_to_sequences(self, texts):
            return [[self.word_index[c] for c in t.lower() if c in self.word_index] for t in texts]

    tf = types.ModuleType("tensorflow")
    tf.keras = types.ModuleType("tensorflow.keras")
    tf.keras.utils = types.SimpleNamespace(Sequence=object)
    mods = {
        "tensorflow": tf, "tensorflow.keras": tf.keras,
        "tensorflow.keras.preprocessing": types.ModuleType("tensorflow.keras.preprocessing"),
        "tensorflow.keras.preprocessing.sequence": types.ModuleType("tensorflow.keras.preprocessing.sequence"),
        "keras": types.ModuleType("keras"),
        "keras.preprocessing": types.ModuleType("keras.preprocessing"),
        "keras.preprocessing.text": types.ModuleType("keras.preprocessing.text"),
    }
    mods["tensorflow.keras.preprocessing.sequence"].pad_sequences = pad_sequences
    mods["keras.preprocessing.text"].Tokenizer = Tokenizer
    sys.modules.update(mods)


def event_golden(out: Path):
    sys.path.insert(0, REF)
    from event_detection.event_detector import EventDetector
    rng = np.random.default_rng(20261018)
    cases = [  # (n_samples, w1, w2, thr1, thr2, peak_height)
        (12000, 6, 9, 1.4, 9.0, 0.2),     # the data path's setting (data_loader.py:12-13,71)
        (6000, 3, 6, 1.4, 9.0, 0.2),      # EventDetector defaults (event_detector.py:27-28)
        (6000, 6, 9, 2.0, 6.0, 0.5),
        (4000, 4, 4, 1.4, 9.0, 0.2),      # w1 == w2: the long detector masks itself
        (4000, 5, 12, 1.4, 9.0, 0.2),
        (3000, 2, 9, 1.4, 9.0, 0.2),      # u32-wrapped first event swallows the read
        (4000, 9, 6, 1.4, 9.0, 0.2),      # w1 > w2
        (40, 6, 9, 1.4, 9.0, 0.2),
        (18, 6, 9, 1.4, 9.0, 0.2),
        (1, 6, 9, 1.4, 9.0, 0.2),
        (0, 6, 9, 1.4, 9.0, 0.2),
    ]
    blob = {"n_cases": np.int64(len(cases))}
    for k, (n, w1, w2, t1, t2, ph) in enumerate(cases):
        raw = synth_read(rng, n) if n else np.zeros(0, np.int32)
        if k == 2:  # flat stretches / constant signal: variance floor path
            raw[1000:1400] = 400
        ev = EventDetector(w1, w2, t1, t2, ph).run(raw.astype(int))
        blob[f"raw_{k}"] = raw.astype(np.int16)
        blob[f"par_{k}"] = np.array([w1, w2, t1, t2, ph], dtype=np.float64)
        blob[f"start_{k}"] = np.array([e.start for e in ev], dtype=np.int64)
        blob[f"length_{k}"] = np.array([e.length for e in ev], dtype=np.int64)
        blob[f"mean_{k}"] = np.array([e.mean for e in ev], dtype=np.float64)
        blob[f"stdv_{k}"] = np.array([e.stdv for e in ev], dtype=np.float64)
        print(f"event case {k}: n={n} w=({w1},{w2}) -> {len(ev)} events")
    np.savez_compressed(out, **blob)


def snippets_golden(out: Path):
    _stub_tf_keras()
    sys.path.insert(0, REF)
    import data_loader as dl
    rng = np.random.default_rng(22)
    blob = {}
    for k, (n, lab0, lab1, stride) in enumerate([(9000, 0, 9000, 6), (5000, 37, 4800, 6), (4000, 0, 4000, 4)]):
        raw = synth_read(rng, n)
        # label rows only contribute their first start / last end to the inference outputs
        edges = np.arange(lab0, lab1 + 1, 8)
        edges[-1] = lab1
        ranges = np.column_stack((edges[:-1], edges[1:])).astype(int)
        syms = rng.choice(np.array(list("ACGT"), dtype=object), size=ranges.shape[0])
        raw_s, ev_s, _ = dl.prepare_snippets(raw.astype(int), ranges, syms, stride)
        blob[f"raw_{k}"] = raw.astype(np.int16)
        blob[f"par_{k}"] = np.array([lab0, lab1, stride], dtype=np.int64)
        blob[f"raw_snips_{k}"] = dl.pad_input_snippets(raw_s, dl.MAX_RAW_LEN)
        blob[f"event_snips_{k}"] = dl.pad_input_snippets(ev_s, dl.MAX_EVENT_LEN)
        # the on-disk entry point with the same read: also pins the target token rows
        import tempfile, os
        with tempfile.TemporaryDirectory() as td:
            sp, lp = os.path.join(td, "r.signal"), os.path.join(td, "r.label")
            np.savetxt(sp, raw.reshape(1, -1), fmt="%d")
            with open(lp, "w") as f:
                for (a, b), c in zip(ranges, syms):
                    f.write(f"{a} {b} {c}\n")
            r3, e3, tk = dl.load_data_from_single_signal_label(sp, lp, stride)
        assert np.array_equal(r3, blob[f"raw_snips_{k}"])
        blob[f"tokens_{k}"] = tk.astype(np.int64)
        blob[f"label_ranges_{k}"] = ranges.astype(np.int64)
        blob[f"label_syms_{k}"] = np.array([ord(c) for c in syms], dtype=np.uint8)
        blob[f"raw_lens_{k}"] = np.array([len(s) for s in raw_s], dtype=np.int64)
        blob[f"event_lens_{k}"] = np.array([len(s) for s in ev_s], dtype=np.int64)
        print(f"snippet case {k}: n={n} -> {len(raw_s)} snippets, raw len {blob[f'raw_lens_{k}'].min()}-"
              f"{blob[f'raw_lens_{k}'].max()}, events {blob[f'event_lens_{k}'].min()}-{blob[f'event_lens_{k}'].max()}")
    blob["n_cases"] = np.int64(3)
    np.savez_compressed(out, **blob)


if __name__ == "__main__":
    gold = ROOT / "tests" / "golden"
    gold.mkdir(parents=True, exist_ok=True)
    event_golden(gold / "event_golden.npz")
    snippets_golden(gold / "snippets_golden.npz")
