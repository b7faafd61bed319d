"""Greedy validation step (Basecaller.test_step, reference basecaller.py:264-279) -- SURVEY §8 f-4."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import model_ref as mr
from ravvent_basecaller_b200.basecaller import Basecaller
from ravvent_basecaller_b200.data_loader import masked_accuracy


def test_masked_accuracy_hand_case():
    y_true = np.array([[3, 4, 1, 0], [5, 2, 6, 0]])
    y_pred = np.array([[3, 5, 1, 0], [5, 2, 3, 9]])
    # omit start(2) / end(1): 6 positions count (padding is NOT omitted, as in the reference's _val_step), 3 match
    assert masked_accuracy(y_true, y_pred, [2, 1]) == pytest.approx(3 / 6)
    assert mr.masked_accuracy(y_true, y_pred, [2, 1]) == pytest.approx(3 / 6)
    assert masked_accuracy(y_true, y_pred, [0, 2, 1]) == pytest.approx(2 / 4)


def test_loss_function_hand_case():
    me = SimpleNamespace(output_padding_token=np.int32(0))
    logits = np.zeros((1, 3, 7), np.float32)
    logits[0, 0, 3] = 2.0
    real = np.array([[3, 4, 0]])
    want = ((np.log(6 + np.exp(2.0)) - 2.0) + np.log(7.0)) / 2.0          # padding position excluded
    assert Basecaller.loss_function(me, real, logits) == pytest.approx(want, rel=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["joint", "raw"])
def test_test_step_matches_oracle(kind):
    import ravvent_basecaller_b200 as rb
    rng = np.random.default_rng(9)
    w = mr.init_weights(22)
    raw, ev = mr.synth_chunks(rng, 24)
    L = 14
    target = rng.integers(3, 7, size=(24, L)).astype(np.int32)
    target[:, 0] = 2
    for b in range(24):                                        # ragged targets: end token then padding
        n = int(rng.integers(5, L - 1))
        target[b, n] = 1
        target[b, n + 1:] = 0
    bc = rb.Basecaller(128, 128, 24, rb.nuc_tk, kind, 0.0).load_weights(w)
    got = bc.test_step((raw, ev, target))
    enc, mask = mr.encode_input(w, (raw, ev) if kind == "joint" else raw, kind)
    want = mr.val_step(w, enc, mask, target)
    assert got["acc"] == pytest.approx(want["acc"], abs=1e-12)
    assert got["loss"] == pytest.approx(want["loss"], rel=1e-5)
