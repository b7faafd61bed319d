"""CPU-side checks of the product's boundary: the C-ABI library loads, exports every
symbol include/ravvent_b200.h declares, refuses to compute without a GPU, and the host
layer mirrors the reference's class surface."""
import inspect
import re

import numpy as np
import pytest

from conftest import ROOT, _has_gpu


def test_library_exports_every_declared_symbol():
    import ctypes
    header = (ROOT / "include" / "ravvent_b200.h").read_text()
    declared = set(re.findall(r"\b(rvb_[a-z0-9_]+)\s*\(", header))
    declared -= {"rvb_model_t"}
    assert len(declared) >= 17
    lib = ctypes.CDLL(str(ROOT / "ravvent_basecaller_b200" / "libravvent_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from ravvent_basecaller_b200 import _lib
    assert set(_lib.EXPORTS) == declared


def test_no_cpu_fallback():
    import ravvent_basecaller_b200 as rb
    if _has_gpu():
        pytest.skip("GPU present")
    assert rb.device_count() == 0
    with pytest.raises(rb.RavventError):
        rb.EventDetector(6, 9)
    with pytest.raises(rb.RavventError):
        rb.Basecaller(128, 128, 128, rb.nuc_tk, 'joint', 0.)
    with pytest.raises(rb.RavventError):
        rb.Merger()


def test_product_does_not_import_oracle():
    pkg = ROOT / "ravvent_basecaller_b200"
    for f in list(pkg.glob("*.py")) + [q for q in (pkg / "csrc").glob("*") if q.is_file()]:
        txt = f.read_text()
        assert "oracle" not in txt.replace("oracle/model_ref.py init_weights", ""), f


def test_basecaller_signature_matches_reference():
    """basecaller.py:158 -- same positional order and defaults."""
    import ravvent_basecaller_b200 as rb
    sig = inspect.signature(rb.Basecaller.__init__)
    names = list(sig.parameters)[1:13]
    assert names == ['enc_units', 'dec_units', 'batch_sz', 'tokenizer', 'input_data_type', 'input_padding_value',
                     'encoder_depth', 'decoder_depth', 'rnn_type', 'teacher_forcing', 'attention_type', 'beam_width']
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d['encoder_depth'], d['decoder_depth'], d['rnn_type'], d['attention_type'], d['beam_width']) == (2, 1, 'bilstm', 'luong', 5)
    for m in ('compile', 'load_weights', '_encode_input', 'greedy_search_prediction', 'beam_search_prediction',
              'tokens_to_nuc_sequences'):
        assert callable(getattr(rb.Basecaller, m))
    ed = inspect.signature(rb.EventDetector.__init__)
    assert [ed.parameters[k].default for k in ('window_length1', 'window_length2', 'threshold1', 'threshold2', 'peak_height')] == [3, 6, 1.4, 9., 0.2]


def test_random_weights_match_oracle_initialiser():
    from oracle import model_ref
    from ravvent_basecaller_b200 import weights
    a, b = weights.random_weights(22), model_ref.init_weights(22)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert sum(v.size for v in a.values()) == 1276807          # SURVEY §8a model size


def test_tokenizer_and_postprocessing():
    from oracle import model_ref
    from ravvent_basecaller_b200 import data_loader as dl
    assert dl.nuc_tk.word_index == model_ref.VOCAB
    ids = np.array([[3, 4, 5, 6, 1, 1], [2, 6, 0, 3, 1, 4]])
    txt = [t.replace(' ', '').replace('^', '').replace('$', '').upper() for t in dl.nuc_tk.sequences_to_texts(ids)]
    assert txt == model_ref.tokens_to_nuc_sequences(ids) == ["ACGT", "TAC"]
    s = np.log(np.array([[0.5, 0.25, 0.125]]))
    np.testing.assert_allclose(dl.calc_prob_logits_beam_search_scores(s), model_ref.beam_scores_to_probs(s))
    np.testing.assert_allclose(dl.calc_prob_logits_beam_search_scores(s), [[0.5, 0.5, 0.5]])
