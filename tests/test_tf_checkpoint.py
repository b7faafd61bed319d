"""TF-format checkpoint reader (ravvent_basecaller_b200/tf_checkpoint.py) -- SURVEY §8 f-2.

No TensorFlow here: the reader is checked against known-answer CRC-32C values, a table file assembled
byte by byte in this test from the published LevelDB/TensorFlow table layout, and round trips through
the module's own writer (multi-block index, prefix compression, restart points)."""
import struct

import numpy as np
import pytest

from ravvent_basecaller_b200 import tf_checkpoint as tc
from ravvent_basecaller_b200 import weights as W


def test_crc32c_known_answers():
    assert tc.crc32c(b"123456789") == 0xE3069283                 # the CRC-32C check value
    assert tc.crc32c(b"\x00" * 32) == 0x8A9136AA                 # RFC 3720 B.4 test patterns
    assert tc.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tc.crc32c(bytes(range(32))) == 0x46DD794E
    assert tc.crc32c(b"hello world") == tc.crc32c(b" world", tc.crc32c(b"hello"))   # incremental form
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert tc.unmask_crc(tc.mask_crc(v)) == v
    assert tc.mask_crc(0) == 0xA282EAD8


def _block(entries, restarts):
    body = b"".join(entries)
    return body + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def _with_trailer(contents):
    return contents + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(contents + b"\x00")))


def test_hand_assembled_table(tmp_path):
    # one data block, keys "ab" -> "1", "abc" -> "22" (shares 2 bytes), "b" -> "" ; entry = shared, non_shared, vlen, key, value
    e = [bytes([0, 2, 1]) + b"ab" + b"1", bytes([2, 1, 2]) + b"c" + b"22", bytes([0, 1, 0]) + b"b"]
    data_blk = _block(e, [0])
    meta_blk = _block([], [0])
    f = bytearray()
    f += _with_trailer(data_blk)
    meta_off = len(f)
    f += _with_trailer(meta_blk)
    handle = bytes([0, len(data_blk)])
    index_blk = _block([bytes([0, 1, len(handle)]) + b"b" + handle], [0])
    idx_off = len(f)
    f += _with_trailer(index_blk)
    footer = bytes([meta_off, len(meta_blk), idx_off, len(index_blk)])
    f += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", tc.TABLE_MAGIC)
    p = tmp_path / "t.index"
    p.write_bytes(bytes(f))
    assert tc.read_table(p) == [(b"ab", b"1"), (b"abc", b"22"), (b"b", b"")]
    bad = bytearray(f)
    bad[3] ^= 0x40                                               # flip one bit inside the data block
    p.write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tc.read_table(p)
    assert tc.read_table(p, verify=False)[1][1] == b"22"
    p.write_bytes(bytes(f[:-8]) + b"\x00" * 8)
    with pytest.raises(ValueError, match="magic"):
        tc.read_table(p)


def test_bundle_round_trip_multi_block(tmp_path):
    rng = np.random.default_rng(5)
    tensors = {f"layer/{i:03d}/kernel/.ATTRIBUTES/VARIABLE_VALUE": rng.normal(size=(3, i % 5 + 1)).astype(np.float32)
               for i in range(70)}
    tensors["scalar"] = np.array(7, dtype=np.int64)
    tensors["empty"] = np.zeros((0, 4), dtype=np.float32)
    tensors["halves"] = rng.normal(size=(5,)).astype(np.float16)
    tc.write_bundle(tmp_path / "ckpt", tensors, block_size=256)          # many data blocks, restart points every 16 keys
    got = tc.read_bundle(tmp_path / "ckpt")
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    raw = bytearray((tmp_path / "ckpt.data-00000-of-00001").read_bytes())
    raw[10] ^= 1
    (tmp_path / "ckpt.data-00000-of-00001").write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        tc.read_bundle(tmp_path / "ckpt")


@pytest.mark.parametrize("enc_depth,dec_depth", [(2, 1), (3, 2)])
def test_keras_key_mapping_round_trip(tmp_path, enc_depth, dec_depth):
    w = W.random_weights(3, encoder_depth=enc_depth, decoder_depth=dec_depth)
    tc.export_keras_checkpoint(tmp_path / "model_chp", w)
    assert tc.is_checkpoint_prefix(tmp_path / "model_chp") and not tc.is_checkpoint_prefix(tmp_path / "nope")
    back = tc.load_keras_checkpoint(tmp_path / "model_chp")
    assert set(back) == set(w)
    for k in w:
        assert np.array_equal(back[k], w[k]), k


def test_keras_key_mapping_ignores_optimizer_and_reports_unknown():
    s = "/.ATTRIBUTES/VARIABLE_VALUE"
    t = {
        "encoder_raw/rnn_layers/0/forward_layer/cell/kernel" + s: np.zeros((1, 512), np.float32),
        "encoder_raw/rnn_layers/0/forward_layer/cell/kernel/.OPTIMIZER_SLOT/optimizer/m" + s: np.ones((1, 512), np.float32),
        "optimizer/iter" + s: np.array(5, np.int64),
        "save_counter" + s: np.array(1, np.int64),
        "decoder/rnn_cell/_attention_layers/0/kernel" + s: np.zeros((384, 128), np.float32),
        "decoder/attention_mechanism/memory_layer/kernel" + s: np.zeros((256, 128), np.float32),
        "decoder/something_new/kernel" + s: np.zeros((2, 2), np.float32),
    }
    w, unmapped = tc.map_keras_keys(t)
    assert set(w) == {"encoder_raw/layer0/forward/kernel", "decoder/attention_layer/kernel", "decoder/memory_layer/kernel"}
    assert not w["encoder_raw/layer0/forward/kernel"].any()          # the optimizer slot did not overwrite the variable
    assert unmapped == ["decoder/something_new/kernel" + s]


@pytest.mark.gpu
def test_load_weights_from_checkpoint_prefix(tmp_path):
    import ravvent_basecaller_b200 as rb
    from oracle import model_ref as mr
    w = W.random_weights(11)
    tc.export_keras_checkpoint(tmp_path / "model_chp", w)
    tok = rb.nuc_tk
    a = rb.Basecaller(128, 128, 8, tok, "joint", 0.0).load_weights(w)
    b = rb.Basecaller(128, 128, 8, tok, "joint", 0.0).load_weights(str(tmp_path / "model_chp"))
    raw, ev = mr.synth_chunks(np.random.default_rng(2), 8)
    ia, sa = a.beam_search_prediction((raw, ev), 3, 12)
    ib, sb = b.beam_search_prediction((raw, ev), 3, 12)
    assert np.array_equal(ia, ib) and np.array_equal(sa, sb)
    with pytest.raises(FileNotFoundError):
        a.load_weights(str(tmp_path / "missing_prefix"))
