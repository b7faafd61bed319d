"""Reader (and minimal writer) for TensorFlow "TF-format" checkpoints, without TensorFlow.

The reference restores trained weights with Keras `model.load_weights(prefix)`
(ravvent_performance_evaluator.py:107, ravvent.py:57-59) from the files
`ModelCheckpoint(save_weights_only=True)` wrote (ravvent.py:61-70):

    <prefix>.index                    an SSTable (TensorFlow's copy of the LevelDB table format)
    <prefix>.data-00000-of-00001      the raw little-endian tensor bytes

This module restates the two published formats:

  * table file: data blocks of prefix-compressed entries
        varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value
    followed by a uint32 restart array and its length; every block carries a 5-byte trailer
    (compression type, masked CRC-32C); the 48-byte footer holds the metaindex and index
    block handles (varint64 offset, size) and the magic 0xdb4775248b80fb57.
  * tensor bundle: key "" -> BundleHeaderProto {num_shards=1, endianness=2, version=3},
    every other key -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5,
    crc32c=6 (masked CRC-32C of the tensor bytes)}.

Keras object-graph checkpoints name a variable by its attribute path from the model plus
"/.ATTRIBUTES/VARIABLE_VALUE" (SURVEY Appendix A.6); `map_keras_keys` translates those
paths into the flat interchange names of weights.py.

PARITY STATUS: no TensorFlow and no TF-written checkpoint exist in the build container, so the
reader is checked against this module's own writer (round trip, multi-block tables, CRCs) and
against hand-assembled byte strings only.  The object-graph key patterns are the ones the
reference's attribute names imply; `map_keras_keys` reports every key it could not place.
"""
from __future__ import annotations

import re
import struct
from pathlib import Path

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER = 5
_MASK_DELTA = 0xA282EAD8

# tensorflow DataType enum -> numpy
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           19: np.float16}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------- CRC-32C (Castagnoli), slicing-by-8
def _make_tables():
    poly = 0x82F63B78
    t0 = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t0.append(c)
    tables = [t0]
    for k in range(1, 8):
        prev = tables[k - 1]
        tables.append([(prev[n] >> 8) ^ t0[prev[n] & 0xFF] for n in range(256)])
    return tables


_T = _make_tables()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    mv = memoryview(data)
    n8 = len(mv) // 8
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    if n8:
        for (w,) in struct.iter_unpack("<Q", mv[:8 * n8]):
            w ^= c
            c = (t7[w & 0xFF] ^ t6[(w >> 8) & 0xFF] ^ t5[(w >> 16) & 0xFF] ^ t4[(w >> 24) & 0xFF] ^
                 t3[(w >> 32) & 0xFF] ^ t2[(w >> 40) & 0xFF] ^ t1[(w >> 48) & 0xFF] ^ t0[(w >> 56) & 0xFF])
    for b in mv[8 * n8:]:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------- varints / protobuf wire format
def _get_varint(buf, pos):
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> list of (field, wire_type, value); value is int (varint / fixed) or bytes (length-delimited)."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise ValueError("truncated protobuf field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((field, wt, v))
    return out


def _parse_shape(buf):
    dims = []
    for field, wt, v in _parse_proto(buf):
        if field == 2 and wt == 2:                      # Dim { int64 size = 1; string name = 2 }
            size = 0
            for f2, w2, v2 in _parse_proto(v):
                if f2 == 1 and w2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
        elif field == 3 and wt == 0 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, wt, v in _parse_proto(buf):
        if field == 1 and wt == 0:
            e["dtype"] = v
        elif field == 2 and wt == 2:
            e["shape"] = _parse_shape(v)
        elif field == 3 and wt == 0:
            e["shard_id"] = v
        elif field == 4 and wt == 0:
            e["offset"] = v
        elif field == 5 and wt == 0:
            e["size"] = v
        elif field == 6 and wt == 5:
            e["crc32c"] = v
        elif field == 7:
            e["sliced"] = True
    return e


# ---------------------------------------------------------------- table (.index) reader
def _read_block(data, offset, size, verify):
    end = offset + size
    if end + BLOCK_TRAILER > len(data):
        raise ValueError("block handle points outside the index file")
    ctype = data[end]
    if verify:
        stored = struct.unpack_from("<I", data, end + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:end + 1]):
            raise ValueError("index block checksum mismatch")
    if ctype != 0:
        raise NotImplementedError("compressed index blocks (snappy) are not supported; TensorFlow's BundleWriter writes them uncompressed")
    return data[offset:end]


def _block_entries(block):
    if len(block) < 4:
        raise ValueError("index block too short")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise ValueError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise ValueError("corrupt index block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """-> ordered list of (key bytes, value bytes) of an SSTable file."""
    data = Path(path).read_bytes()
    if len(data) < FOOTER_LEN:
        raise ValueError(f"{path}: too short to be a table file")
    footer = data[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow table file (bad magic)")
    pos = 0
    _mi_off, pos = _get_varint(footer, pos)
    _mi_size, pos = _get_varint(footer, pos)
    ix_off, pos = _get_varint(footer, pos)
    ix_size, pos = _get_varint(footer, pos)
    out = []
    for _sep, handle in _block_entries(_read_block(data, ix_off, ix_size, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ---------------------------------------------------------------- tensor bundle reader
def read_bundle(prefix, verify=True):
    """Read every numeric tensor of the checkpoint `<prefix>.index` / `<prefix>.data-*`.
    -> dict key (str) -> numpy array.  String tensors (the serialized object graph) are skipped."""
    prefix = str(prefix)
    entries = read_table(prefix + ".index", verify)
    if not entries or entries[0][0] != b"":
        raise ValueError("checkpoint index has no header entry")
    num_shards, endianness = 1, 0
    for field, wt, v in _parse_proto(entries[0][1]):
        if field == 1 and wt == 0:
            num_shards = v
        elif field == 2 and wt == 0:
            endianness = v
    if endianness != 0:
        raise NotImplementedError("big-endian checkpoint")
    shards = {}
    out = {}
    for key, value in entries[1:]:
        e = _parse_entry(value)
        if e["dtype"] not in _DTYPES:                   # DT_STRING etc.: not a weight
            continue
        if e["sliced"]:
            raise NotImplementedError(f"partitioned variable {key!r}")
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = Path(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}").read_bytes()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if len(raw) != e["size"] or e["size"] != n * dt.itemsize:
            raise ValueError(f"{key!r}: size {e['size']} does not match shape {e['shape']} / data file")
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError(f"{key!r}: tensor checksum mismatch")
        out[key.decode("utf-8")] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).reshape(e["shape"]).astype(dt)
    return out


# ---------------------------------------------------------------- Keras object-graph keys -> interchange names
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_VAR = r"(kernel|recurrent_kernel|bias)"
_PATTERNS = [
    (re.compile(rf"^encoder_(raw|event)/rnn_layers/(\d+)/(forward|backward)_layer/cell/{_VAR}$"),
     lambda m: f"encoder_{m[1]}/layer{m[2]}/{m[3]}/{m[4]}"),
    # unidirectional encoders (rnn_type 'lstm' / 'gru'): tf.keras.layers.RNN(cell) without the Bidirectional wrapper
    (re.compile(rf"^encoder_(raw|event)/rnn_layers/(\d+)/cell/{_VAR}$"), lambda m: f"encoder_{m[1]}/layer{m[2]}/forward/{m[3]}"),
    (re.compile(rf"^decoder/decoder_rnn_cell/cells/(\d+)/{_VAR}$"), lambda m: f"decoder/cell{m[1]}/{m[2]}"),
    (re.compile(rf"^decoder/rnn_cell/_cell/cells/(\d+)/{_VAR}$"), lambda m: f"decoder/cell{m[1]}/{m[2]}"),
    (re.compile(r"^decoder/attention_mechanism/memory_layer/kernel$"), lambda m: "decoder/memory_layer/kernel"),
    (re.compile(r"^decoder/rnn_cell/_attention_mechanisms/0/memory_layer/kernel$"), lambda m: "decoder/memory_layer/kernel"),
    (re.compile(r"^decoder/rnn_cell/_attention_layers/0/kernel$"), lambda m: "decoder/attention_layer/kernel"),
    (re.compile(r"^decoder/fc/(kernel|bias)$"), lambda m: f"decoder/fc/{m[1]}"),
    (re.compile(r"^decoder/decoder/output_layer/(kernel|bias)$"), lambda m: f"decoder/fc/{m[1]}"),
]


def map_keras_keys(tensors):
    """Checkpoint key -> interchange name (weights.py).  Optimizer slots, counters and the object graph are dropped.
    -> (weights dict, list of keys that look like model variables but matched no pattern)."""
    out, unmapped = {}, []
    for key, arr in tensors.items():
        if not key.endswith(_SUFFIX) or ".OPTIMIZER_SLOT" in key:
            continue
        path = key[:-len(_SUFFIX)]
        if path.startswith("optimizer") or path in ("save_counter", "_CHECKPOINTABLE_OBJECT_GRAPH"):
            continue
        for rx, name in _PATTERNS:
            m = rx.match(path)
            if m:
                out.setdefault(name(m), np.asarray(arr, dtype=np.float32))
                break
        else:
            unmapped.append(key)
    return out, unmapped


def load_keras_checkpoint(prefix, verify=True):
    """`<prefix>` as passed to Keras load_weights -> interchange weight dict for Basecaller.load_weights."""
    w, unmapped = map_keras_keys(read_bundle(prefix, verify))
    if not w:
        raise ValueError(f"{prefix}: no Ravvent model variables found" +
                         (f"; unrecognised keys: {unmapped[:8]}" if unmapped else ""))
    if unmapped:
        raise ValueError(f"{prefix}: unrecognised model variables {unmapped[:8]} (extend tf_checkpoint._PATTERNS)")
    return w


def is_checkpoint_prefix(path) -> bool:
    return Path(str(path) + ".index").is_file()


# ---------------------------------------------------------------- writer (export / test fixtures)
def _proto_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v & ((1 << 64) - 1))


def _proto_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count and self.count % self.interval == 0:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _emit_block(out, contents):
    offset = len(out)
    out += contents + b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))
    return _put_varint(offset) + _put_varint(len(contents))


def write_table(path, items, block_size=4096):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order."""
    out, index, blk, prev = bytearray(), _BlockBuilder(1), _BlockBuilder(), None
    for key, value in items:
        if prev is not None and key <= prev:
            raise ValueError("table keys must be strictly increasing")
        blk.add(key, value)
        prev = key
        if len(blk.buf) >= block_size:
            index.add(prev, _emit_block(out, blk.finish()))
            blk = _BlockBuilder()
    if blk.count:
        index.add(prev, _emit_block(out, blk.finish()))
    meta = _emit_block(out, _BlockBuilder().finish())
    idx = _emit_block(out, index.finish())
    footer = meta + idx
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    Path(path).write_bytes(bytes(out))


def write_bundle(prefix, tensors, block_size=4096):
    """tensors: dict key -> numeric numpy array.  Writes `<prefix>.index` and `<prefix>.data-00000-of-00001`."""
    prefix = str(prefix)
    header = _proto_varint(1, 1) + _proto_varint(2, 0) + _proto_bytes(3, _proto_varint(1, 1))
    items, data = [(b"", header)], bytearray()
    for key in sorted(tensors, key=lambda k: k.encode("utf-8")):
        a = np.asarray(tensors[key])                     # (ascontiguousarray would promote 0-d to 1-d)
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes(order="C")
        shape = b"".join(_proto_bytes(2, _proto_varint(1, d)) for d in a.shape)
        entry = _proto_varint(1, _DTYPE_IDS[np.dtype(a.dtype)]) + _proto_bytes(2, shape)
        if len(data):
            entry += _proto_varint(4, len(data))
        entry += _proto_varint(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack("<I", mask_crc(crc32c(raw)))
        items.append((key.encode("utf-8"), entry))
        data += raw
    write_table(prefix + ".index", items, block_size)
    Path(prefix + ".data-00000-of-00001").write_bytes(bytes(data))


def export_keras_checkpoint(prefix, weights):
    """Inverse of load_keras_checkpoint: interchange dict -> TF-format checkpoint with the reference's object-graph keys."""
    out = {}
    for name, arr in weights.items():
        parts = name.split("/")
        if parts[0].startswith("encoder_"):
            key = f"{parts[0]}/rnn_layers/{parts[1][5:]}/{parts[2]}_layer/cell/{parts[3]}"
        elif parts[1].startswith("cell"):
            key = f"decoder/decoder_rnn_cell/cells/{parts[1][4:]}/{parts[2]}"
        elif parts[1] == "memory_layer":
            key = "decoder/attention_mechanism/memory_layer/kernel"
        elif parts[1] == "attention_layer":
            key = "decoder/rnn_cell/_attention_layers/0/kernel"
        else:
            key = f"decoder/fc/{parts[2]}"
        out[key + _SUFFIX] = np.asarray(arr, dtype=np.float32)
    write_bundle(prefix, out)
