"""TensorFlow tensor-bundle (checkpoint "V2") reader / writer and the `checkpoint` state file.

Replaces `tf.train.Saver().save / .restore` and `tf.train.get_checkpoint_state` as the reference uses them
(edgegan/models/edgegan.py:421,547 create the Saver over ALL global variables; :635-639 save to
`<checkpoint_dir>/EdgeGAN-Model-<step>`; :641-657 read the `checkpoint` state file, restore, and parse the step out
of the file name).  TensorFlow is not available here, so the format is restated from its published definition
(tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{format,block_builder,table_builder}.cc, which follow
LevelDB's table format):

  <prefix>.index                  an SSTable: key "" -> BundleHeaderProto, key <tensor name> -> BundleEntryProto
  <prefix>.data-00000-of-00001    the raw little-endian tensor bytes, back to back (entry.offset / entry.size)

  table   = data blocks, metaindex block, index block, 48-byte footer (two block handles, padding, magic)
  block   = entries (varint shared, varint non_shared, varint value_len, key suffix, value), restart offsets (fixed32
            each), restart count (fixed32); followed on disk by a 5-byte trailer: compression type (0 = none; bundles
            are written uncompressed) and the masked CRC-32C of block + type
  mask(c) = rotr(c, 15) + 0xa282ead8

Checksums run in the native library (`eg_crc32c`, csrc/host_util.cu).  Parity status: unpinned against TensorFlow
itself (no TF in this environment) -- pinned by the format's published constants and known-answer values
(tests/test_checkpoint_cpu.py) and by write -> read round trips.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import struct

import numpy as np

from . import _lib

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
BLOCK_RESTART_INTERVAL = 16      # table::Options default
BLOCK_SIZE = 262144              # table::Options default in TensorFlow

# tensorflow/core/framework/types.proto
DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
      9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"), 22: np.dtype("<u4"), 23: np.dtype("<u8")}
DT_OF = {v: k for k, v in DT.items()}


def crc32c(data, crc=0) -> int:
    buf = data if isinstance(data, (bytes, bytearray)) else memoryview(np.ascontiguousarray(data)).cast("B")
    n = len(buf)
    if n == 0:
        return crc
    if isinstance(buf, bytes):
        return int(_lib.load().eg_crc32c(buf, n, crc))
    arr = np.frombuffer(buf, np.uint8)
    return int(_lib.load().eg_crc32c(C.c_void_p(arr.ctypes.data), n, crc))


def mask_crc(c: int) -> int:
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m: int) -> int:
    r = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf wire format ------------------------------------------------------------------------
def _put_varint(out: bytearray, v: int):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(b, pos):
    v, shift = 0, 0
    while True:
        c = b[pos]
        pos += 1
        v |= (c & 0x7F) << shift
        if not c & 0x80:
            return v, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _pb_fields(b):
    """Yield (field number, wire type, value) of one protobuf message; length-delimited values are bytes."""
    pos, n = 0, len(b)
    while pos < n:
        tag, pos = _get_varint(b, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(b, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", b, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(b, pos)
            v = bytes(b[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", b, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def _pb_varint(out, field, v):
    _put_varint(out, field << 3)
    _put_varint(out, v)


def _pb_bytes(out, field, payload):
    _put_varint(out, (field << 3) | 2)
    _put_varint(out, len(payload))
    out += payload


def encode_shape(shape) -> bytes:
    """TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }."""
    out = bytearray()
    for d in shape:
        dim = bytearray()
        _pb_varint(dim, 1, int(d))      # proto3 would skip a zero; TensorFlow (proto3) does too, and so does this
        if int(d) == 0:
            dim = bytearray()
        _pb_bytes(out, 2, dim)
    return bytes(out)


def decode_shape(b) -> tuple:
    dims = []
    for f, wt, v in _pb_fields(b):
        if f == 2:
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = v2 - (1 << 64) if v2 >> 63 else v2
            dims.append(size)
        elif f == 3 and v:
            raise ValueError("tensor of unknown rank in a checkpoint")
    return tuple(dims)


def encode_header(num_shards=1) -> bytes:
    """BundleHeaderProto { int32 num_shards = 1; Endianness endianness = 2 (LITTLE = 0, omitted); VersionDef version = 3 {producer = 1} }."""
    out = bytearray()
    _pb_varint(out, 1, num_shards)
    ver = bytearray()
    _pb_varint(ver, 1, 1)
    _pb_bytes(out, 3, ver)
    return bytes(out)


def encode_entry(dtype_enum, shape, shard_id, offset, size, crc_masked) -> bytes:
    """BundleEntryProto { dtype = 1; shape = 2; shard_id = 3; offset = 4; size = 5; fixed32 crc32c = 6 } (zero fields omitted)."""
    out = bytearray()
    _pb_varint(out, 1, dtype_enum)
    _pb_bytes(out, 2, encode_shape(shape))
    if shard_id:
        _pb_varint(out, 3, shard_id)
    if offset:
        _pb_varint(out, 4, offset)
    if size:
        _pb_varint(out, 5, size)
    if crc_masked:
        _put_varint(out, (6 << 3) | 5)
        out += struct.pack("<I", crc_masked)
    return bytes(out)


def decode_entry(b) -> dict:
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0, "slices": 0}
    for f, wt, v in _pb_fields(b):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = decode_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


# ---- table (SSTable) ------------------------------------------------------------------------------------------------
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""
        self.empty = True

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.counter < self.interval:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf += key[shared:]
        self.buf += value
        self.last_key = key
        self.counter += 1
        self.empty = False

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        out = bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))
        return out


def _block_with_trailer(block: bytes) -> bytes:
    trailer_type = b"\x00"                      # kNoCompression
    crc = mask_crc(crc32c(trailer_type, crc32c(block)))
    return block + trailer_type + struct.pack("<I", crc)


def _handle(offset, size) -> bytes:
    out = bytearray()
    _put_varint(out, offset)
    _put_varint(out, size)
    return bytes(out)


def write_table(path, items):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order."""
    f = bytearray()
    index = _BlockBuilder(1)
    data = _BlockBuilder(BLOCK_RESTART_INTERVAL)
    prev = None

    def flush():
        nonlocal data
        if data.empty:
            return
        block = data.finish()
        index.add(data.last_key, _handle(len(f), len(block)))     # any key >= the block's last key separates it
        f.extend(_block_with_trailer(block))
        data = _BlockBuilder(BLOCK_RESTART_INTERVAL)

    for k, v in items:
        if prev is not None and not k > prev:
            raise ValueError("table keys must be strictly increasing")
        prev = k
        data.add(k, v)
        if data.size_estimate() >= BLOCK_SIZE:
            flush()
    flush()
    meta = _BlockBuilder(BLOCK_RESTART_INTERVAL).finish()
    meta_handle = _handle(len(f), len(meta))
    f.extend(_block_with_trailer(meta))
    ib = index.finish()
    index_handle = _handle(len(f), len(ib))
    f.extend(_block_with_trailer(ib))
    footer = bytearray(meta_handle + index_handle)
    footer += b"\x00" * (40 - len(footer))
    footer += struct.pack("<Q", TABLE_MAGIC)
    f.extend(footer)
    with open(path, "wb") as fh:
        fh.write(bytes(f))


def _read_block(buf, offset, size, verify=True) -> bytes:
    block = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(bytes([ctype]), crc32c(bytes(block))):
        raise ValueError("checkpoint index: block checksum mismatch")
    if ctype != 0:
        raise ValueError("checkpoint index: compressed table blocks are not supported (tensor bundles are written uncompressed)")
    return bytes(block)


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """-> list of (key, value) in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow checkpoint index (bad table magic)")
    foot = buf[-48:]
    pos = 0
    _, pos = _get_varint(foot, pos)
    _, pos = _get_varint(foot, pos)
    ioff, pos = _get_varint(foot, pos)
    isize, pos = _get_varint(foot, pos)
    out = []
    for _, h in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p2 = _get_varint(h, 0)
        bsize, _ = _get_varint(h, p2)
        out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return out


# ---- bundle ---------------------------------------------------------------------------------------------------------
def _data_path(prefix, shard, num_shards):
    return f"{prefix}.data-{shard:05d}-of-{num_shards:05d}"


def write_bundle(prefix, tensors):
    """Write {name: array} as <prefix>.index + <prefix>.data-00000-of-00001 (float32 / int / bool arrays, little endian)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode())
    entries, offset = [(b"", encode_header(1))], 0
    with open(_data_path(prefix, 0, 1) + ".tmp", "wb") as fh:
        for name in names:
            if not name:
                raise ValueError("empty tensor name")
            a = np.asarray(tensors[name])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in DT_OF:
                raise TypeError(f"{name}: dtype {a.dtype} has no TensorFlow checkpoint encoding here")
            a = np.asarray(a, dtype=dt)              # (ascontiguousarray would turn a scalar into shape (1,))
            raw = a.tobytes(order="C")
            fh.write(raw)
            entries.append((name.encode(), encode_entry(DT_OF[np.dtype(dt)], a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + ".index.tmp", entries)
    os.replace(_data_path(prefix, 0, 1) + ".tmp", _data_path(prefix, 0, 1))
    os.replace(prefix + ".index.tmp", prefix + ".index")


class BundleReader:
    """Read access to a checkpoint prefix: `.keys()`, `.shape(name)`, `.tensor(name)`, `.tensors()`."""

    def __init__(self, prefix, verify=True):
        self.prefix, self.verify = prefix, verify
        items = read_table(prefix + ".index", verify)
        if not items or items[0][0] != b"":
            raise ValueError(f"{prefix}.index: missing bundle header")
        self.num_shards, endian = 1, 0
        for f, _, v in _pb_fields(items[0][1]):
            if f == 1:
                self.num_shards = v
            elif f == 2:
                endian = v
        if endian != 0:
            raise ValueError("big-endian checkpoints are not supported")
        self.entries = {k.decode(): decode_entry(v) for k, v in items[1:]}
        self._files = {}

    def keys(self):
        return list(self.entries)

    def __contains__(self, name):
        return name in self.entries

    def shape(self, name):
        return self.entries[name]["shape"]

    def dtype(self, name):
        return DT[self.entries[name]["dtype"]]

    def tensor(self, name) -> np.ndarray:
        e = self.entries[name]
        if e["slices"]:
            raise ValueError(f"{name}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in DT:
            raise TypeError(f"{name}: TensorFlow dtype enum {e['dtype']} is not supported")
        path = _data_path(self.prefix, e["shard_id"], self.num_shards)
        fh = self._files.get(path)
        if fh is None:
            fh = self._files[path] = open(path, "rb")
        fh.seek(e["offset"])
        raw = fh.read(e["size"])
        if len(raw) != e["size"]:
            raise ValueError(f"{name}: data file truncated")
        if self.verify and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError(f"{name}: tensor checksum mismatch")
        a = np.frombuffer(raw, DT[e["dtype"]])
        want = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if a.size != want:
            raise ValueError(f"{name}: {a.size} elements on disk, shape {e['shape']}")
        return a.reshape(e["shape"]).copy()

    def tensors(self):
        return {k: self.tensor(k) for k in self.entries}

    def close(self):
        for fh in self._files.values():
            fh.close()
        self._files = {}


# ---- the `checkpoint` state file (CheckpointState text proto) ---------------------------------------------------------
def get_checkpoint_state(checkpoint_dir):
    """tf.train.get_checkpoint_state: -> {'model_checkpoint_path': str, 'all_model_checkpoint_paths': [str]} or None."""
    path = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(path):
        return None
    state = {"model_checkpoint_path": None, "all_model_checkpoint_paths": []}
    for line in open(path):
        m = re.match(r'\s*(\w+)\s*:\s*"((?:[^"\\]|\\.)*)"', line)
        if not m:
            continue
        val = m.group(2).encode().decode("unicode_escape")
        if m.group(1) == "model_checkpoint_path":
            state["model_checkpoint_path"] = val
        elif m.group(1) == "all_model_checkpoint_paths":
            state["all_model_checkpoint_paths"].append(val)
    return state if state["model_checkpoint_path"] else None


def update_checkpoint_state(checkpoint_dir, model_checkpoint_path, max_to_keep=5):
    """What Saver.save does to `<dir>/checkpoint`: paths relative to the directory, the newest last, at most
    `max_to_keep` kept (older bundles are deleted, like the Saver default)."""
    rel = os.path.basename(model_checkpoint_path)
    st = get_checkpoint_state(checkpoint_dir) or {"all_model_checkpoint_paths": []}
    paths = [p for p in st["all_model_checkpoint_paths"] if p != rel] + [rel]
    while max_to_keep and len(paths) > max_to_keep:
        old = paths.pop(0)
        for fn in os.listdir(checkpoint_dir):
            if fn == os.path.basename(old) + ".index" or fn.startswith(os.path.basename(old) + ".data-"):
                os.remove(os.path.join(checkpoint_dir, fn))
    with open(os.path.join(checkpoint_dir, "checkpoint"), "w") as f:
        f.write(f'model_checkpoint_path: "{rel}"\n')
        for p in paths:
            f.write(f'all_model_checkpoint_paths: "{p}"\n')


def step_of(ckpt_name) -> int:
    """edgegan.py:651-652: the last run of digits in the checkpoint file name."""
    return int(next(re.finditer(r"(\d+)(?!.*\d)", ckpt_name)).group(0))
