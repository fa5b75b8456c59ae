"""TensorBoard event files for the training loop's scalar summaries.

Replaces `nn.SummaryWriter(logdir, graph)` + the `scalar_summary` ops of `define_summaries` / `add_summary`
(edgegan/models/edgegan.py:344-411, 427-433, 443; `nn/__init__.py:5-15` aliases tf.summary.*).  The reference
evaluates its merged summaries with two extra graph runs per iteration; here the scalars are the loss values the step
itself produced (`EdgeGAN.read_losses`), written under the reference's tags.  Image and histogram summaries are not
written.

File format (tensorflow/core/lib/io/record_writer.cc, tensorflow/core/util/event.proto, framework/summary.proto):
  events.out.tfevents.<unix seconds>.<hostname>  = a sequence of records
  record  = uint64 length | uint32 masked_crc32c(length bytes) | data | uint32 masked_crc32c(data)      (little endian)
  data    = Event { double wall_time = 1; int64 step = 2; string file_version = 3 | Summary summary = 5 }
  Summary = repeated Value value = 1 { string tag = 1; float simple_value = 2 }
The first record carries file_version "brain.Event:2".  Checksums: edgegan_b200.checkpoint.crc32c (native).
"""
from __future__ import annotations

import os
import socket
import struct
import time

from .checkpoint import _pb_bytes, _pb_fields, _pb_varint, _put_varint, crc32c, mask_crc, unmask_crc

# tags of the reference's scalar summaries (edgegan.py:351-366, 384-387, 401-404) -> EdgeGAN.read_losses() keys
SCALAR_TAGS = {
    "edge_gloss": "edge_gloss", "image_gloss": "image_gloss", "joint_dis_dloss": "joint_dis_dloss",
    "zl_loss": "zl_loss", "loss_g_ac": "loss_g_ac", "loss_d_ac": "loss_d_ac",
    "image_dis_dloss": "image_dis_dloss", "edge_dis_dloss": "edge_dis_dloss",
}


def _record(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data)))


def encode_event(wall_time, step=0, file_version=None, scalars=None) -> bytes:
    out = bytearray()
    out += bytes([(1 << 3) | 1]) + struct.pack("<d", float(wall_time))
    if step:
        _pb_varint(out, 2, int(step))
    if file_version is not None:
        _pb_bytes(out, 3, file_version.encode())
    if scalars:
        summary = bytearray()
        for tag, value in scalars.items():
            val = bytearray()
            _pb_bytes(val, 1, tag.encode())
            _put_varint(val, (2 << 3) | 5)
            val += struct.pack("<f", float(value))
            _pb_bytes(summary, 1, bytes(val))
        _pb_bytes(out, 5, bytes(summary))
    return bytes(out)


class SummaryWriter:
    def __init__(self, logdir, graph=None, filename_suffix=""):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s%s" % (int(time.time()), socket.gethostname(), filename_suffix))
        self._fh = open(self.path, "wb")
        self._fh.write(_record(encode_event(time.time(), file_version="brain.Event:2")))
        self._fh.flush()

    def add_scalars(self, scalars: dict, global_step: int):
        self._fh.write(_record(encode_event(time.time(), step=global_step, scalars=scalars)))

    def add_losses(self, losses: dict, global_step: int):
        """losses: EdgeGAN.read_losses(); written under the reference's tags"""
        self.add_scalars({tag: losses[key] for tag, key in SCALAR_TAGS.items() if key in losses}, global_step)

    def flush(self):
        self._fh.flush()

    def close(self):
        self._fh.close()


def read_events(path, verify=True):
    """-> list of {'wall_time', 'step', 'file_version', 'scalars': {tag: value}} (what this module writes)."""
    buf = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(buf):
        (n,) = struct.unpack_from("<Q", buf, pos)
        (c1,) = struct.unpack_from("<I", buf, pos + 8)
        data = buf[pos + 12:pos + 12 + n]
        (c2,) = struct.unpack_from("<I", buf, pos + 12 + n)
        if verify and (unmask_crc(c1) != crc32c(buf[pos:pos + 8]) or unmask_crc(c2) != crc32c(data)):
            raise ValueError(f"{path}: record checksum mismatch at byte {pos}")
        pos += 16 + n
        ev = {"wall_time": None, "step": 0, "file_version": None, "scalars": {}}
        for f, wt, v in _pb_fields(data):
            if f == 1:
                ev["wall_time"] = struct.unpack("<d", struct.pack("<Q", v))[0]
            elif f == 2:
                ev["step"] = v
            elif f == 3:
                ev["file_version"] = v.decode()
            elif f == 5:
                for f2, _, val in _pb_fields(v):
                    if f2 == 1:
                        tag, x = None, None
                        for f3, _, v3 in _pb_fields(val):
                            if f3 == 1:
                                tag = v3.decode()
                            elif f3 == 2:
                                x = struct.unpack("<f", struct.pack("<I", v3))[0]
                        ev["scalars"][tag] = x
        out.append(ev)
    return out
