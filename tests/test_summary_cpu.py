"""TensorBoard event files (edgegan_b200/summary.py; reference: nn.SummaryWriter + scalar_summary,
edgegan/models/edgegan.py:344-411,443).  The record framing and the Event / Summary wire format are pinned by bytes
assembled by hand from the TFRecord and proto definitions."""
import struct

from edgegan_b200 import checkpoint as ck
from edgegan_b200 import summary as sm


def test_event_bytes_assembled_by_hand():
    # Event { wall_time = 2.5 (field 1, fixed64); step = 7 (field 2); summary (field 5) { value { tag "zl_loss"; simple_value 0.5 } } }
    value = bytes([0x0A, 7]) + b"zl_loss" + bytes([0x15]) + struct.pack("<f", 0.5)
    summary = bytes([0x0A, len(value)]) + value
    want = bytes([0x09]) + struct.pack("<d", 2.5) + bytes([0x10, 7]) + bytes([0x2A, len(summary)]) + summary
    assert sm.encode_event(2.5, step=7, scalars={"zl_loss": 0.5}) == want
    first = bytes([0x09]) + struct.pack("<d", 1.0) + bytes([0x1A, 13]) + b"brain.Event:2"
    assert sm.encode_event(1.0, file_version="brain.Event:2") == first
    rec = sm._record(first)
    assert rec[:8] == struct.pack("<Q", len(first))
    assert struct.unpack("<I", rec[8:12])[0] == ck.mask_crc(ck.crc32c(rec[:8]))
    assert rec[12:-4] == first and struct.unpack("<I", rec[-4:])[0] == ck.mask_crc(ck.crc32c(first))


def test_writer_round_trip_and_corruption(tmp_path):
    w = sm.SummaryWriter(str(tmp_path / "logs"))
    losses = {"joint_dis_dloss": 1.5, "image_dis_dloss": -2.0, "edge_dis_dloss": 0.25, "loss_d_ac": 0.0, "edge_gloss": 3.0,
              "image_gloss": 4.0, "zl_loss": 0.125, "loss_g_ac": 0.0, "edge_gloss_b": 9.0}
    w.add_losses(losses, 1)
    w.add_losses(losses, 2)
    w.close()
    ev = sm.read_events(w.path)
    assert ev[0]["file_version"] == "brain.Event:2" and [e["step"] for e in ev] == [0, 1, 2]
    assert ev[1]["scalars"] == {t: losses[k] for t, k in sm.SCALAR_TAGS.items()}
    assert "edge_gloss_b" not in ev[1]["scalars"]                       # not a summary of the reference
    raw = bytearray(open(w.path, "rb").read())
    raw[40] ^= 1
    open(w.path, "wb").write(bytes(raw))
    try:
        sm.read_events(w.path)
        assert False
    except ValueError as e:
        assert "checksum" in str(e)
