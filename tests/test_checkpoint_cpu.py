"""TensorFlow tensor-bundle checkpoints (edgegan_b200/checkpoint.py; reference: tf.train.Saver at
edgegan/models/edgegan.py:421,547,635-657).  -m "not gpu": the format code is host-side; the CRC runs in the native
library, which loads without a GPU.

TensorFlow cannot run here, so the format is pinned by published constants and hand-assembled bytes:
the CRC-32C vectors of RFC 3720 B.4, LevelDB's mask constant and table magic, and an index file written out by hand in
this test from the protobuf / block layout definitions (independent of the encoder under test)."""
import os
import struct

import numpy as np
import pytest

from edgegan_b200 import checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # incremental == one shot, unaligned starts, numpy input
    data = np.random.RandomState(0).bytes(100003)
    assert ck.crc32c(data[37:], ck.crc32c(data[:37])) == ck.crc32c(data)
    assert ck.crc32c(np.frombuffer(data, np.uint8)[1:]) == ck.crc32c(data[1:])
    assert ck.crc32c(b"") == 0


def test_crc_mask_is_leveldbs():
    # leveldb/util/crc32c.h: Mask(crc) = ((crc >> 15) | (crc << 17)) + 0xa282ead8
    assert ck.mask_crc(0) == 0xA282EAD8
    assert ck.mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    for c in (0, 1, 0xFFFFFFFF, 0x12345678, 0xE3069283):
        assert ck.unmask_crc(ck.mask_crc(c)) == c
        assert ck.mask_crc(c) != c


def _trailer(block):
    return block + b"\x00" + struct.pack("<I", ck.mask_crc(ck.crc32c(b"\x00", ck.crc32c(block))))


def test_index_bytes_of_a_one_tensor_bundle(tmp_path):
    """Bundle {"a": float32 [1.5]} assembled by hand from the format definitions."""
    prefix = str(tmp_path / "m-7")
    ck.write_bundle(prefix, {"a": np.array([1.5], np.float32)})
    raw = struct.pack("<f", 1.5)
    assert open(prefix + ".data-00000-of-00001", "rb").read() == raw
    header = bytes([0x08, 0x01,                    # BundleHeaderProto.num_shards = 1
                    0x1A, 0x02, 0x08, 0x01])       # .version { producer = 1 }   (endianness LITTLE = 0 is omitted)
    entry = bytes([0x08, 0x01,                     # BundleEntryProto.dtype = DT_FLOAT
                   0x12, 0x04, 0x12, 0x02, 0x08, 0x01,   # .shape { dim { size: 1 } }
                   0x28, 0x04,                     # .size = 4   (shard_id 0 and offset 0 are omitted)
                   0x35]) + struct.pack("<I", ck.mask_crc(ck.crc32c(raw)))   # .crc32c (fixed32, masked)
    data = (bytes([0, 0, len(header)]) + header              # shared, non_shared, value_len, key "", value
            + bytes([0, 1, len(entry)]) + b"a" + entry       # key "a"
            + struct.pack("<II", 0, 1))                       # restart[0] = 0, one restart
    meta = struct.pack("<II", 0, 1)                           # empty metaindex block
    off_meta = len(data) + 5
    off_index = off_meta + len(meta) + 5
    index = (bytes([0, 1, 2]) + b"a" + bytes([0, len(data)])  # key = last key of the block, value = handle(0, len)
             + struct.pack("<II", 0, 1))
    footer = bytes([off_meta, len(meta), off_index, len(index)])
    footer += bytes(40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    want = _trailer(data) + _trailer(meta) + _trailer(index) + footer
    assert open(prefix + ".index", "rb").read() == want
    rd = ck.BundleReader(prefix)
    assert rd.keys() == ["a"] and rd.shape("a") == (1,) and rd.tensor("a")[0] == 1.5


def test_round_trip_many_tensors_prefix_compression_and_blocks(tmp_path, monkeypatch):
    rs = np.random.RandomState(1)
    tensors = {}
    for net in ("G1", "G2", "D", "D_patch2"):
        for i in range(9):
            tensors[f"{net}/layer_{i}/conv2d/w"] = rs.standard_normal((3, 3, 4, 5)).astype(np.float32)
            tensors[f"{net}/layer_{i}/conv2d/w/RMSProp"] = rs.uniform(size=(3, 3, 4, 5)).astype(np.float32)
            tensors[f"{net}/layer_{i}/conv2d/b"] = rs.standard_normal(5).astype(np.float32)
    tensors["D2/Conv/prelu/param"] = np.float32(0.2)                       # scalar
    tensors["empty"] = np.zeros((0, 3), np.float32)                        # zero-size dimension
    tensors["step"] = np.array(12345678901, np.int64)
    tensors["mask"] = np.array([True, False, True])
    tensors["wide"] = rs.standard_normal((70, 1000)).astype(np.float64)
    for block_size in (ck.BLOCK_SIZE, 300):                                # 300 B -> dozens of data blocks
        monkeypatch.setattr(ck, "BLOCK_SIZE", block_size)
        prefix = str(tmp_path / f"ckpt{block_size}" / "EdgeGAN-Model-502")
        ck.write_bundle(prefix, tensors)
        rd = ck.BundleReader(prefix)
        assert sorted(rd.keys()) == sorted(tensors)
        assert rd.keys() == sorted(tensors, key=lambda s: s.encode())      # table order = bytewise key order
        for k, a in tensors.items():
            got = rd.tensor(k)
            assert got.dtype == np.asarray(a).dtype and got.shape == np.asarray(a).shape, k
            assert np.array_equal(got, a), k
        rd.close()


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c-1")
    ck.write_bundle(prefix, {"w": np.arange(100, dtype=np.float32), "b": np.ones(3, np.float32)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[17] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    rd = ck.BundleReader(prefix)
    with pytest.raises(ValueError, match="checksum"):
        rd.tensor("w") if rd.entries["w"]["offset"] <= 17 < rd.entries["w"]["offset"] + 400 else rd.tensor("b")
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[5] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="checksum"):
        ck.BundleReader(prefix)
    open(prefix + ".index", "wb").write(b"not a table at all, but long enough to hold a footer ........")
    with pytest.raises(ValueError, match="magic"):
        ck.BundleReader(prefix)


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert ck.get_checkpoint_state(d) is None
    for step in (2, 502, 1002):
        p = os.path.join(d, f"EdgeGAN-Model-{step}")
        ck.write_bundle(p, {"x": np.float32(step)})
        ck.update_checkpoint_state(d, p, max_to_keep=2)
    st = ck.get_checkpoint_state(d)
    assert st["model_checkpoint_path"] == "EdgeGAN-Model-1002"
    assert st["all_model_checkpoint_paths"] == ["EdgeGAN-Model-502", "EdgeGAN-Model-1002"]
    assert not os.path.exists(os.path.join(d, "EdgeGAN-Model-2.index"))          # max_to_keep
    assert open(os.path.join(d, "checkpoint")).readline() == 'model_checkpoint_path: "EdgeGAN-Model-1002"\n'
    assert ck.step_of("EdgeGAN-Model-1002") == 1002 and ck.step_of("v2-model-17") == 17


@pytest.mark.parametrize("multiclass", [False, True])
def test_model_save_load_round_trip(tmp_path, multiclass):
    """EdgeGAN.save / load (edgegan.py:635-657) on the CPU reference operator set: names follow the TF graph."""
    import torch
    from ref_ops import RefOps
    from test_host_step_cpu import small_cfg
    from edgegan_b200.models.edgegan import EdgeGAN
    _, flags = small_cfg(multiclass)
    ops = RefOps(torch.float32)
    m = EdgeGAN(None, flags, None, ops=ops, seed=5)
    m.build_train_model()
    # make the RMSProp slots distinguishable from their initial value
    for st in m.stores.values():
        st.load({n: np.random.RandomState(len(n)).uniform(0.5, 2.0, s.shape).astype(np.float32)
                 for n, s in zip(st.offsets, st.specs)}, what="ms")
    want_v, want_ms = m.export_variables("var"), m.export_variables("ms")
    prefix = m.save(None, str(tmp_path / "checkpoints"), 502)
    assert os.path.basename(prefix) == "EdgeGAN-Model-502"
    rd = ck.BundleReader(prefix)
    names = set(rd.keys())
    rd.close()
    for n in ("G1/g_lin_0/Matrix", "G1/g_lin_0/Matrix/RMSProp", "G1/g_lin_0/Matrix/RMSProp_1",
              "G2/batch_norm/BatchNorm/moving_variance", "D_patch3/d_conv_4/conv2d/w", "E/FC8_sigma/b/RMSProp"):
        assert n in names, n
    if multiclass:
        assert "D2/Conv/D2/Conv/u" in names and "D2/Conv/u" not in names            # doubled scope path of the reference
        assert "D2/fully_connected/weights/RMSProp" in names
        assert "D2/Conv_1/weights" in names and "D2/Conv_1/weights/RMSProp" not in names   # unused head: no slots

    m2 = EdgeGAN(None, flags, None, ops=ops, seed=6)                                    # different initial weights
    m2.build_train_model()
    found, step = m2.load(None, str(tmp_path / "checkpoints"))
    assert found and step == 502
    got_v, got_ms = m2.export_variables("var"), m2.export_variables("ms")
    for k in want_v:
        assert np.array_equal(got_v[k], want_v[k]), k
    for k in want_ms:
        if not (multiclass and (k.endswith("/u") or k.startswith("D2/Conv_1/"))):
            assert np.array_equal(got_ms[k], want_ms[k]), k
    # a test-time model (E, G1, G2 only) restores from the training checkpoint
    m3 = EdgeGAN(None, flags, None, ops=ops, seed=7)
    m3.build_test_model()
    assert m3.load(None, str(tmp_path / "checkpoints")) == (True, 502)
    assert np.array_equal(m3.export_variables("var")["E/FC8_mu/w"], want_v["E/FC8_mu/w"])
    assert m3.load(None, str(tmp_path / "nothing_here")) == (False, 0)
