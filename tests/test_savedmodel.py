"""TensorFlow-free checkpoint reader (multilingual_kws_b200/savedmodel.py).  No TF-written file is
available offline, so these tests pin the pieces that have independent ground truth (CRC-32C test vector, protobuf wire
format, table magic / footer layout) and the reader against the writer that follows the same format description."""
import os
import struct

import numpy as np
import pytest

from multilingual_kws_b200 import savedmodel as SM
from multilingual_kws_b200 import weights as W


def test_crc32c_known_vectors():
    assert SM.crc32c(b"123456789") == 0xE3069283                   # the standard CRC-32C check value
    assert SM.crc32c(b"") == 0
    assert SM.crc32c(bytes(32)) == 0x8A9136AA                       # RFC 3720 B.4: 32 bytes of zeros
    assert SM.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43              # RFC 3720 B.4: 32 bytes of ones
    assert SM.crc32c(b"6789", SM.crc32c(b"12345")) == 0xE3069283   # incremental
    m = SM.mask_crc(0xE3069283)
    assert m != 0xE3069283 and ((m - 0xA282EAD8) & 0xFFFFFFFF) == (((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF)


def test_protobuf_wire_format():
    # field 1 varint 300, field 2 bytes "hi", field 6 fixed32 0xdeadbeef, field 4 varint -1 (int64 two's complement)
    msg = bytes([0x08, 0xAC, 0x02, 0x12, 0x02]) + b"hi" + bytes([0x35]) + struct.pack("<I", 0xDEADBEEF) + \
        bytes([0x20]) + bytes([0xFF] * 9 + [0x01])
    got = list(SM.proto_fields(msg))
    assert got[0] == (1, 0, 300) and got[1] == (2, 2, b"hi")
    assert got[2][:2] == (6, 5) and struct.unpack("<I", got[2][2])[0] == 0xDEADBEEF
    assert got[3] == (4, 0, (1 << 64) - 1)
    with pytest.raises(ValueError):
        list(SM.proto_fields(bytes([0x12, 0x05]) + b"abc"))


def test_table_round_trip_multi_block(tmp_path):
    rng = np.random.default_rng(0)
    items = {b"": b"header"}
    for i in range(300):
        items[f"layer_with_weights-{i}/kernel/.ATTRIBUTES/VARIABLE_VALUE".encode()] = rng.bytes(int(rng.integers(0, 60)))
    path = tmp_path / "t.index"
    SM.write_table(path, items, block_entries=7)
    assert SM.read_table(path) == items
    raw = path.read_bytes()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 48
    bad = bytearray(raw)
    bad[10] ^= 0x40
    (tmp_path / "bad.index").write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        SM.read_table(tmp_path / "bad.index")
    (tmp_path / "junk.index").write_bytes(b"x" * 100)
    with pytest.raises(ValueError, match="magic"):
        SM.read_table(tmp_path / "junk.index")


def test_keras_checkpoint_round_trip(tmp_path):
    full = W.random_init(0, randomize_bn=True, dense_units=(32, 32, 16))
    w = {k: v for k, v in full.items() if k.startswith(("normalization", "stem", "block1a", "block2a", "block7a", "top_", "dense"))}
    SM.write_keras_checkpoint(tmp_path / "model", w)
    assert os.path.isfile(tmp_path / "model" / "variables" / "variables.index")
    got = SM.load_keras_variables(tmp_path / "model", verify_data_crc=False)
    assert set(got) == set(w)
    for k in w:
        assert got[k].dtype == w[k].dtype and got[k].shape == w[k].shape and np.array_equal(got[k], w[k]), k
    from multilingual_kws_b200.model import cut_at
    assert "dense_2/kernel" in cut_at(got, "dense_2") and "dense_2/kernel" not in cut_at(got, "dense_1")
    small = {"a/kernel": np.arange(6, dtype=np.float32).reshape(2, 3), "a/count": np.array(7, dtype=np.int64)}
    SM.write_keras_checkpoint(tmp_path / "small", small)
    got = SM.load_keras_variables(tmp_path / "small", verify_data_crc=True)
    assert got["a/count"].shape == () and int(got["a/count"]) == 7 and np.array_equal(got["a/kernel"], small["a/kernel"])
    data = tmp_path / "small" / "variables" / "variables.data-00000-of-00001"
    raw = bytearray(data.read_bytes())
    raw[3] ^= 1
    data.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        SM.load_keras_variables(tmp_path / "small", verify_data_crc=True)
