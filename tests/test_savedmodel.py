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


FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_bundle_fixture")


def test_reader_against_independently_built_bundle():
    """tests/golden/tf_bundle_fixture was assembled by tests/golden/make_tf_bundle_fixture.py WITHOUT this package's
    writer: protos encoded by the google.protobuf runtime, its own multi-block table builder (prefix compression, restart
    interval 16, shortened index separators), a bitwise CRC-32C, two data shards, a nested Keras object graph."""
    import json
    meta = json.load(open(os.path.join(FIX, "expected.json")))
    assert meta["data_blocks"] >= 4                                    # really a multi-block index
    with np.load(os.path.join(FIX, "expected.npz")) as z:
        want = {k.replace("__", "/"): z[k] for k in z.files}
    prefix = os.path.join(FIX, "plain", "variables")
    table = SM.read_table(prefix + ".index")                          # block checksums verified
    assert len(table) == meta["n_entries"]
    raw = SM.load_checkpoint(prefix, verify_data_crc=True)
    assert np.array_equal(raw["extras/half"], want["extras/half"]) and raw["extras/half"].dtype == np.float16
    assert np.array_equal(raw["extras/bf16"], want["extras/bf16"])
    assert [x.decode() for x in raw["extras/labels"]] == meta["labels"]
    got = SM.load_keras_variables(prefix, verify_data_crc=True)
    assert sorted(got) == meta["keras_names"]                          # optimizer slots / iter skipped, Keras names kept
    for k in got:
        assert got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), k
    assert got["normalization/count"].dtype == np.int64 and int(got["normalization/count"]) == 12345
    # a few-shot SavedModel with session-suffixed Dense names: tower by order, trailing Dense(18) -> Dense(3) = the head
    emb, head = SM.split_fewshot_variables(got)
    assert sorted(k for k in emb if k.startswith("dense")) == ["dense/bias", "dense/kernel", "dense_1/bias", "dense_1/kernel",
                                                             "dense_2/bias", "dense_2/kernel"]
    assert np.array_equal(emb["dense/kernel"], want["dense_7/kernel"]) and np.array_equal(emb["dense_2/bias"], want["dense_9/bias"])
    assert head is not None and head["w1"].shape == (4, 18) and np.array_equal(head["w2"], want["dense_11/kernel"])
    with pytest.raises(ValueError, match="partitioned/embeddings"):
        SM.load_checkpoint(os.path.join(FIX, "sliced", "variables"))


def test_export_in_keras_object_graph_layout(tmp_path):
    """model.save (reference run.py:300): checkpoint keys follow Keras' object graph, few-shot model = Sequential[embedding,
    Dense, Dense]; read back by the loader, head recovered, tower names by order."""
    full = W.random_init(0, randomize_bn=True, dense_units=(32, 32, 16))
    w = {k: v for k, v in full.items() if k.startswith(("normalization", "stem", "block1a", "top_", "dense"))}
    head = dict(w1=np.full((16, 18), 0.5, np.float32), b1=np.arange(18, dtype=np.float32),
                w2=np.full((18, 3), -1.0, np.float32), b2=np.zeros(3, np.float32))
    SM.save_keras_model(tmp_path / "m", w, head)
    keys = sorted(k.decode() for k in SM.read_table(tmp_path / "m" / "variables" / "variables.index"))
    assert "layer_with_weights-0/layer_with_weights-0/mean/.ATTRIBUTES/VARIABLE_VALUE" in keys           # Normalization first
    assert "layer_with_weights-0/layer_with_weights-1/kernel/.ATTRIBUTES/VARIABLE_VALUE" in keys        # stem_conv
    assert "layer_with_weights-1/kernel/.ATTRIBUTES/VARIABLE_VALUE" in keys and "layer_with_weights-2/bias/.ATTRIBUTES/VARIABLE_VALUE" in keys
    got = SM.load_keras_variables(tmp_path / "m", verify_data_crc=True)
    assert "dense_3/kernel" in got and "dense_4/bias" in got                                            # Keras' names of the new layers
    emb, hp = SM.split_fewshot_variables(got)
    assert set(emb) == set(w) and all(np.array_equal(emb[k], w[k]) for k in w)
    assert all(np.array_equal(hp[k], head[k]) for k in head)
    SM.save_keras_model(tmp_path / "e", w)                                                              # embedding alone
    emb2, hp2 = SM.split_fewshot_variables(SM.load_keras_variables(tmp_path / "e"))
    assert hp2 is None and set(emb2) == set(w)
