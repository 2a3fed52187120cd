"""Builds tests/golden/tf_bundle_fixture/: a TensorFlow tensor bundle (variables.index + variables.data-00000-of-00002 /
-00001-of-00002) assembled WITHOUT multilingual_kws_b200/savedmodel.py, to test that module's reader by an independent
route (no TensorFlow offline):

  * BundleHeaderProto / BundleEntryProto / TensorShapeProto / TensorSliceProto / TrackableObjectGraph are encoded by the
    google.protobuf runtime from descriptors written out below from the published .proto definitions
    (tensorflow/core/protobuf/tensor_bundle.proto, framework/tensor_shape.proto, framework/tensor_slice.proto,
    protobuf/trackable_object_graph.proto) — not by the module's hand-rolled wire-format helpers;
  * the table (LevelDB format, tensorflow/core/lib/io/table_format.txt) is built here: data blocks cut by SIZE (several
    blocks), restart interval 16, keys prefix-compressed against the previous key, an index block whose separators are
    SHORTENED keys (>= last key of the block, < first key of the next), an empty metaindex block, 48-byte footer;
  * CRC-32C is a bitwise implementation (the module's is table-driven);
  * two data shards, a float32 / int64 scalar / float16 / bfloat16 / string tensor, a nested object graph (Sequential ->
    functional model -> layers -> variables) with session-suffixed Dense names, optimizer slot variables that a loader
    must skip, and one partitioned (sliced) variable that the reader must reject by name.
Run from the repo root: python tests/golden/make_tf_bundle_fixture.py
"""
import json
import os
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "tf_bundle_fixture")
T = descriptor_pb2.FieldDescriptorProto


def _msg(file, name, fields, nested=()):
    m = file.message_type.add() if not isinstance(file, descriptor_pb2.DescriptorProto) else file.nested_type.add()
    m.name = name
    for fname, num, ftype, label, type_name in fields:
        f = m.field.add()
        f.name, f.number, f.type, f.label = fname, num, ftype, label
        if type_name:
            f.type_name = type_name
    return m


def build_messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "kws_fixture.proto", "kwsfix", "proto3"
    OPT, REP = T.LABEL_OPTIONAL, T.LABEL_REPEATED
    shape = _msg(fd, "TensorShapeProto", [("dim", 2, T.TYPE_MESSAGE, REP, ".kwsfix.TensorShapeProto.Dim"),
                                          ("unknown_rank", 3, T.TYPE_BOOL, OPT, None)])
    _msg(shape, "Dim", [("size", 1, T.TYPE_INT64, OPT, None), ("name", 2, T.TYPE_STRING, OPT, None)])
    sl = _msg(fd, "TensorSliceProto", [("extent", 1, T.TYPE_MESSAGE, REP, ".kwsfix.TensorSliceProto.Extent")])
    _msg(sl, "Extent", [("start", 1, T.TYPE_INT64, OPT, None), ("length", 2, T.TYPE_INT64, OPT, None)])
    _msg(fd, "VersionDef", [("producer", 1, T.TYPE_INT32, OPT, None)])
    _msg(fd, "BundleHeaderProto", [("num_shards", 1, T.TYPE_INT32, OPT, None), ("endianness", 2, T.TYPE_INT32, OPT, None),
                                   ("version", 3, T.TYPE_MESSAGE, OPT, ".kwsfix.VersionDef")])
    _msg(fd, "BundleEntryProto", [("dtype", 1, T.TYPE_INT32, OPT, None), ("shape", 2, T.TYPE_MESSAGE, OPT, ".kwsfix.TensorShapeProto"),
                                  ("shard_id", 3, T.TYPE_INT32, OPT, None), ("offset", 4, T.TYPE_INT64, OPT, None),
                                  ("size", 5, T.TYPE_INT64, OPT, None), ("crc32c", 6, T.TYPE_FIXED32, OPT, None),
                                  ("slices", 7, T.TYPE_MESSAGE, REP, ".kwsfix.TensorSliceProto")])
    og = _msg(fd, "TrackableObjectGraph", [("nodes", 1, T.TYPE_MESSAGE, REP, ".kwsfix.TrackableObjectGraph.TrackableObject")])
    to = _msg(og, "TrackableObject", [
        ("children", 1, T.TYPE_MESSAGE, REP, ".kwsfix.TrackableObjectGraph.TrackableObject.ObjectReference"),
        ("attributes", 2, T.TYPE_MESSAGE, REP, ".kwsfix.TrackableObjectGraph.TrackableObject.SerializedTensor"),
        ("slot_variables", 3, T.TYPE_MESSAGE, REP, ".kwsfix.TrackableObjectGraph.TrackableObject.SlotVariableReference")])
    _msg(to, "ObjectReference", [("node_id", 1, T.TYPE_INT32, OPT, None), ("local_name", 2, T.TYPE_STRING, OPT, None)])
    _msg(to, "SerializedTensor", [("name", 1, T.TYPE_STRING, OPT, None), ("full_name", 2, T.TYPE_STRING, OPT, None),
                                  ("checkpoint_key", 3, T.TYPE_STRING, OPT, None)])
    _msg(to, "SlotVariableReference", [("original_variable_node_id", 1, T.TYPE_INT32, OPT, None),
                                       ("slot_name", 2, T.TYPE_STRING, OPT, None), ("slot_variable_node_id", 3, T.TYPE_INT32, OPT, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("kwsfix." + n))   # noqa: E731
    return {n: get(n) for n in ("BundleHeaderProto", "BundleEntryProto", "TrackableObjectGraph")}


def crc32c_bitwise(data: bytes) -> int:
    crc = 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def masked(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


class TableBuilder:
    """LevelDB-format table: blocks of <= block_size bytes, restart interval 16, shortened index separators."""

    def __init__(self, block_size=260, restart_interval=16):
        self.block_size, self.ri = block_size, restart_interval
        self.file = bytearray()
        self.index = []          # (separator key, handle bytes)
        self._reset()
        self.pending = None      # (last key of the finished block, its handle)

    def _reset(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        if self.pending is not None:                       # separator between the previous block and this key
            last, handle = self.pending
            sep = self._shortest_separator(last, key)
            self.index.append((sep, handle))
            self.pending = None
        shared = 0
        if self.count % self.ri == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count % self.ri:
            while shared < min(len(self.last), len(key)) and self.last[shared] == key[shared]:
                shared += 1
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1
        if len(self.buf) >= self.block_size:
            self._flush()

    @staticmethod
    def _shortest_separator(a: bytes, b: bytes) -> bytes:
        n = 0
        while n < min(len(a), len(b)) and a[n] == b[n]:
            n += 1
        if n < len(a) and n < len(b) and a[n] + 1 < b[n]:
            return a[:n] + bytes([a[n] + 1])
        return a

    def _emit(self, block: bytes) -> bytes:
        off = len(self.file)
        self.file += block + b"\x00" + struct.pack("<I", masked(crc32c_bitwise(block + b"\x00")))
        return varint(off) + varint(len(block))

    def _finish_block(self, buf, restarts) -> bytes:
        return bytes(buf) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))

    def _flush(self):
        if not self.count:
            return
        handle = self._emit(self._finish_block(self.buf, self.restarts))
        self.pending = (self.last, handle)
        self._reset()

    def finish(self) -> bytes:
        self._flush()
        if self.pending is not None:
            last, handle = self.pending
            self.index.append((last + b"\x00", handle))    # a key >= every key of the last block
        meta = self._emit(self._finish_block(bytearray(), [0]))
        ib, restarts = bytearray(), []
        for k, h in self.index:                            # index block: restart interval 1
            restarts.append(len(ib))
            ib += varint(0) + varint(len(k)) + varint(len(h)) + k + h
        idx = self._emit(self._finish_block(ib, restarts))
        footer = meta + idx
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
        return bytes(self.file) + footer


def main():
    M = build_messages()
    rng = np.random.default_rng(20261017)
    os.makedirs(OUT, exist_ok=True)
    DT = {"float32": 1, "int64": 9, "float16": 19, "bfloat16": 14, "string": 7}
    # Keras-style few-shot model: Sequential[functional embedding (3 layers here), Dense, Dense], session-suffixed names
    layers = [("stem_conv", {"kernel": rng.normal(size=(3, 3, 1, 4)).astype(np.float32)}),
              ("stem_bn", {"gamma": rng.normal(size=4).astype(np.float32), "beta": rng.normal(size=4).astype(np.float32),
                           "moving_mean": rng.normal(size=4).astype(np.float32), "moving_variance": rng.uniform(0.5, 2, 4).astype(np.float32)}),
              ("normalization", {"count": np.array(12345, dtype=np.int64)}),
              ("dense_7", {"kernel": rng.normal(size=(4, 6)).astype(np.float32), "bias": rng.normal(size=6).astype(np.float32)}),
              ("dense_8", {"kernel": rng.normal(size=(6, 5)).astype(np.float32), "bias": rng.normal(size=5).astype(np.float32)}),
              ("dense_9", {"kernel": rng.normal(size=(5, 4)).astype(np.float32), "bias": rng.normal(size=4).astype(np.float32)})]
    head = [("dense_10", {"kernel": rng.normal(size=(4, 18)).astype(np.float32), "bias": rng.normal(size=18).astype(np.float32)}),
            ("dense_11", {"kernel": rng.normal(size=(18, 3)).astype(np.float32), "bias": rng.normal(size=3).astype(np.float32)})]
    half = rng.normal(size=(5,)).astype(np.float16)
    bf_src = rng.normal(size=(3, 2)).astype(np.float32)
    bf_bits = (bf_src.view(np.uint32) >> 16).astype(np.uint16)
    shards = [bytearray(), bytearray()]
    entries = {}
    expect = {}

    def put(key, dtype_id, shape, raw, shard):
        e = M["BundleEntryProto"]()
        e.dtype = dtype_id
        for d in shape:
            e.shape.dim.add().size = int(d)
        e.shard_id, e.offset, e.size, e.crc32c = shard, len(shards[shard]), len(raw), masked(crc32c_bitwise(raw))
        shards[shard] += raw
        entries[key.encode()] = e.SerializeToString()

    graph = M["TrackableObjectGraph"]()
    root = graph.nodes.add()

    def add_layer(parent, local, lname, variables, path, shard):
        node_id = len(graph.nodes)
        node = graph.nodes.add()
        ref = parent.children.add()
        ref.node_id, ref.local_name = node_id, local
        var_ids = {}
        for attr, arr in variables.items():
            vid = len(graph.nodes)
            vnode = graph.nodes.add()
            r = graph.nodes[node_id].children.add()
            r.node_id, r.local_name = vid, attr
            key = f"{path}/{attr}/.ATTRIBUTES/VARIABLE_VALUE"
            a = vnode.attributes.add()
            a.name, a.full_name, a.checkpoint_key = "VARIABLE_VALUE", f"{lname}/{attr}", key
            put(key, DT[str(arr.dtype)], arr.shape, arr.astype(arr.dtype.newbyteorder("<")).tobytes(), shard)
            expect[f"{lname}/{attr}"] = arr
            var_ids[attr] = vid
        return node_id, var_ids

    emb_id = len(graph.nodes)
    emb = graph.nodes.add()
    r = root.children.add()
    r.node_id, r.local_name = emb_id, "layer_with_weights-0"
    first_kernel_node = None
    for i, (lname, variables) in enumerate(layers):
        _, ids = add_layer(graph.nodes[emb_id], f"layer_with_weights-{i}", lname, variables, f"layer_with_weights-0/layer_with_weights-{i}", i % 2)
        if first_kernel_node is None and "kernel" in ids:
            first_kernel_node = ids["kernel"]
    for j, (lname, variables) in enumerate(head):
        add_layer(root, f"layer_with_weights-{j + 1}", lname, variables, f"layer_with_weights-{j + 1}", 1)
    # optimizer with one slot variable and its hyper-parameters: a weight loader has to skip all of these
    opt_id = len(graph.nodes)
    opt = graph.nodes.add()
    r = root.children.add()
    r.node_id, r.local_name = opt_id, "optimizer"
    slot_id = len(graph.nodes)
    slot = graph.nodes.add()
    skey = "layer_with_weights-0/layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"
    a = slot.attributes.add()
    a.name, a.full_name, a.checkpoint_key = "VARIABLE_VALUE", "Adam/stem_conv/kernel/m", skey
    put(skey, DT["float32"], (3, 3, 1, 4), np.zeros((3, 3, 1, 4), np.float32).tobytes(), 0)
    sv = graph.nodes[opt_id].slot_variables.add()
    sv.original_variable_node_id, sv.slot_name, sv.slot_variable_node_id = first_kernel_node, "m", slot_id
    it_id = len(graph.nodes)
    it = graph.nodes.add()
    r = graph.nodes[opt_id].children.add()
    r.node_id, r.local_name = it_id, "iter"
    a = it.attributes.add()
    a.name, a.full_name, a.checkpoint_key = "VARIABLE_VALUE", "Adam/iter", "optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"
    put("optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE", DT["int64"], (), np.array(256, np.int64).tobytes(), 1)
    # raw (name-based) extras: float16, bfloat16 and a string tensor
    put("extras/half", DT["float16"], half.shape, half.tobytes(), 0)
    put("extras/bf16", DT["bfloat16"], bf_bits.shape, bf_bits.astype("<u2").tobytes(), 1)
    strings = [b"silence", b"", b"unknown word"]
    lens = b"".join(varint(len(x)) for x in strings)
    put("extras/labels", DT["string"], (3,), lens + struct.pack("<I", masked(crc32c_bitwise(lens))) + b"".join(strings), 0)
    # the object graph itself (a scalar string tensor)
    g = graph.SerializeToString()
    lens = varint(len(g))
    put("_CHECKPOINTABLE_OBJECT_GRAPH", DT["string"], (), lens + struct.pack("<I", masked(crc32c_bitwise(lens))) + g, 0)
    tb_entries = dict(entries)
    hdr = M["BundleHeaderProto"]()
    hdr.num_shards, hdr.endianness = 2, 0
    hdr.version.producer = 1
    tb_entries[b""] = hdr.SerializeToString()

    def write(dirname, extra=None):
        d = os.path.join(OUT, dirname)
        os.makedirs(d, exist_ok=True)
        items = dict(tb_entries)
        if extra:
            items.update(extra)
        tb = TableBuilder()
        for k in sorted(items):
            tb.add(k, items[k])
        with open(os.path.join(d, "variables.index"), "wb") as f:
            f.write(tb.finish())
        for i, sh in enumerate(shards):
            with open(os.path.join(d, f"variables.data-{i:05d}-of-00002"), "wb") as f:
                f.write(bytes(sh))
        return len(tb.index)

    n_blocks = write("plain")
    # the same bundle + one partitioned variable (BundleEntryProto.slices set): unsupported, must be reported by name
    e = M["BundleEntryProto"]()
    e.dtype = DT["float32"]
    e.shape.dim.add().size = 8
    s1 = e.slices.add()
    ex = s1.extent.add()
    ex.start, ex.length = 0, 4
    write("sliced", {b"partitioned/embeddings": e.SerializeToString()})
    np.savez(os.path.join(OUT, "expected.npz"), **{k.replace("/", "__"): v for k, v in expect.items()},
             extras__half=half, extras__bf16=(bf_bits.astype(np.uint32) << 16).view(np.float32))
    with open(os.path.join(OUT, "expected.json"), "w") as f:
        json.dump({"data_blocks": n_blocks, "labels": [x.decode() for x in strings],
                   "keras_names": sorted(expect), "n_entries": len(tb_entries)}, f, indent=1)
    print("wrote", OUT, "data blocks:", n_blocks, "entries:", len(tb_entries))


if __name__ == "__main__":
    main()
