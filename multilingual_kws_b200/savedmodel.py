"""TensorFlow-free reader (and minimal writer) of the variables of a Keras SavedModel / TF2 checkpoint
(SURVEY.md §8 f2): ``<dir>/variables/variables.index`` + ``variables.data-00000-of-00001``.

The reference loads its base model with ``tf.keras.models.load_model(base_model_path)``
(multilingual_kws/embedding/transfer_learning.py:36, batch_streaming_analysis.py:202); only the variable VALUES are
needed here, so this module reads the checkpoint ("tensor bundle") inside the SavedModel directory:

* ``variables.index`` is an immutable sorted string table in TensorFlow's ``table`` format (the LevelDB table format):
  prefix-compressed key/value blocks with restart arrays, each block followed by a 1-byte compression tag and a masked
  CRC-32C, an index block of block handles, and a 48-byte footer ending in the magic number ``0xdb4775248b80fb57``.
  Key ``""`` maps to a ``BundleHeaderProto``, every other key (a checkpoint key) to a ``BundleEntryProto``
  (dtype, shape, shard id, offset, size, crc32c).
* ``variables.data-XXXXX-of-YYYYY`` holds the raw little-endian tensor bytes at those offsets.
* key ``_CHECKPOINTABLE_OBJECT_GRAPH`` holds a serialized ``TrackableObjectGraph``; its ``SerializedTensor`` records give,
  for every variable, the Keras variable name (``full_name``, e.g. ``block1a_dwconv/depthwise_kernel``) and the checkpoint
  key (``layer_with_weights-3/depthwise_kernel/.ATTRIBUTES/VARIABLE_VALUE``) — which is how ``load_keras_variables``
  returns a Keras-named dict without having to guess the layer order.

EXPORT (`save_keras_model`, used by ``FewShotModel.save`` / ``EmbeddingModel.save`` for the reference's
``model.save(output)``, run.py:300): writes ``<dir>/variables/variables.{index,data-00000-of-00001}`` as a TF2
object-based checkpoint whose object graph has the shape Keras gives these models — root -> ``layer_with_weights-i`` ->
``kernel`` / ``bias`` / ``gamma`` / ... -> ``.ATTRIBUTES/VARIABLE_VALUE``, the few-shot model as
Sequential[Functional embedding, Dense(18), Dense(3)] with the embedding nested under ``layer_with_weights-0`` — i.e. what
``tf.keras.Model.load_weights(<dir>/variables/variables)`` / ``tf.train.Checkpoint.restore`` match against a freshly built
Keras model of the same architecture.  ``saved_model.pb`` (the traced TensorFlow graph + Keras metadata) cannot be produced
without TensorFlow and is NOT written; ``weights.npz`` is written next to it for this package's own loader.

STATUS: written from the published format descriptions (tensorflow/core/lib/io/table_format.txt, tensor_bundle.proto,
trackable_object_graph.proto).  No TensorFlow-written file exists in this environment.  The reader is tested against (a)
files produced by the writer below and (b) a fixture assembled by an independent route — tests/golden/make_tf_bundle_fixture.py
encodes the protos with the google.protobuf runtime from descriptors written out from the .proto definitions, builds the
table with its own block builder (several data blocks, restart interval 16, prefix-compressed keys, nested object graph,
a string tensor, a partitioned variable) and its own bitwise CRC-32C — committed under tests/golden/tf_bundle_fixture/.
It is still NOT validated against a file TensorFlow wrote.  Snappy-compressed index blocks (TensorFlow writes the bundle
index uncompressed) are rejected with a clear error.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_STRING, _DT_BFLOAT16 = 7, 14
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------ CRC-32C (Castagnoli), masked as TensorFlow stores it
def _make_crc_table() -> List[int]:
    table = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    t = _CRC_TABLE
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / protobuf wire format
def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
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


def _write_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yields (field number, wire type, value) of one protobuf message: varint -> int, 64-bit / 32-bit -> bytes of that
    width, length-delimited -> bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
            if len(v) != n:
                raise ValueError("truncated length-delimited field")
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def _pb(field: int, wt: int, payload: bytes) -> bytes:
    return _write_varint((field << 3) | wt) + payload


def _pb_bytes(field: int, data: bytes) -> bytes:
    return _pb(field, 2, _write_varint(len(data)) + data)


def _pb_varint(field: int, v: int) -> bytes:
    return _pb(field, 0, _write_varint(v))


# ------------------------------------------------------------------ table (SSTable) reader
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    block = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(block) != size or len(trailer) != 5:
        raise ValueError("table block outside the file")
    if trailer[0] == 1:
        raise ValueError("snappy-compressed table block: not supported (TensorFlow writes the bundle index uncompressed)")
    if trailer[0] != 0:
        raise ValueError(f"unknown table block compression {trailer[0]}")
    if verify:
        want = struct.unpack("<I", trailer[1:5])[0]
        if mask_crc(crc32c(block + trailer[:1])) != want:
            raise ValueError("table block checksum mismatch")
    return block


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("table block too small")
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise ValueError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        unshared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        if shared > len(key):
            raise ValueError("bad prefix compression")
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        value = block[pos:pos + vlen]
        pos += vlen
        yield key, value


def _block_handle(buf: bytes, pos: int = 0) -> Tuple[int, int, int]:
    off, pos = _read_varint(buf, pos)
    size, pos = _read_varint(buf, pos)
    return off, size, pos


def read_table(path: os.PathLike, verify: bool = True) -> Dict[bytes, bytes]:
    """All key/value pairs of a TensorFlow / LevelDB table file."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow table file (bad magic)")
    footer = data[-48:]
    _, _, pos = _block_handle(footer, 0)                     # metaindex handle (unused)
    idx_off, idx_size, _ = _block_handle(footer, pos)
    out: Dict[bytes, bytes] = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, size, _ = _block_handle(handle)
        for k, v in _block_entries(_read_block(data, off, size, verify)):
            out[k] = v
    return out


# ------------------------------------------------------------------ tensor bundle
def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for field, _, v in proto_fields(buf):
        if field == 2:                                       # repeated Dim dim
            size = 0
            for f2, _, v2 in proto_fields(v):
                if f2 == 1:
                    size = v2 - (1 << 64) if v2 >> 63 else v2
            dims.append(size)
        elif field == 3 and v:
            raise ValueError("tensor of unknown rank in a checkpoint")
    return tuple(dims)


def parse_bundle_entry(buf: bytes) -> dict:
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for field, wt, v in proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            e["shape"] = _parse_shape(v)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = struct.unpack("<I", v)[0] if wt == 5 else v
        elif field == 7:
            e["sliced"] = True
    return e


def _decode_tensor(raw: bytes, entry: dict) -> np.ndarray:
    dt = entry["dtype"]
    shape = entry["shape"]
    if dt == _DT_STRING:
        n = int(np.prod(shape)) if shape else 1
        pos, lengths = 0, []
        for _ in range(n):
            ln, pos = _read_varint(raw, pos)
            lengths.append(ln)
        pos += 4                                             # masked crc32c of the lengths
        items = []
        for ln in lengths:
            items.append(raw[pos:pos + ln])
            pos += ln
        arr = np.empty(n, dtype=object)
        arr[:] = items
        return arr.reshape(shape)
    if dt == _DT_BFLOAT16:
        u16 = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
        return u16.view(np.float32).reshape(shape)
    if dt not in _DTYPES:
        raise ValueError(f"unsupported checkpoint dtype {dt}")
    return np.frombuffer(raw, dtype=np.dtype(_DTYPES[dt]).newbyteorder("<")).reshape(shape).copy()


def load_checkpoint(prefix: os.PathLike, verify_data_crc: bool = False) -> Dict[str, np.ndarray]:
    """All tensors of the bundle ``<prefix>.index`` / ``<prefix>.data-*`` keyed by checkpoint key."""
    prefix = os.fspath(prefix)
    table = read_table(prefix + ".index")
    header = table.get(b"")
    num_shards = 1
    if header is not None:
        for field, _, v in proto_fields(header):
            if field == 1:
                num_shards = v
            elif field == 2 and v != 0:
                raise ValueError("big-endian tensor bundle: not supported")
    shards: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for key, value in table.items():
        if key == b"":
            continue
        e = parse_bundle_entry(value)
        if e["sliced"]:
            raise ValueError(f"{key.decode()}: partitioned (sliced) variables are not supported")
        sid = e["shard_id"]
        if sid not in shards:
            with open(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError(f"{key.decode()}: data shard too short")
        if verify_data_crc and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError(f"{key.decode()}: tensor checksum mismatch")
        out[key.decode("utf-8")] = _decode_tensor(raw, e)
    return out


def parse_object_graph(buf: bytes) -> List[dict]:
    """TrackableObjectGraph -> [{children: {local_name: node_id}, attributes: [(name, full_name, checkpoint_key)]}]."""
    nodes = []
    for field, _, v in proto_fields(buf):
        if field != 1:
            continue
        node = dict(children={}, attributes=[])
        for f2, _, v2 in proto_fields(v):
            if f2 == 1:                                      # ObjectReference {node_id = 1, local_name = 2}
                nid, name = 0, ""
                for f3, _, v3 in proto_fields(v2):
                    if f3 == 1:
                        nid = v3
                    elif f3 == 2:
                        name = v3.decode("utf-8")
                node["children"][name] = nid
            elif f2 == 2:                                    # SerializedTensor {name = 1, full_name = 2, checkpoint_key = 3}
                name = full = key = ""
                for f3, _, v3 in proto_fields(v2):
                    if f3 == 1:
                        name = v3.decode("utf-8")
                    elif f3 == 2:
                        full = v3.decode("utf-8")
                    elif f3 == 3:
                        key = v3.decode("utf-8")
                node["attributes"].append((name, full, key))
        nodes.append(node)
    return nodes


def load_keras_variables(model_dir: os.PathLike, verify_data_crc: bool = False) -> Dict[str, np.ndarray]:
    """Keras-named variables (``stem_conv/kernel``, ``block1a_bn/moving_mean``, ``dense_2/bias``, ...) of the SavedModel
    directory (or checkpoint prefix) `model_dir`.  Optimizer slots and bookkeeping tensors are skipped."""
    p = os.fspath(model_dir)
    prefix = os.path.join(p, "variables", "variables") if os.path.isdir(p) else p
    tensors = load_checkpoint(prefix, verify_data_crc)
    if OBJECT_GRAPH_KEY not in tensors:
        raise ValueError(f"{prefix}: no {OBJECT_GRAPH_KEY} entry (not a TF2 object-based checkpoint)")
    graph = parse_object_graph(bytes(tensors[OBJECT_GRAPH_KEY].reshape(-1)[0]))
    out: Dict[str, np.ndarray] = {}
    for node in graph:
        for name, full, key in node["attributes"]:
            if name != "VARIABLE_VALUE" or not full or key not in tensors or "/.OPTIMIZER_SLOT/" in key:
                continue
            full = full[:-2] if full.endswith(":0") else full
            if full.startswith(("Adam/", "training/", "SGD/", "RMSprop/")) or full in ("iter", "beta_1", "beta_2", "decay", "learning_rate"):
                continue
            arr = tensors[key]
            if arr.dtype == object:
                continue
            if full in out and not np.array_equal(out[full], arr):
                raise ValueError(f"two different variables are both named {full}")
            out[full] = arr
    if not out:
        raise ValueError(f"{prefix}: the object graph names no variables")
    return out


# ------------------------------------------------------------------ minimal writer (fixtures for the reader's tests, and
# export of a weight dict in the same container format)
def _table_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out = bytearray()
    restarts = []
    prev = b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _write_varint(shared) + _write_varint(len(k) - shared) + _write_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: os.PathLike, items: Dict[bytes, bytes], block_entries: int = 24) -> None:
    keys = sorted(items)
    data = bytearray()
    index_entries = []

    def emit(block: bytes) -> bytes:
        off = len(data)
        data.extend(block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _write_varint(off) + _write_varint(len(block))

    for i in range(0, len(keys), block_entries):
        chunk = keys[i:i + block_entries]
        handle = emit(_table_block([(k, items[k]) for k in chunk]))
        index_entries.append((chunk[-1], handle))            # separator = last key of the block
    meta = emit(_table_block([]))
    index = emit(_table_block(index_entries, restart_interval=1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(path, "wb") as f:
        f.write(bytes(data) + footer)


def write_keras_checkpoint(model_dir: os.PathLike, variables: Dict[str, np.ndarray]) -> None:
    """Writes `variables` (Keras name -> array) as ``<model_dir>/variables/variables.{index,data-00000-of-00001}`` with an
    object graph that names them, in the layout `load_keras_variables` reads."""
    vdir = os.path.join(os.fspath(model_dir), "variables")
    os.makedirs(vdir, exist_ok=True)
    prefix = os.path.join(vdir, "variables")
    blob = bytearray()
    items: Dict[bytes, bytes] = {b"": _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1))}
    root_children = b""
    nodes = []
    for n, (name, arr) in enumerate(sorted(variables.items())):
        arr = np.asarray(arr, order="C")                  # (ascontiguousarray would turn a 0-d array into shape (1,))
        if arr.dtype not in _DTYPE_IDS:
            raise ValueError(f"{name}: dtype {arr.dtype} cannot be written")
        key = f"layer_with_weights-{n}/v/.ATTRIBUTES/VARIABLE_VALUE"
        raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        shape = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in arr.shape)
        items[key.encode()] = (_pb_varint(1, _DTYPE_IDS[arr.dtype]) + _pb_bytes(2, shape) + _pb_varint(4, len(blob)) +
                               _pb_varint(5, len(raw)) + _pb(6, 5, struct.pack("<I", mask_crc(crc32c(raw)))))
        blob += raw
        root_children += _pb_bytes(1, _pb_varint(1, n + 1) + _pb_bytes(2, f"layer_with_weights-{n}".encode()))
        nodes.append(_pb_bytes(2, _pb_bytes(1, b"VARIABLE_VALUE") + _pb_bytes(2, name.encode()) + _pb_bytes(3, key.encode())))
    graph = _pb_bytes(1, root_children) + b"".join(_pb_bytes(1, nd) for nd in nodes)
    lengths = _write_varint(len(graph))
    raw = lengths + struct.pack("<I", mask_crc(crc32c(lengths))) + graph
    items[OBJECT_GRAPH_KEY.encode()] = (_pb_varint(1, _DT_STRING) + _pb_bytes(2, b"") + _pb_varint(4, len(blob)) +
                                       _pb_varint(5, len(raw)) + _pb(6, 5, struct.pack("<I", mask_crc(crc32c(raw)))))
    blob += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(blob))
    write_table(prefix + ".index", items)


# ------------------------------------------------------------------ export in Keras' own object-graph layout
def _layer_groups(variables: Dict[str, np.ndarray]) -> List[Tuple[str, List[Tuple[str, np.ndarray]]]]:
    """[(layer name, [(attribute, array)])] in model order: the order of `weights.param_shapes()` (= Keras' layer order
    for EfficientNetB0 + tower) for known names, then anything else alphabetically."""
    from . import weights as W
    order = {}
    for k in W.param_shapes(tuple(8 for _ in range(8))):     # names only; 8 dense layers cover any tower length
        order.setdefault(k.rsplit("/", 1)[0], len(order))
    groups: Dict[str, List[Tuple[str, np.ndarray]]] = {}
    for k, v in variables.items():
        if k.startswith(("fewshot_head/", "kws_meta/")):
            continue
        layer, attr = k.rsplit("/", 1)
        groups.setdefault(layer, []).append((attr, np.asarray(v)))
    return sorted(groups.items(), key=lambda kv: (order.get(kv[0], 10 ** 6), kv[0]))


def write_object_checkpoint(prefix: os.PathLike, tree: dict) -> None:
    """TF2 object-based checkpoint from a nested dict: {child local name: subtree | (full_name, array)}.  Checkpoint keys
    are the child paths + "/.ATTRIBUTES/VARIABLE_VALUE", the object graph mirrors the tree."""
    prefix = os.fspath(prefix)
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    blob = bytearray()
    items: Dict[bytes, bytes] = {b"": _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1))}
    nodes: List[bytes] = []

    def visit(sub, path: str) -> int:
        nid = len(nodes)
        nodes.append(b"")
        if isinstance(sub, dict):
            body = b""
            for name, child in sub.items():
                cid = visit(child, f"{path}/{name}" if path else name)
                body += _pb_bytes(1, _pb_varint(1, cid) + _pb_bytes(2, name.encode()))
            nodes[nid] = body
            return nid
        full_name, arr = sub
        arr = np.asarray(arr, order="C")
        if arr.dtype not in _DTYPE_IDS:
            raise ValueError(f"{full_name}: dtype {arr.dtype} cannot be written")
        key = path + "/.ATTRIBUTES/VARIABLE_VALUE"
        raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        shape = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in arr.shape)
        items[key.encode()] = (_pb_varint(1, _DTYPE_IDS[arr.dtype]) + _pb_bytes(2, shape) + _pb_varint(4, len(blob)) +
                               _pb_varint(5, len(raw)) + _pb(6, 5, struct.pack("<I", mask_crc(crc32c(raw)))))
        blob.extend(raw)
        nodes[nid] = _pb_bytes(2, _pb_bytes(1, b"VARIABLE_VALUE") + _pb_bytes(2, full_name.encode()) + _pb_bytes(3, key.encode()))
        return nid

    visit(tree, "")
    graph = b"".join(_pb_bytes(1, nd) for nd in nodes)
    lengths = _write_varint(len(graph))
    raw = lengths + struct.pack("<I", mask_crc(crc32c(lengths))) + graph
    items[OBJECT_GRAPH_KEY.encode()] = (_pb_varint(1, _DT_STRING) + _pb_bytes(2, b"") + _pb_varint(4, len(blob)) +
                                       _pb_varint(5, len(raw)) + _pb(6, 5, struct.pack("<I", mask_crc(crc32c(raw)))))
    blob.extend(raw)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(blob))
    write_table(prefix + ".index", items)


def keras_tree(variables: Dict[str, np.ndarray]) -> dict:
    """Object tree of a Keras functional model: layer_with_weights-i -> attribute -> (full name, array)."""
    tree = {}
    for i, (layer, attrs) in enumerate(_layer_groups(variables)):
        tree[f"layer_with_weights-{i}"] = {attr: (f"{layer}/{attr}", arr) for attr, arr in attrs}
    return tree


def save_keras_model(model_dir: os.PathLike, variables: Dict[str, np.ndarray], head: Optional[Dict[str, np.ndarray]] = None) -> None:
    """``model.save(model_dir)`` without TensorFlow: the variables checkpoint of the embedding (functional model) or, with
    `head` = {w1, b1, w2, b2}, of the few-shot Sequential[embedding, Dense(hidden, tanh), Dense(classes, softmax)] the
    reference builds (transfer_learning.py:47-53), whose new Dense layers Keras names after the tower's (dense_3, dense_4)."""
    emb = keras_tree(variables)
    if head is None:
        tree = emb
    else:
        n_dense = len({k.rsplit("/", 1)[0] for k in variables if k.startswith("dense")})
        names = [f"dense_{n_dense}", f"dense_{n_dense + 1}"]
        tree = {"layer_with_weights-0": emb,
                "layer_with_weights-1": {"kernel": (names[0] + "/kernel", np.asarray(head["w1"], np.float32)),
                                         "bias": (names[0] + "/bias", np.asarray(head["b1"], np.float32))},
                "layer_with_weights-2": {"kernel": (names[1] + "/kernel", np.asarray(head["w2"], np.float32)),
                                         "bias": (names[1] + "/bias", np.asarray(head["b2"], np.float32))}}
    write_object_checkpoint(os.path.join(os.fspath(model_dir), "variables", "variables"), tree)


def _dense_index(layer: str) -> int:
    tail = layer.rsplit("_", 1)[-1]
    return int(tail) if tail.isdigit() else 0


def split_fewshot_variables(variables: Dict[str, np.ndarray]) -> Tuple[Dict[str, np.ndarray], Optional[Dict[str, np.ndarray]]]:
    """Separates a few-shot SavedModel's variables into (embedding variables, head or None).  Dense layers are taken in
    index order whatever their suffix (Keras numbers layers per session: ``dense_3``, ``dense_17`` ...): the tower is
    renamed to ``dense``, ``dense_1``, ...; if the last two Dense layers look like the reference's head (hidden <= 32
    units feeding <= 8 classes, transfer_learning.py:47-53) they become {w1, b1, w2, b2}."""
    dense = sorted({k.rsplit("/", 1)[0] for k in variables if k.split("/")[0].startswith("dense") and k.endswith("/kernel")},
                   key=_dense_index)
    out = {k: v for k, v in variables.items() if not k.split("/")[0].startswith("dense")}
    head = None
    if len(dense) >= 3:
        k1, k2 = variables[dense[-2] + "/kernel"], variables[dense[-1] + "/kernel"]
        if k1.ndim == 2 and k2.ndim == 2 and k1.shape[1] == k2.shape[0] and k1.shape[1] <= 32 and k2.shape[1] <= 8:
            head = dict(w1=k1, b1=variables[dense[-2] + "/bias"], w2=k2, b2=variables[dense[-1] + "/bias"])
            dense = dense[:-2]
    for i, layer in enumerate(dense):
        new = "dense" if i == 0 else f"dense_{i}"
        out[new + "/kernel"] = variables[layer + "/kernel"]
        out[new + "/bias"] = variables[layer + "/bias"]
    return out, head
