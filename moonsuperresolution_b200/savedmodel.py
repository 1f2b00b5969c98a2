"""Reads the weights of the reference's trained models straight from their Keras SavedModel directories -- no TensorFlow.

The reference stores a trained GauGAN / CNNSpade as ``<path>/generator`` and ``<path>/encoder`` (plus ``discriminator``,
unused at inference) with ``keras.Model.save`` (spade/models/model.py:569-605, 798-820) and reloads them with
``models.load_model`` (:607-610, :822-824; process_full_tiles.py:30,50).  A SavedModel directory holds the variables as
a TensorBundle:

    <dir>/variables/variables.index                 an immutable sorted string table (LevelDB table format)
    <dir>/variables/variables.data-00000-of-0000N   raw little-endian tensor bytes

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table, tensor_bundle.proto; restated from the published
LevelDB ``table_format`` description -- PARITY UNPINNED: no TensorFlow-written file exists in this environment, the
reader is exercised against an independent writer in tests/tf_bundle_writer.py):

  * table footer = last 48 bytes: metaindex BlockHandle, index BlockHandle (varint64 offset, varint64 size each),
    zero padding to 40 bytes, magic 0xdb4775248b80fb57 (little endian);
  * block = payload + 1 type byte (0 raw, 1 snappy) + 4 bytes masked CRC-32C of payload+type; payload = entries
    (varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value) followed by the uint32 restart
    offsets and their count;
  * index block: value = BlockHandle of a data block; data blocks: key = checkpoint key, value = BundleEntryProto
    (1 dtype, 2 shape{2 dim{1 size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c fixed32); key "" = BundleHeaderProto.

Checkpoint keys follow Keras' object graph: ``layer_with_weights-<k>/<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`` with
k counting the weighted layers of ``model.layers`` in construction order, attribute names as written in the
reference's layer classes (blocks.py:17-26, spade.py:9-11).  ``load_gaugan_weights`` maps them onto this package's
tensor names (weights.py) and checks every shape.

Which key is which tensor is decided from the bundle's own object graph when it is there: the string tensor
``_CHECKPOINTABLE_OBJECT_GRAPH`` holds a serialized ``TrackableObjectGraph`` (tensorflow/core/protobuf/
trackable_object_graph.proto: nodes[] with children {node_id, local_name} and attributes {name, full_name,
checkpoint_key}).  ``resolve_generator_keys`` / ``resolve_encoder_keys`` walk it from the root: a weighted layer is a
ResidualBlock when it has the children ``spade_1`` / ``conv_1`` (blocks.py:17-20), a Sequential encoder block when it has
its own ``layer_with_weights-0``, and the two same-shaped Dense heads are told apart by the variables' ``full_name``
(``mean/kernel`` vs ``variance/kernel``, networks.py:32-33) -- not by their position.  Only when the bundle carries no
object graph do the positional maps below (``generator_key_map`` / ``encoder_key_map``) apply.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, List, Tuple

import numpy as np

from . import weights as W

TABLE_MAGIC = 0xDB4775248B80FB57
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 19: np.dtype("<f2")}


class BundleError(ValueError):
    pass


# ----------------------------------------------------------------------------------------------------------------------
# primitives: varints, CRC-32C, snappy
# ----------------------------------------------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint too long")


_CRC_TABLE = None


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C (Castagnoli, reflected polynomial 0x82F63B78)."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for n in range(256):
            c = n
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    """LevelDB / TensorFlow store CRCs rotated and offset so that a CRC of data containing CRCs stays well mixed."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def snappy_decompress(buf: bytes) -> bytes:
    """Raw snappy block: varint uncompressed length, then literal (tag & 3 == 0) and copy (1: 11-bit offset, 2: 16-bit,
    3: 32-bit) elements; copies may overlap their own output."""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                extra = ln - 59
                ln = int.from_bytes(buf[pos:pos + extra], "little")
                pos += extra
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise BundleError("snappy: bad copy offset")
        start = len(out) - off
        for k in range(ln):                       # byte-wise: a copy may read what it has just written
            out.append(out[start + k])
    if len(out) != n:
        raise BundleError("snappy: length mismatch")
    return bytes(out)


# ----------------------------------------------------------------------------------------------------------------------
# table
# ----------------------------------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise BundleError("block handle points outside the file")
    payload = data[offset:offset + size]
    kind = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if stored != masked_crc(data[offset:offset + size + 1]):
            raise BundleError("block checksum mismatch")
    if kind == 0:
        return payload
    if kind == 1:
        return snappy_decompress(payload)
    raise BundleError(f"unknown block compression {kind}")


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise BundleError("block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise BundleError("bad restart array")
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key):
            raise BundleError("bad key prefix length")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(data: bytes, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a LevelDB-format table, in key order."""
    if len(data) < 48:
        raise BundleError("file too small for a table footer")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise BundleError("not a TensorFlow / LevelDB table (bad magic)")
    _, pos = _varint(footer, 0)          # metaindex offset
    _, pos = _varint(footer, pos)        # metaindex size
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    pairs: List[Tuple[bytes, bytes]] = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        pairs += _block_entries(_read_block(data, boff, bsize, verify))
    return pairs


# ----------------------------------------------------------------------------------------------------------------------
# bundle
# ----------------------------------------------------------------------------------------------------------------------
def _proto_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """Minimal protobuf wire decoder: [(field number, wire type, value)]."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise BundleError(f"unsupported protobuf wire type {wt}")
        out.append((field, wt, v))
    return out


def _parse_entry(buf: bytes) -> dict:
    e = dict(dtype=0, shape=(), shard=0, offset=0, size=0, crc=None, sliced=False)
    for field, _, v in _proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            dims = []
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif field == 3:
            e["shard"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc"] = v
        elif field == 7:
            e["sliced"] = True
    return e


def read_bundle(prefix: str, verify_blocks: bool = True, verify_tensors: bool = False) -> Dict[str, np.ndarray]:
    """``prefix`` = ``<dir>/variables/variables``.  Returns checkpoint key -> array for every numeric tensor of the
    bundle (string tensors such as the object graph are skipped)."""
    index_path = prefix + ".index"
    if not os.path.exists(index_path):
        raise BundleError(f"{index_path} does not exist")
    with open(index_path, "rb") as f:
        pairs = read_table(f.read(), verify_blocks)
    num_shards, shards = 1, {}
    out: Dict[str, np.ndarray] = {}
    for key, value in pairs:
        if key == b"":
            for field, _, v in _proto_fields(value):
                if field == 1:
                    num_shards = v
                elif field == 2 and v != 0:
                    raise BundleError("big-endian bundles are not supported")
            continue
        e = _parse_entry(value)
        if e["dtype"] not in _DTYPES or e["sliced"]:
            continue
        dt = _DTYPES[e["dtype"]]
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise BundleError(f"{key!r}: {e['size']} bytes do not hold shape {e['shape']} of {dt}")
        sid = e["shard"]
        if sid not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, sid, num_shards)
            if not os.path.exists(path):
                raise BundleError(f"{path} does not exist")
            shards[sid] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise BundleError(f"{key!r}: data shard too short")
        if verify_tensors and e["crc"] is not None and masked_crc(raw.tobytes()) != e["crc"]:
            raise BundleError(f"{key!r}: tensor checksum mismatch")
        out[key.decode("utf-8")] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e["shape"]).copy()
    return out


def read_saved_model_variables(directory: str, **kw) -> Dict[str, np.ndarray]:
    """Variables of a Keras SavedModel directory, keyed by object-graph path without the ``.ATTRIBUTES`` suffix."""
    tensors = read_bundle(os.path.join(directory, "variables", "variables"), **kw)
    return {k[:-len(SUFFIX)]: v for k, v in tensors.items() if k.endswith(SUFFIX)}


# ----------------------------------------------------------------------------------------------------------------------
# object graph (trackable_object_graph.proto)
# ----------------------------------------------------------------------------------------------------------------------
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_BOOKKEEPING = {"keras_api", "variables", "trainable_variables", "non_trainable_variables", "regularization_losses",
                "layers", "signatures", "metrics", "layer_metrics", "layer_regularization_losses", "optimizer",
                "non_trainable_variables", "call_and_return_all_conditional_losses", "__call__", "_default_save_signature"}


def _string_tensor(raw: bytes, count: int) -> List[bytes]:
    """tensor_bundle.cc WriteStringTensor: [varint64 length] * count, 4-byte checksum of the lengths, then the bytes."""
    pos, lengths = 0, []
    for _ in range(count):
        n, pos = _varint(raw, pos)
        lengths.append(n)
    pos += 4
    out = []
    for n in lengths:
        if pos + n > len(raw):
            raise BundleError("string tensor shorter than its lengths say")
        out.append(bytes(raw[pos:pos + n]))
        pos += n
    return out


def read_object_graph(prefix: str):
    """The TrackableObjectGraph of a bundle as a list of nodes ``{"children": {local_name: node_id}, "attributes":
    [{"name", "full_name", "checkpoint_key"}]}`` (node 0 = the saved object), or None when the bundle has none."""
    index_path = prefix + ".index"
    if not os.path.exists(index_path):
        raise BundleError(f"{index_path} does not exist")
    with open(index_path, "rb") as f:
        pairs = read_table(f.read(), True)
    num_shards, entry = 1, None
    for key, value in pairs:
        if key == b"":
            for field, _, v in _proto_fields(value):
                if field == 1:
                    num_shards = v
        elif key == OBJECT_GRAPH_KEY.encode():
            entry = _parse_entry(value)
    if entry is None or entry["dtype"] != 7:
        return None
    path = "%s.data-%05d-of-%05d" % (prefix, entry["shard"], num_shards)
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        f.seek(entry["offset"])
        raw = f.read(entry["size"])
    try:
        blob = _string_tensor(raw, 1)[0]
        nodes = []
        for field, wt, v in _proto_fields(blob):
            if field != 1 or wt != 2:
                continue
            node = {"children": {}, "attributes": []}
            for f2, w2, v2 in _proto_fields(v):
                if f2 == 1 and w2 == 2:
                    ref = {"node_id": 0, "local_name": ""}
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            ref["node_id"] = v3
                        elif f3 == 2:
                            ref["local_name"] = v3.decode("utf-8", "replace")
                    node["children"][ref["local_name"]] = ref["node_id"]
                elif f2 == 2 and w2 == 2:
                    att = {"name": "", "full_name": "", "checkpoint_key": ""}
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 in (1, 2, 3) and isinstance(v3, (bytes, bytearray)):
                            att[("name", "full_name", "checkpoint_key")[f3 - 1]] = v3.decode("utf-8", "replace")
                    node["attributes"].append(att)
            nodes.append(node)
    except (BundleError, struct.error, IndexError):
        return None
    return nodes or None


def _weighted_children(nodes, node_id: int) -> List[Tuple[int, int]]:
    """[(k, node_id)] of the ``layer_with_weights-k`` children of a node, by k."""
    out = []
    for name, nid in nodes[node_id]["children"].items():
        m = re.fullmatch(r"layer_with_weights-(\d+)", name)
        if m and 0 <= nid < len(nodes):
            out.append((int(m.group(1)), nid))
    return sorted(out)


def _variable_leaves(nodes, node_id: int, prefix: str = "", seen=None) -> Dict[str, Tuple[str, str]]:
    """attribute path below a node -> (checkpoint key without the .ATTRIBUTES suffix, variable full_name)."""
    seen = set() if seen is None else seen
    out: Dict[str, Tuple[str, str]] = {}
    if node_id in seen:
        return out
    seen.add(node_id)
    for att in nodes[node_id]["attributes"]:
        if att["name"] == "VARIABLE_VALUE" and att["checkpoint_key"].endswith(SUFFIX):
            out[prefix.rstrip("/")] = (att["checkpoint_key"][:-len(SUFFIX)], att["full_name"])
    for name, nid in nodes[node_id]["children"].items():
        if name in _BOOKKEEPING or name.startswith("layer-") or name.startswith("_") or not (0 <= nid < len(nodes)):
            continue
        out.update(_variable_leaves(nodes, nid, prefix + name + "/", seen))
    return out


def resolve_generator_keys(nodes, found: Dict[str, np.ndarray]) -> Dict[str, str]:
    """checkpoint key -> weights.py name for build_generator (networks.py:37-57), from the object graph: ResidualBlocks
    by their attribute children (blocks.py:17-26) in layer order, the Dense by its rank-2 kernel, the output Conv2D by
    its rank-4 kernel."""
    m: Dict[str, str] = {}
    rb = 0
    for _, nid in _weighted_children(nodes, 0):
        kids = nodes[nid]["children"]
        leaves = _variable_leaves(nodes, nid)
        if "spade_1" in kids and "conv_1" in kids:
            rb += 1
            for path, (key, _) in leaves.items():
                m[key] = f"gen.rb{rb}." + path.replace("/", ".")
        elif "kernel" in leaves:
            rank = found[leaves["kernel"][0]].ndim if leaves["kernel"][0] in found else 0
            pre = "gen.dense" if rank == 2 else "gen.out" if rank == 4 else None
            if pre is None:
                raise BundleError("generator: a weighted layer is neither Dense, ResidualBlock nor Conv2D")
            for leaf in ("kernel", "bias"):
                if leaf in leaves:
                    m[leaves[leaf][0]] = f"{pre}.{leaf}"
    if rb != len(W.RB_FILTERS):
        raise BundleError(f"generator: the object graph holds {rb} ResidualBlocks, expected {len(W.RB_FILTERS)}")
    return m


def resolve_encoder_keys(nodes, found: Dict[str, np.ndarray]) -> Dict[str, str]:
    """checkpoint key -> weights.py name for build_encoder (networks.py:8-34): Sequential blocks in layer order
    (Conv2D kernel, then gamma / beta of the InstanceNormalization), Dense heads by variable name (``mean`` /
    ``variance``, networks.py:32-33) and only if the names are absent by order."""
    m: Dict[str, str] = {}
    block, heads = 0, []
    for _, nid in _weighted_children(nodes, 0):
        inner = _weighted_children(nodes, nid)
        if inner:
            block += 1
            for j, (_, sub) in enumerate(inner):
                for path, (key, _) in _variable_leaves(nodes, sub).items():
                    if path == "kernel":
                        m[key] = f"enc.down{block}.kernel"
                    elif path in ("gamma", "beta"):
                        m[key] = f"enc.down{block}.in_{path}"
        else:
            leaves = _variable_leaves(nodes, nid)
            if "kernel" in leaves:
                heads.append(leaves)
    if len(heads) != 2:
        raise BundleError(f"encoder: expected the two Dense heads, found {len(heads)}")
    names = [h["kernel"][1].split("/")[0].split(":")[0] for h in heads]
    if sorted(names) == ["mean", "variance"]:
        order = names
    else:                                   # unnamed variables: construction order (mean first, networks.py:32-33)
        order = ["mean", "variance"]
    for head, leaves in zip(order, heads):
        for leaf in ("kernel", "bias"):
            if leaf in leaves:
                m[leaves[leaf][0]] = f"enc.{head}.{leaf}"
    return m


# ----------------------------------------------------------------------------------------------------------------------
# positional keys -> this package's tensor names (bundles without an object graph)
# ----------------------------------------------------------------------------------------------------------------------
def generator_key_map(image_size: int) -> Dict[str, str]:
    """checkpoint key -> weights.py name for build_generator (networks.py:37-57): weighted layers in construction order
    are Dense (0), the six ResidualBlocks (1..6) and the output Conv2D (7)."""
    m = {"layer_with_weights-0/kernel": "gen.dense.kernel", "layer_with_weights-0/bias": "gen.dense.bias",
         "layer_with_weights-7/kernel": "gen.out.kernel", "layer_with_weights-7/bias": "gen.out.bias"}
    cin = 1024
    for k, cout in enumerate(W.RB_FILTERS, start=1):
        pre = f"layer_with_weights-{k}"
        spades = ["spade_1", "spade_2"] + (["spade_3"] if cin != cout else [])
        convs = ["conv_1", "conv_2"] + (["conv_3"] if cin != cout else [])
        for sp in spades:
            for sub in ("conv", "conv_gamma", "conv_beta"):                      # spade.py:9-11
                for leaf in ("kernel", "bias"):
                    m[f"{pre}/{sp}/{sub}/{leaf}"] = f"gen.rb{k}.{sp}.{sub}.{leaf}"
        for cv in convs:                                                          # blocks.py:19-26
            for leaf in ("kernel", "bias"):
                m[f"{pre}/{cv}/{leaf}"] = f"gen.rb{k}.{cv}.{leaf}"
        cin = cout
    return m


def encoder_key_map() -> Dict[str, str]:
    """build_encoder (networks.py:8-34): five Sequential blocks (0..4; Conv2D then, from the second block on,
    InstanceNormalization, blocks.py:50-63), then the Dense heads ``mean`` (5) and ``variance`` (6)."""
    m = {}
    for k in range(1, 6):
        pre = f"layer_with_weights-{k - 1}"
        m[f"{pre}/layer_with_weights-0/kernel"] = f"enc.down{k}.kernel"
        if k > 1:
            m[f"{pre}/layer_with_weights-1/gamma"] = f"enc.down{k}.in_gamma"
            m[f"{pre}/layer_with_weights-1/beta"] = f"enc.down{k}.in_beta"
    for k, head in ((5, "mean"), (6, "variance")):
        m[f"layer_with_weights-{k}/kernel"] = f"enc.{head}.kernel"
        m[f"layer_with_weights-{k}/bias"] = f"enc.{head}.bias"
    return m


def _map_variables(found: Dict[str, np.ndarray], key_map: Dict[str, str], what: str) -> Dict[str, np.ndarray]:
    out = {}
    for key, name in key_map.items():
        if key not in found:
            near = sorted(k for k in found if k.split("/")[0] == key.split("/")[0])[:6]
            raise BundleError(f"{what}: variable {key!r} ({name}) is missing; the bundle has e.g. {near}")
        out[name] = np.ascontiguousarray(found[key], dtype=np.float32)
    return out


def load_gaugan_weights(generator_dir: str, encoder_dir: str, image_size: int, arch: str = "spade",
                        **kw) -> Dict[str, np.ndarray]:
    """The Keras-layout weight dict of weights.model_spec(arch, image_size) from the two SavedModel directories the
    reference writes; every tensor is checked against the spec's shape."""
    def load(directory, resolve, positional, what):
        found = read_saved_model_variables(directory, **kw)
        nodes = read_object_graph(os.path.join(directory, "variables", "variables"))
        key_map = None
        if nodes is not None:
            resolved = resolve(nodes, found)
            if sorted(resolved.values()) == sorted(positional.values()):   # every tensor of the spec located by name
                key_map = resolved
        if key_map is None:
            key_map = positional
        inverse = {name: key for key, name in key_map.items()}
        return _map_variables(found, {inverse[name]: name for name in positional.values()}, what)
    gen = load(generator_dir, resolve_generator_keys, generator_key_map(image_size), "generator")
    enc = load(encoder_dir, resolve_encoder_keys, encoder_key_map(), "encoder")
    gen.update(enc)
    W.check_weights(arch, image_size, gen)
    return gen


def is_saved_model_dir(path: str) -> bool:
    return os.path.exists(os.path.join(path, "variables", "variables.index"))
