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
# object-graph keys -> this package's tensor names
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
    gen = _map_variables(read_saved_model_variables(generator_dir, **kw), generator_key_map(image_size), "generator")
    enc = _map_variables(read_saved_model_variables(encoder_dir, **kw), encoder_key_map(), "encoder")
    gen.update(enc)
    W.check_weights(arch, image_size, gen)
    return gen


def is_saved_model_dir(path: str) -> bool:
    return os.path.exists(os.path.join(path, "variables", "variables.index"))
