"""TEST INFRASTRUCTURE: writes a TensorFlow TensorBundle (``<prefix>.index`` + ``<prefix>.data-00000-of-00001``) the way
tensorflow/core/util/tensor_bundle's BundleWriter does, from the published LevelDB table format -- an independent
counterpart of moonsuperresolution_b200/savedmodel.py (no TensorFlow in this environment).  Also lays out the Keras
SavedModel directories of the reference's ``GauGAN.save`` (spade/models/model.py:569-605) with the object-graph keys
Keras derives from the layer classes (blocks.py:17-26, spade.py:9-11, networks.py:8-57)."""
import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57


def varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    return c ^ 0xFFFFFFFF


def mask(c: int) -> int:
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def snappy_compress(data: bytes) -> bytes:
    """A deliberately simple snappy encoder: greedy 4-byte hash matching, emits literals and 2-byte-offset copies (plus
    1-byte-offset copies for short near matches), enough to exercise every element type but the 4-byte-offset copy."""
    out = bytearray(varint(len(data)))
    table, i, lit_start, n = {}, 0, 0, len(data)

    def emit_literal(lo, hi):
        while lo < hi:
            ln = min(hi - lo, 1 << 16)
            if ln <= 60:
                out.append((ln - 1) << 2)
            elif ln <= 256:
                out.append(60 << 2)
                out.append(ln - 1)
            else:
                out.append(61 << 2)
                out.extend(struct.pack("<H", ln - 1))
            out.extend(data[lo:lo + ln])
            lo += ln

    while i + 4 <= n:
        key = data[i:i + 4]
        j = table.get(key)
        table[key] = i
        if j is not None and 0 < i - j < 65536:
            ln = 4
            while i + ln < n and ln < 64 and data[j + ln] == data[i + ln]:
                ln += 1
            emit_literal(lit_start, i)
            off = i - j
            if 4 <= ln <= 11 and off < 2048:
                out.append(1 | ((ln - 4) << 2) | ((off >> 8) << 5))
                out.append(off & 0xFF)
            else:
                out.append(2 | ((ln - 1) << 2))
                out += struct.pack("<H", off)
            i += ln
            lit_start = i
        else:
            i += 1
    emit_literal(lit_start, n)
    return bytes(out)


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_table(path: str, pairs, block_size=4096, compress=False):
    """pairs: sorted [(key bytes, value bytes)]."""
    f = bytearray()

    def emit(block: bytes):
        kind, payload = 0, block
        if compress:
            c = snappy_compress(block)
            if len(c) < len(block) - len(block) // 8:
                kind, payload = 1, c
        off = len(f)
        trailer = bytes([kind])
        f.extend(payload + trailer + struct.pack("<I", mask(crc32c(payload + trailer))))
        return varint(off) + varint(len(payload))

    index = _BlockBuilder(restart_interval=1)
    cur = _BlockBuilder()
    for key, value in pairs:
        cur.add(key, value)
        if len(cur.buf) >= block_size:
            index.add(key, emit(cur.finish()))
            cur = _BlockBuilder()
    if cur.count:
        index.add(cur.last, emit(cur.finish()))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    f.extend(footer)
    with open(path, "wb") as fh:
        fh.write(f)


def _field(num, wt, payload: bytes) -> bytes:
    return varint((num << 3) | wt) + payload


def entry_proto(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(_field(2, 2, varint(len(d)) + d) for d in (_field(1, 0, varint(int(s))) for s in shape))
    out = _field(1, 0, varint(dtype)) + _field(2, 2, varint(len(dims)) + dims)
    if offset:
        out += _field(4, 0, varint(offset))
    out += _field(5, 0, varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


def write_bundle(prefix: str, tensors: dict, compress=False, with_crc=False, extra_string_keys=(), string_tensors=None):
    """tensors: checkpoint key -> float32 ndarray; string_tensors: key -> bytes (scalar DT_STRING tensors in
    tensor_bundle.cc's layout: varint64 length, 4-byte masked CRC-32C of the lengths, the bytes)."""
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    pairs = [(b"", _field(1, 0, varint(1)) + _field(3, 2, varint(2) + _field(1, 0, varint(1))))]   # header: 1 shard
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for key in sorted(tensors):
            a = np.ascontiguousarray(tensors[key], dtype="<f4")
            raw = a.tobytes()
            data.write(raw)
            crc = mask(crc32c(raw)) if with_crc else 0
            pairs.append((key.encode(), entry_proto(1, a.shape, offset, len(raw), crc)))
            offset += len(raw)
        for key in extra_string_keys:                       # a DT_STRING (= 7) entry with an unreadable payload
            data.write(b"\x03abc")
            pairs.append((key.encode(), entry_proto(7, (), offset, 4, 0)))
            offset += 4
        for key, blob in (string_tensors or {}).items():
            raw = varint(len(blob)) + struct.pack("<I", mask(crc32c(struct.pack("<Q", len(blob))))) + blob
            data.write(raw)
            pairs.append((key.encode(), entry_proto(7, (), offset, len(raw), 0)))
            offset += len(raw)
    pairs.sort(key=lambda kv: kv[0])
    write_table(prefix + ".index", pairs, compress=compress)


SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def keras_generator_keys(weights: dict) -> dict:
    """Object-graph checkpoint keys of build_generator (networks.py:37-57) for this package's tensor names."""
    out = {}
    for name, a in weights.items():
        if not name.startswith("gen."):
            continue
        parts = name.split(".")
        if parts[1] == "dense":
            key = f"layer_with_weights-0/{parts[2]}"
        elif parts[1] == "out":
            key = f"layer_with_weights-7/{parts[2]}"
        else:
            key = f"layer_with_weights-{int(parts[1][2:])}/" + "/".join(parts[2:])
        out[key + SUFFIX] = a
    return out


def keras_encoder_keys(weights: dict) -> dict:
    out = {}
    for name, a in weights.items():
        if not name.startswith("enc."):
            continue
        _, layer, leaf = name.split(".")
        if layer.startswith("down"):
            k = int(layer[4:]) - 1
            if leaf == "kernel":
                key = f"layer_with_weights-{k}/layer_with_weights-0/kernel"
            else:
                key = f"layer_with_weights-{k}/layer_with_weights-1/" + leaf[3:]      # in_gamma -> gamma
        else:
            key = f"layer_with_weights-{5 if layer == 'mean' else 6}/{leaf}"
        out[key + SUFFIX] = a
    return out


def object_graph_proto(keys: dict, full_names: dict = None) -> bytes:
    """A serialized TrackableObjectGraph (tensorflow/core/protobuf/trackable_object_graph.proto) for the checkpoint keys
    ``a/b/c/.ATTRIBUTES/VARIABLE_VALUE``: one node per path component, children {node_id, local_name}, the variable
    nodes carry attributes {name: VARIABLE_VALUE, full_name, checkpoint_key}.  Every non-leaf node also gets the
    bookkeeping children Keras adds (``keras_api``, ``variables`` -> a list node that points back at the variables), so
    that a reader has to skip them."""
    nodes = [{"children": {}, "attributes": []}]

    def child(parent, name):
        kids = nodes[parent]["children"]
        if name not in kids:
            nodes.append({"children": {}, "attributes": []})
            kids[name] = len(nodes) - 1
        return kids[name]
    for key in sorted(keys):
        assert key.endswith(SUFFIX)
        node = 0
        for part in key[:-len(SUFFIX)].split("/"):
            node = child(node, part)
        full = (full_names or {}).get(key, key[:-len(SUFFIX)].replace("layer_with_weights-", "layer_"))
        nodes[node]["attributes"].append(("VARIABLE_VALUE", full, key))
    for nid in range(len(nodes)):
        if nodes[nid]["children"] and not nodes[nid]["attributes"]:
            variables = [c for c in nodes[nid]["children"].values()]
            nodes.append({"children": {}, "attributes": []})                   # keras_api
            nodes[nid]["children"]["keras_api"] = len(nodes) - 1
            nodes.append({"children": {str(j): c for j, c in enumerate(variables)}, "attributes": []})   # list wrapper
            nodes[nid]["children"]["variables"] = len(nodes) - 1
    out = b""
    for node in nodes:
        body = b""
        for name, nid in node["children"].items():
            ref = _field(1, 0, varint(nid)) + _field(2, 2, varint(len(name.encode())) + name.encode())
            body += _field(1, 2, varint(len(ref)) + ref)
        for name, full, ckpt in node["attributes"]:
            st = b"".join(_field(k, 2, varint(len(v.encode())) + v.encode()) for k, v in ((1, name), (2, full), (3, ckpt)))
            body += _field(2, 2, varint(len(st)) + st)
        out += _field(1, 2, varint(len(body)) + body)
    return out


def write_gaugan_saved_models(root: str, weights: dict, compress=False, object_graph=False, swap_heads=False,
                              shift_generator_layers=0):
    """<root>/generator and <root>/encoder as Keras SavedModel directories (variables only + a stub saved_model.pb).
    ``object_graph``: also store the TrackableObjectGraph.  ``swap_heads``: the encoder's Dense heads are numbered the
    other way round (variance = layer_with_weights-5, mean = -6) and ``shift_generator_layers`` renumbers the
    generator's weighted layers -- both legal in a checkpoint whose object graph names the variables, both fatal for a
    reader that goes by position."""
    gen_keys, enc_keys = keras_generator_keys(weights), keras_encoder_keys(weights)
    full_names = {}
    if swap_heads:
        swapped = {}
        for key, a in enc_keys.items():
            if key.startswith("layer_with_weights-5/"):
                key = key.replace("layer_with_weights-5/", "layer_with_weights-6/")
            elif key.startswith("layer_with_weights-6/"):
                key = key.replace("layer_with_weights-6/", "layer_with_weights-5/")
            swapped[key] = a
        enc_keys = swapped
    for key in enc_keys:
        k = int(key.split("/")[0].split("-")[1])
        if k >= 5:
            head = ("mean", "variance")[(k - 5) ^ (1 if swap_heads else 0)]
            full_names[key] = head + "/" + key.split("/")[1]
    if shift_generator_layers:
        import re as _re
        gen_keys = {_re.sub(r"^layer_with_weights-(\d+)", lambda m: "layer_with_weights-%d" %
                            (int(m.group(1)) + shift_generator_layers), k): a for k, a in gen_keys.items()}
    for sub, keys in (("generator", gen_keys), ("encoder", enc_keys)):
        d = os.path.join(root, sub)
        if object_graph:
            write_bundle(os.path.join(d, "variables", "variables"), keys, compress=compress,
                         string_tensors={"_CHECKPOINTABLE_OBJECT_GRAPH": object_graph_proto(keys, full_names)})
        else:
            write_bundle(os.path.join(d, "variables", "variables"), keys, compress=compress,
                         extra_string_keys=("_CHECKPOINTABLE_OBJECT_GRAPH",))
        with open(os.path.join(d, "saved_model.pb"), "wb") as f:
            f.write(b"")
