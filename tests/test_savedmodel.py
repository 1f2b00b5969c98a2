"""SavedModel / TensorBundle reader against an independent writer of the same published format
(tests/tf_bundle_writer.py).  No TensorFlow-written file exists here: PARITY UNPINNED for the container format."""
import os
import struct

import numpy as np
import pytest

import tf_bundle_writer as TW
from moonsuperresolution_b200 import savedmodel as SM
from moonsuperresolution_b200 import weights as W


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors for CRC-32C
    assert SM.crc32c(b"123456789") == 0xE3069283
    assert SM.crc32c(bytes(32)) == 0x8A9136AA
    assert SM.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert SM.crc32c(bytes(range(32))) == 0x46DD794E
    assert SM.crc32c(b"123456789") == TW.crc32c(b"123456789")


def test_snappy_decoder():
    # hand-assembled stream: literal "abcd", copy(offset 4, len 8, 1-byte-offset form) -> overlapping run, literal "xyz"
    stream = TW.varint(15) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([1 | ((8 - 4) << 2) | (0 << 5), 4]) + \
        bytes([(3 - 1) << 2]) + b"xyz"
    assert SM.snappy_decompress(stream) == b"abcdabcdabcdxyz"
    # 2-byte-offset copy and a long literal (length byte form)
    body = bytes(range(200))
    stream = TW.varint(200 + 10) + bytes([60 << 2, 199]) + body + bytes([2 | ((10 - 1) << 2)]) + struct.pack("<H", 200)
    assert SM.snappy_decompress(stream) == body + body[:10]
    rng = np.random.default_rng(0)
    for blob in (b"", b"a", b"layer_with_weights-1/spade_1/conv/kernel" * 50, rng.integers(0, 4, 5000, dtype=np.uint8).tobytes()):
        assert SM.snappy_decompress(TW.snappy_compress(blob)) == blob
    with pytest.raises(SM.BundleError):
        SM.snappy_decompress(TW.varint(8) + bytes([2 | (7 << 2)]) + struct.pack("<H", 9))      # copy before any output


@pytest.mark.parametrize("compress", [False, True])
def test_bundle_round_trip_many_blocks(tmp_path, compress):
    rng = np.random.default_rng(1)
    tensors = {f"layer_with_weights-{k}/sub_{k % 7}/kernel" + TW.SUFFIX: rng.standard_normal((k % 5 + 1, 3, k % 4 + 1)).astype(np.float32)
               for k in range(300)}                                         # > 4 KB of index: several data blocks
    tensors["scalar" + TW.SUFFIX] = np.float32(3.5).reshape(())
    prefix = str(tmp_path / "variables" / "variables")
    TW.write_bundle(prefix, tensors, compress=compress, with_crc=True, extra_string_keys=("_CHECKPOINTABLE_OBJECT_GRAPH",))
    got = SM.read_bundle(prefix, verify_blocks=True, verify_tensors=True)
    assert set(got) == set(tensors)
    for k, a in tensors.items():
        np.testing.assert_array_equal(got[k], a)
    # corruption is detected
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[10] ^= 0x40
    open(prefix + ".index", "wb").write(raw)
    with pytest.raises(SM.BundleError):
        SM.read_bundle(prefix)
    open(prefix + ".index", "wb").write(b"short")
    with pytest.raises(SM.BundleError):
        SM.read_bundle(prefix)


def test_gaugan_saved_model_directories_round_trip(tmp_path):
    """Layout of GauGAN.save (spade/models/model.py:569-605): <path>/generator, <path>/encoder as SavedModel directories
    with Keras object-graph keys -> the weight dict of weights.model_spec, shapes checked."""
    i = 64
    weights = W.random_init("spade", i, seed=3, perturb_affine=True)
    small = {k: v for k, v in weights.items()}
    TW.write_gaugan_saved_models(str(tmp_path), small, compress=True)
    assert SM.is_saved_model_dir(str(tmp_path / "generator")) and not SM.is_saved_model_dir(str(tmp_path))
    got = SM.load_gaugan_weights(str(tmp_path / "generator"), str(tmp_path / "encoder"), i)
    assert set(got) == set(weights)
    for k in weights:
        np.testing.assert_array_equal(got[k], weights[k])
    # key inventory: 8 weighted layers in the generator, 7 in the encoder
    gk = SM.generator_key_map(i)
    assert len({k.split("/")[0] for k in gk}) == 8 and len(gk) == len(W.spade_generator_spec(i))
    ek = SM.encoder_key_map()
    assert len({k.split("/")[0] for k in ek}) == 7 and len(ek) == len(W.encoder_spec(i))
    # a missing variable is reported by name
    os.remove(str(tmp_path / "encoder" / "variables" / "variables.index"))
    with pytest.raises(SM.BundleError):
        SM.load_gaugan_weights(str(tmp_path / "generator"), str(tmp_path / "encoder"), i)


def test_keys_are_resolved_from_the_object_graph_not_by_position(tmp_path):
    """A checkpoint whose weighted layers are numbered differently from Keras' usual construction order (Dense heads the
    other way round, generator layers shifted) still loads correctly, because the TrackableObjectGraph names the
    variables: ResidualBlocks by their attribute children (blocks.py:17-26), the same-shaped ``mean`` / ``variance`` heads
    by the variables' full_name (networks.py:32-33).  Without the graph the positional map would swap the heads
    silently -- which is why the graph, when present, wins."""
    i = 64
    weights = W.random_init("spade", i, seed=4, perturb_affine=True)
    TW.write_gaugan_saved_models(str(tmp_path), weights, compress=True, object_graph=True, swap_heads=True,
                                 shift_generator_layers=3)
    nodes = SM.read_object_graph(str(tmp_path / "encoder" / "variables" / "variables"))
    assert nodes is not None and "layer_with_weights-6" in nodes[0]["children"] and "keras_api" in nodes[0]["children"]
    got = SM.load_gaugan_weights(str(tmp_path / "generator"), str(tmp_path / "encoder"), i)
    assert set(got) == set(weights)
    for k in weights:
        np.testing.assert_array_equal(got[k], weights[k])
    # the same encoder bundle read by position alone swaps the heads (same shapes, so nothing would catch it)
    found = SM.read_saved_model_variables(str(tmp_path / "encoder"))
    by_position = SM._map_variables(found, SM.encoder_key_map(), "encoder")
    np.testing.assert_array_equal(by_position["enc.mean.kernel"], weights["enc.variance.kernel"])
    # usual numbering with a graph: both routes agree
    other = tmp_path / "plain"
    TW.write_gaugan_saved_models(str(other), weights, object_graph=True)
    again = SM.load_gaugan_weights(str(other / "generator"), str(other / "encoder"), i)
    for k in weights:
        np.testing.assert_array_equal(again[k], weights[k])


def test_user_supplied_keras_checkpoint(tmp_path):
    """MSR_KERAS_CHECKPOINT=<dir holding generator/ and encoder/ SavedModel directories written by the reference's
    GauGAN.save under real TensorFlow> [MSR_KERAS_IMAGE_SIZE=512]: the reader must locate every tensor of the spec with
    the right shape.  No such artefact exists in the build image (no TensorFlow), so this is skipped there."""
    root = os.environ.get("MSR_KERAS_CHECKPOINT")
    if not root:
        pytest.skip("set MSR_KERAS_CHECKPOINT to a directory written by the reference's GauGAN.save to run this")
    i = int(os.environ.get("MSR_KERAS_IMAGE_SIZE", "512"))
    got = SM.load_gaugan_weights(os.path.join(root, "generator"), os.path.join(root, "encoder"), i)
    W.check_weights("spade", i, got)
    assert SM.read_object_graph(os.path.join(root, "generator", "variables", "variables")) is not None


def test_against_tensorflows_own_proto_schema_and_crc(tmp_path):
    """tensorboard ships TensorFlow's compiled .proto schemas and a TensorFlow-team implementation of the masked CRC-32C
    (its TFRecord writer): independent of this repository.  (1) the reader's CRC equals theirs; (2) a TrackableObjectGraph
    built and serialised with the OFFICIAL trackable_object_graph_pb2 classes -- heads numbered the 'wrong' way round --
    is embedded in a bundle and resolved correctly by savedmodel.py's hand-written parser; (3) the test writer's own
    object graph parses under the official schema to the same structure."""
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    tog = pytest.importorskip("tensorboard.compat.proto.trackable_object_graph_pb2")
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 4096):
        blob = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert SM.crc32c(blob) == pw.crc32c(blob) and SM.masked_crc(blob) == pw.masked_crc32c(blob)
    i = 64
    weights = W.random_init("spade", i, seed=6, perturb_affine=True)
    enc_keys = TW.keras_encoder_keys(weights)
    swapped = {}
    for key, a in enc_keys.items():                       # variance = layer_with_weights-5, mean = -6
        key = key.replace("layer_with_weights-5/", "layer_with_weights-X/").replace("layer_with_weights-6/", "layer_with_weights-5/")
        swapped[key.replace("layer_with_weights-X/", "layer_with_weights-6/")] = a
    graph = tog.TrackableObjectGraph()
    graph.nodes.add()                                      # root
    ids = {(): 0}
    for key in sorted(swapped):
        parts = tuple(key[:-len(TW.SUFFIX)].split("/"))
        for depth in range(1, len(parts) + 1):
            path = parts[:depth]
            if path not in ids:
                graph.nodes.add()
                ids[path] = len(graph.nodes) - 1
                ref = graph.nodes[ids[path[:-1]]].children.add()
                ref.node_id, ref.local_name = ids[path], path[-1]
        att = graph.nodes[ids[parts]].attributes.add()
        att.name, att.checkpoint_key = "VARIABLE_VALUE", key
        top = int(parts[0].split("-")[1])
        att.full_name = (("variance/" if top == 5 else "mean/") + parts[-1]) if top >= 5 else "/".join(parts)
    d = tmp_path / "encoder"
    TW.write_bundle(str(d / "variables" / "variables"), swapped,
                    string_tensors={"_CHECKPOINTABLE_OBJECT_GRAPH": graph.SerializeToString()})
    nodes = SM.read_object_graph(str(d / "variables" / "variables"))
    assert nodes is not None and len(nodes) == len(graph.nodes)
    found = SM.read_saved_model_variables(str(d))
    resolved = SM.resolve_encoder_keys(nodes, found)
    assert resolved["layer_with_weights-5/kernel"] == "enc.variance.kernel"
    assert resolved["layer_with_weights-6/bias"] == "enc.mean.bias"
    assert sorted(resolved.values()) == sorted(SM.encoder_key_map().values())
    # and the other direction: the writer's hand-encoded graph under the official schema
    mine = tog.TrackableObjectGraph()
    mine.ParseFromString(TW.object_graph_proto(TW.keras_generator_keys(weights)))
    names = {c.local_name for c in mine.nodes[0].children}
    assert {f"layer_with_weights-{k}" for k in range(8)} <= names and "keras_api" in names
    keys = {a.checkpoint_key for n in mine.nodes for a in n.attributes}
    assert keys == set(TW.keras_generator_keys(weights))
