"""A THIRD-PARTY implementation of TensorFlow's graph semantics as an anchor for the generator oracle: OpenCV's
TensorFlow importer (cv2.dnn.readNetFromTensorflow) executes hand-assembled GraphDef files (tests/golden/tf_graphdef.py;
no TensorFlow needed to write them).  TensorFlow itself is not installable in the build image; OpenCV's importer exists to
run real TensorFlow models, so its reading of the conventions below is independent of this repository's authors:

  * Conv2D padding='SAME' for 3x3 / 4x4 kernels, strides 1 and 2, even and odd sizes (k4s1 pads
    (1, 2), k3s2 on even sizes (0, 1)), Conv2DBackpropInput (= Keras Conv2DTranspose) 4x4 stride 2,
    ResizeNearestNeighbor with half_pixel_centers (tf.image.resize(method='nearest'), App. B.3), FusedBatchNorm at
    inference with epsilon 1e-3 and LeakyRelu alpha (App. B.7) -- against the numpy op shim AND the torch oracle;
  * the WHOLE pix2pix generator: the deferred graph that the UNMODIFIED pix2pix.py records on the shim
    (Pix2Pix.buildGenerator, pix2pix.py:87-108) is emitted layer by layer as a GraphDef and executed by OpenCV --
    structure from the reference's source, arithmetic from OpenCV -- and must agree with oracle/generator.pix2pix_call;
  * the WHOLE GauGAN.call / CNNSpade.call: the bodies cut out of model.py run on the shim with TRACED tensors, so every
    op the reference's code performs inside build_encoder, GaussianSampler.call, build_generator, ResidualBlock.call and
    SPADE.call is written down as a GraphDef node (~340 nodes, 400 MB with the 100 M weights) and OpenCV executes that
    graph (batch of one: OpenCV reduces over spatial axes only; I = 64: its Reshape to 4-D is layout-free only for
    sw = 1).  The outputs are committed as tests/golden/generator_opencv_tf.npz (tests/golden/make_golden_tf.py
    --opencv), so the oracle is held to them on any box; where the reference checkout is present the run is repeated live.
"""
import os
import sys

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)

import tf_graphdef as TG           # noqa: E402
import tf_numpy_shim as SH         # noqa: E402
from oracle import generator as OG  # noqa: E402


def nchw(a):
    return torch.from_numpy(np.ascontiguousarray(a)).permute(0, 3, 1, 2)


@pytest.mark.parametrize("k,s,size", [(3, 1, 5), (4, 1, 6), (3, 2, 6), (3, 2, 5), (4, 2, 6), (4, 2, 7), (4, 1, 9)])
def test_same_convolution_as_opencv_reads_tensorflow(tmp_path, k, s, size):
    rng = np.random.default_rng(k * 100 + s * 10 + size)
    x = rng.standard_normal((2, size, size, 3)).astype(np.float32)
    w = rng.standard_normal((k, k, 3, 2)).astype(np.float32)
    got = TG.run_opencv(TG.placeholder("input") + TG.conv2d("y", "input", w, s), x, str(tmp_path / "g.pb"))
    SH.set_dtype(np.float64)
    np.testing.assert_allclose(got, SH.conv2d(x, w, s, "same"), atol=2e-5)
    np.testing.assert_allclose(got, OG.conv2d_same(nchw(x), w, None, stride=s).permute(0, 2, 3, 1).numpy(), atol=2e-5)


def test_transposed_convolution_as_opencv_reads_tensorflow(tmp_path):
    rng = np.random.default_rng(5)
    for h in (3, 4):
        x = rng.standard_normal((2, h, h, 3)).astype(np.float32)
        w = rng.standard_normal((4, 4, 2, 3)).astype(np.float32)              # Keras layout [kh, kw, cout, cin]
        g = TG.placeholder("input") + TG.conv2d_transpose("y", "input", w, 2, [2, 2 * h, 2 * h, 2])
        got = TG.run_opencv(g, x, str(tmp_path / "g.pb"))
        SH.set_dtype(np.float64)
        np.testing.assert_allclose(got, SH.conv2d_transpose(x, w, 2, "same"), atol=2e-5)
        np.testing.assert_allclose(got, OG.conv2d_transpose_same_s2(nchw(x), w, None).permute(0, 2, 3, 1).numpy(), atol=2e-5)


def test_half_pixel_nearest_resize_as_opencv_reads_tensorflow(tmp_path):
    x = np.arange(2 * 8 * 8 * 2, dtype=np.float32).reshape(2, 8, 8, 2)
    for out in (1, 2, 4, 8):
        g = TG.placeholder("input") + TG.resize_nearest("y", "input", [out, out], half_pixel_centers=True)
        got = TG.run_opencv(g, x, str(tmp_path / "g.pb"))
        np.testing.assert_array_equal(got, SH.resize_nearest(x, (out, out)).astype(np.float32))
        np.testing.assert_array_equal(got, OG.resize_nearest_tf(nchw(x), out).permute(0, 2, 3, 1).numpy())


def test_batchnorm_inference_and_leaky_relu_as_opencv_reads_tensorflow(tmp_path):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 4, 4, 5)).astype(np.float32)
    gamma, beta, mean = (rng.standard_normal(5).astype(np.float32) for _ in range(3))
    var = rng.uniform(0.5, 2.0, 5).astype(np.float32)
    g = (TG.placeholder("input") + TG.fused_batch_norm("bn", "input", gamma, beta, mean, var, 1e-3) +
         TG.leaky_relu("y", "bn", 0.3))
    got = TG.run_opencv(g, x, str(tmp_path / "g.pb"))
    t = (x - mean) / np.sqrt(var + 1e-3) * gamma + beta
    np.testing.assert_allclose(got, np.where(t > 0, t, 0.3 * t), atol=2e-5)


def test_pix2pix_generator_built_by_the_reference_and_run_by_opencv(tmp_path):
    """Pix2Pix.buildGenerator from the UNMODIFIED pix2pix.py on the shim -> GraphDef -> OpenCV's TensorFlow importer,
    against the torch oracle with the same weights (54 M parameters; the reference's fixed 256 x 256 x 2 input)."""
    import make_golden_tf as MG
    import reference_graph as RG
    if not RG.available():
        pytest.skip("reference checkout not present on this box")
    name, arch, i, b, ws, xs = MG.CASES[3]
    assert arch == "pix2pix"
    w, x, _ = MG.case_inputs(arch, i, b, ws, xs)
    ref = RG.Reference(backend="shim", dtype=np.float32)
    try:
        gen = ref.build_pix2pix(w)
        graph, final = TG.emit_functional_model(gen, SH, batch=b, in_hw=i)
        want_shim = ref._np(gen(x, training=False))
    finally:
        ref.close()
    got = TG.run_opencv(graph, x, str(tmp_path / "pix2pix.pb"))
    want = OG.pix2pix_call(x, w)
    assert got.shape == want.shape == (b, 256, 256, 1)
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), np.abs(got - want).max()
    assert np.abs(got - want_shim).max() <= 1e-4


@pytest.mark.parametrize("case", [c for c in __import__("make_golden_tf").OPENCV_CASES], ids=lambda c: c[0])
def test_oracle_matches_the_reference_graph_run_by_opencv(case):
    """oracle/generator.py against the committed outputs of the reference's own generator code executed by OpenCV's
    TensorFlow importer (float32 engine: agreement to a few float32 roundings of O(1) outputs)."""
    import make_golden_tf as MG
    name, arch, i, b, ws, xs = case
    z = np.load(os.path.join(GOLDEN, "generator_opencv_tf.npz"))
    assert str(z["backend"]).startswith("opencv-dnn-tensorflow-importer")
    w, x, eps = MG.case_inputs(arch, i, b, ws, xs)
    want = z[f"{name}.out"]
    got = OG.pix2pix_call(x, w) if arch == "pix2pix" else OG.gaugan_call(x, w, eps, arch)
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err <= 2e-5 * max(1.0, np.abs(want).max()), (name, err)


@pytest.mark.parametrize("arch", ["spade", "cnn"])
def test_gaugan_call_traced_from_the_reference_and_run_by_opencv(tmp_path, arch):
    """Live: model.py's call body + networks / blocks / spade / sampling.py on traced tensors -> GraphDef -> OpenCV."""
    import make_golden_tf as MG
    import reference_graph as RG
    if not RG.available():
        pytest.skip("reference checkout not present on this box")
    name, _, i, b, ws, xs = [c for c in MG.OPENCV_CASES if c[1] == arch][0]
    w, x, eps = MG.case_inputs(arch, i, b, ws, xs)
    ref = RG.Reference(backend="shim", dtype=np.float32)
    try:
        got, shim = ref.run_spade_opencv(arch, i, w, x, eps, str(tmp_path))
    finally:
        ref.close()
    want = OG.gaugan_call(x, w, eps, arch)
    assert got.shape == want.shape == (1, i, i, 1)
    assert np.abs(got - want).max() <= 2e-5 and np.abs(got - shim).max() <= 2e-5
    z = np.load(os.path.join(GOLDEN, "generator_opencv_tf.npz"))
    np.testing.assert_allclose(got, z[f"{name}.out"], atol=1e-6)


def test_reference_residual_block_with_learned_skip_run_by_opencv(tmp_path):
    """blocks.py's ResidualBlock (filters != input channels: the spade_3 / conv_3 branch) and spade.py's SPADE on a small
    tensor with two Placeholders (features and mask): a graph small enough to read, same route as the full model."""
    import reference_graph as RG
    if not RG.available():
        pytest.skip("reference checkout not present on this box")
    rng = np.random.default_rng(0)
    ref = RG.Reference(backend="shim", dtype=np.float32)
    try:
        rb = ref.blocks.ResidualBlock(filters=32, alpha=0.2)
        x = rng.standard_normal((1, 8, 8, 16)).astype(np.float32)
        mask = rng.uniform(-0.5, 0.5, (1, 64, 64, 2)).astype(np.float32)
        rb(x, mask)
        assert rb.learned_skip
        for layer in [getattr(sp, c) for sp in (rb.spade_1, rb.spade_2, rb.spade_3) for c in ("conv", "conv_gamma", "conv_beta")] + \
                [rb.conv_1, rb.conv_2, rb.conv_3]:
            layer.set_weights([rng.standard_normal(layer.kernel.shape).astype(np.float32) * 0.1,
                               rng.standard_normal(layer.bias.shape).astype(np.float32) * 0.1])
        want = np.asarray(rb(x, mask))
        tx = SH.start_trace(TG, x, "input")
        tm = SH.trace_input(mask, "mask")
        y = rb(tx, tm)
        graph = SH.stop_trace()
    finally:
        SH.TRACER = None
        ref.close()
    got = TG.run_opencv(graph, {"input": x, "mask": mask}, str(tmp_path / "rb.pb"), output=y.tf)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_hand_encoded_graphdefs_parse_under_tensorflows_own_schema(tmp_path):
    """tests/golden/tf_graphdef.py writes the GraphDef wire format by hand; tensorboard ships TensorFlow's compiled
    graph.proto / node_def.proto / attr_value.proto / tensor.proto.  Parsed with those, the nodes carry the intended ops,
    inputs, attributes and tensor contents -- so what OpenCV executes is what TensorFlow would read."""
    gp = pytest.importorskip("tensorboard.compat.proto.graph_pb2")
    rng = np.random.default_rng(2)
    k = rng.standard_normal((4, 4, 3, 2)).astype(np.float32)
    raw = (TG.placeholder("input") + TG.conv2d("c", "input", k, 2) + TG.bias_add("cb", "c", np.arange(2, dtype=np.float32)) +
           TG.leaky_relu("l", "cb", 0.2) + TG.resize_nearest("r", "l", [8, 8]) + TG.mean("m", "r", [1, 2]) +
           TG.concat("cat", ["r", "r"]) + TG.conv2d_transpose("t", "cat", rng.standard_normal((4, 4, 1, 4)).astype(np.float32), 2,
                                                               [1, 16, 16, 1]))
    g = gp.GraphDef()
    g.ParseFromString(raw)
    nodes = {n.name: n for n in g.node}
    assert [nodes[n].op for n in ("input", "c", "cb", "l", "r", "m", "cat", "t")] == \
        ["Placeholder", "Conv2D", "BiasAdd", "LeakyRelu", "ResizeNearestNeighbor", "Mean", "ConcatV2", "Conv2DBackpropInput"]
    c = nodes["c"]
    assert list(c.input) == ["input", "c/w"] and list(c.attr["strides"].list.i) == [1, 2, 2, 1]
    assert c.attr["padding"].s == b"SAME" and c.attr["data_format"].s == b"NHWC" and c.attr["T"].type == 1
    w = nodes["c/w"].attr["value"].tensor
    assert [d.size for d in w.tensor_shape.dim] == [4, 4, 3, 2] and w.dtype == 1
    np.testing.assert_array_equal(np.frombuffer(w.tensor_content, np.float32).reshape(4, 4, 3, 2), k)
    assert abs(nodes["l"].attr["alpha"].f - 0.2) < 1e-7
    assert nodes["r"].attr["half_pixel_centers"].b is True and nodes["r"].attr["align_corners"].b is False
    assert nodes["m"].attr["keep_dims"].b is True
    assert list(np.frombuffer(nodes["m/axes"].attr["value"].tensor.tensor_content, np.int32)) == [1, 2]
    assert nodes["cat"].attr["N"].i == 2 and list(nodes["cat/axis"].attr["value"].tensor.int_val) == [3]
    assert list(nodes["t"].input) == ["t/shape", "t/w", "cat"]
    assert list(np.frombuffer(nodes["t/shape"].attr["value"].tensor.tensor_content, np.int32)) == [1, 16, 16, 1]
