"""Hand-assembled TensorFlow GraphDef protobufs (TEST INFRASTRUCTURE ONLY), so that a THIRD-PARTY implementation of
TensorFlow's graph semantics -- OpenCV's TensorFlow importer, cv2.dnn.readNetFromTensorflow -- can execute the ops whose
conventions the generator oracle depends on (SAME padding of strided / even-kernel convolutions, Conv2DBackpropInput,
half-pixel nearest resize, FusedBatchNorm at inference, LeakyRelu alpha).  No TensorFlow is needed to write these files:
the wire format of graph.proto / node_def.proto / attr_value.proto / tensor.proto is encoded directly.

`emit_functional_model` walks the deferred graph that tests/golden/tf_numpy_shim.py records while the UNMODIFIED
reference code builds a functional Keras model (e.g. Pix2Pix.buildGenerator, pix2pix.py:87-108) and emits one TF node
chain per Keras layer -- the graph structure comes from the reference's source, the op semantics from OpenCV."""
from __future__ import annotations

import os

import numpy as np

DT_FLOAT, DT_INT32 = 1, 3


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


def _fld(num, wt, payload):
    return varint((num << 3) | wt) + payload


def _ld(num, b):
    return _fld(num, 2, varint(len(b)) + b)


def attr_s(s: bytes):
    return _ld(2, s)


def attr_f(v: float):
    import struct
    return _fld(4, 5, struct.pack("<f", v))


def attr_b(v: bool):
    return _fld(5, 0, varint(1 if v else 0))


def attr_type(t: int):
    return _fld(6, 0, varint(t))


def attr_list_i(vals):
    return _ld(1, _ld(3, b"".join(varint(v) for v in vals)))


def _shape(dims):
    return b"".join(_ld(2, _fld(1, 0, varint(int(d)))) for d in dims)


def attr_tensor(arr: np.ndarray):
    dt = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32}[arr.dtype]
    return _ld(8, _fld(1, 0, varint(dt)) + _ld(2, _shape(arr.shape)) + _ld(4, np.ascontiguousarray(arr).tobytes()))


def node(name: str, op: str, inputs=(), **attrs) -> bytes:
    b = _ld(1, name.encode()) + _ld(2, op.encode())
    for i in inputs:
        b += _ld(3, i.encode())
    for k, v in attrs.items():
        b += _ld(5, _ld(1, k.encode()) + _ld(2, v))
    return _ld(1, b)


def const(name: str, arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    arr = arr.astype(np.int32) if arr.dtype.kind in "iu" else arr.astype(np.float32)
    return node(name, "Const", dtype=attr_type(DT_INT32 if arr.dtype == np.int32 else DT_FLOAT), value=attr_tensor(arr))


def placeholder(name: str) -> bytes:
    return node(name, "Placeholder", dtype=attr_type(DT_FLOAT))


def conv2d(name, x, kernel, stride, padding=b"SAME") -> bytes:
    return const(name + "/w", kernel) + node(name, "Conv2D", [x, name + "/w"], T=attr_type(DT_FLOAT),
                                             strides=attr_list_i([1, stride, stride, 1]), padding=attr_s(padding),
                                             data_format=attr_s(b"NHWC"), dilations=attr_list_i([1, 1, 1, 1]))


def conv2d_transpose(name, x, kernel, stride, out_shape) -> bytes:
    """Keras Conv2DTranspose(padding='same') = Conv2DBackpropInput(output_shape, filter [kh, kw, cout, cin], x)."""
    return (const(name + "/shape", np.asarray(out_shape, np.int32)) + const(name + "/w", kernel) +
            node(name, "Conv2DBackpropInput", [name + "/shape", name + "/w", x], T=attr_type(DT_FLOAT),
                 strides=attr_list_i([1, stride, stride, 1]), padding=attr_s(b"SAME"), data_format=attr_s(b"NHWC"),
                 dilations=attr_list_i([1, 1, 1, 1])))


def bias_add(name, x, bias) -> bytes:
    return const(name + "/b", bias) + node(name, "BiasAdd", [x, name + "/b"], T=attr_type(DT_FLOAT),
                                           data_format=attr_s(b"NHWC"))


def fused_batch_norm(name, x, gamma, beta, mean, var, eps) -> bytes:
    """Keras BatchNormalization at training=False."""
    out = b""
    for tag, arr in (("gamma", gamma), ("beta", beta), ("mean", mean), ("var", var)):
        out += const(f"{name}/{tag}", arr)
    return out + node(name, "FusedBatchNorm", [x, name + "/gamma", name + "/beta", name + "/mean", name + "/var"],
                      T=attr_type(DT_FLOAT), epsilon=attr_f(eps), is_training=attr_b(False), data_format=attr_s(b"NHWC"))


def leaky_relu(name, x, alpha) -> bytes:
    return node(name, "LeakyRelu", [x], T=attr_type(DT_FLOAT), alpha=attr_f(alpha))


def unary(name, op, x) -> bytes:
    return node(name, op, [x], T=attr_type(DT_FLOAT))


def const_scalar_int(name: str, v: int) -> bytes:
    """Scalar int32 constant the way TensorFlow serialises it (empty shape, `int_val`): OpenCV's importer reads the axis
    of ConcatV2 from that field (a `tensor_content` scalar crashes it)."""
    t = _fld(1, 0, varint(DT_INT32)) + _ld(2, b"") + _ld(7, varint(v))
    return node(name, "Const", dtype=attr_type(DT_INT32), value=_ld(8, t))


def concat(name, xs, axis=3) -> bytes:
    return const_scalar_int(name + "/axis", axis) + node(name, "ConcatV2", list(xs) + [name + "/axis"],
                                                         T=attr_type(DT_FLOAT), N=_fld(3, 0, varint(len(xs))),
                                                         Tidx=attr_type(DT_INT32))


def resize_nearest(name, x, size, half_pixel_centers=True) -> bytes:
    return const(name + "/size", np.asarray(size, np.int32)) + node(
        name, "ResizeNearestNeighbor", [x, name + "/size"], T=attr_type(DT_FLOAT), align_corners=attr_b(False),
        half_pixel_centers=attr_b(half_pixel_centers))


def binary(name, op, a, b) -> bytes:
    return node(name, op, [a, b], T=attr_type(DT_FLOAT))


def mean(name, x, axes, keep_dims=True) -> bytes:
    return const(name + "/axes", np.asarray(axes, np.int32)) + node(name, "Mean", [x, name + "/axes"], T=attr_type(DT_FLOAT),
                                                                    Tidx=attr_type(DT_INT32), keep_dims=attr_b(keep_dims))


def matmul(name, x, kernel) -> bytes:
    return const(name + "/w", kernel) + node(name, "MatMul", [x, name + "/w"], T=attr_type(DT_FLOAT),
                                             transpose_a=attr_b(False), transpose_b=attr_b(False))


def reshape(name, x, shape) -> bytes:
    return const(name + "/shape", np.asarray(shape, np.int32)) + node(name, "Reshape", [x, name + "/shape"],
                                                                      T=attr_type(DT_FLOAT), Tshape=attr_type(DT_INT32))


def run_opencv(graph: bytes, x_nhwc, path: str, output: str = None) -> np.ndarray:
    """Executes the graph with OpenCV's TensorFlow importer; NHWC in, NHWC out.  ``x_nhwc``: one array for the
    Placeholder ``input``, or a dict placeholder name -> array."""
    import cv2
    with open(path, "wb") as f:
        f.write(graph)
    net = cv2.dnn.readNetFromTensorflow(path)
    try:
        os.remove(path)          # the graphs of whole generators are hundreds of MB: do not leave them in pytest's tmp dirs
    except OSError:
        pass
    feeds = x_nhwc if isinstance(x_nhwc, dict) else {"input": x_nhwc}
    for name, a in feeds.items():
        a = np.asarray(a, np.float32)
        net.setInput(np.ascontiguousarray(a.transpose(0, 3, 1, 2)) if a.ndim == 4 else a, name)
    if output and output not in net.getLayerNames() and output.endswith("_bias"):
        output = output[:-len("_bias")]       # the importer folds BiasAdd into the Conv2D / MatMul layer before it
    y = net.forward(output) if output else net.forward()
    return y.transpose(0, 2, 3, 1) if y.ndim == 4 else y


def emit_functional_model(model, shim, batch: int, in_hw: int) -> bytes:
    """GraphDef for a functional model built on the numpy shim (after its layers have been built and given weights):
    walks the recorded Sym graph from the output, one TF node chain per Keras layer.  Supports what
    Pix2Pix.buildGenerator uses: Sequential[Conv2D | Conv2DTranspose, BatchNormalization, LeakyReLU | ReLU, Dropout],
    Concatenate, Conv2DTranspose with bias and tanh."""
    out, names, shapes, counter = [placeholder("input")], {}, {}, [0]

    def fresh(tag):
        counter[0] += 1
        return f"{tag}_{counter[0]}"

    def emit_layer(layer, x, hw):
        """-> (output node name, output spatial size)"""
        if isinstance(layer, shim.Sequential):
            for sub in layer.layers:
                x, hw = emit_layer(sub, x, hw)
            return x, hw
        if isinstance(layer, shim.Conv2DTranspose):
            n = fresh("convT")
            hw2 = hw * layer.s
            out.append(conv2d_transpose(n, x, np.asarray(layer.kernel, np.float32), layer.s, [batch, hw2, hw2, layer.filters]))
            x = n
            if layer.use_bias:
                out.append(bias_add(n + "_bias", x, np.asarray(layer.bias, np.float32)))
                x = n + "_bias"
            hw = hw2
        elif isinstance(layer, shim.Conv2D):
            n = fresh("conv")
            out.append(conv2d(n, x, np.asarray(layer.kernel, np.float32), layer.s, layer.padding.upper().encode()))
            x = n
            if layer.use_bias:
                out.append(bias_add(n + "_bias", x, np.asarray(layer.bias, np.float32)))
                x = n + "_bias"
            hw = -(-hw // layer.s)
        elif isinstance(layer, shim.BatchNormalization):
            n = fresh("bn")
            out.append(fused_batch_norm(n, x, layer.gamma, layer.beta, layer.moving_mean, layer.moving_variance,
                                        layer.epsilon))
            return n, hw
        elif isinstance(layer, shim.LeakyReLU):
            n = fresh("lrelu")
            out.append(leaky_relu(n, x, layer.alpha))
            return n, hw
        elif isinstance(layer, shim.ReLU):
            n = fresh("relu")
            out.append(unary(n, "Relu", x))
            return n, hw
        elif isinstance(layer, shim.Dropout):
            return x, hw
        else:
            raise NotImplementedError(type(layer).__name__)
        act = getattr(layer, "act_name", None)
        if act == "tanh":
            out.append(unary(x + "_tanh", "Tanh", x))
            x = x + "_tanh"
        elif act == "relu":
            out.append(unary(x + "_relu", "Relu", x))
            x = x + "_relu"
        return x, hw

    def visit(sym):
        if id(sym) in names:
            return names[id(sym)], shapes[id(sym)]
        if sym.layer is None and not sym.args:           # Input
            names[id(sym)], shapes[id(sym)] = "input", in_hw
            return "input", in_hw
        if isinstance(sym.layer, shim.Concatenate):
            parts = [visit(s) for s in sym.args[0]]
            n = fresh("concat")
            out.append(concat(n, [p[0] for p in parts]))
            names[id(sym)], shapes[id(sym)] = n, parts[0][1]
            return n, parts[0][1]
        x, hw = visit(sym.args[0])
        n, hw = emit_layer(sym.layer, x, hw)
        names[id(sym)], shapes[id(sym)] = n, hw
        return n, hw

    final, _ = visit(model.outputs)
    return b"".join(out), final
