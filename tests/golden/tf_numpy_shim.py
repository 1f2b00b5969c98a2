"""A numpy stand-in for the slice of TensorFlow 2.x / Keras / tensorflow-addons that the reference's model files import
(TEST INFRASTRUCTURE ONLY).

Why: TensorFlow is not installable in the build image, so the generator oracle (oracle/generator.py, a torch
restatement) could only be checked against itself.  With this shim registered as ``tensorflow`` / ``tensorflow_addons``
the UNMODIFIED reference modules -- spade/models/{networks,blocks,spade,sampling}.py and pix2pix.py -- import and run, so
the *graph* (which layer feeds which, ResidualBlock.call, SPADE.call, build_generator, build_encoder,
Pix2Pix.buildGenerator) is executed from the reference's own source.  What this file restates is only the op-level
semantics of the ~20 Keras layers / tf functions those modules call, each as plain numpy (loops over kernel taps, no
torch, no scipy), from the public TensorFlow documentation:

  * Conv2D / Conv2DTranspose 'same' padding:  out = ceil(in / s); total = max((out - 1) * s + k - in, 0);
    before = total // 2 (tf.nn.convolution docs);  Conv2DTranspose = gradient of that convolution (scatter form)
  * tf.image.resize(method="nearest"): half-pixel centres, src = min(floor((dst + 0.5) * in / out), in - 1)
  * tf.nn.moments: mean and biased variance;  BatchNormalization(training=False): moving statistics, eps 1e-3
  * tfa InstanceNormalization: per-sample, per-channel moments over (H, W), eps 1e-3, gamma / beta
  * LeakyReLU() default alpha 0.3;  UpSampling2D nearest;  Flatten / Reshape row-major over (h, w, c)

The functional API (Input -> layers -> Model) is a deferred graph: calling a layer on a symbolic tensor records a
node, calling the Model on arrays evaluates the nodes.  Arrays are float64 by default (``set_dtype``).
"""
from __future__ import annotations

import sys
import types

import numpy as np

_DTYPE = np.float64


def set_dtype(dt) -> None:
    global _DTYPE
    _DTYPE = np.dtype(dt).type


# ----------------------------------------------------------------------------------------------------------------------
# deferred graph
# ----------------------------------------------------------------------------------------------------------------------
class Sym:
    """Symbolic tensor of the functional API: ``fn(*args)`` evaluated when the owning Model is called."""
    _counter = 0

    def __init__(self, fn, args, layer=None, pick=None):
        self.fn, self.args, self.layer, self.pick = fn, args, layer, pick
        Sym._counter += 1
        self.order = Sym._counter

    def __add__(self, o):
        return Sym(lambda a, b: a + b, (self, o))

    def __radd__(self, o):
        return Sym(lambda a, b: b + a, (self, o))

    def __sub__(self, o):
        return Sym(lambda a, b: a - b, (self, o))

    def __mul__(self, o):
        return Sym(lambda a, b: a * b, (self, o))

    def __truediv__(self, o):
        return Sym(lambda a, b: a / b, (self, o))


def _has_sym(x) -> bool:
    if isinstance(x, Sym):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_sym(v) for v in x)
    return False


def _evaluate(x, env):
    if isinstance(x, Sym):
        if id(x) not in env:
            args = [_evaluate(a, env) for a in x.args]
            env[id(x)] = x.fn(*args)
        return env[id(x)]
    if isinstance(x, list):
        return [_evaluate(v, env) for v in x]
    if isinstance(x, tuple):
        return tuple(_evaluate(v, env) for v in x)
    return x


def _symbolic(fn):
    """tf.* function usable on symbolic tensors (Keras wraps such calls in TFOpLambda layers)."""
    def wrapped(*args, **kwargs):
        if _has_sym(args):
            return Sym(lambda *a: fn(*a, **kwargs), args)
        return fn(*args, **kwargs)
    wrapped.__name__ = fn.__name__
    return wrapped


def _arr(x):
    return np.asarray(x, dtype=_DTYPE)


# ----------------------------------------------------------------------------------------------------------------------
# tracing: while the reference's code runs on arrays, also write down the TensorFlow graph it would have built
# ----------------------------------------------------------------------------------------------------------------------
class Tracer:
    """Collects TensorFlow GraphDef nodes (tests/golden/tf_graphdef.py) for every op executed on traced tensors, so that
    a third-party TensorFlow-graph executor (OpenCV's importer) can run the graph the reference's ``call`` bodies
    describe.  Tracing is value-carrying: every traced tensor also holds the shim's numpy result."""

    def __init__(self, tg):
        self.tg, self.nodes, self.count = tg, [], 0

    def fresh(self, tag):
        self.count += 1
        return f"{tag}_{self.count}"

    def emit(self, b):
        self.nodes.append(b)

    def name_of(self, x):
        """TF node name of an operand: traced tensors carry one, anything else becomes a Const."""
        if isinstance(x, TT):
            return x.tf
        n = self.fresh("const")
        self.emit(self.tg.const(n, np.asarray(x, np.float32)))
        return n

    def graph(self):
        return b"".join(self.nodes)


TRACER = None


class TT(np.ndarray):
    """Traced tensor: an ndarray that knows the name of the TF node producing it.  numpy ufuncs are switched off on it
    (``__array_ufunc__ = None``), so only the operations spelled out here -- each of which emits its TF node -- can
    consume one; anything else raises instead of silently dropping out of the trace."""
    __array_ufunc__ = None

    def __new__(cls, arr, name):
        obj = np.asarray(arr).view(cls)
        obj.tf = name
        return obj

    def __array_finalize__(self, obj):
        self.tf = getattr(obj, "tf", None)

    def _bin(self, o, op, fn, swap=False):
        a, b = (o, self) if swap else (self, o)
        n = TRACER.fresh(op.lower())
        # AddV2 / Mul commute: the traced operand goes first (TensorFlow writes `0.5 * t` as Mul(t, 0.5) too once the
        # constant is folded to the right; OpenCV's importer only accepts a constant as the second input)
        ga, gb = (b, a) if (swap and op in ("AddV2", "Mul")) else (a, b)
        TRACER.emit(TRACER.tg.binary(n, op, TRACER.name_of(ga), TRACER.name_of(gb)))
        return TT(fn(np.asarray(a, _DTYPE), np.asarray(b, _DTYPE)), n)

    def __add__(self, o):
        return self._bin(o, "AddV2", np.add)

    def __radd__(self, o):
        return self._bin(o, "AddV2", np.add, swap=True)

    def __sub__(self, o):
        return self._bin(o, "Sub", np.subtract)

    def __mul__(self, o):
        return self._bin(o, "Mul", np.multiply)

    def __rmul__(self, o):
        return self._bin(o, "Mul", np.multiply, swap=True)

    def __truediv__(self, o):
        return self._bin(o, "RealDiv", np.divide)


def start_trace(tg, x, name="input"):
    """Begins a trace: returns ``x`` as the traced Placeholder ``name``."""
    global TRACER
    TRACER = Tracer(tg)
    return trace_input(x, name)


def trace_input(x, name):
    """A further Placeholder of the running trace."""
    TRACER.emit(TRACER.tg.placeholder(name))
    return TT(_arr(x), name)


def stop_trace():
    global TRACER
    g, TRACER = TRACER.graph(), None
    return g


def _traced(x):
    return TRACER is not None and isinstance(x, TT)


def _unary_traced(op, x, value):
    n = TRACER.fresh(op.lower())
    TRACER.emit(TRACER.tg.unary(n, op, x.tf))
    return TT(value, n)


# ----------------------------------------------------------------------------------------------------------------------
# ops
# ----------------------------------------------------------------------------------------------------------------------
def same_padding(n_in: int, k: int, s: int):
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return out, total // 2, total - total // 2


def conv2d(x, kernel, stride: int, padding: str):
    """NHWC convolution (cross-correlation), kernel [kh, kw, cin, cout]: one tensordot per kernel tap."""
    x = _arr(x)
    kh, kw, cin, cout = kernel.shape
    n, h, w, c = x.shape
    assert c == cin, (x.shape, kernel.shape)
    if padding == "same":
        oh, pt, pb = same_padding(h, kh, stride)
        ow, pl, pr = same_padding(w, kw, stride)
    elif padding == "valid":
        oh, ow, pt, pb, pl, pr = (h - kh) // stride + 1, (w - kw) // stride + 1, 0, 0, 0, 0
    else:
        raise ValueError(padding)
    xp = np.zeros((n, h + pt + pb, w + pl + pr, c), _DTYPE)
    xp[:, pt:pt + h, pl:pl + w, :] = x
    y = np.zeros((n, oh, ow, cout), _DTYPE)
    for ky in range(kh):
        for kx in range(kw):
            win = xp[:, ky:ky + (oh - 1) * stride + 1:stride, kx:kx + (ow - 1) * stride + 1:stride, :]
            y += np.tensordot(win, _arr(kernel[ky, kx]), axes=([3], [0]))
    return y


def conv2d_transpose(x, kernel, stride: int, padding: str):
    """Keras Conv2DTranspose, kernel [kh, kw, cout, cin]: the gradient of conv2d w.r.t. its input, written as the
    scatter y[i*s + k - pad_before] += x[i] * w[k]; output size = in * stride for 'same'."""
    x = _arr(x)
    kh, kw, cout, cin = kernel.shape
    n, h, w, c = x.shape
    assert c == cin and padding == "same"
    oh, ow = h * stride, w * stride
    _, pt, _ = same_padding(oh, kh, stride)         # padding of the forward convolution oh -> h
    _, pl, _ = same_padding(ow, kw, stride)
    full = np.zeros((n, (h - 1) * stride + kh, (w - 1) * stride + kw, cout), _DTYPE)
    for ky in range(kh):
        for kx in range(kw):
            full[:, ky:ky + (h - 1) * stride + 1:stride, kx:kx + (w - 1) * stride + 1:stride, :] += \
                np.tensordot(x, _arr(kernel[ky, kx]), axes=([3], [1]))
    return full[:, pt:pt + oh, pl:pl + ow, :]


def resize_nearest(x, size):
    x = _arr(x)
    oh, ow = int(size[0]), int(size[1])
    h, w = x.shape[1:3]
    iy = [min(int(np.floor((d + 0.5) * (h / oh))), h - 1) for d in range(oh)]
    ix = [min(int(np.floor((d + 0.5) * (w / ow))), w - 1) for d in range(ow)]
    return x[:, iy][:, :, ix]


def _activation(name):
    if name is None or name == "linear":
        return lambda v: v
    if name == "relu":
        return lambda v: np.maximum(v, 0)
    if name == "tanh":
        return np.tanh
    raise ValueError(f"activation {name!r} not in the shim")


# ----------------------------------------------------------------------------------------------------------------------
# Keras layers
# ----------------------------------------------------------------------------------------------------------------------
class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.built = False
        self.weights_ = []          # names of the weight attributes, in Keras' get_weights() order

    def build(self, input_shape):
        pass

    def call(self, *args, **kwargs):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        kwargs.pop("training", None)
        if _has_sym(args):
            return Sym(lambda *a: self(*a, **kwargs), args, layer=self)
        if not self.built:
            first = args[0][0] if isinstance(args[0], (list, tuple)) else args[0]
            self.build(tuple(np.shape(first)))
            self.built = True
        return self.call(*args, **kwargs)

    def get_weights(self):
        return [getattr(self, n) for n in self.weights_]

    def set_weights(self, values):
        assert self.built, f"{type(self).__name__}: set_weights before the layer was built"
        assert len(values) == len(self.weights_), (type(self).__name__, len(values), self.weights_)
        for n, v in zip(self.weights_, values):
            assert tuple(np.shape(v)) == tuple(getattr(self, n).shape), (type(self).__name__, n, np.shape(v),
                                                                         getattr(self, n).shape)
            setattr(self, n, _arr(v))


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True,
                 kernel_initializer=None, **kwargs):
        super().__init__(**kwargs)
        self.filters, self.k = int(filters), int(kernel_size if np.isscalar(kernel_size) else kernel_size[0])
        self.s = int(strides if np.isscalar(strides) else strides[0])
        self.padding, self.act, self.use_bias = padding, _activation(activation), use_bias
        self.act_name = activation

    def build(self, input_shape):
        self.kernel = np.zeros((self.k, self.k, input_shape[-1], self.filters), _DTYPE)
        self.weights_ = ["kernel"]
        if self.use_bias:
            self.bias = np.zeros((self.filters,), _DTYPE)
            self.weights_.append("bias")

    def call(self, x):
        y = conv2d(x, self.kernel, self.s, self.padding)
        if self.use_bias:
            y = y + self.bias
        y = self.act(y)
        if _traced(x):
            tg, n = TRACER.tg, TRACER.fresh("conv")
            TRACER.emit(tg.conv2d(n, x.tf, np.asarray(self.kernel, np.float32), self.s, self.padding.upper().encode()))
            if self.use_bias:
                TRACER.emit(tg.bias_add(n + "_bias", n, np.asarray(self.bias, np.float32)))
                n += "_bias"
            if self.act_name in ("relu", "tanh"):
                TRACER.emit(tg.unary(n + "_act", self.act_name.capitalize(), n))
                n += "_act"
            return TT(y, n)
        return y


class Conv2DTranspose(Conv2D):
    def build(self, input_shape):
        self.kernel = np.zeros((self.k, self.k, self.filters, input_shape[-1]), _DTYPE)
        self.weights_ = ["kernel"]
        if self.use_bias:
            self.bias = np.zeros((self.filters,), _DTYPE)
            self.weights_.append("bias")

    def call(self, x):
        y = conv2d_transpose(x, self.kernel, self.s, self.padding)
        if self.use_bias:
            y = y + self.bias
        return self.act(y)


class Dense(Layer):
    def __init__(self, units, activation=None, **kwargs):
        super().__init__(**kwargs)
        self.units, self.act, self.act_name = int(units), _activation(activation), activation

    def build(self, input_shape):
        self.kernel = np.zeros((input_shape[-1], self.units), _DTYPE)
        self.bias = np.zeros((self.units,), _DTYPE)
        self.weights_ = ["kernel", "bias"]

    def call(self, x):
        y = self.act(_arr(x) @ self.kernel + self.bias)
        if _traced(x):
            assert self.act_name is None
            tg, n = TRACER.tg, TRACER.fresh("dense")
            TRACER.emit(tg.matmul(n, x.tf, np.asarray(self.kernel, np.float32)) +       # Keras Dense: MatMul + BiasAdd
                        tg.bias_add(n + "_bias", n, np.asarray(self.bias, np.float32)))
            return TT(y, n + "_bias")
        return y


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super().__init__(**kwargs)
        self.target = tuple(target_shape)

    def call(self, x):
        y = _arr(x).reshape((np.shape(x)[0],) + self.target)
        if _traced(x):
            n = TRACER.fresh("reshape")
            TRACER.emit(TRACER.tg.reshape(n, x.tf, (-1,) + self.target))
            return TT(y, n)
        return y


class Flatten(Layer):
    def call(self, x):
        y = _arr(x).reshape(np.shape(x)[0], -1)
        if _traced(x):
            n = TRACER.fresh("flatten")
            TRACER.emit(TRACER.tg.reshape(n, x.tf, (-1, y.shape[1])))
            return TT(y, n)
        return y


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), interpolation="nearest", **kwargs):
        super().__init__(**kwargs)
        assert interpolation == "nearest"
        self.size = (size, size) if np.isscalar(size) else tuple(size)

    def call(self, x):
        y = np.repeat(np.repeat(_arr(x), self.size[0], axis=1), self.size[1], axis=2)
        if _traced(x):   # Keras: backend.resize_images(..., interpolation="nearest") -> ResizeNearestNeighbor
            n = TRACER.fresh("upsample")
            TRACER.emit(TRACER.tg.resize_nearest(n, x.tf, [y.shape[1], y.shape[2]], half_pixel_centers=True))
            return TT(y, n)
        return y


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def call(self, x):
        return _leaky_relu(x, self.alpha)


class ReLU(Layer):
    def call(self, x):
        return np.maximum(_arr(x), 0)


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)

    def call(self, x):          # inference: identity
        return _arr(x)


class BatchNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma, self.beta = np.ones((c,), _DTYPE), np.zeros((c,), _DTYPE)
        self.moving_mean, self.moving_variance = np.zeros((c,), _DTYPE), np.ones((c,), _DTYPE)
        self.weights_ = ["gamma", "beta", "moving_mean", "moving_variance"]

    def call(self, x):          # training=False: moving statistics
        return (_arr(x) - self.moving_mean) / np.sqrt(self.moving_variance + self.epsilon) * self.gamma + self.beta


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs):
        return np.concatenate([_arr(v) for v in xs], axis=self.axis)


class ZeroPadding2D(Layer):
    def call(self, x):
        return np.pad(_arr(x), ((0, 0), (1, 1), (1, 1), (0, 0)))


class InstanceNormalization(Layer):
    """tensorflow_addons.layers.InstanceNormalization() defaults: GroupNormalization with groups = channels, axis -1,
    epsilon 1e-3, center and scale."""
    def __init__(self, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma, self.beta = np.ones((c,), _DTYPE), np.zeros((c,), _DTYPE)
        self.weights_ = ["gamma", "beta"]

    def call(self, x):
        if _traced(x):   # the same arithmetic through the traced operators: every step lands in the graph
            mu, var = _moments(x, (1, 2), keepdims=True)
            std = var + self.epsilon
            std = _unary_traced("Sqrt", std, np.sqrt(_arr(std)))
            return (x - mu) / std * np.asarray(self.gamma, np.float32) + np.asarray(self.beta, np.float32)
        x = _arr(x)
        mu = x.mean(axis=(1, 2), keepdims=True)
        var = ((x - mu) ** 2).mean(axis=(1, 2), keepdims=True)
        return (x - mu) / np.sqrt(var + self.epsilon) * self.gamma + self.beta


class Sequential(Layer):
    def __init__(self, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.layers = list(layers or [])

    def add(self, layer):
        self.layers.append(layer)

    def call(self, x):
        for l in self.layers:
            x = l(x)
        return x


def Input(shape=None, name=None, **kwargs):
    return Sym(None, (), layer=None)


class Model(Layer):
    """Functional model: Model(inputs, outputs).  (Subclassed models of the reference are not built by the harness.)"""
    def __init__(self, inputs=None, outputs=None, name=None, **kwargs):
        super().__init__(name=name)
        self.inputs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs
        self.built = True
        seen, order = set(), []

        def walk(s):
            if isinstance(s, (list, tuple)):
                for v in s:
                    walk(v)
            elif isinstance(s, Sym) and id(s) not in seen:
                seen.add(id(s))
                for a in s.args:
                    walk(a)
                if s.layer is not None:
                    order.append(s)
        walk(outputs)
        order.sort(key=lambda s: s.order)
        self.layers = []
        for s in order:
            if all(s.layer is not l for l in self.layers):
                self.layers.append(s.layer)

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def call(self, x):
        xs = x if isinstance(x, (list, tuple)) else [x]
        assert len(xs) == len(self.inputs)
        env = {id(s): (v if isinstance(v, TT) else _arr(v)) for s, v in zip(self.inputs, xs)}
        return _evaluate(self.outputs, env)


# ----------------------------------------------------------------------------------------------------------------------
# module objects
# ----------------------------------------------------------------------------------------------------------------------
class _Anything:
    """Placeholder for training-only API the model files touch at import / construction time (optimizers, losses,
    metrics): constructible, never used on the inference path."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("training-only API of the shim was called")


def _normal(shape, mean=0.0, stddev=1.0, **kwargs):
    raise RuntimeError("tf.random.normal is unseeded in the reference (sampling.py:13); the harness injects epsilon")


def install() -> types.ModuleType:
    """Registers the shim as ``tensorflow`` / ``tensorflow_addons`` (refuses to shadow a real TensorFlow)."""
    if "tensorflow" in sys.modules and not getattr(sys.modules["tensorflow"], "__msr_shim__", False):
        raise RuntimeError("a real tensorflow is already imported")

    def mod(name):
        m = types.ModuleType(name)
        m.__msr_shim__ = True
        sys.modules[name] = m
        return m
    tf = mod("tensorflow")
    keras = mod("tensorflow.keras")
    layers = mod("tensorflow.keras.layers")
    inits = mod("tensorflow.keras.initializers")
    tfa = mod("tensorflow_addons")
    tfa_layers = mod("tensorflow_addons.layers")
    for cls in (Layer, Conv2D, Conv2DTranspose, Dense, Reshape, Flatten, UpSampling2D, LeakyReLU, ReLU, Dropout,
                BatchNormalization, Concatenate, ZeroPadding2D):
        setattr(layers, cls.__name__, cls)
    layers.Input = Input
    layers.concatenate = lambda xs, axis=-1: Concatenate(axis=axis)(xs)
    inits.GlorotNormal = _Anything
    keras.layers, keras.initializers = layers, inits
    keras.Sequential, keras.Model, keras.Input = Sequential, Model, Input
    keras.optimizers = types.SimpleNamespace(Adam=_Anything)
    keras.losses = types.SimpleNamespace(BinaryCrossentropy=_Anything)
    keras.metrics = types.SimpleNamespace(Mean=_Anything)
    tf.keras = keras
    tf.nn = types.SimpleNamespace(leaky_relu=_symbolic(_leaky_relu), moments=_moments)
    tf.image = types.SimpleNamespace(resize=lambda x, size, method="bilinear": _resize(x, size, method))
    tf.sqrt = _symbolic(lambda x: _unary_traced("Sqrt", x, np.sqrt(_arr(x))) if _traced(x) else np.sqrt(_arr(x)))
    tf.exp = _symbolic(lambda x: _unary_traced("Exp", x, np.exp(_arr(x))) if _traced(x) else np.exp(_arr(x)))
    tf.random = types.SimpleNamespace(normal=_normal)
    tf.random_normal_initializer = _Anything
    tf.function = lambda f=None, **k: f if f is not None else (lambda g: g)
    tf.ones_like, tf.zeros_like, tf.reduce_mean, tf.abs = np.ones_like, np.zeros_like, np.mean, np.abs
    tfa_layers.InstanceNormalization = InstanceNormalization
    tfa.layers = tfa_layers
    return tf


def _resize(x, size, method):
    if method != "nearest":
        raise ValueError("only method='nearest' is in the shim (spade.py:17)")
    y = resize_nearest(x, size)
    if _traced(x):
        n = TRACER.fresh("resize")
        TRACER.emit(TRACER.tg.resize_nearest(n, x.tf, [int(size[0]), int(size[1])], half_pixel_centers=True))
        return TT(y, n)
    return y


def _leaky_relu(x, alpha=0.2):
    a = _arr(x)
    y = np.where(a > 0, a, alpha * a)
    if _traced(x):
        n = TRACER.fresh("lrelu")
        TRACER.emit(TRACER.tg.leaky_relu(n, x.tf, alpha))
        return TT(y, n)
    return y


def _moments(x, axes, keepdims=False):
    a = _arr(x)
    m, v = a.mean(axis=tuple(axes), keepdims=keepdims), a.var(axis=tuple(axes), keepdims=keepdims)
    if _traced(x):
        # tf.nn.moments = Mean, SquaredDifference, Mean.  The executor at hand (OpenCV) reduces over the spatial axes only,
        # so traces run with a batch of ONE, where axes (0, 1, 2) and (1, 2) are the same reduction.
        axes = tuple(axes)
        assert keepdims and (axes == (1, 2) or (axes == (0, 1, 2) and a.shape[0] == 1)), "trace moments with batch 1"
        tg, nm, nd, ns, nv = TRACER.tg, TRACER.fresh("mean"), TRACER.fresh("diff"), TRACER.fresh("sq"), TRACER.fresh("var")
        TRACER.emit(tg.mean(nm, x.tf, [1, 2]) + tg.binary(nd, "Sub", x.tf, nm) + tg.binary(ns, "Mul", nd, nd) +
                    tg.mean(nv, ns, [1, 2]))
        return TT(m, nm), TT(v, nv)
    return m, v


def uninstall() -> None:
    for name in [n for n, m in list(sys.modules.items()) if getattr(m, "__msr_shim__", False)]:
        del sys.modules[name]
    for name in [n for n in list(sys.modules) if n == "pix2pix" or n == "spade" or n.startswith("spade.")]:
        del sys.modules[name]
