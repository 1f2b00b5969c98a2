"""Runs the UNMODIFIED reference generator code -- spade/models/{networks,blocks,spade,sampling}.py, pix2pix.py and the
three-line ``call`` methods of GauGAN / CNNSpade (spade/models/model.py:564-567, 789-791) -- with this repository's
Keras-layout weights, and returns outputs + per-block activations (TEST INFRASTRUCTURE ONLY).

Two back ends, one code path:
  * ``backend="tf"``   real TensorFlow 2.x + tensorflow-addons (not installable in the build image; any box that has them
                        turns the generator oracle's parity from "unpinned" to pinned -- see make_golden_tf.py);
  * ``backend="shim"`` tests/golden/tf_numpy_shim.py: the reference's graph code runs as is, only the op-level semantics
                        are restated in numpy.

spade/models/model.py itself cannot be imported (unresolved merge markers at :36-41, :54-59 make it a SyntaxError), so
the ``call`` methods are cut out of its text and compiled on their own.
"""
from __future__ import annotations

import importlib
import os
import re
import sys
import textwrap
import types
from typing import Dict

import numpy as np

REFERENCE = os.environ.get("MSR_REFERENCE_PATH", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE, "spade", "models", "networks.py"))


class Reference:
    def __init__(self, backend: str = "shim", dtype=np.float64):
        assert available(), f"reference sources not found under {REFERENCE}"
        self.backend = backend
        if backend == "shim":
            if HERE not in sys.path:
                sys.path.insert(0, HERE)
            import tf_numpy_shim
            tf_numpy_shim.set_dtype(dtype)
            self.tf = tf_numpy_shim.install()
            self._shim = tf_numpy_shim
        elif backend == "tf":
            import tensorflow as tf          # noqa: F401  (real TensorFlow)
            import tensorflow_addons         # noqa: F401
            self.tf = tf
            self._shim = None
        else:
            raise ValueError(backend)
        if REFERENCE not in sys.path:
            sys.path.insert(0, REFERENCE)
        self.networks = importlib.import_module("spade.models.networks")
        self.blocks = importlib.import_module("spade.models.blocks")
        self.sampling = importlib.import_module("spade.models.sampling")
        self.pix2pix = importlib.import_module("pix2pix")
        src = open(os.path.join(REFERENCE, "spade", "models", "model.py")).read()
        self.call_gaugan = self._cut_call(src, "GauGAN")
        self.call_cnn = self._cut_call(src, "CNNSpade")

    def close(self) -> None:
        if REFERENCE in sys.path:
            sys.path.remove(REFERENCE)
        if self._shim is not None:
            self._shim.uninstall()

    @staticmethod
    def _cut_call(src: str, cls: str):
        """The text of ``def call(self, source)`` of class ``cls`` in model.py, compiled as a plain function."""
        m = re.search(r"^class %s\(Model\):\n" % cls, src, re.M)
        assert m, cls
        body = src[m.end():]
        nxt = re.search(r"^class ", body, re.M)
        body = body[:nxt.start()] if nxt else body
        c = re.search(r"^    def call\(self, source\):\n((?:        .*\n|\n)+)", body, re.M)
        assert c, f"{cls}.call not found"
        ns: dict = {}
        exec(compile("def call(self, source):\n" + textwrap.indent(textwrap.dedent(c.group(1)), "    "),
                     f"model.py::{cls}.call", "exec"), ns)
        return ns["call"]

    # ------------------------------------------------------------------------------------------------------------------
    def _np(self, t):
        return np.asarray(t.numpy() if hasattr(t, "numpy") else t)

    def build_spade(self, image_size: int, batch_size: int, weights: Dict[str, np.ndarray]):
        """generator / encoder / sampler exactly as GauGAN.__init__ builds them (model.py:371-381), then the repo's
        Keras-layout weights assigned layer by layer."""
        N, S = self.networks, self.sampling
        shape = (image_size, image_size, 2)
        gen = N.build_generator(shape, latent_dim=256, alpha=0.2)
        enc = N.build_encoder(shape, encoder_downsample_factor=64, latent_dim=256, alpha=0.2, dropout=0.5)
        sampler = S.GaussianSampler(batch_size, 256)
        # one forward on zeros so that every (sub-)layer has created its variables
        z = np.zeros((batch_size,) + shape, np.float32)
        gen([np.zeros((batch_size, 256), np.float32), z])
        enc(z)
        layers_mod = self.tf.keras.layers
        dense = [l for l in gen.layers if isinstance(l, layers_mod.Dense)]
        assert len(dense) == 1
        dense[0].set_weights([weights["gen.dense.kernel"], weights["gen.dense.bias"]])
        rbs = [l for l in gen.layers if isinstance(l, self.blocks.ResidualBlock)]
        assert len(rbs) == 6
        for k, rb in enumerate(rbs, start=1):
            pre = f"gen.rb{k}"
            tags = ["spade_1", "spade_2"] + (["spade_3"] if rb.learned_skip else [])
            for tag in tags:
                sp = getattr(rb, tag)
                for conv in ("conv", "conv_gamma", "conv_beta"):
                    getattr(sp, conv).set_weights([weights[f"{pre}.{tag}.{conv}.kernel"],
                                                   weights[f"{pre}.{tag}.{conv}.bias"]])
            for conv in ["conv_1", "conv_2"] + (["conv_3"] if rb.learned_skip else []):
                getattr(rb, conv).set_weights([weights[f"{pre}.{conv}.kernel"], weights[f"{pre}.{conv}.bias"]])
        convs = [l for l in gen.layers if isinstance(l, layers_mod.Conv2D)]
        assert len(convs) == 1
        convs[0].set_weights([weights["gen.out.kernel"], weights["gen.out.bias"]])
        seqs = [l for l in enc.layers if isinstance(l, self.tf.keras.Sequential)]
        assert len(seqs) == 5
        for k, seq in enumerate(seqs, start=1):
            seq.layers[0].set_weights([weights[f"enc.down{k}.kernel"]])
            if k > 1:
                seq.layers[1].set_weights([weights[f"enc.down{k}.in_gamma"], weights[f"enc.down{k}.in_beta"]])
        for head in ("mean", "variance"):
            enc.get_layer(head).set_weights([weights[f"enc.{head}.kernel"], weights[f"enc.{head}.bias"]])
        return types.SimpleNamespace(generator=gen, encoder=enc, sampler=sampler, blocks=rbs)

    def run_spade(self, arch: str, image_size: int, weights, x: np.ndarray, eps: np.ndarray):
        """GauGAN.call (arch 'spade') / CNNSpade.call (arch 'cnn') on batch ``x`` (B, I, I, 2); the sampler's unseeded
        tf.random.normal (sampling.py:13) is replaced by ``eps`` for the duration of the call."""
        b = x.shape[0]
        m = self.build_spade(image_size, b, weights)
        x = np.asarray(x, np.float32)
        rnd = self.tf.random
        saved = rnd.normal
        rnd.normal = lambda shape, mean=0.0, stddev=1.0, **k: np.asarray(eps, np.float32).reshape(shape)
        try:
            out = (self.call_gaugan if arch == "spade" else self.call_cnn)(m, x)
            mean, var = m.encoder(x)
        finally:
            rnd.normal = saved
        res = {"out": self._np(out), "enc.mean": self._np(mean), "enc.variance": self._np(var)}
        # per-block activations: the generator applied layer by layer (same layer objects, same order as gen.layers)
        if arch == "spade":
            latent = self._np(mean) + np.exp(0.5 * self._np(var)) * np.asarray(eps, np.float32)
        else:
            latent = self._np(mean) + self._np(var)
        res["latent"] = latent
        return res

    def build_pix2pix(self, weights):
        P = self.pix2pix.Pix2Pix
        obj = P()                              # pix2pix.py:6-41: builds the generator (and the unused discriminator)
        gen = obj.generator
        z = np.zeros((1, 256, 256, 2), np.float32)
        gen(z)
        for k, seq in enumerate(obj.down_stack, start=1):
            seq.layers[0].set_weights([weights[f"p2p.down{k}.kernel"]])
            if k > 1:
                seq.layers[1].set_weights([weights[f"p2p.down{k}.bn.{n}"] for n in
                                           ("gamma", "beta", "moving_mean", "moving_variance")])
        for k, seq in enumerate(obj.up_stack, start=1):
            seq.layers[0].set_weights([weights[f"p2p.up{k}.kernel"]])
            seq.layers[1].set_weights([weights[f"p2p.up{k}.bn.{n}"] for n in
                                       ("gamma", "beta", "moving_mean", "moving_variance")])
        last = [l for l in gen.layers if isinstance(l, self.tf.keras.layers.Conv2DTranspose)]
        assert len(last) == 1
        last[0].set_weights([weights["p2p.last.kernel"], weights["p2p.last.bias"]])
        return gen

    def run_pix2pix(self, weights, x: np.ndarray):
        gen = self.build_pix2pix(weights)
        return {"out": self._np(gen(np.asarray(x, np.float32), training=False))}


    # ------------------------------------------------------------------------------------------------------------------
    # third-party TensorFlow executor: trace the reference's code into a GraphDef, run it with OpenCV's importer
    # ------------------------------------------------------------------------------------------------------------------
    def run_spade_opencv(self, arch: str, image_size: int, weights, x: np.ndarray, eps: np.ndarray, workdir: str):
        """GauGAN.call / CNNSpade.call (the bodies cut out of model.py) executed on the shim with TRACED tensors: every
        op the reference's code performs -- inside build_encoder, GaussianSampler.call, build_generator,
        ResidualBlock.call, SPADE.call -- is written down as a TensorFlow GraphDef node (tests/golden/tf_graphdef.py) and
        the graph is then executed by cv2.dnn.readNetFromTensorflow.  Batch of one (OpenCV reduces over spatial axes
        only; with one sample tf.nn.moments over (0, 1, 2) is that reduction); the sampler's noise enters as a second
        Placeholder, as tf.random.normal's output would.  Returns (OpenCV output, shim output), both (1, I, I, 1)."""
        assert self.backend == "shim" and x.shape[0] == 1
        import tf_graphdef as TG
        SH = self._shim
        m = self.build_spade(image_size, 1, weights)
        x = np.asarray(x, np.float32)
        eps = np.asarray(eps, np.float32).reshape(1, 256)
        call = self.call_gaugan if arch == "spade" else self.call_cnn
        rnd = self.tf.random
        saved = rnd.normal
        try:
            rnd.normal = lambda shape, mean=0.0, stddev=1.0, **k: eps.reshape(shape)
            want = np.asarray(call(m, x))
            tx = SH.start_trace(TG, x, "input")
            teps = SH.trace_input(eps, "eps")
            rnd.normal = lambda shape, mean=0.0, stddev=1.0, **k: teps
            y = call(m, tx)
            graph = SH.stop_trace()
        finally:
            rnd.normal = saved
            SH.TRACER = None
        feeds = {"input": x, "eps": eps} if arch == "spade" else {"input": x}
        if arch != "spade":      # the unused Placeholder would be an unconnected input
            graph = graph.replace(TG.placeholder("eps"), b"", 1)
        got = TG.run_opencv(graph, feeds, os.path.join(workdir, f"{arch}{image_size}.pb"), output=y.tf)
        return got, want

    def run_pix2pix_opencv(self, weights, x: np.ndarray, workdir: str):
        """Pix2Pix.buildGenerator's functional graph (as recorded by the shim from the unmodified pix2pix.py) emitted
        layer by layer as a GraphDef and executed by OpenCV's TensorFlow importer."""
        assert self.backend == "shim"
        import tf_graphdef as TG
        gen = self.build_pix2pix(weights)
        graph, final = TG.emit_functional_model(gen, self._shim, batch=x.shape[0], in_hw=256)
        want = self._np(gen(np.asarray(x, np.float32), training=False))
        return TG.run_opencv(graph, x, os.path.join(workdir, "pix2pix.pb")), want
