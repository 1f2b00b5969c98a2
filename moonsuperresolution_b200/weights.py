"""Weight inventory of the three generator families, in Keras layout, plus seeded random initialisation.

Tensor shapes follow the reference's layer constructors:
  * SPADE generator  -- spade/models/networks.py:37-57, blocks.py:9-38, spade.py:5-25
  * encoder          -- spade/models/networks.py:8-34, blocks.py:41-68
  * pix2pix U-Net    -- pix2pix.py:10-28, 64-108
Layouts: Conv2D kernel [kh, kw, cin, cout]; Conv2DTranspose kernel [kh, kw, cout, cin];
Dense kernel [in, out].  Names are this package's own (Keras auto-names such as ``conv2d_17`` are not stable).

"Random-init weights" = one seeded numpy draw from the reference's initialiser distributions (App. B.8): Keras
default glorot_uniform + zero bias for Conv2D/Dense, GlorotNormal (truncated normal) for the encoder convs
(blocks.py:59), N(0, 0.02) for pix2pix (pix2pix.py:66,77,90).  TensorFlow's own RNG stream cannot be reproduced
without TensorFlow; the same ``.npz`` is loaded into the CUDA path and the CPU oracle.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

LATENT_DIM = 256                      # process_full_tiles.py:28,48
SPADE_HIDDEN = 128                    # spade.py:9
RB_FILTERS = (1024, 1024, 1024, 512, 256, 128)     # networks.py:43-53
ENC_FILTERS = (64, 128, 256, 512, 512)             # networks.py:15-30 with encoder_downsample_factor=64 (model.py:373-379)
P2P_DOWN = (64, 128, 256, 512, 512, 512, 512, 512)  # pix2pix.py:10-19
P2P_UP = (512, 512, 512, 512, 256, 128, 64)         # pix2pix.py:20-28

Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str]]"


def spade_generator_spec(image_size: int) -> Spec:
    """name -> (shape, init) for build_generator (networks.py:37-57)."""
    sw = image_size // 64
    spec: Spec = OrderedDict()
    spec["gen.dense.kernel"] = ((LATENT_DIM, 16 * sw * sw * 64), "glorot_uniform")
    spec["gen.dense.bias"] = ((16 * sw * sw * 64,), "zeros")
    cin = 1024
    for k, cout in enumerate(RB_FILTERS, start=1):
        def spade(tag, c):
            pre = f"gen.rb{k}.{tag}"
            spec[pre + ".conv.kernel"] = ((3, 3, 2, SPADE_HIDDEN), "glorot_uniform")
            spec[pre + ".conv.bias"] = ((SPADE_HIDDEN,), "zeros")
            spec[pre + ".conv_gamma.kernel"] = ((3, 3, SPADE_HIDDEN, c), "glorot_uniform")
            spec[pre + ".conv_gamma.bias"] = ((c,), "zeros")
            spec[pre + ".conv_beta.kernel"] = ((3, 3, SPADE_HIDDEN, c), "glorot_uniform")
            spec[pre + ".conv_beta.bias"] = ((c,), "zeros")
        # attribute creation order of ResidualBlock.build (blocks.py:17-26)
        spade("spade_1", cin)
        spade("spade_2", cout)
        spec[f"gen.rb{k}.conv_1.kernel"] = ((3, 3, cin, cout), "glorot_uniform")
        spec[f"gen.rb{k}.conv_1.bias"] = ((cout,), "zeros")
        spec[f"gen.rb{k}.conv_2.kernel"] = ((3, 3, cout, cout), "glorot_uniform")
        spec[f"gen.rb{k}.conv_2.bias"] = ((cout,), "zeros")
        if cin != cout:
            spade("spade_3", cin)
            spec[f"gen.rb{k}.conv_3.kernel"] = ((3, 3, cin, cout), "glorot_uniform")
            spec[f"gen.rb{k}.conv_3.bias"] = ((cout,), "zeros")
        cin = cout
    spec["gen.out.kernel"] = ((4, 4, cin, 1), "glorot_uniform")
    spec["gen.out.bias"] = ((1,), "zeros")
    return spec


def encoder_spec(image_size: int) -> Spec:
    """name -> (shape, init) for build_encoder (networks.py:8-34)."""
    spec: Spec = OrderedDict()
    cin = 2
    for k, cout in enumerate(ENC_FILTERS, start=1):
        spec[f"enc.down{k}.kernel"] = ((3, 3, cin, cout), "glorot_normal")
        if k > 1:   # apply_norm=False on the first block (networks.py:16-17); tfa InstanceNormalization gamma/beta
            spec[f"enc.down{k}.in_gamma"] = ((cout,), "ones")
            spec[f"enc.down{k}.in_beta"] = ((cout,), "zeros")
        cin = cout
    feat = (image_size // 32) ** 2 * cin
    for head in ("mean", "variance"):
        spec[f"enc.{head}.kernel"] = ((feat, LATENT_DIM), "glorot_uniform")
        spec[f"enc.{head}.bias"] = ((LATENT_DIM,), "zeros")
    return spec


def pix2pix_spec() -> Spec:
    """name -> (shape, init) for Pix2Pix.buildGenerator (pix2pix.py:64-108); fixed 256x256x2 input (pix2pix.py:7)."""
    spec: Spec = OrderedDict()

    def bn(pre, c):
        spec[pre + ".bn.gamma"] = ((c,), "ones")
        spec[pre + ".bn.beta"] = ((c,), "zeros")
        spec[pre + ".bn.moving_mean"] = ((c,), "zeros")
        spec[pre + ".bn.moving_variance"] = ((c,), "ones")

    cin = 2
    for k, cout in enumerate(P2P_DOWN, start=1):
        spec[f"p2p.down{k}.kernel"] = ((4, 4, cin, cout), "normal002")
        if k > 1:
            bn(f"p2p.down{k}", cout)
        cin = cout
    skips = list(P2P_DOWN[:-1])[::-1]                       # reversed(skips[:-1]), pix2pix.py:102
    for k, cout in enumerate(P2P_UP, start=1):
        spec[f"p2p.up{k}.kernel"] = ((4, 4, cout, cin), "normal002")      # Conv2DTranspose layout
        bn(f"p2p.up{k}", cout)
        cin = cout + skips[k - 1]                           # Concatenate([x, skip]), pix2pix.py:106
    spec["p2p.last.kernel"] = ((4, 4, 1, cin), "normal002")
    spec["p2p.last.bias"] = ((1,), "zeros")
    return spec


def model_spec(arch: str, image_size: int) -> Spec:
    """arch in {"spade", "cnn", "pix2pix"}: GauGAN / CNNSpade share generator + encoder shapes (model.py:368-379,
    model.py:640+)."""
    if arch in ("spade", "cnn"):
        spec = spade_generator_spec(image_size)
        spec.update(encoder_spec(image_size))
        return spec
    if arch == "pix2pix":
        if image_size != 256:
            raise ValueError("pix2pix is fixed to 256x256 inputs (pix2pix.py:7)")
        return pix2pix_spec()
    raise ValueError(f"unknown arch {arch!r}")


def _fans(shape: Tuple[int, ...], name: str) -> Tuple[int, int]:
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = shape[0] * shape[1]
    # Keras computes fans from the stored layout: [kh, kw, in, out] -> fan_in = rf * shape[-2]
    return rf * shape[2], rf * shape[3]


def _draw(rng: np.random.Generator, shape, init: str, name: str) -> np.ndarray:
    if init == "zeros":
        return np.zeros(shape, np.float32)
    if init == "ones":
        return np.ones(shape, np.float32)
    if init == "normal002":
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)).astype(np.float32)
    fan_in, fan_out = _fans(shape, name)
    if init == "glorot_uniform":
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)
    if init == "glorot_normal":
        std = np.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2.0                      # truncated normal: resample outside 2 sigma
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2.0
        return (x * std).astype(np.float32)
    raise ValueError(init)


def random_init(arch: str, image_size: int, seed: int = 0, perturb_affine: bool = False) -> Dict[str, np.ndarray]:
    """Seeded random-init weights for ``arch`` in Keras layout.

    ``perturb_affine=True`` additionally randomises the tensors Keras initialises to constants (biases, norm gamma /
    beta, BatchNorm moving statistics) so that parity tests exercise them; it is a test aid, not a reference
    initialiser."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = OrderedDict()
    for name, (shape, init) in model_spec(arch, image_size).items():
        w = _draw(rng, shape, init, name)
        if perturb_affine and init in ("zeros", "ones"):
            jitter = rng.standard_normal(shape).astype(np.float32) * np.float32(0.1)
            w = (w + jitter).astype(np.float32)
            if name.endswith("moving_variance"):
                w = np.abs(w).astype(np.float32) + np.float32(0.5)
        out[name] = w
    return out


def param_count(spec: Spec, prefix: str = "") -> int:
    return int(sum(int(np.prod(s)) for n, (s, _) in spec.items() if n.startswith(prefix)))


def save_npz(path: str, weights: Dict[str, np.ndarray]) -> None:
    np.savez(path, **weights)


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return OrderedDict((k, np.ascontiguousarray(z[k], dtype=np.float32)) for k in z.files)


def check_weights(arch: str, image_size: int, weights: Dict[str, np.ndarray]) -> None:
    """Raises ValueError when a tensor is missing or mis-shaped."""
    for name, (shape, _) in model_spec(arch, image_size).items():
        if name not in weights:
            raise ValueError(f"missing weight tensor {name}")
        if tuple(weights[name].shape) != tuple(shape):
            raise ValueError(f"weight {name}: shape {tuple(weights[name].shape)} != {tuple(shape)}")
