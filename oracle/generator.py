"""torch-CPU restatement of the reference generators (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

PARITY: PINNED to the reference's own generator code executed by a third-party TensorFlow-graph executor; not pinned to
TensorFlow's own kernels.  The arithmetic of the reference lives in TensorFlow 2.5 / Keras / tensorflow-addons 0.16.1
(pip-env.py:15,34,36), none of which is installable here, and the reference ships no golden vectors.  What holds this
file in place:
  * tests/test_tf_semantics_opencv.py -- the bodies of GauGAN.call / CNNSpade.call (cut out of model.py) and the UNMODIFIED
    networks.py, blocks.py, spade.py, sampling.py, pix2pix.py are run on traced tensors, every op they perform is written
    down as a TensorFlow GraphDef node, and the graph is EXECUTED BY OpenCV's TensorFlow importer
    (cv2.dnn.readNetFromTensorflow), an implementation of TensorFlow's graph semantics written by neither the reference's
    nor this repository's authors.  Whole GauGAN-64 / CNNSpade-64 calls (batch of one) and the whole pix2pix-256 generator
    agree with this file to 6e-6 (tests/golden/generator_opencv_tf.npz, tests/golden/make_golden_tf.py --opencv); so do the
    single ops with silent-mismatch potential (SAME padding k3 / k4, stride 1 / 2, even / odd sizes; Conv2DBackpropInput;
    half-pixel nearest resize; FusedBatchNorm inference; LeakyRelu);
  * tests/test_oracle_generator_pinned.py -- the same reference modules executed on a numpy op shim
    (tests/golden/tf_numpy_shim.py) with batches of 2-3 (tf.nn.moments over the batch axis, zero padding slots) and at
    I = 128 reproduce this file to float32 rounding; the shim's ops are checked against scalar-loop restatements;
  * tests/golden/make_golden_tf.py (no flag) writes generator_tf.npz on any TensorFlow-equipped box; the same test file
    then holds this oracle to TensorFlow itself (skipped while that file is absent).
Also pinned: parameter counts / shapes (App. A), fp32-vs-fp64 self agreement, structural invariants
(tests/test_oracle_generator.py).  Each function cites the reference lines it follows.

Weights arrive in Keras layout (moonsuperresolution_b200/weights.py); activations are NHWC at the interface and NCHW
internally (torch).  ``dtype`` selects float32 (the reference's precision) or float64 (error-free yardstick).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

W = Dict[str, np.ndarray]


def _t(a: np.ndarray, dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _same_pad(x: torch.Tensor, k: int, s: int) -> torch.Tensor:
    """TF 'SAME' padding (App. B.2): total = max((ceil(in/s)-1)*s + k - in, 0), before = total // 2, rest after."""
    h, w = x.shape[-2:]
    th = max((-(-h // s) - 1) * s + k - h, 0)
    tw = max((-(-w // s) - 1) * s + k - w, 0)
    return F.pad(x, (tw // 2, tw - tw // 2, th // 2, th - th // 2))


def conv2d_same(x: torch.Tensor, kernel: np.ndarray, bias: Optional[np.ndarray], stride: int = 1) -> torch.Tensor:
    """Keras Conv2D(padding='same'); kernel [kh, kw, cin, cout] (App. B.1)."""
    wt = _t(kernel, x.dtype).permute(3, 2, 0, 1)
    b = None if bias is None else _t(bias, x.dtype)
    return F.conv2d(_same_pad(x, kernel.shape[0], stride), wt, b, stride=stride)


def conv2d_transpose_same_s2(x: torch.Tensor, kernel: np.ndarray, bias: Optional[np.ndarray]) -> torch.Tensor:
    """Keras Conv2DTranspose(k=4, strides=2, padding='same'); kernel [kh, kw, cout, cin]; equals torch
    ConvTranspose2d(k=4, s=2, padding=1) (App. B.2)."""
    assert kernel.shape[0] == 4 and kernel.shape[1] == 4
    wt = _t(kernel, x.dtype).permute(3, 2, 0, 1)          # torch wants [cin, cout, kh, kw]
    b = None if bias is None else _t(bias, x.dtype)
    return F.conv_transpose2d(x, wt, b, stride=2, padding=1)


def resize_nearest_tf(mask: torch.Tensor, size: int) -> torch.Tensor:
    """tf.image.resize(method='nearest') with half-pixel centres (spade.py:17, App. B.3):
    src = min(floor((dst + 0.5) * in / out), in - 1)."""
    n_in = mask.shape[-1]
    idx = torch.clamp(torch.floor((torch.arange(size, dtype=torch.float64) + 0.5) * (n_in / size)).long(), max=n_in - 1)
    return mask[:, :, idx][:, :, :, idx]


def spade(x: torch.Tensor, mask: torch.Tensor, w: W, pre: str, eps: float = 1e-5) -> torch.Tensor:
    """SPADE.call (spade.py:16-25): batch moments over (N, H, W), biased variance, gamma * xhat + beta."""
    m = resize_nearest_tf(mask, x.shape[-1])
    a = torch.relu(conv2d_same(m, w[pre + ".conv.kernel"], w[pre + ".conv.bias"]))
    gamma = conv2d_same(a, w[pre + ".conv_gamma.kernel"], w[pre + ".conv_gamma.bias"])
    beta = conv2d_same(a, w[pre + ".conv_beta.kernel"], w[pre + ".conv_beta.bias"])
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    return gamma * ((x - mean) / torch.sqrt(var + eps)) + beta


def residual_block(x: torch.Tensor, mask: torch.Tensor, w: W, pre: str, alpha: float = 0.2) -> torch.Tensor:
    """ResidualBlock.call (blocks.py:28-38); learned 3x3 skip iff the channel count changes (blocks.py:23-26)."""
    h = spade(x, mask, w, pre + ".spade_1")
    h = conv2d_same(F.leaky_relu(h, alpha), w[pre + ".conv_1.kernel"], w[pre + ".conv_1.bias"])
    h = spade(h, mask, w, pre + ".spade_2")
    h = conv2d_same(F.leaky_relu(h, alpha), w[pre + ".conv_2.kernel"], w[pre + ".conv_2.bias"])
    if (pre + ".conv_3.kernel") in w:
        s = spade(x, mask, w, pre + ".spade_3")
        skip = conv2d_same(F.leaky_relu(s, alpha), w[pre + ".conv_3.kernel"], w[pre + ".conv_3.bias"])
    else:
        skip = x
    return skip + h


def spade_generator(latent: torch.Tensor, source_nchw: torch.Tensor, w: W) -> torch.Tensor:
    """build_generator graph (networks.py:37-57): Dense -> Reshape(sw, sw, 1024) -> 6 x (ResidualBlock, x2 nearest
    upsample) -> leaky_relu(0.2) -> Conv2D(1, 4, 'same').  No tanh."""
    n, _, size, _ = source_nchw.shape
    sw = size // 64
    x = latent @ _t(w["gen.dense.kernel"], latent.dtype) + _t(w["gen.dense.bias"], latent.dtype)
    x = x.reshape(n, sw, sw, 1024).permute(0, 3, 1, 2)     # Reshape is row-major over (h, w, c) (App. B.1)
    for k in range(1, 7):
        x = residual_block(x, source_nchw, w, f"gen.rb{k}")
        x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)       # UpSampling2D((2,2)) nearest (App. B.4)
    x = F.leaky_relu(x, 0.2)
    return conv2d_same(x, w["gen.out.kernel"], w["gen.out.bias"])


def encoder(source_nchw: torch.Tensor, w: W, alpha: float = 0.2, in_eps: float = 1e-3):
    """build_encoder (networks.py:8-34) with downsample_block (blocks.py:41-68): Conv3x3 s2 SAME no bias ->
    tfa InstanceNormalization (eps 1e-3, per sample & channel; skipped on block 1) -> LeakyReLU(0.2); Flatten in
    NHWC order; Dense mean, Dense variance."""
    x = source_nchw
    for k in range(1, 6):
        x = conv2d_same(x, w[f"enc.down{k}.kernel"], None, stride=2)
        if k > 1:
            mu = x.mean(dim=(2, 3), keepdim=True)
            var = ((x - mu) ** 2).mean(dim=(2, 3), keepdim=True)
            g = _t(w[f"enc.down{k}.in_gamma"], x.dtype).view(1, -1, 1, 1)
            b = _t(w[f"enc.down{k}.in_beta"], x.dtype).view(1, -1, 1, 1)
            x = (x - mu) / torch.sqrt(var + in_eps) * g + b
        x = F.leaky_relu(x, alpha)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    mean = flat @ _t(w["enc.mean.kernel"], x.dtype) + _t(w["enc.mean.bias"], x.dtype)
    var = flat @ _t(w["enc.variance.kernel"], x.dtype) + _t(w["enc.variance.bias"], x.dtype)
    return mean, var


def gaugan_call(source_nhwc: np.ndarray, w: W, eps: Optional[np.ndarray], arch: str = "spade",
                dtype=torch.float32, return_latent: bool = False) -> np.ndarray:
    """GauGAN.call (model.py:564-567): encoder -> sampler ``mean + exp(0.5 * variance) * eps`` (sampling.py:11-17)
    -> generator;  CNNSpade.call (model.py:789-791): ``latent = mean + variance``, no sampler.
    ``eps`` (B, 256) replaces the reference's unseeded tf.random.normal (App. B.11).  Returns (B, I, I, 1)."""
    with torch.no_grad():
        src = _t(np.asarray(source_nhwc, dtype=np.float32), dtype).permute(0, 3, 1, 2).contiguous()
        mean, var = encoder(src, w)
        if arch == "spade":
            latent = mean + torch.exp(0.5 * var) * _t(np.asarray(eps, dtype=np.float32), dtype)
        elif arch == "cnn":
            latent = mean + var
        else:
            raise ValueError(arch)
        y = spade_generator(latent, src, w)
        out = y.permute(0, 2, 3, 1).contiguous().numpy()
    if return_latent:
        return out, latent.numpy()
    return out


def pix2pix_call(source_nhwc: np.ndarray, w: W, dtype=torch.float32, bn_eps: float = 1e-3) -> np.ndarray:
    """Pix2Pix.buildGenerator graph at training=False (pix2pix.py:64-108): 8 x [Conv4x4 s2 SAME no bias, BN (not on
    the first), LeakyReLU(0.3)], 7 x [ConvT4x4 s2 SAME no bias, BN, ReLU] each followed by Concatenate([up, skip]),
    ConvT4x4 s2 (bias) + tanh.  BatchNorm uses moving statistics, eps 1e-3; Dropout inactive (App. B.7)."""
    def bn(x, pre):
        g = _t(w[pre + ".bn.gamma"], x.dtype).view(1, -1, 1, 1)
        b = _t(w[pre + ".bn.beta"], x.dtype).view(1, -1, 1, 1)
        mu = _t(w[pre + ".bn.moving_mean"], x.dtype).view(1, -1, 1, 1)
        var = _t(w[pre + ".bn.moving_variance"], x.dtype).view(1, -1, 1, 1)
        return (x - mu) / torch.sqrt(var + bn_eps) * g + b

    with torch.no_grad():
        x = _t(np.asarray(source_nhwc, dtype=np.float32), dtype).permute(0, 3, 1, 2).contiguous()
        skips = []
        for k in range(1, 9):
            x = conv2d_same(x, w[f"p2p.down{k}.kernel"], None, stride=2)
            if k > 1:
                x = bn(x, f"p2p.down{k}")
            x = F.leaky_relu(x, 0.3)
            skips.append(x)
        for k, skip in zip(range(1, 8), reversed(skips[:-1])):
            x = conv2d_transpose_same_s2(x, w[f"p2p.up{k}.kernel"], None)
            x = torch.relu(bn(x, f"p2p.up{k}"))
            x = torch.cat([x, skip], dim=1)
        x = torch.tanh(conv2d_transpose_same_s2(x, w["p2p.last.kernel"], w["p2p.last.bias"]))
        return x.permute(0, 2, 3, 1).contiguous().numpy()


class OracleModel:
    """Callable with the reference's plug-in signature ``m(x, training=False)`` (process_full_tiles.py:338)."""

    def __init__(self, arch: str, weights: W, eps_fn=None, dtype=torch.float32):
        self.arch, self.w, self.eps_fn, self.dtype = arch, weights, eps_fn, dtype
        self.calls = 0

    def __call__(self, x, training=False):
        x = np.asarray(x, dtype=np.float32)
        self.calls += 1
        if self.arch == "pix2pix":
            return pix2pix_call(x, self.w, self.dtype).astype(np.float32)
        eps = None
        if self.arch == "spade":
            eps = self.eps_fn(self.calls - 1, x.shape[0]) if self.eps_fn else np.zeros((x.shape[0], 256), np.float32)
        return gaugan_call(x, self.w, eps, self.arch, self.dtype).astype(np.float32)
