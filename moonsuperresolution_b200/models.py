"""Generator models behind the reference's plug-in interface.

The reference's engine calls ``self.model(np.array(batch), training=False)`` with a float (B, I, I, 2) NHWC array and
takes the last channel of the (B, I, I, C) result (process_full_tiles.py:338-340).  The classes below keep that
signature -- and the reference's class names / constructor arguments (spade/models/model.py:340-350, 640-650,
pix2pix.py:30-40) -- while the arithmetic runs in libmoonsr.so on the GPU:

  * ``m(x, training=False)``          drop-in path: numpy in, numpy out (H2D + forward + D2H).
  * ``m.forward_device(src, out)``    fast path used by ``DEMSuperResolution``: torch CUDA tensors, no host copies.

There is no CPU implementation here; without a CUDA device or without libmoonsr.so the constructors raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, Optional

import numpy as np

from . import _lib
from . import weights as W

LATENT_DIM = W.LATENT_DIM


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.MoonSRError("no CUDA device: the generator runs only on the GPU (there is no CPU fallback)")
    return torch


class _Generator:
    """Owns one ``msr_generator`` handle (include/moonsr.h)."""
    arch = "spade"

    def __init__(self, image_size: int, batch_size: int, latent_dim: int = LATENT_DIM, precision: str = "bf16",
                 max_groups: int = 1, weights: Optional[Dict[str, np.ndarray]] = None, seed: int = 0,
                 eps_fn: Optional[Callable[[int, int], np.ndarray]] = None):
        if latent_dim != LATENT_DIM:
            raise ValueError("latent_dim is fixed to 256 on this path (process_full_tiles.py:28,48)")
        if precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION)}")
        self.image_size, self.batch_size, self.latent_dim = int(image_size), int(batch_size), latent_dim
        self.precision, self.max_groups = precision, int(max_groups)
        self._handle = C.c_void_p()
        self._finalized = False
        self._calls = 0
        self._eps_fn = eps_fn
        self._eps_rng = np.random.default_rng(seed)
        self._lib = _lib.lib()
        _torch()
        _lib.check(self._lib.msr_generator_create(C.byref(self._handle), _lib.ARCH[self.arch], self.image_size,
                                                  self.batch_size, self.max_groups, _lib.PRECISION[precision]),
                   "msr_generator_create")
        if weights is not None:
            self.set_weights(weights)

    # ---- weights ------------------------------------------------------------------------------------------------
    def compile(self, *args, **kwargs):
        """Keras API compatibility (process_full_tiles.py:29,49): optimisers / losses are training-only."""
        return self

    def set_weights(self, weights: Dict[str, np.ndarray]) -> None:
        """Keras-layout tensors named as in ``weights.model_spec``; uploads and builds the execution plan."""
        if self._finalized:
            raise _lib.MoonSRError("weights already set")
        W.check_weights(self.arch, self.image_size, weights)
        for name in W.model_spec(self.arch, self.image_size):
            a = np.ascontiguousarray(weights[name], dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self._lib.msr_generator_set_weight(self._handle, name.encode(), a.ctypes.data, shape, a.ndim),
                       f"msr_generator_set_weight({name})")
        _lib.check(self._lib.msr_generator_finalize(self._handle), "msr_generator_finalize")
        self._finalized = True

    def load(self, *paths: str) -> None:
        """Reference: ``gaugan.load(path+'generator', path+'discriminator', path+'encoder')`` and, for CNNSpade,
        ``load(path+'generator', path+'encoder')`` (process_full_tiles.py:30,50; spade/models/model.py:607-610,
        822-824) -- Keras SavedModel directories written by ``save`` (:569-605).  When the first and last path are such
        directories their variable bundles are read directly (savedmodel.py, no TensorFlow); the discriminator is never
        used at inference.  Otherwise the tensors come from ``<dirname(paths[0])>/weights.npz`` in Keras layout."""
        from . import savedmodel as SM
        if len(paths) >= 2 and SM.is_saved_model_dir(paths[0]) and SM.is_saved_model_dir(paths[-1]):
            if self.arch == "pix2pix":
                raise ValueError("pix2pix weights are loaded from weights.npz (the reference never saves that model)")
            self.set_weights(SM.load_gaugan_weights(paths[0], paths[-1], self.image_size, self.arch))
            return
        base = os.path.dirname(os.path.normpath(paths[0])) if paths else ""
        npz = os.path.join(base, "weights.npz")
        if not os.path.exists(npz):
            raise ValueError(f"neither SavedModel directories {paths[0]!r} / {paths[-1]!r} nor weight file {npz} exist")
        self.set_weights(W.load_npz(npz))

    def load_npz(self, path: str) -> None:
        self.set_weights(W.load_npz(path))

    # ---- forward --------------------------------------------------------------------------------------------------
    def _draw_eps(self, n: int) -> np.ndarray:
        if self._eps_fn is not None:
            e = np.asarray(self._eps_fn(self._calls, n), dtype=np.float32)
        else:  # the reference's tf.random.normal is unseeded (sampling.py:13); here: a seeded numpy stream
            e = self._eps_rng.standard_normal((n, self.latent_dim), dtype=np.float32)
        if e.shape != (n, self.latent_dim):
            raise ValueError(f"eps must be ({n}, {self.latent_dim})")
        return e

    def forward_device(self, source, out, eps=None, n_groups: int = 1, stream=None, repeat_phase: int = 0) -> None:
        """source (n_groups*B, I, I, 2) f32 CUDA, out (n_groups*B, I, I) f32 CUDA, eps (n_groups*B, 256) f32 CUDA or
        None.  Each group of B samples has its own SPADE batch statistics (spade.py:21).  ``repeat_phase``
        (_lib.REPEAT_FIRST / REPEAT_NEXT): repeated-sample mode -- the first generation of a batch stores the encoder
        outputs and every SPADE layer's gamma | beta, further generations of the same ``source`` reuse them (include/
        moonsr.h, msr_generator_forward_repeat)."""
        if not self._finalized:
            raise _lib.MoonSRError("model has no weights: call set_weights / load first")
        torch = _torch()
        n = n_groups * self.batch_size
        i = self.image_size
        if tuple(source.shape) != (n, i, i, 2) or source.dtype != torch.float32 or not source.is_contiguous():
            raise ValueError(f"source must be a contiguous float32 CUDA tensor of shape {(n, i, i, 2)}")
        if out.numel() != n * i * i or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 CUDA tensor with n*I*I elements")
        if self.arch == "spade":
            if eps is None:
                eps = torch.from_numpy(self._draw_eps(n)).to(source.device, non_blocking=True)
            if tuple(eps.shape) != (n, self.latent_dim) or eps.dtype != torch.float32 or not eps.is_contiguous():
                raise ValueError("eps must be contiguous float32 (n, 256)")
        else:
            eps = None
        self._calls += 1
        _lib.check(self._lib.msr_generator_forward_repeat(self._handle, source.data_ptr(), _lib.ptr(eps), out.data_ptr(),
                                                          n_groups, int(repeat_phase), _lib.stream_ptr(stream)),
                   "msr_generator_forward")

    def __call__(self, x, training=False, eps=None):
        """Reference plug-in signature (process_full_tiles.py:338).  ``x`` (B, I, I, 2), any float dtype (the
        reference's padded batches are float64; Keras casts to float32).  Returns float32
        (B, I, I, 1)."""
        torch = _torch()
        x = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
        b, i = self.batch_size, self.image_size
        if x.shape != (b, i, i, 2):
            raise ValueError(f"model built for batches of shape {(b, i, i, 2)}, got {x.shape} "
                             "(the sampler's noise shape is fixed, sampling.py:13-15)")
        src = torch.from_numpy(x).cuda()
        out = torch.empty((b, i, i), dtype=torch.float32, device=src.device)
        eps_t = None if eps is None else torch.from_numpy(np.ascontiguousarray(eps, dtype=np.float32)).cuda()
        self.forward_device(src, out, eps_t, 1)
        return out.cpu().numpy()[..., None]

    # ---- introspection ---------------------------------------------------------------------------------------------
    def read_activation(self, name: str) -> np.ndarray:
        cnt = C.c_int64(0)
        _lib.check(self._lib.msr_generator_read_activation(self._handle, name.encode(), None, 0, C.byref(cnt)),
                   "msr_generator_read_activation")
        a = np.empty(cnt.value, np.float32)
        _lib.check(self._lib.msr_generator_read_activation(self._handle, name.encode(), a.ctypes.data, a.size,
                                                           C.byref(cnt)), "msr_generator_read_activation")
        return a

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.msr_generator_last_launch_count(self._handle))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.msr_generator_device_bytes(self._handle))

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.msr_generator_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IdentityModel:
    """The reference's default plug-in ``lambda x, training=False: x`` (process_full_tiles.py:139-143) with a device
    entry point, so that the identity round trip -- the reference's own built-in check that the tiling / blending code
    reproduces the input DEM -- runs at full raster sizes without a host round trip per batch.  Channel -1 of the input
    (the normalised DEM) is the prediction; the ``+ 0.5`` of processBatch (:340) is applied by the blend kernel as for
    every device model."""
    arch = "identity"
    max_groups = 64

    def __init__(self, image_size: int, batch_size: int):
        self.image_size, self.batch_size = int(image_size), int(batch_size)
        self.last_launch_count = 1

    def __call__(self, x, training=False):
        return x

    def forward_device(self, source, out, eps=None, n_groups: int = 1, stream=None, repeat_phase: int = 0) -> None:
        out.view(source.shape[0], self.image_size, self.image_size).copy_(source[..., 1])


class GauGAN(_Generator):
    """GauGAN(image_size, batch_size, latent_dim) -- inference ``call`` only: encoder -> Gaussian sampler ->
    SPADE generator (spade/models/model.py:340-350, 564-567)."""
    arch = "spade"


class CNNSpade(_Generator):
    """CNNSpade(image_size, batch_size, latent_dim): encoder -> mean + variance -> SPADE generator
    (spade/models/model.py:640-650, 789-791).  Deterministic."""
    arch = "cnn"


class GauGAN_no_KL(CNNSpade):
    """Same inference graph as CNNSpade (spade/models/model.py:265-267)."""


class Pix2Pix(_Generator):
    """Pix2Pix U-Net generator at training=False (pix2pix.py:64-108); fixed 256x256x2 input (pix2pix.py:7).
    The reference wires it into the engine only through the generic ``model=`` callable."""
    arch = "pix2pix"

    def __init__(self, batch_size: int = 16, precision: str = "fp32", **kw):
        super().__init__(256, batch_size, precision=precision, **kw)

    @property
    def generator(self):
        return self


def load_GAN_model(path: str, image_size: int, batch_size: int, **kw) -> GauGAN:
    """process_full_tiles.py:13-31."""
    assert os.path.exists(path), "The path to the neural-network weight is invalid. Please ensure you gave a valid path."
    gaugan = GauGAN(image_size, batch_size, latent_dim=256, **kw)
    gaugan.compile()
    gaugan.load(path + 'generator', path + 'discriminator', path + 'encoder')
    return gaugan


def load_CNN_model(path: str, image_size: int, batch_size: int, **kw) -> CNNSpade:
    """process_full_tiles.py:33-51."""
    assert os.path.exists(path), "The path to the neural-network weight is invalid. Please ensure you gave a valid path."
    gaugan = CNNSpade(image_size, batch_size, latent_dim=256, **kw)
    gaugan.compile()
    gaugan.load(path + 'generator', path + 'encoder')
    return gaugan
