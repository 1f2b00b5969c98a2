"""Cheap deterministic stand-ins for the generator, shared by the golden generator and the tests.

They obey the reference's model plug-in contract (process_full_tiles.py:338-340): ``m(x, training=False)`` with
``x`` (B, I, I, 2) -> array (B, I, I, C); the last channel is the normalised SR DEM."""
import numpy as np


def identity(x, training=False):
    """The reference's default model (process_full_tiles.py:143)."""
    return x


def wobble(x, training=False):
    """Slot- and pixel-dependent perturbation of the normalised DEM so overlapping generations disagree
    (non-zero std) and batch position matters.  Pure float32 numpy, elementwise, no transcendental functions
    (bit-reproducible everywhere)."""
    x = np.asarray(x, dtype=np.float32)
    b = x.shape[0]
    slot = (np.arange(b, dtype=np.float32) - np.float32(0.5 * (b - 1))) / np.float32(max(b, 1))
    slot = slot.reshape(b, 1, 1)
    ortho, dem = x[..., 0], x[..., 1]
    y = dem * np.float32(0.875) + ortho * ortho * np.float32(0.0625) + slot * np.float32(0.03125)
    return y[..., None].astype(np.float32)


def ripple(x, training=False):
    """Per-sample stand-in (no dependence on the slot or on the rest of the batch, float32 out whatever the input dtype):
    the model class for which dedup mode must equal the reference's tile-by-tile result bit for bit."""
    x = np.asarray(x, dtype=np.float32)
    ortho, dem = x[..., 0], x[..., 1]
    y = dem * np.float32(0.875) + ortho * ortho * np.float32(0.0625) + ortho * dem * np.float32(0.03125)
    return y[..., None].astype(np.float32)


class Flicker:
    """Stateful stand-in for a stochastic generator: per-sample like ``ripple``, plus an offset that changes with every
    call (what new sampler noise does), so repeated generations of the same batch disagree."""

    def __init__(self):
        self.calls = 0

    def __call__(self, x, training=False):
        y = ripple(x)
        k = np.float32((self.calls % 7) - 3) * np.float32(0.015625)
        self.calls += 1
        return (y + k).astype(np.float32)
