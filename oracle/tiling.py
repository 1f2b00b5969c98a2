"""numpy restatement of the reference's tiling / blending arithmetic (TEST INFRASTRUCTURE ONLY).

Every function cites the lines of ``/root/reference/process_full_tiles.py`` it restates.  The restatement is
written as pure functions over arrays (the reference is one stateful class) and spells out every dtype
promotion numpy performs implicitly in the reference, because the CUDA kernels reproduce exactly those
roundings.  PINNED against the reference's own class: ``tests/golden/make_golden.py`` runs the unmodified
reference (imported with stub modules for GDAL / TensorFlow) and ``tests/test_oracle_tiling.py`` compares.

Coordinates: ``(x, y)`` = (column, row); "canvas" = the no_value-padded raster of ``padInputs``.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

CANVAS_QUANTUM = 1024  # literal 1024 in process_full_tiles.py:252-253 (NOT tile_size)


@dataclasses.dataclass(frozen=True)
class Geometry:
    """Integer facts of one run; restates the arithmetic of padInputs (process_full_tiles.py:252-256)."""
    height: int
    width: int
    image_size: int
    stride: int
    tile_size: int

    @property
    def off(self) -> int:                      # I - S, the halo (process_full_tiles.py:252)
        return self.image_size - self.stride

    @property
    def purge(self) -> int:                    # process_full_tiles.py:392
        return self.image_size // 16

    @property
    def canvas_h(self) -> int:                 # process_full_tiles.py:253
        return (self.height // CANVAS_QUANTUM + 1) * CANVAS_QUANTUM + 2 * self.off

    @property
    def canvas_w(self) -> int:                 # process_full_tiles.py:252
        return (self.width // CANVAS_QUANTUM + 1) * CANVAS_QUANTUM + 2 * self.off

    @property
    def pad_y(self) -> int:                    # process_full_tiles.py:256
        return self.canvas_h - self.height - self.off

    @property
    def pad_x(self) -> int:                    # process_full_tiles.py:255
        return self.canvas_w - self.width - self.off

    @property
    def acc_side(self) -> int:                 # process_full_tiles.py:386
        return self.tile_size + 2 * self.image_size - 2 * self.stride


def pad_inputs(dem: np.ndarray, img: np.ndarray, geo: Geometry, no_value: float) -> Tuple[np.ndarray, np.ndarray]:
    """process_full_tiles.py:246-267 -- no_value canvases with the data placed at offset ``off``."""
    dem_c = np.full((geo.canvas_h, geo.canvas_w), np.float32(no_value), dtype=np.float32)
    img_c = np.full((geo.canvas_h, geo.canvas_w), np.float32(no_value), dtype=np.float32)
    o = geo.off
    dem_c[o:o + geo.height, o:o + geo.width] = dem
    img_c[o:o + geo.height, o:o + geo.width] = img
    return dem_c, img_c


def tile_list(geo: Geometry) -> List[Tuple[int, int]]:
    """process_full_tiles.py:313-325 -- (xx, yy) origins over the un-padded dims, x fastest."""
    return [(xx, yy) for yy in range(0, geo.height, geo.tile_size) for xx in range(0, geo.width, geo.tile_size)]


def patch_origins(geo: Geometry, px: int, py: int) -> Iterable[Tuple[int, int]]:
    """process_full_tiles.py:453-454 -- canvas-coordinate patch origins of tile (px, py), y outer / x inner."""
    span = geo.tile_size + geo.image_size - geo.stride
    for yy in range(py, py + span, geo.stride):
        for xx in range(px, px + span, geo.stride):
            yield xx, yy


def patch_is_valid(dem_c: np.ndarray, img_c: np.ndarray, x: int, y: int, size: int, no_value: float) -> bool:
    """process_full_tiles.py:286-292 -- a patch is used iff no pixel of either raster is <= no_value.

    numpy slicing clips at the canvas edge, so a window hanging over the edge is judged on its in-canvas part
    (and then crashes later in the reference); the valid parameter domain never produces one."""
    a = img_c[y:y + size, x:x + size]
    b = dem_c[y:y + size, x:x + size]
    return not ((a <= no_value).any() or (b <= no_value).any())


def normalize_patch(img_p: np.ndarray, dem_p: np.ndarray) -> Tuple[np.ndarray, Tuple[np.float32, np.float32]]:
    """process_full_tiles.py:295-311 -- per-patch min/max affine to [-0.5, 0.5], all float32.

    ``(x - min) / (max - min) - 0.5`` with float32 subtract, float32 true division, float32 subtract; channel 0 =
    ortho image, channel 1 = DEM; returns the DEM's (min, max) as np.float32 scalars."""
    f32 = np.float32
    i_lo, i_hi = f32(img_p.min()), f32(img_p.max())
    d_lo, d_hi = f32(dem_p.min()), f32(dem_p.max())
    with np.errstate(divide="ignore", invalid="ignore"):
        i_n = ((img_p.astype(f32) - i_lo) / f32(i_hi - i_lo)) - f32(0.5)
        d_n = ((dem_p.astype(f32) - d_lo) / f32(d_hi - d_lo)) - f32(0.5)
    return np.stack([i_n, d_n], axis=-1).astype(f32), (d_lo, d_hi)


def blend_weights(image_size: int) -> np.ndarray:
    """process_full_tiles.py:347-361 + 391-393 -- the float64 blend table, purge-cropped, +1e-7.

    A unit-height 2-D Gaussian (sigma = I/5) sampled on linspace(-I/2, I/2, I)^2, min-max normalised to [0, 1]
    (which makes it non-separable), then ``+ 1e-7`` and cropped by ``purge = I // 16`` on every side."""
    i = image_size
    s = i / 5
    ax = np.linspace(-i / 2, i / 2, i)
    gx, gy = np.meshgrid(ax, ax)
    g = 1.0 / (2.0 * np.pi * s * s) * np.exp(-((gx - 0) ** 2.0 / (2.0 * s ** 2.0) + (gy - 0) ** 2.0 / (2.0 * s ** 2.0)))
    g = (g - g.min()) / (g.max() - g.min())
    g = g + 1e-7
    p = i // 16
    return np.ascontiguousarray(g[p:-p, p:-p])


def denormalize(pred01: np.ndarray, lo: np.float32, hi: np.float32) -> np.ndarray:
    """process_full_tiles.py:396 -- ``pred * (max - min) + min``; ``min``/``max`` are np.float32 scalars, so a float32
    prediction gives two separately rounded float32 ops, a float64 prediction (see ``process_tile``) stays float64."""
    f32 = np.float32
    return pred01 * f32(f32(hi) - f32(lo)) + f32(lo)


def accumulate_patch(w_sum: np.ndarray, mean: np.ndarray, s_acc: np.ndarray, pred: np.ndarray, lohi, kx: int, ky: int,
                     w: np.ndarray, i: int, p: int) -> None:
    """One iteration of rebuildTile's loop (process_full_tiles.py:395-402) on float32 accumulators of any extent:
    the patch whose origin is (kx, ky) in accumulator coordinates updates rows [ky+p, ky+I-p) x cols [kx+p, kx+I-p)."""
    f32, f64 = np.float32, np.float64
    lo, hi = lohi
    d = denormalize(pred, lo, hi)[p:i - p, p:i - p]                                      # float32 (or float64)
    ys, xs = slice(ky + p, ky + i - p), slice(kx + p, kx + i - p)
    w_sum[ys, xs] = (w_sum[ys, xs].astype(f64) + w).astype(f32)                          # :398
    m_old = mean[ys, xs].copy()
    delta_old = d - m_old                                                                # dtype of d
    m_new = (m_old.astype(f64) + (w / w_sum[ys, xs].astype(f64)) * delta_old.astype(f64)).astype(f32)  # :401
    mean[ys, xs] = m_new
    # :400 binds ``mean_old`` to a *view* of ``mean``; after the store of :401 that view already shows the
    # new mean, so :402 really evaluates  w * (d - mean_new) * (d - mean_new)  -- the textbook
    # (d - mean_old)(d - mean_new) product never happens.  Reproduced because parity is against the code.
    delta_new = d - m_new                                                                # dtype of d
    s_acc[ys, xs] = (s_acc[ys, xs].astype(f64) + (w * delta_new.astype(f64)) * delta_new.astype(f64)).astype(f32)  # :402


def rebuild_tile(generated: Dict[Tuple[int, int], np.ndarray],
                 minmax: Dict[Tuple[int, int], Tuple[np.float32, np.float32]],
                 geo: Geometry, no_value: float,
                 return_accumulators: bool = False):
    """process_full_tiles.py:363-414 -- Gaussian-weighted incremental (West) mean / variance over patches.

    ``generated[(x, y)]`` is the (I, I) model output **after** the ``+ 0.5`` of processBatch (float32 from a real
    network; float64 when a numpy callable saw the float64 padded batch), keyed by the patch origin relative to the
    tile's canvas origin; iteration follows dict insertion order.  Accumulators are float32; every update is
    evaluated in float64 and rounded to float32 on store; the differences ``d - mean`` are float32 - float32 ->
    float32 for float32 predictions and float64 for float64 predictions (numpy promotion).  This is SURVEY.md
    App. C.4 corrected for the aliasing of ``mean_old`` found by the golden test (comment below).  The dead
    ``w_sum2`` accumulator (``:387,399``) is not reproduced."""
    f32, f64 = np.float32, np.float64
    a, i, p, o = geo.acc_side, geo.image_size, geo.purge, geo.off
    w = blend_weights(i)                                   # float64 (I-2p, I-2p)
    w_sum = np.zeros((a, a), f32)
    mean = np.zeros((a, a), f32)
    s_acc = np.zeros((a, a), f32)
    items = generated.items() if isinstance(generated, dict) else generated   # or an ordered [(key, pred)] (repeats)
    for (kx, ky), pred in items:
        accumulate_patch(w_sum, mean, s_acc, np.asarray(pred), minmax[(kx, ky)], kx, ky, w, i, p)
    w_c = w_sum[o:a - o, o:a - o]
    m_c = mean[o:a - o, o:a - o].copy()
    s_c = s_acc[o:a - o, o:a - o]
    good = w_c > 0                                                                       # :409
    with np.errstate(divide="ignore", invalid="ignore"):
        std = np.sqrt(s_c / w_c).astype(f32)                                             # :411 (f32 / f32, f32 sqrt)
    m_c[~good] = f32(no_value)
    std[~good] = f32(no_value)
    if return_accumulators:
        return m_c, std, good.astype(np.uint8), (w_sum, mean, s_acc)
    return m_c, std, good.astype(np.uint8)


def batch_plan(valid_origins: Sequence[Tuple[int, int]], batch_size: int) -> List[List[Tuple[int, int]]]:
    """process_full_tiles.py:459-474 -- chop the valid patches of one tile into batches of ``batch_size`` in visit
    order; the last partial batch is padded with (-1, -1) slots (all-zero inputs)."""
    out: List[List[Tuple[int, int]]] = []
    cur: List[Tuple[int, int]] = []
    for key in valid_origins:
        cur.append(key)
        if len(cur) == batch_size:
            out.append(cur)
            cur = []
    if cur:
        cur = cur + [(-1, -1)] * (batch_size - len(cur))
        out.append(cur)
    return out


ModelFn = Callable[..., np.ndarray]   # m(x, training=False)


def process_tile(dem_c: np.ndarray, img_c: np.ndarray, geo: Geometry, px: int, py: int, batch_size: int,
                 no_value: float, model: Optional[ModelFn] = None, return_patches: bool = False, repeats: int = 1):
    """process_full_tiles.py:431-479 (+327-345) -- one tile: gather valid patches, normalise, batch, run the model,
    take the last output channel ``+ 0.5``, blend.  ``model(batch[B,I,I,2], training=False) -> [B,I,I,C]``; default
    identity.

    ``repeats`` > 1 is the repeated-sample mode (beyond the reference): every batch is
    generated ``repeats`` times (a stochastic model draws new noise per call) and blended batch by batch, repetition
    by repetition, patch by patch; predictions enter the blend as float32."""
    i = geo.image_size
    keys, inputs, mm = [], {}, {}
    for xx, yy in patch_origins(geo, px, py):
        if not patch_is_valid(dem_c, img_c, xx, yy, i, no_value):
            continue
        key = (xx - px, yy - py)
        x, lohi = normalize_patch(img_c[yy:yy + i, xx:xx + i], dem_c[yy:yy + i, xx:xx + i])
        keys.append(key)
        inputs[key] = x
        mm[key] = lohi
    if repeats > 1:
        sequence = []
        for slots in batch_plan(keys, batch_size):
            batch = np.array([inputs[k] if k != (-1, -1) else np.zeros((i, i, 2)) for k in slots])
            for _ in range(repeats):
                pred = batch if model is None else model(batch, training=False)
                pred = (np.array(pred)[:, :, :, -1] + 0.5).astype(np.float32)
                sequence += [(k, y) for k, y in zip(slots, pred) if k != (-1, -1)]
        return rebuild_tile(sequence, mm, geo, no_value)
    generated: Dict[Tuple[int, int], np.ndarray] = {}
    for slots in batch_plan(keys, batch_size):
        # np.array(batch) (:338): float32, or float64 when float64 zero pads (:472) are present -- Keras casts to
        # float32 (App. B.10), a numpy callable sees the float64 array and its output dtype flows into the blend.
        batch = np.array([inputs[k] if k != (-1, -1) else np.zeros((i, i, 2)) for k in slots])
        pred = batch if model is None else model(batch, training=False)
        pred = np.array(pred)[:, :, :, -1] + 0.5                                          # :340
        for k, y in zip(slots, pred):
            if k != (-1, -1):
                generated[k] = y
    out = rebuild_tile(generated, mm, geo, no_value)
    if return_patches:
        return out, generated, mm
    return out


def assemble(tiles: Dict[Tuple[int, int], np.ndarray], geo: Geometry, dtype) -> np.ndarray:
    """process_full_tiles.py:541-545 -- paste tile (xx, yy) at [yy:yy+T, xx:xx+T] of a zero canvas, crop to (H, W)."""
    canvas = np.zeros((geo.canvas_h, geo.canvas_w), dtype=dtype)
    t = geo.tile_size
    for (xx, yy), tile in tiles.items():
        canvas[yy:yy + t, xx:xx + t] = tile
    return np.ascontiguousarray(canvas[:geo.canvas_h - geo.pad_y - geo.off, :geo.canvas_w - geo.pad_x - geo.off])


def process_map(dem: np.ndarray, img: np.ndarray, image_size: int, stride: int, batch_size: int, tile_size: int,
                no_value: float = -32768.0, model: Optional[ModelFn] = None, repeats: int = 1):
    """process_full_tiles.py:568-587 minus file I/O and preprocess(): pad -> tiles -> blend -> assemble.

    Returns (mean f32, std f32, good u8), each exactly (H, W)."""
    geo = Geometry(dem.shape[0], dem.shape[1], image_size, stride, tile_size)
    dem_c, img_c = pad_inputs(dem, img, geo, no_value)
    means, stds, goods = {}, {}, {}
    for (xx, yy) in tile_list(geo):
        m, s, g = process_tile(dem_c, img_c, geo, xx, yy, batch_size, no_value, model, repeats=repeats)
        means[(xx, yy)], stds[(xx, yy)], goods[(xx, yy)] = m, s, g
    return (assemble(means, geo, np.float32), assemble(stds, geo, np.float32), assemble(goods, geo, np.uint8))


# ----------------------------------------------------------------------------------------------------------------------
# Dedup mode (mode B): every patch position of the global stride lattice is generated once.
# Not a restatement of reference code -- the reference recomputes the I - S halo of every tile -- but built from the
# restated pieces above; for a model whose output does not depend on batch composition it is bit-identical to
# ``process_map`` (tests/test_oracle_tiling.py), because every pixel receives the same patches in the same order.
# ----------------------------------------------------------------------------------------------------------------------
def lattice_counts(geo: Geometry) -> Tuple[int, int]:
    """(GY, GX): patch origins per axis of the union of all tiles' windows (process_full_tiles.py:453-454 over
    generateTileList :313-325); needs S | T so that the tiles' lattices coincide."""
    if geo.tile_size % geo.stride != 0:
        raise ValueError("dedup mode needs stride | tile_size")
    t = geo.tile_size
    n_ty, n_tx = -(-geo.height // t), -(-geo.width // t)
    return (len(range(0, n_ty * t + geo.off, geo.stride)), len(range(0, n_tx * t + geo.off, geo.stride)))


def process_map_dedup(dem: np.ndarray, img: np.ndarray, image_size: int, stride: int, batch_size: int, tile_size: int,
                      no_value: float = -32768.0, model: Optional[ModelFn] = None,
                      row_bands: Optional[Sequence[Tuple[int, int]]] = None, return_plan: bool = False,
                      repeats: int = 1):
    """Global-lattice form of process_map: valid patches visited y outer / x inner over the whole canvas, chopped into
    batches of ``batch_size`` per band of lattice rows (``row_bands`` = [(j0, j1)], default one band; the last batch of
    a band is padded with float32 zero inputs), blended into canvas-sized float32 accumulators with the reference's
    update (accumulate_patch), finalised and cropped like rebuildTile :404-413 / rebuildMap :541-545."""
    f32 = np.float32
    geo = Geometry(dem.shape[0], dem.shape[1], image_size, stride, tile_size)
    i, s, p, o = image_size, stride, geo.purge, geo.off
    gy, gx = lattice_counts(geo)
    dem_c, img_c = pad_inputs(dem, img, geo, no_value)
    w = blend_weights(i)
    w_sum = np.zeros((geo.canvas_h, geo.canvas_w), f32)
    mean = np.zeros_like(w_sum)
    s_acc = np.zeros_like(w_sum)
    plan_out = []
    for (j0, j1) in (row_bands if row_bands is not None else [(0, gy)]):
        keys, inputs, mm = [], {}, {}
        for yy in range(j0 * s, j1 * s, s):
            for xx in range(0, gx * s, s):
                if not patch_is_valid(dem_c, img_c, xx, yy, i, no_value):
                    continue
                x, lohi = normalize_patch(img_c[yy:yy + i, xx:xx + i], dem_c[yy:yy + i, xx:xx + i])
                keys.append((xx, yy))
                inputs[(xx, yy)] = x
                mm[(xx, yy)] = lohi
        batches = batch_plan(keys, batch_size)
        plan_out.append(batches)
        for slots in batches:
            batch = np.array([inputs[k] if k != (-1, -1) else np.zeros((i, i, 2), f32) for k in slots])
            for _ in range(repeats):                                                       # repeated-sample mode: see process_tile
                pred = batch if model is None else model(batch, training=False)
                pred = (np.array(pred)[:, :, :, -1] + 0.5).astype(f32, copy=False)        # :340
                for k, y in zip(slots, pred):
                    if k != (-1, -1):
                        accumulate_patch(w_sum, mean, s_acc, y, mm[k], k[0], k[1], w, i, p)
    h, wd = geo.height, geo.width
    w_c, m_c, s_c = w_sum[o:o + h, o:o + wd], mean[o:o + h, o:o + wd].copy(), s_acc[o:o + h, o:o + wd]
    good = w_c > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        std = np.sqrt(s_c / w_c).astype(f32)
    m_c[~good] = f32(no_value)
    std[~good] = f32(no_value)
    out = (m_c, std, good.astype(np.uint8))
    return (out, plan_out) if return_plan else out
