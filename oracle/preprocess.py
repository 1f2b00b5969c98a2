"""CPU restatement of the reference's ``preprocess`` step (TEST INFRASTRUCTURE ONLY).

Reference: ``process_full_tiles.py`` ``preprocess`` :226-244, ``fillNan`` :214-224, ``interpolateMissingValues``
:184-212.  The arithmetic lives in third-party dependencies that are not part of /root/reference:

  * ``cv2.resize`` INTER_AREA / INTER_CUBIC  -- opencv-python 4.6.0.66 (pip-env.py:21); installed here: 4.13.0
  * ``cv2.connectedComponents``              -- same
  * ``scipy.interpolate.griddata(method='cubic')`` (Clough-Tocher on a Delaunay triangulation) -- scipy (conda-env)

Two layers:

  * ``reference_preprocess`` calls those libraries exactly where the reference does.  PINNED: tests/golden/
    make_golden_preprocess.py runs the UNMODIFIED reference method in the build container and
    tests/test_oracle_preprocess.py holds this restatement bit-exact to its output.
  * ``area4`` / ``cubic_resize`` restate, in numpy float32, the arithmetic of OpenCV's own resize code for the two calls
    the path makes (4x4 box mean with OpenCV's edge rule; 4-tap a = -0.75 cubic, horizontal pass accumulated tap 0 -> 3,
    vertical pass tap 3 -> 0 -- the order of OpenCV's vectorised vertical pass; its scalar tail for the last W mod 4
    columns runs 0 -> 3 -- replicated border).  PINNED against cv2 with IPP switched off (bit-exact; the W mod 4 tail
    columns of the cubic to rounding) -- the pip
    wheels route INTER_CUBIC through Intel IPP, whose rounding differs from OpenCV's code by a few ulp and is not
    documented; the CUDA kernels follow OpenCV's code and are held to cv2-with-IPP within 4 ulp of the raster's range.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------------------------------------
# OpenCV's own arithmetic, restated
# ----------------------------------------------------------------------------------------------------------------------
def area4_shape(h: int, w: int) -> Tuple[int, int]:
    """dsize of cv2.resize(src, (0, 0), fx=0.25, fy=0.25): cvRound (half to even) of the scaled extent."""
    return int(np.rint(h * 0.25)), int(np.rint(w * 0.25))


def area4(src: np.ndarray) -> np.ndarray:
    """cv2.resize(src, (0, 0), fx=0.25, fy=0.25, interpolation=cv2.INTER_AREA) for float32 (process_full_tiles.py:232,
    240).  Full 4 x 4 windows: four row sums ((s0 + s1) + s2) + s3 added top to bottom, times 1/16.  Windows cut by the
    raster edge: running sum in scan order divided by the count; windows entirely outside: 0."""
    src = np.asarray(src, f32)
    h, w = src.shape
    dh, dw = area4_shape(h, w)
    out = np.zeros((dh, dw), f32)
    fh, fw = min(dh, h // 4), min(dw, w // 4)
    with np.errstate(invalid="ignore", over="ignore"):
        blk = src[:fh * 4, :fw * 4].reshape(fh, 4, fw, 4)
        rows = ((blk[..., 0] + blk[..., 1]) + blk[..., 2]) + blk[..., 3]            # (fh, 4, fw)
        acc = np.zeros((fh, fw), f32)
        for r in range(4):
            acc = acc + rows[:, r, :]
        out[:fh, :fw] = acc * f32(1.0 / 16.0)
        for dy in range(dh):
            for dx in range(dw):
                if dy < fh and dx < fw:
                    continue
                sy0, sx0 = dy * 4, dx * 4
                if sy0 >= h or sx0 >= w:
                    continue
                s, c = f32(0), 0
                for yy in range(sy0, min(sy0 + 4, h)):
                    for xx in range(sx0, min(sx0 + 4, w)):
                        s = f32(s + src[yy, xx])
                        c += 1
                out[dy, dx] = f32(s / f32(c))
    return out


def cubic_tables(dst: int, src: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per destination index: first-tap-plus-one source index and the four float32 weights of OpenCV's INTER_CUBIC
    (a = -0.75): position (d + 0.5) * src/dst - 0.5 in double, rounded to float32, split into floor and fraction."""
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * (float(src) / float(dst)) - 0.5).astype(f32)
    sx = np.floor(fx).astype(np.int32)
    x = (fx - sx.astype(f32)).astype(f32)
    a = f32(-0.75)
    one, x1 = f32(1), (x + f32(1)).astype(f32)
    c0 = ((a * x1 - f32(5) * a) * x1 + f32(8) * a) * x1 - f32(4) * a
    c1 = ((a + f32(2)) * x - (a + f32(3))) * x * x + one
    y = (one - x).astype(f32)
    c2 = ((a + f32(2)) * y - (a + f32(3))) * y * y + one
    c3 = one - c0 - c1 - c2
    return sx, np.stack([c0, c1, c2, c3], axis=1).astype(f32)


def cubic_resize(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_CUBIC) as OpenCV's own code computes it for float32
    (process_full_tiles.py:241): horizontal pass first (taps accumulated 0 -> 3), then the vertical pass (3 -> 0),
    source indices clamped to the raster."""
    src = np.asarray(src, f32)
    h, w = src.shape
    xo, xc = cubic_tables(dw, w)
    yo, yc = cubic_tables(dh, h)
    with np.errstate(invalid="ignore", over="ignore"):
        xi = np.clip(xo[:, None] - 1 + np.arange(4)[None, :], 0, w - 1)                # (dw, 4)
        tmp = src[:, xi[:, 0]] * xc[None, :, 0]
        for k in (1, 2, 3):
            tmp = tmp + src[:, xi[:, k]] * xc[None, :, k]
        yi = np.clip(yo[:, None] - 1 + np.arange(4)[None, :], 0, h - 1)                # (dh, 4)
        out = tmp[yi[:, 3], :] * yc[:, 3, None]
        for k in (2, 1, 0):
            out = out + tmp[yi[:, k], :] * yc[:, k, None]
    return out.astype(f32)


# ----------------------------------------------------------------------------------------------------------------------
# the reference's step, through the libraries it calls
# ----------------------------------------------------------------------------------------------------------------------
def interpolate_missing(block: np.ndarray, no_value: float, max_fill_area: int) -> np.ndarray:
    """process_full_tiles.py:184-212 on one block (modified in place and returned).  Invalid = ``<= no_value``.  Nothing
    happens when the block has no or only invalid pixels, or when its SMALLEST connected-component count -- the valid
    background is label 0 and takes part, :196-199 -- exceeds ``max_fill_area``.  Otherwise every pixel of a component
    with fewer than ``max_fill_area`` pixels is replaced by the cubic (Clough-Tocher) interpolant through all valid
    pixels of the block (NaN outside their convex hull)."""
    import cv2
    from scipy import interpolate
    invalid = block <= no_value
    if not invalid.any() or invalid.all():
        return block
    _, labels = cv2.connectedComponents((invalid * 255).astype(np.uint8))
    ids, counts = np.unique(labels, return_counts=True)
    if counts.min() > max_fill_area:
        return block
    yy, xx = np.nonzero(~invalid)
    gx, gy = np.meshgrid(np.arange(block.shape[1]), np.arange(block.shape[0]))
    interp = interpolate.griddata((xx, yy), block[~invalid].ravel(), (gx, gy), method="cubic")
    keep = np.isin(labels, ids[counts < max_fill_area])
    block[keep] = interp[keep]
    return block


def fill_nan(image: np.ndarray, no_value: float, tile_size: int, border: int, max_fill_area: int) -> np.ndarray:
    """process_full_tiles.py:214-224 -- overlapping blocks of ``tile_size`` at stride ``tile_size - 2*border``; each
    block is interpolated on its own copy of the ORIGINAL image and only its interior (``border`` cut on every side)
    is written back, so the outer ``border`` frame of the image is never filled."""
    out = image.copy()
    step = tile_size - 2 * border
    h, w = image.shape
    for y in range(0, h, step):
        y1 = min(y + tile_size - border, h - border)
        for x in range(0, w, step):
            x1 = min(x + tile_size - border, w - border)
            blk = interpolate_missing(image[y:y + tile_size, x:x + tile_size].copy(), no_value, max_fill_area)
            out[y + border:y1, x + border:x1] = blk[border:-border, border:-border]
    return out


def reference_preprocess(dem: np.ndarray, no_value: float, fix_shape: bool = True) -> np.ndarray:
    """process_full_tiles.py:226-244, DEM half (the ortho half, :227, stores its result in ``self.image``, which nothing
    reads -- dead code).  1/4 INTER_AREA with no_value as NaN, small-hole fill on the 1/4 raster, another 1/4
    INTER_AREA, INTER_CUBIC back to full size, NaN -> no_value.

    The reference passes ``self.dem_shape`` = (H, W) as cv2's dsize = (width, height), which transposes the extent of a
    non-square raster (and then crashes in padInputs); ``fix_shape`` (default) asks for (W, H), identical for squares."""
    import cv2
    nv = no_value
    d = np.array(dem, dtype=f32, copy=True)
    d[d <= nv] = np.nan
    d = cv2.resize(d, (0, 0), fx=0.25, fy=0.25, interpolation=cv2.INTER_AREA)
    d[np.isnan(d)] = nv
    d = fill_nan(d, nv, tile_size=256, border=32, max_fill_area=24)
    d[d <= nv] = np.nan
    d = cv2.resize(d, (0, 0), fx=0.25, fy=0.25, interpolation=cv2.INTER_AREA)
    h, w = dem.shape
    d = cv2.resize(d, (w, h) if fix_shape else (h, w), interpolation=cv2.INTER_CUBIC)
    d[np.isnan(d)] = nv
    return d
