"""``preprocess`` of the reference (process_full_tiles.py:226-244) with the two resampling passes on the GPU.

What the reference does to the DEM before it is tiled (the ortho half of the method, :227, stores its result in
``self.image``, which nothing reads -- it has no effect on the outputs and is not reproduced):

    no_value -> NaN; 1/4 INTER_AREA; NaN -> no_value            (:228-233)   msr_resize_area4
    fillNan(tile 256, border 32, max_fill_area 24)               (:235)       host, see below
    no_value -> NaN; 1/4 INTER_AREA; INTER_CUBIC to full size;   (:238-243)   msr_resize_area4 + msr_resize_cubic
    NaN -> no_value

i.e. the network is fed a DEM low-passed to 1/16 resolution.  The hole fill works on the 1/4 raster (1/16 of the
pixels) and only on 256 x 256 blocks that contain invalid pixels; which blocks do is decided on the device with the
summed-area-table kernels of the tiling side.  Its arithmetic is scipy's Clough-Tocher interpolant over a Delaunay
triangulation plus cv2's connected components (process_full_tiles.py:184-212): library code in the reference, the same
library calls here, on the host -- there is no device formulation that would reproduce a triangulation-dependent
interpolant value for value.  A raster without invalid pixels never leaves the device.

One deliberate deviation: the reference hands ``self.dem_shape`` = (H, W) to cv2.resize as (width, height) (:241), so a
non-square raster comes back transposed and padInputs crashes; here the upsampling targets (H, W).  Identical for
square rasters.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib

FILL_TILE, FILL_BORDER, FILL_MAX_AREA = 256, 32, 24          # process_full_tiles.py:235
_f32 = np.float32


def area4_shape(h: int, w: int) -> Tuple[int, int]:
    """Shape of cv2.resize(x, (0, 0), fx=0.25, fy=0.25): extents scaled and rounded half to even (cvRound)."""
    return int(np.rint(h * 0.25)), int(np.rint(w * 0.25))


def cubic_tables(dst: int, src: int) -> Tuple[np.ndarray, np.ndarray]:
    """INTER_CUBIC lookup tables for one axis as OpenCV builds them: the sample position ``(d + 0.5) * src / dst - 0.5``
    is evaluated in double, rounded to float32 and split into floor (index of tap 1) and fraction ``t``; the weights of
    the Keys kernel with a = -0.75 are evaluated in float32 in Horner form, the last one as 1 minus the other three."""
    pos = ((np.arange(dst, dtype=np.float64) + 0.5) * (float(src) / float(dst)) - 0.5).astype(_f32)
    first = np.floor(pos).astype(np.int32)
    t = (pos - first.astype(_f32)).astype(_f32)
    a, one = _f32(-0.75), _f32(1)
    u = (t + one).astype(_f32)                 # distance to tap 0
    v = (one - t).astype(_f32)                 # distance to tap 2
    w0 = ((a * u - _f32(5) * a) * u + _f32(8) * a) * u - _f32(4) * a
    w1 = ((a + _f32(2)) * t - (a + _f32(3))) * t * t + one
    w2 = ((a + _f32(2)) * v - (a + _f32(3))) * v * v + one
    w3 = one - w0 - w1 - w2
    return first, np.ascontiguousarray(np.stack([w0, w1, w2, w3], axis=1), dtype=_f32)


# ----------------------------------------------------------------------------------------------------------------------
# small-hole fill (host; process_full_tiles.py:184-224)
# ----------------------------------------------------------------------------------------------------------------------
def fill_blocks(h: int, w: int, tile: int = FILL_TILE, border: int = FILL_BORDER) -> List[Tuple[int, int]]:
    """(x, y) origins of fillNan's blocks (:217-219): stride ``tile - 2*border`` over both axes."""
    step = tile - 2 * border
    return [(x, y) for y in range(0, h, step) for x in range(0, w, step)]


def _fill_block(block: np.ndarray, no_value: float, max_fill_area: int) -> np.ndarray:
    """interpolateMissingValues (:184-212) on a private copy of one block."""
    import cv2
    from scipy.interpolate import griddata
    bad = block <= no_value
    n_bad = int(bad.sum())
    if n_bad == 0 or n_bad == bad.size:                                   # :188-195
        return block
    _, labels = cv2.connectedComponents(bad.astype(np.uint8) * 255)       # :196 (8-connectivity)
    ids, sizes = np.unique(labels, return_counts=True)                    # label 0 = the valid background
    if sizes.min() > max_fill_area:                                       # :199-201
        return block
    ok = ~bad
    rows, cols = np.nonzero(ok)
    grid_x, grid_y = np.meshgrid(np.arange(block.shape[1]), np.arange(block.shape[0]))
    surface = griddata((cols, rows), block[ok].ravel(), (grid_x, grid_y), method="cubic")   # :206
    small = np.isin(labels, ids[sizes < max_fill_area])                   # :208-210 (strictly smaller)
    block[small] = surface[small]
    return block


def fill_small_holes(image: np.ndarray, no_value: float, origins: Sequence[Tuple[int, int]] = None,
                     tile: int = FILL_TILE, border: int = FILL_BORDER, max_fill_area: int = FILL_MAX_AREA,
                     threads: int = 0) -> np.ndarray:
    """fillNan (:214-224).  ``origins`` restricts the work to the blocks known to hold invalid pixels (a block without
    any is returned unchanged by the reference, :188-190).  Every block is interpolated on a copy of the ORIGINAL image
    and only its interior is written back, so blocks are independent and run on a thread pool."""
    h, w = image.shape
    out = image.copy()
    todo = list(fill_blocks(h, w, tile, border) if origins is None else origins)
    if not todo:
        return out

    def work(xy):
        x, y = xy
        return xy, _fill_block(image[y:y + tile, x:x + tile].copy(), no_value, max_fill_area)

    n_threads = threads or min(len(todo), max(1, len(os.sched_getaffinity(0))))
    if n_threads > 1:
        with ThreadPoolExecutor(n_threads) as pool:
            results = list(pool.map(work, todo))
    else:
        results = [work(xy) for xy in todo]
    for (x, y), blk in results:
        y1, x1 = min(y + tile - border, h - border), min(x + tile - border, w - border)
        if y1 > y + border and x1 > x + border:
            out[y + border:y1, x + border:x1] = blk[border:blk.shape[0] - border, border:blk.shape[1] - border]
    return out


# ----------------------------------------------------------------------------------------------------------------------
# device pipeline
# ----------------------------------------------------------------------------------------------------------------------
def _area4(lib, torch, src, no_value: float):
    h, w = int(src.shape[0]), int(src.shape[1])
    dh, dw = area4_shape(h, w)
    if dh <= 0 or dw <= 0:
        raise ValueError("raster too small for preprocess (needs at least 2 pixels per axis at 1/4 scale)")
    dst = torch.empty((dh, dw), dtype=torch.float32, device=src.device)
    _lib.check(lib.msr_resize_area4(src.data_ptr(), h, w, dst.data_ptr(), dh, dw, no_value, _lib.stream_ptr()),
               "msr_resize_area4")
    return dst


def _blocks_with_holes(lib, torch, quarter, no_value: float) -> List[Tuple[int, int]]:
    """fillNan blocks of the 1/4 raster that contain an invalid pixel: summed-area table of the invalid mask
    (msr_validity_sat) probed with one clipped 256 x 256 window per block (msr_patch_validity)."""
    h, w = int(quarter.shape[0]), int(quarter.shape[1])
    origins = fill_blocks(h, w)
    sat = torch.empty((h + 1, w + 1), dtype=torch.int32, device=quarter.device)
    st = _lib.stream_ptr()
    _lib.check(lib.msr_validity_sat(quarter.data_ptr(), quarter.data_ptr(), h, w, no_value, sat.data_ptr(), st),
               "msr_validity_sat")
    d_xy = torch.tensor(origins, dtype=torch.int32, device=quarter.device)
    d_ok = torch.empty((len(origins),), dtype=torch.uint8, device=quarter.device)
    _lib.check(lib.msr_patch_validity(sat.data_ptr(), h, w, d_xy.data_ptr(), len(origins), FILL_TILE, d_ok.data_ptr(),
                                      st), "msr_patch_validity")
    ok = d_ok.cpu().numpy().astype(bool)
    return [o for o, good in zip(origins, ok) if not good]


def preprocess_dem(dem, no_value: float, fill_holes: bool = True):
    """(H, W) float32 CUDA tensor -> preprocessed (H, W) float32 CUDA tensor (process_full_tiles.py:228-244).  Returns
    (tensor, kernel launches issued)."""
    import torch
    lib = _lib.lib()
    if not (torch.is_tensor(dem) and dem.is_cuda and dem.dtype == torch.float32 and dem.dim() == 2):
        raise ValueError("preprocess_dem needs a 2-D float32 CUDA tensor")
    dem = dem.contiguous()
    h, w = int(dem.shape[0]), int(dem.shape[1])
    nv = float(np.float32(no_value))
    launches = 0
    quarter = _area4(lib, torch, dem, nv)                                            # :228-233
    launches += 1
    if fill_holes:
        holes = _blocks_with_holes(lib, torch, quarter, nv)
        launches += 3
        if holes:                                                                    # :235
            host = quarter.cpu().numpy()
            filled = fill_small_holes(host, nv, origins=holes)
            quarter = torch.from_numpy(filled).to(dem.device)
    sixteenth = _area4(lib, torch, quarter, nv)                                      # :238-240
    xo, xc = cubic_tables(w, int(sixteenth.shape[1]))
    yo, yc = cubic_tables(h, int(sixteenth.shape[0]))
    d_xo, d_xc = torch.from_numpy(xo).to(dem.device), torch.from_numpy(xc).to(dem.device)
    d_yo, d_yc = torch.from_numpy(yo).to(dem.device), torch.from_numpy(yc).to(dem.device)
    out = torch.empty((h, w), dtype=torch.float32, device=dem.device)
    _lib.check(lib.msr_resize_cubic(sixteenth.data_ptr(), int(sixteenth.shape[0]), int(sixteenth.shape[1]),
                                    out.data_ptr(), h, w, d_xo.data_ptr(), d_xc.data_ptr(), d_yo.data_ptr(),
                                    d_yc.data_ptr(), nv, _lib.stream_ptr()), "msr_resize_cubic")   # :241-243
    launches += 2
    return out, launches
