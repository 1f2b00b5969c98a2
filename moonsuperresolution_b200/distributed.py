"""Multi-GPU plumbing of the tiled path (one process per GPU, torch.distributed).

The path shards by tile (mode A): every output tile is self-sufficient -- its halo already
contains every patch that touches it (process_full_tiles.py:449-454) and batches never span tiles -- so ranks own
disjoint bands of tile rows, need no collective on the data path, and the only communication is the final gather of the
disjoint output row bands to rank 0.  Works with the NCCL backend (device tensors) and with gloo (CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .planner import Plan, shard_tiles


def band_of_rank(plan: Plan, world_size: int, rank: int) -> Tuple[List[Tuple[int, int]], int, int]:
    """Tiles owned by ``rank`` and the raster row range [r0, r1) they write (empty band: ([], 0, 0))."""
    tiles = plan.tiles()
    rows = sorted({yy for _, yy in tiles})
    mine = set(yy for _, yy in shard_tiles([(0, yy) for yy in rows], world_size, rank))
    own = [t for t in tiles if t[1] in mine]
    if not own:
        return [], 0, 0
    r0 = min(yy for _, yy in own)
    r1 = min(plan.height, max(yy for _, yy in own) + plan.tile_size)
    return own, r0, r1


def assemble_bands(parts: Sequence[Tuple[int, np.ndarray]], height: int, width: int, dtype) -> np.ndarray:
    """Stacks (first_row, band) pieces into the (H, W) raster; bands are disjoint row ranges."""
    out = np.zeros((height, width), dtype)
    for r0, band in parts:
        out[r0:r0 + band.shape[0]] = band
    return out


def gather_bands(band_arrays: Sequence[np.ndarray], r0: int, height: int, width: int, rank: int, world_size: int,
                 group=None) -> Optional[List[np.ndarray]]:
    """Gathers each rank's output bands (same list of rasters on every rank, e.g. mean / std / good) on rank 0 and
    assembles the full rasters there; returns None on the other ranks."""
    if world_size == 1:
        return [assemble_bands([(r0, a)], height, width, a.dtype) for a in band_arrays]
    import torch.distributed as dist
    parts = [None] * world_size if rank == 0 else None
    dist.gather_object((r0, [np.ascontiguousarray(a) for a in band_arrays]), parts, dst=0, group=group)
    if rank != 0:
        return None
    n = len(band_arrays)
    return [assemble_bands([(p[0], p[1][k]) for p in parts], height, width, band_arrays[k].dtype) for k in range(n)]
