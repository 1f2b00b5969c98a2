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


def rows_read_by_band(plan: Plan, world_size: int, rank: int) -> Tuple[int, int]:
    """Raster rows [n0, n1) the tiles of ``rank`` read: its own rows plus the I - S halo on both sides, clipped."""
    tiles, _, _ = band_of_rank(plan, world_size, rank)
    if not tiles:
        return 0, 0
    c0 = min(yy for _, yy in tiles)
    c1 = min(plan.canvas_h, max(yy for _, yy in tiles) + plan.tile_size + 2 * plan.off)
    return max(0, c0 - plan.off), min(plan.height, c1 - plan.off)


def exchange_halo_rows(own, plan: Plan, rank: int, world_size: int, group=None, bounds=None, needs=None):
    """Input-halo exchange for rasters that are loaded sharded (mode A): every rank holds exactly
    the raster rows of its own band of tiles (``own``: torch tensor (r1 - r0, W), CUDA with the NCCL backend, CPU with
    gloo) and receives the ``I - S`` halo rows its border tiles read from the neighbouring ranks by point-to-point
    send / recv -- over NVLink with NCCL.  Returns a tensor with rows [n0, n1) = ``rows_read_by_band``.
    ``bounds`` / ``needs`` (per-rank [r0, r1) lists) override the mode-A row ranges (dedup mode has its own).

    This is the only data that ever crosses ranks on the path; the outputs of different bands are disjoint."""
    import torch
    import torch.distributed as dist
    if bounds is None:
        bounds = [band_of_rank(plan, world_size, r)[1:] for r in range(world_size)]
    if needs is None:
        needs = [rows_read_by_band(plan, world_size, r) for r in range(world_size)]
    r0, r1 = bounds[rank]
    n0, n1 = needs[rank]
    if tuple(own.shape) != (r1 - r0, plan.width):
        raise ValueError(f"rank {rank} must hold rows [{r0}, {r1}) x {plan.width}, got {tuple(own.shape)}")
    out = torch.empty((n1 - n0, plan.width), dtype=own.dtype, device=own.device)
    if r1 > r0:
        out[r0 - n0:r1 - n0] = own
    if world_size == 1:
        return out
    ops, keep = [], []
    for peer in range(world_size):
        if peer == rank:
            continue
        p0, p1 = bounds[peer]
        q0, q1 = needs[peer]
        # rows of mine that the peer reads
        s0, s1 = max(r0, q0), min(r1, q1)
        if s1 > s0:
            t = own[s0 - r0:s1 - r0].contiguous()
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, peer, group))
        # rows of the peer that I read
        g0, g1 = max(p0, n0), min(p1, n1)
        if g1 > g0:
            t = torch.empty((g1 - g0, plan.width), dtype=own.dtype, device=own.device)
            keep.append((t, g0, g1))
            ops.append(dist.P2POp(dist.irecv, t, peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for item in keep:
        if isinstance(item, tuple):
            t, g0, g1 = item
            out[g0 - n0:g1 - n0] = t
    return out


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
