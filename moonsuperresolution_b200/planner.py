"""Integer planning of the tiled inference path (host side, no pixels touched).

Restates, as closed-form integer arithmetic, the loop structure of the reference:
  * canvas geometry     -- padInputs            (process_full_tiles.py:246-267; literal 1024 at :252-253)
  * tile list           -- generateTileList     (process_full_tiles.py:313-325)
  * patch lattice       -- processTile          (process_full_tiles.py:453-454; y outer, x inner)
  * batch plan          -- processTile          (process_full_tiles.py:459-474; (-1, -1) padding slots)
  * band sharding       -- new: contiguous runs of tiles per rank (mode A: tiles are
                           self-sufficient, so ranks never exchange data on the path)
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np

CANVAS_QUANTUM = 1024      # process_full_tiles.py:252-253 uses the literal 1024, not tile_size
PAD_SLOT = (-1, -1)        # process_full_tiles.py:473


@dataclasses.dataclass(frozen=True)
class Plan:
    """Geometry of one run.  (x, y) = (column, row) everywhere."""
    height: int
    width: int
    image_size: int
    stride: int
    tile_size: int
    batch_size: int

    def __post_init__(self):
        i, s, t = self.image_size, self.stride, self.tile_size
        if self.height <= 0 or self.width <= 0:
            raise ValueError("empty raster")
        if i < 16:
            raise ValueError("image_size must be >= 16: purge = image_size // 16 must be >= 1 "
                             "(process_full_tiles.py:392-393 slices w[purge:-purge])")
        if not (0 < s <= i):
            raise ValueError("stride must be in (0, image_size]")
        if (t + i) % s != 0:
            raise ValueError("stride must divide tile_size + image_size, else the last patch overruns the "
                             "accumulator of rebuildTile (process_full_tiles.py:386,398)")
        if self.batch_size <= 0 or t <= 0:
            raise ValueError("batch_size and tile_size must be positive")
        for dim in (self.height, self.width):
            if ((dim - 1) // t + 1) * t > (dim // CANVAS_QUANTUM + 1) * CANVAS_QUANTUM:
                raise ValueError("tile_size is incompatible with the 1024-quantised canvas of padInputs "
                                 "(process_full_tiles.py:252-253): a tile would read beyond the canvas")

    # ---- padInputs ------------------------------------------------------------------------------------------------
    @property
    def off(self) -> int:
        return self.image_size - self.stride

    @property
    def purge(self) -> int:
        return self.image_size // 16

    @property
    def canvas_h(self) -> int:
        return (self.height // CANVAS_QUANTUM + 1) * CANVAS_QUANTUM + 2 * self.off

    @property
    def canvas_w(self) -> int:
        return (self.width // CANVAS_QUANTUM + 1) * CANVAS_QUANTUM + 2 * self.off

    @property
    def pad_x(self) -> int:
        return self.canvas_w - self.width - self.off

    @property
    def pad_y(self) -> int:
        return self.canvas_h - self.height - self.off

    # ---- generateTileList -----------------------------------------------------------------------------------------
    def tiles(self) -> List[Tuple[int, int]]:
        return [(xx, yy) for yy in range(0, self.height, self.tile_size) for xx in range(0, self.width, self.tile_size)]

    # ---- processTile ------------------------------------------------------------------------------------------------
    @property
    def lattice_side(self) -> int:
        """Patches per axis of one tile: len(range(p, p + T + I - S, S))."""
        return -(-(self.tile_size + self.image_size - self.stride) // self.stride)

    def tile_patch_origins(self, px: int, py: int) -> np.ndarray:
        """(G*G, 2) int32 canvas coordinates (x, y) of tile (px, py)'s patches in the reference's visit order."""
        g = self.lattice_side
        ks = np.arange(g, dtype=np.int32) * self.stride
        xs = np.tile(ks + px, g)
        ys = np.repeat(ks + py, g)
        return np.stack([xs, ys], axis=1).astype(np.int32)

    def batch_slots(self, n_valid: int) -> int:
        """Slots executed for a tile with n_valid patches (last batch padded, none when n_valid == 0)."""
        b = self.batch_size
        return -(-n_valid // b) * b

    def tile_window(self, px: int, py: int) -> Tuple[int, int]:
        """(rows, cols) of tile (px, py) that survive the crop of rebuildMap (process_full_tiles.py:541-545)."""
        return min(self.tile_size, self.height - py), min(self.tile_size, self.width - px)


    # ---- dedup mode (mode B) --------------------------------------------------------------------
    def lattice_counts(self) -> Tuple[int, int]:
        """(GY, GX): patch origins per axis of the union of all tiles' windows (process_full_tiles.py:453-454 over the
        tile list :313-325).  Needs S | T so that the lattices of neighbouring tiles coincide."""
        s, t = self.stride, self.tile_size
        if t % s != 0:
            raise ValueError("dedup mode needs stride | tile_size (the tiles' patch lattices must coincide)")
        n_ty, n_tx = -(-self.height // t), -(-self.width // t)
        return len(range(0, n_ty * t + self.off, s)), len(range(0, n_tx * t + self.off, s))

    def dedup_band(self, world_size: int, rank: int) -> "DedupBand":
        """Band of lattice rows owned by ``rank`` and everything that follows from it (all in canvas rows unless the
        name says raster).  Rows are split evenly; a band must be at least as tall as the overlap so that only
        adjacent ranks share accumulator rows."""
        if world_size <= 0 or not (0 <= rank < world_size):
            raise ValueError("bad rank / world_size")
        gy, gx = self.lattice_counts()
        i, s, p, off = self.image_size, self.stride, self.purge, self.off
        # lattice rows whose patches lie inside the raster rows carry the work (the others see only no_value padding):
        # cut so that every rank gets the same number of those
        inside = [1 if (j * s >= off and j * s + i <= off + self.height) else 0 for j in range(gy)]
        total = sum(inside)
        first_in = inside.index(1) if total else 0
        cuts = [0] + [first_in + (r * total) // world_size for r in range(1, world_size)] + [gy]
        j0, j1 = cuts[rank], cuts[rank + 1]
        if world_size > 1 and min(b - a for a, b in zip(cuts[:-1], cuts[1:])) < max(1, -(-(i - 2 * p) // s)):
            raise ValueError(f"{world_size} ranks leave fewer lattice rows per band than one patch spans "
                             f"({gy} rows in total): use fewer ranks for this raster")
        first, last = rank == 0, rank == world_size - 1
        # A band finalises canvas rows [j0*S + q, j1*S + q): no patch of the NEXT band reaches above j1*S + p, and the
        # patches of the PREVIOUS band reach down to j0*S + off - p -- the seam, present iff that lies below j0*S + q.
        q = min(p, off)
        overlap = off - p > q
        return DedupBand(
            j0=j0, j1=j1, gx=gx,
            read=(j0 * s, min(self.canvas_h, (j1 - 1) * s + i)),
            out=(0 if first else j0 * s + q, self.canvas_h if last else j1 * s + q),
            seam_in=None if first or not overlap else (j0 * s + q, j0 * s + off - p),
            seam_out=None if last or not overlap else (j1 * s + q, j1 * s + off - p))


    def dedup_rows(self, world_size: int, rank: int) -> Tuple[Tuple[int, int], Tuple[int, int]]:
        """Dedup mode, in RASTER rows: (rows this rank finalises, rows its patches read -- a superset of the former)."""
        band = self.dedup_band(world_size, rank)
        own = band.raster_rows(band.out, self.off, self.height)
        need = band.raster_rows(band.read, self.off, self.height)
        if own[1] > own[0]:
            need = (min(need[0], own[0]), max(need[1], own[1])) if need[1] > need[0] else own
        return own, need


@dataclasses.dataclass(frozen=True)
class DedupBand:
    """One rank's share of the global patch lattice in dedup mode.  ``read``: canvas rows its patches read (= rows of its
    accumulators); ``out``: canvas rows it finalises; ``seam_in`` / ``seam_out``: accumulator rows received from the
    previous / sent to the next rank (None at the ends or without overlap)."""
    j0: int
    j1: int
    gx: int
    read: Tuple[int, int]
    out: Tuple[int, int]
    seam_in: Optional[Tuple[int, int]]
    seam_out: Optional[Tuple[int, int]]

    def raster_rows(self, rows: Tuple[int, int], off: int, height: int) -> Tuple[int, int]:
        """Canvas row range -> raster row range, clipped to the raster."""
        r0 = min(max(rows[0] - off, 0), height)
        return r0, max(r0, min(rows[1] - off, height))


def plan_batches(valid_keys: Sequence[Tuple[int, int]], batch_size: int) -> List[List[Tuple[int, int]]]:
    """process_full_tiles.py:459-474 -- valid patches in visit order, chopped into batches, last one padded."""
    out, cur = [], []
    for k in valid_keys:
        cur.append(k)
        if len(cur) == batch_size:
            out.append(cur)
            cur = []
    if cur:
        out.append(cur + [PAD_SLOT] * (batch_size - len(cur)))
    return out


def shard_tiles(tiles: Sequence[Tuple[int, int]], world_size: int, rank: int, cost: Sequence[int] = None
                ) -> List[Tuple[int, int]]:
    """Contiguous run of the tile list owned by ``rank`` (mode A of SURVEY.md 8e).

    Tiles are independent work units (generateTileList's own comment, process_full_tiles.py:320), so the split needs
    no data-path collective.  With ``cost`` (e.g. valid patch slots per tile) the cut points balance the summed
    cost; otherwise tile counts.  Every tile is owned by exactly one rank; ranks may own none."""
    n = len(tiles)
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    if cost is None:
        cost = [1] * n
    if len(cost) != n:
        raise ValueError("cost must have one entry per tile")
    total = int(sum(cost))
    if total == 0:
        lo, hi = (rank * n) // world_size, ((rank + 1) * n) // world_size
        return list(tiles[lo:hi])
    prefix = np.concatenate([[0], np.cumsum(np.asarray(cost, dtype=np.int64))])
    # tile j goes to the rank whose interval contains the midpoint of its cost span
    mid = (prefix[:-1] + prefix[1:]) / 2.0
    owner = np.minimum((mid * world_size / total).astype(np.int64), world_size - 1)
    return [t for t, o in zip(tiles, owner) if o == rank]
