"""Tiled full-DEM super-resolution engine: a drop-in for ``DEMSuperResolution`` of the reference
(process_full_tiles.py:129-587) whose pixel work runs in libmoonsr.so on a B200.

Same class name, constructor ``(config, model=callable)``, attribute names and method names as the reference, same tile
/ overlap parameters and output layout (mean SR DEM, weighted population std, good mask; no_value outside ``good``).
What differs is where the data lives:

  * ``padInputs`` uploads the rasters once and builds the no_value canvases, the validity summed-area table and the
    validity of every patch of every tile ON THE DEVICE (process_full_tiles.py:246-293);
  * ``processTile`` gathers + normalises patches, runs the generator and blends, all on the device; the tile result is
    written straight into the (H, W) output rasters, which fuses the paste + crop of ``rebuildMap``
    (process_full_tiles.py:431-479, 541-545).  Per-tile TIFFs -- the reference's RAM-spill mechanism -- are written only
    when ``save_tiles`` is set;
  * a model that is one of this package's generators (``models.GauGAN`` ...) is called through its device entry point;
    any other callable obeying the reference's plug-in contract ``m(x, training=False)`` is fed host numpy batches
    exactly as the reference would (including the float64 promotion of zero-padded batches, :472).

Multi-GPU (mode A): ``rank`` / ``world_size`` give each process a contiguous band of tile rows; a rank
holds only the canvas rows its band needs; tiles are self-sufficient so no data crosses ranks on the path.

There is no CPU fallback: without libmoonsr.so or a CUDA device the engine raises.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .models import IdentityModel, _Generator
from .distributed import band_of_rank, exchange_halo_rows, gather_bands
from .planner import PAD_SLOT, Plan


@dataclasses.dataclass
class DSRConfig:
    """process_full_tiles.py:53-66, field for field; the trailing fields are additions whose defaults reproduce the
    reference's behaviour."""
    image_size: int = 256
    stride: int = 32
    batch_size: int = 16
    tile_size: int = 1024
    no_value: float = -32768.0
    upsample_factor: float = 1.0
    map_name: str = None
    save_path: str = None
    source_folder_path: str = None
    ortho_image_name: str = "run-DRG.tif"
    dem_name: str = "run-DEM.tif"
    model_path: str = None
    # ---- additions ----
    save_tiles: bool = False        # write the per-tile TIFFs of saveTile (process_full_tiles.py:416-429)
    groups_per_call: int = 0        # batches pushed through one generator call (0 = the model's max_groups)
    seed: int = 0                   # seeds the Gaussian sampler's noise (the reference's is unseeded)
    preprocess: bool = True         # run preprocess() inside processMap like the reference (:577); False = rasters are
                                    # taken as already filtered
    samples_per_patch: int = 1      # > 1: repeated-sample mode (beyond the reference): every batch is generated this many
                                    # times (new sampler noise each time) and all generations are blended
    mode: str = "faithful"          # "faithful": tile by tile like the reference (halo patches recomputed per tile);
                                    # "dedup": every position of the global patch lattice generated once
    reuse_spade: bool = True        # samples_per_patch > 1 with a bf16 SPADE device model: the encoder, the mask convolutions
                                    # and gamma | beta of every SPADE layer (spade.py:18-20) depend only on the patch, not on
                                    # the sampler's noise -- compute them for the first generation, reuse them for the rest
    blend: str = "exact"            # "exact": rebuildTile's arithmetic bit for bit (float64 intermediates, :395-402);
                                    # "fast": same placement / order / good mask, float32 update, 128-bit accesses --
                                    # HBM-bound, values within float32 rounding (~1e-6 relative) of "exact"


def parse_args(argv=None) -> DSRConfig:
    """process_full_tiles.py:68-127 -- same flags, defaults and help texts' meaning; unknown flags are ignored
    (parse_known_args, :114).  Additions: --save_tiles, --groups_per_call, --seed, --mode,
    --no_preprocess."""
    import argparse
    parser = argparse.ArgumentParser("DEM Super Resolution config parser.")
    parser.add_argument("--source_folder_path", type=str, required=True, default=None,
                        help="The path to the folder containing both the ortho image and the DEM.")
    parser.add_argument("--map_name", type=str, required=True, default=None, help="The name of the map to be processes.")
    parser.add_argument("--save_path", type=str, required=True, default=None,
                        help="The path to the folder where the reconstructed map will be stored.")
    parser.add_argument("--ortho_image_name", type=str, default="run-DRG.tif", help="The name of the ortho image.")
    parser.add_argument("--dem_name", type=str, default="run-DEM.tif", help="The name of the DEM image.")
    parser.add_argument("--model_path", type=str, default=None,
                        help="The path to the model. Do not specify to run indentity processing.")
    parser.add_argument("--image_size", type=int, default=256, help="The size of the images the model can process.")
    parser.add_argument("--stride", type=int, default=32, help="The amount of displacement between two images.")
    parser.add_argument("--batch_size", type=int, default=16, help="The batch size of the model.")
    parser.add_argument("--tile_size", type=int, default=1024, help="The size of the tiles.")
    parser.add_argument("--no_value", type=int, default=-32768.0, help="The value marking points without data.")
    parser.add_argument("--upsample_factor", type=float, default=1.0, help="Not used for now.")
    parser.add_argument("--save_tiles", action="store_true", help="Also write the per-tile TIFFs of saveTile.")
    parser.add_argument("--groups_per_call", type=int, default=0, help="Batches per generator call (0 = model's maximum).")
    parser.add_argument("--seed", type=int, default=0, help="Seed of the Gaussian sampler's noise.")
    parser.add_argument("--no_preprocess", action="store_true",
                        help="Skip preprocess() (1/16 box filter + cubic upsampling of the DEM) inside processMap.")
    parser.add_argument("--samples_per_patch", type=int, default=1,
                        help="Generations per patch position (N = samples_per_patch * (image_size / stride)^2).")
    parser.add_argument("--mode", type=str, default="faithful", choices=["faithful", "dedup"],
                        help="faithful: tile by tile like the reference; dedup: every patch position generated once.")
    parser.add_argument("--blend", type=str, default="exact", choices=["exact", "fast"],
                        help="exact: the reference's blend arithmetic bit for bit; fast: float32 update, HBM-bound.")
    args, _unknown = parser.parse_known_args(argv)
    return DSRConfig(source_folder_path=args.source_folder_path, map_name=args.map_name, save_path=args.save_path,
                     ortho_image_name=args.ortho_image_name, dem_name=args.dem_name, model_path=args.model_path,
                     image_size=args.image_size, stride=args.stride, batch_size=args.batch_size,
                     tile_size=args.tile_size, no_value=args.no_value, upsample_factor=args.upsample_factor,
                     save_tiles=args.save_tiles, groups_per_call=args.groups_per_call, seed=args.seed, mode=args.mode,
                     preprocess=not args.no_preprocess, samples_per_patch=args.samples_per_patch, blend=args.blend)


def main(argv=None) -> None:
    """process_full_tiles.py:589-594; without --model_path the identity model runs (the reference's own CLI cannot reach
    its identity mode).  It is the reference's host callable (``lambda x, training=False: x``, :143),
    fed numpy batches exactly as the reference would -- including the float64 promotion of the zero-padded last batch of
    a tile (:472) -- so the identity round trip is bit-faithful; ``IdentityModel`` is the same check at device speed
    (float32 throughout) for callers that construct the engine themselves."""
    from .models import load_GAN_model
    cfg = parse_args(argv)
    if cfg.model_path is None:
        model = _identity
    else:
        model = load_GAN_model(cfg.model_path, cfg.image_size, cfg.batch_size, max_groups=max(1, cfg.groups_per_call or 8))
    DEMSuperResolution(cfg, model=model).processMap()


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.MoonSRError("no CUDA device: DEMSuperResolution runs only on the GPU (there is no CPU fallback)")
    return torch


def _identity(x, training=False):
    return x


class DEMSuperResolution:
    """See the module docstring.  Reference: process_full_tiles.py:129-155."""

    def __init__(self, config: DSRConfig, model=_identity, rank: int = 0, world_size: int = 1, device=None) -> None:
        self.map_name = config.map_name
        self.save_path = config.save_path
        self.folder_path = config.source_folder_path
        self.left_image_name = config.ortho_image_name
        self.dem_name = config.dem_name
        self.no_value = config.no_value
        self.stride = config.stride
        self.image_size = config.image_size
        self.batch_size = config.batch_size
        self.upsample_factor = 1
        self.tile_size = config.tile_size
        self.model = model
        # ---- additions
        self.save_tiles = bool(getattr(config, "save_tiles", False))
        self.seed = int(getattr(config, "seed", 0))
        self.rank, self.world_size = int(rank), int(world_size)
        self._groups_cfg = int(getattr(config, "groups_per_call", 0))
        self.mode = str(getattr(config, "mode", "faithful"))
        self._preprocess_cfg = bool(getattr(config, "preprocess", True))
        self.samples_per_patch = int(getattr(config, "samples_per_patch", 1))
        if self.samples_per_patch < 1:
            raise ValueError("samples_per_patch must be >= 1")
        if self.mode not in ("faithful", "dedup"):
            raise ValueError("mode must be 'faithful' or 'dedup'")
        self.reuse_spade = bool(getattr(config, "reuse_spade", True))
        self.blend = str(getattr(config, "blend", "exact"))
        if self.blend not in ("exact", "fast"):
            raise ValueError("blend must be 'exact' or 'fast'")
        if self.mode == "dedup" and self.save_tiles:
            raise ValueError("save_tiles needs mode='faithful': dedup mode has no per-tile accumulators to save")
        self._lib = _lib.lib()
        torch = _torch()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise _lib.MoonSRError("DEMSuperResolution runs only on a CUDA device")
        if self.device.index is not None and self.device.index != torch.cuda.current_device():
            # the C ABI launches on the CURRENT device's stream: pointers of another device would fault
            raise ValueError(f"device {self.device} is not the current CUDA device (cuda:{torch.cuda.current_device()}): "
                             "call torch.cuda.set_device first (one process per GPU)")
        self.dem = self.img = None
        self._row_offset = 0
        self.geo_transform = self.geo_projection = None
        self.plan: Optional[Plan] = None
        self.launches = 0              # kernel launches issued through the C ABI by this engine (excluding the model)
        self.model_launches = 0
        self.slots_executed = 0
        self._weights_dev = None
        self._weights_dev_f32 = None
        self._host_out = None

    # ---------------------------------------------------------------------------------------------------------------
    # inputs
    # ---------------------------------------------------------------------------------------------------------------
    def setRasters(self, dem, img, geo_transform=None, geo_projection=None, row_offset: int = 0,
                   full_height: Optional[int] = None) -> None:
        """In-memory equivalent of loadImages (process_full_tiles.py:158-182).  Accepts host numpy arrays (uploaded by
        padInputs) or float32 CUDA tensors already resident on this engine's device.  A rank may hold only the rows
        [row_offset, row_offset + rows) of a raster that is ``full_height`` rows tall (inputs loaded sharded); they
        must cover the rows its band of tiles reads."""
        def as_raster(a):
            if hasattr(a, "is_cuda"):
                if not (a.is_cuda and str(a.dtype) == "torch.float32"):
                    raise ValueError("tensor rasters must be float32 CUDA tensors")
                return a
            return np.ascontiguousarray(a, dtype=np.float32)
        dem, img = as_raster(dem), as_raster(img)
        if dem.ndim != 2 or tuple(img.shape) != tuple(dem.shape):
            raise ValueError("dem and ortho-image must be 2-D arrays of the same shape")
        self.dem, self.img = dem, img
        self.geo_transform, self.geo_projection = geo_transform, geo_projection
        self._row_offset = int(row_offset)
        h = int(full_height) if full_height is not None else int(dem.shape[0]) + self._row_offset
        if self._row_offset < 0 or self._row_offset + dem.shape[0] > h:
            raise ValueError("row_offset / full_height do not contain the given rows")
        self.dem_shape = self.img_shape = (h, int(dem.shape[1]))

    def setOwnedRows(self, dem_rows, img_rows, full_height: int) -> None:
        """Sharded loading: this rank supplies exactly the raster rows of its own band of tiles (float32 tensors, CUDA
        with NCCL / CPU with gloo); the I - S halo rows its border tiles read are fetched from the neighbouring ranks
        (distributed.exchange_halo_rows) -- the one exchange step of the path."""
        plan = Plan(int(full_height), int(dem_rows.shape[1]), self.image_size, self.stride, self.tile_size, self.batch_size)
        bounds = needs = None
        if self.mode == "dedup":
            bounds = [self._rows_dedup(plan, r)[0] for r in range(self.world_size)]
            needs = [self._rows_dedup(plan, r)[1] for r in range(self.world_size)]
        dem = exchange_halo_rows(dem_rows, plan, self.rank, self.world_size, bounds=bounds, needs=needs)
        img = exchange_halo_rows(img_rows, plan, self.rank, self.world_size, bounds=bounds, needs=needs)
        n0, _ = self.rowsNeeded(plan.height, plan.width)
        self.setRasters(dem if dem.is_cuda else dem.numpy(), img if img.is_cuda else img.numpy(), row_offset=n0,
                        full_height=plan.height)

    def ownedRows(self, height: int, width: int) -> Tuple[int, int]:
        """Raster rows [r0, r1) of this rank's own band of tiles (what setOwnedRows expects)."""
        plan = Plan(height, width, self.image_size, self.stride, self.tile_size, self.batch_size)
        if self.mode == "dedup":
            return self._rows_dedup(plan, self.rank)[0]
        return band_of_rank(plan, self.world_size, self.rank)[1:]

    def _rows_dedup(self, plan: Plan, rank: int):
        """Dedup mode: (raster rows this rank finalises, raster rows its patches read -- a superset)."""
        return plan.dedup_rows(self.world_size, rank)

    def loadImages(self) -> None:
        """process_full_tiles.py:158-182 -- band 1 of both GeoTIFFs as float32 plus the DEM's geo-referencing."""
        from . import geotiff
        img_path = os.path.join(self.folder_path, self.left_image_name)
        dem_path = os.path.join(self.folder_path, self.dem_name)
        if not os.path.exists(img_path):
            raise ValueError("The path given for the ortho-image does not exist. Provided path is: " + img_path)
        if not os.path.exists(dem_path):
            raise ValueError("The path given for the dem does not exist. Provided path is: " + dem_path)
        if self.world_size > 1 and not self._preprocess_cfg:
            # one process per GPU: decode only the strips / tiles of the rows this rank's band reads (preprocess, when on,
            # resamples across band boundaries and needs the whole DEM)
            h, w = geotiff.shape(dem_path)
            r0, r1 = self.rowsNeeded(h, w)
            img, _ = geotiff.read(img_path, rows=(r0, r1))
            dem, geo = geotiff.read(dem_path, rows=(r0, r1))
            self.setRasters(dem.astype(np.float32), img.astype(np.float32), geo, geo, row_offset=r0, full_height=h)
            return
        img, _ = geotiff.read(img_path)
        dem, geo = geotiff.read(dem_path)
        self.setRasters(dem.astype(np.float32), img.astype(np.float32), geo, geo)

    def interpolateMissingValues(self, data: np.ndarray, no_value: int, max_fill_area: int = 256) -> np.ndarray:
        """process_full_tiles.py:184-212 -- same signature; fills, in place, the connected invalid regions of ``data`` that
        are smaller than ``max_fill_area`` with the cubic interpolant through its valid pixels (host, scipy + cv2 like
        the reference; see preprocess.py)."""
        from . import preprocess as PP
        return PP._fill_block(data, no_value, max_fill_area)

    def fillNan(self, image: np.ndarray, no_value: int, tile_size: int = 1024, border: int = 128,
                max_fill_area: int = 256) -> np.ndarray:
        """process_full_tiles.py:214-224 -- same signature and defaults; block-wise interpolateMissingValues."""
        from . import preprocess as PP
        return PP.fill_small_holes(np.asarray(image), no_value, tile=tile_size, border=border,
                                   max_fill_area=max_fill_area)

    def preprocess(self) -> None:
        """process_full_tiles.py:226-244 -- the DEM is low-passed before it is tiled: 1/4 area resize, small-hole fill,
        1/4 area resize, bicubic back to (H, W), with no_value carried as NaN through the resampling.  Both resampling
        passes are kernels (msr_resize_area4 / msr_resize_cubic); see preprocess.py for the hole fill and for the two
        places where this differs from the reference's code (dead ortho half not run; (W, H) / (H, W) mix-up of :241
        not reproduced).  ``self.dem`` becomes a CUDA tensor.  No-op when the config says ``preprocess=False``."""
        if not self._preprocess_cfg:
            return
        torch = _torch()
        from . import preprocess as PP
        if self.dem is None:
            raise ValueError("no rasters loaded: call loadImages() or setRasters() first")
        if self._row_offset != 0 or int(self.dem.shape[0]) != self.dem_shape[0]:
            raise ValueError("preprocess() needs the whole DEM on this rank (it resamples across band boundaries)")
        dem = self.dem if torch.is_tensor(self.dem) else torch.from_numpy(self.dem).to(self.device)
        self.dem, n = PP.preprocess_dem(dem, self.no_value)
        self.launches += n

    # ---------------------------------------------------------------------------------------------------------------
    # padInputs (+ validity of every patch)
    # ---------------------------------------------------------------------------------------------------------------
    def _band(self) -> Tuple[List[Tuple[int, int]], int, int]:
        """Tiles of this rank and the canvas row range [c0, c1) they read."""
        plan = self.plan
        tiles, _, _ = band_of_rank(plan, self.world_size, self.rank)
        if not tiles:
            return [], 0, 0
        y0 = min(yy for _, yy in tiles)
        y1 = max(yy for _, yy in tiles)
        return tiles, y0, min(plan.canvas_h, y1 + plan.tile_size + 2 * plan.off)

    def padInputs(self) -> None:
        """process_full_tiles.py:246-267 on the device, for this rank's band of canvas rows."""
        torch = _torch()
        if self.dem is None or self.img is None:
            raise ValueError("no rasters loaded: call loadImages() or setRasters() first")
        h, w = self.dem_shape
        self.plan = plan = Plan(h, w, self.image_size, self.stride, self.tile_size, self.batch_size)
        self.pad_x, self.pad_y = plan.pad_x, plan.pad_y
        self.dem_padded_shape = self.img_padded_shape = (plan.canvas_h, plan.canvas_w)
        if self.mode == "dedup":
            self._dband = plan.dedup_band(self.world_size, self.rank)
            self.my_tiles = []
            c0, c1 = self._dband.read
        else:
            self.my_tiles, c0, c1 = self._band()
        self._c0, self._ch = c0, max(c1 - c0, 0)
        st = _lib.stream_ptr()
        nv = float(np.float32(self.no_value))
        if self._ch > 0:
            # raster rows that fall inside canvas rows [c0, c1): canvas row = raster row + off
            r0, r1 = max(0, c0 - plan.off), min(h, c1 - plan.off)
            self._r0 = r0
            ro = self._row_offset
            if r1 > r0 and (r0 < ro or r1 > ro + self.dem.shape[0]):
                raise ValueError(f"rank {self.rank} needs raster rows [{r0}, {r1}) but holds "
                                 f"[{ro}, {ro + self.dem.shape[0]})")
            def rows_on_device(a):
                if torch.is_tensor(a):
                    return a[r0 - ro:r1 - ro].contiguous()
                return torch.from_numpy(a[r0 - ro:r1 - ro]).to(self.device, non_blocking=True)
            d_dem, d_img = rows_on_device(self.dem), rows_on_device(self.img)
            self.dem_padded = torch.empty((self._ch, plan.canvas_w), dtype=torch.float32, device=self.device)
            self.img_padded = torch.empty_like(self.dem_padded)
            if r1 > r0:
                _lib.check(self._lib.msr_pad_inputs(d_dem.data_ptr(), d_img.data_ptr(), r1 - r0, w,
                                                    self.dem_padded.data_ptr(), self.img_padded.data_ptr(), self._ch,
                                                    plan.canvas_w, r0 + plan.off - c0, plan.off, nv, st),
                           "msr_pad_inputs")
                self.launches += 1
            else:
                self.dem_padded.fill_(nv)
                self.img_padded.fill_(nv)
            self._sat = torch.empty((self._ch + 1, plan.canvas_w + 1), dtype=torch.int32, device=self.device)
            _lib.check(self._lib.msr_validity_sat(self.img_padded.data_ptr(), self.dem_padded.data_ptr(), self._ch,
                                                  plan.canvas_w, nv, self._sat.data_ptr(), st), "msr_validity_sat")
            self.launches += 2
        # the reference releases the originals here (:265-266)
        self.dem = None
        self.img = None
        if self.mode == "dedup":
            self._plan_band()
            self._out_r0, self._out_r1 = self._dband.raster_rows(self._dband.out, plan.off, h)
            rows = self._out_r1 - self._out_r0
            self.mean_out = torch.empty((rows, w), dtype=torch.float32, device=self.device)
            self.std_out = torch.empty((rows, w), dtype=torch.float32, device=self.device)
            self.good_out = torch.empty((rows, w), dtype=torch.uint8, device=self.device)
            # rebuildTile's accumulators (:386-390) for the canvas rows this band's patches touch
            self._acc = torch.zeros((3, self._ch, plan.canvas_w), dtype=torch.float32, device=self.device)
            return
        self._plan_tiles()
        # output rasters of this rank's band (rows [out_r0, out_r1) of the (H, W) result)
        if self.my_tiles:
            self._out_r0 = min(yy for _, yy in self.my_tiles)
            self._out_r1 = min(h, max(yy for _, yy in self.my_tiles) + plan.tile_size)
        else:
            self._out_r0 = self._out_r1 = 0
        rows = self._out_r1 - self._out_r0
        self.mean_out = torch.zeros((rows, w), dtype=torch.float32, device=self.device)
        self.std_out = torch.zeros((rows, w), dtype=torch.float32, device=self.device)
        self.good_out = torch.zeros((rows, w), dtype=torch.uint8, device=self.device)
        return

    def _plan_tiles(self) -> None:
        """Validity of every patch of every tile of the band (getPatch, process_full_tiles.py:286-292) in one kernel
        call, then the host-side batch plan (process_full_tiles.py:459-474)."""
        torch = _torch()
        plan = self.plan
        self._tile_plan: Dict[Tuple[int, int], dict] = {}
        if not self.my_tiles:
            return
        origins = [plan.tile_patch_origins(px, py) for px, py in self.my_tiles]
        all_xy = np.concatenate(origins, axis=0)
        local = all_xy.copy()
        local[:, 1] -= self._c0
        d_xy = torch.from_numpy(local).to(self.device)
        d_valid = torch.empty((local.shape[0],), dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.msr_patch_validity(self._sat.data_ptr(), self._ch, plan.canvas_w, d_xy.data_ptr(),
                                                local.shape[0], plan.image_size, d_valid.data_ptr(),
                                                _lib.stream_ptr()), "msr_patch_validity")
        self.launches += 1
        valid = d_valid.cpu().numpy().astype(bool)
        g2 = plan.lattice_side ** 2
        for t, (px, py) in enumerate(self.my_tiles):
            v = valid[t * g2:(t + 1) * g2]
            idx = np.nonzero(v)[0]
            lattice = np.full((g2,), -1, np.int32)
            lattice[idx] = np.arange(idx.size, dtype=np.int32)
            self._tile_plan[(px, py)] = dict(xy=origins[t][idx], lattice=lattice, n_valid=int(idx.size))

    # ---------------------------------------------------------------------------------------------------------------
    # reference-compatible per-patch helpers
    # ---------------------------------------------------------------------------------------------------------------
    def getPatch(self, px: int, py: int):
        """process_full_tiles.py:269-293 -- (valid, img_patch, dem_patch) as numpy arrays (a D2H copy; the fast path
        never calls this)."""
        i = self.image_size
        ly = py - self._c0
        img = self.img_padded[ly:ly + i, px:px + i].cpu().numpy()
        dem = self.dem_padded[ly:ly + i, px:px + i].cpu().numpy()
        valid = not ((img <= self.no_value).any() or (dem <= self.no_value).any())
        return valid, img, dem

    def normalize(self, img_patch: np.ndarray, dem_patch: np.ndarray):
        """process_full_tiles.py:295-311 through the device kernel (patches are uploaded as a 1-patch canvas)."""
        torch = _torch()
        i = self.image_size
        d_img = torch.from_numpy(np.ascontiguousarray(img_patch, dtype=np.float32)).to(self.device)
        d_dem = torch.from_numpy(np.ascontiguousarray(dem_patch, dtype=np.float32)).to(self.device)
        xy = torch.zeros((1, 2), dtype=torch.int32, device=self.device)
        out = torch.empty((1, i, i, 2), dtype=torch.float32, device=self.device)
        mm = torch.empty((1, 4), dtype=torch.float32, device=self.device)
        part = torch.empty((1 * 32 * 4,), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.msr_gather_normalize(d_img.data_ptr(), d_dem.data_ptr(), i, i, xy.data_ptr(), 1, i,
                                                  out.data_ptr(), mm.data_ptr(), part.data_ptr(), _lib.stream_ptr()),
                   "msr_gather_normalize")
        self.launches += 2
        mmh = mm.cpu().numpy()[0]
        return out.cpu().numpy()[0], (np.float32(mmh[2]), np.float32(mmh[3]))

    def generateTileList(self) -> List[Tuple[int, int]]:
        """process_full_tiles.py:313-325 (all tiles, not only this rank's)."""
        if self.plan is not None:
            return self.plan.tiles()
        return [(xx, yy) for yy in range(0, self.dem_shape[0], self.tile_size)
                for xx in range(0, self.dem_shape[1], self.tile_size)]

    def processBatch(self, batch, batch_index, patches: dict) -> None:
        """process_full_tiles.py:327-345, host-array form (compatibility; processTile does not use it)."""
        pred_dems = self.model(np.array(batch), training=False)
        pred_dems = np.array(pred_dems)[:, :, :, -1] + 0.5
        for pred_dem, pred_idx in zip(pred_dems, batch_index):
            if pred_idx != PAD_SLOT:
                patches[pred_idx] = pred_dem

    def makeGaussianKernel(self) -> np.ndarray:
        """process_full_tiles.py:347-361 -- float64 host table (built once per run, I*I doubles)."""
        i = self.image_size
        s = i / 5

        def gaus2d(x=0, y=0, mx=0, my=0, sx=1, sy=1):
            return 1. / (2. * np.pi * sx * sy) * np.exp(-((x - mx) ** 2. / (2. * sx ** 2.) + (y - my) ** 2. / (2. * sy ** 2.)))
        x = np.linspace(-i / 2, i / 2, i)
        y = np.linspace(-i / 2, i / 2, i)
        x, y = np.meshgrid(x, y)
        kern = gaus2d(x, y, sx=s, sy=s)
        return (kern - kern.min()) / (kern.max() - kern.min())

    def _blend_weights(self):
        """makeGaussianKernel() + 1e-7, purge-cropped (process_full_tiles.py:391-393), resident on the device."""
        if self._weights_dev is None:
            torch = _torch()
            p = self.image_size // 16
            w = self.makeGaussianKernel() + 1e-7
            w = np.ascontiguousarray(w[p:-p, p:-p])
            self._weights_dev = torch.from_numpy(w).to(self.device)
        return self._weights_dev

    def _blend_weights_f32(self):
        if self._weights_dev_f32 is None:
            self._weights_dev_f32 = self._blend_weights().to(_torch().float32).contiguous()
        return self._weights_dev_f32

    def _blend_weights_separable(self):
        """makeGaussianKernel (:347-361) is A * e(x) * e(y) before its min-max normalisation, so the purge-cropped weight
        is e[ry] * e[rx] * c1 + c0: (device float32 e over the cropped axis, c1, c0), constants formed in float64."""
        if getattr(self, "_weights_sep", None) is None:
            torch = _torch()
            i = self.image_size
            p, s = i // 16, i / 5
            x = np.linspace(-i / 2, i / 2, i)
            e = np.exp(-(x ** 2.) / (2. * s ** 2.))
            a = 1. / (2. * np.pi * s * s)
            kmin, kmax = a * e.min() ** 2, a * e.max() ** 2
            c1, c0 = a / (kmax - kmin), 1e-7 - kmin / (kmax - kmin)
            e_dev = torch.from_numpy(np.ascontiguousarray(e[p:i - p], dtype=np.float32)).to(self.device)
            self._weights_sep = (e_dev, float(c1), float(c0))
        return self._weights_sep

    def _fast_blend_ok(self) -> bool:
        """The float32 / 128-bit blend kernels need 4-pixel groups that share their patches and aligned rows."""
        return (self.blend == "fast" and self.image_size % 64 == 0 and self.stride % 4 == 0 and
                self.plan is not None and self.plan.width % 4 == 0)

    # ---------------------------------------------------------------------------------------------------------------
    # blend
    # ---------------------------------------------------------------------------------------------------------------
    def _blend(self, ptrs, f64flags, lohi, pxy, n, lattice, add_half, mean, std, good, pitch, rows, cols,
               pred=None) -> None:
        plan_i, plan_s, plan_t = self.image_size, self.stride, self.tile_size
        g = -(-(plan_t + plan_i - plan_s) // plan_s)
        if (pred is not None and lattice is not None and f64flags is None and self._fast_blend_ok() and pitch % 4 == 0
                and mean % 16 == 0 and std % 16 == 0 and good % 4 == 0):
            e_dev, c1, c0 = self._blend_weights_separable()
            _lib.check(self._lib.msr_blend_tile_fast(pred.data_ptr(), _lib.ptr(lohi), n, _lib.ptr(lattice), g,
                                                     e_dev.data_ptr(), c1, c0, plan_i, plan_s, plan_t,
                                                     int(add_half), float(np.float32(self.no_value)), mean, std, good,
                                                     pitch, rows, cols, _lib.stream_ptr()), "msr_blend_tile_fast")
            self.launches += 1
            return
        _lib.check(self._lib.msr_blend_tile(_lib.ptr(ptrs), _lib.ptr(f64flags), _lib.ptr(lohi), _lib.ptr(pxy), n,
                                            _lib.ptr(lattice), g, self._blend_weights().data_ptr(), plan_i, plan_s,
                                            plan_t, int(add_half), float(np.float32(self.no_value)), mean, std, good,
                                            pitch, rows, cols, _lib.stream_ptr()), "msr_blend_tile")
        self.launches += 1

    def rebuildTile(self, generated_dems: dict, generated_minmax: dict):
        """process_full_tiles.py:363-414 -- same signature: dicts keyed by (x, y) relative to the tile origin, values
        (I, I) predictions (after the + 0.5) and (min, max); returns (mean, std, good) of the full T x T tile as numpy
        arrays.  The arithmetic runs in msr_blend_tile."""
        torch = _torch()
        t, i = self.tile_size, self.image_size
        keys = list(generated_dems.keys())
        n = len(keys)
        mean = torch.empty((t, t), dtype=torch.float32, device=self.device)
        std = torch.empty_like(mean)
        good = torch.empty((t, t), dtype=torch.uint8, device=self.device)
        keep, ptrs, flags = [], [], []
        for k in keys:
            a = np.asarray(generated_dems[k])
            if a.dtype != np.float64:
                a = a.astype(np.float32, copy=False)
            d = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            keep.append(d)
            ptrs.append(d.data_ptr())
            flags.append(1 if a.dtype == np.float64 else 0)
        if n:
            d_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            d_flags = torch.tensor(flags, dtype=torch.uint8, device=self.device)
            lohi = np.array([[np.float32(generated_minmax[k][0]), np.float32(generated_minmax[k][1])] for k in keys],
                            dtype=np.float32)
            d_lohi = torch.from_numpy(lohi).to(self.device)
            d_xy = torch.tensor([[k[0], k[1]] for k in keys], dtype=torch.int32, device=self.device)
        else:
            d_ptrs = d_flags = d_lohi = d_xy = None
        self._blend(d_ptrs, d_flags, d_lohi, d_xy, n, None, 0, mean.data_ptr(), std.data_ptr(), good.data_ptr(), t, t, t)
        torch.cuda.current_stream().synchronize()
        return mean.cpu().numpy(), std.cpu().numpy(), good.cpu().numpy()

    def saveTile(self, mean: np.ndarray, std: np.ndarray, good: np.ndarray, name: str) -> None:
        """process_full_tiles.py:416-429 -- same directory / file names."""
        from . import geotiff
        d = os.path.join(self.save_path, "tile_" + name)
        os.makedirs(d, exist_ok=True)
        geotiff.write(os.path.join(d, "tile_" + name + "_mean.tif"), np.asarray(mean))
        geotiff.write(os.path.join(d, "tile_" + name + "_std.tif"), np.asarray(std))
        geotiff.write(os.path.join(d, "tile_" + name + "_correct.tif"), np.asarray(good))

    # ---------------------------------------------------------------------------------------------------------------
    # processTile
    # ---------------------------------------------------------------------------------------------------------------
    def _gather(self, d_slot_xy, n, out, minmax, partial) -> None:
        plan = self.plan
        _lib.check(self._lib.msr_gather_normalize(self.img_padded.data_ptr(), self.dem_padded.data_ptr(), self._ch,
                                                  plan.canvas_w, d_slot_xy.data_ptr(), n, plan.image_size,
                                                  out.data_ptr(), minmax.data_ptr(), partial.data_ptr(),
                                                  _lib.stream_ptr()), "msr_gather_normalize")
        self.launches += 2

    def _sampler_noise(self, slots: int, eps, key: int, repeats: int = 1):
        """Standard-normal draws for the Gaussian sampler (sampling.py:13-16) of a GauGAN device model, shape (slots, 256)
        -- (repeats, slots, 256) in repeated-sample mode: the caller's ``eps`` (parity tests) or a stream seeded by the
        config's seed and ``key`` (the reference's tf.random.normal is unseeded).  None for models without a sampler."""
        if getattr(self.model, "arch", None) != "spade":
            return None
        torch = _torch()
        shape = (slots, 256) if repeats == 1 else (repeats, slots, 256)
        if eps is None:
            gen = torch.Generator(device=self.device)
            gen.manual_seed((self.seed * 1000003 + key) & 0x7FFFFFFFFFFF)
            return torch.randn(shape, generator=gen, dtype=torch.float32, device=self.device)
        d_eps = torch.as_tensor(np.ascontiguousarray(eps, dtype=np.float32)).to(self.device)
        if tuple(d_eps.shape) != shape:
            raise ValueError(f"eps must have shape {shape}")
        return d_eps

    def processTile(self, px: int, py: int, eps=None) -> None:
        """process_full_tiles.py:431-479.  ``eps`` optionally supplies the sampler noise for every slot of the tile,
        shape (slots, 256) (parity tests); by default a seeded per-tile stream is drawn on the device."""
        torch = _torch()
        plan = self.plan
        if (px, py) not in self._tile_plan:
            raise ValueError(f"tile ({px}, {py}) is not owned by rank {self.rank}")
        if self.samples_per_patch > 1:
            self._process_tile_repeats(px, py, eps)
            return
        tp = self._tile_plan[(px, py)]
        i, b, t = plan.image_size, plan.batch_size, plan.tile_size
        n_valid = tp["n_valid"]
        slots = plan.batch_slots(n_valid)
        rows, cols = plan.tile_window(px, py)
        dev = self.device
        out_off = (py - self._out_r0) * plan.width + px
        mean_p = self.mean_out.data_ptr() + 4 * out_off
        std_p = self.std_out.data_ptr() + 4 * out_off
        good_p = self.good_out.data_ptr() + out_off
        tile_bufs = None
        if self.save_tiles:
            tile_bufs = (torch.empty((t, t), dtype=torch.float32, device=dev),
                         torch.empty((t, t), dtype=torch.float32, device=dev),
                         torch.empty((t, t), dtype=torch.uint8, device=dev))
        if n_valid == 0:
            self._blend(None, None, None, None, 0, None, 0, mean_p, std_p, good_p, plan.width, rows, cols)
            if tile_bufs:
                self._blend(None, None, None, None, 0, None, 0, tile_bufs[0].data_ptr(), tile_bufs[1].data_ptr(),
                            tile_bufs[2].data_ptr(), t, t, t)
                self._save_tile_bufs(tile_bufs, px, py)
            return
        slot_xy = np.full((slots, 2), -1, np.int32)
        slot_xy[:n_valid] = tp["xy"]
        slot_xy[:n_valid, 1] -= self._c0
        d_slot_xy = torch.from_numpy(slot_xy).to(dev, non_blocking=True)
        key_xy = tp["xy"].copy()
        key_xy[:, 0] -= px
        key_xy[:, 1] -= py
        d_key_xy = torch.from_numpy(key_xy).to(dev, non_blocking=True)
        d_lattice = torch.from_numpy(tp["lattice"]).to(dev, non_blocking=True)
        minmax = torch.empty((slots, 4), dtype=torch.float32, device=dev)
        device_model = isinstance(self.model, (_Generator, IdentityModel))
        if device_model:
            groups = self._groups_cfg or self.model.max_groups
            groups = max(1, min(groups, self.model.max_groups))
            chunk = groups * b
            pred = torch.empty((slots, i, i), dtype=torch.float32, device=dev)
            src = torch.empty((min(chunk, slots), i, i, 2), dtype=torch.float32, device=dev)
            partial = torch.empty((min(chunk, slots) * 32 * 4,), dtype=torch.float32, device=dev)
            d_eps = self._sampler_noise(slots, eps, py * 131071 + px)
            for s0 in range(0, slots, chunk):
                n = min(chunk, slots - s0)
                self._gather(d_slot_xy[s0:s0 + n], n, src, minmax[s0:s0 + n], partial)
                self.model.forward_device(src[:n], pred[s0:s0 + n], None if d_eps is None else d_eps[s0:s0 + n],
                                          n // b)
                self.model_launches += self.model.last_launch_count
            ptrs = pred.data_ptr() + torch.arange(n_valid, dtype=torch.int64, device=dev) * (i * i * 4)
            flags, add_half, keep = None, 1, [pred]
            pred_base = pred
        else:
            ptrs, flags, keep = self._run_host_model(d_slot_xy, slots, n_valid, minmax)
            add_half = 0
            pred_base = None
        lohi = minmax[:n_valid, 2:4].contiguous()
        self.slots_executed += slots
        self._blend(ptrs, flags, lohi, d_key_xy, n_valid, d_lattice, add_half, mean_p, std_p, good_p, plan.width, rows,
                    cols, pred=pred_base)
        if tile_bufs:
            self._blend(ptrs, flags, lohi, d_key_xy, n_valid, d_lattice, add_half, tile_bufs[0].data_ptr(),
                        tile_bufs[1].data_ptr(), tile_bufs[2].data_ptr(), t, t, t, pred=pred_base)
            self._save_tile_bufs(tile_bufs, px, py)
        del keep

    def _run_host_model(self, d_slot_xy, slots, n_valid, minmax):
        """Generic ``model=`` callable: numpy batches on the host, exactly the reference's data flow
        (process_full_tiles.py:327-345, 468-474)."""
        torch = _torch()
        plan = self.plan
        i, b, dev = plan.image_size, plan.batch_size, self.device
        src = torch.empty((b, i, i, 2), dtype=torch.float32, device=dev)
        partial = torch.empty((b * 32 * 4,), dtype=torch.float32, device=dev)
        keep, ptrs, flags = [], [], []
        for s0 in range(0, slots, b):
            self._gather(d_slot_xy[s0:s0 + b], b, src, minmax[s0:s0 + b], partial)
            batch = src.cpu().numpy()
            n_real = min(b, n_valid - s0)
            if n_real < b:
                batch = batch.astype(np.float64)          # np.array(batch) with float64 zero pads (:472) promotes
            pred = self.model(batch, training=False)
            pred = np.array(pred)[:, :, :, -1] + 0.5       # :340
            if pred.dtype != np.float64:
                pred = pred.astype(np.float32, copy=False)
            d = torch.from_numpy(np.ascontiguousarray(pred[:n_real])).to(dev)
            keep.append(d)
            itemsize = 8 if pred.dtype == np.float64 else 4
            for k in range(n_real):
                ptrs.append(d.data_ptr() + k * i * i * itemsize)
                flags.append(1 if itemsize == 8 else 0)
        d_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        d_flags = torch.tensor(flags, dtype=torch.uint8, device=dev)
        return d_ptrs, d_flags, keep

    # ---------------------------------------------------------------------------------------------------------------
    # dedup mode (mode B)
    # ---------------------------------------------------------------------------------------------------------------
    def _plan_band(self) -> None:
        """Validity (getPatch, process_full_tiles.py:286-292) of every lattice position of this rank's band and the
        visit order of the valid ones: y outer, x inner (:453-454), over the whole band instead of tile by tile."""
        torch = _torch()
        plan, band = self.plan, self._dband
        rows, gx, s = band.j1 - band.j0, band.gx, plan.stride
        self._band_plan = dict(xy=np.zeros((0, 2), np.int32), lattice=np.full((max(rows * gx, 1),), -1, np.int32),
                               n_valid=0, row_of=np.zeros((0,), np.int32))
        if rows <= 0 or self._ch <= 0:
            return
        xs = np.tile(np.arange(gx, dtype=np.int32) * s, rows)
        ys = np.repeat((np.arange(rows, dtype=np.int32) + band.j0) * s, gx)
        xy = np.stack([xs, ys - self._c0], axis=1).astype(np.int32)          # rows local to the canvas band
        d_xy = torch.from_numpy(xy).to(self.device)
        d_valid = torch.empty((xy.shape[0],), dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.msr_patch_validity(self._sat.data_ptr(), self._ch, plan.canvas_w, d_xy.data_ptr(),
                                                xy.shape[0], plan.image_size, d_valid.data_ptr(), _lib.stream_ptr()),
                   "msr_patch_validity")
        self.launches += 1
        idx = np.nonzero(d_valid.cpu().numpy())[0]
        lattice = np.full((rows * gx,), -1, np.int32)
        lattice[idx] = np.arange(idx.size, dtype=np.int32)
        self._band_plan = dict(xy=xy[idx], lattice=lattice, n_valid=int(idx.size), row_of=(idx // gx).astype(np.int32))

    def _accumulate(self, pred, lohi, k0: int, n: int, gy_lo: int, gy_hi: int, add_half: int, row_lo: int,
                    row_hi: int, target=None) -> None:
        """msr_blend_accumulate into the band accumulators (dedup mode) or into ``target`` = (acc (3, rows, pitch),
        d_lattice, GY, GX, lattice_y0, acc_y0): rebuildTile's loop body for patches [k0, k0 + n) of that lattice."""
        plan = self.plan
        if target is None:
            band = self._dband
            target = (self._acc, self._d_lattice, band.j1 - band.j0, band.gx, band.j0 * plan.stride, self._c0)
        acc, d_lattice, gy, gx, lattice_y0, acc_y0 = target
        if (self._fast_blend_ok() and int(acc.shape[2]) % 4 == 0 and pred.dtype == _torch().float32
                and pred.data_ptr() % 16 == 0):
            _lib.check(self._lib.msr_blend_accumulate_fast(pred.data_ptr(), lohi.data_ptr(), k0, n, d_lattice.data_ptr(),
                                                           gy, gx, gy_lo, gy_hi, lattice_y0,
                                                           self._blend_weights_f32().data_ptr(), plan.image_size,
                                                           plan.stride, int(add_half), acc[0].data_ptr(),
                                                           acc[1].data_ptr(), acc[2].data_ptr(), int(acc.shape[2]),
                                                           acc_y0, int(acc.shape[1]), int(acc.shape[2]), row_lo, row_hi,
                                                           _lib.stream_ptr()), "msr_blend_accumulate_fast")
            self.launches += 1
            return
        _lib.check(self._lib.msr_blend_accumulate(pred.data_ptr(), lohi.data_ptr(), k0, n, d_lattice.data_ptr(), gy, gx,
                                                  gy_lo, gy_hi, lattice_y0, self._blend_weights().data_ptr(),
                                                  plan.image_size, plan.stride, int(add_half), acc[0].data_ptr(),
                                                  acc[1].data_ptr(), acc[2].data_ptr(), int(acc.shape[2]), acc_y0,
                                                  int(acc.shape[1]), int(acc.shape[2]), row_lo, row_hi,
                                                  _lib.stream_ptr()), "msr_blend_accumulate")
        self.launches += 1

    def _generate_repeats(self, src, n: int, d_eps, s0: int, device_model: bool, n_real: int,
                          promote_pads: bool = True):
        """Repeated-sample mode: the ``n`` slots in ``src`` generated ``samples_per_patch`` times -> float32 CUDA tensor
        (R, n, I, I) and the flag telling the blend whether ``+ 0.5`` (process_full_tiles.py:340) is still to be added.
        Host plug-ins are called batch by batch, repetition by repetition; tile by tile they see the reference's batch
        dtype (float64 when the batch carries zero pads, :472), in dedup mode float32."""
        torch = _torch()
        plan = self.plan
        i, b, r_ = plan.image_size, plan.batch_size, self.samples_per_patch
        preds = torch.empty((r_, n, i, i), dtype=torch.float32, device=self.device)
        if device_model:
            for r in range(r_):
                phase = (_lib.REPEAT_FIRST if r == 0 else _lib.REPEAT_NEXT) if self.reuse_spade else _lib.REPEAT_NONE
                self.model.forward_device(src[:n], preds[r], None if d_eps is None else d_eps[r, s0:s0 + n], n // b,
                                          repeat_phase=phase)
                self.model_launches += self.model.last_launch_count
            return preds, 1
        host = src[:n].cpu().numpy()
        for j in range(0, n, b):
            batch = host[j:j + b]
            if promote_pads and j + b > n_real:
                batch = batch.astype(np.float64)
            for r in range(r_):
                y = self.model(batch, training=False)
                y = (np.array(y)[:, :, :, -1] + 0.5).astype(np.float32)
                preds[r, j:j + b].copy_(torch.from_numpy(np.ascontiguousarray(y)))
        return preds, 0

    def _process_tile_repeats(self, px: int, py: int, eps=None) -> None:
        """processTile in repeated-sample mode (samples_per_patch = R > 1, beyond the reference): every batch of the
        tile is generated R times; the tile's accumulators (rebuildTile :386-390) are updated batch by batch, repetition
        by repetition, patch by patch with msr_blend_accumulate, then finalised / pasted / cropped like any tile.
        ``eps``: optional (R, slots, 256)."""
        torch = _torch()
        plan = self.plan
        if (px, py) not in self._tile_plan:
            raise ValueError(f"tile ({px}, {py}) is not owned by rank {self.rank}")
        tp = self._tile_plan[(px, py)]
        i, b, t, dev = plan.image_size, plan.batch_size, plan.tile_size, self.device
        r_ = self.samples_per_patch
        n_valid = tp["n_valid"]
        slots = plan.batch_slots(n_valid)
        rows, cols = plan.tile_window(px, py)
        side = t + 2 * plan.off                                                # :386
        acc = torch.zeros((3, side, side), dtype=torch.float32, device=dev)
        g = plan.lattice_side
        if n_valid:
            d_lattice = torch.from_numpy(tp["lattice"]).to(dev, non_blocking=True)
            target = (acc, d_lattice, g, g, 0, 0)
            slot_xy = np.full((slots, 2), -1, np.int32)
            slot_xy[:n_valid] = tp["xy"]
            slot_xy[:n_valid, 1] -= self._c0
            d_slot_xy = torch.from_numpy(slot_xy).to(dev, non_blocking=True)
            row_of = ((tp["xy"][:, 1] - py) // plan.stride).astype(np.int32)
            minmax = torch.empty((slots, 4), dtype=torch.float32, device=dev)
            device_model = isinstance(self.model, (_Generator, IdentityModel))
            groups = max(1, min(self._groups_cfg or getattr(self.model, "max_groups", 1),
                                getattr(self.model, "max_groups", 1))) if device_model else 1
            chunk = groups * b
            src = torch.empty((min(chunk, slots), i, i, 2), dtype=torch.float32, device=dev)
            partial = torch.empty((min(chunk, slots) * 32 * 4,), dtype=torch.float32, device=dev)
            d_eps = self._sampler_noise(slots, eps, py * 131071 + px, r_) if device_model else None
            for s0 in range(0, slots, chunk):
                n = min(chunk, slots - s0)
                self._gather(d_slot_xy[s0:s0 + n], n, src, minmax[s0:s0 + n], partial)
                preds, add_half = self._generate_repeats(src, n, d_eps, s0, device_model, n_valid - s0)
                self.slots_executed += n * r_
                for j in range(0, n, b):
                    n_real = min(b, n_valid - (s0 + j))
                    if n_real <= 0:
                        break
                    lohi = minmax[s0 + j:s0 + j + n_real, 2:4].contiguous()
                    gy_lo, gy_hi = int(row_of[s0 + j]), int(row_of[s0 + j + n_real - 1])
                    for r in range(r_):
                        self._accumulate(preds[r, j:j + n_real], lohi, s0 + j, n_real, gy_lo, gy_hi, add_half, 0, side,
                                         target)
        first = (plan.off * side + plan.off) * 4
        out_off = (py - self._out_r0) * plan.width + px
        nv = float(np.float32(self.no_value))
        windows = [(self.mean_out.data_ptr() + 4 * out_off, self.std_out.data_ptr() + 4 * out_off,
                    self.good_out.data_ptr() + out_off, plan.width, rows, cols)]
        tile_bufs = None
        if self.save_tiles:
            tile_bufs = (torch.empty((t, t), dtype=torch.float32, device=dev),
                         torch.empty((t, t), dtype=torch.float32, device=dev),
                         torch.empty((t, t), dtype=torch.uint8, device=dev))
            windows.append((tile_bufs[0].data_ptr(), tile_bufs[1].data_ptr(), tile_bufs[2].data_ptr(), t, t, t))
        for mean_p, std_p, good_p, pitch, wr, wc in windows:                  # :404-413 + paste / crop (:541-545)
            _lib.check(self._lib.msr_blend_finalize(acc[0].data_ptr() + first, acc[1].data_ptr() + first,
                                                    acc[2].data_ptr() + first, side, wr, wc, nv, mean_p, std_p, good_p,
                                                    pitch, _lib.stream_ptr()), "msr_blend_finalize")
            self.launches += 1
        if tile_bufs:
            self._save_tile_bufs(tile_bufs, px, py)

    def bandSlots(self) -> int:
        """Generator slots this rank executes in dedup mode (valid patches rounded up to whole batches)."""
        return self.plan.batch_slots(self._band_plan["n_valid"])

    def processBandMain(self, eps=None) -> None:
        """Dedup mode, phase 1: generate every valid patch of the band once, in visit order, and blend it into the
        accumulators -- except the rows that the previous rank's patches also touch (``seam_in``): rebuildTile's update
        is order-dependent (process_full_tiles.py:400-402), so those rows wait for the neighbour's accumulator strip
        and the predictions that reach into them are kept for phase 2.  ``eps``: optional (bandSlots(), 256) noise."""
        torch = _torch()
        plan, band, bp = self.plan, self._dband, self._band_plan
        i, b, dev = plan.image_size, plan.batch_size, self.device
        n_valid = bp["n_valid"]
        slots = plan.batch_slots(n_valid)
        self._retained = []
        self._d_lattice = torch.from_numpy(bp["lattice"]).to(dev, non_blocking=True)
        if n_valid == 0:
            return
        main_lo = band.seam_in[1] if band.seam_in is not None else 0
        # lattice rows (relative to the band) whose patches reach above main_lo
        seam_rows = 0 if band.seam_in is None else -(-(plan.off - 2 * plan.purge) // plan.stride)
        slot_xy = np.full((slots, 2), -1, np.int32)
        slot_xy[:n_valid] = bp["xy"]
        d_slot_xy = torch.from_numpy(slot_xy).to(dev, non_blocking=True)
        minmax = torch.empty((slots, 4), dtype=torch.float32, device=dev)
        row_of = bp["row_of"]
        device_model = isinstance(self.model, (_Generator, IdentityModel))
        if device_model:
            groups = max(1, min(self._groups_cfg or self.model.max_groups, self.model.max_groups))
            chunk = groups * b
            add_half = 1
        else:
            chunk, add_half = b, 0
        src = torch.empty((min(chunk, slots), i, i, 2), dtype=torch.float32, device=dev)
        partial = torch.empty((min(chunk, slots) * 32 * 4,), dtype=torch.float32, device=dev)
        r_ = self.samples_per_patch
        d_eps = self._sampler_noise(slots, eps, band.j0 * 8191 + 17, r_) if device_model else None
        for s0 in range(0, slots, chunk):
            n = min(chunk, slots - s0)
            n_real = min(n, n_valid - s0)
            self._gather(d_slot_xy[s0:s0 + n], n, src, minmax[s0:s0 + n], partial)
            if r_ > 1:   # repeated-sample mode: blend batch by batch, repetition by repetition, patch by patch
                preds, half = self._generate_repeats(src, n, d_eps, s0, device_model, n_real, promote_pads=False)
                self.slots_executed += n * r_
                for j in range(0, n_real, b):
                    nr = min(b, n_real - j)
                    lohi = minmax[s0 + j:s0 + j + nr, 2:4].contiguous()
                    gy_lo, gy_hi = int(row_of[s0 + j]), int(row_of[s0 + j + nr - 1])
                    for r in range(r_):
                        pr = preds[r, j:j + nr]
                        self._accumulate(pr, lohi, s0 + j, nr, gy_lo, gy_hi, half, main_lo, plan.canvas_h)
                        if gy_lo < seam_rows:
                            self._retained.append((pr, lohi, s0 + j, nr, gy_lo, min(gy_hi, seam_rows - 1), half))
                continue
            if device_model:
                pred = torch.empty((n, i, i), dtype=torch.float32, device=dev)
                self.model.forward_device(src[:n], pred, None if d_eps is None else d_eps[s0:s0 + n], n // b)
                self.model_launches += self.model.last_launch_count
            else:
                y = self.model(src[:n].cpu().numpy(), training=False)                 # plug-in contract (:338)
                y = (np.array(y)[:, :, :, -1] + 0.5).astype(np.float32, copy=False)   # :340
                pred = torch.from_numpy(np.ascontiguousarray(y)).to(dev)
            self.slots_executed += n
            lohi = minmax[s0:s0 + n_real, 2:4].contiguous()
            gy_lo, gy_hi = int(row_of[s0]), int(row_of[s0 + n_real - 1])
            self._accumulate(pred, lohi, s0, n_real, gy_lo, gy_hi, add_half, main_lo, plan.canvas_h)
            if gy_lo < seam_rows:
                self._retained.append((pred, lohi, s0, n_real, gy_lo, min(gy_hi, seam_rows - 1), add_half))

    def seamOut(self):
        """Accumulator rows this rank's patches share with the NEXT rank's: a (3, rows, canvas_w) float32 view
        (w_sum, mean, S), complete after processBandMain; None for the last rank."""
        band = self._dband
        if band.seam_out is None:
            return None
        a, b = band.seam_out[0] - self._c0, min(band.seam_out[1], self._c0 + self._ch) - self._c0
        return self._acc[:, a:b, :]

    def seamIn(self, strip) -> None:
        """Installs the previous rank's accumulator strip as the starting state of this rank's ``seam_in`` rows."""
        band = self._dband
        if band.seam_in is None:
            return
        a, b = band.seam_in[0] - self._c0, band.seam_in[1] - self._c0
        if tuple(strip.shape) != (3, b - a, self.plan.canvas_w):
            raise ValueError(f"seam strip must have shape {(3, b - a, self.plan.canvas_w)}, got {tuple(strip.shape)}")
        self._acc[:, a:b, :].copy_(strip)

    def _exchange_seams(self) -> None:
        """The one exchange step of dedup mode: accumulator strips travel from every rank to the next one by
        point-to-point send / recv (NCCL over NVLink on the GPUs).  No strip depends on another rank's strip (bands are
        taller than the overlap), so all transfers are posted at once."""
        if self.world_size == 1:
            return
        torch = _torch()
        import torch.distributed as dist
        band = self._dband
        ops, recv = [], None
        out = self.seamOut()
        if out is not None:
            send = out.contiguous()
            ops.append(dist.P2POp(dist.isend, send, self.rank + 1))
        if band.seam_in is not None:
            recv = torch.empty((3, band.seam_in[1] - band.seam_in[0], self.plan.canvas_w), dtype=torch.float32,
                               device=self.device)
            ops.append(dist.P2POp(dist.irecv, recv, self.rank - 1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if recv is not None:
            self.seamIn(recv)

    def processBandFinish(self) -> None:
        """Dedup mode, phase 2: blend the kept predictions into the seam rows (now holding the previous rank's state),
        then rebuildTile's tail (:404-413) and the crop of rebuildMap (:541-545) for this rank's output rows."""
        plan, band = self.plan, self._dband
        if band.seam_in is not None:
            for pred, lohi, k0, n, gy_lo, gy_hi, add_half in self._retained:
                self._accumulate(pred, lohi, k0, n, gy_lo, gy_hi, add_half, band.seam_in[0], band.seam_in[1])
        self._retained = []
        rows = self._out_r1 - self._out_r0
        if rows > 0:
            acc = self._acc
            first = ((self._out_r0 + plan.off - self._c0) * plan.canvas_w + plan.off) * 4
            _lib.check(self._lib.msr_blend_finalize(acc[0].data_ptr() + first, acc[1].data_ptr() + first,
                                                    acc[2].data_ptr() + first, plan.canvas_w, rows, plan.width,
                                                    float(np.float32(self.no_value)), self.mean_out.data_ptr(),
                                                    self.std_out.data_ptr(), self.good_out.data_ptr(), plan.width,
                                                    _lib.stream_ptr()), "msr_blend_finalize")
            self.launches += 1

    def processBand(self, eps=None) -> None:
        """Dedup mode: the whole band of this rank (the counterpart of the processTile loop)."""
        if self.mode != "dedup":
            raise ValueError("processBand needs mode='dedup'")
        self.processBandMain(eps)
        self._exchange_seams()
        self.processBandFinish()

    def _save_tile_bufs(self, bufs, px, py) -> None:
        m, s, g = (x.cpu().numpy() for x in bufs)
        self.saveTile(m, s, g, str(px) + "_" + str(py))

    # ---------------------------------------------------------------------------------------------------------------
    # outputs
    # ---------------------------------------------------------------------------------------------------------------
    def saveGTiff(self, data: np.ndarray, data_type, name: str) -> None:
        """process_full_tiles.py:481-531 -- same naming, same dtype promotion (uint8 -> UInt16), NoData = no_value,
        geo-referencing of the input DEM; the container is written by geotiff.write (no GDAL)."""
        from . import geotiff
        if data_type == np.float32:
            pass
        elif data_type == np.uint8:
            data = data.astype(np.uint16)
        elif data_type == np.uint16:
            pass
        else:
            raise ValueError("Unsupported data-type.")
        if len(data.shape) < 2:
            raise ValueError("Data is of incorrect shape. The array must be 2-dimensional at least.")
        elif len(data.shape) > 3:
            raise ValueError("Data is of incorrect shape")
        geo = self.geo_transform
        if geo is not None and not isinstance(geo, dict):
            # the reference's types (:177-178): GDAL's 6-number affine GeoTransform and a WKT projection string
            geo = geotiff.geo_tags_from_gdal(geo, self.geo_projection)
        geotiff.write(os.path.join(self.save_path, self.map_name + "_" + name + ".tiff"), data, geo=geo,
                      nodata=self.no_value)

    def results(self):
        """(mean f32, std f32, good u8) of this rank's band as numpy arrays plus the band's first raster row.  The arrays
        are views of pinned host buffers owned by the engine (reused by the next call; copy them to keep them)."""
        torch = _torch()
        outs = []
        for k, t in enumerate((self.mean_out, self.std_out, self.good_out)):
            buf = self._host_out[k] if self._host_out is not None else None
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            outs.append(buf)
        self._host_out = outs
        torch.cuda.current_stream().synchronize()
        return outs[0].numpy(), outs[1].numpy(), outs[2].numpy(), self._out_r0

    def gatherResults(self):
        """Full (H, W) rasters on rank 0 (None elsewhere).  Bands are disjoint row ranges, so this is a plain
        gather of the output rows -- no arithmetic crosses ranks."""
        mean, std, good, r0 = self.results()
        res = gather_bands([mean, std, good], r0, self.plan.height, self.plan.width, self.rank, self.world_size)
        return None if res is None else tuple(res)

    def rebuildMap(self) -> None:
        """process_full_tiles.py:533-566 -- the assembled rasters already exist on the device (processTile pasted and
        cropped); fetch them and write the three GeoTIFFs."""
        res = self.gatherResults()
        if res is None:
            return
        mean, std, good = res
        self.saveGTiff(mean, mean.dtype, "mean")
        self.saveGTiff(std, std.dtype, "std")
        self.saveGTiff(good, good.dtype, "good")

    def processTiles(self) -> None:
        """Every tile of this rank's band, in tile-list order (process_full_tiles.py:579-581); in dedup mode the band
        of the global patch lattice."""
        if self.mode == "dedup":
            self.processBand()
            return
        for tile in self.my_tiles:
            self.processTile(*tile)

    def processMap(self) -> None:
        """process_full_tiles.py:568-587."""
        self.loadImages()
        self.preprocess()
        self.padInputs()
        tile_list = self.generateTileList()
        print("Cutting the image in", self.dem_shape[1] // self.tile_size + 1, "by",
              self.dem_shape[0] // self.tile_size + 1, "tiles.")
        if self.mode == "dedup":
            print("Processing lattice rows", self._dband.j0, "to", self._dband.j1, "(dedup mode)")
            self.processBand()
            tile_list = []
        for tile in tile_list:
            if tuple(tile) in self._tile_plan:
                print("Processing tile", tile[0], tile[1])
                self.processTile(*tile)
        self.dem_padded = None
        self.img_padded = None
        self.rebuildMap()
        return

    def rowsNeeded(self, height: int, width: int) -> Tuple[int, int]:
        """Raster rows [r0, r1) this rank's band reads (its tiles plus the I - S halo), before any data is loaded."""
        plan = Plan(height, width, self.image_size, self.stride, self.tile_size, self.batch_size)
        if self.mode == "dedup":
            return self._rows_dedup(plan, self.rank)[1]
        tiles, _, _ = band_of_rank(plan, self.world_size, self.rank)
        if not tiles:
            return 0, 0
        c0 = min(yy for _, yy in tiles)
        c1 = min(plan.canvas_h, max(yy for _, yy in tiles) + plan.tile_size + 2 * plan.off)
        return max(0, c0 - plan.off), min(height, c1 - plan.off)

    def run(self, dem, img, row_offset: int = 0, full_height: Optional[int] = None, copy: bool = True):
        """In-memory processMap: rasters in, (mean, std, good) of this rank's band out (numpy).  ``copy=False`` hands
        out views of the engine's pinned staging buffers (valid until the next results() / run() on this engine)."""
        self.setRasters(dem, img, row_offset=row_offset, full_height=full_height)
        self.padInputs()
        self.processTiles()
        out = self.results()[:3]
        return tuple(np.array(a) for a in out) if copy else out
