"""moonsuperresolution_b200 -- B200-native (sm_100a) implementation of MoonSuperResolution's tiled full-DEM inference
path behind the reference's own Python API (process_full_tiles.py).  The pixel work runs in ``libmoonsr.so`` (C ABI:
``include/moonsr.h``); there is no CPU fallback."""
from .engine import DEMSuperResolution, DSRConfig, parse_args  # noqa: F401
from .models import CNNSpade, GauGAN, GauGAN_no_KL, IdentityModel, Pix2Pix, load_CNN_model, load_GAN_model  # noqa: F401
from .planner import Plan, plan_batches, shard_tiles  # noqa: F401

__all__ = ["DEMSuperResolution", "DSRConfig", "parse_args", "GauGAN", "CNNSpade", "GauGAN_no_KL", "Pix2Pix", "IdentityModel", "load_GAN_model",
           "load_CNN_model", "Plan", "plan_batches", "shard_tiles"]
