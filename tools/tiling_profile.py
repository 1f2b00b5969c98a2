"""Runs the tiling side of the path (preprocess, pad, validity, gather/normalise, blend) at the bench geometry with the
device identity model, tile by tile and in dedup mode -- a short command for ncu captures of the HBM-bound kernels, and
a CUDA-event rate table of the same kernels without ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import DEMSuperResolution, DSRConfig, IdentityModel, _lib

h = w = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
i, s, b, t = 512, 128, 16, 1024
gen = torch.Generator(device="cuda").manual_seed(1)
dem = (torch.cumsum(torch.randn((h, w), generator=gen, device="cuda"), 1) * 3.0 + 1500.0).contiguous()
img = (torch.rand((h, w), generator=gen, device="cuda") * 254.0 + 1.0).contiguous()
for mode in ("faithful", "dedup"):
    eng = DEMSuperResolution(DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t, mode=mode),
                             model=IdentityModel(i, b))

    def step():
        eng.setRasters(dem, img)
        eng.preprocess()
        eng.padInputs()
        eng.processTiles()
    for rep in range(2):
        step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    step()
    print("mode", mode)
    for k, v in _lib.profile_read().items():
        if v["launches"]:
            print("  %-10s %8.3f ms  %8.1f GB/s  launches %d" % (k, v["ms"], v["work"] / v["ms"] / 1e6, v["launches"]))
    _lib.profile_enable(False)
