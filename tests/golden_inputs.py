"""Seeded synthetic rasters for the golden tiling cases (shared by make_golden.py and the tests)."""
import numpy as np

NV = -32768.0

CASES = {
    # SURVEY.md App. E known-answer: identity model must reproduce the DEM where good == 1
    "identity_700x900": dict(H=700, W=900, I=64, S=8, B=16, T=256, NV=NV, model="identity", seed=0, holes=False),
    # overlapping generations disagree (std > 0), NV holes + NV border, ragged last batch, full arrays stored
    "wobble_200x260": dict(H=200, W=260, I=32, S=8, B=5, T=128, NV=NV, model="wobble", seed=1, holes=True,
                           store_full=True),
    # S does not divide T (S | T + I holds), odd purge
    "wobble_I48_S16": dict(H=150, W=330, I=48, S=16, B=7, T=128, NV=NV, model="wobble", seed=2, holes=True,
                           store_full=True),
    # raster larger than one canvas quantum, T = 512
    "wobble_1100x1300": dict(H=1100, W=1300, I=64, S=16, B=16, T=512, NV=NV, model="wobble", seed=3, holes=True),
}


def make_rasters(case):
    rng = np.random.default_rng(case["seed"])
    h, w = case["H"], case["W"]
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w)).astype(np.float32)
    if case["holes"]:
        nv = np.float32(case["NV"])
        dem[:3, :] = nv                       # NV stripe at the top of the DEM only
        img[:, w - 5:] = nv                   # NV stripe at the right of the ortho only
        cy, cx = h // 2, w // 3
        dem[cy:cy + 4, cx:cx + 6] = nv        # small hole in the DEM
        img[h // 4, w // 2] = nv - 1.0        # single pixel below no_value in the ortho
    return dem, img
