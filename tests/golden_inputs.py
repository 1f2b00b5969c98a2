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


# ---- preprocess (process_full_tiles.py:226-244) ------------------------------------------------------------------------
PREPROCESS_CASES = {
    # small hole (filled on the 1/4 raster by the Clough-Tocher interpolant), large hole and NV stripe (kept -> NaN spreads
    # through the cubic upsampling), all inside one 256-block of the 1/4 raster
    "holes_320": dict(N=320, seed=7, holes=True),
    # clean raster: pure 1/16 box filter + cubic upsampling
    "clean_192": dict(N=192, seed=8, holes=False),
}


def make_preprocess_case(name):
    c = PREPROCESS_CASES[name]
    n = c["N"]
    rng = np.random.default_rng(c["seed"])
    dem = (np.cumsum(np.cumsum(rng.standard_normal((n, n)), 0), 1) * 0.5 + 1500.0).astype(np.float32)
    img = rng.uniform(1, 255, (n, n)).astype(np.float32)
    if c["holes"]:
        nv = np.float32(NV)
        dem[150:158, 160:172] = nv      # 2 x 3 pixels of the 1/4 raster, inside the interior [32:48) of its fill block
        dem[200:260, 40:120] = nv       # 15 x 20 pixels of the 1/4 raster: too large, stays
        dem[0:3, :] = nv
    return dem, img, NV
