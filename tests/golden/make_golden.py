"""Generates tests/golden/tiling_golden.npz by running the UNMODIFIED reference tiling/blend code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

The reference's ``process_full_tiles.py`` is imported with stub modules for ``osgeo`` (GDAL is not installed)
and ``spade.models.model`` (TensorFlow is not installed and the file has unresolved merge markers); the methods
on the path -- padInputs, generateTileList, processTile{getPatch, normalize, processBatch, rebuildTile},
makeGaussianKernel -- then execute unmodified.  saveTile is redirected to memory (cv2.imwrite round-trips float32 /
uint8 TIFFs bit-exactly) and rebuildMap's paste + crop (process_full_tiles.py:541-545) is applied
to the captured tiles because saveGTiff needs GDAL.

Inputs are regenerated from seeds by ``tests/golden_inputs.py``; only outputs / digests are stored."""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_inputs  # noqa: E402
import toy_models  # noqa: E402


def import_reference():
    osgeo = types.ModuleType("osgeo")
    gdal = types.ModuleType("osgeo.gdal")
    osgeo.gdal = gdal
    sys.modules["osgeo"], sys.modules["osgeo.gdal"] = osgeo, gdal
    for name in ("spade", "spade.models", "spade.models.model"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["spade.models.model"].GauGAN = object
    sys.modules["spade.models.model"].CNNSpade = object
    sys.path.insert(0, "/root/reference")
    import process_full_tiles
    return process_full_tiles


def run_reference(ref, case):
    dem, img = golden_inputs.make_rasters(case)
    cfg = ref.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                        no_value=case["NV"], save_path="/nonexistent")
    model = getattr(toy_models, case["model"])
    eng = ref.DEMSuperResolution(cfg, model=model)
    eng.dem, eng.img = dem.copy(), img.copy()
    eng.dem_shape, eng.img_shape = dem.shape, img.shape
    eng.padInputs()
    tiles = {}

    def capture(mean, std, good, name):
        xx, yy = (int(v) for v in name.split("_"))
        tiles[(xx, yy)] = (mean.copy(), std.copy(), good.copy())

    eng.saveTile = capture
    with np.errstate(all="ignore"):
        for (xx, yy) in eng.generateTileList():
            eng.processTile(xx, yy)
    outs = []
    for k, dt in enumerate((np.float32, np.float32, np.uint8)):
        canvas = np.zeros(eng.dem_padded_shape, dtype=dt)
        for (xx, yy), trip in tiles.items():
            canvas[yy:yy + eng.tile_size, xx:xx + eng.tile_size] = trip[k]
        outs.append(np.ascontiguousarray(
            canvas[:-eng.pad_y - eng.image_size + eng.stride, :-eng.pad_x - eng.image_size + eng.stride]))
    facts = dict(canvas=np.array(eng.dem_padded_shape), pad=np.array([eng.pad_x, eng.pad_y]),
                 ntiles=np.array(len(tiles)))
    return outs, facts


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = import_reference()
    store = {}
    for name, case in golden_inputs.CASES.items():
        (mean, std, good), facts = run_reference(ref, case)
        store[f"{name}/sha_mean"] = np.array(digest(mean))
        store[f"{name}/sha_std"] = np.array(digest(std))
        store[f"{name}/sha_good"] = np.array(digest(good))
        store[f"{name}/good_count"] = np.array(int(good.sum()))
        for k, v in facts.items():
            store[f"{name}/{k}"] = v
        if case.get("store_full"):
            store[f"{name}/mean"], store[f"{name}/std"], store[f"{name}/good"] = mean, std, good
        print(name, mean.shape, "good", int(good.sum()), digest(mean)[:16])
    # blend weight tables (makeGaussianKernel + 1e-7, purge crop), float64 bytes
    for i in (32, 48, 64, 256, 512):
        cfg = ref.DSRConfig(image_size=i)
        w = ref.DEMSuperResolution(cfg).makeGaussianKernel() + 1e-7
        p = i // 16
        w = np.ascontiguousarray(w[p:-p, p:-p])
        store[f"weights/{i}/sha"] = np.array(digest(w))
        store[f"weights/{i}/minmax"] = np.array([w.min(), w.max()])
        print("weights", i, digest(w)[:16], w.min(), w.max())
    np.savez_compressed(os.path.join(HERE, "tiling_golden.npz"), **store)


if __name__ == "__main__":
    main()
