"""Generates tests/golden/preprocess_golden.npz by running the UNMODIFIED reference ``preprocess`` method
(process_full_tiles.py:226-244, with fillNan :214-224 and interpolateMissingValues :184-212).

Run in the build container only (needs /root/reference, cv2 and scipy):   python tests/golden/make_golden_preprocess.py

The raster is square: the reference hands (H, W) to cv2.resize as (width, height), so any other shape comes back
transposed and crashes in padInputs.  Inputs are regenerated from the seed by ``golden_inputs.make_preprocess_case``;
the stored output is what ``self.dem`` holds after ``preprocess()``."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import golden_inputs  # noqa: E402
import make_golden  # noqa: E402


def main():
    ref = make_golden.import_reference()
    store = {}
    for name in golden_inputs.PREPROCESS_CASES:
        dem, img, nv = golden_inputs.make_preprocess_case(name)
        cfg = ref.DSRConfig(no_value=nv, save_path="/nonexistent")
        eng = ref.DEMSuperResolution(cfg)
        eng.dem, eng.img = dem.copy(), img.copy()
        eng.dem_shape, eng.img_shape = dem.shape, img.shape
        with contextlib.redirect_stdout(io.StringIO()):
            eng.preprocess()
        assert eng.dem.shape == dem.shape and eng.dem.dtype == np.float32
        store[f"{name}/dem"] = eng.dem
        store[f"{name}/nv_count"] = np.array(int((eng.dem <= nv).sum()))
        print(name, eng.dem.shape, "no_value pixels:", int((eng.dem <= nv).sum()))
    np.savez_compressed(os.path.join(HERE, "preprocess_golden.npz"), **store)


if __name__ == "__main__":
    main()
