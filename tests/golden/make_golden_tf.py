#!/usr/bin/env python
"""Golden vectors of the reference generators, produced by running the reference's OWN model code (reference_graph.py).

    python tests/golden/make_golden_tf.py           # real TensorFlow (+ tensorflow-addons) -> tests/golden/generator_tf.npz
    python tests/golden/make_golden_tf.py --shim    # numpy op shim                        -> tests/golden/generator_refshim.npz

The TensorFlow variant needs a box with tensorflow>=2.5 and tensorflow-addons (pip-env.py:34,36 of the reference) and the
reference checkout (MSR_REFERENCE_PATH, default /root/reference); neither exists in the build image, so only the shim
file is committed.  The day generator_tf.npz is dropped next to this script, tests/test_oracle_generator_pinned.py pins
oracle/generator.py (and through it every CUDA parity test) to TensorFlow itself.

Inputs are not stored: every case is regenerated from its seeds by ``case_inputs`` (shared with the test).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

# batch-of-one cases executed by OpenCV's TensorFlow importer on the graph traced from the reference's code (--opencv)
# (I = 64 only for the SPADE nets: OpenCV reads a 2-D -> 4-D Reshape as NCHW, so Reshape((sw, sw, 1024)) with sw > 1 --
# networks.py:42 at I >= 128 -- is outside what its importer can represent; sw = 1 is layout-free.)
OPENCV_CASES = [("spade64b1", "spade", 64, 1, 25, 9), ("cnn64b1", "cnn", 64, 1, 26, 10), ("spade64b1_s2", "spade", 64, 1, 28, 12),
                ("pix2pix", "pix2pix", 256, 1, 24, 8)]

# (name, arch, image_size, batch, weight seed, input seed)
CASES = [("cnn64", "cnn", 64, 3, 21, 5), ("spade64", "spade", 64, 2, 22, 6), ("spade128", "spade", 128, 2, 23, 7),
         ("pix2pix", "pix2pix", 256, 1, 24, 8)]


def case_inputs(arch, image_size, batch, wseed, xseed):
    from moonsuperresolution_b200 import weights as W
    w = W.random_init(arch, image_size, seed=wseed, perturb_affine=True)
    rng = np.random.default_rng(xseed)
    x = rng.uniform(-0.5, 0.5, (batch, image_size, image_size, 2)).astype(np.float32)
    if batch > 1:
        x[-1] = 0.0          # a zero padding slot takes part in the batch statistics (process_full_tiles.py:468-474)
    eps = rng.standard_normal((batch, 256)).astype(np.float32)
    return w, x, eps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shim", action="store_true")
    ap.add_argument("--opencv", action="store_true",
                    help="trace the reference's code into TensorFlow GraphDefs and execute them with OpenCV's TensorFlow "
                         "importer -> tests/golden/generator_opencv_tf.npz")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import reference_graph as RG
    if args.opencv:
        import tempfile
        ref = RG.Reference(backend="shim", dtype=np.float32)
        out = {}
        with tempfile.TemporaryDirectory() as tmp:
            for name, arch, i, b, ws, xs in OPENCV_CASES:
                w, x, eps = case_inputs(arch, i, b, ws, xs)
                got, shim = ref.run_pix2pix_opencv(w, x, tmp) if arch == "pix2pix" else ref.run_spade_opencv(arch, i, w, x, eps, tmp)
                out[f"{name}.out"] = np.asarray(got, np.float32)
                print(name, got.shape, "max |opencv - shim| =", float(np.abs(got - shim).max()), "max |out| =", float(np.abs(got).max()))
        ref.close()
        import cv2
        out["backend"] = np.array("opencv-dnn-tensorflow-importer " + cv2.__version__)
        path = args.out or os.path.join(HERE, "generator_opencv_tf.npz")
        np.savez_compressed(path, **out)
        print("wrote", path)
        return
    backend = "shim" if args.shim else "tf"
    ref = RG.Reference(backend=backend, dtype=np.float64)
    out = {}
    for name, arch, i, b, ws, xs in CASES:
        w, x, eps = case_inputs(arch, i, b, ws, xs)
        res = ref.run_pix2pix(w, x) if arch == "pix2pix" else ref.run_spade(arch, i, w, x, eps)
        for k, v in res.items():
            out[f"{name}.{k}"] = np.asarray(v, np.float32)
        print(name, {k: (v.shape, float(np.abs(v).max())) for k, v in res.items()})
    ref.close()
    out["backend"] = np.array(backend)
    path = args.out or os.path.join(HERE, "generator_refshim.npz" if args.shim else "generator_tf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
