"""Converts the Keras SavedModel directories the reference writes (GauGAN.save, spade/models/model.py:569-605:
<path>/generator, <path>/encoder) into one Keras-layout ``weights.npz`` -- without TensorFlow -- and lists what it found.

    python tools/convert_saved_model.py <model_path> <image_size> [--arch spade|cnn] [--out weights.npz] [--list]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonsuperresolution_b200 import savedmodel as SM      # noqa: E402
from moonsuperresolution_b200 import weights as W          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model_path")
    ap.add_argument("image_size", type=int)
    ap.add_argument("--arch", default="spade", choices=["spade", "cnn"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--list", action="store_true", help="print every variable of both bundles and exit")
    args = ap.parse_args()
    gen, enc = os.path.join(args.model_path, "generator"), os.path.join(args.model_path, "encoder")
    if args.list:
        for d in (gen, enc):
            print(d)
            for k, v in sorted(SM.read_saved_model_variables(d).items()):
                print("  %-70s %s" % (k, tuple(v.shape)))
        return
    weights = SM.load_gaugan_weights(gen, enc, args.image_size, args.arch)
    out = args.out or os.path.join(args.model_path, "weights.npz")
    W.save_npz(out, weights)
    n = sum(int(v.size) for v in weights.values())
    print(f"{len(weights)} tensors, {n / 1e6:.2f} M parameters -> {out}")


if __name__ == "__main__":
    main()
