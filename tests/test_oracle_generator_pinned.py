"""Pins oracle/generator.py (the checker of every CUDA generator test) to the reference's OWN model code.

Three anchors, strongest first:
  1. tests/golden/generator_tf.npz -- outputs of the unmodified reference modules under real TensorFlow
     (tests/golden/make_golden_tf.py).  Not producible in the build image (no TensorFlow); consumed when present.
  2. tests/golden/generator_refshim.npz -- outputs of the unmodified reference modules (networks / blocks / spade /
     sampling / pix2pix + the GauGAN / CNNSpade ``call`` bodies cut out of model.py) executed on a numpy op shim
     (tests/golden/tf_numpy_shim.py): pins the GRAPH of the oracle to the reference source; and, where the reference
     checkout is present, the same run repeated live.
  3. the shim's ops against scalar-loop restatements of TensorFlow's documented semantics (SAME padding k4s1 / k3s2 /
     k4s2, Conv2DTranspose k4s2, half-pixel nearest resize, NHWC flatten) on tiny shapes -- no torch, no numpy
     vectorisation -- and the oracle's torch ops against the same loops (SURVEY.md App. B silent-mismatch list).
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)

import make_golden_tf as MG          # noqa: E402
import tf_numpy_shim as SH           # noqa: E402
from oracle import generator as OG   # noqa: E402


def oracle_outputs(arch, i, b, ws, xs, dtype):
    w, x, eps = MG.case_inputs(arch, i, b, ws, xs)
    if arch == "pix2pix":
        return {"out": OG.pix2pix_call(x, w, dtype)}
    out, latent = OG.gaugan_call(x, w, eps, arch, dtype, return_latent=True)
    with torch.no_grad():
        src = torch.from_numpy(x).to(dtype).permute(0, 3, 1, 2)
        mean, var = OG.encoder(src, w)
    return {"out": out, "latent": latent, "enc.mean": mean.numpy(), "enc.variance": var.numpy()}


def compare(got, want_npz, name, tol):
    for k, v in got.items():
        want = want_npz[f"{name}.{k}"]
        assert want.shape == v.shape, (name, k)
        err = np.abs(np.asarray(v, np.float64) - want).max()
        assert err <= tol * max(1.0, np.abs(want).max()), (name, k, err)


@pytest.mark.parametrize("case", MG.CASES, ids=[c[0] for c in MG.CASES])
def test_oracle_matches_reference_modules_on_the_op_shim(case):
    """fp64 oracle vs the committed outputs of the reference graph (stored as float32: agreement to float32 rounding)."""
    name, arch, i, b, ws, xs = case
    z = np.load(os.path.join(GOLDEN, "generator_refshim.npz"))
    compare(oracle_outputs(arch, i, b, ws, xs, torch.float64), z, name, 2e-6)
    # the float32 oracle (what the GPU tests compare with) sits inside the fp32 tolerance of north_star
    compare(oracle_outputs(arch, i, b, ws, xs, torch.float32), z, name, 1e-4)


def test_reference_modules_run_live_on_the_shim():
    """Re-runs the unmodified reference modules here (needs the reference checkout) and checks the committed file."""
    import reference_graph as RG
    if not RG.available():
        pytest.skip("reference checkout not present on this box")
    name, arch, i, b, ws, xs = MG.CASES[1]
    w, x, eps = MG.case_inputs(arch, i, b, ws, xs)
    ref = RG.Reference(backend="shim", dtype=np.float64)
    try:
        res = ref.run_spade(arch, i, w, x, eps)
        assert len(ref.build_spade(i, b, w).blocks) == 6
    finally:
        ref.close()
    z = np.load(os.path.join(GOLDEN, "generator_refshim.npz"))
    for k, v in res.items():
        np.testing.assert_allclose(np.asarray(v, np.float32), z[f"{name}.{k}"], rtol=0, atol=1e-6)
    assert "tensorflow" not in sys.modules or not getattr(sys.modules["tensorflow"], "__msr_shim__", False)


def test_oracle_matches_tensorflow_golden_when_present():
    """generator_tf.npz is written by make_golden_tf.py on a TensorFlow-equipped box (see its docstring)."""
    path = os.path.join(GOLDEN, "generator_tf.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/generator_tf.npz absent: generator parity is pinned to the reference graph on the "
                    "numpy op shim only (TensorFlow is not installable in the build image)")
    z = np.load(path)
    for name, arch, i, b, ws, xs in MG.CASES:
        compare(oracle_outputs(arch, i, b, ws, xs, torch.float32), z, name, 1e-4)


# ----------------------------------------------------------------------------------------------------------------------
# op-level anchors: scalar loops written from TensorFlow's documented semantics
# ----------------------------------------------------------------------------------------------------------------------
def loop_conv_same(x, k, s):
    """tf.nn.conv2d(padding='SAME'): out = ceil(in / s), pad_total = max((out - 1) * s + k - in, 0), pad_before =
    pad_total // 2; input index = out * s + tap - pad_before, out-of-range reads are zero."""
    n, h, w, ci = x.shape
    kh, kw, _, co = k.shape
    oh, ow = -(-h // s), -(-w // s)
    pt = max((oh - 1) * s + kh - h, 0) // 2
    pl = max((ow - 1) * s + kw - w, 0) // 2
    y = np.zeros((n, oh, ow, co))
    for b in range(n):
        for oy in range(oh):
            for ox in range(ow):
                for ky in range(kh):
                    for kx in range(kw):
                        iy, ix = oy * s + ky - pt, ox * s + kx - pl
                        if 0 <= iy < h and 0 <= ix < w:
                            for c in range(ci):
                                for o in range(co):
                                    y[b, oy, ox, o] += x[b, iy, ix, c] * k[ky, kx, c, o]
    return y


def loop_conv_transpose_same_s2(x, k):
    """Conv2DTranspose(k=4, strides=2, 'same') = gradient of the SAME stride-2 convolution 2n -> n: every input pixel
    (iy, ix) adds x * w[ky, kx] to output (2*iy + ky - 1, 2*ix + kx - 1) (pad_before of that convolution is 1)."""
    n, h, w, ci = x.shape
    kh, kw, co, _ = k.shape
    y = np.zeros((n, 2 * h, 2 * w, co))
    for b in range(n):
        for iy in range(h):
            for ix in range(w):
                for ky in range(kh):
                    for kx in range(kw):
                        oy, ox = 2 * iy + ky - 1, 2 * ix + kx - 1
                        if 0 <= oy < 2 * h and 0 <= ox < 2 * w:
                            for c in range(ci):
                                for o in range(co):
                                    y[b, oy, ox, o] += x[b, iy, ix, c] * k[ky, kx, o, c]
    return y


@pytest.mark.parametrize("k,s,size", [(3, 1, 5), (4, 1, 6), (3, 2, 6), (4, 2, 6), (3, 2, 5)])
def test_same_convolution_against_scalar_loops(k, s, size):
    rng = np.random.default_rng(k * 10 + s)
    x = rng.standard_normal((2, size, size, 3))
    w = rng.standard_normal((k, k, 3, 2))
    want = loop_conv_same(x, w, s)
    SH.set_dtype(np.float64)
    np.testing.assert_allclose(SH.conv2d(x, w, s, "same"), want, atol=1e-12)
    got = OG.conv2d_same(torch.from_numpy(x).permute(0, 3, 1, 2), w, None, stride=s).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_transposed_convolution_against_scalar_loops():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 3, 3))
    w = rng.standard_normal((4, 4, 2, 3))                  # Keras layout [kh, kw, cout, cin]
    want = loop_conv_transpose_same_s2(x, w)
    SH.set_dtype(np.float64)
    np.testing.assert_allclose(SH.conv2d_transpose(x, w, 2, "same"), want, atol=1e-12)
    got = OG.conv2d_transpose_same_s2(torch.from_numpy(x).permute(0, 3, 1, 2), w, None).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, want, atol=1e-12)
    # and it is the adjoint of the SAME stride-2 convolution: <conv(u), v> == <u, convT(v)>
    u = rng.standard_normal((1, 6, 6, 2))
    kf = np.transpose(w, (0, 1, 2, 3))                     # forward kernel [kh, kw, cin=2, cout=3] has the same layout
    v = rng.standard_normal((1, 3, 3, 3))
    lhs = (loop_conv_same(u, kf, 2) * v).sum()
    rhs = (u * loop_conv_transpose_same_s2(v, w)).sum()
    assert abs(lhs - rhs) < 1e-10


def test_half_pixel_nearest_resize_upsampling_and_flatten():
    m = np.arange(2 * 8 * 8 * 2, dtype=np.float64).reshape(2, 8, 8, 2)
    for out in (1, 2, 4, 8):
        want = np.zeros((2, out, out, 2))
        for oy in range(out):
            for ox in range(out):
                sy = min(int(np.floor((oy + 0.5) * 8 / out)), 7)
                sx = min(int(np.floor((ox + 0.5) * 8 / out)), 7)
                want[:, oy, ox] = m[:, sy, sx]
        np.testing.assert_array_equal(SH.resize_nearest(m, (out, out)), want)
        got = OG.resize_nearest_tf(torch.from_numpy(m).permute(0, 3, 1, 2), out).permute(0, 2, 3, 1).numpy()
        np.testing.assert_array_equal(got, want)
    up = SH.UpSampling2D((2, 2))(m)
    for y in range(16):
        for x in range(16):
            np.testing.assert_array_equal(up[:, y, x], m[:, y // 2, x // 2])
    flat = SH.Flatten()(m)
    assert flat[1, (3 * 8 + 5) * 2 + 1] == m[1, 3, 5, 1]          # row-major over (h, w, c)
    assert SH.Reshape((8, 8, 2))(flat)[1, 3, 5, 1] == m[1, 3, 5, 1]
