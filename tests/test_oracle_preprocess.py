"""Pins oracle/preprocess.py: the library-level restatement against the output of the reference's own ``preprocess``
method (tests/golden/make_golden_preprocess.py), the numpy restatements of OpenCV's resize arithmetic against cv2 with
IPP switched off, and the host-side pieces of the product (lookup tables, hole fill) against the oracle."""
import numpy as np
import pytest

import golden_inputs
from oracle import preprocess as OP

cv2 = pytest.importorskip("cv2")
pytest.importorskip("scipy")


@pytest.fixture(scope="module")
def golden_pp():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_golden.npz"))


@pytest.mark.parametrize("name", list(golden_inputs.PREPROCESS_CASES))
def test_restatement_equals_reference_output(golden_pp, name):
    dem, _, nv = golden_inputs.make_preprocess_case(name)
    with np.errstate(all="ignore"):
        got = OP.reference_preprocess(dem, nv)
    np.testing.assert_array_equal(got, golden_pp[f"{name}/dem"])
    assert int((got <= nv).sum()) == int(golden_pp[f"{name}/nv_count"])


@pytest.fixture()
def no_ipp():
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    yield
    cv2.ipp.setUseIPP(was)


@pytest.mark.parametrize("h,w", [(64, 64), (65, 70), (66, 67), (67, 62), (101, 99), (30, 31), (3, 9)])
def test_area4_restatement_is_opencv_arithmetic(no_ipp, h, w):
    """Full windows, windows cut by the edge (H, W not multiples of 4, cvRound half-to-even extents), NaN propagation."""
    rng = np.random.default_rng(h * 1000 + w)
    a = (rng.standard_normal((h, w)) * 1000).astype(np.float32)
    if h > 12:
        a[5:9, 7:13] = np.nan
    want = cv2.resize(a, (0, 0), fx=0.25, fy=0.25, interpolation=cv2.INTER_AREA)
    got = OP.area4(a)
    assert got.shape == want.shape == OP.area4_shape(h, w)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("h,w,H,W", [(8, 8, 128, 128), (7, 9, 100, 140), (24, 24, 384, 384), (5, 64, 80, 1024),
                                     (20, 20, 320, 320), (3, 2, 50, 31)])
def test_cubic_restatement_is_opencv_arithmetic(no_ipp, h, w, H, W):
    rng = np.random.default_rng(h * 100 + w)
    a = (np.cumsum(rng.standard_normal((h, w)), 1) * 100 + 1500).astype(np.float32)
    if h > 4:
        a[2, 3] = np.nan
    want = cv2.resize(a, (W, H), interpolation=cv2.INTER_CUBIC)
    got = OP.cubic_resize(a, H, W)
    # OpenCV's vertical pass is vectorised 4 columns at a time (SSE baseline of the pip wheel) and accumulates taps
    # 3 -> 0; the scalar tail that handles the last W mod 4 columns accumulates 0 -> 3.  The restatement (and the CUDA
    # kernel) use the vector order everywhere: bit-exact on the vectorised columns, rounding-level on the tail.
    body = W - W % 4
    np.testing.assert_array_equal(got[:, :body], want[:, :body])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    if body < W:
        ulp = np.float32(np.nanmax(np.abs(want))) * np.float32(2.0 ** -23)
        assert np.nanmax(np.abs(got[:, body:] - want[:, body:])) <= 4 * ulp


def test_ipp_cubic_differs_only_by_rounding():
    """The pip wheel's default path (Intel IPP) and OpenCV's own code agree to a few ulp of the raster's magnitude and
    produce the same NaN footprint -- the bar the CUDA kernel is held to against stock cv2."""
    rng = np.random.default_rng(0)
    a = (np.cumsum(rng.standard_normal((20, 20)), 1) * 100 + 1500).astype(np.float32)
    a[4, 5] = np.nan
    stock = cv2.resize(a, (320, 320), interpolation=cv2.INTER_CUBIC)
    own = OP.cubic_resize(a, 320, 320)
    np.testing.assert_array_equal(np.isnan(stock), np.isnan(own))
    ulp = np.float32(np.nanmax(np.abs(stock))) * np.float32(2.0 ** -23)
    assert np.nanmax(np.abs(stock - own)) <= 8 * ulp


def test_product_tables_and_hole_fill_equal_the_oracle():
    from moonsuperresolution_b200 import preprocess as P
    for dst, src in [(320, 20), (100, 7), (70000, 4375), (15000, 938), (31, 2)]:
        a, b = P.cubic_tables(dst, src), OP.cubic_tables(dst, src)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    assert P.area4_shape(1002, 1006) == OP.area4_shape(1002, 1006) == (250, 252)
    rng = np.random.default_rng(3)
    h, w = 300, 280                                     # 2 x 2 fill blocks of the 1/4 raster
    d = (np.cumsum(np.cumsum(rng.standard_normal((h, w)), 0), 1) + 1500).astype(np.float32)
    nv = -32768.0
    for (y, x, a, b) in [(50, 60, 2, 3), (200, 190, 4, 4), (100, 150, 10, 10), (222, 222, 2, 2), (10, 10, 2, 2)]:
        d[y:y + a, x:x + b] = nv
    want = OP.fill_nan(d, nv, 256, 32, 24)
    np.testing.assert_array_equal(P.fill_small_holes(d, nv), want)
    assert (want[50:52, 60:63] > nv).all()              # small hole in a block interior: filled
    assert (want[100:110, 150:160] <= nv).all()         # 100 pixels: too large
    assert (want[10:12, 10:12] <= nv).all()             # inside the outer border frame: never filled
    # restricting the work to the blocks that hold invalid pixels changes nothing
    blocks = [(x, y) for (x, y) in P.fill_blocks(h, w) if (d[y:y + 256, x:x + 256] <= nv).any()]
    np.testing.assert_array_equal(P.fill_small_holes(d, nv, origins=blocks), want)


def test_engine_exposes_the_reference_hole_fill_methods():
    """DEMSuperResolution.fillNan / interpolateMissingValues (process_full_tiles.py:184-224) keep the reference's
    signatures; host code, so callable without a device (unbound here: the constructor needs CUDA)."""
    from moonsuperresolution_b200.engine import DEMSuperResolution as E
    rng = np.random.default_rng(4)
    d = (np.cumsum(np.cumsum(rng.standard_normal((150, 140)), 0), 1) + 1500).astype(np.float32)
    nv = -32768.0
    d[70:72, 60:63] = nv
    d[5:7, 5:7] = nv
    np.testing.assert_array_equal(E.fillNan(None, d, nv, tile_size=128, border=16, max_fill_area=24),
                                  OP.fill_nan(d, nv, 128, 16, 24))
    blk = d[:128, :128]
    np.testing.assert_array_equal(E.interpolateMissingValues(None, blk.copy(), nv, 24),
                                  OP.interpolate_missing(blk.copy(), nv, 24))
