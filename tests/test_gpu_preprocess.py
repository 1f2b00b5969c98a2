"""preprocess (process_full_tiles.py:226-244) on the device: msr_resize_area4 / msr_resize_cubic against the oracle's
restatement of OpenCV's arithmetic (bit-exact), against stock cv2 (Intel IPP rounding: <= 8 ulp of the raster's
magnitude, same no_value footprint) and, end to end, against the output of the reference's own method (golden)."""
import os

import numpy as np
import pytest

import golden_inputs
from oracle import preprocess as OP

pytestmark = pytest.mark.gpu
NV = -32768.0


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def msr(torch):
    import moonsuperresolution_b200 as m
    return m


def area4_device(torch, a):
    from moonsuperresolution_b200 import _lib
    h, w = a.shape
    dh, dw = OP.area4_shape(h, w)
    src = torch.from_numpy(a).cuda()
    dst = torch.empty((dh, dw), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().msr_resize_area4(src.data_ptr(), h, w, dst.data_ptr(), dh, dw, NV, _lib.stream_ptr()),
               "msr_resize_area4")
    return dst.cpu().numpy()


def cubic_device(torch, a, H, W):
    from moonsuperresolution_b200 import _lib
    from moonsuperresolution_b200 import preprocess as P
    h, w = a.shape
    xo, xc = P.cubic_tables(W, w)
    yo, yc = P.cubic_tables(H, h)
    t = [torch.from_numpy(x).cuda() for x in (a, xo, xc, yo, yc)]
    dst = torch.empty((H, W), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().msr_resize_cubic(t[0].data_ptr(), h, w, dst.data_ptr(), H, W, t[1].data_ptr(),
                                           t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), NV, _lib.stream_ptr()),
               "msr_resize_cubic")
    return dst.cpu().numpy()


@pytest.mark.parametrize("h,w", [(64, 64), (65, 70), (66, 67), (67, 62), (101, 99), (30, 31), (3, 9), (1000, 1500),
                                 (1023, 2050)])
def test_area4_kernel_bit_exact(torch, h, w):
    rng = np.random.default_rng(h * 1000 + w)
    a = (rng.standard_normal((h, w)) * 1000).astype(np.float32)
    if h > 12:
        a[5:9, 7:13] = NV
        a[h // 2, w // 3] = NV - 5.0
    nan_in = a.copy()
    nan_in[nan_in <= NV] = np.nan                                     # :230
    want = OP.area4(nan_in)
    want[np.isnan(want)] = NV                                         # :233
    np.testing.assert_array_equal(area4_device(torch, a), want)


@pytest.mark.parametrize("h,w,H,W", [(8, 8, 128, 128), (7, 9, 100, 140), (24, 24, 384, 384), (5, 64, 80, 1024),
                                     (3, 2, 50, 31), (63, 94, 1000, 1500), (64, 128, 1023, 2050)])
def test_cubic_kernel_bit_exact_and_close_to_stock_cv2(torch, h, w, H, W):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(h * 100 + w)
    a = (np.cumsum(rng.standard_normal((h, w)), 1) * 100 + 1500).astype(np.float32)
    if h > 4:
        a[2, 3] = NV
    nan_in = a.copy()
    nan_in[nan_in <= NV] = np.nan
    got = cubic_device(torch, a, H, W)
    want = OP.cubic_resize(nan_in, H, W)
    want[np.isnan(want)] = NV
    np.testing.assert_array_equal(got, want)
    stock = cv2.resize(nan_in, (W, H), interpolation=cv2.INTER_CUBIC)  # Intel IPP in the pip wheels
    np.testing.assert_array_equal(np.isnan(stock), got <= NV)
    ok = ~np.isnan(stock)
    ulp = np.float32(np.abs(stock[ok]).max()) * np.float32(2.0 ** -23)
    # IPP evaluates the sample position in float32 (OpenCV's code in double): at non-integer scale factors the position
    # is off by up to ~2^-21 * source index, which the local slope turns into a value difference
    step = max(np.nanmax(np.abs(np.diff(nan_in, axis=0))), np.nanmax(np.abs(np.diff(nan_in, axis=1))))
    tol = 8 * ulp + 2.0 ** -21 * max(h, w) * step
    assert np.abs(stock[ok] - got[ok]).max() <= tol


@pytest.mark.parametrize("name", list(golden_inputs.PREPROCESS_CASES))
def test_preprocess_matches_reference_golden(msr, torch, name):
    """End to end against what the UNMODIFIED reference method produced: same no_value footprint (hole fill included --
    the interpolant is the reference's own library call), values within 8 ulp of the raster's magnitude (IPP)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_golden.npz"))
    dem, img, nv = golden_inputs.make_preprocess_case(name)
    eng = msr.DEMSuperResolution(msr.DSRConfig(image_size=64, stride=16, batch_size=4, tile_size=256, no_value=nv))
    eng.setRasters(dem, img)
    eng.preprocess()
    got = eng.dem.cpu().numpy()
    want = g[f"{name}/dem"]
    assert got.shape == want.shape
    np.testing.assert_array_equal(got <= nv, want <= nv)
    ok = want > nv
    ulp = np.float32(np.abs(want[ok]).max()) * np.float32(2.0 ** -23)
    assert np.abs(got[ok] - want[ok]).max() <= 8 * ulp
    # and bit-exact against the oracle assembled from the restatements of OpenCV's own arithmetic
    q = dem.copy()
    q[q <= nv] = np.nan
    q = OP.area4(q)
    q[np.isnan(q)] = nv
    with np.errstate(all="ignore"):
        q = OP.fill_nan(q, nv, 256, 32, 24)
    q[q <= nv] = np.nan
    q = OP.cubic_resize(OP.area4(q), *dem.shape)
    q[np.isnan(q)] = nv
    np.testing.assert_array_equal(got, q)


def test_non_square_raster_and_process_map(msr, torch, tmp_path):
    """The reference transposes the extent of a non-square raster (:241) and crashes; here preprocess returns (H, W).
    processMap with preprocess on equals run() on the preprocessed DEM."""
    from moonsuperresolution_b200 import geotiff
    import toy_models
    rng = np.random.default_rng(9)
    h, w = 300, 420
    dem = (np.cumsum(np.cumsum(rng.standard_normal((h, w)), 0), 1) * 0.5 + 1500.0).astype(np.float32)
    img = rng.uniform(1, 255, (h, w)).astype(np.float32)
    dem[:2] = NV
    geotiff.write(str(tmp_path / "run-DEM.tif"), dem)
    geotiff.write(str(tmp_path / "run-DRG.tif"), img)
    (tmp_path / "out").mkdir()
    cfg = msr.DSRConfig(image_size=32, stride=8, batch_size=5, tile_size=128, no_value=NV, map_name="m",
                        save_path=str(tmp_path / "out"), source_folder_path=str(tmp_path))
    eng = msr.DEMSuperResolution(cfg, model=toy_models.ripple)
    eng.processMap()
    mean, _ = geotiff.read(str(tmp_path / "out" / "m_mean.tiff"))
    eng2 = msr.DEMSuperResolution(cfg, model=toy_models.ripple)
    eng2.setRasters(dem, img)
    eng2.preprocess()
    pre = eng2.dem.cpu().numpy()
    assert pre.shape == (h, w)
    cfg3 = msr.DSRConfig(image_size=32, stride=8, batch_size=5, tile_size=128, no_value=NV, preprocess=False)
    want = msr.DEMSuperResolution(cfg3, model=toy_models.ripple).run(pre, img)
    np.testing.assert_array_equal(mean, want[0])
    # preprocess=False leaves the DEM alone
    eng4 = msr.DEMSuperResolution(cfg3)
    eng4.setRasters(dem, img)
    eng4.preprocess()
    assert eng4.dem is not None and not torch.is_tensor(eng4.dem)
    # a rank holding only part of the raster cannot preprocess
    eng5 = msr.DEMSuperResolution(cfg)
    eng5.setRasters(dem[10:200], img[10:200], row_offset=10, full_height=h)
    with pytest.raises(ValueError):
        eng5.preprocess()


def test_preprocess_large_non_aligned_raster_against_stock_cv2(msr, torch):
    """A raster whose sides are not multiples of 16 (5003 x 7001: cvRound'ed intermediate extents, edge windows of the
    box filter, non-integer cubic scale) against the reference's sequence run through stock cv2 on the host, plus
    size-independent properties: a constant raster comes back constant, no_value regions grow by the filter footprint."""
    cv2 = pytest.importorskip("cv2")
    h, w = 5003, 7001
    gen = torch.Generator(device="cuda").manual_seed(5)
    dem = (torch.cumsum(torch.randn((h, w), generator=gen, device="cuda"), 1) * 0.7 + 2000.0).contiguous()
    dem[1000:1400, 2000:2600] = NV                                     # large hole: stays (no fill candidate)
    host = dem.cpu().numpy()
    eng = msr.DEMSuperResolution(msr.DSRConfig(no_value=NV))
    eng.setRasters(dem, dem)
    eng.preprocess()
    got = eng.dem.cpu().numpy()
    assert got.shape == (h, w)
    with np.errstate(all="ignore"):
        want = OP.reference_preprocess(host, NV)                        # cv2 (IPP) + scipy, fix_shape=True
    np.testing.assert_array_equal(got <= NV, want <= NV)
    ok = want > NV
    q = host.copy()
    q[q <= NV] = np.nan
    q = OP.area4(OP.area4(q))
    step = max(np.nanmax(np.abs(np.diff(q, axis=0))), np.nanmax(np.abs(np.diff(q, axis=1))))
    ulp = np.float32(np.abs(want[ok]).max()) * np.float32(2.0 ** -23)
    assert np.abs(got[ok] - want[ok]).max() <= 8 * ulp + 2.0 ** -21 * max(q.shape) * step
    assert (got[1100:1300, 2100:2500] == NV).all() and (got[:900, :1900] > NV).all()
    # constant raster: box mean and cubic weights (which sum to 1 up to rounding) keep it within 2 ulp
    const = torch.full((h, w), 1234.5, device="cuda")
    eng.setRasters(const, const)
    eng.preprocess()
    assert (eng.dem - 1234.5).abs().max().item() <= 2 * 1234.5 * 2.0 ** -23
