"""TIFF container + native LZW / predictor codec (host code of libmoonsr.so; no GPU needed).  OpenCV's libtiff is the
independent implementation on the other side of each round trip."""
import numpy as np
import pytest

from moonsuperresolution_b200 import geotiff

cv2 = pytest.importorskip("cv2")


def rasters():
    rng = np.random.default_rng(0)
    smooth = np.cumsum(np.cumsum(rng.standard_normal((301, 517)), 0), 1).astype(np.float32)
    smooth[40:60, 100:130] = -32768.0
    noisy = rng.uniform(-1e4, 1e4, (97, 1031)).astype(np.float32)
    mask = (rng.uniform(0, 1, (513, 260)) > 0.3).astype(np.uint16)
    flat = np.zeros((70, 5000), np.uint8)
    return {"smooth": smooth, "noisy": noisy, "mask16": mask, "flat8": flat}


@pytest.mark.parametrize("name", ["smooth", "noisy", "mask16", "flat8"])
@pytest.mark.parametrize("compress,predictor", [("lzw", 2), ("lzw", 1), ("none", 1)])
def test_write_read_round_trip_and_opencv_reads_it(tmp_path, name, compress, predictor):
    a = rasters()[name]
    path = str(tmp_path / f"{name}.tif")
    geotiff.write(path, a, nodata=-32768.0, compress=compress, predictor=predictor, rows_per_strip=37)
    back, geo = geotiff.read(path)
    assert back.dtype == a.dtype
    np.testing.assert_array_equal(back, a)
    assert geotiff.TAG_GDAL_NODATA in geo
    # libtiff (through OpenCV) decodes our LZW + horizontal-predictor strips to the same samples
    other = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert other is not None and other.dtype == a.dtype
    np.testing.assert_array_equal(other, a)


@pytest.mark.parametrize("name", ["smooth", "noisy", "mask16", "flat8"])
def test_reads_what_libtiff_writes(tmp_path, name):
    """OpenCV writes LZW strips (with libtiff's predictor choice for the sample type); our decoder must agree."""
    a = rasters()[name]
    path = str(tmp_path / f"{name}_cv.tif")
    assert cv2.imwrite(path, a)
    back, _ = geotiff.read(path)
    np.testing.assert_array_equal(back, a)


def test_lzw_compresses_and_handles_table_resets(tmp_path):
    a = np.tile(np.arange(256, dtype=np.uint8), (300, 40))          # long repetitive rows: many table resets
    path = str(tmp_path / "rep.tif")
    geotiff.write(path, a, rows_per_strip=300, predictor=1)
    assert (tmp_path / "rep.tif").stat().st_size < a.nbytes // 4
    np.testing.assert_array_equal(geotiff.read(path)[0], a)
    np.testing.assert_array_equal(cv2.imread(path, cv2.IMREAD_UNCHANGED), a)


def test_geo_tags_pass_through(tmp_path):
    import struct
    a = rasters()["smooth"]
    geo = {33550: (12, 3, struct.pack("<3d", 5.0, 5.0, 0.0)),
           33922: (12, 6, struct.pack("<6d", 0, 0, 0, 1000.0, 2000.0, 0)),
           34735: (3, 4, struct.pack("<4H", 1, 1, 0, 0))}
    p1 = str(tmp_path / "a.tif")
    geotiff.write(p1, a, geo=geo, nodata=-32768.0)
    back, g = geotiff.read(p1)
    for tag in geo:
        assert g[tag] == geo[tag]
    assert g[geotiff.TAG_GDAL_NODATA][2].rstrip(b"\0") == b"-32768"


@pytest.mark.parametrize("name", ["smooth", "mask16", "flat8"])
@pytest.mark.parametrize("scheme", [8, 32946])
def test_reads_deflate_compressed_files(tmp_path, name, scheme):
    """GDAL's COMPRESS=DEFLATE (TIFF compression 8, legacy 32946), written here by libtiff through OpenCV."""
    a = rasters()[name]
    path = str(tmp_path / f"{name}_deflate.tif")
    assert cv2.imwrite(path, a, [cv2.IMWRITE_TIFF_COMPRESSION, scheme])
    back, _ = geotiff.read(path)
    assert back.dtype == a.dtype
    np.testing.assert_array_equal(back, a)
