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


def _write_deflate_tiff(path, a, scheme, rows_per_strip, predictor, tiled=False, tile=64):
    """Minimal classic-TIFF writer with zlib-compressed strips or tiles (what GDAL's COMPRESS=DEFLATE produces); horizontal
    differencing (predictor 2) for integer samples.  OpenCV ignores IMWRITE_TIFF_COMPRESSION=8 for these dtypes, so the
    file is assembled by hand."""
    import struct
    import zlib
    h, w = a.shape
    fmt = {"f": 3, "u": 1, "i": 2}[a.dtype.kind]
    chunks = []
    if tiled:
        for y in range(0, h, tile):
            for x in range(0, w, tile):
                t = np.zeros((tile, tile), a.dtype)
                blk = a[y:y + tile, x:x + tile]
                t[:blk.shape[0], :blk.shape[1]] = blk
                chunks.append(t)
    else:
        chunks = [a[y:y + rows_per_strip] for y in range(0, h, rows_per_strip)]
    payloads = []
    for c in chunks:
        c = np.ascontiguousarray(c)
        if predictor == 2:
            d = c.copy()
            d[:, 1:] = c[:, 1:] - c[:, :-1]          # modular arithmetic of the unsigned sample type
            c = d
        payloads.append(zlib.compress(c.tobytes(), 6))
    entries = [(256, 4, w), (257, 4, h), (258, 3, a.dtype.itemsize * 8), (259, 3, scheme), (262, 3, 1), (277, 3, 1),
               (317, 3, predictor), (339, 3, fmt)]
    if tiled:
        entries += [(322, 4, tile), (323, 4, tile)]
    else:
        entries += [(278, 4, rows_per_strip)]
    off_tag, cnt_tag = (324, 325) if tiled else (273, 279)
    n = len(payloads)
    n_tags = len(entries) + 2
    data_off = 8
    offsets, pos = [], data_off
    for p_ in payloads:
        offsets.append(pos)
        pos += len(p_)
    pos += pos & 1
    ifd_off = pos
    arrays_off = ifd_off + 2 + 12 * n_tags + 4
    body = b"II" + struct.pack("<HI", 42, ifd_off) + b"".join(payloads)
    body += b"\0" * (ifd_off - len(body))
    recs = [(t, ty, 1, struct.pack("<I", v) if ty == 4 else struct.pack("<HH", v, 0)) for t, ty, v in entries]
    if n == 1:
        recs += [(off_tag, 4, 1, struct.pack("<I", offsets[0])), (cnt_tag, 4, 1, struct.pack("<I", len(payloads[0])))]
        tail = b""
    else:
        recs += [(off_tag, 4, n, struct.pack("<I", arrays_off)), (cnt_tag, 4, n, struct.pack("<I", arrays_off + 4 * n))]
        tail = struct.pack("<%dI" % n, *offsets) + struct.pack("<%dI" % n, *[len(p_) for p_ in payloads])
    recs.sort(key=lambda r: r[0])
    ifd = struct.pack("<H", n_tags) + b"".join(struct.pack("<HHI", t, ty, c) + v for t, ty, c, v in recs) + struct.pack("<I", 0)
    with open(path, "wb") as f:
        f.write(body + ifd + tail)


@pytest.mark.parametrize("name,predictor", [("smooth", 1), ("mask16", 2), ("flat8", 2), ("noisy", 1)])
@pytest.mark.parametrize("scheme,tiled", [(8, False), (32946, False), (8, True)])
def test_reads_deflate_compressed_files(tmp_path, name, predictor, scheme, tiled):
    """GDAL's COMPRESS=DEFLATE (TIFF compression 8, legacy 32946), strips and tiles, with and without the horizontal
    predictor; libtiff (through OpenCV) reads the same hand-assembled file to the same samples."""
    a = rasters()[name]
    path = str(tmp_path / f"{name}_deflate.tif")
    _write_deflate_tiff(path, a, scheme, rows_per_strip=41, predictor=predictor, tiled=tiled)
    back, _ = geotiff.read(path)
    assert back.dtype == a.dtype
    np.testing.assert_array_equal(back, a)
    other = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert other is not None
    np.testing.assert_array_equal(other, a)


@pytest.mark.parametrize("layout", ["lzw_strips", "deflate_tiles", "plain_strips"])
def test_reading_a_band_of_rows(tmp_path, layout):
    """read(path, rows=(r0, r1)): only the strips / tiles that intersect the band are decoded; chunks that start above
    the band are clipped at the top, the last ones at the bottom."""
    a = rasters()["smooth"]                                   # 301 x 517
    path = str(tmp_path / "band.tif")
    if layout == "lzw_strips":
        geotiff.write(path, a, compress="lzw", predictor=2, rows_per_strip=37)
    elif layout == "plain_strips":
        geotiff.write(path, a, compress="none", predictor=1, rows_per_strip=16)
    else:
        _write_deflate_tiff(path, a, 8, rows_per_strip=0, predictor=1, tiled=True, tile=64)
    for r0, r1 in [(0, 301), (0, 1), (40, 41), (36, 38), (100, 250), (290, 301), (300, 301), (120, 120)]:
        band, geo = geotiff.read(path, rows=(r0, r1))
        assert band.shape == (r1 - r0, a.shape[1]) and band.dtype == a.dtype
        np.testing.assert_array_equal(band, a[r0:r1])
    with pytest.raises(ValueError):
        geotiff.read(path, rows=(10, 400))


def test_shape_from_header(tmp_path):
    a = rasters()["noisy"]
    path = str(tmp_path / "s.tif")
    geotiff.write(path, a)
    assert geotiff.shape(path) == a.shape
