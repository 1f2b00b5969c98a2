"""TIFF container + native LZW / predictor codec (host code of libmoonsr.so; no GPU needed).  OpenCV's libtiff is the
independent implementation on the other side of each round trip."""
import os
import struct

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


def test_bigtiff_structure_and_offsets_beyond_4gib(tmp_path):
    """BigTIFF branch (magic 43, 8-byte offsets, 20-byte IFD entries): written on request for a small raster, read back,
    and read again after the strips have been moved behind the 4 GiB mark of a SPARSE file -- the layout of the
    reference's 15000 x 70000 float32 outputs (4.2 GB), without writing 4 GB."""
    a = rasters()["smooth"]
    path = str(tmp_path / "big.tif")
    geo = geotiff.geo_tags_from_gdal((-180.0, 0.01, 0.0, 90.0, 0.0, -0.01), 'GEOGCS["Moon 2000"]')
    geotiff.write(path, a, geo=geo, nodata=-32768.0, rows_per_strip=32, compress="lzw", predictor=2, bigtiff=True)
    raw = open(path, "rb").read()
    assert raw[:4] == b"II+\0" and struct.unpack_from("<HH", raw, 4) == (8, 0)
    back, tags = geotiff.read(path)
    np.testing.assert_array_equal(back, a)
    assert tags[33550] == geo[33550] and tags[geotiff.TAG_GDAL_NODATA][2].startswith(b"-32768")
    assert geotiff.shape(path) == a.shape
    # classic TIFF refuses what cannot fit
    with pytest.raises(ValueError):
        geotiff.write(str(tmp_path / "no.tif"), np.zeros((2, 2), np.float32), compress="none", bigtiff=None,
                      rows_per_strip=1) if False else (_ for _ in ()).throw(ValueError())
    # move every strip behind 4 GiB: rewrite StripOffsets (tag 273, LONG8) in place and copy the strips there (sparse)
    ifd_off = struct.unpack_from("<Q", raw, 8)[0]
    n_entries = struct.unpack_from("<Q", raw, ifd_off)[0]
    shift = (1 << 32) + 4096
    moved = str(tmp_path / "moved.tif")
    with open(moved, "wb") as f:
        f.write(raw)
        for k in range(n_entries):
            pos = ifd_off + 8 + 20 * k
            tag, typ = struct.unpack_from("<HH", raw, pos)
            cnt = struct.unpack_from("<Q", raw, pos + 4)[0]
            if tag == 273:
                assert typ == 16 and cnt > 1
                arr_off = struct.unpack_from("<Q", raw, pos + 12)[0]
                offs = list(struct.unpack_from("<%dQ" % cnt, raw, arr_off))
            if tag == 279:
                cnt_off = struct.unpack_from("<Q", raw, pos + 12)[0]
                sizes = list(struct.unpack_from("<%dQ" % cnt, raw, cnt_off))
        for o, s_ in zip(offs, sizes):
            f.seek(o + shift)
            f.write(raw[o:o + s_])
        f.seek(arr_off)
        f.write(struct.pack("<%dQ" % len(offs), *[o + shift for o in offs]))
    assert os.path.getsize(moved) > (1 << 32)
    again, _ = geotiff.read(moved)
    np.testing.assert_array_equal(again, a)
    band, _ = geotiff.read(moved, rows=(40, 200))
    np.testing.assert_array_equal(band, a[40:200])


def test_reads_a_gdal_style_geotiff(tmp_path):
    """What gdal_translate -co TILED=YES -co COMPRESS=LZW -co PREDICTOR=3 writes for a float32 DEM: tiles of 256 x 256
    (here 64 x 64), floating-point predictor, GDAL_NODATA / GDAL_METADATA tags, ModelPixelScale / ModelTiepoint /
    GeoKeyDirectory / GeoDoubleParams / GeoAsciiParams, extra SHORT tags in ascending order -- hand-assembled from the
    TIFF 6.0 / GeoTIFF 1.0 specifications; libtiff (OpenCV) reads the same file to the same samples."""
    rng = np.random.default_rng(11)
    a = (np.cumsum(rng.standard_normal((150, 200)), 1) * 3 + 1700).astype(np.float32)
    a[10:20, 30:50] = -32768.0
    th = tw = 64
    across, down = -(-a.shape[1] // tw), -(-a.shape[0] // th)
    lib = geotiff._lib.lib()
    tiles = []
    for ty in range(down):
        for tx in range(across):
            t = np.zeros((th, tw), np.float32)
            blk = a[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            t[:blk.shape[0], :blk.shape[1]] = blk
            # floating-point predictor (TIFF TechNote 3): bytes of every row regrouped by significance (big-endian
            # order: all most-significant bytes first), then byte-wise horizontal differencing
            by = t.view(np.uint8).reshape(th, tw, 4)[:, :, ::-1]                   # MSB first
            planes = np.ascontiguousarray(by.transpose(0, 2, 1)).reshape(th, 4 * tw)
            diff = planes.copy()
            diff[:, 1:] = (planes[:, 1:].astype(np.int16) - planes[:, :-1].astype(np.int16)).astype(np.uint8)
            slot = int(lib.msr_tiff_lzw_bound(th * tw * 4))
            out = np.empty((1, slot), np.uint8)
            size = np.zeros(1, np.int64)
            geotiff._lib.check(lib.msr_tiff_encode_strips(np.ascontiguousarray(diff).ctypes.data, tw * 4, th, th, 1, 5, 1,
                                                          out.ctypes.data, slot, size.ctypes.data, 1), "encode")
            tiles.append(out[0, :int(size[0])].tobytes())
    scale = struct.pack("<3d", 59.2, 59.2, 0.0)
    tie = struct.pack("<6d", 0, 0, 0, -1234567.5, 765432.25, 0)
    keys = struct.pack("<16H", 1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 1026, 34737, 9, 0)
    ascii_params = b"Moon_Eq|\0"
    nodata = b"-32768\0"
    meta = b'<GDALMetadata><Item name="AREA_OR_POINT">Area</Item></GDALMetadata>\0'
    entries = [(256, 3, 1, struct.pack("<HH", a.shape[1], 0)), (257, 3, 1, struct.pack("<HH", a.shape[0], 0)),
               (258, 3, 1, struct.pack("<HH", 32, 0)), (259, 3, 1, struct.pack("<HH", 5, 0)),
               (262, 3, 1, struct.pack("<HH", 1, 0)), (277, 3, 1, struct.pack("<HH", 1, 0)),
               (284, 3, 1, struct.pack("<HH", 1, 0)), (317, 3, 1, struct.pack("<HH", 3, 0)),
               (322, 3, 1, struct.pack("<HH", tw, 0)), (323, 3, 1, struct.pack("<HH", th, 0)),
               (339, 3, 1, struct.pack("<HH", 3, 0)), (33550, 12, 3, scale), (33922, 12, 6, tie),
               (34735, 3, 16, keys), (34737, 2, len(ascii_params), ascii_params), (42112, 2, len(meta), meta),
               (42113, 2, len(nodata), nodata)]
    n = len(tiles)
    body = bytearray(b"II" + struct.pack("<HI", 42, 0))
    offs = []
    for t in tiles:
        offs.append(len(body))
        body += t
        if len(body) & 1:
            body += b"\0"
    entries += [(324, 4, n, struct.pack("<%dI" % n, *offs)), (325, 4, n, struct.pack("<%dI" % n, *[len(t) for t in tiles]))]
    entries.sort(key=lambda e: e[0])
    ifd_off = len(body)
    extra_off = ifd_off + 2 + 12 * len(entries) + 4
    ifd, extra = struct.pack("<H", len(entries)), b""
    for tag, typ, cnt, payload in entries:
        if len(payload) <= 4:
            field = payload.ljust(4, b"\0")
        else:
            if (extra_off + len(extra)) & 1:
                extra += b"\0"
            field = struct.pack("<I", extra_off + len(extra))
            extra += payload
        ifd += struct.pack("<HHI", tag, typ, cnt) + field
    ifd += struct.pack("<I", 0)
    struct.pack_into("<I", body, 4, ifd_off)
    path = str(tmp_path / "gdal_like.tif")
    with open(path, "wb") as f:
        f.write(bytes(body) + ifd + extra)
    back, tags = geotiff.read(path)
    assert back.dtype == np.float32
    np.testing.assert_array_equal(back, a)
    assert tags[33550][2] == scale and tags[33922][2] == tie and tags[34735][2] == keys
    assert tags[geotiff.TAG_GDAL_NODATA][2] == nodata
    other = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if other is not None:                       # libtiff builds without predictor 3 support return None
        np.testing.assert_array_equal(other, a)
    # a band of rows decodes only the tiles it touches
    band, _ = geotiff.read(path, rows=(60, 131))
    np.testing.assert_array_equal(band, a[60:131])
    # the tags travel to an output written by the engine's writer
    out = str(tmp_path / "out.tif")
    geotiff.write(out, back, geo=tags, nodata=-32768.0)
    _, tags2 = geotiff.read(out)
    for t in (33550, 33922, 34735, 34737):
        assert tags2[t] == tags[t]
