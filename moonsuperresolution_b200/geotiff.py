"""TIFF / BigTIFF raster container I/O for the engine's boundary (no GDAL in this stack).

The reference reads band 1 of two GeoTIFFs with GDAL (process_full_tiles.py:158-182) and writes
``COMPRESS=LZW, PREDICTOR=2`` GeoTIFFs (process_full_tiles.py:481-531).  This module handles the container --
single-band, little-endian, classic TIFF below 4 GiB and BigTIFF above, strips or tiles on read, strips on write -- and
carries the GeoTIFF tags of the input DEM through to the outputs verbatim.  The byte transforms (LZW, horizontal and
floating-point predictors) run multi-threaded in libmoonsr.so (csrc/tiff_codec.cu, ``msr_tiff_*``).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib

# tags carried from the input DEM to the outputs (GeoTIFF + GDAL metadata)
GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737)
TAG_GDAL_NODATA = 42113

_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_SAMPLE = {  # numpy dtype -> (SampleFormat, BitsPerSample)
    np.dtype(np.uint8): (1, 8), np.dtype(np.uint16): (1, 16), np.dtype(np.int16): (2, 16),
    np.dtype(np.uint32): (1, 32), np.dtype(np.int32): (2, 32), np.dtype(np.float32): (3, 32),
    np.dtype(np.float64): (3, 64),
}


def _threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def geo_tags_from_gdal(geo_transform, projection_wkt: Optional[str] = None) -> Dict[int, Tuple[int, int, bytes]]:
    """GeoTIFF tags for the geo-referencing types the reference holds (process_full_tiles.py:177-178): GDAL's affine
    ``GetGeoTransform()`` 6-tuple (x0, dx, rx, y0, ry, dy) and the ``GetProjection()`` WKT string.  North-up rasters get
    ModelPixelScale + ModelTiepoint, rotated ones ModelTransformation; the WKT travels as the GTCitation key of a
    GeoKeyDirectory (model type from the WKT's root keyword, RasterPixelIsArea) -- no EPSG lookup is attempted."""
    gt = [float(v) for v in geo_transform]
    if len(gt) != 6:
        raise ValueError("geo_transform must be GDAL's 6-number affine transform (or the tag dict geotiff.read returns)")
    x0, dx, rx, y0, ry, dy = gt
    tags: Dict[int, Tuple[int, int, bytes]] = {}
    if rx == 0.0 and ry == 0.0:
        tags[33550] = (12, 3, struct.pack("<3d", dx, -dy, 0.0))
        tags[33922] = (12, 6, struct.pack("<6d", 0.0, 0.0, 0.0, x0, y0, 0.0))
    else:
        tags[34264] = (12, 16, struct.pack("<16d", dx, rx, 0.0, x0, ry, dy, 0.0, y0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                                           1.0))
    keys = [(1025, 0, 1, 1)]                                   # GTRasterTypeGeoKey = RasterPixelIsArea
    ascii_params = b""
    if projection_wkt:
        wkt = str(projection_wkt)
        model = 1 if wkt.lstrip().upper().startswith("PROJCS") else 2 if wkt.lstrip().upper().startswith("GEOGCS") else 32767
        keys.insert(0, (1024, 0, 1, model))                    # GTModelTypeGeoKey
        ascii_params = wkt.encode("ascii", "replace") + b"|"
        keys.append((1026, 34737, len(ascii_params), 0))       # GTCitationGeoKey -> GeoAsciiParams
    keys.sort()
    directory = [1, 1, 0, len(keys)] + [v for k in keys for v in k]
    tags[34735] = (3, len(directory), struct.pack("<%dH" % len(directory), *directory))
    if ascii_params:
        tags[34737] = (2, len(ascii_params) + 1, ascii_params + b"\0")
    return tags


def write(path: str, data: np.ndarray, geo: Optional[Dict[int, Tuple[int, int, bytes]]] = None,
          nodata: Optional[float] = None, rows_per_strip: int = 64, compress: str = "lzw", predictor: int = 2,
          bigtiff: Optional[bool] = None) -> None:
    """Writes a 2-D array as a single-band stripped TIFF (BigTIFF when it would not fit in 4 GiB, or when ``bigtiff`` is
    True -- GDAL's BIGTIFF=YES).  Defaults follow the reference's GDAL options ``COMPRESS=LZW, PREDICTOR=2``
    (process_full_tiles.py:521); ``compress="none"`` writes raw strips."""
    a = np.ascontiguousarray(data)
    if a.ndim == 3 and a.shape[2] == 1:
        a = np.ascontiguousarray(a[:, :, 0])
    if a.ndim != 2:
        raise ValueError("only single-band 2-D rasters are supported")
    if np.dtype(a.dtype.name) not in _SAMPLE:
        raise ValueError(f"unsupported dtype {a.dtype}")
    if compress not in ("lzw", "none"):
        raise ValueError("compress must be 'lzw' or 'none'")
    a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False))
    fmt, bits = _SAMPLE[np.dtype(a.dtype.name)]
    h, w = a.shape
    n_strips = -(-h // rows_per_strip)
    row_bytes = w * a.dtype.itemsize
    compression = 5 if compress == "lzw" else 1
    if compression == 1:
        predictor = 1
    lib = _lib.lib()
    slot = int(lib.msr_tiff_lzw_bound(rows_per_strip * row_bytes)) if compression == 5 else rows_per_strip * row_bytes
    packed = np.empty((n_strips, slot), np.uint8)
    sizes = np.zeros(n_strips, np.int64)
    _lib.check(lib.msr_tiff_encode_strips(a.ctypes.data, row_bytes, h, rows_per_strip, a.dtype.itemsize, compression,
                                          predictor, packed.ctypes.data, slot, sizes.ctypes.data, _threads()),
               "msr_tiff_encode_strips")
    counts = [int(c) for c in sizes]
    big = sum(counts) + 65536 + 16 * n_strips >= (1 << 32) if bigtiff is None else bool(bigtiff)
    if not big and sum(counts) + 65536 + 16 * n_strips >= (1 << 32):
        raise ValueError("raster does not fit a classic TIFF: pass bigtiff=None (auto) or True")
    off_t, off_code = ("<Q", 16) if big else ("<I", 4)
    osz = 8 if big else 4
    header = 16 if big else 8
    offsets, cur = [], header
    for c in counts:
        offsets.append(cur)
        cur += c
    data_end = cur

    entries = []  # (tag, type, count, payload bytes)

    def short(tag, v):
        entries.append((tag, 3, 1, struct.pack("<H", v)))

    def long_(tag, v):
        entries.append((tag, 4, 1, struct.pack("<I", v)))

    long_(256, w)
    long_(257, h)
    short(258, bits)
    short(259, compression)
    short(262, 1)        # BlackIsZero
    entries.append((273, off_code, n_strips, b"".join(struct.pack(off_t, o) for o in offsets)))
    short(277, 1)
    long_(278, rows_per_strip)
    entries.append((279, off_code, n_strips, b"".join(struct.pack(off_t, c) for c in counts)))
    short(284, 1)
    if predictor != 1:
        short(317, predictor)
    short(339, fmt)
    if geo:
        for tag in GEO_TAGS:
            if tag in geo:
                typ, cnt, payload = geo[tag]
                entries.append((tag, typ, cnt, payload))
    if nodata is not None:
        txt = (repr(float(nodata)) if float(nodata) != int(nodata) else str(int(nodata))).encode() + b"\0"
        entries.append((TAG_GDAL_NODATA, 2, len(txt), txt))
    entries.sort(key=lambda e: e[0])

    ifd_off = data_end + (data_end & 1)
    entry_sz = 20 if big else 12
    ifd_len = (8 if big else 2) + entry_sz * len(entries) + osz
    extra_off = ifd_off + ifd_len
    extra = b""
    ifd = struct.pack("<Q" if big else "<H", len(entries))
    for tag, typ, cnt, payload in entries:
        if len(payload) <= osz:
            field = payload + b"\0" * (osz - len(payload))
        else:
            if (extra_off + len(extra)) & 1:
                extra += b"\0"
            field = struct.pack(off_t, extra_off + len(extra))
            extra += payload
        ifd += struct.pack("<HH", tag, typ) + struct.pack("<Q" if big else "<I", cnt) + field
    ifd += struct.pack(off_t, 0)

    with open(path, "wb") as f:
        if big:
            f.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_off))
        else:
            f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        for s_, c in enumerate(counts):
            f.write(memoryview(packed[s_, :c]))
        if data_end & 1:
            f.write(b"\0")
        f.write(ifd)
        f.write(extra)


def _open(path: str):
    """Maps the file and parses its first IFD: (buffer, {tag: (type, count, payload bytes)}, ints(tag, default))."""
    import mmap
    with open(path, "rb") as f:
        # mapped, not read: a rank that decodes a band of rows only touches the pages of its own strips / tiles
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) if os.fstat(f.fileno()).st_size else b""
    if buf[:2] != b"II":
        raise ValueError("only little-endian TIFF is supported")
    magic = struct.unpack_from("<H", buf, 2)[0]
    if magic == 42:
        big, ifd_off = False, struct.unpack_from("<I", buf, 4)[0]
    elif magic == 43:
        big, ifd_off = True, struct.unpack_from("<Q", buf, 8)[0]
    else:
        raise ValueError("not a TIFF file")
    osz = 8 if big else 4
    n = struct.unpack_from("<Q" if big else "<H", buf, ifd_off)[0]
    pos = ifd_off + (8 if big else 2)
    tags = {}
    for _ in range(n):
        tag, typ = struct.unpack_from("<HH", buf, pos)
        cnt = struct.unpack_from("<Q" if big else "<I", buf, pos + 4)[0]
        fpos = pos + (12 if big else 8)
        size = _TYPE_SIZE[typ] * cnt
        if size > osz:
            fpos = struct.unpack_from("<Q" if big else "<I", buf, fpos)[0]
        tags[tag] = (typ, cnt, bytes(buf[fpos:fpos + size]))
        pos += 20 if big else 12

    def ints(tag, default=None):
        if tag not in tags:
            return default
        typ, cnt, payload = tags[tag]
        code = {1: "B", 3: "H", 4: "I", 16: "Q"}[typ]
        return list(struct.unpack("<" + code * cnt, payload))
    return buf, tags, ints


def shape(path: str) -> Tuple[int, int]:
    """(rows, cols) of the raster, from the header alone."""
    _, _, ints = _open(path)
    return ints(257)[0], ints(256)[0]


def read(path: str, rows=None):
    """Reads band 1 of a little-endian TIFF / BigTIFF: strips or tiles, uncompressed, LZW or Deflate, predictor 1 / 2 / 3.
    Returns (array, geo) where geo maps the GeoTIFF tag numbers present to (type, count, payload bytes).
    ``rows`` = (r0, r1) reads only that band of raster rows (a rank of a multi-GPU run decodes only the strips / tiles
    its band of tiles touches); the array then has r1 - r0 rows."""
    buf, tags, ints = _open(path)
    w, h = ints(256)[0], ints(257)[0]
    compression = ints(259, [1])[0]
    if compression not in (1, 5, 8, 32946):
        raise ValueError(f"TIFF compression {compression} is not supported (only none, LZW and Deflate)")
    predictor = ints(317, [1])[0]
    spp = ints(277, [1])[0]
    if spp != 1 and ints(284, [1])[0] != 2:
        raise ValueError("only band-sequential or single-band rasters are supported")
    bits, fmt = ints(258)[0], ints(339, [1])[0]
    dtype = {v: k for k, v in _SAMPLE.items()}[(fmt, bits)]
    itemsize = dtype.itemsize
    tiled = 322 in tags
    if tiled:
        tw, th = ints(322)[0], ints(323)[0]
        offs, cnts = ints(324), ints(325)
        across, down = -(-w // tw), -(-h // th)
        n = across * down                      # band 1 only
        rows0 = [(k // across) * th for k in range(n)]
        cols0 = [(k % across) * tw * itemsize for k in range(n)]
        chunk_rows, chunk_row_bytes = th, tw * itemsize
    else:
        rps = min(ints(278, [h])[0], h)
        offs, cnts = ints(273), ints(279)
        n = -(-h // rps)
        rows0 = [k * rps for k in range(n)]
        cols0 = [0] * n
        chunk_rows, chunk_row_bytes = rps, w * itemsize
    offs, cnts = list(offs[:n]), list(cnts[:n])
    r_lo, r_hi = 0, h
    if rows is not None:
        r_lo, r_hi = int(rows[0]), int(rows[1])
        if not (0 <= r_lo <= r_hi <= h):
            raise ValueError(f"rows {rows} outside the raster (0, {h})")
        keep = [k for k in range(n) if rows0[k] < r_hi and rows0[k] + chunk_rows > r_lo]
        offs, cnts = [offs[k] for k in keep], [cnts[k] for k in keep]
        rows0, cols0 = [rows0[k] - r_lo for k in keep], [cols0[k] for k in keep]
        n = len(keep)
    out = np.empty((r_hi - r_lo, w), dtype)
    geo = {t: tags[t] for t in GEO_TAGS + (TAG_GDAL_NODATA,) if t in tags}
    if n == 0 or out.size == 0:
        return out, geo
    if compression in (8, 32946):
        # Deflate (GDAL's COMPRESS=DEFLATE): inflate every chunk with zlib (releases the GIL: thread pool), then hand the
        # inflated image to the native codec as an uncompressed file -- it still undoes the predictor and pastes
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        raw = chunk_rows * chunk_row_bytes

        def inflate(k):
            if offs[k] < 0 or offs[k] + cnts[k] > len(buf):
                raise ValueError("TIFF chunk outside the file")
            return zlib.decompress(buf[offs[k]:offs[k] + cnts[k]])[:raw].ljust(raw, b"\0")
        with ThreadPoolExecutor(_threads()) as pool:
            buf = b"".join(pool.map(inflate, range(n)))
        offs, cnts, compression = [k * raw for k in range(n)], [raw] * n, 1
    fbuf = np.frombuffer(buf, np.uint8)
    a_off = np.asarray(offs, np.int64)
    a_cnt = np.asarray(cnts, np.int64)
    a_row = np.asarray(rows0, np.int64)
    a_col = np.asarray(cols0, np.int64)
    _lib.check(_lib.lib().msr_tiff_decode_chunks(fbuf.ctypes.data, fbuf.size, a_off.ctypes.data, a_cnt.ctypes.data,
                                                 a_row.ctypes.data, a_col.ctypes.data, n, chunk_rows, chunk_row_bytes,
                                                 itemsize, compression, predictor, out.ctypes.data, w * itemsize,
                                                 r_hi - r_lo, _threads()), "msr_tiff_decode_chunks")
    return out, geo
