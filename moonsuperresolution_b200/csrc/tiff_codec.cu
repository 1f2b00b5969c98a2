// Host-side TIFF strip / tile codec of the raster boundary: LZW (TIFF flavour: MSB-first codes, 9..12 bits, early
// change) and the horizontal / floating-point predictors, multi-threaded over strips.
//
// Replaces what GDAL does for the reference: reading band 1 of LZW GeoTIFFs (process_full_tiles.py:158-182) and writing
// 'COMPRESS=LZW', 'PREDICTOR=2' GeoTIFFs (process_full_tiles.py:481-531).  The container (IFD, tags, GeoTIFF keys) is
// handled by moonsuperresolution_b200/geotiff.py; this file only transforms bytes.  No GPU work here.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kClear = 256, kEoi = 257, kFirst = 258, kMaxCode = 4094;  // libtiff resets at 4094 (CODE_MAX - 1)

// ---- LZW encoder ------------------------------------------------------------------------------------------------------
struct BitWriter {
  uint8_t* out;
  int64_t cap, n = 0;
  uint64_t acc = 0;
  int bits = 0;
  bool overflow = false;
  void put(int code, int width) {
    acc = (acc << width) | (uint64_t)code;
    bits += width;
    while (bits >= 8) {
      if (n >= cap) { overflow = true; bits -= 8; continue; }
      out[n++] = (uint8_t)(acc >> (bits - 8));
      bits -= 8;
    }
  }
  void flush() {
    if (bits > 0) {
      if (n >= cap) overflow = true;
      else out[n++] = (uint8_t)(acc << (8 - bits));
      bits = 0;
    }
  }
};

int64_t lzw_encode(const uint8_t* src, int64_t len, uint8_t* dst, int64_t cap) {
  // string table: open-addressing hash of (prefix code, byte) -> code.  One 32-bit word per slot -- key (20 bits) above
  // the code (12 bits) -- in 16384 slots: 64 KB, at most 23 % full (the table is reset at 4094 codes), and
  // a reset is one 64 KB memset per ~10-20 KB of input.
  constexpr int kSlots = 1 << 14;
  constexpr uint32_t kEmpty = 0xFFFFFFFFu;
  std::vector<uint32_t> table(kSlots);
  auto reset = [&] { memset(table.data(), 0xFF, sizeof(uint32_t) * kSlots); };
  BitWriter bw{dst, cap};
  reset();
  int next = kFirst, width = 9;
  bw.put(kClear, width);
  if (len == 0) {
    bw.put(kEoi, width);
    bw.flush();
    return bw.overflow ? -1 : bw.n;
  }
  int prefix = src[0];
  for (int64_t i = 1; i < len; ++i) {
    const int c = src[i];
    const uint32_t key = ((uint32_t)prefix << 8) | (uint32_t)c;
    uint32_t h = (key * 2654435761u) >> 18;
    uint32_t e;
    bool found = false;
    while ((e = table[h]) != kEmpty) {
      if ((e >> 12) == key) { found = true; break; }
      h = (h + 1) & (kSlots - 1);
    }
    if (found) {
      prefix = (int)(e & 0xFFFu);
      continue;
    }
    bw.put(prefix, width);
    table[h] = (key << 12) | (uint32_t)next++;
    // libtiff's rule ("early change"): the decoder is one table entry behind the encoder and widens its codes when its
    // next free entry reaches 2^width - 1, i.e. when the encoder's reaches 2^width; the table is reset at 4094
    if (next == kMaxCode) {
      bw.put(kClear, width);
      reset();
      next = kFirst;
      width = 9;
    } else if (next == (1 << width)) {
      ++width;
    }
    prefix = c;
  }
  bw.put(prefix, width);
  // the decoder adds one more table entry for this last code before it reads EOI (libtiff LZWPostEncode)
  ++next;
  if (next == kMaxCode) {
    bw.put(kClear, width);
    width = 9;
  } else if (next == (1 << width)) {
    ++width;
  }
  bw.put(kEoi, width);
  bw.flush();
  return bw.overflow ? -1 : bw.n;
}

// ---- LZW decoder ------------------------------------------------------------------------------------------------------
int64_t lzw_decode(const uint8_t* src, int64_t len, uint8_t* dst, int64_t cap) {
  std::vector<uint16_t> prefix(4096);
  std::vector<uint8_t> suffix(4096), first(4096);
  std::vector<uint16_t> length(4096);
  for (int i = 0; i < 256; ++i) {
    suffix[i] = first[i] = (uint8_t)i;
    length[i] = 1;
  }
  uint64_t acc = 0;
  int bits = 0, width = 9, next = kFirst, prev = -1;
  int64_t in = 0, out = 0;
  while (true) {
    while (bits < width && in < len) {
      acc = (acc << 8) | src[in++];
      bits += 8;
    }
    if (bits < width) break;   // truncated stream: stop quietly (libtiff does the same with a warning)
    const int code = (int)((acc >> (bits - width)) & ((1u << width) - 1));
    bits -= width;
    if (code == kEoi) break;
    if (code == kClear) {
      width = 9;
      next = kFirst;
      prev = -1;
      continue;
    }
    if (prev < 0) {
      if (code >= 256) return -1;
      if (out < cap) dst[out] = (uint8_t)code;
      ++out;
      prev = code;
      continue;
    }
    int cur = code;
    uint8_t fc;
    int64_t n;
    if (code < next) {
      n = length[code];
      fc = first[code];
    } else if (code == next) {   // KwKwK
      n = length[prev] + 1;
      fc = first[prev];
      cur = prev;
    } else {
      return -1;
    }
    // write the string backwards
    int64_t pos = out + n - 1;
    if (code == next) {
      if (pos < cap) dst[pos] = fc;
      --pos;
    }
    int c2 = cur;
    while (true) {
      if (pos < cap) dst[pos] = suffix[c2];
      --pos;
      if (length[c2] == 1) break;
      c2 = prefix[c2];
    }
    out += n;
    if (next < 4096) {
      prefix[next] = (uint16_t)prev;
      suffix[next] = fc;
      first[next] = first[prev];
      length[next] = (uint16_t)(length[prev] + 1);
      ++next;
      if (next == (1 << width) - 1 && width < 12) ++width;
    }
    prev = code;
  }
  return out;
}

// ---- predictors (in place, per row of `row_bytes`) ------------------------------------------------------------------------
template <typename T>
void diff_rows(uint8_t* buf, int64_t row_bytes, int rows, bool encode) {
  const int64_t n = row_bytes / (int64_t)sizeof(T);
  for (int r = 0; r < rows; ++r) {
    T* p = reinterpret_cast<T*>(buf + (int64_t)r * row_bytes);
    if (encode) {
      for (int64_t i = n - 1; i > 0; --i) p[i] = (T)(p[i] - p[i - 1]);
    } else {
      for (int64_t i = 1; i < n; ++i) p[i] = (T)(p[i] + p[i - 1]);
    }
  }
}

void horizontal_predictor(uint8_t* buf, int64_t row_bytes, int rows, int sample_bytes, bool encode) {
  switch (sample_bytes) {
    case 1: diff_rows<uint8_t>(buf, row_bytes, rows, encode); break;
    case 2: diff_rows<uint16_t>(buf, row_bytes, rows, encode); break;
    case 4: diff_rows<uint32_t>(buf, row_bytes, rows, encode); break;
    default: diff_rows<uint64_t>(buf, row_bytes, rows, encode); break;
  }
}

// TIFF predictor 3 (floating point), decode only: bytes were shuffled into big-endian byte planes, then differenced
void float_predictor_decode(uint8_t* buf, int64_t row_bytes, int rows, int sample_bytes) {
  std::vector<uint8_t> tmp((size_t)row_bytes);
  const int64_t n = row_bytes / sample_bytes;
  for (int r = 0; r < rows; ++r) {
    uint8_t* p = buf + (int64_t)r * row_bytes;
    for (int64_t i = 1; i < row_bytes; ++i) p[i] = (uint8_t)(p[i] + p[i - 1]);
    memcpy(tmp.data(), p, (size_t)row_bytes);
    for (int64_t i = 0; i < n; ++i)
      for (int b = 0; b < sample_bytes; ++b) p[i * sample_bytes + b] = tmp[(size_t)(sample_bytes - 1 - b) * n + i];
  }
}

template <typename F>
void parallel_for(int n, int n_threads, F&& fn) {
  n_threads = std::max(1, std::min(n_threads, n));
  if (n_threads == 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t)
    pool.emplace_back([&] {
      for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    });
  for (auto& th : pool) th.join();
}

}  // namespace

using namespace msr;

extern "C" int64_t msr_tiff_lzw_bound(int64_t raw_bytes) { return raw_bytes + raw_bytes / 2 + 1024; }

// Compress the rows of a (rows, row_bytes) host raster as strips of rows_per_strip rows.
//   compression 1 (none) / 5 (LZW); predictor 1 / 2 (horizontal differencing on sample_bytes-wide integers)
//   h_out holds n_strips slots of `slot_bytes` each (>= msr_tiff_lzw_bound of a strip); h_sizes receives the byte counts
extern "C" int msr_tiff_encode_strips(const uint8_t* h_raster, int64_t row_bytes, int rows, int rows_per_strip,
                                      int sample_bytes, int compression, int predictor, uint8_t* h_out,
                                      int64_t slot_bytes, int64_t* h_sizes, int n_threads) {
  MSR_REQUIRE(h_raster && h_out && h_sizes && row_bytes > 0 && rows > 0 && rows_per_strip > 0, "tiff_encode: bad arguments");
  MSR_REQUIRE(compression == 1 || compression == 5, "tiff_encode: compression must be 1 (none) or 5 (LZW)");
  MSR_REQUIRE(predictor == 1 || predictor == 2, "tiff_encode: predictor must be 1 or 2");
  MSR_REQUIRE(sample_bytes == 1 || sample_bytes == 2 || sample_bytes == 4 || sample_bytes == 8, "tiff_encode: sample size");
  const int n_strips = (rows + rows_per_strip - 1) / rows_per_strip;
  std::atomic<int> failed(0);
  parallel_for(n_strips, n_threads, [&](int s) {
    const int r0 = s * rows_per_strip, nr = std::min(rows_per_strip, rows - r0);
    const int64_t raw = (int64_t)nr * row_bytes;
    uint8_t* dst = h_out + (int64_t)s * slot_bytes;
    std::vector<uint8_t> tmp;
    const uint8_t* src = h_raster + (int64_t)r0 * row_bytes;
    if (predictor == 2) {
      tmp.assign(src, src + raw);
      horizontal_predictor(tmp.data(), row_bytes, nr, sample_bytes, true);
      src = tmp.data();
    }
    if (compression == 1) {
      if (raw > slot_bytes) { failed = 1; return; }
      memcpy(dst, src, (size_t)raw);
      h_sizes[s] = raw;
    } else {
      const int64_t n = lzw_encode(src, raw, dst, slot_bytes);
      if (n < 0) { failed = 1; return; }
      h_sizes[s] = n;
    }
  });
  if (failed) return fail(MSR_E_INVALID, "tiff_encode: output slot too small");
  return MSR_OK;
}

// Decode n chunks (strips or tiles) of a TIFF file image into a (rows, row_bytes) host raster.
//   chunk i: file bytes [offsets[i], offsets[i] + counts[i]); it decodes to chunk_rows x chunk_row_bytes and is pasted at
//   raster row dst_row[i], byte column dst_col[i], clipped to the raster.
extern "C" int msr_tiff_decode_chunks(const uint8_t* h_file, int64_t file_bytes, const int64_t* offsets,
                                      const int64_t* counts, const int64_t* dst_row, const int64_t* dst_col, int n_chunks,
                                      int chunk_rows, int64_t chunk_row_bytes, int sample_bytes, int compression,
                                      int predictor, uint8_t* h_raster, int64_t row_bytes, int64_t rows, int n_threads) {
  MSR_REQUIRE(h_file && offsets && counts && dst_row && dst_col && h_raster, "tiff_decode: null pointer");
  MSR_REQUIRE(compression == 1 || compression == 5, "tiff_decode: only uncompressed and LZW TIFFs are supported");
  MSR_REQUIRE(predictor >= 1 && predictor <= 3, "tiff_decode: unknown predictor");
  MSR_REQUIRE(chunk_rows > 0 && chunk_row_bytes > 0 && n_chunks >= 0, "tiff_decode: bad geometry");
  std::atomic<int> failed(0);
  parallel_for(n_chunks, n_threads, [&](int i) {
    if (offsets[i] < 0 || counts[i] < 0 || offsets[i] + counts[i] > file_bytes) { failed = 1; return; }
    const int64_t raw = (int64_t)chunk_rows * chunk_row_bytes;
    std::vector<uint8_t> buf((size_t)raw, 0);
    const uint8_t* src = h_file + offsets[i];
    if (compression == 1) {
      memcpy(buf.data(), src, (size_t)std::min<int64_t>(raw, counts[i]));
    } else if (lzw_decode(src, counts[i], buf.data(), raw) < 0) {
      failed = 2;
      return;
    }
    if (predictor == 2) horizontal_predictor(buf.data(), chunk_row_bytes, chunk_rows, sample_bytes, false);
    else if (predictor == 3) float_predictor_decode(buf.data(), chunk_row_bytes, chunk_rows, sample_bytes);
    // paste, clipped to the raster on every side (a chunk may start above row 0 when only a band of rows is read)
    const int64_t nr = std::min<int64_t>(chunk_rows, rows - dst_row[i]);
    const int64_t nb = std::min<int64_t>(chunk_row_bytes, row_bytes - dst_col[i]);
    if (dst_col[i] < 0 || nb <= 0) return;
    for (int64_t r = std::max<int64_t>(0, -dst_row[i]); r < nr; ++r)
      memcpy(h_raster + (dst_row[i] + r) * row_bytes + dst_col[i], buf.data() + r * chunk_row_bytes, (size_t)nb);
  });
  if (failed == 1) return fail(MSR_E_INVALID, "tiff_decode: chunk outside the file");
  if (failed == 2) return fail(MSR_E_INVALID, "tiff_decode: corrupt LZW stream");
  return MSR_OK;
}
