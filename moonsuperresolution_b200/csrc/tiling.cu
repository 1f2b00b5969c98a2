// Tiling / blending kernels of the tiled full-DEM inference path (HBM-bound byte / fp32 / fp64 work).
//
// Reference: process_full_tiles.py  padInputs :246-267, getPatch :269-293, normalize :295-311,
//            makeGaussianKernel :347-361, rebuildTile :363-414, rebuildMap :541-545.
// All arithmetic that the reference performs in numpy is reproduced with explicitly rounded IEEE operations
// (__fsub_rn / __fdiv_rn / __dmul_rn ...) so that no FMA contraction changes a bit.
#include "common.cuh"

namespace msr {

// ------------------------------------------------------------------------------------------------------------------
// K0  padInputs: canvas = no_value everywhere, raster pasted at (off, off).  One float4 store per thread.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_inputs_kernel(const float* __restrict__ dem, const float* __restrict__ img,
                                                         int H, int W, float* __restrict__ dem_c,
                                                         float* __restrict__ img_c, int CH, int CW, int off_y,
                                                         int off_x, float nv) {
  const int64_t quads_per_row = (CW + 3) / 4;
  const int64_t total = quads_per_row * CH;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(q / quads_per_row);
    const int x0 = (int)(q % quads_per_row) * 4;
    const int sy = y - off_y;
    const bool row_in = (sy >= 0) && (sy < H);
    float d[4], m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sx = x0 + j - off_x;
      const bool in = row_in && sx >= 0 && sx < W;
      d[j] = in ? __ldg(dem + (int64_t)sy * W + sx) : nv;
      m[j] = in ? __ldg(img + (int64_t)sy * W + sx) : nv;
    }
    const int64_t o = (int64_t)y * CW + x0;
    if ((CW & 3) == 0) {
      *reinterpret_cast<float4*>(dem_c + o) = make_float4(d[0], d[1], d[2], d[3]);
      *reinterpret_cast<float4*>(img_c + o) = make_float4(m[0], m[1], m[2], m[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (x0 + j < CW) {
          dem_c[o + j] = d[j];
          img_c[o + j] = m[j];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1  validity summed-area table.  Pass A: per-row inclusive prefix of the invalid mask into sat[y+1][x+1];
//     pass B: per-column running sum.  int32, exact.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sat_rows_kernel(const float* __restrict__ img, const float* __restrict__ dem,
                                                       int CH, int CW, float nv, int32_t* __restrict__ sat) {
  const int y = blockIdx.x;
  if (y >= CH) return;
  __shared__ int warp_sums[8];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* out = sat + (int64_t)(y + 1) * (CW + 1);
  if (threadIdx.x == 0) out[0] = 0;
  const float* irow = img + (int64_t)y * CW;
  const float* drow = dem + (int64_t)y * CW;
  __syncthreads();
  for (int base = 0; base < CW; base += 256 * 4) {
    const int x0 = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      v[j] = (x < CW) ? ((irow[x] <= nv) || (drow[x] <= nv) ? 1 : 0) : 0;
    }
    int local = v[0] + v[1] + v[2] + v[3];
    int incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int wprefix = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < warp) wprefix += warp_sums[k];
    int run = carry_s + wprefix + incl - local;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      run += v[j];
      if (x0 + j < CW) out[x0 + j + 1] = run;
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = run;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sat_cols_kernel(int CH, int CW, int32_t* __restrict__ sat) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x > CW) return;
  const int64_t pitch = CW + 1;
  sat[x] = 0;
  int acc = 0;
  int y = 1;
  for (; y + 7 <= CH; y += 8) {
    int t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = sat[(int64_t)(y + j) * pitch + x];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc += t[j];
      sat[(int64_t)(y + j) * pitch + x] = acc;
    }
  }
  for (; y <= CH; ++y) {
    acc += sat[(int64_t)y * pitch + x];
    sat[(int64_t)y * pitch + x] = acc;
  }
}

__global__ void patch_validity_kernel(const int32_t* __restrict__ sat, int CH, int CW, const int32_t* __restrict__ xy,
                                      int n, int I, uint8_t* __restrict__ valid) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int x0 = xy[2 * k], y0 = xy[2 * k + 1];
  int x1 = min(x0 + I, CW), y1 = min(y0 + I, CH);
  x0 = max(x0, 0);
  y0 = max(y0, 0);
  if (x1 <= x0 || y1 <= y0) {  // empty window: numpy's .any() of an empty slice is False -> "valid"
    valid[k] = 1;
    return;
  }
  const int64_t pitch = CW + 1;
  const int s = sat[(int64_t)y1 * pitch + x1] - sat[(int64_t)y0 * pitch + x1] - sat[(int64_t)y1 * pitch + x0] +
                sat[(int64_t)y0 * pitch + x0];
  valid[k] = (s == 0) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// K2  gather + normalise.  Pass A: 32 row-chunks per patch reduce (min, max) of both rasters; pass B: reduce the 32
//     partials, then write float2 {ortho, dem} = ((v - lo) / (hi - lo)) - 0.5  (IEEE fp32, process_full_tiles.py:307-310)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMinMaxSplit = 32;

__global__ void __launch_bounds__(256) patch_minmax_kernel(const float* __restrict__ img,
                                                           const float* __restrict__ dem, int CW,
                                                           const int32_t* __restrict__ xy, int I,
                                                           float* __restrict__ partial) {
  const int k = blockIdx.x, split = blockIdx.y;
  const int x0 = xy[2 * k], y0 = xy[2 * k + 1];
  float v[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
  if (x0 >= 0 && y0 >= 0) {
    const int rows_per = (I + kMinMaxSplit - 1) / kMinMaxSplit;
    const int r0 = split * rows_per, r1 = min(I, r0 + rows_per);
    if (((x0 | I | CW) & 3) == 0) {   // 128-bit path: patch rows are 16-byte aligned
      const int q = I >> 2;           // float4 per patch row
      const int count = (r1 - r0) * q;
      for (int e = threadIdx.x; e < count; e += blockDim.x) {
        const int r = r0 + e / q, c = (e - (e / q) * q) << 2;
        const int64_t o = (int64_t)(y0 + r) * CW + x0 + c;
        const float4 a = __ldg(reinterpret_cast<const float4*>(img + o));
        const float4 b = __ldg(reinterpret_cast<const float4*>(dem + o));
        v[0] = fminf(fminf(v[0], a.x), fminf(fminf(a.y, a.z), a.w));
        v[1] = fmaxf(fmaxf(v[1], a.x), fmaxf(fmaxf(a.y, a.z), a.w));
        v[2] = fminf(fminf(v[2], b.x), fminf(fminf(b.y, b.z), b.w));
        v[3] = fmaxf(fmaxf(v[3], b.x), fmaxf(fmaxf(b.y, b.z), b.w));
      }
    } else {
      const int count = (r1 - r0) * I;
      for (int e = threadIdx.x; e < count; e += blockDim.x) {
        const int r = r0 + e / I, c = e % I;
        const int64_t o = (int64_t)(y0 + r) * CW + x0 + c;
        const float a = __ldg(img + o), b = __ldg(dem + o);
        v[0] = fminf(v[0], a);
        v[1] = fmaxf(v[1], a);
        v[2] = fminf(v[2], b);
        v[3] = fmaxf(v[3], b);
      }
    }
  }
  __shared__ float red[8][4];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v[0] = fminf(v[0], __shfl_xor_sync(0xffffffffu, v[0], d));
    v[1] = fmaxf(v[1], __shfl_xor_sync(0xffffffffu, v[1], d));
    v[2] = fminf(v[2], __shfl_xor_sync(0xffffffffu, v[2], d));
    v[3] = fmaxf(v[3], __shfl_xor_sync(0xffffffffu, v[3], d));
  }
  if ((threadIdx.x & 31) == 0)
    for (int j = 0; j < 4; ++j) red[threadIdx.x >> 5][j] = v[j];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      v[0] = fminf(v[0], red[w][0]);
      v[1] = fmaxf(v[1], red[w][1]);
      v[2] = fminf(v[2], red[w][2]);
      v[3] = fmaxf(v[3], red[w][3]);
    }
    float* p = partial + ((int64_t)k * kMinMaxSplit + split) * 4;
    p[0] = v[0];
    p[1] = v[1];
    p[2] = v[2];
    p[3] = v[3];
  }
}

__global__ void __launch_bounds__(256) patch_normalize_kernel(const float* __restrict__ img,
                                                              const float* __restrict__ dem, int CW,
                                                              const int32_t* __restrict__ xy, int I,
                                                              const float* __restrict__ partial,
                                                              float* __restrict__ out, float* __restrict__ minmax) {
  const int k = blockIdx.x;
  const int x0 = xy[2 * k], y0 = xy[2 * k + 1];
  float2* o = reinterpret_cast<float2*>(out) + (int64_t)k * I * I;
  const int chunk = (I * I + gridDim.y - 1) / gridDim.y;
  const int e0 = blockIdx.y * chunk, e1 = min(I * I, e0 + chunk);
  const bool vec = (((x0 | I | CW | chunk) & 3) == 0);
  if (x0 < 0 || y0 < 0) {  // padding slot (process_full_tiles.py:472): zeros
    if (((I | chunk) & 3) == 0) {
      float4* o4 = reinterpret_cast<float4*>(o);
      for (int e = (e0 >> 1) + threadIdx.x; e < (e1 >> 1); e += blockDim.x) o4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) o[e] = make_float2(0.f, 0.f);
    }
    if (blockIdx.y == 0 && threadIdx.x < 4) minmax[4 * k + threadIdx.x] = 0.f;
    return;
  }
  float lo_i = INFINITY, hi_i = -INFINITY, lo_d = INFINITY, hi_d = -INFINITY;
  const float* p = partial + (int64_t)k * kMinMaxSplit * 4;
#pragma unroll 4
  for (int s = 0; s < kMinMaxSplit; ++s) {
    lo_i = fminf(lo_i, p[4 * s + 0]);
    hi_i = fmaxf(hi_i, p[4 * s + 1]);
    lo_d = fminf(lo_d, p[4 * s + 2]);
    hi_d = fmaxf(hi_d, p[4 * s + 3]);
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    minmax[4 * k + 0] = lo_i;
    minmax[4 * k + 1] = hi_i;
    minmax[4 * k + 2] = lo_d;
    minmax[4 * k + 3] = hi_d;
  }
  const float range_i = __fsub_rn(hi_i, lo_i), range_d = __fsub_rn(hi_d, lo_d);
  if (vec) {   // 4 pixels per thread: two 128-bit loads, two 128-bit stores of interleaved {ortho, dem} pairs
    const int q = I >> 2;
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int e = (e0 >> 2) + threadIdx.x; e < (e1 >> 2); e += blockDim.x) {
      const int r = e / q, c = (e - r * q) << 2;
      const int64_t src = (int64_t)(y0 + r) * CW + x0 + c;
      const float4 a = __ldg(reinterpret_cast<const float4*>(img + src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(dem + src));
      float4 u, w;
      u.x = __fsub_rn(__fdiv_rn(__fsub_rn(a.x, lo_i), range_i), 0.5f);
      u.y = __fsub_rn(__fdiv_rn(__fsub_rn(b.x, lo_d), range_d), 0.5f);
      u.z = __fsub_rn(__fdiv_rn(__fsub_rn(a.y, lo_i), range_i), 0.5f);
      u.w = __fsub_rn(__fdiv_rn(__fsub_rn(b.y, lo_d), range_d), 0.5f);
      w.x = __fsub_rn(__fdiv_rn(__fsub_rn(a.z, lo_i), range_i), 0.5f);
      w.y = __fsub_rn(__fdiv_rn(__fsub_rn(b.z, lo_d), range_d), 0.5f);
      w.z = __fsub_rn(__fdiv_rn(__fsub_rn(a.w, lo_i), range_i), 0.5f);
      w.w = __fsub_rn(__fdiv_rn(__fsub_rn(b.w, lo_d), range_d), 0.5f);
      o4[2 * (int64_t)e] = u;
      o4[2 * (int64_t)e + 1] = w;
    }
    return;
  }
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int r = e / I, c = e % I;
    const int64_t src = (int64_t)(y0 + r) * CW + x0 + c;
    const float a = __ldg(img + src), b = __ldg(dem + src);
    float2 v;
    v.x = __fsub_rn(__fdiv_rn(__fsub_rn(a, lo_i), range_i), 0.5f);
    v.y = __fsub_rn(__fdiv_rn(__fsub_rn(b, lo_d), range_d), 0.5f);
    o[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K9  rebuildTile in gather form (+ finalize + paste).  One thread per output pixel of the T x T tile centre.
// ------------------------------------------------------------------------------------------------------------------
struct BlendState {
  float wsum, mean, s;
};

__device__ __forceinline__ void blend_update(BlendState& st, double w, bool is_f64, float d32, double d64) {
  // w_sum[...] += w                                   (:398)  f32 <- f64(f32) + f64
  st.wsum = __double2float_rn(__dadd_rn((double)st.wsum, w));
  // mean = mean_old + (w / w_sum) * (d - mean_old)    (:401)
  const double delta_old = is_f64 ? __dsub_rn(d64, (double)st.mean) : (double)__fsub_rn(d32, st.mean);
  const double ratio = __ddiv_rn(w, (double)st.wsum);
  st.mean = __double2float_rn(__dadd_rn((double)st.mean, __dmul_rn(ratio, delta_old)));
  // S += w * (d - mean_new) * (d - mean_new)          (:402; `mean_old` there is a view that already holds mean_new)
  const double delta_new = is_f64 ? __dsub_rn(d64, (double)st.mean) : (double)__fsub_rn(d32, st.mean);
  st.s = __double2float_rn(__dadd_rn((double)st.s, __dmul_rn(__dmul_rn(w, delta_new), delta_new)));
}

// One patch's contribution to a pixel, split into the loads (issued one contribution ahead of their use: the kernel is
// bound by the latency of these dependent loads and of the float64 chain, see profiles/r01b_ncu_tiling_summary.md)
// and the arithmetic.
struct BlendOperand {
  double w, v64;
  float v32, lo, hi;
  bool is_f64;
};

__device__ __forceinline__ BlendOperand blend_fetch(const void* const* __restrict__ ptrs,
                                                    const uint8_t* __restrict__ f64flags,
                                                    const float* __restrict__ lohi, const double* __restrict__ wtab,
                                                    int k, int ry, int rx, int I, int p) {
  // (ry, rx) = position inside patch k, already known to be in [p, I - p)
  BlendOperand o;
  const float2 lh = __ldg(reinterpret_cast<const float2*>(lohi) + k);
  o.lo = lh.x;
  o.hi = lh.y;
  o.w = __ldg(wtab + (int64_t)(ry - p) * (I - 2 * p) + (rx - p));
  o.is_f64 = f64flags != nullptr && f64flags[k] != 0;
  o.v32 = 0.f;
  o.v64 = 0.0;
  if (o.is_f64) o.v64 = reinterpret_cast<const double*>(ptrs[k])[(int64_t)ry * I + rx];
  else o.v32 = reinterpret_cast<const float*>(ptrs[k])[(int64_t)ry * I + rx];
  return o;
}

__device__ __forceinline__ void blend_apply(BlendState& st, const BlendOperand& o, int add_half) {
  const float range = __fsub_rn(o.hi, o.lo);
  float d32 = 0.f;
  double d64 = 0.0;
  if (o.is_f64) {
    double v = o.v64;
    if (add_half) v = __dadd_rn(v, 0.5);
    d64 = __dadd_rn(__dmul_rn(v, (double)range), (double)o.lo);  // f64 array * f32 scalar + f32 scalar (:396)
  } else {
    float v = o.v32;
    if (add_half) v = __fadd_rn(v, 0.5f);                        // processBatch :340
    d32 = __fadd_rn(__fmul_rn(v, range), o.lo);                  // :396
  }
  blend_update(st, o.w, o.is_f64, d32, d64);
}

__device__ __forceinline__ void blend_contribution(BlendState& st, const void* const* __restrict__ ptrs,
                                                   const uint8_t* __restrict__ f64flags,
                                                   const float* __restrict__ lohi, const double* __restrict__ wtab,
                                                   int k, int ry, int rx, int I, int p, int add_half) {
  blend_apply(st, blend_fetch(ptrs, f64flags, lohi, wtab, k, ry, rx, I, p), add_half);
}

__global__ void __launch_bounds__(256) blend_tile_kernel(const void* const* __restrict__ ptrs,
                                                         const uint8_t* __restrict__ f64flags,
                                                         const float* __restrict__ lohi,
                                                         const int32_t* __restrict__ pxy, int n,
                                                         const int32_t* __restrict__ lattice, int G,
                                                         const double* __restrict__ wtab, int I, int S, int T,
                                                         int add_half, float nv, float* __restrict__ mean_out,
                                                         float* __restrict__ std_out, uint8_t* __restrict__ good_out,
                                                         int64_t pitch, int rows, int cols) {
  // 1-D grid: block = (row, column block); gridDim.y would cap the rows at 65535 (a 70000-row band in dedup mode)
  const int col_blocks = (cols + blockDim.x - 1) / blockDim.x;
  const int row = blockIdx.x / col_blocks;
  const int col = (blockIdx.x - row * col_blocks) * blockDim.x + threadIdx.x;
  if (col >= cols || row >= rows) return;
  const int off = I - S, p = I / 16;
  const int Y = row + off, X = col + off;  // accumulator coordinates (:386, :404)
  BlendState st = {0.f, 0.f, 0.f};
  if (lattice != nullptr) {
    // patches covering Y: ky*S + p <= Y < ky*S + I - p
    int gy0 = (Y - (I - p) + S) / S;  // ceil((Y - (I-p) + 1) / S) for non-negative numerators
    if (Y - (I - p) + 1 <= 0) gy0 = 0;
    int gy1 = (Y - p) / S;
    if (Y - p < 0) gy1 = -1;
    int gx0 = (X - (I - p) + S) / S;
    if (X - (I - p) + 1 <= 0) gx0 = 0;
    int gx1 = (X - p) / S;
    if (X - p < 0) gx1 = -1;
    gy1 = min(gy1, G - 1);
    gx1 = min(gx1, G - 1);
    // lattice cells in the reference's order (gy outer, gx inner); the operands of the NEXT contributing cell are
    // fetched before the float64 chain of the current one runs
    int gy = gy0, gx = gx0 - 1;
    auto next_cell = [&]() -> int {
      for (;;) {
        if (++gx > gx1) {
          gx = gx0;
          if (++gy > gy1) return -1;
        }
        const int k = __ldg(lattice + gy * G + gx);
        if (k >= 0) return k;
      }
    };
    if (gy0 <= gy1 && gx0 <= gx1) {
      int k = next_cell();
      BlendOperand nxt;
      if (k >= 0) nxt = blend_fetch(ptrs, f64flags, lohi, wtab, k, Y - gy * S, X - gx * S, I, p);
      while (k >= 0) {
        const BlendOperand cur = nxt;
        k = next_cell();
        if (k >= 0) nxt = blend_fetch(ptrs, f64flags, lohi, wtab, k, Y - gy * S, X - gx * S, I, p);
        blend_apply(st, cur, add_half);
      }
    }
  } else {
    for (int k = 0; k < n; ++k) {
      const int ry = Y - pxy[2 * k + 1], rx = X - pxy[2 * k];
      if (ry < p || ry >= I - p || rx < p || rx >= I - p) continue;
      blend_contribution(st, ptrs, f64flags, lohi, wtab, k, ry, rx, I, p, add_half);
    }
  }
  const bool good = st.wsum > 0.f;                                    // :409
  const float sd = __fsqrt_rn(__fdiv_rn(st.s, st.wsum));              // :411
  const int64_t o = (int64_t)row * pitch + col;
  mean_out[o] = good ? st.mean : nv;                                  // :412
  std_out[o] = good ? sd : nv;                                        // :413
  good_out[o] = good ? 1 : 0;
}


// ------------------------------------------------------------------------------------------------------------------
// K10  dedup mode (mode B): the same update as K9, but the float32 state (w_sum, mean, S) lives in
//      canvas-shaped accumulators between calls, exactly like the reference's numpy arrays live between two iterations
//      of rebuildTile's loop (:395-402).  One call adds the patches [k0, k0 + n) of a band's lattice; calls are issued
//      in lattice order, so every pixel sees its patches in the reference's order (y outer, x inner).
//      One thread per accumulator pixel of the rows the call touches.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_accumulate_kernel(
    const float* __restrict__ pred, const float* __restrict__ lohi, const int32_t* __restrict__ lattice, int GX,
    int gy_lo, int gy_hi, int lat_y0, int k0, int n, const double* __restrict__ wtab, int I, int S, int add_half,
    float* __restrict__ wsum, float* __restrict__ mean, float* __restrict__ sacc, int64_t pitch, int acc_y0, int y_lo,
    int cols) {
  const int col_blocks = (cols + blockDim.x - 1) / blockDim.x;
  const int brow = blockIdx.x / col_blocks;
  const int X = (blockIdx.x - brow * col_blocks) * blockDim.x + threadIdx.x;
  if (X >= cols) return;
  const int Y = y_lo + brow;  // canvas row
  const int p = I / 16;
  const int ry = Y - lat_y0;        // row relative to the band's first lattice row
  // lattice rows covering Y: gy*S + p <= ry < gy*S + I - p
  int gy0 = (ry - (I - p) + 1 <= 0) ? 0 : (ry - (I - p) + S) / S;
  int gy1 = (ry - p < 0) ? -1 : (ry - p) / S;
  gy0 = max(gy0, gy_lo);
  gy1 = min(gy1, gy_hi);
  int gx0 = (X - (I - p) + 1 <= 0) ? 0 : (X - (I - p) + S) / S;
  int gx1 = (X - p < 0) ? -1 : (X - p) / S;
  gx1 = min(gx1, GX - 1);
  if (gy1 < gy0 || gx1 < gx0) return;
  const int64_t a = (int64_t)(Y - acc_y0) * pitch + X;
  BlendState st;
  bool touched = false;
  const int64_t II = (int64_t)I * I;
  const int wp = I - 2 * p;
  int gy = gy0, gx = gx0 - 1;
  auto next_cell = [&]() -> int {   // next lattice cell, in visit order, that holds a patch of this call
    for (;;) {
      if (++gx > gx1) {
        gx = gx0;
        if (++gy > gy1) return -1;
      }
      const int k = __ldg(lattice + (int64_t)gy * GX + gx) - k0;
      if (k >= 0 && k < n) return k;
    }
  };
  auto fetch = [&](int k) -> BlendOperand {
    const int py = ry - gy * S, px = X - gx * S;
    BlendOperand o;
    const float2 lh = __ldg(reinterpret_cast<const float2*>(lohi) + k);
    o.lo = lh.x;
    o.hi = lh.y;
    o.w = __ldg(wtab + (int64_t)(py - p) * wp + (px - p));
    o.v32 = __ldg(pred + k * II + (int64_t)py * I + px);
    o.v64 = 0.0;
    o.is_f64 = false;
    return o;
  };
  int k = next_cell();
  BlendOperand nxt;
  if (k >= 0) {
    nxt = fetch(k);
    st.wsum = wsum[a];
    st.mean = mean[a];
    st.s = sacc[a];
    touched = true;
  }
  while (k >= 0) {
    const BlendOperand cur = nxt;
    k = next_cell();
    if (k >= 0) nxt = fetch(k);
    blend_apply(st, cur, add_half);
  }
  if (touched) {
    wsum[a] = st.wsum;
    mean[a] = st.mean;
    sacc[a] = st.s;
  }
}

// rebuildTile's tail (:404-413) over a window of the accumulators: good = w_sum > 0, std = sqrt(S / w_sum) in float32,
// mean / std <- no_value where not good.  One thread per 4 adjacent pixels of a row: three 16-byte loads, two 16-byte and
// one 4-byte store when the window is 16-byte aligned (it is whenever off, the canvas width and W are multiples of 4).
__device__ __forceinline__ void finalize_pixel(float w, float m, float s, float nv, float& mo, float& so, uint8_t& go) {
  const bool good = w > 0.f;                                          // :409
  const float sd = __fsqrt_rn(__fdiv_rn(s, w));                       // :411
  mo = good ? m : nv;                                                 // :412
  so = good ? sd : nv;                                                // :413
  go = good ? 1 : 0;
}

__global__ void __launch_bounds__(256) blend_finalize_kernel(const float* __restrict__ wsum,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ sacc, int64_t acc_pitch,
                                                             int rows, int cols, float nv, float* __restrict__ mean_out,
                                                             float* __restrict__ std_out, uint8_t* __restrict__ good_out,
                                                             int64_t out_pitch, int vec_ok) {
  const int col_blocks = ((cols + 3) / 4 + blockDim.x - 1) / blockDim.x;
  const int r = blockIdx.x / col_blocks;
  const int c0 = ((blockIdx.x - r * col_blocks) * blockDim.x + threadIdx.x) * 4;
  if (c0 >= cols || r >= rows) return;
  const int64_t a = (int64_t)r * acc_pitch + c0, o = (int64_t)r * out_pitch + c0;
  if (vec_ok && c0 + 3 < cols) {
    const float4 w = __ldcs(reinterpret_cast<const float4*>(wsum + a));
    const float4 m = __ldcs(reinterpret_cast<const float4*>(mean + a));
    const float4 s = __ldcs(reinterpret_cast<const float4*>(sacc + a));
    float4 mo, so;
    uchar4 go;
    finalize_pixel(w.x, m.x, s.x, nv, mo.x, so.x, go.x);
    finalize_pixel(w.y, m.y, s.y, nv, mo.y, so.y, go.y);
    finalize_pixel(w.z, m.z, s.z, nv, mo.z, so.z, go.z);
    finalize_pixel(w.w, m.w, s.w, nv, mo.w, so.w, go.w);
    __stcs(reinterpret_cast<float4*>(mean_out + o), mo);
    __stcs(reinterpret_cast<float4*>(std_out + o), so);
    *reinterpret_cast<uchar4*>(good_out + o) = go;
    return;
  }
  for (int j = 0; j < 4 && c0 + j < cols; ++j) {
    float mo, so;
    uint8_t go;
    finalize_pixel(__ldg(wsum + a + j), __ldg(mean + a + j), __ldg(sacc + a + j), nv, mo, so, go);
    mean_out[o + j] = mo;
    std_out[o + j] = so;
    good_out[o + j] = go;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K9f / K10f  "fast" blend (DSRConfig(blend="fast")): the same weighted update (:395-402) in float32 -- no float64
// chain, no double-precision division -- so the kernels are bound by the prediction reads, not by the XU pipe.  One
// thread owns FOUR adjacent pixels of a row: 128-bit loads of the predictions and of the (float32) weight table, 128-bit
// stores.  Which patches reach which pixel (placement), their order and `good` are exactly those of the bit-exact
// kernels; the values agree with them to float32 rounding of the update (~1e-6 relative).  Needs I, S, I/16 and the
// window origin to be multiples of 4 (so the four pixels share their set of patches and the loads are aligned).
// ------------------------------------------------------------------------------------------------------------------
struct Blend4 {
  float w[4], m[4], s[4];
};

__device__ __forceinline__ void blend4_update(Blend4& st, const float4 v, const float4 wt, float lo, float range,
                                              float half) {
  const float vv[4] = {v.x, v.y, v.z, v.w}, ww[4] = {wt.x, wt.y, wt.z, wt.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float d = fmaf(vv[j] + half, range, lo);       // :340, :396
    st.w[j] += ww[j];                                    // :398
    const float delta = d - st.m[j];
    st.m[j] = fmaf(__fdividef(ww[j], st.w[j]), delta, st.m[j]);   // :401
    const float dn = d - st.m[j];                        // :402 (mean_old aliases the updated mean)
    st.s[j] = fmaf(ww[j] * dn, dn, st.s[j]);
  }
}

// The weight of patch pixel (ry, rx) is rebuilt from the SEPARABLE part of makeGaussianKernel (:347-361):
// w = e[ry] * e[rx] * c1 + c0 with e[i] = exp(-x_i^2 / 2s^2), c1 = A / (kmax - kmin), c0 = 1e-7 - kmin / (kmax - kmin), so
// the 800 KB 2-D table (as many L2 bytes per contribution as the prediction itself) is replaced by a 1.8 KB 1-D table
// that lives in L1; the result differs from the float32 cast of the float64 table by ~1e-7 relative.
__global__ void __launch_bounds__(256) blend_tile_fast_kernel(const float* __restrict__ pred,
                                                              const float* __restrict__ lohi,
                                                              const int32_t* __restrict__ lattice, int G,
                                                              const float* __restrict__ w1d, float c1, float c0, int I,
                                                              int S, int add_half, float nv, float* __restrict__ mean_out,
                                                              float* __restrict__ std_out, uint8_t* __restrict__ good_out,
                                                              int64_t pitch, int rows, int cols) {
  const int quads = (cols + 3) >> 2;
  const int qblocks = (quads + blockDim.x - 1) / blockDim.x;
  const int row = blockIdx.x / qblocks;
  const int col = ((blockIdx.x - row * qblocks) * blockDim.x + threadIdx.x) * 4;
  if (col >= cols || row >= rows) return;
  const int off = I - S, p = I / 16;
  const int Y = row + off, X = col + off;
  const int64_t II = (int64_t)I * I;
  const float half = add_half ? 0.5f : 0.f;
  int gy0 = (Y - (I - p) + 1 <= 0) ? 0 : (Y - (I - p) + S) / S;
  int gy1 = (Y - p < 0) ? -1 : min((Y - p) / S, G - 1);
  int gx0 = (X - (I - p) + 1 <= 0) ? 0 : (X - (I - p) + S) / S;
  int gx1 = (X - p < 0) ? -1 : min((X - p) / S, G - 1);
  Blend4 st;
#pragma unroll
  for (int j = 0; j < 4; ++j) st.w[j] = st.m[j] = st.s[j] = 0.f;
  // four lattice cells of a row at a time: their lattice entries, then their (independent) prediction / weight / range
  // loads are all in flight before the dependent update chain of the first one starts
  for (int gy = gy0; gy <= gy1; ++gy) {
    const int ry = Y - gy * S;
    const float ey = __ldg(w1d + (ry - p)) * c1;
    for (int gxb = gx0; gxb <= gx1; gxb += 4) {
      int k[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) k[u] = (gxb + u <= gx1) ? __ldg(lattice + gy * G + gxb + u) : -1;
      float4 v[4], wt[4];
      float2 lh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k[u] >= 0) {
          const int rx = X - (gxb + u) * S;
          v[u] = __ldcs(reinterpret_cast<const float4*>(pred + k[u] * II + (int64_t)ry * I + rx));   // read once
          const float4 ex = __ldg(reinterpret_cast<const float4*>(w1d + (rx - p)));
          wt[u] = make_float4(fmaf(ex.x, ey, c0), fmaf(ex.y, ey, c0), fmaf(ex.z, ey, c0), fmaf(ex.w, ey, c0));
          lh[u] = __ldg(reinterpret_cast<const float2*>(lohi) + k[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k[u] >= 0) blend4_update(st, v[u], wt[u], lh[u].x, lh[u].y - lh[u].x, half);
    }
  }
  float mo[4], so[4];
  uint8_t go[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) finalize_pixel(st.w[j], st.m[j], st.s[j], nv, mo[j], so[j], go[j]);
  const int64_t o = (int64_t)row * pitch + col;
  if (col + 3 < cols && (pitch & 3) == 0) {
    __stcs(reinterpret_cast<float4*>(mean_out + o), make_float4(mo[0], mo[1], mo[2], mo[3]));
    __stcs(reinterpret_cast<float4*>(std_out + o), make_float4(so[0], so[1], so[2], so[3]));
    *reinterpret_cast<uchar4*>(good_out + o) = make_uchar4(go[0], go[1], go[2], go[3]);
  } else {
    for (int j = 0; j < 4 && col + j < cols; ++j) {
      mean_out[o + j] = mo[j];
      std_out[o + j] = so[j];
      good_out[o + j] = go[j];
    }
  }
}

__global__ void __launch_bounds__(256) blend_accumulate_fast_kernel(
    const float* __restrict__ pred, const float* __restrict__ lohi, const int32_t* __restrict__ lattice, int GX,
    int gy_lo, int gy_hi, int lat_y0, int k0, int n, const float* __restrict__ wtab, int I, int S, int add_half,
    float* __restrict__ wsum, float* __restrict__ mean, float* __restrict__ sacc, int64_t pitch, int acc_y0, int y_lo,
    int cols) {
  const int quads = cols >> 2;   // cols % 4 == 0 (checked by the launcher)
  const int qblocks = (quads + blockDim.x - 1) / blockDim.x;
  const int brow = blockIdx.x / qblocks;
  const int X = ((blockIdx.x - brow * qblocks) * blockDim.x + threadIdx.x) * 4;
  if (X >= cols) return;
  const int Y = y_lo + brow;
  const int p = I / 16, wp = I - 2 * p;
  const int ry = Y - lat_y0;
  int gy0 = (ry - (I - p) + 1 <= 0) ? 0 : (ry - (I - p) + S) / S;
  int gy1 = (ry - p < 0) ? -1 : (ry - p) / S;
  gy0 = max(gy0, gy_lo);
  gy1 = min(gy1, gy_hi);
  int gx0 = (X - (I - p) + 1 <= 0) ? 0 : (X - (I - p) + S) / S;
  int gx1 = (X - p < 0) ? -1 : min((X - p) / S, GX - 1);
  if (gy1 < gy0 || gx1 < gx0) return;
  const int64_t a = (int64_t)(Y - acc_y0) * pitch + X;
  const int64_t II = (int64_t)I * I;
  const float half = add_half ? 0.5f : 0.f;
  Blend4 st;
  bool loaded = false;
  for (int gy = gy0; gy <= gy1; ++gy) {
    const int py = ry - gy * S;
    for (int gx = gx0; gx <= gx1; ++gx) {
      const int k = __ldg(lattice + (int64_t)gy * GX + gx) - k0;
      if (k < 0 || k >= n) continue;
      if (!loaded) {
        const float4 w4 = *reinterpret_cast<const float4*>(wsum + a);
        const float4 m4 = *reinterpret_cast<const float4*>(mean + a);
        const float4 s4 = *reinterpret_cast<const float4*>(sacc + a);
        st.w[0] = w4.x; st.w[1] = w4.y; st.w[2] = w4.z; st.w[3] = w4.w;
        st.m[0] = m4.x; st.m[1] = m4.y; st.m[2] = m4.z; st.m[3] = m4.w;
        st.s[0] = s4.x; st.s[1] = s4.y; st.s[2] = s4.z; st.s[3] = s4.w;
        loaded = true;
      }
      const int px = X - gx * S;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(pred + k * II + (int64_t)py * I + px));
      const float4 wt = __ldg(reinterpret_cast<const float4*>(wtab + (int64_t)(py - p) * wp + (px - p)));
      const float2 lh = __ldg(reinterpret_cast<const float2*>(lohi) + k);
      blend4_update(st, v, wt, lh.x, lh.y - lh.x, half);
    }
  }
  if (loaded) {
    *reinterpret_cast<float4*>(wsum + a) = make_float4(st.w[0], st.w[1], st.w[2], st.w[3]);
    *reinterpret_cast<float4*>(mean + a) = make_float4(st.m[0], st.m[1], st.m[2], st.m[3]);
    *reinterpret_cast<float4*>(sacc + a) = make_float4(st.s[0], st.s[1], st.s[2], st.s[3]);
  }
}

}  // namespace msr

using namespace msr;

extern "C" int msr_pad_inputs(const float* d_dem, const float* d_img, int H, int W, float* d_dem_canvas,
                              float* d_img_canvas, int CH, int CW, int off_y, int off_x, float no_value,
                              void* stream) {
  MSR_REQUIRE(d_dem && d_img && d_dem_canvas && d_img_canvas, "msr_pad_inputs: null pointer");
  MSR_REQUIRE(H > 0 && W > 0 && CH > 0 && CW > 0 && off_x >= 0 && CW >= W + off_x, "msr_pad_inputs: bad geometry");
  const int64_t quads = (int64_t)((CW + 3) / 4) * CH;
  ProfileScope prof(MSR_PROF_PAD, (cudaStream_t)stream, 8.0 * ((double)H * W + (double)CH * CW));
  const int blocks = (int)std::min<int64_t>((quads + 255) / 256, 148 * 16);
  pad_inputs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_dem, d_img, H, W, d_dem_canvas, d_img_canvas, CH, CW,
                                                              off_y, off_x, no_value);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

extern "C" int msr_validity_sat(const float* d_img_canvas, const float* d_dem_canvas, int CH, int CW, float no_value,
                                int32_t* d_sat, void* stream) {
  MSR_REQUIRE(d_img_canvas && d_dem_canvas && d_sat && CH > 0 && CW > 0, "msr_validity_sat: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  sat_rows_kernel
<<<CH, 256, 0, st>>>(d_img_canvas, d_dem_canvas, CH, CW, no_value, d_sat);
  MSR_LAUNCH_CHECK();
  sat_cols_kernel<<<ceil_div(CW + 1, 256), 256, 0, st>>>(CH, CW, d_sat);
  MSR_LAUNCH_CHECK();
  count_launch(2);
  return MSR_OK;
}

extern "C" int msr_patch_validity(const int32_t* d_sat, int CH, int CW, const int32_t* d_xy, int n, int I,
                                  uint8_t* d_valid, void* stream) {
  MSR_REQUIRE(d_sat && d_xy && d_valid && n >= 0 && I > 0, "msr_patch_validity: bad arguments");
  if (n == 0) return MSR_OK;
  patch_validity_kernel
<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_sat, CH, CW, d_xy, n, I, d_valid);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

extern "C" int msr_gather_normalize(const float* d_img_canvas, const float* d_dem_canvas, int CH, int CW,
                                    const int32_t* d_xy, int n, int I, float* d_out_nhwc, float* d_minmax,
                                    float* d_partial, void* stream) {
  MSR_REQUIRE(d_img_canvas && d_dem_canvas && d_xy && d_out_nhwc && d_minmax && d_partial,
              "msr_gather_normalize: null pointer");
  MSR_REQUIRE(n >= 0 && I > 0 && CH >= I && CW >= I, "msr_gather_normalize: bad geometry");
  if (n == 0) return MSR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(MSR_PROF_GATHER, st, (double)n * I * I * (8.0 + 8.0 + 8.0), 2);
  patch_minmax_kernel<<<dim3(n, kMinMaxSplit), 256, 0, st>>>(d_img_canvas, d_dem_canvas, CW, d_xy, I, d_partial);
  MSR_LAUNCH_CHECK();
  const int ysplit = std::max(1, std::min(64, (I * I) / 2048));
  patch_normalize_kernel<<<dim3(n, ysplit), 256, 0, st>>>(d_img_canvas, d_dem_canvas, CW, d_xy, I, d_partial,
                                                          d_out_nhwc, d_minmax);
  MSR_LAUNCH_CHECK();
  count_launch(2);
  return MSR_OK;
}

extern "C" int msr_blend_tile(const void* const* d_patch_ptr, const uint8_t* d_patch_f64, const float* d_patch_lohi,
                              const int32_t* d_patch_xy, int n, const int32_t* d_lattice, int G,
                              const double* d_weights, int I, int S, int T, int add_half, float no_value,
                              float* d_mean, float* d_std, uint8_t* d_good, int64_t pitch, int rows, int cols,
                              void* stream) {
  MSR_REQUIRE(d_weights && d_mean && d_std && d_good, "msr_blend_tile: null pointer");
  MSR_REQUIRE(n == 0 || (d_patch_ptr && d_patch_lohi && d_patch_xy), "msr_blend_tile: null patch tables");
  MSR_REQUIRE(I >= 16 && S > 0 && S <= I && T > 0, "msr_blend_tile: need I >= 16 (purge >= 1), 0 < S <= I");
  MSR_REQUIRE(rows >= 0 && cols >= 0 && rows <= T && cols <= T && pitch >= cols, "msr_blend_tile: bad output window");
  MSR_REQUIRE(d_lattice == nullptr || G > 0, "msr_blend_tile: lattice needs G > 0");
  if (rows == 0 || cols == 0) return MSR_OK;
  ProfileScope prof(MSR_PROF_BLEND, (cudaStream_t)stream,
                    (double)n * (I - 2 * (I / 16)) * (I - 2 * (I / 16)) * 4.0 + (double)rows * cols * 9.0);
  MSR_REQUIRE((int64_t)ceil_div(cols, 256) * rows < (1ll << 31), "msr_blend_tile: window too large for one launch");
  blend_tile_kernel<<<dim3((unsigned)(ceil_div(cols, 256) * rows)), 256, 0, (cudaStream_t)stream>>>(
      d_patch_ptr, d_patch_f64, d_patch_lohi, d_patch_xy, n, d_lattice, G, d_weights, I, S, T, add_half, no_value,
      d_mean, d_std, d_good, pitch, rows, cols);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

extern "C" int msr_blend_accumulate(const float* d_pred, const float* d_lohi, int k0, int n, const int32_t* d_lattice,
                                    int GY, int GX, int gy_lo, int gy_hi, int lattice_y0, const double* d_weights, int I,
                                    int S, int add_half, float* d_wsum, float* d_mean, float* d_s, int64_t pitch,
                                    int acc_y0, int acc_rows, int cols, int row_lo, int row_hi, void* stream) {
  MSR_REQUIRE(d_pred && d_lohi && d_lattice && d_weights && d_wsum && d_mean && d_s, "msr_blend_accumulate: null pointer");
  MSR_REQUIRE(I >= 16 && S > 0 && S <= I, "msr_blend_accumulate: need I >= 16 (purge >= 1), 0 < S <= I");
  MSR_REQUIRE(GY > 0 && GX > 0 && gy_lo >= 0 && gy_hi < GY && n >= 0 && k0 >= 0, "msr_blend_accumulate: bad lattice range");
  MSR_REQUIRE(acc_rows > 0 && cols > 0 && pitch >= cols, "msr_blend_accumulate: bad accumulator geometry");
  if (n == 0 || gy_hi < gy_lo) return MSR_OK;
  const int p = I / 16;
  // rows the patches of lattice rows [gy_lo, gy_hi] can touch, clipped to the caller's window and the accumulators
  int y0 = lattice_y0 + gy_lo * S + p, y1 = lattice_y0 + gy_hi * S + I - p;
  y0 = std::max(std::max(y0, row_lo), acc_y0);
  y1 = std::min(std::min(y1, row_hi), acc_y0 + acc_rows);
  if (y1 <= y0) return MSR_OK;
  const int wp = I - 2 * p;
  ProfileScope prof(MSR_PROF_BLEND, (cudaStream_t)stream, (double)n * wp * wp * (4.0 + 24.0));
  MSR_REQUIRE((int64_t)ceil_div(cols, 256) * (y1 - y0) < (1ll << 31), "msr_blend_accumulate: window too large");
  blend_accumulate_kernel<<<dim3((unsigned)(ceil_div(cols, 256) * (y1 - y0))), 256, 0, (cudaStream_t)stream>>>(
      d_pred, d_lohi, d_lattice, GX, gy_lo, gy_hi, lattice_y0, k0, n, d_weights, I, S, add_half, d_wsum, d_mean, d_s,
      pitch, acc_y0, y0, cols);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

extern "C" int msr_blend_finalize(const float* d_wsum, const float* d_mean_acc, const float* d_s, int64_t acc_pitch,
                                  int rows, int cols, float no_value, float* d_mean, float* d_std, uint8_t* d_good,
                                  int64_t out_pitch, void* stream) {
  MSR_REQUIRE(d_wsum && d_mean_acc && d_s && d_mean && d_std && d_good, "msr_blend_finalize: null pointer");
  MSR_REQUIRE(rows >= 0 && cols >= 0 && acc_pitch >= cols && out_pitch >= cols, "msr_blend_finalize: bad geometry");
  if (rows == 0 || cols == 0) return MSR_OK;
  const int64_t total = (int64_t)rows * cols;
  ProfileScope prof(MSR_PROF_BLEND, (cudaStream_t)stream, (double)total * 21.0);
  auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const int vec_ok = al(d_wsum, 16) && al(d_mean_acc, 16) && al(d_s, 16) && al(d_mean, 16) && al(d_std, 16) &&
                     al(d_good, 4) && (acc_pitch % 4 == 0) && (out_pitch % 4 == 0);
  MSR_REQUIRE((int64_t)ceil_div(ceil_div(cols, 4), 256) * rows < (1ll << 31), "msr_blend_finalize: window too large");
  blend_finalize_kernel<<<dim3((unsigned)(ceil_div(ceil_div(cols, 4), 256) * rows)), 256, 0, (cudaStream_t)stream>>>(
      d_wsum, d_mean_acc, d_s, acc_pitch, rows, cols, no_value, d_mean, d_std, d_good, out_pitch, vec_ok);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

static bool fast_blend_geometry_ok(int I, int S) { return I % 64 == 0 && S % 4 == 0 && S > 0 && S <= I; }

extern "C" int msr_blend_tile_fast(const float* d_pred, const float* d_lohi, int n, const int32_t* d_lattice, int G,
                                   const float* d_weights_1d, float c1, float c0, int I, int S, int T, int add_half,
                                   float no_value, float* d_mean, float* d_std, uint8_t* d_good, int64_t pitch, int rows,
                                   int cols, void* stream) {
  MSR_REQUIRE(d_weights_1d && d_mean && d_std && d_good && d_lattice && G > 0, "msr_blend_tile_fast: null pointer");
  MSR_REQUIRE(n == 0 || (d_pred && d_lohi), "msr_blend_tile_fast: null patch tables");
  MSR_REQUIRE(fast_blend_geometry_ok(I, S), "msr_blend_tile_fast: needs I % 64 == 0 and S % 4 == 0");
  MSR_REQUIRE(rows >= 0 && cols >= 0 && rows <= T && cols <= T && pitch >= cols, "msr_blend_tile_fast: bad output window");
  MSR_REQUIRE(((reinterpret_cast<uintptr_t>(d_pred) | reinterpret_cast<uintptr_t>(d_weights_1d) |
                reinterpret_cast<uintptr_t>(d_mean) | reinterpret_cast<uintptr_t>(d_std)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d_good) & 3) == 0 && (pitch & 3) == 0,
              "msr_blend_tile_fast: predictions, weights and the output window must be 16-byte aligned (W % 4 == 0)");
  if (rows == 0 || cols == 0) return MSR_OK;
  const int wp = I - 2 * (I / 16);
  // algorithmic bytes: the centre of every prediction read once + 9 bytes written per output pixel
  ProfileScope prof(MSR_PROF_BLEND, (cudaStream_t)stream, (double)n * wp * wp * 4.0 + (double)rows * cols * 9.0);
  const int quads = (cols + 3) / 4;
  MSR_REQUIRE((int64_t)ceil_div(quads, 256) * rows < (1ll << 31), "msr_blend_tile_fast: window too large");
  blend_tile_fast_kernel<<<dim3((unsigned)(ceil_div(quads, 256) * rows)), 256, 0, (cudaStream_t)stream>>>(
      d_pred, d_lohi, d_lattice, G, d_weights_1d, c1, c0, I, S, add_half, no_value, d_mean, d_std, d_good, pitch, rows,
      cols);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

extern "C" int msr_blend_accumulate_fast(const float* d_pred, const float* d_lohi, int k0, int n,
                                         const int32_t* d_lattice, int GY, int GX, int gy_lo, int gy_hi, int lattice_y0,
                                         const float* d_weights_f32, int I, int S, int add_half, float* d_wsum,
                                         float* d_mean, float* d_s, int64_t pitch, int acc_y0, int acc_rows, int cols,
                                         int row_lo, int row_hi, void* stream) {
  MSR_REQUIRE(d_pred && d_lohi && d_lattice && d_weights_f32 && d_wsum && d_mean && d_s,
              "msr_blend_accumulate_fast: null pointer");
  MSR_REQUIRE(fast_blend_geometry_ok(I, S), "msr_blend_accumulate_fast: needs I % 64 == 0 and S % 4 == 0");
  MSR_REQUIRE(GY > 0 && GX > 0 && gy_lo >= 0 && gy_hi < GY && n >= 0 && k0 >= 0, "msr_blend_accumulate_fast: bad lattice range");
  MSR_REQUIRE(acc_rows > 0 && cols > 0 && pitch >= cols && cols % 4 == 0 && pitch % 4 == 0,
              "msr_blend_accumulate_fast: accumulator width and pitch must be multiples of 4");
  MSR_REQUIRE(((reinterpret_cast<uintptr_t>(d_pred) | reinterpret_cast<uintptr_t>(d_weights_f32) |
                reinterpret_cast<uintptr_t>(d_wsum) | reinterpret_cast<uintptr_t>(d_mean) |
                reinterpret_cast<uintptr_t>(d_s)) & 15) == 0, "msr_blend_accumulate_fast: pointers must be 16-byte aligned");
  if (n == 0 || gy_hi < gy_lo) return MSR_OK;
  const int p = I / 16;
  int y0 = lattice_y0 + gy_lo * S + p, y1 = lattice_y0 + gy_hi * S + I - p;
  y0 = std::max(std::max(y0, row_lo), acc_y0);
  y1 = std::min(std::min(y1, row_hi), acc_y0 + acc_rows);
  if (y1 <= y0) return MSR_OK;
  const int wp = I - 2 * p;
  ProfileScope prof(MSR_PROF_BLEND, (cudaStream_t)stream, (double)n * wp * wp * (4.0 + 24.0));
  const int quads = cols / 4;
  MSR_REQUIRE((int64_t)ceil_div(quads, 256) * (y1 - y0) < (1ll << 31), "msr_blend_accumulate_fast: window too large");
  blend_accumulate_fast_kernel<<<dim3((unsigned)(ceil_div(quads, 256) * (y1 - y0))), 256, 0, (cudaStream_t)stream>>>(
      d_pred, d_lohi, d_lattice, GX, gy_lo, gy_hi, lattice_y0, k0, n, d_weights_f32, I, S, add_half, d_wsum, d_mean, d_s,
      pitch, acc_y0, y0, cols);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}
