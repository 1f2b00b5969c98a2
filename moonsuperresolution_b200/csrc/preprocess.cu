// preprocess (process_full_tiles.py:226-244): the DEM is box-filtered to 1/16 resolution in two 1/4 steps and brought
// back to full size by bicubic interpolation before it is tiled.  The two cv2.resize calls of that step as kernels:
//
//   resize_area4_kernel   cv2.resize(x, (0, 0), fx=0.25, fy=0.25, INTER_AREA)   (:232, :240)   read-bound
//   resize_cubic_kernel   cv2.resize(x, (W, H), INTER_CUBIC)                    (:241)         write-bound
//
// with the no_value <-> NaN bookkeeping of :230-233, :238-243 fused into their loads / stores.  The arithmetic follows
// OpenCV's own float32 code operation by operation (explicitly rounded adds / multiplies, no FMA contraction):
//   area:  full 4x4 windows  sum = (((0 + r0) + r1) + r2) + r3,  r = ((s0 + s1) + s2) + s3,  out = sum * (1/16);
//          windows cut by the raster edge: running sum in scan order / count; windows outside the raster: 0.
//   cubic: 4 taps, a = -0.75, weights and first-tap index per destination index precomputed by the host in float32;
//          horizontal pass accumulated tap 0 -> 3, vertical pass tap 3 -> 0, source indices clamped to the raster.
#include "common.cuh"

namespace msr {

__device__ __forceinline__ float nv_to_nan(float v, float nv) { return (v <= nv) ? __int_as_float(0x7fc00000) : v; }
__device__ __forceinline__ float nan_to_nv(float v, float nv) { return (v != v) ? nv : v; }

__global__ void __launch_bounds__(256) resize_area4_kernel(const float* __restrict__ src, int H, int W,
                                                           float* __restrict__ dst, int dh, int dw, float nv,
                                                           int vec_ok) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  if (dx >= dw || dy >= dh) return;
  const int sy0 = dy * 4, sx0 = dx * 4;
  float out;
  if (sy0 + 4 <= H && sx0 + 4 <= W) {
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* p = src + (int64_t)(sy0 + r) * W + sx0;
      float4 v;
      if (vec_ok) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        v = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
      }
      const float row = __fadd_rn(__fadd_rn(__fadd_rn(nv_to_nan(v.x, nv), nv_to_nan(v.y, nv)), nv_to_nan(v.z, nv)),
                                  nv_to_nan(v.w, nv));
      sum = __fadd_rn(sum, row);
    }
    out = __fmul_rn(sum, 0.0625f);
  } else if (sy0 >= H || sx0 >= W) {
    out = 0.f;
  } else {
    float sum = 0.f;
    int count = 0;
    for (int r = 0; r < 4 && sy0 + r < H; ++r)
      for (int c = 0; c < 4 && sx0 + c < W; ++c) {
        sum = __fadd_rn(sum, nv_to_nan(__ldg(src + (int64_t)(sy0 + r) * W + sx0 + c), nv));
        ++count;
      }
    out = __fdiv_rn(sum, (float)count);
  }
  dst[(int64_t)dy * dw + dx] = nan_to_nv(out, nv);
}

// One thread owns 4 adjacent destination columns and walks kCubicRows (16) destination rows; the horizontally interpolated
// values of the 4 source rows under the current destination row are kept in registers and recomputed only when the
// first-tap row changes (every ~16 rows at the path's x16 upsampling), so a destination pixel costs 4 multiplies and
// 3 adds plus one 16-byte store.
constexpr int kCubicRows = 16;

__global__ void __launch_bounds__(128, 8) resize_cubic_kernel(const float* __restrict__ src, int h, int w,
                                                              float* __restrict__ dst, int H, int W,
                                                              const int32_t* __restrict__ xofs,
                                                              const float* __restrict__ xcoef,
                                                              const int32_t* __restrict__ yofs,
                                                              const float* __restrict__ ycoef, float nv) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= W) return;
  const int y0 = blockIdx.y * kCubicRows;
  const int y1 = min(H, y0 + kCubicRows);
  float t[4][4];  // [source row tap][column]: horizontally interpolated source rows under the current destination row
  int cur = INT_MIN;
  const bool vec_ok = ((W & 3) == 0) && (x0 + 3 < W);
  for (int y = y0; y < y1; ++y) {
    const int yo = __ldg(yofs + y);
    if (yo != cur) {  // the 4-row source window moved (every ~16 rows at x16): redo the horizontal pass
      cur = yo;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = min(x0 + j, W - 1);
        const int o = __ldg(xofs + x) - 1;
        const float4 a = __ldg(reinterpret_cast<const float4*>(xcoef) + x);
        const int i0 = min(max(o, 0), w - 1), i1 = min(max(o + 1, 0), w - 1);
        const int i2 = min(max(o + 2, 0), w - 1), i3 = min(max(o + 3, 0), w - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float* row = src + (int64_t)min(max(yo - 1 + r, 0), h - 1) * w;
          float v = __fmul_rn(nv_to_nan(__ldg(row + i0), nv), a.x);
          v = __fadd_rn(v, __fmul_rn(nv_to_nan(__ldg(row + i1), nv), a.y));
          v = __fadd_rn(v, __fmul_rn(nv_to_nan(__ldg(row + i2), nv), a.z));
          v = __fadd_rn(v, __fmul_rn(nv_to_nan(__ldg(row + i3), nv), a.w));
          t[r][j] = v;
        }
      }
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(ycoef) + y);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = __fmul_rn(t[3][j], b.w);
      v = __fadd_rn(v, __fmul_rn(t[2][j], b.z));
      v = __fadd_rn(v, __fmul_rn(t[1][j], b.y));
      v = __fadd_rn(v, __fmul_rn(t[0][j], b.x));
      o[j] = nan_to_nv(v, nv);
    }
    float* q = dst + (int64_t)y * W + x0;
    if (vec_ok) {
      __stcs(reinterpret_cast<float4*>(q), make_float4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (x0 + j < W) q[j] = o[j];
    }
  }
}

}  // namespace msr

using namespace msr;

extern "C" int msr_resize_area4(const float* d_src, int H, int W, float* d_dst, int dh, int dw, float no_value,
                                void* stream) {
  MSR_REQUIRE(d_src && d_dst, "msr_resize_area4: null pointer");
  MSR_REQUIRE(H > 0 && W > 0 && dh > 0 && dw > 0, "msr_resize_area4: empty raster");
  // cv2.resize derives dsize as cvRound(extent * 0.25) (round half to even); anything else is not this operation
  MSR_REQUIRE(dh == (int)nearbyint(H * 0.25) && dw == (int)nearbyint(W * 0.25),
              "msr_resize_area4: destination must be (round(H/4), round(W/4)), half to even");
  const int vec_ok = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 15) == 0);
  ProfileScope prof(MSR_PROF_PREPROCESS, (cudaStream_t)stream, 4.0 * ((double)H * W + (double)dh * dw));
  resize_area4_kernel<<<dim3(ceil_div(dw, 256), dh), 256, 0, (cudaStream_t)stream>>>(d_src, H, W, d_dst, dh, dw,
                                                                                     no_value, vec_ok);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

extern "C" int msr_resize_cubic(const float* d_src, int h, int w, float* d_dst, int H, int W, const int32_t* d_xofs,
                                const float* d_xcoef, const int32_t* d_yofs, const float* d_ycoef, float no_value,
                                void* stream) {
  MSR_REQUIRE(d_src && d_dst && d_xofs && d_xcoef && d_yofs && d_ycoef, "msr_resize_cubic: null pointer");
  MSR_REQUIRE(h > 0 && w > 0 && H > 0 && W > 0, "msr_resize_cubic: empty raster");
  MSR_REQUIRE((reinterpret_cast<uintptr_t>(d_ycoef) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_xcoef) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d_dst) & 15) == 0,
              "msr_resize_cubic: d_xcoef, d_ycoef and d_dst must be 16-byte aligned");
  ProfileScope prof(MSR_PROF_PREPROCESS, (cudaStream_t)stream, 4.0 * ((double)H * W + (double)h * w));
  resize_cubic_kernel<<<dim3(ceil_div(ceil_div(W, 4), 128), ceil_div(H, kCubicRows)), 128, 0, (cudaStream_t)stream>>>(
      d_src, h, w, d_dst, H, W, d_xofs, d_xcoef, d_yofs, d_ycoef, no_value);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}
