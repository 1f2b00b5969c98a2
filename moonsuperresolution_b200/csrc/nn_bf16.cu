// Bandwidth-bound helpers of the bf16 tensor-core mode: everything around the tcgen05 convolutions that is not a GEMM.
//
// Reference semantics: spade/models/spade.py:17-18 (mask resize + 2->128 conv input), blocks.py:53-65 (encoder block),
// networks.py:31-33,41 (dense layers), spade.py:21 (batch moments).
#include "nn.cuh"

namespace msr {

// 256-bit global store (sm_100): one full 32-byte sector per lane per instruction
__device__ __forceinline__ void st_row_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// im2col of the 2-channel source: one thread per output pixel writes one 128-byte row (18 bf16 taps + zeros), so the
// 2->128 (SPADE mask) and 2->64 (encoder block 1) 3x3 convolutions become K = 64 GEMMs on the tensor cores.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) source_patches_kernel(const float* __restrict__ src, int I,
                                                             __nv_bfloat16* __restrict__ out, int n, int r, int mode) {
  const int64_t total = (int64_t)n * r * r;
  const int f = I / r, half = f >> 1;
  const int lr = 31 - __clz(r);                  // r is a power of two (checked by the launcher)
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const int nn = (int)(m >> (2 * lr));
    const int rem = (int)(m & (((int64_t)1 << (2 * lr)) - 1));
    const int h = rem >> lr, x = rem & (r - 1);
    // split-bf16 operand: channels [0,18) = hi, [18,36) = hi again, [36,54) = lo (v = hi + lo to ~2^-17); the matching
    // weight rows are [w_hi | w_lo | w_hi], so the GEMM evaluates a_hi*w_hi + a_hi*w_lo + a_lo*w_hi ~ fp32 product
    __align__(16) __nv_bfloat16 row[64];
#pragma unroll
    for (int j = 54; j < 64; ++j) row[j] = __float2bfloat16_rn(0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float2 s = make_float2(0.f, 0.f);
        if (mode == 0) {  // resized-mask coordinates, SAME pad (1, 1); nearest resize with half-pixel centres (App. B.3)
          const int hh = h + ky - 1, xx = x + kx - 1;
          if (hh >= 0 && hh < r && xx >= 0 && xx < r)
            s = __ldg(reinterpret_cast<const float2*>(src + (((int64_t)nn * I + hh * f + half) * I + xx * f + half) * 2));
        } else {          // stride-2 taps on the full-resolution source, SAME pad (0, 1) (App. B.2)
          const int sy = 2 * h + ky, sx = 2 * x + kx;
          if (sy < I && sx < I) s = __ldg(reinterpret_cast<const float2*>(src + (((int64_t)nn * I + sy) * I + sx) * 2));
        }
        const int j = (ky * 3 + kx) * 2;
        const __nv_bfloat16 hx = __float2bfloat16_rn(s.x), hy = __float2bfloat16_rn(s.y);
        row[j] = hx;
        row[j + 1] = hy;
        row[18 + j] = hx;
        row[18 + j + 1] = hy;
        row[36 + j] = __float2bfloat16_rn(s.x - __bfloat162float(hx));
        row[36 + j + 1] = __float2bfloat16_rn(s.y - __bfloat162float(hy));
      }
    }
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
    for (int q = 0; q < 4; ++q) st_row_v8(out + m * 64 + q * 16, sp + q * 8);
  }
}

// The same rows, stored as whole cache lines: a thread's 128-byte row goes through shared memory (row pitch 144 bytes:
// eight lanes writing 16 bytes at consecutive rows hit distinct banks), then every warp instruction writes 4 complete
// rows (512 contiguous bytes).  One thread per pixel storing its own row issues 32-byte pieces into 32 different lines per
// instruction, and the L1 store path, not DRAM, bounds the kernel (2.8 TB/s).  Grid = ceil(total / 256) blocks exactly.
constexpr int kPatchPitch = 128 + 16;
__global__ void __launch_bounds__(256) source_patches_lines_kernel(const float* __restrict__ src, int I,
                                                                   __nv_bfloat16* __restrict__ out, int n, int r, int mode) {
  __shared__ __align__(16) uint8_t stage[256 * kPatchPitch];
  const int64_t total = (int64_t)n * r * r;
  const int f = I / r, half = f >> 1;
  const int lr = 31 - __clz(r);
  const int64_t m0 = (int64_t)blockIdx.x * 256;
  const int64_t m = m0 + threadIdx.x;
  if (m < total) {
    const int nn = (int)(m >> (2 * lr));
    const int rem = (int)(m & (((int64_t)1 << (2 * lr)) - 1));
    const int h = rem >> lr, x = rem & (r - 1);
    __align__(16) __nv_bfloat16 row[64];
#pragma unroll
    for (int j = 54; j < 64; ++j) row[j] = __float2bfloat16_rn(0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float2 s = make_float2(0.f, 0.f);
        if (mode == 0) {
          const int hh = h + ky - 1, xx = x + kx - 1;
          if (hh >= 0 && hh < r && xx >= 0 && xx < r)
            s = __ldg(reinterpret_cast<const float2*>(src + (((int64_t)nn * I + hh * f + half) * I + xx * f + half) * 2));
        } else {
          const int sy = 2 * h + ky, sx = 2 * x + kx;
          if (sy < I && sx < I) s = __ldg(reinterpret_cast<const float2*>(src + (((int64_t)nn * I + sy) * I + sx) * 2));
        }
        const int j = (ky * 3 + kx) * 2;
        const __nv_bfloat16 hx = __float2bfloat16_rn(s.x), hy = __float2bfloat16_rn(s.y);
        row[j] = hx;
        row[j + 1] = hy;
        row[18 + j] = hx;
        row[18 + j + 1] = hy;
        row[36 + j] = __float2bfloat16_rn(s.x - __bfloat162float(hx));
        row[36 + j + 1] = __float2bfloat16_rn(s.y - __bfloat162float(hy));
      }
    }
    const uint4* rp = reinterpret_cast<const uint4*>(row);
    uint4* sp = reinterpret_cast<uint4*>(stage + threadIdx.x * kPatchPitch);
#pragma unroll
    for (int q = 0; q < 8; ++q) sp[q] = rp[q];
  }
  __syncthreads();
  // 256 rows x 8 chunks of 16 bytes; consecutive threads take consecutive chunks of the block's contiguous 32 KB
  const int rows_here = (total - m0 < 256) ? (int)(total - m0) : 256;
  uint8_t* dst = reinterpret_cast<uint8_t*>(out + m0 * 64);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int e = it * 256 + threadIdx.x;       // 16-byte chunk index inside the block's output
    const int rr = e >> 3, cc = e & 7;
    if (rr < rows_here)
      *reinterpret_cast<uint4*>(dst + (int64_t)e * 16) = *reinterpret_cast<const uint4*>(stage + rr * kPatchPitch + cc * 16);
  }
}

// pix2pix block 1 (pix2pix.py:64-72): Conv2D(64, 4, strides=2, 'same') on the 2-channel source = taps (2y + ky - 1,
// 2x + kx - 1), ky, kx in 0..3.  Row = 32 tap-channel values as hi (channels 0..31) | lo (32..63); with weight rows
// [w | w] the GEMM evaluates (x_hi + x_lo) * w, i.e. the exact input against bf16 weights.
__global__ void __launch_bounds__(256) source_patches4_kernel(const float* __restrict__ src, int I,
                                                              __nv_bfloat16* __restrict__ out, int n, int r) {
  const int64_t total = (int64_t)n * r * r;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const int nn = (int)(m / ((int64_t)r * r));
    const int rem = (int)(m % ((int64_t)r * r));
    const int h = rem / r, x = rem % r;
    __align__(16) __nv_bfloat16 row[64];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        float2 s = make_float2(0.f, 0.f);
        const int sy = 2 * h + ky - 1, sx = 2 * x + kx - 1;
        if (sy >= 0 && sy < I && sx >= 0 && sx < I)
          s = __ldg(reinterpret_cast<const float2*>(src + (((int64_t)nn * I + sy) * I + sx) * 2));
        const int j = (ky * 4 + kx) * 2;
        const __nv_bfloat16 hx = __float2bfloat16_rn(s.x), hy = __float2bfloat16_rn(s.y);
        row[j] = hx;
        row[j + 1] = hy;
        row[32 + j] = __float2bfloat16_rn(s.x - __bfloat162float(hx));
        row[32 + j + 1] = __float2bfloat16_rn(s.y - __bfloat162float(hy));
      }
    }
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
    for (int q = 0; q < 4; ++q) st_row_v8(out + m * 64 + q * 16, sp + q * 8);
  }
}

int source_patches_bf16(const float* source, int I, __nv_bfloat16* out, int n, int r, int mode, cudaStream_t st) {
  MSR_REQUIRE(source && out && n > 0 && r > 0 && I % r == 0 && (r & (r - 1)) == 0,
              "source_patches: bad arguments (r must be a power of two dividing I)");
  MSR_REQUIRE(mode == 0 || ((mode == 1 || mode == 2) && r * 2 == I), "source_patches: modes 1, 2 need r = I / 2");
  const int64_t total = (int64_t)n * r * r;
  ProfileScope prof(MSR_PROF_MASK_CONV, st, (double)total * (128.0 + 8.0));
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  if (mode == 2) source_patches4_kernel<<<blocks, 256, 0, st>>>(source, I, out, n, r);
  else if (total >= 4096 && total / 256 < (1ll << 31))
    source_patches_lines_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(source, I, out, n, r, mode);
  else source_patches_kernel<<<blocks, 256, 0, st>>>(source, I, out, n, r, mode);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Dense layer with bf16 weights for small M (<= 16 rows per pass): weight-bandwidth bound.  Block = 128 threads, each
// owning 4 consecutive output columns; the K range is split across blockIdx.y (deterministic two-stage sum).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDM = 16;    // rows per pass
constexpr int kDKT = 32;   // k rows staged per iteration

__device__ __forceinline__ float4 load_w4(const __nv_bfloat16* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                     __uint_as_float(v.y & 0xffff0000u));
}
__device__ __forceinline__ float4 load_w4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <typename WT>
__global__ void __launch_bounds__(128) dense_partial_kernel(const float* __restrict__ x, const WT* __restrict__ w,
                                                            float* __restrict__ partial, int M, int K, int N,
                                                            int kchunk, int m0) {
  __shared__ float xs[kDM][kDKT + 4];
  const int n4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int split = blockIdx.y;
  const int k0 = split * kchunk, k1 = min(K, k0 + kchunk);
  const int mrows = min(kDM, M - m0);
  float acc[kDM][4];
#pragma unroll
  for (int i = 0; i < kDM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int kb = k0; kb < k1; kb += kDKT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kDM * kDKT; e += 128) {
      const int mi = e / kDKT, kk = e % kDKT;
      xs[mi][kk] = (mi < mrows && kb + kk < k1) ? __ldg(x + (int64_t)(m0 + mi) * K + kb + kk) : 0.f;
    }
    __syncthreads();
    if (n4 < N) {
#pragma unroll
      for (int kq = 0; kq < kDKT; kq += 8) {
        float4 wv[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          wv[kk] = (kb + kq + kk < k1) ? load_w4(w + (int64_t)(kb + kq + kk) * N + n4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
          for (int i = 0; i < kDM; ++i) {
            const float xv = xs[i][kq + kk];
            acc[i][0] = fmaf(xv, wv[kk].x, acc[i][0]);
            acc[i][1] = fmaf(xv, wv[kk].y, acc[i][1]);
            acc[i][2] = fmaf(xv, wv[kk].z, acc[i][2]);
            acc[i][3] = fmaf(xv, wv[kk].w, acc[i][3]);
          }
        }
      }
    }
  }
  if (n4 < N) {
#pragma unroll
    for (int i = 0; i < kDM; ++i)
      if (i < mrows)
        *reinterpret_cast<float4*>(partial + ((int64_t)split * M + m0 + i) * N + n4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

__global__ void dense_bf16w_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                          float* __restrict__ out, int M, int N, int ksplit) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * N) return;
  float s = 0.f;
  for (int k = 0; k < ksplit; ++k) s += partial[(int64_t)k * M * N + e];
  if (bias) s += bias[e % N];
  out[e] = s;
}

template <typename WT>
static int dense_small_m(const float* x, const WT* w, const float* bias, float* out, int M, int K, int N,
                         float* partial, int64_t partial_capacity, cudaStream_t st) {
  MSR_REQUIRE(x && w && out && partial && M > 0 && K > 0 && N > 0 && N % 4 == 0, "dense: bad arguments");
  ProfileScope prof(MSR_PROF_DENSE, st, 2.0 * M * (double)K * N, 2);
  const int gx = ceil_div(N, 512);
  int ksplit = std::max(1, std::min(ceil_div(K, kDKT), (2 * 148) / gx));
  while ((int64_t)ksplit * M * N > partial_capacity && ksplit > 1) --ksplit;
  MSR_REQUIRE((int64_t)ksplit * M * N <= partial_capacity, "dense: partial scratch too small");
  int kchunk = ceil_div(K, ksplit);
  kchunk = ceil_div(kchunk, kDKT) * kDKT;
  ksplit = ceil_div(K, kchunk);
  for (int m0 = 0; m0 < M; m0 += kDM) {
    dense_partial_kernel<WT><<<dim3(gx, ksplit), 128, 0, st>>>(x, w, partial, M, K, N, kchunk, m0);
    MSR_LAUNCH_CHECK();
    count_launch();
  }
  dense_bf16w_reduce_kernel<<<ceil_div((int64_t)M * N, 256), 256, 0, st>>>(partial, bias, out, M, N, ksplit);
  MSR_LAUNCH_CHECK();
  count_launch();
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Dense layer for 16 < M <= 128 rows (several batches per generator call): a split-K SGEMM whose 128 x 128 block tile
// covers ALL rows, so every weight is read from HBM exactly once.  256 threads, 8 x 8 outputs per thread.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGM = 128, kGN = 128, kGK = 16;

__global__ void __launch_bounds__(256) dense_gemm_partial_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ dst,
                                                                 int M, int K, int N, int kchunk, int direct) {
  __shared__ float xs[kGK][kGM + 4];
  __shared__ float ws[kGK][kGN + 4];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int n0 = blockIdx.x * kGN, split = blockIdx.y;
  const int k0 = split * kchunk, k1 = min(K, k0 + kchunk);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // loader roles: x tile = 128 rows x 16 k (thread -> row t >> 1, 8 consecutive k); w tile = 16 k x 128 cols
  const int xr = t >> 1, xk = (t & 1) * 8;
  const int wk = t >> 4, wc = (t & 15) * 8;
  for (int kb = k0; kb < k1; kb += kGK) {
    float xv[8], wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb + xk + j;
      xv[j] = (xr < M && k < k1) ? __ldg(x + (int64_t)xr * K + k) : 0.f;
    }
    {
      const int k = kb + wk;
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = (k < k1 && n0 + wc + j < N) ? __ldg(w + (int64_t)k * N + n0 + wc + j) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) xs[xk + j][xr] = xv[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) ws[wk][wc + j] = wv[j];
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&xs[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&xs[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&ws[kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&ws[kk][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= N) continue;
      if (direct) dst[(int64_t)m * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
      else dst[((int64_t)split * M + m) * N + n] = acc[i][j];
    }
  }
}

int dense_reduce_planes(const float* partial, const float* bias, float* out, int M, int N, int planes, cudaStream_t st) {
  MSR_REQUIRE(partial && out && M > 0 && N > 0 && planes > 0, "dense_reduce_planes: bad arguments");
  ProfileScope prof(MSR_PROF_DENSE, st, (double)M * N * 4.0 * (planes + 1));
  dense_bf16w_reduce_kernel<<<ceil_div((int64_t)M * N, 256), 256, 0, st>>>(partial, bias, out, M, N, planes);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

static int dense_gemm_m128(const float* x, const float* w, const float* bias, float* out, int M, int K, int N,
                           float* partial, int64_t partial_capacity, cudaStream_t st) {
  ProfileScope prof(MSR_PROF_DENSE, st, 2.0 * M * (double)K * N, 2);
  const int gx = ceil_div(N, kGN);
  int ksplit = std::max(1, std::min(ceil_div(K, 4 * kGK), (2 * 148) / gx));
  while ((int64_t)ksplit * M * N > partial_capacity && ksplit > 1) --ksplit;
  int kchunk = ceil_div(ceil_div(K, ksplit), kGK) * kGK;
  ksplit = ceil_div(K, kchunk);
  if (ksplit == 1) {
    dense_gemm_partial_kernel<<<dim3(gx, 1), 256, 0, st>>>(x, w, bias, out, M, K, N, kchunk, 1);
    MSR_LAUNCH_CHECK();
    count_launch();
    return MSR_OK;
  }
  MSR_REQUIRE((int64_t)ksplit * M * N <= partial_capacity, "dense: partial scratch too small");
  dense_gemm_partial_kernel<<<dim3(gx, ksplit), 256, 0, st>>>(x, w, nullptr, partial, M, K, N, kchunk, 0);
  MSR_LAUNCH_CHECK();
  dense_bf16w_reduce_kernel<<<ceil_div((int64_t)M * N, 256), 256, 0, st>>>(partial, bias, out, M, N, ksplit);
  MSR_LAUNCH_CHECK();
  count_launch(2);
  return MSR_OK;
}

int dense_f32w(const float* x, const float* w, const float* bias, float* out, int M, int K, int N, float* partial,
               int64_t partial_capacity, cudaStream_t st) {
  MSR_REQUIRE(x && w && out && partial && M > 0 && K > 0 && N > 0, "dense: bad arguments");
  if (M > 16) {   // several batches per call: all rows share one pass over the weights (chunks of 128 rows)
    for (int m0 = 0; m0 < M; m0 += kGM) {
      const int rows = std::min(kGM, M - m0);
      int rc = dense_gemm_m128(x + (int64_t)m0 * K, w, bias, out + (int64_t)m0 * N, rows, K, N, partial,
                               partial_capacity, st);
      if (rc) return rc;
    }
    return MSR_OK;
  }
  return dense_small_m<float>(x, w, bias, out, M, K, N, partial, partial_capacity, st);
}

// ------------------------------------------------------------------------------------------------------------------
// normalise + affine + activation with bf16 (and optionally fp32) output, 4 channels per thread
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) affine_act_bf16_kernel(const float* __restrict__ x, int ldx,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ yb, float* __restrict__ yf,
                                                              int64_t M, int C, int64_t rows_per_group, int act,
                                                              float slope, int split) {
  const int c4n = C / 4;
  const int64_t total = M * c4n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % c4n) * 4;
    const int64_t m = e / c4n;
    const float4 xv = *reinterpret_cast<const float4*>(x + m * ldx + c);
    float v[4] = {xv.x, xv.y, xv.z, xv.w};
    if (mean) {
      const int64_t g = m / rows_per_group;
      const float4 mu = *reinterpret_cast<const float4*>(mean + g * C + c);
      const float4 rs = *reinterpret_cast<const float4*>(rstd + g * C + c);
      v[0] = (v[0] - mu.x) * rs.x; v[1] = (v[1] - mu.y) * rs.y; v[2] = (v[2] - mu.z) * rs.z; v[3] = (v[3] - mu.w) * rs.w;
    }
    if (gamma) {
      const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
      v[0] *= ga.x; v[1] *= ga.y; v[2] *= ga.z; v[3] *= ga.w;
    }
    if (beta) {
      const float4 be = *reinterpret_cast<const float4*>(beta + c);
      v[0] += be.x; v[1] += be.y; v[2] += be.z; v[3] += be.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (act == ACT_RELU) v[j] = fmaxf(v[j], 0.f);
      else if (act == ACT_LRELU) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
    }
    if (yf) *reinterpret_cast<float4*>(yf + m * C + c) = make_float4(v[0], v[1], v[2], v[3]);
    if (yb) {
      __align__(8) __nv_bfloat16 o[4] = {__float2bfloat16_rn(v[0]), __float2bfloat16_rn(v[1]),
                                          __float2bfloat16_rn(v[2]), __float2bfloat16_rn(v[3])};
      if (!split) {
        *reinterpret_cast<uint2*>(yb + m * C + c) = *reinterpret_cast<const uint2*>(o);
      } else {  // hi | lo halves of a split-bf16 operand, row pitch 2C
        __align__(8) __nv_bfloat16 l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(o[j]));
        *reinterpret_cast<uint2*>(yb + m * 2 * C + c) = *reinterpret_cast<const uint2*>(o);
        *reinterpret_cast<uint2*>(yb + m * 2 * C + C + c) = *reinterpret_cast<const uint2*>(l);
      }
    }
  }
}

// Same operation, 8 channels per thread and 32-bit index arithmetic (C a power of two): two 128-bit loads, one 128-bit
// store per output half.  The generic kernel above spends its time in 64-bit divisions, not on the memory system.
__global__ void __launch_bounds__(256) affine_act_bf16_vec8_kernel(const float* __restrict__ x, int ldx,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   __nv_bfloat16* __restrict__ yb, float* __restrict__ yf,
                                                                   int M, int C, int rows_per_group, int act, float slope,
                                                                   int split) {
  const int c8n = C >> 3;                       // threads per row (power of two, <= 256)
  const int rpb = 256 / c8n;                    // rows per block
  const int c = (threadIdx.x & (c8n - 1)) * 8;
  const int rl = threadIdx.x / c8n;
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gamma ? __ldg(gamma + c + j) : 1.f;
    be[j] = beta ? __ldg(beta + c + j) : 0.f;
  }
  for (int m = blockIdx.x * rpb + rl; m < M; m += gridDim.x * rpb) {
    const float4 a0 = __ldcs(reinterpret_cast<const float4*>(x + (int64_t)m * ldx + c));
    const float4 a1 = __ldcs(reinterpret_cast<const float4*>(x + (int64_t)m * ldx + c + 4));
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (mean) {
      const int g = m / rows_per_group;
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean + (int64_t)g * C + c));
      const float4 m1 = __ldg(reinterpret_cast<const float4*>(mean + (int64_t)g * C + c + 4));
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(rstd + (int64_t)g * C + c));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(rstd + (int64_t)g * C + c + 4));
      const float mu[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      const float rs[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (v[j] - mu[j]) * rs[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (gamma) v[j] *= ga[j];          // same operation order as the generic kernel: (x - mu) * rstd, * gamma, + beta
      if (beta) v[j] += be[j];
      if (act == ACT_RELU) v[j] = fmaxf(v[j], 0.f);
      else if (act == ACT_LRELU) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
    }
    if (yf) {
      *reinterpret_cast<float4*>(yf + (int64_t)m * C + c) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(yf + (int64_t)m * C + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (yb) {
      __align__(16) __nv_bfloat16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(v[j]);
      if (!split) {
        *reinterpret_cast<uint4*>(yb + (int64_t)m * C + c) = *reinterpret_cast<const uint4*>(o);
      } else {
        __align__(16) __nv_bfloat16 l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(o[j]));
        if (split == 1) {
          *reinterpret_cast<uint4*>(yb + (int64_t)m * 2 * C + c) = *reinterpret_cast<const uint4*>(o);
          *reinterpret_cast<uint4*>(yb + (int64_t)m * 2 * C + C + c) = *reinterpret_cast<const uint4*>(l);
        } else {   // flattened image: hi block (K values) then lo block, K = rows_per_group * C
          const int b = m / rows_per_group, pp = m - b * rows_per_group;
          const int64_t K = (int64_t)rows_per_group * C;
          *reinterpret_cast<uint4*>(yb + (int64_t)b * 2 * K + (int64_t)pp * C + c) = *reinterpret_cast<const uint4*>(o);
          *reinterpret_cast<uint4*>(yb + (int64_t)b * 2 * K + K + (int64_t)pp * C + c) = *reinterpret_cast<const uint4*>(l);
        }
      }
    }
  }
}

int affine_act_bf16out(const float* x, int ldx, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, __nv_bfloat16* y_bf16, float* y_f32, int64_t M, int C, int64_t rows_per_group,
                       int act, float slope, int split, cudaStream_t st) {
  MSR_REQUIRE(x && (y_bf16 || y_f32) && M > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && rows_per_group > 0,
              "affine_act_bf16out: bad arguments");
  MSR_REQUIRE(split != 2 || (C % 8 == 0 && (C & (C - 1)) == 0 && C / 8 <= 256),
              "affine_act_bf16out: the flattened split layout needs a power-of-two channel count");
  ProfileScope prof(MSR_PROF_ELEMWISE, st,
                    (double)M * C * (4.0 + (y_bf16 ? (split ? 4.0 : 2.0) : 0.0) + (y_f32 ? 4.0 : 0.0)));
  if (C % 8 == 0 && (C & (C - 1)) == 0 && C / 8 <= 256 && ldx % 4 == 0 && M < (1ll << 31) && rows_per_group < (1ll << 31)) {
    const int c8n = C / 8, rpb = 256 / c8n;
    const int blocks = (int)std::min<int64_t>((M + rpb - 1) / rpb, 148 * 16);
    affine_act_bf16_vec8_kernel<<<blocks, 256, 0, st>>>(x, ldx, mean, rstd, gamma, beta, y_bf16, y_f32, (int)M, C,
                                                        (int)rows_per_group, act, slope, split);
    count_launch();
    MSR_LAUNCH_CHECK();
    return MSR_OK;
  }
  const int64_t total = M * (C / 4);
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  affine_act_bf16_kernel<<<blocks, 256, 0, st>>>(x, ldx, mean, rstd, gamma, beta, y_bf16, y_f32, M, C, rows_per_group,
                                                 act, slope, split);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// batch statistics from the (sum, sumsq) pairs emitted by the tensor-core epilogue (deterministic, fp64 second stage)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_pairs_partial_kernel(const float2* __restrict__ pairs, int64_t rows_p, int C,
                                                                  double* __restrict__ partial, unsigned int* counters,
                                                                  double count, float eps, float* __restrict__ mean,
                                                                  float* __restrict__ rstd) {
  // grid (ceil(C/64), kStatSplit, groups); block = 32 channel pairs x 8 row lanes; one 128-bit load = 2 (sum, sumsq) pairs
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cx) * 2;
  const int split = blockIdx.y, g = blockIdx.z, splits = gridDim.y;
  const int64_t per = (rows_p + splits - 1) / splits;
  const int64_t r0 = split * per, r1 = min(rows_p, r0 + per);
  double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0;
  if (c < C) {
    const float2* base = pairs + ((int64_t)g * rows_p) * C + c;
    int64_t r = r0 + ry;
    // sixteen independent 16-byte loads in flight per thread (the sums stay in row order): the small launches are one
    // DRAM round trip instead of four, the large ones keep 256 bytes per thread in flight
    for (; r + 15 * 8 < r1; r += 16 * 8) {
      float4 v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldcs(reinterpret_cast<const float4*>(base + (r + 8 * k) * C));
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        s0 += (double)v[k].x;
        q0 += (double)v[k].y;
        s1 += (double)v[k].z;
        q1 += (double)v[k].w;
      }
    }
    for (; r + 24 < r1; r += 32) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(base + r * C));
      const float4 b = __ldcs(reinterpret_cast<const float4*>(base + (r + 8) * C));
      const float4 cc = __ldcs(reinterpret_cast<const float4*>(base + (r + 16) * C));
      const float4 d = __ldcs(reinterpret_cast<const float4*>(base + (r + 24) * C));
      s0 += (double)a.x; q0 += (double)a.y; s1 += (double)a.z; q1 += (double)a.w;
      s0 += (double)b.x; q0 += (double)b.y; s1 += (double)b.z; q1 += (double)b.w;
      s0 += (double)cc.x; q0 += (double)cc.y; s1 += (double)cc.z; q1 += (double)cc.w;
      s0 += (double)d.x; q0 += (double)d.y; s1 += (double)d.z; q1 += (double)d.w;
    }
    for (; r < r1; r += 8) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(base + r * C));
      s0 += (double)v.x;
      q0 += (double)v.y;
      s1 += (double)v.z;
      q1 += (double)v.w;
    }
  }
  __shared__ double sh[4][8][32];
  sh[0][ry][cx] = s0;
  sh[1][ry][cx] = q0;
  sh[2][ry][cx] = s1;
  sh[3][ry][cx] = q1;
  __syncthreads();
  if (ry == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      s0 += sh[0][k][cx];
      q0 += sh[1][k][cx];
      s1 += sh[2][k][cx];
      q1 += sh[3][k][cx];
    }
    double* o = partial + (((int64_t)g * splits + split) * C + c) * 2;
    o[0] = s0;
    o[1] = q0;
    o[2] = s1;
    o[3] = q1;
  }
  if (counters != nullptr) {   // single launch: the last block of this (group, channel block) runs the second stage
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int slot = g * gridDim.x + blockIdx.x;
      const unsigned int prev = atomicAdd(counters + slot, 1u);
      s_last = (prev == gridDim.y - 1);
      if (s_last) counters[slot] = 0u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      // second stage over all 256 threads: thread = (split range q of 4, channel): its partial sums are added in split
      // order, the four ranges in range order -> a fixed summation order, one round trip to L2 instead of eight
      const int ch = threadIdx.x & 63, q = threadIdx.x >> 6;
      const int cc = blockIdx.x * 64 + ch;
      const int per_q = (splits + 3) >> 2, k0 = q * per_q, k1 = min(splits, k0 + per_q);
      double s = 0.0, sq = 0.0;
      if (cc < C) {
#pragma unroll 16
        for (int k = k0; k < k1; ++k) {
          const double2 o = __ldcg(reinterpret_cast<const double2*>(partial + (((int64_t)g * splits + k) * C + cc) * 2));
          s += o.x;
          sq += o.y;
        }
      }
      __shared__ double fin[2][4][64];
      fin[0][q][ch] = s;
      fin[1][q][ch] = sq;
      __syncthreads();
      if (threadIdx.x < 64 && cc < C) {
        const double st = ((fin[0][0][ch] + fin[0][1][ch]) + fin[0][2][ch]) + fin[0][3][ch];
        const double qt = ((fin[1][0][ch] + fin[1][1][ch]) + fin[1][2][ch]) + fin[1][3][ch];
        const double mu = st / count;
        double var = qt / count - mu * mu;
        if (var < 0.0) var = 0.0;
        mean[(int64_t)g * C + cc] = (float)mu;
        rstd[(int64_t)g * C + cc] = (float)(1.0 / sqrt(var + (double)eps));
      }
    }
  }
}

__global__ void stats_pairs_finalize_kernel(const double* __restrict__ partial, int C, int64_t count, float eps,
                                            float* __restrict__ mean, float* __restrict__ rstd, int splits) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < splits; ++k) {
    const double* o = partial + (((int64_t)g * splits + k) * C + c) * 2;
    s += o[0];
    q += o[1];
  }
  const double mu = s / (double)count;
  double var = q / (double)count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[(int64_t)g * C + c] = (float)mu;
  rstd[(int64_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

int channel_stats_from_pairs(const float2* pairs, int groups, int64_t rows_p, int64_t count_per_group, int C, float eps,
                             double* partial, float* mean, float* rstd, cudaStream_t st, unsigned int* counters) {
  MSR_REQUIRE(pairs && partial && mean && rstd && groups > 0 && rows_p > 0 && C > 0, "stats_from_pairs: bad arguments");
  MSR_REQUIRE(C % 2 == 0, "stats_from_pairs: channel count must be even");
  ProfileScope prof(MSR_PROF_STATS, st, (double)groups * rows_p * C * 8.0, counters ? 1 : 2);
  const int splits = stat_splits(rows_p, 8 * 32);   // >= 32 rows per row lane and block
  stats_pairs_partial_kernel<<<dim3(ceil_div(C, 64), splits, groups), 256, 0, st>>>(
      pairs, rows_p, C, partial, counters, (double)count_per_group, eps, mean, rstd);
  MSR_LAUNCH_CHECK();
  if (counters == nullptr) {
    stats_pairs_finalize_kernel<<<dim3(ceil_div(C, 128), groups), 128, 0, st>>>(partial, C, count_per_group, eps, mean,
                                                                                 rstd, splits);
    MSR_LAUNCH_CHECK();
  }
  count_launch(counters ? 1 : 2);
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// SPADE modulation from CACHED gamma | beta (repeated-sample mode): spade.py:19-20's two
// convolutions depend only on the source, so across the R generations of one batch they are computed once
// (TC_EPI_ACT_BF16 epilogue, raw bf16 columns in the interleaved 64 gamma | 64 beta order) and every generation only
// runs  act = lrelu(gamma * (x - mean) * rstd + beta)  (spade.py:21-24, blocks.py:30) -- HBM-bound: per pixel and 8
// channels one 16-byte load each of gamma and beta, 32 bytes of x (shared by 4 pixels when x is stored at half
// resolution), one 16-byte store.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spade_modulate_cached_kernel(const __nv_bfloat16* __restrict__ gb,
                                                                    const float* __restrict__ x, int x_shift,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd,
                                                                    __nv_bfloat16* __restrict__ out, int n, int lr, int C,
                                                                    int samples_per_group, float slope) {
  const int c8n = C >> 3;                      // threads per pixel (C is a multiple of 64)
  const int64_t total = ((int64_t)n << (2 * lr)) * c8n;
  const int r = 1 << lr, rs = r >> x_shift;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % c8n) * 8;          // c8n is small; 64-bit modulo by a 32-bit value only once per 8 channels
    const int64_t m = e / c8n;
    const int b = (int)(m >> (2 * lr));
    const int rem = (int)(m & (((int64_t)1 << (2 * lr)) - 1));
    const int h = rem >> lr, w = rem & (r - 1);
    const int j = c >> 6, t = c & 63;          // channel block of 64: columns [128 j, 128 j + 64) gamma, then beta
    const __nv_bfloat16* gp = gb + m * (2 * (int64_t)C) + 128 * j + t;
    const uint4 gq = __ldcs(reinterpret_cast<const uint4*>(gp));
    const uint4 bq = __ldcs(reinterpret_cast<const uint4*>(gp + 64));
    const int64_t xr = ((int64_t)b * rs + (h >> x_shift)) * rs + (w >> x_shift);
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + xr * C + c));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(x + xr * C + c + 4));
    const int g = b / samples_per_group;
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean + (int64_t)g * C + c));
    const float4 m1 = __ldg(reinterpret_cast<const float4*>(mean + (int64_t)g * C + c + 4));
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rstd + (int64_t)g * C + c));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(rstd + (int64_t)g * C + c + 4));
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float mu[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const float rsd[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const __nv_bfloat16* gv = reinterpret_cast<const __nv_bfloat16*>(&gq);
    const __nv_bfloat16* bv = reinterpret_cast<const __nv_bfloat16*>(&bq);
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = fmaf(__bfloat162float(gv[q]), (xv[q] - mu[q]) * rsd[q], __bfloat162float(bv[q]));
      v = v > 0.f ? v : v * slope;
      o[q] = __float2bfloat16_rn(v);
    }
    *reinterpret_cast<uint4*>(out + m * C + c) = *reinterpret_cast<const uint4*>(o);
  }
}

int spade_modulate_cached_bf16(const __nv_bfloat16* gb, const float* x, int x_shift, const float* mean, const float* rstd,
                               __nv_bfloat16* out, int n, int r, int C, int samples_per_group, float slope,
                               cudaStream_t st) {
  MSR_REQUIRE(gb && x && mean && rstd && out && n > 0 && r > 0 && (r & (r - 1)) == 0 && C % 64 == 0 &&
                  samples_per_group > 0, "spade_modulate_cached: bad arguments");
  int lr = 0;
  while ((1 << lr) < r) ++lr;
  const int64_t total = (int64_t)n * r * r * (C / 8);
  ProfileScope prof(MSR_PROF_ELEMWISE, st, (double)n * r * r * C * (4.0 + 2.0 + 4.0 / (1 << (2 * x_shift))));
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  spade_modulate_cached_kernel<<<blocks, 256, 0, st>>>(gb, x, x_shift, mean, rstd, out, n, lr, C, samples_per_group, slope);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

}  // namespace msr
