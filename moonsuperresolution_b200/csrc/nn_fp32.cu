// fp32 CUDA-core operators of the generator graphs (parity mode, and the non-GEMM glue of the bf16 mode).
//
// Reference semantics: spade/models/{networks,blocks,spade,sampling}.py, pix2pix.py:64-108, and the TensorFlow rules of
// SURVEY.md Appendix B (asymmetric SAME padding, half-pixel nearest resize, biased batch moments, ...).
#include "nn.cuh"

namespace msr {

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Generic NHWC conv as implicit GEMM: M = n*Ho*Wo, N = cout, K = kh*kw*cin.  64x64x16 tiles, 4x4 per thread.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_f32_kernel(ConvF32 p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int64_t M = (int64_t)p.n * p.Ho * p.Wo;
  const int K = p.kh * p.kw * p.cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load role: one output pixel (row) and 4 consecutive k per thread
  const int a_row = t >> 2, a_kq = (t & 3) * 4;
  const int64_t am = m0 + a_row;
  const bool a_ok = am < M;
  int an = 0, aho = 0, awo = 0;
  if (a_ok) {
    an = (int)(am / ((int64_t)p.Ho * p.Wo));
    const int rem = (int)(am % ((int64_t)p.Ho * p.Wo));
    aho = rem / p.Wo;
    awo = rem % p.Wo;
  }
  // B-load role
  const int b_row = t >> 4, b_col = (t & 15) * 4;
  // compute role
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  const bool vec_a = (p.cin % 4 == 0) && (p.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
  const bool vec_b = (p.cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0);

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok) {
      const int kg = k0 + a_kq;
      if (vec_a) {
        if (kg < K) {
          const int tap = kg / p.cin, ci = kg % p.cin;
          const int ky = tap / p.kw, kx = tap % p.kw;
          int vy, vx;
          bool ok = true;
          if (!p.transposed) {
            vy = aho * p.stride - p.pad_t + ky;
            vx = awo * p.stride - p.pad_l + kx;
          } else {
            const int ty2 = aho + p.pad_t - ky, tx2 = awo + p.pad_l - kx;
            ok = (ty2 >= 0) && (tx2 >= 0) && (ty2 % p.stride == 0) && (tx2 % p.stride == 0);
            vy = ty2 / p.stride;
            vx = tx2 / p.stride;
          }
          if (ok && vy >= 0 && vy < p.Hv && vx >= 0 && vx < p.Wv) {
            const int sy = (vy * p.in_mul + p.in_add) >> p.in_shift, sx = (vx * p.in_mul + p.in_add) >> p.in_shift;
            const float4 v = *reinterpret_cast<const float4*>(p.x + (((int64_t)an * p.Hs + sy) * p.Ws + sx) * p.ldx + ci);
            av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kgj = kg + j;
          if (kgj >= K) continue;
          const int tap = kgj / p.cin, ci = kgj % p.cin;
          const int ky = tap / p.kw, kx = tap % p.kw;
          int vy, vx;
          bool ok = true;
          if (!p.transposed) {
            vy = aho * p.stride - p.pad_t + ky;
            vx = awo * p.stride - p.pad_l + kx;
          } else {
            const int ty2 = aho + p.pad_t - ky, tx2 = awo + p.pad_l - kx;
            ok = (ty2 >= 0) && (tx2 >= 0) && (ty2 % p.stride == 0) && (tx2 % p.stride == 0);
            vy = ty2 / p.stride;
            vx = tx2 / p.stride;
          }
          if (ok && vy >= 0 && vy < p.Hv && vx >= 0 && vx < p.Wv) {
            const int sy = (vy * p.in_mul + p.in_add) >> p.in_shift, sx = (vx * p.in_mul + p.in_add) >> p.in_shift;
            av[j] = p.x[(((int64_t)an * p.Hs + sy) * p.Ws + sx) * p.ldx + ci];
          }
        }
      }
      if (p.in_slope != 1.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) av[j] = av[j] > 0.f ? av[j] : av[j] * p.in_slope;
      }
    }
    // ---- B tile
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const int kg = k0 + b_row;
      const int c = n0 + b_col;
      if (kg < K) {
        if (vec_b && c + 3 < p.cout) {
          const float4 v = *reinterpret_cast<const float4*>(p.w + (int64_t)kg * p.cout + c);
          bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < p.cout) bv[j] = p.w[(int64_t)kg * p.cout + c + j];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) As[a_kq + j][a_row] = av[j];
    *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int64_t res_row = 0;
    if (p.res) {
      const int nn = (int)(m / ((int64_t)p.Ho * p.Wo));
      const int rem = (int)(m % ((int64_t)p.Ho * p.Wo));
      const int ho = rem / p.Wo, wo = rem % p.Wo;
      const int rh = p.Ho >> p.res_shift, rw = p.Wo >> p.res_shift;
      res_row = ((int64_t)nn * rh + (ho >> p.res_shift)) * rw + (wo >> p.res_shift);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[c];
      if (p.res) v += p.res[res_row * p.ldres + c];
      p.y[m * p.ldy + c] = apply_act(v, p.act, p.act_slope);
    }
  }
}

int conv_f32(const ConvF32& p, cudaStream_t st) {
  MSR_REQUIRE(p.x && p.w && p.y && p.n > 0 && p.cin > 0 && p.cout > 0, "conv_f32: bad arguments");
  const int64_t M = (int64_t)p.n * p.Ho * p.Wo;
  dim3 grid(ceil_div(M, BM), ceil_div(p.cout, BN));
  ProfileScope prof(MSR_PROF_CONV_F32, st, 2.0 * (double)M * p.cout * p.kh * p.kw * p.cin / (p.transposed ? p.stride * p.stride : 1));
  conv_f32_kernel<<<grid, 256, 0, st>>>(p);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Channel statistics: deterministic two-stage reduction, fp64 partials.
// ------------------------------------------------------------------------------------------------------------------
// grid (ceil(C/128), kStatSplit, groups); block = 32 channel quads x 8 row lanes; 128-bit loads (C % 4 == 0, ld % 4 == 0)
// `counters` != null: the last of the kStatSplit blocks of a (group, channel block) to finish also does the second stage
// (fixed summation order, so the result does not depend on which block that is) -- one launch instead of two.
__device__ __forceinline__ void stats_finalize_channel(const double* __restrict__ partial, int g, int c, int C,
                                                       double count, float eps, float* __restrict__ mean,
                                                       float* __restrict__ rstd) {
  const int splits = gridDim.y;
  double s = 0.0, q = 0.0;
#pragma unroll 8
  for (int k = 0; k < splits; ++k) {
    const double2 o = __ldcg(reinterpret_cast<const double2*>(partial + (((int64_t)g * splits + k) * C + c) * 2));
    s += o.x;
    q += o.y;
  }
  const double mu = s / count;
  double var = q / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[(int64_t)g * C + c] = (float)mu;
  rstd[(int64_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

__device__ __forceinline__ bool stats_block_is_last(unsigned int* counters, int slot) {
  __shared__ int s_last;
  __threadfence();               // this block's partials are visible device-wide before the ticket is taken
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counters + slot, 1u);
    s_last = (prev == gridDim.y - 1);
    if (s_last) counters[slot] = 0u;   // ready for the next launch (stream order)
  }
  __syncthreads();
  if (s_last) __threadfence();   // acquire side: the other blocks' partials
  return s_last != 0;
}

template <typename T>
__global__ void __launch_bounds__(256) stats_partial_kernel(const T* __restrict__ x, int ld, int64_t rows, int C,
                                                            double* __restrict__ partial, unsigned int* counters,
                                                            float eps, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cx) * 4;
  const int split = blockIdx.y, g = blockIdx.z, splits = gridDim.y;
  const int64_t per = (rows + splits - 1) / splits;
  const int64_t r0 = split * per, r1 = min(rows, r0 + per);
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < C) {
    const T* base = x + ((int64_t)g * rows) * ld + c;
#pragma unroll 4
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(base + r * ld));
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double d = (double)v[j];
        s[j] += d;
        q[j] += d * d;
      }
    }
  }
  __shared__ double sh[2][8][128];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sh[0][ry][cx * 4 + j] = s[j];
    sh[1][ry][cx * 4 + j] = q[j];
  }
  __syncthreads();
  if (ry == 0 && c < C) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double ss = s[j], qq = q[j];
#pragma unroll
      for (int k = 1; k < 8; ++k) {
        ss += sh[0][k][cx * 4 + j];
        qq += sh[1][k][cx * 4 + j];
      }
      double* o = partial + (((int64_t)g * splits + split) * C + c + j) * 2;
      o[0] = ss;
      o[1] = qq;
    }
  }
  if (counters != nullptr && stats_block_is_last(counters, g * gridDim.x + blockIdx.x)) {
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (threadIdx.x < 128 && cc < C) stats_finalize_channel(partial, g, cc, C, (double)rows, eps, mean, rstd);
  }
}

__global__ void stats_finalize_kernel(const double* __restrict__ partial, int C, int64_t rows, float eps,
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
  const double mu = s / (double)rows;
  double var = q / (double)rows - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[(int64_t)g * C + c] = (float)mu;
  rstd[(int64_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

template <typename T>
static int channel_stats_impl(const T* x, int ld, int groups, int64_t rows, int C, float eps, double* partial,
                              float* mean, float* rstd, unsigned int* counters, cudaStream_t st) {
  MSR_REQUIRE(x && partial && mean && rstd && groups > 0 && rows > 0 && C > 0, "channel_stats: bad arguments");
  MSR_REQUIRE(C % 4 == 0 && ld % 4 == 0, "channel_stats: channel count and pitch must be multiples of 4");
  ProfileScope prof(MSR_PROF_STATS, st, (double)groups * rows * C * sizeof(T), counters ? 1 : 2);
  // a block walks >= 64 rows per row lane (8 lanes): few fat blocks instead of kStatSplit thin ones when the per-group
  // row count is small (per-image moments of the encoder, the 8 x 8 first generator block)
  int splits = stat_splits(rows, 8 * 64);
  {  // ... but never fewer blocks than fill the chip a few times over when the rows allow it (>= 8 rows per row lane): the
     // 8 x 8 first generator block (8 groups x 1024 channels, 1024 rows each) used to run on 128 blocks
    const int base_blocks = ceil_div(C, 128) * groups;
    const int want = ceil_div(4 * 148, base_blocks);
    const int by_rows = (int)std::max<int64_t>(1, rows / 64);
    splits = std::min(kStatSplit, std::max(splits, std::min(want, by_rows)));
  }
  stats_partial_kernel<T><<<dim3(ceil_div(C, 128), splits, groups), 256, 0, st>>>(x, ld, rows, C, partial, counters, eps,
                                                                                  mean, rstd);
  MSR_LAUNCH_CHECK();
  if (counters == nullptr) {
    stats_finalize_kernel<<<dim3(ceil_div(C, 128), groups), 128, 0, st>>>(partial, C, rows, eps, mean, rstd, splits);
    MSR_LAUNCH_CHECK();
  }
  count_launch(counters ? 1 : 2);
  return MSR_OK;
}
int channel_stats_f32(const float* x, int ld, int groups, int64_t rows, int C, float eps, double* partial, float* mean,
                      float* rstd, cudaStream_t st, unsigned int* counters) {
  return channel_stats_impl<float>(x, ld, groups, rows, C, eps, partial, mean, rstd, counters, st);
}

// ------------------------------------------------------------------------------------------------------------------
// SPADE modulation, fp32 path
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spade_modulate_kernel(const float* __restrict__ gb, const float* __restrict__ x,
                                                             int x_shift, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ out,
                                                             int n, int r, int C, int samples_per_group, float slope) {
  const int64_t total = (int64_t)n * r * r * (C / 4);
  const int c4n = C / 4;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % c4n) * 4;
    const int64_t m = e / c4n;
    const int nn = (int)(m / ((int64_t)r * r));
    const int rem = (int)(m % ((int64_t)r * r));
    const int h = rem / r, w = rem % r;
    const int rs = r >> x_shift;
    const int64_t xr = ((int64_t)nn * rs + (h >> x_shift)) * rs + (w >> x_shift);
    const int g = nn / samples_per_group;
    const float4 xv = *reinterpret_cast<const float4*>(x + xr * C + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + (int64_t)g * C + c);
    const float4 rs4 = *reinterpret_cast<const float4*>(rstd + (int64_t)g * C + c);
    const float4 ga = *reinterpret_cast<const float4*>(gb + m * 2 * C + c);
    const float4 be = *reinterpret_cast<const float4*>(gb + m * 2 * C + C + c);
    float4 o;
    o.x = ga.x * ((xv.x - mu.x) * rs4.x) + be.x;
    o.y = ga.y * ((xv.y - mu.y) * rs4.y) + be.y;
    o.z = ga.z * ((xv.z - mu.z) * rs4.z) + be.z;
    o.w = ga.w * ((xv.w - mu.w) * rs4.w) + be.w;
    o.x = o.x > 0.f ? o.x : o.x * slope;
    o.y = o.y > 0.f ? o.y : o.y * slope;
    o.z = o.z > 0.f ? o.z : o.z * slope;
    o.w = o.w > 0.f ? o.w : o.w * slope;
    *reinterpret_cast<float4*>(out + m * C + c) = o;
  }
}

int spade_modulate_f32(const float* gb, const float* x, int x_shift, const float* mean, const float* rstd, float* out,
                       int n, int r, int C, int samples_per_group, float slope, cudaStream_t st) {
  MSR_REQUIRE(gb && x && mean && rstd && out && C % 4 == 0, "spade_modulate: bad arguments");
  ProfileScope prof(MSR_PROF_ELEMWISE, st, (double)n * r * r * C * 4.0 * 4.0);
  const int64_t total = (int64_t)n * r * r * (C / 4);
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  spade_modulate_kernel<<<blocks, 256, 0, st>>>(gb, x, x_shift, mean, rstd, out, n, r, C, samples_per_group, slope);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Normalise + affine + activation (tfa InstanceNormalization / Keras BatchNormalization at inference)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, int ldx,
                                                         const float* __restrict__ mean,
                                                         const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ y, int ldy,
                                                         int64_t M, int C, int64_t rows_per_group, int act,
                                                         float slope) {
  const int64_t total = M * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t m = e / C;
    float v = x[m * ldx + c];
    if (mean) {
      const int64_t g = m / rows_per_group;
      v = (v - mean[g * C + c]) * rstd[g * C + c];
    }
    if (gamma) v *= gamma[c];
    if (beta) v += beta[c];
    y[m * ldy + c] = apply_act(v, act, slope);
  }
}

int affine_act_f32(const float* x, int ldx, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, float* y, int ldy, int64_t M, int C, int64_t rows_per_group, int act,
                   float slope, cudaStream_t st) {
  MSR_REQUIRE(x && y && M > 0 && C > 0 && rows_per_group > 0, "affine_act: bad arguments");
  ProfileScope prof(MSR_PROF_ELEMWISE, st, (double)M * C * 8.0);
  const int64_t total = M * C;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  affine_act_kernel<<<blocks, 256, 0, st>>>(x, ldx, mean, rstd, gamma, beta, y, ldy, M, C, rows_per_group, act, slope);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

__global__ void sampler_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ eps, float* __restrict__ latent, int64_t count) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= count) return;
  latent[e] = eps ? mean[e] + expf(0.5f * var[e]) * eps[e] : mean[e] + var[e];
}

int sampler_f32(const float* mean, const float* var, const float* eps, float* latent, int64_t count, cudaStream_t st) {
  MSR_REQUIRE(mean && var && latent && count > 0, "sampler: bad arguments");
  ProfileScope prof(MSR_PROF_ELEMWISE, st, (double)count * 16.0);
  sampler_kernel<<<ceil_div(count, 256), 256, 0, st>>>(mean, var, eps, latent, count);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

__global__ void sampler_strided_kernel(const float* __restrict__ mv, int ld, const float* __restrict__ eps,
                                       float* __restrict__ latent, int n, int L, __nv_bfloat16* __restrict__ split) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * L) return;
  const int row = (int)(e / L), c = (int)(e % L);
  const float m = mv[(int64_t)row * ld + c], v = mv[(int64_t)row * ld + L + c];
  const float z = eps ? m + expf(0.5f * v) * eps[e] : m + v;
  latent[e] = z;
  if (split) {   // split-bf16 operand of the tensor-core dense layer: row = hi (L) | lo (L)
    const __nv_bfloat16 hi = __float2bfloat16_rn(z);
    split[(int64_t)row * 2 * L + c] = hi;
    split[(int64_t)row * 2 * L + L + c] = __float2bfloat16_rn(z - __bfloat162float(hi));
  }
}

int sampler_strided_f32(const float* mv, int ld, const float* eps, float* latent, int n, int L, cudaStream_t st,
                        __nv_bfloat16* split_out) {
  MSR_REQUIRE(mv && latent && n > 0 && L > 0 && ld >= 2 * L, "sampler: bad arguments");
  ProfileScope prof(MSR_PROF_ELEMWISE, st, (double)n * L * 16.0);
  sampler_strided_kernel<<<ceil_div((int64_t)n * L, 256), 256, 0, st>>>(mv, ld, eps, latent, n, L, split_out);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Final layer: upsample x2 -> leaky_relu(0.2) -> conv 4x4 SAME (pad 1 before, 2 after) -> 1 channel.
// Block = 16x16 output pixels; the (10 x 10) low-res halo tile is staged in shared memory after the leaky relu; each
// warp walks output pixels with lanes over channels (4 per lane) and reduces with shuffles.
// ------------------------------------------------------------------------------------------------------------------
constexpr int FC_TX = 16, FC_TY = 8;   // output tile (x, y)
constexpr int FC_LX = FC_TX / 2 + 2;   // low-res tile incl. halo (rows (o-1)>>1 .. (o+2)>>1)
constexpr int FC_LY = FC_TY / 2 + 2;

__global__ void __launch_bounds__(256) final_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int n, int r) {
  __shared__ float4 tile[FC_LY * FC_LX][32];  // [low-res pixel][lane] -> 4 channels per lane
  const int R = 2 * r;
  const int img = blockIdx.z, oy0 = blockIdx.y * FC_TY, ox0 = blockIdx.x * FC_TX;
  const int ly0 = (oy0 >> 1) - 1, lx0 = (ox0 >> 1) - 1;  // low-res origin of the tile (may be -1)
  for (int e = threadIdx.x; e < FC_LY * FC_LX * 32; e += 256) {
    const int lane = e & 31, pix = e >> 5;
    const int ly = ly0 + pix / FC_LX, lx = lx0 + pix % FC_LX;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ly >= 0 && ly < r && lx >= 0 && lx < r) {
      v = *reinterpret_cast<const float4*>(x + (((int64_t)img * r + ly) * r + lx) * 128 + lane * 4);
      v.x = v.x > 0.f ? v.x : 0.2f * v.x;
      v.y = v.y > 0.f ? v.y : 0.2f * v.y;
      v.z = v.z > 0.f ? v.z : 0.2f * v.z;
      v.w = v.w > 0.f ? v.w : 0.2f * v.w;
    }
    tile[pix][lane] = v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 wr[16];
#pragma unroll
  for (int tp = 0; tp < 16; ++tp) wr[tp] = *reinterpret_cast<const float4*>(w + tp * 128 + lane * 4);
  __syncthreads();
  const float b = bias ? bias[0] : 0.f;
  for (int px = warp; px < FC_TY * FC_TX; px += 8) {
    const int oy = oy0 + px / FC_TX, ox = ox0 + px % FC_TX;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int uy = oy - 1 + ky;  // upsampled-row index, SAME pad (1, 2)
      if (uy < 0 || uy >= R) continue;
      const int ty = (uy >> 1) - ly0;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ux = ox - 1 + kx;
        if (ux < 0 || ux >= R) continue;
        const int tx = (ux >> 1) - lx0;
        const float4 v = tile[ty * FC_LX + tx][lane];
        const float4 ww = wr[ky * 4 + kx];
        acc = fmaf(v.x, ww.x, acc);
        acc = fmaf(v.y, ww.y, acc);
        acc = fmaf(v.z, ww.z, acc);
        acc = fmaf(v.w, ww.w, acc);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) out[((int64_t)img * R + oy) * R + ox] = acc + b;
  }
}

int final_conv_f32(const float* x, const float* w, const float* bias, float* out, int n, int r, cudaStream_t st) {
  MSR_REQUIRE(x && w && out && n > 0 && r > 0, "final_conv: bad arguments");
  MSR_REQUIRE((2 * r) % FC_TX == 0, "final_conv: output side must be a multiple of 16");
  ProfileScope prof(MSR_PROF_FINAL_CONV, st, (double)n * r * r * 128 * 4.0 + (double)n * 4.0 * r * r * 4.0);
  final_conv_kernel<<<dim3(2 * r / FC_TX, 2 * r / FC_TY, n), 256, 0, st>>>(x, w, bias, out, n, r);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

}  // namespace msr
