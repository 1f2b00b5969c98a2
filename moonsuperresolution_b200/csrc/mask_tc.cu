// SPADE's mask convolution on the tensor cores WITHOUT an im2col buffer in HBM.
//
//   mask = tf.image.resize(raw_mask, x.shape[1:3], 'nearest');  a = relu(conv3x3(mask, 2 -> 128))     spade/models/spade.py:17-18
//
// conv_tc.cu runs this layer as a K = 64 GEMM on rows that source_patches_bf16 writes to HBM first (128 bytes per pixel,
// written once per resolution and read by each of the block's two or three SPADE layers).  Here the A operand never
// leaves the SM: producer warps build every 128-pixel x 64 tile directly in shared memory -- each thread reads the
// 9 taps of its pixel from the float32 source at the resized position (8 bytes per tap, served by L1 / L2), splits them
// into bf16 hi + lo, and writes its 128-byte row in the 128-byte-swizzled K-major layout a TMA load would have produced
// (16-byte chunk c of row i at chunk position c ^ (i & 7)); fence.proxy.async makes the generic-proxy stores visible to
// tcgen05.mma.  The weights (128 x 64 bf16, 16 KB) stay in shared memory for the whole kernel.  What remains is the
// unavoidable part: 256 bytes written per pixel -- the eight epilogue warps park relu(acc + bias) as bf16 rows in shared
// memory (swizzled box layout) and one thread hands each tile to the TMA store engine, so the next accumulator is read
// out of TMEM while the previous tiles are still on their way to HBM.
//
// MODE 1 of the same kernel is the encoder's first block (blocks.py:53-60 with apply_norm=False, networks.py:12): Conv2D(64, 3,
// strides=2, 'same', no bias) on the full-resolution source (taps (2h + ky, 2w + kx), SAME padding (0, 1)) -> LeakyReLU(0.2)
// -> split-bf16 row hi (64) | lo (64), the operand of the next encoder convolution.
//
// MODE 2 is pix2pix's first block (pix2pix.py:64-72 without the norm): Conv2D(64, 4, strides=2, 'same', no bias) -> LeakyReLU(0.3)
// (taps (2h + ky - 1, 2w + kx - 1), ky, kx in 0..3), written as 64 bf16 channels into the skip half of a 128-channel
// concat buffer.  Its K layout is the one of the im2col form it replaces: k = 2t + c -> x_hi, k = 32 + 2t + c -> x_lo of
// tap t = ky*4 + kx against [w | w] (the exact input against bf16 weights).
//
// Split-bf16 K layout of MODE 0 / 1 (the three products x_hi*w_hi + x_lo*w_hi + x_hi*w_lo ~ a float32 product; mask_tc_pack_weights):
//   k = 4t + {0, 1, 2, 3}   tap t = ky*3 + kx:  x = (hi0, hi1, lo0, lo1)   w = (whi0, whi1, whi0, whi1)
//   k = 36 + 2t + {0, 1}                        x = (hi0, hi1)             w = (wlo0, wlo1)
//   k = 54, 55                                  x = (1, 1)                 w = (bias_hi, bias_lo): the bias rides in the GEMM
//   k = 56 .. 63                                zero
#include <cstring>
#include <vector>

#include "nn.cuh"
#include "tc_ptx.cuh"

namespace msr {

namespace tc {

constexpr int kMkStages = 5;                          // A tiles (16 KB each) in flight between producers and the MMA thread
constexpr int kMkAcc = 4;                             // 128-column accumulators in TMEM
constexpr int kMkN = 128;
constexpr int kMkEpiWarps = 8, kMkProdWarps = 4;       // producers: one thread per tile row
constexpr int kMkOutBufs = 3;                          // staging buffers of the TMA store (two stores may be in flight)
constexpr int kMkThreads = (kMkEpiWarps + kMkProdWarps + 1) * 32;
constexpr int kMkABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kMkBBytes = kMkN * kBlockK * 2;         // 16 KB
constexpr int kMkStageOut = 2 * kBlockM * 128;        // one output tile: two 128-row x 128-byte boxes (128-byte swizzle)
constexpr int kMkSmemBytes = 1024 + kMkBBytes + kMkStages * kMkABytes + kMkOutBufs * kMkStageOut + 256;

struct MaskGeom {
  int n, r, lr;            // output side r = 2^lr
  int I, f, half;          // source side, I / r, (I / r) / 2: mask pixel (h, w) = source pixel (h*f + half, w*f + half)
  int64_t M;               // n * r * r pixels
  int n_tiles;             // ceil(M / 128)
  float slope;             // MODE 1: LeakyReLU slope
  const float* src;        // [n][I][I][2] float32
  __nv_bfloat16* out;      // [M][128]: MODE 0 the 128 channels; MODE 1 hi (64) | lo (64); MODE 2 channels [64, 128)
};

__device__ __forceinline__ void mbar_arrive_release(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA store of one box (shared -> global, bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kMkThreads, 1)
mask_conv_tc_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o, const MaskGeom g) {
  constexpr int kN = MODE == 0 ? kMkN : 64;          // GEMM columns
  constexpr int kTaps = MODE == 2 ? 16 : 9, kSide = MODE == 2 ? 4 : 3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kOffA = kMkBBytes;
  constexpr int kOffOut = kOffA + kMkStages * kMkABytes;
  constexpr int kOffBar = kOffOut + kMkOutBufs * kMkStageOut;
  const uint32_t bar_base = smem_base + kOffBar;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMkStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMkStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMkStages + kMkAcc + a); };
  const uint32_t w_bar = bar_base + 8u * (2 * kMkStages + 2 * kMkAcc);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBar + 8 * (2 * kMkStages + 2 * kMkAcc + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMkStages; ++s) {
      mbar_init(full_bar(s), kBlockM);             // every thread of the group that builds the tile arrives once
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kMkAcc; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kMkEpiWarps);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  constexpr int kProdWarp0 = kMkEpiWarps, kMmaWarp = kProdWarp0 + kMkProdWarps;
  if (warp == kMmaWarp) {
    if (lane == 0) {
      tma_prefetch_desc(&map_b);
      tma_prefetch_desc(&map_o);
    }
    tmem_alloc(smem_u32((const void*)tmem_slot), kMkAcc * kN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kProdWarp0 && warp < kMmaWarp) {
    // ===================== producers: one thread per tile row (pixel) =====================
    // Every thread keeps the 9 taps of its next TWO tiles in flight while it converts and stores the current one: a tile
    // costs one L2 round trip (~800 clk) of latency against ~250 clk of work, so an unpipelined producer starves the
    // tensor core and the epilogue.
    const int i = threadIdx.x - kProdWarp0 * 32;     // tile row 0..127
    const int r = g.r, lr = g.lr, f = g.f, I = g.I;
    const uint32_t row_off = (uint32_t)i * 128u, sw = (uint32_t)(i & 7);
    auto load_taps = [&](int tile, float2 (&s)[kTaps]) {
      const int64_t m = (int64_t)tile * kBlockM + i;
#pragma unroll
      for (int t = 0; t < kTaps; ++t) s[t] = make_float2(0.f, 0.f);
      if (tile < g.n_tiles && m < g.M) {
        const int b = (int)(m >> (2 * lr));
        const int rem = (int)(m & (((int64_t)1 << (2 * lr)) - 1));
        const int h = rem >> lr, w = rem & (r - 1);
        const float* img = g.src + (int64_t)b * I * I * 2;
#pragma unroll
        for (int ky = 0; ky < kSide; ++ky) {
#pragma unroll
          for (int kx = 0; kx < kSide; ++kx) {
            int sy, sx;
            bool ok;
            if constexpr (MODE == 0) {   // SAME padding (1, 1) on the resized mask; nearest resize, half-pixel centres
              const int hh = h + ky - 1, ww = w + kx - 1;
              ok = hh >= 0 && hh < r && ww >= 0 && ww < r;
              sy = hh * f + g.half;
              sx = ww * f + g.half;
            } else if constexpr (MODE == 1) {   // stride-2 taps on the source itself, SAME padding (0, 1)
              sy = 2 * h + ky;
              sx = 2 * w + kx;
              ok = sy < I && sx < I;
            } else {                     // 4x4 stride-2 taps, SAME padding (1, 1)
              sy = 2 * h + ky - 1;
              sx = 2 * w + kx - 1;
              ok = sy >= 0 && sy < I && sx >= 0 && sx < I;
            }
            if (ok) s[ky * kSide + kx] = __ldg(reinterpret_cast<const float2*>(img + ((int64_t)sy * I + sx) * 2));
          }
        }
      }
    };
    const int tile_step = (int)gridDim.x;
    int stage = 0;
    uint32_t phase = 0;
    // converts the taps in `s` (tile `tile`), refills `s` with the taps of the tile two steps ahead, writes the row
    auto produce = [&](int tile, float2 (&s)[kTaps]) {
      uint32_t ex[kTaps], ey[kTaps];   // per tap: (hi0 | hi1 << 16), (lo0 | lo1 << 16); x = hi + lo to ~2^-17
#pragma unroll
      for (int t = 0; t < kTaps; ++t) {
        const __nv_bfloat162 hi = __floats2bfloat162_rn(s[t].x, s[t].y);
        const float2 hf = __bfloat1622float2(hi);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(s[t].x - hf.x, s[t].y - hf.y);
        ex[t] = *reinterpret_cast<const uint32_t*>(&hi);
        ey[t] = *reinterpret_cast<const uint32_t*>(&lo);
      }
      load_taps(tile + 2 * tile_step, s);            // in flight during the wait and the stores below and the next tile
      mbar_wait(empty_bar(stage), phase ^ 1u);
      uint8_t* row = smem_gen + kOffA + stage * kMkABytes + row_off;
      auto put = [&](uint32_t c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
        *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(a0, a1, a2, a3);
      };
      if constexpr (MODE == 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          put(q, ex[4 * q], ex[4 * q + 1], ex[4 * q + 2], ex[4 * q + 3]);       // k = 2t + c: hi
          put(4 + q, ey[4 * q], ey[4 * q + 1], ey[4 * q + 2], ey[4 * q + 3]);   // k = 32 + 2t + c: lo
        }
      } else {
        put(0, ex[0], ey[0], ex[1], ey[1]);
        put(1, ex[2], ey[2], ex[3], ey[3]);
        put(2, ex[4], ey[4], ex[5], ey[5]);
        put(3, ex[6], ey[6], ex[7], ey[7]);
        put(4, ex[8], ey[8], ex[0], ex[1]);
        put(5, ex[2], ex[3], ex[4], ex[5]);
        put(6, ex[6], ex[7], ex[8], 0x3f803f80u);   // k = 54, 55: bf16 ones against the bias rows of the weights
        put(7, 0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive_release(full_bar(stage));
      if (++stage == kMkStages) {
        stage = 0;
        phase ^= 1u;
      }
    };
    float2 s0[kTaps], s1[kTaps];
    int tile = blockIdx.x;
    load_taps(tile, s0);
    load_taps(tile + tile_step, s1);
    while (tile < g.n_tiles) {
      produce(tile, s0);
      tile += tile_step;
      if (tile >= g.n_tiles) break;
      produce(tile, s1);
      tile += tile_step;
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (+ the one-off weight load) =====================
    if (lane == 0) {
      mbar_expect_tx(w_bar, kN * kBlockK * 2);
      tma_load_2d(smem_base, &map_b, w_bar, 0, 0);
      constexpr uint32_t idesc = make_idesc(kN, kBlockM);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      const uint64_t bdesc = make_smem_desc(smem_base);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_base + kOffA + stage * kMkABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(acc * kN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                    k != 0 ? 1u : 0u);
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(acc));
        if (++stage == kMkStages) {
          stage = 0;
          phase ^= 1u;
        }
        if (++acc == kMkAcc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps 0..7: relu(acc + bias) -> bf16 -> TMA store =====================
    // warp = (TMEM lane quarter, column half): lane l owns tile row i = quarter*32 + l.  The tile leaves as two TMA stores
    // of 128 rows x 128 bytes (columns [0, 64) and [64, 128)), so the rows are parked in shared memory in the 128-byte
    // swizzled box layout (chunk c of row i at c ^ (i & 7)): no bank conflicts on the way in, no shared-memory reads or
    // global store instructions on the way out (the L1 / LSU path, not HBM, bounded the version that copied the rows
    // out with LDS + STG: l1tex 87 % busy at 4.0 TB/s).  Three staging buffers: thread 0 lets at most one store be
    // pending before the barrier of tile k, so the buffer of tile k + 1 (last used by tile k - 2) is free after it.
    const int quarter = warp & 3, csel = warp >> 2;
    const int i = quarter * 32 + lane;
    const uint32_t row_off = (uint32_t)i * 128u, sw = (uint32_t)(i & 7);
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kN + csel * (kN / 2));
      uint32_t v0[32], v1[32];
      tmem_ld32(t_row, v0);
      if constexpr (MODE == 0) tmem_ld32(t_row + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));     // the accumulator is in registers: the tensor core may reuse it
      uint8_t* stg = smem_gen + kOffOut + buf * kMkStageOut;
      auto put = [&](int box, uint32_t c, uint4 val) {  // 16-byte chunk c of this thread's row in box 0 / 1
        *reinterpret_cast<uint4*>(stg + box * (kBlockM * 128) + row_off + ((c ^ sw) << 4)) = val;
      };
      if constexpr (MODE == 0) {
        auto pack8 = [&](const uint32_t* v, int q) {   // relu (the bias is already in the accumulator) -> 8 bf16
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 t2 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * q + 2 * j]), 0.f),
                                                            fmaxf(__uint_as_float(v[8 * q + 2 * j + 1]), 0.f));
            pk[j] = *reinterpret_cast<const uint32_t*>(&t2);
          }
          return make_uint4(pk[0], pk[1], pk[2], pk[3]);
        };
        // columns [64*csel, 64*csel + 64) = box csel, chunks 0..7
#pragma unroll
        for (int q = 0; q < 4; ++q) put(csel, q, pack8(v0, q));
#pragma unroll
        for (int q = 0; q < 4; ++q) put(csel, 4 + q, pack8(v1, q));
      } else if constexpr (MODE == 2) {
        // columns [32*csel, 32*csel + 32): leaky_relu -> bf16 into the one box, chunks 4*csel .. + 3
        const float sl = g.slope;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a0 = __uint_as_float(v0[8 * q + 2 * j]), a1 = __uint_as_float(v0[8 * q + 2 * j + 1]);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(a0, a0 * sl), fmaxf(a1, a1 * sl));
            pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          put(0, (uint32_t)(4 * csel + q), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      } else {
        // columns [32*csel, 32*csel + 32): leaky_relu -> hi into box 0, lo = v - hi into box 1, chunks 4*csel .. + 3
        const float sl = g.slope;   // 0 < slope < 1: leaky_relu(x) = max(x, slope * x)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a0 = __uint_as_float(v0[8 * q + 2 * j]), a1 = __uint_as_float(v0[8 * q + 2 * j + 1]);
            const float o0 = fmaxf(a0, a0 * sl), o1 = fmaxf(a1, a1 * sl);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(o0, o1);
            const float2 hf = __bfloat1622float2(h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(o0 - hf.x, o1 - hf.y);
            ph[j] = *reinterpret_cast<const uint32_t*>(&h2);
            pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          put(0, (uint32_t)(4 * csel + q), make_uint4(ph[0], ph[1], ph[2], ph[3]));
          put(1, (uint32_t)(4 * csel + q), make_uint4(pl[0], pl[1], pl[2], pl[3]));
        }
      }
      fence_proxy_async_smem();                         // generic-proxy stores -> visible to the TMA store
      if (threadIdx.x == 0) tma_store_wait_read<1>();   // the store issued two tiles ago has finished reading its buffer
      asm volatile("bar.sync 1, %0;" ::"n"(kMkEpiWarps * 32) : "memory");
      if (threadIdx.x == 0) {
        const uint32_t s0 = smem_base + kOffOut + buf * kMkStageOut;
        if constexpr (MODE == 2) {
          tma_store_2d(&map_o, s0, 64, tile * kBlockM);                    // the skip half of the concat buffer
        } else {
          tma_store_2d(&map_o, s0, 0, tile * kBlockM);                     // rows beyond M are clipped by the tensor map
          tma_store_2d(&map_o, s0 + kBlockM * 128, 64, tile * kBlockM);
        }
        tma_store_commit();
      }
      if (++acc == kMkAcc) {
        acc = 0;
        acc_phase ^= 1u;
      }
      if (++buf == kMkOutBufs) buf = 0;
    }
    if (threadIdx.x == 0) tma_store_wait_all();         // all rows are in global memory before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kMkAcc * kN);
  }
}

}  // namespace tc

// ---- host side ----------------------------------------------------------------------------------------------------------
void mask_tc_pack_weights(const float* w, const float* bias, int cout, std::vector<uint16_t>* out) {
  auto f2bf = [](float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);   // round to nearest even (finite inputs)
  };
  auto bf2f = [](uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  out->assign((size_t)cout * 64, 0);
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < 2; ++c)
      for (int co = 0; co < cout; ++co) {
        const float v = w[(size_t)(t * 2 + c) * cout + co];
        const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
        uint16_t* row = out->data() + (size_t)co * 64;
        row[4 * t + c] = hi;          // pairs with x_hi
        row[4 * t + 2 + c] = hi;      // pairs with x_lo
        row[36 + 2 * t + c] = lo;     // pairs with x_hi
      }
  if (bias != nullptr)
    for (int co = 0; co < cout; ++co) {  // bias as two more K rows against constant ones in the operand
      const uint16_t hi = f2bf(bias[co]);
      (*out)[(size_t)co * 64 + 54] = hi;
      (*out)[(size_t)co * 64 + 55] = f2bf(bias[co] - bf2f(hi));
    }
}

bool mask_tc_supported(int I, int r) { return r >= 1 && (r & (r - 1)) == 0 && r <= I && I % r == 0; }

// mode 0: SPADE mask convolution at side r; mode 1: encoder block 1 (r = I / 2); mode 2: pix2pix block 1 (r = I / 2)
static int source_conv_tc(int mode, const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out, int n, int r,
                          float slope, cudaStream_t st) {
  MSR_REQUIRE(source && wm && out && n > 0, "source_conv_tc: bad arguments");
  MSR_REQUIRE(mask_tc_supported(I, r), "source_conv_tc: r must be a power of two dividing I");
  MSR_REQUIRE((reinterpret_cast<uintptr_t>(wm) & 127) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(source) & 7) == 0, "source_conv_tc: misaligned operand");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(MSR_E_CUDA, "source_conv_tc: cuTensorMapEncodeTiled entry point not available");
  const int cols = mode == 0 ? tc::kMkN : 64;
  CUtensorMap map_b;
  {
    cuuint64_t dims[2] = {(cuuint64_t)tc::kBlockK, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)tc::kBlockK * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::kBlockK, (cuuint32_t)cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wm), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(MSR_E_CUDA, "source_conv_tc: cuTensorMapEncodeTiled failed with " + std::to_string((int)rc));
  }
  const int64_t M_rows = (int64_t)n * r * r;
  CUtensorMap map_o;   // output [M][128] bf16, stored as boxes of 128 rows x 64 columns
  {
    cuuint64_t dims[2] = {128, (cuuint64_t)M_rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, (cuuint32_t)tc::kBlockM};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&map_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(MSR_E_CUDA, "source_conv_tc: cuTensorMapEncodeTiled(out) failed with " + std::to_string((int)rc));
  }
  tc::MaskGeom g;
  g.n = n; g.r = r; g.lr = 0;
  while ((1 << g.lr) < r) ++g.lr;
  g.I = I; g.f = I / r; g.half = g.f >> 1;
  g.M = (int64_t)n * r * r;
  g.n_tiles = (int)((g.M + tc::kBlockM - 1) / tc::kBlockM);
  g.slope = slope;
  g.src = source; g.out = out;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ProfileScope prof(MSR_PROF_CONV_TC, st, 2.0 * (double)g.M * cols * (mode == 2 ? 32 : 18));
  static bool attr_set = false;
  if (!attr_set) {
    MSR_CUDA_CHECK(cudaFuncSetAttribute(tc::mask_conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kMkSmemBytes));
    MSR_CUDA_CHECK(cudaFuncSetAttribute(tc::mask_conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kMkSmemBytes));
    MSR_CUDA_CHECK(cudaFuncSetAttribute(tc::mask_conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kMkSmemBytes));
    attr_set = true;
  }
  const int grid = std::min(g.n_tiles, sms);
  if (mode == 0) tc::mask_conv_tc_kernel<0><<<grid, tc::kMkThreads, tc::kMkSmemBytes, st>>>(map_b, map_o, g);
  else if (mode == 1) tc::mask_conv_tc_kernel<1><<<grid, tc::kMkThreads, tc::kMkSmemBytes, st>>>(map_b, map_o, g);
  else tc::mask_conv_tc_kernel<2><<<grid, tc::kMkThreads, tc::kMkSmemBytes, st>>>(map_b, map_o, g);
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

int mask_conv_tc(const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out, int n, int r, cudaStream_t st) {
  return source_conv_tc(0, source, I, wm, out, n, r, 0.f, st);
}

int p2p1_conv_tc(const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out128, int n, float slope,
                 cudaStream_t st) {
  MSR_REQUIRE(I >= 2 && I % 2 == 0, "p2p1_conv_tc: the source side must be even");
  return source_conv_tc(2, source, I, wm, out128, n, I / 2, slope, st);
}

int enc1_conv_tc(const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out, int n, float slope, cudaStream_t st) {
  MSR_REQUIRE(I >= 2 && I % 2 == 0, "enc1_conv_tc: the source side must be even");
  return source_conv_tc(1, source, I, wm, out, n, I / 2, slope, st);
}

}  // namespace msr
