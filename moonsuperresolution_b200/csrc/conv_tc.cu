// 3x3 stride-1 SAME convolution as a tcgen05 / TMEM / TMA implicit GEMM for sm_100a (bf16 x bf16 -> fp32).
//
// Replaces the cuDNN calls behind every Conv2D(3, padding="same") of the SPADE generator:
//   main convolutions   spade/models/blocks.py:19-26,30-36      (epilogue: + bias, + residual)
//   gamma / beta convs  spade/models/spade.py:10-11,19-24       (epilogue: fused SPADE normalise-modulate-LeakyReLU)
//
// GEMM view: M = n*r*r output pixels, N = output columns, K = 9*cin with k = (ky*3 + kx)*cin + ci.
//   * an M tile is 128 pixels = NB images x TH rows x TW columns of the NHWC tensor, so for tap (ky, kx) and channel
//     block cb the A operand is ONE 4-D TMA box {64 ch, TW, TH, NB} at coordinates {64cb, w0+kx-1, h0+ky-1, b0}; the
//     SAME-padding halo is the TMA's out-of-bounds zero fill (no im2col buffer, no predicates);
//   * TMA writes the box as 128 rows of 128 bytes with the 128-byte swizzle = the canonical K-major UMMA layout;
//   * B (weights, [N][K] K-major) is a 2-D TMA box {64, BN};
//   * tcgen05.mma cta_group::1, M=128, N=BN, K=16, fp32 accumulators in TMEM, double buffered (2 x BN columns) so
//     the epilogue of tile i overlaps the main loop of tile i+1;
//   * persistent CTAs (one per SM), 10 warps: 0-7 epilogue (TMEM lane quarter = warp % 4, the two warps of a quarter
//     take alternate 32-column chunks), 8 TMA producer, 9 MMA issuer.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "nn.cuh"
#include "tc_ptx.cuh"

namespace msr {

namespace tc {

constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kEpiWarps = 8;                      // two per TMEM lane quarter: they split the accumulator columns
constexpr int kThreads = (kEpiWarps + 2) * 32;   // + TMA producer warp + MMA issuer warp

// CTAS = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x BN tile; each CTA stages its own 128 rows of A and
// only HALF of the B tile, which cuts the L2 -> shared-memory traffic per FLOP (the measured limiter) by a third.
// STRIP: for tiles that are 128 consecutive pixels of ONE image row (r >= 128), a stage holds the 130-pixel input strip
// of one (ky, channel block) and the three weight tiles of kx = 0, 1, 2; the three taps read the SAME strip through
// shared-memory descriptors shifted by kx rows, which cuts the A-operand L2 -> smem traffic (the measured limiter of the
// K = 1152 / 2304 layers) by 3.
constexpr int kStripRows = kBlockM + 2;
constexpr int kStripBytes = ((kStripRows * 128 + 1023) / 1024) * 1024;   // 17 KB, keeps the weight tiles 1024-aligned
// Staging area of the row-transposing epilogue (TC_EPI_RELU_BF16_T): per TMEM lane quarter 32 rows of 256 bytes, rows
// padded by 16 bytes so that 8 lanes writing / reading 16 bytes each at consecutive rows hit distinct banks.
constexpr int kTRowPitch = 256 + 16;
constexpr int kTStageBytes = 4 * 32 * kTRowPitch;
template <int BN, int CTAS = 1, bool STRIP = false, int EXTRA = 0>
struct Cfg {
  static constexpr int kBBytes = (BN / CTAS) * kBlockK * 2;
  static constexpr int kAStage = STRIP ? kStripBytes : kABytes;
  static constexpr int kStageBytes = kAStage + (STRIP ? 3 : 1) * kBBytes;
  static constexpr int kTxBytes = (STRIP ? kStripRows * 128 : kABytes) + (STRIP ? 3 : 1) * kBBytes;   // bytes TMA reports
  // as many stages as fit in 227 KB (the TMA latency of ~3000 cycles must be covered by stages x MMA time per stage)
  static constexpr int kStagesFit = (227 * 1024 - 1024 - 256 - EXTRA) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kAcc = (BN >= 256) ? 2 : 4;   // accumulator buffers in TMEM (epilogue of tile i overlaps i+1..)
  static constexpr int kTmemCols = (kAcc * BN < 32) ? 32 : kAcc * BN;   // power of two >= 32, <= 512
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + EXTRA;
};

struct Geometry {
  int n, r, cin, ncols;      // r = OUTPUT side; the input side is r * stride
  int taps;                  // 9 (3x3), 16 (4x4) or 1 (1x1)
  int split;                 // split-bf16 operands: A has 2*cin channels (hi | lo), B has 3*cin columns per tap
  int stride, pad;           // input coordinate of tap (ky, kx) for output (h, w): (h*stride + ky - pad, w*stride + kx - pad)
  int TW, TH, NB;            // tile = NB images x TH rows x TW cols (all powers of two, product 128)
  int tiles_w, tiles_h, tiles_b, n_tiles_m, n_tiles_n;
  int lw, lh;                // log2(tiles_w), log2(tiles_h) (both are powers of two)
  int strip_base_offset;     // STRIP: 1 = put the start row's swizzle phase into the descriptor base-offset field
  int pref_boxes;            // > 0: L2-prefetch the next tile's input window with this many channel boxes (map_p)
  int pref_chan;             // channels per prefetch box
  int ksplit;                // > 1 (1x1 GEMMs with a long K, e.g. the dense layers): the channel blocks of every part are cut
                             // into ksplit ranges; range ks is a separate work unit that writes its fp32 partial sums to
                             // y + ks * M * ncols.  n_tiles_n counts (N tile, range) pairs: nt = unit % n_tiles_n_real.
  int n_tiles_n_real;
};

struct EpiParams {
  int mode;
  const float* bias;
  float* y;
  const float* res;
  int res_shift;
  float2* stat_pairs;        // TC_EPI_BIAS_F32: optional per-(M tile, warp) column (sum, sum of squares) partials
  const float* sx;
  int sx_shift;
  const float* mean;
  const float* rstd;
  int samples_per_group;
  float slope;
  int act;
  int split_out;             // TC_EPI_ACT_BF16: write hi | lo halves (row pitch 2 * ncols)
  const float* scale;        // optional per-column scale applied before the bias (folded BatchNorm): acc * scale + bias
  int out_pitch;             // TC_EPI_ACT_BF16 / PHASE_ACT: channels per output pixel in memory (0 = ncols resp. cout)
  int phase_cout;            // TC_EPI_PHASE_ACT_BF16: channels per phase (ncols = 4 * phase_cout)
  __nv_bfloat16* out_bf16;
  long long* dbg;            // optional per-CTA cycle counters (msr_debug_tc_counters): 8 values per CTA
};

// Sum over the 32 lanes of t[j] for every j; afterwards lane l holds the total of element l in t[0].  31 shuffles.
__device__ __forceinline__ float warp_transpose_reduce(float (&t)[32], int lane) {
#pragma unroll
  for (int step = 16, n = 32; step >= 1; step >>= 1, n >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? t[i] : t[i + n / 2];
      const float keep = upper ? t[i + n / 2] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return t[0];
}

// ---- the kernel -------------------------------------------------------------------------------------------------------
// KSPLIT (dense layers only) is a template parameter so that the convolution instantiations carry none of its index
// arithmetic: the single producer thread's issue rate bounds the non-strip pipeline (measured: the run-time variant cost
// the r <= 64 layers 5-10 %).
template <int BN, int CTAS, bool STRIP, int EPI, bool KSPLIT = false>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_p, const Geometry g, const EpiParams ep) {
  using C = Cfg<BN, CTAS, STRIP, (EPI == TC_EPI_RELU_BF16_T ? kTStageBytes : 0)>;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // work unit: a 128 x BN tile (CTAS == 1) or a 256 x BN tile shared by the pair (CTAS == 2)
  const int unit = (CTAS == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = (CTAS == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + C::kAcc + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + C::kStages * C::kStageBytes + 8 * (2 * C::kStages + 2 * C::kAcc));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = (g.n_tiles_m / CTAS) * g.n_tiles_n;
  const int chunks_per_part = g.cin / kBlockK;
  const int k_chunks_per_tap = (g.split ? 3 : 1) * chunks_per_part;
  const int k_chunks = KSPLIT ? g.taps * k_chunks_per_tap / g.ksplit
                              : (STRIP ? 3 * chunks_per_part : g.taps * k_chunks_per_tap);   // pipeline stages per tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < C::kAcc; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == kEpiWarps + 1) {
    if constexpr (CTAS == 2) tmem_alloc_pair(smem_u32((const void*)tmem_slot), C::kTmemCols);
    else tmem_alloc(smem_u32((const void*)tmem_slot), C::kTmemCols);
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (M tile, N tile) without divisions in the loops: every role thread walks the same sequence
  // tile = unit, unit + n_units, ... and keeps (mq, nt) = (tile / n_tiles_n, tile % n_tiles_n) incrementally
  const int step_m = n_units / g.n_tiles_n, step_n = n_units % g.n_tiles_n;
  struct TileIter {
    int mq, nt;
  };
  auto tile_first = [&]() {
    TileIter it;
    it.mq = unit / g.n_tiles_n;
    it.nt = unit % g.n_tiles_n;
    return it;
  };
  auto tile_next = [&](TileIter& it) {
    it.mq += step_m;
    it.nt += step_n;
    if (it.nt >= g.n_tiles_n) {
      it.nt -= g.n_tiles_n;
      it.mq += 1;
    }
  };
  auto decode_tile = [&](const TileIter& it, int& b0, int& h0, int& w0, int& n0) {
    const int mt = it.mq * CTAS + (int)cta_rank;
    const int tw = mt & (g.tiles_w - 1), th = (mt >> g.lw) & (g.tiles_h - 1), tb = mt >> (g.lw + g.lh);
    b0 = tb * g.NB;
    h0 = th * g.TH;
    w0 = tw * g.TW;
    n0 = (KSPLIT ? it.nt % g.n_tiles_n_real : it.nt) * BN;
  };
  auto split_of = [&](const TileIter& it) { return KSPLIT ? it.nt / g.n_tiles_n_real : 0; };

  if (warp == kEpiWarps) {
    // ===================== TMA producer =====================
    // (one thread issues both operands: a second producer warp was measured to be slower)
    if (lane == 0) {
      constexpr bool is_a = true;
      int stage = 0;
      uint32_t phase = 0;
      const bool timed = ep.dbg != nullptr && is_a;
      long long t_empty = 0, t_start = timed ? clock64() : 0;
      // (no integer divisions in the steady-state loop)
      const uint32_t lead_full0 = (CTAS == 2) ? mapa_shared(full_bar(0), 0) : 0u;
      const int n_parts = g.split ? 3 : 1;
      const int ksz = (g.taps == 9) ? 3 : (g.taps == 16 ? 4 : 1);
      TileIter it = tile_first();
      for (int tile = unit; tile < total_tiles; tile += n_units, tile_next(it)) {
        int b0, h0, w0, n0;
        decode_tile(it, b0, h0, w0, n0);
        const int nb = n0 + ((CTAS == 2) ? (int)cta_rank * (BN / 2) : 0);
        if (is_a && g.pref_boxes > 0 && tile + n_units < total_tiles) {
          TileIter nx = it;
          tile_next(nx);
          int pb0, ph0, pw0, pn0;
          decode_tile(nx, pb0, ph0, pw0, pn0);
          for (int pc = 0; pc < g.pref_boxes; ++pc)
            tma_prefetch_l2_4d(&map_p, pc * g.pref_chan, pw0 * g.stride, ph0 * g.stride - g.pad, pb0);
        }
        if constexpr (STRIP) {
          // stage = (ky, channel block): the 130-pixel strip [w0 - 1, w0 + 129) of input row h0 + ky - 1 (out-of-range
          // pixels / rows arrive as zeros = SAME padding) + the weight tiles of kx = 0, 1, 2
          for (int ky = 0; ky < 3; ++ky) {
            const int ch = h0 + ky - 1;
            for (int cb = 0; cb < chunks_per_part; ++cb) {
              t_empty += mbar_wait_timed(empty_bar(stage), phase ^ 1u, timed);
              const uint32_t sa = smem_base + stage * C::kStageBytes;
              const int kb = ky * 3 * g.cin + cb * kBlockK;
              if constexpr (CTAS == 2) {
                if (leader) mbar_expect_tx(full_bar(stage), 2 * C::kTxBytes);
                const uint32_t lead_bar = lead_full0 + 8u * stage;
                tma_load_4d_pair(sa, &map_a, lead_bar, cb * kBlockK, w0 - 1, ch, b0);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                  tma_load_2d_pair(sa + C::kAStage + kx * C::kBBytes, &map_b, lead_bar, kb + kx * g.cin, nb);
              } else {
                mbar_expect_tx(full_bar(stage), C::kTxBytes);
                tma_load_4d(sa, &map_a, full_bar(stage), cb * kBlockK, w0 - 1, ch, b0);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                  tma_load_2d(sa + C::kAStage + kx * C::kBBytes, &map_b, full_bar(stage), kb + kx * g.cin, nb);
              }
              if (++stage == C::kStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        } else {
          // split-K work unit (KSPLIT): channel blocks [cb_lo, cb_hi) of every part; otherwise all of them
          const int cb_per = KSPLIT ? chunks_per_part / g.ksplit : chunks_per_part;
          const int cb_lo = KSPLIT ? split_of(it) * cb_per : 0, cb_hi = cb_lo + cb_per;
          int kcol = 0;   // K coordinate of the weight tile
          for (int ky = 0; ky < ksz; ++ky) {
            const int ch = h0 * g.stride + ky - g.pad;
            for (int kx = 0; kx < ksz; ++kx) {
              const int cw = w0 * g.stride + kx - g.pad;
              for (int part = 0; part < n_parts; ++part) {
                // split-bf16: parts (x_hi, x_hi, x_lo) of A pair with (w_hi, w_lo, w_hi) of B
                const int a_base = (part == 2) ? g.cin : 0;
                if constexpr (KSPLIT) kcol = (((ky * ksz + kx) * n_parts + part) * chunks_per_part + cb_lo) * kBlockK;
                for (int cb = cb_lo; cb < cb_hi; ++cb, kcol += kBlockK) {
                  t_empty += mbar_wait_timed(empty_bar(stage), phase ^ 1u, timed);
                  const uint32_t sa = smem_base + stage * C::kStageBytes;
                  if constexpr (CTAS == 2) {
                    // both CTAs report their bytes to the leader's barrier; each loads its own A rows and its half of B
                    const uint32_t lead_bar = lead_full0 + 8u * stage;
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * C::kTxBytes);
                    tma_load_4d_pair(sa, &map_a, lead_bar, a_base + cb * kBlockK, cw, ch, b0);
                    tma_load_2d_pair(sa + kABytes, &map_b, lead_bar, kcol, nb);
                  } else {
                    mbar_expect_tx(full_bar(stage), C::kTxBytes);
                    tma_load_4d(sa, &map_a, full_bar(stage), a_base + cb * kBlockK, cw, ch, b0);
                    tma_load_2d(sa + kABytes, &map_b, full_bar(stage), kcol, nb);
                  }
                  if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1u;
                  }
                }
              }
            }
          }
        }
      }
      if (timed) {
        ep.dbg[blockIdx.x * 8 + 0] = t_empty;
        ep.dbg[blockIdx.x * 8 + 1] = clock64() - t_start;
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && leader) {   // in a pair only the leader CTA issues (for both)
      constexpr uint32_t idesc = make_idesc(BN, kBlockM * CTAS);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const bool timed = ep.dbg != nullptr;
      long long t_full = 0, t_tempty = 0, t_start = timed ? clock64() : 0;
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        t_tempty += mbar_wait_timed(tempty_bar(acc), acc_phase ^ 1u, timed);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kc = 0; kc < k_chunks; ++kc) {
          t_full += mbar_wait_timed(full_bar(stage), phase, timed);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          if constexpr (STRIP) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              // tap kx reads strip rows kx .. kx + 127: same buffer, start shifted by kx rows of 128 bytes (the hardware
              // swizzle XORs address bits [4,7) with [7,10), so a 128-byte-aligned shifted start needs nothing else)
              const uint64_t adesc = make_smem_desc(sa + kx * 128, g.strip_base_offset ? (uint32_t)kx : 0u);
              const uint64_t bdesc = make_smem_desc(sa + C::kAStage + kx * C::kBBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if constexpr (CTAS == 2)
                  umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                 (kc | kx | k) != 0 ? 1u : 0u);
                else
                  umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                            (kc | kx | k) != 0 ? 1u : 0u);
              }
            }
          } else {
            const uint64_t adesc = make_smem_desc(sa);
            const uint64_t bdesc = make_smem_desc(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
              if constexpr (CTAS == 2)
                umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
              else
                umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if constexpr (CTAS == 2) umma_commit_pair(empty_bar(stage));
          else umma_commit(empty_bar(stage));
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CTAS == 2) umma_commit_pair(tfull_bar(acc));
        else umma_commit(tfull_bar(acc));
        if (++acc == C::kAcc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if (timed) {
        ep.dbg[blockIdx.x * 8 + 2] = t_full;
        ep.dbg[blockIdx.x * 8 + 3] = t_tempty;
        ep.dbg[blockIdx.x * 8 + 4] = clock64() - t_start;
      }
    }
  } else {
    // ===================== epilogue warps 0..7 =====================
    // warp w reads TMEM lanes 32*(w % 4) .. +31 (hardware rule); warps w and w + 4 take alternate 32-column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3, csel = warp >> 2;
    const int row = quarter * 32 + lane;  // tile row == TMEM lane
    const int wi = row % g.TW, hi = (row / g.TW) % g.TH, bi = row / (g.TW * g.TH);
    TileIter it = tile_first();
    for (int tile = unit; tile < total_tiles; tile += n_units, tile_next(it)) {
      int b0, h0, w0, n0;
      decode_tile(it, b0, h0, w0, n0);
      const int b = b0 + bi, h = h0 + hi, w = w0 + wi;
      const bool row_ok = b < g.n;
      const int64_t m = ((int64_t)b * g.r + h) * g.r + w;
      const bool timed = ep.dbg != nullptr && threadIdx.x == 0;
      const long long tw0 = timed ? clock64() : 0;
      mbar_wait(tfull_bar(acc), acc_phase);
      if (timed) ep.dbg[blockIdx.x * 8 + 5] += clock64() - tw0;
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (EPI == TC_EPI_BIAS_F32) {
        int64_t res_row = 0;
        if (ep.res != nullptr) {
          const int rs = g.r >> ep.res_shift;
          res_row = ((int64_t)b * rs + (h >> ep.res_shift)) * rs + (w >> ep.res_shift);
        }
        const int mt = it.mq * CTAS + (int)cta_rank;
#pragma unroll 1
        for (int c0 = csel * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          const int col = n0 + c0;
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + col + j));
              o[j] += bb.x; o[j + 1] += bb.y; o[j + 2] += bb.z; o[j + 3] += bb.w;
            }
          }
          if (row_ok) {
            if (ep.res) {
              const float* rs_ptr = ep.res + res_row * g.ncols + col;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float rr[8];
                ld_global_nc_v8(rs_ptr + j, rr);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[j + q] += rr[q];
              }
            }
            // split-K work units write their partial sums to plane `ks` of y (the caller reduces the planes)
            const int64_t plane = KSPLIT ? (int64_t)split_of(it) * ((int64_t)g.n * g.r * g.r) * g.ncols : 0;
            float* dst = ep.y + plane + m * g.ncols + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              st_global_v8(dst + j, __float_as_uint(o[j]), __float_as_uint(o[j + 1]), __float_as_uint(o[j + 2]),
                           __float_as_uint(o[j + 3]), __float_as_uint(o[j + 4]), __float_as_uint(o[j + 5]),
                           __float_as_uint(o[j + 6]), __float_as_uint(o[j + 7]));
          }
          if (ep.stat_pairs != nullptr) {   // per-channel batch statistics of the output (spade.py:21), fused
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = row_ok ? o[j] : 0.f;
            const float sum = warp_transpose_reduce(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = row_ok ? o[j] * o[j] : 0.f;
            const float sq = warp_transpose_reduce(t, lane);
            ep.stat_pairs[((int64_t)mt * 4 + quarter) * g.ncols + col + lane] = make_float2(sum, sq);
          }
        }
      } else if constexpr (EPI == TC_EPI_ACT_BF16) {
        // out_bf16 = act(acc + bias (+ residual)), NHWC bf16: the A operand of a following convolution.  The optional
        // steps are whole loops behind warp-uniform branches (no predicated-off instructions in the issue stream: the
        // epilogue warps are instruction-latency bound).
        int64_t res_row = 0;
        if (ep.res != nullptr) {
          const int rs = g.r >> ep.res_shift;
          res_row = ((int64_t)b * rs + (h >> ep.res_shift)) * rs + (w >> ep.res_shift);
        }
#pragma unroll 1
        for (int c0 = csel * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          if (row_ok) {
            const int col = n0 + c0;
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
            if (ep.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 ss = __ldg(reinterpret_cast<const float4*>(ep.scale + col + j));
                o[j] *= ss.x; o[j + 1] *= ss.y; o[j + 2] *= ss.z; o[j + 3] *= ss.w;
              }
            }
            if (ep.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + col + j));
                o[j] += bb.x; o[j + 1] += bb.y; o[j + 2] += bb.z; o[j + 3] += bb.w;
              }
            }
            if (ep.res != nullptr) {
              const float* rs_ptr = ep.res + res_row * g.ncols + col;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float rr[8];
                ld_global_nc_v8(rs_ptr + j, rr);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[j + q] += rr[q];
              }
            }
            if (ep.act == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
            } else if (ep.act == ACT_LRELU) {
              const float sl = ep.slope;   // 0 < slope < 1: leaky_relu(x) = max(x, slope * x)
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], o[j] * sl);
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat162 t2 = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
              pk[j] = *reinterpret_cast<const uint32_t*>(&t2);
            }
            const int64_t pitch = ep.out_pitch > 0 ? (int64_t)ep.out_pitch
                                                   : (ep.split_out ? 2 * (int64_t)g.ncols : (int64_t)g.ncols);
            __nv_bfloat16* dst = ep.out_bf16 + m * pitch + col;
#pragma unroll
            for (int q = 0; q < 2; ++q)
              st_global_v8(dst + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4],
                           pk[8 * q + 5], pk[8 * q + 6], pk[8 * q + 7]);
            if (ep.split_out) {   // lo half of a split-bf16 operand: v - bf16(v)
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
                const float2 hf = __bfloat1622float2(h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(o[2 * j] - hf.x, o[2 * j + 1] - hf.y);
                pk[j] = *reinterpret_cast<const uint32_t*>(&l2);
              }
#pragma unroll
              for (int q = 0; q < 2; ++q)
                st_global_v8(dst + g.ncols + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4],
                             pk[8 * q + 5], pk[8 * q + 6], pk[8 * q + 7]);
            }
          }
        }
      } else if constexpr (EPI == TC_EPI_RELU_BF16_T) {
        // out_bf16 = act(acc + bias), 128 columns, written as WHOLE 256-byte rows: the TMEM layout gives a lane one row,
        // so the direct epilogue above stores 32 bytes per lane into 32 different cache lines per instruction -- for a
        // K = 64 GEMM (one MMA group per tile) that store pattern, not DRAM, bounds the kernel (ncu: L1/TEX 76 %, DRAM
        // 51 %).  Here the two warps of a lane quarter park their 32-column chunks in shared memory, then each writes
        // 16 complete rows: 2 rows = 4 full lines per instruction.  The tile's 128 rows are consecutive pixels of the
        // NHWC output (checked by the plan), so row i of the tile is pixel m0 + i.
        if constexpr (BN == 128) {
        uint8_t* stg = smem_gen + C::kStages * C::kStageBytes + 256 + quarter * (32 * kTRowPitch);
#pragma unroll 1
        for (int c0 = csel * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          const int col = n0 + c0;
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
          if (ep.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + col + j));
              o[j] += bb.x; o[j + 1] += bb.y; o[j + 2] += bb.z; o[j + 3] += bb.w;
            }
          }
          if (ep.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
          } else if (ep.act == ACT_LRELU) {
            const float sl = ep.slope;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], o[j] * sl);
          }
          uint4* dst = reinterpret_cast<uint4*>(stg + lane * kTRowPitch + c0 * 2);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 t2 = __floats2bfloat162_rn(o[8 * q + 2 * j], o[8 * q + 2 * j + 1]);
              pk[j] = *reinterpret_cast<const uint32_t*>(&t2);
            }
            dst[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        // the accumulator is in registers / shared memory now: hand it back before the global stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");     // both warps of the quarter have parked
        {
          const int64_t m0 = ((int64_t)b0 * g.r + h0) * g.r + w0;           // pixel of tile row 0
          const int rows_ok = min(128, (g.n - b0) * g.TW * g.TH);           // rows of images that exist
          const int sub = lane >> 4, cb16 = (lane & 15) * 16;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rq = csel * 16 + 2 * it + sub;                        // row inside the quarter
            const int row_t = quarter * 32 + rq;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rq * kTRowPitch + cb16);
            if (row_t < rows_ok)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(ep.out_bf16) + ((m0 + row_t) * BN + n0) * 2 + cb16) = val;
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");     // staging may be overwritten by the next tile
        }  // BN == 128
      } else if constexpr (EPI == TC_EPI_PHASE_F32) {
        // a 4x4 conv on the x2-upsampled tensor, or a 4x4 stride-2 transposed conv, as a 3x3 convolution to 4 sub-pixel
        // phases (columns 0..3 = (py, px)) + pixel shuffle:  y[b][2h + py][2w + px] = act(acc[py*2 + px] + bias[0])
        uint32_t v[32];
        if (csel == 0) {
          tmem_ld32(t_row, v);
          tmem_ld_wait();
        }
        if (csel == 0 && row_ok && n0 == 0) {
          const float b0f = ep.bias ? __ldg(ep.bias) : 0.f;
          float o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o4[q] = __uint_as_float(v[q]) + b0f;
            if (ep.act == ACT_TANH) o4[q] = tanhf(o4[q]);
          }
          const int R = 2 * g.r;
          float* dst = ep.y + ((int64_t)b * R + 2 * h) * R + 2 * w;
          *reinterpret_cast<float2*>(dst) = make_float2(o4[0], o4[1]);
          *reinterpret_cast<float2*>(dst + R) = make_float2(o4[2], o4[3]);
        }
      } else if constexpr (EPI == TC_EPI_PHASE_ACT_BF16) {
        // transposed 4x4 stride-2 convolution with cout channels as a 3x3 convolution to 4 * cout phase columns
        // (column = (py*2 + px) * cout + c); out[b][2h+py][2w+px][c] = act(acc * scale[c] + bias[c]) as bf16 into a
        // tensor of side 2r with out_pitch channels per pixel (pix2pix.py:74-86: ConvT -> BatchNorm -> ReLU)
        const int cout = ep.phase_cout;
        const int R = 2 * g.r;
        const int64_t opitch = ep.out_pitch > 0 ? ep.out_pitch : cout;
#pragma unroll 1
        for (int c0 = csel * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          if (row_ok) {
            const int col = n0 + c0;
            const int phase = col / cout, ch = col - phase * cout;   // a 32-column chunk never straddles phases (cout % 32 == 0)
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
            if (ep.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 ss = __ldg(reinterpret_cast<const float4*>(ep.scale + ch + j));
                o[j] *= ss.x; o[j + 1] *= ss.y; o[j + 2] *= ss.z; o[j + 3] *= ss.w;
              }
            }
            if (ep.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + ch + j));
                o[j] += bb.x; o[j + 1] += bb.y; o[j + 2] += bb.z; o[j + 3] += bb.w;
              }
            }
            if (ep.act == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat162 t2 = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
              pk[j] = *reinterpret_cast<const uint32_t*>(&t2);
            }
            __nv_bfloat16* dst = ep.out_bf16 + (((int64_t)b * R + 2 * h + (phase >> 1)) * R + 2 * w + (phase & 1)) * opitch + ch;
#pragma unroll
            for (int q = 0; q < 2; ++q)
              st_global_v8(dst + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4],
                           pk[8 * q + 5], pk[8 * q + 6], pk[8 * q + 7]);
          }
        }
      } else {
        // fused SPADE: a 128-column group holds gamma (64) | beta (64) of channels ch0 .. ch0+63
        if constexpr (BN >= 128) {
        const int Cc = g.ncols >> 1;
        const int rs = g.r >> ep.sx_shift;
        const int64_t x_row = ((int64_t)b * rs + (h >> ep.sx_shift)) * rs + (w >> ep.sx_shift);
        const int grp = row_ok ? b / ep.samples_per_group : 0;
#pragma unroll 1
        for (int gc = 0; gc < BN; gc += 128) {
          const int ch0 = ((n0 + gc) >> 7) * 64;
          {
            const int half = csel;
            uint32_t ga[32], be[32];
            tmem_ld32(t_row + (uint32_t)(gc + half * 32), ga);
            tmem_ld32(t_row + (uint32_t)(gc + 64 + half * 32), be);
            tmem_ld_wait();
            if (row_ok) {
              const int ch = ch0 + half * 32;
              const float* xs = ep.sx + x_row * Cc + ch;
              const float* mu = ep.mean + (int64_t)grp * Cc + ch;
              const float* rsd = ep.rstd + (int64_t)grp * Cc + ch;
              const float* bg = ep.bias + (n0 + gc) + half * 32;        // gamma bias
              const float* bb = bg + 64;                                 // beta bias
              __align__(16) __nv_bfloat16 o[32];
              float xall[32];
#pragma unroll
              for (int j = 0; j < 32; j += 8) ld_global_nc_v8(xs + j, *reinterpret_cast<float(*)[8]>(&xall[j]));
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 xv = make_float4(xall[j], xall[j + 1], xall[j + 2], xall[j + 3]);
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu + j));
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rsd + j));
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(bg + j));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bb + j));
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ma[4] = {m4.x, m4.y, m4.z, m4.w};
                const float ra[4] = {r4.x, r4.y, r4.z, r4.w}, gba[4] = {g4.x, g4.y, g4.z, g4.w};
                const float bba[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float gamma = __uint_as_float(ga[j + q]) + gba[q];
                  const float beta = __uint_as_float(be[j + q]) + bba[q];
                  float t = fmaf(gamma, (xa[q] - ma[q]) * ra[q], beta);
                  t = t > 0.f ? t : t * ep.slope;
                  o[j + q] = __float2bfloat16_rn(t);
                }
              }
              __nv_bfloat16* dst = ep.out_bf16 + m * Cc + ch;
              const uint32_t* src = reinterpret_cast<const uint32_t*>(o);
#pragma unroll
              for (int q = 0; q < 2; ++q)
                st_global_v8(dst + 16 * q, src[8 * q], src[8 * q + 1], src[8 * q + 2], src[8 * q + 3], src[8 * q + 4],
                             src[8 * q + 5], src[8 * q + 6], src[8 * q + 7]);
            }
          }
        }
        }  // BN >= 128
      }
      // release the accumulator buffer (the row-transposing epilogue has already done so)
      if constexpr (EPI != TC_EPI_RELU_BF16_T) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));   // the leader's MMA warp waits
          else mbar_arrive(tempty_bar(acc));
        }
      }
      if (++acc == C::kAcc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();   // no CTA of the pair may exit while its peer can still signal it
  else __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace tc

// ---- host side ----------------------------------------------------------------------------------------------------------
// MSR_TC_PAIRS=0 in the environment forces single-CTA tiles (A/B comparison of the two schedules)
// MSR_TC_PREFETCH=1 switches the L2 prefetch of the next tile's input window on (measured: no gain, slight loss)
static const bool g_disable_prefetch = [] {
  const char* e = getenv("MSR_TC_PREFETCH");
  return !(e != nullptr && e[0] == '1');
}();
// MSR_TC_STRIP: 0 = off, 1 (default) = on; 2 = on AND the start row's phase in the descriptor base-offset field (measured
// on B200: wrong results -- the 128B swizzle is a pure function of the shared-memory address bits, so shifted starts
// need no base offset)
static const int g_strip_mode = [] {
  const char* e = getenv("MSR_TC_STRIP");
  return e != nullptr ? atoi(e) : 1;
}();
static const bool g_disable_tstore = [] {
  const char* e = getenv("MSR_TC_TSTORE");
  return e != nullptr && e[0] == '0';
}();
static const bool g_disable_pairs = [] {
  const char* e = getenv("MSR_TC_PAIRS");
  return e != nullptr && e[0] == '0';
}();

long long* g_tc_dbg = nullptr;   // msr_debug_tc_counters

struct ConvTC {
  CUtensorMap map_a, map_b, map_p;
  tc::Geometry g;
  tc::EpiParams ep;
  int bn;
  int ctas;   // 1, or 2 = CTA pairs (cluster launch)
  int strip;  // 1 = strip mode (one 130-pixel input strip serves the three kx taps)
  int grid;
  double alg_flops;   // what the profiler counts for this launch (algorithmic, not padded)
};

int conv_tc_plan_create(ConvTC** out, const ConvTCArgs& a) {
  MSR_REQUIRE(out && a.x && a.w, "conv_tc: null operand");
  MSR_REQUIRE(a.n > 0 && a.r > 0 && (a.r & (a.r - 1)) == 0, "conv_tc: r must be a power of two");
  MSR_REQUIRE(a.cin % 64 == 0 && a.cin >= 64, "conv_tc: cin must be a multiple of 64");
  MSR_REQUIRE(a.ncols % 32 == 0 && a.ncols >= 32, "conv_tc: output columns must be a multiple of 32");
  MSR_REQUIRE(a.taps == 9 || a.taps == 1 || a.taps == 16, "conv_tc: taps must be 9 (3x3), 16 (4x4) or 1 (1x1)");
  MSR_REQUIRE(a.stride == 1 || a.stride == 2, "conv_tc: stride must be 1 or 2");
  MSR_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 127) == 0,
              "conv_tc: x must be 16-byte and w 128-byte aligned");
  MSR_REQUIRE(a.x_pitch == 0 || (a.x_pitch >= a.cin && a.x_pitch % 8 == 0), "conv_tc: bad x_pitch");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(MSR_E_CUDA, "conv_tc: cuTensorMapEncodeTiled entry point not available");
  ConvTC* p = new ConvTC();
  tc::Geometry& g = p->g;
  g.n = a.n; g.r = a.r; g.cin = a.cin; g.ncols = a.ncols;
  g.taps = a.taps; g.stride = a.stride; g.pad = a.taps == 1 ? 0 : a.pad; g.split = a.split3 ? 1 : 0;
  g.TW = std::min(a.r, 128);
  g.TH = std::min(a.r, 128 / g.TW);
  g.NB = 128 / (g.TW * g.TH);
  g.tiles_w = a.r / g.TW;
  g.tiles_h = a.r / g.TH;
  g.tiles_b = ceil_div(a.n, g.NB);
  g.n_tiles_m = g.tiles_w * g.tiles_h * g.tiles_b;
  g.lw = 0;
  while ((1 << g.lw) < g.tiles_w) ++g.lw;
  g.lh = 0;
  while ((1 << g.lh) < g.tiles_h) ++g.lh;
  p->bn = (a.ncols % 256 == 0) ? 256 : (a.ncols % 128 == 0) ? 128 : (a.ncols % 64 == 0) ? 64 : 32;
  if (a.epilogue == TC_EPI_PHASE_F32) p->bn = 32;
  g.n_tiles_n_real = a.ncols / p->bn;
  g.ksplit = a.ksplit > 1 ? a.ksplit : 1;
  g.n_tiles_n = g.n_tiles_n_real * g.ksplit;
  if (g.ksplit > 1) {
    const char* why = nullptr;
    if (a.taps != 1 || a.epilogue != TC_EPI_BIAS_F32) why = "conv_tc: split-K needs a 1x1 GEMM with the fp32 epilogue";
    else if ((a.cin / 64) % g.ksplit != 0) why = "conv_tc: ksplit must divide cin / 64";
    else if (p->bn != 256) why = "conv_tc: split-K is instantiated for 256-column tiles only";
    else if (a.bias || a.res || a.stat_pairs) why = "conv_tc: split-K planes carry no bias / residual / statistics";
    if (why) {
      delete p;
      return fail(MSR_E_INVALID, why);
    }
  }
  // CTA pairs when the layer is large enough to fill the chip with 256-row tiles
  // (and the K loop long enough to amortise the cross-CTA handshakes)
  const int k_chunks = a.taps * (a.split3 ? 3 : 1) * (a.cin / 64);
  p->ctas = (p->bn >= 128 && g.n_tiles_m % 2 == 0 && (g.n_tiles_m / 2) * g.n_tiles_n >= 74 && k_chunks >= 8 &&
             g.ksplit == 1 && !g_disable_pairs) ? 2 : 1;
  const int rin = a.r * a.stride;
  // strip mode: 3x3 stride-1 convolutions whose M tile is 128 pixels of one image row, on CTA pairs
  // (single-CTA strip tiles exist only for the final sub-pixel layer, whose 32-column tiles are A-traffic bound)
  p->strip = (g_strip_mode != 0 && (p->ctas == 2 || a.epilogue == TC_EPI_PHASE_F32) && a.taps == 9 && a.stride == 1 &&
              a.pad == 1 && !a.split3 && g.TH == 1 && g.NB == 1 && g.TW == 128) ? 1 : 0;
  g.strip_base_offset = (g_strip_mode == 2) ? 1 : 0;

  // A: 4-D NHWC tensor {C, W, H, N}; a stride-2 convolution walks W and H with element stride 2
  {
    const cuuint64_t ca = (cuuint64_t)a.cin * (a.split3 ? 2 : 1);
    const cuuint64_t cp = a.x_pitch > 0 ? (cuuint64_t)a.x_pitch : ca;   // channels per pixel in memory (>= ca: a channel
                                                                         // slice of a wider NHWC tensor, e.g. a concat buffer)
    cuuint64_t dims[4] = {ca, (cuuint64_t)rin, (cuuint64_t)rin, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {cp * 2, (cuuint64_t)rin * cp * 2, (cuuint64_t)rin * rin * cp * 2};
    cuuint32_t box[4] = {(cuuint32_t)tc::kBlockK, (cuuint32_t)(p->strip ? tc::kStripRows : g.TW * a.stride),
                         (cuuint32_t)(g.TH * a.stride), (cuuint32_t)g.NB};
    cuuint32_t estr[4] = {1, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1};
    CUresult r = enc(&p->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.x), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete p;
      return fail(MSR_E_CUDA, "conv_tc: cuTensorMapEncodeTiled(A) failed with " + std::to_string((int)r));
    }
  }
  // L2 prefetch window of one tile: all rows its taps touch x its columns x a block of channels
  g.pref_boxes = 0;
  g.pref_chan = 0;
  p->map_p = p->map_a;
  if (a.stride == 1 && !g_disable_prefetch) {
    const cuuint64_t ca = (cuuint64_t)a.cin * (a.split3 ? 2 : 1);
    const int pc = (int)std::min<cuuint64_t>(ca, 256);
    cuuint64_t dims[4] = {ca, (cuuint64_t)rin, (cuuint64_t)rin, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {ca * 2, (cuuint64_t)rin * ca * 2, (cuuint64_t)rin * rin * ca * 2};
    cuuint32_t box[4] = {(cuuint32_t)pc, (cuuint32_t)g.TW, (cuuint32_t)std::min(256, g.TH + (a.taps == 9 ? 2 : 0)),
                         (cuuint32_t)g.NB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p->map_p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.x), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      g.pref_boxes = (int)((ca + pc - 1) / pc);
      g.pref_chan = pc;
    }
  }
  // B: 2-D weights {K, N}
  {
    const cuuint64_t kb = (cuuint64_t)a.taps * a.cin * (a.split3 ? 3 : 1);
    cuuint64_t dims[2] = {kb, (cuuint64_t)a.ncols};
    cuuint64_t strides[1] = {kb * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::kBlockK, (cuuint32_t)(p->bn / p->ctas)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(a.w), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete p;
      return fail(MSR_E_CUDA, "conv_tc: cuTensorMapEncodeTiled(B) failed with " + std::to_string((int)r));
    }
  }
  tc::EpiParams& e = p->ep;
  e.mode = a.epilogue;
  // bf16 GEMM outputs of exactly one 128-column tile with a plain bias / activation epilogue (the SPADE mask
  // convolutions): whole-row stores through shared memory.  MSR_TC_TSTORE=0 keeps the direct epilogue (A/B runs).
  if (a.epilogue == TC_EPI_ACT_BF16 && p->bn == 128 && a.ncols == 128 && p->ctas == 1 && !p->strip && g.ksplit == 1 &&
      !a.res && !a.scale && !a.split_out && a.out_pitch == 0 && a.act != ACT_TANH && !g_disable_tstore &&
      (g.TW == a.r && (g.TH == a.r || g.NB == 1) || (g.TW == 128 && g.TH == 1 && g.NB == 1)))
    e.mode = TC_EPI_RELU_BF16_T;
  e.bias = a.bias; e.y = a.y; e.res = a.res; e.res_shift = a.res_shift; e.stat_pairs = a.stat_pairs;
  e.sx = a.sx; e.sx_shift = a.sx_shift; e.mean = a.mean; e.rstd = a.rstd;
  e.samples_per_group = a.samples_per_group > 0 ? a.samples_per_group : 1;
  e.slope = a.slope; e.act = a.act; e.split_out = a.split_out; e.out_bf16 = a.out_bf16;
  e.scale = a.scale; e.out_pitch = a.out_pitch; e.phase_cout = a.phase_cout;
  e.dbg = g_tc_dbg;
  const char* bad = nullptr;
  if (a.epilogue == TC_EPI_BIAS_F32) {
    if (!a.y) bad = "conv_tc: y is null";
    if (a.stat_pairs && g.NB != 1) bad = "conv_tc: fused statistics need r*r >= 128";
  } else if (a.epilogue == TC_EPI_SPADE_BF16) {
    if (!(a.sx && a.mean && a.rstd && a.out_bf16 && a.bias)) bad = "conv_tc: SPADE epilogue needs sx, mean, rstd, bias, out_bf16";
    if (a.ncols % 128 != 0) bad = "conv_tc: SPADE epilogue needs a multiple of 128 columns";
  } else if (a.epilogue == TC_EPI_ACT_BF16 || a.epilogue == TC_EPI_RELU_BF16_T) {
    if (!a.out_bf16) bad = "conv_tc: out_bf16 is null";
  } else if (a.epilogue == TC_EPI_PHASE_F32) {
    if (!a.y || a.ncols != 32) bad = "conv_tc: phase epilogue needs y and exactly 32 (padded) columns";
  } else if (a.epilogue == TC_EPI_PHASE_ACT_BF16) {
    if (!a.out_bf16 || a.phase_cout < 32 || a.phase_cout % 32 != 0 || a.ncols != 4 * a.phase_cout)
      bad = "conv_tc: phase-act epilogue needs out_bf16 and ncols = 4 * phase_cout, phase_cout a multiple of 32";
  } else {
    bad = "conv_tc: unknown epilogue";
  }
  if (bad) {
    delete p;
    return fail(MSR_E_INVALID, bad);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (p->ctas == 2) p->grid = 2 * std::min((g.n_tiles_m / 2) * g.n_tiles_n, sms / 2);
  else p->grid = std::min(g.n_tiles_m * g.n_tiles_n, sms);
  p->alg_flops = a.alg_flops > 0.0 ? a.alg_flops : 2.0 * (double)g.n * g.r * g.r * g.ncols * g.taps * g.cin;
  *out = p;
  return MSR_OK;
}

// one instantiation per (tile width, CTAs per tile, strip mode, epilogue) that the graphs actually use
template <int BN, int CTAS, bool STRIP, int EPI, bool KSPLIT = false>
static int launch_variant(const ConvTC* p, cudaStream_t st) {
  using C = tc::Cfg<BN, CTAS, STRIP, (EPI == TC_EPI_RELU_BF16_T ? tc::kTStageBytes : 0)>;
  auto kernel = tc::conv3x3_tc_kernel<BN, CTAS, STRIP, EPI, KSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    MSR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CTAS == 2 ? 1 : 0;
  MSR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, p->map_a, p->map_b, p->map_p, p->g, p->ep));
  return MSR_OK;
}

template <int BN, int EPI>
static int launch_schedule(const ConvTC* p, cudaStream_t st) {
  if constexpr (BN >= 128 && EPI != TC_EPI_PHASE_F32) {
    if (p->ctas == 2 && p->strip) return launch_variant<BN, 2, true, EPI>(p, st);
    if (p->ctas == 2) return launch_variant<BN, 2, false, EPI>(p, st);
  }
  if constexpr (EPI == TC_EPI_PHASE_F32) {
    if (p->strip) return launch_variant<BN, 1, true, EPI>(p, st);
  }
  if constexpr (EPI == TC_EPI_BIAS_F32 && BN == 256) {
    if (p->g.ksplit > 1) return launch_variant<BN, 1, false, EPI, true>(p, st);
  }
  return launch_variant<BN, 1, false, EPI>(p, st);
}

template <int EPI>
static int launch_width(const ConvTC* p, cudaStream_t st) {
  if constexpr (EPI == TC_EPI_PHASE_F32) {
    return launch_schedule<32, EPI>(p, st);
  } else if constexpr (EPI == TC_EPI_SPADE_BF16 || EPI == TC_EPI_PHASE_ACT_BF16) {
    return p->bn == 256 ? launch_schedule<256, EPI>(p, st) : launch_schedule<128, EPI>(p, st);
  } else {
    switch (p->bn) {
      case 256: return launch_schedule<256, EPI>(p, st);
      case 128: return launch_schedule<128, EPI>(p, st);
      case 64: return launch_schedule<64, EPI>(p, st);
      default: return launch_schedule<32, EPI>(p, st);
    }
  }
}

int conv_tc_launch(const ConvTC* p, cudaStream_t st) {
  MSR_REQUIRE(p, "conv_tc_launch: null plan");
  ProfileScope prof(MSR_PROF_CONV_TC, st, p->alg_flops);
  int rc;
  switch (p->ep.mode) {
    case TC_EPI_BIAS_F32: rc = launch_width<TC_EPI_BIAS_F32>(p, st); break;
    case TC_EPI_SPADE_BF16: rc = launch_width<TC_EPI_SPADE_BF16>(p, st); break;
    case TC_EPI_ACT_BF16: rc = launch_width<TC_EPI_ACT_BF16>(p, st); break;
    case TC_EPI_RELU_BF16_T: rc = launch_variant<128, 1, false, TC_EPI_RELU_BF16_T>(p, st); break;
    case TC_EPI_PHASE_ACT_BF16: rc = launch_width<TC_EPI_PHASE_ACT_BF16>(p, st); break;
    default: rc = launch_width<TC_EPI_PHASE_F32>(p, st);
  }
  if (rc) return rc;
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

// The tensor maps of a cached plan capture x and w (internal, stable buffers); every other pointer is re-read from the
// arguments at each launch, because the caller's output buffer changes from call to call.
void conv_tc_update_pointers(ConvTC* p, const ConvTCArgs& a) {
  tc::EpiParams& e = p->ep;
  e.bias = a.bias; e.y = a.y; e.res = a.res; e.stat_pairs = a.stat_pairs;
  e.sx = a.sx; e.mean = a.mean; e.rstd = a.rstd; e.out_bf16 = a.out_bf16; e.scale = a.scale;
}

void conv_tc_plan_destroy(ConvTC* p) { delete p; }

}  // namespace msr
