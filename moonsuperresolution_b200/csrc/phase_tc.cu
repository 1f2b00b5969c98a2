// The one-channel sub-pixel phase layers on the tensor cores, with the 3x3 stencil taken AFTER the channel contraction.
//
//   UpSampling2D -> leaky_relu -> Conv2D(1, 4, 'same')     spade/models/networks.py:54-56   (last layer of the generator)
//   Conv2DTranspose(1, 4, strides=2, 'same') + tanh         pix2pix.py:91-95                 (last layer of the U-Net)
//
// Both are a 3x3 convolution of the low-resolution tensor x [n][r][r][cin] to the 4 sub-pixel phases (py, px) of the
// output [n][2r][2r] (generator.cu builds the phase-combined filters w4[q][tap][c]).  As an implicit GEMM with
// K = 9 * cin and 4 real output columns (TC_EPI_PHASE_F32 of conv_tc.cu) the tensor core fetches every pixel's channels
// from shared memory once per TAP: 72 k-steps per 128 pixels, each bounded by its 4 KB A-operand read whatever N is.
//
// Here the contraction over the channels is done ONCE per pixel,
//     G[p][j] = sum_c x[p][c] * w4[q_j][tap_j][c]          j = the (tap, phase) pairs with a non-zero filter (25 / 16 of 36)
// -- a 1x1 GEMM with K = cin, N = 32: 8 k-steps per 128 pixels -- and the stencil is a sum of scalars,
//     y[b][2h + py][2w + px] = act(bias + sum_{j : q_j = (py, px)} G[(h + ky_j - 1, w + kx_j - 1)][j]).
// G never goes to HBM: the eight epilogue warps move each 128 x 32 accumulator from TMEM into a shared-memory ring of
// 1024 pixels (column-major, so that neighbouring lanes touch neighbouring banks both when a lane writes its own pixel
// and when it reads its 3x3 neighbourhood), and after every 256 pixels the same 256 threads produce the output rows
// whose three input rows are complete.  A work unit is a block of RB output rows of one image; its RB + 2 input rows
// are streamed by TMA (out-of-range rows and the tensor's borders arrive as zeros = SAME padding).  The kernel is
// bound by reading x once from HBM.
#include <cstring>
#include <vector>

#include "nn.cuh"
#include "tc_ptx.cuh"

namespace msr {

namespace tc {

constexpr int kPhStages = 7;                       // 16 KB A tiles in flight
constexpr int kPhAcc = 8;                          // 32-column accumulators in TMEM (one per 128-pixel tile)
constexpr int kPhN = 32;                           // GEMM columns (MMA N)
constexpr int kPhRing = 1024;                      // pixels per ring column: 4 rows at r = 256, 8 rows at r = 128
constexpr int kPhPitch = kPhRing + 16;             // + one zero guard pixel left and right of every ring row (r + 2 per row)
constexpr int kPhABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kPhBBytes = kPhN * kBlockK * 2;      // 4 KB per 64-channel block of the weights
constexpr int kPhMaxChunks = 2;                    // cin <= 128
constexpr int kPhEpiThreads = 256;
constexpr int kPhThreads = kPhEpiThreads + 64;     // + TMA producer warp + MMA issuer warp
constexpr int kPhSmemBytes = 1024 + kPhStages * kPhABytes + kPhMaxChunks * kPhBBytes + kPhaseMaxCols * kPhPitch * 4 + 256;

// Which taps of the 3x3 low-resolution stencil feed sub-pixel phase p (per axis): KIND 0 = 4-tap convolution of the x2
// nearest-upsampled tensor with SAME padding (1, 2): phase 0 reads taps 0..2, phase 1 taps 1..2 (generator.cu);
// KIND 1 = 4-tap stride-2 transposed convolution with SAME padding: phase 0 reads taps 0..1, phase 1 taps 1..2.
// Columns of G are ordered phase-major (q = py*2 + px), then ty, then tx over the valid pairs: 25 resp. 16 columns.
__host__ __device__ constexpr bool phase_tap_valid(int kind, int p, int tp) {
  return kind == 0 ? (p == 0 || tp >= 1) : (p == 0 ? tp <= 1 : tp >= 1);
}
__host__ __device__ constexpr int phase_cols(int kind) { return kind == 0 ? 25 : 16; }

struct PhaseGeom {
  int n, r, lr;              // lr = log2(r); r is 128 or 256
  int chunks;                // cin / 64
  int RB;                    // output (low-resolution) rows per work unit
  int units_per_image, units;
  const float* bias;
  float* out;
  int act;
};

template <int KIND>
__global__ void __launch_bounds__(kPhThreads, 1)
phase_stencil_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const PhaseGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 1024-byte alignment for the swizzle atoms
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kOffB = kPhStages * kPhABytes;
  constexpr int kOffRing = kOffB + kPhMaxChunks * kPhBBytes;
  constexpr int kOffBar = kOffRing + kPhaseMaxCols * kPhPitch * 4;
  constexpr int kCols = phase_cols(KIND);
  float* ring = reinterpret_cast<float*>(smem_gen + kOffRing);
  const uint32_t bar_base = smem_base + kOffBar;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPhStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kPhStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kPhStages + kPhAcc + a); };
  const uint32_t w_bar = bar_base + 8u * (2 * kPhStages + 2 * kPhAcc);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffBar + 8 * (2 * kPhStages + 2 * kPhAcc + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPhStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kPhAcc; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);   // the four warps (TMEM lane quarters) that read one accumulator
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 9) tmem_alloc(smem_u32((const void*)tmem_slot), kPhAcc * kPhN);
  if (threadIdx.x < kPhEpiThreads)   // the guard pixels of the ring rows stay zero for the whole kernel
    for (int e = threadIdx.x; e < kCols * kPhPitch; e += kPhEpiThreads) ring[e] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int r = g.r, lr = g.lr;
  const int P = 256 >> lr;                       // input rows per group of 256 pixels (1 or 2)
  const int groups = (g.RB + 2) / P;             // groups per work unit (RB is even)
  const int ring_rows_mask = (kPhRing >> lr) - 1;
  const int rp = r + 2;                          // ring row pitch: guard | r pixels | guard

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)g.chunks * kPhBBytes);
      for (int c = 0; c < g.chunks; ++c) tma_load_2d(smem_base + kOffB + c * kPhBBytes, &map_b, w_bar, c * kBlockK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < g.units; unit += gridDim.x) {
        const int b = unit / g.units_per_image, h0 = (unit - b * g.units_per_image) * g.RB;
        for (int gi = 0; gi < groups; ++gi) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = h0 - 1 + gi * P + ((j * kBlockM) >> lr), w0 = (j * kBlockM) & (r - 1);
            for (int c = 0; c < g.chunks; ++c) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              mbar_expect_tx(full_bar(stage), kPhABytes);
              tma_load_4d(smem_base + stage * kPhABytes, &map_a, full_bar(stage), c * kBlockK, w0, row, b);
              if (++stage == kPhStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kPhN, kBlockM);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = blockIdx.x; unit < g.units; unit += gridDim.x) {
        for (int t = 0; t < 2 * groups; ++t) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kPhN);
          for (int c = 0; c < g.chunks; ++c) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(smem_base + stage * kPhABytes);
            const uint64_t bdesc = make_smem_desc(smem_base + kOffB + c * kPhBBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (c | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar(stage));   // the stage is free once these MMAs have read it
            if (++stage == kPhStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(tfull_bar(acc));
          if (++acc == kPhAcc) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue + stencil: warps 0..7 = 256 threads = the 256 pixels of a group ================
    // thread t moves pixel t of the group from TMEM to the ring (warps 0-3: first accumulator, 4-7: second; warp % 4 =
    // TMEM lane quarter) and afterwards computes the output pixel (row t >> lr of the group's output rows, column t & (r-1))
    const int t = threadIdx.x, quarter = warp & 3;
    const int prow = t >> lr, pw = t & (r - 1);
    int acc = warp >> 2;
    uint32_t acc_phase = 0;
    int grow = 0;   // ring row of the unit's first input row (h0 - 1); rows are numbered on across units
    const float bias = g.bias ? __ldg(g.bias) : 0.f;
    const int R2 = 2 * r;
    for (int unit = blockIdx.x; unit < g.units; unit += gridDim.x) {
      const int b = unit / g.units_per_image, h0 = (unit - b * g.units_per_image) * g.RB;
      for (int gi = 0; gi < groups; ++gi) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kPhN), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        acc += 2;
        if (acc >= kPhAcc) {
          acc -= kPhAcc;
          acc_phase ^= 1u;
        }
        {
          float* dst = ring + ((grow + gi * P + prow) & ring_rows_mask) * rp + 1 + pw;
#pragma unroll
          for (int j = 0; j < kCols; ++j) dst[j * kPhPitch] = __uint_as_float(v[j]);
        }
        // rows grow + gi*P .. + P-1 are complete for all 256 pixels; the rows read below were completed earlier.  One
        // barrier per group is enough: the next group writes ring rows that no output row of this group reads
        // (live window 2P + 2 rows <= ring rows).
        asm volatile("bar.sync 1, %0;" ::"n"(kPhEpiThreads) : "memory");
        const int lo = gi * P - 1 + prow;   // local input row (0 = h0 - 1) at the centre of this thread's output row
        if (lo >= 1 && lo <= g.RB) {
          // ring rows of the three taps ty (pixel pw + tx - 1 sits at + pw + tx because of the left guard pixel)
          const float* rowp[3];
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) rowp[ty] = ring + ((grow + lo + ty - 1) & ring_rows_mask) * rp + pw;
          float o[4];
          int j = 0;   // compile-time after unrolling: the column order of phase_tc_pack
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s = bias;
#pragma unroll
            for (int ty = 0; ty < 3; ++ty) {
              if (!phase_tap_valid(KIND, q >> 1, ty)) continue;
#pragma unroll
              for (int tx = 0; tx < 3; ++tx) {
                if (!phase_tap_valid(KIND, q & 1, tx)) continue;
                s += rowp[ty][j * kPhPitch + tx];
                ++j;
              }
            }
            o[q] = g.act == ACT_TANH ? tanhf(s) : s;
          }
          const int h = h0 + lo - 1;
          float* dst = g.out + ((int64_t)b * R2 + 2 * h) * R2 + 2 * pw;
          *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
          *reinterpret_cast<float2*>(dst + R2) = make_float2(o[2], o[3]);
        }
      }
      grow += g.RB + 2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kPhAcc * kPhN);
  }
}

}  // namespace tc

// ---- host side ----------------------------------------------------------------------------------------------------------
struct PhaseTC {
  CUtensorMap map_a, map_b;
  tc::PhaseGeom g;
  int kind;
  int grid;
  double alg_flops;
};

bool phase_tc_supported(int r, int cin) { return (r == 128 || r == 256) && (cin == 64 || cin == 128); }

int phase_tc_pack(const uint16_t* w4, int cin, std::vector<uint16_t>* wg, PhaseTable* tab) {
  // which (phase, tap) filters are non-zero decides the layer kind; the column order is the kernel's unrolled loop order
  const size_t K = (size_t)9 * cin;
  bool nz[4][9];
  for (int q = 0; q < 4; ++q)
    for (int tp = 0; tp < 9; ++tp) {
      const uint16_t* src = w4 + q * K + (size_t)tp * cin;
      bool any = false;
      for (int c = 0; c < cin; ++c) any = any || (src[c] & 0x7fffu) != 0;
      nz[q][tp] = any;
    }
  int kind = -1;
  for (int k = 0; k < 2 && kind < 0; ++k) {
    bool same = true;
    for (int q = 0; q < 4; ++q)
      for (int tp = 0; tp < 9; ++tp)
        same = same && nz[q][tp] == (tc::phase_tap_valid(k, q >> 1, tp / 3) && tc::phase_tap_valid(k, q & 1, tp % 3));
    if (same) kind = k;
  }
  if (kind < 0) return -1;
  wg->assign((size_t)32 * cin, 0);
  int j = 0;
  for (int q = 0; q < 4; ++q)
    for (int tp = 0; tp < 9; ++tp)
      if (nz[q][tp]) memcpy(wg->data() + (size_t)(j++) * cin, w4 + q * K + (size_t)tp * cin, (size_t)cin * 2);
  tab->kind = kind;
  tab->ncols = j;
  return j;
}

int phase_tc_plan_create(PhaseTC** out, const PhaseTCArgs& a) {
  MSR_REQUIRE(out && a.x && a.wg && a.tab && a.y, "phase_tc: null operand");
  MSR_REQUIRE(phase_tc_supported(a.r, a.cin), "phase_tc: needs r in {128, 256} and cin in {64, 128}");
  MSR_REQUIRE(a.n > 0 && (a.tab->kind == 0 || a.tab->kind == 1) && a.tab->ncols == tc::phase_cols(a.tab->kind),
              "phase_tc: bad table");
  MSR_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.wg) & 127) == 0,
              "phase_tc: x must be 16-byte and the weights 128-byte aligned");
  MSR_REQUIRE(a.x_pitch == 0 || (a.x_pitch >= a.cin && a.x_pitch % 8 == 0), "phase_tc: bad x_pitch");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(MSR_E_CUDA, "phase_tc: cuTensorMapEncodeTiled entry point not available");
  PhaseTC* p = new PhaseTC();
  tc::PhaseGeom& g = p->g;
  g.n = a.n; g.r = a.r; g.lr = a.r == 256 ? 8 : 7; g.chunks = a.cin / 64;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // rows per work unit: 32 (6 % halo rows) once that gives every SM several units, fewer for small batches
  g.RB = 32;
  while (g.RB > 2 && (int64_t)a.n * (a.r / g.RB) < 4 * (int64_t)sms) g.RB /= 2;
  g.units_per_image = a.r / g.RB;
  g.units = a.n * g.units_per_image;
  p->kind = a.tab->kind;
  g.bias = a.bias; g.out = a.y; g.act = a.act;
  {
    const cuuint64_t cp = a.x_pitch > 0 ? (cuuint64_t)a.x_pitch : (cuuint64_t)a.cin;
    cuuint64_t dims[4] = {(cuuint64_t)a.cin, (cuuint64_t)a.r, (cuuint64_t)a.r, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {cp * 2, (cuuint64_t)a.r * cp * 2, (cuuint64_t)a.r * a.r * cp * 2};
    cuuint32_t box[4] = {(cuuint32_t)tc::kBlockK, (cuuint32_t)tc::kBlockM, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = enc(&p->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.x), dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
      delete p;
      return fail(MSR_E_CUDA, "phase_tc: cuTensorMapEncodeTiled(A) failed with " + std::to_string((int)rc));
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.cin, (cuuint64_t)tc::kPhN};
    cuuint64_t strides[1] = {(cuuint64_t)a.cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::kBlockK, (cuuint32_t)tc::kPhN};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&p->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(a.wg), dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
      delete p;
      return fail(MSR_E_CUDA, "phase_tc: cuTensorMapEncodeTiled(B) failed with " + std::to_string((int)rc));
    }
  }
  p->grid = std::min(g.units, sms);
  p->alg_flops = a.alg_flops > 0.0 ? a.alg_flops : 2.0 * (double)a.n * a.r * a.r * a.cin * a.tab->ncols;
  *out = p;
  return MSR_OK;
}

template <int KIND>
static int launch_phase(const PhaseTC* p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MSR_CUDA_CHECK(cudaFuncSetAttribute(tc::phase_stencil_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        tc::kPhSmemBytes));
    attr_set = true;
  }
  tc::phase_stencil_tc_kernel<KIND><<<p->grid, tc::kPhThreads, tc::kPhSmemBytes, st>>>(p->map_a, p->map_b, p->g);
  return MSR_OK;
}

int phase_tc_launch(const PhaseTC* p, cudaStream_t st) {
  MSR_REQUIRE(p, "phase_tc_launch: null plan");
  ProfileScope prof(MSR_PROF_CONV_TC, st, p->alg_flops);
  int rc = p->kind == 0 ? launch_phase<0>(p, st) : launch_phase<1>(p, st);
  if (rc) return rc;
  count_launch();
  MSR_LAUNCH_CHECK();
  return MSR_OK;
}

void phase_tc_set_output(PhaseTC* p, float* y) { p->g.out = y; }

void phase_tc_plan_destroy(PhaseTC* p) { delete p; }

}  // namespace msr
