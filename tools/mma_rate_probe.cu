// How fast can tcgen05.mma.kind::f16 run from shared-memory operands when NOTHING else touches shared memory?
// (optimisation aid; build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I moonsuperresolution_b200/csrc
//  -o tools/bin/mma_rate_probe tools/mma_rate_probe.cu)
//
// One thread per CTA (per CTA pair) issues a long run of K = 16 MMAs on fixed, zero-filled, 128-byte-swizzled operand
// tiles -- no TMA, no epilogue, every SM busy -- and the cycles per MMA are compared with the arithmetic time
// M * N * 16 / 4096 MAC/clk/SM (64 clk for 128 x 128 per SM, 128 clk for 128 x 256 per SM).  The question behind it
// (DESIGN.md section 4.5): are the cout = 128 layers (CTA pair, N = 128) held back by the TMA fill of shared memory or
// by the operand fetch of the MMA itself?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace msr::tc;

// fill_gap > 0: a second thread streams 16 KB TMA boxes from an L2-resident buffer into a separate shared-memory ring, one
// box every `fill_gap` clocks (as fast as they come when the gap is shorter than that), while the MMAs run.
template <int CTAS, int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* cycles, const __grid_constant__ CUtensorMap map_f,
                                                         int fill_gap, long long* fill_boxes, volatile int* stop) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kA = 128 * 128;                 // A tile: 128 rows x 64 bf16
  constexpr int kB = (N / CTAS) * 128;          // this CTA's part of the B tile
  constexpr int kStages = 2;                    // rotate over 2 operand buffers
  constexpr int kFill = 8;                      // fill ring: 8 x 16 KB behind the operand buffers (TMA latency ~3000 clk)
  const uint32_t fill_base = smem_base + kStages * (kA + kB);
  const uint32_t bar = fill_base + kFill * 16384;
  auto fbar = [&](int s) { return bar + 32u + 8u * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * (kA + kB) + kFill * 16384 + 16);
  volatile int* done = stop + (blockIdx.x / CTAS);   // set by the MMA thread of the (leader) CTA, polled by the fill threads
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < kStages * (kA + kB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_gen)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    for (int s = 0; s < kFill; ++s) mbar_init(fbar(s), 1);
    fence_barrier_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    if constexpr (CTAS == 2) tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
    else tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0 && cta_rank == 0) {
    constexpr uint32_t idesc = make_idesc(N, 128 * CTAS);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = smem_base + (it & (kStages - 1)) * (kA + kB);
      const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + kA);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (CTAS == 2) umma_bf16_pair(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        else umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
      }
    }
    if constexpr (CTAS == 2) {
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                   "h"((uint16_t)1)
                   : "memory");
    } else {
      umma_commit(bar);
    }
    mbar_wait(bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
    *done = 1;
    __threadfence();
  }
  if (threadIdx.x == 32 && fill_gap > 0) {
    // the fill stream of this CTA (both CTAs of a pair fill their own shared memory, as the convolution kernel does)
    long long boxes = 0, next = clock64();
    uint32_t ph[kFill] = {};
    int s = 0;
    bool primed = false;
    while (!*done) {
      if (primed) {
        mbar_wait(fbar(s), ph[s]);
        ph[s] ^= 1u;
      }
      while (clock64() < next && !*done) {}
      next += fill_gap;
      mbar_expect_tx(fbar(s), 16384);
      tma_load_2d(fill_base + s * 16384, &map_f, fbar(s), 0, (int)((boxes * 128 + blockIdx.x * 512) & 8191));
      ++boxes;
      if (++s == kFill) {
        s = 0;
        primed = true;
      }
    }
    // drain what is still in flight before the shared memory goes away
    for (int k = 0; k < kFill; ++k) {
      const int q = (s + k) % kFill;
      if (primed || q < s) mbar_wait(fbar(q), ph[q]);
    }
    if (cta_rank == 0) fill_boxes[blockIdx.x] = boxes;
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();
  else __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

static CUtensorMap g_map_f;

template <int CTAS, int N>
static void run(const char* what, int sms, int iters, long long* d_cycles, int fill_gap = 0) {
  constexpr int smem = 2 * (128 * 128 + (N / CTAS) * 128) + 8 * 16384 + 1024 + 128;
  auto kernel = mma_rate_kernel<CTAS, N>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((sms / CTAS) * CTAS);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CTAS == 2 ? 1 : 0;
  long long* d_boxes = d_cycles + 256;
  int* d_stop = reinterpret_cast<int*>(d_cycles + 512);
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 2 && e == cudaSuccess; ++rep) {
    cudaMemset(d_cycles, 0, sizeof(long long) * 512 + sizeof(int) * 256);
    cudaLaunchKernelEx(&cfg, kernel, iters, d_cycles, g_map_f, fill_gap, d_boxes, d_stop);
    e = cudaDeviceSynchronize();
  }
  long long h[512];
  cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
  double sum = 0, boxes = 0;
  int cnt = 0;
  for (int i = 0; i < 256; ++i)
    if (h[i] > 0) {
      sum += (double)h[i];
      boxes += (double)h[256 + i];
      ++cnt;
    }
  const double cyc = sum / cnt, per = cyc / (4.0 * iters);
  const double ideal = 128.0 * N * 16 / 4096.0;   // per SM: 128 rows x N columns x K = 16 at 4096 MAC/clk
  const int opb = 128 * 32 + (N / CTAS) * 32;
  printf("%-22s fill %5.1f B/clk/SM: %6.1f clk per MMA (arithmetic %5.1f, %5.1f %% of the tensor peak; operand reads %5.1f B/clk)  %s\n",
         what, boxes / cnt * 16384.0 / cyc, per, ideal, 100.0 * ideal / per, opb / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 512 + sizeof(int) * 256);
  // fill source: 8192 rows x 128 bytes = 1 MB, L2-resident
  void* d_src;
  cudaMalloc(&d_src, 8192 * 128 + 128 * 128);
  cudaMemset(d_src, 0, 8192 * 128 + 128 * 128);
  {
    msr::EncodeTiledFn enc = msr::get_encode_fn();
    cuuint64_t dims[2] = {64, 8192 + 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&g_map_f, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      printf("tensor map encode failed\n");
      return 1;
    }
  }
  const int iters = 20000;
  run<1, 64>("one CTA  M=128 N=64", sms, iters, d_cycles);
  run<1, 128>("one CTA  M=128 N=128", sms, iters, d_cycles);
  run<1, 256>("one CTA  M=128 N=256", sms, iters, d_cycles);
  run<2, 128>("CTA pair M=256 N=128", sms, iters, d_cycles);
  run<2, 256>("CTA pair M=256 N=256", sms, iters, d_cycles);
  printf("-- with a concurrent TMA fill stream into other shared-memory buffers (one 16 KB box every `gap` clocks per CTA)\n");
  for (int gap : {1024, 512, 384, 320, 256, 224, 192, 160, 128, 64}) run<2, 128>("CTA pair M=256 N=128", sms, iters, d_cycles, gap);
  for (int gap : {1024, 512, 384, 320, 256, 224, 192, 160, 128, 64}) run<2, 256>("CTA pair M=256 N=256", sms, iters, d_cycles, gap);
  return 0;
}
