// HBM write-bandwidth probe with NON-constant data (optimisation aid): is a write-only stream bounded near the copy
// rate's write half (~3.3 TB/s) or near the full pin rate?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void write16(uint4* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t v = (uint32_t)i * 2654435761u + seed;
    p[i] = make_uint4(v, v ^ 0x9e3779b9u, v + 12345u, v * 3u);
  }
}
// one block = one contiguous 32 KB tile at a time (the epilogue pattern of the tensor-core kernels): 256 threads x 16 B x 8
__global__ void write_tiles(uint4* p, size_t tiles, uint32_t seed) {
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    uint4* base = p + t * 2048;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t v = (uint32_t)(t * 2048 + k * 256 + threadIdx.x) * 2654435761u + seed;
      base[k * 256 + threadIdx.x] = make_uint4(v, v ^ 0x9e3779b9u, v + 12345u, v * 3u);
    }
  }
}
__global__ void read16(const uint4* p, size_t n, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}

template <typename F>
static float time_ms(F f, int reps = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)2 << 30, n = bytes / 16;
  uint4* p; uint32_t* o;
  cudaMalloc(&p, bytes); cudaMalloc(&o, 4);
  for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    float ms = time_ms([&] { write16<<<blocks, 256>>>(p, n, 1); });
    printf("write16 grid-stride  %5d blocks x 256: %.3f ms  %.0f GB/s\n", blocks, ms, bytes / ms / 1e6);
  }
  for (int blocks : {148, 148 * 2, 148 * 4, 148 * 8}) {
    float ms = time_ms([&] { write_tiles<<<blocks, 256>>>(p, bytes / 32768, 1); });
    printf("write 32 KB tiles    %5d blocks x 256: %.3f ms  %.0f GB/s\n", blocks, ms, bytes / ms / 1e6);
  }
  {
    float ms = time_ms([&] { read16<<<148 * 16, 256>>>(p, n, o); });
    printf("read16 grid-stride    %5d blocks x 256: %.3f ms  %.0f GB/s\n", 148 * 16, ms, bytes / ms / 1e6);
  }
  {
    float ms = time_ms([&] { cudaMemsetAsync(p, 0, bytes); });
    printf("cudaMemsetAsync: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
