// Shared helpers for libmoonsr (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/moonsr.h"

namespace msr {

// thread-local last error text, surfaced through msr_last_error()
void set_error(const std::string& s);
int fail(int code, const std::string& s);

#define MSR_CUDA_CHECK(expr)                                                                             \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return ::msr::fail(MSR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                                         std::to_string(__LINE__) + ")");                                \
  } while (0)

#define MSR_LAUNCH_CHECK() MSR_CUDA_CHECK(cudaGetLastError())

#define MSR_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) return ::msr::fail(MSR_E_INVALID, (msg));       \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// launch counter (per thread) so the generator can report gpu_launches
extern thread_local int64_t g_launch_count;
static inline void count_launch(int n = 1) { g_launch_count += n; }

// ---- per-family event timing (msr_profile_enable / msr_profile_read) ---------------------------------------------
extern bool g_profiling;
void profile_begin(int family, cudaStream_t st);
void profile_end(int family, cudaStream_t st, double work, int launches);
struct ProfileScope {
  int family, launches;
  cudaStream_t st;
  double work;
  ProfileScope(int f, cudaStream_t s, double w, int n = 1) : family(f), launches(n), st(s), work(w) {
    if (g_profiling) profile_begin(family, st);
  }
  ~ProfileScope() {
    if (g_profiling) profile_end(family, st, work, launches);
  }
};

}  // namespace msr
