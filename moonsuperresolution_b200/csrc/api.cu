// C-ABI glue: error reporting, device queries and the single-operator entry points used by the parity tests.
#include <algorithm>
#include <vector>

#include "nn.cuh"

namespace msr {
thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;
void set_error(const std::string& s) { g_last_error = s; }
int fail(int code, const std::string& s) {
  g_last_error = s;
  return code;
}

// ---- profiler -----------------------------------------------------------------------------------------------------
bool g_profiling = false;
namespace {
struct ProfRecord {
  int family;
  cudaEvent_t e0, e1;
  double work;
  int launches;
};
std::vector<ProfRecord> g_records;
std::vector<cudaEvent_t> g_open;  // begin events of the scopes in flight (scopes nest at most a few deep)
}  // namespace
void profile_begin(int family, cudaStream_t st) {
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  g_open.push_back(e);
}
void profile_end(int family, cudaStream_t st, double work, int launches) {
  if (g_open.empty()) return;
  ProfRecord r;
  r.family = family;
  r.e0 = g_open.back();
  g_open.pop_back();
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e1, st);
  r.work = work;
  r.launches = launches;
  g_records.push_back(r);
}
static void profile_clear() {
  for (auto& r : g_records) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_records.clear();
  for (auto e : g_open) cudaEventDestroy(e);
  g_open.clear();
}
}  // namespace msr

using namespace msr;

extern "C" int msr_profile_enable(int on) {
  profile_clear();
  g_profiling = on != 0;
  return MSR_OK;
}

extern "C" int msr_profile_read(double* ms, double* work, int64_t* launches) {
  MSR_REQUIRE(ms && work && launches, "msr_profile_read: null pointer");
  MSR_CUDA_CHECK(cudaDeviceSynchronize());
  for (int f = 0; f < MSR_PROF_COUNT; ++f) {
    ms[f] = 0.0;
    work[f] = 0.0;
    launches[f] = 0;
  }
  for (auto& r : g_records) {
    float t = 0.f;
    MSR_CUDA_CHECK(cudaEventElapsedTime(&t, r.e0, r.e1));
    if (r.family >= 0 && r.family < MSR_PROF_COUNT) {
      ms[r.family] += t;
      work[r.family] += r.work;
      launches[r.family] += r.launches;
    }
  }
  return MSR_OK;
}

extern "C" int msr_version(void) { return 100; }
extern "C" const char* msr_last_error(void) { return g_last_error.c_str(); }

extern "C" int msr_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(MSR_E_CUDA, "cudaGetDevice failed");
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail(MSR_E_CUDA, "cudaDeviceGetAttribute failed");
  return sms;
}

extern "C" int msr_op_conv3x3_bf16(const uint16_t* d_x, const uint16_t* d_w, const float* d_bias, float* d_y, int n,
                                   int r, int cin, int cout, void* stream) {
  ConvTCArgs a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(d_x);
  a.w = reinterpret_cast<const __nv_bfloat16*>(d_w);
  a.n = n; a.r = r; a.cin = cin; a.ncols = cout;
  a.epilogue = TC_EPI_BIAS_F32; a.bias = d_bias; a.y = d_y;
  ConvTC* plan = nullptr;
  int rc = conv_tc_plan_create(&plan, a);
  if (rc) return rc;
  rc = conv_tc_launch(plan, (cudaStream_t)stream);
  conv_tc_plan_destroy(plan);  // the tensor maps were copied into the launch parameters
  return rc;
}

extern "C" int msr_profile_records(int family, double* ms, double* work, int64_t capacity, int64_t* count) {
  MSR_REQUIRE(ms && work && count && capacity >= 0, "msr_profile_records: bad arguments");
  MSR_CUDA_CHECK(cudaDeviceSynchronize());
  int64_t k = 0;
  for (auto& r : g_records) {
    if (r.family != family) continue;
    if (k < capacity) {
      float t = 0.f;
      MSR_CUDA_CHECK(cudaEventElapsedTime(&t, r.e0, r.e1));
      ms[k] = t;
      work[k] = r.work;
    }
    ++k;
  }
  *count = k;
  return MSR_OK;
}

extern "C" int msr_op_spade_tc(const uint16_t* d_a, const uint16_t* d_w, const float* d_bias, const float* d_x,
                               int x_shift, const float* d_mean, const float* d_rstd, int samples_per_group,
                               uint16_t* d_out, int n, int r, int C, void* stream) {
  ConvTCArgs a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(d_a);
  a.w = reinterpret_cast<const __nv_bfloat16*>(d_w);
  a.n = n; a.r = r; a.cin = 128; a.ncols = 2 * C;
  a.epilogue = TC_EPI_SPADE_BF16; a.bias = d_bias; a.sx = d_x; a.sx_shift = x_shift; a.mean = d_mean; a.rstd = d_rstd;
  a.samples_per_group = samples_per_group; a.slope = 0.2f; a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d_out);
  ConvTC* plan = nullptr;
  int rc = conv_tc_plan_create(&plan, a);
  if (rc) return rc;
  rc = conv_tc_launch(plan, (cudaStream_t)stream);
  conv_tc_plan_destroy(plan);
  return rc;
}

namespace msr { extern long long* g_tc_dbg; }
extern "C" int msr_debug_tc_counters(long long* d_counters) {
  msr::g_tc_dbg = d_counters;
  return MSR_OK;
}

extern "C" int msr_op_conv_tc(const uint16_t* d_x, const uint16_t* d_w, const float* d_bias, float* d_y_f32,
                              uint16_t* d_y_bf16, int n, int r_out, int cin, int cout, int taps, int stride, int pad,
                              int act, float slope, float* d_stat_pairs, void* stream) {
  MSR_REQUIRE((d_y_f32 != nullptr) != (d_y_bf16 != nullptr), "msr_op_conv_tc: exactly one output must be given");
  ConvTCArgs a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(d_x);
  a.w = reinterpret_cast<const __nv_bfloat16*>(d_w);
  a.n = n; a.r = r_out; a.cin = cin; a.ncols = cout; a.taps = taps; a.stride = stride; a.pad = pad;
  a.bias = d_bias;
  if (d_y_f32) {
    a.epilogue = TC_EPI_BIAS_F32; a.y = d_y_f32; a.stat_pairs = reinterpret_cast<float2*>(d_stat_pairs);
  } else {
    a.epilogue = TC_EPI_ACT_BF16; a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d_y_bf16); a.act = act; a.slope = slope;
  }
  ConvTC* plan = nullptr;
  int rc = conv_tc_plan_create(&plan, a);
  if (rc) return rc;
  rc = conv_tc_launch(plan, (cudaStream_t)stream);
  conv_tc_plan_destroy(plan);
  return rc;
}

extern "C" int msr_op_phase_tc(const uint16_t* d_x, const uint16_t* h_w4, const float* d_bias, float* d_y, int n, int r,
                               int cin, int act, void* stream) {
  MSR_REQUIRE(d_x && h_w4 && d_y, "msr_op_phase_tc: null pointer");
  MSR_REQUIRE(phase_tc_supported(r, cin), "msr_op_phase_tc: needs r in {128, 256} and cin in {64, 128}");
  std::vector<uint16_t> wg;
  PhaseTable tab;
  if (phase_tc_pack(h_w4, cin, &wg, &tab) <= 0)
    return fail(MSR_E_INVALID, "msr_op_phase_tc: no or too many non-zero (tap, phase) filters");
  void* d_wg = nullptr;
  MSR_CUDA_CHECK(cudaMalloc(&d_wg, wg.size() * 2));
  int rc = MSR_OK;
  PhaseTC* plan = nullptr;
  if (cudaMemcpy(d_wg, wg.data(), wg.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    rc = fail(MSR_E_CUDA, "msr_op_phase_tc: weight upload failed");
  } else {
    PhaseTCArgs a;
    a.x = reinterpret_cast<const __nv_bfloat16*>(d_x);
    a.wg = reinterpret_cast<const __nv_bfloat16*>(d_wg);
    a.tab = &tab; a.bias = d_bias; a.y = d_y; a.n = n; a.r = r; a.cin = cin; a.act = act;
    rc = phase_tc_plan_create(&plan, a);
    if (!rc) rc = phase_tc_launch(plan, (cudaStream_t)stream);
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = fail(MSR_E_CUDA, "msr_op_phase_tc: kernel failed");
  }
  if (plan) phase_tc_plan_destroy(plan);
  cudaFree(d_wg);
  return rc;
}

extern "C" int msr_op_mask_tc(const float* d_source, int I, const float* h_w, const float* h_bias, uint16_t* d_out, int n,
                              int r, void* stream) {
  MSR_REQUIRE(d_source && h_w && h_bias && d_out, "msr_op_mask_tc: null pointer");
  std::vector<uint16_t> wm;
  mask_tc_pack_weights(h_w, h_bias, 128, &wm);
  void* d_wm = nullptr;
  MSR_CUDA_CHECK(cudaMalloc(&d_wm, wm.size() * 2));
  int rc = MSR_OK;
  if (cudaMemcpy(d_wm, wm.data(), wm.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = fail(MSR_E_CUDA, "msr_op_mask_tc: weight upload failed");
  if (!rc) rc = mask_conv_tc(d_source, I, reinterpret_cast<const __nv_bfloat16*>(d_wm), reinterpret_cast<__nv_bfloat16*>(d_out),
                             n, r, (cudaStream_t)stream);
  if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = fail(MSR_E_CUDA, "msr_op_mask_tc: kernel failed");
  cudaFree(d_wm);
  return rc;
}

extern "C" int msr_op_enc1_tc(const float* d_source, int I, const float* h_w, uint16_t* d_out, int n, float slope,
                              void* stream) {
  MSR_REQUIRE(d_source && h_w && d_out, "msr_op_enc1_tc: null pointer");
  std::vector<uint16_t> wm;
  mask_tc_pack_weights(h_w, nullptr, 64, &wm);
  void* d_wm = nullptr;
  MSR_CUDA_CHECK(cudaMalloc(&d_wm, wm.size() * 2));
  int rc = MSR_OK;
  if (cudaMemcpy(d_wm, wm.data(), wm.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = fail(MSR_E_CUDA, "msr_op_enc1_tc: weight upload failed");
  if (!rc) rc = enc1_conv_tc(d_source, I, reinterpret_cast<const __nv_bfloat16*>(d_wm), reinterpret_cast<__nv_bfloat16*>(d_out),
                             n, slope, (cudaStream_t)stream);
  if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = fail(MSR_E_CUDA, "msr_op_enc1_tc: kernel failed");
  cudaFree(d_wm);
  return rc;
}

// ---- host-only halves of the tensor-core layers (no GPU needed): exported so that the CPU test suite covers them -------
extern "C" int msr_host_pack_mask_weights(const float* h_w, const float* h_bias, int cout, uint16_t* h_out) {
  MSR_REQUIRE(h_w && h_out && (cout == 64 || cout == 128), "msr_host_pack_mask_weights: bad arguments");
  std::vector<uint16_t> wm;
  mask_tc_pack_weights(h_w, h_bias, cout, &wm);
  std::copy(wm.begin(), wm.end(), h_out);
  return MSR_OK;
}

extern "C" int msr_host_pack_phase_weights(const uint16_t* h_w4, int cin, uint16_t* h_wg, int* kind, int* ncols) {
  MSR_REQUIRE(h_w4 && h_wg && kind && ncols && cin > 0, "msr_host_pack_phase_weights: bad arguments");
  std::vector<uint16_t> wg;
  PhaseTable tab;
  if (phase_tc_pack(h_w4, cin, &wg, &tab) <= 0) {
    *kind = -1;
    *ncols = 0;
    return MSR_OK;   // neither layer kind: the caller keeps the 9-tap form
  }
  std::copy(wg.begin(), wg.end(), h_wg);
  *kind = tab.kind;
  *ncols = tab.ncols;
  return MSR_OK;
}

extern "C" int msr_op_conv3x3_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n, int r,
                                  int cin, int cout, void* stream) {
  ConvF32 c;
  c.x = d_x; c.w = d_w; c.bias = d_bias; c.y = d_y;
  c.n = n; c.Hs = r; c.Ws = r; c.cin = cin; c.ldx = cin; c.Hv = r; c.Wv = r;
  c.Ho = r; c.Wo = r; c.cout = cout; c.ldy = cout;
  return conv_f32(c, (cudaStream_t)stream);
}
