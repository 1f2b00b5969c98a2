// C-ABI glue: error reporting, device queries and the single-operator entry points used by the parity tests.
#include <algorithm>

#include "nn.cuh"

namespace msr {
thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;
void set_error(const std::string& s) { g_last_error = s; }
int fail(int code, const std::string& s) {
  g_last_error = s;
  return code;
}
}  // namespace msr

using namespace msr;

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

extern "C" int msr_op_conv3x3_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n, int r,
                                  int cin, int cout, void* stream) {
  ConvF32 c;
  c.x = d_x; c.w = d_w; c.bias = d_bias; c.y = d_y;
  c.n = n; c.Hs = r; c.Ws = r; c.cin = cin; c.ldx = cin; c.Hv = r; c.Wv = r;
  c.Ho = r; c.Wo = r; c.cout = cout; c.ldy = cout;
  return conv_f32(c, (cudaStream_t)stream);
}
