// Generator graphs: GauGAN / CNNSpade (encoder -> sampler -> SPADE generator) and the pix2pix U-Net, built from the
// operators of nn_fp32.cu (fp32 mode) and conv_tc.cu (bf16 tensor-core mode).
//
// Reference: spade/models/networks.py:8-57, blocks.py:9-68, spade.py:5-25, sampling.py:5-17, model.py:564-567 and
// 789-791, pix2pix.py:64-108.  Nearest x2 upsampling (networks.py:44-54) is never materialised: consumers index
// (h >> 1, w >> 1), and the batch moments of an upsampled tensor equal those of its source.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "nn.cuh"

using namespace msr;

namespace {

constexpr int kLatent = 256;
constexpr int kHidden = 128;
const int kRB[6] = {1024, 1024, 1024, 512, 256, 128};
const int kEnc[5] = {64, 128, 256, 512, 512};
const int kP2PDown[8] = {64, 128, 256, 512, 512, 512, 512, 512};
const int kP2PUp[7] = {512, 512, 512, 512, 256, 128, 64};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct ActRef {
  const void* ptr;
  int64_t count;
  int is_bf16;
};

struct SpadeW {
  const float* conv_w = nullptr;   // [18][128]
  const float* conv_b = nullptr;   // [128]
  const float* gb_w = nullptr;     // fp32 [1152][2C]
  const float* gb_b = nullptr;     // fp32 [2C] (gamma | beta)
  const __nv_bfloat16* gb_wt = nullptr;  // bf16 [2C][1152], rows interleaved per 64 channels
  const float* gb_bt = nullptr;          // fp32 [2C] in the interleaved order
  const __nv_bfloat16* conv_wt = nullptr;  // bf16 [128][64]: k = (ky*3+kx)*2 + c for k < 18, zero beyond (im2col GEMM)
  const __nv_bfloat16* conv_wm = nullptr;  // bf16 [128][64] in the K layout of mask_tc.cu (A tile built inside the kernel)
  int C = 0;
};

struct ConvW {
  const float* w = nullptr;  // [9cin][cout]
  const float* b = nullptr;
  const __nv_bfloat16* wt = nullptr;  // bf16 [cout][9cin]
  int cin = 0, cout = 0;
};

struct BlockW {
  SpadeW s1, s2, s3;
  ConvW c1, c2, c3;
  bool learned = false;
  int cin = 0, cout = 0;
};

}  // namespace

struct msr_generator {
  int arch = 0, I = 0, B = 0, maxG = 1, precision = 0;
  bool finalized = false;
  std::map<std::string, HostTensor> host;
  std::vector<void*> allocs;
  int64_t dev_bytes = 0;
  int64_t last_launches = 0;
  std::map<std::string, ActRef> acts;

  // SPADE / CNN
  const float* dense_w = nullptr; const float* dense_b = nullptr;
  BlockW rb[6];
  const float* out_w = nullptr; const float* out_b = nullptr;
  const float* enc_w[5] = {}; const float* enc_g[5] = {}; const float* enc_bt[5] = {};
  const float* enc_mean_w = nullptr; const float* enc_mean_b = nullptr;
  const float* enc_var_w = nullptr; const float* enc_var_b = nullptr;
  // bf16 mode extras
  const __nv_bfloat16* out_wt = nullptr;         // [32][9*128] sub-pixel phase weights of the final 4x4 conv
  const __nv_bfloat16* enc1_wt = nullptr;        // [64][64] im2col weights of encoder block 1
  const __nv_bfloat16* enc1_wm = nullptr;        // the same in the K layout of mask_tc.cu (operand tile built in the kernel)
  const __nv_bfloat16* enc_wt[5] = {};           // [cout][9*cin] for blocks 2..5
  const __nv_bfloat16* enc_head_wt = nullptr;    // [512 = mean | variance][3 * feat]: split-bf16 rows (w_hi | w_lo | w_hi)
  const float* enc_head_b = nullptr;             // [512]
  const __nv_bfloat16* dense_wt = nullptr;       // [sw*sw*1024][3 * 256]: generator dense layer, same row structure
  __nv_bfloat16* enc_feat_split = nullptr;       // [n][2 * feat]: flattened encoder output as hi | lo
  __nv_bfloat16* latent_split = nullptr;         // [n][2 * 256]
  int head_ksplit = 1;
  __nv_bfloat16* patches = nullptr;              // im2col'd source [n][r][r][64]
  __nv_bfloat16* enc_b0 = nullptr; __nv_bfloat16* enc_b1 = nullptr; float* enc_y = nullptr;
  float* lat_mv = nullptr;                       // [n][512] mean | variance
  float2* stat_pairs = nullptr;
  // pix2pix
  const float* pd_w[8] = {}; const float* pd_mean[8] = {}; const float* pd_rstd[8] = {}; const float* pd_g[8] = {}; const float* pd_b[8] = {};
  const float* pu_w[7] = {}; const float* pu_mean[7] = {}; const float* pu_rstd[7] = {}; const float* pu_g[7] = {}; const float* pu_b[7] = {};
  const float* pl_w = nullptr; const float* pl_b = nullptr;

  // workspace
  float* enc_buf[2] = {};      // encoder ping-pong
  float* enc_stats_mean = nullptr; float* enc_stats_rstd = nullptr;
  float* lat_mean = nullptr; float* lat_var = nullptr; float* latent = nullptr;
  float* dense_partial = nullptr; int64_t dense_partial_cap = 0;
  double* stat_partial = nullptr;
  unsigned int* stat_counters = nullptr;   // tickets of the single-launch statistics kernels (zero between launches)
  float* xbuf[2] = {};         // residual stream ping-pong (fp32)
  float* h1 = nullptr; float* s3 = nullptr;
  float* st_mean[3] = {}; float* st_rstd[3] = {};   // stats of: x_prev, h1, (spare)
  float* a_f32 = nullptr; float* gb_f32 = nullptr; float* act_f32 = nullptr;      // fp32 mode
  __nv_bfloat16* a_bf16 = nullptr; __nv_bfloat16* act_bf16 = nullptr;               // bf16 mode
  std::map<int, std::vector<ConvTC*>> plans;   // (n_groups, repeat phase) -> plans in launch order
  // last layer (one output channel, 4 sub-pixel phases) in its "contract once per pixel, then stencil" form (phase_tc.cu)
  const __nv_bfloat16* ph_wg = nullptr;        // [32][cin] columns = (tap, phase) pairs with a non-zero filter
  PhaseTable ph_tab;
  bool ph_ok = false;
  std::map<int64_t, PhaseTC*> phase_plans;     // patches per call -> plan
  // repeated-sample mode: gamma | beta of the 15 SPADE layers for the batch being repeated (allocated on first use)
  __nv_bfloat16* gb_cache[15] = {};
  int64_t gb_cache_slots = 0;                  // patches the cache was allocated for
  int gb_cache_valid_groups = 0;               // groups of the batch whose gamma | beta the cache currently holds
  // pix2pix workspace
  float* cat[7] = {}; float* d8 = nullptr; float* p2p_raw = nullptr;
  // pix2pix, bf16 tensor-core mode: weights [N][K] bf16, BatchNorm folded into per-channel scale / shift
  const __nv_bfloat16* pdt_w[8] = {}; const float* pdt_scale[8] = {}; const float* pdt_shift[8] = {};
  const __nv_bfloat16* put_w[7] = {}; const float* put_scale[7] = {}; const float* put_shift[7] = {};
  const __nv_bfloat16* plt_w = nullptr;
  __nv_bfloat16* catb[7] = {}; __nv_bfloat16* d8b = nullptr;

  ~msr_generator() {
    for (auto& kv : plans)
      for (auto* p : kv.second) conv_tc_plan_destroy(p);
    for (auto& kv : phase_plans) phase_tc_plan_destroy(kv.second);
    for (void* p : allocs) cudaFree(p);
  }
};

namespace {

int dev_alloc(msr_generator* g, void** out, int64_t bytes) {
  if (bytes <= 0) bytes = 16;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return fail(MSR_E_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  g->allocs.push_back(p);
  g->dev_bytes += bytes;
  *out = p;
  return MSR_OK;
}

template <typename T>
int upload(msr_generator* g, const std::vector<T>& h, const T** out) {
  void* p = nullptr;
  int rc = dev_alloc(g, &p, (int64_t)h.size() * sizeof(T));
  if (rc) return rc;
  MSR_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T*>(p);
  return MSR_OK;
}

int need(msr_generator* g, const std::string& name, std::vector<int64_t> shape, const HostTensor** out) {
  auto it = g->host.find(name);
  if (it == g->host.end()) return fail(MSR_E_STATE, "missing weight tensor " + name);
  if (it->second.shape != shape) return fail(MSR_E_INVALID, "weight " + name + " has the wrong shape");
  *out = &it->second;
  return MSR_OK;
}

int upload_named(msr_generator* g, const std::string& name, std::vector<int64_t> shape, const float** out) {
  const HostTensor* t;
  int rc = need(g, name, shape, &t);
  if (rc) return rc;
  return upload(g, t->data, out);
}

float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

uint16_t f2bf(float f);
uint16_t f2bf(float f) {  // round-to-nearest-even, like __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

int upload_bf16(msr_generator* g, const std::vector<uint16_t>& h, const __nv_bfloat16** out) {
  void* p = nullptr;
  int rc = dev_alloc(g, &p, (int64_t)h.size() * 2);
  if (rc) return rc;
  MSR_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const __nv_bfloat16*>(p);
  return MSR_OK;
}

// MSR_TC_PHASE=0 keeps the last layer on the 9-tap implicit GEMM of conv_tc.cu (A/B runs)
static const bool g_disable_phase_tc = [] {
  const char* e = getenv("MSR_TC_PHASE");
  return e != nullptr && e[0] == '0';
}();

// MSR_TC_MASK=0 keeps SPADE's mask convolution on the im2col buffer + K = 64 GEMM of conv_tc.cu (A/B runs)
static const bool g_disable_mask_tc = [] {
  const char* e = getenv("MSR_TC_MASK");
  return e != nullptr && e[0] == '0';
}();

// last layer: columns of the per-pixel GEMM from the [32][9 * cin] phase-combined filters (rows 0..3 real)
int pack_phase_layer(msr_generator* g, const std::vector<uint16_t>& w32, int cin) {
  g->ph_ok = false;
  if (g_disable_phase_tc || (cin != 64 && cin != 128)) return MSR_OK;
  std::vector<uint16_t> wg;
  if (phase_tc_pack(w32.data(), cin, &wg, &g->ph_tab) <= 0) return MSR_OK;   // the 9-tap form stays
  int rc = upload_bf16(g, wg, &g->ph_wg);
  if (rc) return rc;
  g->ph_ok = true;
  return MSR_OK;
}

// the last layer through phase_tc.cu when its shape allows (r = 128 / 256); false = the caller runs the 9-tap form
int run_phase_layer(msr_generator* g, const __nv_bfloat16* x, int n, int r, int cin, int x_pitch, const float* bias,
                    int act, float* out, double alg_flops, cudaStream_t st, bool* done) {
  *done = false;
  if (!g->ph_ok || !phase_tc_supported(r, cin)) return MSR_OK;
  PhaseTC*& plan = g->phase_plans[n];
  if (plan == nullptr) {
    PhaseTCArgs a;
    a.x = x; a.wg = g->ph_wg; a.tab = &g->ph_tab; a.bias = bias; a.y = out; a.n = n; a.r = r; a.cin = cin;
    a.x_pitch = x_pitch; a.act = act; a.alg_flops = alg_flops;
    int rc = phase_tc_plan_create(&plan, a);
    if (rc) {
      g->phase_plans.erase(n);
      return rc;
    }
  }
  phase_tc_set_output(plan, out);
  *done = true;
  return phase_tc_launch(plan, st);
}

int load_spade(msr_generator* g, const std::string& pre, int C, SpadeW* s) {
  s->C = C;
  int rc;
  if ((rc = upload_named(g, pre + ".conv.kernel", {3, 3, 2, kHidden}, &s->conv_w))) return rc;
  if ((rc = upload_named(g, pre + ".conv.bias", {kHidden}, &s->conv_b))) return rc;
  const HostTensor *gw, *gbias, *bw, *bbias;
  if ((rc = need(g, pre + ".conv_gamma.kernel", {3, 3, kHidden, C}, &gw))) return rc;
  if ((rc = need(g, pre + ".conv_gamma.bias", {C}, &gbias))) return rc;
  if ((rc = need(g, pre + ".conv_beta.kernel", {3, 3, kHidden, C}, &bw))) return rc;
  if ((rc = need(g, pre + ".conv_beta.bias", {C}, &bbias))) return rc;
  const int K = 9 * kHidden;
  if (g->precision == MSR_PRECISION_FP32) {
    std::vector<float> w((size_t)K * 2 * C), b(2 * C);
    for (int k = 0; k < K; ++k) {
      memcpy(&w[(size_t)k * 2 * C], &gw->data[(size_t)k * C], sizeof(float) * C);
      memcpy(&w[(size_t)k * 2 * C + C], &bw->data[(size_t)k * C], sizeof(float) * C);
    }
    memcpy(&b[0], gbias->data.data(), sizeof(float) * C);
    memcpy(&b[C], bbias->data.data(), sizeof(float) * C);
    if ((rc = upload(g, w, &s->gb_w))) return rc;
    if ((rc = upload(g, b, &s->gb_b))) return rc;
  } else {
    // rows: per block j of 64 channels, 64 gamma rows then 64 beta rows; each row K-major
    std::vector<uint16_t> wt((size_t)2 * C * K);
    std::vector<float> bt(2 * C);
    for (int c = 0; c < C; ++c) {
      const int j = c / 64, t = c % 64;
      const size_t rg = (size_t)(128 * j + t), rbeta = (size_t)(128 * j + 64 + t);
      for (int k = 0; k < K; ++k) {
        wt[rg * K + k] = f2bf(gw->data[(size_t)k * C + c]);
        wt[rbeta * K + k] = f2bf(bw->data[(size_t)k * C + c]);
      }
      bt[rg] = gbias->data[c];
      bt[rbeta] = bbias->data[c];
    }
    if ((rc = upload_bf16(g, wt, &s->gb_wt))) return rc;
    if ((rc = upload(g, bt, &s->gb_bt))) return rc;
    const HostTensor* cw;
    if ((rc = need(g, pre + ".conv.kernel", {3, 3, 2, kHidden}, &cw))) return rc;
    std::vector<uint16_t> cwt((size_t)kHidden * 64, 0);   // split-bf16: [w_hi | w_lo | w_hi] against (x_hi, x_hi, x_lo)
    for (int k = 0; k < 18; ++k)
      for (int co = 0; co < kHidden; ++co) {
        const float v = cw->data[(size_t)k * kHidden + co];
        const uint16_t hi = f2bf(v);
        cwt[(size_t)co * 64 + k] = hi;
        cwt[(size_t)co * 64 + 18 + k] = f2bf(v - bf2f(hi));
        cwt[(size_t)co * 64 + 36 + k] = hi;
      }
    if ((rc = upload_bf16(g, cwt, &s->conv_wt))) return rc;
    std::vector<uint16_t> cwm;
    const HostTensor* cb;
    if ((rc = need(g, pre + ".conv.bias", {kHidden}, &cb))) return rc;
    mask_tc_pack_weights(cw->data.data(), cb->data.data(), kHidden, &cwm);
    if ((rc = upload_bf16(g, cwm, &s->conv_wm))) return rc;
  }
  return MSR_OK;
}

int load_conv(msr_generator* g, const std::string& pre, int cin, int cout, ConvW* c) {
  c->cin = cin;
  c->cout = cout;
  const HostTensor *w, *b;
  int rc;
  if ((rc = need(g, pre + ".kernel", {3, 3, cin, cout}, &w))) return rc;
  if ((rc = need(g, pre + ".bias", {cout}, &b))) return rc;
  if ((rc = upload(g, b->data, &c->b))) return rc;
  if (g->precision == MSR_PRECISION_FP32) return upload(g, w->data, &c->w);
  const int K = 9 * cin;
  std::vector<uint16_t> wt((size_t)cout * K);
  for (int k = 0; k < K; ++k)
    for (int co = 0; co < cout; ++co) wt[(size_t)co * K + k] = f2bf(w->data[(size_t)k * cout + co]);
  return upload_bf16(g, wt, &c->wt);
}

template <typename T>
int ws(msr_generator* g, T** out, int64_t count) {
  void* p = nullptr;
  int rc = dev_alloc(g, &p, count * (int64_t)sizeof(T));
  if (rc) return rc;
  *out = reinterpret_cast<T*>(p);
  return MSR_OK;
}

// bf16 mode: repacked weights for the tensor-core encoder / dense / final layer and their workspace
int finalize_spade_bf16_extras(msr_generator* g) {
  const int I = g->I, sw = I / 64;
  const int64_t N = (int64_t)g->B * g->maxG;
  const int64_t half = (int64_t)(I / 2) * (I / 2);
  int rc;
  const HostTensor* t;
  {  // final Conv2D(1, 4, 'same') on the x2-upsampled tensor == 3x3 conv to 4 sub-pixel phases on the low-res tensor:
     // output row 2h+py reads upsampled rows 2h+py-1+ky (pad before = 1), i.e. low-res rows h + ((py-1+ky) >> 1)
    if ((rc = need(g, "gen.out.kernel", {4, 4, 128, 1}, &t))) return rc;
    std::vector<float> acc((size_t)32 * 9 * 128, 0.f);
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        for (int ky = 0; ky < 4; ++ky)
          for (int kx = 0; kx < 4; ++kx) {
            const int ty = ((py - 1 + ky) >> 1) + 1, tx = ((px - 1 + kx) >> 1) + 1;  // arithmetic shift: -1 >> 1 = -1
            for (int c = 0; c < 128; ++c)
              acc[((size_t)(py * 2 + px) * 9 + ty * 3 + tx) * 128 + c] += t->data[((size_t)ky * 4 + kx) * 128 + c];
          }
    std::vector<uint16_t> w(acc.size());
    for (size_t e = 0; e < w.size(); ++e) w[e] = f2bf(acc[e]);
    if ((rc = upload_bf16(g, w, &g->out_wt))) return rc;
    if ((rc = pack_phase_layer(g, w, 128))) return rc;
  }
  {  // encoder block 1: im2col GEMM weights [64][64]
    if ((rc = need(g, "enc.down1.kernel", {3, 3, 2, kEnc[0]}, &t))) return rc;
    std::vector<uint16_t> w((size_t)kEnc[0] * 64, 0);
    for (int k = 0; k < 18; ++k)
      for (int co = 0; co < kEnc[0]; ++co) {
        const float v = t->data[(size_t)k * kEnc[0] + co];
        const uint16_t hi = f2bf(v);
        w[(size_t)co * 64 + k] = hi;
        w[(size_t)co * 64 + 18 + k] = f2bf(v - bf2f(hi));
        w[(size_t)co * 64 + 36 + k] = hi;
      }
    if ((rc = upload_bf16(g, w, &g->enc1_wt))) return rc;
    std::vector<uint16_t> wm;
    mask_tc_pack_weights(t->data.data(), nullptr, kEnc[0], &wm);
    if ((rc = upload_bf16(g, wm, &g->enc1_wm))) return rc;
  }
  for (int k = 1; k < 5; ++k) {
    const int cin = kEnc[k - 1], cout = kEnc[k], K = 9 * cin;
    if ((rc = need(g, "enc.down" + std::to_string(k + 1) + ".kernel", {3, 3, cin, cout}, &t))) return rc;
    // split-bf16 (the encoder feeds the latent, whose error reaches every later layer): per tap [w_hi | w_lo | w_hi]
    std::vector<uint16_t> w((size_t)cout * 3 * K);
    for (int tap = 0; tap < 9; ++tap)
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co) {
          const float v = t->data[((size_t)tap * cin + ci) * cout + co];
          const uint16_t hi = f2bf(v);
          uint16_t* row = &w[(size_t)co * 3 * K + (size_t)tap * 3 * cin];
          row[ci] = hi;
          row[cin + ci] = f2bf(v - bf2f(hi));
          row[2 * cin + ci] = hi;
        }
    if ((rc = upload_bf16(g, w, &g->enc_wt[k]))) return rc;
  }
  {  // encoder heads (networks.py:31-33) merged into one GEMM [n][feat] x [feat][512 = mean | variance] on the tensor
     // cores with split-bf16 operands (x = hi + lo, w = hi + lo; products hi*hi + hi*lo + lo*hi ~ fp32): weight rows are
     // K-major (w_hi | w_lo | w_hi), 3 * feat columns
    const int64_t feat = (int64_t)(I / 32) * (I / 32) * 512;
    const HostTensor *wm, *wv, *bm, *bv;
    if ((rc = need(g, "enc.mean.kernel", {feat, kLatent}, &wm))) return rc;
    if ((rc = need(g, "enc.variance.kernel", {feat, kLatent}, &wv))) return rc;
    if ((rc = need(g, "enc.mean.bias", {kLatent}, &bm))) return rc;
    if ((rc = need(g, "enc.variance.bias", {kLatent}, &bv))) return rc;
    std::vector<uint16_t> w((size_t)2 * kLatent * 3 * feat);
    std::vector<float> b(2 * kLatent);
    for (int head = 0; head < 2; ++head) {
      const HostTensor* src = head == 0 ? wm : wv;
      for (int64_t f = 0; f < feat; ++f)
        for (int c = 0; c < kLatent; ++c) {
          const float v = src->data[(size_t)f * kLatent + c];
          const uint16_t hi = f2bf(v);
          uint16_t* row = &w[(size_t)(head * kLatent + c) * 3 * feat];
          row[f] = hi;
          row[feat + f] = f2bf(v - bf2f(hi));
          row[2 * feat + f] = hi;
        }
    }
    memcpy(&b[0], bm->data.data(), sizeof(float) * kLatent);
    memcpy(&b[kLatent], bv->data.data(), sizeof(float) * kLatent);
    if ((rc = upload_bf16(g, w, &g->enc_head_wt))) return rc;
    if ((rc = upload(g, b, &g->enc_head_b))) return rc;
    // split-K: ~128 work units; the channel blocks of a part (feat / 64) must divide evenly
    int ks = 1;
    while (ks < 64 && (feat / 64) % (2 * ks) == 0) ks *= 2;
    g->head_ksplit = ks;
  }
  {  // generator dense layer (networks.py:41): [n][256] x [256][sw*sw*1024], same split-bf16 structure
    const int64_t nout = (int64_t)16 * sw * sw * 64;
    if ((rc = need(g, "gen.dense.kernel", {kLatent, nout}, &t))) return rc;
    std::vector<uint16_t> w((size_t)nout * 3 * kLatent);
    for (int k = 0; k < kLatent; ++k)
      for (int64_t o = 0; o < nout; ++o) {
        const float v = t->data[(size_t)k * nout + o];
        const uint16_t hi = f2bf(v);
        uint16_t* row = &w[(size_t)o * 3 * kLatent];
        row[k] = hi;
        row[kLatent + k] = f2bf(v - bf2f(hi));
        row[2 * kLatent + k] = hi;
      }
    if ((rc = upload_bf16(g, w, &g->dense_wt))) return rc;
  }
  if ((rc = ws(g, &g->lat_mv, N * 2 * kLatent))) return rc;
  if ((rc = ws(g, &g->patches, N * half * 64))) return rc;
  if ((rc = ws(g, &g->enc_b0, N * half * 128))) return rc;          // hi | lo halves
  if ((rc = ws(g, &g->enc_b1, N * (half / 4) * 256))) return rc;
  if ((rc = ws(g, &g->enc_y, N * (half / 4) * 128))) return rc;
  if ((rc = ws(g, &g->enc_feat_split, N * 2 * (int64_t)(I / 32) * (I / 32) * 512))) return rc;
  if ((rc = ws(g, &g->latent_split, N * 2 * kLatent))) return rc;
  if ((rc = ws(g, &g->stat_pairs, 4 * N * half))) return rc;
  // split-K planes of the head GEMM: rows are padded to whole 128-row M tiles by the kernel's masking, not in memory
  const int64_t need_partial = std::max<int64_t>((int64_t)g->head_ksplit * N * 2 * kLatent, 296 * N * 2 * kLatent);
  if (need_partial > g->dense_partial_cap) {
    g->dense_partial_cap = need_partial;
    if ((rc = ws(g, &g->dense_partial, need_partial))) return rc;
  }
  return MSR_OK;
}

int finalize_spade(msr_generator* g) {
  const int I = g->I, sw = I / 64;
  const int64_t N = (int64_t)g->B * g->maxG;
  int rc;
  const bool fp32 = g->precision == MSR_PRECISION_FP32;
  // fp32 mode: fp32 dense weights (CUDA cores); bf16 mode: split-bf16 weights for the tensor-core GEMM (extras below)
  if (fp32 && (rc = upload_named(g, "gen.dense.kernel", {kLatent, 16 * sw * sw * 64}, &g->dense_w))) return rc;
  if ((rc = upload_named(g, "gen.dense.bias", {16 * sw * sw * 64}, &g->dense_b))) return rc;
  int cin = 1024;
  for (int k = 0; k < 6; ++k) {
    BlockW& b = g->rb[k];
    const int cout = kRB[k];
    b.cin = cin;
    b.cout = cout;
    b.learned = cin != cout;
    const std::string pre = "gen.rb" + std::to_string(k + 1);
    if ((rc = load_spade(g, pre + ".spade_1", cin, &b.s1))) return rc;
    if ((rc = load_spade(g, pre + ".spade_2", cout, &b.s2))) return rc;
    if ((rc = load_conv(g, pre + ".conv_1", cin, cout, &b.c1))) return rc;
    if ((rc = load_conv(g, pre + ".conv_2", cout, cout, &b.c2))) return rc;
    if (b.learned) {
      if ((rc = load_spade(g, pre + ".spade_3", cin, &b.s3))) return rc;
      if ((rc = load_conv(g, pre + ".conv_3", cin, cout, &b.c3))) return rc;
    }
    cin = cout;
  }
  if (fp32 && (rc = upload_named(g, "gen.out.kernel", {4, 4, 128, 1}, &g->out_w))) return rc;
  if ((rc = upload_named(g, "gen.out.bias", {1}, &g->out_b))) return rc;
  int ec = 2;
  for (int k = 0; k < 5; ++k) {
    const std::string pre = "enc.down" + std::to_string(k + 1);
    if (fp32 && (rc = upload_named(g, pre + ".kernel", {3, 3, ec, kEnc[k]}, &g->enc_w[k]))) return rc;
    if (k > 0) {
      if ((rc = upload_named(g, pre + ".in_gamma", {kEnc[k]}, &g->enc_g[k]))) return rc;
      if ((rc = upload_named(g, pre + ".in_beta", {kEnc[k]}, &g->enc_bt[k]))) return rc;
    }
    ec = kEnc[k];
  }
  const int64_t feat = (int64_t)(I / 32) * (I / 32) * 512;
  if (fp32) {
    if ((rc = upload_named(g, "enc.mean.kernel", {feat, kLatent}, &g->enc_mean_w))) return rc;
    if ((rc = upload_named(g, "enc.mean.bias", {kLatent}, &g->enc_mean_b))) return rc;
    if ((rc = upload_named(g, "enc.variance.kernel", {feat, kLatent}, &g->enc_var_w))) return rc;
    if ((rc = upload_named(g, "enc.variance.bias", {kLatent}, &g->enc_var_b))) return rc;
  }

  // ---- workspace
  const int64_t half = (int64_t)(I / 2) * (I / 2);
  // encoder: largest activation is the first block's output (I/2)^2 * 64
  if (fp32) {
    if ((rc = ws(g, &g->enc_buf[0], N * half * 64))) return rc;
    if ((rc = ws(g, &g->enc_buf[1], N * (half / 4) * 128))) return rc;
  }
  if ((rc = ws(g, &g->enc_stats_mean, N * 512))) return rc;
  if ((rc = ws(g, &g->enc_stats_rstd, N * 512))) return rc;
  if ((rc = ws(g, &g->lat_mean, N * kLatent))) return rc;
  if ((rc = ws(g, &g->lat_var, N * kLatent))) return rc;
  if ((rc = ws(g, &g->latent, N * kLatent))) return rc;
  g->dense_partial_cap = std::max<int64_t>(296 * N * kLatent, 4 * N * 16 * sw * sw * 64);
  if ((rc = ws(g, &g->dense_partial, g->dense_partial_cap))) return rc;
  if ((rc = ws(g, &g->stat_partial, (int64_t)std::max<int64_t>(N, g->maxG) * kStatSplit * 1024 * 2))) return rc;
  {
    const int64_t slots = std::max<int64_t>(N, g->maxG) * (1024 / 64);
    if ((rc = ws(g, &g->stat_counters, slots))) return rc;
    MSR_CUDA_CHECK(cudaMemset(g->stat_counters, 0, (size_t)slots * sizeof(unsigned int)));
  }
  const int64_t stream_elems = N * half * 128;  // max over blocks of r^2 * cout (and r_prev^2 * cin)
  if ((rc = ws(g, &g->xbuf[0], std::max<int64_t>(stream_elems, N * sw * sw * 1024)))) return rc;
  if ((rc = ws(g, &g->xbuf[1], std::max<int64_t>(stream_elems, N * sw * sw * 1024)))) return rc;
  if ((rc = ws(g, &g->h1, stream_elems))) return rc;
  if ((rc = ws(g, &g->s3, stream_elems))) return rc;
  for (int i = 0; i < 3; ++i) {
    if ((rc = ws(g, &g->st_mean[i], (int64_t)g->maxG * 1024))) return rc;
    if ((rc = ws(g, &g->st_rstd[i], (int64_t)g->maxG * 1024))) return rc;
  }
  if (g->precision == MSR_PRECISION_BF16) {
    if ((rc = finalize_spade_bf16_extras(g))) return rc;
  }
  if (g->precision == MSR_PRECISION_FP32) {
    if ((rc = ws(g, &g->a_f32, N * half * kHidden))) return rc;
    if ((rc = ws(g, &g->gb_f32, N * half * 512))) return rc;   // max r^2 * 2C = (I/2)^2 * 512
    if ((rc = ws(g, &g->act_f32, N * half * 256))) return rc;  // max r^2 * C
  } else {
    if ((rc = ws(g, &g->a_bf16, N * half * kHidden))) return rc;
    if ((rc = ws(g, &g->act_bf16, N * half * 256))) return rc;
  }
  return MSR_OK;
}

// pix2pix in bf16 tensor-core mode: every (transposed) convolution runs in conv_tc.cu.
//   down k (pix2pix.py:64-72):  Conv4x4 s2 SAME -> BN -> LeakyReLU(0.3) = 16-tap stride-2 implicit GEMM, BN folded into
//                               the epilogue's per-channel scale / shift; block 1 (2 input channels) is an im2col GEMM
//   up k (pix2pix.py:74-86):    ConvT4x4 s2 SAME -> BN -> ReLU = 3x3 convolution to 4 sub-pixel phases x cout columns
//                               (output row 2y+py reads input rows y-1 (ky=3), y (ky=1) for py=0; y (ky=2), y+1 (ky=0)
//                               for py=1) + pixel shuffle in the epilogue; the skip concat is the channel layout of
//                               the cat buffers
int finalize_pix2pix_bf16(msr_generator* g) {
  const int64_t N = (int64_t)g->B * g->maxG;
  int rc;
  auto fold_bn = [&](const std::string& pre, int c, const float** scale, const float** shift) -> int {
    const HostTensor *ga, *be, *mm, *mv;
    int r;
    if ((r = need(g, pre + ".bn.gamma", {c}, &ga))) return r;
    if ((r = need(g, pre + ".bn.beta", {c}, &be))) return r;
    if ((r = need(g, pre + ".bn.moving_mean", {c}, &mm))) return r;
    if ((r = need(g, pre + ".bn.moving_variance", {c}, &mv))) return r;
    std::vector<float> sc(c), sh(c);
    for (int i = 0; i < c; ++i) {
      const double rs = 1.0 / sqrt((double)mv->data[i] + 1e-3);   // Keras BN eps (App. B.7)
      sc[i] = (float)(rs * ga->data[i]);
      sh[i] = (float)(be->data[i] - mm->data[i] * rs * ga->data[i]);
    }
    if ((r = upload(g, sc, scale))) return r;
    return upload(g, sh, shift);
  };
  const HostTensor* t;
  {  // block 1: [64][64], columns [w | w] against (x_hi | x_lo), k = (ky*4 + kx)*2 + c
    if ((rc = need(g, "p2p.down1.kernel", {4, 4, 2, kP2PDown[0]}, &t))) return rc;
    std::vector<uint16_t> w((size_t)kP2PDown[0] * 64);
    for (int k = 0; k < 32; ++k)
      for (int co = 0; co < kP2PDown[0]; ++co) {
        const uint16_t v = f2bf(t->data[(size_t)k * kP2PDown[0] + co]);
        w[(size_t)co * 64 + k] = v;
        w[(size_t)co * 64 + 32 + k] = v;
      }
    if ((rc = upload_bf16(g, w, &g->pdt_w[0]))) return rc;
  }
  int cin = kP2PDown[0];
  for (int k = 1; k < 8; ++k) {
    const int cout = kP2PDown[k], K = 16 * cin;
    const std::string pre = "p2p.down" + std::to_string(k + 1);
    if ((rc = need(g, pre + ".kernel", {4, 4, cin, cout}, &t))) return rc;
    std::vector<uint16_t> w((size_t)cout * K);
    for (int kk = 0; kk < K; ++kk)
      for (int co = 0; co < cout; ++co) w[(size_t)co * K + kk] = f2bf(t->data[(size_t)kk * cout + co]);
    if ((rc = upload_bf16(g, w, &g->pdt_w[k]))) return rc;
    if ((rc = fold_bn(pre, cout, &g->pdt_scale[k], &g->pdt_shift[k]))) return rc;
    cin = cout;
  }
  // transposed convs: Keras kernel [ky][kx][cout][cin] -> phase weights [(py*2+px)*cout + co][(ty*3+tx)*cin + ci]
  std::vector<uint16_t> last_pw;
  auto phase_weights = [&](const std::string& name, int cout, int cin_, int rows_padded, const __nv_bfloat16** out,
                           std::vector<uint16_t>* keep = nullptr) -> int {
    const HostTensor* w;
    int r = need(g, name, {4, 4, cout, cin_}, &w);
    if (r) return r;
    const size_t K = (size_t)9 * cin_;
    std::vector<uint16_t> pw((size_t)rows_padded * K, 0);
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        for (int ty = 0; ty < 3; ++ty)
          for (int tx = 0; tx < 3; ++tx) {
            // input row y + ty - 1 feeds output row 2y + py through kernel row ky = py + 1 - 2 * (ty - 1)
            const int ky = py + 1 - 2 * (ty - 1), kx = px + 1 - 2 * (tx - 1);
            if (ky < 0 || ky > 3 || kx < 0 || kx > 3) continue;
            for (int co = 0; co < cout; ++co) {
              const float* src = &w->data[(((size_t)ky * 4 + kx) * cout + co) * cin_];
              uint16_t* dst = &pw[((size_t)(py * 2 + px) * cout + co) * K + (size_t)(ty * 3 + tx) * cin_];
              for (int ci = 0; ci < cin_; ++ci) dst[ci] = f2bf(src[ci]);
            }
          }
    if (keep) *keep = pw;
    return upload_bf16(g, pw, out);
  };
  for (int k = 0; k < 7; ++k) {
    const std::string pre = "p2p.up" + std::to_string(k + 1);
    if ((rc = phase_weights(pre + ".kernel", kP2PUp[k], cin, 4 * kP2PUp[k], &g->put_w[k]))) return rc;
    if ((rc = fold_bn(pre, kP2PUp[k], &g->put_scale[k], &g->put_shift[k]))) return rc;
    cin = kP2PUp[k] + kP2PDown[6 - k];
  }
  if ((rc = phase_weights("p2p.last.kernel", 1, cin, 32, &g->plt_w, &last_pw))) return rc;
  if ((rc = pack_phase_layer(g, last_pw, cin))) return rc;
  if ((rc = upload_named(g, "p2p.last.bias", {1}, &g->pl_b))) return rc;
  // cat[k] (k = 0..6) holds [up_{k+1} | down_{7-k}] at spatial side 2^(k+1), bf16
  for (int k = 0; k < 7; ++k) {
    const int s = 2 << k;
    if ((rc = ws(g, &g->catb[k], N * s * s * (kP2PUp[k] + kP2PDown[6 - k])))) return rc;
  }
  if ((rc = ws(g, &g->d8b, N * 512))) return rc;
  if ((rc = ws(g, &g->patches, N * 128 * 128 * 64))) return rc;
  return MSR_OK;
}

int finalize_pix2pix(msr_generator* g) {
  if (g->precision == MSR_PRECISION_BF16) return finalize_pix2pix_bf16(g);
  const int64_t N = (int64_t)g->B * g->maxG;
  int rc;
  auto bn = [&](const std::string& pre, int c, const float** mean, const float** rstd, const float** gamma,
                const float** beta) -> int {
    const HostTensor *mm, *mv;
    int r;
    if ((r = upload_named(g, pre + ".bn.gamma", {c}, gamma))) return r;
    if ((r = upload_named(g, pre + ".bn.beta", {c}, beta))) return r;
    if ((r = need(g, pre + ".bn.moving_mean", {c}, &mm))) return r;
    if ((r = need(g, pre + ".bn.moving_variance", {c}, &mv))) return r;
    std::vector<float> rs(c);
    for (int i = 0; i < c; ++i) rs[i] = (float)(1.0 / sqrt((double)mv->data[i] + 1e-3));  // Keras BN eps (App. B.7)
    if ((r = upload(g, mm->data, mean))) return r;
    return upload(g, rs, rstd);
  };
  int cin = 2;
  for (int k = 0; k < 8; ++k) {
    const std::string pre = "p2p.down" + std::to_string(k + 1);
    if ((rc = upload_named(g, pre + ".kernel", {4, 4, cin, kP2PDown[k]}, &g->pd_w[k]))) return rc;
    if (k > 0 && (rc = bn(pre, kP2PDown[k], &g->pd_mean[k], &g->pd_rstd[k], &g->pd_g[k], &g->pd_b[k]))) return rc;
    cin = kP2PDown[k];
  }
  auto load_convT = [&](const std::string& name, int cout, int cin_, const float** out) -> int {
    const HostTensor* w;
    int r = need(g, name, {4, 4, cout, cin_}, &w);
    if (r) return r;
    std::vector<float> t((size_t)16 * cin_ * cout);
    for (int tap = 0; tap < 16; ++tap)
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin_; ++ci)
          t[((size_t)tap * cin_ + ci) * cout + co] = w->data[((size_t)tap * cout + co) * cin_ + ci];
    return upload(g, t, out);
  };
  for (int k = 0; k < 7; ++k) {
    const std::string pre = "p2p.up" + std::to_string(k + 1);
    if ((rc = load_convT(pre + ".kernel", kP2PUp[k], cin, &g->pu_w[k]))) return rc;
    if ((rc = bn(pre, kP2PUp[k], &g->pu_mean[k], &g->pu_rstd[k], &g->pu_g[k], &g->pu_b[k]))) return rc;
    cin = kP2PUp[k] + kP2PDown[6 - k];
  }
  if ((rc = load_convT("p2p.last.kernel", 1, cin, &g->pl_w))) return rc;
  if ((rc = upload_named(g, "p2p.last.bias", {1}, &g->pl_b))) return rc;
  // concat buffers: cat[k] (k = 0..6) holds [up_{k+1} | down_{7-k}] at spatial 2^(k+1)
  for (int k = 0; k < 7; ++k) {
    const int s = 2 << k;
    if ((rc = ws(g, &g->cat[k], N * s * s * (kP2PUp[k] + kP2PDown[6 - k])))) return rc;
  }
  if ((rc = ws(g, &g->d8, N * 512))) return rc;
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
struct Fwd {
  msr_generator* g;
  cudaStream_t st;
  int groups;
  int64_t N;
  int phase = 0;        // MSR_REPEAT_*: 0 plain forward, 1 first generation of a repeated batch (fills the gamma | beta
                        // cache), 2 further generation (encoder, mask convs and gamma | beta convs are reused)
  int spade_index = 0;  // running index of the SPADE layer (cache slot)
  const float* source = nullptr;   // the call's input patches [N][I][I][2] (read by the mask convolutions)
  std::vector<ConvTC*>* plans = nullptr;
  size_t plan_cursor = 0;
  bool building = false;
};

int tc_conv(Fwd& f, const ConvTCArgs& a) {
  if (f.building) {
    ConvTC* p = nullptr;
    int rc = conv_tc_plan_create(&p, a);
    if (rc) return rc;
    f.plans->push_back(p);
  }
  MSR_REQUIRE(f.plan_cursor < f.plans->size(), "internal: tensor-core plan list out of sync");
  ConvTC* p = (*f.plans)[f.plan_cursor++];
  if (!f.building) conv_tc_update_pointers(p, a);   // e.g. the caller's output buffer differs from call to call
  return conv_tc_launch(p, f.st);
}

// SPADE + LeakyReLU(0.2): result in g->act_f32 (fp32 mode) or g->act_bf16 (bf16 mode)
int run_spade(Fwd& f, const SpadeW& s, const float* source, const float* x, int x_shift, const float* mean,
              const float* rstd, int r) {
  msr_generator* g = f.g;
  const int n = (int)f.N, I = g->I;
  int rc;
  if (g->precision == MSR_PRECISION_FP32) {
    ConvF32 c;
    c.x = source; c.w = s.conv_w; c.bias = s.conv_b; c.y = g->a_f32;
    c.n = n; c.Hs = I; c.Ws = I; c.cin = 2; c.ldx = 2; c.Hv = r; c.Wv = r;
    c.in_mul = I / r; c.in_add = (I / r) / 2; c.in_shift = 0;     // tf.image.resize nearest, half-pixel centres (App. B.3)
    c.Ho = r; c.Wo = r; c.cout = kHidden; c.ldy = kHidden; c.act = ACT_RELU;
    if ((rc = conv_f32(c, f.st))) return rc;
    ConvF32 d;
    d.x = g->a_f32; d.w = s.gb_w; d.bias = s.gb_b; d.y = g->gb_f32;
    d.n = n; d.Hs = r; d.Ws = r; d.cin = kHidden; d.ldx = kHidden; d.Hv = r; d.Wv = r;
    d.Ho = r; d.Wo = r; d.cout = 2 * s.C; d.ldy = 2 * s.C;
    if ((rc = conv_f32(d, f.st))) return rc;
    return spade_modulate_f32(g->gb_f32, x, x_shift, mean, rstd, g->act_f32, n, r, s.C, g->B, 0.2f, f.st);
  }
  return fail(MSR_E_STATE, "internal: run_spade is the fp32-mode operator");
}

// main 3x3 conv on the SPADE output (act buffer) -> y (fp32), optional residual
int run_conv(Fwd& f, const ConvW& w, int r, float* y, const float* res, int res_shift) {
  msr_generator* g = f.g;
  const int n = (int)f.N;
  if (g->precision == MSR_PRECISION_FP32) {
    ConvF32 c;
    c.x = g->act_f32; c.w = w.w; c.bias = w.b; c.y = y;
    c.n = n; c.Hs = r; c.Ws = r; c.cin = w.cin; c.ldx = w.cin; c.Hv = r; c.Wv = r;
    c.Ho = r; c.Wo = r; c.cout = w.cout; c.ldy = w.cout;
    c.res = res; c.res_shift = res_shift; c.ldres = w.cout;
    return conv_f32(c, f.st);
  }
  return fail(MSR_E_STATE, "internal: run_conv is the fp32-mode operator");
}

int forward_spade(Fwd& f, const float* source, const float* eps, float* out) {
  msr_generator* g = f.g;
  const int I = g->I, sw = I / 64, n = (int)f.N;
  cudaStream_t st = f.st;
  int rc;
  // ---- encoder (networks.py:8-34): conv3x3 s2, SAME pad (0, 1) on even inputs (App. B.2)
  const float* ex = source;
  int ecin = 2, er = I;
  for (int k = 0; k < 5; ++k) {
    float* ey = g->enc_buf[k & 1];
    ConvF32 c;
    c.x = ex; c.w = g->enc_w[k]; c.y = ey;
    c.n = n; c.Hs = er; c.Ws = er; c.cin = ecin; c.ldx = ecin; c.Hv = er; c.Wv = er;
    c.Ho = er / 2; c.Wo = er / 2; c.cout = kEnc[k]; c.ldy = kEnc[k];
    c.stride = 2; c.pad_t = 0; c.pad_l = 0;
    if (k == 0) { c.act = ACT_LRELU; c.act_slope = 0.2f; }
    if ((rc = conv_f32(c, st))) return rc;
    er /= 2;
    if (k > 0) {  // tfa InstanceNormalization (eps 1e-3, per sample and channel) + LeakyReLU(0.2), blocks.py:62-65
      const int64_t rows = (int64_t)er * er;
      if ((rc = channel_stats_f32(ey, kEnc[k], n, rows, kEnc[k], 1e-3f, g->stat_partial, g->enc_stats_mean,
                                  g->enc_stats_rstd, st, g->stat_counters))) return rc;
      if ((rc = affine_act_f32(ey, kEnc[k], g->enc_stats_mean, g->enc_stats_rstd, g->enc_g[k], g->enc_bt[k], ey,
                               kEnc[k], (int64_t)n * rows, kEnc[k], rows, ACT_LRELU, 0.2f, st))) return rc;
    }
    ex = ey;
    ecin = kEnc[k];
  }
  const int feat = er * er * 512;
  if ((rc = dense_f32w(ex, g->enc_mean_w, g->enc_mean_b, g->lat_mean, n, feat, kLatent, g->dense_partial,
                       g->dense_partial_cap, st))) return rc;
  if ((rc = dense_f32w(ex, g->enc_var_w, g->enc_var_b, g->lat_var, n, feat, kLatent, g->dense_partial,
                       g->dense_partial_cap, st))) return rc;
  // ---- sampler (sampling.py:11-17) or mean + variance (model.py:789-791)
  if ((rc = sampler_f32(g->lat_mean, g->lat_var, g->arch == MSR_ARCH_SPADE ? eps : nullptr, g->latent,
                        (int64_t)n * kLatent, st))) return rc;
  g->acts["enc.mean"] = {g->lat_mean, (int64_t)n * kLatent, 0};
  g->acts["enc.variance"] = {g->lat_var, (int64_t)n * kLatent, 0};
  g->acts["latent"] = {g->latent, (int64_t)n * kLatent, 0};
  // ---- generator (networks.py:37-57)
  float* x = g->xbuf[0];
  if ((rc = dense_f32w(g->latent, g->dense_w, g->dense_b, x, n, kLatent, sw * sw * 1024, g->dense_partial,
                       g->dense_partial_cap, st))) return rc;
  g->acts["x0"] = {x, (int64_t)n * sw * sw * 1024, 0};
  int r = sw, x_shift = 0;
  // statistics of the block input (shared by spade_1 and spade_3; invariant under nearest upsampling)
  if ((rc = channel_stats_f32(x, 1024, f.groups, (int64_t)g->B * sw * sw, 1024, 1e-5f, g->stat_partial, g->st_mean[0],
                              g->st_rstd[0], st, g->stat_counters))) return rc;
  for (int k = 0; k < 6; ++k) {
    const BlockW& b = g->rb[k];
    float* y = g->xbuf[(k + 1) & 1];
    const int64_t rows = (int64_t)g->B * r * r;
    // x = conv_1(lrelu(spade_1(in)))                                  blocks.py:29-30
    if ((rc = run_spade(f, b.s1, source, x, x_shift, g->st_mean[0], g->st_rstd[0], r))) return rc;
    if ((rc = run_conv(f, b.c1, r, g->h1, nullptr, 0))) return rc;
    if ((rc = channel_stats_f32(g->h1, b.cout, f.groups, rows, b.cout, 1e-5f, g->stat_partial, g->st_mean[1],
                                g->st_rstd[1], st, g->stat_counters))) return rc;
    const float* res = x;
    int res_shift = x_shift;
    if (b.learned) {  // skip = conv_3(lrelu(spade_3(in)))             blocks.py:34-36
      if ((rc = run_spade(f, b.s3, source, x, x_shift, g->st_mean[0], g->st_rstd[0], r))) return rc;
      if ((rc = run_conv(f, b.c3, r, g->s3, nullptr, 0))) return rc;
      res = g->s3;
      res_shift = 0;
    }
    // x = conv_2(lrelu(spade_2(x))); out = skip + x                   blocks.py:31-32,38
    if ((rc = run_spade(f, b.s2, source, g->h1, 0, g->st_mean[1], g->st_rstd[1], r))) return rc;
    if ((rc = run_conv(f, b.c2, r, y, res, res_shift))) return rc;
    g->acts["rb" + std::to_string(k + 1) + ".h1"] = {g->h1, (int64_t)n * r * r * b.cout, 0};
    g->acts["rb" + std::to_string(k + 1) + ".out"] = {y, (int64_t)n * r * r * b.cout, 0};
    if (k < 5) {
      if ((rc = channel_stats_f32(y, b.cout, f.groups, rows, b.cout, 1e-5f, g->stat_partial, g->st_mean[0],
                                  g->st_rstd[0], st, g->stat_counters))) return rc;
    }
    x = y;
    x_shift = 1;  // UpSampling2D((2, 2)) after every block (networks.py:44-54), fused into the consumers
    r *= 2;
  }
  // leaky_relu(0.2) + Conv2D(1, 4, 'same') on the upsampled rb6 output (networks.py:54-56)
  if ((rc = final_conv_f32(x, g->out_w, g->out_b, out, n, r / 2, st))) return rc;
  g->acts["out"] = {out, (int64_t)n * I * I, 0};
  return MSR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// bf16 tensor-core mode: every convolution and the im2col'd 2-channel convolutions run in conv_tc.cu
// ------------------------------------------------------------------------------------------------------------------
// a = relu(conv3x3(resized mask)) (spade.py:17-18) -> g->a_bf16: mask_tc.cu builds the operand tile from the source inside
// the kernel; the older form is a K = 64 GEMM on the im2col rows that forward_spade_bf16 writes once per resolution
bool mask_in_kernel(const msr_generator* g, int r) { return !g_disable_mask_tc && mask_tc_supported(g->I, r); }

int mask_bf16(Fwd& f, const SpadeW& s, int r) {
  msr_generator* g = f.g;
  const int n = (int)f.N;
  if (mask_in_kernel(g, r)) return mask_conv_tc(f.source, g->I, s.conv_wm, g->a_bf16, n, r, f.st);
  ConvTCArgs m;
  m.x = g->patches; m.w = s.conv_wt; m.n = n; m.r = r; m.cin = 64; m.ncols = kHidden; m.taps = 1; m.pad = 0;
  m.epilogue = TC_EPI_ACT_BF16; m.bias = s.conv_b; m.act = ACT_RELU; m.out_bf16 = g->a_bf16;
  m.alg_flops = 2.0 * (double)n * r * r * kHidden * 18;          // 3x3 taps x 2 channels (the GEMM's K is padded to 64)
  return tc_conv(f, m);
}

int spade_bf16(Fwd& f, const SpadeW& s, const float* x, int x_shift, const float* mean, const float* rstd, int r) {
  msr_generator* g = f.g;
  const int n = (int)f.N;
  int rc;
  const int layer = f.spade_index++;
  if (f.phase != MSR_REPEAT_NONE) {
    // repeated-sample mode: gamma | beta (spade.py:19-20) depend only on the source -> computed once per batch
    // (phase 1, raw bf16 columns into the cache), every generation modulates from the cache (spade.py:21-24)
    __nv_bfloat16* cache = g->gb_cache[layer];
    MSR_REQUIRE(cache != nullptr, "internal: gamma | beta cache not allocated");
    if (f.phase == MSR_REPEAT_FIRST) {
      if ((rc = mask_bf16(f, s, r))) return rc;
      ConvTCArgs a;
      a.x = g->a_bf16; a.w = s.gb_wt; a.n = n; a.r = r; a.cin = kHidden; a.ncols = 2 * s.C;
      a.epilogue = TC_EPI_ACT_BF16; a.bias = s.gb_bt; a.act = ACT_NONE; a.out_bf16 = cache;
      if ((rc = tc_conv(f, a))) return rc;
    }
    return spade_modulate_cached_bf16(cache, x, x_shift, mean, rstd, g->act_bf16, n, r, s.C, g->B, 0.2f, f.st);
  }
  if ((rc = mask_bf16(f, s, r))) return rc;
  // gamma | beta convolution with the fused normalise - modulate - LeakyReLU epilogue (spade.py:19-24, blocks.py:30)
  ConvTCArgs a;
  a.x = g->a_bf16; a.w = s.gb_wt; a.n = n; a.r = r; a.cin = kHidden; a.ncols = 2 * s.C;
  a.epilogue = TC_EPI_SPADE_BF16; a.bias = s.gb_bt; a.sx = x; a.sx_shift = x_shift; a.mean = mean; a.rstd = rstd;
  a.samples_per_group = g->B; a.slope = 0.2f; a.out_bf16 = g->act_bf16;
  return tc_conv(f, a);
}

// main conv on act_bf16 -> y fp32 (+ residual) and the batch statistics of y into (mean, rstd) when requested
int conv_bf16(Fwd& f, const ConvW& w, int r, float* y, const float* res, int res_shift, float* mean, float* rstd) {
  msr_generator* g = f.g;
  const int n = (int)f.N;
  int rc;
  const bool fused = mean != nullptr && r * r >= 128;
  ConvTCArgs a;
  a.x = g->act_bf16; a.w = w.wt; a.n = n; a.r = r; a.cin = w.cin; a.ncols = w.cout;
  a.epilogue = TC_EPI_BIAS_F32; a.bias = w.b; a.y = y; a.res = res; a.res_shift = res_shift;
  a.stat_pairs = fused ? g->stat_pairs : nullptr;
  if ((rc = tc_conv(f, a))) return rc;
  if (mean == nullptr) return MSR_OK;
  if (fused) {
    const int64_t rows_p = (int64_t)g->B * (r * r / 128) * 4;
    return channel_stats_from_pairs(g->stat_pairs, f.groups, rows_p, (int64_t)g->B * r * r, w.cout, 1e-5f,
                                    g->stat_partial, mean, rstd, f.st, g->stat_counters);
  }
  return channel_stats_f32(y, w.cout, f.groups, (int64_t)g->B * r * r, w.cout, 1e-5f, g->stat_partial, mean, rstd, f.st,
                           g->stat_counters);
}

int forward_spade_bf16(Fwd& f, const float* source, const float* eps, float* out) {
  msr_generator* g = f.g;
  const int I = g->I, sw = I / 64, n = (int)f.N;
  cudaStream_t st = f.st;
  int rc;
  f.spade_index = 0;
  f.source = source;
  const bool reuse = f.phase == MSR_REPEAT_NEXT;   // a further generation of the same batch: only the noise is new
  // ---- encoder (networks.py:8-34); its mean | variance rows (lat_mv) are kept across the generations of a batch
  if (!reuse) {
  // block 1: conv3x3 s2 (2 -> 64, no bias, no norm) + LeakyReLU(0.2) as an im2col GEMM
  if (mask_in_kernel(g, I / 2)) {
    if ((rc = enc1_conv_tc(source, I, g->enc1_wm, g->enc_b0, n, 0.2f, st))) return rc;
  } else {
    if ((rc = source_patches_bf16(source, I, g->patches, n, I / 2, 1, st))) return rc;
    ConvTCArgs a;
    a.x = g->patches; a.w = g->enc1_wt; a.n = n; a.r = I / 2; a.cin = 64; a.ncols = kEnc[0]; a.taps = 1; a.pad = 0;
    a.epilogue = TC_EPI_ACT_BF16; a.act = ACT_LRELU; a.slope = 0.2f; a.out_bf16 = g->enc_b0; a.split_out = 1;
    a.alg_flops = 2.0 * (double)n * (I / 2) * (I / 2) * kEnc[0] * 18;
    if ((rc = tc_conv(f, a))) return rc;
  }
  const __nv_bfloat16* ex = g->enc_b0;
  int er = I / 2;
  for (int k = 1; k < 5; ++k) {   // blocks 2..5: conv3x3 s2 SAME pad (0, 1) -> InstanceNorm (eps 1e-3) -> LeakyReLU(0.2)
    er /= 2;
    ConvTCArgs a;
    a.x = ex; a.w = g->enc_wt[k]; a.n = n; a.r = er; a.cin = kEnc[k - 1]; a.ncols = kEnc[k]; a.stride = 2; a.pad = 0;
    a.split3 = 1;
    a.epilogue = TC_EPI_BIAS_F32; a.y = g->enc_y;
    const int64_t rows = (int64_t)er * er;
    // InstanceNorm moments (per image, per channel) from the epilogue's per-tile column sums: a 128-pixel tile never
    // spans two images once r*r >= 128, so image b owns tile rows [b, b + 1) * (r*r / 128) * 4 of the pair buffer
    const bool fused = rows >= 128 && rows % 128 == 0;
    a.stat_pairs = fused ? g->stat_pairs : nullptr;
    if ((rc = tc_conv(f, a))) return rc;
    if (fused) {
      if ((rc = channel_stats_from_pairs(g->stat_pairs, n, rows / 128 * 4, rows, kEnc[k], 1e-3f, g->stat_partial,
                                         g->enc_stats_mean, g->enc_stats_rstd, st, g->stat_counters))) return rc;
    } else if ((rc = channel_stats_f32(g->enc_y, kEnc[k], n, rows, kEnc[k], 1e-3f, g->stat_partial, g->enc_stats_mean,
                                       g->enc_stats_rstd, st, g->stat_counters))) return rc;
    __nv_bfloat16* eb = (k & 1) ? g->enc_b1 : g->enc_b0;
    // the last block's output is Flatten()ed (NHWC order) into the heads' operand: hi | lo blocks per image
    if ((rc = affine_act_bf16out(g->enc_y, kEnc[k], g->enc_stats_mean, g->enc_stats_rstd, g->enc_g[k], g->enc_bt[k],
                                 k < 4 ? eb : g->enc_feat_split, nullptr, (int64_t)n * rows, kEnc[k], rows,
                                 ACT_LRELU, 0.2f, k < 4 ? 1 : 2, st))) return rc;
    ex = eb;
  }
  const int feat = er * er * 512;
  {  // Dense "mean" | Dense "variance" (networks.py:31-33): one split-K tcgen05 GEMM + the plane reduction (+ bias)
    ConvTCArgs a;
    a.x = g->enc_feat_split; a.w = g->enc_head_wt; a.n = n; a.r = 1; a.cin = feat; a.ncols = 2 * kLatent; a.taps = 1;
    a.pad = 0; a.split3 = 1; a.ksplit = g->head_ksplit; a.epilogue = TC_EPI_BIAS_F32;
    if (g->head_ksplit > 1) {
      a.y = g->dense_partial;
      if ((rc = tc_conv(f, a))) return rc;
      if ((rc = dense_reduce_planes(g->dense_partial, g->enc_head_b, g->lat_mv, n, 2 * kLatent, g->head_ksplit, st)))
        return rc;
    } else {
      a.y = g->lat_mv; a.bias = g->enc_head_b;
      if ((rc = tc_conv(f, a))) return rc;
    }
  }
  }  // !reuse
  // ---- sampler (sampling.py:11-17) or mean + variance (model.py:789-791); rows hold mean | variance
  if ((rc = sampler_strided_f32(g->lat_mv, 2 * kLatent, g->arch == MSR_ARCH_SPADE ? eps : nullptr, g->latent, n, kLatent,
                                st, g->latent_split))) return rc;
  g->acts["latent"] = {g->latent, (int64_t)n * kLatent, 0};
  // ---- generator (networks.py:37-57)
  float* x = g->xbuf[0];
  {  // Dense(16 * sw * sw * 64) (networks.py:41) on the tensor cores, bias in the epilogue
    ConvTCArgs a;
    a.x = g->latent_split; a.w = g->dense_wt; a.n = n; a.r = 1; a.cin = kLatent; a.ncols = sw * sw * 1024; a.taps = 1;
    a.pad = 0; a.split3 = 1; a.epilogue = TC_EPI_BIAS_F32; a.bias = g->dense_b; a.y = x;
    if ((rc = tc_conv(f, a))) return rc;
  }
  g->acts["x0"] = {x, (int64_t)n * sw * sw * 1024, 0};
  int r = sw, x_shift = 0;
  if ((rc = channel_stats_f32(x, 1024, f.groups, (int64_t)g->B * sw * sw, 1024, 1e-5f, g->stat_partial, g->st_mean[0],
                              g->st_rstd[0], st, g->stat_counters))) return rc;
  for (int k = 0; k < 6; ++k) {
    const BlockW& b = g->rb[k];
    float* y = g->xbuf[(k + 1) & 1];
    // im2col rows of the source at this resolution, shared by the block's SPADEs (only when the mask convolution is
    // not the kernel that builds its operand itself)
    if (!reuse && !mask_in_kernel(g, r) && (rc = source_patches_bf16(source, I, g->patches, n, r, 0, st))) return rc;
    // x = conv_1(lrelu(spade_1(in)))                                  blocks.py:29-30
    if ((rc = spade_bf16(f, b.s1, x, x_shift, g->st_mean[0], g->st_rstd[0], r))) return rc;
    if ((rc = conv_bf16(f, b.c1, r, g->h1, nullptr, 0, g->st_mean[1], g->st_rstd[1]))) return rc;
    const float* res = x;
    int res_shift = x_shift;
    if (b.learned) {  // skip = conv_3(lrelu(spade_3(in)))             blocks.py:34-36
      if ((rc = spade_bf16(f, b.s3, x, x_shift, g->st_mean[0], g->st_rstd[0], r))) return rc;
      if ((rc = conv_bf16(f, b.c3, r, g->s3, nullptr, 0, nullptr, nullptr))) return rc;
      res = g->s3;
      res_shift = 0;
    }
    // x = conv_2(lrelu(spade_2(x))); out = skip + x                   blocks.py:31-32,38
    if ((rc = spade_bf16(f, b.s2, g->h1, 0, g->st_mean[1], g->st_rstd[1], r))) return rc;
    if (k < 5) {
      if ((rc = conv_bf16(f, b.c2, r, y, res, res_shift, g->st_mean[0], g->st_rstd[0]))) return rc;
      g->acts["rb" + std::to_string(k + 1) + ".out"] = {y, (int64_t)n * r * r * b.cout, 0};
    } else {
      // last block: its output only feeds leaky_relu(0.2) + the final conv (networks.py:54-56): emit that directly
      ConvTCArgs a;
      a.x = g->act_bf16; a.w = b.c2.wt; a.n = n; a.r = r; a.cin = b.c2.cin; a.ncols = b.c2.cout;
      a.epilogue = TC_EPI_ACT_BF16; a.bias = b.c2.b; a.res = res; a.res_shift = res_shift; a.act = ACT_LRELU;
      a.slope = 0.2f; a.out_bf16 = g->a_bf16;
      if ((rc = tc_conv(f, a))) return rc;
    }
    x = y;
    x_shift = 1;  // UpSampling2D((2, 2)) after every block (networks.py:44-54), fused into the consumers
    r *= 2;
  }
  // UpSampling2D -> leaky_relu(0.2) -> Conv2D(1, 4, 'same') == 3x3 conv to 4 sub-pixel phases on the low-res tensor
  bool phase_done = false;
  if ((rc = run_phase_layer(g, g->a_bf16, n, r / 2, 128, 0, g->out_b, ACT_NONE, out, 2.0 * (double)n * I * I * 16 * 128, st,
                            &phase_done))) return rc;
  if (!phase_done) {
    ConvTCArgs a;
    a.x = g->a_bf16; a.w = g->out_wt; a.n = n; a.r = r / 2; a.cin = 128; a.ncols = 32;
    a.epilogue = TC_EPI_PHASE_F32; a.bias = g->out_b; a.y = out;
    a.alg_flops = 2.0 * (double)n * I * I * 16 * 128;            // Conv2D(1, 4) on the I x I tensor: 16 taps x 128 channels
    if ((rc = tc_conv(f, a))) return rc;
  }
  g->acts["out"] = {out, (int64_t)n * I * I, 0};
  return MSR_OK;
}

int forward_pix2pix(Fwd& f, const float* source, float* out) {
  msr_generator* g = f.g;
  const int n = (int)f.N;
  cudaStream_t st = f.st;
  int rc;
  // down path (pix2pix.py:64-72, 97-101): Conv4x4 s2 SAME (pad 1,1) no bias, BN (not first), LeakyReLU(0.3)
  const float* x = source;
  int cin = 2, ldx = 2, s = 256;
  for (int k = 0; k < 8; ++k) {
    const int cout = kP2PDown[k];
    float* y;
    int ldy;
    if (k < 7) {  // skip d_{k+1} lives in cat[6-k] behind the up-sampled channels
      y = g->cat[6 - k] + kP2PUp[6 - k];
      ldy = kP2PUp[6 - k] + cout;
    } else {
      y = g->d8;
      ldy = cout;
    }
    ConvF32 c;
    c.x = x; c.w = g->pd_w[k]; c.y = y;
    c.n = n; c.Hs = s; c.Ws = s; c.cin = cin; c.ldx = ldx; c.Hv = s; c.Wv = s;
    c.Ho = s / 2; c.Wo = s / 2; c.cout = cout; c.ldy = ldy;
    c.kh = 4; c.kw = 4; c.stride = 2; c.pad_t = 1; c.pad_l = 1;
    if (k == 0) { c.act = ACT_LRELU; c.act_slope = 0.3f; }
    if ((rc = conv_f32(c, st))) return rc;
    s /= 2;
    if (k > 0) {
      const int64_t M = (int64_t)n * s * s;
      if ((rc = affine_act_f32(y, ldy, g->pd_mean[k], g->pd_rstd[k], g->pd_g[k], g->pd_b[k], y, ldy, M, cout, M,
                               ACT_LRELU, 0.3f, st))) return rc;
    }
    x = y; cin = cout; ldx = ldy;
  }
  // up path (pix2pix.py:74-86, 103-106): ConvT4x4 s2 SAME no bias, BN, ReLU, then Concatenate([up, skip])
  for (int k = 0; k < 7; ++k) {
    const int cout = kP2PUp[k];
    const int ldy = cout + kP2PDown[6 - k];
    float* y = g->cat[k];
    ConvF32 c;
    c.x = x; c.w = g->pu_w[k]; c.y = y;
    c.n = n; c.Hs = s; c.Ws = s; c.cin = cin; c.ldx = ldx; c.Hv = s; c.Wv = s;
    c.Ho = 2 * s; c.Wo = 2 * s; c.cout = cout; c.ldy = ldy;
    c.kh = 4; c.kw = 4; c.stride = 2; c.pad_t = 1; c.pad_l = 1; c.transposed = 1;
    if ((rc = conv_f32(c, st))) return rc;
    s *= 2;
    const int64_t M = (int64_t)n * s * s;
    if ((rc = affine_act_f32(y, ldy, g->pu_mean[k], g->pu_rstd[k], g->pu_g[k], g->pu_b[k], y, ldy, M, cout, M,
                             ACT_RELU, 0.f, st))) return rc;
    x = y; cin = ldy; ldx = ldy;
  }
  ConvF32 c;
  c.x = x; c.w = g->pl_w; c.bias = g->pl_b; c.y = out;
  c.n = n; c.Hs = s; c.Ws = s; c.cin = cin; c.ldx = ldx; c.Hv = s; c.Wv = s;
  c.Ho = 2 * s; c.Wo = 2 * s; c.cout = 1; c.ldy = 1;
  c.kh = 4; c.kw = 4; c.stride = 2; c.pad_t = 1; c.pad_l = 1; c.transposed = 1; c.act = ACT_TANH;
  if ((rc = conv_f32(c, st))) return rc;
  g->acts["out"] = {out, (int64_t)n * 256 * 256, 0};
  return MSR_OK;
}

int forward_pix2pix_bf16(Fwd& f, const float* source, float* out) {
  msr_generator* g = f.g;
  const int n = (int)f.N;
  cudaStream_t st = f.st;
  int rc;
  // ---- down path; skip d_{k+1} lives in cat[6-k] behind the up-sampled channels
  // block 1 (2 input channels): the operand tile is built inside the kernel when the skip half of its concat buffer is
  // channels [64, 128) of a 128-channel tensor (mask_tc.cu MODE 2); otherwise im2col rows + a K = 64 GEMM
  const bool b1_in_kernel = !g_disable_mask_tc && kP2PUp[6] == 64 && kP2PDown[0] == 64;
  if (b1_in_kernel) {
    if ((rc = p2p1_conv_tc(source, 256, g->pdt_w[0], g->catb[6], n, 0.3f, st))) return rc;
  } else if ((rc = source_patches_bf16(source, 256, g->patches, n, 128, 2, st))) return rc;
  const __nv_bfloat16* x = b1_in_kernel ? g->catb[6] + kP2PUp[6] : g->patches;
  int cin = 64, x_pitch = b1_in_kernel ? kP2PUp[6] + kP2PDown[0] : 64, s = b1_in_kernel ? 128 : 256;
  for (int k = b1_in_kernel ? 1 : 0; k < 8; ++k) {
    const int cout = kP2PDown[k];
    __nv_bfloat16* y;
    int pitch;
    if (k < 7) {
      y = g->catb[6 - k] + kP2PUp[6 - k];
      pitch = kP2PUp[6 - k] + cout;
    } else {
      y = g->d8b;
      pitch = cout;
    }
    ConvTCArgs a;
    a.x = x; a.w = g->pdt_w[k]; a.n = n; a.r = s / 2; a.cin = cin; a.x_pitch = x_pitch; a.ncols = cout;
    if (k == 0) { a.taps = 1; a.pad = 0; } else { a.taps = 16; a.stride = 2; a.pad = 1; }
    a.epilogue = TC_EPI_ACT_BF16; a.scale = g->pdt_scale[k]; a.bias = g->pdt_shift[k];   // null for block 1 (no BN)
    a.act = ACT_LRELU; a.slope = 0.3f; a.out_bf16 = y; a.out_pitch = pitch;
    if (k == 0) a.alg_flops = 2.0 * (double)n * (s / 2) * (s / 2) * cout * 32;   // 4x4 taps x 2 channels (K padded to 64)
    if ((rc = tc_conv(f, a))) return rc;
    s /= 2;
    x = y; cin = cout; x_pitch = pitch;
  }
  // ---- up path: ConvT -> BN -> ReLU written in front of the skip inside cat[k]
  for (int k = 0; k < 7; ++k) {
    const int cout = kP2PUp[k];
    const int pitch = cout + kP2PDown[6 - k];
    ConvTCArgs a;
    a.x = x; a.w = g->put_w[k]; a.n = n; a.r = s; a.cin = cin; a.x_pitch = x_pitch; a.ncols = 4 * cout;
    a.epilogue = TC_EPI_PHASE_ACT_BF16; a.scale = g->put_scale[k]; a.bias = g->put_shift[k]; a.act = ACT_RELU;
    a.phase_cout = cout; a.out_bf16 = g->catb[k]; a.out_pitch = pitch;
    a.alg_flops = 2.0 * (double)n * (2 * s) * (2 * s) * 4 * cin * cout;   // ConvT 4x4 s2: k^2/s^2 = 4 taps per output pixel
    if ((rc = tc_conv(f, a))) return rc;
    s *= 2;
    x = g->catb[k]; cin = pitch; x_pitch = pitch;
  }
  // ---- last ConvT(1, 4, s2) + bias + tanh
  bool phase_done = false;
  if ((rc = run_phase_layer(g, x, n, s, cin, x_pitch, g->pl_b, ACT_TANH, out, 2.0 * (double)n * (2 * s) * (2 * s) * 4 * cin,
                            st, &phase_done))) return rc;
  if (!phase_done) {
    ConvTCArgs a;
    a.x = x; a.w = g->plt_w; a.n = n; a.r = s; a.cin = cin; a.x_pitch = x_pitch; a.ncols = 32;
    a.epilogue = TC_EPI_PHASE_F32; a.bias = g->pl_b; a.act = ACT_TANH; a.y = out;
    a.alg_flops = 2.0 * (double)n * (2 * s) * (2 * s) * 4 * cin;
    if ((rc = tc_conv(f, a))) return rc;
  }
  g->acts["out"] = {out, (int64_t)n * 256 * 256, 0};
  return MSR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
extern "C" int msr_generator_create(msr_generator** out, int arch, int image_size, int batch_size, int max_groups,
                                    int precision) {
  MSR_REQUIRE(out, "msr_generator_create: null out");
  MSR_REQUIRE(arch == MSR_ARCH_SPADE || arch == MSR_ARCH_CNN || arch == MSR_ARCH_PIX2PIX, "unknown arch");
  MSR_REQUIRE(precision == MSR_PRECISION_FP32 || precision == MSR_PRECISION_BF16, "unknown precision");
  MSR_REQUIRE(batch_size > 0 && max_groups > 0, "batch_size and max_groups must be positive");
  if (arch == MSR_ARCH_PIX2PIX) {
    MSR_REQUIRE(image_size == 256, "pix2pix is fixed to 256x256 inputs (pix2pix.py:7)");
  } else {
    MSR_REQUIRE(image_size >= 64 && image_size % 64 == 0, "image_size must be a multiple of 64 (networks.py:40)");
    if (precision == MSR_PRECISION_BF16)
      MSR_REQUIRE((image_size & (image_size - 1)) == 0, "bf16 path needs a power-of-two image_size");
  }
  msr_generator* g = new msr_generator();
  g->arch = arch; g->I = image_size; g->B = batch_size; g->maxG = max_groups; g->precision = precision;
  *out = g;
  return MSR_OK;
}

extern "C" int msr_generator_set_weight(msr_generator* g, const char* name, const float* h_data, const int64_t* shape,
                                        int ndim) {
  MSR_REQUIRE(g && name && h_data && shape && ndim > 0 && ndim <= 4, "msr_generator_set_weight: bad arguments");
  if (g->finalized) return fail(MSR_E_STATE, "msr_generator_set_weight after finalize");
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  MSR_REQUIRE(n > 0, "empty tensor");
  t.data.assign(h_data, h_data + n);
  g->host[name] = std::move(t);
  return MSR_OK;
}

extern "C" int msr_generator_finalize(msr_generator* g) {
  MSR_REQUIRE(g, "null generator");
  if (g->finalized) return fail(MSR_E_STATE, "already finalized");
  int rc = (g->arch == MSR_ARCH_PIX2PIX) ? finalize_pix2pix(g) : finalize_spade(g);
  if (rc) return rc;
  MSR_CUDA_CHECK(cudaDeviceSynchronize());
  g->host.clear();
  g->finalized = true;
  return MSR_OK;
}

// gamma | beta cache of repeated-sample mode: one [n * r * r][2C] bf16 tensor per SPADE layer, in call order
// (spade_1, spade_3 when the block has a learned skip, spade_2)
static int ensure_gb_cache(msr_generator* g, int64_t n) {
  if (g->gb_cache_slots >= n) return MSR_OK;
  MSR_REQUIRE(g->gb_cache_slots == 0, "internal: gamma | beta cache cannot grow (it is sized for batch_size * max_groups)");
  const int64_t N = (int64_t)g->B * g->maxG;
  int layer = 0, r = g->I / 64, rc;
  for (int k = 0; k < 6; ++k, r *= 2) {
    const BlockW& b = g->rb[k];
    const int chans[3] = {b.cin, b.learned ? b.cin : 0, b.cout};
    for (int c : chans) {
      if (c == 0) continue;
      if ((rc = ws(g, &g->gb_cache[layer], N * r * r * 2 * c))) return rc;
      ++layer;
    }
  }
  g->gb_cache_slots = N;
  return MSR_OK;
}

extern "C" int msr_generator_forward_repeat(msr_generator* g, const float* d_source, const float* d_eps, float* d_out,
                                            int n_groups, int repeat_phase, void* stream) {
  MSR_REQUIRE(g && d_source && d_out, "msr_generator_forward: null pointer");
  if (!g->finalized) return fail(MSR_E_STATE, "forward before finalize");
  MSR_REQUIRE(n_groups >= 1 && n_groups <= g->maxG, "n_groups out of range");
  MSR_REQUIRE(repeat_phase == MSR_REPEAT_NONE || repeat_phase == MSR_REPEAT_FIRST || repeat_phase == MSR_REPEAT_NEXT,
              "repeat_phase must be MSR_REPEAT_NONE, _FIRST or _NEXT");
  if (g->arch == MSR_ARCH_SPADE) MSR_REQUIRE(d_eps, "GauGAN forward needs eps (sampling.py:13-16)");
  Fwd f;
  f.g = g; f.st = (cudaStream_t)stream; f.groups = n_groups; f.N = (int64_t)n_groups * g->B;
  // only the bf16 SPADE graphs have the cache; every other model simply recomputes (same results as MSR_REPEAT_NONE)
  const bool cached = g->precision == MSR_PRECISION_BF16 && g->arch != MSR_ARCH_PIX2PIX;
  f.phase = cached ? repeat_phase : MSR_REPEAT_NONE;
  if (f.phase != MSR_REPEAT_NONE) {
    int rc = ensure_gb_cache(g, f.N);
    if (rc) return rc;
    if (f.phase == MSR_REPEAT_FIRST) g->gb_cache_valid_groups = n_groups;
    else if (g->gb_cache_valid_groups != n_groups)
      return fail(MSR_E_STATE, "MSR_REPEAT_NEXT without a preceding MSR_REPEAT_FIRST call of the same size");
  } else if (cached) {
    g->gb_cache_valid_groups = 0;   // a plain forward overwrites the encoder outputs the cache belongs to
  }
  if (g->precision == MSR_PRECISION_BF16) {
    const int key = n_groups * 4 + f.phase;
    auto it = g->plans.find(key);
    f.building = (it == g->plans.end());
    f.plans = &g->plans[key];
  }
  const int64_t before = g_launch_count;
  int rc = (g->arch == MSR_ARCH_PIX2PIX)          ? (g->precision == MSR_PRECISION_BF16 ? forward_pix2pix_bf16(f, d_source, d_out)
                                                                                          : forward_pix2pix(f, d_source, d_out))
           : (g->precision == MSR_PRECISION_BF16) ? forward_spade_bf16(f, d_source, d_eps, d_out)
                                                  : forward_spade(f, d_source, d_eps, d_out);
  g->last_launches = g_launch_count - before;
  if (rc && f.building) {  // never keep a half-built plan list
    for (auto* p : *f.plans) conv_tc_plan_destroy(p);
    g->plans.erase(n_groups * 4 + f.phase);
  }
  return rc;
}

extern "C" int msr_generator_forward(msr_generator* g, const float* d_source, const float* d_eps, float* d_out,
                                     int n_groups, void* stream) {
  return msr_generator_forward_repeat(g, d_source, d_eps, d_out, n_groups, MSR_REPEAT_NONE, stream);
}

extern "C" int64_t msr_generator_last_launch_count(const msr_generator* g) { return g ? g->last_launches : -1; }
extern "C" int64_t msr_generator_device_bytes(const msr_generator* g) { return g ? g->dev_bytes : -1; }

extern "C" int msr_generator_read_activation(msr_generator* g, const char* name, float* h_dst, int64_t capacity,
                                             int64_t* count) {
  MSR_REQUIRE(g && name && count, "msr_generator_read_activation: bad arguments");
  auto it = g->acts.find(name);
  if (it == g->acts.end()) return fail(MSR_E_INVALID, std::string("unknown activation ") + name);
  *count = it->second.count;
  if (!h_dst) return MSR_OK;
  MSR_REQUIRE(capacity >= it->second.count, "destination too small");
  MSR_CUDA_CHECK(cudaDeviceSynchronize());
  MSR_CUDA_CHECK(cudaMemcpy(h_dst, it->second.ptr, it->second.count * sizeof(float), cudaMemcpyDeviceToHost));
  return MSR_OK;
}

extern "C" int msr_generator_destroy(msr_generator* g) {
  delete g;
  return MSR_OK;
}
