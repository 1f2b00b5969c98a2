// Device-side operator launchers shared by the generator graphs (declarations).
#pragma once
#include <vector>

#include "common.cuh"

namespace msr {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };

// Generic NHWC convolution on CUDA cores, fp32 (implicit GEMM, 64x64x16 tiles).
struct ConvF32 {
  const float* x = nullptr;  // stored input [n][Hs][Ws][ldx]
  const float* w = nullptr;  // [kh*kw*cin][cout] row-major, k = (ky*kw + kx)*cin + ci
  const float* bias = nullptr;
  float* y = nullptr;        // [n][Ho][Wo][ldy]
  int n = 0, Hs = 0, Ws = 0, cin = 0, ldx = 0;
  int Hv = 0, Wv = 0;        // virtual input size; stored index = (v * in_mul + in_add) >> in_shift
  int in_mul = 1, in_add = 0, in_shift = 0;
  float in_slope = 1.f;      // leaky-relu slope applied to the input on load (1 = identity)
  int Ho = 0, Wo = 0, cout = 0, ldy = 0;
  int kh = 3, kw = 3, stride = 1, pad_t = 1, pad_l = 1;
  int transposed = 0;        // 1: stride-`stride` transposed convolution (gather form)
  int act = ACT_NONE;
  float act_slope = 0.f;
  const float* res = nullptr;  // optional residual [n][Ho >> res_shift][Wo >> res_shift][ldres]
  int res_shift = 0, ldres = 0;
};
int conv_f32(const ConvF32& p, cudaStream_t st);

// Per-(group, channel) mean and 1/sqrt(var + eps) over `rows` consecutive rows of x[groups*rows][ld] (biased variance).
// `partial` is scratch of groups * kStatSplit * C * 2 doubles.
// `counters` (optional): groups * ceil(C / 64) zero-initialised uints; with them the second stage runs inside the same
// launch (the last block of every channel block finalises), without them a second small kernel is launched.
constexpr int kStatSplit = 64;   // upper bound of the row splits (sizes the `partial` scratch)
static inline int stat_splits(int64_t rows, int64_t rows_per_block) {
  const int64_t s = (rows + rows_per_block - 1) / rows_per_block;
  return (int)(s < 1 ? 1 : (s > kStatSplit ? kStatSplit : s));
}
int channel_stats_f32(const float* x, int ld, int groups, int64_t rows, int C, float eps, double* partial, float* mean,
                      float* rstd, cudaStream_t st, unsigned int* counters = nullptr);

// SPADE modulation (spade.py:21-24 + blocks.py:30-34): out[m][c] = lrelu(gamma * (x[src(m)][c] - mean) * rstd + beta)
// gb [M][2C] = gamma | beta; x is [n][r >> x_shift][r >> x_shift][C]; statistics per group of rows_per_group rows of M.
int spade_modulate_f32(const float* gb, const float* x, int x_shift, const float* mean, const float* rstd, float* out,
                       int n, int r, int C, int samples_per_group, float slope, cudaStream_t st);

// y[m][c] = act((x[m][c] - mean[g][c]) * rstd[g][c] * gamma[c] + beta[c]); g = m / rows_per_group (mean/rstd [G][C]).
// gamma / beta may be null (1 / 0).  ldx / ldy are row pitches in elements.
int affine_act_f32(const float* x, int ldx, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, float* y, int ldy, int64_t M, int C, int64_t rows_per_group, int act,
                   float slope, cudaStream_t st);


// latent = mean + exp(0.5 * var) * eps (sampling.py:16) when eps != null, else mean + var (model.py:791).
int sampler_f32(const float* mean, const float* var, const float* eps, float* latent, int64_t count, cudaStream_t st);

// same, with mean | variance stored side by side in rows of pitch ld (mean at [0, L), variance at [L, 2L))
// split_out (optional): the latent again as a split-bf16 row hi (L) | lo (L), the A operand of the tensor-core dense layer
int sampler_strided_f32(const float* mv, int ld, const float* eps, float* latent, int n, int L, cudaStream_t st,
                        __nv_bfloat16* split_out = nullptr);

// Final generator layer (networks.py:54-56): UpSampling2D(2) -> leaky_relu(0.2) -> Conv2D(1, 4, 'same') on
// x [n][r][r][128] fp32 -> out [n][2r][2r] fp32.  w [4][4][128] (Keras [4,4,128,1]), bias scalar.
int final_conv_f32(const float* x, const float* w, const float* bias, float* out, int n, int r, cudaStream_t st);

// ---- tensor-core path (conv_tc.cu) -------------------------------------------------------------------------------
struct ConvTC;  // opaque plan
enum TcEpilogue {
  TC_EPI_BIAS_F32 = 0,    // y_f32 = acc + bias (+ residual); optional fused per-channel statistics partials
  TC_EPI_SPADE_BF16 = 1,  // columns are gamma|beta interleaved per 64; out_bf16 = lrelu(gamma * xhat + beta)
  TC_EPI_ACT_BF16 = 2,    // out_bf16 = act(acc + bias (+ residual))
  TC_EPI_PHASE_F32 = 3,   // 4 sub-pixel phase columns -> y[b][2h+py][2w+px] (final generator layer)
  TC_EPI_PHASE_ACT_BF16 = 4,  // 4 * cout phase columns -> out_bf16[b][2h+py][2w+px][c] (transposed 4x4 s2 convolution)
  TC_EPI_RELU_BF16_T = 5,     // internal: TC_EPI_ACT_BF16 for one 128-column tile, whole-row stores through shared memory
};
struct ConvTCArgs {
  const __nv_bfloat16* x = nullptr;  // [n][r*stride][r*stride][cin] bf16
  const __nv_bfloat16* w = nullptr;  // [ncols][taps*cin] bf16 (K-major), k = tap*cin + ci
  int n = 0, r = 0, cin = 0, ncols = 0;   // r = output side
  int taps = 9;                      // 9: 3x3, 16: 4x4, 1: 1x1
  int x_pitch = 0;                   // channels per pixel of x in memory (0 = cin): x may be a slice of a wider tensor
  int split3 = 0;                    // 1: split-bf16 operands (~fp32 products): x has 2*cin channels (hi | lo), w has
                                     //    3*cin columns per tap (w_hi | w_lo | w_hi) pairing with (x_hi, x_hi, x_lo)
  int split_out = 0;                 // TC_EPI_ACT_BF16: also write lo = bf16(v - hi) at column ncols + c (pitch 2*ncols)
  int ksplit = 1;                    // > 1 (taps = 1, TC_EPI_BIAS_F32 only): split-K -- the channel blocks are cut into ksplit
                                     //     ranges, range ks writes its partial sums to y + ks * (n*r*r) * ncols
  int stride = 1, pad = 1;           // input coordinate = out*stride + k - pad
  int epilogue = TC_EPI_BIAS_F32;
  const float* bias = nullptr;       // [ncols]
  // TC_EPI_BIAS_F32 / TC_EPI_PHASE_F32
  float* y = nullptr;                // [n*r*r][ncols]  (PHASE: [n][2r][2r])
  const float* res = nullptr;        // residual [n][r >> res_shift][r >> res_shift][ncols] or null (also ACT_BF16)
  int res_shift = 0;
  float2* stat_pairs = nullptr;      // [n*r*r/128 * 4][ncols] (sum, sumsq) partials; requires r*r >= 128
  // TC_EPI_SPADE_BF16 (ncols = 2C, column tile of 128 = 64 gamma | 64 beta of the same channels)
  const float* sx = nullptr;         // normalised tensor [n][r >> sx_shift][r >> sx_shift][C] fp32
  int sx_shift = 0;
  const float* mean = nullptr;       // [groups][C]
  const float* rstd = nullptr;
  int samples_per_group = 1;
  float slope = 0.2f;
  int act = ACT_NONE;                // TC_EPI_ACT_BF16 / PHASE epilogues
  const float* scale = nullptr;      // optional per-column scale before the bias (folded BatchNorm)
  int out_pitch = 0;                 // channels per output pixel in memory (0 = dense)
  int phase_cout = 0;                // TC_EPI_PHASE_ACT_BF16: channels per phase
  __nv_bfloat16* out_bf16 = nullptr; // [n*r*r][C] (SPADE) / [n*r*r][ncols] (ACT)
  double alg_flops = 0.0;            // algorithmic FLOPs of the layer this launch computes when they differ from
                                     // 2*M*ncols*taps*cin (zero-padded K or N: 18 real taps of 64, 4 real columns of 32);
                                     // 0 = the GEMM formula.  Only the profiler's roofline accounting reads it.
};
int conv_tc_plan_create(ConvTC** plan, const ConvTCArgs& a);
int conv_tc_launch(const ConvTC* plan, cudaStream_t st);
void conv_tc_update_pointers(ConvTC* plan, const ConvTCArgs& a);   // refresh the epilogue pointers of a cached plan
void conv_tc_plan_destroy(ConvTC* plan);

// ---- one-channel sub-pixel phase layers (phase_tc.cu): 1x1 GEMM per pixel + 3x3 stencil of scalars ---------------------
// networks.py:54-56 (UpSampling2D -> leaky_relu -> Conv2D(1, 4)) and pix2pix.py:91-95 (Conv2DTranspose(1, 4, s2) + tanh)
constexpr int kPhaseMaxCols = 25;    // (tap, phase) pairs with a non-zero filter: 25 for the former, 16 for the latter
struct PhaseTable {
  int kind = -1;                     // 0: 4x4 conv of the x2-upsampled tensor, 1: 4x4 stride-2 transposed conv
  int ncols = 0;                     // 25 / 16 columns, ordered phase-major, then ty, then tx
};
struct PhaseTC;  // opaque plan
struct PhaseTCArgs {
  const __nv_bfloat16* x = nullptr;   // [n][r][r][x_pitch] bf16, channels [0, cin) are read
  const __nv_bfloat16* wg = nullptr;  // [32][cin] bf16: row j = the cin weights of column j (phase_tc_pack), rest zero
  const PhaseTable* tab = nullptr;
  const float* bias = nullptr;        // one value or null
  float* y = nullptr;                 // [n][2r][2r] fp32
  int n = 0, r = 0, cin = 0, x_pitch = 0;
  int act = ACT_NONE;                 // ACT_NONE or ACT_TANH
  double alg_flops = 0.0;
};
bool phase_tc_supported(int r, int cin);
// w4: the [4][9*cin] bf16 phase-combined 3x3 filters (the real rows of the TC_EPI_PHASE_F32 weight matrix) -> wg, table.
// Returns the number of columns, or -1 when the non-zero pattern is neither of the two layer kinds.
int phase_tc_pack(const uint16_t* w4, int cin, std::vector<uint16_t>* wg, PhaseTable* tab);
int phase_tc_plan_create(PhaseTC** plan, const PhaseTCArgs& a);
int phase_tc_launch(const PhaseTC* plan, cudaStream_t st);
void phase_tc_set_output(PhaseTC* plan, float* y);
void phase_tc_plan_destroy(PhaseTC* plan);

// ---- SPADE mask convolution without an im2col buffer (mask_tc.cu) --------------------------------------------------------
// out[n][r][r][128] bf16 = relu(conv3x3(nearest_resize(source [n][I][I][2] fp32 -> r x r), w) + bias)   (spade.py:17-18)
// wm: [128][64] bf16 from mask_tc_pack_weights (w: Keras kernel [3][3][2][128] float32, bias [128]): split-bf16 K layout,
// the bias folded in as two more K rows.
void mask_tc_pack_weights(const float* w, const float* bias, int cout, std::vector<uint16_t>* out);
bool mask_tc_supported(int I, int r);
int mask_conv_tc(const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out, int n, int r, cudaStream_t st);
// Encoder block 1 with the same kernel (blocks.py:53-60, no norm; networks.py:12): out [n][I/2][I/2][128] bf16 = hi (64) |
// lo (64) of leaky_relu(conv3x3 stride 2, SAME (0, 1), 2 -> 64, no bias); wm [64][64] from mask_tc_pack_weights(w, null, 64).
int enc1_conv_tc(const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out, int n, float slope, cudaStream_t st);
// pix2pix block 1 with the same kernel (pix2pix.py:64-72, no norm): channels [64, 128) of out128 [n][I/2][I/2][128] bf16 =
// leaky_relu(conv4x4 stride 2, SAME (1, 1), 2 -> 64, no bias); wm [64][64]: k = (ky*4 + kx)*2 + c, columns [w | w].
int p2p1_conv_tc(const float* source, int I, const __nv_bfloat16* wm, __nv_bfloat16* out128, int n, float slope,
                 cudaStream_t st);

// ---- small helpers for the bf16 path (nn_bf16.cu) -----------------------------------------------------------------
// im2col of the 2-channel source for a 3x3 convolution at output side r: out [n][r][r][64] bf16, channel (ky*3+kx)*2+c
// for c in {ortho, dem}, channels 18..63 zero.  mode 0: SPADE's mask path = nearest resize (half-pixel centres,
// spade.py:17) to r x r followed by SAME padding (1, 1); mode 1: encoder block 1 = stride-2 taps on the full
// resolution source with SAME padding (0, 1) (r = I / 2, blocks.py:53-60); mode 2: pix2pix block 1 = 4x4 stride-2 taps
// with SAME padding (1, 1), 32 values as hi (channels 0..31) | lo (32..63) (pix2pix.py:64-72).
int source_patches_bf16(const float* source, int I, __nv_bfloat16* out, int n, int r, int mode, cudaStream_t st);

// out[m][n] = sum_k x[m][k] * w[k][n] + bias[n], bf16 weights, fp32 activations / accumulation.  ldo = row pitch of
// out.  `partial` scratch of ksplit * M * N floats.
int dense_f32w(const float* x, const float* w, const float* bias, float* out, int M, int K, int N, float* partial,
               int64_t partial_capacity, cudaStream_t st);

// (x - mean) * rstd * gamma + beta -> activation -> bf16 (and / or fp32) output; same conventions as affine_act_f32.
int affine_act_bf16out(const float* x, int ldx, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, __nv_bfloat16* y_bf16, float* y_f32, int64_t M, int C, int64_t rows_per_group,
                       int act, float slope, int split, cudaStream_t st);
// split = 1: y_bf16 has row pitch 2C and receives hi = bf16(v) at [0, C) and lo = bf16(v - hi) at [C, 2C)
// split = 2: flattened-image layout for the dense heads: image b = m / rows_per_group owns 2 * rows_per_group * C values,
//            hi of (pixel p, channel c) at b*2K + p*C + c and lo at b*2K + K + p*C + c with K = rows_per_group * C

// sum of the `planes` split-K partial planes [planes][M][N] (+ bias[N]) -> out[M][N]
int dense_reduce_planes(const float* partial, const float* bias, float* out, int M, int N, int planes, cudaStream_t st);

// SPADE modulation + LeakyReLU from cached gamma | beta columns (bf16, [n*r*r][2C], 64 gamma | 64 beta per channel block)
int spade_modulate_cached_bf16(const __nv_bfloat16* gb, const float* x, int x_shift, const float* mean, const float* rstd,
                               __nv_bfloat16* out, int n, int r, int C, int samples_per_group, float slope,
                               cudaStream_t st);

// statistics from the fused (sum, sumsq) pairs written by the tensor-core epilogue: pairs [groups*rows_p][C]
int channel_stats_from_pairs(const float2* pairs, int groups, int64_t rows_p, int64_t count_per_group, int C, float eps,
                             double* partial, float* mean, float* rstd, cudaStream_t st,
                             unsigned int* counters = nullptr);

}  // namespace msr
