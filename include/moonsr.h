/*
 * moonsr.h -- C ABI of libmoonsr.so: the B200 (sm_100a) implementation of MoonSuperResolution's tiled full-DEM
 * inference path.  This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MSR_E_* code otherwise; msr_last_error() gives the text of the
 *     last failure on the calling thread.  No C++ exception crosses this boundary.
 *   - pointers whose name starts with d_ are DEVICE pointers (cudaMalloc / torch-allocated), h_ are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is asynchronous on it unless
 *     stated otherwise; no function allocates device memory except *_create / *_finalize.
 *   - rasters are row-major (rows, cols) float32; activations are NHWC.
 *   - (x, y) = (column, row).  I = image_size, S = stride, T = tile_size, p = purge = I / 16, off = I - S.
 *
 * Each entry point names the reference code it replaces (file:line in AntoineRichard/MoonSuperResolution).
 */
#ifndef MOONSR_H_
#define MOONSR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSR_OK 0
#define MSR_E_INVALID (-1) /* bad argument */
#define MSR_E_CUDA (-2)    /* CUDA runtime / driver error */
#define MSR_E_STATE (-3)   /* call order violated (e.g. forward before finalize) */
#define MSR_E_NOMEM (-4)

#define MSR_ARCH_SPADE 0   /* GauGAN:   encoder -> Gaussian sampler -> SPADE generator (spade/models/model.py:564-567) */
#define MSR_ARCH_CNN 1     /* CNNSpade: encoder -> mean + variance  -> SPADE generator (spade/models/model.py:789-791) */
#define MSR_ARCH_PIX2PIX 2 /* Pix2Pix U-Net generator, training=False (pix2pix.py:64-108) */

#define MSR_PRECISION_FP32 0 /* CUDA-core fp32 everywhere (parity mode, <= 1e-4) */
#define MSR_PRECISION_BF16 1 /* tcgen05 implicit-GEMM convolutions, bf16 operands, fp32 accumulation */

int msr_version(void);
const char* msr_last_error(void);
/* Device properties the host side needs; returns the SM count of the current device (or <0). */
int msr_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Per-kernel-family timing (measurement aid for bench.py; adds a cudaEvent pair around every launch group while on)
 * ------------------------------------------------------------------------------------------------------------- */
#define MSR_PROF_CONV_TC 0   /* tcgen05 implicit-GEMM convolutions (main convs + fused SPADE gamma/beta convs) */
#define MSR_PROF_CONV_F32 1  /* CUDA-core fp32 convolutions (encoder, fp32 mode, pix2pix) */
#define MSR_PROF_MASK_CONV 2 /* SPADE's 2->128 mask conv (bf16 mode) */
#define MSR_PROF_STATS 3     /* batch / instance statistics */
#define MSR_PROF_ELEMWISE 4  /* SPADE modulation (fp32 mode), affine+activation, sampler */
#define MSR_PROF_DENSE 5     /* dense layers */
#define MSR_PROF_FINAL_CONV 6
#define MSR_PROF_PAD 7
#define MSR_PROF_VALIDITY 8
#define MSR_PROF_GATHER 9    /* gather + normalise */
#define MSR_PROF_BLEND 10
#define MSR_PROF_PREPROCESS 11 /* preprocess: 1/4 area resize, cubic upsampling */
#define MSR_PROF_COUNT 12

/* Turns event timing on (clears the counters) or off. */
int msr_profile_enable(int on);
/* Synchronises the device and returns, per family f < MSR_PROF_COUNT: ms[f] summed event time, work[f] summed
 * algorithmic work (FLOPs for the convolution / dense families, bytes for the others), launches[f]. */
int msr_profile_read(double* ms, double* work, int64_t* launches);
/* Per-launch-group records of one family, in launch order: fills up to `capacity` entries of ms / work and returns the
 * total number of records in *count. */
int msr_profile_records(int family, double* ms, double* work, int64_t capacity, int64_t* count);

/* ---------------------------------------------------------------------------------------------------------------
 * preprocess  (process_full_tiles.py:226-244): DEM -> 1/4 -> (small-hole fill on the host) -> 1/16 -> bicubic to full size
 * ------------------------------------------------------------------------------------------------------------- */

/* cv2.resize(x, (0, 0), fx=0.25, fy=0.25, interpolation=cv2.INTER_AREA) of a (H, W) float32 raster
 * (process_full_tiles.py:232, 240) with the bookkeeping around it fused: pixels <= no_value count as NaN (:230, :238), a
 * NaN result is stored as no_value (:233).  (dh, dw) must be (round(H/4), round(W/4)), rounding half to even as cv2 does.
 * Bit-exact with OpenCV's own float32 code: 4x4 box mean; windows cut by the raster edge are averaged over the pixels
 * that exist. */
int msr_resize_area4(const float* d_src, int H, int W, float* d_dst, int dh, int dw, float no_value, void* stream);

/* cv2.resize(x, (W, H), interpolation=cv2.INTER_CUBIC) of a (h, w) float32 raster (process_full_tiles.py:241), source
 * pixels <= no_value read as NaN (:238), NaN results stored as no_value (:243).  Per destination column / row the host
 * supplies the source index of tap 1 (d_xofs (W), d_yofs (H); taps are ofs-1 .. ofs+2, clamped to the raster) and the
 * four float32 weights (d_xcoef (W, 4), d_ycoef (H, 4), 16-byte aligned) -- moonsuperresolution_b200/preprocess.py
 * builds them as OpenCV does.  Bit-exact with OpenCV's own code (pip wheels dispatch this call to Intel IPP, which
 * rounds differently by a few ulp). */
int msr_resize_cubic(const float* d_src, int h, int w, float* d_dst, int H, int W, const int32_t* d_xofs,
                     const float* d_xcoef, const int32_t* d_yofs, const float* d_ycoef, float no_value, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Tiling / blending  (process_full_tiles.py)
 * ------------------------------------------------------------------------------------------------------------- */

/* padInputs (process_full_tiles.py:246-267): fill both (CH, CW) canvases with no_value and paste the (H, W) rasters
 * at rows [off_y, off_y + H) x columns [off_x, off_x + W), clipped to the canvas.  The reference has off_y = off_x =
 * off; a rank that holds only a band of canvas rows passes the band's row offset (off_y may be negative). */
int msr_pad_inputs(const float* d_dem, const float* d_img, int H, int W, float* d_dem_canvas, float* d_img_canvas,
                   int CH, int CW, int off_y, int off_x, float no_value, void* stream);

/* getPatch's validity test (process_full_tiles.py:286-292), factored: summed-area table of the invalid mask
 * (img <= no_value || dem <= no_value).  d_sat is int32 (CH + 1, CW + 1), first row / column zero. */
int msr_validity_sat(const float* d_img_canvas, const float* d_dem_canvas, int CH, int CW, float no_value,
                     int32_t* d_sat, void* stream);

/* valid[k] = 1 iff the I x I window at canvas origin (xy[2k], xy[2k+1]) holds no invalid pixel (clipped to the canvas
 * like a numpy slice).  Exact integer arithmetic. */
int msr_patch_validity(const int32_t* d_sat, int CH, int CW, const int32_t* d_xy, int n, int I, uint8_t* d_valid,
                       void* stream);

/* normalize (process_full_tiles.py:295-311) for n patches: out[k] = (I, I, 2) float32, channel 0 = ortho, channel 1 =
 * DEM, each ((v - min) / (max - min)) - 0.5 in float32 with IEEE division; minmax[k] = {img_min, img_max, dem_min,
 * dem_max}.  A slot with origin (-1, -1) is a padding slot (process_full_tiles.py:468-474): all-zero output.
 * d_partial is scratch of n * 32 * 4 floats. */
int msr_gather_normalize(const float* d_img_canvas, const float* d_dem_canvas, int CH, int CW, const int32_t* d_xy,
                         int n, int I, float* d_out_nhwc, float* d_minmax, float* d_partial, void* stream);

/* rebuildTile (process_full_tiles.py:363-414) fused with the tile's paste + crop of rebuildMap (:541-545), in gather
 * form: one thread per output pixel replays, in the reference's patch order, the Gaussian-weighted incremental
 * mean / variance update with float64 intermediates and float32 state, then writes mean, std, good.
 *   d_patch_ptr[k]   (I, I) prediction of patch k, float32 (or float64 where d_patch_f64[k] != 0; may be NULL)
 *   d_patch_lohi[k]  {dem_min, dem_max} of patch k (float32)
 *   d_patch_xy[k]    patch origin (x, y) relative to the tile's canvas origin, i.e. the reference's dict key
 *   n                number of patches, in dict insertion order
 *   d_lattice        NULL, or (G, G) int32 with d_lattice[gy * G + gx] = k for the patch whose key is (gx*S, gy*S), -1
 *                    if none; legal only when keys lie on that lattice in row-major insertion order.  With NULL the
 *                    kernel scans all n patches per pixel (any keys, any order).
 *   d_weights        float64 (I - 2p, I - 2p): makeGaussianKernel() + 1e-7 cropped by p (:391-393)
 *   add_half         1: predictions are raw network outputs, add 0.5f first (processBatch, :340); 0: already added
 *   outputs          pixel (row, col) of the T x T tile centre goes to out[row * pitch + col] for row < rows,
 *                    col < cols (rows, cols <= T clip the tile to the raster).  good is uint8.
 */
int msr_blend_tile(const void* const* d_patch_ptr, const uint8_t* d_patch_f64, const float* d_patch_lohi,
                   const int32_t* d_patch_xy, int n, const int32_t* d_lattice, int G, const double* d_weights, int I,
                   int S, int T, int add_half, float no_value, float* d_mean, float* d_std, uint8_t* d_good,
                   int64_t pitch, int rows, int cols, void* stream);

/* Dedup mode (mode B: every patch position of the global stride lattice is generated once).
 * The state of rebuildTile's loop (process_full_tiles.py:386-402) -- float32 w_sum, mean, S -- lives in canvas-shaped
 * accumulators (acc_rows, pitch) whose row 0 is canvas row acc_y0; one call replays the loop body (:395-402, float64
 * intermediates, float32 stores) for the patches [k0, k0 + n) of a band:
 *   d_pred      (n, I, I) float32 predictions of those patches (raw network outputs when add_half = 1)
 *   d_lohi      (n, 2) {dem_min, dem_max}
 *   d_lattice   (GY, GX) int32: index of the patch at canvas origin (gx*S, lattice_y0 + gy*S) in the band's visit order
 *               (y outer, x inner, process_full_tiles.py:453-454), -1 where the position holds no valid patch
 *   gy_lo/gy_hi lattice rows (inclusive) that contain patches [k0, k0 + n)
 *   row_lo/hi   only canvas rows [row_lo, row_hi) are updated (a rank defers the rows that first need its neighbour's
 *               contributions: the update is order-dependent, :400-402)
 * Calls must be issued in visit order; then every pixel receives its patches in the reference's order and the result
 * is bit-identical to rebuildTile for the same predictions. */
int msr_blend_accumulate(const float* d_pred, const float* d_lohi, int k0, int n, const int32_t* d_lattice, int GY,
                         int GX, int gy_lo, int gy_hi, int lattice_y0, const double* d_weights, int I, int S,
                         int add_half, float* d_wsum, float* d_mean, float* d_s, int64_t pitch, int acc_y0,
                         int acc_rows, int cols, int row_lo, int row_hi, void* stream);

/* rebuildTile's tail (process_full_tiles.py:404-413) on a (rows, cols) window of the accumulators (pointers already
 * offset to the window's first pixel): good = w_sum > 0, std = sqrt(S / w_sum) in float32, mean / std = no_value where
 * not good; written to (rows, out_pitch) rasters -- the crop of rebuildMap (:541-545) is the choice of window. */
int msr_blend_finalize(const float* d_wsum, const float* d_mean_acc, const float* d_s, int64_t acc_pitch, int rows,
                       int cols, float no_value, float* d_mean, float* d_std, uint8_t* d_good, int64_t out_pitch,
                       void* stream);

/* "Fast" variants of msr_blend_tile / msr_blend_accumulate (DSRConfig(blend="fast")): the same update rule
 * (process_full_tiles.py:395-402), the same patch-to-pixel placement, visiting order and `good` mask, but the running
 * w_sum / mean / S are updated in float32 (no float64 intermediates), four adjacent pixels per thread with 128-bit
 * loads and stores, so the kernels are bound by HBM instead of the double-precision pipe.  Values agree with the
 * bit-exact kernels to float32 rounding of the update (~1e-6 relative; tests/test_gpu_tiling.py).  Restrictions:
 * predictions are one contiguous float32 array (n, I, I) indexed through d_lattice; I % 64 == 0, S % 4 == 0; the output
 * window / accumulators are 16-byte aligned with pitches that are multiples of 4.  d_weights_f32 is the float32 copy of
 * msr_blend_tile's weight table; the tile form rebuilds the weight from the separable part of makeGaussianKernel
 * (process_full_tiles.py:347-361): w(ry, rx) = d_weights_1d[ry - p] * d_weights_1d[rx - p] * c1 + c0 with
 * d_weights_1d[i] = exp(-x_i^2 / 2 s^2) over the purge-cropped axis (I - 2p floats). */
int msr_blend_tile_fast(const float* d_pred, const float* d_lohi, int n, const int32_t* d_lattice, int G,
                        const float* d_weights_1d, float c1, float c0, int I, int S, int T, int add_half,
                        float no_value, float* d_mean, float* d_std, uint8_t* d_good, int64_t pitch, int rows, int cols,
                        void* stream);
int msr_blend_accumulate_fast(const float* d_pred, const float* d_lohi, int k0, int n, const int32_t* d_lattice, int GY,
                              int GX, int gy_lo, int gy_hi, int lattice_y0, const float* d_weights_f32, int I, int S,
                              int add_half, float* d_wsum, float* d_mean, float* d_s, int64_t pitch, int acc_y0,
                              int acc_rows, int cols, int row_lo, int row_hi, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Raster container codec (host side, multi-threaded): what GDAL does for the reference when it reads band 1 of the input
 * GeoTIFFs (process_full_tiles.py:158-182) and writes 'COMPRESS=LZW', 'PREDICTOR=2' GeoTIFFs (:481-531).  The IFD / tags
 * are handled by moonsuperresolution_b200/geotiff.py; these functions only transform strip / tile bytes.
 * ------------------------------------------------------------------------------------------------------------- */

/* Upper bound of the LZW output for raw_bytes of input. */
int64_t msr_tiff_lzw_bound(int64_t raw_bytes);

/* Compress a (rows, row_bytes) host raster as strips of rows_per_strip rows.  compression: 1 none, 5 LZW; predictor: 1
 * none, 2 horizontal differencing on sample_bytes-wide integers.  h_out holds one slot of slot_bytes per strip; h_sizes
 * receives the compressed byte count of every strip. */
int msr_tiff_encode_strips(const uint8_t* h_raster, int64_t row_bytes, int rows, int rows_per_strip, int sample_bytes,
                           int compression, int predictor, uint8_t* h_out, int64_t slot_bytes, int64_t* h_sizes,
                           int n_threads);

/* Decode n_chunks strips or tiles of a TIFF file image into a (rows, row_bytes) host raster.  Chunk i occupies file
 * bytes [offsets[i], offsets[i] + counts[i]), decodes to chunk_rows x chunk_row_bytes and is pasted at raster row
 * dst_row[i], byte column dst_col[i] (clipped).  compression: 1 none, 5 LZW; predictor: 1, 2 (horizontal), 3 (float). */
int msr_tiff_decode_chunks(const uint8_t* h_file, int64_t file_bytes, const int64_t* offsets, const int64_t* counts,
                           const int64_t* dst_row, const int64_t* dst_col, int n_chunks, int chunk_rows,
                           int64_t chunk_row_bytes, int sample_bytes, int compression, int predictor, uint8_t* h_raster,
                           int64_t row_bytes, int64_t rows, int n_threads);

/* ---------------------------------------------------------------------------------------------------------------
 * Generators  (spade/models/{networks,blocks,spade,sampling}.py, model.py:564-567 / 789-791, pix2pix.py:64-108)
 * ------------------------------------------------------------------------------------------------------------- */

typedef struct msr_generator msr_generator;

/* Builds an empty model.  `batch_size` is the reference's batch (the unit over which SPADE's batch statistics are
 * taken, spade.py:21); `max_groups` batches can be pushed through one forward call, each with its own statistics. */
int msr_generator_create(msr_generator** out, int arch, int image_size, int batch_size, int max_groups, int precision);

/* Supplies one tensor in Keras layout (names: moonsuperresolution_b200/weights.py).  Host float32, copied. */
int msr_generator_set_weight(msr_generator* g, const char* name, const float* h_data, const int64_t* shape, int ndim);

/* Checks that every tensor is present, repacks (gamma|beta concatenation, [N][K] bf16 ...), uploads, allocates the
 * workspace and builds the TMA descriptors.  Synchronous. */
int msr_generator_finalize(msr_generator* g);

/* One forward pass over n_groups * batch_size patches.
 *   d_source  (n_groups * B, I, I, 2) float32, as produced by msr_gather_normalize
 *   d_eps     (n_groups * B, 256) float32 standard-normal draws for the Gaussian sampler (sampling.py:13-16);
 *             required for MSR_ARCH_SPADE, ignored otherwise
 *   d_out     (n_groups * B, I, I) float32: last (only) output channel, before the + 0.5 of processBatch */
int msr_generator_forward(msr_generator* g, const float* d_source, const float* d_eps, float* d_out, int n_groups,
                          void* stream);

/* Repeated-sample mode (beyond the reference): the same n_groups * batch_size patches are generated
 * several times with new sampler noise.  spade.py:19-20's gamma / beta convolutions, the mask convolution (spade.py:18)
 * and the encoder (networks.py:8-34) depend only on d_source, so:
 *   MSR_REPEAT_FIRST  computes them, stores gamma | beta of all 15 SPADE layers (bf16) and the encoder's mean | variance,
 *                     and produces the first generation from the stored values;
 *   MSR_REPEAT_NEXT   produces a further generation of the SAME d_source from the stored values: only the sampler, the
 *                     dense layer, the modulation (spade.py:21-24) and the main convolutions (blocks.py:30-36) run;
 *   MSR_REPEAT_NONE   = msr_generator_forward.
 * FIRST and NEXT give bit-identical outputs for identical d_eps.  Models without the cache (fp32 mode, pix2pix) treat
 * every phase as MSR_REPEAT_NONE. */
enum { MSR_REPEAT_NONE = 0, MSR_REPEAT_FIRST = 1, MSR_REPEAT_NEXT = 2 };
int msr_generator_forward_repeat(msr_generator* g, const float* d_source, const float* d_eps, float* d_out, int n_groups,
                                 int repeat_phase, void* stream);

/* Number of kernel launches issued by the last forward call (for bench.py's gpu_launches). */
int64_t msr_generator_last_launch_count(const msr_generator* g);

/* Bytes of device memory held (weights + workspace). */
int64_t msr_generator_device_bytes(const msr_generator* g);

/* Debug / parity aid: copies a named intermediate of the last forward (e.g. "latent", "rb3.out") to the host as
 * float32; *count receives the element count (call with h_dst == NULL to query).  Synchronous. */
int msr_generator_read_activation(msr_generator* g, const char* name, float* h_dst, int64_t capacity, int64_t* count);

int msr_generator_destroy(msr_generator* g);

/* ---------------------------------------------------------------------------------------------------------------
 * Single operators, exported for the parity tests of the tensor-core kernels
 * ------------------------------------------------------------------------------------------------------------- */

/* 3x3 stride-1 SAME convolution as tcgen05 implicit GEMM.  d_x (n, r, r, cin) bf16 (uint16 bits), d_w (cout, 9*cin)
 * bf16 with k = (ky*3 + kx)*cin + ci, d_bias (cout) float32 or NULL, d_y (n, r, r, cout) float32. */
int msr_op_conv3x3_bf16(const uint16_t* d_x, const uint16_t* d_w, const float* d_bias, float* d_y, int n, int r,
                        int cin, int cout, void* stream);

/* General form of the tensor-core convolution (taps = 9: 3x3, taps = 1: 1x1; stride 1 or 2; input coordinate of tap
 * (ky, kx) for output (h, w) = (h*stride + ky - pad, w*stride + kx - pad), zero outside).  d_x (n, r_out*stride,
 * r_out*stride, cin) bf16, d_w (cout, taps*cin) bf16.  Exactly one of d_y_f32 (n, r_out, r_out, cout; acc + bias) and
 * d_y_bf16 (act(acc + bias), act: 0 none, 1 relu, 2 leaky-relu(slope)) is non-NULL.  d_stat_pairs (optional, with
 * d_y_f32, needs r_out^2 >= 128): (n*r_out^2/128*4, cout, 2) float32 per-(128-pixel tile, warp) column sums and sums of
 * squares, the fused form of SPADE's batch moments (spade.py:21). */
int msr_op_conv_tc(const uint16_t* d_x, const uint16_t* d_w, const float* d_bias, float* d_y_f32, uint16_t* d_y_bf16,
                   int n, int r_out, int cin, int cout, int taps, int stride, int pad, int act, float slope,
                   float* d_stat_pairs, void* stream);

/* The fused SPADE operator (spade.py:19-24 + blocks.py:30): gamma | beta 3x3 convolution of the 128-channel mask
 * features d_a (n, r, r, 128) bf16 with d_w (2C, 1152) bf16 whose rows are interleaved per 64 channels (64 gamma rows,
 * 64 beta rows, ...) and d_bias (2C) in the same order; epilogue out = leaky_relu(gamma * (x - mean) * rstd + beta, 0.2)
 * as bf16 (n, r, r, C), with x = d_x (n, r >> x_shift, r >> x_shift, C) float32 read at (h >> x_shift, w >> x_shift)
 * (nearest upsampling fused) and d_mean / d_rstd (n / samples_per_group, C). */
int msr_op_spade_tc(const uint16_t* d_a, const uint16_t* d_w, const float* d_bias, const float* d_x, int x_shift,
                    const float* d_mean, const float* d_rstd, int samples_per_group, uint16_t* d_out, int n, int r,
                    int C, void* stream);

/* The one-channel sub-pixel phase layer (networks.py:54-56 / pix2pix.py:91-95 after the phase decomposition) in its
 * "contract once per pixel, then stencil" form (csrc/phase_tc.cu).  d_x (n, r, r, cin) bf16 on the device; h_w4 (4, 9*cin)
 * bf16 bits on the HOST: row q = py*2 + px holds the 3x3 filter of sub-pixel phase q, k = (ky*3 + kx)*cin + c (all-zero
 * taps are dropped); d_bias one float32 or NULL; act 0 none / 3 tanh; d_y (n, 2r, 2r) float32 with
 * y[b][2h+py][2w+px] = act(bias + sum_{ky,kx,c} x[b][h+ky-1][w+kx-1][c] * w4[q][ky][kx][c]), zero outside the tensor.
 * r must be 128 or 256, cin 64 or 128.  Synchronises the stream before it returns. */
int msr_op_phase_tc(const uint16_t* d_x, const uint16_t* h_w4, const float* d_bias, float* d_y, int n, int r, int cin,
                    int act, void* stream);

/* SPADE's mask convolution (spade.py:17-18) with the operand tile built inside the kernel (csrc/mask_tc.cu):
 * d_out (n, r, r, 128) bf16 = relu(conv3x3_same(nearest_resize(d_source (n, I, I, 2) float32 -> r x r), h_w) + h_bias),
 * nearest resize with half-pixel centres (mask pixel (h, w) = source pixel (h*I/r + I/(2r), ...)); h_w (3, 3, 2, 128)
 * float32 Keras kernel and h_bias (128) float32, both on the HOST (the bias is folded into the packed weights);
 * split-bf16 operands (~float32 products).  r a power of two dividing I.  Synchronises the stream before it returns. */
int msr_op_mask_tc(const float* d_source, int I, const float* h_w, const float* h_bias, uint16_t* d_out, int n, int r,
                   void* stream);

/* Encoder block 1 (blocks.py:53-60 without the norm, networks.py:12) with the same kernel: d_out (n, I/2, I/2, 128) bf16 =
 * hi (64 channels) | lo (64 channels) of v = leaky_relu(conv3x3(d_source (n, I, I, 2) float32, strides 2, SAME = pad
 * (0, 1), h_w (3, 3, 2, 64) float32 Keras kernel on the HOST, no bias), slope), hi = bf16(v), lo = bf16(v - hi): the
 * split-bf16 operand of the next encoder convolution.  Synchronises the stream before it returns. */
int msr_op_enc1_tc(const float* d_source, int I, const float* h_w, uint16_t* d_out, int n, float slope, void* stream);

/* Host-only halves of the two layers above (no GPU involved; the CPU test suite checks them against numpy):
 * msr_host_pack_mask_weights: Keras kernel h_w (3, 3, 2, cout) float32 (+ h_bias (cout) or NULL), cout 64 or 128 ->
 *   h_out (cout, 64) bf16 bits in the split-bf16 K layout of csrc/mask_tc.cu: k = 4t + {0,1,2,3} -> (whi0, whi1, whi0, whi1)
 *   of tap t = ky*3 + kx, k = 36 + 2t + {0,1} -> (wlo0, wlo1), k = 54 / 55 -> bias hi / lo, rest zero.
 * msr_host_pack_phase_weights: h_w4 (4, 9*cin) bf16 bits -> h_wg (32, cin): one row per (phase, tap) pair with a non-zero
 *   filter, phase-major then tap order; *kind 0 (4x4 conv of the x2-upsampled tensor: 25 rows), 1 (4x4 stride-2
 *   transposed conv: 16 rows) or -1 (neither pattern: nothing written). */
int msr_host_pack_mask_weights(const float* h_w, const float* h_bias, int cout, uint16_t* h_out);
int msr_host_pack_phase_weights(const uint16_t* h_w4, int cin, uint16_t* h_wg, int* kind, int* ncols);

/* Optimisation aid: when d_counters != NULL (148 * 8 int64, zero-initialised by the caller), every tensor-core convolution
 * planned afterwards records per-CTA cycle counts: [0] producer wait on empty stages, [1] producer total, [2] MMA issuer
 * wait on full stages, [3] MMA issuer wait on free accumulators, [4] MMA issuer total, [5] epilogue wait on accumulators.
 * Pass NULL to switch off. */
int msr_debug_tc_counters(long long* d_counters);

/* Same operator on CUDA cores in float32 (the fp32-mode kernel); d_w (3, 3, cin, cout) Keras layout. */
int msr_op_conv3x3_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n, int r, int cin,
                       int cout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOONSR_H_ */
