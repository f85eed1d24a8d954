/*
 * hyres_b200.h — C-ABI of the B200-native HyRES residual-codec hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point returns 0 on success or a negative HYRES_ERR_* code; no
 * exception crosses the boundary. Device pointers are raw CUDA device
 * addresses; `stream` is a cudaStream_t passed as void*.
 *
 * Each group of entry points cites the reference interface it replaces
 * (paths relative to the upstream repository, tmkhang1999/HyRES-...):
 *
 *   hyres_conv_*            nn.Conv2d / nn.ConvTranspose2d call sites of
 *                           models/checkerboard.py:35-88 (g_a, g_s, h_a, h_s,
 *                           context_prediction, param_aggregation),
 *                           models/layers/attention.py:16-47 (ResidualUnit, gate),
 *                           models/layers/enhancement.py:60-85 (MultiScaleRefine),
 *                           compressai GDN (conv2d(x^2, gamma, beta)) and
 *                           ResidualBottleneckBlock.
 *   hyres_residual_*        models/hyres.py:48,62,66-67,96,127,131-132
 *   hyres_gc_*              models/checkerboard.py:106-142,149-165 +
 *                           compressai GaussianConditional.{quantize,
 *                           _likelihood,build_indexes,dequantize}
 *   hyres_eb_*              compressai EntropyBottleneck.{forward,compress,
 *                           decompress} via models/checkerboard.py:96-101,172-173,206
 *   hyres_refine_*          models/layers/enhancement.py:15-21,36-40,87-112
 *   hyres_reduce_*          src/losses/rd_loss.py:23-26,39
 *   hyres_jpeg_*            models/utils/turbo_jpeg_compression.py:17-40,62-77
 *                           (TurboJPEG.encode / .decode of PyTurboJPEG 1.7.7 over
 *                           libjpeg-turbo; decoded pixels + file size)
 *   hyres_rans_*, hyres_pmf_to_quantized_cdf
 *                           compressai.ans.{RansEncoder.encode_with_indexes,
 *                           RansDecoder.decode_with_indexes} and
 *                           compressai._CXX.pmf_to_quantized_cdf, reached from
 *                           models/checkerboard.py:159-165,172-173,206,261-267
 */
#ifndef HYRES_B200_H
#define HYRES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYRES_OK 0
#define HYRES_ERR_ARG -1
#define HYRES_ERR_CUDA -2
#define HYRES_ERR_DRIVER -3
#define HYRES_ERR_UNSUPPORTED -4
#define HYRES_ERR_STATE -5

/* Library / device probes. hyres_device_check returns HYRES_OK only on an
 * sm_100 device; there is no CPU fallback behind any compute entry point. */
int hyres_version(void);
int hyres_device_check(int device);
const char* hyres_last_error(void);
/* Number of CUDA kernels this library has launched since it was loaded (host-side count;
 * launches replayed from a captured CUDA graph are counted once, at capture). */
long long hyres_launch_count(void);

/* ------------------------------------------------------------------------- */
/* Implicit-GEMM convolution on tcgen05 tensor cores (NHWC bf16, fp32 accum). */
/* ------------------------------------------------------------------------- */

typedef struct hyres_conv hyres_conv;

/* kind */
#define HYRES_CONV 0        /* nn.Conv2d, weight [Cout][Cin][R][S]                         */
#define HYRES_DECONV_K5S2 1 /* nn.ConvTranspose2d(k=5,s=2,p=2,output_padding=1),
                               weight [Cin][Cout][5][5] (compressai `deconv`)              */

/* epilogue modes: v = f(acc + bias) */
#define HYRES_EPI_LINEAR 0   /* v = acc + bias                                             */
#define HYRES_EPI_ADD 1      /* v = acc + bias + aux0           (residual skip)            */
#define HYRES_EPI_GATE 2     /* v = aux1 * sigmoid(acc+bias) + aux0  (AttentionBlock)      */
#define HYRES_EPI_GDN 3      /* v = aux0 * rsqrt(acc + bias)    (GDN; input is x^2)        */
#define HYRES_EPI_IGDN 4     /* v = aux0 * sqrt(acc + bias)     (inverse GDN)              */
#define HYRES_EPI_PIXSCALE 5 /* v = acc * pixscale[pixel] + bias (spatial attention)       */
#define HYRES_EPI_STATS 6    /* internal: hyres_refine_stats3_tc                           */

/* activation applied after the epilogue mode */
#define HYRES_ACT_NONE 0
#define HYRES_ACT_RELU 1
#define HYRES_ACT_PRELU 2   /* single-slope PReLU */
#define HYRES_ACT_CLAMP01 3

/* Create a layer: packs `weight` (host fp32, PyTorch layout) into the K-major
 * bf16 B operand, uploads it with `bias` (host fp32, may be NULL) and builds the
 * tap table. `w_cin_total` is the Cin extent of the weight array; the layer
 * consumes its first cin0+cin1 input channels (cin1 > 0: a second input tensor
 * is concatenated along channels). `tap_mask` (R*S bytes, may be NULL) drops
 * taps whose entry is 0 (checkerboard context conv). */
int hyres_conv_create(hyres_conv** out, int kind, int cin0, int cin1, int w_cin_total, int cout,
                      int R, int S, int stride, int pad, int dil, const float* weight,
                      const float* bias, const uint8_t* tap_mask);
/* Split-precision variant (fp32-equivalent arithmetic on the bf16 tensor cores) for the layers that decide
 * integer symbols and CDF indexes -- g_a, h_a, h_s, context_prediction, param_aggregation
 * (models/checkerboard.py:35-45,61-88), whose fp32 results the reference rounds at
 * models/checkerboard.py:159-165. nsplit = 1 is hyres_conv_create. nsplit = 2 / 3: the layer's inputs are bf16
 * NHWC tensors [B,H,W,nsplit*cin] holding nsplit bf16 parts of every fp32 activation (part p in channels
 * [p*cin,(p+1)*cin), see hyres_split_f32); the weights are packed as nsplit parts as well and part i of the
 * activations is multiplied with parts 0..nsplit-1-i of the weights (3 / 6 tensor-core products per MAC), all
 * accumulated in one fp32 TMEM tile. Such layers take HYRES_EPI_LINEAR (+ bias, ReLU) and write out_f32 only.
 * nsplit = 2 | HYRES_SPLIT_F16: the two parts are IEEE half values, p0 = half(v), p1 = half((v - p0) * 2^11)
 * (11 + 11 significand bits: v = p0 + p1 * 2^-11 to 2^-23 relative, |v| < 65504); the three products
 * p0*w0 | p0*w1 + p1*w0 go to separate accumulators and the epilogue adds the second one times 2^-11:
 * fp32-equivalent at half the tensor-core work of nsplit = 3. */
#define HYRES_SPLIT_F16 16
int hyres_conv_create_split(hyres_conv** out, int kind, int cin0, int cin1, int w_cin_total, int cout,
                            int R, int S, int stride, int pad, int dil, const float* weight,
                            const float* bias, const uint8_t* tap_mask, int nsplit);
/* Deployment export (the analogue of src/updata.py:50-78, which saves the CDF tables next to the weights): the
 * packed device operands of a layer -- K-major bf16 weights (which = 0), the tap-major copy the 3-output-channel
 * kernel reads (which = 1, 0 elements for other layers) and the padded fp32 bias (which = 2) -- can be copied
 * out once and copied back into a layer created with weight == NULL, so a deployment never re-packs. */
int64_t hyres_conv_packed_elems(const hyres_conv* c, int which);
int hyres_conv_export_packed(const hyres_conv* c, void* w_bf16, void* w_tap_bf16, float* bias);
int hyres_conv_import_packed(hyres_conv* c, const void* w_bf16, const void* w_tap_bf16, const float* bias);
/* Re-pack new weights into an existing layer (same geometry). */
int hyres_conv_update(hyres_conv* c, const float* weight, const float* bias);
/* The same from DEVICE fp32 tensors, packed by a kernel on `stream` (training: optimizer.step() changes the
 * weights every iteration, src/utils/engine.py:56-84; nothing crosses PCIe). bias_dev may be NULL. */
int hyres_conv_update_device(hyres_conv* c, const float* weight_dev, const float* bias_dev, void* stream);
void hyres_conv_destroy(hyres_conv* c);
/* MACs per output position the packed layer executes (padded K and N included). */
int64_t hyres_conv_macs_per_pos(const hyres_conv* c);

/* Weight gradient of the convolution `c` (training: loss.backward() through the nn.Conv2d / nn.ConvTranspose2d
 * modules of models/checkerboard.py:35-88, src/utils/engine.py:50-53) on the tensor cores:
 *   dw[co][ci][r][s] = sum_pos gout[pos][co] * x[pos*stride + tap][ci]      (fp32, PyTorch weight layout of `c`)
 * x: bf16 NHWC [B,H,W,cin] (the layer's input), gout: bf16 NHWC [B,OH,OW,cout] (gradient of its output).
 * Supported: HYRES_CONV layers with one input (any kernel / stride / dilation / tap mask the forward supports);
 * masked taps receive zero.  A transposed convolution's weight gradient is this call on its data-gradient
 * convolution (stride 2, same weight tensor) with x := the transposed layer's output gradient and gout := its
 * input.  `workspace`: hyres_wgrad_workspace_bytes() bytes of device memory. */
int hyres_wgrad_supported(const hyres_conv* c);
int64_t hyres_wgrad_workspace_bytes(const hyres_conv* c, int B, int H, int W);
int hyres_wgrad_run(hyres_conv* c, const void* x, const void* gout, int B, int H, int W, float* dw, void* workspace,
                    void* stream);

/* Bias gradient db[c] = sum over rows of g[row][c] (g: bf16 [rows][C], the NHWC output gradient of a layer;
 * C a multiple of 8): deterministic two-pass column sums. workspace: hyres_colsum_workspace_bytes() bytes. */
int64_t hyres_colsum_workspace_bytes(int64_t rows, int C);
int hyres_colsum_bf16(const void* g, int64_t rows, int C, float* out, void* workspace, void* stream);

typedef struct {
  const void* x0; /* bf16 NHWC [B,H,W,cin0] */
  const void* x1; /* bf16 NHWC [B,H,W,cin1] or NULL */
  int B, H, W;    /* input extent */
  int epi, act;
  float slope;
  const void* aux0; /* bf16 NHWC at output resolution, channel stride ld_aux0 */
  int ld_aux0;
  const void* aux1;
  int ld_aux1;
  const float* pixscale; /* fp32 [B,OH,OW] */
  void* out_bf16;        /* bf16 NHWC, channel stride ld_out (>= cout), may be NULL */
  int ld_out;
  void* out_sq; /* bf16 NHWC of v*v (GDN operand), may be NULL */
  int ld_sq;
  float* out_f32; /* fp32, arbitrary strides (elements), may be NULL */
  int64_t f32_sb, f32_sh, f32_sw, f32_sc;
  int mt_hint; /* 0 = auto; else force 1/2/4 sub-tiles per CTA */
  int ld_x0;   /* channel stride of x0 in elements; 0 = dense (cin0) */
  int x0_square; /* 1: the layer consumes x0*x0 (GDN: conv2d(x^2, gamma, beta)), squared on chip;
                    1x1 layers only */
  int out_pad;   /* > 0: out_bf16 points at a tensor [B,OH+2p,OW+2p,ld_out] and the result is stored
                    into its interior (the caller fills the border, hyres_replicate_border) */
  /* up-add (MultiScaleRefine fusion, models/layers/enhancement.py:101-110): acc += bilinear x2
   * up-sampling of up_t2 + bilinear x4 up-sampling of up_t3 before the epilogue. Both are bf16 NHWC,
   * 64 channels, padded by one replicated pixel: [B,OH/2+2,OW/2+2,64] and [B,OH/4+2,OW/4+2,64].
   * The interpolation runs on the tensor cores (two small GEMMs per tile). NULL = off. */
  const void* up_t2;
  const void* up_t3;
  int cta_limit; /* > 0: launch at most this many CTAs (the persistent kernels use one per SM); lets two
                    independent layers -- the branches of an AttentionBlock -- share the GPU on two streams */
  /* Split-precision layers only (hyres_conv_create_split, nsplit > 1): the fp32 element-wise stage that follows
   * the convolution, fused into its epilogue.  v = split_mode(acc + bias, aux0_f32, aux1_f32) (HYRES_SPLIT_COPY /
   * _ADD / _GATE / _GDN / _IGDN, see hyres_split_f32), then `act` (none / ReLU).  v goes to out_f32 (optional) and,
   * as out_nsplit bf16 parts, to out_split [B,OH,OW,out_nsplit*cout] (optional); split_square != 0 stores the
   * parts of v*v instead (the x^2 operand of the GDN that follows).  aux*: fp32 NHWC [B,OH,OW,cout], dense.
   * out_nsplit carries its own format (1..3, or 2 | HYRES_SPLIT_F16): the consumer of the parts may be a layer
   * of the other format (squares keep bf16 parts: half parts would overflow at |v| >= 256). */
  int split_mode;
  const float* aux0_f32;
  const float* aux1_f32;
  void* out_split;
  int out_nsplit;
  int split_square;
} hyres_conv_io;

int hyres_conv_out_size(const hyres_conv* c, int H, int W, int* OH, int* OW);
int hyres_conv_run(hyres_conv* c, const hyres_conv_io* io, void* stream);

/* Fused bottleneck residual unit at C = 128 (one persistent kernel instead of three
 * convolution launches; the two 64-channel intermediates never leave the SM):
 *   out = [ReLU](x + c3(ReLU(c2(ReLU(c1(x))))))
 * c1: 1x1 128->64, c2: 3x3 64->64 (stride 1, pad 1), c3: 1x1 64->128, created with
 * hyres_conv_create. Replaces ResidualUnit.forward (models/layers/attention.py:16-33,
 * final_relu = 1) and compressai ResidualBottleneckBlock.forward
 * (models/checkerboard.py:38,42,51,55, final_relu = 0). */
typedef struct {
  const void* x; /* bf16 NHWC [B,H,W,128], channel stride ld_x */
  int ld_x;
  void* out;     /* bf16 NHWC [B,H,W,128], channel stride ld_out; must not alias x */
  int ld_out;
  int B, H, W;
  int final_relu;
} hyres_ru_io;
int hyres_ru_supported(const hyres_conv* c1, const hyres_conv* c2, const hyres_conv* c3);
int hyres_ru_run(const hyres_conv* c1, const hyres_conv* c2, const hyres_conv* c3,
                 const hyres_ru_io* io, void* stream);

/* ------------------------------------------------------------------------- */
/* Memory-bound kernels                                                       */
/* ------------------------------------------------------------------------- */

/* The two 3-channel first-layer convolutions fused with the residual arithmetic around them
 * (no im2col tensor in HBM; the im2col row is built on chip):
 *   src = a + sign * b        (fp32 NCHW [B,3,H,W]; b may be NULL)
 *   sum_out = src             (optional fp32 NCHW; requires b)
 *   out = act(conv(src) + bias)   bf16 NHWC [B,OH,OW,cout], channel stride ld_out
 * ksize 5 / stride 2 (g_a.0 on residual = x - jpeg: models/hyres.py:48,96 +
 * models/checkerboard.py:36, cout 128) or ksize 3 / stride 1 (refine.conv_in + PReLU on
 * x0 = jpeg + r_hat: models/hyres.py:62,127 + models/layers/enhancement.py:60,89, cout 64).
 * `c` is the layer created as the 1x1 GEMM over the im2col ordering k = (r*ksize+s)*3 + c
 * (cin 128 / 64, zero beyond 75 / 27). */
int hyres_conv3ch_run(hyres_conv* c, int ksize, int stride, const float* a, const float* b, int sign,
                      float* sum_out, void* out_bf16, int ld_out, int B, int H, int W, int act,
                      float slope, void* stream);
/* x_hat = clamp(x0 + refined, 0, 1), all fp32 NCHW. models/hyres.py:66-67,131-132 */
int hyres_final_clamp(const float* x0, const float* refined, float* x_hat, int64_t n, void* stream);

/* GaussianConditional / checkerboard quantiser. y: fp32 NHWC [B,h,w,M];
 * params: fp32 NHWC [B,h,w,2M] with scales in channels [0,M) and means in
 * [M,2M) (models/checkerboard.py:118). pass 0 = anchor, 1 = non-anchor.
 * mode 0: STE/round (eval), mode 1: additive uniform noise (seeded Philox). */
int hyres_gc_quant_pass(const float* y, const float* params, int pass, int mode, uint64_t seed,
                        float* yq_f32, void* yq_bf16, int B, int h, int w, int M, void* stream);
/* y_hat = yq_a + yq_na (bf16 NHWC out); likelihoods of `y` under
 * (scales_a+scales_na, means_a+means_na) written fp32 NCHW [B,M,h,w];
 * partial sums of log2(lik) accumulated (deterministically) into *sum_log2
 * (double, device). models/checkerboard.py:136-142 */
int hyres_gc_merge_likelihood(const float* y, const float* params_a, const float* params_na,
                              const float* yq_a, const float* yq_na, int mode, uint64_t seed,
                              void* y_hat_bf16, float* lik_nchw, double* sum_log2, int B, int h,
                              int w, int M, void* stream);
/* symbols = int32(round(y_part - mean)), indexes = bucketize(max(scale, bound)),
 * both written in (B,M,h,w) order; yq = symbol + mean (dequantize).
 * models/checkerboard.py:159-161 */
int hyres_gc_symbols(const float* y, const float* params, int pass, const float* scale_table,
                     int n_scales, float scale_bound, int32_t* symbols, int32_t* indexes,
                     float* yq_f32, void* yq_bf16, int B, int h, int w, int M,
                     const int32_t* coder_rows, int32_t* slots, void* stream);
/* coder_rows / slots (both or neither): the host coder's table layout (hyres_rans_table_layout, on the device) and,
 * in the indexes' order, the coder SLOT of every symbol (see hyres_rans_encode_slots_batch).
 * hyres_gc_codes is the decoder's counterpart for checkerboard pass `pass`: indexes (optional) and the decoder CODES --
 * at the structurally zero positions of the pass the symbol is round(-mean) and the code carries its packed entry
 * (hyres_rans_decode_codes_batch); hyres_gc_dequant with the same `pass` then recomputes those symbols instead of
 * reading them (pass = -1: every symbol is read). */
int hyres_gc_codes(const float* params, int pass, const float* scale_table, int n_scales, float scale_bound,
                   const int32_t* coder_rows, int32_t* indexes, int32_t* codes, int B, int h, int w, int M,
                   void* stream);
/* decoder side: indexes from scales only. models/checkerboard.py:163-165 */
int hyres_gc_indexes(const float* params, const float* scale_table, int n_scales,
                     float scale_bound, int32_t* indexes, int B, int h, int w, int M, void* stream);
/* decoder side: yq = float(symbol) + mean; symbols in (B,M,h,w) order. */
int hyres_gc_dequant(const int32_t* symbols, const float* params, float* yq_f32, void* yq_bf16,
                     int B, int h, int w, int M, int pass, void* stream);
/* y_hat = a + b (fp32 NHWC in, bf16 NHWC out). models/checkerboard.py:234 */
int hyres_add_to_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------- */
/* Split-precision trunk: fp32 element-wise work between two split convolutions */
/* ------------------------------------------------------------------------- */
#define HYRES_SPLIT_COPY 0       /* v = in                                                      */
#define HYRES_SPLIT_ADD 1        /* v = in + aux0           (ResidualUnit / RBB skip)           */
#define HYRES_SPLIT_GATE 2       /* v = aux1 * sigmoid(in) + aux0  (models/layers/attention.py:44-47) */
#define HYRES_SPLIT_GDN 3        /* v = aux0 * (1 / sqrt(in))      (compressai GDN)             */
#define HYRES_SPLIT_IGDN 4       /* v = aux0 * sqrt(in)                                         */
#define HYRES_SPLIT_SQUARE 5     /* v = in * in             (the GDN operand x^2)               */
#define HYRES_SPLIT_ROUND_CHAN 6 /* v = round(in - chan[c]) + chan[c]  (models/checkerboard.py:99-101) */
/* in, aux0, aux1: fp32 [rows][C] (NHWC rows); chan: fp32 [C]. v (then ReLU if relu != 0) is written as fp32
 * (out_f32, optional) and as nsplit bf16 parts [rows][nsplit*C] (out_split, optional): part 0 = bf16(v),
 * part 1 = bf16(v - part 0), part 2 = bf16(v - part 0 - part 1); nsplit = 2 | HYRES_SPLIT_F16: two half parts,
 * half(v) and half((v - part 0) * 2^11) (see hyres_conv_create_split). IEEE fp32 arithmetic throughout. */
int hyres_split_f32(const float* in, int64_t rows, int C, int mode, const float* aux0, const float* aux1,
                    const float* chan, int relu, float* out_f32, void* out_split, int nsplit, void* stream);
/* residual = x - jpeg (fp32 NCHW; jpeg may be NULL: x is the residual) and the 5x5/stride-2 im2col of the
 * residual as nsplit bf16 parts [B,H/2,W/2,nsplit*128] (k = (r*5+s)*3+c within a part, 75 live): g_a.0 as a
 * split 1x1 GEMM. models/hyres.py:48,96 + models/checkerboard.py:36 */
int hyres_residual_im2col5s2_split(const float* x, const float* jpeg, float* residual, void* a_out, int nsplit,
                                   int B, int H, int W, void* stream);
/* out[b,i,j,c] = float(symbols[b,c,i,j]) + chan[c] (chan may be NULL): decoded integer symbols in the coder's
 * (B,C,h,w) order -> fp32 NHWC (EntropyBottleneck.dequantize, compressai; models/checkerboard.py:206). */
int hyres_symbols_to_nhwc_f32(const int32_t* symbols, const float* chan, float* out, int B, int h, int w, int C,
                              void* stream);

/* EntropyBottleneck. z: fp32 NHWC [B,h,w,C]. `eb_params`: fp32 [C][58] =
 * softplus(matrices) (3,9,9,9,3), biases (3,3,3,3,1), tanh(factors) (3,3,3,3);
 * medians fp32 [C]. mode 0: dequantize (round about the median), 1: noise.
 * z_hat written bf16 NHWC (+ fp32 NCHW optional), likelihood fp32 NCHW,
 * symbols int32 (B,C,h,w) optional. */
int hyres_eb_forward(const float* z, const float* eb_params, const float* medians, int mode,
                     uint64_t seed, float lik_bound, void* zhat_bf16, float* zhat_nchw,
                     float* lik_nchw, int32_t* symbols, double* sum_log2, int B, int h, int w,
                     int C, void* stream);
/* z_hat = float(symbol) + median, symbols (B,C,h,w) -> bf16 NHWC. */
int hyres_eb_dequant(const int32_t* symbols, const float* medians, void* zhat_bf16, int B, int h,
                     int w, int C, void* stream);

/* MultiScaleRefine memory ops (feat: bf16 NHWC, C=64). */
/* scratch: fp32 [B*64*C] partial sums (two-stage, fixed-order reduction). */
int hyres_refine_se_pool(const void* feat, float* scratch, float* pooled /*[B,C]*/, int B, int H,
                         int W, int C, void* stream);
int hyres_refine_se_scale_down(const void* feat, const float* pooled, const float* fc1,
                               const float* fc2, int C, int Cr, void* feat_s, void* feat_h,
                               void* feat_q, int B, int H, int W, void* stream);
/* Channel mean / max over the virtual concat [f1 | up2(s2) | up4(s3)] (models/layers/enhancement.py:101-106,
 * 15-21) without materialising it: s2 / s3 are bf16 NHWC tensors padded by one replicated
 * pixel, [B,H/2+2,W/2+2,64] / [B,H/4+2,W/4+2,64]; the bilinear up-samplings are GEMMs with a constant
 * interpolation matrix, the epilogue reduces over the channels. H, W multiples of 32. */
int hyres_refine_stats3_tc(const void* f1, const void* s2_padded, const void* s3_padded, float* stats,
                           int B, int H, int W, void* stream);
/* Fill the one-pixel border of a padded bf16 NHWC tensor [B,Hp,Wp,C] with the nearest interior pixel. */
int hyres_replicate_border(void* t, int B, int Hp, int Wp, int C, void* stream);
int hyres_refine_spatial_att(const float* stats, const float* w7x7, float* att /*[B,H,W]*/,
                             int B, int H, int W, void* stream);

/* ------------------------------------------------------------------------- */
/* JPEG stage on the device. Replaces TurboJPEGCompression.forward / .compress  */
/* (models/utils/turbo_jpeg_compression.py:17-40,62-77: TurboJPEG.encode with  */
/* PyTurboJPEG's defaults -- RGB array read as BGR, 4:2:2, baseline Huffman --  */
/* then TurboJPEG.decode) with libjpeg-turbo's integer algorithm reproduced bit */
/* for bit: decoded pixels and file sizes are identical to the library's.       */
/* ------------------------------------------------------------------------- */

/* Scratch bytes hyres_jpeg_forward needs (0 if the size is unsupported: H % 8, W % 16). */
int64_t hyres_jpeg_workspace_bytes(int B, int H, int W);
/* 32-bit words per image of the scan buffer (worst case of baseline Huffman). */
int64_t hyres_jpeg_scan_words(int H, int W);
/* Bytes of everything before the scan (SOI, APP0, 2 DQT, SOF0, 4 DHT, SOS): 623. */
int hyres_jpeg_header_bytes(void);
/* x: fp32 NCHW [B,3,H,W] in [0,1] (clamped, then `.byte()`-truncated like the reference).
 * decoded (optional): fp32 NCHW [B,3,H,W] = decoded u8 / 255.
 * scan_words (optional, device, [B][hyres_jpeg_scan_words]): the entropy-coded scan of every image
 * as big-endian bit strings, without byte stuffing; scan_bits (device, [B]): their lengths in bits;
 * sizes (optional, device, [B]): the size in bytes of the JPEG file of every image. */
int hyres_jpeg_forward(const float* x, int B, int H, int W, int quality, void* workspace,
                       float* decoded, int64_t* sizes, uint32_t* scan_words, int64_t* scan_bits,
                       void* stream);
/* bpp[0] = 8 * sum(sizes[0..B)) / pixels as fp32 (device): `jpeg_bpp` of TurboJPEGCompression.forward
 * (models/utils/turbo_jpeg_compression.py:66-71) without a host round trip. */
int hyres_jpeg_bpp(const int64_t* sizes, int B, int64_t pixels, float* bpp, void* stream);
/* Host: the complete JPEG file of one image (markers + stuffed scan + EOI) from its scan bits. */
int hyres_jpeg_assemble(const uint32_t* scan_words_host, int64_t nbits, int H, int W, int quality,
                        uint8_t* out, int64_t cap, int64_t* len);

/* layout helpers */
int hyres_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W,
                                void* stream);
int hyres_nhwc_to_nchw_f32(const float* in, float* out, int B, int C, int H, int W, void* stream);
int hyres_nhwc_bf16_to_nchw_f32(const void* in, float* out, int B, int C, int H, int W,
                                void* stream);

/* sum((a-b)^2) and sum(log2(x)) accumulated into a device double. */
int hyres_reduce_sqdiff(const float* a, const float* b, int64_t n, double* out, void* stream);
int hyres_reduce_log2(const float* x, int64_t n, double* out, void* stream);
/* RateDistortionLoss.forward (src/losses/rd_loss.py:23-44, alpha = 0) from the three device sums:
 * out6 = [y_bpp, z_bpp, residual_bpp, bpp, mse * 255^2, lambda * mse + bpp] (fp32, device).
 * jpeg_bpp: device fp32 scalar or NULL. */
int hyres_rd_loss_finalize(const double* sum_log2_y, const double* sum_log2_z, const double* sum_sq_err,
                           const float* jpeg_bpp, double num_pixels, double num_elems, float lmbda,
                           float* out6, void* stream);

/* ------------------------------------------------------------------------- */
/* Entropy coder (host). Byte-identical to compressai's rANS64 interface.      */
/* ------------------------------------------------------------------------- */

/* compressai._CXX.pmf_to_quantized_cdf: out has n+1 entries. */
int hyres_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* out);

/* RansEncoder.encode_with_indexes. cdfs: [n_cdfs][cdf_stride] int32. Writes at most
 * out_cap bytes; *out_len receives the length (HYRES_ERR_ARG if out_cap too small;
 * 4*(n_symbols_incl_bypass)+8 always suffices, see hyres_rans_encode_bound). */
int64_t hyres_rans_encode_bound(int64_t n);
int hyres_rans_encode(const int32_t* symbols, const int32_t* indexes, int64_t n,
                      const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                      const int32_t* offsets, uint8_t* out, int64_t out_cap, int64_t* out_len);
/* RansDecoder.decode_with_indexes. */
int hyres_rans_decode(const uint8_t* in, int64_t in_len, const int32_t* indexes, int64_t n,
                      const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                      const int32_t* offsets, int32_t* symbols_out);
/* Batched variants: `count` independent strings coded on up to `threads` host threads. */
int hyres_rans_encode_batch(int count, const int32_t* const* symbols, const int32_t* const* indexes,
                            const int64_t* n, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* const* out,
                            const int64_t* out_cap, int64_t* out_len, int threads);
int hyres_rans_decode_batch(int count, const uint8_t* const* in, const int64_t* in_len,
                            const int32_t* const* indexes, const int64_t* n, const int32_t* cdfs,
                            int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, int32_t* const* symbols_out, int threads);
/* Device-side coder front-end ("GPU emits compact streams per (image, pass)").
 * hyres_rans_table_layout: rows_out[3 * n_cdfs] = for every CDF row the first entry of the row in the coder's packed
 * tables, its offset, and its escape bin (number of in-table values; 0 for an unusable row).  With these three vectors
 * on the device, hyres_gc_symbols emits per symbol a SLOT (>= 0: packed entry = base + value of an in-table value;
 * < 0: -(row + 1), the value is outside the table and is coded through the escape path from symbols[i]) and
 * hyres_gc_codes emits per symbol a CODE for the decoder (bit 30 set: the symbol is known to the caller -- a
 * structurally zero position of the checkerboard pass, models/checkerboard.py:106-110 -- low bits = its packed entry:
 * the decoder only advances the range-coder state and leaves symbols_out[i] untouched; otherwise a plain row index).
 * The byte strings are identical to hyres_rans_encode_batch / consumed exactly like hyres_rans_decode_batch. */
int hyres_rans_table_layout(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, int32_t* rows_out);
int hyres_rans_encode_slots_batch(int count, const int32_t* const* symbols, const int32_t* const* slots,
                                  const int64_t* n, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                                  const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* const* out,
                                  const int64_t* out_cap, int64_t* out_len, int threads);
int hyres_rans_decode_codes_batch(int count, const uint8_t* const* in, const int64_t* in_len,
                                  const int32_t* const* codes, const int64_t* n, const int32_t* cdfs, int n_cdfs,
                                  int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                                  int32_t* const* symbols_out, int threads);

/* ------------------------------------------------------------------------- */
/* Entropy coder (device): the same byte strings, coded on the GPU.           */
/* ------------------------------------------------------------------------- */
/* A string is one dependency chain, so each is coded by one warp (csrc/rans_dev.cu); dozens of strings (all images in
 * flight) run beside the convolution kernels and no host core is involved -- compress + decompress then scale with the
 * number of GPUs of a box instead of with its host cores.  Same format as hyres_rans_encode / hyres_rans_decode
 * (RansEncoder.encode_with_indexes / RansDecoder.decode_with_indexes of compressai 1.2.6).
 * hyres_rans_table_entries / _export: the host coder's packed tables for upload -- enc_out: 16 bytes per entry,
 * sf_out: one uint32 per entry, rows_out: [4][n_cdfs] = first entry, offset, escape bin, usable (0 / 1).
 * hyres_rans_dev_encode: ONE launch for up to four groups (e.g. the z strings and both y passes of a batch), a group
 * = `count` strings of n symbols with one table set; one warp per string, all warps in as few blocks as possible.
 * scratch: n / 2 + 192 words per string always suffice without escapes; 3 n + 192 with.  The strings land back to back,
 * in no particular order, in dst; meta (int32, zeroed by the caller before the first launch that shares it): [0] words
 * used in dst, [1] status (0 ok, 1 bad table / symbol, 2 buffer too small), [2 + 2 k] first word and [3 + 2 k] word
 * count of string k = meta_base + (index of the string in this launch, groups in order).
 * hyres_set_reserved_sms(n): every persistent kernel of the library sizes its grid to (SMs - n) from now on (returns
 * the previous value).  The convolution kernels take a whole SM per CTA with a fixed share of the tiles each, so a CTA
 * that has to wait for an SM held by a (long-running) coder block delays its whole launch: a pipeline that keeps n
 * coder launches resident reserves n SMs for them.
 * hyres_rans_dev_decode: words = all strings in one device buffer, str_off / str_len = first word and word count of
 * each; codes = [count][n] CDF rows, or decoder CODES from hyres_gc_codes when has_codes != 0 (symbols_out is left
 * untouched at known symbols); *status (device int32, zeroed by the caller) receives 1 for a malformed stream. */
int64_t hyres_rans_table_entries(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                                 const int32_t* offsets);
int hyres_rans_table_export(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, void* enc_out, uint32_t* sf_out, int32_t* rows_out);
typedef struct hyres_rans_group {
  const int32_t* symbols;  /* [count][n] device int32 (may be NULL with slots when no slot is negative) */
  const int32_t* index;    /* [count][n] CDF row per symbol, or coder SLOTS (hyres_gc_symbols) when slots != 0 */
  const void* enc;         /* packed encoder entries (hyres_rans_table_export), 16 bytes each */
  const int32_t* rows;     /* [4][n_rows] */
  uint32_t* scratch;       /* [count][cap_words] working space */
  int64_t n, cap_words, n_entries;
  int32_t n_rows, count, slots;
} hyres_rans_group;
int hyres_rans_dev_encode(int n_groups, const hyres_rans_group* groups, uint32_t* dst, int64_t dst_cap_words,
                          int32_t* meta, int meta_base, void* stream);
int hyres_set_reserved_sms(int n);
int hyres_rans_dev_decode(const uint32_t* words, const int64_t* str_off, const int64_t* str_len, const int32_t* codes,
                          int count, int64_t n, int has_codes, const uint32_t* sf, const int32_t* rows, int n_rows,
                          int64_t n_entries, int32_t* symbols_out, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYRES_B200_H */
