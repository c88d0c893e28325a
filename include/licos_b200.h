/*
 * licos_b200 -- C ABI of the B200-native learned-codec hot path.
 *
 * The reference (gomezzz/LICOS) has no FFI of its own: it drives the path through the CompressAI
 * nn.Module API.  Every entry point below therefore names the upstream
 * CompressAI call it replaces and the reference line that reaches it.  The Python host layer
 * (the licos_b200 Python package) mirrors the CompressAI module API on top of these calls.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; device pointers unless the parameter says "host";
 *   - returns LICOS_OK (0) or a negative LICOS_ERR_* code; never throws, never allocates device
 *     memory, never synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - all buffers are caller-owned and must stay alive until the stream has drained.
 */
#ifndef LICOS_B200_H
#define LICOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LICOS_ABI_VERSION 3 /* 3: licos_conv_args.pre_act, licos_conv_wgrad_image */

enum {
    LICOS_OK = 0,
    LICOS_ERR_INVALID = -1,     /* bad argument (shape, null pointer, unsupported combination) */
    LICOS_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed: see licos_last_cuda_error() */
    LICOS_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for */
    LICOS_ERR_NO_DEVICE = -4,   /* no sm_100 device */
    LICOS_ERR_DOMAIN = -5,      /* pmf_to_quantized_cdf: negative / non-finite / all-zero pmf */
    LICOS_ERR_NOMEM = -6,       /* host allocation failed */
    LICOS_ERR_BUFFER = -7       /* caller's output buffer too small */
};

int licos_abi_version(void);
const char* licos_strerror(int code);
int licos_last_cuda_error(void); /* cudaError_t of the last LICOS_ERR_CUDA on this thread */
/* LICOS_OK when `device` is a compute-capability 10.x GPU (B200), else LICOS_ERR_NO_DEVICE. */
int licos_device_ok(int device);

/* ------------------------------------------------------------------------------------------ */
/* Tensor layouts                                                                             */
/* ------------------------------------------------------------------------------------------ */
#define LICOS_LAYOUT_NCHW_F32 0  /* the CompressAI-facing layout (x, y, x_hat, likelihoods)     */
#define LICOS_LAYOUT_NHWC_BF16 1 /* inter-layer activations; C must be a multiple of 64         */
/* Integer pixel tiles (N4; /root/reference/licos/raw_image_folder.py:192-196 scales DN / DN_MAX on the host and ships
 * fp32): the first layer reads them directly and scales in its patch builders, the last layer can write them.
 *   input : x = v / int_max, bit-identical (after the bf16 operand rounding) to feeding float32(float64(v) / int_max);
 *           U16_Q8 adds the reference's default 8-bit step, x = rint(v / int_max * 255) / 255 (img_as_ubyte, :195).
 *   output: v = rint(clamp(x_hat, 0, 1) * int_max).
 * int_max = licos_conv_args.int_max, 0 = 255 (U8) / 4095 (U16: the 12-bit DN_MAX of raw_utils.py:128). */
#define LICOS_LAYOUT_NCHW_U8 2
#define LICOS_LAYOUT_NCHW_U16 3
#define LICOS_LAYOUT_NCHW_U16_Q8 4 /* input only */
/* Host-only: 1 when the single-multiply scaling the first layer applies to integer pixels reproduces
 * float32(float64(v) / int_max) (requant8: float32(rint(v / int_max * 255) / 255)) after the bf16 operand rounding for
 * EVERY v in [0, int_max] (checked exhaustively); licos_conv_forward refuses any other int_max. */
int licos_pixel_scale_exact(int int_max, int requant8);

/* fp32 NCHW -> bf16 NHWC (optionally |x|: ScaleHyperprior.forward's `h_a(torch.abs(y))`). */
int licos_nchw_f32_to_nhwc_bf16(const float* in, void* out, int batch, int channels, int64_t hw,
                                int take_abs, void* stream);
int licos_nhwc_bf16_to_nchw_f32(const void* in, float* out, int batch, int channels, int64_t hw,
                                void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Convolutions (replaces torch.nn.Conv2d / ConvTranspose2d / compressai.layers.GDN forward    */
/* inside g_a / g_s / h_a / h_s; reached from licos/train.py:190 and eval_utils.py:200-201)    */
/* ------------------------------------------------------------------------------------------ */
#define LICOS_CONV_5X5_S2 0   /* compressai.models.utils.conv:   Conv2d(k=5, s=2, p=2)          */
#define LICOS_DECONV_5X5_S2 1 /* compressai.models.utils.deconv: ConvTranspose2d(k=5,s=2,p=2,op=1) */
#define LICOS_CONV_3X3_S1 2   /* conv(.., stride=1, kernel_size=3): Conv2d(k=3, s=1, p=1)       */
#define LICOS_CONV_1X1 3      /* Conv2d(k=1): the GDN channel mix as its own layer (training path) */

#define LICOS_EPI_NONE 0 /* + bias                                                             */
#define LICOS_EPI_GDN 1  /* + bias, then GDN:  v * rsqrt(beta + gamma . v^2)                    */
#define LICOS_EPI_IGDN 2 /* + bias, then IGDN: v *  sqrt(beta + gamma . v^2)                    */
#define LICOS_EPI_RELU 3 /* + bias, then max(v, 0)                                             */

/* Bytes needed for the packed bf16 weight of one layer (see licos_pack_conv_weight). */
int64_t licos_packed_weight_bytes(int kind, int out_c, int in_c, int in_layout);

/* Re-packs a torch-layout fp32 weight for the kernels.
 *   kind = CONV_*  : w is (out_c, in_c, KH, KW)   [torch.nn.Conv2d.weight]
 *   kind = DECONV_*: w is (in_c, out_c, KH, KW)   [torch.nn.ConvTranspose2d.weight]
 * in_layout selects the kernel family the weight is for: NHWC_BF16 -> [tap][out_c_pad][in_c_pad64],
 * NCHW_F32 (first layer, in_c <= 16) -> [out_c][K_pad64] with K = (c, kh, kw). */
int licos_pack_conv_weight(const float* w, int kind, int out_c, int in_c, int in_layout, void* packed,
                           void* stream);

/* GDN parameters after NonNegativeParametrizer: beta_hat = max(beta, beta_bound)^2 - pedestal (fp32),
 * gamma_hat = max(gamma, gamma_bound)^2 - pedestal (bf16, [C][C] row = output channel). */
int licos_gdn_pack(const float* beta, const float* gamma, int channels, float beta_bound, float gamma_bound,
                   float pedestal, float* beta_hat, void* gamma_hat_bf16, void* stream);

typedef struct licos_conv_args {
    int kind;       /* LICOS_CONV_* / LICOS_DECONV_*                                           */
    int epilogue;   /* LICOS_EPI_*                                                             */
    int batch;
    int in_h, in_w; /* spatial size of `in`                                                    */
    int in_c;       /* logical input channels                                                  */
    int out_c;      /* logical output channels                                                 */
    int in_layout;  /* LICOS_LAYOUT_*                                                          */
    int out_layout; /* LICOS_LAYOUT_*                                                          */
    const void* in;
    void* out;
    const void* weight; /* from licos_pack_conv_weight                                         */
    const float* bias;  /* [out_c] fp32, may be NULL                                           */
    const float* beta;  /* [out_c] fp32 from licos_gdn_pack  (EPI_GDN / EPI_IGDN)              */
    const void* gamma;  /* [out_c][out_c] bf16 from licos_gdn_pack                             */
    void* workspace;    /* scratch for in_layout == NCHW_F32 (licos_conv_workspace_bytes)        */
    int64_t workspace_bytes;
    int sm_count;       /* 0 = query the device                                                */
    int int_max;        /* integer pixel layouts: full-scale value (0 = the layout's default)   */
    void* pre_act;      /* optional, EPI_GDN / EPI_IGDN only: also write v = conv + bias (the GDN's input,
                           bf16 NHWC [batch][out_h][out_w][out_c]) -- what the backward pass keeps
                           (train.py:193); NULL in inference                                   */
} licos_conv_args;

/* Scratch bytes licos_conv_forward needs for these args (0 unless in_layout == NCHW_F32). */
int64_t licos_conv_workspace_bytes(const licos_conv_args* args);

/* One conv / deconv layer with its fused epilogue.  Output spatial size:
 * CONV_5X5_S2 -> ceil(in/2), DECONV_5X5_S2 -> 2*in, CONV_3X3_S1 -> in. */
int licos_conv_forward(const licos_conv_args* args, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Backward pass of the transforms (torch autograd of Conv2d / ConvTranspose2d / GDN / ReLU;  */
/* reached from licos/train.py:193 `out_criterion["loss"].backward()`)                         */
/* ------------------------------------------------------------------------------------------ */
/* DATA gradients need no entry point of their own: the gradient of CONV_5X5_S2 with respect to its input is
 * DECONV_5X5_S2 over the output gradient with the SAME (out_c, in_c, 5, 5) weight read as a ConvTranspose2d weight, and
 * vice versa; CONV_3X3_S1 takes the flipped, transposed weight.  They run through licos_conv_forward. */

typedef struct licos_wgrad_args {
    int kind;        /* CONV_5X5_S2 / DECONV_5X5_S2 (`big` has twice the resolution of `small`), CONV_3X3_S1, CONV_1X1 */
    int batch;
    int h, w;         /* spatial size of `small_t`                                                  */
    int big_h, big_w; /* spatial size of `big_t`: h == ceil(big_h / 2) for the stride-2 kinds       */
    int small_c, big_c; /* channels, multiples of 64 (CONV_1X1: big_c any multiple of 16)           */
    const void* small_t; /* bf16 NHWC: Conv2d -> the OUTPUT gradient; ConvTranspose2d -> the layer INPUT */
    const void* big_t;   /* bf16 NHWC: Conv2d -> the layer INPUT; ConvTranspose2d -> the OUTPUT gradient */
    float* out;       /* fp32 [KH*KW][small_c][big_c], 16-byte aligned, ACCUMULATED into with red.add: the caller zeroes it */
    int sm_count;     /* 0 = query the device                                                      */
    int reserved;
} licos_wgrad_args;
/* out[kh*KW + kw][cs][cb] += sum_{b,i,j} small[b][i][j][cs] * big[b][s*i + kh - pad][s*j + kw - pad][cb]
 * (s = 2, pad = 2 for the 5x5 kinds; s = 1, pad = 1 for 3x3; a plain [P x cs]^T [P x cb] product for 1x1).
 * Conv2d.weight.grad[o][i][kh][kw] = out[..][o][i]; ConvTranspose2d.weight.grad[i][o][kh][kw] = out[..][i][o]. */
int licos_conv_wgrad(const licos_wgrad_args* args, void* stream);

/* The whole GDN / IGDN backward in ONE pass over x and g (channels == 128; other widths return LICOS_ERR_UNSUPPORTED and
 * take the sequence below): x = the layer's pre-activation, g = the gradient of its output, both bf16 [n_pixels][channels];
 * gamma_hat_bf16 / beta_hat from licos_gdn_pack.  Writes dx (bf16) and ACCUMULATES d_gamma_hat [C][C], d_beta_hat [C] and,
 * when non-NULL, d_bias [C] (the column sums of dx = the preceding conv's bias.grad); the caller zeroes them
 * (d_gamma_hat 16-byte aligned: vector red.add). */
int licos_gdn_backward(const void* x, const void* g, const void* gamma_hat_bf16, const float* beta_hat, int inverse,
                       int64_t n_pixels, int channels, void* dx, float* d_gamma_hat, float* d_beta_hat, float* d_bias,
                       int sm_count, void* stream);

/* GDN backward, elementwise parts over n bf16 elements (n % 8 == 0); the two channel mixes between them are
 * LICOS_CONV_1X1 layers:  x2 = x^2;  norm = conv1x1(x2, gamma_hat, beta_hat);
 *   mid:  d_norm = g * dy/dnorm, d_direct = g * dy/dx|norm   (inverse != 0: IGDN)
 *   t = conv1x1(d_norm, gamma_hat^T);  out:  dx = d_direct + 2 x t  (dx may alias d_direct)
 *   gamma_hat.grad = licos_conv_wgrad(CONV_1X1, small = d_norm, big = x2);  beta_hat.grad = licos_colsum_bf16(d_norm) */
int licos_square_bf16(const void* x, void* x2, int64_t n, void* stream);
int licos_gdn_bwd_mid(const void* x, const void* g, const void* norm, int inverse, int64_t n, int channels, void* d_norm,
                      void* d_direct, float* sum_d_norm, void* stream);
int licos_gdn_bwd_out(const void* x, const void* t, const void* d_direct, int64_t n, int channels, void* dx, float* sum_dx,
                      void* stream);
/* sum_d_norm / sum_dx (fp32 [channels], accumulated, may be NULL): per-channel sums of the tensor just written, i.e.
 * beta_hat.grad and the preceding conv's bias.grad, without a second pass (channels % 8 == 0, <= 512).
 * Gradient through NonNegativeParametrizer with LowerBound's rule: d_p = d_hat * 2 max(p, bound) where p >= bound or < 0. */
int licos_gdn_param_grad(const float* beta, const float* gamma, const float* d_beta_hat, const float* d_gamma_hat, int channels,
                         float beta_bound, float gamma_bound, float* d_beta, float* d_gamma, void* stream);
/* ReLU backward: dx = y > 0 ? g : 0 (dx may alias g) */
int licos_relu_bwd(const void* y, const void* g, int64_t n, void* dx, void* stream);
/* acc[c] += sum over rows of x[row][c]  (bias / beta gradients); x bf16 [rows][channels], channels % 8 == 0, <= 512 */
int licos_colsum_bf16(const void* x, int64_t rows, int channels, float* acc, void* stream);
/* Patch matrix of a 5x5 stride-2 window over a fp32 NCHW tensor (weight gradients of g_a[0] and g_s[6]):
 * rows[(b, oh, ow)][k] = x[b][c][2 oh + kh - 2][2 ow + kw - 2], k = (c*5 + kh)*5 + kw, zero padded to
 * licos_im2col5x5s2_kpad(channels) bf16 columns (the next multiple of 16). */
int64_t licos_im2col5x5s2_kpad(int channels);
int licos_im2col5x5s2(const float* x, int batch, int channels, int h, int w, void* rows, void* stream);
/* The same weight gradient WITHOUT the patch matrix in HBM (the im2col tile is built in shared memory):
 *   out[cs][k] += sum_{b,oh,ow} small[b][oh][ow][cs] * image[b][c][2 oh + kh - 2][2 ow + kw - 2],  k = (c*5 + kh)*5 + kw
 * small_t bf16 NHWC [batch][ceil(h/2)][ceil(w/2)][small_c] (g_a[0]: the conv output's gradient -> Conv2d.weight.grad[cs][c][kh][kw];
 * g_s[6]: the layer input -> ConvTranspose2d.weight.grad[cs][c][kh][kw]), image fp32 NCHW [batch][channels][h][w],
 * out fp32 [small_c][licos_im2col5x5s2_kpad(channels)], 16-byte aligned, accumulated (the caller zeroes it).
 * Built for channels in {1, 3}, small_c a multiple of 64 <= 256, w % 4 == 0, 16-byte aligned tensors; anything else
 * returns LICOS_ERR_UNSUPPORTED (take licos_im2col5x5s2 + licos_conv_wgrad(CONV_1X1)). */
int licos_conv_wgrad_image(const void* small_t, const float* image, int batch, int channels, int h, int w, int small_c,
                           float* out, int sm_count, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* EntropyBottleneck (compressai.entropy_models.EntropyBottleneck)   */
/* ------------------------------------------------------------------------------------------ */
#define LICOS_EB_MAX_LAYERS 8
#define LICOS_EB_FORM_PLAIN 0  /* sigmoid(upper) - sigmoid(lower)            (CompressAI >= 1.2) */
#define LICOS_EB_FORM_STABLE 1 /* |sigmoid(s*upper) - sigmoid(s*lower)|, s = -sign(lower+upper)   */
#define LICOS_EB_LUT_RADIUS 128 /* eval-mode table covers symbols in [-128, 128]                 */

typedef struct licos_eb_params {
    int channels;
    int n_layers;                        /* len(filters) + 1                                    */
    int widths[LICOS_EB_MAX_LAYERS + 1]; /* (1,) + filters + (1,)                               */
    int params_per_channel;              /* floats per channel in `packed`                      */
    /* [channels][params_per_channel]; per layer i: softplus(_matrix_i) row-major (out, in),
     * _bias_i, then tanh(_factor_i) for i < n_layers - 1 */
    const float* packed;
    const float* medians; /* [channels] = quantiles[:, 0, 1]                                    */
    int form;             /* LICOS_EB_FORM_*                                                    */
    float likelihood_bound; /* <= 0 disables the LowerBound                                     */
} licos_eb_params;

int64_t licos_eb_lut_floats(int channels); /* size of the eval-mode workspace, in floats */

/* EntropyBottleneck.forward(x, training=False): y_hat = round(x - med) + med, likelihood(y_hat).
 * x, y_hat, lik: fp32 [batch][channels][hw].  lut_ws: licos_eb_lut_floats(channels) floats. */
int licos_eb_forward_eval(const licos_eb_params* p, const float* x, int batch, int64_t hw, float* lut_ws,
                          float* y_hat, float* lik, void* stream);
/* The [channels][257] likelihood table of the eval-mode passes: a function of the parameters only, so an inference
 * loop builds it once per parameter version (licos_eb_lut_floats(channels) floats). */
int licos_eb_build_lut(const licos_eb_params* p, float* lut, void* stream);

/* Eval-mode bottleneck in ONE pass over x, producing whichever of the outputs are non-NULL:
 *   y_hat, lik            EntropyBottleneck.forward(x, training=False)
 *   symbols               EntropyModel.quantize(x, "symbols", medians), int32 [batch][channels][hw]
 *   symbols_i16           the same saturated to int16 (what leaves the device in the codec loop: half the bytes)
 *   y_hat_nhwc_bf16       y_hat in the layout licos_conv_forward reads
 *   sum_ln                *sum_ln += sum(ln(lik)) in float64 -- the rate term of compute_bpp (eval_utils.py:172-186)
 *                         and of RateDistortionLoss, without a second pass over the likelihoods
 * hw % 4 == 0, channels even. */
typedef struct licos_eb_fused_args {
    const float* x;
    int batch;
    int lut_ready; /* 1: `lut` already holds licos_eb_build_lut's output for these parameters; 0: build it first */
    int64_t hw;
    float* lut;
    float* y_hat;
    float* lik;
    int32_t* symbols;
    int16_t* symbols_i16;
    void* y_hat_nhwc_bf16;
    double* sum_ln;
} licos_eb_fused_args;
int licos_eb_eval_fused(const licos_eb_params* p, const licos_eb_fused_args* args, void* stream);
/* EntropyBottleneck.forward(x, training=True): y_hat = x + noise.  noise == NULL draws U(-0.5, 0.5)
 * from an in-kernel Philox stream keyed by (seed, element index). */
int licos_eb_forward_noise(const licos_eb_params* p, const float* x, const float* noise, uint64_t seed,
                           int batch, int64_t hw, float* y_hat, float* lik, void* stream);
/* Backward of EntropyBottleneck.forward(x, training=True) (train.py:193): y_hat is the forward's output (x + noise).
 *   d_x[e] = g_yhat[e] + g_lik[e] * d likelihood / d y_hat   (either gradient may be NULL = zero; LowerBound's rule applied)
 *   d_packed[c][:] += gradient with respect to the PACKED parameter block of channel c (softplus(_matrix), _bias,
 *   tanh(_factor) as laid out in licos_eb_params.packed); the caller zeroes it and applies d softplus / d tanh. */
int licos_eb_backward(const licos_eb_params* p, const float* y_hat, const float* g_lik, const float* g_yhat, int batch,
                      int64_t hw, float* d_x, float* d_packed, void* stream);
/* The raw parameters of the density network (device pointers; entries >= n_layers, and factor[n_layers - 1], unused):
 * _matrix_i (C, F[i+1], F[i]), _bias_i (C, F[i+1], 1), _factor_i (C, F[i+1], 1). */
typedef struct licos_eb_raw_params {
    const float* matrix[LICOS_EB_MAX_LAYERS];
    const float* bias[LICOS_EB_MAX_LAYERS];
    const float* factor[LICOS_EB_MAX_LAYERS];
} licos_eb_raw_params;
typedef struct licos_eb_raw_grads {
    float* matrix[LICOS_EB_MAX_LAYERS];
    float* bias[LICOS_EB_MAX_LAYERS];
    float* factor[LICOS_EB_MAX_LAYERS];
} licos_eb_raw_grads;
/* packed[c][:] = (softplus(_matrix_i), _bias_i, tanh(_factor_i))_i, the block licos_eb_params.packed points to, in one launch. */
int licos_eb_pack_params(const licos_eb_raw_params* raw, int channels, int n_layers, const int* widths, float* packed,
                         void* stream);
/* Raw-parameter gradients from licos_eb_backward's d_packed: d_matrix = d * sigmoid(_matrix), d_bias = d,
 * d_factor = d * (1 - tanh(_factor)^2); written (not accumulated) into `grads`. */
int licos_eb_param_grads(const licos_eb_raw_params* raw, const float* d_packed, int channels, int n_layers, const int* widths,
                         const licos_eb_raw_grads* grads, void* stream);
/* EntropyBottleneck.loss() (CompressionModel.aux_loss, train.py:198): loss[0] += sum |logits(quantiles) - target| with the
 * parameters held constant (stop_gradient); d_quantiles[c][k] = its gradient.  quantiles: [C][3], target: [3]. */
int licos_eb_aux_loss(const float* packed, int channels, int n_layers, const int* widths, const float* quantiles,
                      const float* target, float* loss, float* d_quantiles, void* stream);
/* EntropyModel.quantize(x, "symbols", medians) and EntropyBottleneck._build_indexes.
 * symbols / indexes: int32 [batch][channels][hw]; indexes may be NULL. */
int licos_eb_symbols(const float* x, const float* medians, int batch, int channels, int64_t hw,
                     int32_t* symbols, int32_t* indexes, void* stream);
/* EntropyModel.dequantize(symbols, medians) */
int licos_eb_dequantize(const int32_t* symbols, const float* medians, int batch, int channels, int64_t hw,
                        float* y_hat, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* GaussianConditional (compressai.entropy_models.GaussianConditional)   */
/* ------------------------------------------------------------------------------------------ */
/* forward(y, scales, means, training): mode 0 = dequantize, 1 = noise (noise may be NULL -> Philox). */
int licos_gc_forward(const float* y, const float* scales, const float* means, const float* noise,
                     uint64_t seed, int64_t n, int training, float scale_bound, float likelihood_bound,
                     float* y_hat, float* lik, void* stream);
/* Backward of the training-mode forward (y_hat = y + noise is the forward's output): d_y = g_yhat + g_lik * d lik / d y_hat,
 * d_scales = g_lik * d lik / d scale, d_means (may be NULL) = -(the likelihood part of d_y); either incoming gradient may be
 * NULL; LowerBound's gradient rule is applied to the likelihood bound and to the scale bound. */
int licos_gc_backward(const float* y_hat, const float* scales, const float* means, const float* g_lik, const float* g_yhat,
                      int64_t n, float scale_bound, float likelihood_bound, float* d_y, float* d_scales, float* d_means,
                      void* stream);
/* build_indexes(scales): idx = n_table-1 - #{k < n_table-1 : max(scale, bound) <= table[k]} */
int licos_gc_build_indexes(const float* scales, int64_t n, const float* table, int n_table,
                           float scale_bound, int32_t* indexes, void* stream);
/* quantize(y, "symbols", means): int32 round-half-even of (y - means); means may be NULL. */
int licos_gc_symbols(const float* y, const float* means, int64_t n, int32_t* symbols, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Rate-distortion reductions (compressai.losses.RateDistortionLoss, eval_utils.py:172-186)    */
/* ------------------------------------------------------------------------------------------ */
/* acc[0] += sum(ln(lik[i])) ; caller zeroes acc (double, device).  bpp = acc / (-ln2 * N*H*W). */
int licos_sum_log(const float* lik, int64_t n, double* acc, void* stream);
/* acc[0] += sum((a[i]-b[i])^2) */
int licos_sum_sq_err(const float* a, const float* b, int64_t n, double* acc, void* stream);

/* Their backward (RateDistortionLoss under autograd, train.py:192-193); g_dev is a DEVICE scalar, the incoming gradient:
 * out[i] = coef * g_dev[0] / lik[i]   and   out[i] = coef * g_dev[0] * (a[i] - b[i]). */
int licos_scaled_reciprocal(const float* lik, int64_t n, float coef, const float* g_dev, float* out, void* stream);
int licos_scaled_diff(const float* a, const float* b, int64_t n, float coef, const float* g_dev, float* out, void* stream);

/* Raw-tile input scaling (raw_image_folder.py:192-196 `_open_band_`): out = dn / dn_max (dn_max = 4095, raw_utils.py:128),
 * and, when requant8 != 0 (use_full_range = False), out = rint(out * 255) / 255; evaluated in float64, stored as fp32. */
int licos_raw_dn_to_unit(const uint16_t* dn, int64_t n, int dn_max, int requant8, float* out, void* stream);

/* MS-SSIM building blocks (eval_utils.py:159-169 compute_msssim = pytorch_msssim.ms_ssim): one pyramid level --
 * sums[2*i] += sum of the SSIM map, sums[2*i+1] += sum of the contrast-structure map of image-channel i over its
 * (h-10) x (w-10) valid positions (11-tap separable window `win11`, a HOST pointer; c1 = (0.01*L)^2, c2 = (0.03*L)^2).
 * x, y: fp32 [bc][h][w]; workspace: licos_msssim_workspace_floats(bc, h, w) floats; sums: double [bc][2], caller zeroes. */
int64_t licos_msssim_workspace_floats(int64_t bc, int h, int w);
int licos_msssim_level(const float* x, const float* y, int64_t bc, int h, int w, const float* win11, float c1, float c2,
                       float* workspace, double* sums, void* stream);
/* F.avg_pool2d(x, 2, padding=(h % 2, w % 2)): out is [bc][(h + 2*(h%2) - 2)/2 + 1][(w + 2*(w%2) - 2)/2 + 1]. */
int licos_avgpool2(const float* x, int64_t bc, int h, int w, float* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Host-side integer path (compressai._CXX.pmf_to_quantized_cdf, compressai.ans)               */
/* ------------------------------------------------------------------------------------------ */
/* pmf (host, n floats) -> cdf (host, n + 1 uint32), strictly increasing, cdf[n] = 1 << precision. */
int licos_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf);

/* rANS (64-bit state, 32-bit renormalisation, 16-bit precision, 4-bit bypass), bitstream-compatible
 * with compressai.ans.RansEncoder.encode_with_indexes.  All pointers are host pointers.
 * cdfs is [n_cdfs][cdf_stride] int32.  Returns the number of bytes written or a negative error. */
int64_t licos_rans_encode(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                          int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                          uint8_t* out, int64_t out_capacity);
int licos_rans_decode(const uint8_t* encoded, int64_t n_bytes, const int32_t* indexes, int64_t n,
                      const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                      const int32_t* offsets, int32_t* symbols);
/* The per-image loop of EntropyModel.compress, run on `threads` host threads.  symbols/indexes are
 * [batch][n]; image i is written at out + i*out_stride, its length to out_sizes[i]. */
int licos_rans_encode_batch(const int32_t* symbols, const int32_t* indexes, int batch, int64_t n,
                            int64_t index_batch_stride, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out,
                            int64_t out_stride, int64_t* out_sizes, int threads);
int licos_rans_decode_batch(const uint8_t* const* encoded, const int64_t* n_bytes, const int32_t* indexes,
                            int batch, int64_t n, int64_t index_batch_stride, const int32_t* cdfs,
                            int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                            int32_t* symbols, int threads);

/* Device-resident rANS ENCODER (same bitstream as licos_rans_encode; replaces the symbol D2H copy + host coding of
 * EntropyModel.compress).  One thread per image; symbols, indexes and the integer tables are device pointers.
 *   indexes == NULL: index of symbol i is i / n_spatial (EntropyBottleneck: one table per channel), otherwise an int32
 *   array [batch][n] (index_stride = n) or [n] shared by every image (index_stride = 0).
 *   rcp_ws: n_cdfs * cdf_stride uint64 of scratch.  work: batch * cap_words uint32 of scratch; image b's stream is the
 *   last lengths[b] words of its row (lengths[b] = -1: cap_words too small or an invalid index -> use the host coder). */
int licos_rans_encode_device(const int32_t* symbols, const int32_t* indexes, int64_t index_stride, int batch, int64_t n,
                             int64_t n_spatial, const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                             const int32_t* offsets, uint64_t* rcp_ws, uint32_t* work, int64_t cap_words, int32_t* lengths,
                             void* stream);
/* Device-resident rANS DECODER (mirror of licos_rans_decode; one thread per image).  packed: all streams back to back as
 * 32-bit words (a stream's length is a multiple of 4 bytes), stream b = packed[word_offsets[b] .. + n_words[b]).
 * symbols: int32 [batch][n] out.  status[b] = 0, or -1 for a malformed stream / invalid index. */
int licos_rans_decode_device(const uint32_t* packed, const int64_t* word_offsets, const int32_t* n_words, const int32_t* indexes,
                             int64_t index_stride, int batch, int64_t n, int64_t n_spatial, const int32_t* cdfs, int n_cdfs,
                             int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* symbols, int32_t* status,
                             void* stream);
/* Packs the streams back to back: out[word_offsets[b] .. + lengths[b]) = stream b (word_offsets: exclusive prefix sum). */
int licos_rans_pack_device(const uint32_t* work, int64_t cap_words, const int32_t* lengths, const int64_t* word_offsets,
                           int batch, uint32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Federated merge (licos/federation_utils.py:47-53)                                           */
/* ------------------------------------------------------------------------------------------ */
/* dst[i] = w_a * a[i] + w_b * b[i] over a flat fp32 parameter buffer. */
int licos_weighted_sum2(const float* a, const float* b, float w_a, float w_b, int64_t n, float* dst,
                        void* stream);
/* buf[i] *= w */
int licos_scale_inplace(float* buf, float w, int64_t n, void* stream);

/* N-way merge over NCCL / NVLink (BASELINE.json configs[4]: one rank per GPU replaces the reference's file + lock-file
 * merge, federation_utils.py:27-85).  NCCL is bound at run time from the libnccl.so.2 the process already carries.
 *   licos_nccl_unique_id     rank 0 fills 128 host bytes; the caller ships them to the other ranks (any transport)
 *   licos_nccl_comm_create   collective over all ranks (ncclCommInitRank) on the current CUDA device
 *   licos_nccl_weighted_allreduce
 *       flat[0..n) <- sum_r w_r flat_r[0..n) / sum_r w_r,   w_r = *weight_dev if given, else 1 / *loss_dev (a loss that is
 *       not a positive finite number gets a vanishing weight) -- both DEVICE scalars, nothing is read back by the host.
 *       One prep kernel, ONE ncclAllReduce with an ncclRedOpCreatePreMulSum operator (scalar dereferenced on the device
 *       while the collective runs) over n + 1 elements, one normalising kernel.  `flat` must have room for n + 1 floats
 *       (the spare element carries sum_r w_r) and be 16-byte aligned; scalar_dev is one float of scratch.
 * With two ranks and weights (w, 1 - w) this is federation_utils.py:47-53 exactly. */
int licos_nccl_version(void); /* e.g. 22809, or LICOS_ERR_UNSUPPORTED when no libnccl can be loaded */
int licos_nccl_unique_id(void* id128_host);
int licos_nccl_comm_create(const void* id128_host, int world, int rank, void** comm_out);
int licos_nccl_comm_destroy(void* comm);
int licos_nccl_weighted_allreduce(void* comm, float* flat, int64_t n, const float* loss_dev, const float* weight_dev,
                                  float* scalar_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LICOS_B200_H */
