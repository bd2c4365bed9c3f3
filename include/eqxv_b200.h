/*
 * eqxv_b200.h — C ABI of libeqxv_b200.so: the B200 (sm_100a) implementation of eqxvision's
 * vmapped inference forward pass.
 *
 * The reference (paganpasta/eqxvision) has no FFI: its boundary is the Python API, and all
 * arithmetic is delegated to equinox.nn / jax.lax (third party). Each entry point below therefore
 * cites the reference call site(s) whose arithmetic it replaces. Paths are relative to the
 * reference checkout.
 *
 * Conventions
 *   - every function returns 0 (EQXV_OK) or a negative eqxv_status; the text of the last failure
 *     on the calling thread is returned by eqxv_last_error().
 *   - the caller owns every buffer (device pointers; PyTorch CUDA tensors are used as holders),
 *     nothing is allocated inside, nothing synchronises the host: all launches go to `stream`
 *     (a cudaStream_t passed as void*) and are CUDA-Graph capturable.
 *   - activations are channels-last bf16: [N, H, W, pitch] with `pitch >= C` elements per pixel
 *     (pitch % 8 == 0, base 16-byte aligned). Token matrices are row-major [rows, pitch].
 *   - convolution / linear weights are bf16, K-major: [Cout, kh*kw*Cin] with the BatchNorm scale
 *     already folded in (see eqxvision_b200/_pack.py); `bias` is the fp32 folded shift.
 */
#ifndef EQXV_B200_H_
#define EQXV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum eqxv_status {
  EQXV_OK = 0,
  EQXV_ERR_INVALID_ARGUMENT = -1,
  EQXV_ERR_UNSUPPORTED = -2,
  EQXV_ERR_CUDA = -3,
  EQXV_ERR_NO_DEVICE = -4
} eqxv_status;

/* epilogue activations (jax.nn.* used by the reference models) */
typedef enum eqxv_act {
  EQXV_ACT_NONE = 0,
  EQXV_ACT_RELU = 1,        /* jnn.relu         resnet.py:70, vgg.py:141 */
  EQXV_ACT_SILU = 2,        /* jnn.silu         efficientnet.py:134 */
  EQXV_ACT_GELU_TANH = 3,   /* jnn.gelu (approximate=True default)  vit.py:96, mlps.py:62 */
  EQXV_ACT_HARDSWISH = 4,   /* jnn.hard_swish   mobilenetv3.py:79 */
  EQXV_ACT_SIGMOID = 5,     /* jnn.sigmoid      squeeze.py:42 */
  EQXV_ACT_HARDSIGMOID = 6, /* jnn.hard_sigmoid mobilenetv3.py:57 */
  EQXV_ACT_RELU6 = 7
} eqxv_act;

enum {
  EQXV_FLAG_OUT_F32 = 1,        /* y is fp32 (logits); residual not allowed */
  EQXV_FLAG_RES_AFTER_ACT = 2,  /* y = act(conv + bias) + residual  (default: act(conv+bias+res)) */
  /* grouped convolution (equinox.nn.Conv2d(groups=G), resnet.py:19-23,83: ResNeXt) stored block-diagonally:
   * output channels [64b, 64b+64) read only input channels [64b, 64b+64); wgt is [cout, kh*kw*64]
   * (zero outside a group's own channels). Requires cin == cout, cin % 64 == 0, 64 % (cin/G) == 0. */
  EQXV_FLAG_GROUPED_BLOCK64 = 4,
  /* Dense filters with cin > 64 and cin % 64 != 0: a TMA box that reaches past the tensor's inner extent is served at
   * a fraction of the normal rate (measured on B200: the same rows run 2-3.6x faster when every box is in bounds), so
   * the LAST 64-wide K chunk of the activation is fetched from channels [cin - 64, cin) instead of [64 (kc-1), 64 kc).
   * wgt is then [cout, kh*kw * 64 kc] (kc = ceil(cin / 64)): per tap, chunks 0 .. kc-2 as usual, the last chunk holds
   * the filter of channels [cin - 64, cin) with the 64 kc - cin columns that repeat the previous chunk set to ZERO
   * (eqxvision_b200/_pack.py: pack_conv_weight(tail_shift=True)). Results are unchanged. */
  EQXV_FLAG_K_TAIL_SHIFT = 8
};

const char* eqxv_version(void);
const char* eqxv_last_error(void);
/* Binds the calling thread to `device`, raises the kernels' dynamic shared-memory limits and
 * resolves cuTensorMapEncodeTiled. Must be called once per process before any other entry. */
int eqxv_init(int device);
int eqxv_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1/K2: dense convolution as implicit GEMM on tcgen05 tensor cores, fused epilogue
 *   y = act( conv(x, w) + bias [+ residual] )      (or act(...) + residual with RES_AFTER_ACT)
 * Replaces  equinox.nn.Conv2d -> lax.conv_general_dilated  +  experimental.BatchNorm(inference)
 *           + jnn.<act>  (+ `out += identity`)  as composed in
 *   resnet.py:144-162 (_ResNetBottleneck.__call__), resnet.py:80-92, resnet.py:344-346 (stem),
 *   layers/conv_norm_activation.py:61-85, vgg.py:137-145, deeplabv3.py:43-53, densenet.py:66.
 * groups must be 1. kh*kw taps are walked over 64-channel K blocks; A tiles are 4-D TMA boxes of
 * the NHWC input (zero fill gives the padding, box traversal stride gives the conv stride).
 * ------------------------------------------------------------------------------------------- */
typedef struct eqxv_conv_desc {
  const void* x;        /* bf16 [n, h, w, x_pitch]            */
  const void* wgt;      /* bf16 [cout, kh*kw*cin]             */
  const float* bias;    /* fp32 [cout] or NULL                */
  const void* residual; /* bf16 [n, ho, wo, res_pitch] or NULL */
  void* y;              /* bf16 (fp32 with OUT_F32) [n, ho, wo, y_pitch] */
  int32_t n, h, w, cin, cout;
  int32_t kh, kw, stride, pad, dil;
  int32_t x_pitch, y_pitch, res_pitch;
  int32_t act;   /* eqxv_act */
  int32_t flags; /* EQXV_FLAG_* */
} eqxv_conv_desc;
int eqxv_conv2d_igemm_bf16(const eqxv_conv_desc* d, void* stream);

/* K5: out[m, :n] = act(a[m, :k] @ w[n, :k]^T + bias [+ residual])  — the same tcgen05 kernel with
 * one tap. Replaces equinox.nn.Linear under jax.vmap: vit.py:64,74 (qkv, proj), mlps.py:61-65
 * (fc1/fc2), resnet.py:356 (fc), vgg.py:97-106, and every 1x1 stride-1 Conv2d. */
int eqxv_gemm_bias_act_res_bf16(const void* a, int64_t lda, const void* w, const float* bias,
                                const void* residual, int64_t ldr, void* out, int64_t ldo, int64_t m,
                                int32_t n, int32_t k, int32_t act, int32_t flags, void* stream);

/* K5 + K7 fused: the LayerNorm between two GEMMs never touches HBM (vit.py:149,154: `x + attn(norm1(x))`,
 * `x + mlp(norm2(x))`; mlps.py:61-65). The GEMM that PRODUCES the LayerNorm input (attention projection / fc2 with
 * the residual add in its epilogue) also writes, per row and 64-column chunk, (sum, sum of squares) of the values it
 * stores: row_stats is fp32 [m][ceil(n/64)][2]. */
int eqxv_gemm_res_rowstats_bf16(const void* a, int64_t lda, const void* w, const float* bias, const void* residual,
                                int64_t ldr, void* out, int64_t ldo, float* row_stats, int64_t m, int32_t n, int32_t k,
                                void* stream);
/* ... and the GEMM that CONSUMES LayerNorm(x) reads x itself: with w' = w * gamma (per input column, packed by the
 * caller), bias' = bias + w @ beta and wsum[j] = sum_k w'[j, k],
 *   out[i, j] = act( rstd_i * (x_i . w'_j - mean_i * wsum[j]) + bias'[j] ),
 * mean_i / rstd_i from the producer's row_stats (slots = ceil(k/64) chunks per row; biased variance, eps as
 * equinox.nn.LayerNorm). act: none or tanh-GELU. */
int eqxv_gemm_ln_act_bf16(const void* a, int64_t lda, const void* w, const float* bias, const float* wsum,
                          const float* row_stats, int32_t slots, float eps, void* out, int64_t ldo, int64_t m, int32_t n,
                          int32_t k, int32_t act, void* stream);

/* K5 + K11 fused: SqueezeExcitation's `x * scale` (layers/squeeze.py:61) applied to the A operand of the projection that
 * consumes it (efficientnet.py:161-170, mobilenetv3.py:113-121): out = (a * gate[row / rows_per_image]) @ w^T + bias
 * (+ residual). The staged A tile is scaled in shared memory before the MMA (bf16 x bf16 rounded to bf16, exactly what
 * the separate gate pass stored), so the expanded tensor is read ONCE and never rewritten. gate: bf16 [images, ldg]. */
int eqxv_gemm_gated_bf16(const void* a, int64_t lda, const void* gate, int64_t ldg, int32_t rows_per_image, const void* w,
                         const float* bias, const void* residual, int64_t ldr, void* out, int64_t ldo, int64_t m,
                         int32_t n, int32_t k, int32_t flags /* 0 or EQXV_FLAG_K_TAIL_SHIFT */, void* stream);

/* K1 + K2 + K4 fused across layer boundaries: one whole ResNet bottleneck with a 64-channel trunk (resnet.py:144-162
 * as instantiated by resnet.py:288-296: layer1 of ResNet-50/101/152), minus its first 1x1 convolution, plus - optionally -
 * the first 1x1 convolution of the block that follows:
 *     t2   = relu(conv3x3(t1; w2) + b2)                                 64 -> 64, stride 1, pad 1   (never stored)
 *     y    = relu(conv1x1(t2; w3) + b3 + residual)                      64 -> 256                   identity shortcut
 *          | relu([t2 | x0] @ [w3 | wd]^T + (b3 + bd))                                              downsample shortcut
 *     next = relu(conv1x1(y; w1n) + b1n)                                256 -> 64 | 128             (optional)
 * BatchNorm folded into w / b by the caller as for eqxv_conv2d_igemm_bf16. All tensors NHWC bf16; biases fp32.
 * Exactly one of `residual` (256 channels) / `x0` (the 64-channel input of the block's downsample convolution, w3 then
 * holds [w3 | wd] along K: [256, 128], b3 = b3 + bd) must be given. Rounding points are those of the layer-by-layer
 * path (t2, y and next are rounded to bf16 where that path stored them), so results agree with it to accumulation
 * order. One CTA pair per two 8 x 16 pixel tiles, filters resident in shared memory (csrc/bottleneck.cu). */
typedef struct eqxv_bottleneck64_desc {
  const void* t1;       /* bf16 [n, h, w, t1_pitch], 64 channels: output of the block's first 1x1 convolution */
  const void* w2;       /* bf16 [64, 9*64] (tap-major K, as eqxv_conv2d_igemm_bf16) */
  const float* b2;      /* fp32 [64] */
  const void* w3;       /* bf16 [256, 64], or [256, 128] = [w3 | wd] with x0 */
  const float* b3;      /* fp32 [256] */
  const void* residual; /* bf16 [n, h, w, res_pitch], 256 channels, or NULL */
  const void* x0;       /* bf16 [n, h, w, x0_pitch], 64 channels, or NULL */
  void* y;              /* bf16 [n, h, w, y_pitch], 256 channels */
  const void* w1n;      /* bf16 [next_channels, 256] or NULL */
  const float* b1n;     /* fp32 [next_channels] or NULL */
  void* next;           /* bf16 [n, h, w, next_pitch], next_channels channels, or NULL */
  int32_t n, h, w;
  int32_t t1_pitch, res_pitch, x0_pitch, y_pitch, next_pitch;
  int32_t next_channels; /* 64 (the next block of the same stage) or 128 (first block of the next stage); 0 without next */
} eqxv_bottleneck64_desc;
int eqxv_bottleneck64_fused_bf16(const eqxv_bottleneck64_desc* d, void* stream);

/* First-layer ("stem") convolution on the raw image, cin <= 8: resnet.py:243-251 (7x7 s2 p3),
 * vgg.py:137 (3x3 s1 p1), efficientnet.py:327-337 / mobilenetv3.py:196-206 (3x3 s2 p1), densenet.py:175.
 * Input is the padded 8-channel image written by eqxv_pack_stem_input; weights are [cout, kh, 8, 8] bf16
 * (filter row, 8 columns of which kw are real, 8 channels of which cin are real). One filter row is
 * one 64-wide K block of the tcgen05 GEMM (kh <= 8, kw <= 8, stride 1/2, 2*pad <= kw). */
int eqxv_conv_stem_bf16(const void* xpad, const void* wgt, const float* bias, void* y, int32_t n, int32_t h,
                        int32_t w, int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                        int32_t y_pitch, int32_t act, void* stream);
/* Pixel-pair variant for stride-2 first layers with <= 4 input channels (the ResNet / EfficientNet / MobileNet stems): the
 * padded image is bf16 [n, h+2*pad, (w+8)/2, 8] - one 16-byte unit = padded columns (2u, 2u+1) x 4 channels, written by
 * eqxv_pack_stem_input_c4 / eqxv_u8hwc_pack_stem_input_c4 - and the filter is [cout, kh, 64] with K index 4*s + c
 * (s < kw, c < cin; zeros elsewhere). A stride-2 convolution walks the units with stride 1, a 7-tap filter row is 4 unit
 * taps = 2 K steps: half the MMAs, half the staged bytes and half the packed image of the 8-channel layout. Same
 * arithmetic, same results. w must be even. */
int eqxv_conv_stem_c4_bf16(const void* xpad4, const void* wgt, const float* bias, void* y, int32_t n, int32_t h, int32_t w,
                           int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t y_pitch, int32_t act,
                           void* stream);
int eqxv_pack_stem_input_c4(const float* x_nchw, void* xpad4, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad,
                            void* stream);
/* ... with the max-pool that follows it in the ResNet stem (resnet.py:243-253: conv1 -> bn1 -> relu -> maxpool 3x3 /
 * stride 2 / pad 1) in the kernel's epilogue: y_pooled is bf16 [n, ho/2, wo/2, y_pitch]; the conv output itself (411 MB
 * for a 256-image batch) is never written. ReLU is implied (0 is then the identity of max, which is what lets the pooled
 * pixels on tile borders be combined with red.global.max). Needs cout == 64 and conv output extents ho % 16 == 0,
 * wo % 8 == 0; anything else fails with EQXV_ERR_INVALID_ARGUMENT (use the two separate entries). */
int eqxv_conv_stem_maxpool_bf16(const void* xpad, const void* wgt, const float* bias, void* y_pooled, int32_t n, int32_t h,
                                int32_t w, int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                                int32_t y_pitch, void* stream);
/* fp32 NCHW [n,c<=8,h,w] (the reference's input layout, README.md:45) -> bf16 [n, h+2*pad, w+8, 8]:
 * the image sits at rows pad..h+pad-1, columns pad..w+pad-1; border and channels >= c are zero. */
int eqxv_pack_stem_input(const float* x_nchw, void* xpad, int32_t n, int32_t c, int32_t h, int32_t w,
                         int32_t pad, void* stream);

/* input boundary: fp32 NCHW -> bf16 NHWC with channels zero-padded to c_pad (multiple of 8) */
int eqxv_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t c_pad, void* stream);
/* output boundary: bf16 NHWC -> fp32 NCHW (feature maps returned to the caller, test_vgg.py:30) */
int eqxv_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t x_pitch, void* stream);

/* K9: equinox.nn.MaxPool2d(k, stride, padding) (-inf padding): resnet.py:254, vgg.py:134 */
int eqxv_maxpool2d_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                             int32_t k, int32_t stride, int32_t pad, int32_t x_pitch, int32_t y_pitch,
                             void* stream);
/* K9 with use_ceil=True (squeezenet.py:84,89,95; googlenet.py:95,98,103,112 and the inception pooling branch
 * googlenet.py:228): output extent ceil((size + 2*pad - k) / stride) + 1, minus one when the last window would start
 * beyond the input (torch's rule, which the reference's tests pin at 1e-4); partial windows are clipped. */
int eqxv_maxpool2d_ceil_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                                  int32_t k, int32_t stride, int32_t pad, int32_t x_pitch, int32_t y_pitch,
                                  void* stream);
/* K10: equinox.nn.AvgPool2d(k, stride) (densenet.py:128) */
int eqxv_avgpool2d_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                             int32_t k, int32_t stride, int32_t x_pitch, int32_t y_pitch, void* stream);
/* K10: AdaptiveAvgPool2d((oh, ow)), resnet.py:283, vgg.py:90; (1,1) is the global pool that feeds the classifier.
 * h % oh == 0 and w % ow == 0: block mean. Otherwise EQUINOX's rule (not torch's overlapping windows): the axis is cut
 * into consecutive blocks, the first dim % t of them dim // t + 1 long, the rest dim // t (GoogLeNet's auxiliary heads,
 * googlenet.py:265-268: 14x14 -> 4x4; AlexNet / VGG on inputs other than 224 px). */
int eqxv_adaptive_avgpool_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w,
                                    int32_t c, int32_t oh, int32_t ow, int32_t x_pitch,
                                    int32_t y_pitch, void* stream);

/* K7: equinox.nn.LayerNorm(d, eps) per row (vit.py:149,154,272 under jax.vmap) */
int eqxv_layernorm_bf16(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y,
                        int64_t ldy, int64_t rows, int32_t d, float eps, void* stream);

/* K6: softmax(q k^T * scale) v for `heads` heads of dim 64 over `tokens` tokens per image,
 * vit.py:62-73. qkv is [images*tokens, 3*heads*64] with column order (3, heads, 64) as produced by
 * reshape(N,3,H,d) (vit.py:65); out is [images*tokens, heads*64] (vit.py:73). If `attn_out` is
 * non-NULL the fp32 probabilities [images, heads, tokens, tokens] are also written
 * (return_attention path, vit.py:151-152). */
int eqxv_attention_fwd_bf16(const void* qkv, void* out, float* attn_out, int32_t images,
                            int32_t tokens, int32_t heads, int32_t head_dim, float scale,
                            void* stream);

/* K12: patch rows for PatchEmbed (layers/patch_embed.py:79-82): fp32 NCHW [n,c,h,w] ->
 * bf16 [n*gh*gw, c*p*p] with the K order (c, py, px) of the conv weight flattened. */
int eqxv_patchify_nchw_f32_bf16(const float* x, void* rows, int32_t n, int32_t c, int32_t h,
                                int32_t w, int32_t p, void* stream);
/* tokens = concat([cls_token, patches]) + pos_embed  (vit.py:269); patches [n*np, d] bf16,
 * cls [d] / pos [(np+1), d] fp32, out [n*(np+1), d] bf16 */
int eqxv_vit_assemble_tokens_bf16(const void* patches, const float* cls, const float* pos, void* out,
                                  int32_t n, int32_t np, int32_t d, void* stream);
/* gather row `row` of every image's token block: [n*tokens, d] -> [n, d]  (x[0], vit.py:273) */
int eqxv_gather_rows_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t n,
                          int32_t tokens, int32_t row, int32_t d, void* stream);

/* K3: depthwise KxK convolution (groups == channels; k in {3,5,7}, stride 1/2) + folded BatchNorm +
 * activation: layers/conv_norm_activation.py:61-85 with groups=C as built by efficientnet.py:140-151 and
 * mobilenetv3.py:88-101. wgt: fp32 [k*k, w_pitch] (tap-major, BN scale folded), bias: fp32 [c]. */
int eqxv_dwconv_bn_act_bf16(const void* x, const float* wgt, const float* bias, void* y, int32_t n,
                            int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride, int32_t pad,
                            int32_t dil, int32_t x_pitch, int32_t y_pitch, int32_t w_pitch, int32_t act,
                            void* stream);
/* K3 + K11: the same depthwise convolution with the SqueezeExcitation squeeze fused in (layers/squeeze.py:52 `avgpool(x)`
 * on the output of the depthwise ConvNormActivation, efficientnet.py:138-160, mobilenetv3.py:88-112): besides y the kernel
 * writes pooled[n, c] = mean over the output map of the bf16-rounded activation, so no separate pass re-reads the widest
 * tensor of the block. One thread block works on ONE image; per-block partial sums go to `workspace` and the last block of
 * an image (integer ticket) adds them in a fixed order: bitwise reproducible, independent of the batch composition.
 * k in {3,5}, stride 1/2, dilation 1, act in {none, relu, silu, hard-swish}. `workspace`: zero-initialised once by the
 * caller (the kernel leaves it zeroed), eqxv_dwconv_pool_workspace_bytes() bytes (0 = not needed). */
int eqxv_dwconv_pool_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride, int32_t pad,
                                     int64_t* bytes);
int eqxv_dwconv_bn_act_pool_bf16(const void* x, const float* wgt, const float* bias, void* y, void* pooled,
                                 void* workspace, int64_t workspace_bytes, int32_t n, int32_t h, int32_t w, int32_t c,
                                 int32_t k, int32_t stride, int32_t pad, int32_t x_pitch, int32_t y_pitch,
                                 int32_t w_pitch, int32_t pool_pitch, int32_t act, void* stream);
/* Same contract, forced through the shared-memory stencil kernel (TMA-staged halo tile of a 64-channel block,
 * sliding accumulator window; k in {3,5}, dilation 1). eqxv_dwconv_bn_act_bf16 picks it when EQXV_DWTILE=1. */
int eqxv_dwconv_tile_bf16(const void* x, const float* wgt, const float* bias, void* y, int32_t n, int32_t h,
                          int32_t w, int32_t c, int32_t k, int32_t stride, int32_t pad, int32_t dil,
                          int32_t x_pitch, int32_t y_pitch, int32_t w_pitch, int32_t act, void* stream);
/* K8/K11/K15: y = act(x * scale[c] + shift[c] + other) * gate[row / rows_per_image, c]; every operand
 * except x may be NULL. Standalone BatchNorm+ReLU (densenet.py:64-65,118,211), unfused residual adds,
 * the SqueezeExcitation gate `x * scale` (layers/squeeze.py:61). */
int eqxv_eltwise_bf16(const void* x, const float* scale, const float* shift, const void* other,
                      const void* gate, void* y, int64_t rows, int32_t c, int32_t x_pitch,
                      int32_t other_pitch, int32_t gate_pitch, int32_t y_pitch, int32_t rows_per_image,
                      int32_t act, void* stream);
/* K14: jax.image.resize(method="bilinear") upsampling (half-pixel centres, edge clamp):
 * models/segmentation/_utils.py:52,57 (-> fp32 NCHW model output) and deeplabv3.py:74 (-> bf16 NHWC). */
int eqxv_resize_bilinear_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t n, int32_t c, int32_t h,
                                               int32_t w, int32_t oh, int32_t ow, int32_t x_pitch,
                                               void* stream);
int eqxv_resize_bilinear_nhwc_bf16(const void* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w,
                                   int32_t oh, int32_t ow, int32_t x_pitch, int32_t y_pitch, void* stream);
/* K16/K6: Swin shifted-window attention (swin.py:90-255): softmax(q*d^-1/2 k^T + rel_pos_bias + shift_mask) v
 * per (window, head); the cyclic roll, window partition/reverse and the -100 shift mask are index
 * arithmetic. qkv: [n*h*w, 3*heads*32] in spatial row order, columns (3, heads, 32); bias: fp32
 * [heads, window^2, window^2] = relative_position_bias_table[relative_position_index] (swin.py:46-57);
 * out: [n*h*w, heads*32]. */
/* Implementation: tcgen05 - two windows x one head form one 128-row MMA tile (q k^T: M=128, N=128, K=32; p v: M=128, N=32,
 * K=128) on operand tiles the CTA gathers into 128B-swizzled shared memory itself; softmax one thread per row.
 * EQXV_WATTN_TC=0 selects the CUDA-core kernel (A/B reference). */
int eqxv_window_attention_bf16(const void* qkv, const float* bias, void* out, int32_t n, int32_t h, int32_t w,
                               int32_t heads, int32_t head_dim, int32_t window, int32_t shift_h,
                               int32_t shift_w, float scale, void* stream);
/* Swin-V2 cosine attention, first half (swin.py:158-166, _ShiftedWindowAttentionV2): q and k of the spatial-order qkv
 * matrix are divided IN PLACE by their L2 norm over axis 0 of the reference's (num_windows, heads, tokens, d) arrays
 * - the windows of one image, per (head, window token, channel); the reference's quirk, torchvision normalises over d -
 * and q is multiplied by scale_q[head] = exp(min(logit_scale, log 100)). Follow with eqxv_window_attention_bf16
 * (scale = 1, bias = 16 * sigmoid(cpb_mlp(...)), swin.py:494-504). */
int eqxv_swin_v2_qk_normalize_bf16(void* qkv, const float* scale_q, int32_t n, int32_t h, int32_t w, int32_t heads,
                                   int32_t head_dim, int32_t window, int32_t shift_h, int32_t shift_w, void* stream);
/* K16: patch merging gather (swin.py:23-33): [n,h,w,c] -> [n,h/2,w/2,4c] = concat(x[0::2,0::2],
 * x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]) along channels. */
int eqxv_patch_merge_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t x_pitch,
                          int32_t y_pitch, void* stream);
/* debug aid: phase timestamps (clock64) of CTA 0 of the tcgen05 attention kernel into ts[16][16]; NULL = off */
int eqxv_debug_attention_timeline(long long* ts);
/* same for the fused bottleneck kernel (int64 [16 tiles][16 events]; tools/bneck_timeline.py); NULL switches it off */
int eqxv_debug_bottleneck_timeline(long long* ts);
/* same for the first-layer kernel (int64 [24 tiles][16 events]; tools/stem_timeline.py) */
int eqxv_debug_stem_timeline(long long* ts);
/* K13 fallback: strided device-to-device copy (channel slices of a concat buffer) */
int eqxv_copy2d_async(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                      int64_t width_bytes, int64_t rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input edge: the reference's own fixture feeds the models with
 *   transforms.Compose([Resize(s), ToTensor(), Normalize(mean, std)])(PIL image)   (tests/conftest.py:20-41).
 * Here the uint8 HWC image is what crosses PCIe; ToTensor + Normalize are fused into the layout kernels of the first
 * layer. x: uint8 [n, h, w, c] (c <= 4); lut: fp32 [c, 256] with lut[ch][v] = (float(v)/255 - mean[ch]) / std[ch]
 * evaluated by the caller with the reference's fp32 operations (eqxvision_b200/transforms.py), which makes the device
 * result bit-identical to Normalize(ToTensor(img)) before the one bf16 rounding.
 * ------------------------------------------------------------------------------------------- */
/* -> fp32 NCHW [n, c, h, w]: exactly the tensor the reference pipeline hands to the model */
int eqxv_u8hwc_to_nchw_f32(const uint8_t* x, const float* lut, float* y, int32_t n, int32_t h, int32_t w, int32_t c,
                           void* stream);
/* -> bf16 [n, h+2*pad, w+8, 8], the layout eqxv_conv_stem_bf16 reads (same as eqxv_pack_stem_input) */
int eqxv_u8hwc_pack_stem_input(const uint8_t* x, const float* lut, void* xpad, int32_t n, int32_t h, int32_t w,
                               int32_t c, int32_t pad, void* stream);
/* -> the pixel-pair layout of eqxv_conv_stem_c4_bf16 (same as eqxv_pack_stem_input_c4) */
int eqxv_u8hwc_pack_stem_input_c4(const uint8_t* x, const float* lut, void* xpad4, int32_t n, int32_t h, int32_t w,
                                  int32_t c, int32_t pad, void* stream);
/* -> bf16 NHWC [n, h, w, 8] (channels zero-padded to 8), same as eqxv_nchw_f32_to_nhwc_bf16(c_pad = 8) */
int eqxv_u8hwc_to_nhwc_bf16(const uint8_t* x, const float* lut, void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                            void* stream);
/* -> PatchEmbed rows bf16 [n*gh*gw, c*p*p], K order (c, py, px), same as eqxv_patchify_nchw_f32_bf16 */
int eqxv_u8hwc_patchify_bf16(const uint8_t* x, const float* lut, void* rows, int32_t n, int32_t h, int32_t w,
                             int32_t c, int32_t p, void* stream);
/* transforms.Resize on the uint8 image: bilinear, half-pixel centres, no antialiasing (torchvision's tensor path
 * F.interpolate(mode="bilinear", align_corners=False) in fp32), rounded half-to-even to uint8. [n,h,w,c] -> [n,oh,ow,c] */
int eqxv_u8hwc_resize_bilinear(const uint8_t* x, uint8_t* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t oh,
                               int32_t ow, void* stream);

/* ---------------------------------------------------------------------------------------------
 * C1: all-gather of the fp32 logits over NVLink peer memory: the only collective of the path,
 * issued when the caller wants the gathered [B, classes] output of a batch sharded over the GPUs of one box
 * (resnet.py:356 / vit.py:273 produce [B/N, classes] per rank). One process per GPU; each rank owns a "window"
 * (flags + two gather buffers) allocated with eqxv_p2p_alloc and mapped into every peer through CUDA IPC.
 * eqxv_allgather_push copies `bytes` of this rank's rows into slot `rank` of the current buffer of EVERY window with
 * peer stores, publishes a flag per peer and returns (on the stream) once all `world` slices have landed in its own
 * window. Buffers alternate per launch (epoch parity): read the result from eqxv_p2p_buffer_offset(launches & 1).
 * Every rank must issue the same sequence of pushes; consumers of a buffer must be stream-ordered before the next push.
 * ------------------------------------------------------------------------------------------- */
int eqxv_p2p_window_bytes(int64_t buf_bytes, int64_t* total);
int eqxv_p2p_alloc(void** ptr, int64_t bytes);   /* cudaMalloc + zero fill (IPC needs a whole allocation) */
int eqxv_p2p_free(void* ptr);
int eqxv_ipc_get_handle(const void* ptr, uint8_t handle[64]);
int eqxv_ipc_open_handle(const uint8_t handle[64], void** ptr);
int eqxv_ipc_close_handle(void* ptr);
int eqxv_allgather_push(const void* src, int64_t bytes, void* const* windows /* host array [world] */, int32_t rank,
                        int32_t world, int64_t slot_bytes, int64_t buf_bytes, void* stream);
int eqxv_p2p_buffer_offset(int32_t parity, int64_t buf_bytes, int64_t* offset);

/* ---------------------------------------------------------------------------------------------
 * plumbing: streams, CUDA graphs, events (cudaStream_t / cudaGraphExec_t / cudaEvent_t as void*)
 * ------------------------------------------------------------------------------------------- */
int eqxv_stream_create(void** stream);
int eqxv_stream_destroy(void* stream);
int eqxv_stream_sync(void* stream);
int eqxv_graph_begin(void* stream);
int eqxv_graph_end(void* stream, void** graph_exec);
int eqxv_graph_launch(void* graph_exec, void* stream);
int eqxv_graph_destroy(void* graph_exec);
int eqxv_event_create(void** ev);
int eqxv_event_destroy(void* ev);
int eqxv_event_record(void* ev, void* stream);
int eqxv_event_sync(void* ev);
int eqxv_stream_wait_event(void* stream, void* ev);
int eqxv_event_elapsed_ms(void* start, void* stop, float* ms);
/* h2d / d2h name the usual direction; the copy kind is cudaMemcpyDefault (unified addressing), so device sources work */
int eqxv_memcpy_h2d_async(void* dst, const void* src, int64_t bytes, void* stream);
int eqxv_memcpy_d2h_async(void* dst, const void* src, int64_t bytes, void* stream);
/* direction inferred from the pointers (device-to-device copies of results into caller-owned tensors) */
int eqxv_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
int eqxv_memset_async(void* dst, int value, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EQXV_B200_H_ */
