/*
 * cic.h - C ABI of libcic.so, the B200 (sm_100a) learned-image-compression hot path.
 *
 * The reference (hassanrizwank/Contextual-Image-Compression) has no FFI: its boundary is a set
 * of Python callables backed by TensorFlow/Keras and scikit-image.  Each entry point below
 * names the reference call site (file:line under /root/reference) whose arithmetic it replaces;
 * the Python modules at the repository root (GAN_functions.py, GAN_test.py, train_autoencoder.py,
 * test_autoencoder.py) keep the reference's names and signatures and bind these symbols with
 * ctypes.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a HOST pointer;
 *   - tensors are dense NHWC (Keras layout), float32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions return 0 on success or a negative CIC_ERR_* code; cic_last_error() returns a
 *     thread-local message for the last failure;
 *   - the library allocates nothing on the caller's behalf except inside an opaque cic_plan
 *     (device copies of packed weights + TMA descriptors); activations live in a caller-supplied
 *     workspace whose size cic_plan_workspace_bytes() reports;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     CIC_ERR_CUDA.
 */
#ifndef CIC_H_
#define CIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CIC_VERSION 100 /* major*1000 + minor*100 + patch */

/* ---- status codes ---------------------------------------------------------------------- */
#define CIC_OK 0
#define CIC_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define CIC_ERR_CUDA (-2)    /* CUDA runtime or driver error */
#define CIC_ERR_MISSING (-3) /* plan is missing a named weight tensor */
#define CIC_ERR_WORKSPACE (-4)

/* ---- enums ----------------------------------------------------------------------------- */
enum cic_activation { CIC_ACT_NONE = 0, CIC_ACT_RELU = 1, CIC_ACT_LRELU02 = 2, CIC_ACT_SIGMOID = 3, CIC_ACT_TANH = 4 };

/* Arithmetic of the conv/dense layers of a plan.
 *   CIC_PREC_FP32 : fp32 operands, fp32 FMA on the CUDA cores (reference-grade, slow).
 *   CIC_PREC_TC   : tcgen05 tensor cores, fp32 accumulate in TMEM.  Encoder chain (conv2..Dense,
 *                   attention) uses error-compensated 3-term split-bf16 operands so the quantised
 *                   symbols match the fp32 reference; decoders use single-pass bf16. */
enum cic_precision { CIC_PREC_FP32 = 0, CIC_PREC_TC = 1 };

enum cic_plan_kind {
  CIC_PLAN_AUTOENCODER = 1, /* train_autoencoder.py:9-40  build_autoencoder            */
  CIC_PLAN_ENCODER = 2,     /* GAN_functions.py:280-331    build_encoder                */
  CIC_PLAN_GENERATOR = 3,   /* GAN_functions.py:236-278    build_generator              */
  CIC_PLAN_SALIENCY = 4,    /* GAN_functions.py:210-234    build_latent_saliency_model  */
  CIC_PLAN_RD = 5,          /* GAN_functions.py:495-557    build_rate_distortion_optimizer */
  CIC_PLAN_ADAPTIVE = 6     /* GAN_functions.py:559-722    build_adaptive_compression_model */
};

/* A named host tensor in Keras layout (see contextual-image-compression_b200/weights.py):
 * Conv2D (kh,kw,Cin,Cout); Conv2DTranspose (kh,kw,Cout,Cin); Dense (in,out); vectors (C,). */
typedef struct cic_tensor {
  const char* name;
  const float* h_data;
  int32_t ndim;
  int64_t shape[4];
} cic_tensor;

typedef struct cic_plan_opts {
  int32_t precision;   /* enum cic_precision */
  int32_t img_h;       /* model input height (GAN: tile size, 256 in the reference) */
  int32_t img_w;
  int32_t img_c;       /* 3 */
  int32_t latent_dim;  /* ENCODER/GENERATOR/SALIENCY: latent size; ADAPTIVE: base_latent_dim */
  int32_t add_attention; /* ENCODER only */
  int32_t reserved[8];
} cic_plan_opts;

typedef struct cic_plan cic_plan;

/* ---- library ---------------------------------------------------------------------------- */
int cic_version(void);
const char* cic_last_error(void);
/* sm_count / compute capability of the current device. */
int cic_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- plans: weights are copied, packed (BN folded, bf16 hi/lo split, per-phase transposed-conv
 *      matrices) and uploaded once; not part of any timed region --------------------------- */
cic_plan* cic_plan_create(int kind, const cic_tensor* tensors, int n_tensors, const cic_plan_opts* opts);
void cic_plan_destroy(cic_plan* plan);
/* Bytes of device workspace a forward call needs for `batch` model inputs (tiles for the GAN
 * plans, images of h x w for the autoencoder, whose graph is shape-generic: h%4==w%4==0). */
size_t cic_plan_workspace_bytes(const cic_plan* plan, int batch, int h, int w);
/* Number of this library's kernels the last forward call on this plan launched. */
int cic_plan_last_launch_count(const cic_plan* plan);

/* Per-layer device timing of forward calls on this plan (CUDA events on the launch stream around every
 * layer).  cic_plan_get_profile synchronises and writes one "name,ms,flops,bytes" line per layer of the
 * last forward call into buf (NUL-terminated, truncated to cap); returns the size needed. */
int cic_plan_set_profiling(cic_plan* plan, int on);
size_t cic_plan_get_profile(cic_plan* plan, char* buf, size_t cap);

/* build_autoencoder(...).predict  (train_autoencoder.py:9-40, test_autoencoder.py:85-88).
 * d_x (B,H,W,3) in [0,1] -> d_y (B,H,W,3) in (0,1); optional d_y_u8 = (y*255).astype(uint8)
 * (truncation, test_autoencoder.py:88). */
int cic_autoencoder_forward(cic_plan* plan, const float* d_x, float* d_y, uint8_t* d_y_u8, int batch, int h, int w,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* build_encoder(...)(img) -> [latent, x1, x2, x3]  (GAN_functions.py:280-331). */
int cic_encoder_forward(cic_plan* plan, const float* d_img, float* d_latent, float* d_x1, float* d_x2, float* d_x3,
                        int batch, void* d_workspace, size_t workspace_bytes, void* stream);

/* build_generator(...)([latent, skip1, skip2, skip3]) -> image in (-1,1)  (GAN_functions.py:236-278). */
int cic_generator_forward(cic_plan* plan, const float* d_latent, const float* d_skip1, const float* d_skip2,
                          const float* d_skip3, float* d_out, int batch, void* d_workspace, size_t workspace_bytes,
                          void* stream);

/* build_latent_saliency_model(...)(latent) -> (B,1)  (GAN_functions.py:210-234). */
int cic_saliency_forward(cic_plan* plan, const float* d_latent, float* d_score, int batch, void* d_workspace,
                         size_t workspace_bytes, void* stream);

/* build_rate_distortion_optimizer(...)([img, mask, bpp]) -> rd_params (B,3)  (GAN_functions.py:495-557;
 * the image input is unused by the reference graph, :500). */
int cic_rd_forward(cic_plan* plan, const float* d_mask, const float* d_bpp, float* d_rd_params, int batch,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/* Outputs of the adaptive model (GAN_functions.py:690-696) plus optional diagnostics. Any pointer may
 * be NULL to skip that output.  n_tiles = n_img * (img_h/tile) * (img_w/tile). */
typedef struct cic_adaptive_io {
  /* inputs */
  const float* d_img;  /* (n_img, img_h, img_w, 3) in [-1,1]; img_h, img_w multiples of the tile size */
  const float* d_mask; /* (n_img, img_h, img_w, 1) in [0,1] */
  const float* d_bpp;  /* (n_img,) target bits per pixel */
  /* model outputs */
  float* d_blended;     /* (n_img, img_h, img_w, 3) */
  float* d_hq_latent_q; /* (n_tiles, 2*base) dequantised */
  float* d_lq_latent_q; /* (n_tiles, base) */
  float* d_rd_params;   /* (n_tiles, 3) */
  float* d_dt;          /* (n_img, img_h, img_w, 1) dynamic_threshold / bit allocation */
  /* diagnostics */
  int32_t* d_hq_symbols; /* round(latent*scale), (n_tiles, 2*base) */
  int32_t* d_lq_symbols;
  float* d_hq_latent;    /* pre-quantisation latents */
  float* d_lq_latent;
  float* d_hq_scale;     /* (n_tiles,) exp(3*qs*(1-sal)) */
  float* d_lq_scale;
  float* d_hq_out;       /* un-blended generator outputs (n_img, img_h, img_w, 3) */
  float* d_lq_out;
  double* d_hq_ratio_sum; /* (n_img,) sum of dt over the image (hq_ratio = sum / (img_h*img_w)), GAN_test.py:312 */
} cic_adaptive_io;

/* adaptive_model.predict([img, mask, bpp])  (GAN_functions.py:604-696, GAN_test.py:292).  Images larger
 * than the model's tile are processed as independent tiles (the reference graph is fixed at 256x256). */
int cic_adaptive_forward(cic_plan* plan, const cic_adaptive_io* io, int n_img, int img_h, int img_w,
                         void* d_workspace, size_t workspace_bytes, void* stream);

/* The same forward in three phases, for callers that stream a large batch through the GPU in chunks while host<->device
 * copies are in flight (adaptive_model.predict_pipelined): the convolutions run per chunk, the Dense layers - which stream
 * 1.2 GB of weights whatever the batch - once per batch.  Results are identical to cic_adaptive_forward.
 *   CIC_PHASE_ENCODE  per chunk   conv layers of both encoders (+ attention) and the RD network: image / mask / bpp -> state, rd_params
 *   CIC_PHASE_LATENT  per batch   encoder Dense, latent saliency, quantiser, generator Dense: state -> latents, state.g0
 *   CIC_PHASE_DECODE  per chunk   transposed convs, conv_out, ROI blend: state -> blended, dt, hq_ratio_sum
 * The state buffers are batch-wide device arrays owned by the caller (bf16 elements per tile of the model's H x W input, index
 * [0] = hq, [1] = lq encoder / generator): x1 (H/2)(W/2)64, x2 (H/4)(W/4)128, x3 (H/8)(W/8)256, x4_hi and x4_lo (H/16)(W/16)512,
 * g0 (H/16)(W/16)512.  For ENCODE / DECODE the io pointers address the chunk (its first image / tile) and tile0 is the index of
 * the chunk's first tile in the state buffers; for LATENT io addresses the whole batch (d_bpp, d_hq_latent_q, d_lq_latent_q and
 * the optional latent / symbol / scale outputs).  Tensor-core plans (CIC_PREC_TC) only. */
enum { CIC_PHASE_ENCODE = 1, CIC_PHASE_LATENT = 2, CIC_PHASE_DECODE = 3 };
typedef struct cic_adaptive_state {
  void* x1[2];
  void* x2[2];
  void* x3[2];
  void* x4_hi[2];
  void* x4_lo[2];
  void* g0[2];
} cic_adaptive_state;
int cic_adaptive_forward_phase(cic_plan* plan, const cic_adaptive_io* io, const cic_adaptive_state* state, int phase, int tile0,
                               int n_img, int img_h, int img_w, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- stand-alone operators (also used inside the plans) ----------------------------------- */

/* Keras Conv2D(padding='same') + optional per-channel affine (folded BatchNorm) + activation, fp32 CUDA
 * cores.  d_kernel is (kh,kw,Cin,Cout).  y = act((conv(x)+bias)*scale + shift). */
int cic_conv2d_nhwc_f32(const float* d_x, const float* d_kernel, const float* d_bias, const float* d_scale,
                        const float* d_shift, float* d_y, int batch, int h, int w, int cin, int cout, int kh, int kw,
                        int stride, int act, void* stream);

/* Keras Conv2DTranspose(kernel 4, stride 2, padding='same'); d_kernel is (4,4,Cout,Cin). */
int cic_conv2d_transpose4x4s2_nhwc_f32(const float* d_x, const float* d_kernel, const float* d_bias,
                                       const float* d_scale, const float* d_shift, float* d_y, int batch, int h, int w,
                                       int cin, int cout, int act, void* stream);

/* Keras Dense: y = act(x @ kernel + bias); d_kernel (in,out).  d_workspace may be NULL when
 * cic_dense_workspace_bytes() returns 0. */
size_t cic_dense_workspace_bytes(int batch, int in_dim, int out_dim);
int cic_dense_f32(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int in_dim,
                  int out_dim, int act, void* d_workspace, size_t workspace_bytes, void* stream);

/* Tensor-core (tcgen05, bf16 operands, fp32 accumulate in TMEM) forms of the three operators above, the
 * arithmetic the CIC_PREC_TC plans use.  fp32 tensors in and out; operands are converted to bf16 (and the
 * kernel packed to [Cout][K]) per call, so these are operator-level entry points for parity tests and
 * one-off use - the plans keep packed weights and bf16 activations resident instead.
 *   split != 0 : error-compensated 3-term split-bf16 (hi*hi + lo*hi + hi*lo), the encoder's arithmetic;
 *   split == 0 : single-pass bf16, the decoders' arithmetic.
 * cic_conv2d_nhwc_tc: Conv2D(padding='same', stride 1 or 2) on the channel concatenation of d_x (cin) and
 * the optional d_x2 (cin2; Concatenate + Conv2D, GAN_functions.py:256-268), or with transpose != 0
 * Conv2DTranspose(kernel 4, stride 2, 'same') with d_kernel (4,4,Cout,Cin+Cin2).  Channel counts of the
 * sources must be multiples of 32. */
size_t cic_conv2d_tc_workspace_bytes(int batch, int h, int w, int cin, int cin2, int cout, int kh, int kw, int stride,
                                     int transpose);
int cic_conv2d_nhwc_tc(const float* d_x, const float* d_x2, const float* d_kernel, const float* d_bias,
                       const float* d_scale, const float* d_shift, float* d_y, int batch, int h, int w, int cin, int cin2,
                       int cout, int kh, int kw, int stride, int transpose, int act, int split, void* d_workspace,
                       size_t workspace_bytes, void* stream);
size_t cic_dense_tc_workspace_bytes(int batch, int in_dim, int out_dim);
int cic_dense_tc(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int in_dim, int out_dim,
                 int act, int split, void* d_workspace, size_t workspace_bytes, void* stream);

/* SelfAttention.call (GAN_functions.py:344-369): x (B,h,w,C) -> gamma*softmax(q k^T) v + x.
 * d_wq/d_wk (C, C/8), d_wv (C, C) are the 1x1 conv kernels; workspace from cic_attention_workspace_bytes. */
size_t cic_attention_workspace_bytes(int batch, int tokens, int channels);
int cic_self_attention_f32(const float* d_x, const float* d_wq, const float* d_bq, const float* d_wk,
                           const float* d_bk, const float* d_wv, const float* d_bv, float gamma, float* d_y, int batch,
                           int tokens, int channels, void* d_workspace, size_t workspace_bytes, void* stream);

/* AdaptiveQuantizationLayer.call (GAN_functions.py:435-446): scale = exp(3*qs*(1-sal)),
 * symbols = rint(latent*scale) (half-to-even), deq = symbols/scale.  d_sal, d_qs are (B,).
 * d_deq / d_symbols / d_pre / d_scale may be NULL. */
int cic_quantize_latent(const float* d_latent, const float* d_sal, const float* d_qs, float* d_deq,
                        int32_t* d_symbols, float* d_pre, float* d_scale, int batch, int latent_dim, void* stream);

/* Rate scalars of GAN_functions.py:631-649: t = clip(bpp/5,0,1); thr = 0.9-0.85t; qs = 0.9-0.8t. */
int cic_rate_scalars(const float* d_bpp, float* d_t, float* d_thr, float* d_qs, int n, void* stream);

/* dt = sigmoid((mask^0.7 - thr(bpp))*20); out = hq*dt + lq*(1-dt)  (GAN_functions.py:651-684).
 * d_hq/d_lq/d_out (B,HW,C), d_mask/d_dt (B,HW); d_bpp (B,).  d_out or d_dt may be NULL; when d_hq is
 * NULL only dt is produced.  d_dt_sum (B,) doubles, optional: sum of dt per image (hq_ratio numerator). */
int cic_roi_mask_blend(const float* d_hq, const float* d_lq, const float* d_mask, const float* d_bpp, float* d_out,
                       float* d_dt, double* d_dt_sum, int batch, int hw, int channels, void* stream);

/* hq_ratio for every (image, target bpp) pair in one pass over the mask (GAN_test.py:565-573 runs the
 * whole model per level only to take mean(dt)).  d_ratio is (batch, n_levels) doubles. */
int cic_hq_ratio_sweep(const float* d_mask, const float* d_bpp_levels, int n_levels, double* d_ratio, int batch,
                       int hw, void* stream);

/* Zeroth-order entropy, in bits, of each row of integer symbols (B, L).  No reference counterpart
 * (the reference's bpp is analytic, GAN_test.py:310-325); provided for the north-star's
 * "symbol-histogram bpp estimation".  Symbols are clamped to [-CIC_SYM_MAX, CIC_SYM_MAX]. */
#define CIC_SYM_MAX 1023
int cic_symbol_entropy_bits(const int32_t* d_symbols, double* d_bits, int batch, int latent_dim, void* stream);

/* create_saliency_mask(saliency_map, smooth=True) (GAN_functions.py:199-203; recomputed on the CPU for every image and target bpp
 * at GAN_test.py:279-280): cv2.bilateralFilter(map, 9, 75, 75) -> cv2.GaussianBlur(31x31, sigma 0 = 5.0) -> / max (if max > 0), with
 * OpenCV's semantics (circular 9-tap-wide neighbourhood, BORDER_REFLECT_101, float32).  d_saliency, d_mask (B,H,W) float32; the two
 * may not alias. */
size_t cic_saliency_mask_workspace_bytes(int batch, int h, int w);
int cic_saliency_mask_smooth(const float* d_saliency, float* d_mask, int batch, int h, int w, void* d_workspace,
                             size_t workspace_bytes, void* stream);

/* enhance_saliency_map(saliency_map) (GAN_functions.py:123-157; defined by the reference, called nowhere in it):
 * cv2.bilateralFilter(map, 9, 75, 75) -> cv2.GaussianBlur with 3x3, 9x9 and 15x15 kernels (sigma 0) mixed 0.5 / 0.3 / 0.2 -> ^0.8 ->
 * clip to [0, 1].  d_saliency, d_out (B,H,W) float32, may not alias. */
size_t cic_saliency_enhance_workspace_bytes(int batch, int h, int w);
int cic_saliency_enhance(const float* d_saliency, float* d_out, int batch, int h, int w, void* d_workspace, size_t workspace_bytes,
                         void* stream);

/* create_saliency_mask(saliency_map, threshold, smooth=False) (GAN_functions.py:172-197, :204-206): mask = (map > threshold) as
 * float32 0 / 1.  adaptive != 0: the threshold of every map is min(Otsu of the uint8 map / 255 (cv2.threshold THRESH_OTSU), lower
 * edge of the first of 50 np.histogram bins on [0, 1] whose cumulative share exceeds 0.7) clamped to [0.05, 0.5] - computed on the
 * device and, when d_threshold_out (B,) is not NULL, also returned; adaptive == 0: `threshold` for every map (no workspace). */
size_t cic_saliency_mask_binary_workspace_bytes(int batch);
int cic_saliency_mask_binary(const float* d_saliency, float* d_mask, double threshold, int adaptive, double* d_threshold_out, int batch,
                             int h, int w, void* d_workspace, size_t workspace_bytes, void* stream);

/* compute_saliency_map(image, method) (GAN_functions.py:52-121; called per image and target bpp at GAN_test.py:279, :552 and
 * GAN_train.py:84): cv2.saliency.StaticSaliencySpectralResidual / StaticSaliencyFineGrained (opencv-contrib) on the uint8 image,
 * method 'combined' = 0.6 * spectral + 0.4 * fine (:95), every result divided by its maximum when that is positive (:98-99,
 * :118-119).  d_images (B,H,W,3) uint8 - what GAN_functions.py:63-67 hands to OpenCV - in RGB (rgb = 1: the reference's RGB -> BGR
 * swap of :70-71 happens inside) or BGR order; d_map (B,H,W) float32 in [0, 1]. */
enum cic_saliency_method { CIC_SALIENCY_SPECTRAL_RESIDUAL = 0, CIC_SALIENCY_FINE_GRAINED = 1, CIC_SALIENCY_COMBINED = 2 };
size_t cic_saliency_map_workspace_bytes(int batch, int h, int w);
int cic_saliency_map_u8(const uint8_t* d_images, float* d_map, int batch, int h, int w, int rgb, int method, void* d_workspace,
                        size_t workspace_bytes, void* stream);

/* Entropy coder for the integer latent symbols: the bitstream the reference never writes (its bitrate is the nominal 32 bits per
 * latent element of GAN_test.py:310-325).  Static model per call (histogram over all symbols, normalised to 2^14),
 * rANS with 32 interleaved states per row (one warp per row; rows - tiles - stay independently decodable), symbols clamped to
 * [-CIC_SYM_MAX, CIC_SYM_MAX].  Stream layout (little endian): u32 magic "CICR", version 1, rows, latent_dim, prob_bits 14, alphabet
 * 2047, 2 reserved | u16 freq[2048] | u32 row_offset[rows + 1] | per row: u32 state[32], u16 words[], padded to 4 bytes.
 * cic_rans_encode writes the stream and its length in bytes (*d_nbytes, device); cic_rans_decode restores the symbols exactly.
 * d_symbols (rows, latent_dim) int32; d_stream 4-byte aligned with cic_rans_max_bytes() capacity. */
size_t cic_rans_max_bytes(int rows, int latent_dim);
size_t cic_rans_workspace_bytes(int rows, int latent_dim);
int cic_rans_encode(const int32_t* d_symbols, int rows, int latent_dim, uint8_t* d_stream, size_t stream_capacity,
                    unsigned long long* d_nbytes, void* d_workspace, size_t workspace_bytes, void* stream);
int cic_rans_decode(const uint8_t* d_stream, size_t nbytes, int32_t* d_symbols, int rows, int latent_dim, void* stream);

/* Output stage: the file cv2.imwrite("*.jpg", img) writes (test_autoencoder.py:88-93; GAN_functions.py:41-50 save_image, called at
 * GAN_test.py:390), produced on the device - baseline JPEG, 4:2:0, Annex K Huffman tables, JFIF 1.01 header, libjpeg's integer
 * arithmetic throughout, so the bytes equal OpenCV's (cv2.imencode is the tests' oracle).  d_img (batch, h, w, 3) uint8, channel
 * order BGR like cv2.imwrite's input (rgb = 0) or RGB (rgb = 1: save_image's cvtColor folded in); quality as IMWRITE_JPEG_QUALITY
 * (OpenCV's default is 95).  File b starts at d_out + b * capacity and is d_sizes[b] bytes long; cic_jpeg_max_bytes is the strict
 * worst case, a smaller capacity is allowed: bytes beyond it are dropped and d_sizes[b] still holds the size the file needs.
 * d_workspace 256-byte aligned. */
size_t cic_jpeg_max_bytes(int h, int w);
size_t cic_jpeg_workspace_bytes(int batch, int h, int w);
int cic_jpeg_encode_u8(const uint8_t* d_img, int batch, int h, int w, int rgb, int quality, uint8_t* d_out, size_t capacity,
                       int32_t* d_sizes, void* d_workspace, size_t workspace_bytes, void* stream);

/* (y*255).astype(uint8) - truncation toward zero (test_autoencoder.py:88,96). */
int cic_f32_to_u8_trunc(const float* d_x, uint8_t* d_y, size_t n, float mul, void* stream);

/* GAN pixel conventions on the device (GAN_functions.py:31-37 load, :41-50 save): y = (u8 - 127.5) / 127.5 and
 * y = ((x + 1) * 127.5).astype(uint8) (float32 arithmetic, truncation).  Lets callers move 1 byte per sample over PCIe instead of 4
 * (adaptive_model.predict_phased(..., u8_io=True)). */
int cic_u8_to_f32_signed(const uint8_t* d_x, float* d_y, size_t n, void* stream);
int cic_f32_signed_to_u8(const float* d_x, uint8_t* d_y, size_t n, void* stream);

/* compute_metrics (GAN_functions.py:724-759): inputs (B,H,W,C) float32; v = (x + pre_add) * pre_mul maps
 * [-1,1] to [0,1] (pre_add = 1, pre_mul = 0.5) or is the identity (0, 1).  d_out is (B,4) doubles:
 * psnr (skimage, data_range, all channels jointly), ssim (7x7 uniform window, sample covariance, mean of the
 * per-channel means, cropped by 3 px), mse, sum of squared error. */
int cic_metrics_psnr_ssim_f32(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w,
                              int channels, float pre_add, float pre_mul, float data_range, void* stream);
/* Same outputs with the SSIM window sums in float32 on centred data (every 32x32 tile subtracts its first pixel before the
 * products: variances are shift invariant, and the cancellation in E[x^2] - E[x]^2 that makes scipy accumulate in double
 * disappears where it matters).  HBM-bound instead of conversion-bound; psnr / mse / sse are identical to the call above, ssim
 * agrees with scikit-image to ~1e-6 (tests: 1e-5).  Images with more than 4 channels fall back to the exact kernels. */
int cic_metrics_psnr_ssim_f32_fast(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w, int channels,
                                   float pre_add, float pre_mul, float data_range, void* stream);

/* MS-SSIM per image (BASELINE.json configs[4]).  The reference has no MS-SSIM call site (it evaluates single-scale SSIM,
 * GAN_functions.py:745-748); this is Wang-Simoncelli-Bovik 2003 as tf.image.ssim_multiscale / pytorch_msssim implement it: five
 * scales (weights 0.0448, 0.2856, 0.3001, 0.2363, 0.1333), 11x11 Gaussian window sigma 1.5 as a valid correlation, 2x2 average
 * pooling between scales, per channel, mean over channels.  Inputs (B,H,W,C) float32, v = (x + pre_add) * pre_mul as above; h, w >=
 * 176.  d_out (B,) doubles.  PARITY UNPINNED: checked against a float64 numpy restatement of this definition only. */
size_t cic_msssim_workspace_bytes(int batch, int h, int w, int channels);
int cic_msssim_f32(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w, int channels, float pre_add,
                   float pre_mul, float data_range, void* d_workspace, size_t workspace_bytes, void* stream);

/* Metric sums of one evaluated batch in one launch: d_metrics (n,4) doubles as written by cic_metrics_psnr_ssim_f32*, d_dt_sum (n,)
 * doubles as written by cic_roi_mask_blend / cic_adaptive_forward (sum of dt per image of img_px pixels).  Per image
 * hq_ratio = dt_sum / img_px, total_bits = hq_ratio * latent_hq * 32 + (1 - hq_ratio) * latent_lq * 32, actual_bpp = total_bits /
 * tile_px (GAN_test.py:310-325).  d_out (8,) doubles = {sum psnr, sum ssim, sum mse, sum actual_bpp, sum hq_ratio, 0, n, 0}: the row
 * the ranks all-reduce. */
int cic_metric_sums(const double* d_metrics, const double* d_dt_sum, int n, int img_px, int latent_hq, int latent_lq, int tile_px,
                    double* d_out, void* stream);

/* calculate_mse/psnr/ssim on uint8 BGR images (test_autoencoder.py:49-66).  d_out is (B,4) doubles:
 * psnr (data_range 255), ssim of the cv2 BGR2GRAY images (float64 arithmetic, as scikit-image does for
 * uint8), true mse, and the reference's wrapped uint8 "mse" (mean of ((a-b)**2 mod 256), App. D.1). */
int cic_metrics_psnr_ssim_gray_u8(const uint8_t* d_a, const uint8_t* d_b, double* d_out, int batch, int h, int w,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CIC_H_ */
