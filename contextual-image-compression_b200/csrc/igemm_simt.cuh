// fp32 implicit-GEMM on the CUDA cores: the CIC_PREC_FP32 arithmetic of every conv / transposed
// conv / dense / attention matmul, and the on-device reference the tcgen05 kernels are checked
// against layer by layer.
#pragma once
#include "common.cuh"

namespace cic {

struct ConvSrc {
  const float* ptr;  // NHWC, `ld` floats per pixel, this source starts at channel 0 of its pixel record
  int C;             // channels taken from this source
  int ld;            // floats per pixel of the underlying buffer
  int up;            // 1: nearest-neighbour x2 up-sampled view (UpSampling2D, train_autoencoder.py:22,29)
};

struct IGemmParams {
  ConvSrc src[2];  // channel-concatenated sources (Concatenate of [x, skip])
  int nsrc;
  int Cin;         // src[0].C + src[1].C
  int batch;
  int H, W;        // logical input size per batch item (after up-sampling)
  int Ho, Wo;      // output positions iterated per batch item
  int kh, kw, stride, pad_t, pad_l;  // input row = oy*stride + ky - pad_t
  // B operand: row-major [K][N] (K ordered ky,kx,ci) or, when b_trans, [N][K]
  const float* Bmat;
  int N, ldb, b_trans;
  long long b_batch_stride;  // != 0: one B matrix per batch item (attention)
  // epilogue: v = act((alpha*acc + bias)*scale + shift) + residual
  const float* bias;
  const float* scale;
  const float* shift;
  const float* residual;  // same addressing as out
  float alpha;
  int act;
  float* out;
  int out_ld, out_coff;  // floats per output pixel record, channel offset
  int out_H, out_W;      // output image size
  int out_ys, out_xs, out_y0, out_x0;  // output pixel = (oy*out_ys + out_y0, ox*out_xs + out_x0)
  // split-K: partial sums to `partial` [splits][M][N], epilogue deferred to splitk_reduce
  int splits;
  float* partial;
};

int launch_igemm(const IGemmParams& p, cudaStream_t st);
int launch_splitk_reduce(const IGemmParams& p, cudaStream_t st);

// small-Cout direct convolution (Cout <= 4): L7 (64->3, sigmoid), G5 (32->3, tanh)
struct SmallNParams {
  ConvSrc src[2];
  int nsrc, Cin, batch, H, W, kh, kw, pad_t, pad_l;
  const float* Wmat;  // [K][N]
  const float* bias;
  int N, act;
  float* out;  // (batch, H, W, N)
};
int launch_conv_small_n(const SmallNParams& p, cudaStream_t st);

int launch_maxpool2x2(const float* x, float* y, int batch, int H, int W, int C, cudaStream_t st);
int launch_softmax_rows(float* x, long long rows, int cols, cudaStream_t st);
// mean over HW of (B,HW,C) -> (B, ldo) at column 0.., used by GlobalAveragePooling2D (GAN_functions.py:515)
int launch_global_avg_pool(const float* x, float* y, int batch, int hw, int C, int ldo, cudaStream_t st);
int launch_tile_gather(const float* img, float* tiles, int n_img, int H, int W, int C, int tile, cudaStream_t st);
int launch_tile_scatter(const float* tiles, float* img, int n_img, int H, int W, int C, int tile, cudaStream_t st);

}  // namespace cic
