// Direct convolutions on the CUDA cores for the layers a tensor-core tile cannot take: 3-channel inputs
// (K = taps x 3 is far below one 64-wide K block), written for the FMA pipe instead of a generic GEMM.
//
// conv_k4s2_c3: Conv2D(64, k4, s2, 'same') + bias + LeakyReLU(0.2) of the GAN encoder (GAN_functions.py:300-302),
// fp32 NHWC image in, bf16 (hi, lo) NHWC feature map out - the operand format of the split-bf16 tensor-core
// layers that follow - in one pass (the first version ran a generic fp32 GEMM, wrote fp32, then re-read it to
// split).  Block = 32 x 16 output pixels; the (66 x 34 x 3) input patch and the 48 x 64 weights sit in shared
// memory; a thread owns 2 pixels x 64 channels (128 fp32 accumulators), so one broadcast weight float4 feeds
// 8 FMAs and the kernel is FMA-bound rather than LDS-bound.  Accumulation is fp32 FMA in ascending (ky, kx, c).
#include "common.cuh"
#include "plan.cuh"

namespace cic {

constexpr int DC_TX = 32, DC_TY = 16;  // output tile
constexpr int DC_PW = 2 * DC_TX + 2, DC_PH = 2 * DC_TY + 2;

__global__ void __launch_bounds__(256)
conv_k4s2_c3_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32,
                    int H, int W, int pad_t, int pad_l, int act) {
  __shared__ float patch[DC_PH][DC_PW * 3];
  __shared__ __align__(16) float wsm[48][64];
  const int Ho = H >> 1, Wo = W >> 1;
  const int b = blockIdx.z, oy0 = blockIdx.y * DC_TY, ox0 = blockIdx.x * DC_TX;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 64; i += 256) (&wsm[0][0])[i] = wgt[i];
  // input patch: rows 2*oy0 - pad_t .. +DC_PH, columns (2*ox0 - pad_l) .. +DC_PW, zero outside the image
  const int iy0 = 2 * oy0 - pad_t, ix0 = 2 * ox0 - pad_l;
  const float* xb = x + (size_t)b * H * W * 3;
  for (int i = tid; i < DC_PH * DC_PW * 3; i += 256) {
    const int r = i / (DC_PW * 3), cix = i % (DC_PW * 3);
    const int iy = iy0 + r, ix = ix0 + cix / 3;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xb + ((size_t)iy * W + ix) * 3 + cix % 3);
    patch[r][cix] = v;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;  // pixels (ty, tx) and (ty + 8, tx) of the tile
  float acc[2][64];
#pragma unroll
  for (int j = 0; j < 64; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
#pragma unroll 1
  for (int tap = 0; tap < 16; ++tap) {
    const int ky = tap >> 2, kx = tap & 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a0 = patch[2 * ty + ky][(2 * tx + kx) * 3 + c];
      const float a1 = patch[2 * (ty + 8) + ky][(2 * tx + kx) * 3 + c];
      const float4* wr = reinterpret_cast<const float4*>(&wsm[tap * 3 + c][0]);
#pragma unroll
      for (int n4 = 0; n4 < 16; ++n4) {
        const float4 w4 = wr[n4];
        acc[0][4 * n4] = fmaf(a0, w4.x, acc[0][4 * n4]); acc[0][4 * n4 + 1] = fmaf(a0, w4.y, acc[0][4 * n4 + 1]);
        acc[0][4 * n4 + 2] = fmaf(a0, w4.z, acc[0][4 * n4 + 2]); acc[0][4 * n4 + 3] = fmaf(a0, w4.w, acc[0][4 * n4 + 3]);
        acc[1][4 * n4] = fmaf(a1, w4.x, acc[1][4 * n4]); acc[1][4 * n4 + 1] = fmaf(a1, w4.y, acc[1][4 * n4 + 1]);
        acc[1][4 * n4 + 2] = fmaf(a1, w4.z, acc[1][4 * n4 + 2]); acc[1][4 * n4 + 3] = fmaf(a1, w4.w, acc[1][4 * n4 + 3]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int oy = oy0 + ty + 8 * h, ox = ox0 + tx;
    if (oy >= Ho || ox >= Wo) continue;
    const size_t o = (((size_t)b * Ho + oy) * Wo + ox) * 64;
#pragma unroll
    for (int j8 = 0; j8 < 8; ++j8) {
      float f[8];
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = acc[h][8 * j8 + j];
        if (bias) v = __fadd_rn(v, __ldg(bias + 8 * j8 + j));
        f[j] = act == CIC_ACT_LRELU02 ? fmaxf(v, __fmul_rn(v, 0.2f)) : (act == CIC_ACT_RELU ? fmaxf(v, 0.f) : v);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * j] - __uint_as_float(hi[j] << 16), f[2 * j + 1] - __uint_as_float(hi[j] & 0xFFFF0000u));
        lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      if (out_hi) reinterpret_cast<uint4*>(out_hi + o)[j8] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (out_lo) reinterpret_cast<uint4*>(out_lo + o)[j8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      if (out_f32) {
        reinterpret_cast<float4*>(out_f32 + o)[2 * j8] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4*>(out_f32 + o)[2 * j8 + 1] = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
  }
}

// x (B,H,W,3) fp32 -> Conv2D(64, k4, s2, 'same') + bias + act -> bf16 hi (+ lo) and/or fp32, (B,H/2,W/2,64)
int launch_conv_k4s2_c3(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                        float* out_f32, int batch, int H, int W, int act, cudaStream_t st) {
  CIC_REQUIRE(H % 2 == 0 && W % 2 == 0 && H > 0 && W > 0, "conv_k4s2_c3: H and W must be even");
  CIC_REQUIRE(act == CIC_ACT_NONE || act == CIC_ACT_RELU || act == CIC_ACT_LRELU02, "conv_k4s2_c3: unsupported activation");
  if (batch == 0) return CIC_OK;
  const int Ho = H / 2, Wo = W / 2;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((Wo + DC_TX - 1) / DC_TX, (Ho + DC_TY - 1) / DC_TY, nb);
    const size_t oo = (size_t)b0 * Ho * Wo * 64;
    conv_k4s2_c3_kernel<<<grid, 256, 0, st>>>(x + (size_t)b0 * H * W * 3, wgt, bias, out_hi ? out_hi + oo : nullptr,
                                              out_lo ? out_lo + oo : nullptr, out_f32 ? out_f32 + oo : nullptr, H, W,
                                              same_pad_before(H, 4, 2), same_pad_before(W, 4, 2), act);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("conv_k4s2_c3_kernel");
  }
  return CIC_OK;
}

}  // namespace cic
