// Direct convolutions on the CUDA cores for the layers a tensor-core tile cannot take: 1- and 3-channel inputs
// (K = taps x Cin is far below one 32-wide K block), written for the FMA pipe instead of a generic GEMM, and
// producing the bf16 operand format of the tensor-core layers that follow in the same pass.
//
//   conv_k4s2_c3  Conv2D(64, k4, s2, 'same') + LeakyReLU   GAN encoder conv1     (GAN_functions.py:300-302)
//   conv_k3s1_c3  Conv2D(32, k3, 'same') + ReLU + MaxPool  autoencoder conv1     (train_autoencoder.py:14-15)
//   conv_k3s2_c1  Conv2D(32, k3, s2, 'same') + LeakyReLU   RD-optimizer conv1    (GAN_functions.py:511-512)
//
// One block = 256 threads = 32 x 8 thread grid; the input patch of the block's output tile and the whole
// weight matrix sit in shared memory; a thread owns NPX pixels x COUT channels of fp32 accumulators, so one
// broadcast weight float4 feeds 4 * NPX FMAs and the kernel is FMA-bound rather than LDS-bound.  Accumulation
// is fp32 FMA in ascending (ky, kx, c) - the order of the fp32 reference path.
#include "common.cuh"
#include "plan.cuh"

namespace cic {

template <int CIN, int KS, int STRIDE, int COUT, int NPX, bool POOL>
struct DcCfg {
  static constexpr int kTileW = POOL ? 64 : 32;
  static constexpr int kTileH = POOL ? 16 : 8 * NPX;
  static constexpr int kPW = (kTileW - 1) * STRIDE + KS;
  static constexpr int kPH = (kTileH - 1) * STRIDE + KS;
  static constexpr int kK = KS * KS * CIN;
  static_assert(!POOL || NPX == 4, "pooling threads own a 2x2 pixel block");
};

template <int CIN, int KS, int STRIDE, int COUT, int NPX, bool POOL>
__global__ void __launch_bounds__(256)
direct_conv_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                   __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32,
                   __nv_bfloat16* __restrict__ pool_hi, int H, int W, int Ho, int Wo, int pad_t, int pad_l, int act, const TileMap tm) {
  using Cfg = DcCfg<CIN, KS, STRIDE, COUT, NPX, POOL>;
  __shared__ float patch[Cfg::kPH][Cfg::kPW * CIN];
  __shared__ __align__(16) float wsm[Cfg::kK][COUT];
  const int b = blockIdx.z, oy0 = blockIdx.y * Cfg::kTileH, ox0 = blockIdx.x * Cfg::kTileW;
  const int tid = threadIdx.x;
  for (int i = tid; i < Cfg::kK * COUT; i += 256) (&wsm[0][0])[i] = wgt[i];
  const int iy0 = STRIDE * oy0 - pad_t, ix0 = STRIDE * ox0 - pad_l;
  // input item b: a dense (H, W, CIN) image, or tile (ty, tx) of a larger image (zero padding at the tile border:
  // tiles are coded independently)
  const float* xb;
  size_t row_stride;
  int vh = H, vw = W;  // rows / columns of this item that exist in the image: a ragged last tile replicates the image edge
  if (tm.tiles_x) {
    const int tpi = tm.tiles_x * tm.tiles_y, img = b / tpi, t = b % tpi;
    const int gy0 = (t / tm.tiles_x) * H, gx0 = (t % tm.tiles_x) * W;
    xb = x + (((size_t)img * tm.IH + (size_t)gy0) * tm.IW + (size_t)gx0) * CIN;
    row_stride = (size_t)tm.IW * CIN;
    vh = min(H, tm.IH - gy0);
    vw = min(W, tm.IW - gx0);
  } else {
    xb = x + (size_t)b * H * W * CIN;
    row_stride = (size_t)W * CIN;
  }
  {
    // all loads of the patch in flight before the first shared-memory store: with one dependent load per loop iteration the block
    // paid ~9 serial DRAM latencies and the RD conv1 launch was latency-bound at 0.40 ms for 0.55 GB (profiles/r01_ncu_final.md)
    constexpr int kElems = Cfg::kPH * Cfg::kPW * CIN, kPer = (kElems + 255) / 256;
    float v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * 256;
      const int r = i / (Cfg::kPW * CIN), cix = i % (Cfg::kPW * CIN);
      const int iy = iy0 + r, ix = ix0 + cix / CIN;
      v[k] = 0.f;
      if (i < kElems && iy >= 0 && iy < H && ix >= 0 && ix < W)
        v[k] = __ldg(xb + (size_t)min(iy, vh - 1) * row_stride + (size_t)min(ix, vw - 1) * CIN + cix % CIN);
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * 256;
      if (i < kElems) patch[i / (Cfg::kPW * CIN)][i % (Cfg::kPW * CIN)] = v[k];
    }
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  int py[NPX], pxx[NPX];  // tile-local output pixels of this thread
#pragma unroll
  for (int j = 0; j < NPX; ++j) {
    py[j] = POOL ? 2 * ty + (j >> 1) : ty + 8 * j;
    pxx[j] = POOL ? 2 * tx + (j & 1) : tx;
  }
  float acc[NPX][COUT];
#pragma unroll
  for (int j = 0; j < NPX; ++j)
#pragma unroll
    for (int n = 0; n < COUT; ++n) acc[j][n] = 0.f;
#pragma unroll 1
  for (int tap = 0; tap < KS * KS; ++tap) {
    const int ky = tap / KS, kx = tap % KS;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      float a[NPX];
#pragma unroll
      for (int j = 0; j < NPX; ++j) a[j] = patch[STRIDE * py[j] + ky][(STRIDE * pxx[j] + kx) * CIN + c];
      const float4* wr = reinterpret_cast<const float4*>(&wsm[tap * CIN + c][0]);
#pragma unroll
      for (int n4 = 0; n4 < COUT / 4; ++n4) {
        const float4 w4 = wr[n4];
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
          acc[j][4 * n4] = fmaf(a[j], w4.x, acc[j][4 * n4]);
          acc[j][4 * n4 + 1] = fmaf(a[j], w4.y, acc[j][4 * n4 + 1]);
          acc[j][4 * n4 + 2] = fmaf(a[j], w4.z, acc[j][4 * n4 + 2]);
          acc[j][4 * n4 + 3] = fmaf(a[j], w4.w, acc[j][4 * n4 + 3]);
        }
      }
    }
  }
  // bias + activation in place
#pragma unroll
  for (int j = 0; j < NPX; ++j)
#pragma unroll
    for (int n = 0; n < COUT; ++n) {
      float v = acc[j][n];
      if (bias) v = __fadd_rn(v, __ldg(bias + n));
      acc[j][n] = act == CIC_ACT_LRELU02 ? fmaxf(v, __fmul_rn(v, 0.2f)) : (act == CIC_ACT_RELU ? fmaxf(v, 0.f) : v);
    }
#pragma unroll
  for (int j = 0; j < NPX; ++j) {
    const int oy = oy0 + py[j], ox = ox0 + pxx[j];
    if (oy >= Ho || ox >= Wo) continue;
    const size_t o = (((size_t)b * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
    for (int j16 = 0; j16 < COUT / 16; ++j16) {  // 16 channels = one whole 32-byte sector per 256-bit store
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float f0 = acc[j][16 * j16 + 2 * q], f1 = acc[j][16 * j16 + 2 * q + 1];
        const __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
        hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(f0 - __uint_as_float(hi[q] << 16), f1 - __uint_as_float(hi[q] & 0xFFFF0000u));
        lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      if (out_hi) st_global_v8(out_hi + o + 16 * j16, hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]);
      if (out_lo) st_global_v8(out_lo + o + 16 * j16, lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]);
      if (out_f32) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          st_global_v8(out_f32 + o + 16 * j16 + 8 * h, __float_as_uint(acc[j][16 * j16 + 8 * h]), __float_as_uint(acc[j][16 * j16 + 8 * h + 1]),
                       __float_as_uint(acc[j][16 * j16 + 8 * h + 2]), __float_as_uint(acc[j][16 * j16 + 8 * h + 3]),
                       __float_as_uint(acc[j][16 * j16 + 8 * h + 4]), __float_as_uint(acc[j][16 * j16 + 8 * h + 5]),
                       __float_as_uint(acc[j][16 * j16 + 8 * h + 6]), __float_as_uint(acc[j][16 * j16 + 8 * h + 7]));
      }
    }
  }
  if (POOL && pool_hi) {  // MaxPooling2D((2,2), 'same') of this thread's 2x2 block (even H, W: no padding)
    const int oy = (oy0 >> 1) + ty, ox = (ox0 >> 1) + tx;
    const int Hp = Ho >> 1, Wp = Wo >> 1;
    if (oy < Hp && ox < Wp) {
      const size_t o = (((size_t)b * Hp + oy) * Wp + ox) * COUT;
#pragma unroll
      for (int j8 = 0; j8 < COUT / 8; ++j8) {
        uint32_t hi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = 8 * j8 + 2 * q;
          const float m0 = fmaxf(fmaxf(acc[0][n], acc[1][n]), fmaxf(acc[2][n], acc[3][n]));
          const float m1 = fmaxf(fmaxf(acc[0][n + 1], acc[1][n + 1]), fmaxf(acc[2][n + 1], acc[3][n + 1]));
          const __nv_bfloat162 hh = __floats2bfloat162_rn(m0, m1);
          hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        reinterpret_cast<uint4*>(pool_hi + o)[j8] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      }
    }
  }
}

template <int CIN, int KS, int STRIDE, int COUT, int NPX, bool POOL>
static int launch_direct(const char* name, const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi,
                         __nv_bfloat16* out_lo, float* out_f32, __nv_bfloat16* pool_hi, int batch, int H, int W, int act,
                         const TileMap& tm, cudaStream_t st) {
  using Cfg = DcCfg<CIN, KS, STRIDE, COUT, NPX, POOL>;
  CIC_REQUIRE(H > 0 && W > 0, "%s: bad image size", name);
  CIC_REQUIRE(act == CIC_ACT_NONE || act == CIC_ACT_RELU || act == CIC_ACT_LRELU02, "%s: unsupported activation", name);
  CIC_REQUIRE(!POOL || (H % 2 == 0 && W % 2 == 0), "%s: pooling needs even H, W", name);
  if (batch == 0) return CIC_OK;
  const int Ho = same_out(H, STRIDE), Wo = same_out(W, STRIDE);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((Wo + Cfg::kTileW - 1) / Cfg::kTileW, (Ho + Cfg::kTileH - 1) / Cfg::kTileH, nb);
    const size_t oo = (size_t)b0 * Ho * Wo * COUT, po = (size_t)b0 * (Ho / 2) * (Wo / 2) * COUT;
    CIC_REQUIRE(!tm.tiles_x || batch <= 65535, "%s: tiled input supports at most 65535 tiles per launch", name);
    direct_conv_kernel<CIN, KS, STRIDE, COUT, NPX, POOL><<<grid, 256, 0, st>>>(
        x + (tm.tiles_x ? 0 : (size_t)b0 * H * W * CIN), wgt, bias, out_hi ? out_hi + oo : nullptr, out_lo ? out_lo + oo : nullptr,
        out_f32 ? out_f32 + oo : nullptr, pool_hi ? pool_hi + po : nullptr, H, W, Ho, Wo, same_pad_before(H, KS, STRIDE),
        same_pad_before(W, KS, STRIDE), act, tm);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH(name);
    g_last_kernel_kind = KK_DIRECT;
  }
  return CIC_OK;
}

// x (B,H,W,3) fp32 -> Conv2D(64, k4, s2, 'same') + bias + act -> bf16 hi (+ lo) and/or fp32, (B,H/2,W/2,64)
int launch_conv_k4s2_c3(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                        float* out_f32, int batch, int H, int W, int act, const TileMap& tm, cudaStream_t st) {
  CIC_REQUIRE(H % 2 == 0 && W % 2 == 0, "conv_k4s2_c3: H and W must be even");
  return launch_direct<3, 4, 2, 64, 2, false>("conv_k4s2_c3_kernel", x, wgt, bias, out_hi, out_lo, out_f32, nullptr, batch, H, W, act, tm, st);
}

// x (B,H,W,3) fp32 -> Conv2D(32, k3, 'same') + bias + act -> bf16 (B,H,W,32) and its 2x2 max-pool (B,H/2,W/2,32)
int launch_conv_k3s1_c3_pool(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* pool_hi,
                             int batch, int H, int W, int act, cudaStream_t st) {
  return launch_direct<3, 3, 1, 32, 4, true>("conv_k3s1_c3_pool_kernel", x, wgt, bias, out_hi, nullptr, nullptr, pool_hi, batch, H, W, act, TileMap(), st);
}

// x (B,H,W,1) fp32 -> Conv2D(32, k3, s2, 'same') + bias + act -> bf16 hi (+ lo), (B,ceil(H/2),ceil(W/2),32)
int launch_conv_k3s2_c1(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                        int batch, int H, int W, int act, const TileMap& tm, cudaStream_t st) {
  // one pixel per thread: half the accumulator registers, twice the resident CTAs - the kernel is latency-bound, not FMA-bound
  // (r02 ncu: 24 % of the warp slots occupied, long-scoreboard stalls lead)
  if (CIC_KNOB("CIC_RD_NPX", 1) == 1)
    return launch_direct<1, 3, 2, 32, 1, false>("conv_k3s2_c1_kernel", x, wgt, bias, out_hi, out_lo, nullptr, nullptr, batch, H, W, act, tm, st);
  return launch_direct<1, 3, 2, 32, 2, false>("conv_k3s2_c1_kernel", x, wgt, bias, out_hi, out_lo, nullptr, nullptr, batch, H, W, act, tm, st);
}

}  // namespace cic
