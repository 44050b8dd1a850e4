// MS-SSIM per image (BASELINE.json configs[4]: "PSNR/MS-SSIM per image").
//
// The reference evaluates single-scale SSIM only (GAN_functions.py:745-748, cic_metrics_psnr_ssim_f32); MS-SSIM has no call site
// there.  This is the published algorithm (Wang, Simoncelli, Bovik 2003) in the form tf.image.ssim_multiscale / pytorch_msssim use:
// five scales with weights {0.0448, 0.2856, 0.3001, 0.2363, 0.1333}, an 11x11 Gaussian window (sigma 1.5) applied as a VALID
// separable correlation, K1 = 0.01, K2 = 0.03, 2x2 average pooling between scales, per channel
//     msssim_c = prod_{j<5} relu(mean cs_j)^w_j * relu(mean ssim_5)^w_5,        result = mean over channels.
// PARITY UNPINNED: checked against a float64 numpy restatement of this definition only (tests).
//
// Per scale one kernel: a CTA owns a 32x32 tile of the valid output, stages the 42x42 input tile of one channel of both images
// in shared memory CENTRED on the tile's first pixel (the Gaussian-weighted variances are shift invariant, so float32 sums of
// centred values do not suffer the E[x^2] - E[x]^2 cancellation), runs the horizontal then the vertical 11-tap pass for the five
// quantities a, b, aa, bb, ab, and reduces cs and ssim with warp shuffles: one atomicAdd(double) per CTA and quantity.
#include "common.cuh"

namespace cic {

constexpr int MSS_WIN = 11;
constexpr int MSS_T = 32;                     // output tile edge
constexpr int MSS_P = MSS_T + MSS_WIN - 1;    // 42: input tile edge
constexpr int MSS_SCALES = 5;

// exp(-x^2 / (2 * 1.5^2)), x = -5..5, normalised to sum 1 (float64 values rounded to float32)
__device__ constexpr float mss_gauss[MSS_WIN] = {1.0283800845e-03f, 7.5987581352e-03f, 3.6000772128e-02f, 1.0936068951e-01f, 2.1300553771e-01f, 2.6601172486e-01f, 2.1300553771e-01f, 1.0936068951e-01f, 3.6000772128e-02f, 7.5987581352e-03f, 1.0283800845e-03f};

__global__ void __launch_bounds__(256)
msssim_scale_kernel(const float* __restrict__ A, const float* __restrict__ Bm, double* __restrict__ acc, int H, int W, int C, int scale,
                    float pre_add, float pre_mul, float c1, float c2) {
  __shared__ float sa[MSS_P][MSS_P + 1];
  __shared__ float sb[MSS_P][MSS_P + 1];
  __shared__ float sh[5][MSS_P][MSS_T + 1];
  __shared__ double red[2][8];
  const int Ho = H - (MSS_WIN - 1), Wo = W - (MSS_WIN - 1);
  const int tiles_x = (Wo + MSS_T - 1) / MSS_T;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c = blockIdx.y, b = blockIdx.z;
  const int x0 = tx * MSS_T, y0 = ty * MSS_T;
  const size_t base = (size_t)b * H * W * C;
  // centring reference: the tile's first pixel (always inside the image)
  const float ma = __fmul_rn(__fadd_rn(__ldg(A + base + ((size_t)y0 * W + x0) * C + c), pre_add), pre_mul);
  const float mb = __fmul_rn(__fadd_rn(__ldg(Bm + base + ((size_t)y0 * W + x0) * C + c), pre_add), pre_mul);
  for (int i = threadIdx.x; i < MSS_P * MSS_P; i += blockDim.x) {
    const int ly = i / MSS_P, lx = i % MSS_P;
    const int gy = y0 + ly, gx = x0 + lx;
    float va = 0.f, vb = 0.f;
    if (gy < H && gx < W) {
      const size_t idx = base + ((size_t)gy * W + gx) * C + c;
      va = __fmul_rn(__fadd_rn(__ldg(A + idx), pre_add), pre_mul) - ma;
      vb = __fmul_rn(__fadd_rn(__ldg(Bm + idx), pre_add), pre_mul) - mb;
    }
    sa[ly][lx] = va;
    sb[ly][lx] = vb;
  }
  __syncthreads();
  // horizontal pass: 42 rows x 32 columns x 5 quantities
  for (int i = threadIdx.x; i < MSS_P * MSS_T; i += blockDim.x) {
    const int ly = i / MSS_T, ox = i % MSS_T;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int k = 0; k < MSS_WIN; ++k) {
      const float g = mss_gauss[k], va = sa[ly][ox + k], vb = sb[ly][ox + k];
      s0 = fmaf(g, va, s0); s1 = fmaf(g, vb, s1);
      s2 = fmaf(g * va, va, s2); s3 = fmaf(g * vb, vb, s3); s4 = fmaf(g * va, vb, s4);
    }
    sh[0][ly][ox] = s0; sh[1][ly][ox] = s1; sh[2][ly][ox] = s2; sh[3][ly][ox] = s3; sh[4][ly][ox] = s4;
  }
  __syncthreads();
  // vertical pass + the cs / ssim maps of this tile
  float cs_sum = 0.f, ss_sum = 0.f;
  for (int i = threadIdx.x; i < MSS_T * MSS_T; i += blockDim.x) {
    const int oy = i / MSS_T, ox = i % MSS_T;
    if (y0 + oy >= Ho || x0 + ox >= Wo) continue;
    float e[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < MSS_WIN; ++k) {
      const float g = mss_gauss[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) e[q] = fmaf(g, sh[q][oy + k][ox], e[q]);
    }
    const float sxx = e[2] - e[0] * e[0], syy = e[3] - e[1] * e[1], sxy = e[4] - e[0] * e[1];
    const float mux = e[0] + ma, muy = e[1] + mb;
    const float cs = (2.f * sxy + c2) / (sxx + syy + c2);
    const float lum = (2.f * mux * muy + c1) / (mux * mux + muy * muy + c1);
    cs_sum += cs;
    ss_sum += lum * cs;
  }
  double dcs = warp_sum((double)cs_sum), dss = warp_sum((double)ss_sum);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = dcs; red[1][warp] = dss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    double* o = acc + (((size_t)b * C + c) * MSS_SCALES + scale) * 2;
    atomicAdd(o, t0);
    atomicAdd(o + 1, t1);
  }
}

// 2x2 average pooling (floor), NHWC; the first level also applies the pixel normalisation v = (x + pre_add) * pre_mul
__global__ void msssim_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int batch, int H, int W, int C, float pre_add, float pre_mul) {
  const int H2 = H / 2, W2 = W / 2;
  const size_t total = (size_t)batch * H2 * W2 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % W2); r /= W2;
    const int oy = (int)(r % H2);
    const int b = (int)(r / H2);
    const float* p = x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    const float v00 = __fmul_rn(__fadd_rn(__ldg(p), pre_add), pre_mul), v01 = __fmul_rn(__fadd_rn(__ldg(p + C), pre_add), pre_mul);
    const float v10 = __fmul_rn(__fadd_rn(__ldg(p + (size_t)W * C), pre_add), pre_mul), v11 = __fmul_rn(__fadd_rn(__ldg(p + (size_t)W * C + C), pre_add), pre_mul);
    y[i] = 0.25f * ((v00 + v01) + (v10 + v11));
  }
}

__global__ void msssim_finalize_kernel(const double* __restrict__ acc, double* __restrict__ out, int batch, int H, int W, int C) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double wgt[MSS_SCALES] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
  double mean = 0.0;
  for (int c = 0; c < C; ++c) {
    double prod = 1.0;
    int h = H, w = W;
    for (int j = 0; j < MSS_SCALES; ++j) {
      const double n = (double)(h - (MSS_WIN - 1)) * (double)(w - (MSS_WIN - 1));
      const double* a = acc + (((size_t)b * C + c) * MSS_SCALES + j) * 2;
      const double v = (j < MSS_SCALES - 1 ? a[0] : a[1]) / n;
      prod *= pow(v > 0.0 ? v : 0.0, wgt[j]);
      h /= 2; w /= 2;
    }
    mean += prod;
  }
  out[b] = mean / C;
}

}  // namespace cic

using namespace cic;

static size_t mss_align(size_t n) { return (n + 255) & ~(size_t)255; }

extern "C" size_t cic_msssim_workspace_bytes(int batch, int h, int w, int channels) {
  if (batch <= 0 || h <= 0 || w <= 0 || channels <= 0) return 0;
  size_t n = mss_align((size_t)batch * channels * MSS_SCALES * 2 * sizeof(double));
  int hh = h, ww = w;
  for (int j = 1; j < MSS_SCALES; ++j) {
    hh /= 2; ww /= 2;
    n += 2 * mss_align((size_t)batch * hh * ww * channels * sizeof(float));
  }
  return n + 256;
}

extern "C" int cic_msssim_f32(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w, int channels, float pre_add,
                              float pre_mul, float data_range, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch == 0 || (d_a && d_b && d_out), "cic_msssim_f32: null pointer");
  CIC_REQUIRE(batch >= 0 && channels >= 1 && channels <= 65535, "cic_msssim_f32: bad shape");
  CIC_REQUIRE((h >> (MSS_SCALES - 1)) >= MSS_WIN && (w >> (MSS_SCALES - 1)) >= MSS_WIN,
              "cic_msssim_f32: five scales with an 11x11 window need h, w >= 176, got %dx%d", h, w);
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(batch <= 65535, "cic_msssim_f32: at most 65535 images per call");
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_msssim_workspace_bytes(batch, h, w, channels), "cic_msssim_f32: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_workspace;
  double* acc = (double*)ws;
  size_t off = mss_align((size_t)batch * channels * MSS_SCALES * 2 * sizeof(double));
  CIC_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)batch * channels * MSS_SCALES * 2 * sizeof(double), st));
  const float c1 = (float)((0.01 * (double)data_range) * (0.01 * (double)data_range));
  const float c2 = (float)((0.03 * (double)data_range) * (0.03 * (double)data_range));
  const float* pa = d_a;
  const float* pb = d_b;
  float padd = pre_add, pmul = pre_mul;
  int hh = h, ww = w;
  for (int j = 0; j < MSS_SCALES; ++j) {
    const int Ho = hh - (MSS_WIN - 1), Wo = ww - (MSS_WIN - 1);
    dim3 grid(((Wo + MSS_T - 1) / MSS_T) * ((Ho + MSS_T - 1) / MSS_T), channels, batch);
    msssim_scale_kernel<<<grid, 256, 0, st>>>(pa, pb, acc, hh, ww, channels, j, padd, pmul, c1, c2);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("msssim_scale_kernel");
    if (j == MSS_SCALES - 1) break;
    const int h2 = hh / 2, w2 = ww / 2;
    const size_t bytes = mss_align((size_t)batch * h2 * w2 * channels * sizeof(float));
    float* na = (float*)(ws + off);
    float* nb = (float*)(ws + off + bytes);
    off += 2 * bytes;
    const size_t total = (size_t)batch * h2 * w2 * channels;
    const int blocks = (int)((total + 255) / 256 < (size_t)sm_count() * 16 ? (total + 255) / 256 : (size_t)sm_count() * 16);
    msssim_pool_kernel<<<blocks, 256, 0, st>>>(pa, na, batch, hh, ww, channels, padd, pmul);
    msssim_pool_kernel<<<blocks, 256, 0, st>>>(pb, nb, batch, hh, ww, channels, padd, pmul);
    CIC_COUNT_LAUNCH();
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("msssim_pool_kernel");
    pa = na; pb = nb; padd = 0.f; pmul = 1.f;
    hh = h2; ww = w2;
  }
  msssim_finalize_kernel<<<(batch + 127) / 128, 128, 0, st>>>(acc, d_out, batch, h, w, channels);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("msssim_finalize_kernel");
  return CIC_OK;
}
