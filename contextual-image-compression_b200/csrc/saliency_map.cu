// compute_saliency_map(image, method) on the GPU (GAN_functions.py:52-121): the two opencv-contrib detectors the
// reference runs on the CPU for every image and target bpp (GAN_test.py:279, :552; GAN_train.py:84), and their 0.6 / 0.4 mix.
//
//   StaticSaliencySpectralResidual (Hou & Zhang 2007; opencv_contrib modules/saliency/src/staticSaliencySpectralResidual.cpp)
//     gray u8 -> 64x64 (INTER_LINEAR_EXACT, bit-exact fixed point) -> DFT -> log amplitude minus its 3x3 mean -> inverse DFT with
//     the original phase -> |.| -> Gaussian 5x5 sigma 8 -> square -> / max -> float32 -> bilinear resize to the image
//       sal_gray_kernel, sal_spectral_kernel (one CTA per image, the whole 64x64 pipeline in shared memory, float64)
//   StaticSaliencyFineGrained (Montabone & Soto 2010; .../staticSaliencyFineGrained.cpp)
//     gray u8 -> GaussianBlur 3x3 twice (fixed point) -> float32 integral image -> centre-surround "on" / "off" differences at six
//     neighbourhoods (uchar truncation) -> sums -> / max sum -> on + off -> / max -> uint8 -> float32 / 255
//       sal_gauss3_kernel x2, sal_rowscan_kernel, sal_colscan_kernel, sal_fg_scales_kernel, sal_fg_norm_kernel
//   mix: 0.6 * spectral + 0.4 * fine (float32), / max                      sal_combine_kernel, sal_map_scale_kernel
//
// opencv-contrib is not installed anywhere this code is tested, so the composition is a restatement of the published source
// (the tests' CPU restatement says so: parity unpinned); every OpenCV *core* routine inside it (cvtColor, resize, GaussianBlur, integral,
// dft, blur) is pinned: the integer stages are bit-exact against the real cv2 calls, the float64 stage agrees to ~5e-5 (cv2's
// cartToPolar / polarToCart work in float32 with a polynomial arctangent; this kernel keeps the phase exactly).
#include "common.cuh"

#include <cfloat>
#include <cmath>

namespace cic {
namespace {

constexpr int SR = 64, SR2 = SR * SR;

__device__ __forceinline__ int refl101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// cv2.cvtColor(BGR2GRAY) on uint8: 15-bit fixed point, rounded (OpenCV 4.x)
__device__ __forceinline__ int gray_of(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15; }

__global__ void __launch_bounds__(256)
sal_gray_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ gray, size_t npix, int rgb) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = img[3 * i], c1 = img[3 * i + 1], c2 = img[3 * i + 2];
    gray[i] = (uint8_t)(rgb ? gray_of(c2, c1, c0) : gray_of(c0, c1, c2));
  }
}

// interpolationLinear<ufixedpoint16>::getCoeffs of resize_bitExact: source offset and the 8-bit weight of the second tap
__device__ __forceinline__ void exact_coeff(int d, int src, int dst, int& ofs, int& c1) {
  const double scale = (double)src / (double)dst;
  const double f = __dsub_rn(__dmul_rn(scale, (double)d + 0.5), 0.5);
  const int i = (int)floor(f);
  ofs = 0;
  c1 = 0;
  if (i >= 0 && src > 1) {
    if (i < src - 1) {
      ofs = i;
      c1 = (int)rint(__dmul_rn(__dsub_rn(f, (double)i), 256.0));
    } else {
      ofs = src - 1;
    }
  }
}

struct Gauss5 { double k[5]; };

// One CTA per image.  Shared memory: re, im, tr, ti (4 x 32 KB float64) + twiddles.
__global__ void __launch_bounds__(1024)
sal_spectral_kernel(const uint8_t* __restrict__ gray, float* __restrict__ small, int H, int W, Gauss5 gk) {
  extern __shared__ double sm[];
  double* re = sm;
  double* im = sm + SR2;
  double* tr = sm + 2 * SR2;
  double* ti = sm + 3 * SR2;
  double* twc = sm + 4 * SR2;
  double* tws = twc + SR;
  __shared__ double red[32];
  const int tid = threadIdx.x;
  const uint8_t* g = gray + (size_t)blockIdx.x * H * W;
  if (tid < SR) sincospi((double)tid / 32.0, &tws[tid], &twc[tid]);       // e^{2 pi i t / 64}

  // 64x64 INTER_LINEAR_EXACT: rows to 8.8 fixed point, columns to 16.16, round half up
  for (int e = tid; e < SR2; e += blockDim.x) {
    const int dy = e >> 6, dx = e & 63;
    int ox, cx1, oy, cy1;
    exact_coeff(dx, W, SR, ox, cx1);
    exact_coeff(dy, H, SR, oy, cy1);
    const int ox1 = min(ox + 1, W - 1), oy1 = min(oy + 1, H - 1);
    const int r0 = g[(size_t)oy * W + ox] * (256 - cx1) + g[(size_t)oy * W + ox1] * cx1;
    const int r1 = g[(size_t)oy1 * W + ox] * (256 - cx1) + g[(size_t)oy1 * W + ox1] * cx1;
    const unsigned v = ((unsigned)r0 * (unsigned)(256 - cy1) + (unsigned)r1 * (unsigned)cy1 + (1u << 15)) >> 16;
    re[e] = (double)v;
    im[e] = 0.0;
  }
  __syncthreads();

  // forward DFT (e^{-i}), rows then columns; inverse (e^{+i}, unscaled like cv2.dft(DFT_INVERSE)) after the spectral edit
  auto dft_rows = [&](const double* ar, const double* ai, double* br, double* bi, double sign) {
    for (int e = tid; e < SR2; e += blockDim.x) {
      const int y = e >> 6, k = e & 63;
      double sr_ = 0.0, si_ = 0.0;
      for (int x = 0; x < SR; ++x) {
        const double c = twc[(k * x) & 63], s = sign * tws[(k * x) & 63];
        const double a = ar[y * SR + x], b = ai[y * SR + x];
        sr_ += a * c - b * s;
        si_ += a * s + b * c;
      }
      br[e] = sr_;
      bi[e] = si_;
    }
  };
  auto dft_cols = [&](const double* ar, const double* ai, double* br, double* bi, double sign) {
    for (int e = tid; e < SR2; e += blockDim.x) {
      const int k2 = e >> 6, k = e & 63;
      double sr_ = 0.0, si_ = 0.0;
      for (int y = 0; y < SR; ++y) {
        const double c = twc[(k2 * y) & 63], s = sign * tws[(k2 * y) & 63];
        const double a = ar[y * SR + k], b = ai[y * SR + k];
        sr_ += a * c - b * s;
        si_ += a * s + b * c;
      }
      br[e] = sr_;
      bi[e] = si_;
    }
  };
  dft_rows(re, im, tr, ti, -1.0);
  __syncthreads();
  dft_cols(tr, ti, re, im, -1.0);
  __syncthreads();

  // log amplitude -> tr
  for (int e = tid; e < SR2; e += blockDim.x) tr[e] = log(hypot(re[e], im[e]));
  __syncthreads();
  // gain = exp(logA - blur3x3(logA)); spectrum <- gain * e^{i phase}
  for (int e = tid; e < SR2; e += blockDim.x) {
    const int y = e >> 6, x = e & 63;
    const int xl = refl101(x - 1, SR), xr = refl101(x + 1, SR), yu = refl101(y - 1, SR), yd = refl101(y + 1, SR);
    auto hs = [&](int yy) { return __dadd_rn(__dadd_rn(tr[yy * SR + xl], tr[yy * SR + x]), tr[yy * SR + xr]); };
    const double blur = __dmul_rn(__dadd_rn(__dadd_rn(hs(yu), hs(y)), hs(yd)), 1.0 / 9.0);
    const double gain = exp(__dsub_rn(tr[e], blur));
    const double amp = hypot(re[e], im[e]);
    const double ur = amp > 0.0 ? re[e] / amp : 1.0, ui = amp > 0.0 ? im[e] / amp : 0.0;
    re[e] = ur * gain;
    im[e] = ui * gain;
  }
  __syncthreads();
  dft_rows(re, im, tr, ti, 1.0);
  __syncthreads();
  dft_cols(tr, ti, re, im, 1.0);
  __syncthreads();
  for (int e = tid; e < SR2; e += blockDim.x) tr[e] = hypot(re[e], im[e]);
  __syncthreads();
  // Gaussian 5x5, sigma 8, BORDER_REFLECT_101: rows -> ti, columns -> re; then squared
  for (int e = tid; e < SR2; e += blockDim.x) {
    const int y = e >> 6, x = e & 63;
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < 5; ++t) a = __dadd_rn(a, __dmul_rn(gk.k[t], tr[y * SR + refl101(x + t - 2, SR)]));
    ti[e] = a;
  }
  __syncthreads();
  double mx = 0.0;
  for (int e = tid; e < SR2; e += blockDim.x) {
    const int y = e >> 6, x = e & 63;
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < 5; ++t) a = __dadd_rn(a, __dmul_rn(gk.k[t], ti[refl101(y + t - 2, SR) * SR + x]));
    a = __dmul_rn(a, a);
    re[e] = a;
    mx = fmax(mx, a);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, red[i]);
  float* out = small + (size_t)blockIdx.x * SR2;
  for (int e = tid; e < SR2; e += blockDim.x) out[e] = (float)__ddiv_rn(re[e], mx);
}

// cv2.GaussianBlur(u8, (3, 3), 0): [1 2 1] x [1 2 1] / 16, round half up, BORDER_REFLECT_101
__global__ void __launch_bounds__(256)
sal_gauss3_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ y, int H, int W) {
  const uint8_t* src = x + (size_t)blockIdx.y * H * W;
  uint8_t* dst = y + (size_t)blockIdx.y * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int py = i / W, px = i % W;
    const int xl = refl101(px - 1, W), xr = refl101(px + 1, W);
    const uint8_t* r0 = src + (size_t)refl101(py - 1, H) * W;
    const uint8_t* r1 = src + (size_t)py * W;
    const uint8_t* r2 = src + (size_t)refl101(py + 1, H) * W;
    const int v = (r0[xl] + 2 * r0[px] + r0[xr]) + 2 * (r1[xl] + 2 * r1[px] + r1[xr]) + (r2[xl] + 2 * r2[px] + r2[xr]);
    dst[i] = (uint8_t)((v + 8) >> 4);
  }
}

// cv2.integral(u8, CV_32F), part 1: the running sum of every row (exact integers; float32 holds them for W <= 65793) goes to row
// y + 1 of the (H+1, W+1) image; one warp per row
__global__ void __launch_bounds__(256)
sal_rowscan_kernel(const uint8_t* __restrict__ gray, float* __restrict__ integ, int H, int W, int rows_total) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows_total) return;
  const int b = row / H, y = row % H;
  const uint8_t* src = gray + ((size_t)b * H + y) * W;
  float* dst = integ + ((size_t)b * (H + 1) + y + 1) * (W + 1);
  if (lane == 0) dst[0] = 0.f;
  int carry = 0;
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    int v = x < W ? src[x] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    v += carry;
    if (x < W) dst[x + 1] = (float)v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}

// part 2: sum[y+1][x] = sum[y][x] + rowsum[y][x] in float32, top to bottom (one rounding per element, OpenCV's order)
__global__ void __launch_bounds__(128)
sal_colscan_kernel(float* __restrict__ integ, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x > W) return;
  float* p = integ + (size_t)blockIdx.y * (H + 1) * (W + 1) + x;
  const size_t pitch = W + 1;
  p[0] = 0.f;
  float prev = 0.f;
  int y = 1;
  for (; y + 7 <= H; y += 8) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = p[(size_t)(y + j) * pitch];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      prev = __fadd_rn(prev, r[j]);
      p[(size_t)(y + j) * pitch] = prev;
    }
  }
  for (; y <= H; ++y) {
    prev = __fadd_rn(prev, p[(size_t)y * pitch]);
    p[(size_t)y * pitch] = prev;
  }
}

__constant__ int c_fg_nb[6] = {12, 24, 48, 28, 56, 112};        // 3*4, 3*4*2, 3*4*2*2, 7*4, 7*4*2, 7*4*2*2

// getIntensityScaled + getMean for the six neighbourhoods and the per-pixel sums of mixScales; stats[b] = {max on sum, max off
// sum, max on u8, max off u8}
__global__ void __launch_bounds__(256)
sal_fg_scales_kernel(const uint8_t* __restrict__ gray, const float* __restrict__ integ, uint16_t* __restrict__ on_sum,
                     uint16_t* __restrict__ off_sum, int* __restrict__ stats, int H, int W) {
  const int b = blockIdx.y;
  const uint8_t* g = gray + (size_t)b * H * W;
  const float* I = integ + (size_t)b * (H + 1) * (W + 1);
  const int pitch = W + 1;
  int mon = 0, moff = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i % W;
    const float c = (float)g[i];
    int son = 0, soff = 0;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int nb = c_fg_nb[s];
      const int x1 = min(max(x - nb + 1, 0), W), y1 = min(max(y - nb + 1, 0), H);
      const int x2 = min(max(x + nb + 1, 0), W), y2 = min(max(y + nb + 1, 0), H);
      float v = __fadd_rn(__ldg(I + (size_t)y2 * pitch + x2), __ldg(I + (size_t)y1 * pitch + x1));
      v = __fsub_rn(v, __ldg(I + (size_t)y2 * pitch + x1));
      v = __fsub_rn(v, __ldg(I + (size_t)y1 * pitch + x2));
      const float mean = __fdiv_rn(__fsub_rn(v, c), (float)((x2 - x1) * (y2 - y1) - 1));
      const float on = __fsub_rn(c, mean), off = __fsub_rn(mean, c);
      if (on > 0.f) son += __float2int_rz(on) & 0xFF;
      if (off > 0.f) soff += __float2int_rz(off) & 0xFF;
    }
    on_sum[(size_t)b * H * W + i] = (uint16_t)son;
    off_sum[(size_t)b * H * W + i] = (uint16_t)soff;
    mon = max(mon, son);
    moff = max(moff, soff);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mon = max(mon, __shfl_xor_sync(0xffffffffu, mon, o));
    moff = max(moff, __shfl_xor_sync(0xffffffffu, moff, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(stats + 4 * b, mon);
    atomicMax(stats + 4 * b + 1, moff);
  }
}

// mixScales: (uchar)(255. * (float)(sum / (float)max_sum)) for both polarities; maxima of the two uint8 maps for mixOnOff
__global__ void __launch_bounds__(256)
sal_fg_norm_kernel(const uint16_t* __restrict__ on_sum, const uint16_t* __restrict__ off_sum, uint8_t* __restrict__ on_u8,
                   uint8_t* __restrict__ off_u8, int* __restrict__ stats, int hw) {
  const int b = blockIdx.y;
  const int pon = stats[4 * b], poff = stats[4 * b + 1];
  int mon = 0, moff = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const size_t j = (size_t)b * hw + i;
    const int a = pon > 0 ? __double2int_rz(255.0 * (double)__fdiv_rn((float)on_sum[j], (float)pon)) : 0;
    const int c = poff > 0 ? __double2int_rz(255.0 * (double)__fdiv_rn((float)off_sum[j], (float)poff)) : 0;
    on_u8[j] = (uint8_t)a;
    off_u8[j] = (uint8_t)c;
    mon = max(mon, a);
    moff = max(moff, c);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mon = max(mon, __shfl_xor_sync(0xffffffffu, mon, o));
    moff = max(moff, __shfl_xor_sync(0xffffffffu, moff, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(stats + 4 * b + 2, mon);
    atomicMax(stats + 4 * b + 3, moff);
  }
}

// cv2.resize(INTER_LINEAR) coefficient for one float32 axis
__device__ __forceinline__ void lin_coeff(int d, int src, int dst, int& i0, float& f) {
  const double scale = 1.0 / ((double)dst / (double)src);
  f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  i0 = (int)floorf(f);
  f = __fsub_rn(f, (float)i0);
  if (i0 < 0) { f = 0.f; i0 = 0; }
  if (i0 >= src - 1) { f = 0.f; i0 = src - 1; }
}

// method 0: spectral residual resized to the image; 1: fine grained (mixOnOff, / 255); 2: 0.6 * spectral + 0.4 * fine.
// Writes the un-normalised map and its per-image maximum (all values are >= 0: int ordering of the float bits).
__global__ void __launch_bounds__(256)
sal_combine_kernel(const float* __restrict__ small, const uint8_t* __restrict__ on_u8, const uint8_t* __restrict__ off_u8,
                   const int* __restrict__ stats, float* __restrict__ map, int* __restrict__ peak_bits, int H, int W, int method) {
  const int b = blockIdx.y;
  const float* s = small + (size_t)b * SR2;
  int fg_peak = 0;
  if (method != 0) fg_peak = max(stats[4 * b + 2], stats[4 * b + 3]);
  float mx = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i % W;
    float sr = 0.f, fg = 0.f;
    if (method != 1) {
      int ix, iy;
      float fx, fy;
      lin_coeff(x, SR, W, ix, fx);
      lin_coeff(y, SR, H, iy, fy);
      const int ix1 = min(ix + 1, SR - 1), iy1 = min(iy + 1, SR - 1);
      const float ax0 = __fsub_rn(1.f, fx), ay0 = __fsub_rn(1.f, fy);
      const float r0 = __fadd_rn(__fmul_rn(s[iy * SR + ix], ax0), __fmul_rn(s[iy * SR + ix1], fx));
      const float r1 = __fadd_rn(__fmul_rn(s[iy1 * SR + ix], ax0), __fmul_rn(s[iy1 * SR + ix1], fx));
      sr = __fadd_rn(__fmul_rn(r0, ay0), __fmul_rn(r1, fy));
    }
    if (method != 0) {
      const size_t j = (size_t)b * H * W + i;
      int v = 0;
      if (fg_peak > 0) v = __double2int_rz(255.0 * (double)(float)(on_u8[j] + off_u8[j]) / (double)(float)fg_peak) & 0xFF;
      fg = __fmul_rn((float)v, 1.0f / 255.0f);
    }
    const float o = method == 0 ? sr : method == 1 ? fg : __fadd_rn(__fmul_rn(0.6f, sr), __fmul_rn(0.4f, fg));
    map[(size_t)b * H * W + i] = o;
    mx = fmaxf(mx, o);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(peak_bits + b, __float_as_int(mx));
}

// GAN_functions.py:98-99, :118-119: map / map.max() when the maximum is positive
__global__ void __launch_bounds__(256)
sal_map_scale_kernel(float* __restrict__ map, const int* __restrict__ peak_bits, int hw) {
  const int b = blockIdx.y;
  const float mx = __int_as_float(peak_bits[b]);
  if (!(mx > 0.f)) return;
  float* p = map + (size_t)b * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) p[i] = __fdiv_rn(p[i], mx);
}

inline size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }

struct SalWs {
  size_t gray, t1, t2, integ, on_sum, off_sum, on_u8, off_u8, small, stats, total;
  SalWs(int b, int h, int w) {
    const size_t hw = (size_t)b * h * w;
    size_t o = 0;
    auto take = [&](size_t n) { const size_t at = o; o += up256(n); return at; };
    gray = take(hw);
    t1 = take(hw);
    t2 = take(hw);
    integ = take((size_t)b * (h + 1) * (w + 1) * sizeof(float));
    on_sum = take(hw * 2);
    off_sum = take(hw * 2);
    on_u8 = take(hw);
    off_u8 = take(hw);
    small = take((size_t)b * SR2 * sizeof(float));
    stats = take((size_t)b * 5 * sizeof(int));
    total = o;
  }
};

DeviceOnce g_spectral_attr;

}  // namespace
}  // namespace cic

using namespace cic;

extern "C" size_t cic_saliency_map_workspace_bytes(int batch, int h, int w) {
  if (batch <= 0 || h <= 0 || w <= 0) return 0;
  return SalWs(batch, h, w).total;
}

extern "C" int cic_saliency_map_u8(const uint8_t* d_images, float* d_map, int batch, int h, int w, int rgb, int method,
                                   void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0, "cic_saliency_map_u8: bad shape");
  CIC_REQUIRE(method == CIC_SALIENCY_SPECTRAL_RESIDUAL || method == CIC_SALIENCY_FINE_GRAINED || method == CIC_SALIENCY_COMBINED,
              "cic_saliency_map_u8: unsupported saliency method %d", method);
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_images && d_map, "cic_saliency_map_u8: null pointer");
  CIC_REQUIRE(batch <= 65535, "cic_saliency_map_u8: at most 65535 images per call");
  CIC_REQUIRE((size_t)h * w <= (size_t)1 << 30, "cic_saliency_map_u8: image too large");
  const SalWs ws(batch, h, w);
  CIC_REQUIRE(d_workspace && workspace_bytes >= ws.total, "cic_saliency_map_u8: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)d_workspace;
  uint8_t* gray = (uint8_t*)(base + ws.gray);
  uint8_t* t1 = (uint8_t*)(base + ws.t1);
  uint8_t* t2 = (uint8_t*)(base + ws.t2);
  float* integ = (float*)(base + ws.integ);
  uint16_t* on_sum = (uint16_t*)(base + ws.on_sum);
  uint16_t* off_sum = (uint16_t*)(base + ws.off_sum);
  uint8_t* on_u8 = (uint8_t*)(base + ws.on_u8);
  uint8_t* off_u8 = (uint8_t*)(base + ws.off_u8);
  float* small = (float*)(base + ws.small);
  int* stats = (int*)(base + ws.stats);
  int* peak = stats + 4 * batch;
  const int hw = h * w;
  const size_t npix = (size_t)batch * hw;
  const int blocks = (hw + 255) / 256 < sm_count() * 4 ? (hw + 255) / 256 : sm_count() * 4;
  int launches = 0;

  CIC_CHECK_CUDA(cudaMemsetAsync(stats, 0, (size_t)batch * 5 * sizeof(int), st));
  const size_t gblocks = (npix + 255) / 256;
  sal_gray_kernel<<<(unsigned)(gblocks < (size_t)sm_count() * 8 ? gblocks : (size_t)sm_count() * 8), 256, 0, st>>>(d_images, gray, npix, rgb);
  ++launches;
  if (method != CIC_SALIENCY_FINE_GRAINED) {
    const size_t smem = (size_t)(4 * SR2 + 2 * SR) * sizeof(double);
    if (g_spectral_attr.todo()) {
      CIC_CHECK_CUDA(cudaFuncSetAttribute(sal_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      g_spectral_attr.done();
    }
    Gauss5 gk;                                            // cv::getGaussianKernel(5, 8.0)
    double sum = 0.0;
    for (int t = 0; t < 5; ++t) { gk.k[t] = std::exp(-(double)((t - 2) * (t - 2)) / (2.0 * 8.0 * 8.0)); sum += gk.k[t]; }
    for (int t = 0; t < 5; ++t) gk.k[t] /= sum;
    sal_spectral_kernel<<<batch, 1024, smem, st>>>(gray, small, h, w, gk);
    ++launches;
  }
  if (method != CIC_SALIENCY_SPECTRAL_RESIDUAL) {
    sal_gauss3_kernel<<<dim3(blocks, batch), 256, 0, st>>>(gray, t1, h, w);
    sal_gauss3_kernel<<<dim3(blocks, batch), 256, 0, st>>>(t1, t2, h, w);
    const int rows = batch * h;
    sal_rowscan_kernel<<<(rows + 7) / 8, 256, 0, st>>>(t2, integ, h, w, rows);
    sal_colscan_kernel<<<dim3((w + 1 + 127) / 128, batch), 128, 0, st>>>(integ, h, w);
    sal_fg_scales_kernel<<<dim3(blocks, batch), 256, 0, st>>>(t2, integ, on_sum, off_sum, stats, h, w);
    sal_fg_norm_kernel<<<dim3(blocks, batch), 256, 0, st>>>(on_sum, off_sum, on_u8, off_u8, stats, hw);
    launches += 6;
  }
  sal_combine_kernel<<<dim3(blocks, batch), 256, 0, st>>>(small, on_u8, off_u8, stats, d_map, peak, h, w, method);
  sal_map_scale_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_map, peak, hw);
  launches += 2;
  for (int i = 0; i < launches; ++i) CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("saliency map kernels");
  return CIC_OK;
}
