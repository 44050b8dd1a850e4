// create_saliency_mask(saliency_map, smooth=True) on the GPU (GAN_functions.py:199-203): the step that turns a
// saliency map into the model's mask input and that the reference recomputes on the CPU for every image and target bpp
// (GAN_test.py:279-280, :553).
//
//   mask = cv2.bilateralFilter(sal.astype(float32), 9, 75, 75)      -> bilateral_kernel
//   mask = cv2.GaussianBlur(mask, (31, 31), 0)                      -> gauss31_kernel (rows, then columns)
//   mask = mask / mask.max()  (if max > 0)                          -> max_kernel + scale_kernel
//
// OpenCV semantics reproduced (cv2 is the oracle of the tests - these two filters are in core OpenCV, not in contrib):
//   * bilateralFilter, CV_32F, d = 9: radius 4, the CIRCULAR neighbourhood r = sqrt(i^2 + j^2) <= 4 (49 taps), space weight
//     exp(-r^2 / (2 sigma_space^2)), colour weight exp(-dv^2 / (2 sigma_color^2)), BORDER_REFLECT_101; an image whose value
//     range is below FLT_EPSILON is copied.  OpenCV evaluates the colour weight through a 4096-bin interpolated table; with
//     sigma_color = 75 on values in [0, 1] the weight is within 1e-4 of 1 and the table error is ~1e-9, far below the 1e-5 the
//     tests ask for.
//   * GaussianBlur, ksize 31, sigma 0 -> sigma = 0.3 * ((31 - 1) * 0.5 - 1) + 0.8 = 5.0, float32 coefficients of the normalised
//     kernel, separable, BORDER_REFLECT_101.
// create_saliency_mask(smooth=False) - the binary mask at a given threshold or at the adaptive one of :172-194 (OpenCV's Otsu on the
// uint8 map, the 70 % share of a 50-bin histogram, clamped to [0.05, 0.5]) - is at the end of this file.
// The saliency MAP itself (GAN_functions.py:52-121) is saliency_map.cu.
#include "common.cuh"

#include <cfloat>
#include <cmath>

namespace cic {

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// per-image min and max -> mm[2 * b], mm[2 * b + 1] (ordered-int atomics on the float bit patterns; inputs are finite)
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}

__global__ void sal_minmax_init_kernel(float* mm, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) { mm[2 * i] = FLT_MAX; mm[2 * i + 1] = -FLT_MAX; }
}

__global__ void __launch_bounds__(256)
sal_minmax_kernel(const float* __restrict__ x, float* __restrict__ mm, int hw) {
  const int b = blockIdx.y;
  const float* p = x + (size_t)b * hw;
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const float v = __ldg(p + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f(mm + 2 * b, lo);
    atomic_max_f(mm + 2 * b + 1, hi);
  }
}

constexpr int BL_T = 32, BL_R = 4, BL_P = BL_T + 2 * BL_R;

__global__ void __launch_bounds__(256)
sal_bilateral_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mm, int H, int W, float space_coeff,
                     float color_coeff) {
  __shared__ float t[BL_P][BL_P + 1];
  const int b = blockIdx.z, x0 = blockIdx.x * BL_T, y0 = blockIdx.y * BL_T;
  const float* src = x + (size_t)b * H * W;
  for (int i = threadIdx.x; i < BL_P * BL_P; i += blockDim.x) {
    const int ly = i / BL_P, lx = i % BL_P;
    t[ly][lx] = __ldg(src + (size_t)reflect101(y0 + ly - BL_R, H) * W + reflect101(x0 + lx - BL_R, W));
  }
  __syncthreads();
  const bool flat = (mm[2 * b + 1] - mm[2 * b]) < FLT_EPSILON;
  for (int i = threadIdx.x; i < BL_T * BL_T; i += blockDim.x) {
    const int oy = i / BL_T, ox = i % BL_T;
    if (y0 + oy >= H || x0 + ox >= W) continue;
    const float v0 = t[oy + BL_R][ox + BL_R];
    float sum = 0.f, wsum = 0.f;
#pragma unroll
    for (int dy = -BL_R; dy <= BL_R; ++dy)
#pragma unroll
      for (int dx = -BL_R; dx <= BL_R; ++dx) {
        if (dy * dy + dx * dx > BL_R * BL_R) continue;      // OpenCV: r = sqrt(i*i + j*j) > radius -> skipped
        const float v = t[oy + BL_R + dy][ox + BL_R + dx];
        const float d = v - v0;
        const float w = expf((float)(dy * dy + dx * dx) * space_coeff) * expf(d * d * color_coeff);
        sum = fmaf(v, w, sum);
        wsum += w;
      }
    y[(size_t)b * H * W + (size_t)(y0 + oy) * W + x0 + ox] = flat ? v0 : sum / wsum;
  }
}

constexpr int GS_K = 31, GS_R = 15;

// one separable pass: horizontal (axis = 0) or vertical (axis = 1); coefficients in constant-size shared memory
__global__ void __launch_bounds__(256)
sal_gauss31_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int axis) {
  __shared__ float g[GS_K];
  if (threadIdx.x < GS_K) {
    // cv::getGaussianKernel(31, sigma = 5): exp(-x^2 / (2 sigma^2)) in double, normalised, stored as float32
    double s = 0.0;
    for (int k = 0; k < GS_K; ++k) { const double d = k - GS_R; s += exp(-d * d / 50.0); }
    const double d = (int)threadIdx.x - GS_R;
    g[threadIdx.x] = (float)(exp(-d * d / 50.0) / s);
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* src = x + (size_t)b * H * W;
  float* dst = y + (size_t)b * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int py = i / W, px = i % W;
    float acc = 0.f;
    if (axis == 0) {
      const float* row = src + (size_t)py * W;
#pragma unroll
      for (int k = 0; k < GS_K; ++k) acc = fmaf(g[k], __ldg(row + reflect101(px + k - GS_R, W)), acc);
    } else {
#pragma unroll
      for (int k = 0; k < GS_K; ++k) acc = fmaf(g[k], __ldg(src + (size_t)reflect101(py + k - GS_R, H) * W + px), acc);
    }
    dst[i] = acc;
  }
}

__global__ void __launch_bounds__(256)
sal_scale_kernel(float* __restrict__ y, const float* __restrict__ mm, int hw) {
  const int b = blockIdx.y;
  const float mx = mm[2 * b + 1];
  if (!(mx > 0.f)) return;                               // GAN_functions.py:202: only when mask.max() > 0
  float* p = y + (size_t)b * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) p[i] = __fdiv_rn(p[i], mx);
}

// ---- enhance_saliency_map (GAN_functions.py:123-157): bilateral -> Gaussian 3 / 9 / 15 -> 0.5 / 0.3 / 0.2 mix -> ^0.8 -> clip ------
struct GaussTaps { float g[15]; int k; };

// one separable pass with up to 15 float32 taps (cv2.GaussianBlur on CV_32F: float32 coefficients, BORDER_REFLECT_101)
__global__ void __launch_bounds__(256)
sal_gauss_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int axis, GaussTaps t) {
  const int b = blockIdx.y, r = t.k >> 1;
  const float* src = x + (size_t)b * H * W;
  float* dst = y + (size_t)b * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int py = i / W, px = i % W;
    float acc = 0.f;
    for (int k = 0; k < t.k; ++k) {
      const float v = axis == 0 ? __ldg(src + (size_t)py * W + reflect101(px + k - r, W)) : __ldg(src + (size_t)reflect101(py + k - r, H) * W + px);
      acc = fmaf(t.g[k], v, acc);
    }
    dst[i] = acc;
  }
}

// acc (+)= w * g (float32, the reference's `enhanced_map += weights[i] * scale_map`); last: ^0.8 and clip to [0, 1]
__global__ void __launch_bounds__(256)
sal_enhance_mix_kernel(const float* __restrict__ g, float* __restrict__ acc, float w, int first, int last, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = __fadd_rn(first ? 0.f : acc[i], __fmul_rn(w, __ldg(g + i)));
    if (last) a = fminf(fmaxf(powf(a, 0.8f), 0.f), 1.f);
    acc[i] = a;
  }
}

// ---- create_saliency_mask(smooth=False) and the adaptive threshold (GAN_functions.py:172-194, :204-206) -----------------------
// hist[b]: 256 bins of the uint8 map ((uchar)(x * 255) when the map's maximum is <= 1, else (uchar)x) and 50 bins of
// np.histogram(x, 50, range=(0, 1)) (a value belongs to the bin whose float32 edges float32(i * 0.02) enclose it; 1.0 goes to the last bin)
__global__ void __launch_bounds__(256)
sal_hist_kernel(const float* __restrict__ x, const float* __restrict__ mm, unsigned* __restrict__ hist, int hw) {
  __shared__ unsigned h[256 + 50];
  for (int i = threadIdx.x; i < 306; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const int b = blockIdx.y;
  const bool unit = mm[2 * b + 1] <= 1.0f;
  const float* p = x + (size_t)b * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const float v = __ldg(p + i);
    const int u = (unit ? __float2int_rz(__fmul_rn(v, 255.f)) : __float2int_rz(v)) & 0xFF;
    atomicAdd(&h[u], 1u);
    if (v >= 0.f && v <= 1.f) {
      // float32 data -> numpy makes float32 bin edges (linspace in float64, cast) and compares in float32
      int k = (int)__fmul_rn(v, 50.f);
      if (k == 50) k = 49;
      if (v < (float)((double)k * 0.02)) --k;
      else if (k != 49 && v >= (float)((double)(k + 1) * 0.02)) ++k;
      atomicAdd(&h[256 + k], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 306; i += blockDim.x)
    if (h[i]) atomicAdd(hist + (size_t)b * 306 + i, h[i]);
}

// one thread per image: OpenCV's getThreshVal_Otsu_8u, the 70 % histogram share, the clamp to [0.05, 0.5]
__global__ void sal_threshold_kernel(const unsigned* __restrict__ hist, double* __restrict__ thr, int batch, int hw) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const unsigned* h = hist + (size_t)b * 306;
  const double scale = 1.0 / (double)hw;
  double mu = 0.0;
  for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)h[i]));
  mu = __dmul_rn(mu, scale);
  double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
  int max_val = 0;
  for (int i = 0; i < 256; ++i) {
    const double p_i = __dmul_rn((double)h[i], scale);
    mu1 = __dmul_rn(mu1, q1);
    q1 = __dadd_rn(q1, p_i);
    const double q2 = __dsub_rn(1.0, q1);
    if (fmin(q1, q2) < (double)FLT_EPSILON || fmax(q1, q2) > 1.0 - (double)FLT_EPSILON) continue;
    mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
    const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
    const double d = __dsub_rn(mu1, mu2);
    const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
    if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
  }
  const double otsu = (double)max_val / 255.0;
  unsigned long long total = 0, run = 0;
  for (int k = 0; k < 50; ++k) total += h[256 + k];
  int first = 0;                                        // np.argmax of an all-False array is 0
  for (int k = 0; k < 50; ++k) {
    run += h[256 + k];
    if (total && (double)run / (double)total > 0.7) { first = k; break; }
  }
  const double by_share = (double)(float)((double)first * 0.02);   // a float32 bin edge
  thr[b] = fmax(0.05, fmin(0.5, fmin(otsu, by_share)));
}

__global__ void __launch_bounds__(256)
sal_binary_kernel(const float* __restrict__ x, float* __restrict__ y, const double* __restrict__ thr, double fixed, int use_fixed, int hw) {
  const int b = blockIdx.y;
  const double t = use_fixed ? fixed : thr[b];
  const float* p = x + (size_t)b * hw;
  float* q = y + (size_t)b * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) q[i] = __ldg(p + i) > (float)t ? 1.f : 0.f;   // numpy compares the float32 map with the threshold cast to float32
}

}  // namespace cic

using namespace cic;

extern "C" size_t cic_saliency_enhance_workspace_bytes(int batch, int h, int w) {
  if (batch <= 0 || h <= 0 || w <= 0) return 0;
  return 3 * (((size_t)batch * h * w * sizeof(float) + 255) & ~(size_t)255) + (((size_t)batch * 2 * sizeof(float) + 255) & ~(size_t)255) + 256;
}

extern "C" int cic_saliency_enhance(const float* d_saliency, float* d_out, int batch, int h, int w, void* d_workspace, size_t workspace_bytes,
                                    void* stream) {
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0, "cic_saliency_enhance: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_saliency && d_out, "cic_saliency_enhance: null pointer");
  CIC_REQUIRE(batch <= 65535, "cic_saliency_enhance: at most 65535 maps per call");
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_saliency_enhance_workspace_bytes(batch, h, w), "cic_saliency_enhance: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t plane = ((size_t)batch * h * w * sizeof(float) + 255) & ~(size_t)255;
  float* filt = (float*)d_workspace;                       // bilateral output
  float* t0 = (float*)((char*)d_workspace + plane);
  float* t1 = (float*)((char*)d_workspace + 2 * plane);
  float* mm = (float*)((char*)d_workspace + 3 * plane);
  const int hw = h * w;
  const size_t n = (size_t)batch * hw;
  const int blocks = (hw + 255) / 256 < sm_count() * 4 ? (hw + 255) / 256 : sm_count() * 4;
  sal_minmax_init_kernel<<<(batch + 127) / 128, 128, 0, st>>>(mm, batch);
  sal_minmax_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_saliency, mm, hw);
  const float coeff = (float)(-0.5 / (75.0 * 75.0));       // cv2.bilateralFilter(map, 9, 75, 75)
  sal_bilateral_kernel<<<dim3((w + BL_T - 1) / BL_T, (h + BL_T - 1) / BL_T, batch), 256, 0, st>>>(d_saliency, filt, mm, h, w, coeff, coeff);
  int launches = 3;
  // cv2.getGaussianKernel(k, 0, CV_32F): fixed tables for k = 3 and 9 (OpenCV's bit-exact small kernels), the formula with
  // sigma = 0.3 ((k - 1) / 2 - 1) + 0.8 = 2.6 for k = 15
  const int ks[3] = {3, 9, 15};
  const float wts[3] = {0.5f, 0.3f, 0.2f};
  for (int s = 0; s < 3; ++s) {
    GaussTaps t;
    t.k = ks[s];
    if (ks[s] == 3) {
      const float g3[3] = {0.25f, 0.5f, 0.25f};
      for (int i = 0; i < 3; ++i) t.g[i] = g3[i];
    } else if (ks[s] == 9) {
      const int g9[9] = {4, 13, 30, 51, 60, 51, 30, 13, 4};
      for (int i = 0; i < 9; ++i) t.g[i] = (float)g9[i] / 256.f;
    } else {
      double e[15], sum = 0.0;
      for (int i = 0; i < 15; ++i) { const double d = i - 7; e[i] = std::exp(-d * d / (2.0 * 2.6 * 2.6)); sum += e[i]; }
      for (int i = 0; i < 15; ++i) t.g[i] = (float)(e[i] / sum);
    }
    sal_gauss_kernel<<<dim3(blocks, batch), 256, 0, st>>>(filt, t0, h, w, 0, t);
    sal_gauss_kernel<<<dim3(blocks, batch), 256, 0, st>>>(t0, t1, h, w, 1, t);
    const size_t mb = (n + 255) / 256;
    sal_enhance_mix_kernel<<<(unsigned)(mb < (size_t)sm_count() * 8 ? mb : (size_t)sm_count() * 8), 256, 0, st>>>(t1, d_out, wts[s], s == 0, s == 2, n);
    launches += 3;
  }
  for (int i = 0; i < launches; ++i) CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("saliency enhance kernels");
  return CIC_OK;
}

extern "C" size_t cic_saliency_mask_binary_workspace_bytes(int batch) {
  if (batch <= 0) return 0;
  return (((size_t)batch * 306 * sizeof(unsigned) + 255) & ~(size_t)255) + (((size_t)batch * 2 * sizeof(float) + 255) & ~(size_t)255) +
         (((size_t)batch * sizeof(double) + 255) & ~(size_t)255) + 256;
}

extern "C" int cic_saliency_mask_binary(const float* d_saliency, float* d_mask, double threshold, int adaptive, double* d_threshold_out,
                                        int batch, int h, int w, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0, "cic_saliency_mask_binary: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_saliency && d_mask, "cic_saliency_mask_binary: null pointer");
  CIC_REQUIRE(batch <= 65535, "cic_saliency_mask_binary: at most 65535 maps per call");
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = h * w;
  const int blocks = (hw + 255) / 256 < sm_count() * 4 ? (hw + 255) / 256 : sm_count() * 4;
  double* thr = nullptr;
  if (adaptive) {
    CIC_REQUIRE(d_workspace && workspace_bytes >= cic_saliency_mask_binary_workspace_bytes(batch), "cic_saliency_mask_binary: workspace too small");
    unsigned* hist = (unsigned*)d_workspace;
    const size_t hist_bytes = ((size_t)batch * 306 * sizeof(unsigned) + 255) & ~(size_t)255;
    float* mm = (float*)((char*)d_workspace + hist_bytes);
    thr = (double*)((char*)mm + (((size_t)batch * 2 * sizeof(float) + 255) & ~(size_t)255));
    CIC_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)batch * 306 * sizeof(unsigned), st));
    sal_minmax_init_kernel<<<(batch + 127) / 128, 128, 0, st>>>(mm, batch);
    sal_minmax_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_saliency, mm, hw);
    sal_hist_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_saliency, mm, hist, hw);
    sal_threshold_kernel<<<(batch + 63) / 64, 64, 0, st>>>(hist, thr, batch, hw);
    for (int i = 0; i < 4; ++i) CIC_COUNT_LAUNCH();
    if (d_threshold_out) CIC_CHECK_CUDA(cudaMemcpyAsync(d_threshold_out, thr, (size_t)batch * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  sal_binary_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_saliency, d_mask, thr, threshold, adaptive ? 0 : 1, hw);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("saliency binary mask kernels");
  return CIC_OK;
}


extern "C" size_t cic_saliency_mask_workspace_bytes(int batch, int h, int w) {
  if (batch <= 0 || h <= 0 || w <= 0) return 0;
  return (((size_t)batch * h * w * sizeof(float) + 255) & ~(size_t)255) + (((size_t)batch * 2 * sizeof(float) + 255) & ~(size_t)255) + 256;
}

extern "C" int cic_saliency_mask_smooth(const float* d_saliency, float* d_mask, int batch, int h, int w, void* d_workspace,
                                        size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0, "cic_saliency_mask_smooth: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_saliency && d_mask, "cic_saliency_mask_smooth: null pointer");
  CIC_REQUIRE(batch <= 65535, "cic_saliency_mask_smooth: at most 65535 maps per call");
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_saliency_mask_workspace_bytes(batch, h, w), "cic_saliency_mask_smooth: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* tmp = (float*)d_workspace;
  float* mm = (float*)((char*)d_workspace + (((size_t)batch * h * w * sizeof(float) + 255) & ~(size_t)255));
  const int hw = h * w;
  const int blocks = (hw + 255) / 256 < sm_count() * 4 ? (hw + 255) / 256 : sm_count() * 4;
  sal_minmax_init_kernel<<<(batch + 127) / 128, 128, 0, st>>>(mm, batch);
  sal_minmax_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_saliency, mm, hw);
  // cv2.bilateralFilter(src, 9, 75, 75): gauss coefficients -0.5 / sigma^2
  const float coeff = (float)(-0.5 / (75.0 * 75.0));
  sal_bilateral_kernel<<<dim3((w + BL_T - 1) / BL_T, (h + BL_T - 1) / BL_T, batch), 256, 0, st>>>(d_saliency, d_mask, mm, h, w, coeff, coeff);
  sal_gauss31_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_mask, tmp, h, w, 0);
  sal_gauss31_kernel<<<dim3(blocks, batch), 256, 0, st>>>(tmp, d_mask, h, w, 1);
  sal_minmax_init_kernel<<<(batch + 127) / 128, 128, 0, st>>>(mm, batch);
  sal_minmax_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_mask, mm, hw);
  sal_scale_kernel<<<dim3(blocks, batch), 256, 0, st>>>(d_mask, mm, hw);
  for (int i = 0; i < 8; ++i) CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("saliency mask kernels");
  return CIC_OK;
}
