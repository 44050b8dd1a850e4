// Bandwidth-bound kernels of the codec: latent quantiser, rate scalars, ROI mask -> dynamic
// threshold -> HQ/LQ blend, multi-level hq_ratio sweep, symbol-histogram entropy, uint8 cast.
// All are HBM-bound: 128-bit loads/stores, one pass, warp-shuffle + shared-memory reductions,
// grids sized as multiples of the SM count.
#include "common.cuh"

namespace cic {

// ---------------------------------------------------------------------------------------------
// AdaptiveQuantizationLayer.call (GAN_functions.py:435-446)
// 12 B per latent element algorithmic traffic (read fp32, write fp32 dequant + int32 symbol).
// ---------------------------------------------------------------------------------------------
__global__ void quantize_kernel(const float* __restrict__ latent, const float* __restrict__ sal,
                                const float* __restrict__ qs, float* __restrict__ deq, int32_t* __restrict__ sym,
                                float* __restrict__ pre_out, float* __restrict__ scale_out, int batch, int L) {
  const int row = blockIdx.y;
  // effective_quant = qs * (1 - sal); scale = exp(effective_quant * 3)     (:438-441)
  const float eff = __fmul_rn(qs[row], __fsub_rn(1.0f, sal[row]));
  const float scale = expf(__fmul_rn(eff, 3.0f));
  if (scale_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) scale_out[row] = scale;
  const size_t base = (size_t)row * L;
  const int nvec = L >> 2;
  const bool aligned = ((L & 3) == 0);
  if (aligned) {
    const float4* lat4 = reinterpret_cast<const float4*>(latent + base);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
      float4 v = __ldg(lat4 + i);
      float4 p = make_float4(__fmul_rn(v.x, scale), __fmul_rn(v.y, scale), __fmul_rn(v.z, scale), __fmul_rn(v.w, scale));
      float4 r = make_float4(rintf(p.x), rintf(p.y), rintf(p.z), rintf(p.w));  // tf.round: half-to-even (:444)
      if (deq) reinterpret_cast<float4*>(deq + base)[i] =
          make_float4(__fdiv_rn(r.x, scale), __fdiv_rn(r.y, scale), __fdiv_rn(r.z, scale), __fdiv_rn(r.w, scale));
      if (sym) reinterpret_cast<int4*>(sym + base)[i] = make_int4((int)r.x, (int)r.y, (int)r.z, (int)r.w);
      if (pre_out) reinterpret_cast<float4*>(pre_out + base)[i] = p;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
      float p = __fmul_rn(latent[base + i], scale);
      float r = rintf(p);
      if (deq) deq[base + i] = __fdiv_rn(r, scale);
      if (sym) sym[base + i] = (int)r;
      if (pre_out) pre_out[base + i] = p;
    }
  }
}

__global__ void rate_scalars_kernel(const float* __restrict__ bpp, float* __restrict__ t_out, float* __restrict__ thr,
                                    float* __restrict__ qs, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = rate_t(bpp[i]);
  if (t_out) t_out[i] = t;
  if (thr) thr[i] = rate_thr(t);
  if (qs) qs[i] = rate_qs(t);
}

// ---------------------------------------------------------------------------------------------
// ROI mask -> dt -> blend (GAN_functions.py:651-684).  44 B/pixel algorithmic traffic for C=3
// (read hq 12 + lq 12 + mask 4; write out 12 + dt 4).  One thread handles 4 pixels: one float4 of
// mask and three float4 of each RGB stream.
// ---------------------------------------------------------------------------------------------
template <bool BLEND>
__global__ void __launch_bounds__(256)
roi_blend_c3_kernel(const float* __restrict__ hq, const float* __restrict__ lq, const float* __restrict__ mask,
                    const float* __restrict__ bpp, float* __restrict__ out, float* __restrict__ dt_out,
                    double* __restrict__ dt_sum, int hw) {
  const int img = blockIdx.y;
  const float thr = rate_thr(rate_t(bpp[img]));
  const size_t pbase = (size_t)img * hw;
  const int nquad = hw >> 2;
  double local = 0.0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += gridDim.x * blockDim.x) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mask + pbase) + q);
    float d[4] = {dyn_threshold(m.x, thr), dyn_threshold(m.y, thr), dyn_threshold(m.z, thr), dyn_threshold(m.w, thr)};
    local += ((double)d[0] + (double)d[1]) + ((double)d[2] + (double)d[3]);
    if (dt_out) reinterpret_cast<float4*>(dt_out + pbase)[q] = make_float4(d[0], d[1], d[2], d[3]);
    if (BLEND) {
      const float4* h4 = reinterpret_cast<const float4*>(hq + pbase * 3) + (size_t)q * 3;
      const float4* l4 = reinterpret_cast<const float4*>(lq + pbase * 3) + (size_t)q * 3;
      float4* o4 = reinterpret_cast<float4*>(out + pbase * 3) + (size_t)q * 3;
      float hv[12], lv[12], ov[12];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float4 a = __ldg(h4 + j), b = __ldg(l4 + j);
        hv[4 * j] = a.x; hv[4 * j + 1] = a.y; hv[4 * j + 2] = a.z; hv[4 * j + 3] = a.w;
        lv[4 * j] = b.x; lv[4 * j + 1] = b.y; lv[4 * j + 2] = b.z; lv[4 * j + 3] = b.w;
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const float w = d[e / 3];
        // weighted_hq + weighted_lq with (1 - dt) formed first (:682-684); separate roundings
        ov[e] = __fadd_rn(__fmul_rn(hv[e], w), __fmul_rn(lv[e], __fsub_rn(1.0f, w)));
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) o4[j] = make_float4(ov[4 * j], ov[4 * j + 1], ov[4 * j + 2], ov[4 * j + 3]);
    }
  }
  if (dt_sum) {
    __shared__ double wsum[8];
    double s = warp_sum(local);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0.0;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(dt_sum + img, v);
    }
  }
}

// generic (any C, any hw) fallback for the stand-alone operator
__global__ void roi_blend_generic_kernel(const float* __restrict__ hq, const float* __restrict__ lq,
                                         const float* __restrict__ mask, const float* __restrict__ bpp,
                                         float* __restrict__ out, float* __restrict__ dt_out, double* __restrict__ dt_sum,
                                         int hw, int C) {
  const int img = blockIdx.y;
  const float thr = rate_thr(rate_t(bpp[img]));
  const size_t pbase = (size_t)img * hw;
  float local = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const float d = dyn_threshold(mask[pbase + p], thr);
    local += d;
    if (dt_out) dt_out[pbase + p] = d;
    if (hq && out)
      for (int c = 0; c < C; ++c) {
        size_t i = (pbase + p) * C + c;
        out[i] = __fadd_rn(__fmul_rn(hq[i], d), __fmul_rn(lq[i], __fsub_rn(1.0f, d)));
      }
  }
  if (dt_sum) {
    double s = warp_sum((double)local);
    if ((threadIdx.x & 31) == 0) atomicAdd(dt_sum + img, s);
  }
}

// ---------------------------------------------------------------------------------------------
// hq_ratio for all target-bpp levels in one pass over the mask: 4 B/pixel for every level.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLevels = 32;
__global__ void __launch_bounds__(256)
hq_ratio_sweep_kernel(const float* __restrict__ mask, const float* __restrict__ bpp_levels, int n_levels,
                      double* __restrict__ ratio_sum, int hw) {
  __shared__ float s_thr[kMaxLevels];
  __shared__ double s_part[8][kMaxLevels];
  const int img = blockIdx.y;
  if (threadIdx.x < n_levels) s_thr[threadIdx.x] = rate_thr(rate_t(bpp_levels[threadIdx.x]));
  __syncthreads();
  float acc[kMaxLevels];
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) acc[l] = 0.f;
  const size_t pbase = (size_t)img * hw;
  const int nquad = (hw & 3) == 0 ? (hw >> 2) : 0;  // 128-bit loads need every image base 16-byte aligned
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += gridDim.x * blockDim.x) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mask + pbase) + q);
    const float es[4] = {powf(m.x, 0.7f), powf(m.y, 0.7f), powf(m.z, 0.7f), powf(m.w, 0.7f)};
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      if (l < n_levels) {
        const float thr = s_thr[l];
        acc[l] += (sigmoidf_(__fmul_rn(__fsub_rn(es[0], thr), 20.0f)) + sigmoidf_(__fmul_rn(__fsub_rn(es[1], thr), 20.0f))) +
                  (sigmoidf_(__fmul_rn(__fsub_rn(es[2], thr), 20.0f)) + sigmoidf_(__fmul_rn(__fsub_rn(es[3], thr), 20.0f)));
      }
    }
  }
  // scalar path when hw is not a multiple of 4
  {
    for (int p = (nquad << 2) + blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
      const float es = powf(mask[pbase + p], 0.7f);
      for (int l = 0; l < n_levels; ++l) acc[l] += sigmoidf_(__fmul_rn(__fsub_rn(es, s_thr[l]), 20.0f));
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    if (l < n_levels) {
      double s = warp_sum((double)acc[l]);
      if (lane == 0) s_part[warp][l] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < n_levels) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_part[w][threadIdx.x];
    atomicAdd(ratio_sum + (size_t)img * n_levels + threadIdx.x, s);
  }
}

__global__ void scale_f64_kernel(double* __restrict__ v, double mul, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= mul;
}

// ---------------------------------------------------------------------------------------------
// zeroth-order entropy of integer symbols: one CTA per row, shared-memory histogram.
// ---------------------------------------------------------------------------------------------
constexpr int kBins = 2 * CIC_SYM_MAX + 1;
__global__ void __launch_bounds__(256)
symbol_entropy_kernel(const int32_t* __restrict__ sym, double* __restrict__ bits, int L) {
  __shared__ int hist[kBins];
  __shared__ double wsum[8];
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int32_t* s = sym + (size_t)row * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    int v = s[i];
    v = v < -CIC_SYM_MAX ? -CIC_SYM_MAX : (v > CIC_SYM_MAX ? CIC_SYM_MAX : v);
    atomicAdd(&hist[v + CIC_SYM_MAX], 1);
  }
  __syncthreads();
  double local = 0.0;
  const double invL = 1.0 / (double)L;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
    int c = hist[i];
    if (c > 0) local -= (double)c * log2((double)c * invL);
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w];
    bits[row] = t;
  }
}

// (x*mul).astype(uint8): C cast semantics = truncation toward zero (test_autoencoder.py:88)
__global__ void f32_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, size_t n, float mul) {
  size_t nvec = n >> 2;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uchar4 o;
    o.x = (uint8_t)(int)__fmul_rn(v.x, mul);
    o.y = (uint8_t)(int)__fmul_rn(v.y, mul);
    o.z = (uint8_t)(int)__fmul_rn(v.z, mul);
    o.w = (uint8_t)(int)__fmul_rn(v.w, mul);
    reinterpret_cast<uchar4*>(y)[i] = o;
  }
  if (blockIdx.x == 0)
    for (size_t i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) y[i] = (uint8_t)(int)__fmul_rn(x[i], mul);
}

// GAN pixel conventions (GAN_functions.py:24-50): load (u8 - 127.5) / 127.5, save ((x + 1) * 127.5).astype(uint8) (truncation)
__global__ void u8_to_f32_signed_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, size_t n) {
  const size_t nvec = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(x) + i);
    float4 o;
    o.x = __fdiv_rn(__fsub_rn((float)v.x, 127.5f), 127.5f);
    o.y = __fdiv_rn(__fsub_rn((float)v.y, 127.5f), 127.5f);
    o.z = __fdiv_rn(__fsub_rn((float)v.z, 127.5f), 127.5f);
    o.w = __fdiv_rn(__fsub_rn((float)v.w, 127.5f), 127.5f);
    reinterpret_cast<float4*>(y)[i] = o;
  }
  if (blockIdx.x == 0)
    for (size_t i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) y[i] = __fdiv_rn(__fsub_rn((float)x[i], 127.5f), 127.5f);
}

__global__ void f32_signed_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, size_t n) {
  const size_t nvec = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uchar4 o;
    o.x = (uint8_t)(int)__fmul_rn(__fadd_rn(v.x, 1.f), 127.5f);
    o.y = (uint8_t)(int)__fmul_rn(__fadd_rn(v.y, 1.f), 127.5f);
    o.z = (uint8_t)(int)__fmul_rn(__fadd_rn(v.z, 1.f), 127.5f);
    o.w = (uint8_t)(int)__fmul_rn(__fadd_rn(v.w, 1.f), 127.5f);
    reinterpret_cast<uchar4*>(y)[i] = o;
  }
  if (blockIdx.x == 0)
    for (size_t i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) y[i] = (uint8_t)(int)__fmul_rn(__fadd_rn(x[i], 1.f), 127.5f);
}

static inline int grid_for(size_t work_items, int threads, int per_sm = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = (size_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace cic

using namespace cic;

extern "C" int cic_quantize_latent(const float* d_latent, const float* d_sal, const float* d_qs, float* d_deq,
                                   int32_t* d_symbols, float* d_pre, float* d_scale, int batch, int latent_dim,
                                   void* stream) {
  CIC_REQUIRE(batch >= 0 && latent_dim > 0, "cic_quantize_latent: bad shape (%d,%d)", batch, latent_dim);
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_latent && d_sal && d_qs, "cic_quantize_latent: null input");
  int gx = (latent_dim / 4 + 255) / 256;
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, batch);
  quantize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_latent, d_sal, d_qs, d_deq, d_symbols, d_pre, d_scale, batch,
                                                          latent_dim);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("quantize_kernel");
  return CIC_OK;
}

extern "C" int cic_rate_scalars(const float* d_bpp, float* d_t, float* d_thr, float* d_qs, int n, void* stream) {
  CIC_REQUIRE(n >= 0, "cic_rate_scalars: bad args");
  if (n == 0) return CIC_OK;
  CIC_REQUIRE(d_bpp, "cic_rate_scalars: null input");
  rate_scalars_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_bpp, d_t, d_thr, d_qs, n);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("rate_scalars_kernel");
  return CIC_OK;
}

extern "C" int cic_roi_mask_blend(const float* d_hq, const float* d_lq, const float* d_mask, const float* d_bpp,
                                  float* d_out, float* d_dt, double* d_dt_sum, int batch, int hw, int channels,
                                  void* stream) {
  CIC_REQUIRE(batch >= 0 && hw > 0 && channels > 0, "cic_roi_mask_blend: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_mask && d_bpp, "cic_roi_mask_blend: null mask/bpp");
  CIC_REQUIRE((d_hq == nullptr) == (d_lq == nullptr), "cic_roi_mask_blend: hq and lq must both be given or both NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_dt_sum) CIC_CHECK_CUDA(cudaMemsetAsync(d_dt_sum, 0, sizeof(double) * batch, st));
  const bool blend = d_hq != nullptr && d_out != nullptr;
  if (channels == 3 && (hw & 3) == 0) {
    int gx = grid_for((size_t)hw / 4, 256, 8);
    int per_img = (sm_count() * 8 + batch - 1) / batch;
    if (gx > per_img) gx = per_img < 1 ? 1 : per_img;
    dim3 grid(gx, batch);
    if (blend)
      roi_blend_c3_kernel<true><<<grid, 256, 0, st>>>(d_hq, d_lq, d_mask, d_bpp, d_out, d_dt, d_dt_sum, hw);
    else
      roi_blend_c3_kernel<false><<<grid, 256, 0, st>>>(nullptr, nullptr, d_mask, d_bpp, nullptr, d_dt, d_dt_sum, hw);
  } else {
    dim3 grid(grid_for(hw, 256, 4), batch);
    roi_blend_generic_kernel<<<grid, 256, 0, st>>>(blend ? d_hq : nullptr, d_lq, d_mask, d_bpp, d_out, d_dt, d_dt_sum, hw,
                                                    channels);
  }
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("roi_blend_kernel");
  return CIC_OK;
}

extern "C" int cic_hq_ratio_sweep(const float* d_mask, const float* d_bpp_levels, int n_levels, double* d_ratio,
                                  int batch, int hw, void* stream) {
  CIC_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "cic_hq_ratio_sweep: n_levels must be in [1,%d]", kMaxLevels);
  CIC_REQUIRE(batch >= 0 && hw > 0, "cic_hq_ratio_sweep: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_mask && d_bpp_levels && d_ratio, "cic_hq_ratio_sweep: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CIC_CHECK_CUDA(cudaMemsetAsync(d_ratio, 0, sizeof(double) * batch * n_levels, st));
  int gx = grid_for((size_t)hw / 4 + 1, 256, 4);
  int per_img = (sm_count() * 4 + batch - 1) / batch;
  if (gx > per_img) gx = per_img < 1 ? 1 : per_img;
  dim3 grid(gx, batch);
  hq_ratio_sweep_kernel<<<grid, 256, 0, st>>>(d_mask, d_bpp_levels, n_levels, d_ratio, hw);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("hq_ratio_sweep_kernel");
  int n = batch * n_levels;
  scale_f64_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_ratio, 1.0 / (double)hw, n);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("scale_f64_kernel");
  return CIC_OK;
}

extern "C" int cic_symbol_entropy_bits(const int32_t* d_symbols, double* d_bits, int batch, int latent_dim,
                                       void* stream) {
  CIC_REQUIRE(batch >= 0 && latent_dim > 0, "cic_symbol_entropy_bits: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_symbols && d_bits, "cic_symbol_entropy_bits: null pointer");
  symbol_entropy_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(d_symbols, d_bits, latent_dim);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("symbol_entropy_kernel");
  return CIC_OK;
}

extern "C" int cic_f32_to_u8_trunc(const float* d_x, uint8_t* d_y, size_t n, float mul, void* stream) {
  if (n == 0) return CIC_OK;
  CIC_REQUIRE(d_x && d_y, "cic_f32_to_u8_trunc: null pointer");
  f32_to_u8_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n, mul);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("f32_to_u8_kernel");
  return CIC_OK;
}

extern "C" int cic_u8_to_f32_signed(const uint8_t* d_x, float* d_y, size_t n, void* stream) {
  if (n == 0) return CIC_OK;
  CIC_REQUIRE(d_x && d_y, "cic_u8_to_f32_signed: null pointer");
  CIC_REQUIRE(((uintptr_t)d_x & 3) == 0 && ((uintptr_t)d_y & 15) == 0, "cic_u8_to_f32_signed: unaligned buffer");
  u8_to_f32_signed_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("u8_to_f32_signed_kernel");
  return CIC_OK;
}

extern "C" int cic_f32_signed_to_u8(const float* d_x, uint8_t* d_y, size_t n, void* stream) {
  if (n == 0) return CIC_OK;
  CIC_REQUIRE(d_x && d_y, "cic_f32_signed_to_u8: null pointer");
  CIC_REQUIRE(((uintptr_t)d_x & 15) == 0 && ((uintptr_t)d_y & 3) == 0, "cic_f32_signed_to_u8: unaligned buffer");
  f32_signed_to_u8_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("f32_signed_to_u8_kernel");
  return CIC_OK;
}
