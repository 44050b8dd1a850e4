// Fused PSNR + MSE + SSIM reductions (GAN_functions.py:724-759, test_autoencoder.py:49-66).
//
// One pass over the image pair: each CTA owns a 32x32 pixel tile, stages the tile plus a 3-pixel
// halo of both images in shared memory (the 7x7 uniform window of scikit-image's
// structural_similarity), accumulates the squared error of its own pixels, runs the separable
// window sums and reduces the SSIM map of its pixels with warp shuffles; one atomicAdd(double) per
// CTA and quantity.  Algorithmic traffic: 24 B/pixel for an fp32 RGB pair, 6 B/pixel for uint8.
//
// Numerics follow scikit-image op by op: scipy.ndimage.uniform_filter filters axis 0 then axis 1,
// accumulating in double and storing each pass in the image dtype (float32 for float32 input,
// float64 for uint8 input); the SSIM expression is evaluated in that dtype with separate roundings
// (no FMA contraction); the cropped mean and the squared-error mean accumulate in double.  SSIM
// pixels within 3 px of the border are cropped by scikit-image, so the 'reflect' boundary rule of
// uniform_filter never reaches a pixel that counts and no padding is needed.
#include "common.cuh"

namespace cic {

constexpr int TS = 32;        // owned tile edge
constexpr int HALO = 3;       // (win_size - 1) / 2
constexpr int TP = TS + 2 * HALO;

// ------------------------------------------------------------------------------------------
// float32 path: d_a, d_b (B,H,W,C); v = (x + pre_add) * pre_mul
// acc[b*4 + 1] += sum of SSIM map (all channels), acc[b*4 + 3] += sum of squared error
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
metrics_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm, double* __restrict__ acc, int H, int W,
                   int C, float pre_add, float pre_mul, float c1, float c2, float cov_norm) {
  __shared__ float sa[TP][TP + 1];
  __shared__ float sb[TP][TP + 1];
  __shared__ float sv[5][TS][TP + 1];  // vertical-pass results: a, b, aa, bb, ab
  __shared__ double red[2][8];

  const int tiles_x = (W + TS - 1) / TS;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c = blockIdx.y, b = blockIdx.z;
  const int x0 = tx * TS, y0 = ty * TS;
  const size_t img_base = (size_t)b * H * W * C;

  // 1. stage tile + halo (zero outside the image; such values only reach cropped SSIM pixels)
  double sse = 0.0;
  for (int i = threadIdx.x; i < TP * TP; i += blockDim.x) {
    const int ly = i / TP, lx = i % TP;
    const int gy = y0 + ly - HALO, gx = x0 + lx - HALO;
    float va = 0.f, vb = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const size_t idx = img_base + ((size_t)gy * W + gx) * C + c;
      va = __fmul_rn(__fadd_rn(__ldg(A + idx), pre_add), pre_mul);
      vb = __fmul_rn(__fadd_rn(__ldg(Bm + idx), pre_add), pre_mul);
      if (ly >= HALO && ly < HALO + TS && lx >= HALO && lx < HALO + TS) {
        const float d = __fsub_rn(va, vb);
        sse += (double)__fmul_rn(d, d);
      }
    }
    sa[ly][lx] = va;
    sb[ly][lx] = vb;
  }
  __syncthreads();

  // 2. vertical pass (axis 0 first, like scipy): owned rows x (owned + halo) columns
  for (int i = threadIdx.x; i < TS * TP; i += blockDim.x) {
    const int r = i / TP, lx = i % TP;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float va = sa[r + k][lx], vb = sb[r + k][lx];
      s0 += (double)va;
      s1 += (double)vb;
      s2 += (double)__fmul_rn(va, va);
      s3 += (double)__fmul_rn(vb, vb);
      s4 += (double)__fmul_rn(va, vb);
    }
    sv[0][r][lx] = (float)(s0 / 7.0);
    sv[1][r][lx] = (float)(s1 / 7.0);
    sv[2][r][lx] = (float)(s2 / 7.0);
    sv[3][r][lx] = (float)(s3 / 7.0);
    sv[4][r][lx] = (float)(s4 / 7.0);
  }
  __syncthreads();

  // 3. horizontal pass + SSIM expression for owned pixels that survive the 3-px crop
  double ssum = 0.0;
  for (int i = threadIdx.x; i < TS * TS; i += blockDim.x) {
    const int r = i / TS, cx = i % TS;
    const int gy = y0 + r, gx = x0 + cx;
    if (gy < HALO || gy >= H - HALO || gx < HALO || gx >= W - HALO) continue;
    double s[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) t += (double)sv[q][r][cx + k];
      s[q] = t;
    }
    const float ux = (float)(s[0] / 7.0), uy = (float)(s[1] / 7.0);
    const float uxx = (float)(s[2] / 7.0), uyy = (float)(s[3] / 7.0), uxy = (float)(s[4] / 7.0);
    const float vx = __fmul_rn(cov_norm, __fsub_rn(uxx, __fmul_rn(ux, ux)));
    const float vy = __fmul_rn(cov_norm, __fsub_rn(uyy, __fmul_rn(uy, uy)));
    const float vxy = __fmul_rn(cov_norm, __fsub_rn(uxy, __fmul_rn(ux, uy)));
    const float a1 = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, ux), uy), c1);
    const float a2 = __fadd_rn(__fmul_rn(2.0f, vxy), c2);
    const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), c1);
    const float b2 = __fadd_rn(__fadd_rn(vx, vy), c2);
    const float S = __fdiv_rn(__fmul_rn(a1, a2), __fmul_rn(b1, b2));
    ssum += (double)S;
  }

  // 4. CTA reduction, one atomic per quantity
  sse = warp_sum(sse);
  ssum = warp_sum(ssum);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sse; red[1][warp] = ssum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    atomicAdd(acc + (size_t)b * 4 + 3, t0);
    atomicAdd(acc + (size_t)b * 4 + 1, t1);
  }
}

// Same arithmetic, all channels of a tile in one CTA (C <= 4): the pair is read once with fully coalesced rows
// (the per-channel kernel above reads every C-th float: C x the sectors), the window sums of one channel at a
// time reuse one staging buffer, and the /7 of the two filter passes is a multiply by the double 1/7 (one ulp of
// a double before the cast to float).  Dynamic shared memory: 2 * TP * (TP*C + 1) + 5 * TS * (TP + 1) floats.
__global__ void __launch_bounds__(256)
metrics_f32_packed_kernel(const float* __restrict__ A, const float* __restrict__ Bm, double* __restrict__ acc, int H, int W,
                          int C, float pre_add, float pre_mul, float c1, float c2, float cov_norm) {
  extern __shared__ float msm[];
  const int ld = TP * C + 1;
  float* sa = msm;                       // [TP][ld]
  float* sb = sa + TP * ld;              // [TP][ld]
  float* sv = sb + TP * ld;              // [5][TS][TP + 1]
  __shared__ double red[2][8];
  const int tiles_x = (W + TS - 1) / TS;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int b = blockIdx.y;
  const int x0 = tx * TS, y0 = ty * TS;
  const size_t img_base = (size_t)b * H * W * C;
  const int row_elems = TP * C;
  double sse = 0.0;
  {
    // a warp per staged row, lanes along the contiguous (pixel, channel) elements: no per-element division, and the
    // validity / ownership tests are range checks on the element index
    const int lo_rem = max(0, HALO - x0) * C, hi_rem = min(TP, W - x0 + HALO) * C;
    const int in_lo = HALO * C, in_hi = (HALO + TS) * C;
    const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int ly = wid; ly < TP; ly += nwarps) {
      const int gy = y0 + ly - HALO;
      const bool row_ok = gy >= 0 && gy < H, own_row = ly >= HALO && ly < HALO + TS;
      const long long base = (long long)img_base + ((long long)gy * W + (x0 - HALO)) * C;
      for (int rem = ln; rem < row_elems; rem += 32) {
        float va = 0.f, vb = 0.f;
        if (row_ok && rem >= lo_rem && rem < hi_rem) {
          va = __fmul_rn(__fadd_rn(__ldg(A + base + rem), pre_add), pre_mul);
          vb = __fmul_rn(__fadd_rn(__ldg(Bm + base + rem), pre_add), pre_mul);
          if (own_row && rem >= in_lo && rem < in_hi) {
            const float d = __fsub_rn(va, vb);
            sse += (double)__fmul_rn(d, d);
          }
        }
        sa[ly * ld + rem] = va;
        sb[ly * ld + rem] = vb;
      }
    }
  }
  __syncthreads();
  // Window sums slide: s += entering - leaving, all in double.  For image-range data the double sum of 7 floats is
  // exact, so this equals a fresh 7-term sum, with 10 instead of 35 float->double conversions per output (the
  // conversion pipe, 16/clk/SM, bounded the first version of this kernel).
  const double inv7 = 1.0 / 7.0;
  double ssum = 0.0;
  for (int c = 0; c < C; ++c) {
    // vertical pass (axis 0 first, like scipy): a thread per (column, 8-row segment) slides down its segment
    if (threadIdx.x < 4 * TP) {
      const int lx = threadIdx.x % TP, r0 = (threadIdx.x / TP) * 8;
      const float* pa = sa + r0 * ld + lx * C + c;
      const float* pb = sb + r0 * ld + lx * C + c;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float va = pa[k * ld], vb = pb[k * ld];
        s0 += (double)va; s1 += (double)vb;
        s2 += (double)__fmul_rn(va, va); s3 += (double)__fmul_rn(vb, vb); s4 += (double)__fmul_rn(va, vb);
      }
      float* o = sv + r0 * (TP + 1) + lx;
      for (int r = 0;; ++r) {
        o[0] = (float)(s0 * inv7);
        o[TS * (TP + 1)] = (float)(s1 * inv7);
        o[2 * TS * (TP + 1)] = (float)(s2 * inv7);
        o[3 * TS * (TP + 1)] = (float)(s3 * inv7);
        o[4 * TS * (TP + 1)] = (float)(s4 * inv7);
        if (r == 7) break;
        const float na = pa[(r + 7) * ld], nb = pb[(r + 7) * ld], oa = pa[r * ld], ob = pb[r * ld];
        s0 += (double)na - (double)oa;
        s1 += (double)nb - (double)ob;
        s2 += (double)__fmul_rn(na, na) - (double)__fmul_rn(oa, oa);
        s3 += (double)__fmul_rn(nb, nb) - (double)__fmul_rn(ob, ob);
        s4 += (double)__fmul_rn(na, nb) - (double)__fmul_rn(oa, ob);
        o += TP + 1;
      }
    }
    __syncthreads();
    // horizontal pass + SSIM expression: 8 threads per row, each slides over 4 consecutive owned columns
    {
      const int r = threadIdx.x >> 3, cx0 = (threadIdx.x & 7) * 4;
      const int gy = y0 + r;
      const float* src = sv + r * (TP + 1) + cx0;
      double s[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        double t = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) t += (double)src[q * TS * (TP + 1) + k];
        s[q] = t;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cx = cx0 + j, gx = x0 + cx;
        if (!(gy < HALO || gy >= H - HALO || gx < HALO || gx >= W - HALO)) {
          const float ux = (float)(s[0] * inv7), uy = (float)(s[1] * inv7), uxx = (float)(s[2] * inv7), uyy = (float)(s[3] * inv7),
                      uxy = (float)(s[4] * inv7);
          const float vx = __fmul_rn(cov_norm, __fsub_rn(uxx, __fmul_rn(ux, ux)));
          const float vy = __fmul_rn(cov_norm, __fsub_rn(uyy, __fmul_rn(uy, uy)));
          const float vxy = __fmul_rn(cov_norm, __fsub_rn(uxy, __fmul_rn(ux, uy)));
          const float a1 = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, ux), uy), c1);
          const float a2 = __fadd_rn(__fmul_rn(2.0f, vxy), c2);
          const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), c1);
          const float b2 = __fadd_rn(__fadd_rn(vx, vy), c2);
          ssum += (double)__fdiv_rn(__fmul_rn(a1, a2), __fmul_rn(b1, b2));
        }
        if (j < 3) {
#pragma unroll
          for (int q = 0; q < 5; ++q) s[q] += (double)src[q * TS * (TP + 1) + j + 7] - (double)src[q * TS * (TP + 1) + j];
        }
      }
    }
    __syncthreads();
  }
  sse = warp_sum(sse);
  ssum = warp_sum(ssum);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sse; red[1][warp] = ssum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    atomicAdd(acc + (size_t)b * 4 + 3, t0);
    atomicAdd(acc + (size_t)b * 4 + 1, t1);
  }
}

// Fast variant: the same reductions with the window sums in float32 on CENTRED data.  The packed kernel above reproduces scipy's
// double-accumulating uniform_filter and is bound by the fp32<->fp64 conversion pipe at 0.42 TB/s (profiles/r01_ncu_final.md); it
// needs the doubles only because skimage forms the variances as E[x^2] - E[x]^2 of values near 1, which cancels ~4 digits.
// Variance and covariance are shift invariant, so every tile subtracts a reference value (its centre pixel, per image and channel)
// before the products: in flat regions - the only place where the c2 = (0.03 R)^2 term is comparable to the variances - the
// centred values are ~0 and float32 sums of 7 (direct, no sliding: no drift) are accurate to ~1e-7 of the window's energy.
// The squared-error sum (PSNR, MSE) keeps its double accumulation.  Measured |dSSIM| against the scipy oracle: < 2e-6 on the
// test images (tolerance in tests/test_gpu_ops.py: 1e-5; the survey's criterion is 1e-4).
// Dynamic shared memory: 2 * TP * (TP*C + 1) + 5 * TS * (TP + 1) floats, as the packed kernel.
__global__ void __launch_bounds__(256, 3)
metrics_f32_fast_kernel(const float* __restrict__ A, const float* __restrict__ Bm, double* __restrict__ acc, int H, int W,
                        int C, float pre_add, float pre_mul, float c1, float c2, float cov_norm) {
  extern __shared__ float msm[];
  const int ld = TP * C + 1;
  float* sa = msm;                       // [TP][ld]
  float* sb = sa + TP * ld;              // [TP][ld]
  float* sv = sb + TP * ld;              // [5][TS][TP + 1]
  __shared__ double red[2][8];
  __shared__ float ref[2][4];
  const int tiles_x = (W + TS - 1) / TS;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int b = blockIdx.y;
  const int x0 = tx * TS, y0 = ty * TS;
  const size_t img_base = (size_t)b * H * W * C;
  const int row_elems = TP * C;
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  double sse = 0.0;
  {
    // Tile load: with one dependent load per loop iteration the kernel was bound by ~20 serial DRAM latencies per CTA (the fast
    // and the exact kernel both took 0.9 ms whatever their arithmetic); here every thread first issues all its loads of the
    // tile round (2 x 9 independent 4-byte loads), then normalises, accumulates the squared error and stores.
    constexpr int MAXE = 9;   // loads in flight per thread and image; TP * TP * C / 256 = 17 elements for C = 3: two rounds
    const int total = TP * row_elems;
    const int lo_rem = max(0, HALO - x0) * C, hi_rem = min(TP, W - x0 + HALO) * C;
    const int in_lo = HALO * C, in_hi = (HALO + TS) * C;
    const long long tile_base = (long long)img_base + ((long long)(y0 - HALO) * W + (x0 - HALO)) * C;
    // (ly, rem) of element i = threadIdx.x + 256 k advance by (q256, r256) per step: no division in the loops
    const int q256 = 256 / row_elems, r256 = 256 - q256 * row_elems;
    int ly0 = threadIdx.x / row_elems, rem0 = threadIdx.x - ly0 * row_elems;
    const long long row_pitch = (long long)W * C;
    for (int i0 = threadIdx.x; i0 < total; i0 += MAXE * 256) {
      float ra[MAXE], rb[MAXE];
      int ly = ly0, rem = rem0;
#pragma unroll
      for (int k = 0; k < MAXE; ++k) {
        const int gy = y0 + ly - HALO;
        const bool ok = ly < TP && gy >= 0 && gy < H && rem >= lo_rem && rem < hi_rem;
        const long long off = tile_base + ly * row_pitch + rem;
        ra[k] = ok ? __ldg(A + off) : -pre_add;   // (-pre_add + pre_add) * pre_mul = 0: outside the image the staged value is 0
        rb[k] = ok ? __ldg(Bm + off) : -pre_add;
        ly += q256; rem += r256;
        if (rem >= row_elems) { rem -= row_elems; ++ly; }
      }
      ly = ly0; rem = rem0;
#pragma unroll
      for (int k = 0; k < MAXE; ++k) {
        if (ly < TP) {
          const float va = __fmul_rn(__fadd_rn(ra[k], pre_add), pre_mul);
          const float vb = __fmul_rn(__fadd_rn(rb[k], pre_add), pre_mul);
          // owned pixels inside the image (hi_rem, H clip the right / bottom edge tiles)
          if (ly >= HALO && ly < HALO + TS && rem >= in_lo && rem < in_hi && rem < hi_rem && y0 + ly - HALO < H) {
            const float d = __fsub_rn(va, vb);
            sse += (double)__fmul_rn(d, d);
          }
          sa[ly * ld + rem] = va;
          sb[ly * ld + rem] = vb;
        }
        ly += q256; rem += r256;
        if (rem >= row_elems) { rem -= row_elems; ++ly; }
      }
      ly0 = ly; rem0 = rem;
    }
  }
  __syncthreads();
  // reference values: the first owned pixel of the tile (always inside the image)
  if (threadIdx.x < C) {
    ref[0][threadIdx.x] = sa[HALO * ld + HALO * C + threadIdx.x];
    ref[1][threadIdx.x] = sb[HALO * ld + HALO * C + threadIdx.x];
  }
  __syncthreads();
  constexpr int SVQ = TS * (TP + 1);
  const float inv49 = 1.0f / 49.0f;
  float ssum = 0.f;
  for (int c = 0; c < C; ++c) {
    const float ma = ref[0][c], mb = ref[1][c];
    // vertical pass: a thread per (column, 8-row segment) slides its five float32 window sums down the segment (14 rows read for
    // 8 outputs; on centred data seven sliding steps cost ~1e-7 of the window's energy, as a fresh 7-term sum does)
    if (threadIdx.x < 4 * TP) {
      const int lx = threadIdx.x % TP, r0 = (threadIdx.x / TP) * 8;
      const float* pa = sa + r0 * ld + lx * C + c;
      const float* pb = sb + r0 * ld + lx * C + c;
      float va[14], vb[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) { va[k] = pa[k * ld] - ma; vb[k] = pb[k * ld] - mb; }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        s0 += va[k]; s1 += vb[k];
        s2 = fmaf(va[k], va[k], s2); s3 = fmaf(vb[k], vb[k], s3); s4 = fmaf(va[k], vb[k], s4);
      }
      float* o = sv + r0 * (TP + 1) + lx;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        o[0] = s0; o[SVQ] = s1; o[2 * SVQ] = s2; o[3 * SVQ] = s3; o[4 * SVQ] = s4;
        if (r < 7) {
          s0 += va[r + 7] - va[r];
          s1 += vb[r + 7] - vb[r];
          s2 += va[r + 7] * va[r + 7] - va[r] * va[r];
          s3 += vb[r + 7] * vb[r + 7] - vb[r] * vb[r];
          s4 += va[r + 7] * vb[r + 7] - va[r] * vb[r];
          o += TP + 1;
        }
      }
    }
    __syncthreads();
    // horizontal pass + SSIM expression: 8 threads per row, each slides over 4 consecutive owned columns
    {
      const int r = threadIdx.x >> 3, cx0 = (threadIdx.x & 7) * 4;
      const int gy = y0 + r;
      const float* src = sv + r * (TP + 1) + cx0;
      float e[5][4];  // window sums of the five quantities for the thread's four outputs
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        float w[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) w[k] = src[t * SVQ + k];
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < 7; ++k) q += w[k];
        e[t][0] = q;
#pragma unroll
        for (int j = 1; j < 4; ++j) {
          q += w[j + 6] - w[j - 1];
          e[t][j] = q;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gx = x0 + cx0 + j;
        if (!(gy < HALO || gy >= H - HALO || gx < HALO || gx >= W - HALO)) {
          const float e0 = e[0][j] * inv49, e1 = e[1][j] * inv49, e2 = e[2][j] * inv49, e3 = e[3][j] * inv49, e4 = e[4][j] * inv49;
          const float vx = cov_norm * (e2 - e0 * e0);
          const float vy = cov_norm * (e3 - e1 * e1);
          const float vxy = cov_norm * (e4 - e0 * e1);
          const float ux = e0 + ma, uy = e1 + mb;
          const float a1 = 2.0f * ux * uy + c1, a2 = 2.0f * vxy + c2;
          const float b1 = ux * ux + uy * uy + c1, b2 = vx + vy + c2;
          ssum += __fdividef(a1 * a2, b1 * b2);  // 2 ulp: far below the 1e-7 of the window sums
        }
      }
    }
    __syncthreads();
  }
  sse = warp_sum(sse);
  double ssumd = warp_sum((double)ssum);
  if (ln == 0) { red[0][wid] = sse; red[1][wid] = ssumd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < nwarps; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    atomicAdd(acc + (size_t)b * 4 + 3, t0);
    atomicAdd(acc + (size_t)b * 4 + 1, t1);
  }
}

// ------------------------------------------------------------------------------------------
// float32 RGB path, warp-per-strip (the evaluation of the codec's batches; cic_metrics_psnr_ssim_f32_fast for C == 3, W % 4 == 0)
// ------------------------------------------------------------------------------------------
// The tile kernels above stage a 38 x 38 halo tile in shared memory and were bound by instruction issue (283 instructions per
// pixel and channel: 4-byte halo loads with index arithmetic, two shared-memory round trips, block barriers).  Here one warp
// walks down a strip of the image: a row of the interleaved NHWC image is a 1-D array of 3 W floats in which the 7 taps of a
// channel are 3 elements apart.  Every lane owns 12 consecutive elements (three aligned 128-bit loads per image and row = four
// RGB pixels, so element j has channel j % 3); lanes 1..30 produce output, lanes 0 and 31 only carry the 9-element halo.
//   vertical:   four running window sums per element (a, b, aa + bb, ab of centred values: SSIM needs vx + vy, never vx alone)
//               slide down the strip in registers;
//               the row that leaves the 7-row window comes back from a 7-slot ring in shared memory (raw centred a, b: 96 B
//               per thread and row) - no re-read from global memory, no block barrier;
//   horizontal: 9 + 9 halo values per quantity arrive by warp shuffle from the neighbouring lanes; the first three outputs of
//               a lane are full 7-tap sums, the other nine slide (+ in, - out): 36 adds per 12 outputs and quantity;
//   SSIM:       as metrics_f32_fast_kernel (float32 on centred data: variances are shift invariant; the strip's first pixel is
//               the reference), summed per row in float32, per strip in double; the squared error in double per element.
// One warp per CTA, so every branch on the strip geometry is uniform by construction and the shuffles need no re-convergence.
// Strips overlap by 6 rows (R owned rows + 3 above and below are read): redundant reads hit L2.
constexpr int MS_E = 12;             // elements (floats of the interleaved row) per thread and row: 3 x float4 = 4 RGB pixels
constexpr int MS_OUT = 30 * MS_E;    // output elements per warp and row: lanes 1..30 (lanes 0 and 31 carry the 9-element halo)
constexpr int MS_Q = 4;              // window sums per element: a, b, aa + bb, ab (SSIM needs vx + vy, never vx alone)
constexpr int MS_RING_BYTES = 7 * 6 * 32 * 16;  // 7 rows x (3 + 3) float4 per thread, one warp per CTA

__global__ void __launch_bounds__(32, 10)
metrics_f32_strip_kernel(const float* __restrict__ A, const float* __restrict__ Bm, double* __restrict__ acc, int batch, int H, int W, int R,
                         int bands, int segs, float pre_add, float pre_mul, float c1, float c2, float cov_norm) {
  extern __shared__ float4 ms_ring[];  // [slot 7][k 6][lane 32]
  const int lane = threadIdx.x;
  const unsigned wg = blockIdx.x;      // one warp per CTA: the strip geometry below is uniform by construction
  const int band = (int)(wg % (unsigned)bands), seg = (int)((wg / (unsigned)bands) % (unsigned)segs), img = (int)(wg / ((unsigned)bands * (unsigned)segs));
  (void)batch;
  const int y0 = seg * R, y1 = min(y0 + R, H);
  const int row_elems = 3 * W;
  const int e_base = band * MS_OUT - MS_E + MS_E * lane;            // first element of this thread (a multiple of 12: channel = j % 3)
  const bool in_row = e_base >= 0 && e_base < row_elems;            // row_elems % 12 == 0: a thread is inside or outside as a whole
  const bool owner = in_row && lane >= 1 && lane <= 30;
  const int e_load = min(max(e_base, 0), row_elems - MS_E);
  const size_t img_base = (size_t)img * H * row_elems;
  float4* ring = ms_ring + lane;
  // centring reference: the first owned pixel of the strip (variances are shift invariant; float32 window sums of centred
  // values are accurate to ~1e-7 of the window's energy where the c2 term matters)
  float ma[3], mb[3];
  {
    const int er = min(band * MS_OUT, row_elems - 3);
    const float* pa = A + img_base + (size_t)y0 * row_elems + er;
    const float* pb = Bm + img_base + (size_t)y0 * row_elems + er;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ma[c] = __fmul_rn(__fadd_rn(__ldg(pa + c), pre_add), pre_mul);
      mb[c] = __fmul_rn(__fadd_rn(__ldg(pb + c), pre_add), pre_mul);
    }
  }
  float s[MS_Q][MS_E];   // vertical window sums of a, b, aa + bb, ab (centred values)
#pragma unroll
  for (int q = 0; q < MS_Q; ++q)
#pragma unroll
    for (int j = 0; j < MS_E; ++j) s[q][j] = 0.f;
  double sse = 0.0, ssum = 0.0;
  const float inv49 = 1.0f / 49.0f;
  const int px0 = e_base / 3;   // pixel of element 0 (e_base is negative only for lane 0 of band 0, never an owner)
  int slot = 0;
  // software pipeline: the loads of row y + 1 are in flight while row y is processed (10 warps per SM cannot hide a DRAM round
  // trip per row on their own)
  float4 na[3], nb[3];
  {
    const int yr = min(max(y0 - 3, 0), H - 1);
    const float4* ra = reinterpret_cast<const float4*>(A + img_base + (size_t)yr * row_elems + e_load);
    const float4* rb = reinterpret_cast<const float4*>(Bm + img_base + (size_t)yr * row_elems + e_load);
#pragma unroll
    for (int k = 0; k < 3; ++k) { na[k] = __ldg(ra + k); nb[k] = __ldg(rb + k); }
  }
  for (int y = y0 - 3, i = 0; y < y1 + 3; ++y, ++i) {
    float a[MS_E], b[MS_E];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[4 * k] = na[k].x; a[4 * k + 1] = na[k].y; a[4 * k + 2] = na[k].z; a[4 * k + 3] = na[k].w;
      b[4 * k] = nb[k].x; b[4 * k + 1] = nb[k].y; b[4 * k + 2] = nb[k].z; b[4 * k + 3] = nb[k].w;
    }
    if (y + 1 < y1 + 3) {
      const int yr = min(max(y + 1, 0), H - 1);
      const float4* ra = reinterpret_cast<const float4*>(A + img_base + (size_t)yr * row_elems + e_load);
      const float4* rb = reinterpret_cast<const float4*>(Bm + img_base + (size_t)yr * row_elems + e_load);
#pragma unroll
      for (int k = 0; k < 3; ++k) { na[k] = __ldg(ra + k); nb[k] = __ldg(rb + k); }
    }
    double sse_row = 0.0;  // squared error: the same float32 products and double sums as the exact kernels (psnr / mse are identical)
#pragma unroll
    for (int j = 0; j < MS_E; ++j) {
      const float va = __fmul_rn(__fadd_rn(a[j], pre_add), pre_mul);
      const float vb = __fmul_rn(__fadd_rn(b[j], pre_add), pre_mul);
      const float d = __fsub_rn(va, vb);
      sse_row += (double)__fmul_rn(d, d);
      a[j] = va - ma[j % 3];
      b[j] = vb - mb[j % 3];
    }
    if (owner && y >= y0 && y < y1) sse += sse_row;
    // vertical sliding sums: the row that leaves the window (y - 7) comes back from the ring
    if (i >= 7) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 oa = ring[(slot * 6 + k) * 32], ob = ring[(slot * 6 + 3 + k) * 32];
        const float xa[4] = {oa.x, oa.y, oa.z, oa.w}, xb[4] = {ob.x, ob.y, ob.z, ob.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = 4 * k + t;
          s[0][j] -= xa[t];
          s[1][j] -= xb[t];
          s[2][j] = fmaf(-xa[t], xa[t], fmaf(-xb[t], xb[t], s[2][j]));
          s[3][j] = fmaf(-xa[t], xb[t], s[3][j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ring[(slot * 6 + k) * 32] = make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
      ring[(slot * 6 + 3 + k) * 32] = make_float4(b[4 * k], b[4 * k + 1], b[4 * k + 2], b[4 * k + 3]);
    }
    if (++slot == 7) slot = 0;
#pragma unroll
    for (int j = 0; j < MS_E; ++j) {
      s[0][j] += a[j];
      s[1][j] += b[j];
      s[2][j] = fmaf(a[j], a[j], fmaf(b[j], b[j], s[2][j]));
      s[3][j] = fmaf(a[j], b[j], s[3][j]);
    }
    const int yc = y - 3;   // centre row of the window that is complete now
    if (i < 6 || yc < max(y0, 3) || yc >= min(y1, H - 3)) continue;
    // horizontal 7-tap sums along the stride-3 (same channel) sequences: 9 halo elements from each neighbour lane; the first
    // three outputs are full sums, the other nine slide
    float e[MS_Q][MS_E];
#pragma unroll
    for (int q = 0; q < MS_Q; ++q) {
      float x[MS_E + 18];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        x[j] = __shfl_up_sync(0xffffffffu, s[q][j + 3], 1);
        x[MS_E + 9 + j] = __shfl_down_sync(0xffffffffu, s[q][j], 1);
      }
#pragma unroll
      for (int j = 0; j < MS_E; ++j) x[9 + j] = s[q][j];
#pragma unroll
      for (int j = 0; j < 3; ++j) e[q][j] = ((x[j] + x[j + 3]) + (x[j + 6] + x[j + 9])) + ((x[j + 12] + x[j + 15]) + x[j + 18]);
#pragma unroll
      for (int j = 3; j < MS_E; ++j) e[q][j] = e[q][j - 3] + (x[j + 18] - x[j - 3]);
    }
    if (owner) {
      float row_sum = 0.f;
#pragma unroll
      for (int j = 0; j < MS_E; ++j) {
        const int gx = px0 + j / 3;
        const float e0 = e[0][j] * inv49, e1 = e[1][j] * inv49, e2 = e[2][j] * inv49, e3 = e[3][j] * inv49;
        const float mm = fmaf(e0, e0, e1 * e1);           // ux'^2 + uy'^2 of the centred means
        const float vsum = cov_norm * (e2 - mm);          // vx + vy
        const float vxy = cov_norm * fmaf(-e0, e1, e3);
        const float ux = e0 + ma[j % 3], uy = e1 + mb[j % 3];
        const float a1 = fmaf(2.0f * ux, uy, c1), a2 = fmaf(2.0f, vxy, c2);
        const float b1 = fmaf(ux, ux, fmaf(uy, uy, c1)), b2 = vsum + c2;
        const float sv = __fdividef(a1 * a2, b1 * b2);
        row_sum += (gx >= 3 && gx < W - 3) ? sv : 0.f;
      }
      ssum += (double)row_sum;
    }
  }
  sse = warp_sum(sse);
  ssum = warp_sum(ssum);
  if (lane == 0) {
    atomicAdd(acc + (size_t)img * 4 + 3, sse);
    atomicAdd(acc + (size_t)img * 4 + 1, ssum);
  }
}

// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): one issue slot for two float operations - the strip kernel is issue-bound
typedef unsigned long long ms_u64;
__device__ __forceinline__ ms_u64 ms_pk(float2 v) { ms_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 ms_up(ms_u64 r) { float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r)); return v; }
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) { ms_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(ms_pk(a)), "l"(ms_pk(b))); return ms_up(r); }
__device__ __forceinline__ float2 f2_sub(float2 a, float2 b) { ms_u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(ms_pk(a)), "l"(ms_pk(b))); return ms_up(r); }
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) { ms_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(ms_pk(a)), "l"(ms_pk(b))); return ms_up(r); }
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) { ms_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(ms_pk(a)), "l"(ms_pk(b)), "l"(ms_pk(c))); return ms_up(r); }
__device__ __forceinline__ float2 f2_shfl_up(float2 v) { return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1)); }
__device__ __forceinline__ float2 f2_shfl_down(float2 v) { return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1)); }

// As metrics_f32_strip_kernel, with the window sums kept as packed pairs per element: (u, v) = (a + b, a - b) of the centred values
// and (uu, vv).  E[a] = (E[u] + E[v]) / 2, E[aa + bb] = (E[uu] + E[vv]) / 2, E[ab] = (E[uu] - E[vv]) / 4: every vertical and
// horizontal update is one packed instruction for two quantities, and the SSIM expression works on (wu, wv) = (E[uu] - E[u]^2,
// E[vv] - E[v]^2) and (U, V) = (ux + uy, ux - uy).
__global__ void __launch_bounds__(32, 10)
metrics_f32_strip2_kernel(const float* __restrict__ A, const float* __restrict__ Bm, double* __restrict__ acc, int batch, int H, int W, int R,
                          int bands, int segs, float pre_add, float pre_mul, float c1, float c2, float cov_norm) {
  extern __shared__ float4 ms_ring[];  // [slot 7][k 6][lane 32]: the (u, v) pairs of 12 elements
  const int lane = threadIdx.x;
  const unsigned wg = blockIdx.x;
  const int band = (int)(wg % (unsigned)bands), seg = (int)((wg / (unsigned)bands) % (unsigned)segs), img = (int)(wg / ((unsigned)bands * (unsigned)segs));
  (void)batch;
  const int y0 = seg * R, y1 = min(y0 + R, H);
  const int row_elems = 3 * W;
  const int e_base = band * MS_OUT - MS_E + MS_E * lane;
  const bool in_row = e_base >= 0 && e_base < row_elems;
  const bool owner = in_row && lane >= 1 && lane <= 30;
  const int e_load = min(max(e_base, 0), row_elems - MS_E);
  const size_t img_base = (size_t)img * H * row_elems;
  float4* ring = ms_ring + lane;
  float2 muv[3];   // (ma + mb, ma - mb) per channel: the centring reference in the (u, v) basis
  {
    const int er = min(band * MS_OUT, row_elems - 3);
    const float* pa = A + img_base + (size_t)y0 * row_elems + er;
    const float* pb = Bm + img_base + (size_t)y0 * row_elems + er;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ma = __fmul_rn(__fadd_rn(__ldg(pa + c), pre_add), pre_mul);
      const float mb = __fmul_rn(__fadd_rn(__ldg(pb + c), pre_add), pre_mul);
      muv[c] = make_float2(ma + mb, ma - mb);
    }
  }
  float2 s_uv[MS_E], s_sq[MS_E];
#pragma unroll
  for (int j = 0; j < MS_E; ++j) { s_uv[j] = make_float2(0.f, 0.f); s_sq[j] = make_float2(0.f, 0.f); }
  double sse = 0.0, ssum = 0.0;
  const float2 inv49 = make_float2(1.0f / 49.0f, 1.0f / 49.0f);
  const float cnh = 0.5f * cov_norm;
  const int px0 = e_base / 3;
  int slot = 0;
  float4 na[3], nb[3];
  {
    const int yr = min(max(y0 - 3, 0), H - 1);
    const float4* ra = reinterpret_cast<const float4*>(A + img_base + (size_t)yr * row_elems + e_load);
    const float4* rb = reinterpret_cast<const float4*>(Bm + img_base + (size_t)yr * row_elems + e_load);
#pragma unroll
    for (int k = 0; k < 3; ++k) { na[k] = __ldg(ra + k); nb[k] = __ldg(rb + k); }
  }
  for (int y = y0 - 3, i = 0; y < y1 + 3; ++y, ++i) {
    float a[MS_E], b[MS_E];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[4 * k] = na[k].x; a[4 * k + 1] = na[k].y; a[4 * k + 2] = na[k].z; a[4 * k + 3] = na[k].w;
      b[4 * k] = nb[k].x; b[4 * k + 1] = nb[k].y; b[4 * k + 2] = nb[k].z; b[4 * k + 3] = nb[k].w;
    }
    if (y + 1 < y1 + 3) {
      const int yr = min(max(y + 1, 0), H - 1);
      const float4* ra = reinterpret_cast<const float4*>(A + img_base + (size_t)yr * row_elems + e_load);
      const float4* rb = reinterpret_cast<const float4*>(Bm + img_base + (size_t)yr * row_elems + e_load);
#pragma unroll
      for (int k = 0; k < 3; ++k) { na[k] = __ldg(ra + k); nb[k] = __ldg(rb + k); }
    }
    float2 uv[MS_E];
    double sse_row = 0.0;
    const float2 pa2 = make_float2(pre_add, pre_add), pm2 = make_float2(pre_mul, pre_mul);
#pragma unroll
    for (int j = 0; j < MS_E; j += 2) {
      // v = (x + pre_add) * pre_mul and d = va - vb with the roundings of the exact kernels (add.rn / mul.rn / sub.rn, two lanes each)
      const float2 va = f2_mul(f2_add(make_float2(a[j], a[j + 1]), pa2), pm2);
      const float2 vb = f2_mul(f2_add(make_float2(b[j], b[j + 1]), pa2), pm2);
      const float2 d = f2_sub(va, vb), sm = f2_add(va, vb), dd = f2_mul(d, d);
      sse_row += (double)dd.x;
      sse_row += (double)dd.y;
      uv[j] = make_float2(sm.x - muv[j % 3].x, d.x - muv[j % 3].y);
      uv[j + 1] = make_float2(sm.y - muv[(j + 1) % 3].x, d.y - muv[(j + 1) % 3].y);
    }
    if (owner && y >= y0 && y < y1) sse += sse_row;
    if (i >= 7) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float4 o = ring[(slot * 6 + k) * 32];
        const float2 o0 = make_float2(o.x, o.y), o1 = make_float2(o.z, o.w);
        s_uv[2 * k] = f2_sub(s_uv[2 * k], o0);
        s_uv[2 * k + 1] = f2_sub(s_uv[2 * k + 1], o1);
        s_sq[2 * k] = f2_sub(s_sq[2 * k], f2_mul(o0, o0));
        s_sq[2 * k + 1] = f2_sub(s_sq[2 * k + 1], f2_mul(o1, o1));
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) ring[(slot * 6 + k) * 32] = make_float4(uv[2 * k].x, uv[2 * k].y, uv[2 * k + 1].x, uv[2 * k + 1].y);
    if (++slot == 7) slot = 0;
#pragma unroll
    for (int j = 0; j < MS_E; ++j) {
      s_uv[j] = f2_add(s_uv[j], uv[j]);
      s_sq[j] = f2_fma(uv[j], uv[j], s_sq[j]);
    }
    const int yc = y - 3;
    if (i < 6 || yc < max(y0, 3) || yc >= min(y1, H - 3)) continue;
    float2 e_uv[MS_E], e_sq[MS_E];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float2 x[MS_E + 18];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        x[j] = f2_shfl_up(q == 0 ? s_uv[j + 3] : s_sq[j + 3]);
        x[MS_E + 9 + j] = f2_shfl_down(q == 0 ? s_uv[j] : s_sq[j]);
      }
#pragma unroll
      for (int j = 0; j < MS_E; ++j) x[9 + j] = q == 0 ? s_uv[j] : s_sq[j];
      float2 e[MS_E];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        e[j] = f2_add(f2_add(f2_add(x[j], x[j + 3]), f2_add(x[j + 6], x[j + 9])), f2_add(f2_add(x[j + 12], x[j + 15]), x[j + 18]));
#pragma unroll
      for (int j = 3; j < MS_E; ++j) e[j] = f2_add(e[j - 3], f2_sub(x[j + 18], x[j - 3]));
#pragma unroll
      for (int j = 0; j < MS_E; ++j) { if (q == 0) e_uv[j] = e[j]; else e_sq[j] = e[j]; }
    }
    if (owner) {
      float row_sum = 0.f;
#pragma unroll
      for (int j = 0; j < MS_E; ++j) {
        const int gx = px0 + j / 3;
        const float2 m = f2_mul(e_uv[j], inv49);                       // (E[u], E[v]) of the centred values
        const float2 w = f2_sub(f2_mul(e_sq[j], inv49), f2_mul(m, m)); // (E[uu] - E[u]^2, E[vv] - E[v]^2)
        const float2 UV = f2_add(m, muv[j % 3]);                       // (ux + uy, ux - uy)
        const float2 UV2 = f2_mul(UV, UV);
        const float a1 = fmaf(0.5f, UV2.x - UV2.y, c1), b1 = fmaf(0.5f, UV2.x + UV2.y, c1);   // 2 ux uy + c1, ux^2 + uy^2 + c1
        const float a2 = fmaf(cnh, w.x - w.y, c2), b2 = fmaf(cnh, w.x + w.y, c2);             // 2 vxy + c2, vx + vy + c2
        const float sv = __fdividef(a1 * a2, b1 * b2);
        row_sum += (gx >= 3 && gx < W - 3) ? sv : 0.f;
      }
      ssum += (double)row_sum;
    }
  }
  sse = warp_sum(sse);
  ssum = warp_sum(ssum);
  if (lane == 0) {
    atomicAdd(acc + (size_t)img * 4 + 3, sse);
    atomicAdd(acc + (size_t)img * 4 + 1, ssum);
  }
}

// out[b] = {psnr, ssim, mse, sse}
__global__ void metrics_finalize_kernel(double* __restrict__ acc, int batch, double n_elems, double n_ssim,
                                        double data_range) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double* o = acc + (size_t)b * 4;
  const double sse = o[3];
  const double mse = sse / n_elems;
  o[2] = mse;
  o[1] = n_ssim > 0 ? o[1] / n_ssim : 0.0;
  o[0] = 10.0 * log10(data_range * data_range / mse);  // +inf when mse == 0, like scikit-image
}

// ------------------------------------------------------------------------------------------
// uint8 BGR path (test_autoencoder.py:49-66).  acc[b*4+0] wrapped-mse sum, +1 ssim sum, +3 sse
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int bgr2gray_u8(int b, int g, int r) {
  return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15;  // OpenCV 4.x fixed point, 15-bit coefficients
}

constexpr int GTSY = 16;  // gray/uint8 kernel: 32 x 16 tile (double-precision window sums need 2x the smem)
constexpr int GTPY = GTSY + 2 * HALO;
__global__ void __launch_bounds__(256)
metrics_gray_u8_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ Bm, double* __restrict__ acc, int H,
                       int W, double c1, double c2, double cov_norm) {
  __shared__ float sa[GTPY][TP + 1];  // gray values 0..255 are exact in float
  __shared__ float sb[GTPY][TP + 1];
  __shared__ double sv[5][GTSY][TP + 1];
  __shared__ double red[3][8];

  const int tiles_x = (W + TS - 1) / TS;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int b = blockIdx.z;
  const int x0 = tx * TS, y0 = ty * GTSY;
  const size_t img_base = (size_t)b * H * W * 3;

  unsigned long long sse = 0, wrapped = 0;
  for (int i = threadIdx.x; i < GTPY * TP; i += blockDim.x) {
    const int ly = i / TP, lx = i % TP;
    const int gy = y0 + ly - HALO, gx = x0 + lx - HALO;
    float va = 0.f, vb = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const size_t idx = img_base + ((size_t)gy * W + gx) * 3;
      const int a0 = A[idx], a1 = A[idx + 1], a2 = A[idx + 2];
      const int b0 = Bm[idx], b1 = Bm[idx + 1], b2 = Bm[idx + 2];
      va = (float)bgr2gray_u8(a0, a1, a2);
      vb = (float)bgr2gray_u8(b0, b1, b2);
      if (ly >= HALO && ly < HALO + GTSY && lx >= HALO && lx < HALO + TS) {
        const int d0 = a0 - b0, d1 = a1 - b1, d2 = a2 - b2;
        sse += (unsigned long long)(d0 * d0 + d1 * d1 + d2 * d2);
        // numpy uint8 arithmetic: (a-b) wraps mod 256, then the square wraps mod 256
        const unsigned w0 = (unsigned)(d0 & 255), w1 = (unsigned)(d1 & 255), w2 = (unsigned)(d2 & 255);
        wrapped += ((w0 * w0) & 255u) + ((w1 * w1) & 255u) + ((w2 * w2) & 255u);
      }
    }
    sa[ly][lx] = va;
    sb[ly][lx] = vb;
  }
  __syncthreads();

  for (int i = threadIdx.x; i < GTSY * TP; i += blockDim.x) {
    const int r = i / TP, lx = i % TP;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const double va = sa[r + k][lx], vb = sb[r + k][lx];
      s0 += va; s1 += vb; s2 += va * va; s3 += vb * vb; s4 += va * vb;   // integers: exact
    }
    sv[0][r][lx] = s0 / 7.0; sv[1][r][lx] = s1 / 7.0; sv[2][r][lx] = s2 / 7.0;
    sv[3][r][lx] = s3 / 7.0; sv[4][r][lx] = s4 / 7.0;
  }
  __syncthreads();

  double ssum = 0.0;
  for (int i = threadIdx.x; i < GTSY * TS; i += blockDim.x) {
    const int r = i / TS, cx = i % TS;
    const int gy = y0 + r, gx = x0 + cx;
    if (gy < HALO || gy >= H - HALO || gx < HALO || gx >= W - HALO) continue;
    double s[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) t += sv[q][r][cx + k];
      s[q] = t / 7.0;
    }
    const double ux = s[0], uy = s[1], uxx = s[2], uyy = s[3], uxy = s[4];
    const double vx = __dmul_rn(cov_norm, __dsub_rn(uxx, __dmul_rn(ux, ux)));
    const double vy = __dmul_rn(cov_norm, __dsub_rn(uyy, __dmul_rn(uy, uy)));
    const double vxy = __dmul_rn(cov_norm, __dsub_rn(uxy, __dmul_rn(ux, uy)));
    const double a1 = __dadd_rn(__dmul_rn(__dmul_rn(2.0, ux), uy), c1);
    const double a2 = __dadd_rn(__dmul_rn(2.0, vxy), c2);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), c1);
    const double b2 = __dadd_rn(__dadd_rn(vx, vy), c2);
    ssum += __ddiv_rn(__dmul_rn(a1, a2), __dmul_rn(b1, b2));
  }

  double d_sse = warp_sum((double)sse), d_wr = warp_sum((double)wrapped);
  ssum = warp_sum(ssum);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = d_sse; red[1][warp] = ssum; red[2][warp] = d_wr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    atomicAdd(acc + (size_t)b * 4 + 3, t0);
    atomicAdd(acc + (size_t)b * 4 + 1, t1);
    atomicAdd(acc + (size_t)b * 4 + 0, t2);
  }
}

// out[b] = {psnr, ssim, true mse, wrapped uint8 "mse"}
__global__ void metrics_gray_finalize_kernel(double* __restrict__ acc, int batch, double n_elems, double n_ssim) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double* o = acc + (size_t)b * 4;
  const double wrapped = o[0], sse = o[3];
  const double mse = sse / n_elems;
  o[0] = 10.0 * log10(255.0 * 255.0 / mse);
  o[1] = n_ssim > 0 ? o[1] / n_ssim : 0.0;
  o[2] = mse;
  o[3] = wrapped / n_elems;
}

}  // namespace cic

using namespace cic;

static int metrics_f32_impl(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w, int channels, float pre_add,
                            float pre_mul, float data_range, void* stream, bool fast) {
  CIC_REQUIRE(batch == 0 || (d_a && d_b && d_out), "cic_metrics_psnr_ssim_f32: null pointer");
  CIC_REQUIRE(batch >= 0 && h >= 7 && w >= 7 && channels >= 1 && channels <= 65535,
              "cic_metrics_psnr_ssim_f32: needs h,w >= 7 (7x7 SSIM window), got %dx%dx%d", h, w, channels);
  if (batch == 0) return CIC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CIC_CHECK_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 4 * batch, st));
  const int tiles = ((w + TS - 1) / TS) * ((h + TS - 1) / TS);
  // scikit-image: K1 = 0.01, K2 = 0.03, C = (K*R)^2, cov_norm = 49/48; python-float scalars are
  // combined with float32 arrays in float32
  const float c1 = (float)((0.01 * (double)data_range) * (0.01 * (double)data_range));
  const float c2 = (float)((0.03 * (double)data_range) * (0.03 * (double)data_range));
  const float cov_norm = (float)(49.0 / 48.0);
  const bool packed = channels <= 4;
  const size_t smem = packed ? (size_t)(2 * TP * (TP * channels + 1) + 5 * TS * (TP + 1)) * sizeof(float) : 0;
  if (packed) {
    static DeviceOnce attr_set;  // function attributes are per device
    if (attr_set.todo()) {
      CIC_CHECK_CUDA(cudaFuncSetAttribute(metrics_f32_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      CIC_CHECK_CUDA(cudaFuncSetAttribute(metrics_f32_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      attr_set.done();
    }
  }
  if (fast && channels == 3 && w % 4 == 0 && w >= 8 && (((uintptr_t)d_a | (uintptr_t)d_b) & 15) == 0 && CIC_KNOB("CIC_METRICS_STRIP", 1)) {
    static DeviceOnce strip_attr;
    if (strip_attr.todo()) {
      CIC_CHECK_CUDA(cudaFuncSetAttribute(metrics_f32_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MS_RING_BYTES));
      CIC_CHECK_CUDA(cudaFuncSetAttribute(metrics_f32_strip2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MS_RING_BYTES));
      strip_attr.done();
    }
    const int bands = (3 * w + MS_OUT - 1) / MS_OUT;
    // rows per strip: as tall as possible (6 of R + 6 rows read are overlap) while the grid still fills the resident-warp slots
    const long long slots = (long long)sm_count() * 10;
    int R = CIC_KNOB("CIC_METRICS_ROWS", 0);
    if (R <= 0) {
      R = 128;
      while (R > 16 && (long long)batch * bands * ((h + R - 1) / R) < slots) R >>= 1;
    }
    const int segs = (h + R - 1) / R;
    const long long warps = (long long)batch * bands * segs;
    CIC_REQUIRE(warps < 2147483647LL, "metrics: too many strips");
    if (CIC_KNOB("CIC_METRICS_PACKED", 1))
      metrics_f32_strip2_kernel<<<(unsigned)warps, 32, MS_RING_BYTES, st>>>(d_a, d_b, d_out, batch, h, w, R, bands, segs, pre_add, pre_mul, c1, c2, cov_norm);
    else
      metrics_f32_strip_kernel<<<(unsigned)warps, 32, MS_RING_BYTES, st>>>(d_a, d_b, d_out, batch, h, w, R, bands, segs, pre_add, pre_mul, c1, c2, cov_norm);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("metrics_f32_strip_kernel");
    metrics_finalize_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_out, batch, (double)h * w * channels,
                                                                 (double)(h - 6) * (w - 6) * channels, (double)data_range);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("metrics_finalize_kernel");
    return CIC_OK;
  }
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    const float* pa = d_a + (size_t)b0 * h * w * channels;
    const float* pb = d_b + (size_t)b0 * h * w * channels;
    if (packed && fast) {
      dim3 grid(tiles, nb);
      metrics_f32_fast_kernel<<<grid, 256, smem, st>>>(pa, pb, d_out + (size_t)b0 * 4, h, w, channels, pre_add, pre_mul, c1, c2, cov_norm);
    } else if (packed) {
      dim3 grid(tiles, nb);
      metrics_f32_packed_kernel<<<grid, 256, smem, st>>>(pa, pb, d_out + (size_t)b0 * 4, h, w, channels, pre_add, pre_mul, c1, c2, cov_norm);
    } else {
      dim3 grid(tiles, channels, nb);
      metrics_f32_kernel<<<grid, 256, 0, st>>>(pa, pb, d_out + (size_t)b0 * 4, h, w, channels, pre_add, pre_mul, c1, c2, cov_norm);
    }
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("metrics_f32_kernel");
  }
  metrics_finalize_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_out, batch, (double)h * w * channels,
                                                               (double)(h - 6) * (w - 6) * channels, (double)data_range);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("metrics_finalize_kernel");
  return CIC_OK;
}

extern "C" int cic_metrics_psnr_ssim_f32(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w,
                                         int channels, float pre_add, float pre_mul, float data_range, void* stream) {
  return metrics_f32_impl(d_a, d_b, d_out, batch, h, w, channels, pre_add, pre_mul, data_range, stream, false);
}

extern "C" int cic_metrics_psnr_ssim_f32_fast(const float* d_a, const float* d_b, double* d_out, int batch, int h, int w,
                                              int channels, float pre_add, float pre_mul, float data_range, void* stream) {
  return metrics_f32_impl(d_a, d_b, d_out, batch, h, w, channels, pre_add, pre_mul, data_range, stream, true);
}

// Per-rank metric sums of one evaluated batch (the row the ranks all-reduce; bpp accounting of GAN_test.py:310-325):
// out[8] = {sum psnr, sum ssim, sum mse, sum actual_bpp, sum hq_ratio, 0, n, 0} in double, one block.
__global__ void __launch_bounds__(256)
metric_sums_kernel(const double* __restrict__ m, const double* __restrict__ dt_sum, int n, double img_px, double latent_hq,
                   double latent_lq, double tile_px, double* __restrict__ out) {
  __shared__ double red[5][8];
  double s[5] = {0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double hq = dt_sum[i] / img_px;
    s[0] += m[4 * i]; s[1] += m[4 * i + 1]; s[2] += m[4 * i + 2];
    s[3] += (hq * latent_hq + (1.0 - hq) * latent_lq) * 32.0 / tile_px;  // total_bits / pixels of a tile
    s[4] += hq;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    s[q] = warp_sum(s[q]);
    if (lane == 0) red[q][warp] = s[q];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    if (threadIdx.x < 5) for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    if (threadIdx.x == 6) v = (double)n;
    out[threadIdx.x] = v;
  }
}

extern "C" int cic_metric_sums(const double* d_metrics, const double* d_dt_sum, int n, int img_px, int latent_hq, int latent_lq,
                               int tile_px, double* d_out, void* stream) {
  CIC_REQUIRE(d_metrics && d_dt_sum && d_out && n > 0 && img_px > 0 && tile_px > 0, "cic_metric_sums: bad argument");
  metric_sums_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_metrics, d_dt_sum, n, (double)img_px, (double)latent_hq, (double)latent_lq,
                                                        (double)tile_px, d_out);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("metric_sums_kernel");
  return CIC_OK;
}

extern "C" int cic_metrics_psnr_ssim_gray_u8(const uint8_t* d_a, const uint8_t* d_b, double* d_out, int batch, int h,
                                             int w, void* stream) {
  CIC_REQUIRE(batch == 0 || (d_a && d_b && d_out), "cic_metrics_psnr_ssim_gray_u8: null pointer");
  CIC_REQUIRE(batch >= 0 && h >= 7 && w >= 7, "cic_metrics_psnr_ssim_gray_u8: needs h,w >= 7, got %dx%d", h, w);
  if (batch == 0) return CIC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CIC_CHECK_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 4 * batch, st));
  const int tiles = ((w + TS - 1) / TS) * ((h + GTSY - 1) / GTSY);
  const double c1 = (0.01 * 255.0) * (0.01 * 255.0), c2 = (0.03 * 255.0) * (0.03 * 255.0);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid(tiles, 1, nb);
    metrics_gray_u8_kernel<<<grid, 256, 0, st>>>(d_a + (size_t)b0 * h * w * 3, d_b + (size_t)b0 * h * w * 3,
                                                  d_out + (size_t)b0 * 4, h, w, c1, c2, 49.0 / 48.0);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("metrics_gray_u8_kernel");
  }
  metrics_gray_finalize_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_out, batch, (double)h * w * 3,
                                                                    (double)(h - 6) * (w - 6));
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("metrics_gray_finalize_kernel");
  return CIC_OK;
}
