// RD-optimizer tail (GAN_functions.py:518-541) as one kernel: concat clip(bpp/5) -> Dense128 LeakyReLU -> Dense3 -> three
// biased sigmoids (fp32 FMA, ascending k like the Dense layers it replaces).  As four launches (set t, two Dense, finalise)
// the tail was latency-bound at ~25 us per launch whatever the batch - and again for every chunk of the pipelined predict.
// (A fused latent-saliency MLP was tried and dropped: one CTA per 8 rows streams the 2 MB Dense512 kernel from L2 and was
// L2-latency-bound at 0.21 ms against 0.09 ms for the three Dense launches.)
#include "common.cuh"
#include "plan.cuh"

namespace cic {

// feat: [B][ld] with the 64 pooled channels in columns 0..63; one block of 128 threads per tile
__global__ void __launch_bounds__(128)
rd_tail_kernel(const float* __restrict__ feat, int ld, const float* __restrict__ bpp, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out, int B) {
  __shared__ float f[65];
  __shared__ float d1[128];
  __shared__ float red[3][4];
  const int b = blockIdx.x, n = threadIdx.x;
  const float t = rate_t(bpp[b]);
  if (n < 64) f[n] = feat[(size_t)b * ld + n];
  if (n == 64) f[64] = t;  // :518 concat
  __syncthreads();
  float acc = 0.f;
#pragma unroll 5
  for (int k = 0; k < 65; ++k) acc = fmaf(f[k], __ldg(w1 + k * 128 + n), acc);
  acc = __fadd_rn(acc, __ldg(b1 + n));
  d1[n] = fmaxf(acc, __fmul_rn(acc, 0.2f));  // LeakyReLU(0.2) :521-522
  __syncthreads();
  const int warp = n >> 5, lane = n & 31;
  float p[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    p[j] = __fmul_rn(d1[n], __ldg(w2 + n * 3 + j));
    p[j] = warp_sum(p[j]);
    if (lane == 0) red[j][warp] = p[j];
  }
  __syncthreads();
  if (n < 3) {
    const float base = __fadd_rn(__fadd_rn(__fadd_rn(red[n][0], red[n][1]), __fadd_rn(red[n][2], red[n][3])), __ldg(b2 + n));
    const float k = n == 2 ? 1.5f : 2.0f;
    out[b * 3 + n] = sigmoidf_(__fsub_rn(__fadd_rn(base, 1.0f), __fmul_rn(k, t)));  // :529-541
  }
}

int launch_rd_tail_fused(const float* feat, int ld, const float* bpp, const float* w1, const float* b1, const float* w2, const float* b2,
                         float* rd_params, int B, cudaStream_t st) {
  if (B == 0) return CIC_OK;
  rd_tail_kernel<<<B, 128, 0, st>>>(feat, ld, bpp, w1, b1, w2, b2, rd_params, B);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("rd_tail_kernel");
  g_last_kernel_kind = KK_SIMT;
  return CIC_OK;
}

}  // namespace cic
