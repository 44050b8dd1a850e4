// Shared epilogue of the tcgen05 kernels: one accumulator chunk (CH fp32 columns of one output row, already in
// registers) -> alpha, bias, folded BatchNorm, activation, residual -> global memory in one of four forms.
#pragma once
#include "common.cuh"

namespace cic {

enum TcOutMode { TC_OUT_BF16 = 0, TC_OUT_F32 = 1, TC_OUT_PARTIAL = 2, TC_OUT_BF16_T = 3 };

struct TcEpi {
  const float* bias;
  const float* scale;
  const float* shift;
  float alpha;
  int act;
  int out_mode;
  void* out_hi;                 // bf16 (or fp32 for TC_OUT_F32 / TC_OUT_PARTIAL)
  void* out_lo;                 // optional bf16 low part
  const __nv_bfloat16* res_hi;  // optional residual (same addressing as the bf16 output)
  const __nv_bfloat16* res_lo;
  int N;                        // real output channels
  int out_ld, out_coff;
  int out_H, out_W;
  int out_ys, out_xs;
  int8_t out_y0[4], out_x0[4];  // per phase
  int up2;                      // 1: replicate every output pixel 2x2 (nearest up-sampling fused into the store)
  int Ho, Wo;                   // output positions per batch item (of one phase)
  long long m_total;            // rows of the split-K partial buffer
  int tm_tx, tm_ty, tm_IH, tm_IW;  // tm_tx != 0: batch item b is tile (ty, tx) of image b / (tm_tx * tm_ty) in (n, IH, IW, ld)
};

// Per-channel epilogue vectors (global memory).  Staging them in shared memory was measured neutral (r01): a uniform 128-bit
// load costs four l1tex data-pipe wavefronts whether it is LDG or LDS (profiles/r01_epilogue_data_pipe.md); what helps is
// needing fewer vectors - the plans fold BatchNorm's scale into the packed weights and bias / shift into one vector.
struct TcEpiVec {
  const float* bias;
  const float* scale;
  const float* shift;
};

struct TcRow {  // the output row this thread owns
  int b, oy, ox, phase, split;
};

// v: CH accumulator columns starting at output channel n0; nv = number of valid channels (> 0).
// Written for a small *executed* instruction footprint: every branch is on a kernel-uniform value and sits
// outside the element loops, per-channel vectors are fetched as float4, and only the taken variant runs
// (the first version interleaved null checks, a per-element activation switch and scalar loads, and the
// epilogue warps stalled on instruction fetch: profiles/r01_ncu_raster_issue_bound.md).
// 4 x 4 transpose of 16-byte pieces inside every quad of lanes: in: p[4 k + w] = word w of piece k of this lane's row; out:
// p[4 j + w] = word w of piece (lane & 3) of the row of quad lane j.  The four lanes of a quad then store 64 contiguous bytes of
// one row per instruction instead of 32-byte pieces of four rows: 8 lines per warp store instead of 32
// (profiles/r01_epilogue_data_pipe.md).
__device__ __forceinline__ void tc_quad_transpose16(uint32_t* p, int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    const int k0 = 2 * pr, k1 = 2 * pr + 1;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, b0 ? p[4 * k0 + w] : p[4 * k1 + w], 1);
      if (b0) p[4 * k0 + w] = recv; else p[4 * k1 + w] = recv;
    }
  }
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    const int k0 = pr, k1 = pr + 2;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, b1 ? p[4 * k0 + w] : p[4 * k1 + w], 2);
      if (b1) p[4 * k0 + w] = recv; else p[4 * k1 + w] = recv;
    }
  }
}

// Called by ALL lanes of the warp (nv is warp-uniform); `valid` = this lane's row exists.  tw = tile width in rows (0 if consecutive
// rows are not consecutive output positions): with tw % 4 == 0 the rows 4q .. 4q+3 of a warp are neighbours along x and the bf16
// stores go through the quad transpose.
template <int CH>
__device__ __forceinline__ void tc_epilogue_store(const TcEpi& e, const TcEpiVec& ev, const TcRow& r, const uint32_t (&v)[32], int n0, int nv,
                                                  bool valid, int tw) {
  // (hi, lo) outputs keep the direct stores: those layers (encoder conv2-4) are tensor-bound and the second transpose only
  // lengthened their epilogue (conv2 +3 %), measured r01
  const bool quads = CH == 32 && nv == CH && e.out_mode == TC_OUT_BF16 && !e.out_lo && !e.up2 && !e.tm_tx && !e.res_hi && tw > 0 &&
                     (tw & 3) == 0 && ((e.out_ld | e.out_coff) & 15) == 0;
  if (!quads && !valid) return;
  long long opix;
  if (e.tm_tx) {  // scatter the tile into the image layout
    const int tpi = e.tm_tx * e.tm_ty, img = r.b / tpi, t = r.b % tpi;
    const int gy = (t / e.tm_tx) * e.out_H + (r.oy * e.out_ys + e.out_y0[r.phase]), gx = (t % e.tm_tx) * e.out_W + (r.ox * e.out_xs + e.out_x0[r.phase]);
    if (gy >= e.tm_IH || gx >= e.tm_IW) return;  // ragged last tile row / column: cropped on store (never with the quad path: !e.tm_tx)
    opix = ((long long)img * e.tm_IH + gy) * e.tm_IW + gx;
  } else {
    opix = ((long long)r.b * e.out_H + (r.oy * e.out_ys + e.out_y0[r.phase])) * e.out_W + (r.ox * e.out_xs + e.out_x0[r.phase]);
  }
  const bool full = nv == CH;
  if (e.out_mode == TC_OUT_PARTIAL) {
    const long long mrow = ((long long)r.b * e.Ho + r.oy) * e.Wo + r.ox;
    float* dst = reinterpret_cast<float*>(e.out_hi) + ((long long)r.split * e.m_total + mrow) * e.N + n0;
    if (full && (e.N & 7) == 0) {
#pragma unroll
      for (int j = 0; j < CH / 8; ++j)
        st_global_v8(dst + 8 * j, v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3], v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]);
    } else if (full && (e.N & 3) == 0) {
#pragma unroll
      for (int j = 0; j < CH / 4; ++j)
        reinterpret_cast<uint4*>(dst)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (j < nv) dst[j] = __uint_as_float(v[j]);
    }
    return;
  }
  float f[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[j]);
  if (e.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < CH; ++j) f[j] = __fmul_rn(e.alpha, f[j]);
  }
  // per-channel affine: (x + bias) * scale + shift  (conv bias, folded BatchNorm)
  if (full && (n0 & 3) == 0) {
    if (ev.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(ev.bias + n0);
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        f[4 * j] = __fadd_rn(f[4 * j], b.x); f[4 * j + 1] = __fadd_rn(f[4 * j + 1], b.y);
        f[4 * j + 2] = __fadd_rn(f[4 * j + 2], b.z); f[4 * j + 3] = __fadd_rn(f[4 * j + 3], b.w);
      }
    }
    if (ev.scale) {
      const float4* s4 = reinterpret_cast<const float4*>(ev.scale + n0);
      const float4* t4 = reinterpret_cast<const float4*>(ev.shift + n0);
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) {
        const float4 sc = __ldg(s4 + j), sh = __ldg(t4 + j);
        f[4 * j] = __fadd_rn(__fmul_rn(f[4 * j], sc.x), sh.x); f[4 * j + 1] = __fadd_rn(__fmul_rn(f[4 * j + 1], sc.y), sh.y);
        f[4 * j + 2] = __fadd_rn(__fmul_rn(f[4 * j + 2], sc.z), sh.z); f[4 * j + 3] = __fadd_rn(__fmul_rn(f[4 * j + 3], sc.w), sh.w);
      }
    }
  } else {
    if (ev.bias) {
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = __fadd_rn(f[j], j < nv ? __ldg(ev.bias + n0 + j) : 0.f);
    }
    if (ev.scale) {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        f[j] = __fadd_rn(__fmul_rn(f[j], j < nv ? __ldg(ev.scale + n0 + j) : 1.f), j < nv ? __ldg(ev.shift + n0 + j) : 0.f);
    }
  }
  switch (e.act) {
    case CIC_ACT_RELU:
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
      break;
    case CIC_ACT_LRELU02:  // x > 0 ? x : 0.2 x  ==  max(x, 0.2 x)
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], __fmul_rn(f[j], 0.2f));
      break;
    case CIC_ACT_SIGMOID:  // fully unrolled (register-resident f[]); only the valid channels are evaluated
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (j < nv) f[j] = 1.f / (1.f + expf(-f[j]));
      break;
    case CIC_ACT_TANH:
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (j < nv) f[j] = tanhf(f[j]);
      break;
    default: break;
  }
  if (e.out_mode == TC_OUT_F32) {
    float* dst = reinterpret_cast<float*>(e.out_hi) + opix * e.out_ld + e.out_coff + n0;
    if (full && ((e.out_ld | e.out_coff) & 7) == 0) {
#pragma unroll
      for (int j = 0; j < CH / 8; ++j)
        st_global_v8(dst + 8 * j, __float_as_uint(f[8 * j]), __float_as_uint(f[8 * j + 1]), __float_as_uint(f[8 * j + 2]), __float_as_uint(f[8 * j + 3]),
                     __float_as_uint(f[8 * j + 4]), __float_as_uint(f[8 * j + 5]), __float_as_uint(f[8 * j + 6]), __float_as_uint(f[8 * j + 7]));
    } else if (full && ((e.out_ld | e.out_coff) & 3) == 0) {
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (j < nv) dst[j] = f[j];
    }
  } else if (e.out_mode == TC_OUT_BF16) {  // host guarantees N % CH == 0 and 16-byte aligned records
    const long long idx = opix * e.out_ld + e.out_coff + n0;
    if (e.res_hi) {
      const uint4* rh = reinterpret_cast<const uint4*>(e.res_hi + idx);
#pragma unroll
      for (int j = 0; j < CH / 8; ++j) {
        const uint4 q = rh[j];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          f[8 * j + 2 * k] = __fadd_rn(f[8 * j + 2 * k], __uint_as_float(w[k] << 16));
          f[8 * j + 2 * k + 1] = __fadd_rn(f[8 * j + 2 * k + 1], __uint_as_float(w[k] & 0xFFFF0000u));
        }
      }
      if (e.res_lo) {
        const uint4* rl = reinterpret_cast<const uint4*>(e.res_lo + idx);
#pragma unroll
        for (int j = 0; j < CH / 8; ++j) {
          const uint4 q = rl[j];
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            f[8 * j + 2 * k] = __fadd_rn(f[8 * j + 2 * k], __uint_as_float(w[k] << 16));
            f[8 * j + 2 * k + 1] = __fadd_rn(f[8 * j + 2 * k + 1], __uint_as_float(w[k] & 0xFFFF0000u));
          }
        }
      }
    }
    uint32_t hi[CH / 2];
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    uint32_t lo[CH / 2];
    if (e.out_lo) {
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        const __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * j] - __uint_as_float(hi[j] << 16),
                                                       f[2 * j + 1] - __uint_as_float(hi[j] & 0xFFFF0000u));
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
      }
    }
    if constexpr (CH == 32) if (quads) {
      const int lane = threadIdx.x & 31, l4 = lane & 3, q0 = lane & ~3;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      tc_quad_transpose16(hi, lane);
      const long long rstride = (long long)e.out_xs * e.out_ld;  // elements between the records of neighbouring rows
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(e.out_hi) + idx;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!((vmask >> (q0 + j)) & 1u)) continue;
        const long long d = (long long)(j - l4) * rstride + l4 * 8;  // row of quad lane j, 16-byte piece l4
        *reinterpret_cast<uint4*>(ph + d) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
      }
      return;
    }
    const int reps = e.up2 ? 2 : 1;
    const bool wide_ok = ((e.out_ld | e.out_coff) & 15) == 0;
#pragma unroll 1
    for (int rr = 0; rr < reps * reps; ++rr) {
      long long o = idx;
      if (e.up2) o = ((((long long)r.b * e.out_H + (r.oy * 2 + (rr >> 1))) * e.out_W) + (r.ox * 2 + (rr & 1))) * e.out_ld + e.out_coff + n0;
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(e.out_hi) + o;
      __nv_bfloat16* pl = e.out_lo ? reinterpret_cast<__nv_bfloat16*>(e.out_lo) + o : nullptr;
      if (wide_ok) {  // 32-byte aligned records: whole-sector 256-bit stores
#pragma unroll
        for (int j = 0; j < CH / 16; ++j)
          st_global_v8(ph + 16 * j, hi[8 * j], hi[8 * j + 1], hi[8 * j + 2], hi[8 * j + 3], hi[8 * j + 4], hi[8 * j + 5], hi[8 * j + 6], hi[8 * j + 7]);
        if (pl) {
#pragma unroll
          for (int j = 0; j < CH / 16; ++j)
            st_global_v8(pl + 16 * j, lo[8 * j], lo[8 * j + 1], lo[8 * j + 2], lo[8 * j + 3], lo[8 * j + 4], lo[8 * j + 5], lo[8 * j + 6], lo[8 * j + 7]);
        }
      } else {
        uint4* dh = reinterpret_cast<uint4*>(ph);
#pragma unroll
        for (int j = 0; j < CH / 8; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        if (pl) {
          uint4* dl = reinterpret_cast<uint4*>(pl);
#pragma unroll
          for (int j = 0; j < CH / 8; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      }
    }
  } else {  // TC_OUT_BF16_T: out[b][n][position] (V^T for the attention PV product)
    const long long how = (long long)e.Ho * e.Wo;
    const long long pos = (long long)r.oy * e.Wo + r.ox;
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(e.out_hi) + ((long long)r.b * e.N + n0) * how + pos;
    __nv_bfloat16* ol = e.out_lo ? reinterpret_cast<__nv_bfloat16*>(e.out_lo) + ((long long)r.b * e.N + n0) * how + pos : nullptr;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (j < nv) {
        const __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
        oh[(long long)j * how] = h;
        if (ol) ol[(long long)j * how] = __float2bfloat16_rn(f[j] - __bfloat162float(h));
      }
    }
  }
}

}  // namespace cic
