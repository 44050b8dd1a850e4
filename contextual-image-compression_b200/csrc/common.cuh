// Shared helpers for libcic.so (sm_100a).  Host-side error plumbing + device-side math that
// must round exactly like the reference's op-by-op fp32 graph (no FMA contraction where the
// reference executes separate TF ops).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cic.h"

// Kernel-selection switches used while tuning (profiles/*.md record what each one measured).  The shipped library has none: the
// macro is its default value unless the library is built with -DCIC_TUNING_KNOBS (build.py: CIC_BUILD_KNOBS=1), which reads the
// environment once per call site.
#ifdef CIC_TUNING_KNOBS
#include <stdlib.h>
#define CIC_KNOB(name, dflt) (getenv(name) ? atoi(getenv(name)) : (dflt))
#else
#define CIC_KNOB(name, dflt) (dflt)
#endif

namespace cic {

void set_error(const char* fmt, ...);

// "once per device" guard for cudaFuncSetAttribute (the attribute is per device: a process that switches devices must opt every
// device in).  todo() / done() may race between threads; the worst case is a harmless repeated cudaFuncSetAttribute.
struct DeviceOnce {
  unsigned long long mask = 0;
  static unsigned long long bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
  }
  bool todo() const { return !(__atomic_load_n(&mask, __ATOMIC_ACQUIRE) & bit()); }
  void done() { __atomic_fetch_or(&mask, bit(), __ATOMIC_RELEASE); }
};

#define CIC_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      cic::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CIC_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define CIC_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      cic::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CIC_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define CIC_REQUIRE(cond, ...)                                                                 \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      cic::set_error(__VA_ARGS__);                                                             \
      return CIC_ERR_INVALID;                                                                  \
    }                                                                                          \
  } while (0)

// launch counter: every kernel launch of this library goes through CIC_COUNT_LAUNCH so that
// bench.py can report `gpu_launches` from the library's own bookkeeping.
extern thread_local long long g_launch_count;
#define CIC_COUNT_LAUNCH() (++cic::g_launch_count)
// kernel class of the last heavy launch (the per-layer profiler tags its records with it)
enum KernelKind { KK_NONE = 0, KK_TC_GEMM = 1, KK_TC_CONV = 2, KK_TC_CONV1 = 3, KK_DIRECT = 4, KK_SIMT = 5, KK_TC_ROWS = 6, KK_TC_GEMM2 = 7, KK_ATTN = 8 };
extern thread_local int g_last_kernel_kind;

int sm_count();

// TF 'same' padding: out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total/2.
inline int same_out(int in, int s) { return (in + s - 1) / s; }
inline int same_pad_before(int in, int k, int s) {
  int out = same_out(in, s);
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;
}

// ---- device math ---------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case CIC_ACT_RELU: return fmaxf(v, 0.f);
    case CIC_ACT_LRELU02: return v > 0.f ? v : __fmul_rn(v, 0.2f);
    case CIC_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case CIC_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// t = clip(bpp/5, 0, 1) (GAN_functions.py:631-633); separate roundings like the TF ops.
__device__ __forceinline__ float rate_t(float bpp) { return fminf(fmaxf(__fdiv_rn(bpp, 5.0f), 0.f), 1.f); }
__device__ __forceinline__ float rate_thr(float t) { return __fsub_rn(0.9f, __fmul_rn(0.85f, t)); }  // :642-644
__device__ __forceinline__ float rate_qs(float t) { return __fsub_rn(0.9f, __fmul_rn(0.8f, t)); }    // :647-649
// dt = sigmoid((mask^0.7 - thr) * 20) (GAN_functions.py:651-657)
__device__ __forceinline__ float dyn_threshold(float m, float thr) {
  float es = powf(m, 0.7f);
  return sigmoidf_(__fmul_rn(__fsub_rn(es, thr), 20.0f));
}

// 256-bit global store (sm_100: STG.E.ENL2.256): a lane writes one whole 32-byte sector, so strided per-lane
// records are not written as half-sector pieces.  p must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g,
                                             uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace cic
