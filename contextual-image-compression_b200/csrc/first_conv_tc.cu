// First layers with a 1- or 3-channel input and 32 outputs on the tensor cores (r02; the same scheme as conv1_tc.cu):
//
//   autoencoder conv1   Conv2D(32, k3, 'same') + ReLU + MaxPooling2D(2)   train_autoencoder.py:14-15   <3, 3, 1, POOL>
//   RD-optimizer conv1  Conv2D(32, k3, s2, 'same') + LeakyReLU(0.2)       GAN_functions.py:511-512     <1, 3, 2, SPLIT>
//
// K = taps x Cin (27, 9) is far too small for a TMA im2col view, so builder warps make the A operand: the fp32 input patch of a
// 128-pixel output tile is staged in shared memory (16-byte cp.async where the geometry allows it, the next tile's loads in flight
// while this one is split), every builder thread gathers the patch of ITS output pixel, splits it into bf16 (hi, lo) and stores
// it as one K-major row of the SWIZZLE_128B UMMA layout (K padded to 64 with zeros).  The weights ([32][64] bf16 hi / lo,
// pre-swizzled at plan creation) stay resident; one 3-term split MMA group (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM) makes
// 128 pixels x 32 channels.  Epilogue warps: tcgen05.ld -> bias -> activation -> bf16 (hi [, lo]) -> whole-sector stores.
// With POOL the lanes 4k .. 4k+3 of the tile are the four pixels of pooling window k, so the 2x2 maximum is two shuffles on the
// packed bf16 pairs (rounding is monotonic: max of the rounded values = rounded max) and every lane of the quad stores a quarter
// of the pooled pixel.
// These layers ran on the CUDA cores before (direct_conv.cu: FMA-issue-bound, 0.25 ms for 32 x 256x256 against 0.03 ms of HBM
// time); the direct kernels stay for shapes this one does not take and as the fp32 path's first layers.
#include "plan.cuh"
#include "tc_gemm.cuh"

namespace cic {

constexpr int FC_N = 32;                       // output channels
constexpr int FC_ROWB = 128;                   // bytes per A / B row (K = 64 bf16)
constexpr int FC_ABYTES = TC_BM * FC_ROWB;     // one A part (16 KB)
constexpr int FC_MAX_PATCH = 2304;             // floats
constexpr int FC_CTAS = 3;                     // CTAs per SM: one A buffer (32 KB) + weights + patch = 51 KB of shared memory, <= 75 registers

struct FirstConvParams {
  const float* x;
  TileMap tm;
  int batch, H, W, Ho, Wo;
  int TW, TH, tw_log2, tiles_x, tiles_y, total_tiles;
  int pad_t, pad_l, act;
  int PH, PWF, pitch, fast;
  const uint8_t* wimg;   // [hi | lo][32][128 B], rows pre-swizzled
  const float* bias;     // [32]
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  __nv_bfloat16* pool_hi;
};

__device__ __forceinline__ void fc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fc_named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// tile-local output pixel of A row / TMEM lane r
template <bool POOL>
__device__ __forceinline__ void fc_pixel(int r, int tw_log2, int& yl, int& xl) {
  if (POOL) {
    const int k = r >> 2, hw_log2 = tw_log2 - 1;
    yl = 2 * (k >> hw_log2) + ((r >> 1) & 1);
    xl = 2 * (k & ((1 << hw_log2) - 1)) + (r & 1);
  } else {
    yl = r >> tw_log2;
    xl = r & ((1 << tw_log2) - 1);
  }
}

__device__ __forceinline__ uint32_t fc_bf16x2_max(uint32_t a, uint32_t b) {
  const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&m);
}

template <int CIN, int KS, int STRIDE, bool POOL, bool SPLIT>
__global__ void __launch_bounds__(288, FC_CTAS)
first_conv_tc_kernel(const __grid_constant__ FirstConvParams p) {
  constexpr int NK = KS * KS * CIN, KCH = (NK + 7) / 8, KSTEPS = (NK + 15) / 16, ROWF = KS * CIN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                   // [hi | lo][128 rows x 128 B] (one buffer: the 3-9 MMAs of a tile retire long before the next tile is split)
  uint8_t* b_img = smem + 2 * FC_ABYTES;                   // [hi | lo][32 rows x 128 B]
  float* patch = reinterpret_cast<float*>(b_img + 2 * FC_N * FC_ROWB);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(patch + FC_MAX_PATCH + 8);
  uint64_t* a_empty = a_full + 2;
  uint64_t* tmem_full_bar = a_empty + 2;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // resident weights; zero A buffers (rows of invalid pixels and the K padding are never written again)
  for (int i = threadIdx.x; i < 2 * FC_N * FC_ROWB / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(b_img)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
  for (int i = threadIdx.x; i < 2 * FC_ABYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(a_buf)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1);   // (only [0] is used)
        mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * FC_N);
    tmem_relinquish();
  }
  fc_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < 4) {
    // ===== builders: thread r owns row r of the A tile =====
    const int r = threadIdx.x;
    int yl, xl;
    fc_pixel<POOL>(r, p.tw_log2, yl, xl);
    const bool row_ok = yl < p.TH;
    const uint32_t patch_s = smem_u32(patch);
    int lead_cur = 0, lead_next = 0;  // position of the patch row's first float in the staged row (fast staging starts 16-byte aligned)

    auto stage = [&](int t) -> int {
      const int b = t / tiles_per_img, ti = t % tiles_per_img;
      const int oy0 = (ti / p.tiles_x) * p.TH, ox0 = (ti % p.tiles_x) * p.TW;
      const float* xb;
      size_t row_stride;
      int vh = p.H, vw = p.W;  // rows / columns of this item that exist in the image (ragged last tile row / column: fewer)
      if (p.tm.tiles_x) {
        const int tpi = p.tm.tiles_x * p.tm.tiles_y, img = b / tpi, tt = b % tpi;
        const int gy0 = (tt / p.tm.tiles_x) * p.H, gx0 = (tt % p.tm.tiles_x) * p.W;
        xb = p.x + (((size_t)img * p.tm.IH + (size_t)gy0) * p.tm.IW + (size_t)gx0) * CIN;
        row_stride = (size_t)p.tm.IW * CIN;
        vh = min(p.H, p.tm.IH - gy0);
        vw = min(p.W, p.tm.IW - gx0);
      } else {
        xb = p.x + (size_t)b * p.H * p.W * CIN;
        row_stride = (size_t)p.W * CIN;
      }
      const int iy0 = STRIDE * oy0 - p.pad_t, if0 = (STRIDE * ox0 - p.pad_l) * CIN;  // first patch row; first patch float of a row
      const bool ragged = vh < p.H || vw < p.W;
      if (p.fast && !ragged) {
        // whole 16-byte chunks from the aligned float at or below if0: a chunk lies entirely inside or entirely outside the item's
        // row (W * CIN is a multiple of 4), outside chunks are zero fill ('same' padding; tiles are coded independently)
        const int s0 = if0 & ~3, lead = if0 - s0;
        const int nch = (lead + p.PWF + 3) >> 2, wf = p.W * CIN;
        for (int i = r; i < p.PH * nch; i += 128) {
          const int pr = i / nch, c = i - pr * nch;
          const int iy = iy0 + pr, f0 = s0 + 4 * c;
          const bool ok = iy >= 0 && iy < p.H && f0 >= 0 && f0 < wf;
          const float* src = ok ? xb + (size_t)iy * row_stride + f0 : p.x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(patch_s + (uint32_t)(pr * p.pitch + 4 * c) * 4u), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        return lead;
      }
      // generic staging: 4-byte cp.async, zero fill outside the item, edge replication beyond the image of a ragged tile
      for (int i = r; i < p.PH * p.PWF; i += 128) {
        const int pr = i / p.PWF, cix = i - pr * p.PWF;
        const int fx = if0 + cix;                               // float index in the item's row (may be negative)
        const int px = fx >= 0 ? fx / CIN : -1, ch = fx - px * CIN;
        const int iy = iy0 + pr;
        const bool ok = iy >= 0 && iy < p.H && fx >= 0 && px < p.W;
        const float* src = p.x;
        if (ok) src = xb + (size_t)min(iy, vh - 1) * row_stride + (size_t)min(px, vw - 1) * CIN + ch;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(patch_s + (uint32_t)(pr * p.pitch + cix) * 4u), "l"(src), "r"(ok ? 4 : 0) : "memory");
      }
      return 0;
    };

    int lt = 0;
    if ((int)blockIdx.x < p.total_tiles) lead_next = stage(blockIdx.x);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      lead_cur = lead_next;
      asm volatile("cp.async.wait_all;" ::: "memory");
      fc_named_bar_sync(1, 128);  // the whole patch of tile t has landed
      float v[8 * KCH];
#pragma unroll
      for (int i = NK; i < 8 * KCH; ++i) v[i] = 0.f;
      if (row_ok) {
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const float* row = patch + (STRIDE * yl + ky) * p.pitch + lead_cur + STRIDE * xl * CIN;
#pragma unroll
          for (int j = 0; j < ROWF; ++j) v[ky * ROWF + j] = row[j];
        }
      }
      fc_named_bar_sync(1, 128);  // everyone holds its patch values in registers: the staging buffer is free again
      if (t + (int)gridDim.x < p.total_tiles) lead_next = stage(t + gridDim.x);
      mbar_wait(&a_empty[0], ((uint32_t)lt & 1u) ^ 1u);  // the MMAs of the previous tile have retired
      if (row_ok) {
        uint8_t* row_hi = a_buf + (size_t)r * FC_ROWB;
        uint8_t* row_lo = row_hi + FC_ABYTES;
#pragma unroll
        for (int c = 0; c < KCH; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float f0 = v[8 * c + 2 * q], f1 = v[8 * c + 2 * q + 1];
            const __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
            hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(f0 - __uint_as_float(hi[q] << 16), f1 - __uint_as_float(hi[q] & 0xFFFF0000u));
            lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const int off = (c ^ (r & 7)) << 4;  // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
          *reinterpret_cast<uint4*>(row_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(row_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fc_fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[0]);
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_bf16(FC_N);
    const uint32_t a_lo0 = (smem_u32(a_buf) & 0x3FFFF) >> 4, b_lo0 = (smem_u32(b_img) & 0x3FFFF) >> 4;
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t ph = ((uint32_t)lt >> 1) & 1u;
      mbar_wait(&tmem_empty_bar[buf], ph ^ 1u);
      mbar_wait(&a_full[0], (uint32_t)lt & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(buf * FC_N);
        const uint32_t a_hi = a_lo0, a_lo = a_hi + (FC_ABYTES >> 4);
        const uint32_t b_hi = b_lo0, b_lo = b_lo0 + (uint32_t)(FC_N * FC_ROWB >> 4);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) umma_bf16(d, umma_desc_from_lo<64>(a_hi + 2 * k), umma_desc_from_lo<64>(b_hi + 2 * k), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) umma_bf16(d, umma_desc_from_lo<64>(a_lo + 2 * k), umma_desc_from_lo<64>(b_hi + 2 * k), idesc, 1u);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) umma_bf16(d, umma_desc_from_lo<64>(a_hi + 2 * k), umma_desc_from_lo<64>(b_lo + 2 * k), idesc, 1u);
        umma_commit(&a_empty[0]);
        umma_commit(&tmem_full_bar[buf]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 5..8, warp w owns TMEM lanes 32 * (w % 4) .. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int yl, xl;
    fc_pixel<POOL>(r, p.tw_log2, yl, xl);
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const int b = t / tiles_per_img, ti = t % tiles_per_img;
      const int oy = (ti / p.tiles_x) * p.TH + yl, ox = (ti % p.tiles_x) * p.TW + xl;
      const bool valid = yl < p.TH && oy < p.Ho && ox < p.Wo;
      const int buf = lt & 1;
      mbar_wait_relaxed(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * FC_N), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 bb = __ldg(reinterpret_cast<const float2*>(p.bias) + j);
        float f0 = __fadd_rn(__uint_as_float(v[2 * j]), bb.x), f1 = __fadd_rn(__uint_as_float(v[2 * j + 1]), bb.y);
        if (p.act == CIC_ACT_LRELU02) {
          f0 = fmaxf(f0, __fmul_rn(f0, 0.2f));
          f1 = fmaxf(f1, __fmul_rn(f1, 0.2f));
        } else if (p.act == CIC_ACT_RELU) {
          f0 = fmaxf(f0, 0.f);
          f1 = fmaxf(f1, 0.f);
        }
        const __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
        hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
        if (SPLIT) {
          const __nv_bfloat162 ll = __floats2bfloat162_rn(f0 - __uint_as_float(hi[j] << 16), f1 - __uint_as_float(hi[j] & 0xFFFF0000u));
          lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
      }
      if (valid) {
        const size_t o = (((size_t)b * p.Ho + oy) * p.Wo + ox) * FC_N;   // 64-byte aligned: whole-sector 256-bit stores
        st_global_v8(p.out_hi + o, hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]);
        st_global_v8(p.out_hi + o + 16, hi[8], hi[9], hi[10], hi[11], hi[12], hi[13], hi[14], hi[15]);
        if (SPLIT) {
          st_global_v8(p.out_lo + o, lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]);
          st_global_v8(p.out_lo + o + 16, lo[8], lo[9], lo[10], lo[11], lo[12], lo[13], lo[14], lo[15]);
        }
      }
      if (POOL) {
        // MaxPooling2D((2, 2)): lanes 4k .. 4k+3 hold the window (H, W even: a window is inside the image iff its first pixel is)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hi[j] = fc_bf16x2_max(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], 1));
          hi[j] = fc_bf16x2_max(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], 2));
        }
        if (valid) {
          const int l4 = lane & 3;
          const uint32_t w0 = l4 == 0 ? hi[0] : l4 == 1 ? hi[4] : l4 == 2 ? hi[8] : hi[12];
          const uint32_t w1 = l4 == 0 ? hi[1] : l4 == 1 ? hi[5] : l4 == 2 ? hi[9] : hi[13];
          const uint32_t w2 = l4 == 0 ? hi[2] : l4 == 1 ? hi[6] : l4 == 2 ? hi[10] : hi[14];
          const uint32_t w3 = l4 == 0 ? hi[3] : l4 == 1 ? hi[7] : l4 == 2 ? hi[11] : hi[15];
          const size_t o = (((size_t)b * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1)) * FC_N + 8 * l4;
          *reinterpret_cast<uint4*>(p.pool_hi + o) = make_uint4(w0, w1, w2, w3);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 2 * FC_N);
}

// (ks, ks, cin, 32) fp32 -> [hi | lo][32][64] bf16 rows in the swizzled shared-memory image; k = (ky * ks + kx) * cin + c
__global__ void first_conv_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ img, int nk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= FC_N * 64) return;
  const int n = i / 64, k = i % 64;
  const float v = k < nk ? w[k * FC_N + n] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  const size_t off = (size_t)n * FC_ROWB + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(img + (size_t)FC_N * FC_ROWB + off) = l;
}

size_t first_conv_image_bytes() { return (size_t)2 * FC_N * FC_ROWB; }

int first_conv_pack(const float* w, int nk, uint8_t* img, cudaStream_t st) {
  CIC_REQUIRE(w && img && nk > 0 && nk <= 64, "first_conv_pack: bad arguments");
  first_conv_pack_kernel<<<(FC_N * 64 + 255) / 256, 256, 0, st>>>(w, img, nk);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("first_conv_pack_kernel");
  return CIC_OK;
}

template <int CIN, int KS, int STRIDE, bool POOL, bool SPLIT>
static int launch_first_conv(FirstConvParams& p, const char* name, cudaStream_t st) {
  p.Ho = same_out(p.H, STRIDE);
  p.Wo = same_out(p.W, STRIDE);
  const int cap = POOL ? 64 : 128;
  int tw = POOL ? 2 : 1, lg = POOL ? 1 : 0;
  while (tw * 2 <= p.Wo && tw * 2 <= cap) { tw *= 2; ++lg; }
  p.TW = tw; p.tw_log2 = lg; p.TH = TC_BM / tw;
  if (p.TH > p.Ho) {
    int th = POOL ? 2 : 1;
    while (th * 2 <= p.Ho) th *= 2;
    p.TH = th;
  }
  p.tiles_x = (p.Wo + p.TW - 1) / p.TW; p.tiles_y = (p.Ho + p.TH - 1) / p.TH;
  const long long total = (long long)p.batch * p.tiles_x * p.tiles_y;
  CIC_REQUIRE(total < 2147483647LL, "%s: too many tiles", name);
  p.total_tiles = (int)total;
  p.pad_t = same_pad_before(p.H, KS, STRIDE); p.pad_l = same_pad_before(p.W, KS, STRIDE);
  p.PH = (p.TH - 1) * STRIDE + KS;
  p.PWF = ((p.TW - 1) * STRIDE + KS) * CIN;
  p.pitch = ((p.PWF + 3 + 3) >> 2) << 2;        // room for a lead of up to three floats, rows 16-byte aligned
  CIC_REQUIRE(p.PH * p.pitch <= FC_MAX_PATCH, "%s: patch too large (%d x %d floats)", name, p.PH, p.pitch);
  const size_t row_floats = (size_t)(p.tm.tiles_x ? p.tm.IW : p.W) * CIN;
  p.fast = ((p.W * CIN) & 3) == 0 && (row_floats & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 &&
           (p.tm.tiles_x || (((size_t)p.H * p.W * CIN) & 3) == 0);
  const size_t smem = 2 * FC_ABYTES + 2 * FC_N * FC_ROWB + (FC_MAX_PATCH + 8) * sizeof(float) + 128 + 1024;
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(first_conv_tc_kernel<CIN, KS, STRIDE, POOL, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.done();
  }
  const int slots = FC_CTAS * sm_count();
  first_conv_tc_kernel<CIN, KS, STRIDE, POOL, SPLIT><<<p.total_tiles < slots ? p.total_tiles : slots, 288, smem, st>>>(p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH(name);
  g_last_kernel_kind = KK_TC_CONV1;
  return CIC_OK;
}

// x (B,H,W,3) fp32 -> Conv2D(32, k3, 'same') + bias + ReLU -> bf16 (B,H,W,32) and its 2x2 max-pool (B,H/2,W/2,32); H, W even
int launch_first_conv_tc_pool(const float* x, const uint8_t* wimg, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* pool_hi,
                              int batch, int H, int W, int act, cudaStream_t st) {
  CIC_REQUIRE(H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "first_conv_tc (pool): H and W must be even");
  CIC_REQUIRE(act == CIC_ACT_RELU, "first_conv_tc (pool): the shuffled maximum needs a non-negative activation");
  if (batch == 0) return CIC_OK;
  FirstConvParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.batch = batch; p.H = H; p.W = W; p.act = act; p.wimg = wimg; p.bias = bias; p.out_hi = out_hi; p.pool_hi = pool_hi;
  return launch_first_conv<3, 3, 1, true, false>(p, "first_conv_tc_kernel<3,3,1,pool>", st);
}

// x (B,H,W,1) fp32 (or tiles of larger maps) -> Conv2D(32, k3, s2, 'same') + bias + act -> bf16 hi + lo, (B,H/2,W/2,32)
int launch_first_conv_tc_c1s2(const float* x, const uint8_t* wimg, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                              int batch, int H, int W, int act, const TileMap& tm, cudaStream_t st) {
  CIC_REQUIRE(H > 0 && W > 0, "first_conv_tc (c1 s2): bad image size");
  CIC_REQUIRE(act == CIC_ACT_NONE || act == CIC_ACT_RELU || act == CIC_ACT_LRELU02, "first_conv_tc (c1 s2): unsupported activation");
  if (batch == 0) return CIC_OK;
  FirstConvParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.tm = tm; p.batch = batch; p.H = H; p.W = W; p.act = act; p.wimg = wimg; p.bias = bias; p.out_hi = out_hi; p.out_lo = out_lo;
  return launch_first_conv<1, 3, 2, false, true>(p, "first_conv_tc_kernel<1,3,2,split>", st);
}

}  // namespace cic
