// Encoder conv1 on the tensor cores: Conv2D(64, k4, s2, 'same') + bias + LeakyReLU(0.2) of a 3-channel image
// (GAN_functions.py:300-302) for ONE or BOTH encoders of the adaptive model in a single pass.
//
// K = 4*4*3 = 48 is too small for a TMA im2col view (a pixel record is 12 bytes), so the A operand is built in
// shared memory: four builder warps stage the fp32 input patch of a 128-pixel output tile, and every builder
// thread gathers the 48 patch values of its output pixel, splits them into bf16 (hi, lo) and stores them as one
// 128-byte K-major row of the canonical SWIZZLE_128B UMMA layout (K padded to 64 with zeros, generic-proxy stores
// made visible to the tensor core with fence.proxy.async).  The weights of both encoders ([128][64] bf16 hi/lo,
// pre-swizzled at plan creation) stay resident in shared memory, so one 3-term split MMA group
// (hi*hi + lo*hi + hi*lo, 12 tcgen05.mma, fp32 accumulate in TMEM) produces 128 pixels x (64 + 64) channels and
// the image is read once for both encoders.  Builders, the MMA warp and the four epilogue warps (tcgen05.ld -> bias -> LeakyReLU ->
// bf16 hi/lo -> global) of a CTA work on consecutive tiles; r02: ONE A buffer and ONE accumulator per CTA (the 12 MMAs of a tile
// take ~0.3 us of the ~2 us a tile spends in the builders and the epilogue) and 2 x 64 pixel tiles (9 KB patch) keep a CTA at
// 74 KB of shared memory, 128 TMEM columns and 72 registers, so THREE CTAs share an SM - the kernel is bound by how many warps
// hide the epilogue's dependent chains (profiles/r02_ncu_conv1.md), and the co-resident CTAs overlap what the buffers did not.
// This replaces two runs of the CUDA-core direct kernel (direct_conv.cu), which is FFMA-issue-bound
// (3-register FFMA issues at half rate) at ~16 TFLOP/s.
#include "plan.cuh"
#include "tc_gemm.cuh"

namespace cic {

constexpr int C1_K = 64;             // padded K
constexpr int C1_ROWB = C1_K * 2;    // bytes per A / B row
constexpr int C1_ABYTES = TC_BM * C1_ROWB;  // one A part (16 KB)
constexpr int C1_MAX_PATCH = 2400;          // floats: TW = 64, TH = 2 staged as 6 rows of 98 16-byte chunks (2352)
constexpr int C1_CTAS = 3;                  // CTAs per SM: one A buffer + both encoders' weights + patch = 74 KB of shared memory

struct Conv1Params {
  const float* x;
  TileMap tm;
  int batch, H, W, Ho, Wo;
  int TW, TH, tiles_x, tiles_y, total_tiles;
  int pad_t, pad_l;
  const uint8_t* wimg;   // [2 parts][N][128 B], rows pre-swizzled
  const float* bias;     // [N]
  __nv_bfloat16* out_hi[2];
  __nv_bfloat16* out_lo[2];
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 4 x 4 transpose of 16-byte pieces inside every quad of lanes: in: p[4 k + w] = word w of piece k of this lane's row; out:
// p[4 j + w] = word w of piece (lane & 3) of the row of quad lane j.  After it the four lanes of a quad hold 64 contiguous
// bytes of one row per store instruction instead of one 32-byte piece of four different rows: a warp store then touches 8 lines
// instead of 32 (the l1tex data pipe was 69 % busy with this kernel's stores, profiles/r01_epilogue_data_pipe.md).
__device__ __forceinline__ void quad_transpose16(uint32_t (&p)[16], int lane) {
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

template <int N>
__global__ void __launch_bounds__(288, C1_CTAS)
conv1_tc_kernel(const __grid_constant__ Conv1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                   // [hi | lo][128 rows x 128 B] (one buffer: the 12 MMAs of a tile retire long before the next tile is split)
  uint8_t* b_img = smem + 2 * C1_ABYTES;                   // [hi | lo][N rows x 128 B]
  float* patch = reinterpret_cast<float*>(b_img + 2 * N * C1_ROWB);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(patch + C1_MAX_PATCH + 8);
  uint64_t* a_empty = a_full + 2;
  uint64_t* tmem_full_bar = a_empty + 2;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int PW3 = (2 * p.TW + 2) * 3, PH = 2 * p.TH + 2;

  // resident weights + zero K padding (chunks 6, 7 of every A row are never written again)
  for (int i = threadIdx.x; i < 2 * N * C1_ROWB / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(b_img)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
  for (int i = threadIdx.x; i < 2 * C1_ABYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(a_buf)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1);
        mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, N);   // one accumulator (three CTAs x 128 columns fit the SM's 512)
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < 4) {
    // ===== builders: thread r owns row r of the A tile =====
    const int r = threadIdx.x;
    const int xl = r % p.TW, yl = r / p.TW;
    const bool row_ok = yl < p.TH;
    const uint32_t patch_s = smem_u32(patch);
    // Fast staging (the codec's 256-wide tiles and every other 16-byte-friendly geometry): the patch row starts one pixel left of
    // the tile (pad_l = 1), i.e. 12 bytes before a 16-byte boundary, so staging starts ONE FLOAT earlier and moves whole 16-byte
    // chunks: the first chunk holds the left padding pixel and the last one the right padding pixel - both entirely outside the
    // tile (zero fill), every other chunk entirely inside.  6 cp.async per thread and tile instead of 24.
    const int PWC = (PW3 + 1 + 3) >> 2;                      // 16-byte chunks per staged row (one leading float)
    const bool fast = p.pad_l == 1 && (p.W & 3) == 0 && (p.TW & 1) == 0 && (!p.tm.tiles_x || (p.tm.IW & 3) == 0) &&
                      ((size_t)PH * PWC * 4 <= (size_t)C1_MAX_PATCH) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    // floats per staged row; position of the row's first patch float (image float -3 of a row staged from float -4: odd, so the
    // gather below cannot use 8-byte loads on it directly)
    const int pitch = fast ? 4 * PWC : PW3, lead = fast ? 1 : 0;

    // stage the patch of tile t (asynchronously; the caller waits with cp.async.wait_all + the named barrier)
    auto stage = [&](int t) {
      const int b = t / tiles_per_img, ti = t % tiles_per_img;
      const int oy0 = (ti / p.tiles_x) * p.TH, ox0 = (ti % p.tiles_x) * p.TW;
      const float* xb;
      size_t row_stride;
      int vh = p.H, vw = p.W;  // rows / columns of this item that exist in the image (ragged last tile row / column: fewer)
      if (p.tm.tiles_x) {
        const int tpi = p.tm.tiles_x * p.tm.tiles_y, img = b / tpi, tt = b % tpi;
        const int gy0 = (tt / p.tm.tiles_x) * p.H, gx0 = (tt % p.tm.tiles_x) * p.W;
        xb = p.x + (((size_t)img * p.tm.IH + (size_t)gy0) * p.tm.IW + (size_t)gx0) * 3;
        row_stride = (size_t)p.tm.IW * 3;
        vh = min(p.H, p.tm.IH - gy0);
        vw = min(p.W, p.tm.IW - gx0);
      } else {
        xb = p.x + (size_t)b * p.H * p.W * 3;
        row_stride = (size_t)p.W * 3;
      }
      const int iy0 = 2 * oy0 - p.pad_t, ix0 = 2 * ox0 - p.pad_l;
      const bool ragged = vh < p.H || vw < p.W;
      if (fast && !ragged) {
        // the row's first patch float is float 3 ix0 = 6 ox0 - 3 of the item's row (ox0 a multiple of the even TW): staging starts
        // at the aligned float below it; chunk c of row pr covers floats 3 ix0 - 1 + 4c .. + 3 and lies entirely inside or outside
        const int s0 = 3 * ix0 - 1;
        for (int i = r; i < PH * PWC; i += 128) {
          const int pr = i / PWC, c = i - pr * PWC;
          const int iy = iy0 + pr, f0 = s0 + 4 * c;
          const bool ok = iy >= 0 && iy < p.H && f0 >= 0 && f0 < 3 * p.W;
          const float* src = ok ? xb + (size_t)iy * row_stride + f0 : p.x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(patch_s + (uint32_t)i * 16u), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        return;
      }
      // generic staging: 4-byte cp.async with zero fill outside the tile ('same' padding: tiles are coded independently); pixels
      // of a ragged tile beyond the image replicate the image's last row / column
      const float* row0 = xb + (long long)ix0 * 3;
      int pr = 0, cix = r;  // r < 128 < PW3
      for (int i = r; i < PH * PW3; i += 128) {
        const int px = (int)(((unsigned)cix * 43691u) >> 17);  // cix / 3 for cix < 98304
        const int iy = iy0 + pr, ix = ix0 + px;
        const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        const float* src = ok ? row0 + (size_t)iy * row_stride + cix : p.x;
        if (ragged && ok && (iy >= vh || ix >= vw))
          src = xb + (size_t)min(iy, vh - 1) * row_stride + (size_t)min(ix, vw - 1) * 3 + (cix - 3 * px);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(patch_s + (uint32_t)(pr * pitch + lead + cix) * 4u), "l"(src), "r"(ok ? 4 : 0) : "memory");
        cix += 128;
        if (cix >= PW3) { cix -= PW3; ++pr; }
      }
    };

    int lt = 0;
    if ((int)blockIdx.x < p.total_tiles) stage(blockIdx.x);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      named_bar_sync(1, 128);  // the whole patch of tile t has landed
      float v[48];
      if (row_ok) {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const float* row = patch + (2 * yl + ky) * pitch + lead + 6 * xl;
          if (lead) {  // odd start: one word, five aligned pairs, one word
            v[12 * ky] = row[0];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const float2 f = *reinterpret_cast<const float2*>(row + 1 + 2 * j);
              v[12 * ky + 1 + 2 * j] = f.x;
              v[12 * ky + 2 + 2 * j] = f.y;
            }
            v[12 * ky + 11] = row[11];
          } else {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const float2 f = *reinterpret_cast<const float2*>(row + 2 * j);
              v[12 * ky + 2 * j] = f.x;
              v[12 * ky + 2 * j + 1] = f.y;
            }
          }
        }
      }
      named_bar_sync(1, 128);  // everyone holds its patch values in registers: the staging buffer is free again
      // the next tile's loads fly while this tile is split and stored (one staging buffer: no shared memory left for two)
      if (t + (int)gridDim.x < p.total_tiles) stage(t + gridDim.x);
      mbar_wait(&a_empty[0], ((uint32_t)lt & 1u) ^ 1u);  // the MMAs of the previous tile have retired
      if (row_ok) {
        uint8_t* row_hi = a_buf + (size_t)r * C1_ROWB;
        uint8_t* row_lo = row_hi + C1_ABYTES;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
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
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[0]);
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_bf16(N);
    const uint32_t a_lo0 = (smem_u32(a_buf) & 0x3FFFF) >> 4, b_lo0 = (smem_u32(b_img) & 0x3FFFF) >> 4;
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const uint32_t ph = (uint32_t)lt & 1u;
      mbar_wait(&tmem_empty_bar[0], ph ^ 1u);
      mbar_wait(&a_full[0], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base;
        const uint32_t a_hi = a_lo0, a_lo = a_hi + (C1_ABYTES >> 4);
        const uint32_t b_hi = b_lo0, b_lo = b_lo0 + (uint32_t)(N * C1_ROWB >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, umma_desc_from_lo<64>(a_hi + 2 * k), umma_desc_from_lo<64>(b_hi + 2 * k), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, umma_desc_from_lo<64>(a_lo + 2 * k), umma_desc_from_lo<64>(b_hi + 2 * k), idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, umma_desc_from_lo<64>(a_hi + 2 * k), umma_desc_from_lo<64>(b_lo + 2 * k), idesc, 1u);
        umma_commit(&a_empty[0]);
        umma_commit(&tmem_full_bar[0]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 5..8, warp w owns TMEM lanes 32*(w%4).. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int xl = r % p.TW, yl = r / p.TW;
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const int b = t / tiles_per_img, ti = t % tiles_per_img;
      const int oy = (ti / p.tiles_x) * p.TH + yl, ox = (ti % p.tiles_x) * p.TW + xl;
      const bool valid = yl < p.TH && oy < p.Ho && ox < p.Wo;
      mbar_wait_relaxed(&tmem_full_bar[0], (uint32_t)lt & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const size_t opix = (((size_t)b * p.Ho + oy) * p.Wo + ox) * 64;
#pragma unroll 1
      for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (c == N / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[0]);
        }
        const bool quads = (p.TW & 3) == 0;  // rows 4q .. 4q+3 are consecutive pixels of one image row
        if (!quads && !valid) continue;
        const int enc = c >> 1, ch0 = (c & 1) * 32;
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + c * 32);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          float f[4] = {__fadd_rn(__uint_as_float(v[4 * j]), bb.x), __fadd_rn(__uint_as_float(v[4 * j + 1]), bb.y),
                        __fadd_rn(__uint_as_float(v[4 * j + 2]), bb.z), __fadd_rn(__uint_as_float(v[4 * j + 3]), bb.w)};
#pragma unroll
          for (int k = 0; k < 4; ++k) f[k] = fmaxf(f[k], __fmul_rn(f[k], 0.2f));  // LeakyReLU(0.2)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            const uint32_t hw = *reinterpret_cast<const uint32_t*>(&hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * k] - __uint_as_float(hw << 16), f[2 * k + 1] - __uint_as_float(hw & 0xFFFF0000u));
            hi[2 * j + k] = hw;
            lo[2 * j + k] = *reinterpret_cast<const uint32_t*>(&ll);
          }
        }
        __nv_bfloat16* dh = p.out_hi[enc] + opix + ch0;  // 64-byte aligned: whole-sector 256-bit stores
        __nv_bfloat16* dl = p.out_lo[enc] + opix + ch0;
        if (quads) {
          const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
          quad_transpose16(hi, lane);
          quad_transpose16(lo, lane);
          const int l4 = lane & 3, q0 = lane & ~3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!((vmask >> (q0 + j)) & 1u)) continue;
            const long long d = (long long)(j - l4) * 64 + l4 * 8;  // row of quad lane j, piece l4 (8 bf16 = 16 bytes)
            *reinterpret_cast<uint4*>(dh + d) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4*>(dl + d) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          st_global_v8(dh + 16 * j, hi[8 * j], hi[8 * j + 1], hi[8 * j + 2], hi[8 * j + 3], hi[8 * j + 4], hi[8 * j + 5], hi[8 * j + 6], hi[8 * j + 7]);
          st_global_v8(dl + 16 * j, lo[8 * j], lo[8 * j + 1], lo[8 * j + 2], lo[8 * j + 3], lo[8 * j + 4], lo[8 * j + 5], lo[8 * j + 6], lo[8 * j + 7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, N);
}

// weights of one or two encoders, (4,4,3,64) fp32 each -> [hi | lo][N][64] bf16 rows in the swizzled smem image
__global__ void conv1_pack_kernel(const float* __restrict__ w0, const float* __restrict__ w1, uint8_t* __restrict__ img, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C1_K) return;
  const int n = i / C1_K, k = i % C1_K;
  const float* w = n < 64 ? w0 : w1;
  const float v = k < 48 ? w[k * 64 + (n & 63)] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  const size_t off = (size_t)n * C1_ROWB + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(img + (size_t)N * C1_ROWB + off) = l;
}

size_t conv1_tc_image_bytes(int n_enc) { return (size_t)2 * 64 * n_enc * C1_ROWB; }

int conv1_tc_pack(const float* w0, const float* w1, uint8_t* img, cudaStream_t st) {
  const int N = w1 ? 128 : 64;
  conv1_pack_kernel<<<(N * C1_K + 255) / 256, 256, 0, st>>>(w0, w1, img, N);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("conv1_pack_kernel");
  return CIC_OK;
}

template <int N>
static int launch_conv1_n(const Conv1Params& p, cudaStream_t st) {
  const size_t smem = 2 * C1_ABYTES + 2 * N * C1_ROWB + (C1_MAX_PATCH + 8) * sizeof(float) + 128 + 1024;
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.done();
  }
  const int slots = C1_CTAS * sm_count();
  conv1_tc_kernel<N><<<p.total_tiles < slots ? p.total_tiles : slots, 288, smem, st>>>(p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("conv1_tc_kernel");
  g_last_kernel_kind = KK_TC_CONV1;
  return CIC_OK;
}

// x: (batch, H, W, 3) fp32 (or tiles of larger images, see TileMap); out_*[e]: (batch, H/2, W/2, 64) bf16 of encoder e
int launch_conv1_tc(const float* x, const uint8_t* wimg, const float* bias, int n_enc, __nv_bfloat16* const* out_hi,
                    __nv_bfloat16* const* out_lo, int batch, int H, int W, const TileMap& tm, cudaStream_t st) {
  CIC_REQUIRE(n_enc == 1 || n_enc == 2, "conv1_tc: one or two encoders");
  CIC_REQUIRE(H % 2 == 0 && W % 2 == 0 && H > 0 && W > 0, "conv1_tc: H and W must be even");
  if (batch == 0) return CIC_OK;
  Conv1Params p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.tm = tm; p.batch = batch; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
  int tw = 64;
  while (tw > p.Wo && tw > 1) tw >>= 1;  // largest power of two <= Wo (capped at 64: a 2 x 64 tile needs a 6 x 130 pixel patch)
  p.TW = tw; p.TH = TC_BM / tw;
  if (p.TH > p.Ho) {
    int th = 1;
    while (th * 2 <= p.Ho) th *= 2;
    p.TH = th;
  }
  while (p.TH > 1 && (2 * p.TH + 2) * (2 * p.TW + 2) * 3 > C1_MAX_PATCH) p.TH >>= 1;  // one-pixel-wide outputs: fewer rows per tile
  CIC_REQUIRE((2 * p.TH + 2) * (2 * p.TW + 2) * 3 <= C1_MAX_PATCH, "conv1_tc: patch too large");
  p.tiles_x = (p.Wo + p.TW - 1) / p.TW; p.tiles_y = (p.Ho + p.TH - 1) / p.TH;
  const long long total = (long long)batch * p.tiles_x * p.tiles_y;
  CIC_REQUIRE(total < 2147483647LL, "conv1_tc: too many tiles");
  p.total_tiles = (int)total;
  p.pad_t = same_pad_before(H, 4, 2); p.pad_l = same_pad_before(W, 4, 2);
  p.wimg = wimg; p.bias = bias;
  for (int e = 0; e < n_enc; ++e) { p.out_hi[e] = out_hi[e]; p.out_lo[e] = out_lo[e]; }
  return n_enc == 2 ? launch_conv1_n<128>(p, st) : launch_conv1_n<64>(p, st);
}

}  // namespace cic
