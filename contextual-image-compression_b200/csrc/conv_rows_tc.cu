// Tiny-Cout convolution on the tensor cores by column strips: the output convs of the generator
// (Conv2D(3, k4, 'same', tanh), GAN_functions.py:273) and of the autoencoder (Conv2D(3, k3, 'same', sigmoid) on
// concat(up(y5), x1r), train_autoencoder.py:33-35).
//
// With Cout = 3 the generic implicit GEMM wastes the tensor core twice: N is padded 3 -> 16 and every one of the
// kh*kw taps is its own 128 x 16 x Cin MMA group that re-reads the A tile from shared memory (the raster kernel runs
// this layer SMEM-read- and issue-bound at 9 % tensor-pipe activity).  Here the filter is factored the other way:
//   * kx is folded into K: the A operand of tap column kx is the SAME one-row raster started kx pixels later
//     (row-shifted UMMA descriptor), so K = kw * Cin per input row;
//   * ky is folded into N: column n = ky*Cout + o of the accumulator holds the partial sum that input row r
//     contributes to output row r + pad - ky, so N = kh * Cout (12 or 9, padded to 16);
// one input row therefore costs kw * Cin/16 MMAs of N = 16 (8 for the generator) instead of kh*kw*Cin/16 (32), and
// is read from L2 exactly once.  The y-direction sum needs no data exchange: every epilogue thread owns one image
// COLUMN and walks down the strip keeping the kh partially summed output rows in registers; when the last
// contribution of an output row has arrived it applies bias + activation and stores the pixel.
// CTA = one 128-column strip of STRIP_H output rows (persistent over strips); warp 0 = TMA producer (one raster =
// one input row of 128 + kw - 1 pixels per 32-channel block), warp 1 = MMA issuer, warps 2-5 = column owners.
#include "plan.cuh"
#include "tc_gemm.cuh"
#include "tc_host.cuh"

namespace cic {

constexpr int CR_SLOTS = 8;
constexpr int CR_STRIP_H = 128;
constexpr int CR_MAXK = 5;

struct ConvRowsParams {
  int H, W, batch;
  int ks, pad;               // square kernel, 'same' padding before
  int nblk;                  // 32-channel blocks (over both sources)
  int blk_src[4], blk_c0[4]; // source and first channel of each block
  int cout, act;
  int strips_x, strips_y, total_strips;
  int slot_bytes;            // one raster slot: nblk rasters of rast_bytes each
  int rast_bytes;            // rw pixels x 64 B rounded up to 1 KB (swizzle phase of every raster starts at 0)
  int rw;                    // raster width in pixels = 128 + ks - 1
  const uint8_t* wimg;       // [nblk][ks (kx)][16 rows][64 B] pre-swizzled (SWIZZLE_64B)
  const float* bias;
  float* out;                // fp32, (batch, H, W, cout) or the image layout of `tm`
  int tm_tx, tm_ty, tm_IH, tm_IW;
};

template <int KS, int COUT>
__global__ void __launch_bounds__(192, 2)
conv_rows_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ ConvRowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_img = smem + (size_t)CR_SLOTS * p.slot_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_img + (size_t)p.nblk * KS * 1024);
  uint64_t* a_empty = a_full + CR_SLOTS;
  uint64_t* tmem_full_bar = a_empty + CR_SLOTS;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.nblk * KS * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(b_img)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0][0]);
    if (p.blk_src[p.nblk - 1]) prefetch_tmap(&maps.a[1][0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < CR_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the weight image was written with generic stores
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int strips_per_item = p.strips_x * p.strips_y;
  const int rows_in = CR_STRIP_H + KS - 1;  // input rows a strip touches

  if (warp == 0) {
    // ===== TMA producer: one slot = the rasters (one per 32-channel block) of one input row =====
    uint32_t s = 0, ph = 0;
    const uint32_t tx_bytes = (uint32_t)(p.nblk * p.rw * 64);
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
      const int b = t / strips_per_item, ti = t % strips_per_item;
      const int y0 = (ti / p.strips_x) * CR_STRIP_H, x0 = (ti % p.strips_x) * TC_BM;
      for (int i = 0; i < rows_in; ++i) {
        const int r = y0 - p.pad + i;
        if (r < 0 || r >= p.H) continue;  // zero padding row: contributes nothing, skipped by every role
        mbar_wait_relaxed(&a_empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[s], tx_bytes);
          for (int k = 0; k < p.nblk; ++k)
            tma_load_4d(a_ring + (size_t)s * p.slot_bytes + (size_t)k * p.rast_bytes, &maps.a[p.blk_src[k]][0], &a_full[s], p.blk_c0[k], x0 - p.pad, r, b);
        }
        __syncwarp();
        if (++s == CR_SLOTS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per input row, D[128 x 16] = sum over (block, kx) of A(raster shifted kx rows) * W(block, kx) =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t a_lo0 = (smem_u32(a_ring) & 0x3FFFF) >> 4, b_lo0 = (smem_u32(b_img) & 0x3FFFF) >> 4;
    const uint32_t slot_lo = (uint32_t)p.slot_bytes >> 4, rast_lo = (uint32_t)p.rast_bytes >> 4;
    uint32_t s = 0, ph = 0;
    int lt = 0;  // valid rows processed by this CTA (TMEM stage = lt & 1)
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
      const int ti = t % strips_per_item;
      const int y0 = (ti / p.strips_x) * CR_STRIP_H;
      for (int i = 0; i < rows_in; ++i) {
        const int r = y0 - p.pad + i;
        if (r < 0 || r >= p.H) continue;
        const int as = lt & 1;
        mbar_wait(&tmem_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);
        mbar_wait(&a_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(as * 16);
          uint32_t first = 1;
          for (int k = 0; k < p.nblk; ++k) {
            const uint32_t a_blk = a_lo0 + s * slot_lo + (uint32_t)k * rast_lo;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              const uint32_t a = a_blk + (uint32_t)kx * 4;            // + kx pixels (64 B each)
              const uint32_t bw = b_lo0 + (uint32_t)((k * KS + kx) * 64);  // 1 KB per (block, kx)
              umma_bf16(d, umma_desc_from_lo<32>(a), umma_desc_from_lo<32>(bw), idesc, first ? 0u : 1u);
              umma_bf16(d, umma_desc_from_lo<32>(a + 2), umma_desc_from_lo<32>(bw + 2), idesc, 1u);
              first = 0;
            }
          }
          umma_commit(&a_empty[s]);
          umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (++s == CR_SLOTS) { s = 0; ph ^= 1u; }
        ++lt;
      }
    }
  } else {
    // ===== column owners: thread x accumulates output column x0 + x down the strip =====
    const int q = warp & 3;
    const int xl = q * 32 + lane;
    float bias[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) bias[o] = p.bias ? __ldg(p.bias + o) : 0.f;
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
      const int b = t / strips_per_item, ti = t % strips_per_item;
      const int y0 = (ti / p.strips_x) * CR_STRIP_H, x0 = (ti % p.strips_x) * TC_BM;
      const int gx = x0 + xl;
      float* out_col;
      size_t row_stride;
      int vh = p.H, vw = p.W;  // rows / columns of this item that exist in the image (a ragged last tile is cropped on store)
      if (p.tm_tx) {
        const int tpi = p.tm_tx * p.tm_ty, img = b / tpi, tt = b % tpi;
        const int gy0 = (tt / p.tm_tx) * p.H, gx0 = (tt % p.tm_tx) * p.W;
        out_col = p.out + ((((size_t)img * p.tm_IH + (size_t)gy0) * p.tm_IW) + (size_t)gx0 + gx) * COUT;
        row_stride = (size_t)p.tm_IW * COUT;
        vh = min(p.H, p.tm_IH - gy0);
        vw = min(p.W, p.tm_IW - gx0);
      } else {
        out_col = p.out + (((size_t)b * p.H) * p.W + gx) * COUT;
        row_stride = (size_t)p.W * COUT;
      }
      float acc[KS][COUT];  // acc[s]: output row (r + pad - s) while input row r is being added
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[s][o] = 0.f;
#pragma unroll 1
      for (int i = 0; i < rows_in; ++i) {
        const int r = y0 - p.pad + i;
        if (r >= 0 && r < p.H) {
          const int as = lt & 1;
          mbar_wait_relaxed(&tmem_full_bar[as], ((uint32_t)lt >> 1) & 1u);
          tc_fence_after();
          uint32_t v[32];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 16), v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
          ++lt;
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int o = 0; o < COUT; ++o) acc[s][o] += __uint_as_float(v[s * COUT + o]);  // column n = ky * Cout + o, ky = s
        }
        // output row r + pad - (KS - 1) has received its last contribution
        const int yo = r + p.pad - (KS - 1);
        if (yo >= y0 && yo < y0 + CR_STRIP_H && yo < vh && gx < vw) {
          float* dst = out_col + (size_t)yo * row_stride;
#pragma unroll
          for (int o = 0; o < COUT; ++o) {
            const float z = acc[KS - 1][o] + bias[o];
            dst[o] = p.act == CIC_ACT_TANH ? tanhf(z) : (p.act == CIC_ACT_SIGMOID ? 1.f / (1.f + expf(-z)) : z);
          }
        }
#pragma unroll
        for (int s = KS - 1; s > 0; --s)
#pragma unroll
          for (int o = 0; o < COUT; ++o) acc[s][o] = acc[s - 1][o];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[0][o] = 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32);
}

// (ks, ks, Cin, Cout) fp32 -> [blk][kx][16 rows n = ky*Cout + o][32 channels] bf16, rows swizzled (SWIZZLE_64B)
__global__ void conv_rows_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ img, int ks, int cin, int cout) {
  const int nblk = cin / 32;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblk * ks * 16 * 32) return;
  const int c = i % 32, n = (i / 32) % 16, kx = (i / 512) % ks, blk = i / (512 * ks);
  const int ky = n / cout, o = n % cout;
  float v = 0.f;
  if (ky < ks) v = w[(((size_t)ky * ks + kx) * cin + blk * 32 + c) * cout + o];
  const size_t off = (size_t)(blk * ks + kx) * 1024 + (size_t)n * 64 + ((((c >> 3) ^ ((n >> 1) & 3))) << 4) + (c & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
}

size_t conv_rows_image_bytes(int ks, int cin) { return (size_t)(cin / 32) * ks * 1024; }

int conv_rows_pack(const float* w, uint8_t* img, int ks, int cin, int cout, cudaStream_t st) {
  CIC_REQUIRE(cin % 32 == 0 && ks * cout <= 16 && ks <= CR_MAXK, "conv_rows: needs Cin %% 32 == 0 and ks * Cout <= 16");
  const int n = (cin / 32) * ks * 16 * 32;
  conv_rows_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, img, ks, cin, cout);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("conv_rows_pack_kernel");
  return CIC_OK;
}

template <int KS, int COUT>
static int launch_rows(const TcMaps& maps, const ConvRowsParams& p, cudaStream_t st) {
  const size_t smem = (size_t)CR_SLOTS * p.slot_bytes + (size_t)p.nblk * KS * 1024 + 512 + 1024;
  CIC_REQUIRE(smem <= 227 * 1024, "conv_rows: %zu bytes of shared memory", smem);
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(conv_rows_tc_kernel<KS, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set.done();
  }
  const int slots = sm_count() * (smem <= 112 * 1024 ? 2 : 1);  // two co-resident CTAs (32 TMEM columns each) when shared memory allows
  conv_rows_tc_kernel<KS, COUT><<<p.total_strips < slots ? p.total_strips : slots, 192, smem, st>>>(maps, p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("conv_rows_tc_kernel");
  g_last_kernel_kind = KK_TC_ROWS;
  return CIC_OK;
}

// Conv2D(cout <= 4, ks x ks, stride 1, 'same') + bias + activation on channel-concatenated bf16 sources (each C % 32 == 0)
int launch_conv_rows_tc(const TcAct* srcs, int nsrc, const uint8_t* wimg, const float* bias, int ks, int cout, int act, float* out,
                        int batch, int H, int W, const TileMap& tm, cudaStream_t st) {
  CIC_REQUIRE(nsrc == 1 || nsrc == 2, "conv_rows: one or two sources");
  CIC_REQUIRE((ks == 3 || ks == 4) && cout == 3, "conv_rows: built for 3x3 / 4x4 kernels with 3 output channels");
  if (batch == 0) return CIC_OK;
  ConvRowsParams p;
  memset(&p, 0, sizeof(p));
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.H = H; p.W = W; p.batch = batch; p.ks = ks; p.pad = same_pad_before(H, ks, 1);
  p.rw = TC_BM + ks - 1;
  for (int s = 0; s < nsrc; ++s) {
    const TcAct& a = srcs[s];
    CIC_REQUIRE(a.C % 32 == 0 && a.ld % 8 == 0 && a.coff % 8 == 0, "conv_rows: source %d needs C %% 32 == 0", s);
    for (int c0 = 0; c0 < a.C; c0 += 32) {
      CIC_REQUIRE(p.nblk < 4, "conv_rows: at most 128 input channels");
      p.blk_src[p.nblk] = s; p.blk_c0[p.nblk] = a.coff + c0; ++p.nblk;
    }
    const uint64_t dims[4] = {(uint64_t)a.ld, (uint64_t)W, (uint64_t)H, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)a.ld * 2, (uint64_t)W * a.ld * 2, (uint64_t)H * W * a.ld * 2};
    const uint32_t box[4] = {32, (uint32_t)p.rw, 1, 1};
    int rc = tc_encode_map(&maps.a[s][0], a.hi, 4, dims, str, box);
    if (rc) return rc;
  }
  p.rast_bytes = (p.rw * 64 + 1023) & ~1023;
  p.slot_bytes = p.nblk * p.rast_bytes;
  p.cout = cout; p.act = act; p.wimg = wimg; p.bias = bias; p.out = out;
  p.tm_tx = tm.tiles_x; p.tm_ty = tm.tiles_y; p.tm_IH = tm.IH; p.tm_IW = tm.IW;
  p.strips_x = (W + TC_BM - 1) / TC_BM; p.strips_y = (H + CR_STRIP_H - 1) / CR_STRIP_H;
  const long long total = (long long)batch * p.strips_x * p.strips_y;
  CIC_REQUIRE(total < 2147483647LL, "conv_rows: too many strips");
  p.total_strips = (int)total;
  if (ks == 4) return launch_rows<4, 3>(maps, p, st);
  return launch_rows<3, 3>(maps, p, st);
}

}  // namespace cic
