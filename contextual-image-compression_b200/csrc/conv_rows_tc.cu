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

constexpr int CR_SLOTS = 8;          // at most this many raster slots (input rows in flight)
constexpr int CR_STRIP_H = 128;
constexpr int CR_MAXK = 5;

struct ConvRowsParams {
  int H, W, batch;
  int ks, pad;               // square kernel, 'same' padding before
  int nblk;                  // 32-channel blocks (over both sources)
  int blk_src[4], blk_c0[4]; // source and first channel of each block
  int cout, act;
  int strips_x, strips_y, total_strips;
  int slots;                 // raster slots in the ring (3..CR_SLOTS)
  int slot_bytes;            // one raster slot: nblk rasters of rast_bytes each
  int rast_bytes;            // rw pixels x 64 B rounded up to 1 KB (swizzle phase of every raster starts at 0)
  int rw;                    // raster width in pixels = 128 + ks - 1
  const uint8_t* wimg;       // [nblk][ks (kx)][16 rows][64 B] pre-swizzled (SWIZZLE_64B)
  const float* bias;
  float* out;                // fp32, (batch, H, W, cout) or the image layout of `tm`
  int tm_tx, tm_ty, tm_IH, tm_IW;
  // FUSE2 (generator tail, GAN_functions.py:273 of both generators + :651-657, :682-684): block 0 = HQ generator, block 1 = LQ
  // generator, separate accumulators; the column owners apply tanh to both, form dt from the mask and write the BLEND
  const float* bias2;        // conv_out bias of the second generator
  const float* mask;         // (n_img, IH, IW) saliency mask
  const float* bpp;          // (n_img,) target bpp
  float* dt_out;             // (n_img, IH, IW) or NULL
  double* dt_sum;            // (n_img,) += sum of dt (zeroed by the launcher) or NULL
  float* out_hq;             // un-blended generator outputs (diagnostics) or NULL
  float* out_lq;
};

// tanh(z) = 1 - 2 / (e^(2z) + 1) on the SFU (ex2.approx, rcp.approx): absolute error ~1e-6 against tanhf - the fused tail's four
// column-owner warps do the work of three kernels' epilogues, and tanhf's ~20 instructions per value were what paced it
__device__ __forceinline__ float tanh_fast(float z) { return 1.0f - __fdividef(2.0f, __expf(2.0f * z) + 1.0f); }

template <int KS, int COUT, bool FUSE2>
__global__ void __launch_bounds__(192, 3)
conv_rows_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ ConvRowsParams p) {
  constexpr int STG = FUSE2 ? 32 : 16;   // TMEM columns per accumulator stage
  const int SLOTS = p.slots;   // 3..8 raster slots, chosen by the launcher so that three CTAs fit an SM
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_img = smem + (size_t)SLOTS * p.slot_bytes;
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
      for (int s = 0; s < SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * STG);
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
        if (++s == SLOTS) { s = 0; ph ^= 1u; }
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
          uint32_t d = tmem_base + (uint32_t)(as * STG);
          uint32_t first = 1;
          for (int k = 0; k < p.nblk; ++k) {
            if (FUSE2 && k == 1) { d += 16; first = 1; }   // the second generator accumulates in its own 16 columns
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
        if (++s == SLOTS) { s = 0; ph ^= 1u; }
        ++lt;
      }
    }
  } else {
    // ===== column owners: thread x accumulates output column x0 + x down the strip =====
    const int q = warp & 3;
    const int xl = q * 32 + lane;
    float bias[COUT], bias2[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      bias[o] = p.bias ? __ldg(p.bias + o) : 0.f;
      bias2[o] = (FUSE2 && p.bias2) ? __ldg(p.bias2 + o) : 0.f;
    }
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
      const int b = t / strips_per_item, ti = t % strips_per_item;
      const int y0 = (ti / p.strips_x) * CR_STRIP_H, x0 = (ti % p.strips_x) * TC_BM;
      const int gx = x0 + xl;
      size_t pix0;             // pixel index of (row 0, column gx) of this item in the output / mask layout
      size_t row_pix;          // pixels per output row
      int vh = p.H, vw = p.W;  // rows / columns of this item that exist in the image (a ragged last tile is cropped on store)
      int img = b;
      if (p.tm_tx) {
        const int tpi = p.tm_tx * p.tm_ty, tt = b % tpi;
        img = b / tpi;
        const int gy0 = (tt / p.tm_tx) * p.H, gx0 = (tt % p.tm_tx) * p.W;
        pix0 = (((size_t)img * p.tm_IH + (size_t)gy0) * p.tm_IW) + (size_t)gx0 + gx;
        row_pix = (size_t)p.tm_IW;
        vh = min(p.H, p.tm_IH - gy0);
        vw = min(p.W, p.tm_IW - gx0);
      } else {
        pix0 = ((size_t)b * p.H) * p.W + gx;
        row_pix = (size_t)p.W;
      }
      float* out_col = p.out + pix0 * COUT;
      const size_t row_stride = row_pix * COUT;
      float thr = 0.f;
      if (FUSE2) thr = rate_thr(rate_t(__ldg(p.bpp + img)));   // GAN_functions.py:631-644
      double dt_local = 0.0;
      if (FUSE2 && gx < vw) {
        for (int yp = y0; yp < y0 + 12 && yp < vh; ++yp) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.mask + pix0 + (size_t)yp * row_pix));
      }
      float acc[KS][COUT];   // acc[s]: output row (r + pad - s) while input row r is being added
      float acc2[FUSE2 ? KS : 1][COUT];
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int o = 0; o < COUT; ++o) { acc[s][o] = 0.f; if (FUSE2) acc2[s][o] = 0.f; }
#pragma unroll 1
      for (int i = 0; i < rows_in; ++i) {
        const int r = y0 - p.pad + i;
        if (FUSE2) {
          // the mask value of an output row is a dependent global load in this thread's serial walk down the strip: without
          // the prefetch every row paid a DRAM round trip (measured: 0.71 ms for the fused tail against an HBM floor of 0.39)
          const int yp = r + p.pad - (KS - 1) + 12;
          if (yp >= y0 && yp < y0 + CR_STRIP_H && yp < vh && gx < vw)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.mask + pix0 + (size_t)yp * row_pix));
        }
        if (r >= 0 && r < p.H) {
          const int as = lt & 1;
          mbar_wait_relaxed(&tmem_full_bar[as], ((uint32_t)lt >> 1) & 1u);
          tc_fence_after();
          uint32_t v[32];
          if (FUSE2) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * STG), v);
          else tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * STG), v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
          ++lt;
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
              acc[s][o] += __uint_as_float(v[s * COUT + o]);  // column n = ky * Cout + o, ky = s
              if (FUSE2) acc2[s][o] += __uint_as_float(v[16 + s * COUT + o]);
            }
        }
        // output row r + pad - (KS - 1) has received its last contribution
        const int yo = r + p.pad - (KS - 1);
        if (yo >= y0 && yo < y0 + CR_STRIP_H && yo < vh && gx < vw) {
          float* dst = out_col + (size_t)yo * row_stride;
          if (!FUSE2) {
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
              const float z = acc[KS - 1][o] + bias[o];
              dst[o] = p.act == CIC_ACT_TANH ? tanhf(z) : (p.act == CIC_ACT_SIGMOID ? 1.f / (1.f + expf(-z)) : z);
            }
          } else {
            const size_t pix = pix0 + (size_t)yo * row_pix;
            const float w = dyn_threshold(__ldg(p.mask + pix), thr);          // :651-657
            const float w1 = __fsub_rn(1.0f, w);
            dt_local += (double)w;
            if (p.dt_out) p.dt_out[pix] = w;
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
              const float h = tanh_fast(acc[KS - 1][o] + bias[o]), l = tanh_fast(acc2[KS - 1][o] + bias2[o]);   // :273 of both generators
              dst[o] = __fadd_rn(__fmul_rn(h, w), __fmul_rn(l, w1));                                    // :682-684
              if (p.out_hq) p.out_hq[pix * COUT + o] = h;
              if (p.out_lq) p.out_lq[pix * COUT + o] = l;
            }
          }
        }
#pragma unroll
        for (int s = KS - 1; s > 0; --s)
#pragma unroll
          for (int o = 0; o < COUT; ++o) { acc[s][o] = acc[s - 1][o]; if (FUSE2) acc2[s][o] = acc2[s - 1][o]; }
#pragma unroll
        for (int o = 0; o < COUT; ++o) { acc[0][o] = 0.f; if (FUSE2) acc2[0][o] = 0.f; }
      }
      if (FUSE2 && p.dt_sum) {   // hq_ratio numerator (GAN_test.py:312): one atomic per warp and strip
        const double sdt = warp_sum(dt_local);
        if (lane == 0) atomicAdd(p.dt_sum + img, sdt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * STG);
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

template <int KS, int COUT, bool FUSE2>
static int launch_rows(const TcMaps& maps, ConvRowsParams& p, cudaStream_t st) {
  // The column owners walk their strips serially (TMEM load, accumulate, store per row), so what hides their latencies is the
  // number of strips in flight per SM: size the raster ring for THREE co-resident CTAs (32 / 64 TMEM columns each) whenever at
  // least three slots fit (r02: the fused generator tail went 0.67 -> 0.56 ms from two to three CTAs per SM).
  const size_t fixed = (size_t)p.nblk * KS * 1024 + 512 + 1024;
  int slots = (int)((75 * 1024 - fixed) / p.slot_bytes);
  slots = slots > CR_SLOTS ? CR_SLOTS : slots;
  if (slots < 3) {   // wide inputs: two CTAs per SM, or one
    slots = (int)((112 * 1024 - fixed) / p.slot_bytes);
    if (slots < 3) slots = (int)((227 * 1024 - fixed) / p.slot_bytes);
    slots = slots > CR_SLOTS ? CR_SLOTS : slots;
    CIC_REQUIRE(slots >= 2, "conv_rows: input rows of %d bytes do not fit the raster ring", p.slot_bytes);
  }
  p.slots = slots;
  const size_t smem = (size_t)slots * p.slot_bytes + fixed;
  CIC_REQUIRE(smem <= 227 * 1024, "conv_rows: %zu bytes of shared memory", smem);
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(conv_rows_tc_kernel<KS, COUT, FUSE2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set.done();
  }
  const int per_sm = smem <= 75 * 1024 ? 3 : (smem <= 112 * 1024 ? 2 : 1);
  const int ctas = sm_count() * per_sm;
  conv_rows_tc_kernel<KS, COUT, FUSE2><<<p.total_strips < ctas ? p.total_strips : ctas, 192, smem, st>>>(maps, p);
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
  if (ks == 4) return launch_rows<4, 3, false>(maps, p, st);
  return launch_rows<3, 3, false>(maps, p, st);
}

// Generator tail in one kernel: Conv2D(3, k4, 'same', tanh) of the HQ and of the LQ generator (GAN_functions.py:273) on their
// deconv4 outputs (32 channels each), dynamic threshold dt = sigmoid((mask^0.7 - thr) * 20) (:651-657) and the blend
// hq * dt + lq * (1 - dt) (:682-684).  Replaces two conv_out launches + roi_blend: the un-blended fp32 images (24 B / pixel written,
// 24 B / pixel read back) never reach HBM.  wimg: the two generators' pre-swizzled conv_out images back to back.
int launch_gen_tail(const TcAct& hq, const TcAct& lq, const uint8_t* wimg, const float* bias_hq, const float* bias_lq, const float* mask,
                    const float* bpp, float* out, float* dt_out, double* dt_sum, float* out_hq, float* out_lq, int n_img, int batch, int H,
                    int W, const TileMap& tm, cudaStream_t st) {
  CIC_REQUIRE(hq.C == 32 && lq.C == 32 && wimg && mask && bpp && out, "gen_tail: two 32-channel sources, weights, mask, bpp and output");
  if (batch == 0) return CIC_OK;
  if (dt_sum) CIC_CHECK_CUDA(cudaMemsetAsync(dt_sum, 0, sizeof(double) * n_img, st));
  ConvRowsParams p;
  memset(&p, 0, sizeof(p));
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.H = H; p.W = W; p.batch = batch; p.ks = 4; p.pad = same_pad_before(H, 4, 1);
  p.rw = TC_BM + 4 - 1;
  const TcAct* srcs[2] = {&hq, &lq};
  for (int s = 0; s < 2; ++s) {
    const TcAct& a = *srcs[s];
    CIC_REQUIRE(a.ld % 8 == 0 && a.coff % 8 == 0, "gen_tail: source %d needs 16-byte aligned pixel records", s);
    p.blk_src[p.nblk] = s; p.blk_c0[p.nblk] = a.coff; ++p.nblk;
    const uint64_t dims[4] = {(uint64_t)a.ld, (uint64_t)W, (uint64_t)H, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)a.ld * 2, (uint64_t)W * a.ld * 2, (uint64_t)H * W * a.ld * 2};
    const uint32_t box[4] = {32, (uint32_t)p.rw, 1, 1};
    int rc = tc_encode_map(&maps.a[s][0], a.hi, 4, dims, str, box);
    if (rc) return rc;
  }
  p.rast_bytes = (p.rw * 64 + 1023) & ~1023;
  p.slot_bytes = p.nblk * p.rast_bytes;
  p.cout = 3; p.act = CIC_ACT_TANH; p.wimg = wimg; p.bias = bias_hq; p.bias2 = bias_lq; p.out = out;
  p.mask = mask; p.bpp = bpp; p.dt_out = dt_out; p.dt_sum = dt_sum; p.out_hq = out_hq; p.out_lq = out_lq;
  p.tm_tx = tm.tiles_x; p.tm_ty = tm.tiles_y; p.tm_IH = tm.IH; p.tm_IW = tm.IW;
  p.strips_x = (W + TC_BM - 1) / TC_BM; p.strips_y = (H + CR_STRIP_H - 1) / CR_STRIP_H;
  const long long total = (long long)batch * p.strips_x * p.strips_y;
  CIC_REQUIRE(total < 2147483647LL, "gen_tail: too many strips");
  p.total_strips = (int)total;
  return launch_rows<4, 3, true>(maps, p, st);
}

}  // namespace cic
