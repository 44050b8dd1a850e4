// Host-side assembly of tcgen05 layer launches + conversion kernels (see tc_host.cuh).
#include "tc_host.cuh"

namespace cic {

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p *= 2;
  return p;
}

// M tile = TW x TH x TB output positions (<= 128): minimise the padded position count
static void pick_tile(int Wo, int Ho, int batch, bool one_item, int& TW, int& TH, int& TB) {
  long long best = -1;
  for (int tw = 128; tw >= 1; tw /= 2) {
    if (tw > pow2_ceil(Wo)) continue;
    int th = 128 / tw;
    if (th > pow2_ceil(Ho)) th = pow2_ceil(Ho);
    int tb = one_item ? 1 : 128 / (tw * th);
    if (tb > pow2_ceil(batch)) tb = pow2_ceil(batch);
    const long long cost = (long long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((batch + tb - 1) / tb);  // tiles
    // every tile costs a full 128-row MMA: fewer tiles is better; ties -> wider rows
    if (best < 0 || cost < best) { best = cost; TW = tw; TH = th; TB = tb; }
  }
}

int tc_pick_block_n(int n_pad, bool split, int bk) {
  const int cap = (split || bk != 64) ? 128 : 256;
  for (int bn = cap; bn >= 16; bn /= 2)
    if (n_pad % bn == 0) return bn;
  return 0;
}

int tc_pick_block_k(const TcLayer& L) {
  for (int s = 0; s < L.nsrc; ++s)
    if (L.src[s].C % 64 != 0) return 32;
  return 64;
}

int tc_run_layer(const TcLayer& L, cudaStream_t st) {
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  TcParams p;
  memset(&p, 0, sizeof(p));
  const bool s2 = L.kind == TC_CONV_S2;
  const bool dc = L.kind == TC_DECONV_K4S2;
  CIC_REQUIRE(L.nsrc == 1 || L.nsrc == 2, "tc layer: nsrc must be 1 or 2");
  const int BK = tc_pick_block_k(L);
  for (int s = 0; s < L.nsrc; ++s) {
    CIC_REQUIRE(L.src[s].C % BK == 0 && L.src[s].ld % 8 == 0 && L.src[s].coff % 8 == 0,
                "tc layer: source %d needs C %% 32 == 0 and 16-byte aligned pixel records (C=%d, ld=%d, coff=%d)", s, L.src[s].C,
                L.src[s].ld, L.src[s].coff);
    CIC_REQUIRE(!L.split || L.src[s].lo, "tc layer: split mode needs the low part of source %d", s);
  }
  CIC_REQUIRE(!L.split || L.w.lo, "tc layer: split mode needs the low part of the weights");
  CIC_REQUIRE(!s2 || (L.H % 2 == 0 && L.W % 2 == 0), "tc layer: stride-2 path needs even H, W");
  const int Ho = s2 ? L.H / 2 : L.H, Wo = s2 ? L.W / 2 : L.W;
  pick_tile(Wo, Ho, L.batch, L.b_batched, p.TW, p.TH, p.TB);
  p.tiles_x = (Wo + p.TW - 1) / p.TW;
  p.tiles_y = (Ho + p.TH - 1) / p.TH;
  p.tiles_b = (L.batch + p.TB - 1) / p.TB;
  p.Wo = Wo; p.Ho = Ho; p.batch = L.batch;
  p.a5d = s2 ? 1 : 0;
  p.nsrc = L.nsrc;
  // activation tensor maps
  for (int s = 0; s < L.nsrc; ++s) {
    const TcAct& a = L.src[s];
    p.src_blocks[s] = a.C / BK;
    p.src_coff[s] = a.coff;
    for (int part = 0; part < (L.split ? 2 : 1); ++part) {
      const bf16* base = part ? a.lo : a.hi;
      int rc;
      if (!s2) {
        const uint64_t dims[4] = {(uint64_t)a.ld, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.batch};
        const uint64_t str[3] = {(uint64_t)a.ld * 2, (uint64_t)L.W * a.ld * 2, (uint64_t)L.H * L.W * a.ld * 2};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TB};
        rc = tc_encode_map(&maps.a[s][part], base, 4, dims, str, box);
      } else {
        // (x-parity, channel) merged | x/2 | y-parity | y/2 | batch
        const uint64_t dims[5] = {(uint64_t)2 * a.ld, (uint64_t)L.W / 2, 2, (uint64_t)L.H / 2, (uint64_t)L.batch};
        const uint64_t str[4] = {(uint64_t)2 * a.ld * 2, (uint64_t)L.W * a.ld * 2, (uint64_t)2 * L.W * a.ld * 2,
                                 (uint64_t)L.H * L.W * a.ld * 2};
        const uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.TW, 1, (uint32_t)p.TH, (uint32_t)p.TB};
        rc = tc_encode_map(&maps.a[s][part], base, 5, dims, str, box);
      }
      if (rc) return rc;
    }
  }
  // taps
  p.nphases = dc ? 4 : 1;
  if (dc) {
    p.ntaps = 4;
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          TcTap& t = p.taps[ph][ty * 2 + tx];
          t.dc = 0; t.pz = 0;
          t.dy = (int16_t)(ty - (py == 0 ? 1 : 0));
          t.dx = (int16_t)(tx - (px == 0 ? 1 : 0));
        }
      p.out_y0[ph] = (int8_t)py;
      p.out_x0[ph] = (int8_t)px;
    }
    p.out_ys = 2; p.out_xs = 2; p.out_H = 2 * L.H; p.out_W = 2 * L.W;
  } else {
    p.ntaps = L.kh * L.kw;
    CIC_REQUIRE(p.ntaps <= TC_MAX_TAPS, "tc layer: too many taps");
    for (int ky = 0; ky < L.kh; ++ky)
      for (int kx = 0; kx < L.kw; ++kx) {
        TcTap& t = p.taps[0][ky * L.kw + kx];
        const int dy = ky - L.pad_t, dx = kx - L.pad_l;
        if (s2) {
          const int yb = dy >= 0 ? dy / 2 : -((-dy + 1) / 2), xb = dx >= 0 ? dx / 2 : -((-dx + 1) / 2);
          t.dy = (int16_t)yb; t.pz = (int16_t)(dy - 2 * yb);
          t.dx = (int16_t)xb; t.dc = (int16_t)((dx - 2 * xb) * L.src[0].ld);
          CIC_REQUIRE(L.nsrc == 1, "tc layer: stride-2 path takes one source");
        } else {
          t.dy = (int16_t)dy; t.dx = (int16_t)dx; t.dc = 0; t.pz = 0;
        }
      }
    p.out_ys = 1; p.out_xs = 1; p.out_H = Ho * (L.epi.up2 ? 2 : 1); p.out_W = Wo * (L.epi.up2 ? 2 : 1);
  }
  const int cpt = p.src_blocks[0] + (L.nsrc > 1 ? p.src_blocks[1] : 0);
  p.kblocks = p.ntaps * cpt;
  CIC_REQUIRE((long long)p.kblocks * BK == L.w.K, "tc layer: weight K=%d does not match taps x channels = %d", L.w.K,
              p.kblocks * BK);
  p.splits = L.splits;
  p.N = L.N;
  p.N_pad = dc ? L.w.rows / 4 : L.w.rows;
  p.b_batched = L.b_batched ? 1 : 0;
  const int bn = tc_pick_block_n(p.N_pad, L.split, BK);
  CIC_REQUIRE(bn > 0 && L.N <= p.N_pad, "tc layer: N=%d (padded %d) has no supported tile", L.N, p.N_pad);
  CIC_REQUIRE(L.w.row_stride % 8 == 0 && L.w.batch_stride % 8 == 0, "tc layer: B rows must be 16-byte aligned");
  // weight / B map
  for (int part = 0; part < (L.split ? 2 : 1); ++part) {
    const uint64_t dims[3] = {(uint64_t)L.w.K, (uint64_t)L.w.rows, (uint64_t)L.w.batches};
    const uint64_t str[2] = {(uint64_t)L.w.row_stride * 2, (uint64_t)L.w.batch_stride * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)bn, 1};
    int rc = tc_encode_map(&maps.b[part], part ? L.w.lo : L.w.hi, 3, dims, str, box);
    if (rc) return rc;
  }
  // epilogue
  const TcEpilogue& e = L.epi;
  CIC_REQUIRE(L.splits == 1 || e.out_mode == TC_OUT_PARTIAL, "tc layer: split-K needs the partial output mode");
  CIC_REQUIRE(e.out_hi, "tc layer: null output");
  if (e.out_mode == TC_OUT_BF16) {
    const int ld = e.out_ld ? e.out_ld : L.N;
    CIC_REQUIRE(L.N % 16 == 0 && (bn >= 32 ? L.N % 32 == 0 : true) && ld % 8 == 0 && e.out_coff % 8 == 0,
                "tc layer: bf16 output needs N %% 16 == 0 and 16-byte aligned records (N=%d, ld=%d, coff=%d)", L.N, ld, e.out_coff);
  }
  p.bias = e.bias; p.scale = e.scale; p.shift = e.shift; p.alpha = e.alpha; p.act = e.act;
  p.out_mode = e.out_mode; p.out_hi = e.out_hi; p.out_lo = e.out_lo; p.res_hi = e.res_hi; p.res_lo = e.res_lo;
  p.out_ld = e.out_ld ? e.out_ld : L.N; p.out_coff = e.out_coff; p.up2 = e.up2;
  p.m_total = (long long)L.batch * Ho * Wo;
  return launch_tc_gemm(maps, p, bn, BK, L.split, st);
}

// ------------------------------------------------------------------------------------------------
// conversion kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float v, bf16& h, bf16& l) {
  h = __float2bfloat16_rn(v);
  l = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ src, int K, int N, int N_pad, int ld, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + tx;
    tile[i][tx] = (k < K && n < N) ? src[(size_t)k * ld + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + tx;
    if (n < N_pad && k < K) {
      bf16 h, l;
      split2(tile[tx][i], h, l);
      hi[(size_t)n * K + k] = h;
      if (lo) lo[(size_t)n * K + k] = l;
    }
  }
}

int tc_pack_weight(const float* src, int K, int N, int N_pad, int ld, bf16* hi, bf16* lo, cudaStream_t st) {
  dim3 grid((K + 31) / 32, (N_pad + 31) / 32);
  CIC_REQUIRE(grid.y <= 65535, "pack_weight: N too large");
  pack_weight_kernel<<<grid, 256, 0, st>>>(src, K, N, N_pad, ld, hi, lo);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("pack_weight_kernel");
  return CIC_OK;
}

__global__ void split_f32_kernel(const float* __restrict__ src, bf16* __restrict__ hi, bf16* __restrict__ lo, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bf16 h, l;
    split2(src[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

int tc_split_f32(const float* src, bf16* hi, bf16* lo, size_t n, cudaStream_t st) {
  if (n == 0) return CIC_OK;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  split_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, hi, lo, n);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("split_f32_kernel");
  return CIC_OK;
}

__global__ void join_to_f32_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, float* __restrict__ dst,
                                   size_t pixels, int C, int ld, int coff) {
  const size_t total = pixels * (size_t)C;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t px = i / C;
    const int c = (int)(i % C);
    const size_t j = px * ld + coff + c;
    float v = __bfloat162float(hi[j]);
    if (lo) v += __bfloat162float(lo[j]);
    dst[i] = v;
  }
}

int tc_join_to_f32(const bf16* hi, const bf16* lo, float* dst, size_t pixels, int C, int ld, int coff, cudaStream_t st) {
  const size_t total = pixels * (size_t)C;
  if (total == 0) return CIC_OK;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  join_to_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(hi, lo, dst, pixels, C, ld, coff);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("join_to_f32_kernel");
  return CIC_OK;
}

__global__ void tc_splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long M, int N,
                                        const float* __restrict__ bias, const float* __restrict__ scale,
                                        const float* __restrict__ shift, int act, float* __restrict__ out_f32,
                                        bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
  const long long total = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s = __fadd_rn(s, partial[(long long)sp * total + i]);  // fixed order
    if (bias) s = __fadd_rn(s, bias[n]);
    if (scale) s = __fadd_rn(__fmul_rn(s, scale[n]), shift[n]);
    s = act_apply(s, act);
    if (out_f32) out_f32[i] = s;
    if (out_hi) {
      bf16 h, l;
      split2(s, h, l);
      out_hi[i] = h;
      if (out_lo) out_lo[i] = l;
    }
  }
}

int tc_splitk_reduce(const float* partial, int splits, long long M, int N, const float* bias, const float* scale,
                     const float* shift, int act, float* out_f32, bf16* out_hi, bf16* out_lo, cudaStream_t st) {
  const long long total = M * N;
  if (total == 0) return CIC_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  tc_splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(partial, splits, M, N, bias, scale, shift, act, out_f32, out_hi, out_lo);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("tc_splitk_reduce_kernel");
  return CIC_OK;
}

// one warp per row; the row lives in registers when cols <= 1024 (32 values per lane)
__global__ void __launch_bounds__(256)
softmax_rows_split_kernel(const float* __restrict__ x, bf16* __restrict__ p_hi, bf16* __restrict__ p_lo, long long rows, int cols) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* r = x + row * cols;
  float m = -INFINITY;
  for (int i = lane; i < cols; i += 32) m = fmaxf(m, r[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) s += expf(r[i] - m);
  s = warp_sum(s);
  for (int i = lane; i < cols; i += 32) {
    const float pv = __fdiv_rn(expf(r[i] - m), s);
    bf16 h, l;
    split2(pv, h, l);
    p_hi[row * cols + i] = h;
    if (p_lo) p_lo[row * cols + i] = l;
  }
}

int tc_softmax_rows_split(const float* logits, bf16* p_hi, bf16* p_lo, long long rows, int cols, cudaStream_t st) {
  if (rows == 0) return CIC_OK;
  const long long blocks = (rows + 7) / 8;
  CIC_REQUIRE(blocks < 2147483647LL, "softmax: too many rows");
  softmax_rows_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, p_hi, p_lo, rows, cols);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("softmax_rows_split_kernel");
  return CIC_OK;
}

}  // namespace cic
