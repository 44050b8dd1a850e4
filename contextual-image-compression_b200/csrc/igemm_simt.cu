// fp32 implicit-GEMM convolution / dense / batched matmul on the CUDA cores (CIC_PREC_FP32).
//
// GEMM view: M = batch*Ho*Wo output positions, N = Cout, K = kh*kw*Cin with K ordered (ky,kx,ci),
// which is exactly the flattening of a Keras Conv2D kernel (kh,kw,Cin,Cout) to a row-major [K][N]
// matrix.  A is gathered on the fly from NHWC activations (TF 'same' zero padding by bounds check,
// optional nearest x2 up-sampling and channel concatenation of two sources folded into the gather);
// transposed 4x4/stride-2 convolutions run as four 2x2 output-phase convolutions.
// Tile 128 x BN x 16, 256 threads, 8 x (BN/16) register tile, register-staged prefetch of the next
// K chunk.  fp32 FMA accumulation in ascending-K order inside a split; split-K partials are combined
// in a fixed order by splitk_reduce_kernel so results are bit-reproducible run to run.
#include "igemm_simt.cuh"

namespace cic {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int AS_LD = BM + 4;  // 16-byte aligned rows

struct RowInfo {
  int b, oy, ox;
  bool valid;
};

__device__ __forceinline__ RowInfo decode_row(const IGemmParams& p, int tile_x, int r) {
  RowInfo ri;
  const int how = p.Ho * p.Wo;
  int m_in;
  if (p.b_batch_stride != 0) {
    const int tpb = (how + BM - 1) / BM;
    ri.b = tile_x / tpb;
    m_in = (tile_x % tpb) * BM + r;
    ri.valid = m_in < how;
  } else {
    const long long m = (long long)tile_x * BM + r;
    ri.valid = m < (long long)p.batch * how;
    ri.b = (int)(m / how);
    m_in = (int)(m % how);
  }
  ri.oy = m_in / p.Wo;
  ri.ox = m_in % p.Wo;
  return ri;
}

__device__ __forceinline__ const float* a_address(const IGemmParams& p, const RowInfo& ri, int ky, int kx, int ci,
                                                  bool& inb) {
  const int iy = ri.oy * p.stride + ky - p.pad_t;
  const int ix = ri.ox * p.stride + kx - p.pad_l;
  inb = ri.valid && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
  const int s = (p.nsrc > 1 && ci >= p.src[0].C) ? 1 : 0;
  const int c = s ? ci - p.src[0].C : ci;
  const ConvSrc& src = p.src[s];
  const int sy = src.up ? (iy >> 1) : iy, sx = src.up ? (ix >> 1) : ix;
  const int Hs = src.up ? (p.H >> 1) : p.H, Ws = src.up ? (p.W >> 1) : p.W;
  return src.ptr + (((long long)ri.b * Hs + sy) * Ws + sx) * src.ld + c;
}

__device__ __forceinline__ void epilogue_store(const IGemmParams& p, const RowInfo& ri, int n, float acc) {
  float v = __fmul_rn(p.alpha, acc);
  if (p.bias) v = __fadd_rn(v, p.bias[n]);
  if (p.scale) v = __fadd_rn(__fmul_rn(v, p.scale[n]), p.shift[n]);
  v = act_apply(v, p.act);
  const long long pix = ((long long)ri.b * p.out_H + (ri.oy * p.out_ys + p.out_y0)) * p.out_W + (ri.ox * p.out_xs + p.out_x0);
  const long long idx = pix * p.out_ld + p.out_coff + n;
  if (p.residual) v = __fadd_rn(v, p.residual[idx]);
  p.out[idx] = v;
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(256)
igemm_f32_kernel(const IGemmParams p) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.y * BN;
  const int K = p.kh * p.kw * p.Cin;
  const int nchunks = (K + BK - 1) / BK;
  const int cps = (nchunks + p.splits - 1) / p.splits;
  const int split = blockIdx.z;
  const int kc_begin = split * cps;
  const int kc_end = min(nchunks, kc_begin + cps);

  // rows this thread gathers for A
  RowInfo rowA[2];
  if (VEC) {
    rowA[0] = decode_row(p, blockIdx.x, tid >> 2);
    rowA[1] = decode_row(p, blockIdx.x, (tid >> 2) + 64);
  } else {
    rowA[0] = decode_row(p, blockIdx.x, tid & 127);
  }
  const float* Bbase = p.Bmat;
  if (p.b_batch_stride != 0) {
    const int tpb = (p.Ho * p.Wo + BM - 1) / BM;
    Bbase += (long long)(blockIdx.x / tpb) * p.b_batch_stride;
  }

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2];
  float rs[8];
  float rb[4];

  auto load_global = [&](int kc) {
    const int kbase = kc * BK;
    if (VEC) {
      const int tap = kbase / p.Cin;
      const int ci = kbase % p.Cin + (tid & 3) * 4;
      const int ky = tap / p.kw, kx = tap % p.kw;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        bool inb;
        const float* ap = a_address(p, rowA[j], ky, kx, ci, inb);
        ra[j] = inb ? __ldg(reinterpret_cast<const float4*>(ap)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = kbase + (tid >> 7) + 2 * j;
        float v = 0.f;
        if (kk < K) {
          const int tap = kk / p.Cin, ci = kk % p.Cin;
          bool inb;
          const float* ap = a_address(p, rowA[0], tap / p.kw, tap % p.kw, ci, inb);
          if (inb) v = __ldg(ap);
        }
        rs[j] = v;
      }
    }
    // B tile: BK x BN
    if (!p.b_trans) {
      // 16 x BN floats; thread -> (k = tid / (BN/4), n4 = tid % (BN/4)) for BN=64; for BN=32 half the threads
      const int per_row = BN / 4;
      const int k = tid / per_row, n4 = tid % per_row;
      if (k < BK) {
        const int kk = kbase + k, n = n0 + n4 * 4;
        const float* bp = Bbase + (long long)kk * p.ldb + n;
        if (kk < K && n + 3 < p.N && (p.ldb & 3) == 0) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(bp));
          rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) rb[q] = (kk < K && n + q < p.N) ? __ldg(bp + q) : 0.f;
        }
      }
    } else {
      // B[k][n] = Bmat[n*ldb + k]: thread -> (n = tid / 4, k4 = tid % 4) covers 64 n x 16 k
      const int n = tid >> 2, k4 = (tid & 3) * 4;
      if (n < BN) {
        const int nn = n0 + n;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kk = kbase + k4 + q;
          rb[q] = (nn < p.N && kk < K) ? __ldg(Bbase + (long long)nn * p.ldb + kk) : 0.f;
        }
      }
    }
  };

  auto store_smem = [&]() {
    if (VEC) {
      const int kq = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = (tid >> 2) + 64 * j;
        As[kq + 0][r] = ra[j].x; As[kq + 1][r] = ra[j].y; As[kq + 2][r] = ra[j].z; As[kq + 3][r] = ra[j].w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) As[(tid >> 7) + 2 * j][tid & 127] = rs[j];
    }
    if (!p.b_trans) {
      const int per_row = BN / 4;
      const int k = tid / per_row, n4 = tid % per_row;
      if (k < BK) *reinterpret_cast<float4*>(&Bs[k][n4 * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
      const int n = tid >> 2, k4 = (tid & 3) * 4;
      if (n < BN) {
#pragma unroll
        for (int q = 0; q < 4; ++q) Bs[k4 + q][n] = rb[q];
      }
    }
  };

  if (kc_begin < kc_end) {
    load_global(kc_begin);
    store_smem();
  }
  __syncthreads();
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const bool has_next = kc + 1 < kc_end;
    if (has_next) load_global(kc + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      if (TN == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
        b[0] = t.x; b[1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (has_next) {
      store_smem();
      __syncthreads();
    }
  }

  // epilogue
  const int how = p.Ho * p.Wo;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const RowInfo ri = decode_row(p, blockIdx.x, ty * 8 + i);
    if (!ri.valid) continue;
    if (p.splits > 1) {
      const long long mg = (long long)ri.b * how + ri.oy * p.Wo + ri.ox;
      const long long Mtot = (long long)p.batch * how;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < p.N) p.partial[((long long)split * Mtot + mg) * p.N + n] = acc[i][j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < p.N) epilogue_store(p, ri, n, acc[i][j]);
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const IGemmParams p) {
  const int how = p.Ho * p.Wo;
  const long long Mtot = (long long)p.batch * how;
  const long long total = Mtot * p.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / p.N;
    const int n = (int)(i % p.N);
    float s = 0.f;
    for (int sp = 0; sp < p.splits; ++sp) s = __fadd_rn(s, p.partial[(long long)sp * total + i]);  // fixed order
    RowInfo ri;
    ri.valid = true;
    ri.b = (int)(m / how);
    const int r = (int)(m % how);
    ri.oy = r / p.Wo;
    ri.ox = r % p.Wo;
    epilogue_store(p, ri, n, s);
  }
}

int launch_igemm(const IGemmParams& p, cudaStream_t st) {
  const int how = p.Ho * p.Wo;
  long long mtiles;
  if (p.b_batch_stride != 0) mtiles = (long long)p.batch * ((how + BM - 1) / BM);
  else mtiles = ((long long)p.batch * how + BM - 1) / BM;
  if (mtiles == 0) return CIC_OK;
  CIC_REQUIRE(mtiles < 2147483647LL, "igemm: too many M tiles");
  const bool vec = (p.Cin % 16 == 0) && (p.src[0].C % 4 == 0) && (p.src[0].ld % 4 == 0) &&
                   (p.nsrc < 2 || (p.src[1].ld % 4 == 0 && p.src[1].C % 4 == 0)) &&
                   ((reinterpret_cast<uintptr_t>(p.src[0].ptr) & 15) == 0) &&
                   (p.nsrc < 2 || (reinterpret_cast<uintptr_t>(p.src[1].ptr) & 15) == 0);
  const bool bn32 = p.N <= 32;
  const int BNv = bn32 ? 32 : 64;
  dim3 grid((unsigned)mtiles, (p.N + BNv - 1) / BNv, p.splits);
  CIC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "igemm: grid too large");
  if (bn32) {
    if (vec) igemm_f32_kernel<32, true><<<grid, 256, 0, st>>>(p);
    else igemm_f32_kernel<32, false><<<grid, 256, 0, st>>>(p);
  } else {
    if (vec) igemm_f32_kernel<64, true><<<grid, 256, 0, st>>>(p);
    else igemm_f32_kernel<64, false><<<grid, 256, 0, st>>>(p);
  }
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("igemm_f32_kernel");
  g_last_kernel_kind = KK_SIMT;
  if (p.splits > 1) return launch_splitk_reduce(p, st);
  return CIC_OK;
}

int launch_splitk_reduce(const IGemmParams& p, cudaStream_t st) {
  const long long total = (long long)p.batch * p.Ho * p.Wo * p.N;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("splitk_reduce_kernel");
  return CIC_OK;
}

// ---------------------------------------------------------------------------------------------
// small-Cout direct conv: one thread per output pixel, weights in shared memory
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
conv_small_n_kernel(const SmallNParams p) {
  extern __shared__ float sw[];  // [K][4]
  const int K = p.kh * p.kw * p.Cin;
  for (int i = threadIdx.x; i < K * 4; i += blockDim.x) {
    const int k = i >> 2, n = i & 3;
    sw[i] = n < p.N ? p.Wmat[(long long)k * p.N + n] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)p.batch * p.H * p.W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int ox = (int)(pix % p.W);
  const int oy = (int)((pix / p.W) % p.H);
  const int b = (int)(pix / ((long long)p.W * p.H));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < p.kh; ++ky) {
    const int iy = oy + ky - p.pad_t;
    if (iy < 0 || iy >= p.H) continue;
    for (int kx = 0; kx < p.kw; ++kx) {
      const int ix = ox + kx - p.pad_l;
      if (ix < 0 || ix >= p.W) continue;
      const float* wk = sw + (size_t)((ky * p.kw + kx) * p.Cin) * 4;
      int cbase = 0;
      for (int s = 0; s < p.nsrc; ++s) {
        const ConvSrc& src = p.src[s];
        const int sy = src.up ? (iy >> 1) : iy, sx = src.up ? (ix >> 1) : ix;
        const int Hs = src.up ? (p.H >> 1) : p.H, Ws = src.up ? (p.W >> 1) : p.W;
        const float* xp = src.ptr + (((long long)b * Hs + sy) * Ws + sx) * src.ld;
        for (int c = 0; c < src.C; c += 4) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(xp + c));
          const float4* w4 = reinterpret_cast<const float4*>(wk + (size_t)(cbase + c) * 4);
          const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
          acc[0] = fmaf(x.x, w0.x, acc[0]); acc[1] = fmaf(x.x, w0.y, acc[1]); acc[2] = fmaf(x.x, w0.z, acc[2]); acc[3] = fmaf(x.x, w0.w, acc[3]);
          acc[0] = fmaf(x.y, w1.x, acc[0]); acc[1] = fmaf(x.y, w1.y, acc[1]); acc[2] = fmaf(x.y, w1.z, acc[2]); acc[3] = fmaf(x.y, w1.w, acc[3]);
          acc[0] = fmaf(x.z, w2.x, acc[0]); acc[1] = fmaf(x.z, w2.y, acc[1]); acc[2] = fmaf(x.z, w2.z, acc[2]); acc[3] = fmaf(x.z, w2.w, acc[3]);
          acc[0] = fmaf(x.w, w3.x, acc[0]); acc[1] = fmaf(x.w, w3.y, acc[1]); acc[2] = fmaf(x.w, w3.z, acc[2]); acc[3] = fmaf(x.w, w3.w, acc[3]);
        }
        cbase += src.C;
      }
    }
  }
  for (int n = 0; n < p.N; ++n) {
    float v = acc[n];
    if (p.bias) v = __fadd_rn(v, p.bias[n]);
    p.out[pix * p.N + n] = act_apply(v, p.act);
  }
}

int launch_conv_small_n(const SmallNParams& p, cudaStream_t st) {
  CIC_REQUIRE(p.N >= 1 && p.N <= 4, "conv_small_n: N must be <= 4");
  CIC_REQUIRE(p.src[0].C % 4 == 0 && p.src[0].ld % 4 == 0 && (p.nsrc < 2 || (p.src[1].C % 4 == 0 && p.src[1].ld % 4 == 0)),
              "conv_small_n: channel counts must be multiples of 4");
  const int K = p.kh * p.kw * p.Cin;
  const size_t smem = (size_t)K * 4 * sizeof(float);
  CIC_REQUIRE(smem <= 48 * 1024, "conv_small_n: K too large");
  const long long total = (long long)p.batch * p.H * p.W;
  if (total == 0) return CIC_OK;
  conv_small_n_kernel<<<(unsigned)((total + 127) / 128), 128, smem, st>>>(p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("conv_small_n_kernel");
  return CIC_OK;
}

// ---------------------------------------------------------------------------------------------
// MaxPooling2D((2,2), padding='same') (train_autoencoder.py:15,18)
// ---------------------------------------------------------------------------------------------
__global__ void maxpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, int batch, int H, int W, int C) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, C4 = C >> 2;
  const long long total = (long long)batch * Ho * Wo * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long r = i / C4;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = oy * 2 + dy, ix = ox * 2 + dx;
        if (iy < H && ix < W) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + iy) * W + ix) * C) + c4);
          m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
      }
    reinterpret_cast<float4*>(y + (((long long)b * Ho + oy) * Wo + ox) * C)[c4] = m;
  }
}

int launch_maxpool2x2(const float* x, float* y, int batch, int H, int W, int C, cudaStream_t st) {
  CIC_REQUIRE(C % 4 == 0, "maxpool2x2: C must be a multiple of 4");
  const long long total = (long long)batch * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
  if (total == 0) return CIC_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  maxpool2x2_kernel<<<(int)blocks, 256, 0, st>>>(x, y, batch, H, W, C);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("maxpool2x2_kernel");
  return CIC_OK;
}

// ---------------------------------------------------------------------------------------------
// row softmax (tf.nn.softmax(axis=-1), GAN_functions.py:359): one warp per row, in place
// ---------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float* __restrict__ x, long long rows, int cols) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* r = x + row * cols;
  float m = -INFINITY;
  for (int i = lane; i < cols; i += 32) m = fmaxf(m, r[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) {
    const float e = expf(r[i] - m);
    r[i] = e;
    s += e;
  }
  s = warp_sum(s);
  for (int i = lane; i < cols; i += 32) r[i] = __fdiv_rn(r[i], s);
}

int launch_softmax_rows(float* x, long long rows, int cols, cudaStream_t st) {
  if (rows == 0) return CIC_OK;
  const long long blocks = (rows + 7) / 8;
  CIC_REQUIRE(blocks < 2147483647LL, "softmax: too many rows");
  softmax_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, rows, cols);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("softmax_rows_kernel");
  return CIC_OK;
}

// GlobalAveragePooling2D: one CTA per (batch, 32-channel group); fixed-order reduction
__global__ void __launch_bounds__(256)
global_avg_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int hw, int C, int ldo) {
  __shared__ float part[8][32];
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int p = w; p < hw; p += 8) s += x[((long long)b * hw + p) * C + c];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
    y[(long long)b * ldo + c] = t / (float)hw;
  }
}

int launch_global_avg_pool(const float* x, float* y, int batch, int hw, int C, int ldo, cudaStream_t st) {
  if (batch == 0) return CIC_OK;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((C + 31) / 32, nb);
    global_avg_pool_kernel<<<grid, 256, 0, st>>>(x + (long long)b0 * hw * C, y + (long long)b0 * ldo, hw, C, ldo);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("global_avg_pool_kernel");
  }
  return CIC_OK;
}

// image (n,H,W,C) <-> tiles (n*ty*tx, tile, tile, C); tile order (image, tile row, tile col).  H, W need not be multiples of
// the tile: the gather replicates the image's last row / column into a ragged last tile, the scatter crops it.
template <bool GATHER>
__global__ void tile_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_img, int H, int W, int C,
                                 int tile) {
  const int tyn = (H + tile - 1) / tile, txn = (W + tile - 1) / tile;
  const long long total = (long long)n_img * tyn * txn * tile * tile * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the tile-major buffer
    const int c = (int)(i % C);
    long long r = i / C;
    const int lx = (int)(r % tile); r /= tile;
    const int ly = (int)(r % tile); r /= tile;
    const int tx = (int)(r % txn); r /= txn;
    const int ty = (int)(r % tyn);
    const int img = (int)(r / tyn);
    const int gy = ty * tile + ly, gx = tx * tile + lx;
    if (GATHER) {
      dst[i] = src[((((long long)img * H + min(gy, H - 1)) * W) + min(gx, W - 1)) * C + c];
    } else if (gy < H && gx < W) {
      dst[((((long long)img * H + gy) * W) + gx) * C + c] = src[i];
    }
  }
}

static int launch_tile_copy(bool gather, const float* src, float* dst, int n_img, int H, int W, int C, int tile, cudaStream_t st) {
  const long long total = (long long)n_img * ((H + tile - 1) / tile) * ((W + tile - 1) / tile) * tile * tile * C;
  if (total == 0) return CIC_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (gather) tile_copy_kernel<true><<<(int)blocks, 256, 0, st>>>(src, dst, n_img, H, W, C, tile);
  else tile_copy_kernel<false><<<(int)blocks, 256, 0, st>>>(src, dst, n_img, H, W, C, tile);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH(gather ? "tile_gather_kernel" : "tile_scatter_kernel");
  return CIC_OK;
}

int launch_tile_gather(const float* img, float* tiles, int n_img, int H, int W, int C, int tile, cudaStream_t st) {
  return launch_tile_copy(true, img, tiles, n_img, H, W, C, tile, st);
}

int launch_tile_scatter(const float* tiles, float* img, int n_img, int H, int W, int C, int tile, cudaStream_t st) {
  return launch_tile_copy(false, tiles, img, n_img, H, W, C, tile, st);
}

}  // namespace cic
