// tcgen05 implicit-GEMM kernel + host-side launch (tensor-map encoding, grid).
#include "tc_gemm.cuh"
#include "tc_host.cuh"

#include <mutex>

namespace cic {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int BK, bool SPLIT>
__global__ void __launch_bounds__(192, TcCfg<BN, BK, SPLIT>::kMinCtas)
tc_gemm_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<BN, BK, SPLIT>;
  constexpr int kStages = Cfg::kStages;
  constexpr int CH = Cfg::kChunk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  const int tile = blockIdx.x;
  const int tix = tile % p.tiles_x;
  const int tiy = (tile / p.tiles_x) % p.tiles_y;
  const int tib = tile / (p.tiles_x * p.tiles_y);
  const int ox0 = tix * p.TW, oy0 = tiy * p.TH, b0 = tib * p.TB;
  const int n_tile = blockIdx.y;
  const int phase = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int per_split = (p.kblocks + p.splits - 1) / p.splits;
  const int kb0 = split * per_split;
  const int kb1 = min(p.kblocks, kb0 + per_split);
  const int nkb = kb1 - kb0;  // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.b[0]);
    if (SPLIT) { prefetch_tmap(&maps.a[0][1]); prefetch_tmap(&maps.b[1]); }
    if (p.nsrc > 1) { prefetch_tmap(&maps.a[1][0]); if (SPLIT) prefetch_tmap(&maps.a[1][1]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int cpt = p.src_blocks[0] + p.src_blocks[1];
      const uint32_t rows = (uint32_t)(p.TW * p.TH * p.TB);
      const uint32_t tx_bytes = (SPLIT ? 2u : 1u) * (rows + (uint32_t)BN) * (uint32_t)(2 * BK);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (uint32_t)(i / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        const int kb = kb0 + i;
        const int tap = kb / cpt, ch = kb % cpt;
        const int src = ch >= p.src_blocks[0] ? 1 : 0;
        const int cblk = src ? ch - p.src_blocks[0] : ch;
        const TcTap t = p.taps[phase][tap];
        const int c = p.src_coff[src] + cblk * BK + t.dc;
        uint8_t* st = smem + s * Cfg::kStageBytes;
        uint8_t* a_hi = st;
        uint8_t* a_lo = st + Cfg::kABytes;
        uint8_t* b_hi = st + (SPLIT ? 2 : 1) * Cfg::kABytes;
        uint8_t* b_lo = b_hi + Cfg::kBBytes;
        if (p.a5d) {
          tma_load_5d(a_hi, &maps.a[src][0], &full_bar[s], c, ox0 + t.dx, t.pz, oy0 + t.dy, b0);
          if (SPLIT) tma_load_5d(a_lo, &maps.a[src][1], &full_bar[s], c, ox0 + t.dx, t.pz, oy0 + t.dy, b0);
        } else {
          tma_load_4d(a_hi, &maps.a[src][0], &full_bar[s], c, ox0 + t.dx, oy0 + t.dy, b0);
          if (SPLIT) tma_load_4d(a_lo, &maps.a[src][1], &full_bar[s], c, ox0 + t.dx, oy0 + t.dy, b0);
        }
        const int bn = phase * p.N_pad + n_tile * BN;
        const int bz = p.b_batched ? b0 : 0;
        tma_load_3d(b_hi, &maps.b[0], &full_bar[s], kb * BK, bn, bz);
        if (SPLIT) tma_load_3d(b_lo, &maps.b[1], &full_bar[s], kb * BK, bn, bz);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (uint32_t)(i / kStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t a_hi = st, a_lo = st + Cfg::kABytes;
        const uint32_t b_hi = st + (SPLIT ? 2 : 1) * Cfg::kABytes, b_lo = b_hi + Cfg::kBBytes;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, umma_desc_kmajor<BK>(a_hi + k * 32), umma_desc_kmajor<BK>(b_hi + k * 32), idesc, (i | k) != 0);
        if (SPLIT) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base, umma_desc_kmajor<BK>(a_lo + k * 32), umma_desc_kmajor<BK>(b_hi + k * 32), idesc, 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base, umma_desc_kmajor<BK>(a_hi + k * 32), umma_desc_kmajor<BK>(b_lo + k * 32), idesc, 1u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4).. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile
    const int xl = r % p.TW, yl = (r / p.TW) % p.TH, bl = r / (p.TW * p.TH);
    const int ox = ox0 + xl, oy = oy0 + yl, b = b0 + bl;
    const bool valid = (bl < p.TB) && ox < p.Wo && oy < p.Ho && b < p.batch;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long opix = ((long long)b * p.out_H + (oy * p.out_ys + p.out_y0[phase])) * p.out_W + (ox * p.out_xs + p.out_x0[phase]);
    const long long mrow = ((long long)b * p.Ho + oy) * p.Wo + ox;
#pragma unroll 1
    for (int c = 0; c < BN / CH; ++c) {
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores of the previous chunk
      if (CH == 32) tmem_ld32(taddr + (uint32_t)(c * CH), v);
      else tmem_ld16(taddr + (uint32_t)(c * CH), v);
      tmem_ld_wait();
      const int n0 = n_tile * BN + c * CH;
      const int nv = min(CH, p.N - n0);  // valid channels of this chunk
      if (!valid || nv <= 0) continue;
      if (p.out_mode == TC_OUT_PARTIAL) {
        float* dst = reinterpret_cast<float*>(p.out_hi) + ((long long)split * p.m_total + mrow) * p.N + n0;
        if (nv == CH && (p.N & 3) == 0) {
#pragma unroll
          for (int j = 0; j < CH / 4; ++j)
            reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (j < nv) dst[j] = __uint_as_float(v[j]);
        }
        continue;
      }
      float f[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float x = __fmul_rn(p.alpha, __uint_as_float(v[j]));
        if (j < nv) {
          if (p.bias) x = __fadd_rn(x, __ldg(p.bias + n0 + j));
          if (p.scale) x = __fadd_rn(__fmul_rn(x, __ldg(p.scale + n0 + j)), __ldg(p.shift + n0 + j));
        }
        f[j] = act_apply(x, p.act);
      }
      if (p.out_mode == TC_OUT_F32) {
        float* dst = reinterpret_cast<float*>(p.out_hi) + opix * p.out_ld + p.out_coff + n0;
        if (nv == CH && ((p.out_ld | p.out_coff) & 3) == 0) {
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (j < nv) dst[j] = f[j];
        }
      } else if (p.out_mode == TC_OUT_BF16) {  // host guarantees N % CH == 0 and 16-byte aligned records
        const long long idx = opix * p.out_ld + p.out_coff + n0;
        if (p.res_hi) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float rv = __bfloat162float(p.res_hi[idx + j]);
            if (p.res_lo) rv += __bfloat162float(p.res_lo[idx + j]);
            f[j] = __fadd_rn(f[j], rv);
          }
        }
        uint32_t hi[CH / 2], lo[CH / 2];
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
          hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
          lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const int reps = p.up2 ? 2 : 1;
        for (int ry = 0; ry < reps; ++ry)
          for (int rx = 0; rx < reps; ++rx) {
            long long o = idx;
            if (p.up2) o = ((((long long)b * p.out_H + (oy * 2 + ry)) * p.out_W) + (ox * 2 + rx)) * p.out_ld + p.out_coff + n0;
            uint4* dh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_hi) + o);
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            if (p.out_lo) {
              uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_lo) + o);
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
      } else {  // TC_OUT_BF16_T: out[b][n][position] (V^T for the attention PV product)
        const long long how = (long long)p.Ho * p.Wo;
        const long long pos = (long long)oy * p.Wo + ox;
        __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.out_hi);
        __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(p.out_lo);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (j >= nv) break;
          const long long o = ((long long)b * p.N + n0 + j) * how + pos;
          const __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
          oh[o] = h;
          if (ol) ol[o] = __float2bfloat16_rn(f[j] - __bfloat162float(h));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

int tc_encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return CIC_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  // the inner box is one swizzle atom: 64 bf16 -> 128-byte swizzle, 32 bf16 -> 64-byte swizzle
  const CUtensorMapSwizzle sw = box[0] == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] base %p", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, base);
    return CIC_ERR_CUDA;
  }
  return CIC_OK;
}

template <int BN, int BK, bool SPLIT>
static int launch_one(const TcMaps& maps, const TcParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN, BK, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, BK, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  tc_gemm_kernel<BN, BK, SPLIT><<<grid, 192, Cfg::kSmemBytes, st>>>(maps, p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("tc_gemm_kernel");
  return CIC_OK;
}

template <int BK, bool SPLIT>
static int launch_bn(const TcMaps& maps, const TcParams& p, int block_n, dim3 grid, cudaStream_t st) {
  switch (block_n) {
    case 16: return launch_one<16, BK, SPLIT>(maps, p, grid, st);
    case 32: return launch_one<32, BK, SPLIT>(maps, p, grid, st);
    case 64: return launch_one<64, BK, SPLIT>(maps, p, grid, st);
    case 128: return launch_one<128, BK, SPLIT>(maps, p, grid, st);
    case 256:
      if (!SPLIT && BK == 64) return launch_one<256, 64, false>(maps, p, grid, st);
      break;
  }
  set_error("tc_gemm: unsupported N tile %d (K block %d, split=%d)", block_n, BK, (int)SPLIT);
  return CIC_ERR_INVALID;
}

int launch_tc_gemm(const TcMaps& maps, const TcParams& p, int block_n, int block_k, bool split, cudaStream_t st) {
  CIC_REQUIRE(block_n > 0 && p.N_pad % block_n == 0 && p.N <= p.N_pad, "tc_gemm: padded N=%d is not a multiple of the N tile %d", p.N_pad, block_n);
  CIC_REQUIRE(p.kblocks >= p.splits && p.splits >= 1, "tc_gemm: bad split-K %d for %d K blocks", p.splits, p.kblocks);
  const int per = (p.kblocks + p.splits - 1) / p.splits;
  CIC_REQUIRE((p.splits - 1) * per < p.kblocks, "tc_gemm: split-K leaves an empty split");
  CIC_REQUIRE(p.TW * p.TH * p.TB <= TC_BM && p.TW >= 1, "tc_gemm: bad M tile");
  const long long mt = (long long)p.tiles_x * p.tiles_y * p.tiles_b;
  CIC_REQUIRE(mt > 0 && mt < 2147483647LL, "tc_gemm: bad tile count");
  dim3 grid((unsigned)mt, (p.N + block_n - 1) / block_n, p.nphases * p.splits);
  CIC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "tc_gemm: grid too large");
  if (block_k == 64) return split ? launch_bn<64, true>(maps, p, block_n, grid, st) : launch_bn<64, false>(maps, p, block_n, grid, st);
  if (block_k == 32) return split ? launch_bn<32, true>(maps, p, block_n, grid, st) : launch_bn<32, false>(maps, p, block_n, grid, st);
  set_error("tc_gemm: unsupported K block %d", block_k);
  return CIC_ERR_INVALID;
}

}  // namespace cic
