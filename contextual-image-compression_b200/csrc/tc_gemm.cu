// tcgen05 implicit-GEMM kernel + host-side launch (tensor-map encoding, grid).
#include "tc_gemm.cuh"
#include "tc_host.cuh"

#include <mutex>

namespace cic {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TileCoord {
  int ox0, oy0, b0;  // first output position of the M tile
  int n_tile, phase, split;
  int kb0, nkb;      // K-block range of this split
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t) {
  const int mt = p.tiles_x * p.tiles_y * p.tiles_b;
  int m, n, z;
  if (p.m_fast) {
    m = t % mt;
    const int r = t / mt;
    n = r % p.n_tiles;
    z = r / p.n_tiles;
  } else {
    n = t % p.n_tiles;
    const int r = t / p.n_tiles;
    const int zt = p.nphases * p.splits;
    z = r % zt;
    m = r / zt;
  }
  TileCoord c;
  c.ox0 = (m % p.tiles_x) * p.TW;
  c.oy0 = ((m / p.tiles_x) % p.tiles_y) * p.TH;
  c.b0 = (m / (p.tiles_x * p.tiles_y)) * p.TB;
  c.n_tile = n;
  c.phase = z / p.splits;
  c.split = z % p.splits;
  const int per_split = (p.kblocks + p.splits - 1) / p.splits;
  c.kb0 = c.split * per_split;
  c.nkb = min(p.kblocks, c.kb0 + per_split) - c.kb0;  // host guarantees >= 1
  return c;
}

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
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const TcEpiVec ev{p.epi.bias, p.epi.scale, p.epi.shift};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.b[0]);
    if (SPLIT) { prefetch_tmap(&maps.a[0][1]); prefetch_tmap(&maps.b[1]); }
    if (p.nsrc > 1) { prefetch_tmap(&maps.a[1][0]); if (SPLIT) prefetch_tmap(&maps.a[1][1]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
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
    // ===== TMA producer: warp-convergent, one elected lane issues; runs ahead across tile boundaries =====
    const int cpt = p.src_blocks[0] + p.src_blocks[1];
    const uint32_t rows = (uint32_t)(p.TW * p.TH * p.TB);
    const uint32_t tx_bytes = (SPLIT ? 2u : 1u) * (rows + (uint32_t)BN) * (uint32_t)(2 * BK);
    uint32_t s = 0, ph = 0;  // ring slot / phase
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      const int bn = tc.phase * p.N_pad + tc.n_tile * BN;
      const int bz = p.b_batched ? tc.b0 : 0;
      int tap = tc.kb0 / cpt, ch = tc.kb0 % cpt;
      for (int i = 0; i < tc.nkb; ++i) {
        mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          const int kb = tc.kb0 + i;
          const int src = ch >= p.src_blocks[0] ? 1 : 0;
          const int cblk = src ? ch - p.src_blocks[0] : ch;
          const TcTap tp = p.taps[tc.phase][tap];
          const int c = p.src_coff[src] + cblk * BK + tp.dc;
          uint8_t* st = smem + s * Cfg::kStageBytes;
          uint8_t* a_hi = st;
          uint8_t* a_lo = st + Cfg::kABytes;
          uint8_t* b_hi = st + (SPLIT ? 2 : 1) * Cfg::kABytes;
          uint8_t* b_lo = b_hi + Cfg::kBBytes;
          if (p.a5d) {
            tma_load_5d(a_hi, &maps.a[src][0], &full_bar[s], c, tc.ox0 + tp.dx, tp.pz, tc.oy0 + tp.dy, tc.b0);
            if (SPLIT) tma_load_5d(a_lo, &maps.a[src][1], &full_bar[s], c, tc.ox0 + tp.dx, tp.pz, tc.oy0 + tp.dy, tc.b0);
          } else {
            tma_load_4d(a_hi, &maps.a[src][0], &full_bar[s], c, tc.ox0 + tp.dx, tc.oy0 + tp.dy, tc.b0);
            if (SPLIT) tma_load_4d(a_lo, &maps.a[src][1], &full_bar[s], c, tc.ox0 + tp.dx, tc.oy0 + tp.dy, tc.b0);
          }
          tma_load_3d(b_hi, &maps.b[0], &full_bar[s], kb * BK, bn, bz);
          if (SPLIT) tma_load_3d(b_lo, &maps.b[1], &full_bar[s], kb * BK, bn, bz);
        }
        __syncwarp();
        if (++ch == cpt) { ch = 0; ++tap; }
        if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-convergent, one elected lane issues =====
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    const uint32_t ring_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
    constexpr uint32_t kStageLo = Cfg::kStageBytes >> 4, kALo = Cfg::kABytes >> 4, kBLo = Cfg::kBBytes >> 4;
    uint32_t s = 0, ph = 0;
    int lt = 0;  // tiles processed by this CTA
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const TileCoord tc = decode_tile(p, t);
      const int as = lt & 1;
      mbar_wait(&tmem_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(as * BN);
      for (int i = 0; i < tc.nkb; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = ring_lo + s * kStageLo, a_lo = a_hi + kALo;
          const uint32_t b_hi = a_hi + (SPLIT ? 2 : 1) * kALo, b_lo = b_hi + kBLo;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, (i | k) != 0);
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(d, umma_desc_from_lo<BK>(a_lo + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_lo + 2 * k), idesc, 1u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[as]);  // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4).. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile
    const int xl = r % p.TW, yl = (r / p.TW) % p.TH, bl = r / (p.TW * p.TH);
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
      const TileCoord tc = decode_tile(p, t);
      const int as = lt & 1;
      const int ox = tc.ox0 + xl, oy = tc.oy0 + yl, b = tc.b0 + bl;
      const bool valid = (bl < p.TB) && ox < p.Wo && oy < p.Ho && b < p.batch;
      mbar_wait_relaxed(&tmem_full_bar[as], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      const TcRow row{b, oy, ox, tc.phase, tc.split};
#pragma unroll 1
      for (int c = 0; c < BN / CH; ++c) {
        uint32_t v[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores of the previous chunk
        if (CH == 32) tmem_ld32(taddr + (uint32_t)(c * CH), v);
        else tmem_ld16(taddr + (uint32_t)(c * CH), v);
        tmem_ld_wait();
        if (c == BN / CH - 1) {  // accumulator is in registers: hand the TMEM buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
        const int n0 = tc.n_tile * BN + c * CH;
        const int nv = min(CH, p.N - n0);  // valid channels of this chunk
        if (nv > 0) tc_epilogue_store<CH>(p.epi, ev, row, v, n0, nv, valid, p.TW);
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
static int launch_one(const TcMaps& maps, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN, BK, SPLIT>;
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, BK, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.done();
  }
  // persistent: at most one wave of CTAs, each walking the tile list with stride gridDim.x
  const int slots = sm_count() * Cfg::kMinCtas;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  tc_gemm_kernel<BN, BK, SPLIT><<<grid, 192, Cfg::kSmemBytes, st>>>(maps, p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("tc_gemm_kernel");
  g_last_kernel_kind = KK_TC_GEMM;
  return CIC_OK;
}

template <int BK, bool SPLIT>
static int launch_bn(const TcMaps& maps, const TcParams& p, int block_n, cudaStream_t st) {
  switch (block_n) {
    case 16: return launch_one<16, BK, SPLIT>(maps, p, st);
    case 32: return launch_one<32, BK, SPLIT>(maps, p, st);
    case 64: return launch_one<64, BK, SPLIT>(maps, p, st);
    case 128: return launch_one<128, BK, SPLIT>(maps, p, st);
    case 256:
      if (!SPLIT && BK == 64) return launch_one<256, 64, false>(maps, p, st);
      break;
  }
  set_error("tc_gemm: unsupported N tile %d (K block %d, split=%d)", block_n, BK, (int)SPLIT);
  return CIC_ERR_INVALID;
}

int launch_tc_gemm(const TcMaps& maps, TcParams& p, int block_n, int block_k, bool split, cudaStream_t st) {
  CIC_REQUIRE(block_n > 0 && p.N_pad % block_n == 0 && p.N <= p.N_pad, "tc_gemm: padded N=%d is not a multiple of the N tile %d", p.N_pad, block_n);
  CIC_REQUIRE(p.kblocks >= p.splits && p.splits >= 1, "tc_gemm: bad split-K %d for %d K blocks", p.splits, p.kblocks);
  const int per = (p.kblocks + p.splits - 1) / p.splits;
  CIC_REQUIRE((p.splits - 1) * per < p.kblocks, "tc_gemm: split-K leaves an empty split");
  CIC_REQUIRE(p.TW * p.TH * p.TB <= TC_BM && p.TW >= 1, "tc_gemm: bad M tile");
  const long long mt = (long long)p.tiles_x * p.tiles_y * p.tiles_b;
  CIC_REQUIRE(mt > 0 && mt < 2147483647LL, "tc_gemm: bad tile count");
  p.n_tiles = (p.N + block_n - 1) / block_n;
  const long long total = mt * p.n_tiles * p.nphases * p.splits;
  CIC_REQUIRE(total < 2147483647LL, "tc_gemm: too many tiles");
  p.total_tiles = (int)total;
  if (block_k == 64) return split ? launch_bn<64, true>(maps, p, block_n, st) : launch_bn<64, false>(maps, p, block_n, st);
  if (block_k == 32) return split ? launch_bn<32, true>(maps, p, block_n, st) : launch_bn<32, false>(maps, p, block_n, st);
  set_error("tc_gemm: unsupported K block %d", block_k);
  return CIC_ERR_INVALID;
}

}  // namespace cic
