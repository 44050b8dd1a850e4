// CTA-pair (cta_group::2) variant of the per-tap implicit-GEMM kernel of tc_gemm.cuh.
//
// Why: measured on B200 (profiles/r01_smem_pipe_model.md) the one-CTA kernel is bound by the SM's 128 B/clk
// shared-memory data pipe, which carries BOTH the TMA fills of the stage ring and the tensor core's operand reads:
//   cycles per K block  ~  (fill bytes + MMAs * (A 4 KB + B N*32 B)) / 128
// e.g. split-bf16 conv3 (N tile 128): 512 + 768 wavefronts = 1280 clk against 768 clk of tensor math.  With a CTA pair
// one tcgen05.mma.cta_group::2 computes D[256 x BN]: each CTA supplies its own 128 A rows and HALF of the B tile, so per SM
// the B fill and the B operand reads halve and the N tile can be 256 wide in split mode as well.
//
// Structure (per CTA: 192 or 320 threads): warp 0 = TMA producer (both CTAs: own A tile + own half of B; every
// complete_tx goes to the LEADER's full barrier), warp 1 = TMEM allocator (both CTAs, tcgen05.alloc.cta_group::2) and, in
// the leader (cluster rank 0) only, the MMA issuer; its tcgen05.commit multicasts the arrive to both CTAs' empty /
// tmem_full barriers.  Warps 2-5 (+ 6-9 with a second epilogue group) drain the CTA's own TMEM half and arrive on the
// leader's tmem_empty barrier (remote arrive from the peer).  M tiles are taken in pairs (2m, 2m + 1); an odd
// tile count leaves a phantom tile whose loads are out of bounds (zero fill) and whose rows are masked by the epilogue.
#include "tc_gemm.cuh"
#include "tc_host.cuh"

#include <cstring>

namespace cic {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in both CTAs of the pair once all prior MMAs of this thread have retired
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// arrive on the leader CTA's barrier at the same offset as `bar` (local arrive in the leader itself)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

// D fp32, A/B bf16 K-major, M = 256 over the CTA pair
__host__ __device__ constexpr uint32_t umma2_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int BK, bool SPLIT>
struct Tc2Cfg {
  static constexpr int kABytes = TC_BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;  // this CTA's half of the B tile
  static constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kABytes + kBBytes);
  static constexpr int kBudget = 200 * 1024;
  static constexpr int kStages = (kBudget / kStageBytes) < 2 ? 2 : ((kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // double-buffered accumulator
  static constexpr int kChunk = BN < 32 ? 16 : 32;
  static_assert(kTmemCols <= 512, "TMEM over-subscribed");
};

struct PairTile {
  int ox0, oy0, b0;  // this CTA's M tile
  int n_tile, phase, split;
  int kb0, nkb;      // K-block range of this split
};

// pair index t -> (M-tile pair, N tile, phase); same orders as decode_tile of the one-CTA kernel
__device__ __forceinline__ PairTile decode_pair(const TcParams& p, int t, int rank) {
  const int mt = p.tiles_x * p.tiles_y * p.tiles_b;
  const int mp = (mt + 1) >> 1;
  int m2, n, z;
  if (p.m_fast) {
    m2 = t % mp;
    const int r = t / mp;
    n = r % p.n_tiles;
    z = r / p.n_tiles;
  } else {
    n = t % p.n_tiles;
    const int r = t / p.n_tiles;
    const int zt = p.nphases * p.splits;
    z = r % zt;
    m2 = r / zt;
  }
  const int m = 2 * m2 + rank;  // may be == mt (phantom tile of an odd count): b0 >= batch, everything out of bounds
  PairTile c;
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

template <int BN, int BK, bool SPLIT, int THREADS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_gemm2_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p, const int total_pairs) {
  using Cfg = Tc2Cfg<BN, BK, SPLIT>;
  constexpr int kStages = Cfg::kStages;
  constexpr int CH = Cfg::kChunk;
  constexpr int NE = THREADS > 192 ? 2 : 1;  // epilogue groups of four warps
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2] (the leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const TcEpiVec ev{p.epi.bias, p.epi.scale, p.epi.shift};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, npairs_grid = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.b[0]);
    if (SPLIT) { prefetch_tmap(&maps.a[0][1]); prefetch_tmap(&maps.b[1]); }
    if (p.nsrc > 1) { prefetch_tmap(&maps.a[1][0]); if (SPLIT) prefetch_tmap(&maps.a[1][1]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * 4 * NE); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and the TMEM allocation of both CTAs are visible before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A tile, own half of the B tile; all bytes are counted on the leader's barrier =====
    if (elect_one()) {
      const int cpt = p.src_blocks[0] + p.src_blocks[1];
      const uint32_t rows = (uint32_t)(p.TW * p.TH * p.TB);
      const uint32_t tx_bytes = 2u * (SPLIT ? 2u : 1u) * (rows + (uint32_t)(BN / 2)) * (uint32_t)(2 * BK);  // both CTAs
      uint32_t s = 0, ph = 0;
      for (int t = pair0; t < total_pairs; t += npairs_grid) {
        const PairTile tc = decode_pair(p, t, (int)rank);
        const int bn = tc.phase * p.N_pad + tc.n_tile * BN + (int)rank * (BN / 2);
        int tap = tc.kb0 / cpt, ch = tc.kb0 % cpt;
        for (int i = 0; i < tc.nkb; ++i) {
          mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          const uint32_t bar = smem_u32(&full_bar[s]) & kPeerBitMask;
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
            tma2_load_5d(a_hi, &maps.a[src][0], bar, c, tc.ox0 + tp.dx, tp.pz, tc.oy0 + tp.dy, tc.b0);
            if (SPLIT) tma2_load_5d(a_lo, &maps.a[src][1], bar, c, tc.ox0 + tp.dx, tp.pz, tc.oy0 + tp.dy, tc.b0);
          } else {
            tma2_load_4d(a_hi, &maps.a[src][0], bar, c, tc.ox0 + tp.dx, tc.oy0 + tp.dy, tc.b0);
            if (SPLIT) tma2_load_4d(a_lo, &maps.a[src][1], bar, c, tc.ox0 + tp.dx, tc.oy0 + tp.dy, tc.b0);
          }
          tma2_load_3d(b_hi, &maps.b[0], bar, (tc.kb0 + i) * BK, bn, 0);
          if (SPLIT) tma2_load_3d(b_lo, &maps.b[1], bar, (tc.kb0 + i) * BK, bn, 0);
          if (++ch == cpt) { ch = 0; ++tap; }
          if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA's elected lane issues for the pair =====
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma2_idesc_bf16(BN);
      const uint32_t ring_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
      constexpr uint32_t kStageLo = Cfg::kStageBytes >> 4, kALo = Cfg::kABytes >> 4, kBLo = Cfg::kBBytes >> 4;
      uint32_t s = 0, ph = 0;
      int lt = 0;
      for (int t = pair0; t < total_pairs; t += npairs_grid, ++lt) {
        const int nkb = decode_pair(p, t, 0).nkb;
        const int as = lt & 1;
        mbar_wait(&tmem_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);  // both CTAs' epilogues have drained this stage
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(as * BN);
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_hi = ring_lo + s * kStageLo, a_lo = a_hi + kALo;
          const uint32_t b_hi = a_hi + (SPLIT ? 2 : 1) * kALo, b_lo = b_hi + kBLo;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma2_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, (i | k) != 0);
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma2_bf16(d, umma_desc_from_lo<BK>(a_lo + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma2_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_lo + 2 * k), idesc, 1u);
          }
          umma2_commit_mc(&empty_bar[s]);  // frees the stage in both CTAs when these MMAs retire
          if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
        }
        umma2_commit_mc(&tmem_full_bar[as]);  // accumulator complete, in both CTAs
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: group eg of four warps drains every NE-th chunk of this CTA's 128 accumulator rows =====
    const int eg = warp >= 6 ? 1 : 0;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int xl = r % p.TW, yl = (r / p.TW) % p.TH, bl = r / (p.TW * p.TH);
    constexpr int kChunks = BN / CH;
    int lt = 0;
    for (int t = pair0; t < total_pairs; t += npairs_grid, ++lt) {
      const PairTile tc = decode_pair(p, t, (int)rank);
      const int as = lt & 1;
      const int ox = tc.ox0 + xl, oy = tc.oy0 + yl, b = tc.b0 + bl;
      const bool valid = (bl < p.TB) && ox < p.Wo && oy < p.Ho && b < p.batch;
      mbar_wait_relaxed(&tmem_full_bar[as], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      const TcRow row{b, oy, ox, tc.phase, tc.split};
      constexpr int kLastMine0 = kChunks - 1 - ((kChunks - 1) % NE);  // last chunk of group 0
      const int last_mine = eg == 0 ? kLastMine0 : (kChunks - 1 - ((kChunks - 1 - 1 + NE) % NE));
      if (eg >= kChunks) {  // more groups than chunks: nothing to drain
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[as]);
      }
#pragma unroll 1
      for (int c = eg; c < kChunks; c += NE) {
        uint32_t v[32];
        __syncwarp();
        if (CH == 32) tmem_ld32(taddr + (uint32_t)(c * CH), v);
        else tmem_ld16(taddr + (uint32_t)(c * CH), v);
        tmem_ld_wait();
        if (c == last_mine) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[as]);
        }
        const int n0 = tc.n_tile * BN + c * CH;
        const int nv = min(CH, p.N - n0);
        if (nv > 0) tc_epilogue_store<CH>(p.epi, ev, row, v, n0, nv, valid, p.TW);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory and barriers stay alive until every MMA / remote arrive has landed
  if (warp == 1) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Conv2DTranspose(k4, s2) with the four output phases merged per input shift, on CTA pairs.
//
// The 16 (phase, tap) products of the layer read only 9 distinct shifted input windows.  With the accumulators ordered
// (py, px) = (0,0) (0,1) (1,1) (1,0) the phases sharing a shift are adjacent columns of one D[256 x 4*Cout] accumulator, so
// one K block = (shift, 64 channels) is ONE A tile and one MMA group of N = nph * Cout (nph = 4 for the centre shift, 2 for
// three of the edge shifts, 1 otherwise: 10 ops).  Per K block each SM then moves A once per shift instead of once per
// (phase, tap) - 9 instead of 16 fills and operand read streams - and, as in tc_gemm2_kernel, only half of the B tile
// (model of profiles/r01_smem_pipe_model.md, deconv3: 16 x 384 -> 3328 clk per channel block and 128 pixels).
// Requires N_pad == Cout in {32, 64} (4 * Cout <= 256 TMEM columns per stage), K block 64, no split operands.
struct Dc2Op {
  int8_t dy, dx;
  uint8_t acc0, nph;
  uint8_t tap[4];  // weight tap index (of its phase) of accumulator acc0 + j
};
struct Dc2Params {
  int nops;
  Dc2Op op[10];
  int8_t acc_phase[4];
};

template <int COUT, int THREADS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_deconv2_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p, const __grid_constant__ Dc2Params q,
                  const int total_pairs) {
  constexpr int BK = 64;
  constexpr int BNT = 4 * COUT;                    // accumulator columns of one tile (all four phases)
  constexpr int kABytes = TC_BM * BK * 2;          // 16 KB
  constexpr int kHalfBytes = (COUT / 2) * BK * 2;  // one half phase block of B: Cout/2 rows
  constexpr int kBBytes = 4 * kHalfBytes;          // this CTA's half of the widest op (nph = 4)
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  constexpr int NE = THREADS > 192 ? 2 : 1;
  constexpr int kChunks = BNT / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  // per-op words fetched by the two single-thread roles (a thread sustains ~5 clk per dependent instruction: the first version,
  // which decoded q.op[o] in the loop, spent ~700 clk per K block in the producer - profiles/r01_epilogue_data_pipe.md):
  //   tab_a[o] = {dx, dy, nph, expect_tx bytes of the pair}; tab_b[o][i] = {K offset of the tap, B row} of this CTA's i-th half block;
  //   tab_m[o] = {instruction descriptor, accumulator column offset}
  uint4* tab_a = reinterpret_cast<uint4*>(smem + kStages * kStageBytes + 256);
  uint2* tab_b = reinterpret_cast<uint2*>(tab_a + 10);
  uint2* tab_m = tab_b + 40;
  const TcEpiVec ev{p.epi.bias, p.epi.scale, p.epi.shift};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, npairs_grid = gridDim.x >> 1;
  const int cpt = p.src_blocks[0] + p.src_blocks[1];
  const int mt = p.tiles_x * p.tiles_y * p.tiles_b;
  if (threadIdx.x < q.nops) {
    const int o = threadIdx.x;
    const Dc2Op op = q.op[o];
    tab_a[o] = make_uint4((uint32_t)(int)op.dx, (uint32_t)(int)op.dy, op.nph, 2u * (uint32_t)(kABytes + op.nph * kHalfBytes));
    for (int i = 0; i < op.nph; ++i) {
      const int h = (int)rank * op.nph + i, j = h >> 1, half = h & 1;
      tab_b[o * 4 + i] = make_uint2((uint32_t)(op.tap[j] * cpt * BK), (uint32_t)(q.acc_phase[op.acc0 + j] * p.N_pad + half * (COUT / 2)));
    }
    tab_m[o] = make_uint2(umma2_idesc_bf16(op.nph * COUT), (uint32_t)(op.acc0 * COUT));
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.b[0]);
    if (p.nsrc > 1) prefetch_tmap(&maps.a[1][0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * 4 * NE); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_slot, 2 * BNT);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A window of the shift, own half of the op's B rows =====
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      const uint32_t ta = smem_u32(tab_a), tb = smem_u32(tab_b);
      const int nops = q.nops;
      for (int t = pair0; t < total_pairs; t += npairs_grid) {
        const int m = 2 * t + (int)rank;  // == mt for the phantom tile of an odd count: everything out of bounds
        const int ox0 = (m % p.tiles_x) * p.TW, oy0 = ((m / p.tiles_x) % p.tiles_y) * p.TH, b0 = (m / (p.tiles_x * p.tiles_y)) * p.TB;
        for (int cb = 0; cb < cpt; ++cb) {
          const int src = cb >= p.src_blocks[0] ? 1 : 0;
          const int c = p.src_coff[src] + (src ? cb - p.src_blocks[0] : cb) * BK;
          const int kcb = cb * BK;
          uint4 a = lds_v4(ta);
          for (int o = 0; o < nops; ++o) {
            const uint4 nx = lds_v4(ta + (uint32_t)(o + 1 < nops ? o + 1 : 0) * 16u);  // one op ahead
            mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], a.w);
            const uint32_t bar = smem_u32(&full_bar[s]) & kPeerBitMask;
            uint8_t* st = smem + s * kStageBytes;
            tma2_load_4d(st, &maps.a[src][0], bar, c, ox0 + (int)a.x, oy0 + (int)a.y, b0);
            for (uint32_t i = 0; i < a.z; ++i) {
              const uint2 bw = lds_v2(tb + ((uint32_t)o * 4u + i) * 8u);
              tma2_load_3d(st + kABytes + i * kHalfBytes, &maps.b[0], bar, (int)bw.x + kcb, (int)bw.y, 0);
            }
            if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
            a = nx;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (rank == 0 && elect_one()) {
      const uint32_t ring_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
      constexpr uint32_t kStageLo = kStageBytes >> 4, kALo = kABytes >> 4;
      uint32_t s = 0, ph = 0;
      int lt = 0;
      const uint32_t tm = smem_u32(tab_m);
      const int nops = q.nops;
      for (int t = pair0; t < total_pairs; t += npairs_grid, ++lt) {
        const int as = lt & 1;
        mbar_wait(&tmem_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(as * BNT);
        for (int cb = 0; cb < cpt; ++cb) {
          uint2 mw = lds_v2(tm);
          for (int o = 0; o < nops; ++o) {
            const uint2 nx = lds_v2(tm + (uint32_t)(o + 1 < nops ? o + 1 : 0) * 8u);
            const uint32_t idesc = mw.x;
            const uint32_t d = d0 + mw.y;
            mw = nx;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_lo = ring_lo + s * kStageLo, b_lo = a_lo + kALo;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma2_bf16(d, umma_desc_from_lo<BK>(a_lo + 2 * k), umma_desc_from_lo<BK>(b_lo + 2 * k), idesc, (cb | o | k) != 0);
            umma2_commit_mc(&empty_bar[s]);
            if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; }
          }
        }
        umma2_commit_mc(&tmem_full_bar[as]);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: this CTA's 128 rows; chunk c = 32 columns of accumulator c * 32 / COUT =====
    const int eg = warp >= 6 ? 1 : 0;
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const int xl = r % p.TW, yl = (r / p.TW) % p.TH, bl = r / (p.TW * p.TH);
    int lt = 0;
    for (int t = pair0; t < total_pairs; t += npairs_grid, ++lt) {
      const int m = 2 * t + (int)rank;
      const int ox = (m % p.tiles_x) * p.TW + xl, oy = ((m / p.tiles_x) % p.tiles_y) * p.TH + yl, b = (m / (p.tiles_x * p.tiles_y)) * p.TB + bl;
      const bool valid = m < mt && (bl < p.TB) && ox < p.Wo && oy < p.Ho && b < p.batch;
      const int as = lt & 1;
      mbar_wait_relaxed(&tmem_full_bar[as], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BNT);
      constexpr int kLast0 = kChunks - 1 - ((kChunks - 1) % NE);
      const int last_mine = eg == 0 ? kLast0 : (kChunks - 1 - ((kChunks - 1 - 1 + NE) % NE));
#pragma unroll 1
      for (int c = eg; c < kChunks; c += NE) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (c == last_mine) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[as]);
        }
        const int acc = (c * 32) / COUT, n0 = (c * 32) % COUT;
        const TcRow row{b, oy, ox, q.acc_phase[acc], 0};
        tc_epilogue_store<32>(p.epi, ev, row, v, n0, 32, valid, p.TW);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, 2 * BNT);
}

template <int COUT>
static int launch_dc2(const TcMaps& maps, const TcParams& p, const Dc2Params& q, int total_pairs, cudaStream_t st) {
  constexpr int kStageBytes = TC_BM * 64 * 2 + 4 * (COUT / 2) * 64 * 2;
  constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 1024 /*barriers + op tables*/;
  constexpr int THREADS = 320;  // two epilogue groups: 4 * Cout / 32 >= 4 chunks per tile
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(tc_deconv2_kernel<COUT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set.done();
  }
  const int pairs = sm_count() / 2;
  const int grid = 2 * (total_pairs < pairs ? total_pairs : pairs);
  tc_deconv2_kernel<COUT, THREADS><<<grid, THREADS, kSmemBytes, st>>>(maps, p, q, total_pairs);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("tc_deconv2_kernel");
  g_last_kernel_kind = KK_TC_GEMM2;
  return CIC_OK;
}

bool tc_deconv2_ok(int bk, bool split, long long m_tiles, int n_pad, int n) {
  return bk == 64 && !split && n == n_pad && (n == 32 || n == 64) && m_tiles >= 2 && (m_tiles % 2 == 0 || m_tiles >= 16);
}

// p as prepared for launch_tc_gemm for a transposed conv (taps[phase][tap], A maps with box {64, TW, TH, TB}); maps.b must have
// been encoded with box {64, Cout / 2}
int launch_tc_deconv2(const TcMaps& maps, TcParams& p, cudaStream_t st) {
  CIC_REQUIRE(p.nphases == 4 && p.ntaps == 4 && p.splits == 1 && !p.b_batched && !p.a5d, "tc_deconv2: not a transposed-conv layer");
  CIC_REQUIRE(p.N == p.N_pad && (p.N == 32 || p.N == 64), "tc_deconv2: Cout must be 32 or 64 (got %d, padded %d)", p.N, p.N_pad);
  Dc2Params q;
  memset(&q, 0, sizeof(q));
  static const int acc_phase[4] = {0, 1, 3, 2};
  struct MOp { int dy, dx, acc0, nph; };
  static const MOp mops[10] = {{0, 0, 0, 4},  {-1, 0, 0, 2},  {1, 0, 2, 2},  {0, 1, 1, 2}, {0, -1, 3, 1},
                               {0, -1, 0, 1}, {-1, -1, 0, 1}, {-1, 1, 1, 1}, {1, 1, 2, 1}, {1, -1, 3, 1}};
  for (int a = 0; a < 4; ++a) q.acc_phase[a] = (int8_t)acc_phase[a];
  q.nops = 10;
  for (int o = 0; o < 10; ++o) {
    Dc2Op& op = q.op[o];
    op.dy = (int8_t)mops[o].dy; op.dx = (int8_t)mops[o].dx; op.acc0 = (uint8_t)mops[o].acc0; op.nph = (uint8_t)mops[o].nph;
    for (int j = 0; j < mops[o].nph; ++j) {
      const int ph = acc_phase[mops[o].acc0 + j];
      int found = -1;
      for (int t = 0; t < 4; ++t)
        if (p.taps[ph][t].dy == mops[o].dy && p.taps[ph][t].dx == mops[o].dx) found = t;
      CIC_REQUIRE(found >= 0, "tc_deconv2: phase %d has no tap at shift (%d, %d)", ph, mops[o].dy, mops[o].dx);
      op.tap[j] = (uint8_t)found;
    }
  }
  const long long mt = (long long)p.tiles_x * p.tiles_y * p.tiles_b;
  const long long total = (mt + 1) / 2;
  CIC_REQUIRE(total > 0 && total < 2147483647LL, "tc_deconv2: bad tile count");
  p.n_tiles = 1;
  p.total_tiles = (int)total;
  if (p.N == 64) return launch_dc2<64>(maps, p, q, (int)total, st);
  return launch_dc2<32>(maps, p, q, (int)total, st);
}

template <int BN, int BK, bool SPLIT, int THREADS>
static int launch2_thr(const TcMaps& maps, const TcParams& p, int total_pairs, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN, BK, SPLIT>;
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<BN, BK, SPLIT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.done();
  }
  const int pairs = sm_count() / 2;
  const int grid = 2 * (total_pairs < pairs ? total_pairs : pairs);
  tc_gemm2_kernel<BN, BK, SPLIT, THREADS><<<grid, THREADS, Cfg::kSmemBytes, st>>>(maps, p, total_pairs);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("tc_gemm2_kernel");
  g_last_kernel_kind = KK_TC_GEMM2;
  return CIC_OK;
}

template <int BN, int BK, bool SPLIT>
static int launch2_one(const TcMaps& maps, const TcParams& p, int total_pairs, cudaStream_t st) {
  // a second epilogue group when the K loop is short: MMA time per tile ~ 2 * kblocks * terms * BN clk against ~75 * BN clk for
  // four warps to drain it (one warp per scheduler is latency-bound at ~2400 clk per 32-column chunk)
  const bool two = (BN / Tc2Cfg<BN, BK, SPLIT>::kChunk) >= 2 && (p.kblocks / p.splits) * (SPLIT ? 3 : 1) < 64;
  if (two) return launch2_thr<BN, BK, SPLIT, 320>(maps, p, total_pairs, st);
  return launch2_thr<BN, BK, SPLIT, 192>(maps, p, total_pairs, st);
}

// N tile of the pair kernel: the widest of 256 / 128 / 64 / 32 that divides the padded N
int tc2_pick_block_n(int n_pad) {
  for (int bn = 256; bn >= 32; bn /= 2)
    if (n_pad % bn == 0) return bn;
  return 0;
}

// p as prepared for launch_tc_gemm (taps, maps for A); maps.b must have been encoded with box {BK, block_n / 2}
int launch_tc_gemm2(const TcMaps& maps, TcParams& p, int block_n, int block_k, bool split, cudaStream_t st) {
  CIC_REQUIRE(block_k == 64, "tc_gemm2: K block must be 64");
  CIC_REQUIRE(!p.b_batched, "tc_gemm2: no batched B");
  CIC_REQUIRE(p.kblocks >= p.splits && p.splits >= 1, "tc_gemm2: bad split-K %d for %d K blocks", p.splits, p.kblocks);
  CIC_REQUIRE((p.splits - 1) * ((p.kblocks + p.splits - 1) / p.splits) < p.kblocks, "tc_gemm2: split-K leaves an empty split");
  CIC_REQUIRE(block_n > 0 && p.N_pad % block_n == 0 && p.N <= p.N_pad, "tc_gemm2: bad N tile %d", block_n);
  const long long mt = (long long)p.tiles_x * p.tiles_y * p.tiles_b;
  p.n_tiles = (p.N + block_n - 1) / block_n;
  const long long total = ((mt + 1) / 2) * p.n_tiles * p.nphases * p.splits;
  CIC_REQUIRE(total > 0 && total < 2147483647LL, "tc_gemm2: bad tile count");
  p.total_tiles = (int)total;
  switch (block_n) {
    case 32: return split ? launch2_one<32, 64, true>(maps, p, (int)total, st) : launch2_one<32, 64, false>(maps, p, (int)total, st);
    case 64: return split ? launch2_one<64, 64, true>(maps, p, (int)total, st) : launch2_one<64, 64, false>(maps, p, (int)total, st);
    case 128: return split ? launch2_one<128, 64, true>(maps, p, (int)total, st) : launch2_one<128, 64, false>(maps, p, (int)total, st);
    case 256: return split ? launch2_one<256, 64, true>(maps, p, (int)total, st) : launch2_one<256, 64, false>(maps, p, (int)total, st);
  }
  set_error("tc_gemm2: unsupported N tile %d", block_n);
  return CIC_ERR_INVALID;
}

}  // namespace cic
