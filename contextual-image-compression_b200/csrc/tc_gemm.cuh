// tcgen05 / TMEM / TMA implicit-GEMM for sm_100a: the CIC_PREC_TC arithmetic of the conv, transposed
// conv, dense and attention matmuls.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x BK] (bf16, smem, K-major, swizzled) * B[BN x BK]^T (same)
//
// A is the im2col view of an NHWC bf16 activation tensor: one TMA box {BK channels, TW, TH, TB}
// per (tap, BK-channel chunk) lands as 128 rows x (2*BK) bytes, which *is* the canonical K-major
// swizzled UMMA operand layout (SWIZZLE_128B for BK = 64, SWIZZLE_64B for BK = 32, the latter for
// the 32-channel layers), and TMA's out-of-bounds zero fill implements TF 'same' padding.  Stride-2
// convs read through a 5-D view (x-parity folded into the channel dim, y-parity as its own dim) so
// no traversal strides are needed; transposed 4x4/stride-2 convs run as four 2x2 phases selected by
// blockIdx.z.  B is the layer's weight matrix pre-packed to bf16 [N][K] (K ordered tap-major).
// In SPLIT mode A and B are (hi, lo) bf16 pairs and each K block issues three MMAs
// (hi*hi + lo*hi + hi*lo): error-compensated bf16 that reproduces fp32 products to ~2^-17 relative.
//
// Persistent kernel: the grid is one (or two) CTAs per SM and every CTA walks the tile list
// t = blockIdx.x, blockIdx.x + gridDim.x, ...  Warp roles (192 threads): warp 0 = TMA producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2-5 = epilogue (tcgen05.ld -> bias / folded BN /
// activation -> global).  A ring of kStages smem slots with full/empty mbarriers decouples TMA from the
// tensor pipe and runs ahead across tile boundaries; the accumulator is double-buffered in TMEM
// (tmem_full / tmem_empty mbarriers) so the MMAs of tile i+1 overlap the epilogue of tile i.
#pragma once
#include "common.cuh"
#include "tc_epilogue.cuh"

#include <cuda.h>

namespace cic {

constexpr int TC_BM = 128;
constexpr int TC_MAX_TAPS = 16;

struct TcTap {  // TMA coordinate deltas of one filter tap
  int16_t dc;   // added to the channel coordinate (x-parity * ld for the stride-2 view)
  int16_t dx, dy;
  int16_t pz;   // y-parity coordinate of the stride-2 view
};

struct TcParams {
  // geometry of the M tile: rows = (tb, th, tw)
  int TW, TH, TB;
  int tiles_x, tiles_y, tiles_b;
  int Wo, Ho, batch;      // output positions per batch item and batch size (row validity)
  int a5d;                // 1: A maps are the 5-D stride-2 view
  int nsrc;               // channel-concatenated sources
  int src_blocks[2];      // BK-channel blocks per tap from each source
  int src_coff[2];        // channel offset inside each source's pixel record
  int ntaps;
  TcTap taps[4][TC_MAX_TAPS];  // [phase][tap]
  int nphases;            // 4 for transposed conv
  int splits;             // split-K factor (blockIdx.z = phase * splits + split)
  int kblocks;            // total K blocks = ntaps * (src_blocks[0] + src_blocks[1])
  int N, N_pad;           // real / padded output channels (B rows per phase = N_pad)
  int b_batched;          // 1: B has one matrix per batch item (attention); tile never spans items
  int n_tiles;            // N tiles per (M tile, phase, split)
  int total_tiles;        // M tiles x n_tiles x nphases x splits
  int m_fast;             // tile order: 1 = M tiles fastest (weight-dominated GEMMs), 0 = N tiles fastest
  TcEpi epi;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug traps (CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (it == 1024) t0 = clock64();
    if (it > 1024 && (it & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();  // ~2 s
  }
}

// One lane of a fully converged warp (the same lane every time).  The producer and MMA warps run their loops
// warp-convergent and elect a lane only around the asynchronous instruction issue: inside a divergent
// `if (lane == 0)` region the compiler wraps every uniform-datapath instruction (tcgen05.*, TMA, mbarrier) in
// vote / elect / branch sequences, which made the issuing thread itself the bottleneck (ncu source view,
// profiles/r01_ncu_raster_issue_bound.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Same wait for warps that are not on the critical path (epilogue, producers): the try_wait carries a
// suspend-time hint so the warp sleeps in hardware instead of hot-polling the scheduler it shares with the
// MMA-issuing warp.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if (it == 64) t0 = clock64();
    if (it > 64 && (it & 63) == 0 && clock64() - t0 > 4000000000LL) __trap();  // ~2 s
  }
}

// 16-byte shared-memory load by 32-bit shared address (LDS.128; tables that are constant after the prologue barrier)
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32, single CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major swizzled shared-memory operand descriptor (sm_100 format).  One swizzle atom along K
// (BK bf16 = 2*BK bytes per row), 8-row groups 8 * 2*BK bytes apart (SBO); LBO unused.  The start address
// may be any row (multiple of 2*BK bytes), not only a multiple of the swizzle repeat: the hardware swizzles
// on absolute shared-memory address bits, exactly like TMA does when it writes the tile (measured on B200,
// profiles/r01_shift_probe.log), so base_offset stays 0.
//   BK = 64: SWIZZLE_128B (layout type 2), SBO 1024;  BK = 32: SWIZZLE_64B (layout type 4), SBO 512.
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr);
// the same descriptor from a pre-shifted start-address field: desc = hi(const) | lo, lo = (addr & 0x3FFFF) >> 4
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo) {
  constexpr uint32_t hi = (uint32_t)((8 * 2 * BK) >> 4) | (1u << 14) | ((BK == 64 ? 2u : 4u) << 29);
  return ((uint64_t)hi << 32) | (uint64_t)lo;
}
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  constexpr uint64_t sbo = 8 * 2 * BK;
  constexpr uint64_t layout = BK == 64 ? 2 : 4;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)(sbo >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= layout << 61;                            // swizzle mode
  return d;
}

// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN, int BK, bool SPLIT>
struct TcCfg {
  static constexpr int kABytes = TC_BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kABytes + kBBytes);
  static constexpr int kBudget = SPLIT || BN >= 256 ? 196608 : 98304;  // 1 or 2 CTAs per SM
  static constexpr int kStages = (kBudget / kStageBytes) < 2 ? 2 : ((kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;          // double-buffered accumulator (32 .. 512 columns)
  static constexpr int kChunk = BN < 32 ? 16 : 32;  // accumulator columns per tcgen05.ld
  static constexpr int kMinCtas = kSmemBytes <= 112 * 1024 ? 2 : 1;
  static_assert(kMinCtas * kTmemCols <= 512, "TMEM over-subscribed");
};

struct TcMaps {
  CUtensorMap a[2][2];  // [source][hi, lo]
  CUtensorMap b[2];     // [hi, lo]
};

int launch_tc_gemm(const TcMaps& maps, TcParams& p, int block_n, int block_k, bool split, cudaStream_t st);
// CTA-pair (cta_group::2) variant, tc_gemm2.cu; maps.b encoded with box {BK, block_n / 2}
int launch_tc_gemm2(const TcMaps& maps, TcParams& p, int block_n, int block_k, bool split, cudaStream_t st);
int tc2_pick_block_n(int n_pad);
// merged-phase CTA-pair kernel for Conv2DTranspose(k4, s2) with Cout in {32, 64} (tc_gemm2.cu); maps.b box {64, Cout / 2}
bool tc_deconv2_ok(int bk, bool split, long long m_tiles, int n_pad, int n);
int launch_tc_deconv2(const TcMaps& maps, TcParams& p, cudaStream_t st);

}  // namespace cic
