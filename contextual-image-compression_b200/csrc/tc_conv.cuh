// "Raster" convolution kernel for sm_100a: implicit GEMM whose filter taps are served from ONE halo tile in
// shared memory instead of one TMA im2col load per tap.
//
// tc_gemm.cuh loads a separate shifted copy of the activation tile for every tap, so a 3x3 conv moves each
// input element L2 -> SMEM nine times (sixteen for the four phases of a transposed 4x4/s2 conv) and the
// small-Cout layers end up L2-bandwidth-bound (ncu: profiles/r01_ncu_tc_v1_nonpersistent.md).  Here one TMA
// box {BK channels, rw, rh, rb} - a raster of pixels including the halo - lands in shared memory as rows of
// 2*BK bytes in the canonical swizzled K-major layout, and every tap is an MMA whose A descriptor simply starts
// `row_shift` rows into that raster (any row offset is legal: the swizzle is a function of absolute address
// bits; measured, see tc_gemm.cuh).  Two raster shapes cover the layers:
//   ROW (tile = TW pixels of one image row): one raster {TW + kw - 1, kh}; tap (ky, kx) starts at ky*rw + kx;
//   COL (tile = TW x TH patch):              one raster {TW, TH + kh - 1} per kx; tap ky starts at ky*TW;
// tiles that span batch items fall back to one raster per tap.  The four output phases of a transposed conv
// share the rasters and accumulate into separate TMEM column ranges in the same pass (Cout <= 64), so the
// input is read once for all of them.  Stride-2 convs use the 5-D parity view of tc_gemm.cuh, one raster set
// per parity plane.
//
// Pipeline: persistent CTAs (one per SM), 224 threads: warp 0 = raster (A) producer, warp 6 = weight (B)
// producer, warp 1 = MMA issuer, warps 2-5 = epilogue.  A rasters and B blocks travel in separate mbarrier
// rings filled by separate warps, because one raster feeds many MMAs and the raster of the next tile must be
// in flight long before the last weight block of this one is issued.  When the whole weight matrix of the
// layer fits beside the raster ring it is loaded once and stays resident (no per-MMA weight traffic at all).
// The accumulator (up to 256 columns per stage) is double-buffered in TMEM.
#pragma once
#include "tc_gemm.cuh"

namespace cic {

constexpr int TCV_MAX_RASTERS = 16;
constexpr int TCV_MAX_OPS = 16;
constexpr int TCV_MAX_PASS = 4;
constexpr int TCV_MAX_SLOTS = 16;
constexpr int TCV_MAX_ISSUERS = 4;
constexpr int TCV_MAX_THREADS = 448;  // warps: 0 rasters, 1 issuer 0, 2-5 epilogue 0, 6 weights, 7 issuer 1, 8-11 epilogue 1, 12-13 issuers 2-3
constexpr int TCV_ACC_COLS = 256;  // TMEM columns per accumulator stage

struct TcvRaster {
  int16_t dc;        // channel-coordinate delta (x-parity * ld of the stride-2 view)
  int16_t dx, dy;    // pixel-coordinate deltas of the raster origin relative to the tile origin
  int16_t pz;        // y-parity plane of the stride-2 view
  uint8_t op0, nops; // its MMA ops: op[op0 .. op0 + nops)
  uint8_t pad_[2];
};

struct TcvOp {
  uint16_t row_shift;  // A window starts this many rows (pixels) into the raster
  uint8_t acc;         // accumulator (local phase) index of the pass (the first one of a merged op)
  uint8_t tap;         // tap index in the weight K order: K block = (tap * cpt + channel block)
  uint8_t first;       // first op of its accumulator(s) in the pass (overwrite at channel block 0)
  uint8_t nph;         // merged op: accumulators acc .. acc + nph - 1 share this A window (one MMA of N = nph * BN); else 1
  uint8_t blk0;        // merged op: first weight block of the op within a channel block's group of resident blocks
  uint8_t pad_;
  uint8_t taps[4];     // merged op: tap index of each merged accumulator
};

// flags of the MMA op table the kernel builds in shared memory
enum { TCV_F_FIRST = 1, TCV_F_NEW_RASTER = 2, TCV_F_LAST_OF_RASTER = 4, TCV_F_NEW_BGROUP = 8, TCV_F_LAST_OF_BGROUP = 16, TCV_F_MINE = 32 };

struct TcvPass {
  int nrast;
  int nops;
  int nacc;               // accumulators (output phases) of this pass
  int8_t phase_id[4];     // global phase of each accumulator (weight row block, output offset)
  TcvRaster r[TCV_MAX_RASTERS];
  TcvOp op[TCV_MAX_OPS];
};

// division by a runtime constant as multiply-high + shift (the tile decode runs once per tile in three warps;
// six hardware divisions there cost ~1000 cycles of dependent issue)
struct FastDiv {
  uint32_t d, mul, shr;
#ifdef __CUDACC__
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = d == 1 ? n : (__umulhi(n, mul) >> shr);
    r = n - q * d;
  }
#endif
};
inline FastDiv make_fastdiv(uint32_t d) {  // exact for n < 2^31
  FastDiv f{d, 0, 0};
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.shr = l - 1;
  f.mul = (uint32_t)(((1ull << (32 + l - 1)) + d - 1) / d);
  return f;
}

struct TcvParams {
  int TW, TH, TB;
  int tiles_x, tiles_y, tiles_b;
  FastDiv fd_ntiles, fd_npass, fd_tx, fd_ty;
  int Wo, Ho, batch;
  int a5d;
  int nsrc;
  int src_blocks[2];
  int src_coff[2];
  int rw, rh, rb;         // raster box in pixels (the same for every raster of the layer)
  int npass;
  int N_pad;              // weight rows per phase
  int BN;                 // accumulator width = MMA N (multiple of 16, <= 256, BN * nacc <= 256)
  int n_tiles;            // N tiles per phase (N_pad / BN)
  int total_tiles;        // M tiles x npass x n_tiles
  int ntaps;              // taps per phase in the weight K order
  int b_resident;         // 1: all weight blocks live in shared memory for the whole kernel
  int b_blocks;           // resident blocks: nphases x n_tiles x ntaps x channel blocks
  int a_slots, b_slots;
  int a_slot_bytes;       // per slot; in split mode [hi raster | lo raster], each a_slot_bytes / 2
  int b_slot_bytes;       // one weight block [BN x BK]; in split mode [hi | lo]
  int b_group;            // weight blocks per ring slot (loaded under one mbarrier); ring slot = b_group * b_slot_bytes
  uint32_t a_tx_bytes, b_tx_bytes;
  int merged;             // 1: transposed-conv ops are merged per input shift (resident weights, see tc_host.cu)
  int nblk_cb;            // merged: resident weight blocks per channel block (16)
  int nw;                 // MMA issuer warps (accumulators a % nw == w belong to issuer w); > 1 only with resident weights
  int ne;                 // epilogue groups of four warps (1 or 2)
  int dbg;                // CIC_TC_DBG elimination bits: 1 no MMA issue, 2 no weight loads, 4 no raster loads, 8 no stores
  TcEpi epi;
  TcvPass pass[TCV_MAX_PASS];
};

size_t tcv_smem_bytes(const TcvParams& p);
int launch_tc_conv(const TcMaps& maps, const TcvParams& p, int block_k, bool split, cudaStream_t st);

}  // namespace cic
