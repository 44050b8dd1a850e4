// Raster convolution kernel (see tc_conv.cuh) + launch.
#include "tc_conv.cuh"

namespace cic {

struct TcvTile {
  int ox0, oy0, b0;
  int pass, n_tile;
};

__device__ __forceinline__ TcvTile tcv_decode(const TcvParams& p, int t) {
  // N tiles fastest, then passes, then M tiles: every (pass, N tile) of one M tile re-reads the same rasters
  // back to back, so those re-reads hit L2
  TcvTile c;
  uint32_t r, nt, m, ps, mx, my, q, mb;
  p.fd_ntiles.divmod((uint32_t)t, r, nt);
  p.fd_npass.divmod(r, m, ps);
  p.fd_tx.divmod(m, q, mx);
  p.fd_ty.divmod(q, mb, my);
  c.n_tile = (int)nt;
  c.pass = (int)ps;
  c.ox0 = (int)mx * p.TW;
  c.oy0 = (int)my * p.TH;
  c.b0 = (int)mb * p.TB;
  return c;
}

// instruction descriptor with a runtime N (D fp32, A/B bf16 K-major, M = 128)
__device__ __forceinline__ uint32_t tcv_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

template <int BK, bool SPLIT, bool RESIDENT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
tc_conv_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)p.a_slots * p.a_slot_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + (size_t)(p.b_resident ? p.b_blocks : p.b_slots * p.b_group) * p.b_slot_bytes);
  uint64_t* a_empty = a_full + TCV_MAX_SLOTS;
  uint64_t* b_full = a_empty + TCV_MAX_SLOTS;
  uint64_t* b_empty = b_full + TCV_MAX_SLOTS;
  uint64_t* tmem_full_bar = b_empty + TCV_MAX_SLOTS;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
  uint64_t* first_bar = tmem_empty_bar + 2;           // [2] merged ops: the overwriting first op of the tile has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(first_bar + 2);
  // per issuer warp: [TCV_MAX_ISSUERS][TCV_MAX_PASS][TCV_MAX_OPS] {a_shift, b_off, d_off, flags} and, per raster of a pass,
  // that issuer's op range (first | count << 8)
  uint4* ops_tab = reinterpret_cast<uint4*>(tmem_empty_bar + 6);
  uint32_t* rast_tab = reinterpret_cast<uint32_t*>(ops_tab + TCV_MAX_ISSUERS * TCV_MAX_PASS * TCV_MAX_OPS);
  const TcEpiVec ev{p.epi.bias, p.epi.scale, p.epi.shift};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpt = p.src_blocks[0] + p.src_blocks[1];  // channel blocks per tap
  constexpr uint32_t kRowBytes = 2 * BK;
  const uint32_t b_blk_lo = (uint32_t)p.b_slot_bytes >> 4;  // one weight block in descriptor-address units (16 B)

  // MMA op tables: everything an issuing thread needs per op in one 16-byte shared-memory word.  With nw issuer warps,
  // issuer w owns the accumulators a with a % nw == w (its own TMEM columns, so no ordering between issuers is needed)
  // and gets the ops of those accumulators, raster by raster, in the original order.
  const int nw = p.nw;
  if (threadIdx.x < TCV_MAX_ISSUERS * TCV_MAX_PASS) {
    const int w = threadIdx.x / TCV_MAX_PASS, ps_i = threadIdx.x % TCV_MAX_PASS;
    if (w < nw && ps_i < p.npass) {
      const TcvPass& ps = p.pass[ps_i];
      uint4* tab = ops_tab + (w * TCV_MAX_PASS + ps_i) * TCV_MAX_OPS;
      uint32_t* rt = rast_tab + (w * TCV_MAX_PASS + ps_i) * TCV_MAX_RASTERS;
      int n = 0;
      for (int ri = 0; ri < ps.nrast; ++ri) {
        const int first = n;
        for (int oi = ps.r[ri].op0; oi < ps.r[ri].op0 + ps.r[ri].nops; ++oi) {
          const TcvOp op = ps.op[oi];
          const bool mine = (p.merged ? oi : (int)op.acc) % nw == w;  // merged ops span accumulators: dealt round-robin
          if (p.b_resident && !mine) continue;  // streamed weights: every issuer walks every op (ring hand-shakes), issues its own
          uint32_t flags = (op.first ? TCV_F_FIRST : 0) | (mine ? TCV_F_MINE : 0);
          if (oi == ps.r[ri].op0) flags |= TCV_F_NEW_RASTER;
          if (oi == ps.r[ri].op0 + ps.r[ri].nops - 1) flags |= TCV_F_LAST_OF_RASTER;
          uint32_t b_off;
          if (p.merged && p.b_resident) {
            b_off = (uint32_t)op.blk0 * b_blk_lo;
          } else if (p.merged) {  // streamed: one weight-ring slot per op (its nph blocks back to back)
            flags |= TCV_F_NEW_BGROUP | TCV_F_LAST_OF_BGROUP;
            b_off = 0;
          } else if (p.b_resident) {
            b_off = (uint32_t)((ps.phase_id[op.acc] * p.n_tiles * p.ntaps + op.tap) * cpt) * b_blk_lo;
          } else {  // table order == op order == weight ring order
            if (oi % p.b_group == 0) flags |= TCV_F_NEW_BGROUP;
            if (oi % p.b_group == p.b_group - 1 || oi == ps.nops - 1) flags |= TCV_F_LAST_OF_BGROUP;
            b_off = (uint32_t)(oi % p.b_group) * b_blk_lo;
          }
          flags |= (uint32_t)(((p.merged ? op.nph : 1) * p.BN) >> 3) << 8;  // MMA N of this op (instruction-descriptor field)
          tab[n++] = make_uint4((uint32_t)op.row_shift * (kRowBytes >> 4), b_off, (uint32_t)(op.acc * p.BN), flags);
        }
        rt[ri] = (uint32_t)first | ((uint32_t)(n - first) << 8);
      }
    }
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0][0]);
    prefetch_tmap(&maps.b[0]);
    if (SPLIT) { prefetch_tmap(&maps.a[0][1]); prefetch_tmap(&maps.b[1]); }
    if (p.nsrc > 1) { prefetch_tmap(&maps.a[1][0]); if (SPLIT) prefetch_tmap(&maps.a[1][1]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TCV_MAX_SLOTS; ++s) {
        mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], (uint32_t)nw);
        mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], (uint32_t)nw);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], (uint32_t)nw); mbar_init(&tmem_empty_bar[s], 4u * (uint32_t)p.ne);
        mbar_init(&first_bar[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * TCV_ACC_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== raster (A) producer: one elected lane runs the whole role (waits, tile decode, TMA issue) =====
    if (elect_one()) {
      uint32_t sa = 0, pa = 0;  // ring slot / phase of the next raster
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TcvTile tc = tcv_decode(p, t);
        const TcvPass& ps = p.pass[tc.pass];
        const int nrast = ps.nrast;
        for (int cb = 0; cb < cpt; ++cb) {
          const int src = cb >= p.src_blocks[0] ? 1 : 0;
          const int cbase = p.src_coff[src] + (src ? cb - p.src_blocks[0] : cb) * BK;
          for (int ri = 0; ri < nrast; ++ri) {
            const TcvRaster R = ps.r[ri];
            mbar_wait_relaxed(&a_empty[sa], pa ^ 1u);
            uint8_t* a_hi = a_ring + (size_t)sa * p.a_slot_bytes;
            uint8_t* a_lo = a_hi + (p.a_slot_bytes >> 1);
            const int c = cbase + R.dc;
            if (p.dbg & 4) {
              mbar_arrive(&a_full[sa]);
            } else if (mbar_arrive_expect_tx(&a_full[sa], p.a_tx_bytes), p.a5d) {
              tma_load_5d(a_hi, &maps.a[src][0], &a_full[sa], c, tc.ox0 + R.dx, R.pz, tc.oy0 + R.dy, tc.b0);
              if (SPLIT) tma_load_5d(a_lo, &maps.a[src][1], &a_full[sa], c, tc.ox0 + R.dx, R.pz, tc.oy0 + R.dy, tc.b0);
            } else {
              tma_load_4d(a_hi, &maps.a[src][0], &a_full[sa], c, tc.ox0 + R.dx, tc.oy0 + R.dy, tc.b0);
              if (SPLIT) tma_load_4d(a_lo, &maps.a[src][1], &a_full[sa], c, tc.ox0 + R.dx, tc.oy0 + R.dy, tc.b0);
            }
            if (++sa == (uint32_t)p.a_slots) { sa = 0; pa ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 6) {
    // ===== weight (B) producer =====
    if (elect_one()) {
      if (RESIDENT) {
        // the whole layer: block index = ((phase * n_tiles + n_tile) * ntaps + tap) * cpt + cb, all on one barrier
        mbar_arrive_expect_tx(&b_full[0], p.b_tx_bytes * (uint32_t)p.b_blocks);
        const int nph = p.b_blocks / (p.n_tiles * p.ntaps * cpt);
        int blk = 0;
        if (p.merged) {
          // per channel block: the blocks of every op back to back, so a merged op's B operand is one contiguous N = nph * BN tile
          const TcvPass& ps = p.pass[0];
          for (int cb = 0; cb < cpt; ++cb)
            for (int oi = 0; oi < ps.nops; ++oi) {
              const TcvOp op = ps.op[oi];
              for (int j = 0; j < op.nph; ++j) {
                uint8_t* b_hi = b_ring + (size_t)(cb * p.nblk_cb + op.blk0 + j) * p.b_slot_bytes;
                tma_load_3d(b_hi, &maps.b[0], &b_full[0], (op.taps[j] * cpt + cb) * BK, ps.phase_id[op.acc + j] * p.N_pad, 0);
              }
            }
        } else
        for (int ph = 0; ph < nph; ++ph)
          for (int nt = 0; nt < p.n_tiles; ++nt)
            for (int tap = 0; tap < p.ntaps; ++tap)
              for (int cb = 0; cb < cpt; ++cb, ++blk) {
                uint8_t* b_hi = b_ring + (size_t)blk * p.b_slot_bytes;
                const int k = (tap * cpt + cb) * BK, row = ph * p.N_pad + nt * p.BN;
                tma_load_3d(b_hi, &maps.b[0], &b_full[0], k, row, 0);
                if (SPLIT) tma_load_3d(b_hi + (p.b_slot_bytes >> 1), &maps.b[1], &b_full[0], k, row, 0);
              }
      } else {
        uint32_t sb = 0, pb = 0;
        const int G = p.b_group;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TcvTile tc = tcv_decode(p, t);
          const TcvPass& ps = p.pass[tc.pass];
          const int nops = ps.nops;
          const int brow0 = tc.n_tile * p.BN;
          for (int cb = 0; cb < cpt; ++cb) {
            if (p.merged) {
              for (int oi = 0; oi < nops; ++oi) {
                const TcvOp op = ps.op[oi];
                mbar_wait_relaxed(&b_empty[sb], pb ^ 1u);
                uint8_t* slot = b_ring + (size_t)sb * G * p.b_slot_bytes;
                if (p.dbg & 2) mbar_arrive(&b_full[sb]);
                else mbar_arrive_expect_tx(&b_full[sb], p.b_tx_bytes * (uint32_t)op.nph);
                for (int j = 0; j < op.nph && !(p.dbg & 2); ++j)
                  tma_load_3d(slot + (size_t)j * p.b_slot_bytes, &maps.b[0], &b_full[sb], (op.taps[j] * cpt + cb) * BK,
                              ps.phase_id[op.acc + j] * p.N_pad + brow0, 0);
                if (++sb == (uint32_t)p.b_slots) { sb = 0; pb ^= 1u; }
              }
              continue;
            }
            for (int o0 = 0; o0 < nops; o0 += G) {
              const int cnt = min(G, nops - o0);
              mbar_wait_relaxed(&b_empty[sb], pb ^ 1u);
              uint8_t* slot = b_ring + (size_t)sb * G * p.b_slot_bytes;
              if (p.dbg & 2) mbar_arrive(&b_full[sb]);
              else mbar_arrive_expect_tx(&b_full[sb], p.b_tx_bytes * (uint32_t)cnt);
              for (int j = 0; j < cnt && !(p.dbg & 2); ++j) {
                const TcvOp op = ps.op[o0 + j];
                uint8_t* b_hi = slot + (size_t)j * p.b_slot_bytes;
                const int k = (op.tap * cpt + cb) * BK;
                const int row = ps.phase_id[op.acc] * p.N_pad + brow0;
                tma_load_3d(b_hi, &maps.b[0], &b_full[sb], k, row, 0);
                if (SPLIT) tma_load_3d(b_hi + (p.b_slot_bytes >> 1), &maps.b[1], &b_full[sb], k, row, 0);
              }
              if (++sb == (uint32_t)p.b_slots) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 7 || warp >= 12) {
    // ===== MMA issuers: warp 1 (+ warps 7.. when nw > 1); one elected lane of each runs the whole role.  Every op is one
    // 16-byte shared-memory word fetched one op ahead (ld.shared).  A single thread sustains ~5 cycles per dependent
    // instruction and ~90 per mbarrier try_wait, so with N <= 64 per MMA one issuer cannot keep the tensor pipe busy
    // (elimination runs, profiles/r01_smem_pipe_model.md section 1): the accumulators (output phases) are split over nw
    // issuers when the weights are resident.
    const int w = warp == 1 ? 0 : (warp == 7 ? 1 : warp - 10);
    if (w < nw && elect_one()) {
      const uint32_t idesc0 = tcv_idesc(0);  // N comes from the op table
      const uint32_t b_cb_step = b_blk_lo * (uint32_t)(p.merged ? p.nblk_cb : 1);
      const uint32_t a_ring_lo = (smem_u32(a_ring) & 0x3FFFF) >> 4, b_ring_lo = (smem_u32(b_ring) & 0x3FFFF) >> 4;
      const uint32_t a_slot_lo = (uint32_t)p.a_slot_bytes >> 4, a_half_lo = a_slot_lo >> 1;
      const uint32_t b_slot_lo = b_blk_lo * (uint32_t)p.b_group, b_half_lo = b_blk_lo >> 1;
      const uint32_t n_aslots = (uint32_t)p.a_slots, n_bslots = (uint32_t)p.b_slots;
      const uint32_t res_tile_stride = (uint32_t)(p.ntaps * cpt) * b_blk_lo;
      const uint32_t tab0 = smem_u32(ops_tab) + (uint32_t)(w * TCV_MAX_PASS * TCV_MAX_OPS) * 16u;
      const uint32_t rt0 = smem_u32(rast_tab) + (uint32_t)(w * TCV_MAX_PASS * TCV_MAX_RASTERS) * 4u;
      constexpr bool resident = RESIDENT;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, b_base = 0;
      int lt = 0;
      if (resident) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      if (!resident) {
        // streamed weights: one flat pass over the op table (flags drive the raster / weight-ring hand-shakes)
        uint32_t a_base = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
          const TcvTile tc = tcv_decode(p, t);
          const int nops = p.pass[tc.pass].nops;
          const uint32_t tab = tab0 + (uint32_t)(tc.pass * TCV_MAX_OPS) * 16u;
          const int as = lt & 1;
          mbar_wait(&tmem_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)(as * TCV_ACC_COLS);
          uint4 o = lds_v4(tab);
          for (int cb = 0; cb < cpt; ++cb) {
            const uint32_t fresh_mask = cb == 0 ? (uint32_t)TCV_F_FIRST : 0u;
            for (int i = 0; i < nops; ++i) {
              const uint4 nx = lds_v4(tab + (uint32_t)(i + 1 < nops ? i + 1 : 0) * 16u);  // next op (op 0 of the next channel block)
              if (o.w & TCV_F_NEW_RASTER) {
                mbar_wait(&a_full[sa], pa);
                tc_fence_after();
                a_base = a_ring_lo + sa * a_slot_lo;
              }
              if (o.w & TCV_F_NEW_BGROUP) {
                mbar_wait(&b_full[sb], pb);
                tc_fence_after();
                b_base = b_ring_lo + sb * b_slot_lo;
              }
              if ((o.w & TCV_F_MINE) && !(p.dbg & 1)) {
                const uint32_t idesc = idesc0 | (((o.w >> 8) & 0xFFu) << 17);
                const uint32_t a_hi = a_base + o.x, b_hi = b_base + o.y, d = d0 + o.z;
                const uint32_t keep = (o.w & fresh_mask) ? 0u : 1u;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, k == 0 ? keep : 1u);
                if (SPLIT) {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(d, umma_desc_from_lo<BK>(a_hi + a_half_lo + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, 1u);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_hi + b_half_lo + 2 * k), idesc, 1u);
                }
              }
              if (o.w & TCV_F_LAST_OF_BGROUP) {
                umma_commit(&b_empty[sb]);
                if (++sb == n_bslots) { sb = 0; pb ^= 1u; }
              }
              if (o.w & TCV_F_LAST_OF_RASTER) {
                umma_commit(&a_empty[sa]);
                if (++sa == n_aslots) { sa = 0; pa ^= 1u; }
              }
              o = nx;
            }
          }
          umma_commit(&tmem_full_bar[as]);
        }
      } else {
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
          const TcvTile tc = tcv_decode(p, t);
          const int nrast = p.pass[tc.pass].nrast;
          const uint32_t tab = tab0 + (uint32_t)(tc.pass * TCV_MAX_OPS) * 16u;
          const uint32_t rt = rt0 + (uint32_t)(tc.pass * TCV_MAX_RASTERS) * 4u;
          const int as = lt & 1;
          mbar_wait(&tmem_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (p.merged && w != 0) {
            // merged ops of different issuers accumulate into the same TMEM columns: order does not matter for the sums, but
            // the first op of the tile overwrites - wait until it has retired
            mbar_wait(&first_bar[as], ((uint32_t)lt >> 1) & 1u);
            tc_fence_after();
          }
          const uint32_t d0 = tmem_base + (uint32_t)(as * TCV_ACC_COLS);
          uint32_t b_cb = b_ring_lo + (uint32_t)tc.n_tile * res_tile_stride;  // resident: block (phase 0, n_tile, tap 0, cb)
          for (int cb = 0; cb < cpt; ++cb) {
            const uint32_t fresh_mask = cb == 0 ? (uint32_t)TCV_F_FIRST : 0u;
            for (int ri = 0; ri < nrast; ++ri) {
              const uint32_t rinfo = lds_u32(rt + (uint32_t)ri * 4u);
              const uint32_t op0 = rinfo & 0xFFu, op1 = op0 + (rinfo >> 8);
              uint4 o = lds_v4(tab + op0 * 16u);
              mbar_wait(&a_full[sa], pa);
              tc_fence_after();
              const uint32_t a_base = a_ring_lo + sa * a_slot_lo;
              for (uint32_t i = op0; i < op1; ++i) {
                const uint4 nx = lds_v4(tab + (i + 1 < op1 ? i + 1 : i) * 16u);  // one op ahead
                if (!resident && (o.w & TCV_F_NEW_BGROUP)) {
                  mbar_wait(&b_full[sb], pb);
                  tc_fence_after();
                  b_base = b_ring_lo + sb * b_slot_lo;
                }
                const uint32_t idesc = idesc0 | (((o.w >> 8) & 0xFFu) << 17);
                const uint32_t a_hi = a_base + o.x, b_hi = (resident ? b_cb : b_base) + o.y, d = d0 + o.z;
                const uint32_t keep = (o.w & fresh_mask) ? 0u : 1u;
                if (!(p.dbg & 1)) {
  #pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, k == 0 ? keep : 1u);
                }
                if (SPLIT && !(p.dbg & 1)) {
  #pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(d, umma_desc_from_lo<BK>(a_hi + a_half_lo + 2 * k), umma_desc_from_lo<BK>(b_hi + 2 * k), idesc, 1u);
  #pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(d, umma_desc_from_lo<BK>(a_hi + 2 * k), umma_desc_from_lo<BK>(b_hi + b_half_lo + 2 * k), idesc, 1u);
                }
                if (!resident && (o.w & TCV_F_LAST_OF_BGROUP)) {
                  umma_commit(&b_empty[sb]);
                  if (++sb == n_bslots) { sb = 0; pb ^= 1u; }
                }
                if (p.merged && (o.w & fresh_mask)) umma_commit(&first_bar[as]);  // the overwriting op (issuer 0 owns it)
                o = nx;
              }
              umma_commit(&a_empty[sa]);  // this issuer's share of the raster slot: free once its MMAs have retired
              if (++sa == n_aslots) { sa = 0; pa ^= 1u; }
            }
            b_cb += b_cb_step;  // resident: next channel block
          }
          umma_commit(&tmem_full_bar[as]);
        }
      }
    }
    __syncwarp();
  } else if ((warp >= 2 && warp <= 5) || (warp >= 8 && warp <= 11)) {
    // ===== epilogue: one or two groups of 4 warps; warp w owns TMEM lanes 32*(w%4)...  One warp per scheduler runs the
    // ~350-instruction chunk at ~7 cycles per instruction (latency-bound: ncu source view), so layers with narrow MMAs
    // (deconv4: four 32-column accumulators per 128 MMAs) are epilogue-bound with four warps; with ne == 2 the second
    // group takes every other accumulator chunk.
    const int eg = warp >= 8 ? 1 : 0;
    if (eg < p.ne) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int xl = r % p.TW, yl = (r / p.TW) % p.TH, bl = r / (p.TW * p.TH);
      const bool wide = (p.BN & 31) == 0;  // 32-column accumulator chunks, else 16
      const int chunks = wide ? p.BN / 32 : p.BN / 16;
      int lt = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
        const TcvTile tc = tcv_decode(p, t);
        const TcvPass& ps = p.pass[tc.pass];
        const int as = lt & 1;
        const int ox = tc.ox0 + xl, oy = tc.oy0 + yl, b = tc.b0 + bl;
        const bool valid = (bl < p.TB) && ox < p.Wo && oy < p.Ho && b < p.batch;
        mbar_wait_relaxed(&tmem_full_bar[as], ((uint32_t)lt >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TCV_ACC_COLS);
        // this group's chunks: global chunk index g = j * chunks + c with g % ne == eg; the last one hands TMEM back
        const int total = ps.nacc * chunks;
        const int mine_last = total - 1 - ((total - 1 - eg) % p.ne + p.ne) % p.ne;  // largest g <= total-1 with g % ne == eg (or < 0)
        if (mine_last < 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
#pragma unroll 1
        for (int g = eg; g < total; g += p.ne) {
          const int j = g / chunks, c = g - j * chunks;
          const TcRow row{b, oy, ox, ps.phase_id[j], 0};
          uint32_t v[32];
          __syncwarp();
          if (wide) tmem_ld32(taddr + (uint32_t)(j * p.BN + c * 32), v);
          else tmem_ld16(taddr + (uint32_t)(j * p.BN + c * 16), v);
          tmem_ld_wait();
          if (g == mine_last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
          }
          const int ch = wide ? 32 : 16;
          const int n0 = tc.n_tile * p.BN + c * ch;
          const int nv = min(ch, p.epi.N - n0);
          if (nv > 0 && !(p.dbg & 8)) {
            if (wide) tc_epilogue_store<32>(p.epi, ev, row, v, n0, nv, valid, p.TW);
            else tc_epilogue_store<16>(p.epi, ev, row, v, n0, nv, valid, 0);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * TCV_ACC_COLS);
}

size_t tcv_smem_bytes(const TcvParams& p) {
  size_t n = (size_t)p.a_slots * p.a_slot_bytes + (size_t)(p.b_resident ? p.b_blocks : p.b_slots * p.b_group) * p.b_slot_bytes +
             1024 /*align*/ + 6144 /*barriers + op tables*/;
  // the kernel allocates all 512 TMEM columns: keep it to one CTA per SM whatever the ring sizes are
  const size_t floor_bytes = 120 * 1024;
  return n < floor_bytes ? floor_bytes : n;
}

template <int BK, bool SPLIT, bool RESIDENT, int THREADS>
static int launch_conv_thr(const TcMaps& maps, const TcvParams& p, cudaStream_t st) {
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<BK, SPLIT, RESIDENT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set.done();
  }
  const size_t smem = tcv_smem_bytes(p);
  CIC_REQUIRE(smem <= 227 * 1024, "tc_conv: %zu bytes of shared memory needed", smem);
  const int slots = sm_count();
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  tc_conv_kernel<BK, SPLIT, RESIDENT, THREADS><<<grid, THREADS, smem, st>>>(maps, p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("tc_conv_kernel");
  g_last_kernel_kind = KK_TC_CONV;
  return CIC_OK;
}

// block size by role count: 224 (one issuer, one epilogue group), 384 (<= 2 issuers, 2 epilogue groups), 448 (4 issuers)
template <int BK, bool SPLIT, bool RESIDENT>
static int launch_conv_one(const TcMaps& maps, const TcvParams& p, cudaStream_t st) {
  if (p.nw > 2) return launch_conv_thr<BK, SPLIT, RESIDENT, 448>(maps, p, st);
  if (p.nw > 1 || p.ne > 1) return launch_conv_thr<BK, SPLIT, RESIDENT, 384>(maps, p, st);
  return launch_conv_thr<BK, SPLIT, RESIDENT, 224>(maps, p, st);
}

int launch_tc_conv(const TcMaps& maps, const TcvParams& p, int block_k, bool split, cudaStream_t st) {
  CIC_REQUIRE(p.total_tiles > 0 && p.npass >= 1 && p.npass <= TCV_MAX_PASS, "tc_conv: bad tile list");
  CIC_REQUIRE(p.BN % 16 == 0 && p.BN >= 16 && p.BN <= TCV_ACC_COLS, "tc_conv: bad accumulator width %d", p.BN);
  CIC_REQUIRE(p.a_slots >= 2 && p.a_slots <= TCV_MAX_SLOTS && (p.b_resident || (p.b_slots >= 2 && p.b_slots <= TCV_MAX_SLOTS)),
              "tc_conv: bad ring sizes");
  CIC_REQUIRE(p.b_resident || (p.b_group >= 1 && p.b_group <= TCV_MAX_OPS), "tc_conv: bad weight group");
  CIC_REQUIRE(p.nw >= 1 && p.nw <= TCV_MAX_ISSUERS && true, "tc_conv: bad issuer count %d", p.nw);
  CIC_REQUIRE(p.ne == 1 || p.ne == 2, "tc_conv: bad epilogue group count %d", p.ne);
  for (int i = 0; i < p.npass; ++i)
    CIC_REQUIRE(p.pass[i].nops >= 1 && p.pass[i].nops <= TCV_MAX_OPS && p.pass[i].nacc >= 1 && p.pass[i].nacc * p.BN <= TCV_ACC_COLS && p.pass[i].nrast >= 1 && p.pass[i].nrast <= TCV_MAX_RASTERS,
                "tc_conv: bad pass %d", i);
  const bool res = p.b_resident != 0;
  if (block_k == 64) {
    if (split) return res ? launch_conv_one<64, true, true>(maps, p, st) : launch_conv_one<64, true, false>(maps, p, st);
    return res ? launch_conv_one<64, false, true>(maps, p, st) : launch_conv_one<64, false, false>(maps, p, st);
  }
  if (block_k == 32) {
    if (split) return res ? launch_conv_one<32, true, true>(maps, p, st) : launch_conv_one<32, true, false>(maps, p, st);
    return res ? launch_conv_one<32, false, true>(maps, p, st) : launch_conv_one<32, false, false>(maps, p, st);
  }
  set_error("tc_conv: unsupported K block %d", block_k);
  return CIC_ERR_INVALID;
}

}  // namespace cic
