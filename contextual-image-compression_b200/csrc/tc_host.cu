// Host-side assembly of tcgen05 layer launches + conversion kernels (see tc_host.cuh).
#include "tc_host.cuh"
#include "tc_conv.cuh"

#include <algorithm>
#include <vector>
#include <cstdlib>
#include <cstring>

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


// ------------------------------------------------------------------------------------------------
// raster kernel (tc_conv.cuh): build the tap program of a conv / transposed-conv layer
// ------------------------------------------------------------------------------------------------
struct TapDesc {
  int dc, pz, dx, dy;  // TMA coordinate deltas (as TcTap)
  int tap;             // index in the weight K order of its phase
  int phase;
};

static void layer_taps(const TcLayer& L, std::vector<TapDesc>& taps) {
  taps.clear();
  if (L.kind == TC_DECONV_K4S2) {
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) taps.push_back({0, 0, tx - (px == 0 ? 1 : 0), ty - (py == 0 ? 1 : 0), ty * 2 + tx, ph});
    }
    return;
  }
  for (int ky = 0; ky < L.kh; ++ky)
    for (int kx = 0; kx < L.kw; ++kx) {
      const int dy = ky - L.pad_t, dx = kx - L.pad_l;
      if (L.kind == TC_CONV_S2) {
        const int yb = dy >= 0 ? dy / 2 : -((-dy + 1) / 2), xb = dx >= 0 ? dx / 2 : -((-dx + 1) / 2);
        taps.push_back({(dx - 2 * xb) * L.src[0].ld, dy - 2 * yb, xb, yb, ky * L.kw + kx, 0});
      } else {
        taps.push_back({0, 0, dx, dy, ky * L.kw + kx, 0});
      }
    }
}

// weights too large to stay resident beside 64-channel rasters but small enough beside 32-channel ones
static bool raster_prefers_bk32(const TcLayer& L, int N_pad) {
  if (L.split || L.b_batched || L.splits != 1) return false;
  const bool dc = L.kind == TC_DECONV_K4S2;
  const long long wbytes = (long long)(dc ? 4 : 1) * N_pad * L.w.K * 2;
  return wbytes > 96 * 1024 && wbytes <= 136 * 1024 && (long long)L.batch * L.H * L.W >= 148LL * 128 * 16;
}

// returns CIC_OK and sets *used = true when the layer was launched on the raster kernel; *used = false means
// "not eligible, use the per-tap kernel"
static int try_run_raster(const TcLayer& L, int BK, int N_pad, const TcMaps& act_maps_unused, cudaStream_t st, bool* used) {
  (void)act_maps_unused;
  *used = false;
  const bool s2 = L.kind == TC_CONV_S2, dc = L.kind == TC_DECONV_K4S2;
  static const int force = CIC_KNOB("CIC_TC_RASTER", -1);  // 0: never, 1: whenever possible
  if (force == 0) return CIC_OK;
  if (L.b_batched || L.splits != 1 || L.epi.out_mode == TC_OUT_PARTIAL) return CIC_OK;
  if (!dc && L.kh * L.kw == 1) return CIC_OK;  // 1x1 / dense: nothing to reuse
  int BN = 0;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N_pad % bn == 0) { BN = bn; break; }
  if (!BN) return CIC_OK;
  const int parts = L.split ? 2 : 1;
  const int b_slot_bytes = BN * 2 * BK * parts;
  if (force != 1 && b_slot_bytes > 32768) return CIC_OK;  // wide, K-heavy layers are already tensor-bound on the per-tap kernel

  TcvParams p;
  memset(&p, 0, sizeof(p));
  const int Ho = s2 ? L.H / 2 : L.H, Wo = s2 ? L.W / 2 : L.W;
  pick_tile(Wo, Ho, L.batch, false, p.TW, p.TH, p.TB);
  p.tiles_x = (Wo + p.TW - 1) / p.TW;
  p.tiles_y = (Ho + p.TH - 1) / p.TH;
  p.tiles_b = (L.batch + p.TB - 1) / p.TB;
  p.Wo = Wo; p.Ho = Ho; p.batch = L.batch;
  p.a5d = s2 ? 1 : 0;
  p.nsrc = L.nsrc;
  for (int s = 0; s < L.nsrc; ++s) { p.src_blocks[s] = L.src[s].C / BK; p.src_coff[s] = L.src[s].coff; }
  p.N_pad = N_pad; p.BN = BN; p.n_tiles = N_pad / BN;

  std::vector<TapDesc> taps;
  layer_taps(L, taps);
  const int nph = dc ? 4 : 1;
  const bool fused = dc && p.n_tiles == 1 && 4 * BN <= TCV_ACC_COLS;
  p.npass = fused ? 1 : nph;
  const int mode = (p.TH == 1 && p.TB == 1) ? 0 : (p.TB == 1 ? 1 : 2);  // 0 ROW, 1 COL, 2 per tap

  // pass -> its taps (acc = local phase index)
  struct PTap { TapDesc t; int acc; };
  std::vector<std::vector<PTap>> ptaps(p.npass);
  for (const TapDesc& t : taps) {
    const int ps = fused ? 0 : t.phase, acc = fused ? t.phase : 0;
    ptaps[ps].push_back({t, acc});
  }
  // raster box: the maximum halo over all planes and passes
  int ext_w = 0, ext_h = 0;
  for (auto& pt : ptaps)
    for (size_t i = 0; i < pt.size(); ++i)
      for (size_t j = 0; j < pt.size(); ++j)
        if (pt[i].t.dc == pt[j].t.dc && pt[i].t.pz == pt[j].t.pz) {
          ext_w = std::max(ext_w, pt[i].t.dx - pt[j].t.dx);
          ext_h = std::max(ext_h, pt[i].t.dy - pt[j].t.dy);
        }
  p.rw = mode == 0 ? p.TW + ext_w : p.TW;
  p.rh = mode == 2 ? p.TH : p.TH + ext_h;
  p.rb = p.TB;
  if (p.rw > 256 || p.rh > 256) return CIC_OK;
  int max_end = 0;
  for (int ps = 0; ps < p.npass; ++ps) {
    TcvPass& P = p.pass[ps];
    P.nacc = fused ? 4 : 1;
    for (int a = 0; a < P.nacc; ++a) P.phase_id[a] = (int8_t)(fused ? a : ps);
    bool first_seen[4] = {false, false, false, false};
    std::vector<PTap>& pt = ptaps[ps];
    std::vector<bool> done(pt.size(), false);
    int nops = 0;
    for (size_t i = 0; i < pt.size(); ++i) {
      if (done[i]) continue;
      // raster key: the plane, plus dx (COL) or the tap itself (per tap)
      std::vector<size_t> grp;
      for (size_t j = i; j < pt.size(); ++j) {
        if (done[j] || pt[j].t.dc != pt[i].t.dc || pt[j].t.pz != pt[i].t.pz) continue;
        if (mode == 1 && pt[j].t.dx != pt[i].t.dx) continue;
        if (mode == 2 && (pt[j].t.dx != pt[i].t.dx || pt[j].t.dy != pt[i].t.dy)) continue;
        grp.push_back(j);
      }
      int dx0 = pt[grp[0]].t.dx, dy0 = pt[grp[0]].t.dy;
      for (size_t j : grp) { dx0 = std::min(dx0, pt[j].t.dx); dy0 = std::min(dy0, pt[j].t.dy); }
      if (P.nrast >= TCV_MAX_RASTERS || nops + (int)grp.size() > TCV_MAX_OPS) return CIC_OK;
      TcvRaster& R = P.r[P.nrast++];
      R.dc = (int16_t)pt[i].t.dc; R.pz = (int16_t)pt[i].t.pz; R.dx = (int16_t)dx0; R.dy = (int16_t)dy0;
      R.op0 = (uint8_t)nops; R.nops = (uint8_t)grp.size();
      P.nops = nops + (int)grp.size();
      for (size_t j : grp) {
        TcvOp& o = P.op[nops++];
        const int sh = (pt[j].t.dy - dy0) * p.rw + (pt[j].t.dx - dx0);
        o.row_shift = (uint16_t)sh;
        o.acc = (uint8_t)pt[j].acc;
        o.tap = (uint8_t)pt[j].t.tap;
        o.first = first_seen[pt[j].acc] ? 0 : 1;
        first_seen[pt[j].acc] = true;
        max_end = std::max(max_end, sh + TC_BM);
        done[j] = true;
      }
    }
  }
  const long long mt = (long long)p.tiles_x * p.tiles_y * p.tiles_b;
  const long long total = mt * p.npass * p.n_tiles;
  if (total <= 0 || total >= 2147483647LL) return CIC_OK;
  p.total_tiles = (int)total;
  // shared-memory rings
  const int raster_rows = p.rw * p.rh * p.rb;
  const int a_rows = std::max(raster_rows, max_end);
  const int a_half = ((a_rows * 2 * BK) + 1023) & ~1023;
  p.a_slot_bytes = a_half * parts;
  p.a_tx_bytes = (uint32_t)(raster_rows * 2 * BK * parts);
  p.b_slot_bytes = b_slot_bytes;
  p.b_tx_bytes = (uint32_t)b_slot_bytes;
  const int budget = 218 * 1024;
  p.ntaps = dc ? 4 : L.kh * L.kw;
  p.b_blocks = nph * p.n_tiles * p.ntaps * (p.src_blocks[0] + p.src_blocks[1]);
  const long long resident_bytes = (long long)p.b_blocks * b_slot_bytes;
  static const int no_resident = CIC_KNOB("CIC_TC_NO_RESIDENT", 0);
  if (!no_resident && resident_bytes <= budget - 2LL * p.a_slot_bytes && resident_bytes <= 136 * 1024 &&
      resident_bytes * 148 < (long long)p.total_tiles * 16 * b_slot_bytes) {
    // the whole weight matrix stays in shared memory (only when that is less traffic than streaming it)
    p.b_resident = 1;
    p.b_slots = 0;
    p.a_slots = std::min(8, (int)((budget - resident_bytes) / p.a_slot_bytes));
  } else {
    // weight blocks travel in groups of up to 16 KB under one barrier (one wait per group in the MMA warp)
    static const int gbytes_env = CIC_KNOB("CIC_TC_BGROUP_BYTES", 0);
    p.b_group = std::max(1, std::min(4, (gbytes_env > 0 ? gbytes_env : 32768) / b_slot_bytes));
    const int gbytes = p.b_group * b_slot_bytes;
    p.b_slots = std::min(TCV_MAX_SLOTS, std::max(2, 65536 / gbytes));
    p.a_slots = std::min(8, (budget - p.b_slots * gbytes) / p.a_slot_bytes);
    if (p.a_slots < 2) {  // big rasters: give the weight ring less
      p.b_slots = std::max(2, std::min(TCV_MAX_SLOTS, (budget - 2 * p.a_slot_bytes) / gbytes));
      p.a_slots = (budget - p.b_slots * gbytes) / p.a_slot_bytes;
    }
  }
  if (p.a_slots < 2) return CIC_OK;
  // Resident transposed conv: merge the four phases' MMAs per input shift.  The 16 (phase, tap) products read only 9
  // distinct shifted A windows; with the accumulators ordered (0,0) (0,1) (1,1) (1,0) the phases that share a shift are
  // adjacent (except one wrap-around pair), so one MMA of N = nph * BN serves them: 11 ops instead of 16 and 9/16 of the
  // A operand reads - the shared-memory data pipe is what bounds these N <= 32 layers (profiles/r01_smem_pipe_model.md).
  // CIC_TC_MERGE: 0 off, 1 resident-weight layers (default; deconv4 1.13 -> 0.87 ms), 2 also streamed weights on this kernel
  // (measured slower than the per-tap kernel for deconv3: 0.87 vs 0.70 ms - the raster kernel's per-op hand-shakes)
  static const int merge_env = CIC_KNOB("CIC_TC_MERGE", 1);
  if (merge_env && fused && !L.split && mode != 2 && 4 * BN <= TCV_ACC_COLS && (p.b_resident || merge_env >= 2)) {
    TcvPass& P = p.pass[0];
    static const int acc_phase[4] = {0, 1, 3, 2};                 // accumulator -> phase (py * 2 + px)
    struct MOp { int dy, dx, acc0, nph; };
    static const MOp mops[11] = {{0, 0, 0, 4},  {-1, 0, 0, 2}, {1, 0, 2, 2},  {0, 1, 1, 2},  {0, -1, 3, 1}, {0, -1, 0, 1},
                                 {-1, -1, 0, 1}, {-1, 1, 1, 1}, {1, 1, 2, 1}, {1, -1, 3, 1}};
    const int nm = 10;
    // rasters: ROW mode one raster holds every shift; COL mode one raster per dx, the dx = 0 raster (centre op) first
    const int dx_order[3] = {0, -1, 1};
    memset(&P, 0, sizeof(P));
    P.nacc = 4;
    for (int a = 0; a < 4; ++a) P.phase_id[a] = (int8_t)acc_phase[a];
    int nops = 0, blk = 0;
    max_end = 0;
    const int nr = mode == 0 ? 1 : 3;
    for (int ri = 0; ri < nr; ++ri) {
      TcvRaster& R = P.r[P.nrast++];
      R.dc = 0; R.pz = 0;
      R.dx = (int16_t)(mode == 0 ? -1 : dx_order[ri]);
      R.dy = -1;
      R.op0 = (uint8_t)nops;
      for (int m = 0; m < nm; ++m) {
        if (mode == 1 && mops[m].dx != dx_order[ri]) continue;
        TcvOp& o = P.op[nops];
        const int sh = (mops[m].dy - R.dy) * p.rw + (mops[m].dx - R.dx);
        o.row_shift = (uint16_t)sh;
        o.acc = (uint8_t)mops[m].acc0;
        o.nph = (uint8_t)mops[m].nph;
        o.blk0 = (uint8_t)blk;
        o.first = nops == 0 ? 1 : 0;
        for (int j = 0; j < mops[m].nph; ++j) {
          const int ph = acc_phase[mops[m].acc0 + j], py = ph >> 1, px = ph & 1;
          const int ty = mops[m].dy + (py == 0 ? 1 : 0), tx = mops[m].dx + (px == 0 ? 1 : 0);
          o.taps[j] = (uint8_t)(ty * 2 + tx);
        }
        o.tap = o.taps[0];
        blk += mops[m].nph;
        max_end = std::max(max_end, sh + TC_BM);
        ++nops;
      }
      R.nops = (uint8_t)(nops - R.op0);
    }
    P.nops = nops;
    p.merged = 1;
    p.nblk_cb = blk;  // 16
    if (!p.b_resident) {  // streamed: one ring slot per op, sized for four blocks
      p.b_group = 4;
      const int gbytes = p.b_group * b_slot_bytes;
      p.b_slots = std::min(TCV_MAX_SLOTS, std::max(2, (budget - 3 * p.a_slot_bytes) / gbytes));
      p.a_slots = std::min(8, (budget - p.b_slots * gbytes) / p.a_slot_bytes);
      if (p.a_slots < 2) return CIC_OK;
    }
  }
  // Streamed weights: measured (r01, CIC_TC_RASTER=0 vs default on conv2 / deconv1..3) the per-tap kernel is 10-25 % faster -
  // one barrier pair per K block there against separate raster + weight-group hand-shakes per op here, and both
  // are bound by the shared-memory data pipe rather than by L2 traffic.  The raster kernel keeps the resident-weight layers.
  if (!p.b_resident && force != 1 && !p.merged) return CIC_OK;

  // tensor maps
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int s = 0; s < L.nsrc; ++s) {
    const TcAct& a = L.src[s];
    for (int part = 0; part < parts; ++part) {
      const bf16* base = part ? a.lo : a.hi;
      int rc;
      if (!s2) {
        const uint64_t dims[4] = {(uint64_t)a.ld, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.batch};
        const uint64_t str[3] = {(uint64_t)a.ld * 2, (uint64_t)L.W * a.ld * 2, (uint64_t)L.H * L.W * a.ld * 2};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.rw, (uint32_t)p.rh, (uint32_t)p.rb};
        rc = tc_encode_map(&maps.a[s][part], base, 4, dims, str, box);
      } else {
        const uint64_t dims[5] = {(uint64_t)2 * a.ld, (uint64_t)L.W / 2, 2, (uint64_t)L.H / 2, (uint64_t)L.batch};
        const uint64_t str[4] = {(uint64_t)2 * a.ld * 2, (uint64_t)L.W * a.ld * 2, (uint64_t)2 * L.W * a.ld * 2,
                                 (uint64_t)L.H * L.W * a.ld * 2};
        const uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.rw, 1, (uint32_t)p.rh, (uint32_t)p.rb};
        rc = tc_encode_map(&maps.a[s][part], base, 5, dims, str, box);
      }
      if (rc) return rc;
    }
  }
  for (int part = 0; part < parts; ++part) {
    const uint64_t dims[3] = {(uint64_t)L.w.K, (uint64_t)L.w.rows, (uint64_t)L.w.batches};
    const uint64_t str[2] = {(uint64_t)L.w.row_stride * 2, (uint64_t)L.w.batch_stride * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, 1};
    int rc = tc_encode_map(&maps.b[part], part ? L.w.lo : L.w.hi, 3, dims, str, box);
    if (rc) return rc;
  }
  // epilogue
  const TcEpilogue& e = L.epi;
  TcEpi& pe = p.epi;
  pe.bias = e.bias; pe.scale = e.scale; pe.shift = e.shift; pe.alpha = e.alpha; pe.act = e.act;
  pe.out_mode = e.out_mode; pe.out_hi = e.out_hi; pe.out_lo = e.out_lo; pe.res_hi = e.res_hi; pe.res_lo = e.res_lo;
  pe.N = L.N; pe.out_ld = e.out_ld ? e.out_ld : L.N; pe.out_coff = e.out_coff; pe.up2 = e.up2;
  pe.Ho = Ho; pe.Wo = Wo;
  pe.tm_tx = e.tm_tx; pe.tm_ty = e.tm_ty; pe.tm_IH = e.tm_IH; pe.tm_IW = e.tm_IW;
  pe.m_total = (long long)L.batch * Ho * Wo;
  if (dc) {
    for (int ph = 0; ph < 4; ++ph) { pe.out_y0[ph] = (int8_t)(ph >> 1); pe.out_x0[ph] = (int8_t)(ph & 1); }
    pe.out_ys = 2; pe.out_xs = 2; pe.out_H = 2 * L.H; pe.out_W = 2 * L.W;
  } else {
    pe.out_ys = 1; pe.out_xs = 1; pe.out_H = Ho * (e.up2 ? 2 : 1); pe.out_W = Wo * (e.up2 ? 2 : 1);
  }
  p.fd_ntiles = make_fastdiv((uint32_t)p.n_tiles);
  p.fd_npass = make_fastdiv((uint32_t)p.npass);
  p.fd_tx = make_fastdiv((uint32_t)p.tiles_x);
  p.fd_ty = make_fastdiv((uint32_t)p.tiles_y);
  // issuer warps: one per accumulator (output phase) when the weights are resident and the MMAs are small
  static const int nw_env = CIC_KNOB("CIC_TC_NW", 0);
  p.nw = 1;
  if (fused) p.nw = nw_env > 0 ? nw_env : (p.b_resident ? 2 : 1);  // measured r01: 2 issuers (384 threads, 168 regs) beat 4 (448 threads: 128-register
                                                                  // cap spills the epilogue); streamed weights: extra issuers gave nothing
  // a second epilogue group when one tile has several accumulator chunks to drain per few MMAs
  static const int ne_env = CIC_KNOB("CIC_TC_NE", 0);
  {
    const int chunks_per_tile = (fused ? 4 : 1) * (BN % 32 == 0 ? BN / 32 : BN / 16);
    p.ne = ne_env > 0 ? ne_env : (chunks_per_tile >= 2 ? 2 : 1);
  }
  static const int dbg = CIC_KNOB("CIC_TC_DBG", 0);
  p.dbg = dbg;
  *used = true;
  return launch_tc_conv(maps, p, BK, L.split, st);
}

// CTA-pair kernel eligibility (also used by the split-K planner of the Dense layers): K block 64, N tile >= 128 (an N
// tile of 64 measured slower than one CTA), at least one full pair of M tiles and few phantom tiles
bool tc_pair_ok(int bk, long long m_tiles, int n_pad, int n) {
  const int bn2 = tc2_pick_block_n(n_pad);
  return bk == 64 && bn2 >= 128 && n % 32 == 0 && m_tiles >= 2 && (m_tiles % 2 == 0 || m_tiles >= 16);
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
  {
    const int n_pad = dc ? L.w.rows / 4 : L.w.rows;
    const int cin = L.src[0].C + (L.nsrc > 1 ? L.src[1].C : 0);
    CIC_REQUIRE((long long)(dc ? 4 : L.kh * L.kw) * cin == L.w.K, "tc layer: weight K=%d does not match taps x channels", L.w.K);
    CIC_REQUIRE(L.epi.out_hi && L.N <= n_pad && n_pad % 16 == 0, "tc layer: bad output (N=%d, padded %d)", L.N, n_pad);
    CIC_REQUIRE(L.w.row_stride % 8 == 0 && L.w.batch_stride % 8 == 0, "tc layer: B rows must be 16-byte aligned");
    if (L.epi.out_mode == TC_OUT_BF16) {
      const int ld = L.epi.out_ld ? L.epi.out_ld : L.N;
      CIC_REQUIRE(L.N % 16 == 0 && ld % 8 == 0 && L.epi.out_coff % 8 == 0,
                  "tc layer: bf16 output needs N %% 16 == 0 and 16-byte aligned records (N=%d, ld=%d, coff=%d)", L.N, ld, L.epi.out_coff);
    }
    bool used = false;
    // CIC_TC_DC2: merged-phase CTA-pair kernel for transposed convs with Cout in {32, 64} (tc_gemm2.cu): 0 off, 1 the layers
    // the raster kernel streams weights for (deconv3), 2 also the resident-weight raster layers (deconv4)
    static const int dc2_env = CIC_KNOB("CIC_TC_DC2", 1);
    bool skip_raster = false;
    if (dc && dc2_env >= 2) {
      int tw, th, tb;
      pick_tile(Wo, Ho, L.batch, false, tw, th, tb);
      const long long mt = (long long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((L.batch + tb - 1) / tb);
      skip_raster = tc_deconv2_ok(BK, L.split, mt, n_pad, L.N);
    }
    // 64-channel layers whose whole weight matrix fits in shared memory only next to 32-channel rasters (deconv4:
    // 128 KB of weights + 3 x 25 KB rasters) run with the smaller K block and resident weights
    if (!skip_raster && BK == 64 && raster_prefers_bk32(L, n_pad)) {
      const int rc32 = try_run_raster(L, 32, n_pad, maps, st, &used);
      if (rc32 || used) return rc32;
    }
    if (!skip_raster) {
      const int rc = try_run_raster(L, BK, n_pad, maps, st, &used);
      if (rc || used) return rc;
    }
  }
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
      p.epi.out_y0[ph] = (int8_t)py;
      p.epi.out_x0[ph] = (int8_t)px;
    }
    p.epi.out_ys = 2; p.epi.out_xs = 2; p.epi.out_H = 2 * L.H; p.epi.out_W = 2 * L.W;
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
    p.epi.out_ys = 1; p.epi.out_xs = 1; p.epi.out_H = Ho * (L.epi.up2 ? 2 : 1); p.epi.out_W = Wo * (L.epi.up2 ? 2 : 1);
  }
  const int cpt = p.src_blocks[0] + (L.nsrc > 1 ? p.src_blocks[1] : 0);
  p.kblocks = p.ntaps * cpt;
  CIC_REQUIRE((long long)p.kblocks * BK == L.w.K, "tc layer: weight K=%d does not match taps x channels = %d", L.w.K,
              p.kblocks * BK);
  p.splits = L.splits;
  p.N = L.N;
  p.N_pad = dc ? L.w.rows / 4 : L.w.rows;
  p.b_batched = L.b_batched ? 1 : 0;
  {  // tile order: keep the larger operand's tiles adjacent in time so the smaller one is re-read from L2
    double act_elems = 0;
    for (int s = 0; s < L.nsrc; ++s) act_elems += (double)L.batch * L.H * L.W * L.src[s].C;
    const double w_elems = (double)L.w.rows * L.w.K * L.w.batches;
    p.m_fast = (!L.b_batched && w_elems > act_elems) ? 1 : 0;
  }
  // CTA-pair kernel (tc_gemm2.cu) for the large conv / transposed-conv GEMMs: halves the B fill and B operand reads per SM
  static const int pair_env = CIC_KNOB("CIC_TC_PAIR", 1);
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_b;
  const int bn2 = tc2_pick_block_n(p.N_pad);
  static const int dc2_env = CIC_KNOB("CIC_TC_DC2", 1);
  const bool dc2 = dc && dc2_env && pair_env && L.splits == 1 && tc_deconv2_ok(BK, L.split, m_tiles, p.N_pad, L.N);
  const bool pair = !dc2 && pair_env && !L.b_batched && tc_pair_ok(BK, m_tiles, p.N_pad, L.N);
  const int bn = dc2 ? L.N : (pair ? bn2 : tc_pick_block_n(p.N_pad, L.split, BK));
  CIC_REQUIRE(bn > 0 && L.N <= p.N_pad, "tc layer: N=%d (padded %d) has no supported tile", L.N, p.N_pad);
  CIC_REQUIRE(L.w.row_stride % 8 == 0 && L.w.batch_stride % 8 == 0, "tc layer: B rows must be 16-byte aligned");
  // weight / B map
  for (int part = 0; part < (L.split ? 2 : 1); ++part) {
    const uint64_t dims[3] = {(uint64_t)L.w.K, (uint64_t)L.w.rows, (uint64_t)L.w.batches};
    const uint64_t str[2] = {(uint64_t)L.w.row_stride * 2, (uint64_t)L.w.batch_stride * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)(pair || dc2 ? bn / 2 : bn), 1};
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
  TcEpi& pe = p.epi;
  pe.bias = e.bias; pe.scale = e.scale; pe.shift = e.shift; pe.alpha = e.alpha; pe.act = e.act;
  pe.out_mode = e.out_mode; pe.out_hi = e.out_hi; pe.out_lo = e.out_lo; pe.res_hi = e.res_hi; pe.res_lo = e.res_lo;
  pe.N = L.N; pe.out_ld = e.out_ld ? e.out_ld : L.N; pe.out_coff = e.out_coff; pe.up2 = e.up2;
  pe.Ho = Ho; pe.Wo = Wo;
  pe.tm_tx = e.tm_tx; pe.tm_ty = e.tm_ty; pe.tm_IH = e.tm_IH; pe.tm_IW = e.tm_IW;
  pe.m_total = (long long)L.batch * Ho * Wo;
  if (dc2) return launch_tc_deconv2(maps, p, st);
  if (pair) return launch_tc_gemm2(maps, p, bn, BK, L.split, st);
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
pack_weight_kernel(const float* __restrict__ src, int K, int N, int N_pad, int ld, bf16* __restrict__ hi, bf16* __restrict__ lo,
                   const float* __restrict__ col_scale) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + tx;
    tile[i][tx] = (k < K && n < N) ? (col_scale ? __fmul_rn(src[(size_t)k * ld + n], col_scale[n]) : src[(size_t)k * ld + n]) : 0.f;
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

int tc_pack_weight(const float* src, int K, int N, int N_pad, int ld, bf16* hi, bf16* lo, cudaStream_t st, const float* col_scale) {
  dim3 grid((K + 31) / 32, (N_pad + 31) / 32);
  CIC_REQUIRE(grid.y <= 65535, "pack_weight: N too large");
  pack_weight_kernel<<<grid, 256, 0, st>>>(src, K, N, N_pad, ld, hi, lo, col_scale);
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

__global__ void fuse_bias_kernel(const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                                 float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(__fmul_rn(bias ? bias[i] : 0.f, scale[i]), shift[i]);
}

int tc_fuse_bias(const float* bias, const float* scale, const float* shift, float* out, int n, cudaStream_t st) {
  CIC_REQUIRE(scale && shift && out && n > 0, "fuse_bias: null argument");
  fuse_bias_kernel<<<(n + 255) / 256, 256, 0, st>>>(bias, scale, shift, out, n);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("fuse_bias_kernel");
  return CIC_OK;
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

__global__ void __launch_bounds__(256)
maxpool2x2_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int batch, int H, int W, int C8) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)batch * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const uint4* p0 = x + (((long long)b * H + 2 * oy) * W + 2 * ox) * C8 + c8;
    const uint4 v[4] = {__ldg(p0), __ldg(p0 + C8), __ldg(p0 + (long long)W * C8), __ldg(p0 + (long long)W * C8 + C8)};
    uint4 m;
    uint32_t* mw = reinterpret_cast<uint32_t*>(&m);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&v[0]) + q);
#pragma unroll
      for (int k = 1; k < 4; ++k) a = __hmax2(a, *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&v[k]) + q));
      mw[q] = *reinterpret_cast<const uint32_t*>(&a);
    }
    y[i] = m;
  }
}

int tc_maxpool2x2_bf16(const bf16* x, bf16* y, int batch, int H, int W, int C, cudaStream_t st) {
  CIC_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_bf16: needs C %% 8 == 0 and even H, W");
  const long long total = (long long)batch * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return CIC_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  maxpool2x2_bf16_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), batch, H, W, C / 8);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("maxpool2x2_bf16_kernel");
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
