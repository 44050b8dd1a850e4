// Fused SelfAttention core (GAN_functions.py:353-367): y = gamma * softmax(q k^T) v + x for 32x32-token images, on
// split-bf16 operands (3-term products, fp32 accumulate in TMEM), without materialising the tokens x tokens map.
//
// The unfused path writes the fp32 scores (4 MB per image) and the (hi, lo) probabilities (4 MB) to HBM and reads them
// back: 4.3 GB per 256 images, five launches, 1.5 ms.  Here one CTA owns 128 query rows of one image and streams the
// keys in blocks of 64:
//   pass 1: S_j = Q K_j^T (TMEM, 64 columns, double-buffered) -> running row maximum (no 1/sqrt(d) scale, :358);
//   pass 2: S_j again (K = 32: recomputing costs 1/8 of P V), p = exp(S_j - max) -> (hi, lo) bf16 written to shared
//           memory as the swizzled K-major A operand, row sums in registers, O += P_j V_j (TMEM, 256 columns);
//   epilogue: y = gamma * O / rowsum + x  ->  (hi, lo) bf16.
// Two passes instead of an online softmax: no rescaling of O in TMEM, and the probabilities are final when they are
// rounded to bf16 pairs.  Warps: 0 = TMA producer (Q, K blocks, V^T blocks), 1 = TMEM allocator + MMA issuer,
// 2-5 and 6-9 = two softmax / epilogue groups on alternate key blocks (thread = query row = TMEM lane).
#include "tc_gemm.cuh"
#include "tc_host.cuh"

#include <cstring>

namespace cic {

constexpr int AT_Q = 128;    // query rows per CTA tile
constexpr int AT_KB = 64;    // keys per block
constexpr int AT_D = 32;     // q / k channels
constexpr int AT_C = 256;    // value channels
constexpr int AT_QBYTES = AT_Q * AT_D * 2;     // one part of the Q tile (8 KB, SWIZZLE_64B rows of 64 B)
constexpr int AT_KBYTES = AT_KB * AT_D * 2;    // one part of a K block (4 KB)
constexpr int AT_VBYTES = AT_C * AT_KB * 2;    // one part of a V^T block (32 KB, SWIZZLE_128B rows of 128 B)
constexpr int AT_PBYTES = AT_Q * AT_KB * 2;    // one part of a P block (16 KB, SWIZZLE_128B)
constexpr int AT_SMEM = 2 * AT_QBYTES + 2 * 2 * AT_KBYTES + 2 * 2 * AT_VBYTES + 2 * 2 * AT_PBYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*softmax exchange*/;
constexpr uint32_t AT_S_COLS = 64, AT_O_COL0 = 128;

struct AttnMaps {
  CUtensorMap q[2], k[2], v[2];  // [hi, lo]
};

struct AttnParams {
  int nb, tokens;
  TcEpi epi;  // alpha = gamma, residual x, output y (rows = nb * tokens, 256 channels)
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync_at(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t at_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(320, 1)
attn_fused_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;                            // [hi | lo]
  uint8_t* k_s = q_s + 2 * AT_QBYTES;             // [stage][hi | lo]
  uint8_t* v_s = k_s + 2 * 2 * AT_KBYTES;         // [stage][hi | lo]
  uint8_t* p_s = v_s + 2 * 2 * AT_VBYTES;         // [stage][hi | lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * 2 * AT_PBYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* q_empty = bars + 1;       // [1]
  uint64_t* k_full = bars + 2;        // [2]
  uint64_t* k_empty = bars + 4;       // [2]
  uint64_t* v_full = bars + 6;        // [2]
  uint64_t* v_empty = bars + 8;       // [2]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* s_empty = bars + 12;      // [2]
  uint64_t* p_full = bars + 14;       // [2]
  uint64_t* p_empty = bars + 16;      // [2]
  uint64_t* o_full = bars + 18;       // [1]
  uint64_t* o_empty = bars + 19;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* xch = reinterpret_cast<float*>(bars + 22);  // [2 groups][128 rows]: row maxima / row sums exchanged between the softmax groups

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_q = p.tokens / AT_Q, nblk = p.tokens / AT_KB;
  const int items = p.nb * tiles_q;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) { prefetch_tmap(&maps.q[i]); prefetch_tmap(&maps.k[i]); prefetch_tmap(&maps.v[i]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1); mbar_init(q_empty, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
        mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4);
        mbar_init(&p_full[s], 4); mbar_init(&p_empty[s], 1);
      }
      mbar_init(o_full, 1); mbar_init(o_empty, 8);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      uint32_t kc = 0, vc = 0, ic = 0;  // K blocks, V blocks, items issued so far
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++ic) {
        const int b = it / tiles_q, qt = it - b * tiles_q;
        const int row0 = b * p.tokens;
        mbar_wait_relaxed(q_empty, (ic & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full, 2 * AT_QBYTES);
        tma_load_2d(q_s, &maps.q[0], q_full, 0, row0 + qt * AT_Q);
        tma_load_2d(q_s + AT_QBYTES, &maps.q[1], q_full, 0, row0 + qt * AT_Q);
        for (int pass = 0; pass < 2; ++pass) {
          for (int j = 0; j < nblk; ++j) {
            const uint32_t ks = kc & 1u;
            mbar_wait_relaxed(&k_empty[ks], ((kc >> 1) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&k_full[ks], 2 * AT_KBYTES);
            uint8_t* kd = k_s + ks * 2 * AT_KBYTES;
            tma_load_2d(kd, &maps.k[0], &k_full[ks], AT_D, row0 + j * AT_KB);
            tma_load_2d(kd + AT_KBYTES, &maps.k[1], &k_full[ks], AT_D, row0 + j * AT_KB);
            ++kc;
            if (pass == 1) {
              const uint32_t vs = vc & 1u;
              mbar_wait_relaxed(&v_empty[vs], ((vc >> 1) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&v_full[vs], 2 * AT_VBYTES);
              uint8_t* vd = v_s + vs * 2 * AT_VBYTES;
              tma_load_3d(vd, &maps.v[0], &v_full[vs], j * AT_KB, 0, b);
              tma_load_3d(vd + AT_VBYTES, &maps.v[1], &v_full[vs], j * AT_KB, 0, b);
              ++vc;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc_s = at_idesc(AT_KB), idesc_o = at_idesc(AT_C);
      const uint32_t q_lo16 = (smem_u32(q_s) & 0x3FFFF) >> 4, k_lo16 = (smem_u32(k_s) & 0x3FFFF) >> 4;
      const uint32_t v_lo16 = (smem_u32(v_s) & 0x3FFFF) >> 4, p_lo16 = (smem_u32(p_s) & 0x3FFFF) >> 4;
      uint32_t kc = 0, vc = 0, sc = 0, pc = 0, ic = 0;
      // S_j = Q K_j^T into S stage (sc & 1): three split terms x two K steps
      auto issue_s = [&]() {
        const uint32_t ks = kc & 1u, ss = sc & 1u;
        mbar_wait(&k_full[ks], (kc >> 1) & 1u);
        mbar_wait(&s_empty[ss], ((sc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + ss * AT_S_COLS;
        const uint32_t qh = q_lo16, ql = q_lo16 + (AT_QBYTES >> 4);
        const uint32_t kh = k_lo16 + ks * (2 * AT_KBYTES >> 4), kl = kh + (AT_KBYTES >> 4);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(d, umma_desc_from_lo<32>(qh + 2 * k), umma_desc_from_lo<32>(kh + 2 * k), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(d, umma_desc_from_lo<32>(ql + 2 * k), umma_desc_from_lo<32>(kh + 2 * k), idesc_s, 1u);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(d, umma_desc_from_lo<32>(qh + 2 * k), umma_desc_from_lo<32>(kl + 2 * k), idesc_s, 1u);
        umma_commit(&k_empty[ks]);
        umma_commit(&s_full[ss]);
        ++kc;
        ++sc;
      };
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++ic) {
        mbar_wait(q_full, ic & 1u);
        tc_fence_after();
        for (int j = 0; j < nblk; ++j) issue_s();  // pass 1: scores for the row maxima
        issue_s();                                 // pass 2, block 0
        for (int j = 0; j < nblk; ++j) {
          if (j + 1 < nblk) {
            issue_s();                             // the scores of the next block are computed while this block's exp runs
            if (j + 2 == nblk) umma_commit(q_empty);  // last score MMA of the item issued: Q is free once it retires
          }
          const uint32_t ps = pc & 1u, vs = vc & 1u;
          mbar_wait(&p_full[ps], (pc >> 1) & 1u);
          mbar_wait(&v_full[vs], (vc >> 1) & 1u);
          if (j == 0) mbar_wait(o_empty, (ic & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d = tmem_base + AT_O_COL0;
          const uint32_t ph = p_lo16 + ps * (2 * AT_PBYTES >> 4), pl = ph + (AT_PBYTES >> 4);
          const uint32_t vh = v_lo16 + vs * (2 * AT_VBYTES >> 4), vl = vh + (AT_VBYTES >> 4);
#pragma unroll
          for (int k = 0; k < AT_KB / 16; ++k) umma_bf16(d, umma_desc_from_lo<64>(ph + 2 * k), umma_desc_from_lo<64>(vh + 2 * k), idesc_o, (j | k) != 0);
#pragma unroll
          for (int k = 0; k < AT_KB / 16; ++k) umma_bf16(d, umma_desc_from_lo<64>(pl + 2 * k), umma_desc_from_lo<64>(vh + 2 * k), idesc_o, 1u);
#pragma unroll
          for (int k = 0; k < AT_KB / 16; ++k) umma_bf16(d, umma_desc_from_lo<64>(ph + 2 * k), umma_desc_from_lo<64>(vl + 2 * k), idesc_o, 1u);
          umma_commit(&v_empty[vs]);
          umma_commit(&p_empty[ps]);
          ++pc;
          ++vc;
        }
        umma_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    // ===== softmax + epilogue: thread = query row; two groups of four warps take alternate key blocks (group g owns S
    // stage g and P stage g) - one warp per scheduler runs the ~1000-instruction block at ~5 cycles per instruction =====
    const int g = warp >= 6 ? 1 : 0;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)g * AT_S_COLS;
    const uint32_t ph = smem_u32(p_s) + (uint32_t)g * 2 * AT_PBYTES + (uint32_t)r * 128u, pl = ph + AT_PBYTES;
    const uint32_t sw = (uint32_t)(r & 7);
    const float kLog2e = 1.4426950408889634f;
    uint32_t gc = 0, gp = 0, ic = 0;  // S blocks / P blocks handled by this group, items
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++ic) {
      const int b = it / tiles_q, qt = it - b * tiles_q;
      // pass 1: row maximum over this group's key blocks
      float m = -INFINITY;
      for (int j = g; j < nblk; j += 2, ++gc) {
        mbar_wait_relaxed(&s_full[g], gc & 1u);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        __syncwarp();
        tmem_ld32(s_addr, v0);
        tmem_ld32(s_addr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[g]);
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
      }
      xch[g * AT_Q + r] = m;
      named_bar_sync_at(1, 256);
      m = fmaxf(m, xch[(g ^ 1) * AT_Q + r]);
      // pass 2: probabilities (unnormalised) as bf16 (hi, lo) rows of the P operand, row sum
      const float mb = m * kLog2e;
      float l = 0.f;
      for (int j = g; j < nblk; j += 2, ++gc, ++gp) {
        mbar_wait_relaxed(&s_full[g], gc & 1u);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        __syncwarp();
        tmem_ld32(s_addr, v0);
        tmem_ld32(s_addr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[g]);
        uint32_t hi[32], lo[32];  // 64 keys -> 32 packed pairs each
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s0 = __uint_as_float(i < 16 ? v0[2 * i] : v1[2 * i - 32]);
          const float s1 = __uint_as_float(i < 16 ? v0[2 * i + 1] : v1[2 * i - 31]);
          const float e0 = exp2f(fmaf(s0, kLog2e, -mb)), e1 = exp2f(fmaf(s1, kLog2e, -mb));
          l += e0;
          l += e1;
          const __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
          hi[i] = *reinterpret_cast<const uint32_t*>(&h);
          const __nv_bfloat162 lw = __floats2bfloat162_rn(e0 - __uint_as_float(hi[i] << 16), e1 - __uint_as_float(hi[i] & 0xFFFF0000u));
          lo[i] = *reinterpret_cast<const uint32_t*>(&lw);
        }
        mbar_wait_relaxed(&p_empty[g], (gp & 1u) ^ 1u);  // the MMAs that read this buffer two blocks ago have retired
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 16-byte chunk c (keys 8c .. 8c+7) sits at chunk position c ^ (row & 7) (SWIZZLE_128B)
          const uint32_t off = ((uint32_t)c ^ sw) << 4;
          st_shared_v4(ph + off, hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          st_shared_v4(pl + off, lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      named_bar_sync_at(2, 256);   // the other group has read this group's row maximum
      xch[g * AT_Q + r] = l;
      named_bar_sync_at(1, 256);
      l += xch[(g ^ 1) * AT_Q + r];
      // epilogue: y = gamma * O / l + x; the two groups take alternate 32-column chunks
      const float inv_l = 1.f / l;
      mbar_wait_relaxed(o_full, ic & 1u);
      tc_fence_after();
      const TcRow row{b, qt * AT_Q + r, 0, 0, 0};
#pragma unroll 1
      for (int c = g; c < AT_C / 32; c += 2) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(lane_addr + AT_O_COL0 + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (c + 2 >= AT_C / 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__fmul_rn(__uint_as_float(v[i]), inv_l));
        tc_epilogue_store<32>(p.epi, TcEpiVec{p.epi.bias, p.epi.scale, p.epi.shift}, row, v, c * 32, 32, true, 0);
      }
      named_bar_sync_at(2, 256);   // row sums consumed before the next item overwrites the exchange buffer
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// qk: (hi, lo) [nb * tokens][2 * 32] (q | k); vt: (hi, lo) [nb][256][tokens]; x / y: (hi, lo) [nb * tokens][256]
int launch_attn_fused(const bf16* qk_hi, const bf16* qk_lo, const bf16* vt_hi, const bf16* vt_lo, const bf16* x_hi, const bf16* x_lo,
                      bf16* y_hi, bf16* y_lo, const float* bias_v_scaled, float gamma, int nb, int tokens, cudaStream_t st) {
  CIC_REQUIRE(tokens % AT_Q == 0 && tokens >= AT_Q, "attn_fused: token count %d must be a multiple of %d", tokens, AT_Q);
  if (nb == 0) return CIC_OK;
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int part = 0; part < 2; ++part) {
    const uint64_t qdims[2] = {2 * AT_D, (uint64_t)nb * tokens};
    const uint64_t qstr[1] = {2 * AT_D * 2};
    const uint32_t qbox[2] = {AT_D, AT_Q}, kbox[2] = {AT_D, AT_KB};
    int rc = tc_encode_map(&maps.q[part], part ? qk_lo : qk_hi, 2, qdims, qstr, qbox);
    if (rc) return rc;
    if ((rc = tc_encode_map(&maps.k[part], part ? qk_lo : qk_hi, 2, qdims, qstr, kbox))) return rc;
    const uint64_t vdims[3] = {(uint64_t)tokens, AT_C, (uint64_t)nb};
    const uint64_t vstr[2] = {(uint64_t)tokens * 2, (uint64_t)AT_C * tokens * 2};
    const uint32_t vbox[3] = {AT_KB, AT_C, 1};
    if ((rc = tc_encode_map(&maps.v[part], part ? vt_lo : vt_hi, 3, vdims, vstr, vbox))) return rc;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.nb = nb;
  p.tokens = tokens;
  TcEpi& e = p.epi;
  e.bias = bias_v_scaled;  // gamma * b_v: softmax rows sum to 1, so the value bias passes through the product
  e.alpha = gamma;
  e.act = CIC_ACT_NONE;
  e.out_mode = TC_OUT_BF16;
  e.out_hi = y_hi; e.out_lo = y_lo; e.res_hi = x_hi; e.res_lo = x_lo;
  e.N = AT_C; e.out_ld = AT_C; e.out_coff = 0;
  e.out_H = tokens; e.out_W = 1; e.out_ys = 1; e.out_xs = 1;
  e.Ho = tokens; e.Wo = 1;
  e.m_total = (long long)nb * tokens;
  static DeviceOnce attr_set;  // function attributes are per device
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set.done();
  }
  const int items = nb * (tokens / AT_Q);
  const int grid = items < sm_count() ? items : sm_count();
  attn_fused_kernel<<<grid, 320, AT_SMEM, st>>>(maps, p);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("attn_fused_kernel");
  g_last_kernel_kind = KK_ATTN;
  return CIC_OK;
}

}  // namespace cic
