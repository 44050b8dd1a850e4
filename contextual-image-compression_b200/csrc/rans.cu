// Entropy coder for the quantised latent symbols: the data format on the far side of the path.
//
// The reference stops at the quantiser: its "bitrate" is the nominal 32 bits per latent element of GAN_test.py:310-325, no
// bitstream is ever produced.  This file turns the integer symbols round(latent * scale) (cic_quantize_latent /
// cic_adaptive_forward's d_*_symbols) into an actual byte stream and back, bit-exactly, on the GPU:
//
//   * static model per call: one histogram over all symbols (shared-memory histogram, global atomics), normalised to 2^14 by an
//     integer rule any implementation can repeat (f = max(1, floor(count * (M - K) / total)) for the K present symbols, the
//     remainder goes to the most frequent symbol);
//   * rANS with a 32-bit state and 16-bit renormalisation words; every row (one tile's latent vector) is an independent stream
//     coded by ONE WARP with 32 interleaved states: lane l codes symbols l, l + 32, ...; the lanes that renormalise in a step
//     place their words by ballot rank, so the decoder - also one warp per row - reads them back in lock step.  Rows decode in
//     parallel and in any order (tiles stay independently decodable, like the codec's tiles);
//   * container: 32-byte header, the 2047-entry frequency table, a row offset table, then per row 32 final states + its words.
// The CPU restatement of the same format (tests) produces the same bytes; decode(encode(x)) == x for |x| <= CIC_SYM_MAX.
#include "common.cuh"

namespace cic {

constexpr int RANS_PROB_BITS = 14;
constexpr uint32_t RANS_M = 1u << RANS_PROB_BITS;
constexpr uint32_t RANS_L = 1u << 16;                  // lower bound of the normalised state interval [L, L << 16)
constexpr int RANS_ALPHA = 2 * CIC_SYM_MAX + 1;        // 2047 symbols: -1023 .. 1023
constexpr int RANS_HEADER = 32;
constexpr int RANS_TABLE_BYTES = 4096;                 // 2047 x u16 + 2 bytes of padding
constexpr uint32_t RANS_MAGIC = 0x52434943u;           // "CICR"

__device__ __forceinline__ int rans_clamp_sym(int v) {
  v = v < -CIC_SYM_MAX ? -CIC_SYM_MAX : v;
  v = v > CIC_SYM_MAX ? CIC_SYM_MAX : v;
  return v + CIC_SYM_MAX;
}

__global__ void __launch_bounds__(256)
rans_hist_kernel(const int32_t* __restrict__ sym, size_t n, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[RANS_ALPHA];
  for (int i = threadIdx.x; i < RANS_ALPHA; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(&h[rans_clamp_sym(__ldg(sym + i))], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < RANS_ALPHA; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}

// One CTA: counts -> frequencies summing to 2^14 -> packed (freq | cum << 16) table + the stream header and frequency table.
__global__ void __launch_bounds__(1024)
rans_table_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ tab, uint8_t* __restrict__ stream, uint32_t rows, uint32_t L) {
  __shared__ unsigned long long total_s;
  __shared__ uint32_t present_s, fsum_s, best_s;
  __shared__ uint32_t f[RANS_ALPHA + 1];
  __shared__ uint32_t scan[1024];
  const int t = threadIdx.x;
  if (t == 0) { total_s = 0; present_s = 0; fsum_s = 0; best_s = 0; }
  __syncthreads();
  unsigned long long cnt[2] = {0, 0};
  for (int k = 0; k < 2; ++k) {
    const int s = t + 1024 * k;
    if (s < RANS_ALPHA) cnt[k] = hist[s];
  }
  atomicAdd(&total_s, cnt[0] + cnt[1]);
  atomicAdd(&present_s, (uint32_t)(cnt[0] > 0) + (uint32_t)(cnt[1] > 0));
  __syncthreads();
  const unsigned long long total = total_s;
  const uint32_t K = present_s;
  uint32_t fv[2] = {0, 0};
  for (int k = 0; k < 2; ++k)
    if (cnt[k] > 0) {
      const unsigned long long q = cnt[k] * (unsigned long long)(RANS_M - K) / total;
      fv[k] = q < 1 ? 1u : (uint32_t)q;
    }
  atomicAdd(&fsum_s, fv[0] + fv[1]);
  // most frequent symbol, lowest index on ties: max over (count, -index)
  {
    unsigned long long key = 0;
    for (int k = 0; k < 2; ++k) {
      const int s = t + 1024 * k;
      if (s < RANS_ALPHA && cnt[k] > 0) {
        const unsigned long long kk = (cnt[k] << 12) | (unsigned long long)(4095 - s);
        key = kk > key ? kk : key;
      }
    }
    // block max through shared atomics on two 32-bit halves is awkward: reduce with warp shuffles, then one atomicMax per warp
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    __shared__ unsigned long long wmax[32];
    if ((t & 31) == 0) wmax[t >> 5] = key;
    __syncthreads();
    if (t == 0) {
      unsigned long long m = 0;
      for (int w = 0; w < 32; ++w) m = wmax[w] > m ? wmax[w] : m;
      best_s = m ? (uint32_t)(4095 - (m & 4095)) : 0u;
    }
  }
  __syncthreads();
  for (int k = 0; k < 2; ++k) {
    const int s = t + 1024 * k;
    if (s < RANS_ALPHA) f[s] = fv[k] + ((uint32_t)s == best_s && total > 0 ? RANS_M - fsum_s : 0u);
  }
  if (t == 0 && total == 0) f[CIC_SYM_MAX] = RANS_M;     // empty input: a valid table all the same (symbol 0 with probability 1)
  __syncthreads();
  // exclusive prefix sum over 2047 entries: two entries per thread
  const uint32_t a0 = 2 * t < RANS_ALPHA ? f[2 * t] : 0u, a1 = 2 * t + 1 < RANS_ALPHA ? f[2 * t + 1] : 0u;
  scan[t] = a0 + a1;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = t >= o ? scan[t - o] : 0u;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  const uint32_t excl = scan[t] - (a0 + a1);
  if (2 * t < RANS_ALPHA) tab[2 * t] = a0 | (excl << 16);
  if (2 * t + 1 < RANS_ALPHA) tab[2 * t + 1] = a1 | ((excl + a0) << 16);
  uint16_t* ft = reinterpret_cast<uint16_t*>(stream + RANS_HEADER);
  if (2 * t < RANS_ALPHA) ft[2 * t] = (uint16_t)a0;                                 // f <= 2^14 fits the 16-bit field
  if (2 * t + 1 < RANS_ALPHA) ft[2 * t + 1] = (uint16_t)a1;
  if (t == 1023) ft[RANS_ALPHA] = 0;                                                // padding entry
  if (t == 0) {
    uint32_t* hd = reinterpret_cast<uint32_t*>(stream);
    hd[0] = RANS_MAGIC; hd[1] = 1u; hd[2] = rows; hd[3] = L; hd[4] = RANS_PROB_BITS; hd[5] = RANS_ALPHA; hd[6] = 0u; hd[7] = 0u;
  }
}

// One warp per row: 32 interleaved rANS states.  words: per-row slot of cap_words u16, filled from the END backwards.
__global__ void __launch_bounds__(128)
rans_encode_rows_kernel(const int32_t* __restrict__ sym, const uint32_t* __restrict__ tab, uint16_t* __restrict__ words, uint32_t* __restrict__ states,
                        uint32_t* __restrict__ n_words, int rows, int L, int cap_words) {
  __shared__ uint32_t stab[RANS_ALPHA];
  for (int i = threadIdx.x; i < RANS_ALPHA; i += blockDim.x) stab[i] = tab[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int32_t* rs = sym + (size_t)row * L;
  uint16_t* rw = words + (size_t)row * cap_words;
  uint32_t x = RANS_L;
  int ptr = cap_words;
  const int T = (L + 31) / 32;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int t = T - 1; t >= 0; --t) {
    const int idx = t * 32 + lane;
    const bool active = idx < L;
    uint32_t f = 1, c = 0;
    if (active) {
      const uint32_t e = stab[rans_clamp_sym(__ldg(rs + idx))];
      f = e & 0xffffu;
      c = e >> 16;
    }
    const bool emit = active && (unsigned long long)x >= ((unsigned long long)f << (32 - RANS_PROB_BITS));
    const uint32_t mask = __ballot_sync(0xffffffffu, emit);
    const int base = ptr - __popc(mask);
    if (emit) {
      rw[base + __popc(mask & lt_mask)] = (uint16_t)(x & 0xffffu);
      x >>= 16;
    }
    ptr = base;
    if (active) x = ((x / f) << RANS_PROB_BITS) + (x % f) + c;
  }
  states[(size_t)row * 32 + lane] = x;
  if (lane == 0) n_words[row] = (uint32_t)(cap_words - ptr);
}

// byte offsets of the rows in the payload (4-byte aligned rows): exclusive scan over rows, one CTA
__global__ void __launch_bounds__(1024)
rans_offsets_kernel(const uint32_t* __restrict__ n_words, uint32_t* __restrict__ offsets, unsigned long long* __restrict__ total_bytes, int rows,
                    unsigned long long payload_start) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  const int per = (rows + 1023) / 1024;
  const int lo = t * per, hi = min(lo + per, rows);
  unsigned long long s = 0;
  for (int r = lo; r < hi; ++r) s += 128ull + (((unsigned long long)n_words[r] * 2 + 3) & ~3ull);
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned long long v = t >= o ? part[t - o] : 0ull;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  unsigned long long run = part[t] - s;
  for (int r = lo; r < hi; ++r) {
    offsets[r] = (uint32_t)run;
    run += 128ull + (((unsigned long long)n_words[r] * 2 + 3) & ~3ull);
  }
  if (t == 1023) {
    offsets[rows] = (uint32_t)part[1023];
    *total_bytes = payload_start + part[1023];
  }
}

__global__ void __launch_bounds__(128)
rans_pack_kernel(const uint16_t* __restrict__ words, const uint32_t* __restrict__ states, const uint32_t* __restrict__ n_words,
                 const uint32_t* __restrict__ offsets, uint8_t* __restrict__ payload, int rows, int cap_words) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  uint8_t* dst = payload + offsets[row];
  reinterpret_cast<uint32_t*>(dst)[lane] = states[(size_t)row * 32 + lane];
  const uint32_t n = n_words[row];
  const uint16_t* src = words + (size_t)row * cap_words + (cap_words - n);
  uint16_t* dw = reinterpret_cast<uint16_t*>(dst + 128);
  for (uint32_t i = lane; i < n; i += 32) dw[i] = src[i];
  if ((n & 1u) && lane == 0) dw[n] = 0;               // padding to a 4-byte boundary
}

__global__ void __launch_bounds__(128)
rans_decode_rows_kernel(const uint8_t* __restrict__ stream, int32_t* __restrict__ out, int rows, int L) {
  extern __shared__ uint32_t rsm[];
  {  // a stream whose header does not describe (rows, L) is not walked at all: its offsets would be read out of bounds
    const uint32_t* hd = reinterpret_cast<const uint32_t*>(stream);
    if (hd[0] != RANS_MAGIC || hd[1] != 1u || hd[2] != (uint32_t)rows || hd[3] != (uint32_t)L || hd[4] != RANS_PROB_BITS || hd[5] != RANS_ALPHA) return;
  }
  uint32_t* stab = rsm;                                            // [2047] freq | cum << 16
  uint16_t* lut = reinterpret_cast<uint16_t*>(rsm + RANS_ALPHA + 1);   // [2^14] slot -> symbol
  __shared__ uint32_t scan[128];
  const uint16_t* ft = reinterpret_cast<const uint16_t*>(stream + RANS_HEADER);
  // cumulative table: 2047 entries, 16 per thread
  const int t = threadIdx.x;
  uint32_t loc[16], s = 0;
  for (int k = 0; k < 16; ++k) {
    const int i = t * 16 + k;
    loc[k] = i < RANS_ALPHA ? ft[i] : 0u;
    s += loc[k];
  }
  scan[t] = s;
  __syncthreads();
  for (int o = 1; o < 128; o <<= 1) {
    const uint32_t v = t >= o ? scan[t - o] : 0u;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  uint32_t run = scan[t] - s;
  for (int k = 0; k < 16; ++k) {
    const int i = t * 16 + k;
    if (i < RANS_ALPHA) {
      stab[i] = loc[k] | (run << 16);
      run += loc[k];
    }
  }
  __syncthreads();
  // slot -> symbol: the last symbol whose cumulative count is <= slot (a symbol with frequency 0 shares its cumulative count with
  // the next present one, which has the larger index, so the upper bound is always a present symbol)
  for (uint32_t slot = t; slot < RANS_M; slot += blockDim.x) {
    int lo = 0, hi = RANS_ALPHA - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((stab[mid] >> 16) <= slot) lo = mid; else hi = mid - 1;
    }
    lut[slot] = (uint16_t)lo;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint32_t* offsets = reinterpret_cast<const uint32_t*>(stream + RANS_HEADER + RANS_TABLE_BYTES);
  const uint8_t* payload = stream + RANS_HEADER + RANS_TABLE_BYTES + (((size_t)rows + 1) * 4);
  const uint8_t* src = payload + offsets[row];
  uint32_t x = reinterpret_cast<const uint32_t*>(src)[lane];
  const uint16_t* rw = reinterpret_cast<const uint16_t*>(src + 128);
  int rp = 0;
  const int T = (L + 31) / 32;
  const uint32_t lt_mask = (1u << lane) - 1u;
  int32_t* ro = out + (size_t)row * L;
  for (int tt = 0; tt < T; ++tt) {
    const int idx = tt * 32 + lane;
    const bool active = idx < L;
    if (active) {
      const uint32_t slot = x & (RANS_M - 1u);
      const uint32_t sy = lut[slot];
      const uint32_t e = stab[sy];
      x = (e & 0xffffu) * (x >> RANS_PROB_BITS) + slot - (e >> 16);
      ro[idx] = (int32_t)sy - CIC_SYM_MAX;
    }
    const bool need = active && x < RANS_L;
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (need) x = (x << 16) | rw[rp + __popc(mask & lt_mask)];
    rp += __popc(mask);
  }
}

static size_t rans_align(size_t n) { return (n + 255) & ~(size_t)255; }
static int rans_cap_words(int L) { return (L + 1) & ~1; }   // at most one 16-bit word per symbol

}  // namespace cic

using namespace cic;

extern "C" size_t cic_rans_max_bytes(int rows, int latent_dim) {
  if (rows < 0 || latent_dim <= 0) return 0;
  return (size_t)RANS_HEADER + RANS_TABLE_BYTES + ((size_t)rows + 1) * 4 + (size_t)rows * (128 + (size_t)rans_cap_words(latent_dim) * 2) + 16;
}

extern "C" size_t cic_rans_workspace_bytes(int rows, int latent_dim) {
  if (rows < 0 || latent_dim <= 0) return 0;
  return rans_align(RANS_ALPHA * 4) * 2 + rans_align((size_t)rows * rans_cap_words(latent_dim) * 2) + rans_align((size_t)rows * 128) +
         rans_align((size_t)rows * 4 + 4) + 1024;
}

extern "C" int cic_rans_encode(const int32_t* d_symbols, int rows, int latent_dim, uint8_t* d_stream, size_t stream_capacity,
                               unsigned long long* d_nbytes, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(rows >= 0 && latent_dim > 0, "cic_rans_encode: bad shape");
  CIC_REQUIRE(d_stream && d_nbytes && (rows == 0 || d_symbols), "cic_rans_encode: null pointer");
  CIC_REQUIRE(stream_capacity >= cic_rans_max_bytes(rows, latent_dim), "cic_rans_encode: stream buffer smaller than cic_rans_max_bytes");
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_rans_workspace_bytes(rows, latent_dim), "cic_rans_encode: workspace too small");
  CIC_REQUIRE(((uintptr_t)d_stream & 3) == 0, "cic_rans_encode: stream buffer must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_workspace;
  uint32_t* hist = (uint32_t*)ws; ws += rans_align(RANS_ALPHA * 4);
  uint32_t* tab = (uint32_t*)ws; ws += rans_align(RANS_ALPHA * 4);
  const int capw = rans_cap_words(latent_dim);
  uint16_t* words = (uint16_t*)ws; ws += rans_align((size_t)rows * capw * 2);
  uint32_t* states = (uint32_t*)ws; ws += rans_align((size_t)rows * 128);
  uint32_t* n_words = (uint32_t*)ws;
  const size_t n = (size_t)rows * latent_dim;
  CIC_CHECK_CUDA(cudaMemsetAsync(hist, 0, RANS_ALPHA * 4, st));
  if (n) {
    const size_t want = (n + 255) / 256;
    const int blocks = (int)(want < (size_t)sm_count() * 8 ? want : (size_t)sm_count() * 8);
    rans_hist_kernel<<<blocks, 256, 0, st>>>(d_symbols, n, hist);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("rans_hist_kernel");
  }
  rans_table_kernel<<<1, 1024, 0, st>>>(hist, tab, d_stream, (uint32_t)rows, (uint32_t)latent_dim);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("rans_table_kernel");
  uint32_t* offsets = reinterpret_cast<uint32_t*>(d_stream + RANS_HEADER + RANS_TABLE_BYTES);
  const unsigned long long payload_start = (unsigned long long)RANS_HEADER + RANS_TABLE_BYTES + ((unsigned long long)rows + 1) * 4;
  if (rows) {
    rans_encode_rows_kernel<<<(rows + 3) / 4, 128, 0, st>>>(d_symbols, tab, words, states, n_words, rows, latent_dim, capw);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("rans_encode_rows_kernel");
  }
  rans_offsets_kernel<<<1, 1024, 0, st>>>(n_words, offsets, d_nbytes, rows, payload_start);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("rans_offsets_kernel");
  if (rows) {
    rans_pack_kernel<<<(rows + 3) / 4, 128, 0, st>>>(words, states, n_words, offsets, d_stream + payload_start, rows, capw);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("rans_pack_kernel");
  }
  return CIC_OK;
}

extern "C" int cic_rans_decode(const uint8_t* d_stream, size_t nbytes, int32_t* d_symbols, int rows, int latent_dim, void* stream) {
  CIC_REQUIRE(rows >= 0 && latent_dim > 0, "cic_rans_decode: bad shape");
  CIC_REQUIRE(d_stream && (rows == 0 || d_symbols), "cic_rans_decode: null pointer");
  CIC_REQUIRE(nbytes >= (size_t)RANS_HEADER + RANS_TABLE_BYTES + ((size_t)rows + 1) * 4, "cic_rans_decode: stream shorter than its header");
  CIC_REQUIRE(((uintptr_t)d_stream & 3) == 0, "cic_rans_decode: stream buffer must be 4-byte aligned");
  if (rows == 0) return CIC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(RANS_ALPHA + 1) * 4 + (size_t)RANS_M * 2;
  static DeviceOnce attr_set;
  if (attr_set.todo()) {
    CIC_CHECK_CUDA(cudaFuncSetAttribute(rans_decode_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.done();
  }
  rans_decode_rows_kernel<<<(rows + 3) / 4, 128, smem, st>>>(d_stream, d_symbols, rows, latent_dim);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("rans_decode_rows_kernel");
  return CIC_OK;
}
