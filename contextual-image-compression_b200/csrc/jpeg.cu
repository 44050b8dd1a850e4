// Baseline JPEG encoder on the GPU: the output stage of the reference's path.  The reference writes every
// compressed image with cv2.imwrite("*.jpg", uint8 BGR) (test_autoencoder.py:88-93; GAN_functions.py:41-50 save_image, called at
// GAN_test.py:390); this file produces the SAME FILE BYTES from a uint8 image batch that is already in HBM, so only the compressed
// bytes cross PCIe.
//
// What "the same bytes" means: OpenCV's JPEG writer is libjpeg(-turbo) with its defaults - baseline sequential DCT, quality 95,
// 4:2:0 chroma, the Annex K Huffman tables, no restart markers, JFIF 1.01 header.  Every arithmetic step below follows the
// published libjpeg design (IJG libjpeg 6b; libjpeg-turbo reproduces it bit for bit):
//   colour     Y / Cb / Cr in 16-bit fixed point, rounding constants ONE_HALF and ONE_HALF - 1                     (jccolor.c)
//   chroma     2x2 box, bias alternating 1, 2 along a row; right edge replicated in the input, bottom edge replicated in the input up
//              to an even row count and in the down-sampled rows below that                                        (jcsample.c, jcprepct.c)
//   DCT        the accurate integer 8x8 DCT: 13-bit constants, 2 guard bits after the row pass, output scaled by 8  (jfdctint.c)
//   quantise   divide by 8 q, rounding half away from zero (exact reciprocal multiply here)                        (jcdctmgr.c)
//   dummies    Y blocks completing an MCU beyond the component's block grid: AC = 0, DC = DC of the previous block  (jccoefct.c)
//   entropy    DC prediction per component, run / size symbols, ZRL, EOB, 0xFF stuffing, 1-bit padding             (jchuff.c)
// The tests compare the output byte for byte with cv2.imencode (the real library) on the GPU box and with the numpy restatement.
//
// Kernels (all HBM-bound integer / byte work, no tensor cores):
//   jpeg_dct_kernel      one CTA per 8 MCUs of a row: BGR bytes -> Y/Cb/Cr in shared memory -> 48 blocks, row pass and column pass as
//                        384 one-dimensional DCTs each -> quantise, zig-zag, dummy rule -> int16 coefficients, coalesced
//   jpeg_len_kernel      one thread per block (the CTA's 128 blocks staged in shared memory with coalesced loads): bit length of its code
//   jpeg_scan_kernel     per-image exclusive scan (one CTA per image) -> bit offset of every block; zeroes the words two pack CTAs share
//   jpeg_pack_kernel     one CTA per 128 blocks: every thread ORs its code words into the CTA's span in shared memory, the span goes
//                        to HBM with coalesced word stores (the two boundary words by atomicOr)
//   jpeg_ffcount_kernel / jpeg_scan_kernel / jpeg_stuff_kernel   0xFF stuffing: count per 32-byte chunk, scan, then every CTA
//                        assembles its 8 KB of the file in shared memory and stores whole words behind the header
#include "common.cuh"

#include <initializer_list>
#include <string.h>

namespace cic {

constexpr int JPEG_HEADER = 623;
constexpr int JPEG_BLOCK_MAX_BYTES = 208;   // (9 + 11) + 63 * (16 + 10) = 1658 bits
constexpr int JPEG_PACK_N = 128;            // blocks per pack CTA
constexpr int JPEG_CHUNK = 32;              // bytes per stuffing chunk

struct JpegTables { uint32_t dc[2][16]; uint32_t ac[2][256]; };   // (code << 5) | length
struct JpegQuant { uint16_t q[2][64]; };                          // natural order
struct JpegHeader { uint8_t b[JPEG_HEADER + 1]; };

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one pass of jfdctint.c over 8 values; FIRST: row pass (results keep 2 extra bits), else column pass
template <bool FIRST>
__device__ __forceinline__ void dct8(int* d) {
  constexpr int N = FIRST ? 13 - 2 : 13 + 2;
  const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
  const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  if (FIRST) {
    d[0] = (t10 + t11) << 2;
    d[4] = (t10 - t11) << 2;
  } else {
    d[0] = descale(t10 + t11, 2);
    d[4] = descale(t10 - t11, 2);
  }
  int z1 = (t12 + t13) * 4433;
  d[2] = descale(z1 + t13 * 6270, N);
  d[6] = descale(z1 - t12 * 15137, N);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * 9633;
  const int u4 = t4 * 2446, u5 = t5 * 16819, u6 = t6 * 25172, u7 = t7 * 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 = z3 * -16069 + z5;
  z4 = z4 * -3196 + z5;
  d[7] = descale(u4 + z1 + z3, N);
  d[5] = descale(u5 + z2 + z4, N);
  d[3] = descale(u6 + z2 + z3, N);
  d[1] = descale(u7 + z1 + z4, N);
}

template <bool RGB>
__global__ void __launch_bounds__(256)
jpeg_dct_kernel(const uint8_t* __restrict__ img, int16_t* __restrict__ coef, int H, int W, int mw, int mh, int words_ok, JpegQuant qt) {
  __shared__ uint8_t sY[16][128], sCb[16][128], sCr[16][128];
  __shared__ int s[48][72];
  __shared__ uint32_t recip[2][64];
  __shared__ uint16_t sq[2][64];
  __shared__ uint8_t zz[64];
  const int b = blockIdx.z, my = blockIdx.y, mx0 = blockIdx.x * 8, tid = threadIdx.x;
  if (tid < 128) {
    const int t = tid >> 6, i = tid & 63;
    const uint32_t q = (uint32_t)qt.q[t][i] << 3;
    sq[t][i] = (uint16_t)q;
    recip[t][i] = (uint32_t)((0x100000000ull + q - 1) / q);       // floor(a / q) == umulhi(a, recip) for a * q < 2^32
  }
  if (tid < 64) zz[tid] = c_zigzag[tid];
  const uint8_t* src = img + (size_t)b * H * W * 3;
  auto convert = [&](int ly, int lx, int c0, int c1, int c2) {
    const int r = RGB ? c0 : c2, g = c1, bl = RGB ? c2 : c0;
    sY[ly][lx] = (uint8_t)((19595 * r + 38470 * g + 7471 * bl + 32768) >> 16);
    sCb[ly][lx] = (uint8_t)((-11059 * r - 21709 * g + 32768 * bl + (128 << 16) + 32767) >> 16);
    sCr[ly][lx] = (uint8_t)((32768 * r - 27439 * g - 5329 * bl + (128 << 16) + 32767) >> 16);
  };
  if (words_ok && mx0 * 16 + 128 <= W) {
    // the CTA's 128-pixel row segments are whole, 4-byte aligned runs of 96 words: stage the raw bytes with word loads
    uint32_t* sraw = reinterpret_cast<uint32_t*>(&s[0][0]);       // 16 x 96 words, reused by the DCT afterwards
    for (int i = tid; i < 16 * 96; i += 256) {
      const int ly = i / 96, wx = i - ly * 96;
      const int y = min(my * 16 + ly, H - 1);
      sraw[i] = __ldg(reinterpret_cast<const uint32_t*>(src + ((size_t)y * W + mx0 * 16) * 3) + wx);
    }
    __syncthreads();
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(sraw);
    for (int i = tid; i < 16 * 128; i += 256) {
      const int ly = i >> 7, lx = i & 127;
      const uint8_t* p = sb + ly * 384 + lx * 3;
      convert(ly, lx, p[0], p[1], p[2]);
    }
  } else {
    for (int i = tid; i < 16 * 128; i += 256) {
      const int ly = i >> 7, lx = i & 127;
      const int y = min(my * 16 + ly, H - 1), x = min(mx0 * 16 + lx, W - 1);
      const uint8_t* p = src + ((size_t)y * W + x) * 3;
      convert(ly, lx, __ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
  }
  __syncthreads();
  for (int i = tid; i < 8 * 4 * 64; i += 256) {
    const int m = i >> 8, j = (i >> 6) & 3, ry = (i >> 3) & 7, rx = i & 7;
    s[m * 6 + j][ry * 9 + rx] = (int)sY[(j >> 1) * 8 + ry][m * 16 + (j & 1) * 8 + rx] - 128;
  }
  const int ch = (H + 1) >> 1;                                    // down-sampled rows that exist; below them the last one repeats
  for (int i = tid; i < 8 * 64; i += 256) {
    const int m = i >> 6, ry = (i >> 3) & 7, rx = i & 7;
    const int ly = (min(my * 8 + ry, ch - 1) - my * 8) * 2, lx = m * 16 + rx * 2;
    const int bias = 1 + (rx & 1);
    s[m * 6 + 4][ry * 9 + rx] = ((sCb[ly][lx] + sCb[ly][lx + 1] + sCb[ly + 1][lx] + sCb[ly + 1][lx + 1] + bias) >> 2) - 128;
    s[m * 6 + 5][ry * 9 + rx] = ((sCr[ly][lx] + sCr[ly][lx + 1] + sCr[ly + 1][lx] + sCr[ly + 1][lx + 1] + bias) >> 2) - 128;
  }
  __syncthreads();
  for (int t = tid; t < 384; t += 256) {                          // rows
    int* p = &s[t >> 3][(t & 7) * 9];
    int d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = p[k];
    dct8<true>(d);
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = d[k];
  }
  __syncthreads();
  for (int t = tid; t < 384; t += 256) {                          // columns
    int* p = &s[t >> 3][t & 7];
    int d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = p[k * 9];
    dct8<false>(d);
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k * 9] = d[k];
  }
  __syncthreads();
  const int nm = min(8, mw - mx0);
  const bool bottom = (((H + 7) >> 3) & 1) && my == mh - 1;       // the MCU row's second Y block row does not exist
  const bool right_odd = ((W + 7) >> 3) & 1;
  int16_t* dst = coef + (((size_t)b * mh * mw + (size_t)my * mw + mx0) * 6) * 64;
  // thread = one zig-zag position k of every fourth block: the table entries are loaded once
  const int k = tid & 63, nat_k = zz[k], pos_k = (nat_k >> 3) * 9 + (nat_k & 7);
  const uint32_t rq[2] = {recip[0][nat_k], recip[1][nat_k]}, hq[2] = {(uint32_t)sq[0][nat_k] >> 1, (uint32_t)sq[1][nat_k] >> 1};
  const uint32_t rq0[2] = {recip[0][0], recip[1][0]}, hq0[2] = {(uint32_t)sq[0][0] >> 1, (uint32_t)sq[1][0] >> 1};
  for (int blk = tid >> 6; blk < nm * 6; blk += 4) {
    const int m = blk / 6, j = blk - m * 6, t = j < 4 ? 0 : 1;
    int v = s[blk][pos_k];
    uint32_t r = rq[t], h = hq[t];
    bool zero = false;
    if ((bottom || right_odd) && j < 4) {                         // dummy blocks (jccoefct.c): AC = 0, DC = DC of the previous block
      const bool right = right_odd && (mx0 + m == mw - 1);
      int srcj = j;
      bool dummy = false;
      if (bottom && j >= 2) { srcj = 1; dummy = true; }
      else if (right && (j & 1)) { srcj = j - 1; dummy = true; }
      if (srcj == 1 && right) srcj = 0;
      if (dummy) {
        v = s[m * 6 + srcj][0];
        r = rq0[0];
        h = hq0[0];
        zero = k > 0;
      }
    }
    int qv = (int)__umulhi((uint32_t)abs(v) + h, r);
    qv = v < 0 ? -qv : qv;
    dst[blk * 64 + k] = (int16_t)(zero ? 0 : qv);
  }
}

__device__ __forceinline__ int nbits_of(int v) { return 32 - __clz(abs(v)); }

// DC predictor of block `blk` (index within the image, 6 per MCU): the previous block of the same component in scan order
__device__ __forceinline__ int dc_pred(const int16_t* __restrict__ coef_img, int blk) {
  const int m = blk / 6, j = blk - m * 6;
  int prev;
  if (j >= 1 && j <= 3) prev = blk - 1;
  else if (m == 0) return 0;
  else prev = j == 0 ? blk - 3 : blk - 6;
  return coef_img[(size_t)prev * 64];
}

constexpr int JPEG_CSTRIDE = 33;   // words per staged block (32 + 1: thread t reads word t * 33 + i without bank conflicts)

// the coefficients of JPEG_PACK_N consecutive blocks -> shared memory with coalesced 16-byte loads
__device__ __forceinline__ void stage_blocks(const int16_t* __restrict__ coef_img, int blk0, int nblk, uint32_t* sc) {
  const uint4* src = reinterpret_cast<const uint4*>(coef_img + (size_t)blk0 * 64);
  const int n4 = min(JPEG_PACK_N, nblk - blk0) * 8;
  for (int i = threadIdx.x; i < n4; i += JPEG_PACK_N) {
    const uint4 v = __ldg(src + i);
    uint32_t* d = sc + (i >> 3) * JPEG_CSTRIDE + (i & 7) * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

__global__ void __launch_bounds__(JPEG_PACK_N)
jpeg_len_kernel(const int16_t* __restrict__ coef, uint32_t* __restrict__ len, int nblk, JpegTables tab) {
  __shared__ uint32_t sc[JPEG_PACK_N * JPEG_CSTRIDE];
  __shared__ uint8_t lac[2][256], ldc[2][16];
  const int tid = threadIdx.x, blk0 = blockIdx.x * JPEG_PACK_N, blk = blk0 + tid, b = blockIdx.y;
  for (int i = tid; i < 512; i += JPEG_PACK_N) lac[i >> 8][i & 255] = (uint8_t)(tab.ac[i >> 8][i & 255] & 31);
  if (tid < 32) ldc[tid >> 4][tid & 15] = (uint8_t)(tab.dc[tid >> 4][tid & 15] & 31);
  const int16_t* coef_img = coef + (size_t)b * nblk * 64;
  stage_blocks(coef_img, blk0, nblk, sc);
  __syncthreads();
  if (blk >= nblk) return;
  const uint32_t* mine = sc + tid * JPEG_CSTRIDE;
  const int t = (blk % 6) < 4 ? 0 : 1;
  uint32_t w = mine[0];
  const int diff = (int)(int16_t)(w & 0xFFFFu) - dc_pred(coef_img, blk);
  int n = nbits_of(diff);
  uint32_t bits = ldc[t][n] + n;
  int run = 0;
  auto ac = [&](int v) {
    if (v == 0) { ++run; return; }
    bits += (run >> 4) * lac[t][0xF0];
    const int nb = nbits_of(v);
    bits += lac[t][((run & 15) << 4) | nb] + nb;
    run = 0;
  };
  ac((int)(int16_t)(w >> 16));
#pragma unroll 4
  for (int i = 1; i < 32; ++i) {
    w = mine[i];
    ac((int)(int16_t)(w & 0xFFFFu));
    ac((int)(int16_t)(w >> 16));
  }
  if (run) bits += lac[t][0];
  len[(size_t)b * nblk + blk] = bits;
}

// Per-image exclusive scan, one CTA per image.  ZERO: also clear the raw-stream words that two pack CTAs share (the word holding
// the first bit of every JPEG_PACK_N-th block, and the word holding the end of the stream).  COUNT_FROM_TOTAL: the number of items
// is the number of stuffing chunks of this image's stream.
template <bool ZERO, bool COUNT_FROM_TOTAL>
__global__ void __launch_bounds__(1024)
jpeg_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ total, const uint32_t* __restrict__ bits_total,
                 size_t stride, int n_fixed, uint32_t* __restrict__ raw, size_t raw_words_per_image) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int n = n_fixed;
  if (COUNT_FROM_TOTAL) {
    const uint32_t nbytes = (bits_total[b] + 7) >> 3;
    n = (int)((nbytes + JPEG_CHUNK - 1) / JPEG_CHUNK);
  }
  const uint32_t* src = in + (size_t)b * stride;
  uint32_t* dst = out + (size_t)b * stride;
  uint32_t* raw_img = ZERO ? raw + (size_t)b * raw_words_per_image : nullptr;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    const int i0 = base + tid * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k < n ? src[i0 + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t excl = carry + (wid ? warp_sum[wid - 1] : 0u) + x - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) {
        dst[i0 + k] = excl;
        if (ZERO && ((i0 + k) % JPEG_PACK_N) == 0) raw_img[excl >> 5] = 0u;
      }
      excl += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) {
    total[b] = carry_s;
    if (ZERO) raw_img[carry_s >> 5] = 0u;
  }
}

constexpr int JPEG_SPAN_WORDS = 2048;   // shared-memory span of one pack CTA (64 bytes per block; q95 photographs need ~30)

__global__ void __launch_bounds__(JPEG_PACK_N)
jpeg_pack_kernel(const int16_t* __restrict__ coef, const uint32_t* __restrict__ off, const uint32_t* __restrict__ total,
                 uint32_t* __restrict__ raw, size_t raw_words_per_image, int nblk, JpegTables tab) {
  __shared__ uint32_t sc[JPEG_PACK_N * JPEG_CSTRIDE];
  __shared__ uint32_t span[JPEG_SPAN_WORDS];
  __shared__ uint32_t tac[2][256], tdc[2][16];
  const int tid = threadIdx.x, b = blockIdx.y, blk0 = blockIdx.x * JPEG_PACK_N, blk = blk0 + tid;
  for (int i = tid; i < 512; i += JPEG_PACK_N) tac[i >> 8][i & 255] = tab.ac[i >> 8][i & 255];
  if (tid < 32) tdc[tid >> 4][tid & 15] = tab.dc[tid >> 4][tid & 15];
  const int16_t* coef_img = coef + (size_t)b * nblk * 64;
  stage_blocks(coef_img, blk0, nblk, sc);
  const uint32_t* off_img = off + (size_t)b * nblk;
  const uint32_t begin = off_img[blk0];
  const int blk_end = min(blk0 + JPEG_PACK_N, nblk);
  const uint32_t end = blk_end < nblk ? off_img[blk_end] : total[b];
  const uint32_t w0 = begin >> 5;
  const uint32_t nwords = end > begin ? ((end - 1) >> 5) - w0 + 1 : 0;
  // The CTA's bits form one contiguous span of the raw stream.  A word the span shares with a neighbouring CTA (the span starts or
  // ends inside it) was zeroed by the scan kernel and is merged with atomicOr; every other word belongs to this CTA alone.
  // Usual case: the span is assembled in shared memory and stored with coalesced word stores.  A span too long for it (noise at
  // quality 100) is assembled in place: the CTA zeroes its own words, then every thread ORs straight into HBM.
  const bool in_smem = nwords <= JPEG_SPAN_WORDS;
  uint32_t* dst = raw + (size_t)b * raw_words_per_image + w0;
  auto shared_word = [&](uint32_t i) { return (i == 0 && (begin & 31)) || (i == nwords - 1 && (end & 31)); };
  if (in_smem) {
    for (uint32_t i = tid; i < nwords; i += JPEG_PACK_N) span[i] = 0u;
  } else {
    for (uint32_t i = tid; i < nwords; i += JPEG_PACK_N)
      if (!shared_word(i)) dst[i] = 0u;
  }
  __syncthreads();
  if (blk < nblk) {
    const uint32_t* mine = sc + tid * JPEG_CSTRIDE;
    const int t = (blk % 6) < 4 ? 0 : 1;
    const uint32_t start = off_img[blk];
    uint32_t* target = (in_smem ? span : dst) + ((start >> 5) - w0);
    int nacc = (int)(start & 31);
    unsigned long long acc = 0;
    auto flush = [&](uint32_t word) {
      atomicOr(target++, in_smem ? word : __byte_perm(word, 0, 0x0123));
    };
    auto emit = [&](uint32_t code, int length) {
      acc = (acc << length) | code;
      nacc += length;
      if (nacc >= 32) {
        nacc -= 32;
        flush((uint32_t)(acc >> nacc));
        acc &= (1ull << nacc) - 1;
      }
    };
    uint32_t w = mine[0];
    const int diff = (int)(int16_t)(w & 0xFFFFu) - dc_pred(coef_img, blk);
    int n = nbits_of(diff);
    uint32_t e = tdc[t][n];
    emit(e >> 5, (int)(e & 31));
    if (n) emit((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << n) - 1), n);
    int run = 0;
    auto ac = [&](int v) {
      if (v == 0) { ++run; return; }
      while (run > 15) {
        const uint32_t z = tac[t][0xF0];
        emit(z >> 5, (int)(z & 31));
        run -= 16;
      }
      const int nb = nbits_of(v);
      const uint32_t c = tac[t][(run << 4) | nb];
      emit(c >> 5, (int)(c & 31));
      emit((uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1), nb);
      run = 0;
    };
    ac((int)(int16_t)(w >> 16));
#pragma unroll 4
    for (int i = 1; i < 32; ++i) {
      w = mine[i];
      ac((int)(int16_t)(w & 0xFFFFu));
      ac((int)(int16_t)(w >> 16));
    }
    if (run) {
      e = tac[t][0];
      emit(e >> 5, (int)(e & 31));
    }
    if (nacc) flush((uint32_t)(acc << (32 - nacc)));
  }
  if (!in_smem) return;
  __syncthreads();
  for (uint32_t i = tid; i < nwords; i += JPEG_PACK_N) {
    const uint32_t v = __byte_perm(span[i], 0, 0x0123);           // stream byte order
    if (shared_word(i)) atomicOr(dst + i, v);
    else dst[i] = v;
  }
}

// raw stream -> bytes of one 32-byte chunk with the final byte padded with 1 bits; returns the number of valid bytes
__device__ __forceinline__ int load_chunk(const uint32_t* __restrict__ raw_img, uint32_t bits, int chunk, uint8_t* bytes) {
  const uint32_t nbytes = (bits + 7) >> 3;
  const uint32_t base = (uint32_t)chunk * JPEG_CHUNK;
  const int valid = (int)min((uint32_t)JPEG_CHUNK, nbytes - base);
  const uint4* p = reinterpret_cast<const uint4*>(raw_img + base / 4);
  const uint4 a = p[0], c = p[1];
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bytes[4 * i] = (uint8_t)(w[i] & 0xFF);
    bytes[4 * i + 1] = (uint8_t)((w[i] >> 8) & 0xFF);
    bytes[4 * i + 2] = (uint8_t)((w[i] >> 16) & 0xFF);
    bytes[4 * i + 3] = (uint8_t)(w[i] >> 24);
  }
  if (base + valid == nbytes && (bits & 7)) bytes[valid - 1] |= (uint8_t)((1u << (8 - (bits & 7))) - 1);
  return valid;
}

__global__ void __launch_bounds__(256)
jpeg_ffcount_kernel(const uint32_t* __restrict__ raw, size_t raw_words_per_image, const uint32_t* __restrict__ bits_total,
                    uint32_t* __restrict__ cnt, size_t max_chunks) {
  const int chunk = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  const uint32_t bits = bits_total[b];
  const uint32_t nbytes = (bits + 7) >> 3;
  if ((size_t)chunk * JPEG_CHUNK >= nbytes) return;
  uint8_t bytes[JPEG_CHUNK];
  const int valid = load_chunk(raw + (size_t)b * raw_words_per_image, bits, chunk, bytes);
  uint32_t n = 0;
#pragma unroll
  for (int i = 0; i < JPEG_CHUNK; ++i) n += (i < valid && bytes[i] == 0xFF) ? 1u : 0u;
  cnt[(size_t)b * max_chunks + chunk] = n;
}

__global__ void __launch_bounds__(256)
jpeg_stuff_kernel(const uint32_t* __restrict__ raw, size_t raw_words_per_image, const uint32_t* __restrict__ bits_total,
                  const uint32_t* __restrict__ ffoff, const uint32_t* __restrict__ fftotal, size_t max_chunks, uint8_t* __restrict__ out,
                  size_t capacity, int32_t* __restrict__ sizes) {
  // 256 chunks of 32 raw bytes -> at most 16 KB of stuffed output, contiguous in the file: assembled in shared memory at the file's
  // own 4-byte alignment, then stored as whole words (single bytes only at the two ends, which other CTAs share)
  __shared__ __align__(16) uint8_t sb[256 * 2 * JPEG_CHUNK + 16];
  __shared__ uint32_t s_end;
  const int tid = threadIdx.x, chunk0 = blockIdx.x * 256, chunk = chunk0 + tid, b = blockIdx.y;
  const uint32_t bits = bits_total[b];
  const uint32_t nbytes = (bits + 7) >> 3;
  if ((size_t)chunk0 * JPEG_CHUNK >= nbytes) return;
  uint8_t* dst = out + (size_t)b * capacity;
  const size_t pos0 = (size_t)JPEG_HEADER + (size_t)chunk0 * JPEG_CHUNK + ffoff[(size_t)b * max_chunks + chunk0];
  const uint32_t lead = (uint32_t)((uintptr_t)(dst + pos0) & 3);    // the span starts `lead` bytes into an aligned word
  if (tid == 0) s_end = 0;
  __syncthreads();
  if ((size_t)chunk * JPEG_CHUNK < nbytes) {
    uint8_t bytes[JPEG_CHUNK];
    const int valid = load_chunk(raw + (size_t)b * raw_words_per_image, bits, chunk, bytes);
    uint32_t rel = lead + (uint32_t)((size_t)(chunk - chunk0) * JPEG_CHUNK + ffoff[(size_t)b * max_chunks + chunk] - ffoff[(size_t)b * max_chunks + chunk0]);
#pragma unroll
    for (int i = 0; i < JPEG_CHUNK; ++i) {
      if (i < valid) {
        sb[rel++] = bytes[i];
        if (bytes[i] == 0xFF) sb[rel++] = 0;
      }
    }
    const bool last = (size_t)chunk * JPEG_CHUNK + valid == nbytes;
    if (last) {                                                     // EOI and the file size
      sb[rel++] = 0xFF;
      sb[rel++] = 0xD9;
      sizes[b] = (int32_t)((size_t)JPEG_HEADER + nbytes + fftotal[b] + 2);
    }
    if (last || tid == 255) s_end = rel;
  }
  __syncthreads();
  uint32_t end = s_end;
  if (pos0 >= capacity) return;
  if ((size_t)(end - lead) > capacity - pos0) end = lead + (uint32_t)(capacity - pos0);    // nothing is written beyond the capacity
  uint8_t* base = dst + pos0 - lead;                                // 4-byte aligned
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(sb);
  for (uint32_t i = tid; i * 4 < end; i += 256) {
    const uint32_t lo = i * 4, hi = lo + 4;
    if (lo >= lead && hi <= end) reinterpret_cast<uint32_t*>(base)[i] = sw[i];
    else
      for (uint32_t k = max(lo, lead); k < min(hi, end); ++k) base[k] = sb[k];
  }
}

__global__ void __launch_bounds__(256)
jpeg_header_kernel(uint8_t* __restrict__ out, size_t capacity, JpegHeader hdr) {
  uint8_t* dst = out + (size_t)blockIdx.x * capacity;
  for (int i = threadIdx.x; i < JPEG_HEADER; i += 256)
    if ((size_t)i < capacity) dst[i] = hdr.b[i];
}

// ---- host: tables and header ---------------------------------------------------------------------------------------------------------
static const uint8_t k_luma_q[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                                     69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55,  64,
                                     81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72,  92,  95,  98,  112, 100, 103, 99};
static const uint8_t k_chroma_q[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                       99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                       99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t k_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const uint8_t k_dc_bits[2][16] = {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}};
static const uint8_t k_ac_bits[2][16] = {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D}, {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}};
static const uint8_t k_ac_vals[2][162] = {
    {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xA1,
     0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26,
     0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56,
     0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83, 0x84, 0x85,
     0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA,
     0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6,
     0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9,
     0xFA},
    {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
     0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1, 0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19,
     0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55,
     0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x82, 0x83,
     0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8,
     0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4,
     0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9,
     0xFA}};

static void fill_codes(const uint8_t* bits, const uint8_t* vals, uint32_t* table) {       // Annex C
  uint32_t code = 0;
  int k = 0;
  for (int length = 1; length <= 16; ++length) {
    for (int i = 0; i < bits[length - 1]; ++i) table[vals[k++]] = (code++ << 5) | (uint32_t)length;
    code <<= 1;
  }
}

static void make_tables(int quality, JpegTables* tab, JpegQuant* qt) {
  static const uint8_t dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
  memset(tab, 0, sizeof(*tab));
  for (int t = 0; t < 2; ++t) {
    fill_codes(k_dc_bits[t], dc_vals, tab->dc[t]);
    fill_codes(k_ac_bits[t], k_ac_vals[t], tab->ac[t]);
  }
  const int q = quality < 1 ? 1 : quality > 100 ? 100 : quality;
  const int scale = q < 50 ? 5000 / q : 200 - 2 * q;                                      // jcparam.c jpeg_quality_scaling
  for (int i = 0; i < 64; ++i) {
    const int l = (k_luma_q[i] * scale + 50) / 100, c = (k_chroma_q[i] * scale + 50) / 100;
    qt->q[0][i] = (uint16_t)(l < 1 ? 1 : l > 255 ? 255 : l);
    qt->q[1][i] = (uint16_t)(c < 1 ? 1 : c > 255 ? 255 : c);
  }
}

static void make_header(int h, int w, const JpegQuant& qt, JpegHeader* hdr) {
  uint8_t* p = hdr->b;
  auto put = [&](std::initializer_list<int> v) { for (int x : v) *p++ = (uint8_t)x; };
  put({0xFF, 0xD8, 0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
  for (int t = 0; t < 2; ++t) {
    put({0xFF, 0xDB, 0, 67, t});
    for (int i = 0; i < 64; ++i) *p++ = (uint8_t)qt.q[t][k_zigzag[i]];
  }
  put({0xFF, 0xC0, 0, 17, 8, h >> 8, h & 255, w >> 8, w & 255, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
  static const uint8_t dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
  for (int t = 0; t < 2; ++t) {
    put({0xFF, 0xC4, 0, 3 + 16 + 12, t});
    for (int i = 0; i < 16; ++i) *p++ = k_dc_bits[t][i];
    for (int i = 0; i < 12; ++i) *p++ = dc_vals[i];
    put({0xFF, 0xC4, 0, 3 + 16 + 162, 0x10 | t});
    for (int i = 0; i < 16; ++i) *p++ = k_ac_bits[t][i];
    for (int i = 0; i < 162; ++i) *p++ = k_ac_vals[t][i];
  }
  put({0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3F, 0});
}

struct JpegDims {
  int mh, mw;
  size_t nblk, raw_bytes, max_chunks;
};
static JpegDims jpeg_dims(int h, int w) {
  JpegDims d;
  d.mh = (h + 15) / 16;
  d.mw = (w + 15) / 16;
  d.nblk = (size_t)d.mh * d.mw * 6;
  d.raw_bytes = (d.nblk * JPEG_BLOCK_MAX_BYTES + 64 + 255) & ~(size_t)255;
  d.max_chunks = d.raw_bytes / JPEG_CHUNK;
  return d;
}
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace cic

using namespace cic;

extern "C" size_t cic_jpeg_max_bytes(int h, int w) {
  if (h <= 0 || w <= 0) return 0;
  const JpegDims d = jpeg_dims(h, w);
  return (size_t)JPEG_HEADER + 2 * d.nblk * JPEG_BLOCK_MAX_BYTES + 2;
}

extern "C" size_t cic_jpeg_workspace_bytes(int batch, int h, int w) {
  if (batch <= 0 || h <= 0 || w <= 0) return 0;
  const JpegDims d = jpeg_dims(h, w);
  const size_t b = (size_t)batch;
  return al256(b * d.nblk * 64 * 2) + 2 * al256(b * d.nblk * 4) + b * d.raw_bytes + 2 * al256(b * d.max_chunks * 4) + 3 * al256(b * 4) + 256;
}

extern "C" int cic_jpeg_encode_u8(const uint8_t* d_img, int batch, int h, int w, int rgb, int quality, uint8_t* d_out, size_t capacity,
                                  int32_t* d_sizes, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0 && h <= 65535 && w <= 65535, "cic_jpeg_encode_u8: bad shape (JPEG holds at most 65535 x 65535)");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_img && d_out && d_sizes, "cic_jpeg_encode_u8: null pointer");
  CIC_REQUIRE(batch <= 65535, "cic_jpeg_encode_u8: at most 65535 images per call");
  CIC_REQUIRE(capacity >= (size_t)JPEG_HEADER + 2, "cic_jpeg_encode_u8: capacity below the header size");
  const JpegDims d = jpeg_dims(h, w);
  CIC_REQUIRE(d.nblk * 1658 < 0xFFFFFFFFull, "cic_jpeg_encode_u8: image too large for 32-bit bit offsets");
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_jpeg_workspace_bytes(batch, h, w), "cic_jpeg_encode_u8: workspace too small");
  CIC_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "cic_jpeg_encode_u8: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = (size_t)batch;
  char* ws = (char*)d_workspace;
  int16_t* coef = (int16_t*)ws; ws += al256(B * d.nblk * 64 * 2);
  uint32_t* len = (uint32_t*)ws; ws += al256(B * d.nblk * 4);
  uint32_t* off = (uint32_t*)ws; ws += al256(B * d.nblk * 4);
  uint32_t* raw = (uint32_t*)ws; ws += B * d.raw_bytes;
  uint32_t* ffcnt = (uint32_t*)ws; ws += al256(B * d.max_chunks * 4);
  uint32_t* ffoff = (uint32_t*)ws; ws += al256(B * d.max_chunks * 4);
  uint32_t* bits_total = (uint32_t*)ws; ws += al256(B * 4);
  uint32_t* fftotal = (uint32_t*)ws; ws += al256(B * 4);
  JpegTables tab;
  JpegQuant qt;
  JpegHeader hdr;
  make_tables(quality, &tab, &qt);
  make_header(h, w, qt, &hdr);
  const size_t raw_words = d.raw_bytes / 4;
  const int nblk = (int)d.nblk;
  jpeg_header_kernel<<<batch, 256, 0, st>>>(d_out, capacity, hdr);
  const dim3 g1((d.mw + 7) / 8, d.mh, batch);
  const int words_ok = (w % 4 == 0) && (((uintptr_t)d_img & 3) == 0);
  if (rgb) jpeg_dct_kernel<true><<<g1, 256, 0, st>>>(d_img, coef, h, w, d.mw, d.mh, words_ok, qt);
  else jpeg_dct_kernel<false><<<g1, 256, 0, st>>>(d_img, coef, h, w, d.mw, d.mh, words_ok, qt);
  jpeg_len_kernel<<<dim3((nblk + JPEG_PACK_N - 1) / JPEG_PACK_N, batch), JPEG_PACK_N, 0, st>>>(coef, len, nblk, tab);
  jpeg_scan_kernel<true, false><<<batch, 1024, 0, st>>>(len, off, bits_total, nullptr, d.nblk, nblk, raw, raw_words);
  jpeg_pack_kernel<<<dim3((nblk + JPEG_PACK_N - 1) / JPEG_PACK_N, batch), JPEG_PACK_N, 0, st>>>(coef, off, bits_total, raw, raw_words, nblk, tab);
  const dim3 g5((unsigned)((d.max_chunks + 255) / 256), batch);
  jpeg_ffcount_kernel<<<g5, 256, 0, st>>>(raw, raw_words, bits_total, ffcnt, d.max_chunks);
  jpeg_scan_kernel<false, true><<<batch, 1024, 0, st>>>(ffcnt, ffoff, fftotal, bits_total, d.max_chunks, 0, nullptr, 0);
  jpeg_stuff_kernel<<<g5, 256, 0, st>>>(raw, raw_words, bits_total, ffoff, fftotal, d.max_chunks, d_out, capacity, d_sizes);
  for (int i = 0; i < 8; ++i) CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("jpeg kernels");
  return CIC_OK;
}
