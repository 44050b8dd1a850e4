// Host-side description of a tcgen05 layer launch (activation views, packed weights, epilogue) and the
// small conversion kernels around the tensor-core path.
#pragma once
#include "tc_gemm.cuh"

namespace cic {

typedef __nv_bfloat16 bf16;

struct TcAct {      // NHWC bf16 activation view: channels [coff, coff + C) of pixel records of `ld` elements
  const bf16* hi;
  const bf16* lo;   // nullptr when the tensor has no low part
  int C, ld, coff;
};

struct TcMat {      // B operand: [batches][rows][K] bf16, K contiguous
  const bf16* hi;
  const bf16* lo;
  int K;                   // elements per row used by the GEMM
  int rows;                // rows per batch (phases * N_pad for weights)
  long long row_stride;    // elements
  int batches;
  long long batch_stride;  // elements
};

struct TcEpilogue {
  const float* bias = nullptr;
  const float* scale = nullptr;
  const float* shift = nullptr;
  float alpha = 1.f;
  int act = CIC_ACT_NONE;
  int out_mode = TC_OUT_BF16;
  void* out_hi = nullptr;
  void* out_lo = nullptr;
  const bf16* res_hi = nullptr;
  const bf16* res_lo = nullptr;
  int out_ld = 0, out_coff = 0;
  int up2 = 0;
  int tm_tx = 0, tm_ty = 0, tm_IH = 0, tm_IW = 0;  // output tile map (see TcEpi)
};

enum TcKind { TC_CONV_S1 = 0, TC_CONV_S2 = 1, TC_DECONV_K4S2 = 2 };

struct TcLayer {
  int kind = TC_CONV_S1;
  TcAct src[2];
  int nsrc = 1;
  int batch = 0, H = 0, W = 0;  // input size per batch item
  int kh = 1, kw = 1, pad_t = 0, pad_l = 0;
  TcMat w{};
  int N = 0;          // output channels (rows of w per phase)
  bool split = false;  // 3-term split-bf16 (needs src[].lo and w.lo)
  int splits = 1;      // split-K; > 1 requires epi.out_mode == TC_OUT_PARTIAL
  bool b_batched = false;
  TcEpilogue epi;
};

int tc_encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
int tc_run_layer(const TcLayer& L, cudaStream_t st);
int tc_pick_block_n(int n_pad, bool split, int bk);
int tc_pick_block_k(const TcLayer& L);
bool tc_pair_ok(int bk, long long m_tiles, int n_pad, int n);

// fp32 [K][ld] (row-major, Keras kernel / Dense layout; the first N columns) -> bf16 hi/lo [N_pad][K]
// (K contiguous), zero rows >= N
// col_scale (optional, N floats): column n of src is multiplied by col_scale[n] before rounding - folded BatchNorm scale
int tc_pack_weight(const float* src, int K, int N, int N_pad, int ld, bf16* hi, bf16* lo, cudaStream_t st, const float* col_scale = nullptr);
// out[n] = bias[n] * scale[n] + shift[n]: the one epilogue vector left once the scale lives in the weights
int tc_fuse_bias(const float* bias, const float* scale, const float* shift, float* out, int n, cudaStream_t st);
// fp32 -> bf16 hi (+ lo)
int tc_split_f32(const float* src, bf16* hi, bf16* lo, size_t n, cudaStream_t st);
// bf16 hi (+ lo) -> fp32, with strided source records (ld, coff) -> dense C
int tc_join_to_f32(const bf16* hi, const bf16* lo, float* dst, size_t pixels, int C, int ld, int coff, cudaStream_t st);
// tiny-Cout output convs by column strips (conv_rows_tc.cu)
struct TileMap;
size_t conv_rows_image_bytes(int ks, int cin);
int conv_rows_pack(const float* w, uint8_t* img, int ks, int cin, int cout, cudaStream_t st);
int launch_conv_rows_tc(const TcAct* srcs, int nsrc, const uint8_t* wimg, const float* bias, int ks, int cout, int act, float* out,
                        int batch, int H, int W, const TileMap& tm, cudaStream_t st);
int launch_gen_tail(const TcAct& hq, const TcAct& lq, const uint8_t* wimg, const float* bias_hq, const float* bias_lq, const float* mask,
                    const float* bpp, float* out, float* dt_out, double* dt_sum, float* out_hq, float* out_lq, int n_img, int batch, int H,
                    int W, const TileMap& tm, cudaStream_t st);
// MaxPooling2D((2,2)) on NHWC bf16 (even H, W; C % 8 == 0)
int tc_maxpool2x2_bf16(const bf16* x, bf16* y, int batch, int H, int W, int C, cudaStream_t st);
// split-K partials [splits][M][N] -> epilogue -> fp32 [M][N] and/or bf16 hi/lo [M][N]
int tc_splitk_reduce(const float* partial, int splits, long long M, int N, const float* bias, const float* scale,
                     const float* shift, int act, float* out_f32, bf16* out_hi, bf16* out_lo, cudaStream_t st);
// fused attention core (attn_fused.cu): y = gamma * softmax(q k^T) v + x on (hi, lo) bf16 tensors, 32 q/k channels, 256 value channels
int launch_attn_fused(const bf16* qk_hi, const bf16* qk_lo, const bf16* vt_hi, const bf16* vt_lo, const bf16* x_hi, const bf16* x_lo,
                      bf16* y_hi, bf16* y_lo, const float* bias_v_scaled, float gamma, int nb, int tokens, cudaStream_t st);
// row softmax of fp32 logits -> bf16 hi/lo probabilities (tf.nn.softmax(axis=-1), GAN_functions.py:359)
int tc_softmax_rows_split(const float* logits, bf16* p_hi, bf16* p_lo, long long rows, int cols, cudaStream_t st);

}  // namespace cic
