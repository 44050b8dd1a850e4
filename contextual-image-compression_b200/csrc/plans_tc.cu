// CIC_PREC_TC walkers: the reference graphs (see plans.cu for the file:line map) on the tcgen05 path.
//
// Activations are NHWC bf16.  The encoder chain (conv2..conv4, attention, Dense) carries every tensor as a
// (hi, lo) bf16 pair and multiplies with the 3-term split (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM)
// so that the quantised symbols match the fp32 reference; the decoders (Dense, four
// transposed convs, output conv) and the autoencoder run single-pass bf16.  Layers a tensor-core tile
// cannot take (3-channel inputs, the tiny saliency / RD networks) stay on the fp32 CUDA-core kernels.
#include "plan.cuh"
#include "tc_host.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace cic {

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct ActBuf {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

static ActBuf alloc_act(Ctx& c, size_t elems, bool with_lo) {
  ActBuf a;
  a.hi = (bf16*)c.arena.alloc_bytes(elems * sizeof(bf16));
  if (with_lo) a.lo = (bf16*)c.arena.alloc_bytes(elems * sizeof(bf16));
  return a;
}

static TcAct view(const ActBuf& a, int C, int ld = 0, int coff = 0) { return TcAct{a.hi, a.lo, C, ld ? ld : C, coff}; }

// ---- packed weights ---------------------------------------------------------------------------
// fp32 [K][ld] device matrix (columns [col0, col0 + N)) -> bf16 [N_pad][K] hi (+ lo) owned by the plan
// bn != "": the layer is followed by inference BatchNorm: y = (x W + bias) * scale + shift = x (W * scale) + (bias * scale + shift), so the
// scale is folded into the packed weights and "<name>#fb" holds the fused bias - the epilogue then reads one per-channel vector
// instead of three (a uniform 128-bit load costs four l1tex data-pipe wavefronts, and that pipe bounds the narrow layers:
// profiles/r01_epilogue_data_pipe.md)
static int fuse_bn(cic_plan* pl, const std::string& name, const std::string& bn, int N) {
  const WeightStore& w = pl->w;
  const float* scale = w.ptr(bn + "/scale");
  const float* shift = w.ptr(bn + "/shift");
  CIC_REQUIRE(scale && shift, "tc plan: missing folded BatchNorm '%s'", bn.c_str());
  float* fb = (float*)pl->tcw.alloc(name + "#fb", (size_t)N * sizeof(float));
  CIC_REQUIRE(fb, "tc plan: out of device memory");
  return tc_fuse_bias(w.ptr(name + "/bias"), scale, shift, fb, N, nullptr);
}

static int pack(cic_plan* pl, const std::string& name, const float* src, int K, int N, int N_pad, bool with_lo, int ld = 0, int col0 = 0,
                const std::string& bn = std::string()) {
  CIC_REQUIRE(src, "tc plan: missing fp32 weight for '%s'", name.c_str());
  const size_t bytes = (size_t)N_pad * K * sizeof(bf16);
  bf16* hi = (bf16*)pl->tcw.alloc(name + "#hi", bytes);
  bf16* lo = with_lo ? (bf16*)pl->tcw.alloc(name + "#lo", bytes) : nullptr;
  if (!hi || (with_lo && !lo)) {
    set_error("tc plan: out of device memory packing '%s' (%zu bytes)", name.c_str(), bytes);
    return CIC_ERR_CUDA;
  }
  if (!bn.empty()) {
    const int rc = fuse_bn(pl, name, bn, N);
    if (rc) return rc;
  }
  return tc_pack_weight(src + col0, K, N, N_pad, ld ? ld : N, hi, lo, nullptr, bn.empty() ? nullptr : pl->w.ptr(bn + "/scale"));
}

static const float* fused_bias(const cic_plan* pl, const std::string& name) { return (const float*)pl->tcw.ptr(name + "#fb"); }

static TcMat mat(const cic_plan* pl, const std::string& name, int K, int rows) {
  TcMat m{};
  m.hi = (const bf16*)pl->tcw.ptr(name + "#hi");
  m.lo = (const bf16*)pl->tcw.ptr(name + "#lo");
  m.K = K;
  m.rows = rows;
  m.row_stride = K;
  m.batches = 1;
  m.batch_stride = (long long)rows * K;
  return m;
}

int build_plan_tc(cic_plan* pl, const cic_tensor* tensors, int n, const std::string& prefix) {
  (void)tensors; (void)n; (void)prefix;
  const WeightStore& w = pl->w;
  int rc = CIC_OK;
  switch (pl->kind) {
    case CIC_PLAN_AUTOENCODER: {
      static const struct { const char* name; int cin, cout; } L[] = {
          {"conv2", 32, 64}, {"conv3", 64, 64}, {"conv_x2", 64, 64}, {"conv5", 128, 32}, {"conv_x1", 32, 32}};
      for (auto& l : L)
        if ((rc = pack(pl, l.name, w.ptr(std::string(l.name) + "/kernel"), 9 * l.cin, l.cout, l.cout, false))) return rc;
      if (pl->opts.img_c <= 16) rc = pack(pl, "conv_out", w.ptr("conv_out/kernel"), 9 * 64, pl->opts.img_c, 16, false);
      if (!rc && pl->opts.img_c == 3) {  // column-strip kernel image (conv_rows_tc.cu)
        uint8_t* img = (uint8_t*)pl->tcw.alloc("conv_out#rows", conv_rows_image_bytes(3, 64));
        CIC_REQUIRE(img, "tc plan: out of device memory");
        rc = conv_rows_pack(w.ptr("conv_out/kernel"), img, 3, 64, 3, nullptr);
      }
      if (!rc && pl->opts.img_c == 3) {  // conv1 (3 -> 32, k3) as the resident weight image of first_conv_tc.cu
        uint8_t* img = (uint8_t*)pl->tcw.alloc("conv1#fc", first_conv_image_bytes());
        CIC_REQUIRE(img && w.ptr("conv1/kernel"), "tc plan: conv1 weights");
        rc = first_conv_pack(w.ptr("conv1/kernel"), 27, img, nullptr);
      }
      break;
    }
    case CIC_PLAN_ENCODER: {
      const int ch[5] = {pl->opts.img_c, 64, 128, 256, 512};
      if (pl->opts.img_c == 3) {  // conv1 weights as the pre-swizzled shared-memory image of conv1_tc.cu
        uint8_t* img = (uint8_t*)pl->tcw.alloc("conv1#img", conv1_tc_image_bytes(1));
        CIC_REQUIRE(img && w.ptr("conv1/kernel"), "tc plan: conv1 weights");
        if ((rc = conv1_tc_pack(w.ptr("conv1/kernel"), nullptr, img, nullptr))) return rc;
      }
      for (int i = 2; i <= 4; ++i) {
        const std::string nm = "conv" + std::to_string(i);
        if ((rc = pack(pl, nm, w.ptr(nm + "/kernel"), 16 * ch[i - 1], ch[i], ch[i], true, 0, 0, "bn" + std::to_string(i)))) return rc;
      }
      if (pl->opts.add_attention) {
        const DevTensor* qkv = w.find("attn/qkv/kernel");
        CIC_REQUIRE(qkv, "tc plan: missing attention weights");
        const int C = (int)qkv->shape[0], Nt = (int)qkv->shape[1], dq = (Nt - C) / 2;
        CIC_REQUIRE(dq % 32 == 0 && C % 64 == 0, "tc plan: attention needs C %% 64 == 0 and C/8 %% 32 == 0 (C=%d)", C);
        if ((rc = pack(pl, "attn/qk", qkv->p, C, 2 * dq, 2 * dq, true, Nt, 0))) return rc;
        if ((rc = pack(pl, "attn/v", qkv->p, C, C, C, true, Nt, 2 * dq))) return rc;
        float g = 0.f;
        CIC_CHECK_CUDA(cudaMemcpy(&g, w.ptr("attn/gamma"), sizeof(float), cudaMemcpyDeviceToHost));
        pl->attn_gamma = g;
      }
      const int feat = (pl->opts.img_h / 16) * (pl->opts.img_w / 16) * 512;
      rc = pack(pl, "dense", w.ptr("dense/kernel"), feat, pl->opts.latent_dim, round_up(pl->opts.latent_dim, 16), true);
      break;
    }
    case CIC_PLAN_GENERATOR: {
      const int feat = (pl->opts.img_h / 16) * (pl->opts.img_w / 16) * 512, L = pl->opts.latent_dim;
      if (L % 32 == 0 && (rc = pack(pl, "dense", w.ptr("dense/kernel"), L, feat, feat, false, 0, 0, "bn0"))) return rc;
      const int cin[5] = {0, 512, 512, 256, 128}, cout[5] = {0, 256, 128, 64, 32};
      for (int i = 1; i <= 4; ++i) {
        const std::string nm = "deconv" + std::to_string(i);
        const float* ph = w.ptr(nm + "/phases");  // [4][4*cin][cout]
        CIC_REQUIRE(ph, "tc plan: missing %s phases", nm.c_str());
        const size_t bytes = (size_t)4 * cout[i] * 4 * cin[i] * sizeof(bf16);
        bf16* hi = (bf16*)pl->tcw.alloc(nm + "#hi", bytes);
        CIC_REQUIRE(hi, "tc plan: out of device memory packing %s", nm.c_str());
        const std::string bn = "bn" + std::to_string(i);
        if ((rc = fuse_bn(pl, nm, bn, cout[i]))) return rc;
        for (int p = 0; p < 4; ++p)
          if ((rc = tc_pack_weight(ph + (size_t)p * 4 * cin[i] * cout[i], 4 * cin[i], cout[i], cout[i], cout[i],
                                   hi + (size_t)p * cout[i] * 4 * cin[i], nullptr, nullptr, w.ptr(bn + "/scale")))) return rc;
      }
      if (pl->opts.img_c <= 16) rc = pack(pl, "conv_out", w.ptr("conv_out/kernel"), 16 * 32, pl->opts.img_c, 16, false);
      if (!rc && pl->opts.img_c == 3) {
        uint8_t* img = (uint8_t*)pl->tcw.alloc("conv_out#rows", conv_rows_image_bytes(4, 32));
        CIC_REQUIRE(img, "tc plan: out of device memory");
        rc = conv_rows_pack(w.ptr("conv_out/kernel"), img, 4, 32, 3, nullptr);
      }
      break;
    }
    case CIC_PLAN_RD:  // conv2 (32 -> 64, k3 s2) on the tensor cores, split-bf16 (rd_params are compared at 2e-5)
      rc = pack(pl, "conv2", w.ptr("conv2/kernel"), 9 * 32, 64, 64, true);
      if (!rc) {  // conv1 (1 -> 32, k3 s2) as the resident weight image of first_conv_tc.cu
        uint8_t* img = (uint8_t*)pl->tcw.alloc("conv1#fc", first_conv_image_bytes());
        CIC_REQUIRE(img && w.ptr("conv1/kernel"), "tc plan: RD conv1 weights");
        rc = first_conv_pack(w.ptr("conv1/kernel"), 9, img, nullptr);
      }
      break;
    default: break;  // the saliency MLPs stay fp32
  }
  if (rc) return rc;
  CIC_CHECK_CUDA(cudaDeviceSynchronize());
  return CIC_OK;
}

// adaptive model: one conv1 pass serves both encoders (weights and biases side by side, hq | lq)
int build_adaptive_tc(cic_plan* pl) {
  if (pl->opts.img_c != 3) return CIC_OK;
  const float* w0 = pl->hq_enc->w.ptr("conv1/kernel");
  const float* w1 = pl->lq_enc->w.ptr("conv1/kernel");
  const float* b0 = pl->hq_enc->w.ptr("conv1/bias");
  const float* b1 = pl->lq_enc->w.ptr("conv1/bias");
  CIC_REQUIRE(w0 && w1 && b0 && b1, "tc plan: conv1 weights of both encoders");
  uint8_t* img = (uint8_t*)pl->tcw.alloc("conv1x2#img", conv1_tc_image_bytes(2));
  float* bias = (float*)pl->tcw.alloc("conv1x2#bias", 128 * sizeof(float));
  CIC_REQUIRE(img && bias, "tc plan: out of device memory");
  int rc = conv1_tc_pack(w0, w1, img, nullptr);
  if (rc) return rc;
  CIC_CHECK_CUDA(cudaMemcpy(bias, b0, 64 * sizeof(float), cudaMemcpyDeviceToDevice));
  CIC_CHECK_CUDA(cudaMemcpy(bias + 64, b1, 64 * sizeof(float), cudaMemcpyDeviceToDevice));
  // generator tail (conv_out of both generators + dt + blend in one kernel): the two conv_out images back to back
  const float* k0 = pl->hq_gen->w.ptr("conv_out/kernel");
  const float* k1 = pl->lq_gen->w.ptr("conv_out/kernel");
  CIC_REQUIRE(k0 && k1, "tc plan: conv_out weights of both generators");
  const size_t one = conv_rows_image_bytes(4, 32);
  uint8_t* tail = (uint8_t*)pl->tcw.alloc("gen_tail#rows", 2 * one);
  CIC_REQUIRE(tail, "tc plan: out of device memory");
  if ((rc = conv_rows_pack(k0, tail, 4, 32, 3, nullptr))) return rc;
  if ((rc = conv_rows_pack(k1, tail + one, 4, 32, 3, nullptr))) return rc;
  CIC_CHECK_CUDA(cudaDeviceSynchronize());
  return CIC_OK;
}

// ---- layer helpers ----------------------------------------------------------------------------
static int run_tc(Ctx& c, const char* name, TcLayer& L, double flops, double bytes) {
  if (c.dry) return CIC_OK;
  Scope sc(c, name, flops, bytes);
  return tc_run_layer(L, c.st);
}

static TcEpilogue epi_bf16(const float* bias, const float* scale, const float* shift, int act, const ActBuf& out, int ld = 0, int coff = 0) {
  TcEpilogue e;
  e.bias = bias; e.scale = scale; e.shift = shift; e.act = act;
  e.out_mode = TC_OUT_BF16; e.out_hi = out.hi; e.out_lo = out.lo; e.out_ld = ld; e.out_coff = coff;
  return e;
}

static int conv_tc(Ctx& c, const char* name, int kind, const TcAct& s0, const TcAct* s1, int batch, int H, int W, int kh, int kw,
                   int stride, const TcMat& wm, int N, bool split, const TcEpilogue& e) {
  TcLayer L;
  L.kind = kind;
  L.src[0] = s0;
  L.nsrc = 1;
  if (s1) { L.src[1] = *s1; L.nsrc = 2; }
  L.batch = batch; L.H = H; L.W = W; L.kh = kh; L.kw = kw;
  L.pad_t = kind == TC_DECONV_K4S2 ? 0 : same_pad_before(H, kh, stride);
  L.pad_l = kind == TC_DECONV_K4S2 ? 0 : same_pad_before(W, kw, stride);
  L.w = wm; L.N = N; L.split = split; L.epi = e;
  const int cin = s0.C + (s1 ? s1->C : 0);
  const double mrows = kind == TC_DECONV_K4S2 ? 4.0 * batch * H * W : (double)batch * same_out(H, stride) * same_out(W, stride);
  const double taps = kind == TC_DECONV_K4S2 ? 4 : kh * kw;
  return run_tc(c, name, L, 2.0 * mrows * N * taps * cin, 2.0 * ((double)batch * H * W * cin + mrows * N));
}

// Dense as a GEMM over [batch][K] rows; splits > 1 writes fp32 partials [splits][batch][N]
static int dense_tc(Ctx& c, const char* name, const TcAct& x, int batch, const TcMat& wm, int N, bool split, int splits,
                    const TcEpilogue& e) {
  TcLayer L;
  L.kind = TC_CONV_S1;
  L.src[0] = x; L.nsrc = 1;
  L.batch = batch; L.H = 1; L.W = 1; L.kh = 1; L.kw = 1;
  L.w = wm; L.N = N; L.split = split; L.splits = splits; L.epi = e;
  return run_tc(c, name, L, 2.0 * batch * (double)x.C * N, 2.0 * ((double)x.C * N * (split ? 2 : 1) + (double)batch * (x.C + N)));
}

static int dense_splits(int batch, int N_pad, int K, bool split) {
  const int bk = K % 64 == 0 ? 64 : 32;
  const long long mt = (batch + 127) / 128;
  static const int pair_env = CIC_KNOB("CIC_TC_PAIR", 1);
  const bool pair = pair_env && tc_pair_ok(bk, mt, N_pad, N_pad);
  const int bn = pair ? tc2_pick_block_n(N_pad) : tc_pick_block_n(N_pad, split, bk);
  const long long tiles = (pair ? (mt + 1) / 2 : mt) * (N_pad / (bn > 0 ? bn : 16));
  const int kblocks = K / bk;
  int s = 1;
  const int target = pair ? sm_count() / 2 : 2 * sm_count();  // one wave of CTA pairs / two waves of CTAs
  if (tiles < target) s = pair ? (int)(target / tiles) : (int)((target + tiles - 1) / tiles);
  if (s > kblocks / 4) s = kblocks / 4;
  if (s < 1) s = 1;
  // no empty split: ceil(kblocks / s) * (s - 1) < kblocks
  while (s > 1 && ((kblocks + s - 1) / s) * (s - 1) >= kblocks) --s;
  return s;
}

// ---- SelfAttention (GAN_functions.py:344-369) on split-bf16 operands ----------------------------------
static int attention_tc(cic_plan* pl, Ctx& c, const ActBuf& x, const ActBuf& y, int batch, int tokens, int C) {
  const WeightStore& w = pl->w;
  const int dq = C / 8;
  CIC_REQUIRE(tokens % 32 == 0, "attention (tc): token count %d must be a multiple of 32", tokens);
  static const int fused_env = CIC_KNOB("CIC_ATTN_FUSED", 1);
  if (fused_env && tokens % 128 == 0 && C == 256 && dq == 32) {
    // projections as GEMMs over the whole batch, then one fused kernel for softmax(q k^T) v (attn_fused.cu)
    const size_t mk = c.arena.mark();
    ActBuf qk = alloc_act(c, (size_t)batch * tokens * 2 * dq, true);
    ActBuf vt = alloc_act(c, (size_t)batch * C * tokens, true);
    int rc = CIC_OK;
    const float* bqkv = w.ptr("attn/qkv/bias");
    if (!c.dry) {
      if ((rc = conv_tc(c, "attn_qk", TC_CONV_S1, view(x, C), nullptr, batch * tokens, 1, 1, 1, 1, 1, mat(pl, "attn/qk", C, 2 * dq), 2 * dq, true,
                        epi_bf16(bqkv, nullptr, nullptr, CIC_ACT_NONE, qk)))) return rc;
      {
        TcEpilogue e;
        e.bias = bqkv + 2 * dq; e.out_mode = TC_OUT_BF16_T; e.out_hi = vt.hi; e.out_lo = vt.lo;
        TcLayer L;
        L.kind = TC_CONV_S1; L.src[0] = view(x, C); L.nsrc = 1;
        L.batch = batch; L.H = tokens; L.W = 1; L.kh = 1; L.kw = 1;
        L.w = mat(pl, "attn/v", C, C); L.N = C; L.split = true; L.epi = e;
        if ((rc = run_tc(c, "attn_v", L, 2.0 * batch * tokens * (double)C * C, 0))) return rc;
      }
      Scope sc(c, "attn_core", 2.0 * batch * tokens * (double)tokens * (dq + C), 2.0 * 2 * 2 * (double)batch * tokens * C);
      if ((rc = launch_attn_fused(qk.hi, qk.lo, vt.hi, vt.lo, x.hi, x.lo, y.hi, y.lo, nullptr, pl->attn_gamma, batch, tokens, c.st))) return rc;
    }
    c.arena.release(mk);
    return rc;
  }
  static const int chunk_env = CIC_KNOB("CIC_ATTN_CHUNK", 0);
  const int chunk_max = chunk_env > 0 ? chunk_env : 128;  // images per pass: bounds the tokens x tokens score workspace
  const int chunk = batch < chunk_max ? batch : chunk_max;
  const size_t mk = c.arena.mark();
  ActBuf qk = alloc_act(c, (size_t)chunk * tokens * 2 * dq, true);
  ActBuf vt = alloc_act(c, (size_t)chunk * C * tokens, true);
  float* S = c.arena.f32((size_t)chunk * tokens * tokens);
  ActBuf P = alloc_act(c, (size_t)chunk * tokens * tokens, true);
  int rc = CIC_OK;
  const float* bqkv = w.ptr("attn/qkv/bias");
  for (int b0 = 0; b0 < batch && !c.dry; b0 += chunk) {
    const int nb = batch - b0 < chunk ? batch - b0 : chunk;
    const size_t xo = (size_t)b0 * tokens * C;
    ActBuf xb{x.hi + xo, x.lo + xo}, yb{y.hi + xo, y.lo + xo};
    // q | k = x Wqk + b   (1x1 convs, :346-347)
    if ((rc = conv_tc(c, "attn_qk", TC_CONV_S1, view(xb, C), nullptr, nb * tokens, 1, 1, 1, 1, 1, mat(pl, "attn/qk", C, 2 * dq), 2 * dq, true,
                      epi_bf16(bqkv, nullptr, nullptr, CIC_ACT_NONE, qk)))) break;
    // v^T[b][c][t]  (:348)
    {
      TcEpilogue e;
      e.bias = bqkv + 2 * dq; e.out_mode = TC_OUT_BF16_T; e.out_hi = vt.hi; e.out_lo = vt.lo;
      TcLayer L;
      L.kind = TC_CONV_S1; L.src[0] = view(xb, C); L.nsrc = 1;
      L.batch = nb; L.H = tokens; L.W = 1; L.kh = 1; L.kw = 1;
      L.w = mat(pl, "attn/v", C, C); L.N = C; L.split = true; L.epi = e;
      if ((rc = run_tc(c, "attn_v", L, 2.0 * nb * tokens * (double)C * C, 0))) break;
    }
    // S[b] = q[b] k[b]^T, no 1/sqrt(d) (:358)
    {
      TcEpilogue e;
      e.out_mode = TC_OUT_F32; e.out_hi = S; e.out_ld = tokens;
      TcLayer L;
      L.kind = TC_CONV_S1; L.src[0] = TcAct{qk.hi, qk.lo, dq, 2 * dq, 0}; L.nsrc = 1;
      L.batch = nb; L.H = tokens; L.W = 1; L.kh = 1; L.kw = 1;
      TcMat km{};
      km.hi = qk.hi + dq; km.lo = qk.lo + dq; km.K = dq; km.rows = tokens; km.row_stride = 2 * dq; km.batches = nb;
      km.batch_stride = (long long)tokens * 2 * dq;
      L.w = km; L.N = tokens; L.split = true; L.b_batched = true; L.epi = e;
      if ((rc = run_tc(c, "attn_qkT", L, 2.0 * nb * tokens * (double)tokens * dq, 0))) break;
    }
    {  // softmax over keys (:359)
      Scope sc(c, "attn_softmax", 0, 8.0 * nb * tokens * tokens);
      if ((rc = tc_softmax_rows_split(S, P.hi, P.lo, (long long)nb * tokens, tokens, c.st))) break;
    }
    // y = gamma * (P v) + x  (:362-367)
    {
      TcEpilogue e = epi_bf16(nullptr, nullptr, nullptr, CIC_ACT_NONE, yb);
      e.alpha = pl->attn_gamma; e.res_hi = xb.hi; e.res_lo = xb.lo;
      TcLayer L;
      L.kind = TC_CONV_S1; L.src[0] = TcAct{P.hi, P.lo, tokens, tokens, 0}; L.nsrc = 1;
      L.batch = nb; L.H = tokens; L.W = 1; L.kh = 1; L.kw = 1;
      TcMat vm{};
      vm.hi = vt.hi; vm.lo = vt.lo; vm.K = tokens; vm.rows = C; vm.row_stride = tokens; vm.batches = nb;
      vm.batch_stride = (long long)C * tokens;
      L.w = vm; L.N = C; L.split = true; L.b_batched = true; L.epi = e;
      if ((rc = run_tc(c, "attn_pv", L, 2.0 * nb * tokens * (double)tokens * C, 0))) break;
    }
  }
  c.arena.release(mk);
  return rc;
}

// ---- encoder ----------------------------------------------------------------------------------
struct EncSkips {  // bf16 skip tensors (hi, lo) that outlive the encoder call: x1 (H/2, 64), x2 (H/4, 128), x3 (H/8, 256)
  ActBuf x1, x2, x3;
};

static EncSkips alloc_skips(Ctx& c, size_t px) {
  EncSkips s;
  s.x1 = alloc_act(c, px / 4 * 64, true);
  s.x2 = alloc_act(c, px / 16 * 128, true);
  s.x3 = alloc_act(c, px / 64 * 256, true);
  return s;
}

// part: 0 = whole encoder, 1 = convolutions only (conv4's output goes to *x4_ext), 2 = Flatten + Dense only (reads *x4_ext):
// the phased adaptive forward runs the convolutions per chunk of the batch and the Dense layer once per batch.
enum { ENC_ALL = 0, ENC_CONVS = 1, ENC_DENSE = 2 };
static int encoder_dense_tc(cic_plan* pl, Ctx& c, const ActBuf& x4, float* latent, int B);

static int encoder_core_tc(cic_plan* pl, Ctx& c, const float* img, float* latent, float* x1_f32, const EncSkips& sk, int B,
                           const TileMap& tm = TileMap(), bool x1_ready = false, int part = ENC_ALL, const ActBuf* x4_ext = nullptr) {
  const WeightStore& w = pl->w;
  const int H = pl->opts.img_h, W = pl->opts.img_w, C = pl->opts.img_c, L = pl->opts.latent_dim;
  const size_t px = (size_t)B * H * W;
  const size_t mk = c.arena.mark();
  int rc;
  (void)L;
  if (part == ENC_DENSE) {
    CIC_REQUIRE(x4_ext, "encoder (tc): dense part needs conv4's output");
    rc = encoder_dense_tc(pl, c, *x4_ext, latent, B);
    c.arena.release(mk);
    return rc;
  }
  // conv1 (3 -> 64, k4 s2) + LeakyReLU: K = 48 cannot feed a tensor-core K block; a direct CUDA-core kernel writes
  // the (hi, lo) bf16 pair the split-bf16 layers read (:300-302)
  if (x1_ready) {
    // sk.x1 was produced by the shared conv1 pass of the adaptive model
  } else if (C == 3 && !x1_f32) {
    if (!c.dry) {  // tensor cores, im2col built in shared memory (conv1_tc.cu)
      Scope sc(c, "conv1", 2.0 * (px / 4) * 64 * 16 * C, 4.0 * px * C + 4.0 * px / 4 * 64);
      bf16* oh[1] = {sk.x1.hi};
      bf16* ol[1] = {sk.x1.lo};
      if ((rc = launch_conv1_tc(img, (const uint8_t*)pl->tcw.ptr("conv1#img"), w.ptr("conv1/bias"), 1, oh, ol, B, H, W, tm, c.st))) return rc;
    }
  } else if (C == 3) {
    if (!c.dry) {  // fp32 copy of x1 requested: the CUDA-core kernel writes it in the same pass
      Scope sc(c, "conv1", 2.0 * (px / 4) * 64 * 16 * C, 4.0 * px * C + 4.0 * px / 4 * 64);
      if ((rc = launch_conv_k4s2_c3(img, w.ptr("conv1/kernel"), w.ptr("conv1/bias"), sk.x1.hi, sk.x1.lo, x1_f32, B, H, W, CIC_ACT_LRELU02, tm, c.st))) return rc;
    }
  } else {
    CIC_REQUIRE(!tm.tiles_x, "encoder (tc): tiled input needs a 3-channel image");
    float* x1f = x1_f32 ? x1_f32 : c.arena.f32(px / 4 * 64);
    if (!c.dry) {
      Scope sc(c, "conv1", 2.0 * (px / 4) * 64 * 16 * C, 4.0 * (px * C + px / 4 * 64) + 4.0 * px / 4 * 64);
      IGemmParams p{};
      p.src[0] = ConvSrc{img, C, C, 0};
      p.nsrc = 1; p.Cin = C; p.batch = B; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
      p.kh = 4; p.kw = 4; p.stride = 2; p.pad_t = same_pad_before(H, 4, 2); p.pad_l = same_pad_before(W, 4, 2);
      p.Bmat = w.ptr("conv1/kernel"); p.N = 64; p.ldb = 64; p.bias = w.ptr("conv1/bias"); p.act = CIC_ACT_LRELU02; p.alpha = 1.f;
      p.out = x1f; p.out_ld = 64; p.out_H = p.Ho; p.out_W = p.Wo; p.out_ys = p.out_xs = 1; p.splits = 1;
      if ((rc = launch_igemm(p, c.st))) return rc;
      if ((rc = tc_split_f32(x1f, sk.x1.hi, sk.x1.lo, px / 4 * 64, c.st))) return rc;
    }
  }
  // conv2..conv3: k4 s2 + BN + LeakyReLU, 3-term split-bf16 (:304-312)
  if ((rc = conv_tc(c, "conv2", TC_CONV_S2, view(sk.x1, 64), nullptr, B, H / 2, W / 2, 4, 4, 2, mat(pl, "conv2", 16 * 64, 128), 128, true,
                    epi_bf16(fused_bias(pl, "conv2"), nullptr, nullptr, CIC_ACT_LRELU02, sk.x2)))) return rc;
  if ((rc = conv_tc(c, "conv3", TC_CONV_S2, view(sk.x2, 128), nullptr, B, H / 4, W / 4, 4, 4, 2, mat(pl, "conv3", 16 * 128, 256), 256, true,
                    epi_bf16(fused_bias(pl, "conv3"), nullptr, nullptr, CIC_ACT_LRELU02, sk.x3)))) return rc;
  ActBuf x3a = sk.x3;
  if (pl->opts.add_attention) {  // the skip is tapped before attention (:312 vs :318)
    x3a = alloc_act(c, px / 64 * 256, true);
    if ((rc = attention_tc(pl, c, sk.x3, x3a, B, (H / 8) * (W / 8), 256))) return rc;
  }
  ActBuf x4 = x4_ext ? *x4_ext : alloc_act(c, px / 256 * 512, true);
  if ((rc = conv_tc(c, "conv4", TC_CONV_S2, view(x3a, 256), nullptr, B, H / 8, W / 8, 4, 4, 2, mat(pl, "conv4", 16 * 256, 512), 512, true,
                    epi_bf16(fused_bias(pl, "conv4"), nullptr, nullptr, CIC_ACT_LRELU02, x4)))) return rc;
  if (part == ENC_CONVS) {
    c.arena.release(mk);
    return CIC_OK;
  }
  rc = encoder_dense_tc(pl, c, x4, latent, B);
  c.arena.release(mk);
  return rc;
}

// Flatten (NHWC) + Dense (:325-326): split-K GEMM, partials reduced in a fixed order
static int encoder_dense_tc(cic_plan* pl, Ctx& c, const ActBuf& x4, float* latent, int B) {
  const WeightStore& w = pl->w;
  const int H = pl->opts.img_h, W = pl->opts.img_w, L = pl->opts.latent_dim;
  int rc;
  const size_t mk = c.arena.mark();
  const int feat = (H / 16) * (W / 16) * 512, Lp = round_up(L, 16);
  const int splits = dense_splits(B, Lp, feat, true);
  float* part = splits > 1 ? c.arena.f32((size_t)splits * B * L) : nullptr;
  TcEpilogue e;
  e.bias = w.ptr("dense/bias"); e.act = CIC_ACT_NONE;
  if (splits > 1) { e.out_mode = TC_OUT_PARTIAL; e.out_hi = part; e.bias = nullptr; }
  else { e.out_mode = TC_OUT_F32; e.out_hi = latent; e.out_ld = L; }
  if ((rc = dense_tc(c, "dense", view(x4, feat), B, mat(pl, "dense", feat, Lp), L, true, splits, e))) return rc;
  if (splits > 1 && !c.dry) {
    Scope sc(c, "dense_reduce", 0, 4.0 * (splits + 1) * B * L);
    if ((rc = tc_splitk_reduce(part, splits, B, L, w.ptr("dense/bias"), nullptr, nullptr, CIC_ACT_NONE, latent, nullptr, nullptr, c.st))) return rc;
  }
  c.arena.release(mk);
  return CIC_OK;
}

int encoder_forward_tc(cic_plan* pl, Ctx& c, const float* img, float* latent, float* x1, float* x2, float* x3, int B) {
  const size_t px = (size_t)B * pl->opts.img_h * pl->opts.img_w;
  EncSkips sk = alloc_skips(c, px);
  int rc = encoder_core_tc(pl, c, img, latent, x1, sk, B);
  if (rc || c.dry) return rc;
  if (x2 && (rc = tc_join_to_f32(sk.x2.hi, sk.x2.lo, x2, px / 16, 128, 128, 0, c.st))) return rc;
  if (x3 && (rc = tc_join_to_f32(sk.x3.hi, sk.x3.lo, x3, px / 64, 256, 256, 0, c.st))) return rc;
  return CIC_OK;
}

// ---- generator --------------------------------------------------------------------------------
// latent fp32 (B, L); skips as bf16 NHWC (hi only is read)
// part: 0 = whole generator, 1 = Dense + BN + LeakyReLU only (output to *g0_ext), 2 = transposed convs + conv_out only (reads *g0_ext)
enum { GEN_ALL = 0, GEN_DENSE = 1, GEN_CONVS = 2 };
// g4_ext != nullptr: stop after deconv4 and leave its output (256^2 x 32 bf16 per tile) there - the adaptive model runs conv_out of
// both generators and the blend as one kernel (launch_gen_tail)
static int generator_core_tc(cic_plan* pl, Ctx& c, const float* latent, const bf16* s1, const bf16* s2, const bf16* s3, float* out, int B,
                             const TileMap& tm = TileMap(), int part = GEN_ALL, const ActBuf* g0_ext = nullptr, const ActBuf* g4_ext = nullptr) {
  const WeightStore& w = pl->w;
  const int H = pl->opts.img_h, W = pl->opts.img_w, C = pl->opts.img_c, L = pl->opts.latent_dim;
  const int h16 = H / 16, w16 = W / 16, feat = h16 * w16 * 512;
  const size_t px = (size_t)B * H * W;
  CIC_REQUIRE(C <= 16, "generator (tc): at most 16 output channels");
  const size_t mk = c.arena.mark();
  CIC_REQUIRE(part == GEN_ALL || g0_ext, "generator (tc): the split parts need the Dense output buffer");
  ActBuf g0 = g0_ext ? *g0_ext : alloc_act(c, (size_t)B * feat, false);
  ActBuf g1 = alloc_act(c, px / 64 * 256, false);
  ActBuf g2 = alloc_act(c, px / 16 * 128, false);
  ActBuf g3 = alloc_act(c, px / 4 * 64, false);
  ActBuf g4 = g4_ext ? *g4_ext : alloc_act(c, px * 32, false);
  int rc;
  // :247-250 Dense -> Reshape(h16, w16, 512) NHWC -> BN -> LeakyReLU
  if (part == GEN_CONVS) {
    // g0 was produced by the Dense part
  } else if (L % 32 == 0) {
    ActBuf lat = alloc_act(c, (size_t)B * L, false);
    if (!c.dry && (rc = tc_split_f32(latent, lat.hi, nullptr, (size_t)B * L, c.st))) return rc;
    if ((rc = dense_tc(c, "dense", view(lat, L), B, mat(pl, "dense", L, feat), feat, false, 1,
                       epi_bf16(fused_bias(pl, "dense"), nullptr, nullptr, CIC_ACT_LRELU02, g0)))) return rc;
  } else {  // latent sizes a K block cannot take: fp32 CUDA-core GEMM, then to bf16
    float* g0f = c.arena.f32((size_t)B * feat);
    const size_t wsb = cic_dense_workspace_bytes(B, L, feat);
    float* ws = wsb ? (float*)c.arena.alloc_bytes(wsb) : nullptr;
    if (!c.dry) {
      Scope sc(c, "dense", 2.0 * B * (double)L * feat, 4.0 * ((double)L * feat + (double)B * feat));
      if ((rc = run_dense(latent, w.ptr("dense/kernel"), w.ptr("dense/bias"), w.ptr("bn0/scale"), w.ptr("bn0/shift"), g0f, B, L, feat,
                          CIC_ACT_LRELU02, ws, wsb / sizeof(float), c.st))) return rc;
      if ((rc = tc_split_f32(g0f, g0.hi, nullptr, (size_t)B * feat, c.st))) return rc;
    }
  }
  if (part == GEN_DENSE) {
    c.arena.release(mk);
    return CIC_OK;
  }
  // :253-270 four Conv2DTranspose(k4, s2) + BN + LeakyReLU, the skips concatenated on the channel axis
#define DC(i, s0v, s1p, hh, ww, cin, co, dst)                                                                             \
  if ((rc = conv_tc(c, "deconv" #i, TC_DECONV_K4S2, s0v, s1p, B, hh, ww, 2, 2, 1, mat(pl, "deconv" #i, 4 * (cin), 4 * (co)), co, false, \
                    epi_bf16(fused_bias(pl, "deconv" #i), nullptr, nullptr, CIC_ACT_LRELU02, dst)))) return rc
  TcAct k3{s3, nullptr, 256, 256, 0}, k2{s2, nullptr, 128, 128, 0}, k1{s1, nullptr, 64, 64, 0};
  DC(1, view(g0, 512), nullptr, h16, w16, 512, 256, g1);
  DC(2, view(g1, 256), &k3, 2 * h16, 2 * w16, 512, 128, g2);
  DC(3, view(g2, 128), &k2, 4 * h16, 4 * w16, 256, 64, g3);
  DC(4, view(g3, 64), &k1, 8 * h16, 8 * w16, 128, 32, g4);
#undef DC
  if (g4_ext) {
    c.arena.release(mk);
    return CIC_OK;
  }
  // :273 Conv2D(3, k4, 'same', tanh): pad 1 before / 2 after
  static const int no_rows = CIC_KNOB("CIC_TC_NO_ROWS", 0);
  if (C == 3 && !no_rows) {  // column-strip formulation (kx folded into K, ky into N)
    if (!c.dry) {
      Scope sc(c, "conv_out", 2.0 * px * 16 * 32 * C, 2.0 * px * 32 + 4.0 * px * C);
      TcAct src = view(g4, 32);
      if ((rc = launch_conv_rows_tc(&src, 1, (const uint8_t*)pl->tcw.ptr("conv_out#rows"), w.ptr("conv_out/bias"), 4, C, CIC_ACT_TANH, out, B, H, W, tm, c.st))) return rc;
    }
    c.arena.release(mk);
    return CIC_OK;
  }
  TcEpilogue e;
  e.bias = w.ptr("conv_out/bias"); e.act = CIC_ACT_TANH; e.out_mode = TC_OUT_F32; e.out_hi = out; e.out_ld = C;
  e.tm_tx = tm.tiles_x; e.tm_ty = tm.tiles_y; e.tm_IH = tm.IH; e.tm_IW = tm.IW;  // write straight into the image layout
  if ((rc = conv_tc(c, "conv_out", TC_CONV_S1, view(g4, 32), nullptr, B, H, W, 4, 4, 1, mat(pl, "conv_out", 16 * 32, 16), C, false, e))) return rc;
  c.arena.release(mk);
  return CIC_OK;
}

int generator_forward_tc(cic_plan* pl, Ctx& c, const float* latent, const float* s1, const float* s2, const float* s3,
                         float* out, int B) {
  const size_t px = (size_t)B * pl->opts.img_h * pl->opts.img_w;
  ActBuf b1 = alloc_act(c, px / 4 * 64, false), b2 = alloc_act(c, px / 16 * 128, false), b3 = alloc_act(c, px / 64 * 256, false);
  int rc;
  if (!c.dry) {
    if ((rc = tc_split_f32(s1, b1.hi, nullptr, px / 4 * 64, c.st))) return rc;
    if ((rc = tc_split_f32(s2, b2.hi, nullptr, px / 16 * 128, c.st))) return rc;
    if ((rc = tc_split_f32(s3, b3.hi, nullptr, px / 64 * 256, c.st))) return rc;
  }
  return generator_core_tc(pl, c, latent, b1.hi, b2.hi, b3.hi, out, B);
}

// ---- autoencoder (train_autoencoder.py:9-40), single-pass bf16 ------------------------------------------
int autoencoder_forward_tc(cic_plan* pl, Ctx& c, const float* x, float* y, uint8_t* y_u8, int B, int H, int W) {
  const WeightStore& w = pl->w;
  const int C = pl->opts.img_c;
  if (C != 3) return autoencoder_forward_f32(pl, c, x, y, y_u8, B, H, W);  // the direct first layer is written for 3 channels
  const size_t px = (size_t)B * H * W;
  const size_t mk = c.arena.mark();
  ActBuf x1 = alloc_act(c, px * 32, false), x1p = alloc_act(c, px / 4 * 32, false), x2 = alloc_act(c, px / 4 * 64, false);
  ActBuf enc = alloc_act(c, px / 16 * 64, false), y3u = alloc_act(c, px / 4 * 64, false), x2r = alloc_act(c, px / 4 * 64, false);
  ActBuf y5u = alloc_act(c, px * 32, false), x1r = alloc_act(c, px * 32, false);
  int rc;
  if (!c.dry) {  // :14-15 Conv2D(32, relu) + MaxPooling2D fused: writes x1 and its pooled copy
    Scope sc(c, "conv1", 2.0 * px * 32 * 27, 4.0 * px * 3 + 2.0 * px * 32 * 1.25);
    const uint8_t* fc = (const uint8_t*)pl->tcw.ptr("conv1#fc");
    if (fc && H % 2 == 0 && W % 2 == 0 && CIC_KNOB("CIC_FIRST_TC", 1))
      rc = launch_first_conv_tc_pool(x, fc, w.ptr("conv1/bias"), x1.hi, x1p.hi, B, H, W, CIC_ACT_RELU, c.st);
    else
      rc = launch_conv_k3s1_c3_pool(x, w.ptr("conv1/kernel"), w.ptr("conv1/bias"), x1.hi, x1p.hi, B, H, W, CIC_ACT_RELU, c.st);
    if (rc) return rc;
  }
#define AE_CONV(name, s0, s1p, hh, ww, cin, co, outbuf, up)                                                                   \
  {                                                                                                                           \
    TcEpilogue e = epi_bf16(w.ptr(name "/bias"), nullptr, nullptr, CIC_ACT_RELU, outbuf);                                     \
    e.up2 = up;                                                                                                               \
    if ((rc = conv_tc(c, name, TC_CONV_S1, s0, s1p, B, hh, ww, 3, 3, 1, mat(pl, name, 9 * (cin), co), co, false, e))) return rc; \
  }
  AE_CONV("conv2", view(x1p, 32), nullptr, H / 2, W / 2, 32, 64, x2, 0);                      // :17
  if (!c.dry) {                                                                               // :18
    Scope sc(c, "pool2", 0, 2.0 * px / 4 * 64 * 1.25);
    if ((rc = tc_maxpool2x2_bf16(x2.hi, enc.hi, B, H / 2, W / 2, 64, c.st))) return rc;
  }
  AE_CONV("conv3", view(enc, 64), nullptr, H / 4, W / 4, 64, 64, y3u, 1);                     // :21 + :22 UpSampling2D in the store
  AE_CONV("conv_x2", view(x2, 64), nullptr, H / 2, W / 2, 64, 64, x2r, 0);                    // :25
  TcAct x2rv = view(x2r, 64), x1rv = view(x1r, 32);
  AE_CONV("conv5", view(y3u, 64), &x2rv, H / 2, W / 2, 128, 32, y5u, 1);                      // :26 concat, :28, :29 UpSampling2D
  AE_CONV("conv_x1", view(x1, 32), nullptr, H, W, 32, 32, x1r, 0);                            // :32
#undef AE_CONV
  static const int no_rows = CIC_KNOB("CIC_TC_NO_ROWS", 0);
  if (!no_rows) {                                                                              // :33 concat, :35 Conv2D(3, sigmoid)
    if (!c.dry) {
      Scope sc(c, "conv_out", 2.0 * px * 9 * 64 * C, 2.0 * px * 64 + 4.0 * px * C);
      TcAct srcs[2] = {view(y5u, 32), x1rv};
      if ((rc = launch_conv_rows_tc(srcs, 2, (const uint8_t*)pl->tcw.ptr("conv_out#rows"), w.ptr("conv_out/bias"), 3, C, CIC_ACT_SIGMOID, y, B, H, W, TileMap(), c.st))) return rc;
    }
  } else {
    TcEpilogue e;
    e.bias = w.ptr("conv_out/bias"); e.act = CIC_ACT_SIGMOID; e.out_mode = TC_OUT_F32; e.out_hi = y; e.out_ld = C;
    if ((rc = conv_tc(c, "conv_out", TC_CONV_S1, view(y5u, 32), &x1rv, B, H, W, 3, 3, 1, mat(pl, "conv_out", 9 * 64, 16), C, false, e))) return rc;
  }
  c.arena.release(mk);
  if (c.dry) return CIC_OK;
  Scope sc(c, "cast_u8", 0, 5.0 * px * C);
  if (y_u8) rc = cic_f32_to_u8_trunc(y, y_u8, px * C, 255.0f, c.st);                          // test_autoencoder.py:88
  return rc;
}

// ---- RD optimizer (GAN_functions.py:495-557): conv1 direct, conv2 on the tensor cores, the rest fp32 -------------------
int rd_forward_tc(cic_plan* pl, Ctx& c, const float* mask, const float* bpp, float* rd_params, int B, const TileMap& tm) {
  const int H = pl->opts.img_h, W = pl->opts.img_w;
  if (H % 4 || W % 4) {
    CIC_REQUIRE(!tm.tiles_x, "rd (tc): tiled input needs H, W divisible by 4");
    return rd_forward_f32(pl, c, mask, bpp, rd_params, B);
  }
  const WeightStore& w = pl->w;
  const int h2 = H / 2, w2 = W / 2, h4 = H / 4, w4 = W / 4;
  ActBuf r1 = alloc_act(c, (size_t)B * h2 * w2 * 32, true);
  float* r2 = c.arena.f32((size_t)B * h4 * w4 * 64);
  float* feat = c.arena.f32((size_t)B * 65);
  float* d1 = c.arena.f32((size_t)B * 128);
  float* base = c.arena.f32((size_t)B * 3);
  int rc;
  if (!c.dry) {                                                                                // :511-512
    Scope sc(c, "conv1", 2.0 * B * h2 * w2 * 32 * 9, 4.0 * B * H * W + 4.0 * B * h2 * w2 * 32);
    const uint8_t* fc = (const uint8_t*)pl->tcw.ptr("conv1#fc");
    if (fc && CIC_KNOB("CIC_FIRST_TC", 1))
      rc = launch_first_conv_tc_c1s2(mask, fc, w.ptr("conv1/bias"), r1.hi, r1.lo, B, H, W, CIC_ACT_LRELU02, tm, c.st);
    else
      rc = launch_conv_k3s2_c1(mask, w.ptr("conv1/kernel"), w.ptr("conv1/bias"), r1.hi, r1.lo, B, H, W, CIC_ACT_LRELU02, tm, c.st);
    if (rc) return rc;
  }
  TcEpilogue e;                                                                                // :513-514
  e.bias = w.ptr("conv2/bias"); e.act = CIC_ACT_LRELU02; e.out_mode = TC_OUT_F32; e.out_hi = r2; e.out_ld = 64;
  if ((rc = conv_tc(c, "conv2", TC_CONV_S2, view(r1, 32), nullptr, B, h2, w2, 3, 3, 2, mat(pl, "conv2", 9 * 32, 64), 64, true, e))) return rc;
  if (!c.dry && (rc = launch_global_avg_pool(r2, feat, B, h4 * w4, 64, 65, c.st))) return rc;  // :515
  return launch_rd_tail(pl, c, bpp, feat, d1, base, rd_params, B);                            // :518-541
}

// ---- adaptive model (GAN_functions.py:604-696) --------------------------------------------------------
int adaptive_forward_tc(cic_plan* pl, Ctx& c, const cic_adaptive_io* io, int n_img, int img_h, int img_w, int phase,
                        const cic_adaptive_state* state, int tile0) {
  const int T = pl->opts.img_h, base = pl->opts.latent_dim;
  const int tpi = ((img_h + T - 1) / T) * ((img_w + T - 1) / T);
  const int nt = n_img * tpi;
  // sizes that are not a multiple of the model tile: the last tile row / column replicates the image edge on load (conv1, RD conv1)
  // and is cropped on store (conv_out); dt, blend and hq_ratio run on the real image
  const bool tiled = tpi > 1 || img_h != T || img_w != T;
  const size_t tpx = (size_t)nt * T * T;
  int rc;
  const bool all = phase == 0;
  const bool do_enc = all || phase == CIC_PHASE_ENCODE, do_lat = all || phase == CIC_PHASE_LATENT, do_dec = all || phase == CIC_PHASE_DECODE;
  CIC_REQUIRE(all || state, "adaptive (tc): the phased forward needs the state buffers");
  // tiles are addressed in place in the image layout by the first (conv1, RD conv1) and last (conv_out) layers
  TileMap tm;
  if (tiled) { tm.tiles_x = (img_w + T - 1) / T; tm.tiles_y = (img_h + T - 1) / T; tm.IH = img_h; tm.IW = img_w; }
  const float* img_t = io->d_img;
  const float* mask_t = io->d_mask;
  float* bpp_t = c.arena.f32(nt);
  float* qs_t = c.arena.f32(nt);
  if (!c.dry && (rc = launch_expand_bpp(io->d_bpp, bpp_t, qs_t, nt, tpi, c.st))) return rc;              // :631-649
  float* hq_lat = io->d_hq_latent ? io->d_hq_latent : c.arena.f32((size_t)nt * 2 * base);
  float* lq_lat = io->d_lq_latent ? io->d_lq_latent : c.arena.f32((size_t)nt * base);
  // skip tensors, conv4 outputs and generator Dense outputs: in the workspace for the one-call forward, in the caller's
  // batch-wide state buffers (at the chunk's tile offset) for the phased one
  const size_t e1 = (size_t)(T / 2) * (T / 2) * 64, e2 = (size_t)(T / 4) * (T / 4) * 128, e3 = (size_t)(T / 8) * (T / 8) * 256,
               e4 = (size_t)(T / 16) * (T / 16) * 512;
  EncSkips sk[2];
  ActBuf x4[2], g0[2];
  for (int e = 0; e < 2; ++e) {
    if (all) {
      sk[e] = alloc_skips(c, tpx);
    } else {
      sk[e].x1.hi = (bf16*)state->x1[e] + (size_t)tile0 * e1;
      sk[e].x2.hi = (bf16*)state->x2[e] + (size_t)tile0 * e2;
      sk[e].x3.hi = (bf16*)state->x3[e] + (size_t)tile0 * e3;
      x4[e].hi = (bf16*)state->x4_hi[e] + (size_t)tile0 * e4;
      x4[e].lo = (bf16*)state->x4_lo[e] + (size_t)tile0 * e4;
      g0[e].hi = (bf16*)state->g0[e] + (size_t)tile0 * e4;
      if (do_enc) {  // the low parts only live inside the encoders
        sk[e].x1.lo = (bf16*)c.arena.alloc_bytes((size_t)nt * e1 * sizeof(bf16));
        sk[e].x2.lo = (bf16*)c.arena.alloc_bytes((size_t)nt * e2 * sizeof(bf16));
        sk[e].x3.lo = (bf16*)c.arena.alloc_bytes((size_t)nt * e3 * sizeof(bf16));
      }
    }
  }
  const EncSkips &hs = sk[0], &ls = sk[1];
  size_t mk = c.arena.mark();
  if (do_enc) {
    // 1-2. encoder convolutions (:604-617); skips stay on the device as bf16 for the generators
    const bool shared_conv1 = pl->opts.img_c == 3 && pl->tcw.ptr("conv1x2#img");
    if (shared_conv1 && !c.dry) {  // conv1 of both encoders in one pass over the image (:300-302 of both build_encoder calls)
      if (c.prof) c.prof->prefix = "";
      Scope sc(c, "enc_conv1_x2", 2.0 * 2 * (tpx / 4) * 64 * 48, 4.0 * tpx * 3 + 2 * 4.0 * tpx / 4 * 64);
      bf16* oh[2] = {hs.x1.hi, ls.x1.hi};
      bf16* ol[2] = {hs.x1.lo, ls.x1.lo};
      if ((rc = launch_conv1_tc(img_t, (const uint8_t*)pl->tcw.ptr("conv1x2#img"), (const float*)pl->tcw.ptr("conv1x2#bias"), 2, oh, ol, nt, T, T,
                                tm, c.st))) return rc;
    }
    if (c.prof) c.prof->prefix = "hq_enc/";
    if ((rc = encoder_core_tc(pl->hq_enc.get(), c, img_t, hq_lat, nullptr, hs, nt, tm, shared_conv1, all ? ENC_ALL : ENC_CONVS, all ? nullptr : &x4[0]))) return rc;
    c.arena.release(mk);
    if (c.prof) c.prof->prefix = "lq_enc/";
    if ((rc = encoder_core_tc(pl->lq_enc.get(), c, img_t, lq_lat, nullptr, ls, nt, tm, shared_conv1, all ? ENC_ALL : ENC_CONVS, all ? nullptr : &x4[1]))) return rc;
    c.arena.release(mk);
  }
  if (do_lat && !all) {  // Dense layers of both encoders on the whole batch
    if (c.prof) c.prof->prefix = "hq_enc/";
    if ((rc = encoder_core_tc(pl->hq_enc.get(), c, nullptr, hq_lat, nullptr, hs, nt, tm, true, ENC_DENSE, &x4[0]))) return rc;
    c.arena.release(mk);
    if (c.prof) c.prof->prefix = "lq_enc/";
    if ((rc = encoder_core_tc(pl->lq_enc.get(), c, nullptr, lq_lat, nullptr, ls, nt, tm, true, ENC_DENSE, &x4[1]))) return rc;
    c.arena.release(mk);
  }
  float* hq_q = io->d_hq_latent_q ? io->d_hq_latent_q : c.arena.f32((size_t)nt * 2 * base);
  float* lq_q = io->d_lq_latent_q ? io->d_lq_latent_q : c.arena.f32((size_t)nt * base);
  if (do_lat) {
    // 3. latent saliency (:619-620), fp32
    float* sal_hq = c.arena.f32(nt);
    float* sal_lq = c.arena.f32(nt);
    mk = c.arena.mark();
    if (c.prof) c.prof->prefix = "sal_hq/";
    if ((rc = saliency_forward_f32(pl->sal_hq.get(), c, hq_lat, sal_hq, nt))) return rc;
    c.arena.release(mk);
    if (c.prof) c.prof->prefix = "sal_lq/";
    if ((rc = saliency_forward_f32(pl->sal_lq.get(), c, lq_lat, sal_lq, nt))) return rc;
    c.arena.release(mk);
    // 5. quantise (:661-666)
    if (c.prof) c.prof->prefix = "";
    if (!c.dry) {
      Scope sc(c, "quantize", 0, 12.0 * nt * 3 * base);
      if ((rc = cic_quantize_latent(hq_lat, sal_hq, qs_t, hq_q, io->d_hq_symbols, nullptr, io->d_hq_scale, nt, 2 * base, c.st))) return rc;
      if ((rc = cic_quantize_latent(lq_lat, sal_lq, qs_t, lq_q, io->d_lq_symbols, nullptr, io->d_lq_scale, nt, base, c.st))) return rc;
    }
    if (!all) {  // generator Dense layers on the whole batch
      if (c.prof) c.prof->prefix = "hq_gen/";
      if ((rc = generator_core_tc(pl->hq_gen.get(), c, hq_q, nullptr, nullptr, nullptr, nullptr, nt, tm, GEN_DENSE, &g0[0]))) return rc;
      c.arena.release(mk);
      if (c.prof) c.prof->prefix = "lq_gen/";
      if ((rc = generator_core_tc(pl->lq_gen.get(), c, lq_q, nullptr, nullptr, nullptr, nullptr, nt, tm, GEN_DENSE, &g0[1]))) return rc;
      c.arena.release(mk);
    }
  }
  mk = c.arena.mark();
  // 4. rate-distortion parameters (:624), fp32; an output only
  if (do_enc && (io->d_rd_params || c.dry)) {
    if (c.prof) c.prof->prefix = "rd/";
    if ((rc = rd_forward_tc(pl->rd.get(), c, mask_t, bpp_t, io->d_rd_params, nt, tm))) return rc;
    c.arena.release(mk);
  }
  if (do_dec) {
    const uint8_t* tail_img = (const uint8_t*)pl->tcw.ptr("gen_tail#rows");
    // (the workspace dry run measures the fused path: it keeps both deconv4 outputs alive at once)
    const bool fused_tail = pl->opts.img_c == 3 && tail_img && (io->d_blended || c.dry) && CIC_KNOB("CIC_GEN_TAIL", 1);
    if (fused_tail) {
      // 6-7. generators up to deconv4 (:669-670), then conv_out of both + dynamic threshold + blend in one kernel (:273, :651-657,
      // :682-684): the un-blended images never reach HBM
      ActBuf g4h = alloc_act(c, tpx * 32, false), g4l = alloc_act(c, tpx * 32, false);
      mk = c.arena.mark();
      if (c.prof) c.prof->prefix = "hq_gen/";
      if ((rc = generator_core_tc(pl->hq_gen.get(), c, hq_q, hs.x1.hi, hs.x2.hi, hs.x3.hi, nullptr, nt, tm, all ? GEN_ALL : GEN_CONVS, all ? nullptr : &g0[0], &g4h))) return rc;
      c.arena.release(mk);
      if (c.prof) c.prof->prefix = "lq_gen/";
      if ((rc = generator_core_tc(pl->lq_gen.get(), c, lq_q, ls.x1.hi, ls.x2.hi, ls.x3.hi, nullptr, nt, tm, all ? GEN_ALL : GEN_CONVS, all ? nullptr : &g0[1], &g4l))) return rc;
      c.arena.release(mk);
      if (c.prof) c.prof->prefix = "";
      if (!c.dry) {
        const double px = (double)n_img * img_h * img_w;
        Scope sc(c, "gen_tail", 2.0 * 2 * tpx * 16 * 32 * 3, 2.0 * 2 * tpx * 32 + px * (4.0 + 12.0 + (io->d_dt ? 4.0 : 0.0)));
        if ((rc = launch_gen_tail(view(g4h, 32), view(g4l, 32), tail_img, pl->hq_gen->w.ptr("conv_out/bias"), pl->lq_gen->w.ptr("conv_out/bias"),
                                  io->d_mask, io->d_bpp, io->d_blended, io->d_dt, io->d_hq_ratio_sum, io->d_hq_out, io->d_lq_out, n_img, nt, T, T,
                                  tiled ? tm : TileMap(), c.st))) return rc;
      }
      return CIC_OK;
    }
    // 6. generators (:669-670)
    float* hq_img = io->d_hq_out ? io->d_hq_out : c.arena.f32(tpx * 3);  // image layout
    float* lq_img = io->d_lq_out ? io->d_lq_out : c.arena.f32(tpx * 3);
    mk = c.arena.mark();
    if (c.prof) c.prof->prefix = "hq_gen/";
    if ((rc = generator_core_tc(pl->hq_gen.get(), c, hq_q, hs.x1.hi, hs.x2.hi, hs.x3.hi, hq_img, nt, tm, all ? GEN_ALL : GEN_CONVS, all ? nullptr : &g0[0]))) return rc;
    c.arena.release(mk);
    if (c.prof) c.prof->prefix = "lq_gen/";
    if ((rc = generator_core_tc(pl->lq_gen.get(), c, lq_q, ls.x1.hi, ls.x2.hi, ls.x3.hi, lq_img, nt, tm, all ? GEN_ALL : GEN_CONVS, all ? nullptr : &g0[1]))) return rc;
    c.arena.release(mk);
    // 7. dynamic threshold + blend on whole images (:651-657, :682-684)
    if (c.prof) c.prof->prefix = "";
    if (!c.dry) {
      Scope sc(c, "roi_blend", 0, 44.0 * n_img * img_h * img_w);
      float* blended = io->d_blended;
      if ((rc = cic_roi_mask_blend(blended ? hq_img : nullptr, blended ? lq_img : nullptr, io->d_mask, io->d_bpp, blended,
                                   io->d_dt, io->d_hq_ratio_sum, n_img, img_h * img_w, 3, c.st))) return rc;
    }
  }
  return CIC_OK;
}

}  // namespace cic

using namespace cic;

// ---- stand-alone tensor-core operators (fp32 in / fp32 out; operands converted and packed per call) ------
static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

extern "C" size_t cic_conv2d_tc_workspace_bytes(int batch, int h, int w, int cin, int cin2, int cout, int kh, int kw, int stride,
                                                int transpose) {
  const size_t px = (size_t)batch * h * w;
  const int cp = round_up(cout, 16);
  const size_t K = (size_t)(transpose ? 16 : kh * kw) * (cin + cin2);
  size_t n = 0;
  n += 2 * align256(px * cin * 2) + 2 * align256(px * (cin2 > 0 ? cin2 : 0) * 2 + 256);
  n += 2 * align256(K * cp * 2);
  n += align256(K * cout * 4);  // transposed-conv phase matrices
  (void)stride;
  return n + 4096;
}

// Conv2D('same') [+ Concatenate of a second source] or Conv2DTranspose(k4, s2, 'same') on the tensor cores.
// split != 0 selects the 3-term split-bf16 arithmetic.  The second source (d_x2, cin2) may be NULL / 0.
extern "C" int cic_conv2d_nhwc_tc(const float* d_x, const float* d_x2, const float* d_kernel, const float* d_bias,
                                  const float* d_scale, const float* d_shift, float* d_y, int batch, int h, int w, int cin,
                                  int cin2, int cout, int kh, int kw, int stride, int transpose, int act, int split,
                                  void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(d_x && d_kernel && d_y, "cic_conv2d_nhwc_tc: null pointer");
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "cic_conv2d_nhwc_tc: bad shape");
  CIC_REQUIRE(stride == 1 || stride == 2, "cic_conv2d_nhwc_tc: stride must be 1 or 2");
  CIC_REQUIRE(!transpose || (kh == 4 && kw == 4 && stride == 2), "cic_conv2d_nhwc_tc: transposed conv is 4x4 / stride 2 only");
  if (batch == 0) return CIC_OK;
  if (!d_x2) cin2 = 0;
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_conv2d_tc_workspace_bytes(batch, h, w, cin, cin2, cout, kh, kw, stride, transpose),
              "cic_conv2d_nhwc_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c;
  c.arena.base = (char*)d_workspace;
  c.arena.cap = workspace_bytes;
  c.st = st;
  const size_t px = (size_t)batch * h * w;
  const int cp = round_up(cout, 16), ct = cin + cin2;
  ActBuf a0 = alloc_act(c, px * cin, split != 0), a1;
  int rc;
  if ((rc = tc_split_f32(d_x, a0.hi, a0.lo, px * cin, st))) return rc;
  if (cin2) {
    a1 = alloc_act(c, px * cin2, split != 0);
    if ((rc = tc_split_f32(d_x2, a1.hi, a1.lo, px * cin2, st))) return rc;
  }
  const int taps = transpose ? 4 : kh * kw;
  const int K = taps * ct;
  const int rows = transpose ? 4 * cp : cp;
  bf16* whi = (bf16*)c.arena.alloc_bytes((size_t)rows * K * 2);
  bf16* wlo = split ? (bf16*)c.arena.alloc_bytes((size_t)rows * K * 2) : nullptr;
  if (transpose) {
    // (4,4,Cout,Cin) -> four [K = (ty,tx,ci)][Cout] phase matrices, then to bf16 [Cout_pad][K]
    std::vector<float> hk((size_t)16 * ct * cout), packed;
    CIC_CHECK_CUDA(cudaMemcpyAsync(hk.data(), d_kernel, hk.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    CIC_CHECK_CUDA(cudaStreamSynchronize(st));
    pack_deconv_phases(hk.data(), cout, ct, packed);
    float* d_ph = (float*)c.arena.alloc_bytes(packed.size() * sizeof(float));
    CIC_CHECK_CUDA(cudaMemcpyAsync(d_ph, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CIC_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < 4; ++p)
      if ((rc = tc_pack_weight(d_ph + (size_t)p * K * cout, K, cout, cp, cout, whi + (size_t)p * cp * K, wlo ? wlo + (size_t)p * cp * K : nullptr, st))) return rc;
  } else {
    if ((rc = tc_pack_weight(d_kernel, K, cout, cp, cout, whi, wlo, st))) return rc;
  }
  TcMat wm{};
  wm.hi = whi; wm.lo = wlo; wm.K = K; wm.rows = rows; wm.row_stride = K; wm.batches = 1; wm.batch_stride = (long long)rows * K;
  TcEpilogue e;
  e.bias = d_bias; e.scale = d_scale; e.shift = d_shift; e.act = act; e.out_mode = TC_OUT_F32; e.out_hi = d_y; e.out_ld = cout;
  TcAct s1v = view(a1, cin2 ? cin2 : 32);
  const int kind = transpose ? TC_DECONV_K4S2 : (stride == 2 ? TC_CONV_S2 : TC_CONV_S1);
  rc = conv_tc(c, "conv_tc", kind, view(a0, cin), cin2 ? &s1v : nullptr, batch, h, w, transpose ? 2 : kh, transpose ? 2 : kw,
               transpose ? 1 : stride, wm, cout, split != 0, e);
  if (rc == CIC_OK && c.arena.overflow) {
    set_error("cic_conv2d_nhwc_tc: internal workspace overflow");
    return CIC_ERR_WORKSPACE;
  }
  return rc;
}

extern "C" size_t cic_dense_tc_workspace_bytes(int batch, int in_dim, int out_dim) {
  const int np = round_up(out_dim, 16);
  return 2 * align256((size_t)batch * in_dim * 2) + 2 * align256((size_t)np * in_dim * 2) +
         align256((size_t)64 * batch * out_dim * 4) + 4096;
}

// Dense on the tensor cores (split-K when batch x out_dim gives too few tiles); split != 0: 3-term split-bf16.
extern "C" int cic_dense_tc(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int in_dim,
                            int out_dim, int act, int split, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch >= 0 && in_dim > 0 && out_dim > 0, "cic_dense_tc: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_x && d_kernel && d_y, "cic_dense_tc: null pointer");
  CIC_REQUIRE(in_dim % 32 == 0, "cic_dense_tc: in_dim must be a multiple of 32");
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_dense_tc_workspace_bytes(batch, in_dim, out_dim), "cic_dense_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c;
  c.arena.base = (char*)d_workspace;
  c.arena.cap = workspace_bytes;
  c.st = st;
  const int np = round_up(out_dim, 16);
  ActBuf a = alloc_act(c, (size_t)batch * in_dim, split != 0);
  bf16* whi = (bf16*)c.arena.alloc_bytes((size_t)np * in_dim * 2);
  bf16* wlo = split ? (bf16*)c.arena.alloc_bytes((size_t)np * in_dim * 2) : nullptr;
  int rc;
  if ((rc = tc_split_f32(d_x, a.hi, a.lo, (size_t)batch * in_dim, st))) return rc;
  if ((rc = tc_pack_weight(d_kernel, in_dim, out_dim, np, out_dim, whi, wlo, st))) return rc;
  int splits = dense_splits(batch, np, in_dim, split != 0);
  if (splits > 64) splits = 64;
  float* part = splits > 1 ? c.arena.f32((size_t)splits * batch * out_dim) : nullptr;
  TcMat wm{};
  wm.hi = whi; wm.lo = wlo; wm.K = in_dim; wm.rows = np; wm.row_stride = in_dim; wm.batches = 1; wm.batch_stride = (long long)np * in_dim;
  TcEpilogue e;
  e.act = act;
  if (splits > 1) { e.out_mode = TC_OUT_PARTIAL; e.out_hi = part; }
  else { e.bias = d_bias; e.out_mode = TC_OUT_F32; e.out_hi = d_y; e.out_ld = out_dim; }
  if ((rc = dense_tc(c, "dense_tc", view(a, in_dim), batch, wm, out_dim, split != 0, splits, e))) return rc;
  if (splits > 1) rc = tc_splitk_reduce(part, splits, batch, out_dim, d_bias, nullptr, nullptr, act, d_y, nullptr, nullptr, st);
  if (rc == CIC_OK && c.arena.overflow) {
    set_error("cic_dense_tc: internal workspace overflow");
    return CIC_ERR_WORKSPACE;
  }
  return rc;
}
