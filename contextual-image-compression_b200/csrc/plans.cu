// Plans: weight packing at creation, and the graph walkers that restate the reference models
//   build_autoencoder                (train_autoencoder.py:9-40)
//   build_encoder / SelfAttention    (GAN_functions.py:280-374)
//   build_generator                  (GAN_functions.py:236-278)
//   build_latent_saliency_model      (GAN_functions.py:210-234)
//   build_rate_distortion_optimizer  (GAN_functions.py:495-557)
//   build_adaptive_compression_model (GAN_functions.py:559-722)
// as sequences of this library's kernels on one stream.  CIC_PREC_FP32 walkers live here; the
// tcgen05 walkers are in plans_tc.cu.
#include "plan.cuh"

#include <cmath>
#include <cstring>

namespace cic {

WeightStore::~WeightStore() {
  for (auto& kv : t_) cudaFree(kv.second.p);
}

int WeightStore::upload(const std::string& name, const float* h, const std::vector<int64_t>& shape) {
  DevTensor t;
  t.shape = shape;
  const size_t n = t.numel();
  CIC_CHECK_CUDA(cudaMalloc(&t.p, (n ? n : 1) * sizeof(float)));
  if (n) CIC_CHECK_CUDA(cudaMemcpy(t.p, h, n * sizeof(float), cudaMemcpyHostToDevice));
  auto it = t_.find(name);
  if (it != t_.end()) {
    cudaFree(it->second.p);
    bytes_ -= it->second.numel() * sizeof(float);
  }
  t_[name] = t;
  bytes_ += n * sizeof(float);
  return CIC_OK;
}

const DevTensor* WeightStore::find(const std::string& name) const {
  auto it = t_.find(name);
  return it == t_.end() ? nullptr : &it->second;
}

Profiler::~Profiler() {
  for (auto& r : recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : pool) cudaEventDestroy(e);
}
cudaEvent_t Profiler::get_event() {
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void Profiler::reset() {
  for (auto& r : recs) { pool.push_back(r.e0); pool.push_back(r.e1); }
  recs.clear();
}
int Profiler::begin(const std::string& name, double flops, double bytes, cudaStream_t st) {
  Rec r{prefix + name, get_event(), get_event(), flops, bytes, 0};
  g_last_kernel_kind = KK_NONE;
  cudaEventRecord(r.e0, st);
  recs.push_back(r);
  return (int)recs.size() - 1;
}
void Profiler::end(int idx, cudaStream_t st) {
  if (idx >= 0 && idx < (int)recs.size()) {
    cudaEventRecord(recs[idx].e1, st);
    recs[idx].kind = g_last_kernel_kind;
  }
}
std::string Profiler::report() {
  std::string out;
  char line[256];
  for (auto& r : recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.e1);
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    snprintf(line, sizeof(line), "%s,%.6f,%.6e,%.6e,%d\n", r.name.c_str(), ms, r.flops, r.bytes, r.kind);
    out += line;
  }
  return out;
}

// ---- host-side tensor lookup during plan creation ---------------------------------------------
struct HostTensors {
  const cic_tensor* t;
  int n;
  std::string prefix;
  const cic_tensor* get(const std::string& name) const {
    const std::string full = prefix + name;
    for (int i = 0; i < n; ++i)
      if (t[i].name && full == t[i].name) return &t[i];
    return nullptr;
  }
};

static size_t numel(const cic_tensor* t) {
  size_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= (size_t)t->shape[i];
  return n;
}

#define NEED(var, hs, name)                                                             \
  const cic_tensor* var = (hs).get(name);                                               \
  if (!var) {                                                                           \
    set_error("plan: missing weight tensor '%s%s'", (hs).prefix.c_str(), name);         \
    return CIC_ERR_MISSING;                                                             \
  }

static int upload_raw(cic_plan* pl, const HostTensors& hs, const char* name) {
  NEED(t, hs, name);
  std::vector<int64_t> shape(t->shape, t->shape + t->ndim);
  return pl->w.upload(name, t->h_data, shape);
}

// folded inference BatchNorm (eps 1e-3): scale = gamma/sqrt(var+eps), shift = beta - mean*scale
static int upload_bn(cic_plan* pl, const HostTensors& hs, const std::string& bn, int repeat = 1) {
  NEED(g, hs, (bn + "/gamma").c_str());
  NEED(b, hs, (bn + "/beta").c_str());
  NEED(m, hs, (bn + "/moving_mean").c_str());
  NEED(v, hs, (bn + "/moving_variance").c_str());
  const size_t c = numel(g);
  std::vector<float> scale(c * repeat), shift(c * repeat);
  for (size_t i = 0; i < c; ++i) {
    const double s = (double)g->h_data[i] / std::sqrt((double)v->h_data[i] + 1e-3);
    const float sf = (float)s, tf = (float)((double)b->h_data[i] - (double)m->h_data[i] * s);
    for (int r = 0; r < repeat; ++r) {
      scale[(size_t)r * c + i] = sf;
      shift[(size_t)r * c + i] = tf;
    }
  }
  int rc = pl->w.upload(bn + "/scale", scale.data(), {(int64_t)(c * repeat)});
  if (rc) return rc;
  return pl->w.upload(bn + "/shift", shift.data(), {(int64_t)(c * repeat)});
}

// ---- plan builders ----------------------------------------------------------------------------
static int build_autoencoder(cic_plan* pl, const HostTensors& hs) {
  static const char* layers[] = {"conv1", "conv2", "conv3", "conv_x2", "conv5", "conv_x1", "conv_out"};
  for (const char* l : layers) {
    int rc = upload_raw(pl, hs, (std::string(l) + "/kernel").c_str());
    if (rc) return rc;
    rc = upload_raw(pl, hs, (std::string(l) + "/bias").c_str());
    if (rc) return rc;
  }
  return CIC_OK;
}

static int build_encoder(cic_plan* pl, const HostTensors& hs) {
  for (int i = 1; i <= 4; ++i) {
    const std::string c = "conv" + std::to_string(i);
    int rc = upload_raw(pl, hs, (c + "/kernel").c_str());
    if (rc) return rc;
    rc = upload_raw(pl, hs, (c + "/bias").c_str());
    if (rc) return rc;
    if (i > 1) {
      rc = upload_bn(pl, hs, "bn" + std::to_string(i));
      if (rc) return rc;
    }
  }
  if (pl->opts.add_attention) {
    // one 1x1 conv with the q, k, v kernels side by side: [C][C/8 + C/8 + C]
    NEED(q, hs, "attn/query/kernel");
    NEED(k, hs, "attn/key/kernel");
    NEED(v, hs, "attn/value/kernel");
    NEED(bq, hs, "attn/query/bias");
    NEED(bk, hs, "attn/key/bias");
    NEED(bv, hs, "attn/value/bias");
    NEED(gm, hs, "attn/gamma");
    const int C = (int)q->shape[2], dq = (int)q->shape[3], dv = (int)v->shape[3];
    const int N = 2 * dq + dv;
    std::vector<float> w((size_t)C * N), b(N);
    for (int c = 0; c < C; ++c) {
      for (int j = 0; j < dq; ++j) w[(size_t)c * N + j] = q->h_data[(size_t)c * dq + j];
      for (int j = 0; j < dq; ++j) w[(size_t)c * N + dq + j] = k->h_data[(size_t)c * dq + j];
      for (int j = 0; j < dv; ++j) w[(size_t)c * N + 2 * dq + j] = v->h_data[(size_t)c * dv + j];
    }
    for (int j = 0; j < dq; ++j) { b[j] = bq->h_data[j]; b[dq + j] = bk->h_data[j]; }
    for (int j = 0; j < dv; ++j) b[2 * dq + j] = bv->h_data[j];
    int rc = pl->w.upload("attn/qkv/kernel", w.data(), {C, N});
    if (rc) return rc;
    rc = pl->w.upload("attn/qkv/bias", b.data(), {N});
    if (rc) return rc;
    rc = pl->w.upload("attn/gamma", gm->h_data, {1});
    if (rc) return rc;
  }
  int rc = upload_raw(pl, hs, "dense/kernel");
  if (rc) return rc;
  return upload_raw(pl, hs, "dense/bias");
}

static int build_generator(cic_plan* pl, const HostTensors& hs) {
  int rc = upload_raw(pl, hs, "dense/kernel");
  if (rc) return rc;
  rc = upload_raw(pl, hs, "dense/bias");
  if (rc) return rc;
  const int feat_pix = (pl->opts.img_h / 16) * (pl->opts.img_w / 16);
  rc = upload_bn(pl, hs, "bn0", feat_pix);  // per-channel BN after the NHWC Reshape, expanded over pixels
  if (rc) return rc;
  for (int i = 1; i <= 4; ++i) {
    const std::string d = "deconv" + std::to_string(i);
    NEED(k, hs, (d + "/kernel").c_str());
    const int cout = (int)k->shape[2], cin = (int)k->shape[3];
    std::vector<float> packed;
    pack_deconv_phases(k->h_data, cout, cin, packed);
    rc = pl->w.upload(d + "/phases", packed.data(), {4, 4 * cin, cout});
    if (rc) return rc;
    rc = upload_raw(pl, hs, (d + "/bias").c_str());
    if (rc) return rc;
    rc = upload_bn(pl, hs, "bn" + std::to_string(i));
    if (rc) return rc;
  }
  rc = upload_raw(pl, hs, "conv_out/kernel");
  if (rc) return rc;
  return upload_raw(pl, hs, "conv_out/bias");
}

static int build_saliency(cic_plan* pl, const HostTensors& hs) {
  for (int i = 1; i <= 3; ++i) {
    const std::string d = "dense" + std::to_string(i);
    int rc = upload_raw(pl, hs, (d + "/kernel").c_str());
    if (rc) return rc;
    rc = upload_raw(pl, hs, (d + "/bias").c_str());
    if (rc) return rc;
  }
  return CIC_OK;
}

static int build_rd(cic_plan* pl, const HostTensors& hs) {
  static const char* names[] = {"conv1", "conv2", "dense1", "dense2"};
  for (const char* n : names) {
    int rc = upload_raw(pl, hs, (std::string(n) + "/kernel").c_str());
    if (rc) return rc;
    rc = upload_raw(pl, hs, (std::string(n) + "/bias").c_str());
    if (rc) return rc;
  }
  return CIC_OK;
}


static int build_any(cic_plan* pl, const cic_tensor* tensors, int n, const std::string& prefix) {
  HostTensors hs{tensors, n, prefix};
  int rc = CIC_OK;
  switch (pl->kind) {
    case CIC_PLAN_AUTOENCODER: rc = build_autoencoder(pl, hs); break;
    case CIC_PLAN_ENCODER: rc = build_encoder(pl, hs); break;
    case CIC_PLAN_GENERATOR: rc = build_generator(pl, hs); break;
    case CIC_PLAN_SALIENCY: rc = build_saliency(pl, hs); break;
    case CIC_PLAN_RD: rc = build_rd(pl, hs); break;
    default: set_error("plan: unknown kind %d", pl->kind); return CIC_ERR_INVALID;
  }
  if (rc) return rc;
  if (pl->opts.precision == CIC_PREC_TC) return build_plan_tc(pl, tensors, n, prefix);
  return CIC_OK;
}

// ---- small kernels used only by the walkers ---------------------------------------------------
__global__ void expand_bpp_kernel(const float* __restrict__ bpp, float* __restrict__ bpp_t, float* __restrict__ qs_t,
                                  int n_tiles, int tiles_per_img) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles) return;
  const float b = bpp[i / tiles_per_img];
  bpp_t[i] = b;
  qs_t[i] = rate_qs(rate_t(b));
}

// column 64 of the RD feature row = clip(bpp/5, 0, 1) (GAN_functions.py:518)
__global__ void rd_set_t_kernel(const float* __restrict__ bpp, float* __restrict__ feat, int batch, int ld, int col) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) feat[(size_t)i * ld + col] = rate_t(bpp[i]);
}

// sigmoid(p0 + 1 - 2t), sigmoid(p1 + 1 - 2t), sigmoid(p2 + 1 - 1.5t) (GAN_functions.py:529-541)
__global__ void rd_finalize_kernel(const float* __restrict__ base, const float* __restrict__ bpp, float* __restrict__ out,
                                   int batch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const float t = rate_t(bpp[i]);
  const float k[3] = {2.0f, 2.0f, 1.5f};
#pragma unroll
  for (int j = 0; j < 3; ++j)
    out[i * 3 + j] = sigmoidf_(__fsub_rn(__fadd_rn(base[i * 3 + j], 1.0f), __fmul_rn(k[j], t)));
}

// ---- generic layer helpers --------------------------------------------------------------------
struct Srcs {
  ConvSrc s[2];
  int n;
};
static Srcs one(const float* p, int C, int ld = 0, int up = 0) {
  Srcs r{};
  r.s[0] = ConvSrc{p, C, ld ? ld : C, up};
  r.n = 1;
  return r;
}
static Srcs two(const float* p0, int C0, int up0, const float* p1, int C1) {
  Srcs r{};
  r.s[0] = ConvSrc{p0, C0, C0, up0};
  r.s[1] = ConvSrc{p1, C1, C1, 0};
  r.n = 2;
  return r;
}

static int conv_same(Ctx& c, const char* name, const Srcs& in, int batch, int H, int W, const float* kernel, int kh, int kw,
                     int stride, int cout, const float* bias, const float* scale, const float* shift, int act, float* out) {
  if (c.dry) return CIC_OK;
  const int cin_ = in.s[0].C + (in.n > 1 ? in.s[1].C : 0);
  const double mrows = (double)batch * same_out(H, stride) * same_out(W, stride);
  Scope sc(c, name, 2.0 * mrows * cout * kh * kw * cin_, 4.0 * ((double)batch * H * W * cin_ + mrows * cout));
  IGemmParams p{};
  p.src[0] = in.s[0];
  p.src[1] = in.s[1];
  p.nsrc = in.n;
  p.Cin = in.s[0].C + (in.n > 1 ? in.s[1].C : 0);
  p.batch = batch; p.H = H; p.W = W;
  p.Ho = same_out(H, stride); p.Wo = same_out(W, stride);
  p.kh = kh; p.kw = kw; p.stride = stride;
  p.pad_t = same_pad_before(H, kh, stride); p.pad_l = same_pad_before(W, kw, stride);
  p.Bmat = kernel; p.N = cout; p.ldb = cout;
  p.bias = bias; p.scale = scale; p.shift = shift; p.act = act; p.alpha = 1.f;
  p.out = out; p.out_ld = cout; p.out_H = p.Ho; p.out_W = p.Wo; p.out_ys = p.out_xs = 1;
  p.splits = 1;
  return launch_igemm(p, c.st);
}

static int deconv_k4s2(Ctx& c, const char* name, const Srcs& in, int batch, int H, int W, const float* phases, int cout,
                       const float* bias, const float* scale, const float* shift, int act, float* out) {
  if (c.dry) return CIC_OK;
  const int cin = in.s[0].C + (in.n > 1 ? in.s[1].C : 0);
  Scope sc(c, name, 2.0 * (double)batch * H * W * 4 * 4 * cin * cout, 4.0 * (double)batch * H * W * (cin + 4.0 * cout));
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    IGemmParams p{};
    p.src[0] = in.s[0];
    p.src[1] = in.s[1];
    p.nsrc = in.n;
    p.Cin = cin;
    p.batch = batch; p.H = H; p.W = W; p.Ho = H; p.Wo = W;
    p.kh = 2; p.kw = 2; p.stride = 1;
    p.pad_t = py == 0 ? 1 : 0; p.pad_l = px == 0 ? 1 : 0;
    p.Bmat = phases + (size_t)ph * 4 * cin * cout; p.N = cout; p.ldb = cout;
    p.bias = bias; p.scale = scale; p.shift = shift; p.act = act; p.alpha = 1.f;
    p.out = out; p.out_ld = cout; p.out_H = 2 * H; p.out_W = 2 * W;
    p.out_ys = 2; p.out_xs = 2; p.out_y0 = py; p.out_x0 = px;
    p.splits = 1;
    int rc = launch_igemm(p, c.st);
    if (rc) return rc;
  }
  return CIC_OK;
}

static int dense(Ctx& c, const char* name, const float* x, const float* kernel, const float* bias, const float* scale,
                 const float* shift, float* y, int batch, int in_dim, int out_dim, int act) {
  Scope sc(c, name, 2.0 * batch * (double)in_dim * out_dim, 4.0 * ((double)in_dim * out_dim + (double)batch * (in_dim + out_dim)));
  const size_t wsb = cic_dense_workspace_bytes(batch, in_dim, out_dim);
  const size_t mk = c.arena.mark();
  float* ws = wsb ? (float*)c.arena.alloc_bytes(wsb) : nullptr;
  int rc = CIC_OK;
  if (!c.dry) rc = run_dense(x, kernel, bias, scale, shift, y, batch, in_dim, out_dim, act, ws, wsb / sizeof(float), c.st);
  c.arena.release(mk);  // stream order makes reuse after the reduce safe
  return rc;
}

// SelfAttention.call: qkv 1x1 conv -> S = q k^T -> softmax rows -> gamma * (P v) + x
static int attention_f32(Ctx& c, const char* name, const float* x, const float* wqkv, const float* bqkv, const float* d_gamma, float gamma_host,
                         bool gamma_on_device, float* y, int batch, int tokens, int C) {
  const int dq = C / 8, N = 2 * dq + C;
  Scope sc(c, name, 2.0 * batch * tokens * ((double)C * N + (double)tokens * dq + (double)tokens * C), 8.0 * batch * tokens * C);
  const int chunk_max = 32;  // images per pass: bounds the tokens x tokens score matrix workspace
  const size_t mk = c.arena.mark();
  const int chunk = batch < chunk_max ? batch : chunk_max;
  float* qkv = c.arena.f32((size_t)chunk * tokens * N);
  float* S = c.arena.f32((size_t)chunk * tokens * tokens);
  int rc = CIC_OK;
  float g = gamma_host;
  if (!c.dry && gamma_on_device) {
    CIC_CHECK_CUDA(cudaMemcpyAsync(&g, d_gamma, sizeof(float), cudaMemcpyDeviceToHost, c.st));
    CIC_CHECK_CUDA(cudaStreamSynchronize(c.st));
  }
  for (int b0 = 0; b0 < batch && !c.dry; b0 += chunk) {
    const int nb = batch - b0 < chunk ? batch - b0 : chunk;
    const float* xb = x + (size_t)b0 * tokens * C;
    // q, k, v projections as one GEMM: (nb*tokens, C) x (C, N)
    rc = conv_same(c, "qkv", one(xb, C), nb * tokens, 1, 1, wqkv, 1, 1, 1, N, bqkv, nullptr, nullptr, CIC_ACT_NONE, qkv);
    if (rc) break;
    // S[b] = q[b] k[b]^T  (no 1/sqrt(d) scaling, GAN_functions.py:358)
    IGemmParams p{};
    p.src[0] = ConvSrc{qkv, dq, N, 0};
    p.nsrc = 1; p.Cin = dq; p.batch = nb; p.H = tokens; p.W = 1; p.Ho = tokens; p.Wo = 1;
    p.kh = p.kw = 1; p.stride = 1;
    p.Bmat = qkv + dq; p.N = tokens; p.ldb = N; p.b_trans = 1; p.b_batch_stride = (long long)tokens * N;
    p.alpha = 1.f; p.act = CIC_ACT_NONE;
    p.out = S; p.out_ld = tokens; p.out_H = tokens; p.out_W = 1; p.out_ys = p.out_xs = 1; p.splits = 1;
    rc = launch_igemm(p, c.st);
    if (rc) break;
    rc = launch_softmax_rows(S, (long long)nb * tokens, tokens, c.st);
    if (rc) break;
    // y = gamma * (P v) + x
    IGemmParams q{};
    q.src[0] = ConvSrc{S, tokens, tokens, 0};
    q.nsrc = 1; q.Cin = tokens; q.batch = nb; q.H = tokens; q.W = 1; q.Ho = tokens; q.Wo = 1;
    q.kh = q.kw = 1; q.stride = 1;
    q.Bmat = qkv + 2 * dq; q.N = C; q.ldb = N; q.b_trans = 0; q.b_batch_stride = (long long)tokens * N;
    q.alpha = g; q.act = CIC_ACT_NONE; q.residual = xb;
    q.out = y + (size_t)b0 * tokens * C; q.out_ld = C; q.out_H = tokens; q.out_W = 1; q.out_ys = q.out_xs = 1; q.splits = 1;
    rc = launch_igemm(q, c.st);
    if (rc) break;
  }
  c.arena.release(mk);
  return rc;
}

// ---- walkers (CIC_PREC_FP32) ------------------------------------------------------------------
int autoencoder_forward_f32(cic_plan* pl, Ctx& c, const float* x, float* y, uint8_t* y_u8, int B, int H, int W) {
  const WeightStore& w = pl->w;
  const int C = pl->opts.img_c;
  const size_t px = (size_t)B * H * W;
  float* x1 = c.arena.f32(px * 32);
  float* x1p = c.arena.f32(px / 4 * 32);
  float* x2 = c.arena.f32(px / 4 * 64);
  float* enc = c.arena.f32(px / 16 * 64);
  float* y3 = c.arena.f32(px / 16 * 64);
  float* x2r = c.arena.f32(px / 4 * 64);
  float* y5 = c.arena.f32(px / 4 * 32);
  float* x1r = c.arena.f32(px * 32);
  if (c.dry) return CIC_OK;
  int rc;
#define K(n) w.ptr(n "/kernel"), 3, 3, 1
#define Bv(n) w.ptr(n "/bias"), nullptr, nullptr
  if ((rc = conv_same(c, "conv1", one(x, C), B, H, W, K("conv1"), 32, Bv("conv1"), CIC_ACT_RELU, x1))) return rc;          // :14
  { Scope sc(c, "pool1", 0, 4.0 * px * 32 * 1.25);
    if ((rc = launch_maxpool2x2(x1, x1p, B, H, W, 32, c.st))) return rc; }                                              // :15
  if ((rc = conv_same(c, "conv2", one(x1p, 32), B, H / 2, W / 2, K("conv2"), 64, Bv("conv2"), CIC_ACT_RELU, x2))) return rc;  // :17
  { Scope sc(c, "pool2", 0, 4.0 * px / 4 * 64 * 1.25);
    if ((rc = launch_maxpool2x2(x2, enc, B, H / 2, W / 2, 64, c.st))) return rc; }                                      // :18
  if ((rc = conv_same(c, "conv3", one(enc, 64), B, H / 4, W / 4, K("conv3"), 64, Bv("conv3"), CIC_ACT_RELU, y3))) return rc;  // :21
  if ((rc = conv_same(c, "conv_x2", one(x2, 64), B, H / 2, W / 2, K("conv_x2"), 64, Bv("conv_x2"), CIC_ACT_RELU, x2r))) return rc;  // :25
  // :22 UpSampling2D + :26 concatenate folded into the gather of conv5 (:28)
  if ((rc = conv_same(c, "conv5", two(y3, 64, 1, x2r, 64), B, H / 2, W / 2, K("conv5"), 32, Bv("conv5"), CIC_ACT_RELU, y5))) return rc;
  if ((rc = conv_same(c, "conv_x1", one(x1, 32), B, H, W, K("conv_x1"), 32, Bv("conv_x1"), CIC_ACT_RELU, x1r))) return rc;     // :32
#undef K
#undef Bv
  // :29 UpSampling2D + :33 concatenate + :35 Conv2D(3, sigmoid)
  SmallNParams s{};
  s.src[0] = ConvSrc{y5, 32, 32, 1};
  s.src[1] = ConvSrc{x1r, 32, 32, 0};
  s.nsrc = 2; s.Cin = 64; s.batch = B; s.H = H; s.W = W; s.kh = 3; s.kw = 3; s.pad_t = 1; s.pad_l = 1;
  s.Wmat = w.ptr("conv_out/kernel"); s.bias = w.ptr("conv_out/bias"); s.N = C; s.act = CIC_ACT_SIGMOID; s.out = y;
  { Scope sc(c, "conv_out", 2.0 * px * 9 * 64 * C, 4.0 * px * (64 + C));
    if ((rc = launch_conv_small_n(s, c.st))) return rc; }
  Scope sc(c, "cast_u8", 0, 5.0 * px * C);
  if (y_u8) rc = cic_f32_to_u8_trunc(y, y_u8, px * C, 255.0f, c.st);                                                 // test_autoencoder.py:88
  return rc;
}

int encoder_forward_f32(cic_plan* pl, Ctx& c, const float* img, float* latent, float* x1, float* x2, float* x3, int B) {
  const WeightStore& w = pl->w;
  const int H = pl->opts.img_h, W = pl->opts.img_w, C = pl->opts.img_c, L = pl->opts.latent_dim;
  const size_t px = (size_t)B * H * W;
  if (!x1) x1 = c.arena.f32(px / 4 * 64);
  if (!x2) x2 = c.arena.f32(px / 16 * 128);
  if (!x3) x3 = c.arena.f32(px / 64 * 256);
  float* x3a = pl->opts.add_attention ? c.arena.f32(px / 64 * 256) : x3;
  float* x4 = c.arena.f32(px / 256 * 512);
  int rc;
  if ((rc = conv_same(c, "conv1", one(img, C), B, H, W, w.ptr("conv1/kernel"), 4, 4, 2, 64, w.ptr("conv1/bias"), nullptr, nullptr,
                      CIC_ACT_LRELU02, x1))) return rc;                                                              // :300-301
  if ((rc = conv_same(c, "conv2", one(x1, 64), B, H / 2, W / 2, w.ptr("conv2/kernel"), 4, 4, 2, 128, w.ptr("conv2/bias"),
                      w.ptr("bn2/scale"), w.ptr("bn2/shift"), CIC_ACT_LRELU02, x2))) return rc;                      // :304-306
  if ((rc = conv_same(c, "conv3", one(x2, 128), B, H / 4, W / 4, w.ptr("conv3/kernel"), 4, 4, 2, 256, w.ptr("conv3/bias"),
                      w.ptr("bn3/scale"), w.ptr("bn3/shift"), CIC_ACT_LRELU02, x3))) return rc;                      // :309-311
  if (pl->opts.add_attention) {                                                                                      // :315-318
    if ((rc = attention_f32(c, "attention", x3, w.ptr("attn/qkv/kernel"), w.ptr("attn/qkv/bias"), w.ptr("attn/gamma"), 0.f, true, x3a, B,
                            (H / 8) * (W / 8), 256))) return rc;
  }
  if ((rc = conv_same(c, "conv4", one(x3a, 256), B, H / 8, W / 8, w.ptr("conv4/kernel"), 4, 4, 2, 512, w.ptr("conv4/bias"),
                      w.ptr("bn4/scale"), w.ptr("bn4/shift"), CIC_ACT_LRELU02, x4))) return rc;                      // :320-322
  const int feat = (H / 16) * (W / 16) * 512;
  return dense(c, "dense", x4, w.ptr("dense/kernel"), w.ptr("dense/bias"), nullptr, nullptr, latent, B, feat, L, CIC_ACT_NONE);  // :325-326
}

int generator_forward_f32(cic_plan* pl, Ctx& c, const float* latent, const float* s1, const float* s2, const float* s3,
                          float* out, int B) {
  const WeightStore& w = pl->w;
  const int H = pl->opts.img_h, W = pl->opts.img_w, C = pl->opts.img_c, L = pl->opts.latent_dim;
  const int h16 = H / 16, w16 = W / 16, feat = h16 * w16 * 512;
  const size_t px = (size_t)B * H * W;
  float* g0 = c.arena.f32((size_t)B * feat);
  float* g1 = c.arena.f32(px / 64 * 256);
  float* g2 = c.arena.f32(px / 16 * 128);
  float* g3 = c.arena.f32(px / 4 * 64);
  float* g4 = c.arena.f32(px * 32);
  int rc;
  // :247-250 Dense -> Reshape(h16,w16,512) NHWC -> BN -> LeakyReLU
  if ((rc = dense(c, "dense", latent, w.ptr("dense/kernel"), w.ptr("dense/bias"), w.ptr("bn0/scale"), w.ptr("bn0/shift"), g0, B, L, feat,
                  CIC_ACT_LRELU02))) return rc;
#define DC(i, srcs, hh, ww, co, dst)                                                                             \
  if ((rc = deconv_k4s2(c, "deconv" #i, srcs, B, hh, ww, w.ptr("deconv" #i "/phases"), co, w.ptr("deconv" #i "/bias"),       \
                        w.ptr("bn" #i "/scale"), w.ptr("bn" #i "/shift"), CIC_ACT_LRELU02, dst))) return rc
  DC(1, one(g0, 512), h16, w16, 256, g1);                            // :253-255
  DC(2, two(g1, 256, 0, s3, 256), 2 * h16, 2 * w16, 128, g2);        // :256 concat skip3, :258-260
  DC(3, two(g2, 128, 0, s2, 128), 4 * h16, 4 * w16, 64, g3);         // :261, :263-265
  DC(4, two(g3, 64, 0, s1, 64), 8 * h16, 8 * w16, 32, g4);           // :266, :268-270
#undef DC
  if (c.dry) return CIC_OK;
  SmallNParams s{};                                                  // :273 Conv2D(3, k4, 'same', tanh): pad 1 before / 2 after
  s.src[0] = ConvSrc{g4, 32, 32, 0};
  s.nsrc = 1; s.Cin = 32; s.batch = B; s.H = H; s.W = W; s.kh = 4; s.kw = 4;
  s.pad_t = same_pad_before(H, 4, 1); s.pad_l = same_pad_before(W, 4, 1);
  s.Wmat = w.ptr("conv_out/kernel"); s.bias = w.ptr("conv_out/bias"); s.N = C; s.act = CIC_ACT_TANH; s.out = out;
  Scope sc(c, "conv_out", 2.0 * px * 16 * 32 * C, 4.0 * px * (32 + C));
  return launch_conv_small_n(s, c.st);
}

int saliency_forward_f32(cic_plan* pl, Ctx& c, const float* latent, float* score, int B) {
  const WeightStore& w = pl->w;
  const int L = pl->opts.latent_dim;
  float* h1 = c.arena.f32((size_t)B * 512);
  float* h2 = c.arena.f32((size_t)B * 256);
  int rc;
  if ((rc = dense(c, "dense1", latent, w.ptr("dense1/kernel"), w.ptr("dense1/bias"), nullptr, nullptr, h1, B, L, 512, CIC_ACT_RELU))) return rc;
  if ((rc = dense(c, "dense2", h1, w.ptr("dense2/kernel"), w.ptr("dense2/bias"), nullptr, nullptr, h2, B, 512, 256, CIC_ACT_RELU))) return rc;
  return dense(c, "dense3", h2, w.ptr("dense3/kernel"), w.ptr("dense3/bias"), nullptr, nullptr, score, B, 256, 1, CIC_ACT_SIGMOID);
}

int rd_forward_f32(cic_plan* pl, Ctx& c, const float* mask, const float* bpp, float* rd_params, int B) {
  const WeightStore& w = pl->w;
  const int H = pl->opts.img_h, W = pl->opts.img_w;
  const int h2 = same_out(H, 2), w2 = same_out(W, 2), h4 = same_out(h2, 2), w4 = same_out(w2, 2);
  float* r1 = c.arena.f32((size_t)B * h2 * w2 * 32);
  float* r2 = c.arena.f32((size_t)B * h4 * w4 * 64);
  float* feat = c.arena.f32((size_t)B * 65);
  float* d1 = c.arena.f32((size_t)B * 128);
  float* base = c.arena.f32((size_t)B * 3);
  int rc;
  if ((rc = conv_same(c, "conv1", one(mask, 1), B, H, W, w.ptr("conv1/kernel"), 3, 3, 2, 32, w.ptr("conv1/bias"), nullptr, nullptr,
                      CIC_ACT_LRELU02, r1))) return rc;                                                   // :511-512
  if ((rc = conv_same(c, "conv2", one(r1, 32), B, h2, w2, w.ptr("conv2/kernel"), 3, 3, 2, 64, w.ptr("conv2/bias"), nullptr, nullptr,
                      CIC_ACT_LRELU02, r2))) return rc;                                                   // :513-514
  if (!c.dry && (rc = launch_global_avg_pool(r2, feat, B, h4 * w4, 64, 65, c.st))) return rc;             // :515
  return launch_rd_tail(pl, c, bpp, feat, d1, base, rd_params, B);
}

// concat clip(bpp/5) (:518) -> Dense128 LReLU (:521-522) -> Dense3 (:525) -> three biased sigmoids (:529-541)
int launch_rd_tail(cic_plan* pl, Ctx& c, const float* bpp, float* feat, float* d1, float* base, float* rd_params, int B) {
  const WeightStore& w = pl->w;
  int rc;
  static const int fused_env = CIC_KNOB("CIC_MLP_FUSED", 1);
  if (fused_env) {  // concat + Dense128 + Dense3 + sigmoids in one kernel (mlp_fused.cu)
    if (c.dry) return CIC_OK;
    Scope sc(c, "tail", 2.0 * B * (65.0 * 128 + 128 * 3), 4.0 * (65 * 128 + 128 * 3 + (double)B * 68));
    return launch_rd_tail_fused(feat, 65, bpp, w.ptr("dense1/kernel"), w.ptr("dense1/bias"), w.ptr("dense2/kernel"), w.ptr("dense2/bias"), rd_params, B,
                                c.st);
  }
  if (!c.dry) {
    rd_set_t_kernel<<<(B + 127) / 128, 128, 0, c.st>>>(bpp, feat, B, 65, 64);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("rd_set_t_kernel");
  }
  if ((rc = dense(c, "dense1", feat, w.ptr("dense1/kernel"), w.ptr("dense1/bias"), nullptr, nullptr, d1, B, 65, 128, CIC_ACT_LRELU02))) return rc;
  if ((rc = dense(c, "dense2", d1, w.ptr("dense2/kernel"), w.ptr("dense2/bias"), nullptr, nullptr, base, B, 128, 3, CIC_ACT_NONE))) return rc;
  if (!c.dry) {
    rd_finalize_kernel<<<(B + 127) / 128, 128, 0, c.st>>>(base, bpp, rd_params, B);
    CIC_COUNT_LAUNCH();
    CIC_CHECK_LAUNCH("rd_finalize_kernel");
  }
  return CIC_OK;
}

int launch_expand_bpp(const float* bpp, float* bpp_t, float* qs_t, int n_tiles, int tiles_per_img, cudaStream_t st) {
  expand_bpp_kernel<<<(n_tiles + 127) / 128, 128, 0, st>>>(bpp, bpp_t, qs_t, n_tiles, tiles_per_img);
  CIC_COUNT_LAUNCH();
  CIC_CHECK_LAUNCH("expand_bpp_kernel");
  return CIC_OK;
}

int adaptive_forward_f32(cic_plan* pl, Ctx& c, const cic_adaptive_io* io, int n_img, int img_h, int img_w) {
  const int T = pl->opts.img_h, base = pl->opts.latent_dim;
  const int tpi = ((img_h + T - 1) / T) * ((img_w + T - 1) / T);
  const int nt = n_img * tpi;
  const bool tiled = tpi > 1 || img_h != T || img_w != T;  // ragged sizes: edge-replicated into whole tiles, cropped on the way out
  const size_t tpx = (size_t)nt * T * T;
  int rc;
  // 0. image / mask tiles
  const float* img_t = io->d_img;
  const float* mask_t = io->d_mask;
  if (tiled) {
    float* it = c.arena.f32(tpx * 3);
    float* mt = c.arena.f32(tpx);
    if (!c.dry) {
      if ((rc = launch_tile_gather(io->d_img, it, n_img, img_h, img_w, 3, T, c.st))) return rc;
      if ((rc = launch_tile_gather(io->d_mask, mt, n_img, img_h, img_w, 1, T, c.st))) return rc;
    }
    img_t = it;
    mask_t = mt;
  }
  float* bpp_t = c.arena.f32(nt);
  float* qs_t = c.arena.f32(nt);
  if (!c.dry && (rc = launch_expand_bpp(io->d_bpp, bpp_t, qs_t, nt, tpi, c.st))) return rc;              // :631-649
  // 1-2. encoders (:604-617)
  float* hq_lat = io->d_hq_latent ? io->d_hq_latent : c.arena.f32((size_t)nt * 2 * base);
  float* lq_lat = io->d_lq_latent ? io->d_lq_latent : c.arena.f32((size_t)nt * base);
  float* hx1 = c.arena.f32(tpx / 4 * 64);
  float* hx2 = c.arena.f32(tpx / 16 * 128);
  float* hx3 = c.arena.f32(tpx / 64 * 256);
  float* lx1 = c.arena.f32(tpx / 4 * 64);
  float* lx2 = c.arena.f32(tpx / 16 * 128);
  float* lx3 = c.arena.f32(tpx / 64 * 256);
  size_t mk = c.arena.mark();
  if (c.prof) c.prof->prefix = "hq_enc/";
  if ((rc = encoder_forward_f32(pl->hq_enc.get(), c, img_t, hq_lat, hx1, hx2, hx3, nt))) return rc;
  c.arena.release(mk);
  if (c.prof) c.prof->prefix = "lq_enc/";
  if ((rc = encoder_forward_f32(pl->lq_enc.get(), c, img_t, lq_lat, lx1, lx2, lx3, nt))) return rc;
  c.arena.release(mk);
  // 3. latent saliency (:619-620)
  float* sal_hq = c.arena.f32(nt);
  float* sal_lq = c.arena.f32(nt);
  mk = c.arena.mark();
  if (c.prof) c.prof->prefix = "sal_hq/";
  if ((rc = saliency_forward_f32(pl->sal_hq.get(), c, hq_lat, sal_hq, nt))) return rc;
  c.arena.release(mk);
  if (c.prof) c.prof->prefix = "sal_lq/";
  if ((rc = saliency_forward_f32(pl->sal_lq.get(), c, lq_lat, sal_lq, nt))) return rc;
  c.arena.release(mk);
  // 4. rate-distortion parameters (:624) - an output only; they do not drive the quantiser or blend
  if (io->d_rd_params || c.dry) {
    if (c.prof) c.prof->prefix = "rd/";
    if ((rc = rd_forward_f32(pl->rd.get(), c, mask_t, bpp_t, io->d_rd_params, nt))) return rc;
    c.arena.release(mk);
  }
  // 5. quantise (:661-666)
  float* hq_q = io->d_hq_latent_q ? io->d_hq_latent_q : c.arena.f32((size_t)nt * 2 * base);
  float* lq_q = io->d_lq_latent_q ? io->d_lq_latent_q : c.arena.f32((size_t)nt * base);
  if (c.prof) c.prof->prefix = "";
  if (!c.dry) {
    Scope sc(c, "quantize", 0, 12.0 * nt * 3 * base);
    if ((rc = cic_quantize_latent(hq_lat, sal_hq, qs_t, hq_q, io->d_hq_symbols, nullptr, io->d_hq_scale, nt, 2 * base, c.st))) return rc;
    if ((rc = cic_quantize_latent(lq_lat, sal_lq, qs_t, lq_q, io->d_lq_symbols, nullptr, io->d_lq_scale, nt, base, c.st))) return rc;
  }
  // 6. generators (:669-670)
  float* hq_out_t = (!tiled && io->d_hq_out) ? io->d_hq_out : c.arena.f32(tpx * 3);
  float* lq_out_t = (!tiled && io->d_lq_out) ? io->d_lq_out : c.arena.f32(tpx * 3);
  mk = c.arena.mark();
  if (c.prof) c.prof->prefix = "hq_gen/";
  if ((rc = generator_forward_f32(pl->hq_gen.get(), c, hq_q, hx1, hx2, hx3, hq_out_t, nt))) return rc;
  c.arena.release(mk);
  if (c.prof) c.prof->prefix = "lq_gen/";
  if ((rc = generator_forward_f32(pl->lq_gen.get(), c, lq_q, lx1, lx2, lx3, lq_out_t, nt))) return rc;
  c.arena.release(mk);
  // 7. dynamic threshold + blend on whole images (:651-657, :682-684)
  const float* hq_img = hq_out_t;
  const float* lq_img = lq_out_t;
  if (tiled) {
    float* hi = io->d_hq_out ? io->d_hq_out : c.arena.f32(tpx * 3);
    float* li = io->d_lq_out ? io->d_lq_out : c.arena.f32(tpx * 3);
    if (!c.dry) {
      if ((rc = launch_tile_scatter(hq_out_t, hi, n_img, img_h, img_w, 3, T, c.st))) return rc;
      if ((rc = launch_tile_scatter(lq_out_t, li, n_img, img_h, img_w, 3, T, c.st))) return rc;
    }
    hq_img = hi;
    lq_img = li;
  }
  if (c.prof) c.prof->prefix = "";
  if (!c.dry) {
    Scope sc(c, "roi_blend", 0, 44.0 * n_img * img_h * img_w);
    float* blended = io->d_blended;
    if ((rc = cic_roi_mask_blend(blended ? hq_img : nullptr, blended ? lq_img : nullptr, io->d_mask, io->d_bpp, blended,
                                 io->d_dt, io->d_hq_ratio_sum, n_img, img_h * img_w, 3, c.st))) return rc;
  }
  return CIC_OK;
}

}  // namespace cic

using namespace cic;

// ---- C ABI ------------------------------------------------------------------------------------
static cic_plan* make_sub(int kind, const cic_plan_opts& base, int latent, int attn, const cic_tensor* tensors, int n,
                          const std::string& prefix, int* rc) {
  cic_plan* p = new cic_plan();
  p->kind = kind;
  p->opts = base;
  p->opts.latent_dim = latent;
  p->opts.add_attention = attn;
  *rc = build_any(p, tensors, n, prefix);
  if (*rc) {
    delete p;
    return nullptr;
  }
  return p;
}

extern "C" cic_plan* cic_plan_create(int kind, const cic_tensor* tensors, int n_tensors, const cic_plan_opts* opts) {
  if (!opts || (!tensors && n_tensors > 0)) {
    set_error("cic_plan_create: null argument");
    return nullptr;
  }
  if (opts->precision != CIC_PREC_FP32 && opts->precision != CIC_PREC_TC) {
    set_error("cic_plan_create: unknown precision %d", opts->precision);
    return nullptr;
  }
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    cudaGetLastError();
    set_error("cic_plan_create: no CUDA device (this library has no CPU fallback)");
    return nullptr;
  }
  if (kind != CIC_PLAN_AUTOENCODER && kind != CIC_PLAN_SALIENCY) {
    if (opts->img_h <= 0 || opts->img_w <= 0 || opts->img_h % 16 || opts->img_w % 16) {
      set_error("cic_plan_create: model input size must be a positive multiple of 16, got %dx%d", opts->img_h, opts->img_w);
      return nullptr;
    }
  }
  cic_plan* pl = new cic_plan();
  pl->kind = kind;
  pl->opts = *opts;
  int rc = CIC_OK;
  if (kind == CIC_PLAN_ADAPTIVE) {
    const int base = opts->latent_dim;
    pl->hq_enc.reset(make_sub(CIC_PLAN_ENCODER, *opts, 2 * base, 1, tensors, n_tensors, "hq_encoder/", &rc));
    if (!rc) pl->lq_enc.reset(make_sub(CIC_PLAN_ENCODER, *opts, base, 0, tensors, n_tensors, "lq_encoder/", &rc));
    if (!rc) pl->hq_gen.reset(make_sub(CIC_PLAN_GENERATOR, *opts, 2 * base, 0, tensors, n_tensors, "hq_generator/", &rc));
    if (!rc) pl->lq_gen.reset(make_sub(CIC_PLAN_GENERATOR, *opts, base, 0, tensors, n_tensors, "lq_generator/", &rc));
    if (!rc) pl->sal_hq.reset(make_sub(CIC_PLAN_SALIENCY, *opts, 2 * base, 0, tensors, n_tensors, "latent_saliency_hq/", &rc));
    if (!rc) pl->sal_lq.reset(make_sub(CIC_PLAN_SALIENCY, *opts, base, 0, tensors, n_tensors, "latent_saliency_lq/", &rc));
    if (!rc) pl->rd.reset(make_sub(CIC_PLAN_RD, *opts, base, 0, tensors, n_tensors, "rd_optimizer/", &rc));
    if (!rc && opts->precision == CIC_PREC_TC) rc = build_adaptive_tc(pl);
  } else {
    rc = build_any(pl, tensors, n_tensors, "");
  }
  if (rc) {
    delete pl;
    return nullptr;
  }
  return pl;
}

extern "C" void cic_plan_destroy(cic_plan* plan) { delete plan; }

extern "C" int cic_plan_last_launch_count(const cic_plan* plan) { return plan ? (int)plan->last_launches : 0; }

extern "C" int cic_plan_set_profiling(cic_plan* plan, int on) {
  CIC_REQUIRE(plan, "cic_plan_set_profiling: null plan");
  plan->prof.on = on != 0;
  if (!on) plan->prof.reset();
  return CIC_OK;
}

extern "C" size_t cic_plan_get_profile(cic_plan* plan, char* buf, size_t cap) {
  if (!plan) return 0;
  const std::string rep = plan->prof.report();
  if (buf && cap > 0) {
    const size_t n = rep.size() < cap - 1 ? rep.size() : cap - 1;
    memcpy(buf, rep.data(), n);
    buf[n] = 0;
  }
  return rep.size() + 1;
}

static int dispatch(cic_plan* pl, Ctx& c, int batch, int h, int w, const void* a0, const void* a1, const void* a2,
                    const void* a3, void* o0, void* o1, void* o2, void* o3, void* o4) {
  const bool tc = pl->opts.precision == CIC_PREC_TC;
  switch (pl->kind) {
    case CIC_PLAN_AUTOENCODER:
      return tc ? autoencoder_forward_tc(pl, c, (const float*)a0, (float*)o0, (uint8_t*)o1, batch, h, w)
                : autoencoder_forward_f32(pl, c, (const float*)a0, (float*)o0, (uint8_t*)o1, batch, h, w);
    case CIC_PLAN_ENCODER:
      return tc ? encoder_forward_tc(pl, c, (const float*)a0, (float*)o0, (float*)o1, (float*)o2, (float*)o3, batch)
                : encoder_forward_f32(pl, c, (const float*)a0, (float*)o0, (float*)o1, (float*)o2, (float*)o3, batch);
    case CIC_PLAN_GENERATOR:
      return tc ? generator_forward_tc(pl, c, (const float*)a0, (const float*)a1, (const float*)a2, (const float*)a3, (float*)o0, batch)
                : generator_forward_f32(pl, c, (const float*)a0, (const float*)a1, (const float*)a2, (const float*)a3, (float*)o0, batch);
    case CIC_PLAN_SALIENCY:
      return saliency_forward_f32(pl, c, (const float*)a0, (float*)o0, batch);
    case CIC_PLAN_RD:
      return tc ? rd_forward_tc(pl, c, (const float*)a0, (const float*)a1, (float*)o0, batch)
                : rd_forward_f32(pl, c, (const float*)a0, (const float*)a1, (float*)o0, batch);
    case CIC_PLAN_ADAPTIVE: {
      const cic_adaptive_io* io = (const cic_adaptive_io*)a0;
      return tc ? adaptive_forward_tc(pl, c, io, batch, h, w) : adaptive_forward_f32(pl, c, io, batch, h, w);
    }
  }
  set_error("unknown plan kind %d", pl->kind);
  return CIC_ERR_INVALID;
}

extern "C" size_t cic_plan_workspace_bytes(const cic_plan* plan, int batch, int h, int w) {
  if (!plan || batch <= 0) return 0;
  Ctx c;
  c.dry = true;
  static const cic_adaptive_io empty_io{};
  cic_plan* pl = const_cast<cic_plan*>(plan);
  if (plan->kind != CIC_PLAN_AUTOENCODER && plan->kind != CIC_PLAN_ADAPTIVE) { h = plan->opts.img_h; w = plan->opts.img_w; }
  dispatch(pl, c, batch, h, w, plan->kind == CIC_PLAN_ADAPTIVE ? (const void*)&empty_io : nullptr, nullptr, nullptr, nullptr,
           nullptr, nullptr, nullptr, nullptr, nullptr);
  return c.arena.peak + 256;
}

static int run(cic_plan* pl, int batch, int h, int w, void* ws, size_t ws_bytes, void* stream, const void* a0,
               const void* a1, const void* a2, const void* a3, void* o0, void* o1, void* o2, void* o3) {
  CIC_REQUIRE(pl, "null plan");
  CIC_REQUIRE(batch >= 0, "negative batch");
  if (batch == 0) return CIC_OK;
  const size_t need = cic_plan_workspace_bytes(pl, batch, h, w);
  if (need > 256 && (!ws || ws_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
    return CIC_ERR_WORKSPACE;
  }
  Ctx c;
  c.arena.base = (char*)ws;
  c.arena.cap = ws_bytes;
  c.st = (cudaStream_t)stream;
  c.prof = &pl->prof;
  if (pl->prof.on) { pl->prof.reset(); pl->prof.prefix = ""; }
  const long long before = g_launch_count;
  int rc = dispatch(pl, c, batch, h, w, a0, a1, a2, a3, o0, o1, o2, o3, nullptr);
  pl->last_launches = g_launch_count - before;
  if (rc == CIC_OK && c.arena.overflow) {
    set_error("internal: workspace overflow");
    return CIC_ERR_WORKSPACE;
  }
  return rc;
}

extern "C" int cic_autoencoder_forward(cic_plan* plan, const float* d_x, float* d_y, uint8_t* d_y_u8, int batch, int h,
                                       int w, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_AUTOENCODER, "cic_autoencoder_forward: not an autoencoder plan");
  CIC_REQUIRE(batch == 0 || (d_x && d_y), "cic_autoencoder_forward: null pointer");
  CIC_REQUIRE(h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0, "cic_autoencoder_forward: H and W must be multiples of 4, got %dx%d", h, w);
  return run(plan, batch, h, w, d_workspace, workspace_bytes, stream, d_x, nullptr, nullptr, nullptr, d_y, d_y_u8, nullptr, nullptr);
}

extern "C" int cic_encoder_forward(cic_plan* plan, const float* d_img, float* d_latent, float* d_x1, float* d_x2,
                                   float* d_x3, int batch, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_ENCODER, "cic_encoder_forward: not an encoder plan");
  CIC_REQUIRE(batch == 0 || (d_img && d_latent), "cic_encoder_forward: null pointer");
  return run(plan, batch, plan->opts.img_h, plan->opts.img_w, d_workspace, workspace_bytes, stream, d_img, nullptr, nullptr,
             nullptr, d_latent, d_x1, d_x2, d_x3);
}

extern "C" int cic_generator_forward(cic_plan* plan, const float* d_latent, const float* d_skip1, const float* d_skip2,
                                     const float* d_skip3, float* d_out, int batch, void* d_workspace,
                                     size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_GENERATOR, "cic_generator_forward: not a generator plan");
  CIC_REQUIRE(batch == 0 || (d_latent && d_skip1 && d_skip2 && d_skip3 && d_out), "cic_generator_forward: null pointer");
  return run(plan, batch, plan->opts.img_h, plan->opts.img_w, d_workspace, workspace_bytes, stream, d_latent, d_skip1, d_skip2,
             d_skip3, d_out, nullptr, nullptr, nullptr);
}

extern "C" int cic_saliency_forward(cic_plan* plan, const float* d_latent, float* d_score, int batch, void* d_workspace,
                                    size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_SALIENCY, "cic_saliency_forward: not a saliency plan");
  CIC_REQUIRE(batch == 0 || (d_latent && d_score), "cic_saliency_forward: null pointer");
  return run(plan, batch, 0, 0, d_workspace, workspace_bytes, stream, d_latent, nullptr, nullptr, nullptr, d_score, nullptr,
             nullptr, nullptr);
}

extern "C" int cic_rd_forward(cic_plan* plan, const float* d_mask, const float* d_bpp, float* d_rd_params, int batch,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_RD, "cic_rd_forward: not an rd plan");
  CIC_REQUIRE(batch == 0 || (d_mask && d_bpp && d_rd_params), "cic_rd_forward: null pointer");
  return run(plan, batch, plan->opts.img_h, plan->opts.img_w, d_workspace, workspace_bytes, stream, d_mask, d_bpp, nullptr,
             nullptr, d_rd_params, nullptr, nullptr, nullptr);
}

extern "C" int cic_adaptive_forward(cic_plan* plan, const cic_adaptive_io* io, int n_img, int img_h, int img_w,
                                    void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_ADAPTIVE, "cic_adaptive_forward: not an adaptive plan");
  CIC_REQUIRE(io && (n_img == 0 || (io->d_img && io->d_mask && io->d_bpp)), "cic_adaptive_forward: null input");
  const int T = plan->opts.img_h;
  CIC_REQUIRE(plan->opts.img_h == plan->opts.img_w, "cic_adaptive_forward: square model tiles only");
  CIC_REQUIRE(img_h > 0 && img_w > 0 && T > 0, "cic_adaptive_forward: bad image size %dx%d", img_h, img_w);
  return run(plan, n_img, img_h, img_w, d_workspace, workspace_bytes, stream, io, nullptr, nullptr, nullptr, nullptr, nullptr,
             nullptr, nullptr);
}

extern "C" int cic_adaptive_forward_phase(cic_plan* plan, const cic_adaptive_io* io, const cic_adaptive_state* state, int phase,
                                          int tile0, int n_img, int img_h, int img_w, void* d_workspace, size_t workspace_bytes,
                                          void* stream) {
  CIC_REQUIRE(plan && plan->kind == CIC_PLAN_ADAPTIVE, "cic_adaptive_forward_phase: not an adaptive plan");
  CIC_REQUIRE(plan->opts.precision == CIC_PREC_TC, "cic_adaptive_forward_phase: tensor-core plans only");
  CIC_REQUIRE(phase == CIC_PHASE_ENCODE || phase == CIC_PHASE_LATENT || phase == CIC_PHASE_DECODE,
              "cic_adaptive_forward_phase: unknown phase %d", phase);
  CIC_REQUIRE(io && state && tile0 >= 0 && n_img >= 0, "cic_adaptive_forward_phase: bad argument");
  if (n_img == 0) return CIC_OK;
  CIC_REQUIRE(io->d_bpp, "cic_adaptive_forward_phase: null bpp");
  CIC_REQUIRE(phase != CIC_PHASE_ENCODE || (io->d_img && io->d_mask), "cic_adaptive_forward_phase: ENCODE needs image and mask");
  CIC_REQUIRE(phase != CIC_PHASE_DECODE || io->d_mask, "cic_adaptive_forward_phase: DECODE needs the mask");
  CIC_REQUIRE(phase == CIC_PHASE_ENCODE || (io->d_hq_latent_q && io->d_lq_latent_q) || phase == CIC_PHASE_DECODE,
              "cic_adaptive_forward_phase: LATENT needs the quantised latent outputs");
  for (int e = 0; e < 2; ++e)
    CIC_REQUIRE(state->x1[e] && state->x2[e] && state->x3[e] && state->x4_hi[e] && state->x4_lo[e] && state->g0[e],
                "cic_adaptive_forward_phase: null state buffer");
  const int T = plan->opts.img_h;
  CIC_REQUIRE(plan->opts.img_h == plan->opts.img_w, "cic_adaptive_forward_phase: square model tiles only");
  CIC_REQUIRE(img_h > 0 && img_w > 0 && T > 0, "cic_adaptive_forward_phase: bad image size %dx%d", img_h, img_w);
  const size_t need = cic_plan_workspace_bytes(plan, n_img, img_h, img_w);  // upper bound: the one-call forward of as many images
  if (!d_workspace || workspace_bytes < need) {
    set_error("workspace too small: need %zu bytes, got %zu", need, d_workspace ? workspace_bytes : (size_t)0);
    return CIC_ERR_WORKSPACE;
  }
  Ctx c;
  c.arena.base = (char*)d_workspace;
  c.arena.cap = workspace_bytes;
  c.st = (cudaStream_t)stream;
  c.prof = &plan->prof;
  if (plan->prof.on) { plan->prof.reset(); plan->prof.prefix = ""; }
  const long long before = g_launch_count;
  const int rc = adaptive_forward_tc(plan, c, io, n_img, img_h, img_w, phase, state, tile0);
  plan->last_launches = g_launch_count - before;
  if (rc == CIC_OK && c.arena.overflow) {
    set_error("internal: workspace overflow");
    return CIC_ERR_WORKSPACE;
  }
  return rc;
}

extern "C" size_t cic_attention_workspace_bytes(int batch, int tokens, int channels) {
  if (batch <= 0) return 0;
  const int chunk = batch < 32 ? batch : 32;
  const size_t N = (size_t)channels / 8 * 2 + channels;
  return ((size_t)chunk * tokens * N + (size_t)chunk * tokens * tokens) * sizeof(float) + 1024 +
         ((size_t)channels * N + N) * sizeof(float) + 1024;
}

extern "C" int cic_self_attention_f32(const float* d_x, const float* d_wq, const float* d_bq, const float* d_wk,
                                      const float* d_bk, const float* d_wv, const float* d_bv, float gamma, float* d_y,
                                      int batch, int tokens, int channels, void* d_workspace, size_t workspace_bytes,
                                      void* stream) {
  CIC_REQUIRE(d_x && d_wq && d_wk && d_wv && d_y, "cic_self_attention_f32: null pointer");
  CIC_REQUIRE(channels % 8 == 0 && channels >= 8 && tokens > 0, "cic_self_attention_f32: bad shape");
  if (batch <= 0) return CIC_OK;
  CIC_REQUIRE(d_workspace && workspace_bytes >= cic_attention_workspace_bytes(batch, tokens, channels),
              "cic_self_attention_f32: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c;
  c.arena.base = (char*)d_workspace;
  c.arena.cap = workspace_bytes;
  c.st = st;
  const int dq = channels / 8, N = 2 * dq + channels;
  // pack [C][N] = q | k | v and the bias row with strided 2-D copies
  float* wqkv = c.arena.f32((size_t)channels * N);
  float* bqkv = c.arena.f32(N);
  CIC_CHECK_CUDA(cudaMemcpy2DAsync(wqkv, N * sizeof(float), d_wq, dq * sizeof(float), dq * sizeof(float), channels, cudaMemcpyDeviceToDevice, st));
  CIC_CHECK_CUDA(cudaMemcpy2DAsync(wqkv + dq, N * sizeof(float), d_wk, dq * sizeof(float), dq * sizeof(float), channels, cudaMemcpyDeviceToDevice, st));
  CIC_CHECK_CUDA(cudaMemcpy2DAsync(wqkv + 2 * dq, N * sizeof(float), d_wv, channels * sizeof(float), channels * sizeof(float), channels, cudaMemcpyDeviceToDevice, st));
  CIC_CHECK_CUDA(cudaMemsetAsync(bqkv, 0, N * sizeof(float), st));
  if (d_bq) CIC_CHECK_CUDA(cudaMemcpyAsync(bqkv, d_bq, dq * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_bk) CIC_CHECK_CUDA(cudaMemcpyAsync(bqkv + dq, d_bk, dq * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_bv) CIC_CHECK_CUDA(cudaMemcpyAsync(bqkv + 2 * dq, d_bv, channels * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return attention_f32(c, "attention", d_x, wqkv, bqkv, nullptr, gamma, false, d_y, batch, tokens, channels);
}
