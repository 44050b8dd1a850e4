// cic_plan: device-resident packed weights + the graph walkers of the reference's models.
#pragma once
#include "common.cuh"
#include "igemm_simt.cuh"

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace cic {

struct DevTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  size_t numel() const {
    size_t n = 1;
    for (auto s : shape) n *= (size_t)s;
    return n;
  }
};

// name -> device tensor; owns the allocations
class WeightStore {
 public:
  ~WeightStore();
  int upload(const std::string& name, const float* h, const std::vector<int64_t>& shape);
  const DevTensor* find(const std::string& name) const;
  float* ptr(const std::string& name) const {
    const DevTensor* t = find(name);
    return t ? t->p : nullptr;
  }
  size_t bytes() const { return bytes_; }

 private:
  std::map<std::string, DevTensor> t_;
  size_t bytes_ = 0;
};

// bump allocator over the caller's workspace; with base == nullptr it only measures
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool overflow = false;
  void* alloc_bytes(size_t n) {
    off = (off + 255) & ~(size_t)255;
    void* r = base ? base + off : nullptr;
    off += n;
    if (off > peak) peak = off;
    if (base && off > cap) overflow = true;
    return r;
  }
  float* f32(size_t n) { return (float*)alloc_bytes(n * sizeof(float)); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// Optional per-layer timing: CUDA events recorded on the launch stream around every layer of a
// forward call (cic_plan_set_profiling).  bench.py uses it for the live roofline of the conv GEMMs.
struct Profiler {
  struct Rec {
    std::string name;
    cudaEvent_t e0, e1;
    double flops, bytes;
    int kind;  // KernelKind of the layer's heavy kernel
  };
  bool on = false;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  std::string prefix;
  ~Profiler();
  cudaEvent_t get_event();
  void reset();
  int begin(const std::string& name, double flops, double bytes, cudaStream_t st);
  void end(int idx, cudaStream_t st);
  std::string report();  // synchronises; one "name,ms,flops,bytes" line per layer
};

struct Ctx {
  Arena arena;
  cudaStream_t st = nullptr;
  bool dry = false;  // measure workspace only, launch nothing
  Profiler* prof = nullptr;
};

struct Scope {  // RAII layer bracket
  Ctx& c;
  int idx = -1;
  Scope(Ctx& ctx, const std::string& name, double flops = 0, double bytes = 0) : c(ctx) {
    if (c.prof && c.prof->on && !c.dry) idx = c.prof->begin(name, flops, bytes, c.st);
  }
  ~Scope() {
    if (idx >= 0) c.prof->end(idx, c.st);
  }
  Scope(const Scope&) = delete;
};

int run_dense(const float* x, const float* kernel, const float* bias, const float* scale, const float* shift, float* y,
              int batch, int in_dim, int out_dim, int act, float* ws, size_t ws_floats, cudaStream_t st);
void pack_deconv_phases(const float* k, int cout, int cin, std::vector<float>& out);
// Batch item b of a model is tile (ty, tx) of image b / (tiles_x * tiles_y) in an (n_img, IH, IW, C) tensor, so the
// first and last layers of the tiled codec address the image layout directly (no gather / scatter pass).
// tiles_x == 0: plain dense batch.
struct TileMap {
  int tiles_x = 0, tiles_y = 0, IH = 0, IW = 0;
};
int launch_conv_k4s2_c3(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                        float* out_f32, int batch, int H, int W, int act, const TileMap& tm, cudaStream_t st);
// encoder conv1 on the tensor cores for one or both encoders of the adaptive model (conv1_tc.cu)
size_t conv1_tc_image_bytes(int n_enc);
int conv1_tc_pack(const float* w0, const float* w1, uint8_t* img, cudaStream_t st);
int launch_conv1_tc(const float* x, const uint8_t* wimg, const float* bias, int n_enc, __nv_bfloat16* const* out_hi,
                    __nv_bfloat16* const* out_lo, int batch, int H, int W, const TileMap& tm, cudaStream_t st);
// 1- / 3-channel first layers with 32 outputs on the tensor cores (first_conv_tc.cu): autoencoder conv1 + pool, RD conv1
size_t first_conv_image_bytes();
int first_conv_pack(const float* w, int nk, uint8_t* img, cudaStream_t st);
int launch_first_conv_tc_pool(const float* x, const uint8_t* wimg, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* pool_hi,
                              int batch, int H, int W, int act, cudaStream_t st);
int launch_first_conv_tc_c1s2(const float* x, const uint8_t* wimg, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                              int batch, int H, int W, int act, const TileMap& tm, cudaStream_t st);
int launch_conv_k3s1_c3_pool(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* pool_hi,
                             int batch, int H, int W, int act, cudaStream_t st);
int launch_conv_k3s2_c1(const float* x, const float* wgt, const float* bias, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                        int batch, int H, int W, int act, const TileMap& tm, cudaStream_t st);
int launch_expand_bpp(const float* bpp, float* bpp_t, float* qs_t, int n_tiles, int tiles_per_img, cudaStream_t st);

// named raw device buffers (packed bf16 weights of the tensor-core path)
class BufStore {
 public:
  ~BufStore() {
    for (auto& kv : b_) cudaFree(kv.second);
  }
  void* alloc(const std::string& name, size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    auto it = b_.find(name);
    if (it != b_.end()) cudaFree(it->second);
    b_[name] = p;
    return p;
  }
  void* ptr(const std::string& name) const {
    auto it = b_.find(name);
    return it == b_.end() ? nullptr : it->second;
  }

 private:
  std::map<std::string, void*> b_;
};

}  // namespace cic

struct cic_plan;
namespace cic {
// walkers: CIC_PREC_FP32 (plans.cu) and CIC_PREC_TC (plans_tc.cu)
int build_plan_tc(cic_plan* pl, const cic_tensor* tensors, int n, const std::string& prefix);
int build_adaptive_tc(cic_plan* pl);
int autoencoder_forward_f32(cic_plan* pl, Ctx& c, const float* x, float* y, uint8_t* y_u8, int B, int H, int W);
int encoder_forward_f32(cic_plan* pl, Ctx& c, const float* img, float* latent, float* x1, float* x2, float* x3, int B);
int generator_forward_f32(cic_plan* pl, Ctx& c, const float* latent, const float* s1, const float* s2, const float* s3,
                          float* out, int B);
int saliency_forward_f32(cic_plan* pl, Ctx& c, const float* latent, float* score, int B);
// mlp_fused.cu
int launch_rd_tail_fused(const float* feat, int ld, const float* bpp, const float* w1, const float* b1, const float* w2, const float* b2,
                         float* rd_params, int B, cudaStream_t st);
int rd_forward_f32(cic_plan* pl, Ctx& c, const float* mask, const float* bpp, float* rd_params, int B);
int adaptive_forward_f32(cic_plan* pl, Ctx& c, const cic_adaptive_io* io, int n_img, int img_h, int img_w);
int launch_rd_tail(cic_plan* pl, Ctx& c, const float* bpp, float* feat, float* d1, float* base, float* rd_params, int B);
int rd_forward_tc(cic_plan* pl, Ctx& c, const float* mask, const float* bpp, float* rd_params, int B, const TileMap& tm = TileMap());
int autoencoder_forward_tc(cic_plan* pl, Ctx& c, const float* x, float* y, uint8_t* y_u8, int B, int H, int W);
int encoder_forward_tc(cic_plan* pl, Ctx& c, const float* img, float* latent, float* x1, float* x2, float* x3, int B);
int generator_forward_tc(cic_plan* pl, Ctx& c, const float* latent, const float* s1, const float* s2, const float* s3,
                         float* out, int B);
int adaptive_forward_tc(cic_plan* pl, Ctx& c, const cic_adaptive_io* io, int n_img, int img_h, int img_w, int phase = 0,
                        const cic_adaptive_state* state = nullptr, int tile0 = 0);
}  // namespace cic

struct cic_plan {
  int kind = 0;
  cic_plan_opts opts{};
  cic::WeightStore w;
  cic::BufStore tcw;
  float attn_gamma = 0.f;  // SelfAttention.gamma, read once at plan creation (tensor-core path)
  long long last_launches = 0;
  cic::Profiler prof;
  // CIC_PLAN_ADAPTIVE owns its seven sub-models
  std::unique_ptr<cic_plan> hq_enc, lq_enc, hq_gen, lq_gen, sal_hq, sal_lq, rd;
};
