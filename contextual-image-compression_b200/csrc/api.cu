// Library-level plumbing of libcic.so + the stand-alone fp32 operators of include/cic.h.
#include "common.cuh"
#include "igemm_simt.cuh"

#include <string>
#include <vector>

namespace cic {

thread_local long long g_launch_count = 0;
thread_local int g_last_kernel_kind = 0;
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};  // per device ordinal
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const int slot = dev & 63;
  if (cached[slot] > 0) return cached[slot];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached[slot] = n;
    return n;
  }
  return 148;  // B200
}

// Conv2DTranspose (4,4,Cout,Cin) -> four phase matrices [phase = py*2+px][K = (ty,tx,ci)][Cout].
// Output pixel (2y+py, 2x+px) reads input rows y-1+ty (py = 0) or y+ty (py = 1); the kernel tap that
// connects input row i to output row o is k = o + 1 - 2i (padding 1 of the 'same' crop).
void pack_deconv_phases(const float* k, int cout, int cin, std::vector<float>& out) {
  out.assign((size_t)4 * 4 * cin * cout, 0.f);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          const int ky = py == 0 ? 3 - 2 * ty : 2 - 2 * ty;
          const int kx = px == 0 ? 3 - 2 * tx : 2 - 2 * tx;
          for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
              out[((((size_t)(py * 2 + px) * 2 + ty) * 2 + tx) * cin + ci) * cout + co] =
                  k[(((size_t)ky * 4 + kx) * cout + co) * cin + ci];
        }
}

}  // namespace cic

using namespace cic;

extern "C" int cic_version(void) { return CIC_VERSION; }
extern "C" const char* cic_last_error(void) { return g_err; }

extern "C" int cic_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  CIC_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CIC_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm) *sm = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return CIC_OK;
}

static void fill_conv(IGemmParams& p, const float* x, int batch, int h, int w, int cin) {
  p = IGemmParams{};
  p.src[0] = ConvSrc{x, cin, cin, 0};
  p.nsrc = 1;
  p.Cin = cin;
  p.batch = batch;
  p.H = h;
  p.W = w;
  p.alpha = 1.f;
  p.splits = 1;
  p.out_ys = p.out_xs = 1;
}

extern "C" int cic_conv2d_nhwc_f32(const float* d_x, const float* d_kernel, const float* d_bias, const float* d_scale,
                                   const float* d_shift, float* d_y, int batch, int h, int w, int cin, int cout, int kh,
                                   int kw, int stride, int act, void* stream) {
  CIC_REQUIRE(batch == 0 || (d_x && d_kernel && d_y), "cic_conv2d_nhwc_f32: null pointer");
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && kh > 0 && kw > 0 && stride > 0,
              "cic_conv2d_nhwc_f32: bad shape");
  CIC_REQUIRE((d_scale == nullptr) == (d_shift == nullptr), "cic_conv2d_nhwc_f32: scale and shift go together");
  IGemmParams p;
  fill_conv(p, d_x, batch, h, w, cin);
  p.Ho = same_out(h, stride);
  p.Wo = same_out(w, stride);
  p.kh = kh; p.kw = kw; p.stride = stride;
  p.pad_t = same_pad_before(h, kh, stride);
  p.pad_l = same_pad_before(w, kw, stride);
  p.Bmat = d_kernel; p.N = cout; p.ldb = cout;
  p.bias = d_bias; p.scale = d_scale; p.shift = d_shift; p.act = act;
  p.out = d_y; p.out_ld = cout; p.out_H = p.Ho; p.out_W = p.Wo;
  return launch_igemm(p, (cudaStream_t)stream);
}

extern "C" int cic_conv2d_transpose4x4s2_nhwc_f32(const float* d_x, const float* d_kernel, const float* d_bias,
                                                  const float* d_scale, const float* d_shift, float* d_y, int batch,
                                                  int h, int w, int cin, int cout, int act, void* stream) {
  CIC_REQUIRE(d_x && d_kernel && d_y, "cic_conv2d_transpose4x4s2_nhwc_f32: null pointer");
  CIC_REQUIRE(batch >= 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "cic_conv2d_transpose4x4s2_nhwc_f32: bad shape");
  // stand-alone operator: repack the phase matrices on the fly (plans do this once at creation)
  std::vector<float> hk((size_t)16 * cin * cout), packed;
  cudaStream_t st = (cudaStream_t)stream;
  CIC_CHECK_CUDA(cudaMemcpyAsync(hk.data(), d_kernel, hk.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  CIC_CHECK_CUDA(cudaStreamSynchronize(st));
  pack_deconv_phases(hk.data(), cout, cin, packed);
  float* d_packed = nullptr;
  CIC_CHECK_CUDA(cudaMalloc(&d_packed, packed.size() * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(d_packed, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  int rc = CIC_OK;
  if (e != cudaSuccess) { set_error("memcpy failed: %s", cudaGetErrorString(e)); rc = CIC_ERR_CUDA; }
  for (int ph = 0; ph < 4 && rc == CIC_OK; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    IGemmParams p;
    fill_conv(p, d_x, batch, h, w, cin);
    p.Ho = h; p.Wo = w; p.kh = 2; p.kw = 2; p.stride = 1;
    p.pad_t = py == 0 ? 1 : 0;
    p.pad_l = px == 0 ? 1 : 0;
    p.Bmat = d_packed + (size_t)ph * 4 * cin * cout; p.N = cout; p.ldb = cout;
    p.bias = d_bias; p.scale = d_scale; p.shift = d_shift; p.act = act;
    p.out = d_y; p.out_ld = cout; p.out_H = 2 * h; p.out_W = 2 * w;
    p.out_ys = 2; p.out_xs = 2; p.out_y0 = py; p.out_x0 = px;
    rc = launch_igemm(p, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(d_packed);
  return rc;
}

static int dense_splits(int batch, int in_dim, int out_dim) {
  // enough CTAs to fill the GPU when batch x out_dim alone gives too few tiles
  const long long tiles = (long long)((batch + 127) / 128) * ((out_dim + 63) / 64);
  const int chunks = (in_dim + 15) / 16;
  int splits = 1;
  const int target = 2 * sm_count();
  if (tiles < target) splits = (int)((target + tiles - 1) / tiles);
  if (splits > chunks / 8) splits = chunks / 8;
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  return splits;
}

extern "C" size_t cic_dense_workspace_bytes(int batch, int in_dim, int out_dim) {
  const int s = dense_splits(batch, in_dim, out_dim);
  return s > 1 ? (size_t)s * batch * out_dim * sizeof(float) : 0;
}

namespace cic {
int run_dense(const float* x, const float* kernel, const float* bias, const float* scale, const float* shift, float* y,
              int batch, int in_dim, int out_dim, int act, float* ws, size_t ws_floats, cudaStream_t st) {
  IGemmParams p;
  fill_conv(p, x, batch, 1, 1, in_dim);
  p.Ho = 1; p.Wo = 1; p.kh = 1; p.kw = 1; p.stride = 1;
  p.Bmat = kernel; p.N = out_dim; p.ldb = out_dim;
  p.bias = bias; p.scale = scale; p.shift = shift; p.act = act;
  p.out = y; p.out_ld = out_dim; p.out_H = 1; p.out_W = 1;
  p.splits = dense_splits(batch, in_dim, out_dim);
  if (p.splits > 1) {
    CIC_REQUIRE(ws && ws_floats >= (size_t)p.splits * batch * out_dim, "dense: workspace too small");
    p.partial = ws;
  }
  return launch_igemm(p, st);
}
}  // namespace cic

extern "C" int cic_dense_f32(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch,
                             int in_dim, int out_dim, int act, void* d_workspace, size_t workspace_bytes, void* stream) {
  CIC_REQUIRE(batch >= 0 && in_dim > 0 && out_dim > 0, "cic_dense_f32: bad shape");
  if (batch == 0) return CIC_OK;
  CIC_REQUIRE(d_x && d_kernel && d_y, "cic_dense_f32: null pointer");
  return run_dense(d_x, d_kernel, d_bias, nullptr, nullptr, d_y, batch, in_dim, out_dim, act, (float*)d_workspace,
                   workspace_bytes / sizeof(float), (cudaStream_t)stream);
}
