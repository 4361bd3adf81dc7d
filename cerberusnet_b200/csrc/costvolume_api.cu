// costvolume_api.cu -- the C ABI of libcerberus_costvolume.so (include/cerberus_costvolume.h,
// include/cerberus_trt_plugin.h): argument checking, geometry, dispatch.  No torch, no TensorRT.
#include <atomic>
#include <cmath>
#include <cstring>

#include "../../include/cerberus_trt_plugin.h"
#include "costvolume_launch.h"

namespace cerb {

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static size_t dtype_size(int dtype) { return dtype == CERB_F32 ? 4 : 2; }

static void fill_strides(const int64_t in[4], long long out[3], long long C, long long H, long long W, bool& w_ok) {
  if (in[0] == 0 && in[1] == 0 && in[2] == 0 && in[3] == 0) {
    out[0] = C * H * W; out[1] = H * W; out[2] = W;
    w_ok = true;
  } else {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
    w_ok = in[3] == 1;
  }
}

// correlation_cuda.cpp:6-14
static int build_geom(const cerb_corr_params* p, bool has_flow, Geom& g) {
  if (!p) return CERB_EINVAL;
  if (p->batch < 1 || p->channels < 1 || p->height < 1 || p->width < 1) return CERB_EINVAL;
  if (p->pad_size < 0 || p->kernel_size < 1 || p->max_displacement < 0 || p->stride1 < 1 || p->stride2 < 1)
    return CERB_EINVAL;
  if (p->dtype != CERB_F32 && p->dtype != CERB_F16 && p->dtype != CERB_BF16) return CERB_EINVAL;
  if (p->x2_batch_roll < 0 || p->x2_batch_roll >= p->batch) return CERB_EINVAL;
  if (has_flow) {
    if (p->warp_mode < CERB_WARP_TORCH || p->warp_mode > CERB_WARP_TORCH_CPU) return CERB_EINVAL;
    if (p->height < 2 || p->width < 2) return CERB_EINVAL;  // grid normalisation divides by size-1
  }
  g.B = p->batch; g.C = p->channels; g.H = p->height; g.W = p->width;
  g.pad = p->pad_size; g.k = p->kernel_size; g.md = p->max_displacement; g.s1 = p->stride1; g.s2 = p->stride2;
  g.kr = (g.k - 1) / 2;
  g.r = g.md / g.s2;
  g.D = 2 * g.r + 1;
  g.D2 = g.D * g.D;
  const int border = g.kr + g.md;
  g.outH = (int)std::ceil((float)(g.H + 2 * g.pad - 2 * border) / (float)g.s1);
  g.outW = (int)std::ceil((float)(g.W + 2 * g.pad - 2 * border) / (float)g.s1);
  if (g.outH < 1 || g.outW < 1) return CERB_ESHAPE;
  g.warp_mode = p->warp_mode;
  g.has_act = (p->leaky_slope == p->leaky_slope) && p->leaky_slope >= 0.f;  // NaN / negative = off
  g.slope = g.has_act ? p->leaky_slope : 1.f;
  g.unnorm_fma = 1;
  g.x2roll = p->x2_batch_roll;
  bool ok1, ok2, ok3, ok4;
  fill_strides(p->x1_stride, g.x1s, g.C, g.H, g.W, ok1);
  fill_strides(p->x2_stride, g.x2s, g.C, g.H, g.W, ok2);
  fill_strides(p->flow_stride, g.fls, 2, g.H, g.W, ok3);
  fill_strides(p->out_stride, g.os, g.D2, g.outH, g.outW, ok4);
  if (!(ok1 && ok2 && ok3 && ok4)) return CERB_ESTRIDE;
  // kernels index one (n) slice with 32-bit tap offsets
  const long long lim = 0x7fffffffLL;
  if (g.x1s[1] * g.C >= lim || g.x2s[1] * g.C >= lim || g.os[1] * g.D2 >= lim) return CERB_ESTRIDE;
  for (int i = 0; i < 3; ++i)
    if (g.x1s[i] < 0 || g.x2s[i] < 0 || g.fls[i] < 0 || g.os[i] < 0) return CERB_ESTRIDE;
  return CERB_OK;
}

static bool is_fast(const Geom& g) { return g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md >= 4; }  // forward fast path

}  // namespace cerb

using namespace cerb;

static int trt_params(const cerb_trt_corr_fields* f, const int64_t dims[4], int trt_type, cerb_corr_params* p) {
  if (!f) return CERB_EINVAL;
  memset(p, 0, sizeof(*p));
  p->batch = (int32_t)dims[0]; p->channels = (int32_t)dims[1]; p->height = (int32_t)dims[2]; p->width = (int32_t)dims[3];
  p->pad_size = f->pad_size; p->kernel_size = f->kernel_size; p->max_displacement = f->max_displacement;
  p->stride1 = f->stride1; p->stride2 = f->stride2; p->corr_multiply = f->corr_multiply;
  if (trt_type == CERB_TRT_FLOAT) p->dtype = CERB_F32;
  else if (trt_type == CERB_TRT_HALF) p->dtype = CERB_F16;
  else return CERB_EUNSUPPORTED;  // reference throws std::runtime_error, correlation.cu:159-162
  p->warp_mode = CERB_WARP_TRT;
  p->leaky_slope = NAN;
  return CERB_OK;
}

template <typename Desc>
static int trt_enqueue_impl(const cerb_trt_corr_fields* f, int warp_mode, float slope, bool fused, const Desc* in,
                            const Desc* out, const void* const* inputs, void* const* outputs, cerb_stream_t stream) {
  if (!f || !in || !out || !inputs || !outputs) return CERB_EINVAL;
  if (in[0].dims.nbDims != 4 || in[1].dims.nbDims != 4 || out[0].dims.nbDims != 4) return CERB_EINVAL;
  for (int i = 0; i < 4; ++i)
    if (in[0].dims.d[i] != in[1].dims.d[i]) return CERB_EINVAL;
  if (in[0].type != in[1].type || out[0].type != in[0].type) return CERB_EUNSUPPORTED;
  if (in[0].format != CERB_TRT_LINEAR || in[1].format != CERB_TRT_LINEAR || out[0].format != CERB_TRT_LINEAR)
    return CERB_EUNSUPPORTED;
  const int64_t d[4] = {(int64_t)in[0].dims.d[0], (int64_t)in[0].dims.d[1], (int64_t)in[0].dims.d[2],
                        (int64_t)in[0].dims.d[3]};
  cerb_corr_params p;
  int rc = trt_params(f, d, in[0].type, &p);
  if (rc != CERB_OK) return rc;
  const float* flow = nullptr;
  if (fused) {
    if (in[2].dims.nbDims != 4 || in[2].dims.d[0] != in[0].dims.d[0] || in[2].dims.d[1] != 2 ||
        in[2].dims.d[2] != in[0].dims.d[2] || in[2].dims.d[3] != in[0].dims.d[3] || in[2].type != CERB_TRT_FLOAT)
      return CERB_EINVAL;
    flow = (const float*)inputs[2];
    p.warp_mode = warp_mode;
    p.leaky_slope = slope;
  }
  int32_t oc, oh, ow;
  rc = cerb_corr_output_dims(&p, &oc, &oh, &ow);
  if (rc != CERB_OK) return rc;
  if (out[0].dims.d[0] != in[0].dims.d[0] || out[0].dims.d[1] != oc || out[0].dims.d[2] != oh || out[0].dims.d[3] != ow)
    return CERB_ESHAPE;
  return cerb_warp_corr_forward(&p, inputs[0], inputs[1], flow, outputs[0], stream);
}

extern "C" {

int cerb_abi_version(void) { return CERB_ABI_VERSION; }

const char* cerb_error_string(int code) {
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case CERB_OK: return "ok";
    case CERB_EINVAL: return "invalid argument (null pointer, non-positive size or unknown enum)";
    case CERB_ESHAPE: return "correlation parameters give an empty output";
    case CERB_ESTRIDE: return "unsupported strides (innermost stride must be 1; one batch item must fit 32-bit offsets)";
    case CERB_EUNSUPPORTED: return "combination not implemented";
    case CERB_EWORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

int cerb_corr_output_dims(const cerb_corr_params* p, int32_t* out_channels, int32_t* out_h, int32_t* out_w) {
  Geom g;
  const int rc = build_geom(p, false, g);
  if (rc != CERB_OK) return rc;
  if (out_channels) *out_channels = g.D2;
  if (out_h) *out_h = g.outH;
  if (out_w) *out_w = g.outW;
  return CERB_OK;
}

CERB_API int cerb_warp_corr_forward_variant(const cerb_corr_params* p, const void* x1, const void* x2,
                                            const float* flow, void* out, int variant, cerb_stream_t stream) {
  Geom g;
  const int rc = build_geom(p, flow != nullptr, g);
  if (rc != CERB_OK) return rc;
  if (!x1 || !x2 || !out) return CERB_EINVAL;
  if (variant < CERB_FWD_VARIANT_AUTO || variant > CERB_FWD_VARIANT_TC) return CERB_EINVAL;
  if (variant != CERB_FWD_VARIANT_AUTO && variant != CERB_FWD_VARIANT_GENERIC && !is_fast(g)) return CERB_EUNSUPPORTED;
  const cudaError_t e = launch_warp_corr_forward(g, p->dtype, x1, x2, flow, out, variant, (cudaStream_t)stream);
  if (e == cudaSuccess) count_launches(1);
  return (int)e;
}

int cerb_warp_corr_forward_upflow(const cerb_corr_params* p, const void* x1, const void* x2, const float* flow_coarse,
                                  const int64_t coarse_stride[4], float* flow_up, const int64_t up_stride[4], void* out,
                                  cerb_stream_t stream) {
  Geom g;
  const int rc = build_geom(p, true, g);
  if (rc != CERB_OK) return rc;
  if (!x1 || !x2 || !out || !flow_coarse || !flow_up) return CERB_EINVAL;
  if ((g.H & 1) || (g.W & 1) || g.H < 4 || g.W < 4) return CERB_EINVAL;   // scale_factor = 2 exactly, coarse maps at least 2x2
  if (!is_fast(g) || g.pad != g.md) return CERB_EUNSUPPORTED;             // tiles must cover the whole image
  UpFlow uf;
  uf.coarse = flow_coarse; uf.up = flow_up; uf.Hc = g.H / 2; uf.Wc = g.W / 2;
  bool ok1, ok2;
  static const int64_t contiguous[4] = {0, 0, 0, 0};
  fill_strides(coarse_stride ? coarse_stride : contiguous, uf.cs, 2, uf.Hc, uf.Wc, ok1);
  fill_strides(up_stride ? up_stride : contiguous, uf.us, 2, g.H, g.W, ok2);
  if (!(ok1 && ok2)) return CERB_ESTRIDE;
  const cudaError_t e = launch_warp_corr_forward(g, p->dtype, x1, x2, nullptr, out, CERB_FWD_VARIANT_AUTO,
                                                 (cudaStream_t)stream, &uf);
  if (e == cudaSuccess) count_launches(1);
  return (int)e;
}

int cerb_warp_corr_forward(const cerb_corr_params* p, const void* x1, const void* x2, const float* flow, void* out,
                           cerb_stream_t stream) {
  return cerb_warp_corr_forward_variant(p, x1, x2, flow, out, CERB_FWD_VARIANT_AUTO, stream);
}

size_t cerb_warp_corr_backward_workspace(const cerb_corr_params* p, int has_flow) {
  Geom g;
  if (build_geom(p, has_flow != 0, g) != CERB_OK) return 0;
  if (!has_flow) return 0;
  // the warped second map and the gradient wrt it; 16-bit dtypes add an fp32 accumulator for the warp's splat
  const size_t n = (size_t)g.B * g.C * g.H * g.W, es = dtype_size(p->dtype);
  return 2 * n * es + (es == 2 ? n * sizeof(float) : 0);
}

int cerb_warp_corr_backward(const cerb_corr_params* p, const void* x1, const void* x2, const float* flow,
                            const void* out, const void* grad_out, void* grad_x1, void* grad_x2, float* grad_flow,
                            void* workspace, size_t workspace_bytes, cerb_stream_t stream) {
  Geom g;
  const int rc = build_geom(p, flow != nullptr, g);
  if (rc != CERB_OK) return rc;
  if (!x1 || !x2 || !grad_out || !grad_x1 || !grad_x2) return CERB_EINVAL;
  if (flow != nullptr && grad_flow == nullptr) return CERB_EINVAL;
  if (g.has_act && out == nullptr) return CERB_EINVAL;
  const size_t need = cerb_warp_corr_backward_workspace(p, flow != nullptr);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) return CERB_EWORKSPACE;
  return (int)launch_warp_corr_backward(g, p->dtype, x1, x2, flow, out, grad_out, grad_x1, grad_x2, grad_flow,
                                        workspace, (cudaStream_t)stream);
}

int cerb_flow_warp_forward(const void* image, const float* flow, void* out, int32_t batch, int32_t channels,
                           int32_t height, int32_t width, int32_t dtype, int32_t warp_mode, cerb_stream_t stream) {
  if (!image || !flow || !out || batch < 1 || channels < 1 || height < 2 || width < 2) return CERB_EINVAL;
  if (warp_mode < 0 || warp_mode > 7) return CERB_EINVAL;  // bit 2 (4): experiment switch, un-normalise without FMA
  if ((long long)channels * height * width >= 0x7fffffffLL) return CERB_ESTRIDE;
  return (int)launch_flow_warp_forward(dtype, image, flow, out, batch, channels, height, width, warp_mode,
                                       (cudaStream_t)stream);
}

int cerb_flow_warp_backward(const void* image, const float* flow, const void* grad_out, void* grad_image,
                            float* grad_flow, int32_t batch, int32_t channels, int32_t height, int32_t width,
                            int32_t dtype, int32_t warp_mode, cerb_stream_t stream) {
  if (!image || !flow || !grad_out || !grad_image || !grad_flow || batch < 1 || channels < 1 || height < 2 || width < 2)
    return CERB_EINVAL;
  if (warp_mode < CERB_WARP_TORCH || warp_mode > CERB_WARP_TORCH_CPU) return CERB_EINVAL;
  if ((long long)channels * height * width >= 0x7fffffffLL) return CERB_ESTRIDE;
  return (int)launch_flow_warp_backward(dtype, image, flow, grad_out, grad_image, grad_flow, batch, channels, height,
                                        width, warp_mode, (cudaStream_t)stream);
}

int cerb_grid_sample_forward(const void* input, const void* grid, void* output, int32_t batch, int32_t channels,
                             int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w, int32_t dtype,
                             int32_t interpolation_mode, int32_t padding_mode, int32_t align_corners, int32_t convention,
                             cerb_stream_t stream) {
  if (!input || !grid || !output || batch < 1 || channels < 1 || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1)
    return CERB_EINVAL;
  if (dtype != CERB_F32 && dtype != CERB_F16 && dtype != CERB_BF16) return CERB_EINVAL;
  if (interpolation_mode != CERB_GRID_BILINEAR && interpolation_mode != CERB_GRID_NEAREST) return CERB_EUNSUPPORTED;
  if (padding_mode < CERB_GRID_PAD_ZEROS || padding_mode > CERB_GRID_PAD_REFLECTION) return CERB_EINVAL;
  if (convention != CERB_GRID_CONV_TRT && convention != CERB_GRID_CONV_ATEN) return CERB_EINVAL;
  if ((long long)in_h * in_w >= 0x7fffffffLL) return CERB_ESTRIDE;
  return (int)launch_grid_sampler(dtype, input, grid, output, batch, channels, in_h, in_w, out_h, out_w, interpolation_mode,
                                  padding_mode, align_corners ? 1 : 0, convention, (cudaStream_t)stream);
}

size_t cerb_photometric_workspace(int32_t batch, int32_t height, int32_t width) {
  if (batch < 1 || height < 1 || width < 1) return 0;
  return photometric_workspace_bytes(batch, height, width);
}

static int photometric_args_ok(int32_t batch, int32_t channels, int32_t height, int32_t width, int32_t warp_mode) {
  if (batch < 1 || channels < 1 || height < 4 || width < 4) return CERB_EINVAL;   // ReflectionPad2d(1) + one mirror per axis
  if (warp_mode < CERB_WARP_TORCH || warp_mode > CERB_WARP_TORCH_CPU) return CERB_EINVAL;
  if ((long long)channels * height * width >= 0x7fffffffLL) return CERB_ESTRIDE;
  return CERB_OK;
}

int cerb_photometric_forward(const float* im_orig, const float* im_src, const float* flow, float* loss, void* workspace,
                             size_t workspace_bytes, int32_t batch, int32_t channels, int32_t height, int32_t width,
                             float l1_weight, float ssim_weight, int32_t warp_mode, cerb_stream_t stream) {
  if (!im_orig || !im_src || !flow || !loss || !workspace) return CERB_EINVAL;
  const int rc = photometric_args_ok(batch, channels, height, width, warp_mode);
  if (rc != CERB_OK) return rc;
  if (workspace_bytes < cerb_photometric_workspace(batch, height, width)) return CERB_EWORKSPACE;
  return (int)launch_photometric_forward(im_orig, im_src, flow, loss, workspace, batch, channels, height, width, l1_weight,
                                         ssim_weight, warp_mode, (cudaStream_t)stream);
}

int cerb_photometric_backward(const float* im_orig, const float* im_src, const float* flow, const float* grad_loss,
                              float* grad_flow, int32_t batch, int32_t channels, int32_t height, int32_t width,
                              float l1_weight, float ssim_weight, int32_t warp_mode, cerb_stream_t stream) {
  if (!im_orig || !im_src || !flow || !grad_loss || !grad_flow) return CERB_EINVAL;
  const int rc = photometric_args_ok(batch, channels, height, width, warp_mode);
  if (rc != CERB_OK) return rc;
  return (int)launch_photometric_backward(im_orig, im_src, flow, grad_loss, grad_flow, batch, channels, height, width,
                                          l1_weight, ssim_weight, warp_mode, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- host-buffer end-to-end ---
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t cerb_warp_corr_forward_host_workspace(const cerb_corr_params* p, int has_flow) {
  Geom g;
  if (build_geom(p, has_flow != 0, g) != CERB_OK) return 0;
  const size_t es = dtype_size(p->dtype);
  const size_t in_b = align256((size_t)g.B * g.C * g.H * g.W * es);
  const size_t fl_b = has_flow ? align256((size_t)g.B * 2 * g.H * g.W * 4) : 0;
  const size_t out_b = align256((size_t)g.B * g.D2 * g.outH * g.outW * es);
  return 2 * in_b + fl_b + out_b;
}

int cerb_warp_corr_forward_host(const cerb_corr_params* p, const void* h_x1, const void* h_x2, const float* h_flow,
                                void* h_out, void* dev_workspace, size_t dev_workspace_bytes, cerb_stream_t stream) {
  if (!p) return CERB_EINVAL;
  cerb_corr_params q = *p;  // host buffers are dense
  memset(q.x1_stride, 0, sizeof(q.x1_stride));
  memset(q.x2_stride, 0, sizeof(q.x2_stride));
  memset(q.flow_stride, 0, sizeof(q.flow_stride));
  memset(q.out_stride, 0, sizeof(q.out_stride));
  Geom g;
  const int rc = build_geom(&q, h_flow != nullptr, g);
  if (rc != CERB_OK) return rc;
  if (!h_x1 || !h_x2 || !h_out || !dev_workspace) return CERB_EINVAL;
  if (dev_workspace_bytes < cerb_warp_corr_forward_host_workspace(&q, h_flow != nullptr)) return CERB_EWORKSPACE;
  const size_t es = dtype_size(q.dtype);
  const size_t in_raw = (size_t)g.B * g.C * g.H * g.W * es;
  const size_t fl_raw = (size_t)g.B * 2 * g.H * g.W * 4;
  const size_t out_raw = (size_t)g.B * g.D2 * g.outH * g.outW * es;
  char* base = (char*)dev_workspace;
  void* d_x1 = base;
  void* d_x2 = base + align256(in_raw);
  float* d_flow = h_flow ? (float*)(base + 2 * align256(in_raw)) : nullptr;
  void* d_out = base + 2 * align256(in_raw) + (h_flow ? align256(fl_raw) : 0);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_x1, h_x1, in_raw, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyAsync(d_x2, h_x2, in_raw, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
  if (h_flow && (e = cudaMemcpyAsync(d_flow, h_flow, fl_raw, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
  const int r2 = cerb_warp_corr_forward(&q, d_x1, d_x2, d_flow, d_out, stream);
  if (r2 != CERB_OK) return r2;
  if ((e = cudaMemcpyAsync(h_out, d_out, out_raw, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  return CERB_OK;
}

// Debugging aid, not part of the public ABI: per-CTA clock64() trace of the fast forward kernel.
CERB_API void cerb_debug_set_trace_buffer(void* dev_ptr) { cerb::set_trace_buffer((long long*)dev_ptr); }
CERB_API void cerb_debug_set_trace_iter(int it) { cerb::set_trace_iter(it); }
// dev_ptr: 4 x uint64 on the device, zeroed by the caller: tiles per staging path of the warped forward
// ([1] small raw box, [3] large raw box, [2] direct-gather fallback); NULL switches counting off.
CERB_API void cerb_debug_set_path_counters(void* dev_ptr) { cerb::set_path_counters((unsigned long long*)dev_ptr); }
// which kernel cerb_warp_corr_backward's fast path takes: -1 automatic (default), 0 CUDA cores, 1 tensor cores wherever supported
CERB_API void cerb_debug_set_backward_kernel(int mode) { cerb::set_backward_kernel_mode(mode); }

int cerb_measure_fma_peak(double* tflops, cerb_stream_t stream) {
  if (!tflops) return CERB_EINVAL;
  return (int)cerb::measure_fma_peak(tflops, (cudaStream_t)stream);
}

uint64_t cerb_launch_count(void) { return (uint64_t)g_launches.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------- TensorRT-shaped ABI ------
void cerb_trt_corr_default_fields(cerb_trt_corr_fields* f) {
  if (!f) return;
  f->pad_size = 4; f->kernel_size = 1; f->max_displacement = 4;
  f->stride1 = 1; f->stride2 = 1; f->corr_multiply = 1;
}

size_t cerb_trt_corr_serialization_size(void) { return sizeof(cerb_trt_corr_fields); }

size_t cerb_trt_corr_serialize(const cerb_trt_corr_fields* f, void* buffer) {
  if (!f || !buffer) return 0;
  memcpy(buffer, f, sizeof(*f));  // six raw int32 in declaration order
  return sizeof(*f);
}

int cerb_trt_corr_deserialize(const void* data, size_t length, cerb_trt_corr_fields* f) {
  if (!data || !f || length != sizeof(*f)) return CERB_EINVAL;
  memcpy(f, data, sizeof(*f));
  return CERB_OK;
}

int cerb_trt_corr_output_dims(const cerb_trt_corr_fields* f, const cerb_trt_dims* in0, cerb_trt_dims* out) {
  if (!f || !in0 || !out || in0->nbDims != 4) return CERB_EINVAL;
  const int64_t d[4] = {in0->d[0], in0->d[1], in0->d[2], in0->d[3]};
  cerb_corr_params p;
  int rc = trt_params(f, d, CERB_TRT_FLOAT, &p);
  if (rc != CERB_OK) return rc;
  int32_t oc, oh, ow;
  rc = cerb_corr_output_dims(&p, &oc, &oh, &ow);
  if (rc != CERB_OK) return rc;
  memset(out, 0, sizeof(*out));
  out->nbDims = 4;
  out->d[0] = in0->d[0]; out->d[1] = oc; out->d[2] = oh; out->d[3] = ow;
  return CERB_OK;
}

int cerb_trt_corr_supports_format(int pos, const cerb_trt_tensor_desc* io, int nb_inputs, int nb_outputs) {
  if (!io || nb_inputs != 2 || nb_outputs != 1 || pos < 0 || pos >= 3) return 0;
  bool ok = io[pos].format == CERB_TRT_LINEAR;
  ok = ok && (io[pos].type == CERB_TRT_FLOAT || io[pos].type == CERB_TRT_HALF);
  for (int i = 0; i < io[0].dims.nbDims && i < CERB_TRT_MAX_DIMS; ++i) ok = ok && io[0].dims.d[i] == io[1].dims.d[i];
  if (pos == 1) ok = ok && io[1].type == io[0].type;
  if (pos == 2) ok = ok && io[2].type == io[0].type && io[2].type == io[1].type;
  return ok ? 1 : 0;
}

size_t cerb_trt_corr_workspace_size(const cerb_trt_corr_fields*, const cerb_trt_tensor_desc*, int,
                                    const cerb_trt_tensor_desc*, int) {
  return 0;
}

int cerb_trt_corr_enqueue(const cerb_trt_corr_fields* f, const cerb_trt_tensor_desc* input_desc,
                          const cerb_trt_tensor_desc* output_desc, const void* const* inputs, void* const* outputs,
                          void* /*workspace*/, cerb_stream_t stream) {
  return trt_enqueue_impl(f, CERB_WARP_TRT, NAN, false, input_desc, output_desc, inputs, outputs, stream);
}

int cerb_trt_corr_enqueue_i64(const cerb_trt_corr_fields* f, const cerb_trt_tensor_desc64* input_desc,
                              const cerb_trt_tensor_desc64* output_desc, const void* const* inputs,
                              void* const* outputs, void* /*workspace*/, cerb_stream_t stream) {
  return trt_enqueue_impl(f, CERB_WARP_TRT, NAN, false, input_desc, output_desc, inputs, outputs, stream);
}

int cerb_trt_warp_corr_enqueue(const cerb_trt_corr_fields* f, int32_t warp_mode, float leaky_slope,
                               const cerb_trt_tensor_desc* input_desc, const cerb_trt_tensor_desc* output_desc,
                               const void* const* inputs, void* const* outputs, void* /*workspace*/,
                               cerb_stream_t stream) {
  if (warp_mode < CERB_WARP_TORCH || warp_mode > CERB_WARP_TORCH_CPU) return CERB_EINVAL;
  return trt_enqueue_impl(f, warp_mode, leaky_slope, true, input_desc, output_desc, inputs, outputs, stream);
}

void cerb_trt_warp_corr_default_fields(cerb_trt_warp_corr_fields* f) {
  if (!f) return;
  cerb_trt_corr_default_fields(&f->corr);
  f->warp_mode = CERB_WARP_TRT;
  f->leaky_slope = 0.1f;
}

size_t cerb_trt_warp_corr_serialize(const cerb_trt_warp_corr_fields* f, void* buffer) {
  if (!buffer) return sizeof(cerb_trt_warp_corr_fields);
  if (!f) return 0;
  memcpy(buffer, f, sizeof(*f));   // six int32, int32 warp_mode, float slope
  return sizeof(*f);
}

int cerb_trt_warp_corr_deserialize(const void* data, size_t length, cerb_trt_warp_corr_fields* f) {
  if (!data || !f || length != sizeof(*f)) return CERB_EINVAL;
  memcpy(f, data, sizeof(*f));
  return CERB_OK;
}

void cerb_trt_grid_sampler_default_fields(cerb_trt_grid_sampler_fields* f) {
  if (!f) return;
  f->align_corners = 0; f->interpolation_mode = CERB_GRID_BILINEAR; f->padding_mode = CERB_GRID_PAD_BORDER;   // grid_sampler.cpp:40-42
}

size_t cerb_trt_grid_sampler_serialize(const cerb_trt_grid_sampler_fields* f, void* buffer) {
  const size_t n = 1 + 2 * sizeof(int32_t);   // bool, int, int (grid_sampler.cpp:57-74)
  if (!buffer) return n;
  if (!f) return 0;
  unsigned char* d = (unsigned char*)buffer;
  d[0] = f->align_corners ? 1 : 0;
  memcpy(d + 1, &f->interpolation_mode, 4);
  memcpy(d + 5, &f->padding_mode, 4);
  return n;
}

int cerb_trt_grid_sampler_deserialize(const void* data, size_t length, cerb_trt_grid_sampler_fields* f) {
  if (!data || !f || length != 9) return CERB_EINVAL;
  const unsigned char* d = (const unsigned char*)data;
  f->align_corners = d[0] ? 1 : 0;
  memcpy(&f->interpolation_mode, d + 1, 4);
  memcpy(&f->padding_mode, d + 5, 4);
  return CERB_OK;
}

int cerb_trt_grid_sampler_enqueue(const cerb_trt_grid_sampler_fields* f, const cerb_trt_tensor_desc* in,
                                  const cerb_trt_tensor_desc* out, const void* const* inputs, void* const* outputs,
                                  void* /*workspace*/, cerb_stream_t stream) {
  if (!f || !in || !out || !inputs || !outputs) return CERB_EINVAL;
  if (in[0].dims.nbDims != 4 || in[1].dims.nbDims != 4 || out[0].dims.nbDims != 4 || in[1].dims.d[3] != 2) return CERB_EINVAL;
  if (in[0].type != in[1].type || out[0].type != in[0].type) return CERB_EUNSUPPORTED;
  int dtype;
  if (in[0].type == CERB_TRT_FLOAT) dtype = CERB_F32;
  else if (in[0].type == CERB_TRT_HALF) dtype = CERB_F16;
  else return CERB_EUNSUPPORTED;   // reference throws (grid_sampler.cu:263-266)
  if (out[0].dims.d[0] != in[0].dims.d[0] || out[0].dims.d[1] != in[0].dims.d[1] || out[0].dims.d[2] != in[1].dims.d[1] ||
      out[0].dims.d[3] != in[1].dims.d[2] || in[1].dims.d[0] != in[0].dims.d[0])
    return CERB_ESHAPE;
  return cerb_grid_sample_forward(inputs[0], inputs[1], outputs[0], in[0].dims.d[0], in[0].dims.d[1], in[0].dims.d[2],
                                  in[0].dims.d[3], out[0].dims.d[2], out[0].dims.d[3], dtype, f->interpolation_mode,
                                  f->padding_mode, f->align_corners, CERB_GRID_CONV_TRT, stream);
}

}  // extern "C"
