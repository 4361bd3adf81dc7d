/*
 * cerberus_costvolume.h -- C ABI of the B200-native cost-volume hot path
 * (libcerberus_costvolume.so, sm_100a).
 *
 * This is the drop-in boundary for CerberusNet's correlation + flow-warp + LeakyReLU path.
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference checkout).  Plain pointers and sizes only; no torch / TensorRT types.
 *
 * Conventions (all entry points unless stated otherwise)
 *   - tensors are NCHW; device pointers are borrowed, never freed, never retained;
 *   - the innermost (W) stride must be 1; N/C/H strides are in elements and may be
 *     arbitrary (stride[] all zero = contiguous NCHW), so the activated cost volume can be
 *     written straight into a wider concat buffer;
 *   - work is enqueued on `stream` only: no allocation, no host synchronisation, no private
 *     streams, CUDA-graph capturable (the reference's TensorRT plugin synchronises the host
 *     three times per enqueue, trt_plugins/correlation.cu:105,124-125);
 *   - return value: 0 on success, a positive cudaError_t if the CUDA runtime reported one,
 *     or a negative CERB_E* code for argument errors.  Nothing aborts or throws
 *     (reference: AT_ERROR -> RuntimeError, correlation_cuda.cpp:22-23; NV_CUDA_CHECK ->
 *     abort(), trt_plugins/trt_utils.hpp:7-15);
 *   - thread-safety: entry points may be called concurrently from several host threads / on several devices.  The only
 *     process-wide state is lazily initialised and idempotent (per-device function attributes and SM counts, the driver
 *     entry point of cuTensorMapEncodeTiled, a launch counter updated atomically).  The cerb_debug_* hooks (trace
 *     buffer, path counters) are global switches for single-threaded debugging and are NOT thread-safe.
 */
#ifndef CERBERUS_COSTVOLUME_H_
#define CERBERUS_COSTVOLUME_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CERB_API
#define CERB_API __attribute__((visibility("default")))
#endif

/* cudaStream_t without dragging cuda_runtime.h into C callers (cgo / ctypes / JNI). */
typedef void* cerb_stream_t;

#define CERB_ABI_VERSION 1

/* element types (reference: AT_DISPATCH_FLOATING_TYPES_AND_HALF, correlation_cuda_kernel.cu:269;
 * TensorRT kFLOAT / kHALF, trt_plugins/correlation.cu:109,133).  bf16 is new capability.
 * Accumulation is always fp32. */
enum { CERB_F32 = 0, CERB_F16 = 1, CERB_BF16 = 2 };

/* flow-warp conventions of the second feature map */
enum {
  CERB_WARP_TORCH = 0,    /* nnet_training/loss_functions/UnFlowLoss.py:83-94 as it runs on CUDA (the
                             reference's training path): ATen evaluates `grid / (size-1)` as a
                             multiplication by the rounded reciprocal there */
  CERB_WARP_TRT = 1,      /* runtime/cerberus_net/trt_plugins/grid_sampler.cu:48-59 */
  CERB_WARP_TORCH_CPU = 2 /* same as TORCH but with the true division ATen's CPU kernels perform
                             (1 ulp apart in the grid; lets CPU-generated fixtures be matched) */
};

/* argument errors (negative so they never collide with cudaError_t) */
enum {
  CERB_OK = 0,
  CERB_EINVAL = -1,      /* null pointer / non-positive size / bad enum */
  CERB_ESHAPE = -2,      /* parameters give an empty output (correlation_cuda.cpp:13-14) */
  CERB_ESTRIDE = -3,     /* innermost stride != 1 or 32-bit index overflow */
  CERB_EUNSUPPORTED = -4,/* combination not implemented */
  CERB_EWORKSPACE = -5   /* workspace too small */
};

/* One decoder level's problem description.
 * Mirrors the six ints of `correlation_args` (nnet_models/pwcnet_sfd.py:126-133) that are the
 * reference op's whole configuration surface, plus the fused warp / activation switches. */
typedef struct cerb_corr_params {
  int32_t batch, channels, height, width; /* x1 / x2 dims (N, C, H, W) */
  int32_t pad_size, kernel_size, max_displacement, stride1, stride2;
  int32_t corr_multiply; /* accepted and ignored, like correlation_cuda_kernel.cu:246-247 */
  int32_t dtype;         /* CERB_F32 / CERB_F16 / CERB_BF16: x1, x2, out and their grads */
  int32_t warp_mode;     /* CERB_WARP_*; only read when a flow pointer is given */
  float leaky_slope;     /* LeakyReLU negative slope fused on the output (pwcnet_sfd.py:182);
                            NaN (or negative) = no activation */
  int32_t x2_batch_roll; /* 0 <= roll < batch: item n of x1 / flow / out is paired with item (n + roll) mod batch
                            of x2 (0 = the reference pairing).  With x1 = x2 = features of [image 1; image 2]
                            (batch 2N) and roll = N, one launch computes both flow directions of the reference's
                            consistency=True forward (nnet_models/pwcnet.py:108-113, cerberus.py:131-135) without
                            copying or re-ordering a feature map. */
  int64_t x1_stride[4], x2_stride[4], flow_stride[4], out_stride[4]; /* elements; 0,0,0,0 = contiguous */
} cerb_corr_params;

CERB_API int cerb_abi_version(void);
CERB_API const char* cerb_error_string(int code);

/* Output geometry.  Replaces correlation_cuda.cpp:6-14 and
 * CorrelationPlugin::getOutputDimensions (trt_plugins/correlation.cpp:178-205). */
CERB_API int cerb_corr_output_dims(const cerb_corr_params* p, int32_t* out_channels, int32_t* out_h,
                                   int32_t* out_w);

/* Fused forward of one decoder level:
 *     x2w = flow ? flow_warp(x2, flow) : x2
 *     out = leaky_relu(correlation(x1, x2w), slope)
 * Replaces, in one launch: flow_warp (UnFlowLoss.py:83-94, ~10 kernels + a CPU mesh-grid H2D),
 * correlation_forward_cuda (correlation_cuda.cpp:3-26 -> correlation_cuda_kernel.cu:244-324:
 * 3 memsets + 2 transposes + correlation_forward) and F.leaky_relu_ (pwcnet_sfd.py:182).
 * flow: fp32 (N,2,H,W), channel 0 = x displacement in pixels; NULL = plain correlation.
 * With flow == NULL and leaky_slope = NaN this is exactly torch.ops.cerberus.correlation. */
CERB_API int cerb_warp_corr_forward(const cerb_corr_params* p, const void* x1, const void* x2,
                                    const float* flow, void* out, cerb_stream_t stream);

/* Testing / tuning hook: same as cerb_warp_corr_forward with the kernel variant pinned.
 * variant: 0 auto (what cerb_warp_corr_forward uses), 1 8x32-tile kernel with TMA staging,
 * 2 same without TMA (LDG/STG staging), 3 4x16-tile split-channel kernel with TMA, 4 same
 * without TMA, 5 generic one-thread-per-output kernel, 6 8x16-tile kernel (channels split two ways),
 * 7 tensor-core kernel (tcgen05 / TMEM; fp32 via a 3xTF32 split, fp16 / bf16 via kind::f16; max_displacement 4 or 8: what auto
 * picks from 64 tiles of 8x16).
 * Variants 1-4 and 6 need kernel_size=1, stride1=stride2=1, max_displacement>=4, variant 7 additionally
 * max_displacement 4 or 8 (else CERB_EUNSUPPORTED / cudaErrorNotSupported). */
CERB_API int cerb_warp_corr_forward_variant(const cerb_corr_params* p, const void* x1, const void* x2,
                                            const float* flow, void* out, int variant, cerb_stream_t stream);

/* Same forward with the decoder's flow up-sampling fused in (pwcnet_sfd.py:176):
 *   flow = interpolate(2 * flow_coarse, scale_factor=2, mode='bilinear', align_corners=True)
 * flow_coarse is (B,2,H/2,W/2) fp32 (H, W even), evaluated per sample with ATen's arithmetic; the up-sampled flow
 * the decoder needs afterwards is written to flow_up (B,2,H,W) fp32 -- e.g. the last two channels of the concat
 * buffer.  Strides as in cerb_corr_params (elements, all zero = contiguous).  Needs the fast kernels
 * (kernel_size=1, strides 1, max_displacement >= 4) and pad_size == max_displacement, else CERB_EUNSUPPORTED. */
CERB_API int cerb_warp_corr_forward_upflow(const cerb_corr_params* p, const void* x1, const void* x2,
                                           const float* flow_coarse, const int64_t coarse_stride[4], float* flow_up,
                                           const int64_t up_stride[4], void* out, cerb_stream_t stream);

/* Bytes of scratch cerb_warp_corr_backward needs: 0 without a flow; with a flow the re-materialised
 * warped map and the gradient with respect to it, 2*B*C*H*W elements of the tensor dtype (plus B*C*H*W
 * floats for 16-bit dtypes: the warp's splat is accumulated in fp32). */
CERB_API size_t cerb_warp_corr_backward_workspace(const cerb_corr_params* p, int has_flow);

/* Backward of cerb_warp_corr_forward.
 * Replaces correlation_backward_cuda (correlation_cuda.cpp:28-43 -> .cu:326-429: 4 memsets,
 * 2 transposes, 2*B launches), LeakyReluBackward and ATen GridSampler2DBackward + the
 * norm_grid / mesh_grid backward.
 *   out       : the forward result (activated); only its sign is read.  May be NULL when
 *               leaky_slope is NaN.
 *   grad_out  : dL/d(out), same layout as out (contiguous or out_stride).
 *   grad_x1/2 : written (not accumulated), contiguous NCHW, dtype of x1.
 *   grad_flow : fp32 (N,2,H,W) contiguous; required iff flow != NULL.
 * For kernel_size > 1 or stride1 > 1 the reference kernels index out of range
 * (SURVEY.md section 5 iv); here the gradient is defined as the exact adjoint of the forward. */
CERB_API int cerb_warp_corr_backward(const cerb_corr_params* p, const void* x1, const void* x2,
                                     const float* flow, const void* out, const void* grad_out,
                                     void* grad_x1, void* grad_x2, float* grad_flow, void* workspace,
                                     size_t workspace_bytes, cerb_stream_t stream);

/* Stand-alone flow warp (bilinear, border), the a-2 surface on its own.
 * Replaces flow_warp / mesh_grid / norm_grid (UnFlowLoss.py:11-32,83-94) and, with
 * CERB_WARP_TRT, GridSamplerPlugin::enqueue (trt_plugins/grid_sampler.cu:238-271) for the
 * bilinear/border case.  image/out: (N,C,H,W) of `dtype`, contiguous; flow fp32 (N,2,H,W). */
CERB_API int cerb_flow_warp_forward(const void* image, const float* flow, void* out, int32_t batch,
                                    int32_t channels, int32_t height, int32_t width, int32_t dtype,
                                    int32_t warp_mode, cerb_stream_t stream);

/* grad_image is written (zero-filled first, then splatted with atomics); grad_flow written. */
CERB_API int cerb_flow_warp_backward(const void* image, const float* flow, const void* grad_out,
                                     void* grad_image, float* grad_flow, int32_t batch,
                                     int32_t channels, int32_t height, int32_t width, int32_t dtype,
                                     int32_t warp_mode, cerb_stream_t stream);

/* Stand-alone grid sampler with a GRID input (normalised coordinates), every mode of the reference's plugin.
 * Replaces GridSamplerPlugin::enqueue / grid_sampler_kernel (runtime/cerberus_net/trt_plugins/grid_sampler.cu:146-271),
 * the node the ONNX export emits for F.grid_sample (utilities/onnx_export.py:25-28).  The mode enumerations are the
 * reference's (grid_sampler.hpp:14-15), i.e. PyTorch's.  `convention` picks the un-normalise / rounding rule:
 * CERB_GRID_CONV_TRT = the plugin's own ((g+1)*(size-1))/2 for align_corners=0 and roundf for `nearest`
 * (grid_sampler.cu:55-58,219-220); CERB_GRID_CONV_ATEN = what the same graph computes in PyTorch
 * (((g+1)*size-1)/2, round-half-to-even).  input (N,C,H,W), grid (N,out_h,out_w,2) x then y, output
 * (N,C,out_h,out_w), all contiguous and of `dtype`; coordinates are evaluated in fp32. */
enum { CERB_GRID_BILINEAR = 0, CERB_GRID_NEAREST = 1 };
enum { CERB_GRID_PAD_ZEROS = 0, CERB_GRID_PAD_BORDER = 1, CERB_GRID_PAD_REFLECTION = 2 };
enum { CERB_GRID_CONV_TRT = 0, CERB_GRID_CONV_ATEN = 1 };
CERB_API int cerb_grid_sample_forward(const void* input, const void* grid, void* output, int32_t batch,
                                      int32_t channels, int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w,
                                      int32_t dtype, int32_t interpolation_mode, int32_t padding_mode,
                                      int32_t align_corners, int32_t convention, cerb_stream_t stream);

/* Photometric term of unFlowLoss, fused: flow-warp + L1 + SSIM(3x3, reflection pad) + mean.
 * Replaces, per scale and direction: flow_warp (UnFlowLoss.py:83-94,279-283), unFlowLoss.loss_photometric with the
 * all-ones occlusion mask the reference runs with (UnFlowLoss.py:225-241,285-297), SSIM (loss_functions.py:47-78) and
 * their autograd with respect to the flow.
 *   loss = mean( l1_weight * |im_orig - rec| + ssim_weight * SSIM(rec, im_orig) ),  rec = flow_warp(im_src, flow)
 * im_orig / im_src (N,C,H,W) fp32 contiguous, flow (N,2,H,W) fp32 contiguous, H, W >= 4.  `loss` is ONE float on the
 * device; `workspace` holds cerb_photometric_workspace() bytes (per-tile partial sums, added in a fixed order: the
 * result is bit-reproducible).  Backward: grad_flow (N,2,H,W) = d loss / d flow * grad_loss[0] (grad_loss: one float
 * on the device); the images are inputs of the loss and get no gradient. */
CERB_API size_t cerb_photometric_workspace(int32_t batch, int32_t height, int32_t width);
CERB_API int cerb_photometric_forward(const float* im_orig, const float* im_src, const float* flow, float* loss,
                                      void* workspace, size_t workspace_bytes, int32_t batch, int32_t channels,
                                      int32_t height, int32_t width, float l1_weight, float ssim_weight,
                                      int32_t warp_mode, cerb_stream_t stream);
CERB_API int cerb_photometric_backward(const float* im_orig, const float* im_src, const float* flow,
                                       const float* grad_loss, float* grad_flow, int32_t batch, int32_t channels,
                                       int32_t height, int32_t width, float l1_weight, float ssim_weight,
                                       int32_t warp_mode, cerb_stream_t stream);

/* Same as cerb_warp_corr_forward but with HOST buffers (pinned or pageable): copies x1, x2
 * (and flow) host->device into `dev_workspace`, runs the fused kernel and copies the result
 * back, all asynchronously on `stream`.  This is the end-to-end call bench.py times.
 * dev_workspace must hold cerb_warp_corr_forward_host_workspace(p, has_flow) bytes. */
CERB_API size_t cerb_warp_corr_forward_host_workspace(const cerb_corr_params* p, int has_flow);
CERB_API int cerb_warp_corr_forward_host(const cerb_corr_params* p, const void* h_x1, const void* h_x2,
                                         const float* h_flow, void* h_out, void* dev_workspace,
                                         size_t dev_workspace_bytes, cerb_stream_t stream);

/* Measures the fp32 FMA throughput (TFLOP/s) of the current device with a dependency-free packed-FMA loop on every
 * SM: the compute roof bench.py reports next to the HBM roof ("the harness must measure one").
 * Synchronises `stream`; ~20 ms. */
CERB_API int cerb_measure_fma_peak(double* tflops, cerb_stream_t stream);

/* Number of SMs / kernel launches issued so far by this library in this process (the
 * `gpu_launches` claim in bench.py). */
CERB_API uint64_t cerb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CERBERUS_COSTVOLUME_H_ */
