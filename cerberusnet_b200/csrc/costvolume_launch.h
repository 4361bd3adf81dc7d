// costvolume_launch.h -- internal seam between the C ABI (costvolume_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>

#include "costvolume_common.cuh"

// Forward kernel variants; AUTO is what the product uses, the others exist so tests and the
// bench can pin a path (cerb_warp_corr_forward_variant).
enum {
  CERB_FWD_VARIANT_AUTO = 0,
  CERB_FWD_VARIANT_FAST = 1,        // 8x32 tiles, TMA in/out when alignment allows
  CERB_FWD_VARIANT_FAST_NOTMA = 2,  // 8x32 tiles, LDG/STG staging
  CERB_FWD_VARIANT_SMALL = 3,       // 4x16 tiles, channels split 4-way inside the CTA
  CERB_FWD_VARIANT_SMALL_NOTMA = 4,
  CERB_FWD_VARIANT_GENERIC = 5,     // one thread per output element, any parameters
  CERB_FWD_VARIANT_MID = 6,         // 8x16 tiles, channels split 2-way inside the CTA
  CERB_FWD_VARIANT_TC = 7           // tensor cores (tcgen05 / TMEM): 8x16 tiles as 128 x 384 Gram tiles, 3xTF32 for fp32
};

namespace cerb {

// Fused flow up-sampling: the flow of the next-coarser level, (B,2,H/2,W/2), is doubled and
// bilinearly up-sampled (align_corners=True) inside the warp prologue; the up-sampled flow is also written out.
struct UpFlow {
  const float* coarse;
  long long cs[3];   // N, C, H strides of the coarse flow (elements)
  int Hc, Wc;
  float* up;         // (B,2,H,W) destination, may be a channel slice of a wider tensor
  long long us[3];
};

cudaError_t launch_warp_corr_forward(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow,
                                     void* out, int variant, cudaStream_t stream, const UpFlow* upflow = nullptr);

// tensor-core forward (costvolume_fwd_tc.cu)
bool tc_forward_supported(const Geom& g, int dtype, const UpFlow* upflow);
cudaError_t launch_warp_corr_forward_tc(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow, void* out,
                                        cudaStream_t stream, const UpFlow* upflow = nullptr);

// workspace: fp32 [B,C,H,W] accumulation buffer for grad_x2 when dtype is 16-bit and flow != NULL
cudaError_t launch_warp_corr_backward(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow,
                                      const void* out, const void* gout, void* gx1, void* gx2, float* gflow,
                                      void* workspace, cudaStream_t stream);

cudaError_t launch_flow_warp_forward(int dtype, const void* image, const float* flow, void* out, int B, int C, int H,
                                     int W, int mode, cudaStream_t stream);
cudaError_t launch_flow_warp_backward(int dtype, const void* image, const float* flow, const void* gout, void* gimage,
                                      float* gflow, int B, int C, int H, int W, int mode, cudaStream_t stream);

cudaError_t launch_grid_sampler(int dtype, const void* input, const void* grid, void* out, int N, int C, int H, int W, int oH,
                                int oW, int interp, int padding, int align, int conv, cudaStream_t stream);

size_t photometric_workspace_bytes(int B, int H, int W);
cudaError_t launch_photometric_forward(const float* orig, const float* src, const float* flow, float* loss, void* workspace, int B,
                                       int C, int H, int W, float l1_w, float ssim_w, int mode, cudaStream_t stream);
cudaError_t launch_photometric_backward(const float* orig, const float* src, const float* flow, const float* gloss, float* gflow,
                                        int B, int C, int H, int W, float l1_w, float ssim_w, int mode, cudaStream_t stream);

void set_backward_kernel_mode(int mode);   // -1 automatic, 0 CUDA-core kernels, 1 tensor-core kernel wherever supported
void count_launches(int n);
void set_trace_buffer(long long* p);
void set_trace_iter(int it);
void set_path_counters(unsigned long long* p);

// fp32 FMA throughput of the current device (TFLOP/s), measured with a dependency-free FFMA2 loop over every SM
cudaError_t measure_fma_peak(double* tflops, cudaStream_t stream);

}  // namespace cerb
