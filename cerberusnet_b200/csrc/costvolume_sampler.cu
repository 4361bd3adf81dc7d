// costvolume_sampler.cu -- stand-alone grid sampler (grid input, every mode) for sm_100a.
//
// Replaces (reference paths relative to the reference checkout):
//   GridSamplerPlugin::enqueue / grid_sampler_kernel   runtime/cerberus_net/trt_plugins/grid_sampler.cu:146-271
//   (the node the ONNX export emits for F.grid_sample, nnet_training/utilities/onnx_export.py:25-28)
// Two un-normalise conventions, because the reference has two: its TensorRT plugin maps a
// non-align-corners coordinate with ((g+1)*(size-1))/2 (grid_sampler.cu:55-58), ATen -- what the same
// graph computes in PyTorch -- with ((g+1)*size-1)/2.  Same for `nearest`: roundf in the plugin
// (:219-220), round-half-to-even in ATen.  One thread per output pixel, channels in batches of four
// (16 tap loads in flight); coordinates are always evaluated in fp32 (the plugin's kHALF path evaluates
// them in half precision; this one only stores in half).
#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

// grid_sampler.cu:48-59 (TRT) / ATen grid_sampler_unnormalize
__device__ __forceinline__ float gs_unnormalize(float c, int size, bool align, int conv) {
  if (align) return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(size - 1));
  if (conv == CERB_GRID_CONV_TRT) return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), (float)(size - 1)), 0.5f);
  return __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(c, 1.f), (float)size), -1.f), 0.5f);
}

// grid_sampler.cu:73-86: reflect until inside [low, high], bounds given as twice their value
__device__ __forceinline__ float gs_reflect(float in, int twice_low, int twice_high) {
  if (twice_low == twice_high) return 0.f;
  const float mn = (float)twice_low * 0.5f;
  const float span = (float)(twice_high - twice_low) * 0.5f;
  in = fabsf(in - mn);
  const float extra = fmodf(in, span);
  const int flips = (int)floorf(in / span);
  return (flips & 1) == 0 ? extra + mn : span - extra + mn;
}

// grid_sampler.cu:126-141
__device__ __forceinline__ float gs_source_index(float c, int size, int padding, bool align, int conv) {
  c = gs_unnormalize(c, size, align, conv);
  if (padding == CERB_GRID_PAD_BORDER) {
    c = fminf((float)(size - 1), fmaxf(c, 0.f));
  } else if (padding == CERB_GRID_PAD_REFLECTION) {
    c = align ? gs_reflect(c, 0, 2 * (size - 1)) : gs_reflect(c, -1, 2 * size - 1);
    c = fminf((float)(size - 1), fmaxf(c, 0.f));
  }
  // safe_downgrade_to_int_range (:113-121): anything that cannot be an int index is sent out of bounds
  if (c > 2147483646.f || c < -2147483648.f || !isfinite(c)) c = -100.f;
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256) grid_sampler_kernel(const T* __restrict__ input, const T* __restrict__ grid,
                                                           T* __restrict__ out, int N, int C, int H, int W, int oH, int oW,
                                                           int interp, int padding, int align, int conv) {
  const long long oplane = (long long)oH * oW, iplane = (long long)H * W;
  const long long total = (long long)N * oplane;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / oplane);
    const long long rem = idx - (long long)n * oplane;
    const T* gp = grid + idx * 2;                       // (N, oH, oW, 2): x then y (grid_sampler.cu:172-178)
    const float ix = gs_source_index(to_f32<T>(gp[0]), W, padding, align != 0, conv);
    const float iy = gs_source_index(to_f32<T>(gp[1]), H, padding, align != 0, conv);
    const T* ip = input + (long long)n * C * iplane;
    T* op = out + (long long)n * C * oplane + rem;
    int off[4];
    float w[4];
    int ntap;
    if (interp == CERB_GRID_BILINEAR) {                 // grid_sampler.cu:181-217
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
      const float wx1 = __fsub_rn((float)x1, ix), wx0 = __fsub_rn(ix, (float)x0);
      const float wy1 = __fsub_rn((float)y1, iy), wy0 = __fsub_rn(iy, (float)y0);
      const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
      const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
      const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
      off[0] = cy0 * W + cx0; off[1] = cy0 * W + cx1; off[2] = cy1 * W + cx0; off[3] = cy1 * W + cx1;
      w[0] = (vx0 && vy0) ? __fmul_rn(wx1, wy1) : 0.f;   // nw
      w[1] = (vx1 && vy0) ? __fmul_rn(wx0, wy1) : 0.f;   // ne
      w[2] = (vx0 && vy1) ? __fmul_rn(wx1, wy0) : 0.f;   // sw
      w[3] = (vx1 && vy1) ? __fmul_rn(wx0, wy0) : 0.f;   // se
      ntap = 4;
    } else {                                            // nearest, grid_sampler.cu:218-233
      const float rx = conv == CERB_GRID_CONV_TRT ? roundf(ix) : nearbyintf(ix);
      const float ry = conv == CERB_GRID_CONV_TRT ? roundf(iy) : nearbyintf(iy);
      const int xn = (int)rx, yn = (int)ry;
      const bool ok = xn >= 0 && xn < W && yn >= 0 && yn < H;
      off[0] = ok ? yn * W + xn : 0;
      w[0] = ok ? 1.f : 0.f;
      off[1] = off[2] = off[3] = 0; w[1] = w[2] = w[3] = 0.f;
      ntap = 1;
    }
    for (int c0 = 0; c0 < C; c0 += 4) {
      float v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const T* p = ip + (long long)min(c0 + k, C - 1) * iplane;
        v[k][0] = ldg_f32(p + off[0]);
        if (ntap == 4) { v[k][1] = ldg_f32(p + off[1]); v[k][2] = ldg_f32(p + off[2]); v[k][3] = ldg_f32(p + off[3]); }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c0 + k < C) {
          float r;
          if (ntap == 4) {
            // out = 0; out += v*w per in-bounds tap in the order nw, ne, sw, se (:198-213); a tap outside the image
            // is skipped there -- here it has weight 0 and a clamped address, and NaN/Inf at that address must not leak
            r = 0.f;
            if (w[0] != 0.f) r = __fmaf_rn(v[k][0], w[0], r);
            if (w[1] != 0.f) r = __fmaf_rn(v[k][1], w[1], r);
            if (w[2] != 0.f) r = __fmaf_rn(v[k][2], w[2], r);
            if (w[3] != 0.f) r = __fmaf_rn(v[k][3], w[3], r);
          } else {
            r = w[0] != 0.f ? v[k][0] : 0.f;
          }
          op[(long long)(c0 + k) * oplane] = from_f32<T>(r);
        }
      }
    }
  }
}

template <typename T>
static cudaError_t grid_sampler_t(const void* input, const void* grid, void* out, int N, int C, int H, int W, int oH, int oW,
                                  int interp, int padding, int align, int conv, cudaStream_t stream) {
  const long long total = (long long)N * oH * oW;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (blocks < 1) blocks = 1;
  grid_sampler_kernel<T><<<(int)blocks, 256, 0, stream>>>((const T*)input, (const T*)grid, (T*)out, N, C, H, W, oH, oW, interp,
                                                          padding, align, conv);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_grid_sampler(int dtype, const void* input, const void* grid, void* out, int N, int C, int H, int W, int oH,
                                int oW, int interp, int padding, int align, int conv, cudaStream_t stream) {
  switch (dtype) {
    case CERB_F32: return grid_sampler_t<float>(input, grid, out, N, C, H, W, oH, oW, interp, padding, align, conv, stream);
    case CERB_F16: return grid_sampler_t<__half>(input, grid, out, N, C, H, W, oH, oW, interp, padding, align, conv, stream);
    case CERB_BF16: return grid_sampler_t<__nv_bfloat16>(input, grid, out, N, C, H, W, oH, oW, interp, padding, align, conv, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cerb
