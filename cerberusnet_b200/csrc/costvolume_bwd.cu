// costvolume_bwd.cu -- backward of the fused level op, and the stand-alone flow warp, sm_100a.
//
// Replaces (reference paths relative to the reference checkout):
//   correlation_backward_cuda     nnet_training/correlation_package/correlation_cuda.cpp:28-43
//     correlation_backward_input1/2   correlation_cuda_kernel.cu:97-172, 174-242, launch loop :326-429
//   LeakyReluBackward             autograd of pwcnet_sfd.py:182
//   GridSampler2DBackward + norm_grid / mesh_grid backward   UnFlowLoss.py:22-32,83-94
//
// All batch items in one launch (the reference launches 2*B kernels, .cu:386,407), no padded NHWC
// scratch, no memsets except grad_x2 when it is splatted through the warp.
//
// Fast path (k=1, s1=s2=1, md=4).  Both gradients are the same banded contraction
//     g[c, p] = 1/C * sum_d  G[d, p] * S[c, p (+/-) d]
// so one kernel template serves both: a 256-thread CTA owns an 8x32 pixel tile, every thread
// keeps its pixel's 81 cost-volume gradients (LeakyReLU mask applied) in registers and walks the
// channels, reading the 9x9 window of S from a shared-memory halo tile that is staged per
// 8-channel chunk (for grad_x1 that tile is the *re-warped* x2, gathered on the fly).
// For grad_x2 the per-channel result is pushed straight through the bilinear warp backward:
// 4 atomics into grad_x2 and the flow gradient accumulated over channels in registers.
#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

constexpr int kMDb = 4;
constexpr int kDb = 9;
constexpr int kD2b = 81;
constexpr int BT_Y = 8, BT_X = 32;
constexpr int BH_Y = BT_Y + 2 * kMDb, BH_X = BT_X + 2 * kMDb;  // 16 x 40 halo
constexpr int BH_XS = BH_X + 1;
constexpr int kCB = 8;

template <typename T> __device__ __forceinline__ void atomic_add_t(T* p, float v);
template <> __device__ __forceinline__ void atomic_add_t<float>(float* p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_add_t<__half>(__half* p, float v) { atomicAdd(p, __float2half_rn(v)); }
template <> __device__ __forceinline__ void atomic_add_t<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

struct BwdArgs {
  Geom g;
  const void* x1;
  const void* x2;
  const float* flow;
  const void* out;   // activated forward output (sign only), may be null
  const void* gout;
  void* gx1;
  void* gx2;
  float* gflow;
  int off;           // md - pad
  int tiles_x, tiles_y;
};

// WHICH == 0: grad_x1 (S = warped x2, window p + d).  WHICH == 1: grad_x2 (S = x1, window q - d).
template <typename T, int WHICH>
__global__ void __launch_bounds__(256) corr_bwd_fast_kernel(const BwdArgs a) {
  __shared__ float halo[kCB * BH_Y * BH_XS];
  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const int ty = tid >> 5, tx = tid & 31;
  const int n = blockIdx.z;
  const int iy0 = blockIdx.y * BT_Y, ix0 = blockIdx.x * BT_X;  // input-frame tile origin
  const int iy = iy0 + ty, ix = ix0 + tx;
  const bool pix_ok = iy < g.H && ix < g.W;
  const T* __restrict__ x1 = (const T*)a.x1 + (long long)n * g.x1s[0];
  const T* __restrict__ x2 = (const T*)a.x2 + (long long)n * g.x2s[0];
  const T* __restrict__ gout = (const T*)a.gout + (long long)n * g.os[0];
  const T* __restrict__ outp = a.out ? (const T*)a.out + (long long)n * g.os[0] : nullptr;
  const bool warped = a.flow != nullptr;

  // ---- the pixel's 81 cost-volume gradients, activation mask applied
  float G[kD2b];
#pragma unroll
  for (int dy = 0; dy < kDb; ++dy) {
#pragma unroll
    for (int dx = 0; dx < kDb; ++dx) {
      // WHICH 0: output pixel of this input pixel.  WHICH 1: output pixel of source p = q - d.
      const int py = (WHICH == 0) ? iy : iy - (dy - kMDb);
      const int px = (WHICH == 0) ? ix : ix - (dx - kMDb);
      const int oy = py - a.off, ox = px - a.off;
      float v = 0.f;
      if (pix_ok && py >= 0 && py < g.H && px >= 0 && px < g.W && oy >= 0 && oy < g.outH && ox >= 0 && ox < g.outW) {
        const long long o = (long long)(dy * kDb + dx) * g.os[1] + (long long)oy * g.os[2] + ox;
        v = ldg_f32(gout + o);
        if (g.has_act && outp != nullptr && !(ldg_f32(outp + o) > 0.f)) v *= g.slope;
      }
      G[dy * kDb + dx] = v;
    }
  }

  // ---- staging plan: halo positions handled by this thread
  constexpr int NPOS = BH_Y * BH_X;
  constexpr int PPT = (NPOS + 255) / 256;
  Taps taps[PPT];
  int sdst[PPT];
  unsigned valid_mask = 0;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = tid + j * 256;
    sdst[j] = -1;
    if (i < NPOS) {
      const int hy = i / BH_X, hx = i - hy * BH_X;
      sdst[j] = hy * BH_XS + hx;
      const int qy = iy0 - kMDb + hy, qx = ix0 - kMDb + hx;
      if (qy >= 0 && qy < g.H && qx >= 0 && qx < g.W) {
        valid_mask |= 1u << j;
        if (WHICH == 0 && warped) {
          const float* fp = a.flow + (long long)n * g.fls[0] + (long long)qy * g.fls[2] + qx;
          bool in_x, in_y;
          const float sx = sample_pos(qx, __ldg(fp), g.W, g.warp_mode, in_x);
          const float sy = sample_pos(qy, __ldg(fp + g.fls[1]), g.H, g.warp_mode, in_y);
          taps[j] = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
        } else {
          const long long hs = (WHICH == 0) ? g.x2s[2] : g.x1s[2];
          const int o = (int)(qy * hs) + qx;
          taps[j].off[0] = taps[j].off[1] = taps[j].off[2] = taps[j].off[3] = o;
          taps[j].w[0] = 1.f; taps[j].w[1] = taps[j].w[2] = taps[j].w[3] = 0.f;
        }
      }
    }
  }

  // ---- WHICH 1 + warp: this pixel's own bilinear taps for the splat and the flow gradient
  Taps mytap;
  bool in_x = false, in_y = false;
  float gix = 0.f, giy = 0.f;
  float wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f;
  if (WHICH == 1 && warped && pix_ok) {
    const float* fp = a.flow + (long long)n * g.fls[0] + (long long)iy * g.fls[2] + ix;
    const float sx = sample_pos(ix, __ldg(fp), g.W, g.warp_mode, in_x);
    const float sy = sample_pos(iy, __ldg(fp + g.fls[1]), g.H, g.warp_mode, in_y);
    mytap = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
    const float fx = floorf(sx), fy = floorf(sy);
    wx1 = fx + 1.f - sx; wx0 = sx - fx;
    wy1 = fy + 1.f - sy; wy0 = sy - fy;
  }

  const float inv_c = 1.0f / (float)g.C;
  const T* __restrict__ src = (WHICH == 0) ? x2 : x1;
  const long long src_cs = (WHICH == 0) ? g.x2s[1] : g.x1s[1];

  for (int c0 = 0; c0 < g.C; c0 += kCB) {
    __syncthreads();  // previous chunk fully consumed
#pragma unroll 2
    for (int c = 0; c < kCB; ++c) {
      const int ch = c0 + c;
      const T* plane = src + (long long)ch * src_cs;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        if (sdst[j] < 0) continue;
        float v = 0.f;
        if (ch < g.C && ((valid_mask >> j) & 1u)) {
          if (WHICH == 0 && warped)
            v = blend(ldg_f32(plane + taps[j].off[0]), ldg_f32(plane + taps[j].off[1]), ldg_f32(plane + taps[j].off[2]),
                      ldg_f32(plane + taps[j].off[3]), taps[j]);
          else
            v = ldg_f32(plane + taps[j].off[0]);
        }
        halo[c * (BH_Y * BH_XS) + sdst[j]] = v;
      }
    }
    __syncthreads();
    const int cmax = (g.C - c0) < kCB ? (g.C - c0) : kCB;
    for (int c = 0; c < cmax; ++c) {
      const float* hp = halo + c * (BH_Y * BH_XS);
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < kDb; ++dy) {
#pragma unroll
        for (int dx = 0; dx < kDb; ++dx) {
          // halo origin is (iy0-4, ix0-4).  WHICH 0 reads p + d; WHICH 1 reads q - d.
          const int hy = (WHICH == 0) ? ty + dy : ty + 2 * kMDb - dy;
          const int hx = (WHICH == 0) ? tx + dx : tx + 2 * kMDb - dx;
          s = fmaf(G[dy * kDb + dx], hp[hy * BH_XS + hx], s);
        }
      }
      s *= inv_c;
      if (!pix_ok) continue;
      const int ch = c0 + c;
      if (WHICH == 0) {
        ((T*)a.gx1)[(long long)n * g.C * g.H * g.W + ((long long)ch * g.H + iy) * g.W + ix] = from_f32<T>(s);
      } else if (!warped) {
        ((T*)a.gx2)[(long long)n * g.C * g.H * g.W + ((long long)ch * g.H + iy) * g.W + ix] = from_f32<T>(s);
      } else {
        // grad wrt the warped map -> splat into grad_x2, accumulate d/d(position)
        T* gp = (T*)a.gx2 + (long long)n * g.C * g.H * g.W + (long long)ch * g.H * g.W;
        const T* xp = x2 + (long long)ch * g.x2s[1];
        // grad_x2 is contiguous: rebuild tap offsets with the contiguous row stride
        const int sx0 = mytap.off[0] % (int)g.x2s[2], sy0 = mytap.off[0] / (int)g.x2s[2];
        const int sx1 = mytap.off[3] % (int)g.x2s[2], sy1 = mytap.off[3] / (int)g.x2s[2];
        if (mytap.w[0] != 0.f) atomic_add_t<T>(gp + sy0 * g.W + sx0, s * mytap.w[0]);
        if (mytap.w[1] != 0.f) atomic_add_t<T>(gp + sy0 * g.W + sx1, s * mytap.w[1]);
        if (mytap.w[2] != 0.f) atomic_add_t<T>(gp + sy1 * g.W + sx0, s * mytap.w[2]);
        if (mytap.w[3] != 0.f) atomic_add_t<T>(gp + sy1 * g.W + sx1, s * mytap.w[3]);
        const bool bx1 = sx1 != sx0, by1 = sy1 != sy0;  // far taps inside the image
        const float v_nw = ldg_f32(xp + mytap.off[0]);
        const float v_ne = bx1 ? ldg_f32(xp + mytap.off[1]) : 0.f;
        const float v_sw = by1 ? ldg_f32(xp + mytap.off[2]) : 0.f;
        const float v_se = (bx1 && by1) ? ldg_f32(xp + mytap.off[3]) : 0.f;
        gix += s * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
        giy += s * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
      }
    }
  }
  if (WHICH == 1 && warped && pix_ok) {
    float* gf = a.gflow + (long long)n * 2 * g.H * g.W + (long long)iy * g.W + ix;
    gf[0] = in_x ? gix * pos_scale(g.W, g.warp_mode) : 0.f;
    gf[(long long)g.H * g.W] = in_y ? giy * pos_scale(g.H, g.warp_mode) : 0.f;
  }
}

// ------------------------------------------------------------------ generic backward ------
// Exact adjoint of the generic forward for any parameters.  One thread per (n, c, y, x) input
// element, gathering over every (output pixel, displacement, kernel tap) that touched it.
// grad wrt x2 here means grad wrt the *second correlation input* (the warped map when a flow is
// given); the warp backward kernel below finishes the job in that case.
template <typename T>
__global__ void __launch_bounds__(256) corr_bwd_generic_kernel(const Geom g, const T* __restrict__ x1,
                                                               const T* __restrict__ second, long long sec_ns,
                                                               long long sec_cs, long long sec_hs,
                                                               const T* __restrict__ gout, const T* __restrict__ outp,
                                                               T* __restrict__ gx1, T* __restrict__ gsecond) {
  const long long total = (long long)g.B * g.C * g.H * g.W;
  const float inv = 1.0f / (float)(g.k * g.k * g.C);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % g.W);
    long long t = idx / g.W;
    const int y = (int)(t % g.H);
    t /= g.H;
    const int c = (int)(t % g.C);
    const int n = (int)(t / g.C);
    const T* x1p = x1 + (long long)n * g.x1s[0] + (long long)c * g.x1s[1];
    const T* sp = second + (long long)n * sec_ns + (long long)c * sec_cs;
    const T* gop = gout + (long long)n * g.os[0];
    const T* op = outp ? outp + (long long)n * g.os[0] : nullptr;
    float a1 = 0.f, a2 = 0.f;
    for (int j = -g.kr; j <= g.kr; ++j) {
      for (int i = -g.kr; i <= g.kr; ++i) {
        for (int tj = -g.r; tj <= g.r; ++tj) {
          for (int ti = -g.r; ti <= g.r; ++ti) {
            const int tc = (tj + g.r) * g.D + (ti + g.r);
            // (1) this element as the x1 tap: y = by*s1 + md + j - pad
            {
              const int ny = y + g.pad - g.md - j, nx = x + g.pad - g.md - i;
              if (ny >= 0 && nx >= 0 && ny % g.s1 == 0 && nx % g.s1 == 0) {
                const int by = ny / g.s1, bx = nx / g.s1;
                const int yb = y + tj * g.s2, xb = x + ti * g.s2;
                if (by < g.outH && bx < g.outW && yb >= 0 && yb < g.H && xb >= 0 && xb < g.W) {
                  const long long o = (long long)tc * g.os[1] + (long long)by * g.os[2] + bx;
                  float gv = ldg_f32(gop + o);
                  if (g.has_act && op && !(ldg_f32(op + o) > 0.f)) gv *= g.slope;
                  a1 = fmaf(gv, ldg_f32(sp + (long long)yb * sec_hs + xb), a1);
                }
              }
            }
            // (2) this element as the second-input tap: y = by*s1 + md + j - pad + tj*s2
            {
              const int ya = y - tj * g.s2, xa = x - ti * g.s2;
              const int ny = ya + g.pad - g.md - j, nx = xa + g.pad - g.md - i;
              if (ya >= 0 && ya < g.H && xa >= 0 && xa < g.W && ny >= 0 && nx >= 0 && ny % g.s1 == 0 &&
                  nx % g.s1 == 0) {
                const int by = ny / g.s1, bx = nx / g.s1;
                if (by < g.outH && bx < g.outW) {
                  const long long o = (long long)tc * g.os[1] + (long long)by * g.os[2] + bx;
                  float gv = ldg_f32(gop + o);
                  if (g.has_act && op && !(ldg_f32(op + o) > 0.f)) gv *= g.slope;
                  a2 = fmaf(gv, ldg_f32(x1p + (long long)ya * g.x1s[2] + xa), a2);
                }
              }
            }
          }
        }
      }
    }
    gx1[idx] = from_f32<T>(a1 * inv);
    gsecond[idx] = from_f32<T>(a2 * inv);
  }
}

// ------------------------------------------------------------------ stand-alone warp ------
// flow_warp forward: one thread per (n, y, x), looping channels (coalesced along x).
template <typename T>
__global__ void __launch_bounds__(256) flow_warp_fwd_kernel(const T* __restrict__ img, const float* __restrict__ flow,
                                                            T* __restrict__ out, int B, int C, int H, int W, int mode) {
  const long long plane = (long long)H * W;
  const long long total = (long long)B * plane;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / plane);
    const long long rem = idx - (long long)n * plane;
    const int y = (int)(rem / W), x = (int)(rem - (long long)y * W);
    const float* fp = flow + (long long)n * 2 * plane + rem;
    bool in_x, in_y;
    const float sx = sample_pos(x, __ldg(fp), W, mode & 3, in_x, !(mode & 4));
    const float sy = sample_pos(y, __ldg(fp + plane), H, mode & 3, in_y, !(mode & 4));
    const Taps tp = make_taps(sx, sy, H, W, W);
    const T* ip = img + (long long)n * C * plane;
    T* op = out + (long long)n * C * plane + rem;
    for (int c = 0; c < C; ++c) {
      const T* p = ip + (long long)c * plane;
      op[(long long)c * plane] = from_f32<T>(
          blend(ldg_f32(p + tp.off[0]), ldg_f32(p + tp.off[1]), ldg_f32(p + tp.off[2]), ldg_f32(p + tp.off[3]), tp));
    }
  }
}

// flow_warp backward: grad_image splatted with atomics (pre-zeroed), grad_flow written.
template <typename T>
__global__ void __launch_bounds__(256) flow_warp_bwd_kernel(const T* __restrict__ img, const float* __restrict__ flow,
                                                            const T* __restrict__ gout, T* __restrict__ gimg,
                                                            float* __restrict__ gflow, int B, int C, int H, int W,
                                                            int mode) {
  const long long plane = (long long)H * W;
  const long long total = (long long)B * plane;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / plane);
    const long long rem = idx - (long long)n * plane;
    const int y = (int)(rem / W), x = (int)(rem - (long long)y * W);
    const float* fp = flow + (long long)n * 2 * plane + rem;
    bool in_x, in_y;
    const float sx = sample_pos(x, __ldg(fp), W, mode, in_x);
    const float sy = sample_pos(y, __ldg(fp + plane), H, mode, in_y);
    const Taps tp = make_taps(sx, sy, H, W, W);
    const float fx = floorf(sx), fy = floorf(sy);
    const float wx1 = fx + 1.f - sx, wx0 = sx - fx, wy1 = fy + 1.f - sy, wy0 = sy - fy;
    const bool bx1 = (int)fx + 1 < W, by1 = (int)fy + 1 < H;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long cb = ((long long)n * C + c) * plane;
      const float gv = ldg_f32(gout + cb + rem);
      const T* p = img + cb;
      T* gp = gimg + cb;
      if (tp.w[0] != 0.f) atomic_add_t<T>(gp + tp.off[0], gv * tp.w[0]);
      if (tp.w[1] != 0.f) atomic_add_t<T>(gp + tp.off[1], gv * tp.w[1]);
      if (tp.w[2] != 0.f) atomic_add_t<T>(gp + tp.off[2], gv * tp.w[2]);
      if (tp.w[3] != 0.f) atomic_add_t<T>(gp + tp.off[3], gv * tp.w[3]);
      const float v_nw = ldg_f32(p + tp.off[0]);
      const float v_ne = bx1 ? ldg_f32(p + tp.off[1]) : 0.f;
      const float v_sw = by1 ? ldg_f32(p + tp.off[2]) : 0.f;
      const float v_se = (bx1 && by1) ? ldg_f32(p + tp.off[3]) : 0.f;
      gix += gv * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
      giy += gv * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
    }
    float* gf = gflow + (long long)n * 2 * plane + rem;
    gf[0] = in_x ? gix * pos_scale(W, mode) : 0.f;
    gf[plane] = in_y ? giy * pos_scale(H, mode) : 0.f;
  }
}

// ------------------------------------------------------------------ host launchers -------
static int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
static cudaError_t launch_bwd_t(const Geom& g, const void* x1, const void* x2, const float* flow, const void* out,
                                const void* gout, void* gx1, void* gx2, float* gflow, void* workspace,
                                cudaStream_t stream) {
  const long long in_elems = (long long)g.B * g.C * g.H * g.W;
  const bool fast = g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == kMDb;
  cudaError_t e;
  if (fast) {
    BwdArgs a;
    a.g = g;
    a.x1 = x1; a.x2 = x2; a.flow = flow; a.out = out; a.gout = gout;
    a.gx1 = gx1; a.gx2 = gx2; a.gflow = gflow;
    a.off = g.md - g.pad;
    a.tiles_x = (g.W + BT_X - 1) / BT_X;
    a.tiles_y = (g.H + BT_Y - 1) / BT_Y;
    if (flow != nullptr) {
      e = cudaMemsetAsync(gx2, 0, (size_t)in_elems * sizeof(T), stream);
      if (e != cudaSuccess) return e;
    }
    dim3 grid(a.tiles_x, a.tiles_y, g.B);
    corr_bwd_fast_kernel<T, 0><<<grid, 256, 0, stream>>>(a);
    corr_bwd_fast_kernel<T, 1><<<grid, 256, 0, stream>>>(a);
    count_launches(flow != nullptr ? 3 : 2);
    return cudaGetLastError();
  }
  // generic parameters
  if (flow == nullptr) {
    corr_bwd_generic_kernel<T><<<grid_for(in_elems, 256), 256, 0, stream>>>(
        g, (const T*)x1, (const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], (const T*)gout, (const T*)out, (T*)gx1, (T*)gx2);
    count_launches(1);
    return cudaGetLastError();
  }
  // with a flow: workspace holds the warped map and the gradient wrt it (2 * in_elems of T)
  T* warped = (T*)workspace;
  T* gwarped = warped + in_elems;
  if (g.x2s[2] != g.W || g.x2s[1] != (long long)g.H * g.W || g.x2s[0] != (long long)g.C * g.H * g.W)
    return cudaErrorNotSupported;  // generic + flow needs contiguous x2
  flow_warp_fwd_kernel<T><<<grid_for((long long)g.B * g.H * g.W, 256), 256, 0, stream>>>(
      (const T*)x2, flow, warped, g.B, g.C, g.H, g.W, g.warp_mode);
  corr_bwd_generic_kernel<T><<<grid_for(in_elems, 256), 256, 0, stream>>>(
      g, (const T*)x1, warped, (long long)g.C * g.H * g.W, (long long)g.H * g.W, (long long)g.W, (const T*)gout,
      (const T*)out, (T*)gx1, gwarped);
  e = cudaMemsetAsync(gx2, 0, (size_t)in_elems * sizeof(T), stream);
  if (e != cudaSuccess) return e;
  flow_warp_bwd_kernel<T><<<grid_for((long long)g.B * g.H * g.W, 256), 256, 0, stream>>>(
      (const T*)x2, flow, gwarped, (T*)gx2, gflow, g.B, g.C, g.H, g.W, g.warp_mode);
  count_launches(4);
  return cudaGetLastError();
}

cudaError_t launch_warp_corr_backward(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow,
                                      const void* out, const void* gout, void* gx1, void* gx2, float* gflow,
                                      void* workspace, cudaStream_t stream) {
  switch (dtype) {
    case CERB_F32: return launch_bwd_t<float>(g, x1, x2, flow, out, gout, gx1, gx2, gflow, workspace, stream);
    case CERB_F16: return launch_bwd_t<__half>(g, x1, x2, flow, out, gout, gx1, gx2, gflow, workspace, stream);
    case CERB_BF16: return launch_bwd_t<__nv_bfloat16>(g, x1, x2, flow, out, gout, gx1, gx2, gflow, workspace, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename T>
static cudaError_t warp_fwd_t(const void* image, const float* flow, void* out, int B, int C, int H, int W, int mode,
                              cudaStream_t stream) {
  flow_warp_fwd_kernel<T><<<grid_for((long long)B * H * W, 256), 256, 0, stream>>>((const T*)image, flow, (T*)out, B, C,
                                                                                    H, W, mode);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_flow_warp_forward(int dtype, const void* image, const float* flow, void* out, int B, int C, int H,
                                     int W, int mode, cudaStream_t stream) {
  switch (dtype) {
    case CERB_F32: return warp_fwd_t<float>(image, flow, out, B, C, H, W, mode, stream);
    case CERB_F16: return warp_fwd_t<__half>(image, flow, out, B, C, H, W, mode, stream);
    case CERB_BF16: return warp_fwd_t<__nv_bfloat16>(image, flow, out, B, C, H, W, mode, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename T>
static cudaError_t warp_bwd_t(const void* image, const float* flow, const void* gout, void* gimage, float* gflow, int B,
                              int C, int H, int W, int mode, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(gimage, 0, (size_t)B * C * H * W * sizeof(T), stream);
  if (e != cudaSuccess) return e;
  flow_warp_bwd_kernel<T><<<grid_for((long long)B * H * W, 256), 256, 0, stream>>>(
      (const T*)image, flow, (const T*)gout, (T*)gimage, gflow, B, C, H, W, mode);
  count_launches(2);
  return cudaGetLastError();
}

cudaError_t launch_flow_warp_backward(int dtype, const void* image, const float* flow, const void* gout, void* gimage,
                                      float* gflow, int B, int C, int H, int W, int mode, cudaStream_t stream) {
  switch (dtype) {
    case CERB_F32: return warp_bwd_t<float>(image, flow, gout, gimage, gflow, B, C, H, W, mode, stream);
    case CERB_F16: return warp_bwd_t<__half>(image, flow, gout, gimage, gflow, B, C, H, W, mode, stream);
    case CERB_BF16: return warp_bwd_t<__nv_bfloat16>(image, flow, gout, gimage, gflow, B, C, H, W, mode, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cerb
