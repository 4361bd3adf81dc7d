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
// Fast path (k=1, s1=s2=1, md=4): corr_bwd_tiled_kernel below, one template for both gradients.
#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

long long* get_trace_buffer();
#define BWD_TRACE(slot) do { if (a.dbg && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 0 && threadIdx.x == 0) a.dbg[(slot)] = clock64(); } while (0)

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
  long long* dbg;  // optional clock64 trace of CTA 0 (cerb_debug_set_trace_buffer)
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

// Register-tiled backward.  Both gradients have the form
//     g[c, q] = 1/C * sum_e  G[e, q] * S[c, q + e],      e in [-4,4]^2
//   WHICH 0 (grad_x1):  G[e, q] = gO'[e, q]            S = warped x2      (q = x1 pixel)
//   WHICH 1 (grad_x2w): G[e, q] = gO'[-e, q + e]       S = x1             (q = warped-map pixel)
// with gO' = grad_out masked by the LeakyReLU of the saved output.  One CTA per 8x32 tile of q:
//   * the whole G tile (81 planes x 8 x 32, 83 KB) is built once and stays in shared memory;
//   * 32 channels of the S halo tile (16 x 40, re-warped on the fly for grad_x1) per chunk;
//   * thread = (8-pixel strip, row, 4 channels): 32 accumulators, per row displacement the 9 x 8
//     slice of G in registers (LDS.128 broadcast across the 8 channel-subset lanes) and, per
//     channel, one 16-float row of S (4 LDS.128) -> 72 FFMA;
//   * results are staged through shared memory and written coalesced; for grad_x2 with a flow
//     they go straight through the bilinear-warp backward (4 atomics, flow gradient in registers).
constexpr int BS_XS = 44;                      // S row stride (floats)
constexpr int BS_CH = BH_Y * BS_XS + 4;        // S channel stride: == 4 (mod 32) -> 8 channels hit 8 bank groups
constexpr int BCH = 32;                        // channels per chunk
constexpr int BO_CH = BT_Y * BT_X + 4;         // staged-output channel stride
constexpr int BG_FLOATS = kD2b * BT_Y * BT_X;  // 20736
constexpr size_t BWD_SMEM = sizeof(float) * (BG_FLOATS + BCH * BS_CH);

template <typename T, int WHICH>
__global__ void __launch_bounds__(256, 1) corr_bwd_tiled_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float bsm[];
  float* Gs = bsm;                  // [81][8][32]
  float* Ss = bsm + BG_FLOATS;      // [32][BS_CH]   (aliased by the staged output [32][BO_CH])
  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int iy0 = blockIdx.y * BT_Y, ix0 = blockIdx.x * BT_X;
  const T* __restrict__ x1 = (const T*)a.x1 + (long long)n * g.x1s[0];
  const T* __restrict__ x2 = (const T*)a.x2 + (long long)n * g.x2s[0];
  const T* __restrict__ gout = (const T*)a.gout + (long long)n * g.os[0];
  const T* __restrict__ outp = a.out ? (const T*)a.out + (long long)n * g.os[0] : nullptr;
  const bool warped = a.flow != nullptr;
  const bool mask = g.has_act && outp != nullptr;

  BWD_TRACE(WHICH * 32 + 0);
  // ---- G tile: thread = pixel of the tile, planes in batches of 9 (18 independent loads in flight)
  {
    const int ty = tid >> 5, tx = tid & 31;
    const int qy = iy0 + ty, qx = ix0 + tx;
    const bool q_ok = qy < g.H && qx < g.W;
#pragma unroll 1
    for (int dy0 = 0; dy0 < kDb; dy0 += 3) {
      // Loads are unconditional from clamped (always valid) addresses and validity is applied
      // afterwards: with only 7 predicate registers the compiler otherwise consumes each predicated
      // load right after issuing it and the 54 loads serialise.
      float gvv[3 * kDb], ovv[3 * kDb];
#pragma unroll
      for (int r = 0; r < 3 * kDb; ++r) {
        const int dyi = dy0 + r / kDb, dxi = r % kDb;
        const int plane = dyi * kDb + dxi;
        const int py = (WHICH == 0) ? qy : qy + dyi - kMDb, px = (WHICH == 0) ? qx : qx + dxi - kMDb;
        const int d = (WHICH == 0) ? plane : (kD2b - 1 - plane);
        const int oy = min(max(py - a.off, 0), g.outH - 1), ox = min(max(px - a.off, 0), g.outW - 1);
        const long long o = (long long)d * g.os[1] + (long long)oy * g.os[2] + ox;
        gvv[r] = ldcg_f32(gout + o);
        ovv[r] = mask ? ldcg_f32(outp + o) : 1.f;
      }
#pragma unroll
      for (int r = 0; r < 3 * kDb; ++r) {
        const int dyi = dy0 + r / kDb, dxi = r % kDb;
        const int py = (WHICH == 0) ? qy : qy + dyi - kMDb, px = (WHICH == 0) ? qx : qx + dxi - kMDb;
        const int oy = py - a.off, ox = px - a.off;
        const bool ok = q_ok && py >= 0 && py < g.H && px >= 0 && px < g.W && oy >= 0 && oy < g.outH && ox >= 0 && ox < g.outW;
        float v = ok ? gvv[r] : 0.f;
        if (!(ovv[r] > 0.f)) v *= g.slope;   // ovv == 1 when there is no activation
        Gs[(dy0 * kDb + r) * (BT_Y * BT_X) + tid] = v;
      }
      BWD_TRACE(WHICH * 32 + 6 + dy0 / 3);
    }
  }

  // ---- staging plan for the S halo tile: positions handled by this thread (fixed for the tile)
  constexpr int NPOS = BH_Y * BH_X;
  constexpr int PPT = (NPOS + 255) / 256;
  Taps taps[PPT];
  int sdst[PPT];
  unsigned valid_mask = 0;
  const bool gather = (WHICH == 0) && warped;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = tid + j * 256;
    sdst[j] = -1;
    taps[j].off[0] = taps[j].off[1] = taps[j].off[2] = taps[j].off[3] = 0;
    taps[j].w[0] = taps[j].w[1] = taps[j].w[2] = taps[j].w[3] = 0.f;
    if (i < NPOS) {
      const int hy = i / BH_X, hx = i - hy * BH_X;
      sdst[j] = hy * BS_XS + hx;
      const int qy = iy0 - kMDb + hy, qx = ix0 - kMDb + hx;
      if (qy >= 0 && qy < g.H && qx >= 0 && qx < g.W) {
        valid_mask |= 1u << j;
        if (gather) {
          const float* fp = a.flow + (long long)n * g.fls[0] + (long long)qy * g.fls[2] + qx;
          bool in_x, in_y;
          const float sx = sample_pos(qx, __ldg(fp), g.W, g.warp_mode, in_x);
          const float sy = sample_pos(qy, __ldg(fp + g.fls[1]), g.H, g.warp_mode, in_y);
          taps[j] = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
        } else {
          const long long hs = (WHICH == 0) ? g.x2s[2] : g.x1s[2];
          taps[j].off[0] = (int)(qy * hs) + qx;
          taps[j].w[0] = 1.f;
        }
      }
    }
  }

  // ---- roles
  const int subset = tid & 7, combo = tid >> 3;   // compute: 8 channel subsets x 32 (strip,row) combos
  const int strip = combo & 3, yrow = combo >> 2;
  const int wty = tid >> 5, wtx = tid & 31;       // write-out: one pixel per thread
  const int wy = iy0 + wty, wx = ix0 + wtx;
  const bool wpix_ok = wy < g.H && wx < g.W;

  // grad_x2 with a flow: this pixel's bilinear taps for the splat and the flow gradient
  Taps mytap;
  bool in_x = false, in_y = false;
  float gix = 0.f, giy = 0.f, wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f;
  int sx0 = 0, sy0 = 0, sx1 = 0, sy1 = 0;
  if (WHICH == 1 && warped && wpix_ok) {
    const float* fp = a.flow + (long long)n * g.fls[0] + (long long)wy * g.fls[2] + wx;
    const float sx = sample_pos(wx, __ldg(fp), g.W, g.warp_mode, in_x);
    const float sy = sample_pos(wy, __ldg(fp + g.fls[1]), g.H, g.warp_mode, in_y);
    mytap = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
    const float fx = floorf(sx), fy = floorf(sy);
    wx1 = fx + 1.f - sx; wx0 = sx - fx;
    wy1 = fy + 1.f - sy; wy0 = sy - fy;
    sx0 = (int)fx; sy0 = (int)fy;
    sx1 = (sx0 + 1 < g.W) ? sx0 + 1 : sx0;
    sy1 = (sy0 + 1 < g.H) ? sy0 + 1 : sy0;
  }

  const float inv_c = 1.0f / (float)g.C;
  const T* __restrict__ src = (WHICH == 0) ? x2 : x1;
  const long long src_cs = (WHICH == 0) ? g.x2s[1] : g.x1s[1];
  const long long plane_elems = (long long)g.H * g.W;
  T* gdst = (T*)((WHICH == 0) ? a.gx1 : a.gx2) + (long long)n * g.C * plane_elems;

  for (int c0 = 0; c0 < g.C; c0 += BCH) {
    __syncthreads();  // G tile complete (first pass) / previous chunk's write-out done
    if (c0 == 0) BWD_TRACE(WHICH * 32 + 1);
    // ---- stage 32 channels of the S halo tile, 4 channels of loads in flight per thread
    if (gather) {
      for (int cq = 0; cq < BCH; cq += 4) {
        float tv[4][PPT][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int ch = min(c0 + cq + k, g.C - 1);  // clamped: always a valid plane (masked below)
          const T* plane = src + (long long)ch * src_cs;
#pragma unroll
          for (int j = 0; j < PPT; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q) tv[k][j][q] = ldg_f32(plane + taps[j].off[q]);  // off = 0 for invalid positions
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool ch_ok = c0 + cq + k < g.C;
#pragma unroll
          for (int j = 0; j < PPT; ++j) {
            const float r = (ch_ok && ((valid_mask >> j) & 1u)) ? blend(tv[k][j][0], tv[k][j][1], tv[k][j][2], tv[k][j][3], taps[j]) : 0.f;
            if (sdst[j] >= 0) Ss[(cq + k) * BS_CH + sdst[j]] = r;
          }
        }
      }
    } else {
      for (int cq = 0; cq < BCH; cq += 16) {
        float tv[16][PPT];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int ch = min(c0 + cq + k, g.C - 1);
          const T* plane = src + (long long)ch * src_cs;
#pragma unroll
          for (int j = 0; j < PPT; ++j) tv[k][j] = ldg_f32(plane + taps[j].off[0]);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const bool ch_ok = c0 + cq + k < g.C;
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            if (sdst[j] >= 0) Ss[(cq + k) * BS_CH + sdst[j]] = (ch_ok && ((valid_mask >> j) & 1u)) ? tv[k][j] : 0.f;
        }
      }
    }
    __syncthreads();
    if (c0 == 0) BWD_TRACE(WHICH * 32 + 2);

    // ---- contraction: acc[cb][px] for channels c0 + cb*8 + subset
    float acc[4][8];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb)
#pragma unroll
      for (int px = 0; px < 8; ++px) acc[cb][px] = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < kDb; ++dy) {
      float gv[kDb][8];
#pragma unroll
      for (int dx = 0; dx < kDb; ++dx) {
        const float* gp = Gs + ((dy * kDb + dx) * BT_Y + yrow) * BT_X + strip * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(gp);
        const float4 g1 = *reinterpret_cast<const float4*>(gp + 4);
        gv[dx][0] = g0.x; gv[dx][1] = g0.y; gv[dx][2] = g0.z; gv[dx][3] = g0.w;
        gv[dx][4] = g1.x; gv[dx][5] = g1.y; gv[dx][6] = g1.z; gv[dx][7] = g1.w;
      }
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const float* sp = Ss + (cb * 8 + subset) * BS_CH + (yrow + dy) * BS_XS + strip * 8;
        float sv[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(sp + 4 * q);
          sv[4 * q + 0] = s4.x; sv[4 * q + 1] = s4.y; sv[4 * q + 2] = s4.z; sv[4 * q + 3] = s4.w;
        }
#pragma unroll
        for (int px = 0; px < 8; ++px)
#pragma unroll
          for (int dx = 0; dx < kDb; ++dx) acc[cb][px] = fmaf(gv[dx][px], sv[px + dx], acc[cb][px]);
      }
    }
    __syncthreads();  // every thread is done reading the S tile: reuse it as the output stage
    if (c0 == 0) BWD_TRACE(WHICH * 32 + 3);

    float* Os = Ss;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      float* op = Os + (cb * 8 + subset) * BO_CH + yrow * BT_X + strip * 8;
      *reinterpret_cast<float4*>(op) = make_float4(acc[cb][0] * inv_c, acc[cb][1] * inv_c, acc[cb][2] * inv_c, acc[cb][3] * inv_c);
      *reinterpret_cast<float4*>(op + 4) = make_float4(acc[cb][4] * inv_c, acc[cb][5] * inv_c, acc[cb][6] * inv_c, acc[cb][7] * inv_c);
    }
    __syncthreads();

    if (c0 == 0) BWD_TRACE(WHICH * 32 + 4);
    // ---- write-out: one pixel per thread, channels of the chunk
    if (wpix_ok) {
      const int cmax = (g.C - c0) < BCH ? (g.C - c0) : BCH;
      const float* orow = Os + wty * BT_X + wtx;
      if (!(WHICH == 1 && warped)) {
        T* gp = gdst + (long long)c0 * plane_elems + (long long)wy * g.W + wx;
        for (int c = 0; c < cmax; ++c) gp[(long long)c * plane_elems] = from_f32<T>(orow[c * BO_CH]);
      } else {
        const bool bx1 = sx1 != sx0, by1 = sy1 != sy0;  // far taps inside the image
        // channels in batches of 8: the 32 tap loads of a batch are in flight together
        for (int cb0 = 0; cb0 < cmax; cb0 += 8) {
          float tvv[8][4], sv8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const bool ok = cb0 + k < cmax;
            const T* xp = x2 + (long long)min(c0 + cb0 + k, g.C - 1) * g.x2s[1];
            sv8[k] = ok ? orow[(cb0 + k) * BO_CH] : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) tvv[k][q] = ldg_f32(xp + mytap.off[q]);  // clamped taps: always valid
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // taps outside the image contribute nothing to d/d(position)
            if (!bx1) { tvv[k][1] = 0.f; tvv[k][3] = 0.f; }
            if (!by1) { tvv[k][2] = 0.f; tvv[k][3] = 0.f; }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (cb0 + k < cmax) {
              const float s = sv8[k];
              T* gp = gdst + (long long)(c0 + cb0 + k) * plane_elems;
              if (mytap.w[0] != 0.f) atomic_add_t<T>(gp + sy0 * g.W + sx0, s * mytap.w[0]);
              if (mytap.w[1] != 0.f) atomic_add_t<T>(gp + sy0 * g.W + sx1, s * mytap.w[1]);
              if (mytap.w[2] != 0.f) atomic_add_t<T>(gp + sy1 * g.W + sx0, s * mytap.w[2]);
              if (mytap.w[3] != 0.f) atomic_add_t<T>(gp + sy1 * g.W + sx1, s * mytap.w[3]);
              gix += s * ((tvv[k][1] - tvv[k][0]) * wy1 + (tvv[k][3] - tvv[k][2]) * wy0);
              giy += s * ((tvv[k][2] - tvv[k][0]) * wx1 + (tvv[k][3] - tvv[k][1]) * wx0);
            }
          }
        }
      }
    }
  }
  BWD_TRACE(WHICH * 32 + 5);
  if (WHICH == 1 && warped && wpix_ok) {
    float* gf = a.gflow + (long long)n * 2 * plane_elems + (long long)wy * g.W + wx;
    gf[0] = in_x ? gix * pos_scale(g.W, g.warp_mode) : 0.f;
    gf[plane_elems] = in_y ? giy * pos_scale(g.H, g.warp_mode) : 0.f;
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
    a.dbg = get_trace_buffer();
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
    static bool attr_set = false;  // benign race: idempotent
    if (!attr_set) {
      e = cudaFuncSetAttribute(corr_bwd_tiled_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(corr_bwd_tiled_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    corr_bwd_tiled_kernel<T, 0><<<grid, 256, BWD_SMEM, stream>>>(a);
    corr_bwd_tiled_kernel<T, 1><<<grid, 256, BWD_SMEM, stream>>>(a);
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
