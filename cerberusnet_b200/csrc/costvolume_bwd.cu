// costvolume_bwd.cu -- backward of the fused level op, and the stand-alone flow warp, sm_100a.
//
// Replaces (reference paths relative to the reference checkout):
//   correlation_backward_cuda     nnet_training/correlation_package/correlation_cuda.cpp:28-43
//     correlation_backward_input1/2   correlation_cuda_kernel.cu:97-172, 174-242, launch loop :326-429
//   LeakyReluBackward             autograd of pwcnet_sfd.py:182
//   GridSampler2DBackward + norm_grid / mesh_grid backward   UnFlowLoss.py:22-32,83-94
//
// All batch items in one launch (the reference launches 2*B kernels, .cu:386,407), no padded NHWC
// scratch.  With a flow the backward is warp-forward (re-materialised warped map, workspace) ->
// two correlation-backward kernels -> warp-backward (atomics into a zeroed grad_x2, flow gradient).
//
// Fast path (k=1, s1=s2=1, md=4): corr_bwd_tiled_kernel below, one template for both gradients.
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

long long* get_trace_buffer();
// tensor-core backward (costvolume_bwd_tc.cu)
bool tc_backward_supported(const Geom& g, int dtype, const void* s0, const long long s0s[3], const void* x1);
cudaError_t launch_corr_backward_tc(const Geom& g, const void* s0, const long long s0s[3], int s0_roll, const void* x1,
                                    const void* out, const void* gout, void* gx1, void* gsecond, int gsecond_roll,
                                    cudaStream_t stream);
// which backward kernel the fast path takes: -1 automatic, 0 CUDA cores only, 1 tensor cores wherever supported
// (initialised from CERB_DEBUG_BWD_TC; cerb_debug_set_backward_kernel for tests and the bench)
static int g_bwd_tc_mode = -2;
int get_backward_kernel_mode() {
  if (g_bwd_tc_mode == -2) g_bwd_tc_mode = getenv("CERB_DEBUG_BWD_TC") ? atoi(getenv("CERB_DEBUG_BWD_TC")) : -1;
  return g_bwd_tc_mode;
}
void set_backward_kernel_mode(int mode) { g_bwd_tc_mode = mode < -1 || mode > 1 ? -1 : mode; }
#define BWD_TRACE(slot) do { if (a.dbg && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 0 && threadIdx.x == 0) a.dbg[(slot)] = clock64(); } while (0)

constexpr int kMDb = 4;
constexpr int kDb = 9;
constexpr int kD2b = 81;
constexpr int BT_X = 32;
constexpr int BH_X = BT_X + 2 * kMDb;  // 40-wide halo

template <typename T> __device__ __forceinline__ void atomic_add_t(T* p, float v);
template <> __device__ __forceinline__ void atomic_add_t<float>(float* p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_add_t<__half>(__half* p, float v) { atomicAdd(p, __float2half_rn(v)); }
template <> __device__ __forceinline__ void atomic_add_t<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

struct BwdArgs {
  long long* dbg;  // optional clock64 trace of one CTA (cerb_debug_set_trace_buffer)
  Geom g;
  const void* s[2];        // the "other" operand per gradient: [0] second correlation input (warped x2), [1] x1
  long long s_ns[2], s_cs[2], s_hs[2];   // its N, C, H strides in elements
  const void* out;         // activated forward output (sign only), may be null
  const void* gout;
  void* gdst[2];           // contiguous NCHW gradients: [0] wrt x1, [1] wrt the second correlation input
  int off;                 // md - pad
  int tiles_x, tiles_y;
  int nwin;                // 9 x 9 displacement windows per axis (1 for max_displacement 4)
  int async_ok[2];         // fp32 and 16-byte alignment of s[] rows: stage the halo tiles with cp.async
  int vec4_ok[2];          // 16-bit and 8-byte alignment of s[] rows: stage the halo tiles with 4-element loads
  int s_roll[2], g_roll[2]; // batch-item roll applied when reading s[] / writing gdst[] (x2_batch_roll when they are x2 / grad_x2 themselves)
  // fused warp backward (phase 1 of the fused kernel): the gradient with respect to the warped map stays on chip and is
  // splatted through the bilinear taps straight into grad_x2, the flow gradient is reduced over the channels
  const void* x2;          // the un-warped second map (tap values for the flow gradient)
  const float* flow;
  void* gx2;               // splat target: grad_x2 (fp32 input) or the fp32 accumulator (16-bit inputs), zeroed by the launcher
  float* gflow;
};

// Register-tiled backward of the correlation.  Both gradients have the form
//     g[c, q] = 1/C * sum_e  G[e, q] * S[c, q + e],      e in [-4,4]^2
//   WHICH 0 (grad_x1):     G[e, q] = gO'[e, q]            S = second correlation input   (q = x1 pixel)
//   WHICH 1 (grad_second): G[e, q] = gO'[-e, q + e]       S = x1                         (q = pixel of the second input)
// with gO' = grad_out masked by the LeakyReLU of the saved output.  With a flow the second input
// is the warped map: the launcher materialises it before, and pushes grad_second through the warp
// backward after, so this kernel is a plain correlation backward.  One CTA per 8x32 tile of q:
//   * the whole G tile (81 planes x 8 x 32, 83 KB) is fetched once with per-thread asynchronous
//     copies (all 162 loads of a thread in flight; zero-fill outside the output), the saved output
//     goes through the not-yet-used S buffer and the LeakyReLU mask is applied in place;
//   * 32 channels of the S halo tile (16 x 40) per chunk, 16-byte asynchronous copies;
//   * thread = (8-pixel strip, row, 4 channels): 32 accumulators, per row displacement the 9 x 8
//     slice of G in registers (LDS.128 broadcast across the 8 channel-subset lanes) and, per
//     channel, one 16-float row of S (4 LDS.128) -> 72 FFMA;
//   * results are staged through shared memory and written coalesced.
constexpr int BS_XS = 44;                      // S row stride (floats)
constexpr int BCH = 32;                        // channels per chunk
// TYB rows per tile: 8 (one 256-thread CTA per SM) or 4 (two 128-thread CTAs per SM, so the load
// phases of one overlap the contraction of the other)
template <int TYB>
struct BwdCfg {
  static constexpr int BT_Y = TYB;
  static constexpr int BH_Y = TYB + 2 * kMDb;
  static constexpr int NT = 32 * TYB;                     // threads: one per tile pixel
  static constexpr int BS_CH = BH_Y * BS_XS + 4;          // S channel stride: 8 channels hit 8 distinct 16-byte bank groups
  static constexpr int BO_CH = BT_Y * BT_X + 4;           // staged-output channel stride
  static constexpr int BG_FLOATS = kD2b * BT_Y * BT_X;
  static constexpr size_t SMEM = sizeof(float) * (BG_FLOATS + BCH * BS_CH);
  static_assert(BCH * BS_CH >= BG_FLOATS, "the saved-output tile borrows the S buffer");
};

// SPLAT (WHICH == 1 only): instead of writing the gradient with respect to the (warped) second input, every thread
// pushes its pixel's values through the four bilinear taps into grad_x2 (red.global.add) and accumulates the flow
// gradient over the channels -- the stand-alone warp-backward kernel and its workspace round trip folded in.
template <typename T, typename GT, int WHICH, int TYB, bool SPLAT>
__device__ __forceinline__ void corr_bwd_phase(const BwdArgs& a, float* bsm) {
  using Cfg = BwdCfg<TYB>;
  constexpr int BT_Y = Cfg::BT_Y, BH_Y = Cfg::BH_Y, NT = Cfg::NT, BS_CH = Cfg::BS_CH, BO_CH = Cfg::BO_CH,
                BG_FLOATS = Cfg::BG_FLOATS;
  float* Gs = bsm;                  // [81][8][32]
  float* Ss = bsm + BG_FLOATS;      // [32][BS_CH]   (aliased by the staged output [32][BO_CH])
  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int iy0 = blockIdx.y * BT_Y, ix0 = blockIdx.x * BT_X;
  const int n_src = (n + a.s_roll[WHICH]) % g.B, n_dst = (n + a.g_roll[WHICH]) % g.B;
  const T* __restrict__ src = (const T*)a.s[WHICH] + (long long)n_src * a.s_ns[WHICH];
  const long long src_cs = a.s_cs[WHICH], src_hs = a.s_hs[WHICH];
  const T* __restrict__ gout = (const T*)a.gout + (long long)n * g.os[0];
  const T* __restrict__ outp = a.out ? (const T*)a.out + (long long)n * g.os[0] : nullptr;
  const bool mask = g.has_act && outp != nullptr;
  constexpr bool kF32 = std::is_same<T, float>::value;

  // ---- SPLAT: sampling data of this thread's pixel (thread = pixel of the tile), fixed for all channels and windows
  Taps tp_in, tp_out;             // taps into x2 (its strides) / into the contiguous gradient
  float s_wx1 = 0.f, s_wx0 = 0.f, s_wy1 = 0.f, s_wy0 = 0.f, gix = 0.f, giy = 0.f;
  bool s_bx1 = false, s_by1 = false, s_inx = false, s_iny = false;
  const int n_x2 = (n + g.x2roll) % g.B;   // batch item of x2 / grad_x2 paired with item n
  if constexpr (SPLAT) {
    const int sy_ = min(iy0 + (tid >> 5), g.H - 1), sx_ = min(ix0 + (tid & 31), g.W - 1);
    const float* fp = a.flow + (long long)n * g.fls[0] + (long long)sy_ * g.fls[2] + sx_;
    const float sx = sample_pos(sx_, __ldg(fp), g.W, g.warp_mode, s_inx);
    const float sy = sample_pos(sy_, __ldg(fp + g.fls[1]), g.H, g.warp_mode, s_iny);
    tp_in = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
    tp_out = make_taps(sx, sy, g.H, g.W, g.W);
    const float fx = floorf(sx), fy = floorf(sy);
    s_wx1 = fx + 1.f - sx; s_wx0 = sx - fx; s_wy1 = fy + 1.f - sy; s_wy0 = sy - fy;
    s_bx1 = (int)fx + 1 < g.W; s_by1 = (int)fy + 1 < g.H;
  }

  // max_displacement > 4: the D x D displacement range is covered by 9 x 9 windows (origins 0, 8, ...,
  // D - 9, as in the forward); the windows' contributions are accumulated into the output, rows /
  // columns that an earlier window already covered are zeroed in G.
  const int nwin = a.nwin;
  for (int win = 0; win < nwin * nwin; ++win) {
  const int wyi = win / nwin, wxi = win - wyi * nwin;
  const int woy = min(8 * wyi, g.D - kDb), wox = min(8 * wxi, g.D - kDb);
  const int lo_y = wyi == 0 ? 0 : min(8 * (wyi - 1), g.D - kDb) + kDb - woy;
  const int lo_x = wxi == 0 ? 0 : min(8 * (wxi - 1), g.D - kDb) + kDb - wox;
  const int ey0 = woy - g.md, ex0 = wox - g.md;   // pixel displacement of window entry (0, 0)
  if (win > 0) __syncthreads();                    // previous window's write-out has finished with the S buffer

  BWD_TRACE(WHICH * 32 + 0);
  // ---- G tile: thread = pixel of the tile
  {
    const int ty = tid >> 5, tx = tid & 31;
    const int qy = iy0 + ty, qx = ix0 + tx;
    const bool q_ok = qy < g.H && qx < g.W;
    if constexpr (kF32) {
      const uint32_t gs_u32 = smem_u32(Gs) + 4u * (uint32_t)tid, os_u32 = smem_u32(Ss) + 4u * (uint32_t)tid;
      if (WHICH == 0) {
        // the same output pixel in every displacement plane: one bounds test
        const int oy = qy - a.off, ox = qx - a.off;
        const bool ok = q_ok && oy >= 0 && oy < g.outH && ox >= 0 && ox < g.outW;
        const long long o0 = (long long)min(max(oy, 0), g.outH - 1) * g.os[2] + min(max(ox, 0), g.outW - 1);
#pragma unroll 1
        for (int dyi = 0; dyi < kDb; ++dyi) {
          const long long orow = o0 + (long long)((woy + dyi) * g.D + wox) * g.os[1];
          const bool rok = ok && dyi >= lo_y;
#pragma unroll
          for (int dxi = 0; dxi < kDb; ++dxi) {
            const int plane = dyi * kDb + dxi;
            const uint32_t nb = (rok && dxi >= lo_x) ? 4u : 0u;
            cp_async4(gs_u32 + 4u * (uint32_t)(plane * (BT_Y * BT_X)), gout + orow + (long long)dxi * g.os[1], nb);
            if (mask) cp_async4(os_u32 + 4u * (uint32_t)(plane * (BT_Y * BT_X)), outp + orow + (long long)dxi * g.os[1], nb);
          }
        }
      } else {
        // window entry (dyi, dxi) = pixel displacement e reads the plane of -e at pixel q + e:
        // validity per row / column displacement as bit masks
        unsigned row_ok = 0, col_ok = 0;
#pragma unroll
        for (int d = 0; d < kDb; ++d) {
          const int py = qy + ey0 + d, px = qx + ex0 + d;
          if (q_ok && d >= lo_y && py >= 0 && py < g.H && py - a.off >= 0 && py - a.off < g.outH) row_ok |= 1u << d;
          if (d >= lo_x && px >= 0 && px < g.W && px - a.off >= 0 && px - a.off < g.outW) col_ok |= 1u << d;
        }
        const int oxc0 = qx + ex0 - a.off;
#pragma unroll 1
        for (int dyi = 0; dyi < kDb; ++dyi) {
          const int oy = min(max(qy + ey0 + dyi - a.off, 0), g.outH - 1);
          const bool rok = (row_ok >> dyi) & 1u;
          const long long orow = (long long)((g.D - 1 - (woy + dyi)) * g.D + (g.D - 1 - wox)) * g.os[1] + (long long)oy * g.os[2];
#pragma unroll
          for (int dxi = 0; dxi < kDb; ++dxi) {
            const int plane = dyi * kDb + dxi;
            const long long o = orow - (long long)dxi * g.os[1] + min(max(oxc0 + dxi, 0), g.outW - 1);
            const uint32_t nb = (rok && ((col_ok >> dxi) & 1u)) ? 4u : 0u;
            cp_async4(gs_u32 + 4u * (uint32_t)(plane * (BT_Y * BT_X)), gout + o, nb);
            if (mask) cp_async4(os_u32 + 4u * (uint32_t)(plane * (BT_Y * BT_X)), outp + o, nb);
          }
        }
      }
      cp_async_commit();
      cp_async_wait_all();
      if (mask) {
        __syncthreads();
        // LeakyReLU backward in place: every thread masks the elements it copied itself
#pragma unroll 9
        for (int plane = 0; plane < kD2b; ++plane) {
          const int i = plane * (BT_Y * BT_X) + tid;
          if (!(Ss[i] > 0.f)) Gs[i] *= g.slope;
        }
      }
    } else {
#pragma unroll 1
      for (int dy0 = 0; dy0 < kDb; dy0 += 3) {
        // Loads are unconditional from clamped (always valid) addresses and validity is applied
        // afterwards: with only 7 predicate registers the compiler otherwise consumes each predicated
        // load right after issuing it and the 54 loads serialise.
        float gvv[3 * kDb], ovv[3 * kDb];
#pragma unroll
        for (int r = 0; r < 3 * kDb; ++r) {
          const int dyi = dy0 + r / kDb, dxi = r % kDb;
          const int py = (WHICH == 0) ? qy : qy + ey0 + dyi, px = (WHICH == 0) ? qx : qx + ex0 + dxi;
          const int dg = (woy + dyi) * g.D + wox + dxi;
          const int d = (WHICH == 0) ? dg : (g.D2 - 1 - dg);
          const int oy = min(max(py - a.off, 0), g.outH - 1), ox = min(max(px - a.off, 0), g.outW - 1);
          const long long o = (long long)d * g.os[1] + (long long)oy * g.os[2] + ox;
          gvv[r] = ldcg_f32(gout + o);
          ovv[r] = mask ? ldcg_f32(outp + o) : 1.f;
        }
#pragma unroll
        for (int r = 0; r < 3 * kDb; ++r) {
          const int dyi = dy0 + r / kDb, dxi = r % kDb;
          const int py = (WHICH == 0) ? qy : qy + ey0 + dyi, px = (WHICH == 0) ? qx : qx + ex0 + dxi;
          const int oy = py - a.off, ox = px - a.off;
          const bool ok = q_ok && dyi >= lo_y && dxi >= lo_x && py >= 0 && py < g.H && px >= 0 && px < g.W && oy >= 0 && oy < g.outH && ox >= 0 && ox < g.outW;
          float v = ok ? gvv[r] : 0.f;
          if (!(ovv[r] > 0.f)) v *= g.slope;   // ovv == 1 when there is no activation
          Gs[(dy0 * kDb + r) * (BT_Y * BT_X) + tid] = v;
        }
      }
    }
  }
  BWD_TRACE(WHICH * 32 + 6);

  // ---- roles
  const int subset = tid & 7, combo = tid >> 3;   // compute: 8 channel subsets x 32 (strip,row) combos
  const int strip = combo & 3, yrow = combo >> 2;
  const int wty = tid >> 5, wtx = tid & 31;       // write-out: one pixel per thread
  const int wy = iy0 + wty, wx = ix0 + wtx;
  const bool wpix_ok = wy < g.H && wx < g.W;

  const float inv_c = 1.0f / (float)g.C;
  const long long plane_elems = (long long)g.H * g.W;
  T* gdst = (T*)a.gdst[WHICH] + (long long)n_dst * g.C * plane_elems;
  const bool async_s = kF32 && a.async_ok[WHICH];

  for (int c0 = 0; c0 < g.C; c0 += BCH) {
    __syncthreads();  // G tile complete (first pass) / previous chunk's write-out done
    if (c0 == 0) BWD_TRACE(WHICH * 32 + 1);
    const int cmax = (g.C - c0) < BCH ? (g.C - c0) : BCH;
    // ---- stage the S halo tile of this chunk: [cmax][16][40] at (iy0 - 4, ix0 - 4), zero outside the image
    if (async_s) {
      constexpr int V4_ROW = BH_X / 4;                 // 10 float4 per halo row
      constexpr int U = BH_Y * V4_ROW;                 // float4 units per channel
      // unit index tid, tid + NT, ... decomposed into (channel, unit-in-channel) incrementally
      int c = tid / U, r = tid - c * U;
      const uint32_t ss_u32 = smem_u32(Ss);
      while (c < cmax) {
        const int hy = r / V4_ROW, v = r - hy * V4_ROW;
        const int qy = iy0 + ey0 + hy, qx = ix0 + ex0 + 4 * v;
        const bool row_ok = qy >= 0 && qy < g.H && qx >= 0 && qx < g.W;
        const int nbytes = row_ok ? min(16, 4 * (g.W - qx)) : 0;
        const T* p = src + (long long)(c0 + c) * src_cs + (long long)(row_ok ? qy : 0) * src_hs + (row_ok ? qx : 0);
        cp_async16(ss_u32 + 4u * (uint32_t)(c * BS_CH + hy * BS_XS + 4 * v), p, (uint32_t)nbytes);
        r += NT % U; c += NT / U;
        if (r >= U) { r -= U; ++c; }
      }
      cp_async_commit();
      cp_async_wait_all();
    } else if (sizeof(T) == 2 && a.vec4_ok[WHICH]) {
      // 16-bit: 8-byte loads of 4 elements (a halo row is 10 of them), 15 in flight per thread, widened on the way in
      constexpr int V_ROW = BH_X / 4;
      constexpr int U = BH_Y * V_ROW;
      constexpr int NB = 15;
      const int total = cmax * U;
      for (int u0 = tid; u0 < total; u0 += NT * NB) {
        uint2 pk[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int u = min(u0 + k * NT, total - 1);
          const int c = u / U, r = u - c * U;
          const int hy = r / V_ROW, v = r - hy * V_ROW;
          const int qy = min(max(iy0 + ey0 + hy, 0), g.H - 1), qx = min(max(ix0 + ex0 + 4 * v, 0), g.W - 4);
          pk[k] = __ldg(reinterpret_cast<const uint2*>(src + (long long)(c0 + c) * src_cs + (long long)qy * src_hs + qx));
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int u = u0 + k * NT;
          if (u < total) {
            const int c = u / U, r = u - c * U;
            const int hy = r / V_ROW, v = r - hy * V_ROW;
            const int qy = iy0 + ey0 + hy, qx = ix0 + ex0 + 4 * v;
            const bool ok = qy >= 0 && qy < g.H && qx >= 0 && qx + 3 < g.W;   // W % 4 == 0: a group is all in or all out
            const T* hv = reinterpret_cast<const T*>(&pk[k]);
            const float4 w4 = ok ? make_float4(to_f32<T>(hv[0]), to_f32<T>(hv[1]), to_f32<T>(hv[2]), to_f32<T>(hv[3]))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(Ss + c * BS_CH + hy * BS_XS + 4 * v) = w4;
          }
        }
      }
    } else {
      constexpr int NPOS = BH_Y * BH_X;
      const int total = cmax * NPOS;
      for (int u0 = tid; u0 < total; u0 += NT * 8) {
        float tv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int u = min(u0 + k * NT, total - 1);
          const int c = u / NPOS, i = u - c * NPOS;
          const int hy = i / BH_X, hx = i - hy * BH_X;
          const int qy = min(max(iy0 + ey0 + hy, 0), g.H - 1), qx = min(max(ix0 + ex0 + hx, 0), g.W - 1);
          tv[k] = ldg_f32(src + (long long)(c0 + c) * src_cs + (long long)qy * src_hs + qx);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int u = u0 + k * NT;
          if (u < total) {
            const int c = u / NPOS, i = u - c * NPOS;
            const int hy = i / BH_X, hx = i - hy * BH_X;
            const int qy = iy0 + ey0 + hy, qx = ix0 + ex0 + hx;
            Ss[c * BS_CH + hy * BS_XS + hx] = (qy >= 0 && qy < g.H && qx >= 0 && qx < g.W) ? tv[k] : 0.f;
          }
        }
      }
    }
    __syncthreads();
    if (c0 == 0) BWD_TRACE(WHICH * 32 + 2);

    // ---- contraction: acc[cb][px] for channels c0 + cb*8 + subset (channel groups past C are skipped)
    const int ncb = (cmax + 7) >> 3;
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
        if (cb < ncb) {   // uniform across the CTA
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
    if constexpr (SPLAT) {
      // ---- splat: the chunk's gradient with respect to the warped map goes through the taps of this thread's pixel
      if (wpix_ok) {
        const float* orow = Os + wty * BT_X + wtx;
        const T* x2n = (const T*)a.x2 + (long long)n_x2 * g.x2s[0];
        GT* gx2n = (GT*)a.gx2 + (long long)n_x2 * g.C * plane_elems;
        for (int c4 = 0; c4 < cmax; c4 += 4) {   // 4 channels per batch: 16 independent tap loads in flight
          float gvv[4], v[4][4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = min(c4 + k, cmax - 1);
            gvv[k] = orow[c * BO_CH];
            const T* pch = x2n + (long long)(c0 + c) * g.x2s[1];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[k][q] = ldg_f32(pch + tp_in.off[q]);   // clamped taps: always valid addresses
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (c4 + k < cmax) {
              GT* gp = gx2n + (long long)(c0 + c4 + k) * plane_elems;
              if (tp_out.w[0] != 0.f) atomic_add_t<GT>(gp + tp_out.off[0], gvv[k] * tp_out.w[0]);
              if (tp_out.w[1] != 0.f) atomic_add_t<GT>(gp + tp_out.off[1], gvv[k] * tp_out.w[1]);
              if (tp_out.w[2] != 0.f) atomic_add_t<GT>(gp + tp_out.off[2], gvv[k] * tp_out.w[2]);
              if (tp_out.w[3] != 0.f) atomic_add_t<GT>(gp + tp_out.off[3], gvv[k] * tp_out.w[3]);
              const float v_nw = v[k][0];
              const float v_ne = s_bx1 ? v[k][1] : 0.f;
              const float v_sw = s_by1 ? v[k][2] : 0.f;
              const float v_se = (s_bx1 && s_by1) ? v[k][3] : 0.f;
              gix += gvv[k] * ((v_ne - v_nw) * s_wy1 + (v_se - v_sw) * s_wy0);
              giy += gvv[k] * ((v_sw - v_nw) * s_wx1 + (v_se - v_ne) * s_wx0);
            }
          }
        }
      }
    } else
    // ---- write-out: one pixel per thread, channels of the chunk (rows of 32 pixels: coalesced)
    if (wpix_ok) {
      const float* orow = Os + wty * BT_X + wtx;
      T* gp = gdst + (long long)c0 * plane_elems + (long long)wy * g.W + wx;
      if (win == 0) {
#pragma unroll 8
        for (int c = 0; c < cmax; ++c) gp[(long long)c * plane_elems] = from_f32<T>(orow[c * BO_CH]);
      } else {   // later displacement windows add to what this thread wrote before
#pragma unroll 8
        for (int c = 0; c < cmax; ++c)
          gp[(long long)c * plane_elems] = from_f32<T>(to_f32<T>(gp[(long long)c * plane_elems]) + orow[c * BO_CH]);
      }
    }
  }
  }  // displacement windows
  if constexpr (SPLAT) {
    const int sy_ = iy0 + (tid >> 5), sx_ = ix0 + (tid & 31);
    if (sy_ < g.H && sx_ < g.W) {
      float* gf = a.gflow + (long long)n * 2 * g.H * g.W + (long long)sy_ * g.W + sx_;
      gf[0] = s_inx ? gix * pos_scale(g.W, g.warp_mode) : 0.f;
      gf[(long long)g.H * g.W] = s_iny ? giy * pos_scale(g.H, g.warp_mode) : 0.f;
    }
  }
  BWD_TRACE(WHICH * 32 + 5);
}

// one gradient per launch (kept for the two-launch path: no flow, or debugging)
template <typename T, int WHICH, int TYB>
__global__ void __launch_bounds__(32 * TYB, TYB == 4 ? 2 : 1) corr_bwd_tiled_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float bsm[];
  corr_bwd_phase<T, T, WHICH, TYB, false>(a, bsm);
}

// Both gradients of a tile in one CTA: grad_x1 first, then -- the 81 planes of grad_out / out it needs are the ones just
// read, shifted by at most 4 pixels, so they come from L2 -- the gradient with respect to the second input, which with a
// flow never leaves the chip: it is splatted into grad_x2 and reduced into grad_flow right here.
template <typename T, typename GT, int TYB, bool SPLAT>
__global__ void __launch_bounds__(32 * TYB, TYB == 4 ? 2 : 1) corr_bwd_fused_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float bsm[];
  corr_bwd_phase<T, T, 0, TYB, false>(a, bsm);
  __syncthreads();
  corr_bwd_phase<T, GT, 1, TYB, SPLAT>(a, bsm);
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
                                                               T* __restrict__ gx1, T* __restrict__ gsecond, int sec_roll) {
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
    const int n2 = (n + sec_roll) % g.B;   // batch item of the second input paired with item n
    const T* sp = second + (long long)n2 * sec_ns + (long long)c * sec_cs;
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
    gsecond[idx + (long long)(n2 - n) * g.C * g.H * g.W] = from_f32<T>(a2 * inv);
  }
}

// ------------------------------------------------------------------ stand-alone warp ------
// flow_warp forward: one thread per (n, y, x), looping channels (coalesced along x).  `img` and
// `flow` carry N/C/H strides (W stride 1); the output is contiguous NCHW.
template <typename T>
__global__ void __launch_bounds__(256) flow_warp_fwd_kernel(const T* __restrict__ img, long long in_ns, long long in_cs,
                                                            long long in_hs, const float* __restrict__ flow,
                                                            long long f_ns, long long f_cs, long long f_hs,
                                                            T* __restrict__ out, int B, int C, int H, int W, int mode,
                                                            int roll) {
  const long long plane = (long long)H * W;
  const long long total = (long long)B * plane;
  const AxisConst ax = make_axis(W), ay = make_axis(H);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / plane);
    const long long rem = idx - (long long)n * plane;
    const int y = (int)(rem / W), x = (int)(rem - (long long)y * W);
    const float* fp = flow + (long long)n * f_ns + (long long)y * f_hs + x;
    bool in_x, in_y;
    const float sx = sample_pos(x, __ldg(fp), ax, mode & 3, in_x, !(mode & 4));
    const float sy = sample_pos(y, __ldg(fp + f_cs), ay, mode & 3, in_y, !(mode & 4));
    const Taps tp = make_taps(sx, sy, H, W, in_hs);
    const T* ip = img + (long long)((n + roll) % B) * in_ns;
    T* op = out + (long long)n * C * plane + rem;
    int c = 0;
    for (; c + 4 <= C; c += 4) {   // 16 independent tap loads in flight
      float v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const T* p = ip + (long long)(c + k) * in_cs;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[k][q] = ldg_f32(p + tp.off[q]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) op[(long long)(c + k) * plane] = from_f32<T>(blend(v[k][0], v[k][1], v[k][2], v[k][3], tp));
    }
    for (; c < C; ++c) {
      const T* p = ip + (long long)c * in_cs;
      op[(long long)c * plane] = from_f32<T>(
          blend(ldg_f32(p + tp.off[0]), ldg_f32(p + tp.off[1]), ldg_f32(p + tp.off[2]), ldg_f32(p + tp.off[3]), tp));
    }
  }
}

// flow_warp backward: grad_image splatted with atomics (pre-zeroed, contiguous), grad_flow written
// (contiguous); `img` and `flow` carry strides, `gout` is contiguous.
template <typename T, typename GT>
__global__ void __launch_bounds__(256) flow_warp_bwd_kernel(const T* __restrict__ img, long long in_ns, long long in_cs,
                                                            long long in_hs, const float* __restrict__ flow,
                                                            long long f_ns, long long f_cs, long long f_hs,
                                                            const T* __restrict__ gout, GT* __restrict__ gimg,
                                                            float* __restrict__ gflow, int B, int C, int H, int W,
                                                            int mode, int roll) {
  const long long plane = (long long)H * W;
  const long long total = (long long)B * plane;
  const AxisConst ax = make_axis(W), ay = make_axis(H);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / plane);
    const long long rem = idx - (long long)n * plane;
    const int y = (int)(rem / W), x = (int)(rem - (long long)y * W);
    const float* fp = flow + (long long)n * f_ns + (long long)y * f_hs + x;
    bool in_x, in_y;
    const float sx = sample_pos(x, __ldg(fp), ax, mode, in_x);
    const float sy = sample_pos(y, __ldg(fp + f_cs), ay, mode, in_y);
    const Taps tp = make_taps(sx, sy, H, W, in_hs);   // reads of img
    const Taps to = make_taps(sx, sy, H, W, W);       // splat into the contiguous gradient
    const float fx = floorf(sx), fy = floorf(sy);
    const float wx1 = fx + 1.f - sx, wx0 = sx - fx, wy1 = fy + 1.f - sy, wy0 = sy - fy;
    const bool bx1 = (int)fx + 1 < W, by1 = (int)fy + 1 < H;
    float gix = 0.f, giy = 0.f;
    const int n2 = (n + roll) % B;   // the image (and its gradient) live at the rolled batch item
    const T* ip = img + (long long)n2 * in_ns;
    for (int c0 = 0; c0 < C; c0 += 4) {   // 4 channels per batch: 20 independent loads in flight
      float gv[4], v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = min(c0 + k, C - 1);
        gv[k] = ldg_f32(gout + ((long long)n * C + c) * plane + rem);
        const T* p = ip + (long long)c * in_cs;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[k][q] = ldg_f32(p + tp.off[q]);   // clamped taps: always valid addresses
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c0 + k < C) {
          GT* gp = gimg + ((long long)n2 * C + c0 + k) * plane;
          if (to.w[0] != 0.f) atomic_add_t<GT>(gp + to.off[0], gv[k] * to.w[0]);
          if (to.w[1] != 0.f) atomic_add_t<GT>(gp + to.off[1], gv[k] * to.w[1]);
          if (to.w[2] != 0.f) atomic_add_t<GT>(gp + to.off[2], gv[k] * to.w[2]);
          if (to.w[3] != 0.f) atomic_add_t<GT>(gp + to.off[3], gv[k] * to.w[3]);
          const float v_nw = v[k][0];
          const float v_ne = bx1 ? v[k][1] : 0.f;
          const float v_sw = by1 ? v[k][2] : 0.f;
          const float v_se = (bx1 && by1) ? v[k][3] : 0.f;
          gix += gv[k] * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
          giy += gv[k] * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
        }
      }
    }
    float* gf = gflow + (long long)n * 2 * plane + rem;
    gf[0] = in_x ? gix * pos_scale(W, mode) : 0.f;
    gf[plane] = in_y ? giy * pos_scale(H, mode) : 0.f;
  }
}

// fp32 splat accumulator -> 16-bit gradient (the fused backward accumulates the warp's splat in fp32)
template <typename T>
__global__ void __launch_bounds__(256) cvt_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<T>(src[i]);
}

// ------------------------------------------------------------------ host launchers -------
// flow_warp backward through shared-memory windows (costvolume_splat.cu), fp32
cudaError_t launch_flow_warp_backward_box(const float* img, const long long in_s[3], const float* flow, const long long f_s[3],
                                          const float* gout, float* gimg, float* gflow, int B, int C, int H, int W, int mode,
                                          int roll, cudaStream_t stream);
static int grid_for(long long total, int block);
// splat of a gradient wrt the warped map (contiguous, x1's batch order) into a zeroed grad_image + the flow gradient: the
// shared-memory-window kernel for fp32 when the tensors can be TMA tensors, scattered atomics otherwise
template <typename T, typename GT>
__global__ void flow_warp_bwd_kernel(const T* __restrict__ img, long long in_ns, long long in_cs, long long in_hs,
                                     const float* __restrict__ flow, long long f_ns, long long f_cs, long long f_hs,
                                     const T* __restrict__ gout, GT* __restrict__ gimg, float* __restrict__ gflow, int B, int C,
                                     int H, int W, int mode, int roll);
template <typename T, typename GT>
static cudaError_t splat_backward(const T* img, long long in_ns, long long in_cs, long long in_hs, const float* flow, long long f_ns,
                                  long long f_cs, long long f_hs, const T* gout, GT* gimg, float* gflow, int B, int C, int H, int W,
                                  int mode, int roll, cudaStream_t stream) {
  if constexpr (std::is_same<T, float>::value && std::is_same<GT, float>::value) {
    const long long in_s[3] = {in_ns, in_cs, in_hs}, f_s[3] = {f_ns, f_cs, f_hs};
    const cudaError_t e = launch_flow_warp_backward_box(img, in_s, flow, f_s, gout, gimg, gflow, B, C, H, W, mode, roll, stream);
    if (e != cudaErrorNotSupported) return e;
  }
  flow_warp_bwd_kernel<T, GT><<<grid_for((long long)B * H * W, 256), 256, 0, stream>>>(img, in_ns, in_cs, in_hs, flow, f_ns, f_cs, f_hs,
                                                                                        gout, gimg, gflow, B, C, H, W, mode, roll);
  return cudaGetLastError();
}
static int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// per-device opt-in to > 48 KB of dynamic shared memory (function attributes are per device: one process may drive several GPUs)
template <typename K>
static cudaError_t ensure_smem_attr(K kern, size_t smem, unsigned long long& devs) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
  if (bit != 0ull && (devs & bit)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) devs |= bit;
  return e;
}

template <typename T, int TYB>
static cudaError_t launch_tiled_pair(BwdArgs& a, const Geom& g, cudaStream_t stream) {
  using Cfg = BwdCfg<TYB>;
  a.tiles_y = (g.H + TYB - 1) / TYB;
  dim3 grid(a.tiles_x, a.tiles_y, g.B);
  static unsigned long long devs0 = 0ull, devs1 = 0ull;
  cudaError_t e = ensure_smem_attr(corr_bwd_tiled_kernel<T, 0, TYB>, Cfg::SMEM, devs0);
  if (e != cudaSuccess) return e;
  e = ensure_smem_attr(corr_bwd_tiled_kernel<T, 1, TYB>, Cfg::SMEM, devs1);
  if (e != cudaSuccess) return e;
  corr_bwd_tiled_kernel<T, 0, TYB><<<grid, Cfg::NT, Cfg::SMEM, stream>>>(a);
  corr_bwd_tiled_kernel<T, 1, TYB><<<grid, Cfg::NT, Cfg::SMEM, stream>>>(a);
  return cudaSuccess;
}

// both gradients of a tile in one CTA; with SPLAT the gradient with respect to the warped map never leaves the chip
template <typename T, typename GT, int TYB, bool SPLAT>
static cudaError_t launch_tiled_fused(BwdArgs& a, const Geom& g, cudaStream_t stream) {
  using Cfg = BwdCfg<TYB>;
  a.tiles_y = (g.H + TYB - 1) / TYB;
  dim3 grid(a.tiles_x, a.tiles_y, g.B);
  static unsigned long long devs = 0ull;
  cudaError_t e = ensure_smem_attr(corr_bwd_fused_kernel<T, GT, TYB, SPLAT>, Cfg::SMEM, devs);
  if (e != cudaSuccess) return e;
  corr_bwd_fused_kernel<T, GT, TYB, SPLAT><<<grid, Cfg::NT, Cfg::SMEM, stream>>>(a);
  return cudaSuccess;
}

template <typename T>
static cudaError_t launch_bwd_t(const Geom& g, const void* x1, const void* x2, const float* flow, const void* out,
                                const void* gout, void* gx1, void* gx2, float* gflow, void* workspace,
                                cudaStream_t stream) {
  const long long in_elems = (long long)g.B * g.C * g.H * g.W;
  const bool fast = g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md >= kMDb;
  cudaError_t e;
  if (fast) {
    // with a flow: the workspace holds the warped map and the gradient wrt it (2 * in_elems of T)
    T* warped = (T*)workspace;
    T* gwarped = warped ? warped + in_elems : nullptr;
    if (flow != nullptr) {
      if (workspace == nullptr) return cudaErrorInvalidValue;
      flow_warp_fwd_kernel<T><<<grid_for((long long)g.B * g.H * g.W, 256), 256, 0, stream>>>(
          (const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], flow, g.fls[0], g.fls[1], g.fls[2], warped, g.B, g.C, g.H, g.W,
          g.warp_mode, g.x2roll);
    }
    const long long cs = (long long)g.H * g.W, ns = (long long)g.C * cs;
    BwdArgs a;
    a.dbg = get_trace_buffer();
    a.g = g;
    a.s[0] = flow ? (const void*)warped : x2;
    a.s_ns[0] = flow ? ns : g.x2s[0]; a.s_cs[0] = flow ? cs : g.x2s[1]; a.s_hs[0] = flow ? (long long)g.W : g.x2s[2];
    a.s[1] = x1;
    a.s_ns[1] = g.x1s[0]; a.s_cs[1] = g.x1s[1]; a.s_hs[1] = g.x1s[2];
    a.out = out; a.gout = gout;
    a.gdst[0] = gx1;
    a.gdst[1] = flow ? (void*)gwarped : gx2;
    a.s_roll[0] = flow ? 0 : g.x2roll; a.s_roll[1] = 0;      // the warped workspace is already in x1's batch order
    a.g_roll[0] = 0; a.g_roll[1] = flow ? 0 : g.x2roll;
    a.off = g.md - g.pad;
    a.nwin = (g.D - kDb + 7) / 8 + 1;
    for (int w = 0; w < 2; ++w)   // halo rows start at ix0 + window origin - md: 16-byte aligned only if md % 4 == 0
      a.async_ok[w] = std::is_same<T, float>::value && ((uintptr_t)a.s[w] % 16) == 0 && (a.s_ns[w] % 4) == 0 &&
                      (a.s_cs[w] % 4) == 0 && (a.s_hs[w] % 4) == 0 && (g.md % 4) == 0;
    for (int w = 0; w < 2; ++w)
      a.vec4_ok[w] = sizeof(T) == 2 && ((uintptr_t)a.s[w] % 8) == 0 && (a.s_ns[w] % 4) == 0 && (a.s_cs[w] % 4) == 0 &&
                     (a.s_hs[w] % 4) == 0 && (g.md % 4) == 0 && (g.W % 4) == 0;
    a.tiles_x = (g.W + BT_X - 1) / BT_X;
    static const int force_ty = getenv("CERB_DEBUG_BWD_TY") ? atoi(getenv("CERB_DEBUG_BWD_TY")) : 0;
    static const bool unfused = getenv("CERB_DEBUG_BWD_UNFUSED") != nullptr;
    const bool ty4 = force_ty ? force_ty == 4 : true;
    // Tensor-core kernel (costvolume_bwd_tc.cu): fp32, md = pad = 4, C <= 128, 16-byte aligned rows.  Taken automatically from
    // 512 tiles of 8 x 16 and 48 channels up, where it measured faster than the fused CUDA-core kernel (batch 8: HRNet level 3
    // 459 vs 508 us, HRNet level 2 250 vs 272, PWC level 3 191 vs 195; with 32 channels the two are within a few per cent of
    // each other either way); cerb_debug_set_backward_kernel / CERB_DEBUG_BWD_TC: 1 forces it wherever supported, 0 turns it off
    if (std::is_same<T, float>::value && !unfused && !force_ty) {
      const int tc_mode = get_backward_kernel_mode();
      const long long s0s[3] = {a.s_ns[0], a.s_cs[0], a.s_hs[0]};
      const long long tiles = (long long)g.B * ((g.H + 7) / 8) * ((g.W + 15) / 16);
      if (tc_mode != 0 && (tc_mode == 1 || (tiles >= 512 && g.C >= 48)) && tc_backward_supported(g, CERB_F32, a.s[0], s0s, x1)) {
        if (flow != nullptr) {
          e = cudaMemsetAsync(gx2, 0, (size_t)in_elems * sizeof(T), stream);
          if (e != cudaSuccess) return e;
        }
        if (flow != nullptr) {
          // gradient wrt the warped map into the workspace (coalesced stores), then the splat as its own kernel
          e = launch_corr_backward_tc(g, a.s[0], s0s, a.s_roll[0], x1, out, gout, gx1, gwarped, 0, stream);
          if (e == cudaSuccess) {
            e = splat_backward<T, T>((const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], flow, g.fls[0], g.fls[1], g.fls[2], gwarped, (T*)gx2,
                                     gflow, g.B, g.C, g.H, g.W, g.warp_mode, g.x2roll, stream);
            if (e != cudaSuccess) return e;
            count_launches(4);
            return cudaGetLastError();
          }
        } else {
          e = launch_corr_backward_tc(g, a.s[0], s0s, a.s_roll[0], x1, out, gout, gx1, gx2, g.x2roll, stream);
          if (e == cudaSuccess) {
            count_launches(1);
            return cudaGetLastError();
          }
        }
        if (e != cudaErrorNotSupported) return e;
      }
    }
    if (!unfused && ty4) {
      // one kernel for both gradients.  With a flow the second one is splatted through the bilinear taps straight into
      // grad_x2 (zeroed here; 16-bit: an fp32 accumulator in the workspace, narrowed once) and reduced into grad_flow
      // inside that kernel: no gradient-of-the-warped-map workspace, no separate warp-backward pass.
      a.x2 = x2; a.flow = flow; a.gflow = gflow;
      if (flow == nullptr) {
        a.gx2 = nullptr;
        e = launch_tiled_fused<T, T, 4, false>(a, g, stream);
        if (e != cudaSuccess) return e;
        count_launches(1);
        return cudaGetLastError();
      }
      if (sizeof(T) == 2) {
        float* acc32 = reinterpret_cast<float*>(gwarped + in_elems);
        e = cudaMemsetAsync(acc32, 0, (size_t)in_elems * sizeof(float), stream);
        if (e != cudaSuccess) return e;
        a.gx2 = acc32;
        e = launch_tiled_fused<T, float, 4, true>(a, g, stream);
        if (e != cudaSuccess) return e;
        cvt_from_f32_kernel<T><<<grid_for(in_elems, 256), 256, 0, stream>>>(acc32, (T*)gx2, in_elems);
        count_launches(4);
      } else {
        e = cudaMemsetAsync(gx2, 0, (size_t)in_elems * sizeof(T), stream);
        if (e != cudaSuccess) return e;
        a.gx2 = gx2;
        e = launch_tiled_fused<T, T, 4, true>(a, g, stream);
        if (e != cudaSuccess) return e;
        count_launches(3);
      }
      return cudaGetLastError();
    }
    a.x2 = nullptr; a.flow = nullptr; a.gx2 = nullptr; a.gflow = nullptr;
    e = ty4 ? launch_tiled_pair<T, 4>(a, g, stream) : launch_tiled_pair<T, 8>(a, g, stream);
    if (e != cudaSuccess) return e;
    if (flow != nullptr) {
      if (sizeof(T) == 2) {
        // 16-bit: splat into an fp32 accumulator (third workspace region), then narrow -- 16-bit
        // atomics are several times slower and round at every add
        float* acc32 = reinterpret_cast<float*>(gwarped + in_elems);
        e = cudaMemsetAsync(acc32, 0, (size_t)in_elems * sizeof(float), stream);
        if (e != cudaSuccess) return e;
        e = splat_backward<T, float>((const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], flow, g.fls[0], g.fls[1], g.fls[2], gwarped, acc32, gflow, g.B,
            g.C, g.H, g.W, g.warp_mode, g.x2roll, stream);
        if (e != cudaSuccess) return e;
        cvt_from_f32_kernel<T><<<grid_for(in_elems, 256), 256, 0, stream>>>(acc32, (T*)gx2, in_elems);
      } else {
        e = cudaMemsetAsync(gx2, 0, (size_t)in_elems * sizeof(T), stream);
        if (e != cudaSuccess) return e;
        e = splat_backward<T, T>((const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], flow, g.fls[0], g.fls[1], g.fls[2], gwarped, (T*)gx2, gflow, g.B,
            g.C, g.H, g.W, g.warp_mode, g.x2roll, stream);
        if (e != cudaSuccess) return e;
      }
    }
    count_launches(flow != nullptr ? (sizeof(T) == 2 ? 6 : 5) : 2);
    return cudaGetLastError();
  }
  // generic parameters
  if (flow == nullptr) {
    corr_bwd_generic_kernel<T><<<grid_for(in_elems, 256), 256, 0, stream>>>(
        g, (const T*)x1, (const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], (const T*)gout, (const T*)out, (T*)gx1, (T*)gx2,
        g.x2roll);
    count_launches(1);
    return cudaGetLastError();
  }
  // with a flow: workspace holds the warped map and the gradient wrt it (2 * in_elems of T)
  T* warped = (T*)workspace;
  T* gwarped = warped + in_elems;
  flow_warp_fwd_kernel<T><<<grid_for((long long)g.B * g.H * g.W, 256), 256, 0, stream>>>(
      (const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], flow, g.fls[0], g.fls[1], g.fls[2], warped, g.B, g.C, g.H, g.W,
      g.warp_mode, g.x2roll);
  corr_bwd_generic_kernel<T><<<grid_for(in_elems, 256), 256, 0, stream>>>(
      g, (const T*)x1, warped, (long long)g.C * g.H * g.W, (long long)g.H * g.W, (long long)g.W, (const T*)gout,
      (const T*)out, (T*)gx1, gwarped, 0);
  e = cudaMemsetAsync(gx2, 0, (size_t)in_elems * sizeof(T), stream);
  if (e != cudaSuccess) return e;
  e = splat_backward<T, T>((const T*)x2, g.x2s[0], g.x2s[1], g.x2s[2], flow, g.fls[0], g.fls[1], g.fls[2], gwarped, (T*)gx2, gflow, g.B, g.C,
      g.H, g.W, g.warp_mode, g.x2roll, stream);
        if (e != cudaSuccess) return e;
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
  const long long cs = (long long)H * W;
  flow_warp_fwd_kernel<T><<<grid_for((long long)B * H * W, 256), 256, 0, stream>>>(
      (const T*)image, (long long)C * cs, cs, (long long)W, flow, 2 * cs, cs, (long long)W, (T*)out, B, C, H, W, mode, 0);
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
  const long long cs = (long long)H * W;
  e = splat_backward<T, T>((const T*)image, (long long)C * cs, cs, (long long)W, flow, 2 * cs, cs, (long long)W, (const T*)gout, (T*)gimage,
      gflow, B, C, H, W, mode, 0, stream);
        if (e != cudaSuccess) return e;
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
