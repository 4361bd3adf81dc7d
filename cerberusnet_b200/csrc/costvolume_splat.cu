// costvolume_splat.cu -- flow_warp backward (bilinear, border) for sm_100a with the splat done in shared memory.
//
// Replaces GridSampler2DBackward + the backward of norm_grid / mesh_grid of the reference's flow_warp
// (loss_functions/UnFlowLoss.py:11-32,83-94): grad_image[c, taps(q)] += w_k(q) * grad_out[c, q], grad_flow[q] =
// sum_c grad_out[c, q] * d(sample)/d(position).
//
// Why: four scattered red.global.add.f32 per (position, channel) run at the L2's atomic-transaction rate -- 246 G/s on a
// B200 for a flow that varies from pixel to pixel, 205 us for 8 x 48 x 128 x 256 (tools/microbench/atomics.cu) -- while
// coalesced reductions run at 4 TB/s.  One CTA per 8 x 16 tile of positions: when the tile's taps fit a 22 x 36 window of
// the image (flow variation up to +-6 px inside the tile) the splat goes through shared memory, 8 channels per pass:
//   * the x2 window (tap values for the flow gradient) arrives by TMA, zero outside the image;
//   * the gradient window is accumulated with 32-bit integer shared-memory reductions in fixed point (fp32 shared-memory
//     atomics are compare-and-swap loops, SASS ATOMS.CAST.SPIN; integer ones are native ATOMS.ADD): scale = a power of two
//     chosen from the pass's largest |grad_out| so that 128 terms cannot overflow, i.e. 2^-23 of that value per term -- far
//     inside the fp32 parity bar -- and integer sums are associative, so the window is bit-reproducible;
//   * it is converted in place and added to grad_image by ONE TMA reduce (cp.reduce.async.bulk.tensor .add, SASS UTMAREDG),
//     clipped at the image border by the tensor map.
// Tiles whose taps do not fit (wild flows) use scattered red.global.add like the old kernel.  fp32 only; strided image.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

bool make_tmap_nchw(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                    const long long strides[3], int bx, int by, int bc, bool swizzle128);
unsigned long long* get_path_counters();

namespace splat {

constexpr int TY = 8, TX = 16, NT = TY * TX;
constexpr int MARGIN = 6, CB = 8;
constexpr int BOX_H = TY + 2 * MARGIN + 2;                        // 22
constexpr int BOX_W = (TX + 2 * MARGIN + 2 + 3 + 3) / 4 * 4;      // 36: origin aligned down to 4 elements
constexpr uint32_t BOX_BYTES = CB * BOX_H * BOX_W * 4;            // 25344
constexpr uint32_t SMEM_BYTES = 2 * BOX_BYTES + 128;
static_assert(BOX_BYTES % 128 == 0, "TMA shared-memory alignment");

struct Args {
  const float* img;       // strided (in_s: N, C, H)
  long long in_s[3];
  const float* flow;
  long long f_s[3];
  const float* gout;      // contiguous (B, C, H, W)
  float* gimg;            // contiguous, zeroed by the caller
  float* gflow;           // contiguous (B, 2, H, W)
  int B, C, H, W, mode, roll;
  int tiles_x, tiles_y;
  unsigned long long* path_ctr;
};

__device__ __forceinline__ void red_shared_s32(uint32_t addr, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void tma_reduce_add_4d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(NT, 4)
flow_warp_bwd_box_kernel(const Args a, const __grid_constant__ CUtensorMap tm_img, const __grid_constant__ CUtensorMap tm_gimg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const uint32_t xbox = smem_u32(smem), gbox = xbox + BOX_BYTES;
  __shared__ uint64_t xb_full;
  __shared__ int bbox_red[4 * 4];
  __shared__ uint32_t amax_red[2][4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.z, y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
  const int y = y0 + (tid >> 4), x = x0 + (tid & 15);
  const bool pix_ok = y < a.H && x < a.W;
  const int yc = min(y, a.H - 1), xc = min(x, a.W - 1);
  const long long plane = (long long)a.H * a.W;
  int n2 = n + a.roll;   // the image (and its gradient) live at the rolled batch item
  if (n2 >= a.B) n2 -= a.B;

  if (tid == 0) {
    mbar_init(&xb_full, 1);
    fence_barrier_init();
  }
  // ---- sampling data of this thread's position
  const float* fp = a.flow + (long long)n * a.f_s[0] + (long long)yc * a.f_s[2] + xc;
  bool inx, iny;
  const float sx = sample_pos(xc, __ldg(fp), a.W, a.mode, inx);
  const float sy = sample_pos(yc, __ldg(fp + a.f_s[1]), a.H, a.mode, iny);
  const float fx = floorf(sx), fy = floorf(sy);
  const int ix0 = (int)fx, iy0 = (int)fy;
  const float wx1 = fx + 1.f - sx, wx0 = sx - fx, wy1 = fy + 1.f - sy, wy0 = sy - fy;
  const bool bx1 = ix0 + 1 < a.W, by1 = iy0 + 1 < a.H;
  const float w_nw = wx1 * wy1, w_ne = bx1 ? wx0 * wy1 : 0.f, w_sw = by1 ? wx1 * wy0 : 0.f, w_se = (bx1 && by1) ? wx0 * wy0 : 0.f;
  // ---- do the tile's taps fit one window?
  int box_ox, box_oy;
  bool fits;
  {
    int xmin = pix_ok ? ix0 : 0x7fffffff, xmax = pix_ok ? ix0 + 1 : -0x7fffffff;
    int ymin = pix_ok ? iy0 : 0x7fffffff, ymax = pix_ok ? iy0 + 1 : -0x7fffffff;
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if (lane == 0) *reinterpret_cast<int4*>(bbox_red + 4 * warp) = make_int4(xmin, xmax, ymin, ymax);
    // the gradient window starts out zero
    for (uint32_t o = (uint32_t)tid * 16u; o < BOX_BYTES; o += NT * 16u) sts128(gbox + o, make_float4(0.f, 0.f, 0.f, 0.f));
    __syncthreads();
    xmin = 0x7fffffff; xmax = -0x7fffffff; ymin = 0x7fffffff; ymax = -0x7fffffff;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int4 b = *reinterpret_cast<const int4*>(bbox_red + 4 * w);
      xmin = min(xmin, b.x); xmax = max(xmax, b.y);
      ymin = min(ymin, b.z); ymax = max(ymax, b.w);
    }
    box_ox = xmin & ~3;   // TMA box starts are 16-byte aligned
    box_oy = ymin;
    fits = xmin <= xmax && xmax - box_ox < BOX_W && ymax - box_oy < BOX_H;
  }
  if (tid == 0 && a.path_ctr != nullptr) atomicAdd(&a.path_ctr[fits ? 1 : 2], 1ull);
  const float* gop = a.gout + (long long)n * a.C * plane + (long long)yc * a.W + xc;
  float gix = 0.f, giy = 0.f;
  if (fits) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&xb_full, BOX_BYTES);
      tma_load_4d(smem, &tm_img, &xb_full, box_ox, box_oy, 0, n2);
      // (prefetching the later passes' windows into L2 here with cp.async.bulk.prefetch.tensor made the kernel 5-15 % slower)
    }
    const uint32_t t_nw = (uint32_t)(((iy0 - box_oy) * BOX_W + (ix0 - box_ox)) * 4);
    // the gradients of a pass are requested one pass ahead: a first-touch load from HBM takes several thousand cycles here,
    // and with one request per pass that latency, not the shared-memory work, was the period of a pass
    float gn[CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) gn[i] = __ldg(gop + (long long)min(i, a.C - 1) * plane);
    for (int c0 = 0, ps = 0; c0 < a.C; c0 += CB, ++ps) {
      // this pass's gradients; fixed-point scale from the largest of the tile
      float gv[CB];
      float am = 0.f;
#pragma unroll
      for (int i = 0; i < CB; ++i) {
        gv[i] = (pix_ok && c0 + i < a.C) ? gn[i] : 0.f;
        am = fmaxf(am, fabsf(gv[i]));
      }
      if (c0 + CB < a.C) {
#pragma unroll
        for (int i = 0; i < CB; ++i) gn[i] = __ldg(gop + (long long)min(c0 + CB + i, a.C - 1) * plane);
      }
      const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(am));   // non-negative floats order like integers
      if (lane == 0) amax_red[ps & 1][warp] = wm;
      // tap values of this pass's channels
#ifndef CERB_SPLAT_X_NOX2
      mbar_wait(&xb_full, (uint32_t)(ps & 1));
#endif
      float tv[CB][4];
#pragma unroll
      for (int i = 0; i < CB; ++i) {
        const uint32_t pl = xbox + (uint32_t)(i * BOX_H * BOX_W * 4) + (pix_ok ? t_nw : 0u);
        tv[i][0] = lds_f32(pl); tv[i][1] = lds_f32(pl + 4);
        tv[i][2] = lds_f32(pl + BOX_W * 4); tv[i][3] = lds_f32(pl + BOX_W * 4 + 4);
      }
      __syncthreads();   // everyone is done with the x2 window; the pass's maxima are visible; the gradient window is zero
#ifndef CERB_SPLAT_X_NOX2
      if (tid == 0 && c0 + CB < a.C) {
        mbar_arrive_expect_tx(&xb_full, BOX_BYTES);
        tma_load_4d(smem, &tm_img, &xb_full, box_ox, box_oy, c0 + CB, n2);
      }
#endif
      const uint32_t cm = max(max(amax_red[ps & 1][0], amax_red[ps & 1][1]), max(amax_red[ps & 1][2], amax_red[ps & 1][3]));
      // cm < 2^(e+1) with e its exponent: scale 2^(22-e) keeps every term below 2^23 and a sum of 128 below 2^30
      const int e = min(max((int)(cm >> 23) - 127, -100), 100);
      const float fs = __uint_as_float((uint32_t)(127 + 22 - e) << 23), ifs = __uint_as_float((uint32_t)(127 - 22 + e) << 23);
      if (pix_ok) {
#pragma unroll
        for (int i = 0; i < CB; ++i) {
          const float gs = gv[i] * fs;
          const uint32_t pl = gbox + (uint32_t)(i * BOX_H * BOX_W * 4) + t_nw;
#ifdef CERB_SPLAT_X_NOATOMS   // CERB_SPLAT_X_*: timing-only elimination builds, results are wrong
          if (gs == 123.456f) red_shared_s32(pl, __float2int_rn(gs * w_nw + w_ne + w_sw + w_se));
#else
          red_shared_s32(pl, __float2int_rn(gs * w_nw));
          red_shared_s32(pl + 4, __float2int_rn(gs * w_ne));
          red_shared_s32(pl + BOX_W * 4, __float2int_rn(gs * w_sw));
          red_shared_s32(pl + BOX_W * 4 + 4, __float2int_rn(gs * w_se));
#endif
          const float v_nw = tv[i][0], v_ne = bx1 ? tv[i][1] : 0.f, v_sw = by1 ? tv[i][2] : 0.f;
          const float v_se = (bx1 && by1) ? tv[i][3] : 0.f;
          gix += gv[i] * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
          giy += gv[i] * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
        }
      }
      __syncthreads();
      // fixed point -> fp32 in place
#ifndef CERB_SPLAT_X_NOCONVERT
      for (uint32_t o = (uint32_t)tid * 16u; o < BOX_BYTES; o += NT * 16u) {
        const float4 q = lds128(gbox + o);
        sts128(gbox + o, make_float4((float)__float_as_int(q.x) * ifs, (float)__float_as_int(q.y) * ifs,
                                     (float)__float_as_int(q.z) * ifs, (float)__float_as_int(q.w) * ifs));
      }
#endif
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA's reads
      __syncthreads();
#ifndef CERB_SPLAT_X_NOREDUCE
      if (tid == 0) {
        tma_reduce_add_4d(&tm_gimg, gbox, box_ox, box_oy, c0, n2);
        tma_store_commit();
        tma_store_wait_read0();   // the window has been read: it may be cleared for the next pass
      }
#endif
      __syncthreads();
#ifndef CERB_SPLAT_X_NOZERO
      if (c0 + CB < a.C)
        for (uint32_t o = (uint32_t)tid * 16u; o < BOX_BYTES; o += NT * 16u) sts128(gbox + o, make_float4(0.f, 0.f, 0.f, 0.f));
#endif
    }
    if (tid == 0) tma_store_wait_all0();
  } else if (pix_ok) {
    const Taps tp = make_taps(sx, sy, a.H, a.W, a.in_s[2]);   // reads of img
    const Taps to = make_taps(sx, sy, a.H, a.W, a.W);         // splat into the contiguous gradient
    const float* ip = a.img + (long long)n2 * a.in_s[0];
    float* gb = a.gimg + (long long)n2 * a.C * plane;
    for (int c0 = 0; c0 < a.C; c0 += 4) {   // 4 channels per batch: 20 independent loads in flight
      float gv[4], v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = min(c0 + k, a.C - 1);
        gv[k] = __ldg(gop + (long long)c * plane);
        const float* p = ip + (long long)c * a.in_s[1];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[k][q] = __ldg(p + tp.off[q]);   // clamped taps: always valid addresses
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c0 + k < a.C) {
          float* gp = gb + (long long)(c0 + k) * plane;
          if (to.w[0] != 0.f) red_add(gp + to.off[0], gv[k] * to.w[0]);
          if (to.w[1] != 0.f) red_add(gp + to.off[1], gv[k] * to.w[1]);
          if (to.w[2] != 0.f) red_add(gp + to.off[2], gv[k] * to.w[2]);
          if (to.w[3] != 0.f) red_add(gp + to.off[3], gv[k] * to.w[3]);
          const float v_nw = v[k][0];
          const float v_ne = bx1 ? v[k][1] : 0.f;
          const float v_sw = by1 ? v[k][2] : 0.f;
          const float v_se = (bx1 && by1) ? v[k][3] : 0.f;
          gix += gv[k] * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
          giy += gv[k] * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
        }
      }
    }
  }
  if (pix_ok) {
    float* gf = a.gflow + (long long)n * 2 * plane + (long long)y * a.W + x;
    gf[0] = inx ? gix * pos_scale(a.W, a.mode) : 0.f;
    gf[plane] = iny ? giy * pos_scale(a.H, a.mode) : 0.f;
  }
}

}  // namespace splat

// fp32 flow_warp backward through shared-memory windows; cudaErrorNotSupported when the tensors cannot be TMA tensors
// (base or strides not 16-byte aligned) -- the caller then uses the scattered-atomics kernel.  gimg must be zeroed.
cudaError_t launch_flow_warp_backward_box(const float* img, const long long in_s[3], const float* flow, const long long f_s[3],
                                          const float* gout, float* gimg, float* gflow, int B, int C, int H, int W, int mode,
                                          int roll, cudaStream_t stream) {
  static const bool off = getenv("CERB_DEBUG_NO_SPLAT_BOX") != nullptr;
  if (off || B > 65535 || C < 1) return cudaErrorNotSupported;
  splat::Args a;
  a.img = img; a.flow = flow; a.gout = gout; a.gimg = gimg; a.gflow = gflow;
  for (int i = 0; i < 3; ++i) { a.in_s[i] = in_s[i]; a.f_s[i] = f_s[i]; }
  a.B = B; a.C = C; a.H = H; a.W = W; a.mode = mode; a.roll = roll;
  a.tiles_x = (W + splat::TX - 1) / splat::TX;
  a.tiles_y = (H + splat::TY - 1) / splat::TY;
  a.path_ctr = get_path_counters();
  CUtensorMap tmi, tmg;
  memset(&tmi, 0, sizeof(tmi));
  memset(&tmg, 0, sizeof(tmg));
  const long long gs[3] = {(long long)C * H * W, (long long)H * W, (long long)W};
  if (!make_tmap_nchw(&tmi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, img, W, H, C, B, in_s, splat::BOX_W, splat::BOX_H, splat::CB, false) ||
      !make_tmap_nchw(&tmg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gimg, W, H, C, B, gs, splat::BOX_W, splat::BOX_H, splat::CB, false))
    return cudaErrorNotSupported;
  auto kern = splat::flow_warp_bwd_box_kernel;
  static unsigned long long attr_devs = 0ull;   // function attributes are per device
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit == 0ull || !(attr_devs & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)splat::SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr_devs |= bit;
    }
  }
  dim3 grid(a.tiles_x, a.tiles_y, B);
  kern<<<grid, splat::NT, splat::SMEM_BYTES, stream>>>(a, tmi, tmg);
  return cudaGetLastError();
}

}  // namespace cerb
