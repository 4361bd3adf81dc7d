// costvolume_bwd_tc.cu -- tensor-core (tcgen05 / TMEM) backward of the fused level op for sm_100a: both correlation
// gradients of a tile as banded GEMMs, the gradient with respect to the warped map splatted in the drain.
//
// Replaces the same reference sites as costvolume_bwd.cu (correlation_backward_input1/2 correlation_cuda_kernel.cu:97-242,
// launch loop :326-429; LeakyReluBackward of pwcnet_sfd.py:182; GridSampler2DBackward of UnFlowLoss.py:83-94) for fp32,
// kernel 1, strides 1, max_displacement 4 = pad, C <= 128.
//
// Formulation.  Both gradients are   R[c, t] = 1/C * sum_{e in [0,8]^2}  Gx[e, t] * S[c, t + e - 4]   over an 8 x 16 tile of
// positions t (costvolume_bwd.cu states the two (Gx, S) pairs).  With the 16 x 24 halo of the tile as the contraction index
// this is one GEMM per tile and gradient:  R[t, c] = sum_h Band[t, h] * S[h, c],  Band[t, t + e] = Gx[e, t], zero elsewhere
// (21 % of the band matrix is populated; the tensor core has the headroom, the FMA pipe -- 8 TFLOP/s in the CUDA-core
// kernel -- does not).  M = 128 positions, N = C rounded up to 16, K = 16 halo rows x 24 columns, fp32 accumulation in TMEM.
//   * Band is the A operand and lives in TENSOR MEMORY: a chunk = one halo row = 24 K columns (3 kind::tf32 MMAs of K = 8).
//     TMEM lane t belongs to one thread; lanes are assigned to tile positions so that a warp's 32 lanes cover 8 rows x 4
//     columns: the 9 non-zero entries of a lane's band row then start at column 4*quad + (x & 3), every warp writes only
//     12 of the 24 columns (the others stay zero from the start of the kernel) and the per-lane shift is 0..3 -- it is done
//     by the load address (shared-memory column of the lane, plane k - (x & 3)), six of the twelve values need a select.
//   * S is the B operand: one TMA box (32 x 1 x N, SWIZZLE_128B) per halo row lands as a K-major operand tile, zero
//     outside the image for free.  fp32 parity (1e-5 of max|ref|) needs 3xTF32: Gx is split while the band row is built,
//     S by four "split" warps (hi = tf32(v) in place, lo = v - hi next to it); three products hi*hi + lo*hi + hi*lo.
//   * Two 4-warp groups work side by side, one per gradient, each with its own accumulator, band ring and operand stages:
//     stage Gx of its tile (coalesced loads, LeakyReLU mask applied on the way, private shared-memory columns) -> build 16
//     band chunks -> stage the next tile's Gx -> drain the accumulator (tcgen05.ld): grad_x1 is stored, the gradient with
//     respect to the second input is either stored (no flow) or pushed through the four bilinear taps of its position into
//     grad_x2 (red.global.add) with the flow gradient accumulated over the channels.
// Roles: warps 0-3 gradient wrt x1, 4-7 gradient wrt the second input, 8-11 operand split, 12 MMA issue (one lane),
// 13 TMA issue (one lane).  One CTA per SM, persistent over tiles.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

bool make_tmap_nchw(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                    const long long strides[3], int bx, int by, int bc, bool swizzle128);
int num_sms_current();

namespace btc {

constexpr int TY = 8, TX = 16, M = TY * TX;
constexpr int MD = 4, D = 9, D2 = 81;
constexpr int HY = TY + 2 * MD;                 // 16 halo rows = band chunks per tile and gradient
constexpr int KU = TX + 2 * MD;                 // 24 K columns per chunk (3 MMAs of K = 8)
constexpr int KBOX = 32;                        // TMA box width: one 128-byte swizzle row per channel
constexpr int GROUP_WARPS = 4, SPLIT_WARPS = 4;
constexpr int MMA_WARP = 2 * GROUP_WARPS + SPLIT_WARPS, TMA_WARP = MMA_WARP + 1;
constexpr int NTHREADS = (TMA_WARP + 1) * 32;   // 448
constexpr int PADPL = 3;                        // zero planes before / after the 81 of a Gx buffer (shifted reads)
constexpr int GPLANES = D2 + 2 * PADPL;
constexpr uint32_t GS_BYTES = GPLANES * M * 4;  // 44544
constexpr int MAXSTG = 4, MAXSLOT = 4;
constexpr int SLOT_COLS = 2 * KU;               // band slot in TMEM: 24 hi + 24 lo columns
constexpr int NBARS = 2 * (3 * MAXSTG + 2 * MAXSLOT + 2);

struct Args {
  Geom g;
  const float* gout;
  const float* out;       // saved activated output (sign only), may be null
  float* gx1;             // contiguous NCHW
  float* gsecond;         // no flow: gradient wrt x2 (contiguous NCHW, item order of x2)
  const float* x2;        // with a flow: un-warped second map (tap values for the flow gradient)
  const float* flow;
  float* gx2;             // with a flow: splat target (zeroed by the launcher)
  float* gflow;
  int tiles_x, tiles_y, total_tiles;
  int npad, nstg, nslot;  // N of the MMA, operand stages per gradient, band slots per gradient
  int s0_roll;            // batch roll applied to the TMA item coordinate of the first gradient's S (x2 itself when no flow)
};

// ---- tcgen05 wrappers (same encodings as costvolume_fwd_tc.cu, validated in tools/microbench/umma_probe.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, 8-row atoms 1024 bytes apart
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {   // hi = tf32(v) (round to nearest), lo = v - hi (exact)
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  lo = v - hi;
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) { mbar_wait_hint(bar, parity, 20000); }
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// TMEM lane (= MMA row) of tile position (py, px): a warp's 32 lanes are 8 rows x 4 columns
__device__ __forceinline__ int lane_of_pos(int py, int px) { return 32 * (px >> 2) + 4 * py + (px & 3); }
// index of lane t inside a 128-float plane of a Gx buffer: conflict-free both for the staging threads (2 rows x 16 columns
// per warp) and for the band threads (one lane quadrant per warp)
__device__ __forceinline__ int plane_idx(int t) { return (t & ~31) | ((t + 8 * (t >> 5)) & 31); }

__global__ void __launch_bounds__(NTHREADS, 1)
corr_bwd_tc_kernel(const Args a, const __grid_constant__ CUtensorMap tm_s0, const __grid_constant__ CUtensorMap tm_s1) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const Geom& g = a.g;
  const int npad = a.npad, nstg = a.nstg, nslot = a.nslot;
  const uint32_t STG = (uint32_t)npad * 128u;                       // one operand tile (hi or lo) of a chunk
  const uint32_t OFF_GS = 2u * (uint32_t)nstg * 2u * STG;           // Gx buffers of the two groups
  const uint32_t OFF_BAR = OFF_GS + 2u * GS_BYTES;
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  // per gradient X: raw_full[MAXSTG] s_full[MAXSTG] s_empty[MAXSTG] band_full[MAXSLOT] band_empty[MAXSLOT] d_full d_empty
  auto BAR = [&](int X, int which, int i) -> uint64_t* {
    const int base = X * (3 * MAXSTG + 2 * MAXSLOT + 2);
    const int off = which < 3 ? which * MAXSTG : 3 * MAXSTG + (which - 3) * MAXSLOT;   // which 5 / 6: d_full / d_empty
    return bars + base + (which < 5 ? off + i : 3 * MAXSTG + 2 * MAXSLOT + (which - 5));
  };
  enum { RAW_FULL = 0, S_FULL = 1, S_EMPTY = 2, BAND_FULL = 3, BAND_EMPTY = 4, D_FULL = 5, D_EMPTY = 6 };
  uint32_t* tmem_slot = (uint32_t*)(smem + OFF_BAR + NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int X = 0; X < 2; ++X) {
      for (int s = 0; s < MAXSTG; ++s) {
        mbar_init(BAR(X, RAW_FULL, s), 1);
        mbar_init(BAR(X, S_FULL, s), SPLIT_WARPS);
        mbar_init(BAR(X, S_EMPTY, s), 1);
      }
      for (int s = 0; s < MAXSLOT; ++s) {
        mbar_init(BAR(X, BAND_FULL, s), GROUP_WARPS);
        mbar_init(BAR(X, BAND_EMPTY, s), 1);
      }
      mbar_init(BAR(X, D_FULL, 0), 1);
      mbar_init(BAR(X, D_EMPTY, 0), GROUP_WARPS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_s0);
    tma_prefetch_desc(&tm_s1);
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the pad planes of the Gx buffers are read (and discarded) by the shifted band loads: keep them finite
  for (int i = tid; i < 2 * (int)(GS_BYTES / 4); i += NTHREADS) reinterpret_cast<float*>(smem + OFF_GS)[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t RING0 = 2u * (uint32_t)npad;   // TMEM columns: accumulators [0, npad) and [npad, 2 npad), then the band rings
  if (warp < GROUP_WARPS) {
    // band columns a lane never writes must read as zero: clear both rings once (warp w owns lane quadrant w)
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16) + RING0;
    for (int c = 0; c < 2 * nslot * SLOT_COLS; c += 4) tmem_st4(tl + (uint32_t)c, 0.f, 0.f, 0.f, 0.f);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int per_img = a.tiles_x * a.tiles_y;
  auto decode = [&](int tile, int& n, int& y0, int& x0) {
    n = tile / per_img;
    const int tr = tile - n * per_img;
    y0 = (tr / a.tiles_x) * TY;
    x0 = (tr % a.tiles_x) * TX;
  };

  if (warp < 2 * GROUP_WARPS) {
    // =========================== one gradient per group of four warps ===========================
    const int X = warp / GROUP_WARPS;                 // 0: gradient wrt x1, 1: gradient wrt the second correlation input
    const int wq = warp & 3, gt = tid & 127;          // lane quadrant, thread of the group
    float* Gs = reinterpret_cast<float*>(smem + OFF_GS + (uint32_t)X * GS_BYTES);
    const uint32_t gs_u32 = sbase + OFF_GS + (uint32_t)X * GS_BYTES;
    const int bar_id = 1 + X;
    // staging role: position (spy, spx) of the tile, coalesced rows of 16 pixels
    const int spy = gt >> 4, spx = gt & 15;
    const int sidx = plane_idx(lane_of_pos(spy, spx));
    // band / drain role: TMEM lane t = 32 wq + lane <-> position (bpy, bpx)
    const int bpy = lane >> 2, bpx = 4 * wq + (lane & 3), res = lane & 3;
    const int bidx = plane_idx(32 * wq + lane);
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
    const bool mask = g.has_act && a.out != nullptr;
    const float slope = g.slope;
    const float inv_c = 1.0f / (float)g.C;

    // ---- Gx of a tile into this group's buffer: plane e = ey * 9 + ex, thread-private columns
    auto stage_g = [&](int tile) {
      int n, y0, x0;
      decode(tile, n, y0, x0);
      const float* go = a.gout + (long long)n * g.os[0];
      const float* oo = mask ? a.out + (long long)n * g.os[0] : go;
      const int qy = y0 + spy, qx = x0 + spx;
#pragma unroll 1
      for (int ey0 = 0; ey0 < D; ey0 += 3) {
        float gv[3 * D], ov[3 * D];
        // unconditional loads from clamped addresses, validity applied afterwards (costvolume_bwd.cu: predicated loads serialise)
#pragma unroll
        for (int r = 0; r < 3 * D; ++r) {
          const int ey = ey0 + r / D, ex = r % D;
          const int py = X == 0 ? qy : qy + ey - MD, px = X == 0 ? qx : qx + ex - MD;
          const int d = X == 0 ? ey * D + ex : (D - 1 - ey) * D + (D - 1 - ex);
          const long long o = (long long)d * g.os[1] + (long long)min(max(py, 0), g.outH - 1) * g.os[2] + min(max(px, 0), g.outW - 1);
          gv[r] = __ldcg(go + o);
          ov[r] = __ldcg(oo + o);
        }
#pragma unroll
        for (int r = 0; r < 3 * D; ++r) {
          const int ey = ey0 + r / D, ex = r % D;
          const int py = X == 0 ? qy : qy + ey - MD, px = X == 0 ? qx : qx + ex - MD;
          const bool ok = py >= 0 && py < g.outH && px >= 0 && px < g.outW;
          float v = ok ? gv[r] : 0.f;
          if (mask && !(ov[r] > 0.f)) v *= slope;
          Gs[(PADPL + ey0 * D + r) * M + sidx] = v;
        }
      }
    };

    int cnt = 0;   // band chunks built so far (all tiles)
    int ti = 0;
    if ((int)blockIdx.x < a.total_tiles) stage_g(blockIdx.x);
    named_bar_sync(bar_id, 128);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      // ---- 16 band chunks: chunk r = halo row r; this lane's entries are Gx[(r - bpy, k - res)] at columns 4 wq + k
#pragma unroll 1
      for (int r = 0; r < HY; ++r, ++cnt) {
        const int slot = cnt % nslot, use = cnt / nslot;
        const int ey = r - bpy;
        const bool rowv = ey >= 0 && ey < D;
        const int eyc = min(max(ey, 0), D - 1);
        const uint32_t base = gs_u32 + 4u * (uint32_t)((PADPL + eyc * D - res) * M + bidx);
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = lds_f32(base + (uint32_t)k * (M * 4));
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          const bool colv = k >= 3 && k <= 8 ? true : (k < 3 ? res <= k : res >= k - 8);   // 0 <= k - res <= 8
          v[k] = (rowv && colv) ? v[k] : 0.f;
        }
        float hi[12], lo[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) split_tf32(v[k], hi[k], lo[k]);
        wait_bar(BAR(X, BAND_EMPTY, slot), (uint32_t)((use & 1) ^ 1));
        tc_fence_after();
        const uint32_t col = tlane + RING0 + (uint32_t)((X * nslot + slot) * SLOT_COLS + 4 * wq);
#pragma unroll
        for (int k = 0; k < 12; k += 4) {
          tmem_st4(col + (uint32_t)k, hi[k], hi[k + 1], hi[k + 2], hi[k + 3]);
          tmem_st4(col + (uint32_t)(KU + k), lo[k], lo[k + 1], lo[k + 2], lo[k + 3]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(X, BAND_FULL, slot));
      }
      // ---- the next tile's Gx while this tile's last MMAs run (every warp of the group is done reading the buffer)
      named_bar_sync(bar_id, 128);
      if (tile + (int)gridDim.x < a.total_tiles) stage_g(tile + gridDim.x);

      // ---- drain: TMEM lane = position, column = channel
      int n, y0, x0;
      decode(tile, n, y0, x0);
      const int y = y0 + bpy, x = x0 + bpx;
      const bool pix_ok = y < g.H && x < g.W;
      const long long plane = (long long)g.H * g.W;
      const uint32_t dcol = tlane + (uint32_t)(X * npad);
      wait_bar(BAR(X, D_FULL, 0), (uint32_t)(ti & 1));
      tc_fence_after();
      if (X == 0 || a.flow == nullptr) {
        const int n_dst = X == 0 ? n : x2_item(g, n);
        float* dst = (X == 0 ? a.gx1 : a.gsecond) + (long long)n_dst * g.C * plane + (long long)min(y, g.H - 1) * g.W + min(x, g.W - 1);
        for (int c0 = 0; c0 < npad; c0 += 8) {
          float v[8];
          tmem_ld8(dcol + (uint32_t)c0, v);
          if (pix_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c0 + i < g.C) dst[(long long)(c0 + i) * plane] = v[i] * inv_c;
          }
        }
      } else {
        // gradient wrt the warped map -> through the bilinear taps of this position into grad_x2; flow gradient over channels
        const int yc = min(y, g.H - 1), xc = min(x, g.W - 1);
        const float* fp = a.flow + (long long)n * g.fls[0] + (long long)yc * g.fls[2] + xc;
        bool inx, iny;
        const float sx = sample_pos(xc, __ldg(fp), g.W, g.warp_mode, inx);
        const float sy = sample_pos(yc, __ldg(fp + g.fls[1]), g.H, g.warp_mode, iny);
        const Taps tp_in = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
        const Taps tp_out = make_taps(sx, sy, g.H, g.W, g.W);
        const float fx = floorf(sx), fy = floorf(sy);
        const float wx1 = fx + 1.f - sx, wx0 = sx - fx, wy1 = fy + 1.f - sy, wy0 = sy - fy;
        const bool bx1 = (int)fx + 1 < g.W, by1 = (int)fy + 1 < g.H;
        const int n_x2 = x2_item(g, n);
        const float* x2n = a.x2 + (long long)n_x2 * g.x2s[0];
        float* gx2n = a.gx2 + (long long)n_x2 * g.C * plane;
        float gix = 0.f, giy = 0.f;
        for (int c0 = 0; c0 < npad; c0 += 8) {
          float v[8];
          tmem_ld8(dcol + (uint32_t)c0, v);
          if (pix_ok) {
            float tv[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float* pch = x2n + (long long)min(c0 + i, g.C - 1) * g.x2s[1];
#pragma unroll
              for (int q = 0; q < 4; ++q) tv[i][q] = __ldg(pch + tp_in.off[q]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (c0 + i < g.C) {
                const float gv = v[i] * inv_c;
                float* gp = gx2n + (long long)(c0 + i) * plane;
                if (tp_out.w[0] != 0.f) red_add(gp + tp_out.off[0], gv * tp_out.w[0]);
                if (tp_out.w[1] != 0.f) red_add(gp + tp_out.off[1], gv * tp_out.w[1]);
                if (tp_out.w[2] != 0.f) red_add(gp + tp_out.off[2], gv * tp_out.w[2]);
                if (tp_out.w[3] != 0.f) red_add(gp + tp_out.off[3], gv * tp_out.w[3]);
                const float v_nw = tv[i][0];
                const float v_ne = bx1 ? tv[i][1] : 0.f;
                const float v_sw = by1 ? tv[i][2] : 0.f;
                const float v_se = (bx1 && by1) ? tv[i][3] : 0.f;
                gix += gv * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
                giy += gv * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
              }
            }
          }
        }
        if (pix_ok) {
          float* gf = a.gflow + (long long)n * 2 * plane + (long long)y * g.W + x;
          gf[0] = inx ? gix * pos_scale(g.W, g.warp_mode) : 0.f;
          gf[plane] = iny ? giy * pos_scale(g.H, g.warp_mode) : 0.f;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(X, D_EMPTY, 0));
      // the staged Gx of the next tile is complete in every warp of the group before anyone builds from it
      named_bar_sync(bar_id, 128);
    }
  } else if (warp < MMA_WARP) {
    // =========================== operand split: hi = tf32(v) in place, lo = v - hi ===========================
    const int st = tid - 2 * GROUP_WARPS * 32;
    const int units = npad * (KU / 4);   // 16-byte chunks that the MMAs read (columns 0..23 of every channel row)
    int cnt = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      for (int r = 0; r < HY; ++r, ++cnt) {
        const int stg = cnt % nstg, use = cnt / nstg;
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          const uint32_t hi_t = sbase + (uint32_t)((X * nstg + stg) * 2) * STG;
          wait_bar(BAR(X, RAW_FULL, stg), (uint32_t)(use & 1));
          for (int u = st; u < units; u += SPLIT_WARPS * 32) {
            const int row = u / (KU / 4), c = u - row * (KU / 4);
            const uint32_t ad = hi_t + (uint32_t)row * 128u + ((uint32_t)(c ^ (row & 7)) << 4);
            const float4 v = lds128(ad);
            float4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            sts128(ad, h);
            sts128(ad + STG, l);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(X, S_FULL, stg));
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue: one lane ===========================
    if (lane == 0) {
      // instruction descriptor: D fp32 (bit 4), A / B tf32 (2 at bits 7-9 / 10-12), K-major, N = npad, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      int cnt = 0, ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        for (int r = 0; r < HY; ++r, ++cnt) {
          const int stg = cnt % nstg, suse = cnt / nstg;
          const int slot = cnt % nslot, buse = cnt / nslot;
#pragma unroll
          for (int X = 0; X < 2; ++X) {
            if (r == 0) {   // the previous tile's accumulator of this gradient has been drained
              wait_bar(BAR(X, D_EMPTY, 0), (uint32_t)((ti & 1) ^ 1));
            }
            wait_bar(BAR(X, S_FULL, stg), (uint32_t)(suse & 1));
            wait_bar(BAR(X, BAND_FULL, slot), (uint32_t)(buse & 1));
            tc_fence_after();
            const uint32_t d_t = tmem + (uint32_t)(X * npad);
            const uint32_t a_hi = tmem + RING0 + (uint32_t)((X * nslot + slot) * SLOT_COLS), a_lo = a_hi + KU;
            const uint64_t b_hi = make_desc(sbase + (uint32_t)((X * nstg + stg) * 2) * STG);
            const uint64_t b_lo = make_desc(sbase + (uint32_t)((X * nstg + stg) * 2 + 1) * STG);
#pragma unroll
            for (int ks = 0; ks < KU / 8; ++ks) {
              const uint64_t ko = (uint64_t)(2 * ks);   // 32 bytes per K step inside the 128-byte swizzle atom
              umma_tf32_ta(d_t, a_hi + 8u * ks, b_hi + ko, idesc, (r > 0 || ks > 0) ? 1u : 0u);
              umma_tf32_ta(d_t, a_lo + 8u * ks, b_hi + ko, idesc, 1u);
              umma_tf32_ta(d_t, a_hi + 8u * ks, b_lo + ko, idesc, 1u);
            }
            umma_commit(BAR(X, BAND_EMPTY, slot));
            umma_commit(BAR(X, S_EMPTY, stg));
            if (r == HY - 1) umma_commit(BAR(X, D_FULL, 0));
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== TMA issue: one lane ===========================
    if (lane == 0) {
      int cnt = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        int n, y0, x0;
        decode(tile, n, y0, x0);
        int n0 = n + a.s0_roll;
        if (n0 >= g.B) n0 -= g.B;
        for (int r = 0; r < HY; ++r, ++cnt) {
          const int stg = cnt % nstg, use = cnt / nstg;
#pragma unroll
          for (int X = 0; X < 2; ++X) {
            wait_bar(BAR(X, S_EMPTY, stg), (uint32_t)((use & 1) ^ 1));
            mbar_arrive_expect_tx(BAR(X, RAW_FULL, stg), STG);
            tma_load_4d(smem + (uint32_t)((X * nstg + stg) * 2) * STG, X == 0 ? &tm_s0 : &tm_s1, BAR(X, RAW_FULL, stg), x0 - MD,
                        y0 - MD + r, 0, X == 0 ? n0 : n);
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace btc

// s0: the first gradient's "other" operand (warped second map, or x2 itself without a flow), element strides N, C, H
bool tc_backward_supported(const Geom& g, int dtype, const void* s0, const long long s0s[3], const void* x1) {
  if (dtype != CERB_F32 || g.k != 1 || g.s1 != 1 || g.s2 != 1 || g.md != btc::MD || g.pad != g.md) return false;
  if (g.C > 128 || g.C < 1 || g.outH != g.H || g.outW != g.W) return false;
  if (((uintptr_t)s0 & 15) || ((uintptr_t)x1 & 15)) return false;
  for (int i = 0; i < 3; ++i)
    if ((s0s[i] * 4) % 16 != 0 || (g.x1s[i] * 4) % 16 != 0) return false;
  return true;
}

cudaError_t launch_corr_backward_tc(const Geom& g, const void* s0, const long long s0s[3], int s0_roll, const void* x1,
                                    const void* x2, const float* flow, const void* out, const void* gout, void* gx1,
                                    void* gsecond, void* gx2_splat, float* gflow, cudaStream_t stream) {
  btc::Args a;
  a.g = g;
  a.gout = (const float*)gout; a.out = (const float*)out;
  a.gx1 = (float*)gx1; a.gsecond = (float*)gsecond;
  a.x2 = (const float*)x2; a.flow = flow; a.gx2 = (float*)gx2_splat; a.gflow = gflow;
  a.tiles_x = (g.W + btc::TX - 1) / btc::TX;
  a.tiles_y = (g.H + btc::TY - 1) / btc::TY;
  a.total_tiles = g.B * a.tiles_x * a.tiles_y;
  a.npad = (g.C + 15) / 16 * 16;
  a.s0_roll = s0_roll;
  // TMEM: 2 accumulators of npad columns + 2 band rings of nslot x 48; shared memory: 2 x nstg stages of (hi, lo) tiles
  a.nslot = (512 - 2 * a.npad) / (2 * btc::SLOT_COLS);
  if (a.nslot > btc::MAXSLOT) a.nslot = btc::MAXSLOT;
  const size_t fixed = 2 * (size_t)btc::GS_BYTES + btc::NBARS * 8 + 16 + 1024;
  const size_t stage = 2 * (size_t)a.npad * 128;
  a.nstg = (int)((227 * 1024 - fixed) / (2 * stage));
  if (a.nstg > btc::MAXSTG) a.nstg = btc::MAXSTG;
  if (a.nslot < 2 || a.nstg < 2) return cudaErrorNotSupported;
  const size_t smem = 2 * (size_t)a.nstg * stage + fixed;
  CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  memset(&tm1, 0, sizeof(tm1));
  if (!make_tmap_nchw(&tm0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s0, g.W, g.H, g.C, g.B, s0s, btc::KBOX, 1, a.npad, true) ||
      !make_tmap_nchw(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x1, g.W, g.H, g.C, g.B, g.x1s, btc::KBOX, 1, a.npad, true))
    return cudaErrorNotSupported;
  auto kern = btc::corr_bwd_tc_kernel;
  static unsigned long long attr_devs = 0ull;   // function attributes are per device
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit == 0ull || !(attr_devs & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
      attr_devs |= bit;
    }
  }
  int grid = num_sms_current();
  if (grid > a.total_tiles) grid = a.total_tiles;
  kern<<<grid, btc::NTHREADS, smem, stream>>>(a, tm0, tm1);
  return cudaGetLastError();
}

}  // namespace cerb
