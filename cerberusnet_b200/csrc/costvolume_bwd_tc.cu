// costvolume_bwd_tc.cu -- tensor-core (tcgen05 / TMEM) backward of the fused level op for sm_100a: both correlation
// gradients of a tile as banded GEMMs (the splat of the gradient with respect to the warped map: costvolume_splat.cu).
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
//     S by one "split" warp per gradient (hi = tf32(v) in place, lo = v - hi next to it); three products hi*hi + hi*lo + lo*hi
//     into partial accumulators summed in the drain -- with three of them the two products that share the band's hi part are
//     ONE MMA of N = 2 N_pad over the adjacent hi / lo operand tiles (an MMA costs ~70 cycles here whatever N is).
//   * Two 4-warp groups work side by side, one per gradient, each with its own accumulator, band ring and operand stages:
//     stage Gx of its tile (coalesced loads, LeakyReLU mask applied on the way, private shared-memory columns) -> build 16
//     band chunks -> stage the next tile's Gx -> drain the accumulators (tcgen05.ld) and store.  With a flow the second
//     gradient is the one with respect to the warped map: the launcher has it written to the workspace and runs the
//     shared-memory-window splat (costvolume_splat.cu) afterwards.  (A splat fused into the drain was built and measured --
//     its three windows left room for only two operand stages -- and removed.)
// Roles (16 warps): 0-3 gradient wrt x1, 4-7 gradient wrt the second input; per gradient one operand-split warp (8-9), one
// TMA-issue lane (10-11) and two MMA-issue lanes (12-15).  The two gradients' pipelines are independent of each other and
// run half a tile period apart.  One CTA per SM, persistent over tiles.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

bool make_tmap_nchw(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                    const long long strides[3], int bx, int by, int bc, bool swizzle128);
int num_sms_current();
long long* get_trace_buffer();

namespace btc {

constexpr int TY = 8, TX = 16, M = TY * TX;
constexpr int MD = 4, D = 9, D2 = 81;
constexpr int HY = TY + 2 * MD;                 // 16 halo rows = band chunks per tile and gradient
constexpr int KU = TX + 2 * MD;                 // 24 K columns per chunk (3 MMAs of K = 8)
constexpr int KBOX = 32;                        // TMA box width: one 128-byte swizzle row per channel
constexpr int GROUP_WARPS = 4, SPLIT_WARPS = 2, TMA_WARPS = 2, MMA_WARPS = 4;   // per gradient: one split, one TMA, two MMA-issue warps
static_assert(SPLIT_WARPS % 2 == 0, "split warps alternate between the two gradients");
constexpr int TMA_WARP = 2 * GROUP_WARPS + SPLIT_WARPS, MMA_WARP = TMA_WARP + TMA_WARPS;
constexpr int NTHREADS = (MMA_WARP + MMA_WARPS) * 32;   // 512: 16 warps (registers are allocated per 4 warps: 128 per thread)
constexpr int PADPL = 3;                        // zero planes before / after the 81 of a Gx buffer (shifted reads)
constexpr int GPLANES = D2 + 2 * PADPL;
constexpr uint32_t GS_BYTES = GPLANES * M * 4;  // 44544
constexpr uint32_t ZERO_BYTES = 12 * M * 4;     // twelve zero planes: what a lane reads for a halo row outside its 9 displacement rows
constexpr int MAXSTG = 6, MAXSLOT = 4;
constexpr int SLOT_COLS = 2 * KU;               // band slot in TMEM: 24 hi + 24 lo columns
constexpr int NBARS = 2 * (3 * MAXSTG + 2 * MAXSLOT + 2) + 1;   // + the groups' phase offset
struct Args {
  Geom g;
  const float* gout;
  const float* out;       // saved activated output (sign only), may be null
  float* gx1;             // contiguous NCHW
  float* gsecond;         // gradient wrt the second correlation input (contiguous NCHW): grad_x2 itself, or the workspace
  int tiles_x, tiles_y, total_tiles;
  int npad, nstg, nslot;  // N of the MMA, operand stages per gradient, band slots per gradient
  int nacc;               // partial accumulators per gradient (1..3)
  int no_cat;             // debugging: three separate MMAs per K step even with three accumulators
  int g2_roll;            // batch roll applied when writing gsecond (x2_batch_roll when it is grad_x2 itself)
  int s0_roll;            // batch roll applied to the TMA item coordinate of the first gradient's S (x2 itself when no flow)
  int use_pf;             // tensor maps tm_go / tm_o valid: the next tile's grad_out / out planes are prefetched into L2
  long long* dbg;         // optional per-CTA clock64() trace (cerb_debug_set_trace_buffer; -DCERB_BTC_TRACE), 192 slots per CTA
};

#ifdef CERB_BTC_TRACE
#define BT_TRACE(ti, s) do { if (a.dbg && (ti) < 4) a.dbg[(long long)blockIdx.x * 192 + 1 + (ti) * 40 + (s)] = clock64(); } while (0)
#else
#define BT_TRACE(ti, s) do { } while (0)
#endif

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
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
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
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {   // tile -> L2
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// TMEM lane (= MMA row) of tile position (py, px): a warp's 32 lanes are 8 rows x 4 columns
__device__ __forceinline__ int lane_of_pos(int py, int px) { return 32 * (px >> 2) + 4 * py + (px & 3); }
// index of lane t inside a 128-float plane of a Gx buffer: conflict-free both for the staging threads (2 rows x 16 columns
// per warp) and for the band threads (one lane quadrant per warp)
__device__ __forceinline__ int plane_idx(int t) { return (t & ~31) | ((t + 8 * (t >> 5)) & 31); }

__global__ void __launch_bounds__(NTHREADS, 1)
corr_bwd_tc_kernel(const Args a, const __grid_constant__ CUtensorMap tm_s0, const __grid_constant__ CUtensorMap tm_s1,
                   const __grid_constant__ CUtensorMap tm_go, const __grid_constant__ CUtensorMap tm_o) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const Geom& g = a.g;
  const int npad = a.npad, nstg = a.nstg, nslot = a.nslot, nacc = a.nacc;
  const uint32_t STG = (uint32_t)npad * 128u;                       // one operand tile (hi or lo) of a chunk
  const uint32_t OFF_GS = 2u * (uint32_t)nstg * 2u * STG;           // Gx buffers of the two groups
  const uint32_t OFF_ZERO = OFF_GS + 2u * GS_BYTES;
  const uint32_t OFF_BAR = OFF_ZERO + ZERO_BYTES;
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  // per gradient X: raw_full[MAXSTG] s_full[MAXSTG] s_empty[MAXSTG] band_full[MAXSLOT] band_empty[MAXSLOT] d_full d_empty
  auto BAR = [&](int X, int which, int i) -> uint64_t* {
    const int base = X * (3 * MAXSTG + 2 * MAXSLOT + 2);
    const int off = which < 3 ? which * MAXSTG : 3 * MAXSTG + (which - 3) * MAXSLOT;   // which 5 / 6: d_full / d_empty
    return bars + base + (which < 5 ? off + i : 3 * MAXSTG + 2 * MAXSLOT + (which - 5));
  };
  enum { RAW_FULL = 0, S_FULL = 1, S_EMPTY = 2, BAND_FULL = 3, BAND_EMPTY = 4, D_FULL = 5, D_EMPTY = 6 };
  uint64_t* offset_bar = bars + NBARS - 1;
  uint32_t* tmem_slot = (uint32_t*)(smem + OFF_BAR + NBARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int X = 0; X < 2; ++X) {
      for (int s = 0; s < MAXSTG; ++s) {
        mbar_init(BAR(X, RAW_FULL, s), 1);
        mbar_init(BAR(X, S_FULL, s), 1);
        mbar_init(BAR(X, S_EMPTY, s), min(nacc, 2));
      }
      for (int s = 0; s < MAXSLOT; ++s) {
        mbar_init(BAR(X, BAND_FULL, s), GROUP_WARPS);
        mbar_init(BAR(X, BAND_EMPTY, s), min(nacc, 2));
      }
      mbar_init(BAR(X, D_FULL, 0), min(nacc, 2));
      mbar_init(BAR(X, D_EMPTY, 0), GROUP_WARPS);
    }
    mbar_init(offset_bar, GROUP_WARPS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the pad planes of the Gx buffers are read (and discarded) by the shifted band loads: keep them finite
  for (int i = tid; i < (int)((2 * GS_BYTES + ZERO_BYTES) / 4); i += NTHREADS) reinterpret_cast<float*>(smem + OFF_GS)[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t RING0 = 2u * (uint32_t)(nacc * npad);   // TMEM columns: 2 x nacc accumulators of npad columns, then the band rings
  if (warp < GROUP_WARPS) {
    // band columns a lane never writes must read as zero: clear both rings once (warp w owns lane quadrant w)
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16) + RING0;
    for (int c = 0; c < 2 * nslot * SLOT_COLS; c += 4) tmem_st4(tl + (uint32_t)c, 0.f, 0.f, 0.f, 0.f);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

#ifdef CERB_BTC_TRACE
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * 192] = clock64();
#endif
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
      const int os1 = (int)g.os[1], os2 = (int)g.os[2];   // (the launcher checks that an item's 81 planes fit 31 bits)
      // Entry e = (ey, ex) reads plane e at the position itself (X == 0) or plane (8 - ey, 8 - ex) at position q + e - 4
      // (X == 1).  The address splits into a row part and a column part; loads are unconditional from clamped (always
      // valid) addresses and validity is applied afterwards (costvolume_bwd.cu: predicated loads serialise).
      int xo[D];
      unsigned colok = 0, rowok = 0;
#pragma unroll
      for (int e = 0; e < D; ++e) {
        const int px = X == 0 ? qx : qx + e - MD, py = X == 0 ? qy : qy + e - MD;
        xo[e] = (X == 0 ? e : -e) * os1 + min(max(px, 0), g.outW - 1);
        if (px >= 0 && px < g.outW) colok |= 1u << e;
        if (py >= 0 && py < g.outH) rowok |= 1u << e;
      }
#pragma unroll 1
      for (int ey0 = 0; ey0 < D; ey0 += 3) {
        float gv[3 * D], ov[3 * D];
#pragma unroll
        for (int r3 = 0; r3 < 3; ++r3) {
          const int ey = ey0 + r3;
          const int py = X == 0 ? qy : qy + ey - MD;
          const int rb = (X == 0 ? ey * D : (D - 1 - ey) * D + D - 1) * os1 + min(max(py, 0), g.outH - 1) * os2;
#pragma unroll
          for (int ex = 0; ex < D; ++ex) {
#ifdef CERB_BTC_X_NOSTAGE
            gv[r3 * D + ex] = __int_as_float(rb + xo[ex]);
            ov[r3 * D + ex] = __int_as_float(rb - xo[ex]);
#else
            gv[r3 * D + ex] = __ldcg(go + (rb + xo[ex]));
            ov[r3 * D + ex] = __ldcg(oo + (rb + xo[ex]));
#endif
          }
        }
#pragma unroll
        for (int r = 0; r < 3 * D; ++r) {
          const bool ok = ((rowok >> (ey0 + r / D)) & (colok >> (r % D)) & 1u) != 0u;
          float v = ok ? gv[r] : 0.f;
          if (mask && !(ov[r] > 0.f)) v *= slope;
          Gs[(PADPL + ey0 * D + r) * M + sidx] = v;
        }
      }
    };

    int ti = 0;
    if ((int)blockIdx.x < a.total_tiles) stage_g(blockIdx.x);
    named_bar_sync(bar_id, 128);
#ifndef CERB_BTC_NO_OFFSET
    // The two groups run half a period apart: while one builds (and the tensor pipe works on its gradient) the other
    // stages / drains -- in lockstep both builds share the pipe (20.7 K cycles of MMA per tile pair at ~72 per MMA) and
    // both then wait on memory at the same time.  The second group starts building when the first has built its first tile.
    if (X == 1) wait_bar(offset_bar, 0u);
#endif
    // band ring: slot index and phase parity kept incrementally, barrier / column bases hoisted (the chunk loop is bound by
    // instruction issue: ~230 instructions per chunk and thread)
    uint64_t* const band_full0 = BAR(X, BAND_FULL, 0);
    uint64_t* const band_empty0 = BAR(X, BAND_EMPTY, 0);
    const uint32_t col0 = tlane + RING0 + (uint32_t)(X * nslot * SLOT_COLS + 4 * wq);
    const uint32_t zero_u32 = sbase + OFF_ZERO + 4u * (uint32_t)bidx;
    int slot = 0;
    uint32_t bphase = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      if ((tid & 127) == 0) BT_TRACE(ti, X * 12 + 0);
      // ---- 16 band chunks: chunk r = halo row r; this lane's entries are Gx[(r - bpy, k - res)] at columns 4 wq + k
#pragma unroll 1
      for (int r = 0; r < HY; ++r) {
        const int ey = r - bpy;
        const bool rowv = ey >= 0 && ey < D;
        // a halo row outside this lane's 9 displacement rows reads the zero planes instead (one select on the address)
        const uint32_t base = rowv ? gs_u32 + 4u * (uint32_t)((PADPL + ey * D - res) * M + bidx) : zero_u32;
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = lds_f32(base + (uint32_t)k * (M * 4));
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          if (k < 3) v[k] = res <= k ? v[k] : 0.f;             // 0 <= k - res <= 8: columns 3..8 are always inside
          if (k > 8) v[k] = res >= k - 8 ? v[k] : 0.f;
        }
        float hi[12], lo[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) split_tf32(v[k], hi[k], lo[k]);
        wait_bar(band_empty0 + slot, bphase ^ 1u);
        tc_fence_after();
        const uint32_t col = col0 + (uint32_t)(slot * SLOT_COLS);
#pragma unroll
        for (int k = 0; k < 12; k += 4) {
          tmem_st4(col + (uint32_t)k, hi[k], hi[k + 1], hi[k + 2], hi[k + 3]);
          tmem_st4(col + (uint32_t)(KU + k), lo[k], lo[k + 1], lo[k + 2], lo[k + 3]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(band_full0 + slot);
        if ((tid & 127) == 0 && r == 3) BT_TRACE(ti, X * 12 + 1);
        if (++slot == nslot) { slot = 0; bphase ^= 1u; }
      }
      if ((tid & 127) == 0) BT_TRACE(ti, X * 12 + 2);
#ifndef CERB_BTC_NO_OFFSET
      if (X == 0 && ti == 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(offset_bar);
      }
#endif
      named_bar_sync(bar_id, 128);   // every warp of the group is done reading the Gx buffer
      int n, y0, x0;
      decode(tile, n, y0, x0);
      const int y = y0 + bpy, x = x0 + bpx;
      const bool pix_ok = y < g.H && x < g.W;
      const long long plane = (long long)g.H * g.W;
      const uint32_t dcol = tlane + (uint32_t)(X * nacc * npad);
      // 8 channels of this lane's position: the partial accumulators summed
      auto ld_acc8 = [&](int c0, float* v) {
        tmem_ld8(dcol + (uint32_t)c0, v);
        for (int j = 1; j < nacc; ++j) {
          float w[8];
          tmem_ld8(dcol + (uint32_t)(j * npad + c0), w);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += w[i];
        }
      };
      if ((tid & 127) == 0) BT_TRACE(ti, X * 12 + 3);
      // ---- the next tile's Gx while this tile's last MMAs run
      if (tile + (int)gridDim.x < a.total_tiles) stage_g(tile + gridDim.x);
      if ((tid & 127) == 0) BT_TRACE(ti, X * 12 + 4);

      // ---- drain: TMEM lane = position, column = channel
      wait_bar(BAR(X, D_FULL, 0), (uint32_t)(ti & 1));
      tc_fence_after();
      if ((tid & 127) == 0) BT_TRACE(ti, X * 12 + 5);
      auto release_d = [&]() {   // every TMEM read of this warp has landed: (with the other three) the next tile's MMAs may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(X, D_EMPTY, 0));
      };
      {
        int n_dst = X == 0 ? n : n + a.g2_roll;
        if (n_dst >= g.B) n_dst -= g.B;
        float* dst = (X == 0 ? a.gx1 : a.gsecond) + (long long)n_dst * g.C * plane + (long long)min(y, g.H - 1) * g.W + min(x, g.W - 1);
        for (int c0 = 0; c0 < npad; c0 += 8) {
          float v[8];
          ld_acc8(c0, v);
          if (c0 + 8 >= npad) release_d();
          if (pix_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c0 + i < g.C) dst[(long long)(c0 + i) * plane] = v[i] * inv_c;
          }
        }
      }
      if ((tid & 127) == 0) BT_TRACE(ti, X * 12 + 6);
      // the staged Gx of the next tile is complete in every warp of the group before anyone builds from it
      named_bar_sync(bar_id, 128);
    }
    if (X == 1 && gt == 0) tma_store_wait_all0();   // the last gradient windows have been added to grad_x2
  } else if (warp < TMA_WARP) {
    // =========================== operand split: hi = tf32(v) in place, lo = v - hi ===========================
    // One WARP per operand tile (split warp sw takes gradient sw % 2, every (SPLIT_WARPS / 2)-th row): the tiles of a row pair
    // are split concurrently -- with all warps on one tile the chain wait -> 3 loads -> stores -> fence -> arrive of a
    // single tile (~600 cycles) was the period of the whole operand pipeline (clock64 trace).
    const int sw = warp - 2 * GROUP_WARPS;
    const int X = sw & 1, rstep = SPLIT_WARPS / 2, r0 = sw >> 1;
    const int units = npad * (KU / 4);   // 16-byte chunks that the MMAs read (columns 0..23 of every channel row)
    int cnt = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      for (int r = 0; r < HY; ++r, ++cnt) {
        if ((r % rstep) != r0) continue;
        const int stg = cnt % nstg, use = cnt / nstg;
        const uint32_t hi_t = sbase + (uint32_t)((X * nstg + stg) * 2) * STG;
        wait_bar(BAR(X, RAW_FULL, stg), (uint32_t)(use & 1));
        if (lane == 0 && (r == 0 || r == 15)) BT_TRACE(cnt / HY, 36 + X + (r == 15 ? 2 : 0));
        for (int u0 = lane; u0 < units; u0 += 4 * 32) {   // four independent loads in flight
          uint32_t ad[4];
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int u = min(u0 + 32 * k, units - 1);
            const int row = u / (KU / 4), c = u - row * (KU / 4);
            ad[k] = hi_t + (uint32_t)row * 128u + ((uint32_t)(c ^ (row & 7)) << 4);
            v[k] = lds128(ad[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (u0 + 32 * k < units) {
              float4 h, l;
              split_tf32(v[k].x, h.x, l.x); split_tf32(v[k].y, h.y, l.y); split_tf32(v[k].z, h.z, l.z); split_tf32(v[k].w, h.w, l.w);
              sts128(ad[k], h);
              sts128(ad[k] + STG, l);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(X, S_FULL, stg));
      }
    }
  } else if (warp >= MMA_WARP) {
    // =========================== MMA issue: one lane of up to three warps per gradient ===========================
    // A tcgen05.mma of N <= 128 costs the issuing thread ~140 cycles (clock64 trace, elimination builds: the same for N = 16
    // and N = 48, halved by a second issuing warp), several times what the tensor pipe needs for it.  Warp (X, j) issues the
    // products of the 3xTF32 split that accumulate into partial accumulator j of gradient X -- hi*hi, lo*hi, hi*lo with
    // three accumulators -- and commits onto the shared barriers (counts = nacc).
    // (16 warps in all: two issuing warps per gradient -- one for hi*hi, one for both cross terms)
    const int mw = warp - MMA_WARP, X = mw >> 1, jw = mw & 1;
    // accumulators this warp issues into
    const uint32_t accmask = nacc == 3 ? (jw == 0 ? 1u : 6u) : (jw < nacc ? 1u << jw : 0u);
    if (lane == 0 && accmask != 0u) {
      // instruction descriptor: D fp32 (bit 4), A / B tf32 (2 at bits 7-9 / 10-12), K-major, N = npad, M = 128
#ifdef CERB_BTC_X_N16
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#else
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#endif
      const uint32_t idesc_cat = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * npad) >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const bool cat = nacc == 3 && !a.no_cat;
      // which products this warp issues, and whether a product is the first to write the accumulator in a tile
      uint32_t d_col[3], first_acc[3];
      bool mine[3];
#pragma unroll
      for (int prod = 0; prod < 3; ++prod) {
        const int j = nacc == 1 ? 0 : (prod < nacc ? prod : prod - nacc);
        mine[prod] = ((accmask >> j) & 1u) != 0u;
        d_col[prod] = tmem + (uint32_t)((X * nacc + j) * npad);
        first_acc[prod] = prod < nacc ? 0u : 1u;
      }
      int stg = 0, slot = 0;
      uint32_t sphase = 0, bphase = 0;   // phase parities of the operand-stage / band-slot rings
      int ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        wait_bar(BAR(X, D_EMPTY, 0), (uint32_t)((ti & 1) ^ 1));   // the previous tile's accumulators have been drained
        for (int r = 0; r < HY; ++r) {
          wait_bar(BAR(X, S_FULL, stg), sphase);
          wait_bar(BAR(X, BAND_FULL, slot), bphase);
          tc_fence_after();
          if (jw == 0 && r == 0) BT_TRACE(ti, 24 + X);
          if (jw == 0 && r == 8) BT_TRACE(ti, 26 + X);
          if (jw == 0 && r == 15) BT_TRACE(ti, 28 + X);
          const uint32_t a_hi = tmem + RING0 + (uint32_t)((X * nslot + slot) * SLOT_COLS);
          const uint64_t b_hi = make_desc(sbase + (uint32_t)((X * nstg + stg) * 2) * STG);
          const uint64_t b_lo = make_desc(sbase + (uint32_t)((X * nstg + stg) * 2 + 1) * STG);
          if (cat) {
            // Three partial accumulators per gradient, laid out [hi*hi | hi*lo | lo*hi]: the two products that share the band's
            // hi part are ONE MMA of N = 2 npad -- the hi and lo operand tiles of a stage are adjacent in shared memory, so one
            // descriptor covers both -- and an MMA costs the tensor pipe about the same ~72 cycles for any N <= 128 here
            // (elimination builds: N = 16 and N = 48 take the same time).  Two MMAs per K step instead of three.
#if !defined(CERB_BTC_X_NOMMA)
#pragma unroll
            for (int ks = 0; ks < KU / 8; ++ks) {
              const uint64_t ko = (uint64_t)(2 * ks);
              const uint32_t acc = (r > 0 || ks > 0) ? 1u : 0u;
              if (jw == 0) umma_tf32_ta(tmem + (uint32_t)(X * 3 * npad), a_hi + 8u * ks, b_hi + ko, idesc_cat, acc);
              else umma_tf32_ta(tmem + (uint32_t)((X * 3 + 2) * npad), a_hi + KU + 8u * ks, b_hi + ko, idesc, acc);
            }
#endif
          } else
#pragma unroll
          for (int ks = 0; ks < KU / 8; ++ks) {
            const uint64_t ko = (uint64_t)(2 * ks);   // 32 bytes per K step inside the 128-byte swizzle atom
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
              const uint32_t acc = (r > 0 || ks > 0) ? 1u : first_acc[prod];
              const uint32_t a_t = (prod == 1 ? a_hi + KU : a_hi) + 8u * ks;     // hi*hi, lo*hi, hi*lo
#if defined(CERB_BTC_X_NOMMA)   // CERB_BTC_X_*: timing-only elimination builds (tools/ab_variants.py), results are wrong
              (void)acc; (void)a_t;
#else
              if (mine[prod]) umma_tf32_ta(d_col[prod], a_t, (prod == 2 ? b_lo : b_hi) + ko, idesc, acc);
#endif
            }
          }
          umma_commit(BAR(X, BAND_EMPTY, slot));
          umma_commit(BAR(X, S_EMPTY, stg));
          if (r == HY - 1) umma_commit(BAR(X, D_FULL, 0));
          if (jw == 0 && r == HY - 1) BT_TRACE(ti, 30 + X);
          if (++stg == nstg) { stg = 0; sphase ^= 1u; }
          if (++slot == nslot) { slot = 0; bphase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp < MMA_WARP) {
    // =========================== TMA issue: one lane of one warp per gradient ===========================
    if (lane == 0) {
      const int X = warp - TMA_WARP;
      const CUtensorMap* tm = X == 0 ? &tm_s0 : &tm_s1;
      const int roll = X == 0 ? a.s0_roll : 0;
      int cnt = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        int n, y0, x0;
        decode(tile, n, y0, x0);
        int n0 = n + roll;
        if (n0 >= g.B) n0 -= g.B;
        // the next tile's inputs into L2 while this one is processed: its grad_out / out planes (incl. the 4-pixel halo the
        // second gradient reads) and its operand rows
        if (tile + (int)gridDim.x < a.total_tiles) {
          int nn, ny0, nx0;
          decode(tile + gridDim.x, nn, ny0, nx0);
          int nn0 = nn + roll;
          if (nn0 >= g.B) nn0 -= g.B;
          if (a.use_pf && X == 0) {
            tma_prefetch_4d(&tm_go, nx0 - MD, ny0 - MD, 0, nn);
            if (g.has_act && a.out != nullptr) tma_prefetch_4d(&tm_o, nx0 - MD, ny0 - MD, 0, nn);
          }
          for (int r = 0; r < HY; ++r) tma_prefetch_4d(tm, nx0 - MD, ny0 - MD + r, 0, nn0);
        }
        for (int r = 0; r < HY; ++r, ++cnt) {
          const int stg = cnt % nstg, use = cnt / nstg;
          wait_bar(BAR(X, S_EMPTY, stg), (uint32_t)((use & 1) ^ 1));
          if (r == 0 || r == 15) BT_TRACE(cnt / HY, 32 + X + (r == 15 ? 2 : 0));
          mbar_arrive_expect_tx(BAR(X, RAW_FULL, stg), STG);
          tma_load_4d(smem + (uint32_t)((X * nstg + stg) * 2) * STG, tm, BAR(X, RAW_FULL, stg), x0 - MD, y0 - MD + r, 0, n0);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
#ifdef CERB_BTC_TRACE
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * 192 + 191] = clock64();
#endif
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
  if (81 * g.os[1] + (long long)g.H * g.os[2] >= (1ll << 31)) return false;   // 32-bit offsets inside one item of grad_out
  if (((uintptr_t)s0 & 15) || ((uintptr_t)x1 & 15)) return false;
  for (int i = 0; i < 3; ++i)
    if ((s0s[i] * 4) % 16 != 0 || (g.x1s[i] * 4) % 16 != 0) return false;
  return true;
}

// gsecond: contiguous NCHW destination of the gradient with respect to the second correlation input -- grad_x2 itself
// (gsecond_roll = x2_batch_roll) or, with a flow, the workspace the splat kernel reads (roll 0)
cudaError_t launch_corr_backward_tc(const Geom& g, const void* s0, const long long s0s[3], int s0_roll, const void* x1,
                                    const void* out, const void* gout, void* gx1, void* gsecond, int gsecond_roll,
                                    cudaStream_t stream) {
  btc::Args a;
  a.g = g;
  a.gout = (const float*)gout; a.out = (const float*)out;
  a.gx1 = (float*)gx1; a.gsecond = (float*)gsecond;
  a.tiles_x = (g.W + btc::TX - 1) / btc::TX;
  a.tiles_y = (g.H + btc::TY - 1) / btc::TY;
  a.total_tiles = g.B * a.tiles_x * a.tiles_y;
  a.npad = (g.C + 15) / 16 * 16;
  a.s0_roll = s0_roll;
  a.g2_roll = gsecond_roll;
  // TMEM (512 columns): 2 x nacc accumulators of npad columns + 2 band rings of nslot x 48; as many partial accumulators
  // as leave two band slots per gradient
  static const int force_nacc = getenv("CERB_DEBUG_BWD_TC_NACC") ? atoi(getenv("CERB_DEBUG_BWD_TC_NACC")) : 0;
  a.nacc = 3;
  while (a.nacc > 1 && (512 - 2 * a.nacc * a.npad) / (2 * btc::SLOT_COLS) < 2) --a.nacc;
  if (force_nacc >= 1 && force_nacc < a.nacc) a.nacc = force_nacc;
  a.nslot = (512 - 2 * a.nacc * a.npad) / (2 * btc::SLOT_COLS);
  if (a.nslot > btc::MAXSLOT) a.nslot = btc::MAXSLOT;
  static const bool no_cat = getenv("CERB_DEBUG_BWD_TC_NOCAT") != nullptr;
  a.no_cat = no_cat ? 1 : 0;
  a.dbg = get_trace_buffer();
  CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  memset(&tm1, 0, sizeof(tm1));
  if (!make_tmap_nchw(&tm0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s0, g.W, g.H, g.C, g.B, s0s, btc::KBOX, 1, a.npad, true) ||
      !make_tmap_nchw(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x1, g.W, g.H, g.C, g.B, g.x1s, btc::KBOX, 1, a.npad, true))
    return cudaErrorNotSupported;
  // shared memory: the two Gx buffers, barriers, and as many operand stages (hi + lo tile) per gradient as fit
  const size_t fixed = 2 * (size_t)btc::GS_BYTES + btc::ZERO_BYTES + btc::NBARS * 8 + 16 + 1024;
  const size_t stage = 2 * (size_t)a.npad * 128;
  CUtensorMap tmgo, tmo;
  memset(&tmgo, 0, sizeof(tmgo));
  memset(&tmo, 0, sizeof(tmo));
  static const bool no_pf = getenv("CERB_DEBUG_BWD_TC_NOPF") != nullptr;
  a.use_pf = !no_pf &&
             make_tmap_nchw(&tmgo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gout, g.outW, g.outH, g.D2, g.B, g.os, btc::KU, btc::HY, g.D2, false) &&
             (out == nullptr ||
              make_tmap_nchw(&tmo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, g.outW, g.outH, g.D2, g.B, g.os, btc::KU, btc::HY, g.D2, false));
  a.nstg = (int)((227 * 1024 - fixed) / (2 * stage));
  if (a.nstg > btc::MAXSTG) a.nstg = btc::MAXSTG;
  if (a.nslot < 2 || a.nstg < 2) return cudaErrorNotSupported;
  const size_t smem = 2 * (size_t)a.nstg * stage + fixed;
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
  kern<<<grid, btc::NTHREADS, smem, stream>>>(a, tm0, tm1, tmgo, tmo);
  return cudaGetLastError();
}

}  // namespace cerb
