// costvolume_fwd_tc.cu -- tensor-core (tcgen05 / TMEM) variant of the fused flow-warp + correlation + LeakyReLU
// forward for sm_100a: the banded contraction as one Gram tile per output tile.
//
// Same reference sites as costvolume_fwd.cu (flow_warp UnFlowLoss.py:83-94, correlation_forward
// correlation_cuda_kernel.cu:29-95, leaky_relu pwcnet_sfd.py:182); max_displacement 4, kernel 1, strides 1.
//
// Why: on the CUDA-core kernel every x2 element crosses shared memory twice and the consumers' LDS traffic alone is
// as expensive as their FFMAs (ncu: LSU/shared pipe the limiter, FMA pipe 21 %).  Here the contraction leaves both
// the FMA pipe and the LSU: for an 8 x 16 output tile (M = 128 pixels) and its 16 x 24 halo of the warped second
// map (N = 384 positions) the tensor core computes G[p][q] = sum_c x1[c][p] * x2w[c][q] into TMEM (384 fp32 columns);
// the 81 displacements of pixel p are the entries q = p + (dy, dx) of its row.  21 % of the Gram tile is used; the
// MMA time is still below what the FFMA formulation needs.
//   fp32 inputs:  "3xTF32" -- every operand is split hi = tf32(v), lo = v - hi and three products (hi*hi, lo*hi,
//                 hi*lo) accumulate in fp32: measured 4e-7 of max|ref| (tools/microbench/umma_probe.cu), inside the
//                 1e-5 parity bar.  A single TF32 product (3e-4) is not.
// Roles (18 warps, one CTA per SM, persistent over tiles):
//   warps 0-3   thread = pixel.  Stage the x1 tile K-major (hi / lo) per K step of 8 channels; after the tile's last MMA
//               drain TMEM: warp w owns TMEM lanes 32w..32w+31 (pixel rows 2w, 2w+1), loads the 10 halo rows it needs
//               24 columns at a time (tcgen05.ld 32x32b), shifts the row by its own x with a 4-stage select network
//               (the column offset differs per lane, tcgen05.ld's does not), divides, activates, stores.
//   warps 4-15  thread = halo position.  Flow -> sample position -> 4 taps + weights once per tile; per K step the taps
//               of 8 channels from the raw x2 box in shared memory (or straight from global memory when the tile's
//               footprint does not fit the box), blend in ATen's order, split, four 16-byte stores into the K-major
//               128B-swizzled operand rows.  The warped map never exists in HBM.
//   warp 16     one lane issues tcgen05.mma.kind::tf32 (M 128, N 192 x 2, K 8): 6 per K step; tcgen05.commit releases
//               the operand slot / publishes the accumulator.
//   warp 17     one lane issues the raw-box TMA loads (box origin from the tile's tap bounding box).
// Operand slots: a 128-byte operand row holds 32 channels = 4 K steps; slot j of every row is refilled as soon as the
// MMAs that read it have retired, so staging, MMA and the previous tile's drain overlap.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

// host helpers defined in costvolume_fwd.cu
bool make_tmap_nchw(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                    const long long strides[3], int bx, int by, int bc, bool swizzle128);
int num_sms_current();
long long* get_trace_buffer();
unsigned long long* get_path_counters();

namespace tc {

constexpr int TY = 8, TX = 16, M = TY * TX;           // output tile, MMA M
constexpr int MD = 4;
constexpr int HY = TY + 2 * MD, HX = TX + 2 * MD;     // 16 x 24 halo of the second map
constexpr int N = HY * HX, NH = N / 2;                // 384 accumulator columns, two MMAs of N = 192
constexpr int MARGIN = 6;                             // flow variation (px) inside one halo the raw box absorbs
constexpr int RAW_H = HY + 2 * MARGIN + 2;            // 30
constexpr int RAW_W = (HX + 2 * MARGIN + 2 + 3 + 3) / 4 * 4;   // 44 (box starts are 16-byte aligned)
constexpr int KC = 8;                                 // channels per K step (32 bytes of a K-major row = one tf32 MMA)
constexpr int SLOTS = 4;                              // K steps per 128-byte operand row
constexpr int RS = 2;                                 // raw boxes in flight
constexpr int EPI_WARPS = 4, GATHER_WARPS = 12;
constexpr int GATHER_THREADS = GATHER_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS, TMA_WARP = MMA_WARP + 1;
constexpr int NTHREADS = (TMA_WARP + 1) * 32;         // 576
constexpr uint32_t A_BYTES = M * 128, B_BYTES = N * 128;
constexpr uint32_t OFF_AHI = 0, OFF_ALO = A_BYTES, OFF_BHI = 2 * A_BYTES, OFF_BLO = 2 * A_BYTES + B_BYTES;
constexpr uint32_t OFF_RAW = 2 * A_BYTES + 2 * B_BYTES;
constexpr uint32_t RAW_PLANE = RAW_H * RAW_W * 4;
constexpr uint32_t RAW_STAGE = KC * RAW_PLANE;
constexpr uint32_t OFF_RED = OFF_RAW + RS * RAW_STAGE;
constexpr uint32_t OFF_BAR = OFF_RED + 2 * GATHER_WARPS * 4 * 4;
constexpr int NBARS = 3 * SLOTS + 2 * RS + 2 + 4;
constexpr uint32_t OFF_TMEM = OFF_BAR + NBARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(RAW_STAGE % 128 == 0 && OFF_RAW % 1024 == 0, "TMA destination alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(N == GATHER_THREADS && M == EPI_WARPS * 32, "one halo position / one pixel per thread");

enum { PATH_RAW = 1, PATH_DIRECT = 2 };   // indices match costvolume_fwd.cu's path counters

struct Args {
  Geom g;
  const void* x1;
  const void* x2;
  const float* flow;
  void* out;
  int off;              // md - pad
  int tiles_x, tiles_y, total_tiles;
  int nks;              // K steps per tile: ceil(C / 8)
  int use_raw;          // raw x2 boxes by TMA
  unsigned long long* path_ctr;
  long long* dbg;       // optional per-CTA clock64() trace (cerb_debug_set_trace_buffer), 192 slots per CTA
};
// trace slot `s` of tile iteration `ti` (first four tiles of a CTA; 40 slots each)
#define TC_TRACE(ti, s) do { if (a.dbg && (ti) < 4) a.dbg[(long long)blockIdx.x * 192 + 1 + (ti) * 40 + (s)] = clock64(); } while (0)

// ---- tcgen05 wrappers
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, 8-row atoms 1024 bytes apart
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives once every MMA issued so far has retired
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for the loads into v[0..23]; the registers pass through the statement so that no use can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait24(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
  asm volatile("" : "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) : : "memory");
  asm volatile("" : "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]) : : "memory");
}
__device__ __forceinline__ int slot_of(int kc) { return kc & 3; }

// hi = tf32(v) (round to nearest, low 13 mantissa bits zero), lo = v - hi (exact)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  lo = v - hi;
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 1)
warp_corr_fwd_tc_kernel(const Args a, const __grid_constant__ CUtensorMap tm_raw) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  int* red = (int*)(smem + OFF_RED);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* a_full = bars;                    // [SLOTS] x1 slot written (4 warps)
  uint64_t* b_full = bars + SLOTS;            // [SLOTS] warped-x2 slot written (12 warps)
  uint64_t* slot_empty = bars + 2 * SLOTS;    // [SLOTS] MMAs reading the slot have retired (tcgen05.commit)
  uint64_t* raw_full = bars + 3 * SLOTS;      // [RS]
  uint64_t* raw_empty = raw_full + RS;        // [RS]
  uint64_t* d_full = raw_empty + RS;          // accumulator of the tile complete
  uint64_t* d_empty = d_full + 1;             // accumulator drained (4 warps)
  uint64_t* bbox_full = d_empty + 1;          // [2] every gather warp has published its share of the tile's tap bounding box
  uint64_t* bbox_empty = bbox_full + 2;       // [2] the TMA warp has read it (the slot may be rewritten two tiles later)
  uint32_t* tmem_slot = (uint32_t*)(smem + OFF_TMEM);

  const Geom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* __restrict__ x1 = (const T*)a.x1;
  const T* __restrict__ x2 = (const T*)a.x2;

  if (tid == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&a_full[s], EPI_WARPS);
      mbar_init(&b_full[s], GATHER_WARPS);
      mbar_init(&slot_empty[s], 1);
    }
    for (int s = 0; s < RS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], GATHER_WARPS);
    }
    mbar_init(d_full, 1);
    mbar_init(d_empty, EPI_WARPS);
    mbar_init(&bbox_full[0], GATHER_WARPS);
    mbar_init(&bbox_full[1], GATHER_WARPS);
    mbar_init(&bbox_empty[0], 1);
    mbar_init(&bbox_empty[1], 1);
    fence_barrier_init();
    if (a.use_raw) tma_prefetch_desc(&tm_raw);
  }
  if (warp == MMA_WARP) {   // 512 columns: the 384 of the accumulator rounded up to a power of two
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * 192] = clock64();
  // PDL: everything above overlapped the previous kernel's tail; its outputs may be our inputs
  pdl_wait();

  const int per_img = a.tiles_x * a.tiles_y;
  const int nks = a.nks;

  if (warp < EPI_WARPS) {
    // =========================== x1 staging + accumulator drain: thread = pixel ===========================
    const int py = tid >> 4, px = tid & 15;
    const uint32_t arow = (uint32_t)tid * 128u, sw = (uint32_t)(tid & 7);
    const float fC = (float)g.C, rC = __frcp_rn((float)g.C);
    const bool c_pow2 = (g.C & (g.C - 1)) == 0;   // 1/C exact: the division is one multiply
    const long long os1 = g.os[1];
    // x1 values are requested a group of four K steps (32 channels) at a time, in one burst right after the previous
    // group has been staged: the next tile's first group is in flight while this tile drains.  (Requests interleaved
    // with the consumption of older ones made every use wait for the youngest load: scoreboards count, they do not
    // track individual loads.)
    float nx[SLOTS][KC];
    auto request_group = [&](int tile, int ks0) {
      const int n = tile / per_img, trem = tile - n * per_img;
      const int iy = (trem / a.tiles_x) * TY + a.off + py, ix = (trem % a.tiles_x) * TX + a.off + px;
      const bool inimg = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const T* p = x1 + (long long)n * g.x1s[0] + (long long)min(max(iy, 0), g.H - 1) * g.x1s[2] + min(max(ix, 0), g.W - 1);
#pragma unroll
      for (int b = 0; b < SLOTS; ++b)
#pragma unroll
        for (int c = 0; c < KC; ++c) {   // unconditional loads from clamped addresses, validity applied afterwards
          const int ch = (ks0 + b) * KC + c;
          const float v = ldg_f32(p + (long long)min(ch, g.C - 1) * g.x1s[1]);
          nx[b][c] = (inimg && ch < g.C) ? v : 0.f;
        }
    };
    int kc = 0, ti = 0;
    if ((int)blockIdx.x < a.total_tiles) request_group(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      const bool has_next = tile + (int)gridDim.x < a.total_tiles;
      for (int ks4 = 0; ks4 < nks; ks4 += SLOTS) {
#pragma unroll
        for (int b = 0; b < SLOTS; ++b) {
          const int ks = ks4 + b;
          if (ks < nks) {
            float hi[KC], lo[KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) split_tf32(nx[b][c], hi[c], lo[c]);
            const int slot = kc & (SLOTS - 1), use = kc >> 2;
            mbar_wait(&slot_empty[slot], (uint32_t)((use & 1) ^ 1));
            const uint32_t c0 = (((uint32_t)(2 * slot)) ^ sw) << 4, c1 = (((uint32_t)(2 * slot + 1)) ^ sw) << 4;
            sts128(sbase + OFF_AHI + arow + c0, make_float4(hi[0], hi[1], hi[2], hi[3]));
            sts128(sbase + OFF_AHI + arow + c1, make_float4(hi[4], hi[5], hi[6], hi[7]));
            sts128(sbase + OFF_ALO + arow + c0, make_float4(lo[0], lo[1], lo[2], lo[3]));
            sts128(sbase + OFF_ALO + arow + c1, make_float4(lo[4], lo[5], lo[6], lo[7]));
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[slot]);
            if (tid == 0 && ks < 4) TC_TRACE(ti, 26 + ks);
            ++kc;
          }
        }
        if (ks4 + SLOTS < nks) request_group(tile, ks4 + SLOTS);
        else if (has_next) request_group(tile + gridDim.x, 0);
      }
      // ---- drain: TMEM lane = pixel, column = halo position; this lane's displacements are rows py..py+8, columns px..px+8
      const int n = tile / per_img, trem = tile - n * per_img;
      const int oy = (trem / a.tiles_x) * TY + py, ox = (trem % a.tiles_x) * TX + px;
      const bool pix_ok = oy < g.outH && ox < g.outW;
      T* op = (T*)a.out + (long long)n * g.os[0] + (long long)min(oy, g.outH - 1) * g.os[2] + min(ox, g.outW - 1);
      mbar_wait(d_full, (uint32_t)(ti & 1));
      tc_fence_after();
      if (tid == 0) TC_TRACE(ti, 30);
      const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int rr = 0; rr < 10; ++rr) {
        uint32_t v[24];
        const uint32_t col = (uint32_t)((2 * warp + rr) * HX);
        tmem_ld8(tlane + col, v);
        tmem_ld8(tlane + col + 8, v + 8);
        tmem_ld8(tlane + col + 16, v + 16);
        tmem_ld_wait24(v);
        if (rr == 9) {   // every load of the tile has landed: the next tile's MMAs may overwrite the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty);
          if (tid == 0) TC_TRACE(ti, 31);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (px & 8) ? v[i + 8] : v[i];
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = (px & 4) ? v[i + 4] : v[i];
#pragma unroll
        for (int i = 0; i < 10; ++i) v[i] = (px & 2) ? v[i + 2] : v[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = (px & 1) ? v[i + 1] : v[i];
        const int dy = rr - (py & 1);   // halo row 2w + rr is displacement row (2w + rr) - py of this pixel
        if (pix_ok && dy >= 0 && dy <= 2 * MD) {
          T* orow = op + (long long)(dy * (2 * MD + 1)) * os1;
#pragma unroll
          for (int dx = 0; dx < 2 * MD + 1; ++dx) {
            float r = c_pow2 ? __fmul_rn(__uint_as_float(v[dx]), rC) : div_const(__uint_as_float(v[dx]), fC, rC);
            if (g.has_act) r = leaky(r, g.slope);
            *orow = from_f32<T>(r);
            orow += os1;
          }
        }
      }
      if (tid == 0) TC_TRACE(ti, 32);
    }
  } else if (warp < MMA_WARP) {
    // =========================== warped second map: thread = halo position ===========================
    const int gt = tid - EPI_WARPS * 32, gw = warp - EPI_WARPS;
    const int hy = gt / HX, hx = gt - hy * HX;
    const uint32_t brow = (uint32_t)gt * 128u, sw = (uint32_t)(gt & 7);
    const bool warped = a.flow != nullptr;
    const AxisConst axis_x = make_axis(g.W), axis_y = make_axis(g.H);
    // sampling data of one tile for this thread's halo position
    struct Pos {
      int x0, x1c, y0, y1c;
      float w[4];
      bool valid;
    };
    // halo position of a tile -> clamped pixel (the flow vector is read there), inside-the-image flag
    auto locate = [&](int tile, int& n, int& cy, int& cx) -> bool {
      n = tile / per_img;
      const int trem = tile - n * per_img;
      const int qy = (trem / a.tiles_x) * TY + a.off - g.md + hy, qx = (trem % a.tiles_x) * TX + a.off - g.md + hx;
      cy = min(max(qy, 0), g.H - 1); cx = min(max(qx, 0), g.W - 1);
      return qy >= 0 && qy < g.H && qx >= 0 && qx < g.W;   // outside: the correlation's zero padding
    };
    auto load_flow = [&](int tile, float& fu, float& fv) {
      int n, cy, cx;
      locate(tile, n, cy, cx);
      const float* fp = a.flow + (long long)n * g.fls[0] + (long long)cy * g.fls[2] + cx;
      fu = __ldg(fp); fv = __ldg(fp + g.fls[1]);
    };
    // taps of the tile and this warp's share of their bounding box (published for tile iteration `it`)
    auto prepare = [&](int tile, int it, float fu, float fv) -> Pos {
      int n, cy, cx;
      Pos p;
      p.valid = locate(tile, n, cy, cx);
      float sx = (float)cx, sy = (float)cy;   // un-warped: the pixel itself (weights 1, 0, 0, 0)
      if (warped) {
        bool in_x, in_y;
        sx = sample_pos(cx, fu, axis_x, g.warp_mode, in_x);
        sy = sample_pos(cy, fv, axis_y, g.warp_mode, in_y);
      }
      const Taps tp = make_taps(sx, sy, g.H, g.W, 0);   // hstride 0: off[] = {x0, x1c, x0, x1c}
      p.x0 = tp.off[0]; p.x1c = tp.off[1];
      p.y0 = (int)floorf(sy);
      p.y1c = (p.y0 + 1 < g.H) ? p.y0 + 1 : p.y0;
#pragma unroll
      for (int k = 0; k < 4; ++k) p.w[k] = tp.w[k];
      if (a.use_raw) {
        int xmin = p.valid ? p.x0 : 0x7fffffff, xmax = p.valid ? p.x1c : -0x7fffffff;
        int ymin = p.valid ? p.y0 : 0x7fffffff, ymax = p.valid ? p.y1c : -0x7fffffff;
        xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
        ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
        if (lane == 0) {
          if (it >= 2) mbar_wait(&bbox_empty[it & 1], (uint32_t)(((it >> 1) - 1) & 1));
          int* rp = red + (it & 1) * (GATHER_WARPS * 4) + gw * 4;
          *reinterpret_cast<int4*>(rp) = make_int4(xmin, xmax, ymin, ymax);
          mbar_arrive(&bbox_full[it & 1]);
        }
      }
      return p;
    };
    int kc = 0, rc = 0, ti = 0;
    Pos cur, nxt;
    {
      float fu = 0.f, fv = 0.f;
      if ((int)blockIdx.x < a.total_tiles) {
        if (warped) load_flow(blockIdx.x, fu, fv);
        cur = prepare(blockIdx.x, 0, fu, fv);
      }
    }
    nxt = cur;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      if (gt == 0) TC_TRACE(ti, 0);
      const int n = tile / per_img;
      const bool has_next = tile + (int)gridDim.x < a.total_tiles;
      float nfu = 0.f, nfv = 0.f;   // the next tile's flow vector: requested now, used after the first K step
      if (has_next && warped) load_flow(tile + gridDim.x, nfu, nfv);
      // ---- does the tile's sampling footprint fit the raw box?
      int path = PATH_DIRECT, ox = 0, oy = 0;
      if (a.use_raw) {
        mbar_wait(&bbox_full[ti & 1], (uint32_t)((ti >> 1) & 1));
        const int4* rp = reinterpret_cast<const int4*>(red + (ti & 1) * (GATHER_WARPS * 4));
        int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
        for (int w = 0; w < GATHER_WARPS; ++w) {
          const int4 b = rp[w];
          xmin = min(xmin, b.x); xmax = max(xmax, b.y);
          ymin = min(ymin, b.z); ymax = max(ymax, b.w);
        }
        ox = xmin & ~3;   // TMA box starts must be 16-byte aligned
        oy = ymin;
        if (xmin <= xmax && xmax - ox < RAW_W && ymax - oy < RAW_H) path = PATH_RAW;
      }
      if (gt == 0) TC_TRACE(ti, 2);
      if (gt == 0 && a.path_ctr != nullptr) atomicAdd(&a.path_ctr[path], 1ull);
      const bool valid = cur.valid;
      Taps tp;
#pragma unroll
      for (int k = 0; k < 4; ++k) tp.w[k] = cur.w[k];
      // the operand slot of this K step is free and written: publish it
      auto stage_b = [&](const float (&hi)[KC], const float (&lo)[KC], int ks) {
        const int slot = kc & (SLOTS - 1), use = kc >> 2;
        mbar_wait(&slot_empty[slot], (uint32_t)((use & 1) ^ 1));
        if (gt == 0 && ks < 4) TC_TRACE(ti, 4 + 3 * ks);
        const uint32_t c0 = (((uint32_t)(2 * slot)) ^ sw) << 4, c1 = (((uint32_t)(2 * slot + 1)) ^ sw) << 4;
        sts128(sbase + OFF_BHI + brow + c0, make_float4(hi[0], hi[1], hi[2], hi[3]));
        sts128(sbase + OFF_BHI + brow + c1, make_float4(hi[4], hi[5], hi[6], hi[7]));
        sts128(sbase + OFF_BLO + brow + c0, make_float4(lo[0], lo[1], lo[2], lo[3]));
        sts128(sbase + OFF_BLO + brow + c1, make_float4(lo[4], lo[5], lo[6], lo[7]));
        fence_proxy_async_smem();
        __syncwarp();
      };
      if (path == PATH_RAW) {
        // byte offsets of the four taps inside a channel plane of the box (positions without a sample read offset 0)
        const int r0 = (cur.y0 - oy) * RAW_W - ox, r1 = (cur.y1c - oy) * RAW_W - ox;
        const uint32_t t0 = valid ? (uint32_t)(r0 + cur.x0) << 2 : 0u, t1 = valid ? (uint32_t)(r0 + cur.x1c) << 2 : 0u;
        const uint32_t t2 = valid ? (uint32_t)(r1 + cur.x0) << 2 : 0u, t3 = valid ? (uint32_t)(r1 + cur.x1c) << 2 : 0u;
        for (int ks = 0; ks < nks; ++ks, ++kc, ++rc) {
          const int rs = rc & (RS - 1), ruse = rc / RS;
          mbar_wait(&raw_full[rs], (uint32_t)(ruse & 1));
          if (gt == 0 && ks < 4) TC_TRACE(ti, 3 + 3 * ks);
          const uint32_t rb = sbase + OFF_RAW + (uint32_t)rs * RAW_STAGE;
          float tv[KC][4];
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            tv[c][0] = lds_f32(rb + t0 + (uint32_t)c * RAW_PLANE);
            tv[c][1] = lds_f32(rb + t1 + (uint32_t)c * RAW_PLANE);
            tv[c][2] = lds_f32(rb + t2 + (uint32_t)c * RAW_PLANE);
            tv[c][3] = lds_f32(rb + t3 + (uint32_t)c * RAW_PLANE);
          }
          float hi[KC], lo[KC];
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            const float r = valid ? blend(tv[c][0], tv[c][1], tv[c][2], tv[c][3], tp) : 0.f;
            split_tf32(r, hi[c], lo[c]);
          }
          stage_b(hi, lo, ks);
          if (lane == 0) {
            mbar_arrive(&raw_empty[rs]);
            mbar_arrive(&b_full[slot_of(kc)]);
          }
          if (gt == 0 && ks < 4) TC_TRACE(ti, 5 + 3 * ks);
          if (ks == 0 && has_next) nxt = prepare(tile + gridDim.x, ti + 1, nfu, nfv);
        }
      } else {
        // ---- fallback: taps straight from global memory (any alignment / flow)
        const T* x2n = x2 + (long long)x2_item(g, n) * g.x2s[0];
        const long long r0 = (long long)cur.y0 * g.x2s[2], r1 = (long long)cur.y1c * g.x2s[2];
        const long long t0 = valid ? r0 + cur.x0 : 0, t1 = valid ? r0 + cur.x1c : 0;
        const long long t2 = valid ? r1 + cur.x0 : 0, t3 = valid ? r1 + cur.x1c : 0;
        for (int ks = 0; ks < nks; ++ks, ++kc) {
          float tv[KC][4];
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            const T* plane = x2n + (long long)min(ks * KC + c, g.C - 1) * g.x2s[1];
            tv[c][0] = ldg_f32(plane + t0);
            tv[c][1] = ldg_f32(plane + t1);
            tv[c][2] = ldg_f32(plane + t2);
            tv[c][3] = ldg_f32(plane + t3);
          }
          float hi[KC], lo[KC];
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            const float r = (valid && ks * KC + c < g.C) ? blend(tv[c][0], tv[c][1], tv[c][2], tv[c][3], tp) : 0.f;
            split_tf32(r, hi[c], lo[c]);
          }
          stage_b(hi, lo, ks);
          if (lane == 0) mbar_arrive(&b_full[slot_of(kc)]);
          if (ks == 0 && has_next) nxt = prepare(tile + gridDim.x, ti + 1, nfu, nfv);
        }
      }
      cur = nxt;
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue: one lane ===========================
    if (lane == 0) {
      // instruction descriptor: D fp32, A / B tf32, both K-major, N = 192, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint64_t d_ah = make_desc(sbase + OFF_AHI), d_al = make_desc(sbase + OFF_ALO);
      const uint64_t d_bh0 = make_desc(sbase + OFF_BHI), d_bh1 = make_desc(sbase + OFF_BHI + NH * 128);
      const uint64_t d_bl0 = make_desc(sbase + OFF_BLO), d_bl1 = make_desc(sbase + OFF_BLO + NH * 128);
      int kc = 0, ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        mbar_wait(d_empty, (uint32_t)((ti & 1) ^ 1));   // previous tile drained
        tc_fence_after();
        TC_TRACE(ti, 20);
        for (int ks = 0; ks < nks; ++ks, ++kc) {
          const int slot = kc & (SLOTS - 1), use = kc >> 2;
          mbar_wait(&a_full[slot], (uint32_t)(use & 1));
          mbar_wait(&b_full[slot], (uint32_t)(use & 1));
          tc_fence_after();
          if (ks < 4) TC_TRACE(ti, 21 + ks);
          const uint64_t ko = (uint64_t)(2 * slot);   // 32 bytes per K step inside the 128-byte swizzle atom
          const uint32_t acc = ks > 0 ? 1u : 0u;
          umma_tf32(tmem, d_ah + ko, d_bh0 + ko, idesc, acc);
          umma_tf32(tmem, d_al + ko, d_bh0 + ko, idesc, 1u);
          umma_tf32(tmem, d_ah + ko, d_bl0 + ko, idesc, 1u);
          umma_tf32(tmem + NH, d_ah + ko, d_bh1 + ko, idesc, acc);
          umma_tf32(tmem + NH, d_al + ko, d_bh1 + ko, idesc, 1u);
          umma_tf32(tmem + NH, d_ah + ko, d_bl1 + ko, idesc, 1u);
          umma_commit(&slot_empty[slot]);
        }
        umma_commit(d_full);
        TC_TRACE(ti, 25);
      }
    }
    __syncwarp();
  } else {
    // =========================== raw x2 boxes by TMA: one lane ===========================
    if (a.use_raw && lane == 0) {
      int rc = 0, ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        const int n = tile / per_img;
        mbar_wait(&bbox_full[ti & 1], (uint32_t)((ti >> 1) & 1));
        TC_TRACE(ti, 15);
        const int4* rp = reinterpret_cast<const int4*>(red + (ti & 1) * (GATHER_WARPS * 4));
        int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
        for (int w = 0; w < GATHER_WARPS; ++w) {
          const int4 b = rp[w];
          xmin = min(xmin, b.x); xmax = max(xmax, b.y);
          ymin = min(ymin, b.z); ymax = max(ymax, b.w);
        }
        mbar_arrive(&bbox_empty[ti & 1]);
        const int ox = xmin & ~3, oy = ymin;
        if (xmin <= xmax && xmax - ox < RAW_W && ymax - oy < RAW_H) {
          for (int ks = 0; ks < nks; ++ks, ++rc) {
            const int rs = rc & (RS - 1), ruse = rc / RS;
            mbar_wait(&raw_empty[rs], (uint32_t)((ruse & 1) ^ 1));
            mbar_arrive_expect_tx(&raw_full[rs], RAW_STAGE);
            tma_load_4d(smem + OFF_RAW + (uint32_t)rs * RAW_STAGE, &tm_raw, &raw_full[rs], ox, oy, ks * KC, x2_item(g, n));
            if (ks < 4) TC_TRACE(ti, 16 + ks);
          }
        }
      }
    }
    __syncwarp();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * 192 + 191] = clock64();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace tc

bool tc_forward_supported(const Geom& g, int dtype, const UpFlow* uf) {
  return dtype == CERB_F32 && uf == nullptr && g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == tc::MD && g.outH > 0 && g.outW > 0;
}

cudaError_t launch_warp_corr_forward_tc(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow, void* out,
                                        cudaStream_t stream) {
  if (!tc_forward_supported(g, dtype, nullptr)) return cudaErrorNotSupported;
  tc::Args a;
  a.g = g;
  a.x1 = x1; a.x2 = x2; a.flow = flow; a.out = out;
  a.off = g.md - g.pad;
  a.tiles_x = (g.outW + tc::TX - 1) / tc::TX;
  a.tiles_y = (g.outH + tc::TY - 1) / tc::TY;
  a.total_tiles = g.B * a.tiles_x * a.tiles_y;
  a.nks = (g.C + tc::KC - 1) / tc::KC;
  a.path_ctr = get_path_counters();
  a.dbg = get_trace_buffer();
  CUtensorMap tm_raw;
  memset(&tm_raw, 0, sizeof(tm_raw));
  a.use_raw = 0;
  if (!getenv("CERB_DEBUG_TC_NO_RAW"))
    a.use_raw = make_tmap_nchw(&tm_raw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x2, g.W, g.H, g.C, g.B, g.x2s, tc::RAW_W, tc::RAW_H,
                               tc::KC, false) ? 1 : 0;
  auto kern = tc::warp_corr_fwd_tc_kernel<float>;
  static unsigned long long attr_devs = 0ull;   // function attributes are per device
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit == 0ull || !(attr_devs & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr_devs |= bit;
    }
  }
  int grid = num_sms_current();
  if (grid > a.total_tiles) grid = a.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = tc::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  static const bool use_pdl = getenv("CERB_DEBUG_NO_PDL") == nullptr;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a, tm_raw);
  if (le != cudaSuccess) return le;
  return cudaGetLastError();
}

}  // namespace cerb
