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
// Roles (22 warps, one CTA per SM, persistent over tiles):
//   warps 0-7   accumulator drain, thread = pixel: warps w and w + 4 own TMEM lanes 32(w%4).. (pixel rows 2(w%4), +1) and
//               six / four of the ten halo rows those need; 24 columns at a time (tcgen05.ld 32x32b), the row is shifted
//               by the lane's own x with a 4-stage select network (the column offset differs per lane, tcgen05.ld's does
//               not), then divide, activate, store.  Warps 4-7 also stage the x1 operand: LDG in one burst per 32
//               channels, hi / lo split, tcgen05.st into TMEM columns 384.. (the A operand is read from tensor memory:
//               it never touches shared memory, whose bandwidth is what bounds this kernel).
//   warps 8-19  thread = halo position.  Flow -> sample position -> 4 taps + weights once per tile (the next tile's are
//               prepared during this tile's first K step); per K step of 8 channels the taps from the raw x2 box in
//               shared memory (or straight from global memory when the tile's footprint does not fit the box), blend in
//               ATen's order, hi / lo split, four 16-byte stores into the K-major 128B-swizzled operand rows.  The warped
//               map never exists in HBM.
//   warp 20     one lane issues tcgen05.mma.kind::tf32 (M 128, N 192 x 2, K 8): 6 per K step; tcgen05.commit releases
//               the operand slot / publishes the accumulator.
//   warp 21     one lane issues the TMA load of a K step's raw x2 box (8 ch x 30 x 44, origin from the tile's tap
//               bounding box), three in flight.
// Operand slots: a 128-byte operand row holds 32 channels = 4 K steps; slot j of every row is refilled as soon as the
// MMAs that read it have retired, so staging, MMA and the previous tile's drain overlap.  The accumulator itself is
// single-buffered (384 of TMEM's 512 columns, 64 more hold the x1 operand): MMAs of tile t+1 start when tile t has drained.
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
constexpr int N = 384, NH = N / 2;                    // accumulator columns of one pass, two MMAs of N = 192
#ifndef CERB_TC_MARGIN
#define CERB_TC_MARGIN 6
#endif
constexpr int MARGIN = CERB_TC_MARGIN;                // flow variation (px) inside one halo the raw box absorbs
constexpr int SLOTS = 4;                              // K steps per 128-byte operand row
constexpr int RS_MAX = 3;                             // raw boxes in flight (fp32: 3, 16-bit: 2)
constexpr int EPI_WARPS = 8, GATHER_WARPS = 12, A_WARPS = 4;   // drain / gather / (of the drain warps) x1 staging
constexpr int GATHER_THREADS = GATHER_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS + GATHER_WARPS, TMA_WARP = MMA_WARP + 1, MMA_WARP2 = TMA_WARP + 1;
#ifndef CERB_TC_MMA_WARPS
#define CERB_TC_MMA_WARPS 2
#endif
constexpr int MMA_ISSUERS = CERB_TC_MMA_WARPS;        // 2: one issuing warp per accumulator half (an MMA costs its issuing thread ~140 cycles)
constexpr int NTHREADS = (TMA_WARP + MMA_ISSUERS) * 32;   // 736
constexpr uint32_t B_BYTES = N * 128, A_BYTES = M * 128;
constexpr uint32_t OFF_BHI = 0;
constexpr uint32_t TMEM_A = N;                        // fp32: x1 operand in TMEM columns 384 + 16 slot (8 hi, 8 lo)
constexpr int NBARS = 3 * SLOTS + 2 * RS_MAX + 2 + 4;
static_assert(N == GATHER_THREADS && M == A_WARPS * 32 && EPI_WARPS == 8, "one halo position / one pixel per thread");

// Geometry per max_displacement.  md 4: the 16 x 24 halo of an 8 x 16 tile is one accumulator pass (N = 384).  md 8: the
// halo is 24 x 32 = 768 positions -- two passes of 12 x 32 (N = 384 each): a work unit is (tile, pass), pixel row py finds
// its displacement rows py..py+16 in pass 0 (halo rows 0-11) and pass 1 (rows 12-23).
// ES = element size of the inputs.  fp32 (ES 4): a K step is 8 channels (32 bytes of an operand row = one kind::tf32 MMA),
// operands are hi / lo pairs (3xTF32), x1 lives in TMEM.  16-bit (ES 2): a K step is 16 channels (one kind::f16 MMA); x1 is
// exact in T (a K-major shared-memory tile), the blended x2w is not, so it is kept as hi = T(v), lo = T(v - hi) and two
// products accumulate (x1*hi + x1*lo): the result carries the one rounding of the stored output, as the CUDA-core path's does.
template <int MD_, int ES>
struct Geo {
  static constexpr int MD = MD_, D = 2 * MD_ + 1;
  static constexpr bool F32 = ES == 4;
  static constexpr int KC = 32 / ES;                          // channels per K step: 8 / 16
  static constexpr int AL = 16 / ES;                          // elements per 16 bytes: TMA box starts are 16-byte aligned
  static constexpr int NPASS = MD_ == 4 ? 1 : 2;
  static constexpr int HYB = (TY + 2 * MD_) / NPASS;          // halo rows per pass: 16 / 12
  static constexpr int HX = TX + 2 * MD_;                     // halo columns: 24 / 32
  static constexpr int RAW_H = HYB + 2 * MARGIN + 2;          // 30 / 26
  static constexpr int RAW_W = (HX + 2 * MARGIN + 2 + 2 * (AL - 1)) / AL * AL;   // fp32 44 / 52, 16-bit 48 / 56
  static constexpr uint32_t RAW_PLANE = RAW_H * RAW_W * ES;
  static constexpr uint32_t RAW_STAGE = KC * RAW_PLANE;       // raw x2 box of a K step
  static constexpr int RS = F32 ? 3 : 2;                      // raw boxes in flight
  static constexpr uint32_t OFF_BLO = B_BYTES;                // lo tile of x2w
  static constexpr uint32_t OFF_A = 2 * B_BYTES;              // 16-bit: the x1 tile
  static constexpr uint32_t OFF_RAW = F32 ? 2 * B_BYTES : 2 * B_BYTES + A_BYTES;
  static constexpr uint32_t OFF_RED = OFF_RAW + RS * RAW_STAGE;
  static constexpr uint32_t OFF_BAR = OFF_RED + 2 * GATHER_WARPS * 4 * 4;
  static constexpr uint32_t OFF_TMEM = OFF_BAR + NBARS * 8;
  static constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
  static_assert(HYB * HX == N && (TY + 2 * MD_) % NPASS == 0 && HX == D + 15, "halo pass = 384 positions; shift network");
  static_assert(RAW_W >= HX + 2 * MARGIN + 2 + AL - 1, "box covers the footprint after aligning its start down");
  static_assert(RAW_STAGE % 128 == 0 && OFF_RAW % 1024 == 0, "TMA destination alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

enum { PATH_RAW = 1, PATH_DIRECT = 2 };   // indices match costvolume_fwd.cu's path counters

struct Args {
  Geom g;
  const void* x1;
  const void* x2;
  const float* flow;
  void* out;
  int off;              // md - pad
  int tiles_x, tiles_y, total_tiles;   // total_tiles counts work units: (tile, accumulator pass)
  int nks;              // K steps per tile: ceil(C / 8)
  int use_raw;          // raw x2 boxes by TMA
  // fused flow up-sampling (cerb_warp_corr_forward_upflow): coarse flow in, up-sampled flow out (cflow == nullptr: off)
  const float* cflow;
  long long cfs[3];
  int Hc, Wc;
  float up_sy, up_sx;   // ATen's align_corners=True scale (in - 1) / (out - 1), fp32
  float* flow_up;
  long long fus[3];
  unsigned long long* path_ctr;
  long long* dbg;       // optional per-CTA clock64() trace (cerb_debug_set_trace_buffer), 192 slots per CTA
};
// trace slot `s` of tile iteration `ti` (first four tiles of a CTA; 40 slots each)
// (compiled in only with -DCERB_TC_TRACE: `python tools/ab_variants.py trace:-DCERB_TC_TRACE`; each point costs ~15 instructions
// even when off at run time, and the gather and drain loops are bound by instruction issue)
#ifdef CERB_TC_TRACE
#define TC_TRACE(ti, s) do { if (a.dbg && (ti) < 4) a.dbg[(long long)blockIdx.x * 192 + 1 + (ti) * 40 + (s)] = clock64(); } while (0)
#else
#define TC_TRACE(ti, s) do { } while (0)
#endif

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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from tensor memory (lane = row, 8 consecutive 32-bit columns = the K = 8 tf32 values of one instruction)
__device__ __forceinline__ void umma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// 16-bit operands (kind::f16: fp16 or bf16 per the instruction descriptor), both from shared memory
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);   // two values rounded to T, a in the low half
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<float>(float, float) { return 0u; }   // (never used)
template <typename T> __device__ __forceinline__ float unpack_lo(uint32_t w);   // low / high half of a packed pair, widened
template <typename T> __device__ __forceinline__ float unpack_hi(uint32_t w);
template <> __device__ __forceinline__ float unpack_lo<__half>(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
template <> __device__ __forceinline__ float unpack_hi<__half>(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
template <> __device__ __forceinline__ float unpack_lo<__nv_bfloat16>(uint32_t w) { return __uint_as_float(w << 16); }
template <> __device__ __forceinline__ float unpack_hi<__nv_bfloat16>(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
template <> __device__ __forceinline__ float unpack_lo<float>(uint32_t) { return 0.f; }
template <> __device__ __forceinline__ float unpack_hi<float>(uint32_t) { return 0.f; }
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <typename T> __device__ __forceinline__ float lds_t(uint32_t addr);   // one input element from shared memory, widened
template <> __device__ __forceinline__ float lds_t<float>(uint32_t addr) { return lds_f32(addr); }
template <> __device__ __forceinline__ float lds_t<__half>(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return __half2float(__ushort_as_half(v));
}
template <> __device__ __forceinline__ float lds_t<__nv_bfloat16>(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return __uint_as_float((uint32_t)v << 16);
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
__device__ __forceinline__ void tmem_ld_tie8(uint32_t* v) {   // (extends tmem_ld_wait24 to a fourth group of eight registers)
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) : : "memory");
}
__device__ __forceinline__ int slot_of(int kc) { return kc & 3; }
// Every wait carries a suspend-time hint: the waiting warp sleeps in hardware until the phase completes instead of
// re-polling.  mbarrier polls are shared-memory operations; with 22 warps of which most are waiting at any time, plain
// try_wait loops competed with the gather's LDS / STS traffic for the same pipe.
#ifndef CERB_TC_WAIT_HINT_NS
#define CERB_TC_WAIT_HINT_NS 20000
#endif
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
#if CERB_TC_WAIT_HINT_NS > 0
  mbar_wait_hint(bar, parity, CERB_TC_WAIT_HINT_NS);
#else
  mbar_wait(bar, parity);
#endif
}

// hi = tf32(v): round to nearest (ties away) by adding half an ulp of the 10-bit mantissa to the bit pattern and clearing the
// low 13 bits -- what cvt.rna.tf32.f32 computes for finite values, in 2 instructions instead of the 4 (with an Inf / NaN
// guard) that cvt is expanded to; lo = v - hi (exact).  Non-finite inputs give non-finite outputs either way.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  lo = v - hi;
}

template <typename T, int MDT>
__global__ void __launch_bounds__(NTHREADS, 1)
warp_corr_fwd_tc_kernel(const Args a, const __grid_constant__ CUtensorMap tm_raw) {
  using G = Geo<MDT, (int)sizeof(T)>;
  constexpr bool F32 = G::F32;
  constexpr int MD = G::MD, D = G::D, NPASS = G::NPASS, HYB = G::HYB, HX = G::HX, RAW_H = G::RAW_H, RAW_W = G::RAW_W,
                KC = G::KC, AL = G::AL, ES = (int)sizeof(T);
  constexpr int RS = G::RS;
  constexpr int NB = KC / 4;          // batches of 4 channels per K step (the gather's software-pipeline unit): 2 / 4
  constexpr int AG = F32 ? SLOTS : 2; // K steps per x1 staging burst (32 channels: registers)
  constexpr uint32_t RAW_PLANE = G::RAW_PLANE, RAW_STAGE = G::RAW_STAGE, OFF_RED = G::OFF_RED, OFF_BAR = G::OFF_BAR,
                     OFF_TMEM = G::OFF_TMEM, OFF_RAW = G::OFF_RAW, OFF_BLO = G::OFF_BLO, OFF_A = G::OFF_A;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  int* red = (int*)(smem + OFF_RED);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* a_full = bars;                    // [SLOTS] x1 slot written (drain warps 4-7)
  uint64_t* b_full = bars + SLOTS;            // [SLOTS] warped-x2 slot written (12 warps)
  uint64_t* slot_empty = bars + 2 * SLOTS;    // [SLOTS] MMAs reading the slot have retired (tcgen05.commit)
  uint64_t* raw_full = bars + 3 * SLOTS;      // [RS]
  uint64_t* raw_empty = raw_full + RS;        // [RS]
  uint64_t* d_full = raw_empty + RS;          // accumulator of the tile complete
  uint64_t* d_empty = d_full + 1;             // accumulator drained (8 warps)
  uint64_t* bbox_full = d_empty + 1;          // [2] every gather warp has published its share of the tile's tap bounding box
  uint64_t* bbox_empty = bbox_full + 2;       // [2] the TMA warp has read it (the slot may be rewritten two tiles later)
  uint32_t* tmem_slot = (uint32_t*)(smem + OFF_TMEM);

  const Geom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* __restrict__ x1 = (const T*)a.x1;
  const T* __restrict__ x2 = (const T*)a.x2;

  if (tid == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&a_full[s], A_WARPS);
      mbar_init(&b_full[s], GATHER_WARPS);
      mbar_init(&slot_empty[s], MMA_ISSUERS);
    }
    for (int s = 0; s < RS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], GATHER_WARPS);
    }
    mbar_init(d_full, MMA_ISSUERS);
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
#ifdef CERB_TC_TRACE
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * 192] = clock64();
#endif
  // PDL: everything above overlapped the previous kernel's tail; its outputs may be our inputs
  pdl_wait();

  const int per_img = a.tiles_x * a.tiles_y;
  const int nks = a.nks;
  // work unit -> batch item, tile origin (output coordinates), accumulator pass
  auto decode = [&](int unit, int& n, int& by0, int& bx0, int& pass) {
    pass = NPASS > 1 ? unit % NPASS : 0;
    const int t = NPASS > 1 ? unit / NPASS : unit;
    n = t / per_img;
    const int trem = t - n * per_img;
    by0 = (trem / a.tiles_x) * TY;
    bx0 = (trem % a.tiles_x) * TX;
  };

  if (warp < EPI_WARPS) {
    // =========================== x1 operand (warps 4-7) + accumulator drain: thread = pixel ===========================
    // TMEM lane = pixel, column = halo position; this lane's displacements are halo rows py..py+8, columns px..px+8.
    // Warps w and w + 4 share a lane quadrant: w takes its halo rows 0-5, w + 4 rows 6-9 and the x1 staging.
    const int q = warp & 3;
    const bool a_warp = warp >= A_WARPS;
    const int pl = q * 32 + lane;                   // pixel of the tile = TMEM lane
    const int py = pl >> 4, px = pl & 15;
    const float fC = (float)g.C, rC = __frcp_rn((float)g.C);
    const bool c_pow2 = (g.C & (g.C - 1)) == 0;   // 1/C exact: the division is one multiply
    const float act_slope = g.has_act ? g.slope : 1.f;   // no activation = slope 1
    const long long os1 = g.os[1];
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    // x1: 32 channels of this thread's pixel are requested in one burst and consumed right away (loads interleaved with
    // the consumption of older ones made every use wait for the youngest: scoreboards count, they do not track loads)
    int ka = 0;   // K steps staged so far (all tiles)
    auto stage_a_group = [&](int tile, int ks0) {
      int n, by0, bx0, pass;
      decode(tile, n, by0, bx0, pass);
      const int iy = by0 + a.off + py, ix = bx0 + a.off + px;
      const bool inimg = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const T* p = x1 + (long long)n * g.x1s[0] + (long long)min(max(iy, 0), g.H - 1) * g.x1s[2] + min(max(ix, 0), g.W - 1);
      float nx[AG][KC];
#pragma unroll
      for (int b = 0; b < AG; ++b)
#pragma unroll
        for (int c = 0; c < KC; ++c) {   // unconditional loads from clamped addresses, validity applied afterwards
          const int ch = (ks0 + b) * KC + c;
          const float v = ldg_f32(p + (long long)min(ch, g.C - 1) * g.x1s[1]);
          nx[b][c] = (inimg && ch < g.C) ? v : 0.f;
        }
#pragma unroll
      for (int b = 0; b < AG; ++b) {
        if (ks0 + b < nks) {
          const int slot = ka & (SLOTS - 1), use = ka >> 2;
          if constexpr (F32) {
            float hi[KC], lo[KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) split_tf32(nx[b][c], hi[c], lo[c]);
            tc_wait(&slot_empty[slot], (uint32_t)((use & 1) ^ 1));
            tc_fence_after();
            tmem_st8(tlane + TMEM_A + (uint32_t)slot * 16u, hi);
            tmem_st8(tlane + TMEM_A + (uint32_t)slot * 16u + 8u, lo);
            tmem_st_wait();
            tc_fence_before();
          } else {
            // 16-bit: the pixel's row of the K-major x1 tile (the values are exact in T: they were loaded as T)
            uint32_t w[KC / 2];
#pragma unroll
            for (int c = 0; c < KC / 2; ++c) w[c] = pack2<T>(nx[b][2 * c], nx[b][2 * c + 1]);
            tc_wait(&slot_empty[slot], (uint32_t)((use & 1) ^ 1));
            const uint32_t arow = sbase + OFF_A + (uint32_t)pl * 128u, sw = (uint32_t)(pl & 7);
            sts128u(arow + ((((uint32_t)(2 * slot)) ^ sw) << 4), w[0], w[1], w[2], w[3]);
            sts128u(arow + ((((uint32_t)(2 * slot + 1)) ^ sw) << 4), w[4], w[5], w[6], w[7]);
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[slot]);
          ++ka;
        }
      }
    };
    int ti = 0;
    if (a_warp && (int)blockIdx.x < a.total_tiles) stage_a_group(blockIdx.x, 0);   // first group of the first tile
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      if (a_warp) {
        // the rest of this tile's x1 (slots free up as its MMAs retire), then the first group of the next tile: staged
        // while this tile's MMAs run, so the next tile's can start the moment the accumulator has drained
        for (int ks0 = AG; ks0 < nks; ks0 += AG) stage_a_group(tile, ks0);
        if (tile + (int)gridDim.x < a.total_tiles) stage_a_group(tile + gridDim.x, 0);
        if (warp == A_WARPS && lane == 0) TC_TRACE(ti, 26);
      }
      int n, by0, bx0, pass;
      decode(tile, n, by0, bx0, pass);
      const int oy = by0 + py, ox = bx0 + px;
      const bool pix_ok = oy < g.outH && ox < g.outW;
      T* op = (T*)a.out + (long long)n * g.os[0] + (long long)min(oy, g.outH - 1) * g.os[2] + min(ox, g.outW - 1);
      tc_wait(d_full, (uint32_t)(ti & 1));
      tc_fence_after();
      if (tid == 0) TC_TRACE(ti, 30);
      // halo rows (of the whole halo) this pass holds that the quadrant's pixel rows 2q, 2q+1 need; the two warps of a
      // quadrant split them (the x1-staging warp takes the smaller share)
      const int r_lo = max(pass * HYB, 2 * q), r_hi = min(pass * HYB + HYB, 2 * q + 1 + D);   // [r_lo, r_hi), an even count
      const int n0 = 2 * (((r_hi - r_lo) / 2 + 1) / 2);
      const int rr0 = a_warp ? r_lo + n0 : r_lo, rr1 = a_warp ? r_hi : r_lo + n0;
      // one halo row of the accumulator: shift by the lane's own x (select network: the column offset differs per lane,
      // tcgen05.ld's does not), divide, activate, store the D displacement columns of displacement row r - py
      auto finish_row = [&](uint32_t* v, int r) {
#pragma unroll
        for (int i = 0; i < D + 7; ++i) v[i] = (px & 8) ? v[i + 8] : v[i];
#pragma unroll
        for (int i = 0; i < D + 3; ++i) v[i] = (px & 4) ? v[i + 4] : v[i];
#pragma unroll
        for (int i = 0; i < D + 1; ++i) v[i] = (px & 2) ? v[i + 2] : v[i];
#pragma unroll
        for (int i = 0; i < D; ++i) v[i] = (px & 1) ? v[i + 1] : v[i];
        const int dy = r - py;
        if (pix_ok && dy >= 0 && dy < D) {
          T* orow = op + (long long)(dy * D) * os1;
          if (c_pow2) {
#pragma unroll
            for (int dx = 0; dx < D; ++dx) {
              const float res = __fmul_rn(__uint_as_float(v[dx]), rC);
#ifdef CERB_TCX_NOSTG
              if (res == 123.456f)
#endif
              *orow = from_f32<T>(res > 0.f ? res : res * act_slope);
              orow += os1;
            }
          } else {
#pragma unroll
            for (int dx = 0; dx < D; ++dx) {
              const float res = div_const(__uint_as_float(v[dx]), fC, rC);
              *orow = from_f32<T>(res > 0.f ? res : res * act_slope);
              orow += os1;
            }
          }
        }
      };
      auto release = [&]() {   // every load of this warp has landed: (with the other seven) the next unit's MMAs may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);
        if (tid == 0) TC_TRACE(ti, 31);
      };
      if (rr0 >= rr1) release();
      if constexpr (MDT == 4) {
        // two halo rows per iteration: their select networks and stores are independent instruction streams (one warp
        // per sub-partition drains at a time -- what it lacks is instruction-level parallelism, not issue slots)
#pragma unroll 1
#ifdef CERB_TCX_NODRAIN
        for (int r = rr1 - 2; r < rr1; r += 2) {
#else
        for (int r = rr0; r < rr1; r += 2) {
#endif
          uint32_t v0[24], v1[24];
          const uint32_t col = (uint32_t)((r - pass * HYB) * HX);
          tmem_ld8(tlane + col, v0);
          tmem_ld8(tlane + col + 8, v0 + 8);
          tmem_ld8(tlane + col + 16, v0 + 16);
          tmem_ld8(tlane + col + HX, v1);
          tmem_ld8(tlane + col + HX + 8, v1 + 8);
          tmem_ld8(tlane + col + HX + 16, v1 + 16);
          tmem_ld_wait24(v0);
          tmem_ld_wait24(v1);
          if (r == rr1 - 2) release();
          finish_row(v0, r);
          finish_row(v1, r + 1);
        }
      } else {
#pragma unroll 1
        for (int r = rr0; r < rr1; ++r) {
          uint32_t v[32];
          const uint32_t col = (uint32_t)((r - pass * HYB) * HX);
          tmem_ld8(tlane + col, v);
          tmem_ld8(tlane + col + 8, v + 8);
          tmem_ld8(tlane + col + 16, v + 16);
          tmem_ld8(tlane + col + 24, v + 24);
          tmem_ld_wait24(v);
          tmem_ld_tie8(v + 24);
          if (r == rr1 - 1) release();
          finish_row(v, r);
        }
      }
      if (tid == 0) TC_TRACE(ti, 32);
    }
  } else if (warp < MMA_WARP) {
    // =========================== warped second map: thread = halo position ===========================
    const int gt = tid - EPI_WARPS * 32, gw = warp - EPI_WARPS;
    const int hy = gt / HX, hx = gt - hy * HX;
    const uint32_t brow = (uint32_t)gt * 128u, sw = (uint32_t)(gt & 7);
    const bool upflow = a.cflow != nullptr;
    const bool warped = a.flow != nullptr || upflow;
    const AxisConst axis_x = make_axis(g.W), axis_y = make_axis(g.H);
    // sampling data of one tile for this thread's halo position
    struct Pos {
      int x0, x1c, y0, y1c;
      float w[4];
      bool valid;
    };
    // halo position of a tile -> clamped pixel (the flow vector is read there), inside-the-image flag
    auto locate = [&](int tile, int& n, int& cy, int& cx, int& ghy) -> bool {
      int by0, bx0, pass;
      decode(tile, n, by0, bx0, pass);
      ghy = pass * HYB + hy;   // row inside the tile's whole halo
      const int qy = by0 + a.off - MD + ghy, qx = bx0 + a.off - MD + hx;
      cy = min(max(qy, 0), g.H - 1); cx = min(max(qx, 0), g.W - 1);
      return qy >= 0 && qy < g.H && qx >= 0 && qx < g.W;   // outside: the correlation's zero padding
    };
    auto load_flow = [&](int tile, float& fu, float& fv) {
      int n, cy, cx, ghy;
      const bool inside = locate(tile, n, cy, cx, ghy);
      if (!upflow) {
        const float* fp = a.flow + (long long)n * g.fls[0] + (long long)cy * g.fls[2] + cx;
        fu = __ldg(fp); fv = __ldg(fp + g.fls[1]);
        return;
      }
      // flow = interpolate(2 * coarse, scale_factor=2, bilinear, align_corners=True) (pwcnet_sfd.py:176), evaluated per
      // halo position with ATen's arithmetic (upsample_bilinear2d: src = scale * dst, lambda = src - int(src)); doubling
      // commutes exactly with the blend.  Same expressions as costvolume_fwd.cu: bit-exact against ATen's CUDA kernel.
      const float* cn = a.cflow + (long long)n * a.cfs[0];
      const float h1r = __fmul_rn(a.up_sy, (float)cy), w1r = __fmul_rn(a.up_sx, (float)cx);
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = (h1 < a.Hc - 1) ? 1 : 0, w1p = (w1 < a.Wc - 1) ? 1 : 0;
      const float h1l = __fsub_rn(h1r, (float)h1), w1l = __fsub_rn(w1r, (float)w1);
      const float h0l = __fsub_rn(1.f, h1l), w0l = __fsub_rn(1.f, w1l);
      const float* p0 = cn + (long long)h1 * a.cfs[2] + w1;
      const float* p1 = p0 + (long long)h1p * a.cfs[2];
      float cv[8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        cv[4 * c + 0] = __ldg(p0 + c * a.cfs[1]);
        cv[4 * c + 1] = __ldg(p0 + c * a.cfs[1] + w1p);
        cv[4 * c + 2] = __ldg(p1 + c * a.cfs[1]);
        cv[4 * c + 3] = __ldg(p1 + c * a.cfs[1] + w1p);
      }
      float r[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float t0 = __fmaf_rn(w0l, cv[4 * c + 0], __fmul_rn(w1l, cv[4 * c + 1]));
        const float t1 = __fmaf_rn(w0l, cv[4 * c + 2], __fmul_rn(w1l, cv[4 * c + 3]));
        r[c] = __fmul_rn(2.f, __fmaf_rn(h0l, t0, __fmul_rn(h1l, t1)));
      }
      fu = r[0]; fv = r[1];
      // positions inside the tile itself (each pixel of the image exactly once) also write the up-sampled flow out
      if (inside && ghy >= MD && ghy < MD + TY && hx >= MD && hx < MD + TX) {
        float* up = a.flow_up + (long long)n * a.fus[0] + (long long)cy * a.fus[2] + cx;
        up[0] = r[0];
        up[a.fus[1]] = r[1];
      }
    };
    // taps of the tile and this warp's share of their bounding box (published for tile iteration `it`)
    auto prepare = [&](int tile, int it, float fu, float fv) -> Pos {
      int n, cy, cx, ghy;
      Pos p;
      p.valid = locate(tile, n, cy, cx, ghy);
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
          if (it >= 2) tc_wait(&bbox_empty[it & 1], (uint32_t)(((it >> 1) - 1) & 1));
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
      int n, by0_, bx0_, pass_;
      decode(tile, n, by0_, bx0_, pass_);
      const bool has_next = tile + (int)gridDim.x < a.total_tiles;
      float nfu = 0.f, nfv = 0.f;   // the next tile's flow vector: requested now, used after the first K step
      if (has_next && warped) load_flow(tile + gridDim.x, nfu, nfv);
      // ---- does the tile's sampling footprint fit the raw box?
      int path = PATH_DIRECT, ox = 0, oy = 0;
      if (a.use_raw) {
        tc_wait(&bbox_full[ti & 1], (uint32_t)((ti >> 1) & 1));
        const int4* rp = reinterpret_cast<const int4*>(red + (ti & 1) * (GATHER_WARPS * 4));
        int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
        for (int w = 0; w < GATHER_WARPS; ++w) {
          const int4 b = rp[w];
          xmin = min(xmin, b.x); xmax = max(xmax, b.y);
          ymin = min(ymin, b.z); ymax = max(ymax, b.w);
        }
        ox = xmin & ~(AL - 1);   // TMA box starts must be 16-byte aligned
        oy = ymin;
        if (xmin <= xmax && xmax - ox < RAW_W && ymax - oy < RAW_H) path = PATH_RAW;
      }
      if (gt == 0) TC_TRACE(ti, 2);
      if (gt == 0 && a.path_ctr != nullptr) atomicAdd(&a.path_ctr[path], 1ull);
      const bool valid = cur.valid;
      Taps tp;
#pragma unroll
      for (int k = 0; k < 4; ++k) tp.w[k] = cur.w[k];
      // one batch = 4 channels of this thread's position: blended (fp32, ATen's order; validity as a select -- a branch
      // here diverges per lane), then written into the operand row.  fp32: hi / lo split, one 16-byte chunk each;
      // 16-bit: rounded once to T, 8 bytes of the chunk.
      auto store_batch = [&](const float (&v)[4][4], int slot, int bch, int chan0) {
        float r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float x = blend(v[c][0], v[c][1], v[c][2], v[c][3], tp);
          r[c] = (valid && chan0 + c < g.C) ? x : 0.f;
        }
        if constexpr (F32) {
          float h[4], l[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) split_tf32(r[c], h[c], l[c]);
          const uint32_t ch = (((uint32_t)(2 * slot + bch)) ^ sw) << 4;
#ifdef CERB_TCX_NOSTS
          if (h[0] == 123.456f) sts128(sbase + OFF_BHI + brow + ch, make_float4(h[0], h[1], l[2], l[3]));
#else
          sts128(sbase + OFF_BHI + brow + ch, make_float4(h[0], h[1], h[2], h[3]));
          sts128(sbase + OFF_BLO + brow + ch, make_float4(l[0], l[1], l[2], l[3]));
#endif
        } else {
          const uint32_t ch = ((((uint32_t)(2 * slot + (bch >> 1))) ^ sw) << 4) + (uint32_t)(bch & 1) * 8u;
          const uint32_t h01 = pack2<T>(r[0], r[1]), h23 = pack2<T>(r[2], r[3]);
          float l[4];   // residual of the rounding to T (exact in fp32), itself rounded to T
          l[0] = r[0] - unpack_lo<T>(h01); l[1] = r[1] - unpack_hi<T>(h01);
          l[2] = r[2] - unpack_lo<T>(h23); l[3] = r[3] - unpack_hi<T>(h23);
          sts64(sbase + OFF_BHI + brow + ch, h01, h23);
          sts64(sbase + OFF_BLO + brow + ch, pack2<T>(l[0], l[1]), pack2<T>(l[2], l[3]));
        }
      };
      if (path == PATH_RAW) {
        // byte offsets of the four taps inside a channel plane of the box (positions without a sample read offset 0)
        const int r0 = (cur.y0 - oy) * RAW_W - ox, r1 = (cur.y1c - oy) * RAW_W - ox;
        const uint32_t t0 = valid ? (uint32_t)(r0 + cur.x0) * ES : 0u, t1 = valid ? (uint32_t)(r0 + cur.x1c) * ES : 0u;
        const uint32_t t2 = valid ? (uint32_t)(r1 + cur.x0) * ES : 0u, t3 = valid ? (uint32_t)(r1 + cur.x1c) * ES : 0u;
        // Software pipeline over batches of 4 channels: the 16 tap loads of the next batch -- across the K-step boundary
        // too, the box is usually there already -- are in flight while this one is blended and stored.
        auto load_batch = [&](uint32_t rb, int bch, float (&dst)[4][4]) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t pl = rb + (uint32_t)(bch * 4 + c) * RAW_PLANE;
#ifdef CERB_TCX_NOLDS   // CERB_TCX_*: timing-only elimination builds (tools/ab_variants.py), results are wrong
            dst[c][0] = __uint_as_float(pl + t0); dst[c][1] = __uint_as_float(pl + t1);
            dst[c][2] = __uint_as_float(pl + t2); dst[c][3] = __uint_as_float(pl + t3);
#else
            dst[c][0] = lds_t<T>(pl + t0);
            dst[c][1] = lds_t<T>(pl + t1);
            dst[c][2] = lds_t<T>(pl + t2);
            dst[c][3] = lds_t<T>(pl + t3);
#endif
          }
        };
        float la[4][4], lb[4][4];
        {
          const int rs = rc % RS, ruse = rc / RS;
          tc_wait(&raw_full[rs], (uint32_t)(ruse & 1));
          load_batch(sbase + OFF_RAW + (uint32_t)rs * RAW_STAGE, 0, la);
        }
        for (int ks = 0; ks < nks; ++ks, ++kc, ++rc) {
          const int rs = rc % RS;
          const uint32_t rb = sbase + OFF_RAW + (uint32_t)rs * RAW_STAGE;
          const int slot = kc & (SLOTS - 1), use = kc >> 2;
          if (gt == 0 && ks < 4) TC_TRACE(ti, 3 + 3 * ks);
#pragma unroll
          for (int bch = 0; bch < NB; ++bch) {
            float (&cu)[4][4] = (bch & 1) ? lb : la;
            float (&nx)[4][4] = (bch & 1) ? la : lb;
            if (bch + 1 < NB) {
              load_batch(rb, bch + 1, nx);
            } else if (ks + 1 < nks) {   // first batch of the next K step (NB is even: it lands in `la`)
              const int rn = (rc + 1) % RS, rnuse = (rc + 1) / RS;
              tc_wait(&raw_full[rn], (uint32_t)(rnuse & 1));
              load_batch(sbase + OFF_RAW + (uint32_t)rn * RAW_STAGE, 0, nx);
            }
            if (bch == 0) {
              tc_wait(&slot_empty[slot], (uint32_t)((use & 1) ^ 1));
              if (gt == 0 && ks < 4) TC_TRACE(ti, 4 + 3 * ks);
            }
            store_batch(cu, slot, bch, ks * KC + bch * 4);
          }
#ifndef CERB_TCX_NOFENCE
          fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
#endif
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&raw_empty[rs]);
            mbar_arrive(&b_full[slot]);
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
          const int slot = kc & (SLOTS - 1), use = kc >> 2;
#pragma unroll
          for (int bch = 0; bch < NB; ++bch) {
            float tv[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const T* plane = x2n + (long long)min(ks * KC + bch * 4 + c, g.C - 1) * g.x2s[1];
              tv[c][0] = ldg_f32(plane + t0);
              tv[c][1] = ldg_f32(plane + t1);
              tv[c][2] = ldg_f32(plane + t2);
              tv[c][3] = ldg_f32(plane + t3);
            }
            if (bch == 0) tc_wait(&slot_empty[slot], (uint32_t)((use & 1) ^ 1));
            store_batch(tv, slot, bch, ks * KC + bch * 4);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&b_full[slot]);
          if (ks == 0 && has_next) nxt = prepare(tile + gridDim.x, ti + 1, nfu, nfv);
        }
      }
      cur = nxt;
    }
  } else if (warp == MMA_WARP || warp == MMA_WARP2) {
    // =========================== MMA issue: one lane per accumulator half ===========================
    // (the two N = 192 halves of the accumulator are independent; issuing a tcgen05.mma costs the thread ~140 cycles --
    // measured on the backward kernel, costvolume_bwd_tc.cu -- so two issuing warps shorten the MMA -> drain chain)
    if (lane == 0) {
      const int h0 = MMA_ISSUERS == 2 ? (warp == MMA_WARP ? 0 : 1) : 0, h1 = MMA_ISSUERS == 2 ? h0 + 1 : 2;
      // instruction descriptor: D fp32 (bit 4), A / B format (bits 7-9 / 10-12: tf32 = 2; kind::f16: fp16 = 0, bf16 = 1), both
      // K-major, N = 192, M = 128
      constexpr uint32_t fmt = F32 ? 2u : (std::is_same<T, __half>::value ? 0u : 1u);
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint64_t d_a = make_desc(sbase + OFF_A);                                                     // (16-bit only)
      int kc = 0, ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        tc_wait(d_empty, (uint32_t)((ti & 1) ^ 1));   // previous tile drained
        tc_fence_after();
        if (warp == MMA_WARP) TC_TRACE(ti, 20);
        for (int ks = 0; ks < nks; ++ks, ++kc) {
          const int slot = kc & (SLOTS - 1), use = kc >> 2;
          tc_wait(&a_full[slot], (uint32_t)(use & 1));
          tc_wait(&b_full[slot], (uint32_t)(use & 1));
          tc_fence_after();
          if (warp == MMA_WARP && ks < 4) TC_TRACE(ti, 21 + ks);
          const uint64_t ko = (uint64_t)(2 * slot);   // 32 bytes per K step inside the 128-byte swizzle atom
          const uint32_t acc = ks > 0 ? 1u : 0u;
#ifndef CERB_TCX_NOMMA
          for (int h = h0; h < h1; ++h) {
            const uint64_t d_bh = make_desc(sbase + OFF_BHI + (uint32_t)h * NH * 128), d_bl = make_desc(sbase + OFF_BLO + (uint32_t)h * NH * 128);
            const uint32_t d_t = tmem + (uint32_t)h * NH;
            if constexpr (F32) {
              const uint32_t a_hi = tmem + TMEM_A + (uint32_t)slot * 16u, a_lo = a_hi + 8u;
              umma_tf32_ta(d_t, a_hi, d_bh + ko, idesc, acc);
              umma_tf32_ta(d_t, a_lo, d_bh + ko, idesc, 1u);
              umma_tf32_ta(d_t, a_hi, d_bl + ko, idesc, 1u);
            } else {
              umma_f16(d_t, d_a + ko, d_bh + ko, idesc, acc);
              umma_f16(d_t, d_a + ko, d_bl + ko, idesc, 1u);
            }
          }
#endif
          umma_commit(&slot_empty[slot]);
        }
        umma_commit(d_full);
        if (warp == MMA_WARP) TC_TRACE(ti, 25);
      }
    }
    __syncwarp();
  } else if (warp == TMA_WARP) {
    // =========================== raw x2 boxes by TMA: one lane ===========================
    if (a.use_raw && lane == 0) {
      int rc = 0, ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        int n, by0_, bx0_, pass_;
        decode(tile, n, by0_, bx0_, pass_);
        tc_wait(&bbox_full[ti & 1], (uint32_t)((ti >> 1) & 1));
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
        const int ox = xmin & ~(AL - 1), oy = ymin;
        if (xmin <= xmax && xmax - ox < RAW_W && ymax - oy < RAW_H) {
          for (int ks = 0; ks < nks; ++ks, ++rc) {
            const int rs = rc % RS, ruse = rc / RS;
            tc_wait(&raw_empty[rs], (uint32_t)((ruse & 1) ^ 1));
#ifdef CERB_TCX_NOTMA   // timing experiment only: the stage "lands" at once (stale data)
            mbar_arrive(&raw_full[rs]);
#else
            mbar_arrive_expect_tx(&raw_full[rs], RAW_STAGE);
            tma_load_4d(smem + OFF_RAW + (uint32_t)rs * RAW_STAGE, &tm_raw, &raw_full[rs], ox, oy, ks * KC, x2_item(g, n));
#endif
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
#ifdef CERB_TC_TRACE
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * 192 + 191] = clock64();
#endif
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace tc

bool tc_forward_supported(const Geom& g, int dtype, const UpFlow* uf) {
  if (uf != nullptr && (g.pad != g.md || (g.H & 1) || (g.W & 1))) return false;   // tiles must cover the image exactly once
  return (dtype == CERB_F32 || dtype == CERB_F16 || dtype == CERB_BF16) && g.k == 1 && g.s1 == 1 && g.s2 == 1 &&
         (g.md == 4 || g.md == 8) && g.outH > 0 && g.outW > 0;
}

template <typename T, int MDT>
static cudaError_t launch_tc_md(const Geom& g, const void* x1, const void* x2, const float* flow, void* out, cudaStream_t stream,
                                const UpFlow* uf) {
  using G = tc::Geo<MDT, (int)sizeof(T)>;
  tc::Args a;
  a.g = g;
  a.x1 = x1; a.x2 = x2; a.flow = flow; a.out = out;
  a.off = g.md - g.pad;
  a.tiles_x = (g.outW + tc::TX - 1) / tc::TX;
  a.tiles_y = (g.outH + tc::TY - 1) / tc::TY;
  a.total_tiles = g.B * a.tiles_x * a.tiles_y * G::NPASS;
  a.nks = (g.C + G::KC - 1) / G::KC;
  a.path_ctr = get_path_counters();
  a.dbg = get_trace_buffer();
  a.cflow = nullptr; a.flow_up = nullptr; a.Hc = a.Wc = 0; a.up_sy = a.up_sx = 0.f;
  for (int i = 0; i < 3; ++i) a.cfs[i] = a.fus[i] = 0;
  if (uf != nullptr) {
    a.cflow = uf->coarse; a.flow_up = uf->up; a.Hc = uf->Hc; a.Wc = uf->Wc;
    for (int i = 0; i < 3; ++i) { a.cfs[i] = uf->cs[i]; a.fus[i] = uf->us[i]; }
    // ATen area_pixel_compute_scale<float>(in, out, align_corners=true): (float)(in - 1) / (out - 1)
    a.up_sy = g.H > 1 ? (float)(uf->Hc - 1) / (float)(g.H - 1) : 0.f;
    a.up_sx = g.W > 1 ? (float)(uf->Wc - 1) / (float)(g.W - 1) : 0.f;
  }
  CUtensorMap tm_raw;
  memset(&tm_raw, 0, sizeof(tm_raw));
  a.use_raw = 0;
  const CUtensorMapDataType dt = std::is_same<T, float>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  if (!getenv("CERB_DEBUG_TC_NO_RAW"))   // TMA needs 16-byte aligned base / strides (make_tmap_nchw checks)
    a.use_raw = make_tmap_nchw(&tm_raw, dt, (int)sizeof(T), x2, g.W, g.H, g.C, g.B, g.x2s, G::RAW_W, G::RAW_H, G::KC, false) ? 1 : 0;
  auto kern = tc::warp_corr_fwd_tc_kernel<T, MDT>;
  static unsigned long long attr_devs = 0ull;   // function attributes are per device
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit == 0ull || !(attr_devs & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr_devs |= bit;
    }
  }
  int grid = num_sms_current();
  if (grid > a.total_tiles) grid = a.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
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

template <typename T>
static cudaError_t launch_tc_t(const Geom& g, const void* x1, const void* x2, const float* flow, void* out, cudaStream_t stream,
                               const UpFlow* uf) {
  return g.md == 4 ? launch_tc_md<T, 4>(g, x1, x2, flow, out, stream, uf) : launch_tc_md<T, 8>(g, x1, x2, flow, out, stream, uf);
}

cudaError_t launch_warp_corr_forward_tc(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow, void* out,
                                        cudaStream_t stream, const UpFlow* uf) {
  if (!tc_forward_supported(g, dtype, uf)) return cudaErrorNotSupported;
  switch (dtype) {
    case CERB_F32: return launch_tc_t<float>(g, x1, x2, flow, out, stream, uf);
    case CERB_F16: return launch_tc_t<__half>(g, x1, x2, flow, out, stream, uf);
    case CERB_BF16: return launch_tc_t<__nv_bfloat16>(g, x1, x2, flow, out, stream, uf);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cerb
