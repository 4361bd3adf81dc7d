// costvolume_fwd.cu -- fused flow-warp + correlation + LeakyReLU forward for sm_100a.
//
// Replaces (reference paths relative to the reference checkout):
//   flow_warp                     nnet_training/loss_functions/UnFlowLoss.py:83-94
//   correlation_forward_cuda      nnet_training/correlation_package/correlation_cuda.cpp:3-26
//     channels_first x2, correlation_forward   correlation_cuda_kernel.cu:13-27, 29-95, 244-324
//   F.leaky_relu_(out, 0.1)       nnet_training/nnet_models/pwcnet_sfd.py:182
//
// Fast path (kernel_size=1, stride1=stride2=1, max_displacement>=4; 4 is every model in the reference):
//   persistent, warp-specialised CTAs (16 warps), one output tile of TY x TX pixels at a time.
//   * 1 TMA warp (one elected lane): x1 tile per channel chunk (cp.async.bulk.tensor, zero fill = the
//     correlation padding) and, per chunk, the raw x2 source box that covers every bilinear tap of
//     the tile's (TY+8) x (TX+8) halo (or the halo tile itself when there is no flow);
//   * 6 gather warps: flow -> sample positions -> taps and weights once per tile (registers), then
//     per chunk 4 LDS per sample from the raw box, blend, STS into the warped halo tile -- the
//     warped map never exists in HBM.  16-bit inputs arrive as 16-bit boxes and are widened here;
//   * 9 consumer warps: thread = (8-pixel strip, row y, row displacement dy) holding 8 x 9
//     accumulators; per channel 2 + 4 LDS.128 feed 72 FFMA.  Lane pairs are arranged to read the
//     same x2 row (B200 merges adjacent-lane duplicate LDS.128 addresses: 2.4 instead of 4
//     clk/instr, tools/microbench/pipes.cu);
//   * epilogue: divide by C, LeakyReLU, stage the 81 x TY x TX tile in shared memory (128B
//     swizzle), TMA store through a 5-D tensor map (x, y, dx, dy, n): one box for a CTA's last
//     tile, nine per-column boxes issued as they are written otherwise;
//   * few tiles (coarse pyramid levels): 4 x 16 tiles, channels split over KS consumer groups and
//     over the CTAs of a thread-block cluster, partial tiles pushed through distributed shared
//     memory and summed behind one hardware cluster barrier;
//   * max_displacement > 4: the D x D displacement range is covered by 9 x 9 windows (origins 0, 8,
//     ..., D-9; neighbours overlap by one row/column and write identical values there); a work unit
//     is (tile, window), the halo tile origin shifts with the window;
//   * programmatic dependent launch: setup runs before griddepcontrol.wait, every role triggers
//     launch_dependents after its last tile.
// Generic path (any pad/kernel/stride parameters): one thread per output element.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

static long long* g_trace_buffer = nullptr;  // debugging only
void set_trace_buffer(long long* p) { g_trace_buffer = p; }
long long* get_trace_buffer() { return g_trace_buffer; }
static int g_trace_iter = 0;
static unsigned long long* g_path_counters = nullptr;  // optional: tiles per staging path (bench.py reports the fallback rate)
void set_path_counters(unsigned long long* p) { g_path_counters = p; }
unsigned long long* get_path_counters() { return g_path_counters; }
void set_trace_iter(int it) { g_trace_iter = it; }
// records only the `dbg_iter`-th tile processed by each CTA (every role keeps its own `titer`)
constexpr int kTraceSlots = 192;  // clock64 slots per CTA in the debugging trace
#define CERB_TRACE(slot) do { if (a.dbg && titer == a.dbg_iter) a.dbg[(long long)blockIdx.x * kTraceSlots + (slot)] = clock64(); } while (0)

// ------------------------------------------------------------------ configuration --------
constexpr int kMD = 4;
constexpr int kD = 2 * kMD + 1;  // 9
constexpr int kD2 = kD * kD;     // 81
constexpr int kProducerThreads = 224;  // 7 staging warps (16 warps total = 4 per SMSP -> 128 regs)
constexpr int kProducerWarps = kProducerThreads / 32;
// Two staging warps only issue TMA copies (one elected lane each): a cp.async.bulk.tensor issue occupies its
// thread for 350-700 cycles whatever the box size (tools/microbench/tma.cu: ~566 cycles per box from one
// thread, aggregate rate scales with the number of issuing warps), so the raw x2 boxes and the x1 / plain x2
// tiles go out from different warps.  Staging warp w is CTA warp 9 + w: warp 9 heads the pipeline (raw boxes)
// and sits on a sub-partition with two consumer warps; the x1 issuer runs ahead of everyone and is parked on
// the sub-partition that hosts three consumer warps.
#ifndef CERB_AB_T0
#define CERB_AB_T0 0
#endif
#ifndef CERB_AB_T1
#define CERB_AB_T1 3
#endif
#ifndef CERB_AB_PIPE
#define CERB_AB_PIPE 1
#endif
constexpr int kTmaWarp = CERB_AB_T0;    // raw x2 source boxes (and, for 16-bit inputs, the x1 tile that rides with them)
constexpr int kTmaWarp2 = CERB_AB_T1;   // x1 tiles, un-warped x2 halo tiles
constexpr int kGatherWarps = kProducerWarps - 2;
constexpr int kBboxThreads = (kGatherWarps + 1) * 32;   // gather warps + the raw-box issuer meet on named barrier 2

constexpr int kGatherThreads = kGatherWarps * 32;
// x1 / warped-x2 pipeline stages: 3 for the split-channel 4x16 configuration; the 8x32 configuration stores its
// accumulators straight to global memory (no staged output tile), which leaves room for a deeper pipeline
#ifndef CERB_AB_ST1
#define CERB_AB_ST1 4
#endif
#ifndef CERB_AB_RS1
#define CERB_AB_RS1 3
#endif
#ifndef CERB_AB_CC1
#define CERB_AB_CC1 4
#endif
constexpr int kRawMargin = 6;  // flow variation (px) inside one halo tile the raw box absorbs
constexpr int kRawMarginL = 13; // ... and the large box (8x32 tiles only)
constexpr int kCBatch = 1;     // direct-gather fallback: channels per software-pipelined batch

enum { PATH_TMA_X2 = 0, PATH_RAW = 1, PATH_DIRECT = 2, PATH_RAWL = 3 };

// TY x TX output tile, KS-way in-CTA channel split, CC channels per pipeline stage, RS raw-box stages
template <int TY, int TX, int KS, int CC_, int RS_>
struct FwdCfg {
  static constexpr int CC = CC_;   // channels per pipeline stage
  static constexpr int ST = (KS == 1) ? CERB_AB_ST1 : 3;   // x1 / warped-x2 stages
  static constexpr bool DIRECT_OUT = (KS == 1);   // accumulators -> global memory without a staged tile
  static constexpr int RS = RS_;   // raw x2 source boxes (TMA) in flight
  static constexpr int NSTRIP = TX / 8;
  static constexpr int COMBOS = TY * kD;
  static constexpr int GROUP = NSTRIP * COMBOS;
  static constexpr int NCONS = GROUP * KS;
  static constexpr int NTHREADS = NCONS + kProducerThreads;
  static constexpr int HY = TY + 2 * kMD, HX = TX + 2 * kMD;
  static constexpr int XS = HX + 4;  // 44 / 28: 8 consecutive rows hit 8 distinct 16-byte bank groups
  static constexpr int NPOS = HY * HX;
  static constexpr int POS_PER_THREAD = (NPOS + kGatherThreads - 1) / kGatherThreads;
  // raw source box: halo + flow-variation margin both sides + the far bilinear tap + the
  // (W/(W-1)) stretch of the training-path warp, and 3 columns so the box can start 16B-aligned
  static constexpr int RAW_H = HY + 2 * kRawMargin + 2;
  static constexpr int RAW_W = (HX + 2 * kRawMargin + 2 + 3 + 3) / 4 * 4;
  // 16-bit inputs: the raw box and the x1 tile arrive by TMA as 16-bit data inside one raw stage
  // (box rows padded to 16 bytes, the x1 tile behind the box); the gather warps convert
  static constexpr int RAW_W16 = (RAW_W + 7) / 8 * 8;
  static constexpr int RAW16_BOX_BYTES = CC * RAW_H * RAW_W16 * 2;
  static constexpr int RAW16_X1_OFF = (RAW16_BOX_BYTES + 127) / 128 * 128;   // bytes
  static constexpr int RAW16_X1_BYTES = CC * TY * TX * 2;
  static constexpr int X1_STAGE = CC * TY * TX;       // floats
  static constexpr int X2_STAGE = CC * HY * XS;       // floats
  // second, larger raw box for tiles whose flow varies more than +-kRawMargin px (a x2 up-sampled decoder flow with
  // sigma = 3 px does): +-kRawMarginL px.  Same stage buffer, its own tensor map; only those tiles pay the extra
  // bytes.  The 8x32 configuration has the room since its output no longer passes through shared memory.
  static constexpr bool BIGBOX = DIRECT_OUT && CC <= 4;
  static constexpr int RAWL_H = BIGBOX ? HY + 2 * kRawMarginL + 2 : RAW_H;
  static constexpr int RAWL_W = BIGBOX ? (HX + 2 * kRawMarginL + 2 + 3 + 3) / 4 * 4 : RAW_W;
  static constexpr int RAW_STAGE = CC * (RAWL_H * RAWL_W > RAW_H * RAW_W ? RAWL_H * RAWL_W : RAW_H * RAW_W);  // floats
  static constexpr int OUT_TILE = kD2 * TY * TX;       // floats, one partial buffer
  static constexpr size_t SMEM_X1 = 0;
  static constexpr size_t SMEM_X2 = SMEM_X1 + sizeof(float) * ST * X1_STAGE;
  static constexpr size_t SMEM_RAW = (SMEM_X2 + sizeof(float) * ST * X2_STAGE + 127) / 128 * 128;
  static constexpr size_t SMEM_OUT = (SMEM_RAW + sizeof(float) * RS * RAW_STAGE + 1023) / 1024 * 1024;
  // cluster channel split: slices of the other CTAs' partial tiles are pushed here (KS > 1 only)
  static constexpr size_t SMEM_RECV = SMEM_OUT + (DIRECT_OUT ? 0 : sizeof(float) * KS * OUT_TILE);
  static constexpr size_t SMEM_RED = SMEM_RECV + (KS > 1 ? sizeof(float) * OUT_TILE : 0);
  static constexpr size_t SMEM_BAR = SMEM_RED + sizeof(int) * 2 * kProducerWarps * 4;  // (kGatherWarps rows used)
  static constexpr size_t SMEM_BYTES = SMEM_BAR + (2 * ST + 2 * RS + 2) * sizeof(uint64_t) + 1024;
  static_assert(NCONS % 32 == 0, "consumer threads must be whole warps");
  static_assert((sizeof(float) * X1_STAGE) % 1024 == 0, "x1 stage must keep 1024-byte alignment");
  static_assert((sizeof(float) * X2_STAGE) % 128 == 0 && (sizeof(float) * RAW_STAGE) % 128 == 0, "TMA dst alignment");
  static_assert(CC % KS == 0, "channel groups must divide the stage");
  static_assert(RAW16_X1_OFF + RAW16_X1_BYTES <= (int)sizeof(float) * RAW_STAGE, "16-bit box + x1 tile must fit a raw stage");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// (y, dy) enumeration per strip.  Lane pairs (2j, g+1), (2j+1, g) read the same x2 row r = 2j + g + 1
// (adjacent-lane duplicate LDS.128 addresses are merged), and 8 consecutive entries -- a quarter-warp,
// the unit a 128-bit shared access is processed in -- hold 8 distinct y for TY = 8, so the x1 loads
// and the epilogue's accumulator stores (chunk swizzle keyed on y) are bank-conflict free.  The last
// group takes the unpaired entries (even y with dy = 0, odd y with dy = 8).
template <int TY>
struct ComboLut {
  unsigned char y[TY * kD];
  unsigned char d[TY * kD];
};
template <int TY>
constexpr ComboLut<TY> make_combo_lut() {
  ComboLut<TY> l{};
  int n = 0;
  for (int g = 0; g < kD - 1; ++g)
    for (int j = 0; j < TY / 2; ++j) {
      l.y[n] = (unsigned char)(2 * j); l.d[n] = (unsigned char)(g + 1); ++n;
      l.y[n] = (unsigned char)(2 * j + 1); l.d[n] = (unsigned char)g; ++n;
    }
  for (int j = 0; j < TY / 2; ++j) {
    l.y[n] = (unsigned char)(2 * j); l.d[n] = 0; ++n;
    l.y[n] = (unsigned char)(2 * j + 1); l.d[n] = (unsigned char)(kD - 1); ++n;
  }
  return l;
}
__constant__ ComboLut<8> c_lut8 = make_combo_lut<8>();
__constant__ ComboLut<4> c_lut4 = make_combo_lut<4>();

template <int TY> __device__ __forceinline__ void combo_of(int idx, int& y, int& d);
template <> __device__ __forceinline__ void combo_of<8>(int idx, int& y, int& d) { y = c_lut8.y[idx]; d = c_lut8.d[idx]; }
template <> __device__ __forceinline__ void combo_of<4>(int idx, int& y, int& d) { y = c_lut4.y[idx]; d = c_lut4.d[idx]; }

// 8 consecutive output pixels of one row: two 16-byte stores (fp32) or one (16-bit)
template <typename T>
__device__ __forceinline__ void store_row8(T* p, const float (&v)[8], bool wide) {
  if constexpr (sizeof(T) == 4) {
    if (wide) {   // one 256-bit store (STG.256, sm_100): a full 32-byte sector per thread, a full 128-byte line per 4 strips
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                   "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    } else {
      asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
      asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    }
  } else {
    uint4 pk;
    T* hv = reinterpret_cast<T*>(&pk);
#pragma unroll
    for (int k = 0; k < 8; ++k) hv[k] = from_f32<T>(v[k]);
    *reinterpret_cast<uint4*>(p) = pk;
  }
}

struct FwdArgs {
  Geom g;
  const void* x1;
  const void* x2;
  const float* flow;
  void* out;
  int off;        // md - pad: output pixel (by,bx) looks at input pixel (by+off, bx+off)
  int tiles_x, tiles_y, total_tiles;   // total_tiles counts (tile, displacement window) work units
  int nwin;       // displacement windows per axis (1 for max_displacement 4)
  int nchunks;    // ceil(C / CC)
  int use_tma_in;   // x1 tiles by TMA
  int use_tma_x2;   // un-warped x2 halo tiles by TMA (flow == null)
  int use_tma_raw;  // raw x2 source boxes by TMA, warp gathered from shared memory
  int use_tma_rawL; // the large raw box is available too
  int use_tma_out;  // output tile by TMA store
  int raw16;        // 16-bit inputs: raw x2 box and x1 tile by TMA as 16-bit data, converted by the gather warps
  int out_vec8;     // output rows may be written with 16-byte stores (base and N/C/H strides 16-byte aligned)
  // fused flow up-sampling: coarse flow in, up-sampled flow out (cflow == nullptr: off)
  const float* cflow;
  long long cfs[3];
  int Hc, Wc;
  float up_sy, up_sx;   // ATen's align_corners=True scale (in - 1) / (out - 1), fp32
  float* flow_up;
  long long fus[3];
  int csplit_log2;
  int csplit;       // CTAs per cluster sharing one tile, each taking a slice of the channel chunks (1 = off)
  int dbg_iter;
  long long* dbg;   // optional per-CTA clock64() trace (cerb_debug_set_trace_buffer)
  unsigned long long* path_ctr;   // optional: [PATH_*] tile counters of the warped gather (cerb_debug_set_path_counters)
};

// 16-byte chunk swizzle of a [rows][TX] fp32 tile: CU_TENSOR_MAP_SWIZZLE_128B for TX == 32
// (chunk ^= row & 7 inside 1024-byte atoms); identity for narrower tiles.
template <int TX>
__device__ __forceinline__ int swz_chunk(int row, int chunk) {
  if constexpr (TX == 32) return chunk ^ (row & 7);
  else return chunk;
}
// 16-wide tiles have no TMA swizzle; when the tile is not going through TMA (cluster split) the
// 4 chunks of a row are XOR-ed with two more row bits so the accumulator stores of one warp spread
// over the banks (8-way conflicts otherwise).  Involution: the same function maps back.
template <int TX>
__device__ __forceinline__ int swz_partial(int row, int chunk, bool on) {
  if constexpr (TX == 16) return on ? (chunk ^ ((row >> 1) & 3)) : chunk;
  else return swz_chunk<TX>(row, chunk);
}

// work unit -> batch item, displacement-window origin (in [0, D-9]) and tile origin
struct Unit {
  int n, woy, wox, by0, bx0;
};
template <int TY, int TX>
__device__ __forceinline__ Unit decode_unit(const FwdArgs& a, int tile) {
  Unit u;
  const int per_img = a.tiles_x * a.tiles_y;
  int t = tile;
  int win = 0;
  if (a.nwin > 1) {
    const int per_win = per_img * a.g.B;
    win = t / per_win;
    t -= win * per_win;
  }
  u.n = t / per_img;
  const int trem = t - u.n * per_img;
  u.by0 = (trem / a.tiles_x) * TY;
  u.bx0 = (trem % a.tiles_x) * TX;
  const int wy = win / a.nwin, wx = win - wy * a.nwin;
  u.woy = min(wy * 8, a.g.D - kD);
  u.wox = min(wx * 8, a.g.D - kD);
  return u;
}

// ------------------------------------------------------------------ fast kernel ----------
// UP: the flow is up-sampled on the fly from the next-coarser level (a separate instantiation, so the
// plain path's prologue is not perturbed -- folding it in behind a runtime test cost 0.5 us per level)
template <typename T, int TY, int TX, int KS, int CC, int RS, bool UP>
__global__ void __launch_bounds__(FwdCfg<TY, TX, KS, CC, RS>::NTHREADS, 1)
warp_corr_fwd_kernel(const FwdArgs a, const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                     const __grid_constant__ CUtensorMap tm_raw, const __grid_constant__ CUtensorMap tm_rawL,
                     const __grid_constant__ CUtensorMap tm_out,
                     const __grid_constant__ CUtensorMap tm_outc) {
  using Cfg = FwdCfg<TY, TX, KS, CC, RS>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // keep the pointer derived from smem_raw (so loads/stores stay LDS/STS) while forcing the
  // 1024-byte alignment the 128B TMA swizzle atoms need
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* x1s = (float*)(smem + Cfg::SMEM_X1);
  float* x2s = (float*)(smem + Cfg::SMEM_X2);
  float* raws = (float*)(smem + Cfg::SMEM_RAW);
  float* outs = (float*)(smem + Cfg::SMEM_OUT);
  int* red = (int*)(smem + Cfg::SMEM_RED);
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::SMEM_BAR);
  uint64_t* empty_bar = full_bar + Cfg::ST;
  uint64_t* raw_full = empty_bar + Cfg::ST;
  uint64_t* raw_empty = raw_full + RS;

  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const T* __restrict__ x1 = (const T*)a.x1;
  const T* __restrict__ x2 = (const T*)a.x2;

  if (tid == 0) {
    for (int s = 0; s < Cfg::ST; ++s) {
      mbar_init(&full_bar[s], 1 + kGatherWarps);  // TMA thread (+tx bytes) and one arrival per gather warp
      mbar_init(&empty_bar[s], Cfg::NCONS / 32);
    }
    for (int s = 0; s < RS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], kGatherWarps);
    }
    fence_barrier_init();
    if (a.use_tma_in) tma_prefetch_desc(&tm_x1);
    if (a.use_tma_x2) tma_prefetch_desc(&tm_x2);
    if (a.use_tma_raw) tma_prefetch_desc(&tm_raw);
    if (a.use_tma_rawL) tma_prefetch_desc(&tm_rawL);
    if (a.use_tma_out) { tma_prefetch_desc(&tm_out); if (KS == 1) tma_prefetch_desc(&tm_outc); }
  }
  __syncthreads();
  // cluster channel split: CTA `crank` of a cluster of `csplit` CTAs handles chunks [ck_begin, ck_end)
  // of the cluster's single tile (csplit is a power of two: shifts, not divisions -- this runs on
  // every thread's critical path).  The partial tiles meet in the epilogue: each CTA pushes slice r
  // of its partial tile into CTA r's shared memory, one hardware cluster barrier, each CTA sums
  // and stores its slice.
  const int S = a.csplit, Slog = a.csplit_log2;
  const int crank = S > 1 ? (int)cluster_ctarank() : 0;
  // Distributed shared memory may only be touched once every CTA of the cluster is running: a split-phase
  // cluster barrier -- arrive here (non-blocking), wait right before the first remote store in the epilogue,
  // by which time it has long completed.
  if (S > 1) cluster_arrive_release();
  const int tile0 = (int)blockIdx.x >> Slog, tile_step = (int)gridDim.x >> Slog;
  const int ck_begin = (crank * a.nchunks) >> Slog, ck_end = ((crank + 1) * a.nchunks) >> Slog;
  int titer = 0;
  if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * kTraceSlots + 0] = clock64();
  // PDL: everything above (barrier init, descriptor prefetch, cluster sync) overlapped the tail of
  // the previous kernel in the stream; inputs may be its outputs, so wait before the first read.
  pdl_wait();

  int stage = 0;
  uint32_t phase = 0;

  if (tid >= Cfg::NCONS) {
    // =========================== STAGING WARPS ===========================
    // producer warp 0: one elected lane issues every TMA copy (a TMA issue blocks the issuing
    // thread for hundreds of cycles, so it must not share a warp with the gather);
    // producer warps 1..6: bilinear gather of the warped x2 halo tile.
    const int pt = tid - Cfg::NCONS;
    const int pwarp = pt >> 5;
    const int lane = pt & 31;
    const bool warped = UP || a.flow != nullptr;
    const bool raw16 = sizeof(T) == 2 && a.raw16;           // 16-bit inputs staged through the raw stage
    const bool reduce_bbox = (warped || raw16) && a.use_tma_raw;
    int red_par = 0;

    if (pwarp == kTmaWarp2) {
      // ---------------------------- TMA warp: x1 tiles, un-warped x2 halo tiles ----------------------------
      // Depends on nothing but free stages, so it runs ahead of the flow / bounding-box work of the tile.
      int xs = 0;
      uint32_t xphase = 0;
      const bool plain_x2 = !warped && a.use_tma_x2;
      for (int tile = tile0; tile < a.total_tiles; tile += tile_step) {
        const Unit un = decode_unit<TY, TX>(a, tile);
        const int n = un.n;
        const int iy0 = un.by0 + a.off, ix0 = un.bx0 + a.off;
        const int qy0 = iy0 + un.woy - g.md, qx0 = ix0 + un.wox - g.md;
        if (lane == 0) {
          for (int ck = ck_begin; ck < ck_end; ++ck) {
            mbar_wait(&empty_bar[xs], xphase ^ 1);
            if (ck < 6) CERB_TRACE(64 + 8 * ck + 3);
            const uint32_t tx = (a.use_tma_in ? (uint32_t)(sizeof(float) * Cfg::X1_STAGE) : 0u) +
                                (plain_x2 ? (uint32_t)(sizeof(float) * Cfg::X2_STAGE) : 0u);
            if (tx) mbar_arrive_expect_tx(&full_bar[xs], tx);
            else mbar_arrive(&full_bar[xs]);
            if (a.use_tma_in) tma_load_4d(x1s + xs * Cfg::X1_STAGE, &tm_x1, &full_bar[xs], ix0, iy0, ck * CC, n);
            if (plain_x2) tma_load_4d(x2s + xs * Cfg::X2_STAGE, &tm_x2, &full_bar[xs], qx0, qy0, ck * CC, x2_item(g, n));
            if (++xs == Cfg::ST) { xs = 0; xphase ^= 1; }
            if (ck < 6) CERB_TRACE(64 + 8 * ck + 4);
            if (ck < 8) CERB_TRACE(2 + ck);
          }
        }
        __syncwarp();
        ++titer;
      }
      pdl_launch_dependents();  // this role has issued its last copy: let the next kernel start launching
      if (S > 1) { __syncwarp(); cluster_wait_acquire(); cluster_sync_all(); }  // start-up phase, then the consumers' barrier (every thread of the cluster arrives)
    } else if (pwarp == kTmaWarp) {
      // ---------------------------- TMA warp: raw x2 source boxes ----------------------------
      int ri = 0;
      uint32_t riphase = 0;
      for (int tile = tile0; tile < a.total_tiles; tile += tile_step) {
        const Unit un = decode_unit<TY, TX>(a, tile);
        const int n = un.n;
        const int iy0 = un.by0 + a.off, ix0 = un.bx0 + a.off;
        // halo tile of the second map: displacement window origin w covers displacements w-md .. w-md+8
        const int qy0 = iy0 + un.woy - g.md, qx0 = ix0 + un.wox - g.md;
        int path = PATH_DIRECT, ox = 0, oy = 0;
        // 16-bit inputs without a flow: the box is the halo tile itself, known without any bounding box
        // (origin aligned down to 8 elements; coordinates outside the image are zero-filled by TMA)
        const bool plain16 = raw16 && !warped && (qx0 & 3) == 0;
        if (plain16) { path = PATH_RAW; ox = qx0 & ~7; oy = qy0; }
        if (reduce_bbox && !plain16) {
          named_bar_sync(2, kBboxThreads);
          const int* rp = red + red_par * (kGatherWarps * 4);
          int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
          for (int w = 0; w < kGatherWarps; ++w) {
            xmin = min(xmin, rp[w * 4 + 0]); xmax = max(xmax, rp[w * 4 + 1]);
            ymin = min(ymin, rp[w * 4 + 2]); ymax = max(ymax, rp[w * 4 + 3]);
          }
          red_par ^= 1;
          ox = xmin & ~3;
          oy = ymin;
          if (raw16) ox = xmin & ~7;
          if (xmin <= xmax && xmax - ox < (raw16 ? Cfg::RAW_W16 : Cfg::RAW_W) && ymax - oy < Cfg::RAW_H) path = PATH_RAW;
          else if (Cfg::BIGBOX && a.use_tma_rawL && xmin <= xmax && xmax - ox < Cfg::RAWL_W && ymax - oy < Cfg::RAWL_H) path = PATH_RAWL;
        }
        if (lane == 0 && path == PATH_RAWL) {
          for (int ck = ck_begin; ck < ck_end; ++ck) {
            mbar_wait(&raw_empty[ri], riphase ^ 1);
            mbar_arrive_expect_tx(&raw_full[ri], (uint32_t)(sizeof(float) * CC * Cfg::RAWL_H * Cfg::RAWL_W));
            tma_load_4d(raws + ri * Cfg::RAW_STAGE, &tm_rawL, &raw_full[ri], ox, oy, ck * CC, x2_item(g, n));
            if (++ri == RS) { ri = 0; riphase ^= 1; }
          }
        }
        if (lane == 0 && path == PATH_RAW) {
          for (int ck = ck_begin; ck < ck_end; ++ck) {
            if (ck < 6) CERB_TRACE(64 + 8 * ck + 0);
            mbar_wait(&raw_empty[ri], riphase ^ 1);
            if (ck < 6) CERB_TRACE(64 + 8 * ck + 1);
            if (raw16) {
              unsigned char* rs = (unsigned char*)(raws + ri * Cfg::RAW_STAGE);
              mbar_arrive_expect_tx(&raw_full[ri], (uint32_t)(Cfg::RAW16_BOX_BYTES + Cfg::RAW16_X1_BYTES));
              tma_load_4d(rs, &tm_raw, &raw_full[ri], ox, oy, ck * CC, x2_item(g, n));
              tma_load_4d(rs + Cfg::RAW16_X1_OFF, &tm_x1, &raw_full[ri], ix0, iy0, ck * CC, n);
            } else {
              mbar_arrive_expect_tx(&raw_full[ri], (uint32_t)(sizeof(float) * CC * Cfg::RAW_H * Cfg::RAW_W));
              tma_load_4d(raws + ri * Cfg::RAW_STAGE, &tm_raw, &raw_full[ri], ox, oy, ck * CC, x2_item(g, n));
            }
            if (++ri == RS) { ri = 0; riphase ^= 1; }
            if (ck < 6) CERB_TRACE(64 + 8 * ck + 2);
          }
        }
        __syncwarp();
        ++titer;
      }
      pdl_launch_dependents();
      if (S > 1) { __syncwarp(); cluster_wait_acquire(); cluster_sync_all(); }
    } else {
      // ---------------------------- gather warps ----------------------------
      const int gw = pwarp - (pwarp > kTmaWarp ? 1 : 0) - (pwarp > kTmaWarp2 ? 1 : 0);  // 0 .. kGatherWarps-1
      const int gt = gw * 32 + lane;                        // 0 .. kGatherThreads-1
      int rc = 0;
      uint32_t rcphase = 0;

      // x1 tile without TMA (any dtype / alignment): cooperative loads into the swizzled layout
      auto coop_x1 = [&](float* x1dst, int ix0, int iy0, int c0, int n) {
        constexpr int PER = (Cfg::X1_STAGE + kGatherThreads - 1) / kGatherThreads;
        float v[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int e = gt + k * kGatherThreads;
          const int c = e / (TY * TX), rem = e - c * (TY * TX);
          const int y = rem / TX, x = rem - y * TX;
          const int iy = iy0 + y, ix = ix0 + x, ch = c0 + c;
          v[k] = 0.f;
          if (e < Cfg::X1_STAGE && ch < g.C && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
            v[k] = ldg_f32(x1 + (long long)n * g.x1s[0] + (long long)ch * g.x1s[1] + (long long)iy * g.x1s[2] + ix);
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int e = gt + k * kGatherThreads;
          if (e < Cfg::X1_STAGE) {
            const int c = e / (TY * TX), rem = e - c * (TY * TX);
            const int y = rem / TX, x = rem - y * TX;
            const int row = c * TY + y;
            x1dst[row * TX + swz_chunk<TX>(row, x >> 2) * 4 + (x & 3)] = v[k];
          }
        }
      };
      const AxisConst axis_x = make_axis(g.W), axis_y = make_axis(g.H);
      auto publish_stage = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == Cfg::ST) { stage = 0; phase ^= 1; }
      };

      for (int tile = tile0; tile < a.total_tiles; tile += tile_step) {
        const Unit un = decode_unit<TY, TX>(a, tile);
        const int n = un.n;
        // input-frame origin of the x1 tile and of the x2 halo tile (shifted with the displacement window)
        const int iy0 = un.by0 + a.off, ix0 = un.bx0 + a.off;
        const int qy0 = iy0 + un.woy - g.md, qx0 = ix0 + un.wox - g.md;

        // ------------- un-warped second map: the TMA warp loads the halo tile as a box -------------
        if (!warped && a.use_tma_x2) {
          for (int ck = ck_begin; ck < ck_end; ++ck) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (!a.use_tma_in) coop_x1(x1s + stage * Cfg::X1_STAGE, ix0, iy0, ck * CC, n);
            publish_stage();
          }
          ++titer;
          continue;
        }

        // ------------- 16-bit inputs, no flow: widen the TMA-loaded halo tile (and the x1 tile) -------------
        if constexpr (sizeof(T) == 2) {
          if (raw16 && !warped && (qx0 & 3) == 0) {
            const int dxo = qx0 - (qx0 & ~7);   // 0 or 4: column of the halo origin inside the box
            for (int ck = ck_begin; ck < ck_end; ++ck) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_wait(&raw_full[rc], rcphase);
              const unsigned char* rs = (const unsigned char*)(raws + rc * Cfg::RAW_STAGE);
              const uint4* x1r = reinterpret_cast<const uint4*>(rs + Cfg::RAW16_X1_OFF);
              float* x1dst = x1s + stage * Cfg::X1_STAGE;
              for (int e8 = gt; e8 < Cfg::X1_STAGE / 8; e8 += kGatherThreads) {
                const uint4 pk = x1r[e8];
                const T* hv = reinterpret_cast<const T*>(&pk);
                const int e = e8 * 8, row = e / TX, x = e - row * TX;   // row = c * TY + y
                *reinterpret_cast<float4*>(x1dst + row * TX + swz_chunk<TX>(row, x >> 2) * 4) =
                    make_float4(to_f32<T>(hv[0]), to_f32<T>(hv[1]), to_f32<T>(hv[2]), to_f32<T>(hv[3]));
                *reinterpret_cast<float4*>(x1dst + row * TX + swz_chunk<TX>(row, (x >> 2) + 1) * 4) =
                    make_float4(to_f32<T>(hv[4]), to_f32<T>(hv[5]), to_f32<T>(hv[6]), to_f32<T>(hv[7]));
              }
              const T* __restrict__ src16 = reinterpret_cast<const T*>(rs);
              float* __restrict__ x2dst = x2s + stage * Cfg::X2_STAGE;
              constexpr int V_ROW = Cfg::HX / 4, U = Cfg::HY * V_ROW;
              for (int u = gt; u < CC * U; u += kGatherThreads) {
                const int c = u / U, r = u - c * U;
                const int hy = r / V_ROW, v = r - hy * V_ROW;
                const uint2 pk = *reinterpret_cast<const uint2*>(src16 + c * (Cfg::RAW_H * Cfg::RAW_W16) + hy * Cfg::RAW_W16 +
                                                                 dxo + 4 * v);
                const T* hv = reinterpret_cast<const T*>(&pk);
                *reinterpret_cast<float4*>(x2dst + c * (Cfg::HY * Cfg::XS) + hy * Cfg::XS + 4 * v) =
                    make_float4(to_f32<T>(hv[0]), to_f32<T>(hv[1]), to_f32<T>(hv[2]), to_f32<T>(hv[3]));
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&raw_empty[rc]);
              if (++rc == RS) { rc = 0; rcphase ^= 1; }
              publish_stage();
            }
            ++titer;
            continue;
          }
        }

        // ------------- per-position sampling data, fixed for the whole tile -------------
        Taps taps[Cfg::POS_PER_THREAD];   // off[] first holds {x0, x1c, y0, y1c}, then offsets
        int sdst[Cfg::POS_PER_THREAD];    // smem float offset inside a channel plane (surplus slots: a padding column)
        unsigned valid_mask = 0;
        int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
        // every flow vector this thread needs is requested before the first one is used:
        // unconditional loads from clamped (always valid) addresses -- loads left inside the
        // validity branches are issued one round trip at a time
        float fu[Cfg::POS_PER_THREAD], fv[Cfg::POS_PER_THREAD];
        if constexpr (UP) {
          // flow = interpolate(2 * coarse, scale_factor=2, bilinear, align_corners=True), evaluated per halo
          // position with ATen's arithmetic (upsample_bilinear2d: src = scale * dst, lambda = src - int(src));
          // doubling commutes exactly with the blend.  Tile-interior positions also write the result out.
          const float* cn = a.cflow + (long long)n * a.cfs[0];
          float cv[Cfg::POS_PER_THREAD][8], lam[Cfg::POS_PER_THREAD][2];
#pragma unroll
          for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
            const int i = gt + j * kGatherThreads;
            const int hy = i / Cfg::HX, hx = i - hy * Cfg::HX;
            const int cy = min(max(qy0 + hy, 0), g.H - 1), cx = min(max(qx0 + hx, 0), g.W - 1);
            const float h1r = __fmul_rn(a.up_sy, (float)cy), w1r = __fmul_rn(a.up_sx, (float)cx);
            const int h1 = (int)h1r, w1 = (int)w1r;
            const int h1p = (h1 < a.Hc - 1) ? 1 : 0, w1p = (w1 < a.Wc - 1) ? 1 : 0;
            lam[j][0] = __fsub_rn(h1r, (float)h1);
            lam[j][1] = __fsub_rn(w1r, (float)w1);
            const float* p0 = cn + (long long)h1 * a.cfs[2] + w1;
            const float* p1 = p0 + (long long)h1p * a.cfs[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              cv[j][4 * c + 0] = __ldg(p0 + c * a.cfs[1]);
              cv[j][4 * c + 1] = __ldg(p0 + c * a.cfs[1] + w1p);
              cv[j][4 * c + 2] = __ldg(p1 + c * a.cfs[1]);
              cv[j][4 * c + 3] = __ldg(p1 + c * a.cfs[1] + w1p);
            }
          }
#pragma unroll
          for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
            const float h1l = lam[j][0], w1l = lam[j][1];
            const float h0l = __fsub_rn(1.f, h1l), w0l = __fsub_rn(1.f, w1l);
            float r[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              // h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11) as nvcc contracts it in ATen
              const float t0 = __fmaf_rn(w0l, cv[j][4 * c + 0], __fmul_rn(w1l, cv[j][4 * c + 1]));
              const float t1 = __fmaf_rn(w0l, cv[j][4 * c + 2], __fmul_rn(w1l, cv[j][4 * c + 3]));
              r[c] = __fmul_rn(2.f, __fmaf_rn(h0l, t0, __fmul_rn(h1l, t1)));
            }
            fu[j] = r[0];
            fv[j] = r[1];
            const int i = gt + j * kGatherThreads;
            const int hy = i / Cfg::HX, hx = i - hy * Cfg::HX;
            const int qy = qy0 + hy, qx = qx0 + hx;
            if (i < Cfg::NPOS && qy >= iy0 && qy < iy0 + TY && qx >= ix0 && qx < ix0 + TX && qy < g.H && qx < g.W) {
              float* up = a.flow_up + (long long)n * a.fus[0] + (long long)qy * a.fus[2] + qx;
              up[0] = r[0];
              up[a.fus[1]] = r[1];
            }
          }
        }
        if (!UP && warped) {
          const float* fn = a.flow + (long long)n * g.fls[0];
#pragma unroll
          for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
            const int i = gt + j * kGatherThreads;
            const int hy = i / Cfg::HX, hx = i - hy * Cfg::HX;
            const int cy = min(max(qy0 + hy, 0), g.H - 1), cx = min(max(qx0 + hx, 0), g.W - 1);
            const float* fp = fn + (long long)cy * g.fls[2] + cx;
            fu[j] = __ldg(fp);
            fv[j] = __ldg(fp + g.fls[1]);
          }
        }
        // branch-free per position (selects on validity): the POS_PER_THREAD chains interleave
#pragma unroll
        for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
          const int i = gt + j * kGatherThreads;
          const int hy = i / Cfg::HX, hx = i - hy * Cfg::HX;
          const int qy = qy0 + hy, qx = qx0 + hx;
          const bool in_tile = i < Cfg::NPOS;
          const bool valid = in_tile && qy >= 0 && qy < g.H && qx >= 0 && qx < g.W;
          const int cy = min(max(qy, 0), g.H - 1), cx = min(max(qx, 0), g.W - 1);
          sdst[j] = in_tile ? hy * Cfg::XS + hx : Cfg::HX;   // surplus slots write a padding column nobody reads
          float sx = (float)cx, sy = (float)cy;   // un-warped: the pixel itself (weights 1,0,0,0)
          if (warped) {
            bool in_x, in_y;
            sx = sample_pos(cx, fu[j], axis_x, g.warp_mode, in_x);
            sy = sample_pos(cy, fv[j], axis_y, g.warp_mode, in_y);
          }
          const Taps tp = make_taps(sx, sy, g.H, g.W, 0);  // hstride 0: off[] = {x0, x1c, x0, x1c}
          const int y0 = (int)floorf(sy);
          const int y1c = (y0 + 1 < g.H) ? y0 + 1 : y0;
          taps[j].off[0] = valid ? tp.off[0] : 0;
          taps[j].off[1] = valid ? tp.off[1] : 0;
          taps[j].off[2] = valid ? y0 : 0;
          taps[j].off[3] = valid ? y1c : 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) taps[j].w[k] = valid ? tp.w[k] : 0.f;
          if (valid) valid_mask |= 1u << j;
          xmin = valid ? min(xmin, tp.off[0]) : xmin; xmax = valid ? max(xmax, tp.off[1]) : xmax;
          ymin = valid ? min(ymin, y0) : ymin; ymax = valid ? max(ymax, y1c) : ymax;
        }
        if (gt == 0) CERB_TRACE(58);

        // ------------- does the tile's sampling footprint fit the raw TMA box? -------------
        int path = PATH_DIRECT, ox = 0, oy = 0;
        if (reduce_bbox) {
          xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
          ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
          int* rp = red + red_par * (kGatherWarps * 4);
          if (lane == 0) {
            rp[gw * 4 + 0] = xmin; rp[gw * 4 + 1] = xmax;
            rp[gw * 4 + 2] = ymin; rp[gw * 4 + 3] = ymax;
          }
          named_bar_sync(2, kBboxThreads);
#pragma unroll
          for (int w = 0; w < kGatherWarps; ++w) {
            xmin = min(xmin, rp[w * 4 + 0]); xmax = max(xmax, rp[w * 4 + 1]);
            ymin = min(ymin, rp[w * 4 + 2]); ymax = max(ymax, rp[w * 4 + 3]);
          }
          red_par ^= 1;
          ox = raw16 ? (xmin & ~7) : (xmin & ~3);  // TMA box starts must be 16-byte aligned
          oy = ymin;
          if (xmin <= xmax && xmax - ox < (raw16 ? Cfg::RAW_W16 : Cfg::RAW_W) && ymax - oy < Cfg::RAW_H) path = PATH_RAW;
          else if (Cfg::BIGBOX && a.use_tma_rawL && xmin <= xmax && xmax - ox < Cfg::RAWL_W && ymax - oy < Cfg::RAWL_H) path = PATH_RAWL;
        }
#pragma unroll
        for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
          const int x0 = taps[j].off[0], x1c = taps[j].off[1], y0 = taps[j].off[2], y1c = taps[j].off[3];
          int r0, r1;
          const int raw_w = path == PATH_RAWL ? Cfg::RAWL_W : (raw16 ? Cfg::RAW_W16 : Cfg::RAW_W);
          if (path == PATH_RAW || path == PATH_RAWL) { r0 = (y0 - oy) * raw_w - ox; r1 = (y1c - oy) * raw_w - ox; }
          else { r0 = (int)(y0 * g.x2s[2]); r1 = (int)(y1c * g.x2s[2]); }
          const bool ok = (valid_mask >> j) & 1u;  // positions without a sample read offset 0 (always inside the source)
          taps[j].off[0] = ok ? r0 + x0 : 0; taps[j].off[1] = ok ? r0 + x1c : 0;
          taps[j].off[2] = ok ? r1 + x0 : 0; taps[j].off[3] = ok ? r1 + x1c : 0;
        }
        if (gt == 0) CERB_TRACE(1);
        if (gt == 0 && a.path_ctr != nullptr) atomicAdd(&a.path_ctr[path], 1ull);

        if (path == PATH_RAW || path == PATH_RAWL) {
          // ------------- warp gathered from the raw source box in shared memory -------------
          if constexpr (sizeof(T) == 4) {
            constexpr int CB = 1;
            constexpr int NBAT = CC / CB;
            static_assert(CC % CB == 0 && NBAT % 2 == 0, "an even number of channel batches per stage");
            // channel-plane stride of the box in use (two box sizes): a runtime value, so the tap addresses advance by it
            // per channel (16 adds) instead of sitting in immediates -- two unrolled copies of this loop spilled
            const uint32_t raw_plane = (uint32_t)sizeof(float) * (uint32_t)((Cfg::BIGBOX && path == PATH_RAWL) ? Cfg::RAWL_H * Cfg::RAWL_W : Cfg::RAW_H * Cfg::RAW_W);
            constexpr uint32_t kDstPlane = sizeof(float) * Cfg::HY * Cfg::XS;
            const uint32_t raw0 = smem_u32(raws), dst0 = smem_u32(x2s);
            uint32_t ta[Cfg::POS_PER_THREAD][4], da[Cfg::POS_PER_THREAD];
            float tv[2][CB][Cfg::POS_PER_THREAD][4];
            auto set_sources = [&](int rstage) {
              const uint32_t rbase = raw0 + (uint32_t)rstage * (uint32_t)(sizeof(float) * Cfg::RAW_STAGE);
#pragma unroll
              for (int j = 0; j < Cfg::POS_PER_THREAD; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) ta[j][k] = rbase + ((uint32_t)taps[j].off[k] << 2);
            };
            auto load_batch = [&](int, float (&dst)[CB][Cfg::POS_PER_THREAD][4]) {   // next CB channels of the current box
#pragma unroll
              for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                for (int j = 0; j < Cfg::POS_PER_THREAD; ++j)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    dst[cb][j][k] = lds_f32(ta[j][k]);
                    ta[j][k] += raw_plane;
                  }
            };
            if (ck_begin < ck_end) {
              mbar_wait(&raw_full[rc], rcphase);
              set_sources(rc);
              load_batch(0, tv[0]);
            }
            for (int ck = ck_begin; ck < ck_end; ++ck) {
              int rn = rc + 1;
              uint32_t rnphase = rcphase;
              if (rn == RS) { rn = 0; rnphase ^= 1; }
#pragma unroll
              for (int b = 0; b < NBAT; ++b) {
                if (b + 1 < NBAT) {
                  load_batch(b + 1, tv[(b + 1) & 1]);
                } else if (ck + 1 < ck_end) {
                  mbar_wait(&raw_full[rn], rnphase);
                  set_sources(rn);
                  load_batch(0, tv[0]);
                }
                if (b == 0) {
                  mbar_wait(&empty_bar[stage], phase ^ 1);
                  if (!a.use_tma_in) coop_x1(x1s + stage * Cfg::X1_STAGE, ix0, iy0, ck * CC, n);
                  const uint32_t dbase = dst0 + (uint32_t)stage * (uint32_t)(sizeof(float) * Cfg::X2_STAGE);
#pragma unroll
                  for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) da[j] = dbase + ((uint32_t)sdst[j] << 2);
                  if (gt == 0 && ck < 4) CERB_TRACE(45 + 3 * ck);
                }
#pragma unroll
                for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                  for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
                    const float* v = tv[b & 1][cb][j];
                    const float r = ((valid_mask >> j) & 1u) ? blend(v[0], v[1], v[2], v[3], taps[j]) : 0.f;
                    sts_f32(da[j] + (uint32_t)(b * CB + cb) * kDstPlane, r);
                  }
              }
              __syncwarp();
              if (gt == 0 && ck < 4) CERB_TRACE(46 + 3 * ck);
              if (lane == 0) mbar_arrive(&raw_empty[rc]);
              rc = rn; rcphase = rnphase;
              publish_stage();
            }
            ++titer;
            continue;
          } else {
          for (int ck = ck_begin; ck < ck_end; ++ck) {
            if (lane == 0 && (ck == 2 || ck == 3)) CERB_TRACE(112 + (ck - 2) * 40 + gw * 6 + 0);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (lane == 0 && (ck == 2 || ck == 3)) CERB_TRACE(112 + (ck - 2) * 40 + gw * 6 + 1);
            if (!a.use_tma_in && !raw16) coop_x1(x1s + stage * Cfg::X1_STAGE, ix0, iy0, ck * CC, n);
            if (gt == 0 && ck < 4) CERB_TRACE(44 + 3 * ck);
            mbar_wait(&raw_full[rc], rcphase);
            if (lane == 0 && (ck == 2 || ck == 3)) CERB_TRACE(112 + (ck - 2) * 40 + gw * 6 + 2);
            if (gt == 0 && ck < 4) CERB_TRACE(45 + 3 * ck);
            const float* __restrict__ src = raws + rc * Cfg::RAW_STAGE;
            float* __restrict__ x2dst = x2s + stage * Cfg::X2_STAGE;
            if constexpr (sizeof(T) == 2) {
              if (raw16) {
                // x1 tile: 16-bit [c][y][x] behind the box -> fp32 stage in the consumers' swizzled layout
                const unsigned char* rs = (const unsigned char*)src;
                const uint4* x1r = reinterpret_cast<const uint4*>(rs + Cfg::RAW16_X1_OFF);
                float* x1dst = x1s + stage * Cfg::X1_STAGE;
                for (int e8 = gt; e8 < Cfg::X1_STAGE / 8; e8 += kGatherThreads) {
                  const uint4 pk = x1r[e8];
                  const T* hv = reinterpret_cast<const T*>(&pk);
                  const int e = e8 * 8, row = e / TX, x = e - row * TX;   // row = c * TY + y
                  float4 lo = make_float4(to_f32<T>(hv[0]), to_f32<T>(hv[1]), to_f32<T>(hv[2]), to_f32<T>(hv[3]));
                  float4 hi = make_float4(to_f32<T>(hv[4]), to_f32<T>(hv[5]), to_f32<T>(hv[6]), to_f32<T>(hv[7]));
                  *reinterpret_cast<float4*>(x1dst + row * TX + swz_chunk<TX>(row, x >> 2) * 4) = lo;
                  *reinterpret_cast<float4*>(x1dst + row * TX + swz_chunk<TX>(row, (x >> 2) + 1) * 4) = hi;
                }
                // taps from the 16-bit box
                const T* __restrict__ src16 = reinterpret_cast<const T*>(rs);
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                  const T* __restrict__ sp = src16 + c * (Cfg::RAW_H * Cfg::RAW_W16);
                  float* __restrict__ dp = x2dst + c * (Cfg::HY * Cfg::XS);
                  float tv16[Cfg::POS_PER_THREAD][4];
#pragma unroll
                  for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tv16[j][k] = to_f32<T>(sp[taps[j].off[k]]);
                  }
#pragma unroll
                  for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
                    const float r = ((valid_mask >> j) & 1u) ? blend(tv16[j][0], tv16[j][1], tv16[j][2], tv16[j][3], taps[j]) : 0.f;
                    dp[sdst[j]] = r;
                  }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty[rc]);
                if (++rc == RS) { rc = 0; rcphase ^= 1; }
                publish_stage();
                continue;
              }
            }
            // Lean gather: everything that does not change with the channel is in registers as a shared-memory
            // byte address (tap sources, destination) -- per channel a tap is one LDS with an immediate offset,
            // a sample 1 FMUL + 3 FFMA + 1 select + 1 STS.  (The plain indexed form cost ~40 instructions per
            // sample, mostly address arithmetic and convergence barriers around predicated stores: the gather
            // warps' instruction stream, not shared-memory bandwidth, was what the consumers waited for.)
            // Software pipeline over channel batches: the taps of batch b+1 are requested before batch b is
            // blended and stored (shared loads are not hoisted above shared stores).
#ifdef CERB_AB_CB
            constexpr int CB = Cfg::POS_PER_THREAD >= 4 ? CERB_AB_CB : (Cfg::POS_PER_THREAD >= 2 ? 2 : 4);
#else
            constexpr int CB = Cfg::POS_PER_THREAD >= 4 ? 2 : (Cfg::POS_PER_THREAD >= 2 ? 2 : 4);
#endif
            constexpr int NBAT = CC / CB;
            constexpr bool kPipe = CERB_AB_PIPE != 0 && NBAT > 1;
            static_assert(CC % CB == 0, "channel batches must divide the stage");
            constexpr uint32_t kRawPlane = sizeof(float) * Cfg::RAW_H * Cfg::RAW_W;
            constexpr uint32_t kDstPlane = sizeof(float) * Cfg::HY * Cfg::XS;
            const uint32_t rbase = smem_u32(src), dbase = smem_u32(x2dst);
            uint32_t ta[Cfg::POS_PER_THREAD][4], da[Cfg::POS_PER_THREAD];
#pragma unroll
            for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ta[j][k] = rbase + ((uint32_t)taps[j].off[k] << 2);  // off = 0 for invalid positions
              da[j] = dbase + ((uint32_t)sdst[j] << 2);
            }
            float tv[2][CB][Cfg::POS_PER_THREAD][4];
            auto load_batch = [&](int b, float (&dst)[CB][Cfg::POS_PER_THREAD][4]) {
#pragma unroll
              for (int cb = 0; cb < CB; ++cb) {
#pragma unroll
                for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
#pragma unroll
#if defined(CERB_X_NOGLDS) || defined(CERB_X_NOGATHER)
                  for (int k = 0; k < 4; ++k) dst[cb][j][k] = __int_as_float(ta[j][k] + cb);
#else
                  for (int k = 0; k < 4; ++k) dst[cb][j][k] = lds_f32(ta[j][k] + (uint32_t)(b * CB + cb) * kRawPlane);
#endif
                }
              }
            };
            if (kPipe) load_batch(0, tv[0]);
#pragma unroll
            for (int b = 0; b < NBAT; ++b) {
              if (kPipe) { if (b + 1 < NBAT) load_batch(b + 1, tv[(b + 1) & 1]); }
              else load_batch(b, tv[b & 1]);
#pragma unroll
              for (int cb = 0; cb < CB; ++cb) {
#pragma unroll
                for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
                  const float* v = tv[b & 1][cb][j];
                  const float r = ((valid_mask >> j) & 1u) ? blend(v[0], v[1], v[2], v[3], taps[j]) : 0.f;
#if defined(CERB_X_NOGATHER)
                  if (r == 12345.678f) sts_f32(da[j] + (uint32_t)(b * CB + cb) * kDstPlane, r);
#else
                  sts_f32(da[j] + (uint32_t)(b * CB + cb) * kDstPlane, r);
#endif
                }
              }
            }
            __syncwarp();
            if (lane == 0 && (ck == 2 || ck == 3)) CERB_TRACE(112 + (ck - 2) * 40 + gw * 6 + 3);
            if (gt == 0 && ck < 4) CERB_TRACE(46 + 3 * ck);
            if (lane == 0) mbar_arrive(&raw_empty[rc]);
            if (++rc == RS) { rc = 0; rcphase ^= 1; }
            publish_stage();
          }
          ++titer;
          continue;
          }   // 16-bit inputs
        }

        // ------------- fallback: gather straight from global memory (any dtype / alignment / flow) -------------
        // The loads of batch gb+1 are issued before batch gb is blended and stored, and they run
        // ahead across chunk boundaries -- only the shared-memory stores wait for a free stage.
        constexpr int NB = CC / kCBatch;
        constexpr int NV = Cfg::POS_PER_THREAD * kCBatch * 4;
        const T* x2n = x2 + (long long)x2_item(g, n) * g.x2s[0];
        const int total_batches = (ck_end - ck_begin) * NB;
        const int gb0 = ck_begin * NB;
        float cur[NV], nxt[NV];
        auto issue = [&](int gb, float (&dst)[NV]) {
          const int chb = gb * kCBatch;
#pragma unroll
          for (int cb = 0; cb < kCBatch; ++cb) {
            // unconditional loads from always-valid addresses (tap offsets of invalid positions are
            // 0, channel clamped); validity is applied when the batch is blended -- predicated loads
            // get consumed one by one (7 predicate registers) and serialise
            const int ch = min(chb + cb, g.C - 1);
            const T* plane = x2n + (long long)ch * g.x2s[1];
#pragma unroll
            for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
              float* d = &dst[(cb * Cfg::POS_PER_THREAD + j) * 4];
              d[0] = ldg_f32(plane + taps[j].off[0]);
              if (warped) {
                d[1] = ldg_f32(plane + taps[j].off[1]);
                d[2] = ldg_f32(plane + taps[j].off[2]);
                d[3] = ldg_f32(plane + taps[j].off[3]);
              }
            }
          }
        };
        if (total_batches > 0) issue(gb0, cur);
        for (int gb = 0; gb < total_batches; ++gb) {
          if (gb + 1 < total_batches) issue(gb0 + gb + 1, nxt);
          const int ck = (gb0 + gb) / NB, bi = (gb0 + gb) - ck * NB;
          float* x2dst = x2s + stage * Cfg::X2_STAGE;
          if (bi == 0) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (!a.use_tma_in) coop_x1(x1s + stage * Cfg::X1_STAGE, ix0, iy0, ck * CC, n);
          }
#pragma unroll
          for (int cb = 0; cb < kCBatch; ++cb) {
            float* plane_dst = x2dst + (bi * kCBatch + cb) * (Cfg::HY * Cfg::XS);
#pragma unroll
            for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
              const float* v = &cur[(cb * Cfg::POS_PER_THREAD + j) * 4];
              const bool ok = ((valid_mask >> j) & 1u) && (ck * CC + bi * kCBatch + cb < g.C);
              const float r = !ok ? 0.f : (warped ? blend(v[0], v[1], v[2], v[3], taps[j]) : v[0]);
              plane_dst[sdst[j]] = r;
            }
          }
          if (bi == NB - 1) publish_stage();
#pragma unroll
          for (int i = 0; i < NV; ++i) cur[i] = nxt[i];
        }
        ++titer;
      }
      pdl_launch_dependents();
      if (S > 1) { __syncwarp(); cluster_wait_acquire(); cluster_sync_all(); }
    }
  } else {
    // =========================== CONSUMER WARPS ===========================
    int grp, strip, combo;
    if constexpr (Cfg::DIRECT_OUT && Cfg::NSTRIP == 4 && TY == 8) {
      // lane = pair member | strip << 1 | pair-in-warp << 3: lanes (2k, 2k+1) are the two (y, dy) combos of a
      // pair that reads the same x2 row (merged LDS.128), a quarter-warp is one pair x 4 strips (x1 rows y even /
      // y odd -> disjoint swizzled chunks), and the 4 strips of a combo make the epilogue's stores contiguous
      const int w = tid >> 5, l = tid & 31;
      grp = 0;
      strip = (l >> 1) & 3;
      combo = w * 8 + (l >> 3) * 2 + (l & 1);
    } else {
      grp = tid / Cfg::GROUP;
      const int u = tid - grp * Cfg::GROUP;
      strip = u / Cfg::COMBOS;
      combo = u - strip * Cfg::COMBOS;
    }
    int y, dyi;
    combo_of<TY>(combo, y, dyi);
    const float divisor = (float)g.C;  // k == 1: nelems = C (correlation_cuda_kernel.cu:85)
    const float rdivisor = __frcp_rn(divisor);

    // shared-memory float offsets that do not depend on stage / channel
    const int x2_off = (y + dyi) * Cfg::XS + strip * 8;
    int x1_off[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) x1_off[h] = y * TX + swz_chunk<TX>(y, strip * 2 + h) * 4;  // row = c*TY + y; TY % 8 == 0 or TX < 32

    int tiles_done = 0;
    if (tid == 0) CERB_TRACE(16);
    for (int tile = tile0; tile < a.total_tiles; tile += tile_step) {
      const Unit un = decode_unit<TY, TX>(a, tile);
      const int n = un.n, by0 = un.by0, bx0 = un.bx0;

      float acc[8 * kD];
#pragma unroll
      for (int i = 0; i < 8 * kD; ++i) acc[i] = 0.f;

      for (int ck = ck_begin; ck < ck_end; ++ck) {
#if defined(CERB_AB_CWAIT_HINT)
        mbar_wait_hint(&full_bar[stage], phase, CERB_AB_CWAIT_HINT);
#elif defined(CERB_AB_CWAIT_SLEEP)
        mbar_wait_sleep(&full_bar[stage], phase, CERB_AB_CWAIT_SLEEP);
#else
        mbar_wait(&full_bar[stage], phase);
#endif
        if (tid == 0 && ck < 8) CERB_TRACE(17 + 2 * ck);
        const float* x1p = x1s + stage * Cfg::X1_STAGE;
        const float* x2p = x2s + stage * Cfg::X2_STAGE + x2_off;
#pragma unroll
        for (int cc = 0; cc < CC / KS; ++cc) {
          const int c = cc * KS + grp;
          float av[8], bv[16];
#if defined(CERB_X_NOCLDS)
#pragma unroll
          for (int i = 0; i < 8; ++i) av[i] = __int_as_float(0x3f800000 + i + c + ck);
#pragma unroll
          for (int i = 0; i < 16; ++i) bv[i] = __int_as_float(0x3f800000 + 3 * i + c + ck);
#else
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int off = x1_off[h];
            if constexpr (TX == 32 && (TY % 8) != 0) off = y * TX + swz_chunk<TX>(c * TY + y, strip * 2 + h) * 4;
            const float4 t = *reinterpret_cast<const float4*>(x1p + c * (TY * TX) + off);
            av[4 * h + 0] = t.x; av[4 * h + 1] = t.y; av[4 * h + 2] = t.z; av[4 * h + 3] = t.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(x2p + c * (Cfg::HY * Cfg::XS) + 4 * q);
            bv[4 * q + 0] = t.x; bv[4 * q + 1] = t.y; bv[4 * q + 2] = t.z; bv[4 * q + 3] = t.w;
          }
#endif
#if defined(CERB_X_NOFMA)
#pragma unroll
          for (int px = 0; px < 8; ++px) acc[px * kD] = fmaf(av[px], bv[px] + bv[px + 8], acc[px * kD]);
#else
#pragma unroll
          for (int px = 0; px < 8; ++px)
#pragma unroll
            for (int dx = 0; dx < kD; ++dx) acc[px * kD + dx] = fmaf(av[px], bv[px + dx], acc[px * kD + dx]);
#endif
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
        if (tid == 0 && ck < 8) CERB_TRACE(18 + 2 * ck);
        if (++stage == Cfg::ST) { stage = 0; phase ^= 1; }
      }
      if (tid == 0) CERB_TRACE(40);
      if (tile + tile_step >= a.total_tiles) pdl_launch_dependents();  // last tile: only the epilogue is left

      if constexpr (Cfg::DIRECT_OUT) {
        // ---------------- epilogue: /C, LeakyReLU, straight to global memory ----------------
        // A warp is 4 adjacent strips x 8 (y, dy) rows: every store instruction writes 8 output rows of
        // 4 x 16 bytes, and a thread's two halves complete each 32-byte sector back to back.  No staged
        // tile, no barrier: the consumers go back to the main loop as soon as the stores are issued.
        const int oy = by0 + y, ox = bx0 + strip * 8;
        if (oy < g.outH && ox < g.outW) {
          T* op = (T*)a.out + (long long)n * g.os[0] + (long long)((un.woy + dyi) * g.D + un.wox) * g.os[1] +
                  (long long)oy * g.os[2] + ox;
          const bool vec = a.out_vec8 && ox + 8 <= g.outW;
#pragma unroll
          for (int dx = 0; dx < kD; ++dx) {
            float v[8];
#pragma unroll
            for (int px = 0; px < 8; ++px) {
              float r = div_const(acc[px * kD + dx], divisor, rdivisor);
              if (g.has_act) r = leaky(r, g.slope);
              v[px] = r;
            }
#if defined(CERB_X_NOSTORE)
            if (v[0] == 12345.678f)
#endif
            if (vec) {
              store_row8<T>(op, v, a.out_vec8 > 1);
            } else {
              for (int px = 0; px < 8 && ox + px < g.outW; ++px) op[px] = from_f32<T>(v[px]);
            }
            op += g.os[1];
          }
        }
        ++tiles_done;
        ++titer;
        continue;
      }
      // ---------------- epilogue: /C, LeakyReLU, stage tile, TMA store ----------------
      const bool finalize_local = (KS == 1) || (S == 1);  // the cluster split is only used with KS > 1
      if (S == 1 && a.use_tma_out) {
        // previous tile's stores have finished reading `outs` (bulk groups belong to the thread that committed them)
        if ((tid & 31) == 0 && (KS > 1 ? tid == 0 : tid < kD * 32)) tma_store_wait_read0();
      }
      named_bar_sync(1, Cfg::NCONS);
      if (tid == 0) CERB_TRACE(35);
      float* obuf = outs + grp * Cfg::OUT_TILE;
      // KS == 1 with a TMA store and more tiles to come: the tile is staged displacement-column major
      // ([dx][dy][y][x]) and column dx is stored as soon as every thread has written it (9 boxes of 9
      // planes), so the consumers get back to the next tile's main loop sooner (batch 8: 113 -> 105 us).
      // A CTA's last tile has nothing to overlap with and goes out as one box.
      const bool piped_now = (KS == 1) && a.use_tma_out && (tile + tile_step < a.total_tiles);
#pragma unroll
      for (int dx = 0; dx < kD; ++dx) {
        const int row = (piped_now ? (dx * kD + dyi) : (dyi * kD + dx)) * TY + y;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v;
          float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float r = acc[(4 * h + e) * kD + dx];
            if constexpr (KS == 1) {
              if (finalize_local) {
                r = div_const(r, divisor, rdivisor);
                if (g.has_act) r = leaky(r, g.slope);
              }
            }
            vv[e] = r;
          }
          *reinterpret_cast<float4*>(obuf + row * TX + swz_partial<TX>(row, strip * 2 + h, S > 1) * 4) = v;
        }
        if constexpr (KS == 1) {
          if (piped_now) {
            fence_proxy_async_smem();
            named_bar_sync(1, Cfg::NCONS);
            if (tid == dx * 32) {   // lane 0 of consumer warp dx: a TMA issue stalls its thread, spread the nine over nine warps
              tma_store_5d(&tm_outc, outs + dx * (kD * TY * TX), bx0, by0, un.wox + dx, un.woy, n);
              tma_store_commit();
            }
          }
        }
      }
      if (tid == 0) CERB_TRACE(36);
      if constexpr (KS > 1) {
        named_bar_sync(1, Cfg::NCONS);
        if (tid == 0) CERB_TRACE(37);
        constexpr int kUnits = Cfg::OUT_TILE / 4;       // float4 units of one partial tile
        const float4* o4 = reinterpret_cast<const float4*>(outs);
        if (S == 1) {
          // in-CTA sum of the KS channel groups, finalised in place
          for (int u = tid; u < kUnits; u += Cfg::NCONS) {
            float4 s4 = o4[u];
#pragma unroll
            for (int k2 = 1; k2 < KS; ++k2) {
              const float4 p = o4[k2 * kUnits + u];
              s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
            }
            float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              sv[k] = div_const(sv[k], divisor, rdivisor);
              if (g.has_act) sv[k] = leaky(sv[k], g.slope);
            }
            reinterpret_cast<float4*>(outs)[u] = make_float4(sv[0], sv[1], sv[2], sv[3]);
          }
        } else {
          // ---- cluster reduction: group sums are pushed to the CTA that owns the slice ----
          float* recv = (float*)(smem + Cfg::SMEM_RECV);
          const int units = kUnits >> Slog;             // float4 units per slice (host guarantees divisibility)
          const uint32_t rbase = smem_u32(recv) + 16u * (uint32_t)(crank * units);
          __syncwarp();
          cluster_wait_acquire();                       // start-up phase: every CTA of the cluster is running
          for (int u = tid; u < kUnits; u += Cfg::NCONS) {
            float4 s4 = o4[u];
#pragma unroll
            for (int k2 = 1; k2 < KS; ++k2) {
              const float4 p = o4[k2 * kUnits + u];
              s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
            }
            const int r = u / units, idx = u - r * units;
            st_dsmem_v4(mapa_u32(rbase + 16u * (uint32_t)idx, (uint32_t)r), s4);
          }
          if (tid == 0) CERB_TRACE(38);
          __syncwarp();
          cluster_arrive_release();                     // my pushes are done ...
          if (tid == 0) CERB_TRACE(39);
          cluster_wait_acquire();                       // ... and everyone's have landed here
          if (tid == 0) CERB_TRACE(43);
          const float4* r4 = reinterpret_cast<const float4*>(recv);
          T* outp = (T*)a.out + (long long)n * g.os[0];
          for (int i = tid; i < units; i += Cfg::NCONS) {
            float4 s4 = r4[i];
            for (int r = 1; r < S; ++r) {               // fixed order: results do not depend on timing
              const float4 p = r4[r * units + i];
              s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
            }
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
            // physical chunk -> logical position of the [plane][y][x] tile (undo the partial-tile swizzle)
            const int pchunk = crank * units + i;
            const int row = pchunk / (TX / 4), lchunk = swz_partial<TX>(row, pchunk - row * (TX / 4), true);
            const int wplane = row / TY, yy = row - wplane * TY;
            const int plane = (un.woy + wplane / kD) * g.D + un.wox + wplane % kD;   // window plane -> output plane
            const int oy = by0 + yy;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float s = div_const(sv[k], divisor, rdivisor);
              if (g.has_act) s = leaky(s, g.slope);
              const int ox = bx0 + lchunk * 4 + k;
              if (oy < g.outH && ox < g.outW)
                outp[(long long)plane * g.os[1] + (long long)oy * g.os[2] + ox] = from_f32<T>(s);
            }
          }
          if (tid == 0) CERB_TRACE(41);
        }
      }
      if (S > 1) {
        // output already written by the cluster reduction above
      } else if (a.use_tma_out) {
        if (!piped_now) {   // whole window as one box (measured: nine boxes behind one barrier are slower, 20.1 vs 19.4 us)
          fence_proxy_async_smem();
          named_bar_sync(1, Cfg::NCONS);
          if (tid == 0) {
            tma_store_5d(&tm_out, outs, bx0, by0, un.wox, un.woy, n);
            tma_store_commit();
          }
        }
        if (tid == 0) CERB_TRACE(41);
      } else if (sizeof(T) == 2 && a.out_vec8) {
        // 16-bit output: 8 pixels (16 bytes) per store from the staged fp32 tile
        named_bar_sync(1, Cfg::NCONS);
        T* outp = (T*)a.out + (long long)n * g.os[0];
        for (int u = tid; u < Cfg::OUT_TILE / 8; u += Cfg::NCONS) {
          const int row = u / (TX / 8), xg = u - row * (TX / 8);
          const int wplane = row / TY, yy = row - wplane * TY;
          const int plane = (un.woy + wplane / kD) * g.D + un.wox + wplane % kD;
          const int oy = by0 + yy, ox = bx0 + xg * 8;
          if (oy < g.outH && ox < g.outW) {
            const float4 lo = *reinterpret_cast<const float4*>(outs + row * TX + swz_chunk<TX>(row, xg * 2) * 4);
            const float4 hi = *reinterpret_cast<const float4*>(outs + row * TX + swz_chunk<TX>(row, xg * 2 + 1) * 4);
            const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            T* gp = outp + (long long)plane * g.os[1] + (long long)oy * g.os[2] + ox;
            if (ox + 8 <= g.outW) {
              uint4 pk;
              T* hv = reinterpret_cast<T*>(&pk);
#pragma unroll
              for (int k = 0; k < 8; ++k) hv[k] = from_f32<T>(v[k]);
              *reinterpret_cast<uint4*>(gp) = pk;
            } else {
              for (int k = 0; k < 8 && ox + k < g.outW; ++k) gp[k] = from_f32<T>(v[k]);
            }
          }
        }
        named_bar_sync(1, Cfg::NCONS);  // `outs` is rewritten by the next tile's epilogue
      } else {
        named_bar_sync(1, Cfg::NCONS);
        T* outp = (T*)a.out + (long long)n * g.os[0];
        for (int e = tid; e < Cfg::OUT_TILE; e += Cfg::NCONS) {
          const int row = e / TX, x = e - row * TX;
          const int wplane = row / TY, yy = row - wplane * TY;
          const int plane = (un.woy + wplane / kD) * g.D + un.wox + wplane % kD;
          const int oy = by0 + yy, ox = bx0 + x;
          if (oy < g.outH && ox < g.outW) {
            const float v = outs[row * TX + swz_chunk<TX>(row, x >> 2) * 4 + (x & 3)];
            outp[(long long)plane * g.os[1] + (long long)oy * g.os[2] + ox] = from_f32<T>(v);
          }
        }
        named_bar_sync(1, Cfg::NCONS);  // `outs` is rewritten by the next tile's epilogue
      }
      ++tiles_done;
      ++titer;
    }
    if (S == 1 && a.use_tma_out && (tid & 31) == 0 && (KS > 1 ? tid == 0 : tid < kD * 32)) tma_store_wait_read0();
    if (tid == 0 && a.dbg) a.dbg[(long long)blockIdx.x * kTraceSlots + 42] = clock64();
  }
}

// ------------------------------------------------------------------ generic kernel -------
// Any pad / kernel_size / max_displacement / stride1 / stride2.  One thread per output element;
// taps beyond the padded buffer read zero (the reference reads out of bounds there for
// kernel_size > 1).
template <typename T>
__global__ void __launch_bounds__(256) corr_fwd_generic_kernel(const Geom g, const T* __restrict__ x1,
                                                               const T* __restrict__ x2, const float* __restrict__ flow,
                                                               T* __restrict__ out) {
  const long long total = (long long)g.B * g.D2 * g.outH * g.outW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int bx = (int)(idx % g.outW);
    long long t = idx / g.outW;
    const int by = (int)(t % g.outH);
    t /= g.outH;
    const int tc = (int)(t % g.D2);
    const int n = (int)(t / g.D2);
    const int tj = tc / g.D - g.r, ti = tc % g.D - g.r;
    const T* x1n = x1 + (long long)n * g.x1s[0];
    const T* x2n = x2 + (long long)x2_item(g, n) * g.x2s[0];
    float acc = 0.f;
    for (int j = -g.kr; j <= g.kr; ++j) {
      for (int i = -g.kr; i <= g.kr; ++i) {
        const int ya = by * g.s1 + g.md + j - g.pad, xa = bx * g.s1 + g.md + i - g.pad;
        const int yb = ya + tj * g.s2, xb = xa + ti * g.s2;
        if (ya < 0 || ya >= g.H || xa < 0 || xa >= g.W) continue;
        if (yb < 0 || yb >= g.H || xb < 0 || xb >= g.W) continue;
        const T* pa = x1n + (long long)ya * g.x1s[2] + xa;
        if (flow != nullptr) {
          const float* fp = flow + (long long)n * g.fls[0] + (long long)yb * g.fls[2] + xb;
          bool in_x, in_y;
          const float sx = sample_pos(xb, __ldg(fp), g.W, g.warp_mode, in_x);
          const float sy = sample_pos(yb, __ldg(fp + g.fls[1]), g.H, g.warp_mode, in_y);
          const Taps tp = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
          for (int c = 0; c < g.C; ++c) {
            const T* pb = x2n + (long long)c * g.x2s[1];
            const float b = blend(ldg_f32(pb + tp.off[0]), ldg_f32(pb + tp.off[1]), ldg_f32(pb + tp.off[2]),
                                  ldg_f32(pb + tp.off[3]), tp);
            acc = fmaf(ldg_f32(pa + (long long)c * g.x1s[1]), b, acc);
          }
        } else {
          const T* pb = x2n + (long long)yb * g.x2s[2] + xb;
          for (int c = 0; c < g.C; ++c)
            acc = fmaf(ldg_f32(pa + (long long)c * g.x1s[1]), ldg_f32(pb + (long long)c * g.x2s[1]), acc);
        }
      }
    }
    float v = __fdiv_rn(acc, (float)(g.k * g.k * g.C));
    if (g.has_act) v = leaky(v, g.slope);
    out[(long long)n * g.os[0] + (long long)tc * g.os[1] + (long long)by * g.os[2] + bx] = from_f32<T>(v);
  }
}

// ------------------------------------------------------------------ host launchers -------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return (PFN_encodeTiled)p;
  }();
  return fn;
}

// fp32 NCHW (W-stride 1) tensor map with box {bx, by, bc, 1}
// 4-D (x, y, c, n) tiled tensor map over an NCHW tensor of `esize`-byte elements
bool make_tmap_nchw(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                    const long long strides[3], int bx, int by, int bc, bool swizzle128);
static bool make_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                      const long long strides[3], int bx, int by, int bc, bool swizzle128) {
  return make_tmap_nchw(tm, dt, esize, base, W, H, C, B, strides, bx, by, bc, swizzle128);
}
// Encoded tensor maps are cached: a decoder calls the op with the same buffers every step, and one encode costs about a
// microsecond of host time (up to six per launch).  Key = everything the encode depends on; a small direct-mapped table
// under a mutex (the C ABI may be called from several host threads); a map is a pure function of its key, so a stale
// entry for a freed-and-reallocated pointer with the same geometry is still the right map.
struct TmapKey {
  unsigned long long base;
  long long s0, s1, s2;
  int dt, esize, W, H, C, B, bx, by, bc, kind;   // kind: 0/1 = NCHW map without / with swizzle, 2 = 5-D output map (bc = D)
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool used; };
static TmapSlot g_tmap_cache[256];
static std::mutex g_tmap_mutex;
static unsigned tmap_hash(const TmapKey& k) {
  unsigned long long h = k.base * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)(k.bx * 131 + k.by * 31 + k.bc * 7 + k.kind * 3 + k.dt) * 0xC2B2AE3D27D4EB4Full;
  h ^= (unsigned long long)k.s0 * 0x165667B19E3779F9ull;
  return (unsigned)(h >> 40) & 255u;
}
static bool tmap_lookup(const TmapKey& k, CUtensorMap* tm) {
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  const TmapSlot& s = g_tmap_cache[tmap_hash(k)];
  if (!s.used || memcmp(&s.key, &k, sizeof(TmapKey)) != 0) return false;
  *tm = s.map;
  return true;
}
static void tmap_store(const TmapKey& k, const CUtensorMap& tm) {
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  TmapSlot& s = g_tmap_cache[tmap_hash(k)];
  s.key = k; s.map = tm; s.used = true;
}

// (also used by costvolume_fwd_tc.cu)
bool make_tmap_nchw(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, int W, int H, int C, int B,
                    const long long strides[3], int bx, int by, int bc, bool swizzle128) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  if (((uintptr_t)base & 15) != 0) return false;
  for (int i = 0; i < 3; ++i)
    if (strides[i] <= 0 || (strides[i] * esize) % 16 != 0) return false;
  if (bx > 256 || by > 256 || bc > 256 || (bx * esize) % 16 != 0) return false;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = (unsigned long long)(uintptr_t)base; key.s0 = strides[0]; key.s1 = strides[1]; key.s2 = strides[2];
  key.dt = (int)dt; key.esize = esize; key.W = W; key.H = H; key.C = C; key.B = B; key.bx = bx; key.by = by; key.bc = bc;
  key.kind = swizzle128 ? 1 : 0;
  if (tmap_lookup(key, tm)) return true;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)strides[2] * esize, (cuuint64_t)strides[1] * esize, (cuuint64_t)strides[0] * esize};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bc, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, dt, 4, const_cast<void*>(base), dims, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_store(key, *tm);
  return r == CUDA_SUCCESS;
}
static bool make_tmap_f32(CUtensorMap* tm, const void* base, int W, int H, int C, int B, const long long strides[3],
                          int bx, int by, int bc, bool swizzle128) {
  return make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, W, H, C, B, strides, bx, by, bc, swizzle128);
}

// output as (x, y, dx, dy, n): a 9 x 9 displacement window of a tile is one box
static bool make_tmap_out5d(CUtensorMap* tm, const void* base, const Geom& g, int bx, int by, int bdx) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  if (((uintptr_t)base & 15) != 0) return false;
  for (int i = 0; i < 3; ++i)
    if (g.os[i] <= 0 || (g.os[i] * 4) % 16 != 0) return false;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = (unsigned long long)(uintptr_t)base; key.s0 = g.os[0]; key.s1 = g.os[1]; key.s2 = g.os[2];
  key.dt = (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT32; key.esize = 4; key.W = g.outW; key.H = g.outH; key.C = g.D; key.B = g.B;
  key.bx = bx; key.by = by; key.bc = bdx; key.kind = 2;
  if (tmap_lookup(key, tm)) return true;
  cuuint64_t dims[5] = {(cuuint64_t)g.outW, (cuuint64_t)g.outH, (cuuint64_t)g.D, (cuuint64_t)g.D, (cuuint64_t)g.B};
  cuuint64_t gstr[4] = {(cuuint64_t)g.os[2] * 4, (cuuint64_t)g.os[1] * 4, (cuuint64_t)g.os[1] * 4 * g.D, (cuuint64_t)g.os[0] * 4};
  cuuint32_t box[5] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bdx, (cuuint32_t)kD, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bx == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_store(key, *tm);
  return r == CUDA_SUCCESS;
}

// SM count of the current device (cached per device: one process may drive several GPUs)
static int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  if (v <= 0) v = 148;
  if (dev >= 0 && dev < 64) cache[dev] = v;
  return v;
}

int num_sms_current() { return num_sms(); }   // costvolume_fwd_tc.cu

template <typename T, int TY, int TX, int KS, int CC, int RS, bool UP>
static const void* kern_for_query() { return (const void*)warp_corr_fwd_kernel<T, TY, TX, KS, CC, RS, UP>; }

// co-resident clusters of `csize` CTAs for this kernel on the current device (cached per kernel / size / device)
static int max_active_clusters(const void* kern, int threads, size_t smem, int csize) {
  struct Key { const void* k; int c, dev, val; };
  static Key cache[64];
  static int ncache = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < ncache; ++i)
    if (cache[i].k == kern && cache[i].c == csize && cache[i].dev == dev) return cache[i].val;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * 148);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = (unsigned)csize; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = num_sms() / csize; }
  if (ncache < 64) cache[ncache++] = Key{kern, csize, dev, n};   // benign race: worst case the query repeats
  return n;
}

template <typename T, int TY, int TX, int KS, int CC, int RS, bool UP>
static cudaError_t launch_fast(const Geom& g, const void* x1, const void* x2, const float* flow, void* out,
                               int force_no_tma, bool allow_csplit, cudaStream_t stream, const UpFlow* uf) {
  using Cfg = FwdCfg<TY, TX, KS, CC, RS>;
  FwdArgs a;
  a.g = g;
  a.x1 = x1; a.x2 = x2; a.flow = flow; a.out = out;
  a.cflow = nullptr; a.flow_up = nullptr; a.Hc = a.Wc = 0; a.up_sy = a.up_sx = 0.f;
  for (int i = 0; i < 3; ++i) a.cfs[i] = a.fus[i] = 0;
  if (uf != nullptr) {
    a.cflow = uf->coarse; a.flow_up = uf->up; a.Hc = uf->Hc; a.Wc = uf->Wc;
    for (int i = 0; i < 3; ++i) { a.cfs[i] = uf->cs[i]; a.fus[i] = uf->us[i]; }
    // ATen area_pixel_compute_scale<float>(in, out, align_corners=true): (float)(in - 1) / (out - 1)
    a.up_sy = g.H > 1 ? (float)(uf->Hc - 1) / (float)(g.H - 1) : 0.f;
    a.up_sx = g.W > 1 ? (float)(uf->Wc - 1) / (float)(g.W - 1) : 0.f;
  }
  const bool warped_in = flow != nullptr || uf != nullptr;
  a.off = g.md - g.pad;
  a.tiles_x = (g.outW + TX - 1) / TX;
  a.tiles_y = (g.outH + TY - 1) / TY;
  a.nwin = (g.D - kD + 7) / 8 + 1;   // 9 x 9 displacement windows per axis: origins 0, 8, ..., D - 9
  a.total_tiles = g.B * a.tiles_x * a.tiles_y * a.nwin * a.nwin;
  a.nchunks = (g.C + CC - 1) / CC;
  a.dbg = g_trace_buffer;
  a.path_ctr = g_path_counters;
  a.dbg_iter = g_trace_iter;
  CUtensorMap tm_x1, tm_x2, tm_raw, tm_rawL, tm_out, tm_outc;
  memset(&tm_x1, 0, sizeof(tm_x1));
  memset(&tm_x2, 0, sizeof(tm_x2));
  memset(&tm_raw, 0, sizeof(tm_raw));
  memset(&tm_rawL, 0, sizeof(tm_rawL));
  memset(&tm_out, 0, sizeof(tm_out));
  memset(&tm_outc, 0, sizeof(tm_outc));
  a.use_tma_in = a.use_tma_x2 = a.use_tma_raw = a.use_tma_rawL = a.use_tma_out = 0;
  int tma_mask = 15;  // debugging knob CERB_DEBUG_TMA: bit0 x1 loads, bit1 stores, bit2 plain x2 tiles, bit3 raw x2 boxes
  if (const char* e = getenv("CERB_DEBUG_TMA")) tma_mask = atoi(e);
  if (std::is_same<T, float>::value && !force_no_tma) {
    // the box start of a TMA tile load must be 16-byte aligned in global memory (misaligned
    // starts trap): x1 / x2 tiles begin at bx0 + (md - pad) (- 4), so pad != md (mod 4) stages
    // them with LDG instead.  Raw boxes are aligned by construction.
    if ((tma_mask & 1) && (a.off % 4) == 0)
      a.use_tma_in = make_tmap_f32(&tm_x1, x1, g.W, g.H, g.C, g.B, g.x1s, TX, TY, CC, TX == 32) ? 1 : 0;
    // (x2 halo tiles start at bx0 - pad + window origin; origins are multiples of 8 and D - 9 = 2 md - 8)
    if ((tma_mask & 4) && (g.pad % 4) == 0 && (a.nwin == 1 || (g.md % 2) == 0) && !warped_in)
      a.use_tma_x2 = make_tmap_f32(&tm_x2, x2, g.W, g.H, g.C, g.B, g.x2s, Cfg::XS, Cfg::HY, CC, false) ? 1 : 0;
    if ((tma_mask & 8) && warped_in)
      a.use_tma_raw = make_tmap_f32(&tm_raw, x2, g.W, g.H, g.C, g.B, g.x2s, Cfg::RAW_W, Cfg::RAW_H, CC, false) ? 1 : 0;
    if (Cfg::BIGBOX && a.use_tma_raw && !getenv("CERB_DEBUG_NO_BIGBOX"))
      a.use_tma_rawL = make_tmap_f32(&tm_rawL, x2, g.W, g.H, g.C, g.B, g.x2s, Cfg::RAWL_W, Cfg::RAWL_H, CC, false) ? 1 : 0;
    if (!Cfg::DIRECT_OUT && (tma_mask & 2))
      a.use_tma_out = (make_tmap_out5d(&tm_out, out, g, TX, TY, kD) &&
                       (KS > 1 || make_tmap_out5d(&tm_outc, out, g, TX, TY, 1))) ? 1 : 0;   // KS == 1 also stores per displacement column
  }
  a.raw16 = a.out_vec8 = 0;
  if (sizeof(T) == 2 && !force_no_tma && (tma_mask & 8)) {
    // 16-bit inputs: raw x2 box and x1 tile by TMA as 16-bit data (box starts 16-byte aligned: multiples of 8 elements)
    const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if ((a.off % 8) == 0 &&
        make_tmap(&tm_raw, dt, 2, x2, g.W, g.H, g.C, g.B, g.x2s, Cfg::RAW_W16, Cfg::RAW_H, CC, false) &&
        make_tmap(&tm_x1, dt, 2, x1, g.W, g.H, g.C, g.B, g.x1s, TX, TY, CC, false)) {
      a.raw16 = 1;
      a.use_tma_raw = 1;
    }
  }
  {
    const int per16 = 16 / (int)sizeof(T);
    if (sizeof(T) == 2 || Cfg::DIRECT_OUT)
      a.out_vec8 = (((uintptr_t)out & 15) == 0 && g.os[0] % per16 == 0 && g.os[1] % per16 == 0 && g.os[2] % per16 == 0) ? 1 : 0;
    if (a.out_vec8 && sizeof(T) == 4 && ((uintptr_t)out & 31) == 0 && g.os[0] % 8 == 0 && g.os[1] % 8 == 0 && g.os[2] % 8 == 0)
      a.out_vec8 = 2;   // 32-byte aligned rows: 256-bit stores
  }
  auto kern = warp_corr_fwd_kernel<T, TY, TX, KS, CC, RS, UP>;
  // function attributes are per device: remember which devices have the > 48 KB opt-in (benign race: idempotent call)
  static unsigned long long attr_devs = 0ull;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit == 0ull || !(attr_devs & bit)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
      if (e != cudaSuccess) return e;
      attr_devs |= bit;
    }
  }
  // coarse pyramid levels have fewer tiles than SMs: split the channel chunks of each tile over a
  // thread-block cluster (partial tiles reduced through DSMEM in the epilogue)
  int S = 1;
  if (allow_csplit && TX < 32 && !getenv("CERB_DEBUG_NO_CSPLIT")) {
    while (S < 8 && a.total_tiles * (S * 2) <= num_sms() && a.nchunks >= S * 2 && (Cfg::OUT_TILE % (S * 2 * 4)) == 0) S *= 2;
    // every cluster must be resident at once (one tile per cluster, no second wave): clusters are placed inside
    // one GPC, and not every GPC holds two clusters of 8 -- 16 tiles x 8 CTAs ran as two waves (17 vs 8.5 us)
    while (S > 1 && max_active_clusters(kern_for_query<T, TY, TX, KS, CC, RS, UP>(), Cfg::NTHREADS, Cfg::SMEM_BYTES, S) < a.total_tiles) S /= 2;
  }
  a.csplit = S;
  a.csplit_log2 = S == 8 ? 3 : S == 4 ? 2 : S == 2 ? 1 : 0;
  if (S > 1) a.use_tma_out = 0;
  int nclusters = num_sms() / S;
  if (nclusters > a.total_tiles) nclusters = a.total_tiles;
  const int grid = nclusters * S;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NTHREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool use_pdl = getenv("CERB_DEBUG_NO_PDL") == nullptr;
  if (use_pdl) {  // programmatic dependent launch: our setup overlaps the previous kernel's tail
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (S > 1) {    // (no cluster attribute for plain launches: a 1-CTA "cluster" still switches the CTA
                  // scheduler and costs ~2 us on a one-wave grid)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)S;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a, tm_x1, tm_x2, tm_raw, tm_rawL, tm_out, tm_outc);
  if (le != cudaSuccess) return le;
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_fwd_t(const Geom& g, const void* x1, const void* x2, const float* flow, void* out,
                                int variant, cudaStream_t stream, const UpFlow* uf) {
  const bool fast_ok = g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md >= kMD && variant != CERB_FWD_VARIANT_GENERIC;
  // tensor-core variant (costvolume_fwd_tc.cu): on request, or when there are at least 64 tiles of 8 x 16 (measured on
  // B200, both flow directions per launch: 64 tiles 15.1 vs 15.6 us, 128 tiles 15.4 vs 23.6, 512 tiles 29.9 vs 32.6;
  // 32 tiles 18.9 vs 14.1 -- the cluster-split CUDA-core kernel keeps the coarse levels)
  {
    const int dt = std::is_same<T, float>::value ? CERB_F32 : (std::is_same<T, __half>::value ? CERB_F16 : CERB_BF16);
    if (variant == CERB_FWD_VARIANT_TC) {
      if (!tc_forward_supported(g, dt, uf)) return cudaErrorNotSupported;
      return launch_warp_corr_forward_tc(g, dt, x1, x2, flow, out, stream, uf);
    }
    static const int tc_env = getenv("CERB_FWD_TC") ? atoi(getenv("CERB_FWD_TC")) : -1;   // 0: never, 1: whenever supported
    if (variant == CERB_FWD_VARIANT_AUTO && tc_env != 0 && tc_forward_supported(g, dt, uf) && (flow != nullptr || uf != nullptr)) {
      const long long tiles = (long long)g.B * ((g.outW + 15) / 16) * ((g.outH + 7) / 8);
      if (tc_env == 1 || tiles >= 64) return launch_warp_corr_forward_tc(g, dt, x1, x2, flow, out, stream, uf);
    }
  }
  if (fast_ok) {
    const int no_tma = (variant == CERB_FWD_VARIANT_FAST_NOTMA || variant == CERB_FWD_VARIANT_SMALL_NOTMA) ? 1 : 0;
    bool small = variant == CERB_FWD_VARIANT_SMALL || variant == CERB_FWD_VARIANT_SMALL_NOTMA;
    bool mid = variant == CERB_FWD_VARIANT_MID;
    if (variant == CERB_FWD_VARIANT_AUTO) {
      // Tile shape by how many work units it gives the 148 SMs: 8x32 (least halo per pixel) when that fills 3/4 of
      // the GPU; else 4x16 tiles with the channels split four ways in the CTA (and over a cluster) while they
      // still fit one wave; else 8x16 if that fits one wave (4x16 would need two); else 8x32 after all.
      const long long nwin = (g.D - kD + 7) / 8 + 1;
      const long long per_b = (long long)g.B * nwin * nwin;
      const long long big_tiles = per_b * ((g.outW + 31) / 32) * ((g.outH + 7) / 8);
      const long long mid_tiles = per_b * ((g.outW + 15) / 16) * ((g.outH + 7) / 8);
      const long long small_tiles = per_b * ((g.outW + 15) / 16) * ((g.outH + 3) / 4);
      const long long sms = num_sms();
      if (big_tiles >= sms * 3 / 4) { /* 8x32 */ }
      else if (small_tiles <= sms) small = true;
      else if (mid_tiles <= sms) mid = true;
    }
    if (uf != nullptr) {
      if (small) return launch_fast<T, 4, 16, 4, 8, 2, true>(g, x1, x2, flow, out, no_tma, true, stream, uf);
      if (mid) return launch_fast<T, 8, 16, 2, 4, 2, true>(g, x1, x2, flow, out, no_tma, true, stream, uf);
      return launch_fast<T, 8, 32, 1, CERB_AB_CC1, CERB_AB_RS1, true>(g, x1, x2, flow, out, no_tma, false, stream, uf);
    }
    if (small) return launch_fast<T, 4, 16, 4, 8, 2, false>(g, x1, x2, flow, out, no_tma, true, stream, uf);
    if (mid) return launch_fast<T, 8, 16, 2, 4, 2, false>(g, x1, x2, flow, out, no_tma, true, stream, uf);
    return launch_fast<T, 8, 32, 1, CERB_AB_CC1, CERB_AB_RS1, false>(g, x1, x2, flow, out, no_tma, false, stream, uf);
  }
  if (uf != nullptr) return cudaErrorNotSupported;   // the generic kernel has no fused up-sampling
  const long long total = (long long)g.B * g.D2 * g.outH * g.outW;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  corr_fwd_generic_kernel<T><<<(int)blocks, 256, 0, stream>>>(g, (const T*)x1, (const T*)x2, flow, (T*)out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ FMA roof --------------
// 8 independent packed-FMA chains per thread (fma.rn.f32x2: two FMAs per lane per instruction), 1024 threads per SM.
__global__ void __launch_bounds__(1024) fma_peak_kernel(float* out, int iters) {
  unsigned long long a[8];
  const unsigned long long m = 0x3f8000013f800001ull, c = 0x3a83126f3a83126full;   // {1.0000001, 1.0000001}, {0.001, 0.001}
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0x3f8000003f800000ull + (unsigned long long)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  if (s == 0x1234567812345678ull) out[0] = 1.f;
}

cudaError_t measure_fma_peak(double* tflops, cudaStream_t stream) {
  float* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(float));
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = num_sms() * 2, iters = 20000;
  fma_peak_kernel<<<blocks, 1024, 0, stream>>>(d, 2000);   // warm-up (clocks)
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, stream);
    fma_peak_kernel<<<blocks, 1024, 0, stream>>>(d, iters);
    cudaEventRecord(e1, stream);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)blocks * 1024.0 * (double)iters * 32.0 * 2.0 * 2.0;   // 32 f32x2 FMAs per iteration
    if (ms > 0.f) best = best > flops / (ms * 1e-3) / 1e12 ? best : flops / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e == cudaSuccess) *tflops = best;
  return e;
}

cudaError_t launch_warp_corr_forward(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow,
                                     void* out, int variant, cudaStream_t stream, const UpFlow* uf) {
  switch (dtype) {
    case CERB_F32: return launch_fwd_t<float>(g, x1, x2, flow, out, variant, stream, uf);
    case CERB_F16: return launch_fwd_t<__half>(g, x1, x2, flow, out, variant, stream, uf);
    case CERB_BF16: return launch_fwd_t<__nv_bfloat16>(g, x1, x2, flow, out, variant, stream, uf);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cerb
