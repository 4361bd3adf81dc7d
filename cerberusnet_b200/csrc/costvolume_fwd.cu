// costvolume_fwd.cu -- fused flow-warp + correlation + LeakyReLU forward for sm_100a.
//
// Replaces (reference paths relative to the reference checkout):
//   flow_warp                     nnet_training/loss_functions/UnFlowLoss.py:83-94
//   correlation_forward_cuda      nnet_training/correlation_package/correlation_cuda.cpp:3-26
//     channels_first x2, correlation_forward   correlation_cuda_kernel.cu:13-27, 29-95, 244-324
//   F.leaky_relu_(out, 0.1)       nnet_training/nnet_models/pwcnet_sfd.py:182
//
// Fast path (kernel_size=1, stride1=stride2=1, max_displacement=4: every model in the reference):
//   persistent, warp-specialised CTAs, one output tile of TY x TX pixels at a time.
//   * 3 producer warps: x1 tile per channel chunk by TMA (cp.async.bulk.tensor, zero fill = the
//     correlation padding) and the (TY+8) x (TX+8) halo tile of the *warped* x2 gathered
//     bilinearly straight from global/L2 into shared memory -- the warped map never exists in HBM;
//   * 9 consumer warps: thread = (8-pixel strip, row y, row displacement dy) holding 8 x 9
//     accumulators; per channel 2 + 4 LDS.128 feed 72 FFMA.  Lane pairs are arranged to read the
//     same x2 row (B200 merges adjacent-lane duplicate LDS.128 addresses: 2.4 instead of 4
//     clk/instr, tools/microbench/pipes.cu);
//   * epilogue: divide by C, LeakyReLU, stage the 81 x TY x TX tile in shared memory (128B
//     swizzle) and write it with one TMA store.
// Generic path (any pad/kernel/stride parameters): one thread per output element.
#include <cuda.h>

#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

// ------------------------------------------------------------------ configuration --------
constexpr int kMD = 4;
constexpr int kD = 2 * kMD + 1;  // 9
constexpr int kD2 = kD * kD;     // 81
constexpr int kProducerThreads = 96;
constexpr int kCC = 8;      // channels per pipeline stage
constexpr int kStages = 3;

template <int TY, int TX, int KS>
struct FwdCfg {
  static constexpr int NSTRIP = TX / 8;
  static constexpr int COMBOS = TY * kD;
  static constexpr int GROUP = NSTRIP * COMBOS;
  static constexpr int NCONS = GROUP * KS;
  static constexpr int NTHREADS = NCONS + kProducerThreads;
  static constexpr int HY = TY + 2 * kMD, HX = TX + 2 * kMD;
  static constexpr int XS = HX + 4;  // 44 / 28: 8 consecutive rows hit 8 distinct 16-byte bank groups
  static constexpr int NPOS = HY * HX;
  static constexpr int POS_PER_THREAD = (NPOS + kProducerThreads - 1) / kProducerThreads;
  static constexpr int X1_STAGE = kCC * TY * TX;  // floats
  static constexpr int X2_STAGE = kCC * HY * XS;  // floats
  static constexpr int OUT_TILE = kD2 * TY * TX;  // floats, one partial buffer
  static constexpr size_t SMEM_X1 = 0;
  static constexpr size_t SMEM_X2 = SMEM_X1 + sizeof(float) * kStages * X1_STAGE;
  static constexpr size_t SMEM_OUT = (SMEM_X2 + sizeof(float) * kStages * X2_STAGE + 1023) / 1024 * 1024;
  static constexpr size_t SMEM_BAR = SMEM_OUT + sizeof(float) * KS * OUT_TILE;
  static constexpr size_t SMEM_BYTES = SMEM_BAR + 2 * kStages * sizeof(uint64_t) + 1024;  // + alignment slack
  static_assert(NCONS % 32 == 0, "consumer threads must be whole warps");
  static_assert((sizeof(float) * X1_STAGE) % 1024 == 0, "x1 stage must keep 1024-byte alignment");
  static_assert(kCC % KS == 0, "channel groups must divide the stage");
};

// (y, dy) enumeration per strip: aligned lane pairs share the x2 row r = y + dy wherever possible.
template <int TY>
struct ComboLut {
  unsigned char y[TY * kD];
  unsigned char d[TY * kD];
};
template <int TY>
constexpr ComboLut<TY> make_combo_lut() {
  ComboLut<TY> l{};
  unsigned char ly[TY * kD] = {}, ld[TY * kD] = {};
  int n = 0, nl = 0;
  for (int r = 0; r < TY + 2 * kMD; ++r) {
    const int y0 = r - 2 * kMD > 0 ? r - 2 * kMD : 0;
    const int y1 = r < TY - 1 ? r : TY - 1;
    const int cnt = y1 - y0 + 1;
    int i = 0;
    for (; i + 1 < cnt; i += 2) {
      l.y[n] = (unsigned char)(y0 + i); l.d[n] = (unsigned char)(r - (y0 + i)); ++n;
      l.y[n] = (unsigned char)(y0 + i + 1); l.d[n] = (unsigned char)(r - (y0 + i + 1)); ++n;
    }
    if (i < cnt) { ly[nl] = (unsigned char)(y0 + i); ld[nl] = (unsigned char)(r - (y0 + i)); ++nl; }
  }
  for (int i = 0; i < nl; ++i) { l.y[n] = ly[i]; l.d[n] = ld[i]; ++n; }
  return l;
}
__constant__ ComboLut<8> c_lut8 = make_combo_lut<8>();
__constant__ ComboLut<4> c_lut4 = make_combo_lut<4>();

template <int TY> __device__ __forceinline__ void combo_of(int idx, int& y, int& d);
template <> __device__ __forceinline__ void combo_of<8>(int idx, int& y, int& d) { y = c_lut8.y[idx]; d = c_lut8.d[idx]; }
template <> __device__ __forceinline__ void combo_of<4>(int idx, int& y, int& d) { y = c_lut4.y[idx]; d = c_lut4.d[idx]; }

struct FwdArgs {
  Geom g;
  const void* x1;
  const void* x2;
  const float* flow;
  void* out;
  int off;        // md - pad: output pixel (by,bx) looks at input pixel (by+off, bx+off)
  int tiles_x, tiles_y, total_tiles;
  int nchunks;    // ceil(C / kCC)
  int use_tma_in, use_tma_out;
};

// 16-byte chunk swizzle of a [rows][TX] fp32 tile: CU_TENSOR_MAP_SWIZZLE_128B for TX == 32
// (chunk ^= row & 7 inside 1024-byte atoms); identity for narrower tiles.
template <int TX>
__device__ __forceinline__ int swz_chunk(int row, int chunk) {
  if constexpr (TX == 32) return chunk ^ (row & 7);
  else return chunk;
}

// ------------------------------------------------------------------ fast kernel ----------
template <typename T, int TY, int TX, int KS>
__global__ void __launch_bounds__(FwdCfg<TY, TX, KS>::NTHREADS, 1)
warp_corr_fwd_kernel(const FwdArgs a, const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_out) {
  using Cfg = FwdCfg<TY, TX, KS>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // keep the pointer derived from smem_raw (so loads/stores stay LDS/STS) while forcing the
  // 1024-byte alignment the 128B TMA swizzle atoms need
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* x1s = (float*)(smem + Cfg::SMEM_X1);
  float* x2s = (float*)(smem + Cfg::SMEM_X2);
  float* outs = (float*)(smem + Cfg::SMEM_OUT);
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::SMEM_BAR);
  uint64_t* empty_bar = full_bar + kStages;

  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const T* __restrict__ x1 = (const T*)a.x1;
  const T* __restrict__ x2 = (const T*)a.x2;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1 + kProducerThreads / 32);
      mbar_init(&empty_bar[s], Cfg::NCONS / 32);
    }
    fence_barrier_init();
    if (a.use_tma_in) tma_prefetch_desc(&tm_x1);
    if (a.use_tma_out) tma_prefetch_desc(&tm_out);
  }
  __syncthreads();

  int stage = 0;
  uint32_t phase = 0;

  if (tid >= Cfg::NCONS) {
    // =========================== PRODUCER WARPS ===========================
    const int pt = tid - Cfg::NCONS;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int n = tile / (a.tiles_x * a.tiles_y);
      const int trem = tile - n * (a.tiles_x * a.tiles_y);
      const int by0 = (trem / a.tiles_x) * TY, bx0 = (trem % a.tiles_x) * TX;
      // input-frame origin of the x1 tile and of the x2 halo tile
      const int iy0 = by0 + a.off, ix0 = bx0 + a.off;
      const int qy0 = iy0 - kMD, qx0 = ix0 - kMD;

      // per-position sampling data, fixed for the whole tile (all channels reuse it)
      Taps taps[Cfg::POS_PER_THREAD];
      int sdst[Cfg::POS_PER_THREAD];  // smem float offset inside a channel plane, -1 = no position
      unsigned valid_mask = 0;
#pragma unroll
      for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
        const int i = pt + j * kProducerThreads;
        sdst[j] = -1;
        if (i < Cfg::NPOS) {
          const int hy = i / Cfg::HX, hx = i - hy * Cfg::HX;
          sdst[j] = hy * Cfg::XS + hx;
          const int qy = qy0 + hy, qx = qx0 + hx;
          if (qy >= 0 && qy < g.H && qx >= 0 && qx < g.W) {
            valid_mask |= 1u << j;
            if (a.flow != nullptr) {
              const float* fp = a.flow + (long long)n * g.fls[0] + (long long)qy * g.fls[2] + qx;
              const float u = __ldg(fp), v = __ldg(fp + g.fls[1]);
              bool in_x, in_y;
              const float sx = sample_pos(qx, u, g.W, g.warp_mode, in_x);
              const float sy = sample_pos(qy, v, g.H, g.warp_mode, in_y);
              taps[j] = make_taps(sx, sy, g.H, g.W, g.x2s[2]);
            } else {
              const int o = (int)(qy * g.x2s[2]) + qx;
              taps[j].off[0] = taps[j].off[1] = taps[j].off[2] = taps[j].off[3] = o;
              taps[j].w[0] = 1.f; taps[j].w[1] = taps[j].w[2] = taps[j].w[3] = 0.f;
            }
          }
        }
      }

      for (int ck = 0; ck < a.nchunks; ++ck) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        float* x1dst = x1s + stage * Cfg::X1_STAGE;
        float* x2dst = x2s + stage * Cfg::X2_STAGE;
        const int c0 = ck * kCC;
        if (pt == 0) {
          if (a.use_tma_in) {
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(sizeof(float) * Cfg::X1_STAGE));
            tma_load_4d(x1dst, &tm_x1, &full_bar[stage], ix0, iy0, c0, n);
          } else {
            mbar_arrive(&full_bar[stage]);
          }
        }
        if (!a.use_tma_in) {
          // cooperative x1 tile load (any dtype / alignment), same swizzled layout TMA produces
          for (int e = pt; e < Cfg::X1_STAGE; e += kProducerThreads) {
            const int c = e / (TY * TX), rem = e - c * (TY * TX);
            const int y = rem / TX, x = rem - y * TX;
            const int iy = iy0 + y, ix = ix0 + x, ch = c0 + c;
            float v = 0.f;
            if (ch < g.C && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
              v = ldg_f32(x1 + (long long)n * g.x1s[0] + (long long)ch * g.x1s[1] + (long long)iy * g.x1s[2] + ix);
            const int row = c * TY + y;
            x1dst[row * TX + swz_chunk<TX>(row, x >> 2) * 4 + (x & 3)] = v;
          }
        }
        // warped (or plain) x2 halo tile for this channel chunk
        const T* x2n = x2 + (long long)n * g.x2s[0];
#pragma unroll 2
        for (int c = 0; c < kCC; ++c) {
          const int ch = c0 + c;
          const T* plane = x2n + (long long)ch * g.x2s[1];
          const bool ch_ok = ch < g.C;
          float vals[Cfg::POS_PER_THREAD];
#pragma unroll
          for (int j = 0; j < Cfg::POS_PER_THREAD; ++j) {
            float v = 0.f;
            if (ch_ok && ((valid_mask >> j) & 1u)) {
              if (a.flow != nullptr) {
                const float v0 = ldg_f32(plane + taps[j].off[0]);
                const float v1 = ldg_f32(plane + taps[j].off[1]);
                const float v2 = ldg_f32(plane + taps[j].off[2]);
                const float v3 = ldg_f32(plane + taps[j].off[3]);
                v = blend(v0, v1, v2, v3, taps[j]);
              } else {
                v = ldg_f32(plane + taps[j].off[0]);
              }
            }
            vals[j] = v;
          }
#pragma unroll
          for (int j = 0; j < Cfg::POS_PER_THREAD; ++j)
            if (sdst[j] >= 0) x2dst[c * (Cfg::HY * Cfg::XS) + sdst[j]] = vals[j];
        }
        __syncwarp();
        if ((pt & 31) == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // =========================== CONSUMER WARPS ===========================
    const int grp = tid / Cfg::GROUP;
    const int u = tid - grp * Cfg::GROUP;
    const int strip = u / Cfg::COMBOS;
    int y, dyi;
    combo_of<TY>(u - strip * Cfg::COMBOS, y, dyi);
    const float inv_div = (float)g.C;  // k == 1: nelems = C (correlation_cuda_kernel.cu:85)

    // shared-memory float offsets that do not depend on stage / channel
    const int x2_off = (y + dyi) * Cfg::XS + strip * 8;
    int x1_off[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) x1_off[h] = y * TX + swz_chunk<TX>(y, strip * 2 + h) * 4;  // row = c*TY + y; TY % 8 == 0 or TX < 32

    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int n = tile / (a.tiles_x * a.tiles_y);
      const int trem = tile - n * (a.tiles_x * a.tiles_y);
      const int by0 = (trem / a.tiles_x) * TY, bx0 = (trem % a.tiles_x) * TX;

      float acc[8 * kD];
#pragma unroll
      for (int i = 0; i < 8 * kD; ++i) acc[i] = 0.f;

      for (int ck = 0; ck < a.nchunks; ++ck) {
        mbar_wait(&full_bar[stage], phase);
        const float* x1p = x1s + stage * Cfg::X1_STAGE;
        const float* x2p = x2s + stage * Cfg::X2_STAGE + x2_off;
#pragma unroll
        for (int cc = 0; cc < kCC / KS; ++cc) {
          const int c = cc * KS + grp;
          float av[8], bv[16];
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
#pragma unroll
          for (int px = 0; px < 8; ++px)
#pragma unroll
            for (int dx = 0; dx < kD; ++dx) acc[px * kD + dx] = fmaf(av[px], bv[px + dx], acc[px * kD + dx]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }

      // ---------------- epilogue: /C, LeakyReLU, stage tile, TMA store ----------------
      if (a.use_tma_out) {
        if (tid == 0) tma_store_wait_read0();  // previous tile's store has finished reading `outs`
      }
      named_bar_sync(1, Cfg::NCONS);
      float* obuf = outs + grp * Cfg::OUT_TILE;
#pragma unroll
      for (int dx = 0; dx < kD; ++dx) {
        const int row = (dyi * kD + dx) * TY + y;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v;
          float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float r = acc[(4 * h + e) * kD + dx];
            if constexpr (KS == 1) {
              r = __fdiv_rn(r, inv_div);
              if (g.has_act) r = leaky(r, g.slope);
            }
            vv[e] = r;
          }
          *reinterpret_cast<float4*>(obuf + row * TX + swz_chunk<TX>(row, strip * 2 + h) * 4) = v;
        }
      }
      if constexpr (KS > 1) {
        named_bar_sync(1, Cfg::NCONS);
        for (int e = tid; e < Cfg::OUT_TILE; e += Cfg::NCONS) {
          float s = outs[e];
#pragma unroll
          for (int k2 = 1; k2 < KS; ++k2) s += outs[k2 * Cfg::OUT_TILE + e];
          s = __fdiv_rn(s, inv_div);
          if (g.has_act) s = leaky(s, g.slope);
          outs[e] = s;
        }
      }
      if (a.use_tma_out) {
        fence_proxy_async_smem();
        named_bar_sync(1, Cfg::NCONS);
        if (tid == 0) {
          tma_store_4d(&tm_out, outs, bx0, by0, 0, n);
          tma_store_commit();
        }
      } else {
        named_bar_sync(1, Cfg::NCONS);
        T* outp = (T*)a.out + (long long)n * g.os[0];
        for (int e = tid; e < Cfg::OUT_TILE; e += Cfg::NCONS) {
          const int row = e / TX, x = e - row * TX;
          const int plane = row / TY, yy = row - plane * TY;
          const int oy = by0 + yy, ox = bx0 + x;
          if (oy < g.outH && ox < g.outW) {
            const float v = outs[row * TX + swz_chunk<TX>(row, x >> 2) * 4 + (x & 3)];
            outp[(long long)plane * g.os[1] + (long long)oy * g.os[2] + ox] = from_f32<T>(v);
          }
        }
        named_bar_sync(1, Cfg::NCONS);  // `outs` is rewritten by the next tile's epilogue
      }
    }
    if (a.use_tma_out && tid == 0) tma_store_wait_all0();
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
    const T* x2n = x2 + (long long)n * g.x2s[0];
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
static bool make_tmap_f32(CUtensorMap* tm, const void* base, int W, int H, int C, int B, const long long strides[3],
                          int bx, int by, int bc, bool swizzle128) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  if (((uintptr_t)base & 15) != 0) return false;
  for (int i = 0; i < 3; ++i)
    if (strides[i] <= 0 || (strides[i] * 4) % 16 != 0) return false;
  if (bx > 256 || by > 256 || bc > 256) return false;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)strides[2] * 4, (cuuint64_t)strides[1] * 4, (cuuint64_t)strides[0] * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bc, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int num_sms() {
  static int n = []() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

template <typename T, int TY, int TX, int KS>
static cudaError_t launch_fast(const Geom& g, const void* x1, const void* x2, const float* flow, void* out,
                               int force_no_tma, cudaStream_t stream) {
  using Cfg = FwdCfg<TY, TX, KS>;
  FwdArgs a;
  a.g = g;
  a.x1 = x1; a.x2 = x2; a.flow = flow; a.out = out;
  a.off = g.md - g.pad;
  a.tiles_x = (g.outW + TX - 1) / TX;
  a.tiles_y = (g.outH + TY - 1) / TY;
  a.total_tiles = g.B * a.tiles_x * a.tiles_y;
  a.nchunks = (g.C + kCC - 1) / kCC;
  CUtensorMap tm_x1, tm_out;
  memset(&tm_x1, 0, sizeof(tm_x1));
  memset(&tm_out, 0, sizeof(tm_out));
  a.use_tma_in = 0;
  a.use_tma_out = 0;
  if (std::is_same<T, float>::value && !force_no_tma) {
    a.use_tma_in = make_tmap_f32(&tm_x1, x1, g.W, g.H, g.C, g.B, g.x1s, TX, TY, kCC, TX == 32) ? 1 : 0;
    a.use_tma_out = make_tmap_f32(&tm_out, out, g.outW, g.outH, g.D2, g.B, g.os, TX, TY, kD2, TX == 32) ? 1 : 0;
  }
  auto kern = warp_corr_fwd_kernel<T, TY, TX, KS>;
  static bool attr_set = false;  // benign race: the attribute call is idempotent
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = a.total_tiles < num_sms() ? a.total_tiles : num_sms();
  kern<<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream>>>(a, tm_x1, tm_out);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_fwd_t(const Geom& g, const void* x1, const void* x2, const float* flow, void* out,
                                int variant, cudaStream_t stream) {
  const bool fast_ok = g.k == 1 && g.s1 == 1 && g.s2 == 1 && g.md == kMD && variant != CERB_FWD_VARIANT_GENERIC;
  if (fast_ok) {
    const int no_tma = (variant == CERB_FWD_VARIANT_FAST_NOTMA || variant == CERB_FWD_VARIANT_SMALL_NOTMA) ? 1 : 0;
    bool small = variant == CERB_FWD_VARIANT_SMALL || variant == CERB_FWD_VARIANT_SMALL_NOTMA;
    if (variant == CERB_FWD_VARIANT_AUTO) {
      // not enough 8x32 tiles to fill the GPU: 4x16 tiles with the channels split four ways in-CTA
      const long long big_tiles = (long long)g.B * ((g.outW + 31) / 32) * ((g.outH + 7) / 8);
      small = big_tiles < (long long)num_sms() * 3 / 4;
    }
    if (small) return launch_fast<T, 4, 16, 4>(g, x1, x2, flow, out, no_tma, stream);
    return launch_fast<T, 8, 32, 1>(g, x1, x2, flow, out, no_tma, stream);
  }
  const long long total = (long long)g.B * g.D2 * g.outH * g.outW;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  corr_fwd_generic_kernel<T><<<(int)blocks, 256, 0, stream>>>(g, (const T*)x1, (const T*)x2, flow, (T*)out);
  return cudaGetLastError();
}

cudaError_t launch_warp_corr_forward(const Geom& g, int dtype, const void* x1, const void* x2, const float* flow,
                                     void* out, int variant, cudaStream_t stream) {
  switch (dtype) {
    case CERB_F32: return launch_fwd_t<float>(g, x1, x2, flow, out, variant, stream);
    case CERB_F16: return launch_fwd_t<__half>(g, x1, x2, flow, out, variant, stream);
    case CERB_BF16: return launch_fwd_t<__nv_bfloat16>(g, x1, x2, flow, out, variant, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cerb
