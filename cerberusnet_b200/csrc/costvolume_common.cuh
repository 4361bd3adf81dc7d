// costvolume_common.cuh -- shared device helpers for the sm_100a cost-volume kernels.
//
// Semantics restated from SURVEY.md section 8(a); reference sites (relative to the reference
// checkout) are cited next to the arithmetic they pin.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/cerberus_costvolume.h"

namespace cerb {

// ---------------------------------------------------------------- element I/O -------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ float ldg_f32(const T* p) { return to_f32<T>(__ldg(p)); }
// streaming load that does not allocate in L1 (ld.global.cg): for planes a power-of-two stride
// apart, which would all fight for the same L1 set
template <typename T> __device__ __forceinline__ float ldcg_f32(const T* p) { return to_f32<T>(__ldcg(p)); }

// ---------------------------------------------------------------- problem geometry --------
struct Geom {
  int B, C, H, W;            // inputs
  int pad, k, md, s1, s2;    // correlation parameters
  int kr, r, D, D2;          // kernel radius, displacement radius (md / s2), 2r+1, D*D
  int outH, outW;
  int warp_mode;             // CERB_WARP_*
  int has_act;               // LeakyReLU fused
  float slope;
  int unnorm_fma;            // ATen un-normalise ((g+1)*size-1)/2 contracted to one FMA (nvcc -fmad) or not
  long long x1s[3], x2s[3], fls[3], os[3];  // N, C, H strides in elements (W stride == 1)
  int x2roll;                // batch item n of x1 / flow / out is paired with item (n + x2roll) mod B of x2
};

// Batch item of the second map paired with item n (cerb_corr_params.x2_batch_roll): with x1 = x2 = the features of
// [image 1; image 2] and a roll of B/2, one launch computes both flow directions of the reference's
// `consistency=True` forward (nnet_models/pwcnet.py:108-113, cerberus.py:131-135) without copying a feature map.
__host__ __device__ __forceinline__ int x2_item(const Geom& g, int n) {
  const int m = n + g.x2roll;
  return m >= g.B ? m - g.B : m;
}

// ---------------------------------------------------------------- flow warp ---------------
// a / c correctly rounded (Markstein: q = a*rc, r = a - q*c exact by FMA, q' = q + r*rc) for a
// finite, non-tiny `a`, a positive integer-valued `c` and rc = RN(1/c).  Same bits as IEEE
// division without its multi-branch slow path; used where the reference divides by a constant.
__device__ __forceinline__ float div_const(float a, float c, float rc) {
  const float q = __fmul_rn(a, rc);
  const float r = __fmaf_rn(-q, c, a);
  return __fmaf_rn(r, rc, q);
}

// Sample position along one axis for output pixel `pix` displaced by `disp`.
//   grid  g = 2*(pix+disp)/(size-1) - 1     UnFlowLoss.py:16-19,30-31,89-91: three separate fp32
//                                           roundings (separate torch kernels).  On CUDA -- the
//                                           reference's training path -- ATen evaluates
//                                           `tensor / python_scalar` as tensor * RN(1/scalar);
//                                           its CPU kernels divide.  CERB_WARP_TORCH follows CUDA,
//                                           CERB_WARP_TORCH_CPU the true division (the golden
//                                           fixtures were generated on CPU).
//   TORCH p = ((g+1)*size - 1)/2            ATen grid_sampler_unnormalize(align_corners=False)
//   TRT   p = ((g+1)*(size-1))/2            trt_plugins/grid_sampler.cu:55-58
//   border clip to [0,size-1]               clip_coordinates; grid_sampler.cu:62-64
// inside = un-clipped position strictly inside (0,size-1) -> gradient passes (ATen
// clip_coordinates_set_grad), else zero.
// Per-axis constants of the sample position, computed once per thread (the reciprocal is a
// MUFU + Newton step + denormal branch: not something to redo for every sample).
struct AxisConst {
  float fsize, sm1, rsm1;
};
__device__ __forceinline__ AxisConst make_axis(int size) {
  AxisConst k;
  k.fsize = (float)size;
  k.sm1 = (float)(size - 1);
  k.rsm1 = __frcp_rn(k.sm1);
  return k;
}

__device__ __forceinline__ float sample_pos(int pix, float disp, const AxisConst& k, int mode, bool& inside, int unnorm_fma = 1) {
  float v = __fadd_rn((float)pix, disp);
  v = __fmul_rn(2.0f, v);
  v = (mode == CERB_WARP_TORCH) ? __fmul_rn(v, k.rsm1) : div_const(v, k.sm1, k.rsm1);
  const float g = __fadd_rn(v, -1.0f);
  float p;
  if (mode == CERB_WARP_TRT)
    p = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), k.sm1), 0.5f);  // (.)/2 is exact as *0.5
  else
    p = unnorm_fma ? __fmul_rn(__fmaf_rn(__fadd_rn(g, 1.f), k.fsize, -1.f), 0.5f)
                   : __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), k.fsize), -1.f), 0.5f);
  inside = (p > 0.f) && (p < k.sm1);
  return fminf(k.sm1, fmaxf(p, 0.f));
}

__device__ __forceinline__ float sample_pos(int pix, float disp, int size, int mode, bool& inside, int unnorm_fma = 1) {
  return sample_pos(pix, disp, make_axis(size), mode, inside, unnorm_fma);
}

__device__ __forceinline__ float pos_scale(int size, int mode) {
  return (mode == CERB_WARP_TRT) ? 1.0f : __fdiv_rn((float)size, (float)(size - 1));
}

// Bilinear tap set of one warped position: offsets (within one H*W plane, using the H stride)
// of the 4 taps and their weights, in ATen order nw, ne, sw, se
// (weights: nw=(x_se-x)*(y_se-y) ... ; a tap outside the image gets weight 0 and a clamped,
// always-dereferenceable offset).
struct Taps {
  int off[4];
  float w[4];
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W, long long hstride) {
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const int x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = (float)x1 - ix, wx0 = ix - (float)x0;
  const float wy1 = (float)y1 - iy, wy0 = iy - (float)y0;
  const bool bx1 = x1 < W, by1 = y1 < H;  // x0,y0 are in range after the clip
  const int cx1 = bx1 ? x1 : x0, cy1 = by1 ? y1 : y0;
  Taps t;
  t.off[0] = (int)(y0 * hstride) + x0;
  t.off[1] = (int)(y0 * hstride) + cx1;
  t.off[2] = (int)(cy1 * hstride) + x0;
  t.off[3] = (int)(cy1 * hstride) + cx1;
  t.w[0] = wx1 * wy1;
  t.w[1] = bx1 ? wx0 * wy1 : 0.f;
  t.w[2] = by1 ? wx1 * wy0 : 0.f;
  t.w[3] = (bx1 && by1) ? wx0 * wy0 : 0.f;
  return t;
}

// acc = 0; acc += v_nw*w_nw; ... in ATen's order (each step one FMA under -fmad).
__device__ __forceinline__ float blend(float v0, float v1, float v2, float v3, const Taps& t) {
  float a = __fmul_rn(v0, t.w[0]);
  a = __fmaf_rn(v1, t.w[1], a);
  a = __fmaf_rn(v2, t.w[2], a);
  a = __fmaf_rn(v3, t.w[3], a);
  return a;
}

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

// ---------------------------------------------------------------- PTX wrappers ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Wait with a suspend-time hint (ns): a waiting warp may sleep in hardware up to that long per attempt instead of
// re-polling; used by roles that typically wait long (consumers on `full`) so their polling does not compete with
// the staging warps' shared-memory traffic.
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "HWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni HWAIT_DONE;\n\t"
      "bra.uni HWAIT_LOOP;\n\t"
      "HWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(hint_ns)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(sleep_ns);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Sub-block barrier.  `bar.sync` is `barrier.sync.aligned`: every thread of a warp must execute it
// convergently, which a warp coming out of an `if (lane == 0)` block does not guarantee (found by
// compute-sanitizer synccheck).  The non-aligned form counts per-thread arrivals and is safe.
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  __syncwarp();
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// TMA: 4-D tiled load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA: 4-D tiled store shared -> global (SASS: UTMASTG); out-of-range elements are clipped
__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* tmap, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- programmatic dependent launch: the next kernel in the stream may start its launch and setup
// while this grid is still running; it must not touch global memory produced by its predecessor
// before pdl_wait() (which returns once the predecessor has completed and flushed).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- Ampere-style asynchronous copies (per-thread, no register staging).  src_bytes < size zero-fills the rest.
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- thread-block cluster / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni CWAIT_DONE;\n\t"
      "bra.uni CWAIT_LOOP;\n\t"
      "CWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// arrive without its own release: the caller has issued fence.acq_rel.cluster after its writes
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_dsmem_v4(uint32_t cluster_addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// scalar shared-memory access by 32-bit shared address: a constant added to `addr` folds into the instruction's
// immediate offset (LDS R, [R + imm])
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace cerb
