// TMA tile-load throughput per SM as a function of box shape (fp32, NCHW 1x32x128x256 like the
// finest pyramid level).  One elected thread per CTA keeps `depth` boxes in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma tma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) k_tma(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2, int alt, int box_bytes, int depth, int iters,
                                            int W, int H, int C, int bw, int bh, int bc, int xalign, int seq, long long* cyc, int nw_tma, int lds_iters) {
  extern __shared__ __align__(1024) unsigned char smem_all[];
  __shared__ uint64_t bars_all[64];
  const int warp = threadIdx.x >> 5;
  if (warp >= nw_tma) {
    // competing shared-memory load traffic (like the correlation consumers)
    float4* lp = (float4*)(smem_all + 160 * 1024);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < lds_iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float4 v;
        unsigned a = (unsigned)__cvta_generic_to_shared(lp + ((threadIdx.x * 3 + u * 67 + i * 13) & 1023));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (acc.x + acc.y + acc.z + acc.w == 1234.5f) cyc[0] = 1;
    return;
  }
  uint64_t* bars = bars_all + warp * 8;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if ((threadIdx.x & 31) != 0) return;
  const int stage_bytes = (box_bytes + 1023) / 1024 * 1024;
  unsigned char* smem = smem_all + warp * depth * stage_bytes;
  unsigned seed = (blockIdx.x * 8 + warp) * 7919u + 13u;
  long long t0 = 0;
  uint32_t phase_bits = 0;
  for (int it = 0; it < iters + depth; ++it) {
    if (it == depth) t0 = clock64();
    const int s = it % depth;
    if (it >= depth) {  // wait for the box issued `depth` iterations ago in this slot
      uint32_t par = (phase_bits >> s) & 1u;
      asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni D;\n\tbra.uni W;\n\tD:\n\t}" ::"r"(su32(&bars[s])), "r"(par) : "memory");
      phase_bits ^= 1u << s;
    }
    if (it < iters) {
      seed = seed * 1664525u + 1013904223u;
      int x = (int)((seed >> 8) & 127u) & ~(xalign - 1);
      int y = (int)((seed >> 16) & 63u);
      int c = (int)((seed >> 4) & 15u);
      asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(su32(&bars[s])), "r"(box_bytes) : "memory");
      const CUtensorMap* tmp = (alt && (it & 1)) ? &tm2 : &tm;
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(su32(smem + s * stage_bytes)), "l"(tmp), "r"(su32(&bars[s])), "r"(x), "r"(y), "r"(c), "r"(0) : "memory");
    }
  }
  if (warp == 0) cyc[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int W = 256, H = 128, C = 32;
  float* d; CK(cudaMalloc(&d, sizeof(float) * W * H * C)); CK(cudaMemset(d, 0, sizeof(float) * W * H * C));
  float* d2; CK(cudaMalloc(&d2, sizeof(float) * W * H * C)); CK(cudaMemset(d2, 0, sizeof(float) * W * H * C));
  long long* dc; CK(cudaMalloc(&dc, 8 * 256));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  PFN enc = (PFN)fp;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Shape { int bw, bh, bc, swz, xalign; const char* name; };
  Shape shapes[] = {{32, 8, 4, 1, 32, "x1 32x8x4 swz"}, {60, 30, 4, 0, 4, "raw 60x30x4"}, {60, 30, 8, 0, 4, "raw 60x30x8"}};
  printf("{\"grid\": %d", sms);
  for (auto& sh : shapes) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, 1};
    cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)sh.bw, (cuuint32_t)sh.bh, (cuuint32_t)sh.bc, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sh.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap tm2;
    enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d2, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        sh.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf(", \"%s\": \"encode failed %d\"", sh.name, (int)r); continue; }
    const int bytes = sh.bw * sh.bh * sh.bc * 4;
    for (int nlds : {0}) for (int depth : {3}) for (int seq : {0, 1}) {
      const int nw = 1;
      const int stage = (bytes + 1023) / 1024 * 1024;
      if (stage * depth * nw > 150 * 1024) continue;
      const int iters = 300;
      k_tma<<<sms, 32 * (nw + nlds), 176 * 1024>>>(tm, tm2, seq, bytes, depth, iters, W, H, C, sh.bw, sh.bh, sh.bc, sh.xalign, seq, dc, nw, 40000);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(sms);
      CK(cudaMemcpy(h.data(), dc, 8 * sms, cudaMemcpyDeviceToHost));
      std::sort(h.begin(), h.end());
      double cyc = (double)h[sms / 2] / iters;
      printf(", \"%s alt_desc=%d d%d\": {\"cyc_per_box\": %.0f, \"B_per_cyc\": %.1f}", sh.name, seq, depth, cyc, nw * bytes / cyc);
    }
  }
  printf("}\n");
  return 0;
}
