// Global fp32 reduction throughput on B200 for the backward's bilinear splat (the gradient with respect to the warped
// map goes through 4 taps per position and channel into grad_x2): which form of red.global the L2 sustains.
//   scalar_iid     4 scalar red.global.add.f32 per (position, channel), per-pixel random offsets (+-6 px): the bench's flow
//   scalar_smooth  the same with one offset per 8x16 tile (a decoder-like smooth flow: a warp's addresses are contiguous runs)
//   v2_iid         horizontal tap pairs as red.global.add.v2.f32 (offsets forced even: the aligned best case)
//   v4_stream      coalesced red.global.add.v4.f32 over the whole gradient (what a shared-memory box flush would issue)
//   bulk_reduce    cp.reduce.async.bulk.global.shared::cta .add.f32 of 4 KB rows from shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int B = 8, C = 48, H = 128, W = 256;

__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// thread = pixel (8 x 16 tiles, warp = 2 rows of 16 like a TMEM lane quadrant); mode 0 iid, 1 smooth, 2 v2 iid, 3 v2 smooth
__global__ void __launch_bounds__(128) k_splat(float* g, int mode) {
  const int tiles_x = W / 16, tiles_y = H / 8;
  const int tile = blockIdx.x, n = tile / (tiles_x * tiles_y), tr = tile % (tiles_x * tiles_y);
  const int y = (tr / tiles_x) * 8 + (threadIdx.x >> 4), x = (tr % tiles_x) * 16 + (threadIdx.x & 15);
  const uint32_t h = hash((mode & 1) ? (uint32_t)tile : (uint32_t)(tile * 128 + threadIdx.x));
  int rx = (int)(h % 13u) - 6, ry = (int)((h >> 8) % 13u) - 6;
  int x0 = min(max(x + rx, 0), W - 2);
  const int y0 = min(max(y + ry, 0), H - 2);
  if (mode >= 2) x0 &= ~1;   // (aligned best case: in the kernel only even x0 can use the pair form)
  float* p = g + (long long)n * C * H * W + (long long)y0 * W + x0;
  const float v = 1.0f + (float)(h & 7u);
  if (mode < 2) {
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      float* pc = p + (long long)c * H * W;
      red1(pc, v); red1(pc + 1, v); red1(pc + W, v); red1(pc + W + 1, v);
    }
  } else {
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      float* pc = p + (long long)c * H * W;
      red2(pc, v, v); red2(pc + W, v, v);
    }
  }
}

// coalesced v4 reductions over `frac16`/16 of every plane (box flush: more or fewer elements than positions)
__global__ void __launch_bounds__(256) k_v4(float* g, long long n4) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) red4(g + 4 * i, 1.f, 2.f, 3.f, 4.f);
}

// bulk reduce: every CTA keeps `depth` 4 KB rows in flight from shared memory
__global__ void __launch_bounds__(32) k_bulk(float* g, long long rows, int depth) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (int i = threadIdx.x; i < 4096 * 4 / 4; i += 32) ((float*)sm)[i] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x != 0) return;
  int k = 0;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x, ++k) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + (k & 3) * 4096);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 4096;" ::"l"(g + r * 1024), "r"(s) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const long long elems = (long long)B * C * H * W;
  float* g;
  CK(cudaMalloc(&g, elems * 4));
  CK(cudaMemset(g, 0, elems * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int tiles = B * (H / 8) * (W / 16);
  auto timeit = [&](const char* name, auto fn, double ops, double bytes) -> int {
    for (int i = 0; i < 3; ++i) fn();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) fn();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps;
    printf("{\"case\": \"%s\", \"us\": %.1f, \"Gops_per_s\": %.1f, \"GB_per_s\": %.1f}\n", name, us, ops / us / 1e3, bytes / us / 1e3);
    return 0;
  };
  const double taps = (double)B * H * W * C * 4;
  const char* names[4] = {"scalar_iid", "scalar_smooth", "v2_iid_aligned", "v2_smooth_aligned"};
  for (int mode = 0; mode < 4; ++mode)
    if (timeit(names[mode], [&] { k_splat<<<tiles, 128>>>(g, mode); }, mode < 2 ? taps : taps / 2, taps * 4)) return 1;
  if (timeit("v4_stream_all", [&] { k_v4<<<148 * 8, 256>>>(g, elems / 4); }, elems / 4.0, elems * 4.0)) return 1;
  if (timeit("memset_all", [&] { cudaMemsetAsync(g, 0, elems * 4); }, 0, elems * 4.0)) return 1;
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  if (timeit("bulk_reduce_4KB_rows_depth4", [&] { k_bulk<<<148 * 4, 32, 16384>>>(g, elems / 1024, 4); }, elems / 1024.0, elems * 4.0)) return 1;
  if (timeit("bulk_reduce_4KB_rows_depth1", [&] { k_bulk<<<148 * 4, 32, 16384>>>(g, elems / 1024, 1); }, elems / 1024.0, elems * 4.0)) return 1;
  return 0;
}
