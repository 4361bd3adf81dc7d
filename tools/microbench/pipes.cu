// Microbenchmarks that pin the B200 facts the cost-volume kernels are designed around:
//   1. fp32 FMA peak: scalar FFMA vs packed fma.rn.f32x2 (FFMA2)
//   2. shared-memory LDS.128 / LDS.32 wavefront cost under broadcast patterns
//   3. back-to-back tiny-kernel launch cost (stream, PDL)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------- FFMA ------
template <int NACC>
__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b, int iters) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

// FFMA with two register multiplicands that differ per accumulator (no reuse-cache help)
template <int NACC>
__global__ void __launch_bounds__(256) k_ffma_rr(float* out, const float* in, int iters) {
  float acc[NACC], x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = in[i]; y[i] = in[8 + i]; }
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x[i & 7], y[(i >> 3) & 7], acc[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_ffma2(float* out, const float* in, int iters) {
  unsigned long long acc[NACC], x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 t = make_float2(in[i], in[i + 1]);
    x[i] = *reinterpret_cast<unsigned long long*>(&t);
    float2 u = make_float2(in[8 + i], in[9 + i]);
    y[i] = *reinterpret_cast<unsigned long long*>(&u);
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float2 t = make_float2(threadIdx.x * 0.001f + i, 1.f);
    acc[i] = *reinterpret_cast<unsigned long long*>(&t);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = ffma2(x[i & 7], y[(i >> 3) & 7], acc[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float2 t = *reinterpret_cast<float2*>(&acc[i]);
    s += t.x + t.y;
  }
  if (s == 12345.678f) out[0] = s;
}

// ---------------------------------------------------------------- LDS -------
// pattern -> per-lane 16-byte chunk index
__device__ __forceinline__ int lds_pattern(int pat, int lane) {
  switch (pat) {
    case 0: return lane;            // 32 distinct chunks (512 B)
    case 1: return lane & 7;        // quarter-warps identical (8 distinct)
    case 2: return lane >> 2;       // groups of 4 adjacent lanes share (8 distinct)
    case 3: return 0;               // full broadcast
    case 4: return lane >> 1;       // pairs share (16 distinct)
    case 5: return (lane & 3) + 8 * (lane >> 3);  // each quarter-warp: 4 distinct chunks, different per quarter
    case 6: return (lane >> 3);     // each quarter-warp reads one chunk, 4 distinct total
    case 7: return (lane & 7) * 11; // 8 distinct chunks with row stride 44 floats (=11 chunks): bank-spread
    default: return lane;
  }
}

__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(512) k_lds128(float* out, long long* cycles, int pat, int iters) {
  extern __shared__ float4 sm4[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm4[i] = make_float4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  int lane = threadIdx.x & 31;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm4);
  int idx = lds_pattern(pat, lane);
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a = base + 16u * ((unsigned)(idx + u * 352 + it * 32) & 2047u);
      float4 v = lds128(a);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ int lds32_pattern(int pat, int lane) {
  switch (pat) {
    case 0: return lane;             // 32 distinct words, 32 banks
    case 1: return (lane & 7);       // 8 distinct words
    case 2: return 0;                // broadcast
    case 3: return (lane % 12) * 45; // 12 distinct words in 12 distinct banks (row-like)
    case 4: return lane * 32;        // 32-way conflict
    case 5: return lane * 2;         // 2-way conflict
    default: return lane;
  }
}

__global__ void __launch_bounds__(512) k_lds32(float* out, long long* cycles, int pat, int iters) {
  extern __shared__ float sm1[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm1[i] = i;
  __syncthreads();
  int lane = threadIdx.x & 31;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm1);
  int idx = lds32_pattern(pat, lane);
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += lds32(base + 4u * ((unsigned)(idx + u * 1056 + it * 32) & 8191u));
  }
  long long t1 = clock64();
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// LDG gather from an L1/L2-resident 64 KB window: 4 taps per sample like the bilinear warp
__global__ void __launch_bounds__(512) k_ldg_gather(float* out, const float* __restrict__ src, long long* cycles,
                                                    int iters, int stride) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int p = ((it * 4 + u) * 37 + warp * 131 + lane) & 8191;
      const float* q = src + p;
      acc += __ldg(q) + __ldg(q + 1) + __ldg(q + stride) + __ldg(q + stride + 1);
    }
  }
  long long t1 = clock64();
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- launch ----
__global__ void k_empty(float* out) { if (out == nullptr) printf("x"); }

template <typename F>
static float time_ms(F&& f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* d_out; CK(cudaMalloc(&d_out, 1024));
  float h_in[32]; for (int i = 0; i < 32; ++i) h_in[i] = 1.0f + 1e-6f * i;
  float* d_in; CK(cudaMalloc(&d_in, sizeof(h_in)));
  CK(cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * 4096));

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d", prop.name, sms, clk_khz);

  // ---- FMA peaks: grid = sms*8 blocks of 256 threads (2 blocks/SMSP-ish), long enough to dominate launch
  {
    const int iters = 4096;
    const int NACC = 32;
    int grid = sms * 8;
    double flops = 2.0 * NACC * (double)iters * 256.0 * grid;
    float ms1 = time_ms([&] { k_ffma<NACC><<<grid, 256>>>(d_out, 1.0001f, 0.5f, iters); }, 10);
    float ms2 = time_ms([&] { k_ffma_rr<NACC><<<grid, 256>>>(d_out, d_in, iters); }, 10);
    float ms3 = time_ms([&] { k_ffma2<NACC><<<grid, 256>>>(d_out, d_in, iters); }, 10);
    CK(cudaGetLastError());
    printf(", \"ffma_imm_tflops\": %.2f, \"ffma_rr_tflops\": %.2f, \"ffma2_tflops\": %.2f",
           flops / ms1 * 1e-9, flops / ms2 * 1e-9, 2.0 * flops / ms3 * 1e-9);
    // sustained (2 s) FFMA2 to see power-capped clocks
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int n = 0;
    for (; n < 400; ++n) k_ffma2<NACC><<<grid, 256>>>(d_out, d_in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"ffma2_sustained_tflops\": %.2f, \"ffma2_sustained_window_ms\": %.1f", 2.0 * flops * n / ms * 1e-9, ms);
    cudaEventRecord(e0);
    for (n = 0; n < 200; ++n) k_ffma_rr<NACC><<<grid, 256>>>(d_out, d_in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"ffma_rr_sustained_tflops\": %.2f", flops * n / ms * 1e-9);
  }

  // ---- LDS patterns: one block of W warps per SM, cycles per warp-instruction seen by the SM
  {
    CK(cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CK(cudaFuncSetAttribute(k_lds32, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    const int iters = 2048;
    for (int warps : {1, 4, 16}) {
      printf(", \"lds128_w%d_cyc_per_instr_SM\": [", warps);
      for (int pat = 0; pat < 8; ++pat) {
        k_lds128<<<sms, warps * 32, 32768>>>(d_out, d_cyc, pat, iters);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        double cyc = (double)h[sms / 2] / ((double)iters * 8 * warps);
        printf("%s%.3f", pat ? ", " : "", cyc);
      }
      printf("]");
      printf(", \"lds32_w%d_cyc_per_instr_SM\": [", warps);
      for (int pat = 0; pat < 6; ++pat) {
        k_lds32<<<sms, warps * 32, 32768>>>(d_out, d_cyc, pat, iters);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        double cyc = (double)h[sms / 2] / ((double)iters * 8 * warps);
        printf("%s%.3f", pat ? ", " : "", cyc);
      }
      printf("]");
    }
  }

  {
    float* d_src; CK(cudaMalloc(&d_src, 4 * 32768)); CK(cudaMemset(d_src, 0, 4 * 32768));
    const int iters = 1024;
    for (int warps : {4, 16}) {
      k_ldg_gather<<<sms, warps * 32>>>(d_out, d_src, d_cyc, iters, 256);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(sms);
      CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      std::sort(h.begin(), h.end());
      printf(", \"ldg_gather_w%d_cyc_per_ldg_SM\": %.3f", warps, (double)h[sms / 2] / ((double)iters * 16 * warps));
    }
  }

  // ---- launch costs
  {
    const int n = 1000;
    float ms = time_ms([&] { for (int i = 0; i < n; ++i) k_empty<<<1, 32>>>(d_out); }, 5);
    printf(", \"empty_launch_stream_us\": %.3f", ms * 1000.f / n);
    ms = time_ms([&] { for (int i = 0; i < n; ++i) k_empty<<<sms * 2, 256>>>(d_out); }, 5);
    printf(", \"empty_launch_296x256_stream_us\": %.3f", ms * 1000.f / n);
    // CUDA graph of n empty kernels
    cudaStream_t s; cudaStreamCreate(&s);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < n; ++i) k_empty<<<sms * 2, 256, 0, s>>>(d_out);
    cudaStreamEndCapture(s, &g);
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaEventSynchronize(e1);
      float t; cudaEventElapsedTime(&t, e0, e1); best = std::min(best, t);
    }
    printf(", \"empty_launch_296x256_graph_us\": %.3f", best * 1000.f / n);
    // PDL launches
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms * 2); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0, s);
      for (int i = 0; i < n; ++i) cudaLaunchKernelEx(&cfg, k_empty, d_out);
      cudaEventRecord(e1, s); cudaEventSynchronize(e1);
      float t; cudaEventElapsedTime(&t, e0, e1); best = std::min(best, t);
    }
    CK(cudaGetLastError());
    printf(", \"empty_launch_296x256_pdl_us\": %.3f", best * 1000.f / n);
  }
  printf("}\n");
  return 0;
}
