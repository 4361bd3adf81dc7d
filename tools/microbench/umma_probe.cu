// umma_probe.cu -- stand-alone check of the tcgen05 building blocks the tensor-core forward rests on:
//   K-major SWIZZLE_128B shared-memory operand tiles written by ordinary threads (hi / lo TF32 split of fp32 data),
//   tcgen05.mma.kind::tf32 (M = 128, N = 192 twice, K = 8 per instruction, 3 products per K step = "3xTF32"),
//   tcgen05.commit onto an mbarrier, tcgen05.ld 32x32b from the accumulator.
// D[128][384] = A[128][K] * B[384][K]^T is compared on the host with a float64 product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/microbench/umma_probe tools/microbench/umma_probe.cu
//   tools/microbench/umma_probe [K=32] [mode: 3 = 3xTF32, 1 = single TF32]
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni W_DONE;\n\t"
      "bra.uni W_LOOP;\n\t"
      "W_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

constexpr int M = 128, N = 384, NH = 192;

// K-major SWIZZLE_128B descriptor: rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// A operand from tensor memory (lane = row, 8 consecutive 32-bit columns = the K = 8 tf32 values of one instruction)
__device__ __forceinline__ void umma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                 int K, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* a_hi = (float*)smem;                      // [128][32]
  float* a_lo = a_hi + M * 32;
  float* b_hi = a_lo + M * 32;                     // [384][32]
  float* b_lo = b_hi + N * 32;
  uint64_t* bar = (uint64_t*)(b_lo + N * 32);
  uint32_t* tmem_slot = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t parity = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    // fill the operand tiles: element (row, c) at row * 128 + ((c / 4) ^ (row % 8)) * 16 + (c % 4) * 4 bytes
    for (int e = tid; e < (M + N) * 32; e += blockDim.x) {
      const int row = e >> 5, c = e & 31;
      const bool isA = row < M;
      const int r = isA ? row : row - M;
      float v = 0.f;
      if (k0 + c < K) v = isA ? A[(size_t)r * K + k0 + c] : B[(size_t)r * K + k0 + c];
      uint32_t hib;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hib) : "f"(v));
      const float hi = __uint_as_float(hib);
      const float lo = v - hi;
      const int off = r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3);
      (isA ? a_hi : b_hi)[off] = hi;
      (isA ? a_lo : b_lo)[off] = lo;
    }
    if (mode == 4) {   // thread = row: the 32 channels of this chunk as hi / lo, 8 columns per K step, straight into TMEM
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t h[8], l[8];
        for (int c = 0; c < 8; ++c) {
          const float v = (k0 + ks * 8 + c < K) ? A[(size_t)tid * K + k0 + ks * 8 + c] : 0.f;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[c]) : "f"(v));
          l[c] = __float_as_uint(v - __uint_as_float(h[c]));
        }
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 384 + ks * 16;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(h[0]), "r"(h[1]),
                     "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta + 8), "r"(l[0]), "r"(l[1]),
                     "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0 && mode == 4) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < 4; ++ks) {
        for (int h = 0; h < 2; ++h) {
          const uint64_t bh = make_desc(smem_u32(b_hi) + h * NH * 128 + 32 * ks), bl = make_desc(smem_u32(b_lo) + h * NH * 128 + 32 * ks);
          const uint32_t d = tmem + h * NH, ah = tmem + 384 + ks * 16, al = ah + 8;
          umma_tf32_ta(d, ah, bh, idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
          umma_tf32_ta(d, al, bh, idesc, 1u);
          umma_tf32_ta(d, ah, bl, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    if (tid == 0 && mode != 4) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < 4; ++ks) {
        for (int h = 0; h < 2; ++h) {
          const uint64_t ah = make_desc(smem_u32(a_hi) + 32 * ks), al = make_desc(smem_u32(a_lo) + 32 * ks);
          const uint64_t bh = make_desc(smem_u32(b_hi) + h * NH * 128 + 32 * ks), bl = make_desc(smem_u32(b_lo) + h * NH * 128 + 32 * ks);
          const uint32_t d = tmem + h * NH;
          umma_tf32(d, ah, bh, idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
          if (mode == 3) {
            umma_tf32(d, al, bh, idesc, 1u);
            umma_tf32(d, ah, bl, idesc, 1u);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    mbar_wait(&bar[0], parity);   // the operand tiles may be overwritten once the MMAs have retired
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // accumulator -> global: warp w reads TMEM lanes 32w .. 32w+31 (row = lane), 8 columns per load
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 32;
  const int mode = argc > 2 ? atoi(argv[2]) : 3;
  std::vector<float> A((size_t)M * K), B((size_t)N * K), D((size_t)M * N);
  srand(3);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  const int smem = (M + N) * 32 * 4 * 2 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, K, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxref = 0, maxerr = 0; int bad_m = -1, bad_n = -1;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
      const double err = fabs(r - (double)D[(size_t)m * N + n]);
      if (fabs(r) > maxref) maxref = fabs(r);
      if (!(err <= maxerr)) { maxerr = err; bad_m = m; bad_n = n; }
    }
  printf("umma_probe K=%d mode=%d: max|ref|=%.4f max|err|=%.3e rel=%.3e at (%d,%d) got %.6f\n", K, mode, maxref, maxerr,
         maxerr / maxref, bad_m, bad_n, D[(size_t)bad_m * N + bad_n]);
  printf("D[0][0..3] = %.5f %.5f %.5f %.5f   D[127][380..383] = %.5f %.5f %.5f %.5f\n", D[0], D[1], D[2], D[3],
         D[127 * N + 380], D[127 * N + 381], D[127 * N + 382], D[127 * N + 383]);
  const bool ok = maxerr / maxref < (mode >= 3 ? 2e-6 : 2e-3);
  printf(ok ? "PROBE OK\n" : "PROBE FAIL\n");
  return ok ? 0 : 1;
}
