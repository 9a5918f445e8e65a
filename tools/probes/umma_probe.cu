// Micro-probe: latency/throughput of chains of small tcgen05.mma (kind::f16) instructions on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slotdiffusion_b200/csrc tools/probes/umma_probe.cu -o tools/probes/umma_probe
#include <cstdio>
#include <cuda_fp16.h>
#include "ptx.cuh"
using namespace sdb;

__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void umma_f16_pred(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}

// mode 0: A,B K-major, one accumulator; 1: A MN-major; 2: two alternating accumulators; 3: same smem address every MMA
// mode 4: the whole warp runs the loop convergently, the MMA itself is predicated on lane 0 (uniform operands)
__global__ void probe(int M, int N, int count, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tm;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tm, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tm, 0);
  if (mode == 4) {
    if (threadIdx.x < 32) {
      const uint32_t idesc = umma_idesc_f16(M, N);
      const uint32_t a0 = smem_u32(base), b0 = a0 + 32768;
      const uint32_t issue = threadIdx.x == 0;
      long long t0 = clock64();
#pragma unroll 8
      for (int i = 0; i < count; ++i) {
        const int k = i & 3;
        uint64_t da = umma_desc_kmajor_sw128(a0) + (uint64_t)((k * 32) >> 4);
        uint64_t db = umma_desc_kmajor_sw128(b0) + (uint64_t)((k * 32) >> 4);
        umma_f16_pred(tmem, da, db, idesc, i > 0 ? 1u : 0u, issue);
      }
      long long t1 = clock64();
      if (threadIdx.x == 0) umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t2 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else
  if (threadIdx.x == 0) {
    uint32_t idesc = umma_idesc_f16(M, N) | (mode == 1 ? (1u << 15) : 0u);
    const uint32_t a0 = smem_u32(base), b0 = a0 + 32768;
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const int k = (mode == 3) ? 0 : (i & 3);
      uint64_t da = (mode == 1) ? desc_mn(a0, 16384) + (uint64_t)((k * 2048) >> 4) : umma_desc_kmajor_sw128(a0) + (uint64_t)((k * 32) >> 4);
      uint64_t db = umma_desc_kmajor_sw128(b0) + (uint64_t)((k * 32) >> 4);
      uint32_t d = tmem + ((mode == 2) ? (i & 1) * 256 : 0);
      umma_f16(d, da, db, idesc, i > 1 ? 1u : 0u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ms[] = {128};
  const int Ns[] = {16, 32, 64};
  for (int mode : {0, 4})
    for (int M : Ms)
      for (int N : Ns) {
        if (M == 64 && mode == 1) continue;
        for (int count : {1, 8, 64, 256}) {
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<1, 128, 100 * 1024>>>(M, N, count, mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d M %d N %d count %d: %s\n", mode, M, N, count, cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("mode %d M %3d N %3d count %3d: issue %6lld cyc, done %6lld cyc, %.1f cyc/mma\n", mode, M, N, count, h[0], h[1], (double)h[1] / count);
        }
      }
  return 0;
}
