// Probe: where do the 64 rows of an M = 64 (cta_group::1, kind::f16) tcgen05.mma accumulator live in TMEM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slotdiffusion_b200/csrc tools/probes/umma_m64_probe.cu -o tools/probes/umma_m64_probe.bin
// A[i][0] = i + 1 (other k zero), B[j][0] = 1  ->  D[i][j] = i + 1.  All 128 lanes are cleared first with an M = 128 MMA
// on a zero A tile; then every warp dumps its 32 lanes x 16 columns.
#include <cstdio>
#include <cuda_fp16.h>
#include "ptx.cuh"
using namespace sdb;

__global__ void probe(float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_zero = base;            // [128 rows][128 B] zeros
  uint8_t* a_tile = base + 16384;    // [64 rows][128 B]
  uint8_t* b_tile = base + 32768;    // [16 rows][128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tm;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0u;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int r = threadIdx.x;
    *reinterpret_cast<__half*>(a_tile + r * 128 + ((0 ^ (r & 7)) << 4)) = __float2half((float)(r + 1));
  }
  if (threadIdx.x < 16) {
    const int r = threadIdx.x;
    *reinterpret_cast<__half*>(b_tile + r * 128 + ((0 ^ (r & 7)) << 4)) = __float2half(1.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tm, 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tm;
  if (threadIdx.x == 0) {
    umma_f16(tmem, umma_desc_kmajor_sw128(smem_u32(a_zero)), umma_desc_kmajor_sw128(smem_u32(b_tile)),
             umma_idesc_f16(128, 16), 0u);
    umma_f16(tmem, umma_desc_kmajor_sw128(smem_u32(a_tile)), umma_desc_kmajor_sw128(smem_u32(b_tile)),
             umma_idesc_f16(64, 16), 0u);
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[16];
  tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<1, 128, 64 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static float h[128 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("TMEM lane -> accumulator row (value of column 0 = row + 1; 0 = lane not written by the M=64 MMA); columns equal: ");
  bool eq = true;
  for (int l = 0; l < 128; ++l) for (int j = 1; j < 16; ++j) if (h[l * 16 + j] != h[l * 16]) eq = false;
  printf("%s\n", eq ? "yes" : "NO");
  for (int l = 0; l < 128; ++l) printf("%s%3d:%3.0f", (l % 16) ? " " : "\n", l, h[l * 16]);
  printf("\n");
  return 0;
}
