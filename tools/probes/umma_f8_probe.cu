// Micro-probe for DESIGN.md section 8 item 0: does "fp16 main term + two e4m3 correction terms" (kind::f16 +
// kind::f8f6f4, second accumulator) issue at 2/3 of the cost of today's three fp16 passes, from real shared-memory
// operand tiles (128B-swizzled, distinct tiles per plane so the operand read traffic is the real one)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slotdiffusion_b200/csrc tools/probes/umma_f8_probe.cu -o tools/probes/umma_f8_probe.bin
//   ./tools/probes/umma_f8_probe.bin
// One "unit" = K = 128 of one 128 x N output tile (cta_group::1):
//   scheme 0 (today)   8 k-steps x {A_h W_h, A_l W_h, A_h W_l} = 24 kind::f16 MMAs (K = 16 each), one accumulator
//   scheme 1 (planned) 8 kind::f16 MMAs (A_h W_h) into D1  +  4 k-steps (K = 32) x {A_l8 W_h8, A_h8 W_l8} = 8 kind::f8f6f4
//                      MMAs into D2                                                            = 16 MMAs
// All operands are 1.0 (fp16 0x3c00, e4m3 0x38), so after `units` units D1 = 128 * units (x3 for scheme 0) and
// D2 = 256 * units: the probe also checks that the f8f6f4 descriptors / K advance are right.
#include <cstdio>
#include <cuda_fp16.h>
#include "ptx.cuh"
using namespace sdb;

__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int SMEM_BYTES = 193 * 1024;

__global__ void probe(int N, int units, int scheme, long long* out, float* vals) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tm;
  const uint32_t a_blk = 128 * 128;        // one 64-wide fp16 (or 128-wide fp8) k-block of A: 128 rows x 128 B
  const uint32_t b_blk = (uint32_t)N * 128;
  // scheme 0: [A_h 2 blk][A_l 2 blk][B_h 2 blk][B_l 2 blk]   (fp16)
  // scheme 1: [A_h 2 blk][B_h 2 blk] (fp16) [A_h8][A_l8][B_h8][B_l8] (e4m3, 1 blk each)
  const uint32_t f16_bytes = scheme == 0 ? 4 * a_blk + 4 * b_blk : 2 * a_blk + 2 * b_blk;
  const uint32_t f8_bytes = scheme == 0 ? 0 : 2 * a_blk + 2 * b_blk;
  for (uint32_t i = threadIdx.x; i < f16_bytes / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;           // fp16 1.0
  for (uint32_t i = threadIdx.x; i < f8_bytes / 4; i += blockDim.x) ((uint32_t*)(base + f16_bytes))[i] = 0x38383838u;   // e4m3 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tm, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tm, 0);
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);        // formats 0/0: F16 for kind::f16, E4M3 for kind::f8f6f4
    const uint32_t s0 = smem_u32(base);
    const uint32_t d1 = tmem, d2 = tmem + 256;
    long long t0 = clock64();
    for (int u = 0; u < units; ++u) {
      const uint32_t acc0 = u ? 1u : 0u;
      if (scheme == 0) {
        const uint32_t ah = s0, al = s0 + 2 * a_blk, bh = s0 + 4 * a_blk, bl = bh + 2 * b_blk;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            const uint64_t dah = umma_desc_kmajor_sw128(ah + kb * a_blk) + adv, dal = umma_desc_kmajor_sw128(al + kb * a_blk) + adv;
            const uint64_t dbh = umma_desc_kmajor_sw128(bh + kb * b_blk) + adv, dbl = umma_desc_kmajor_sw128(bl + kb * b_blk) + adv;
            umma_f16(d1, dah, dbh, idesc, (kb | k) ? 1u : acc0);
            umma_f16(d1, dal, dbh, idesc, 1u);
            umma_f16(d1, dah, dbl, idesc, 1u);
          }
        }
      } else {
        const uint32_t ah = s0, bh = s0 + 2 * a_blk;
        const uint32_t ah8 = s0 + f16_bytes, al8 = ah8 + a_blk, bh8 = al8 + a_blk, bl8 = bh8 + b_blk;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_f16(d1, umma_desc_kmajor_sw128(ah + kb * a_blk) + adv, umma_desc_kmajor_sw128(bh + kb * b_blk) + adv, idesc,
                     (kb | k) ? 1u : acc0);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {       // K = 32 e4m3 elements = 32 B per step: same descriptor advance
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_f8(d2, umma_desc_kmajor_sw128(al8) + adv, umma_desc_kmajor_sw128(bh8) + adv, idesc, k ? 1u : acc0);
          umma_f8(d2, umma_desc_kmajor_sw128(ah8) + adv, umma_desc_kmajor_sw128(bl8) + adv, idesc, 1u);
        }
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem, r);
    tmem_ld_wait();
    if (threadIdx.x == 0) vals[0] = __uint_as_float(r[0]);
    tmem_ld_32x32(tmem + 256, r);
    tmem_ld_wait();
    if (threadIdx.x == 0) vals[1] = __uint_as_float(r[0]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- layout check: e4m3 operand tiles written by threads with the swizzle formulas a GEMM producer / TMA would use ----
// sw = 128: 128-byte rows (K = 128 e4m3), 8-row atoms of 1024 B;  sw = 64: 64-byte rows (K = 64), 8-row atoms of 512 B.
// A[r][k] = (r + 2k) % 7 - 3, B[n][k] = (3n + k) % 5 - 2 (exact in e4m3), D = A B^T must match the host sum exactly.
__device__ __forceinline__ uint8_t e4m3_small(int v) {   // -3..3
  const uint8_t mag[4] = {0x00, 0x38, 0x40, 0x44};
  return (uint8_t)(mag[v < 0 ? -v : v] | (v < 0 ? 0x80 : 0));
}
__device__ __forceinline__ uint64_t desc_k_sw(uint32_t addr, int sw) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                                   // LBO: unused for swizzled K-major
  d |= (uint64_t)((sw == 128 ? 1024 : 512) >> 4) << 32;     // SBO: 8 rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw == 128 ? 2 : 4) << 61;                 // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return d;
}
__global__ void layout_check(int sw, float* dout) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tm;
  const int K = sw, N = 64;
  uint8_t* a = base;
  uint8_t* b = base + 128 * sw;
  for (int i = threadIdx.x; i < (128 + N) * K; i += blockDim.x) {
    const bool isb = i >= 128 * K;
    const int e = isb ? i - 128 * K : i, r = e / K, k = e % K;
    const int v = isb ? (3 * r + k) % 5 - 2 : (r + 2 * k) % 7 - 3;
    const int chunk = k / 16, x = sw == 128 ? (r % 8) : ((r % 8) / 2);
    const int off = (r / 8) * (8 * sw) + (r % 8) * sw + ((chunk ^ x) * 16) + k % 16;
    (isb ? b : a)[off] = e4m3_small(v);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tm, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tm, 0);
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    for (int k = 0; k < K / 32; ++k)
      umma_f8(tmem, desc_k_sw(smem_u32(a), sw) + (uint64_t)((k * 32) >> 4), desc_k_sw(smem_u32(b), sw) + (uint64_t)((k * 32) >> 4),
              idesc, k ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) dout[threadIdx.x * N + c0 + c] = __uint_as_float(r[c]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static int run_layout_check(int sw) {
  float* dd; cudaMalloc(&dd, 128 * 64 * 4);
  cudaFuncSetAttribute(layout_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  layout_check<<<1, 128, 64 * 1024>>>(sw, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("layout check sw %d: %s\n", sw, cudaGetErrorString(e)); return 1; }
  static float h[128 * 64];
  cudaMemcpy(h, dd, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 64; ++n) {
      float want = 0.f;
      for (int k = 0; k < sw; ++k) want += (float)((r + 2 * k) % 7 - 3) * (float)((3 * n + k) % 5 - 2);
      if (h[r * 64 + n] != want && bad++ < 4) printf("  sw %d D[%d][%d] = %g, expected %g\n", sw, r, n, h[r * 64 + n], want);
    }
  printf("layout check e4m3 K-major SWIZZLE_%dB (K = %d, 128 x 64): %s (%d mismatches)\n", sw, sw, bad ? "FAILED" : "ok", bad);
  return 0;
}

int main() {
  run_layout_check(128);
  run_layout_check(64);
  long long* d; float* v;
  cudaMalloc(&d, 16); cudaMalloc(&v, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  for (int N : {128, 256})
    for (int units : {1, 16, 64}) {
      double cyc[2] = {0, 0};
      for (int scheme : {0, 1}) {
        long long h[2]; float hv[2];
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<1, 128, SMEM_BYTES>>>(N, units, scheme, d, v);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("N %d units %d scheme %d: %s\n", N, units, scheme, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(hv, v, 8, cudaMemcpyDeviceToHost);
        cyc[scheme] = (double)h[1] / units;
        const float want1 = scheme == 0 ? 384.f * units : 128.f * units, want2 = scheme == 0 ? 0.f : 256.f * units;
        printf("N %3d units %2d scheme %d: %8.1f cycles per K=128 unit (issue %lld, done %lld)  D1 = %.0f (expect %.0f)%s", N, units,
               scheme, cyc[scheme], h[0], h[1], hv[0], want1, hv[0] == want1 ? "" : "  MISMATCH");
        if (scheme == 1) printf("  D2 = %.0f (expect %.0f)%s", hv[1], want2, hv[1] == want2 ? "" : "  MISMATCH");
        printf("\n");
      }
      printf("   -> planned / today = %.3f (ideal 0.667)\n", cyc[1] / cyc[0]);
    }
  return 0;
}
