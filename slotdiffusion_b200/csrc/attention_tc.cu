// Attention core of the UNet transformer blocks on the tensor cores (attention.py:188-205):
//   out = softmax(scale * q k^T) v   per (sample, head), head dim 32, self (Lk = L) and slot cross-attention (Lk = S).
//
// One CTA (128 threads) = one (sample, head, 128-query tile).  Keys are processed in chunks of 64, online softmax:
//   all threads   stage q / k / v: fp32 rows from global -> fp16 hi/lo planes in 128B-swizzled UMMA tiles
//                 Q' = [q_hi | q_lo]      K' = [k_hi | k_hi] , [k_lo | 0]     (K extent 64 per block)
//                 so ONE accumulation chain of 6 tcgen05.mma gives q_hi k_hi + q_lo k_hi + q_hi k_lo  (fp32-faithful):
//                 4 k-steps Q' x K'[0], then the q_hi half of Q' again x K'[1]
//                 V' = [v_hi | v_lo] per key row (MN-major B operand, N = 64)
//   thread 0      S[128 x 64] = Q' K'^T   -> TMEM columns [0,64)
//   all threads   thread <-> query row <-> TMEM lane: chunk max, p = exp(s - m), running sum, rescale of the register
//                 accumulator, P hi/lo planes -> shared memory (A operand of the second product)
//   thread 0      O_c[128 x 64] = P_hi V' + P_lo V'  -> TMEM columns [64,128): cols [0,32) = p v_hi, [32,64) = p v_lo
//   all threads   o += O_c[:, :32] + O_c[:, 32:]
// The output is written in packed GEMM-operand format for to_out.  72 KB of shared memory and 128 TMEM columns per
// CTA: three CTAs per SM overlap each other's load / MMA / softmax phases (the phases of one CTA are serial).
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace sdb {

SDB_DEFINE_PACK_MODE_SETTER(set_pack_mode_attention_tc)

constexpr int AT_THREADS = 128;
constexpr int AT_M = 128;        // queries per CTA
constexpr int AT_KC = 64;        // keys per chunk
constexpr int AT_D = 32;         // head dim
constexpr int AT_TMEM_COLS = 128;

struct AtSmem {
  uint8_t q[AT_M * 128];         // [q_hi | q_lo]
  uint8_t k[2][AT_KC * 128];     // [k_hi | k_hi] , [k_lo | 0]
  uint8_t v[AT_KC * 128];        // [v_hi | v_lo] per key
  uint8_t p[2][AT_M * 128];      // P_hi , P_lo  [128 queries][64 keys]
  uint64_t bar;
  uint32_t tmem_base;
};

// 8 fp32 -> 8 fp16 hi (16 B) and 8 fp16 lo (16 B)
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  auto sp = [](float x, float y, uint32_t& h, uint32_t& l) {
    const __half2 hh = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x - hf.x, y - hf.y);
    h = *reinterpret_cast<const uint32_t*>(&hh);
    l = *reinterpret_cast<const uint32_t*>(&ll);
  };
  sp(a.x, a.y, hi.x, lo.x);
  sp(a.z, a.w, hi.y, lo.y);
  sp(b.x, b.y, hi.z, lo.z);
  sp(b.z, b.w, hi.w, lo.w);
}
__device__ __forceinline__ float ex2_fast(float x) {     // bare MUFU.EX2 (2 ulp; flushes results below 2^-126 to 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte chunk `c` (0..7) of row `r` of a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint4* sw_chunk(uint8_t* tile, int r, int c) {
  return reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
}

__device__ __forceinline__ uint64_t at_desc_k(uint32_t addr) { return umma_desc_kmajor_sw128(addr); }
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t addr) {   // one 64-element MN block, 8-row K groups 1 KB apart
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int G>
__global__ void __launch_bounds__(AT_THREADS, 3)
attention_tc_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                    const float* __restrict__ v, int64_t ldv, __half* __restrict__ out, int Lq, int Lk, int heads,
                    float scale, int64_t plane, int npairs) {
  extern __shared__ uint8_t smem_raw[];
  AtSmem& sm = *reinterpret_cast<AtSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  // G = 1: CTA (q-tile, head, sample).  G = 2 / 4 (small Lq, Lk): G (sample, head) pairs share the tile -- pair g owns
  // query rows [RS g, RS g + Lq) and key slots [KS g, KS g + Lk) of the single key chunk; the off-diagonal blocks of
  // S are masked in the softmax, so P is block-diagonal and P V' is exact.
  constexpr int RS = AT_M / G, KS = AT_KC / G;
  const int pair0 = G == 1 ? (int)(blockIdx.z * heads + blockIdx.y) : (int)blockIdx.x * G;
  const int q0 = G == 1 ? blockIdx.x * AT_M : 0;
  const int C = heads * AT_D;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&sm.tmem_base, AT_TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();        // PDL (common.cuh): barrier init and TMEM allocation above may overlap the previous kernel's tail
  pdl_trigger();
  const int pmode = g_pack_mode;   // operand format of the consumer GEMM: read once (common.cuh)
  // ---- loads of the Q tile and of the first K / V chunk are issued together (one exposed global latency)
  // K / V chunk staging: thread <-> 2 x (key, 8-channel quarter)
  float4 kreg[2][2], vreg[2][2];
  auto load_kv = [&](int j0) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = tid + AT_THREADS * it, kr = item >> 2, qt = item & 3;
      const int g = kr / KS, lk = j0 + kr - g * KS, pi = pair0 + g;
      if (lk < Lk && pi < npairs) {
        const int64_t b = G == 1 ? (int64_t)blockIdx.z : (int64_t)(pi / heads);
        const int h = G == 1 ? (int)blockIdx.y : pi - (int)b * heads;
        const float4* ks = reinterpret_cast<const float4*>(k + (b * Lk + lk) * ldk + h * AT_D + qt * 8);
        const float4* vs = reinterpret_cast<const float4*>(v + (b * Lk + lk) * ldv + h * AT_D + qt * 8);
        kreg[it][0] = ks[0]; kreg[it][1] = ks[1];
        vreg[it][0] = vs[0]; vreg[it][1] = vs[1];
      } else {
        kreg[it][0] = kreg[it][1] = vreg[it][0] = vreg[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero rows: masked keys
      }
    }
  };
  auto store_kv = [&]() {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = tid + AT_THREADS * it, kr = item >> 2, qt = item & 3;
      uint4 khi, klo, vhi, vlo;
      split8(kreg[it][0], kreg[it][1], khi, klo);
      split8(vreg[it][0], vreg[it][1], vhi, vlo);
      *sw_chunk(sm.k[0], kr, qt) = khi;
      *sw_chunk(sm.k[0], kr, 4 + qt) = khi;
      *sw_chunk(sm.k[1], kr, qt) = klo;
      *sw_chunk(sm.v, kr, qt) = vhi;
      *sw_chunk(sm.v, kr, 4 + qt) = vlo;
    }
  };
  {
    float4 qreg[4][2];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int item = tid + AT_THREADS * it, qr = item >> 2, qt = item & 3;
      const int g = qr / RS, lr = q0 + qr - g * RS, pi = pair0 + g;
      if (lr < Lq && pi < npairs) {
        const int64_t b = G == 1 ? (int64_t)blockIdx.z : (int64_t)(pi / heads);
        const int h = G == 1 ? (int)blockIdx.y : pi - (int)b * heads;
        const float4* src = reinterpret_cast<const float4*>(q + (b * Lq + lr) * ldq + h * AT_D + qt * 8);
        qreg[it][0] = src[0]; qreg[it][1] = src[1];
      } else {
        qreg[it][0] = qreg[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    load_kv(0);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int item = tid + AT_THREADS * it, qr = item >> 2, qt = item & 3;
      float4 a = qreg[it][0], c = qreg[it][1];
      // logits in units of log2: softmax(s) = 2^(s' - max s') with s' = s log2(e) -- the exponentials below are bare ex2
      const float sc2 = scale * 1.4426950408889634f;
      a.x *= sc2; a.y *= sc2; a.z *= sc2; a.w *= sc2;
      c.x *= sc2; c.y *= sc2; c.z *= sc2; c.w *= sc2;
      uint4 hi, lo;
      split8(a, c, hi, lo);
      *sw_chunk(sm.q, qr, qt) = hi;
      *sw_chunk(sm.q, qr, 4 + qt) = lo;
    }
  }
  // upper halves of the second K block stay zero for the whole kernel
  for (int i = tid; i < AT_KC * 4; i += AT_THREADS) *sw_chunk(sm.k[1], i >> 2, 4 + (i & 3)) = zero4;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);

  // row state: thread <-> query row
  const int r = tid;
  const int rg = r / RS;                              // pair of this row
  const int jlo = rg * KS;                            // first key slot of that pair
  const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
  float o[AT_D];
#pragma unroll
  for (int i = 0; i < AT_D; ++i) o[i] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;
  constexpr uint32_t phase = 0;   // two commits per chunk: the barrier parity is back at 0 at every chunk start

  const uint32_t idesc_s = umma_idesc_f16(AT_M, AT_KC);                 // A, B K-major
  const uint32_t idesc_o = umma_idesc_f16(AT_M, 64) | (1u << 16);       // A K-major (P), B MN-major (V')
  const uint32_t q_a = smem_u32(sm.q), k_a = smem_u32(sm.k[0]), v_a = smem_u32(sm.v), p_a = smem_u32(sm.p[0]);

  for (int j0 = 0; j0 < Lk; j0 += AT_KC) {
    const int jhi = jlo + min(KS, Lk - j0);          // valid key slots of this row: [jlo, jhi)
#if SDB_AT_NO_PREFETCH
    if (j0 > 0) load_kv(j0);
    store_kv();
#else
    store_kv();
    if (j0 + AT_KC < Lk) load_kv(j0 + AT_KC);   // next chunk's rows travel while this chunk is multiplied and normalised
#endif
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- S = Q' K'^T
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) {
        const uint32_t blk = ks >> 2, kk = ks & 3;       // block 0: 4 k-steps, block 1: 2 k-steps
        const uint64_t da = at_desc_k(q_a) + uint64_t((kk * 32) >> 4);          // block 1 re-reads the q_hi half
        const uint64_t db = at_desc_k(k_a + blk * (AT_KC * 128)) + uint64_t((kk * 32) >> 4);
        umma_f16(tmem, da, db, idesc_s, ks ? 1u : 0u);
      }
      umma_commit(&sm.bar);
    }
    __syncwarp();
    {
      mbar_wait(&sm.bar, phase);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(lane_addr, s0);
      tmem_ld_32x32(lane_addr + 32, s1);
      tmem_ld_wait();
      // full chunk of a single-pair tile (the common case of the L = 256 / 64 self-attention cores): no key masks
      const bool full = (G == 1) && (jhi - jlo == AT_KC);
      float cmax = -INFINITY;
      if (full) {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // four independent max chains
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          m0 = fmaxf(m0, __uint_as_float(s0[j]));
          m1 = fmaxf(m1, __uint_as_float(s0[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(s1[j]));
          m3 = fmaxf(m3, __uint_as_float(s1[j + 1]));
        }
        cmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j >= jlo && j < jhi) cmax = fmaxf(cmax, __uint_as_float(s0[j]));
          if (32 + j >= jlo && 32 + j < jhi) cmax = fmaxf(cmax, __uint_as_float(s1[j]));
        }
      }
      const float mnew = fmaxf(mrun, cmax);
      const float corr = ex2_fast(mrun - mnew);             // 2^(-inf) = 0 on the first chunk
      mrun = mnew;
      float psum = 0.f;
      auto probs = [&](auto fullc) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {                     // 8 keys per 16-byte chunk
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = c * 8 + e;
            const float sv = __uint_as_float(j < 32 ? s0[j] : s1[j - 32]);
            if constexpr (decltype(fullc)::value) pv[e] = ex2_fast(sv - mnew);
            else pv[e] = (j >= jlo && j < jhi) ? ex2_fast(sv - mnew) : 0.f;
          }
          psum += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
          uint4 hi, lo;
          split8(make_float4(pv[0], pv[1], pv[2], pv[3]), make_float4(pv[4], pv[5], pv[6], pv[7]), hi, lo);
          *sw_chunk(sm.p[0], r, c) = hi;
          *sw_chunk(sm.p[1], r, c) = lo;
        }
      };
      if (full) probs(std::true_type{});
      else probs(std::false_type{});
      lrun = lrun * corr + psum;
#pragma unroll
      for (int i = 0; i < AT_D; ++i) o[i] *= corr;
      fence_proxy_async_smem();
      tc_fence_before();
    }
    __syncthreads();
    // ---- O_c = P V'
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
        for (int kk = 0; kk < AT_KC / 16; ++kk) {
          const uint64_t da = at_desc_k(p_a + pl * (AT_M * 128)) + uint64_t((kk * 32) >> 4);
          const uint64_t db = at_desc_mn(v_a) + uint64_t((kk * 16 * 128) >> 4);
          umma_f16(tmem + 64, da, db, idesc_o, (pl | kk) ? 1u : 0u);
        }
      }
      umma_commit(&sm.bar);
    }
    __syncwarp();
    {
      mbar_wait(&sm.bar, phase ^ 1);
      tc_fence_after();
      uint32_t c0[32], c1[32];
      tmem_ld_32x32(lane_addr + 64, c0);
      tmem_ld_32x32(lane_addr + 96, c1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < AT_D; ++i) o[i] += __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
      tc_fence_before();
    }
    // both commits of this chunk consumed: the barrier is back at the starting phase parity
    __syncthreads();   // K' / V' / P and the TMEM columns may be overwritten by the next chunk
  }

  {
    const int lr = q0 + r - rg * RS, pi = pair0 + rg;
    if (lr < Lq && pi < npairs) {
      const int64_t b = G == 1 ? (int64_t)blockIdx.z : (int64_t)(pi / heads);
      const int h = G == 1 ? (int)blockIdx.y : pi - (int)b * heads;
      const float inv = 1.f / lrun;
      const int64_t ob = (b * Lq + lr) * C + h * AT_D;
      if (pmode == SDB_FMT_F8C) {     // operand format of the consumer GEMM (to_out projection), common.cuh
#pragma unroll
        for (int i = 0; i < AT_D; i += 4)
          store_split4(out, out + plane, ob + i, make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv), pmode);
      } else {
#pragma unroll
        for (int i = 0; i < AT_D; i += 8) {
          uint4 hi, lo;
          split8(make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv),
                 make_float4(o[i + 4] * inv, o[i + 5] * inv, o[i + 6] * inv, o[i + 7] * inv), hi, lo);
          *reinterpret_cast<uint4*>(out + ob + i) = hi;
          *reinterpret_cast<uint4*>(out + plane + ob + i) = lo;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, AT_TMEM_COLS);
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_attention_tc_supported(int64_t heads, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv) {
  return d == 32 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && heads > 0;
}

extern "C" int sdb_attention_tc(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                void* out, int64_t B, int64_t Lq, int64_t Lk, int heads, int d, float scale,
                                void* stream) {
  SDB_REQUIRE(q && k && v && out && B > 0 && Lq > 0 && Lk > 0 && heads > 0, "sdb_attention_tc: bad args");
  SDB_REQUIRE(d == AT_D, "sdb_attention_tc: head dim %d unsupported (32)", d);
  SDB_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "sdb_attention_tc: row strides must be multiples of 4");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                reinterpret_cast<uintptr_t>(out)) & 15) == 0 && (heads * d) % 8 == 0,
              "sdb_attention_tc: operands must be 16-byte aligned");
  SDB_REQUIRE(B <= 65535 && heads <= 65535 && Lq < (1 << 30) && Lk < (1 << 30) && B * heads < (1 << 30),
              "sdb_attention_tc: grid too large");
  const size_t smem = sizeof(AtSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    SDB_CHECK(cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  // small problems: several (sample, head) pairs per CTA (block-diagonal masking), else one CTA per 128-query tile
  const int npairs = (int)(B * heads);
  const int64_t plane = B * Lq * heads * d;
  cudaStream_t st = as_stream(stream);
  __half* o = (__half*)out;
  if (Lq <= 32 && Lk <= 16)
    (void)launch_k(attention_tc_kernel<4>, dim3((unsigned)cdiv(npairs, 4)), dim3(AT_THREADS), smem, st, q, ldq, k, ldk, v, ldv,
                   o, (int)Lq, (int)Lk, heads, scale, plane, npairs);
  else if (Lq <= 64 && Lk <= 32)
    (void)launch_k(attention_tc_kernel<2>, dim3((unsigned)cdiv(npairs, 2)), dim3(AT_THREADS), smem, st, q, ldq, k, ldk, v, ldv,
                   o, (int)Lq, (int)Lk, heads, scale, plane, npairs);
  else
    (void)launch_k(attention_tc_kernel<1>, dim3((unsigned)cdiv(Lq, AT_M), (unsigned)heads, (unsigned)B), dim3(AT_THREADS), smem,
                   st, q, ldq, k, ldk, v, ldv, o, (int)Lq, (int)Lk, heads, scale, plane, npairs);
  SDB_LAUNCH_CHECK();
  return 0;
}
