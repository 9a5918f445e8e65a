// Slot Attention inner loop, tensor-core form (slot_attention.py:67-91, sa_diffusion.py:43-56) -- inference path.
//
// The k / v projections are never materialised.  With n = (x - mean) * rstd (LayerNorm without its affine part):
//   logits[n,s] = scale * k[n] . q[s]          = n[n] . qa[s,:Din] + qa[s,Din]          (qa = LN_q(slots) W_qa^T, host-folded
//                                                                                        W_qa = scale * diag(gamma) Wk^T Wq | beta row)
//   updates[s]  = (sum_n a[n,s] v[n]) / sum_n a = (gamma * U[s] + beta) Wv^T,  U[s] = (sum_n a[n,s] n[n]) / sum_n a[n,s]
// so one pass over the RAW features per iteration does LayerNorm statistics, both N x S contractions and the softmax.
//
//   warp TMA     producer: cp.async.bulk of 32-token fp32 chunks straight into the operand slots (mbarrier completion),
//                started before the rest of the prologue; L2 prefetch of the tile after next
//   warps 4..4+CW-1  converters: LayerNorm (8 lanes per token, warp-shuffle reductions) and fp16 hi/lo split IN PLACE:
//                the 6 KB of eight fp32 token rows become the 6 one-KB 128B-swizzled UMMA atoms (2 planes x Din/64
//                blocks) of the same eight tokens.  ONE operand tile [128 tokens][Din] serves BOTH contractions
//                (K-major for the logits, MN-major for the weighted sum); tiles are double-buffered.
//                (A register-staged variant -- ld.global one tile ahead, no fp32 copy in shared memory -- was measured
//                too and lost: see profiles/README.md.)
//   warp G1      MMA issuer (one thread): logits[128 x SP] = X Q^T                      tcgen05.mma kind::f16
//   warp G2      MMA issuer (one thread): U^T[Din x SP] += X^T A
//                fp32-faithful products hi*hi + lo*hi + hi*lo (+ lo*lo) with the hi and lo planes STACKED inside one
//                instruction (along N for the small operand, along M for the features in the second product): small
//                MMAs are bound by their ~50-cycle issue cost (tools/probes/umma_probe.cu), not by the tensor pipe
//   warps 0..3   softmax over slots (thread <-> token <-> TMEM lane), seg-mask store, a = softmax + eps as the B operand
//                of the second contraction, column sums by warp shuffles; finally drain U from TMEM
//
// Grid (chunks, B): a CTA owns a contiguous range of 128-token tiles of one sample; partial sums are combined by
// slot_attend_fused_finalize_kernel (deterministic, no atomics).
#include "common.cuh"
#include "ptx.cuh"

namespace sdb {

constexpr int SF_TILE = 128;          // tokens per tile (UMMA M of the logits product, K extent of the update product)
constexpr int SF_GROUPS = SF_TILE / 8;
constexpr int SF_CT = 32;             // tokens per TMA chunk
constexpr int SF_CPT = SF_TILE / SF_CT;   // chunks per tile
constexpr int SF_SOFT_WARPS = 4;
// converter warps CW: 16 (one 8-token group per warp and tile) when the softmax warps fit 88 registers (SP = 16),
// else 8.  Warps: 0..3 softmax, 4..4+CW-1 converters, then the two MMA issuers and the TMA producer.
constexpr int SF_TMEM_COLS = 512;     // logits set s at [64 s, 64 s + 2 SP) | U block kb at [128 + 2 SP kb, ... + 2 SP)
constexpr float SF_ASCALE = 4096.f;   // a = softmax + eps is scaled before the fp16 split (keeps 1e-6 out of fp16 subnormals)

struct SfCtl {
  uint64_t sfull[2 * SF_CPT];           // TMA chunk landed
  uint64_t xfull[2], xempty[2];         // operand tile converted / consumed by the MMAs
  uint64_t lfull[2], afull[2];          // logits in TMEM / a operand written
  uint64_t ufull;
  uint32_t tmem_base;
  uint32_t pad;
  float cb[32];                         // logit bias per slot (beta row of the folded projection)
  float cs_scr[SF_SOFT_WARPS][32];      // per-warp column sums
};

template <int DIN, int SP>
struct SfCfg {
  static constexpr int NKB = DIN / 64;                       // 64-channel blocks
  static constexpr int GROUP_BYTES = 2 * NKB * 1024;         // 8 tokens: fp32 rows == 2 planes x NKB swizzle atoms of 1 KB
  static constexpr int CHUNK_BYTES = (SF_CT / 8) * GROUP_BYTES;
  static constexpr int TILE_BYTES = SF_GROUPS * GROUP_BYTES;
  static_assert(GROUP_BYTES == 8 * DIN * 4, "in-place conversion needs equal fp32 and fp16x2 footprints");
  static constexpr int QBYTES = NKB * 2 * SP * 128;          // per k-block: [q_hi rows | q_lo rows] x 128 B
  static constexpr int ABYTES = (SF_TILE / 64) * 2 * SP * 128;   // per set: 2 token blocks x [a_hi rows | a_lo rows] x 128 B
  static constexpr int CTL = 1024;
  static constexpr int NSETS = (1024 + 2 * TILE_BYTES + QBYTES + 2 * ABYTES + CTL <= 227 * 1024) ? 2 : 1;
  static constexpr int SMEM = 1024 + NSETS * TILE_BYTES + QBYTES + NSETS * ABYTES + CTL;
  static_assert(SMEM <= 227 * 1024, "slot_attend_fused: tile does not fit shared memory");
  static_assert(sizeof(SfCtl) <= CTL, "control block too large");
  static_assert(NKB * 64 * SP * 4 <= TILE_BYTES, "epilogue scratch must fit an operand tile");
  static_assert(128 + NKB * 2 * SP <= SF_TMEM_COLS, "TMEM columns");
};

// two fp32 -> (hi, lo) fp16 pairs; inputs are LayerNorm outputs / probabilities (bounded, no saturation needed)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// SWIZZLE_128B descriptors with explicit strides.  K-major: sbo = bytes between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// MN-major: lbo = bytes between 64-element MN blocks, sbo = bytes between groups of 8 K rows.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// 2*SP accumulator columns of this thread's TMEM lane, folded: out[s] = col[s] + col[SP + s]
template <int SP>
__device__ __forceinline__ void tmem_ld_folded(uint32_t taddr, float (&out)[SP]) {
  if constexpr (SP == 16) {
    uint32_t r[32];
    tmem_ld_32x32(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int s = 0; s < 16; ++s) out[s] = __uint_as_float(r[s]) + __uint_as_float(r[16 + s]);
  } else {
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(taddr, r0);
    tmem_ld_32x32(taddr + 32, r1);
    tmem_ld_wait();
#pragma unroll
    for (int s = 0; s < 32; ++s) out[s] = __uint_as_float(r0[s]) + __uint_as_float(r1[s]);
  }
}

// Two changes derived offline from the ncu source page of this kernel (profiles/README.md section 12).  Run on the B200 in
// round 2 (profiles/README.md section 13: parity tests green, attend 38.9 -> 36.9 us at B=64, 96.3 -> 92.2 us at B=256) and
// the default since; -DSDB_SF_EXPERIMENTAL=0 builds the earlier form:
//   * bounded-wait loops read the clock once per 256 / 4096 polls instead of on each one (poll loops were 40 % of the
//     executed warp instructions: 18 -> 8 instructions per poll);
//   * the converters zero rows beyond the end of a sample only in the partial last group (48 FSELs of 611 instructions).
#ifndef SDB_SF_EXPERIMENTAL
#define SDB_SF_EXPERIMENTAL 1
#endif
#if SDB_SF_EXPERIMENTAL
constexpr uint32_t SF_SPIN_CLOCK_MASK = 0xfffu;   // read the clock when (polls & mask) == 0
constexpr uint32_t SF_WAIT_CLOCK_MASK = 0xffu;

// latency-critical single-thread wait (MMA issuers): non-blocking test_wait in a tight loop, bounded
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  long long t0 = 0;
  uint32_t polls = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++polls & SF_SPIN_CLOCK_MASK) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        printf("sdb200: mbarrier spin timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
        __trap();
      }
    }
  }
}

// one lane polls, the warp follows (keeps hundreds of threads from hammering the barrier)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) {
    long long t0 = 0;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {     // try_wait itself blocks for a hardware-defined interval (~200 cycles)
      if ((++polls & SF_WAIT_CLOCK_MASK) == 0) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) {
          printf("sdb200: mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
          __trap();
        }
      }
    }
  }
  __syncwarp();
  mbar_wait(bar, parity);   // already complete: one try_wait per thread = its own acquire
}

#else   // the GPU-verified wait loops, verbatim
// latency-critical single-thread wait (MMA issuers): non-blocking test_wait in a tight loop, bounded
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000LL) {
      printf("sdb200: mbarrier spin timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// one lane polls, the warp follows (keeps hundreds of threads from hammering the barrier)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) {
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {     // try_wait itself blocks for a hardware-defined interval
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        printf("sdb200: mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
        __trap();
      }
    }
  }
  __syncwarp();
  mbar_wait(bar, parity);   // already complete: one try_wait per thread = its own acquire
}

#endif

__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}

template <int DIN, int SP, int CW>
__global__ void __launch_bounds__(32 * (SF_SOFT_WARPS + CW + 3), 1)
slot_attend_fused_kernel(const float* __restrict__ x, const float* __restrict__ qa, int ldq,
                         float* __restrict__ seg_mask, float* __restrict__ part_upd, float* __restrict__ part_cs,
                         int N, int S, int chunks, float ln_eps, float eps, long long* __restrict__ dbg) {
  using C = SfCfg<DIN, SP>;
  // optional timeline of CTA (0,0) for tools/sa_timeline.py: dbg[role * 64 + tile * 4 + k] = cycles since kernel entry
  const long long t_entry = clock64();
#define SF_T(role, tile, k)                                                                     \
  do {                                                                                          \
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && (tile) < 16) dbg[(role) * 64 + (tile) * 4 + (k)] = clock64() - t_entry; \
  } while (0)
  constexpr int NKB = C::NKB;
  constexpr int NSETS = C::NSETS;
  constexpr int GB = C::GROUP_BYTES;
  constexpr int SF_G1_WARP = SF_SOFT_WARPS + CW;
  constexpr int SF_G2_WARP = SF_G1_WARP + 1;
  constexpr int SF_TMA_WARP = SF_G2_WARP + 1;
  constexpr int SF_THREADS = 32 * (SF_TMA_WARP + 1);
  constexpr int NV = DIN / 32;                      // float4 per lane per token
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xop = base;                              // NSETS tiles: [16 token groups][2 planes][NKB][8 tokens][128 B]
  uint8_t* qop = xop + NSETS * C::TILE_BYTES;       // [NKB][q_hi SP rows | q_lo SP rows][128 B]
  uint8_t* aop = qop + C::QBYTES;                   // NSETS x [2 token blocks][a_hi SP rows | a_lo SP rows][128 B]
  SfCtl& ctl = *reinterpret_cast<SfCtl*>(aop + NSETS * C::ABYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x;
  const long long b = blockIdx.y;
  const int tiles_total = (N + SF_TILE - 1) / SF_TILE;
  const int tpc = (tiles_total + chunks - 1) / chunks;
  const int n_begin = chunk * tpc * SF_TILE;
  const int n_end = min(N, n_begin + tpc * SF_TILE);
  const int ntok = n_end - n_begin;
  float* my_upd = part_upd + ((b * chunks + chunk) * S) * DIN;
  float* my_cs = part_cs + (b * chunks + chunk) * S;
  if (ntok <= 0) {   // empty chunk (uniform per CTA): contribute zeros
    for (int i = tid; i < S * DIN; i += SF_THREADS) my_upd[i] = 0.f;
    if (tid < S) my_cs[tid] = 0.f;
    return;
  }
  const int ntiles = (ntok + SF_TILE - 1) / SF_TILE;

  const int nchunks = (ntok + SF_CT - 1) / SF_CT;
  const bool is_conv = warp >= SF_SOFT_WARPS && warp < SF_G1_WARP;
  const int cw = warp - SF_SOFT_WARPS, sub = lane >> 3, j = lane & 7;
  // token of (round q, lane) within its 8-token group: a half-warp's two tokens land in disjoint halves of the bank space
  auto tok_in_group = [&](int q) { return 4 * (sub & 1) + (sub >> 1) + 2 * q; };

  // ------------------------------------------------------------------ prologue
  // TMA chunk loads of tile i into its set (+ L2 prefetch of the tile that will follow it into the same set)
  auto issue_tile = [&](int i) {
    const int set = i % NSETS;
    for (int cc = 0; cc < SF_CPT; ++cc) {
      const int c = i * SF_CPT + cc;
      if (c >= nchunks) break;
      const int rows = min(SF_CT, ntok - c * SF_CT);
      const uint32_t bytes = (uint32_t)rows * DIN * 4;
      uint64_t* bar = &ctl.sfull[set * SF_CPT + cc];
      mbar_arrive_expect_tx(bar, bytes);
      bulk_load(xop + set * C::TILE_BYTES + cc * C::CHUNK_BYTES, x + (b * N + n_begin + c * SF_CT) * DIN, bytes, bar);
    }
    const int nxt = (i + NSETS) * SF_TILE;
    if (nxt < ntok) l2_prefetch(x + (b * N + n_begin + nxt) * DIN, (uint32_t)min(SF_TILE, ntok - nxt) * DIN * 4);
  };
  if (warp == SF_TMA_WARP && lane == 0) {
    for (int i = 0; i < 2 * SF_CPT; ++i) mbar_init(&ctl.sfull[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl.xfull[i], CW);
      mbar_init(&ctl.xempty[i], 1);
      mbar_init(&ctl.lfull[i], 1);
      mbar_init(&ctl.afull[i], SF_SOFT_WARPS);
    }
    mbar_init(&ctl.ufull, 1);
    fence_mbar_init();
    for (int i = 0; i < NSETS && i < ntiles; ++i) issue_tile(i);   // the feature stream starts before the rest of the prologue
  }
  if (warp == SF_G1_WARP) {
    tmem_alloc(&ctl.tmem_base, SF_TMEM_COLS);
    tmem_relinquish();
  }
  // slot-side operand of the logits product: qa[s, 0:DIN] split into fp16 planes stacked along N, K-major swizzled rows
  for (int i = tid; i < SP * (DIN / 4); i += SF_THREADS) {
    const int s = i / (DIN / 4), c = (i % (DIN / 4)) * 4;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < S) w = *reinterpret_cast<const float4*>(qa + (b * S + s) * ldq + c);
    uint2 hi, lo;
    split2(w.x, w.y, hi.x, lo.x);
    split2(w.z, w.w, hi.y, lo.y);
    const int kb = c >> 6, cc = c & 63;
    const uint32_t off = kb * (2 * SP * 128) + (((cc >> 3) ^ (s & 7)) << 4) + (cc & 7) * 2;   // SP % 8 == 0: lo row has the same swizzle phase
    *reinterpret_cast<uint2*>(qop + off + s * 128) = hi;
    *reinterpret_cast<uint2*>(qop + off + (SP + s) * 128) = lo;
  }
  if (tid < 32) ctl.cb[tid] = (tid < S) ? qa[(b * S + tid) * ldq + DIN] : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ctl.tmem_base, 0);   // warp-uniform by construction
  if (tid == 0) SF_T(5, 0, 2);   // prologue done

  if (warp == SF_TMA_WARP) {
    // ================================================================ TMA producer
    if (lane == 0) {
      for (int i = NSETS; i < ntiles; ++i) {
        const int set = i % NSETS;
        mbar_wait(&ctl.xempty[set], ((i / NSETS) & 1) ^ 1);    // the MMAs of the tile that used this set have retired
        issue_tile(i);
      }
    }
    __syncwarp();
  } else if (warp == SF_G1_WARP) {
    // ================================================================ MMA issuer 1: logits product
    // (two issuer warps with plain counted loops: every operand of tcgen05.mma stays warp-uniform, so the per-MMA
    //  instruction stream is a handful of uniform-datapath adds -- small MMAs are bound by their issue overhead)
    if (lane == 0) {
      const uint32_t id1 = umma_idesc_f16(SF_TILE, 2 * SP);   // A, B K-major
      const uint32_t x0 = smem_u32(xop), q0 = smem_u32(qop);
      for (int i = 0; i < ntiles; ++i) {
        const int set = i % NSETS;
        // logits[128 tokens x 2SP] = X[128 x DIN] * [Q_hi ; Q_lo]^T for X = hi plane, then lo plane:
        //   cols [0,SP): x_hi q_hi + x_lo q_hi     cols [SP,2SP): x_hi q_lo + x_lo q_lo
        mbar_spin(&ctl.xfull[set], (i / NSETS) & 1);
        tc_fence_after();
        SF_T(1, i, 0);
        const uint32_t xs = x0 + set * C::TILE_BYTES;
        const uint32_t d_tmem = tmem + 64u * set;
        const uint64_t dxh0 = umma_desc_k_sw128(xs, GB), dxl0 = umma_desc_k_sw128(xs + NKB * 1024, GB);
        const uint64_t dq0 = umma_desc_k_sw128(q0, 1024);
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t xadv = uint64_t((kb * 1024 + k * 32) >> 4);
            const uint64_t qadv = uint64_t((kb * (2 * SP * 128) + k * 32) >> 4);
            umma_f16(d_tmem, dxh0 + xadv, dq0 + qadv, id1, (kb | k) ? 1u : 0u);
            umma_f16(d_tmem, dxl0 + xadv, dq0 + qadv, id1, 1u);
          }
        }
        umma_commit(&ctl.lfull[set]);
        SF_T(1, i, 1);
      }
    }
    __syncwarp();
  } else if (warp == SF_G2_WARP) {
    // ================================================================ MMA issuer 2: weighted-sum product
    if (lane == 0) {
      const uint32_t id2 = umma_idesc_f16(SF_TILE, 2 * SP) | (1u << 15);     // A MN-major (features^T), B K-major
      const uint32_t x0 = smem_u32(xop), a0 = smem_u32(aop);
      for (int i = 0; i < ntiles; ++i) {
        const int set = i % NSETS;
        // per 64-channel block kb: D_kb[128 x 2SP] += [X_hi^T ; X_lo^T][(64+64) x 128 tokens] * [A_hi ; A_lo]^T
        //   (rows 0..63: hi plane of channels 64 kb.., rows 64..127: lo plane -- the two MN blocks of the A descriptor)
        mbar_spin(&ctl.afull[set], (i / NSETS) & 1);
        tc_fence_after();
        SF_T(2, i, 0);
        const uint32_t xs = x0 + set * C::TILE_BYTES;
        const uint64_t da0 = umma_desc_k_sw128(a0 + set * C::ABYTES, 1024);
        const uint64_t dx0 = umma_desc_mn_sw128(xs, NKB * 1024, GB);
        const uint32_t first = i ? 1u : 0u;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          const uint32_t d_tmem = tmem + 128u + uint32_t(2 * SP * kb);
#pragma unroll
          for (int k2 = 0; k2 < SF_TILE / 16; ++k2) {
            const uint64_t xadv = uint64_t((kb * 1024 + k2 * 2 * GB) >> 4);   // 16 tokens = two 8-token groups
            const uint64_t aadv = uint64_t(((k2 >> 2) * (2 * SP * 128) + (k2 & 3) * 32) >> 4);
            umma_f16(d_tmem, dx0 + xadv, da0 + aadv, id2, k2 ? 1u : first);
          }
        }
        umma_commit(&ctl.xempty[set]);    // operand tile, a tile and the logits columns of this set may be overwritten
        SF_T(2, i, 1);
      }
      umma_commit(&ctl.ufull);
    }
    __syncwarp();
  } else if (is_conv) {
    // ================================================================ converters: LayerNorm + fp16 split, in place
    // a warp owns whole 8-token groups (the unit of the in-place rewrite); 8 lanes per token, two rounds of 4 tokens
    for (int i = 0; i < ntiles; ++i) {
      const int set = i % NSETS;
      for (int gi = cw; gi < SF_GROUPS; gi += CW) {
        const int cc = gi >> 2;
        const int c = i * SF_CPT + cc;
        uint8_t* grp = xop + set * C::TILE_BYTES + gi * GB;   // fp32 rows in, UMMA atoms out
        float4 v[2][NV];
        if (c < nchunks) mbar_wait_warp(&ctl.sfull[set * SF_CPT + cc], (i / NSETS) & 1, lane);   // CTA-uniform branch
        else mbar_wait_warp(&ctl.xempty[set], ((i / NSETS) & 1) ^ 1, lane);   // zero-filled chunk: the set must still be free
        if (lane == 0 && gi == 0) SF_T(3, i, 0);
#if SDB_SF_EXPERIMENTAL
        // every token of the group inside the sample (warp-uniform; false only in the last group of a ragged N):
        // skips the 48 selects per lane that zero the rows beyond the end (8 % of the converter's instruction stream,
        // which is issue-bound: ncu `stall_not_selected` is its top stall reason, profiles/README.md section 12).
        const bool group_full = (i * SF_TILE + gi * 8 + 8) <= ntok;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int tg = tok_in_group(q);
          const float* row = reinterpret_cast<const float*>(grp) + tg * DIN + 4 * j;
#pragma unroll
          for (int k = 0; k < NV; ++k) v[q][k] = *reinterpret_cast<const float4*>(row + 32 * k);
        }
        if (!group_full) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const bool valid = (i * SF_TILE + gi * 8 + tok_in_group(q)) < ntok;   // stale bytes beyond the end -> zeros
#pragma unroll
            for (int k = 0; k < NV; ++k)
              if (!valid) v[q][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#else   // GPU-verified form
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int tg = tok_in_group(q);
          const bool valid = (i * SF_TILE + gi * 8 + tg) < ntok;
          const float* row = reinterpret_cast<const float*>(grp) + tg * DIN + 4 * j;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            v[q][k] = *reinterpret_cast<const float4*>(row + 32 * k);   // stale bytes if !valid: zeroed below, no branch
            if (!valid) v[q][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#endif
        __syncwarp();                                       // all fp32 rows of the group are in registers
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int tg = tok_in_group(q);
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < NV; ++k) sum += (v[q][k].x + v[q][k].y) + (v[q][k].z + v[q][k].w);
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          sum += __shfl_xor_sync(0xffffffffu, sum, 4);
          const float mean = sum * (1.f / DIN);
          float sq = 0.f;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            v[q][k].x -= mean; v[q][k].y -= mean; v[q][k].z -= mean; v[q][k].w -= mean;
            sq += (v[q][k].x * v[q][k].x + v[q][k].y * v[q][k].y) + (v[q][k].z * v[q][k].z + v[q][k].w * v[q][k].w);
          }
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          const float rstd = rsqrtf(sq * (1.f / DIN) + ln_eps);   // zero rows stay zero
          // swizzled 16-B chunk of channel block (j >> 1) + 4 (k & 1) in row tg: (chunk ^ tg); the k parity only flips bit 6
          const uint32_t o0 = tg * 128 + ((((j >> 1) ^ (tg & 3)) | ((tg >> 2) << 2)) << 4) + (j & 1) * 8;
          uint8_t* w0 = grp + o0;
          uint8_t* w1 = grp + (o0 ^ 64u);
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            uint2 hi, lo;
            split2(v[q][k].x * rstd, v[q][k].y * rstd, hi.x, lo.x);
            split2(v[q][k].z * rstd, v[q][k].w * rstd, hi.y, lo.y);
            uint8_t* dst = ((k & 1) ? w1 : w0) + (k >> 1) * 1024;
            *reinterpret_cast<uint2*>(dst) = hi;
            *reinterpret_cast<uint2*>(dst + NKB * 1024) = lo;
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && cw == 0) SF_T(3, i, 1);
      if (lane == 0 && cw == CW - 1) SF_T(3, i, 3);
      if (lane == 0) mbar_arrive(&ctl.xfull[set]);
    }
  } else {
    // ================================================================ softmax over slots; thread <-> token <-> TMEM lane
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    // a operand stores: tokens (2m, 2m+1) share a 32-bit word of a slot row; even lanes write slot rows s, odd lanes
    // rows s + 4 (other half of the 128-B bank space) -> every 32-lane store is one conflict-free wavefront
    const int col = r & 63, odd = lane & 1;
    const uint32_t a_lane = (r >> 6) * (2 * SP * 128) + (odd ? 4 * 128 : 0) + (((col >> 3) ^ (odd ? 4 : 0)) << 4) +
                            ((col & 7) >> 1) * 4;
    const uint32_t sel_send = odd ? 0x5410u : 0x7632u;
    const uint32_t sel_hi = odd ? 0x3254u : 0x5410u, sel_lo = odd ? 0x3276u : 0x7610u;
    float cs[SP];
#pragma unroll
    for (int s = 0; s < SP; ++s) cs[s] = 0.f;
    for (int i = 0; i < ntiles; ++i) {
      const int set = i % NSETS;
      mbar_wait_warp(&ctl.lfull[set], (i / NSETS) & 1, lane);
      tc_fence_after();
      if (tid == 0) SF_T(4, i, 0);
      float l[SP];
      tmem_ld_folded<SP>(lane_addr + 64u * set, l);
      const int n = n_begin + i * SF_TILE + r;
      const bool valid = n < n_end;
      float mx = -INFINITY;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        l[s] += ctl.cb[s];
        if (s < S) mx = fmaxf(mx, l[s]);
      }
      float sum = 0.f;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        l[s] = (s < S) ? __expf(l[s] - mx) : 0.f;
        sum += l[s];
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int s = 0; s < SP; ++s) l[s] *= inv;                     // softmax over slots
      if (seg_mask && valid) {
        float* mrow = seg_mask + (b * S) * N + n;
#pragma unroll
        for (int s = 0; s < SP; ++s)
          if (s < S) mrow[(long long)s * N] = l[s];
      }
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        l[s] = (valid && s < S) ? l[s] + eps : 0.f;                   // a = attn + eps (zero for padding slots / tokens)
        cs[s] += l[s];
      }
      uint8_t* abase = aop + set * C::ABYTES;
#pragma unroll
      for (int pq = 0; pq < SP / 2; ++pq) {
        const int s = (pq & 3) + 8 * (pq >> 2);                      // slots s and s + 4
        uint32_t H, L;
        split2(l[s] * SF_ASCALE, l[s + 4] * SF_ASCALE, H, L);
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, __byte_perm(H, L, sel_send), 1);
        uint8_t* dst = abase + ((a_lane ^ ((s & 3) << 4)) + s * 128);
        *reinterpret_cast<uint32_t*>(dst) = __byte_perm(H, recv, sel_hi);
        *reinterpret_cast<uint32_t*>(dst + SP * 128) = __byte_perm(L, recv, sel_lo);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (tid == 0) SF_T(4, i, 1);
      if (lane == 0) mbar_arrive(&ctl.afull[set]);
    }
    // ---- drain U^T: block kb, TMEM lane r < 64: hi-plane part of channel 64 kb + r; lane 64 + r: its lo-plane part
    mbar_wait_warp(&ctl.ufull, 0, lane);
    tc_fence_after();
    if (tid == 0) SF_T(5, 0, 0);
    float* scr = reinterpret_cast<float*>(xop);       // [NKB][SP][64] fp32; the operand tiles are dead now
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      float u[SP];
      tmem_ld_folded<SP>(lane_addr + 128u + uint32_t(2 * SP * kb), u);
      if (r >= 64) {
#pragma unroll
        for (int s = 0; s < SP; ++s) scr[(kb * SP + s) * 64 + (r - 64)] = u[s];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(SF_SOFT_WARPS * 32) : "memory");
      if (r < 64) {
#pragma unroll
        for (int s = 0; s < SP; ++s)
          if (s < S) my_upd[s * DIN + kb * 64 + r] = u[s] + scr[(kb * SP + s) * 64 + r];
      }
    }
#pragma unroll
    for (int s = 0; s < SP; ++s) {
      const float w = warp_sum(cs[s]);
      if (lane == 0) ctl.cs_scr[warp][s] = w;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SF_SOFT_WARPS * 32) : "memory");
    if (tid < S) my_cs[tid] = (ctl.cs_scr[0][tid] + ctl.cs_scr[1][tid]) + (ctl.cs_scr[2][tid] + ctl.cs_scr[3][tid]);
    if (tid == 0) SF_T(5, 0, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SF_G1_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, SF_TMEM_COLS);
  }
}

// U[b,s,:] = sum_chunks part_upd / (ascale * sum_chunks part_cs)  -> packed GEMM operand (+ fp32 copy)
__global__ void slot_attend_fused_finalize_kernel(const float* __restrict__ part_upd, const float* __restrict__ part_cs,
                                                  __half* __restrict__ out, float* __restrict__ upd32, int64_t BS, int S,
                                                  int D, int chunks, float ascale) {
  const int64_t bs = blockIdx.x;
  const int64_t b = bs / S;
  const int s = (int)(bs % S);
  float cs = 0.f;
  for (int c = 0; c < chunks; ++c) cs += part_cs[(b * chunks + c) * S + s];
  const float inv = 1.f / (cs * ascale);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = 0.f;
    for (int c = 0; c < chunks; ++c) v += part_upd[((b * chunks + c) * S + s) * D + d];
    v *= inv;
    if (upd32) upd32[bs * D + d] = v;
    __half h, l;
    split_f16(v, h, l);
    out[bs * D + d] = h;
    out[BS * D + bs * D + d] = l;
  }
}

static int sf_chunks(int64_t B, int64_t N) {
  int64_t tiles = cdiv(N, SF_TILE);
  int64_t c = num_sms() / (B > 0 ? B : 1);
  if (c < 1) c = 1;
  if (c > tiles) c = tiles;
  return (int)c;
}

template <int DIN, int SP>
static int launch_fused(const float* x, const float* qa, int ldq, float* seg_mask, float* part_upd, float* part_cs,
                        int64_t B, int N, int S, int chunks, float ln_eps, float eps, long long* dbg, cudaStream_t st) {
  using C = SfCfg<DIN, SP>;
  constexpr int CW = (SP == 16 && DIN <= 192) ? 16 : 8;
  constexpr int SF_THREADS = 32 * (SF_SOFT_WARPS + CW + 3);
  auto kern = slot_attend_fused_kernel<DIN, SP, CW>;
  static bool attr = false;
  if (!attr) {
    SDB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  dim3 grid(chunks, (unsigned)B);
  kern<<<grid, SF_THREADS, C::SMEM, st>>>(x, qa, ldq, seg_mask, part_upd, part_cs, N, S, chunks, ln_eps, eps, dbg);
  SDB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sdb

using namespace sdb;

static long long* g_sf_dbg = nullptr;
/* debug: device buffer of 6*64 int64 that receives the timeline of CTA (0,0) of every following launch (NULL = off) */
extern "C" int sdb_slot_attend_fused_debug(void* buf) {
  g_sf_dbg = reinterpret_cast<long long*>(buf);
  return 0;
}

extern "C" int sdb_slot_attend_fused_supported(int64_t S, int64_t Din) {
  if (S < 1 || S > 32) return 0;
  if (Din == 128 || Din == 192) return 1;
  if (Din == 256) return 1;
  return 0;
}

extern "C" int64_t sdb_slot_attend_fused_workspace(int64_t B, int64_t N, int64_t S, int64_t Din) {
  const int chunks = sf_chunks(B, N);
  return (B * chunks * S * Din + B * chunks * S) * (int64_t)sizeof(float);
}

// finalize = 0: leave the per-chunk partial sums in `work` (sdb_slot_update consumes them)
static int sf_run(const float* x, const float* qa, int64_t ldq, float* seg_mask, void* upd_packed, float* upd32,
                  float* work, int64_t B, int64_t N, int64_t S, int64_t Din, float ln_eps, float eps, void* stream,
                  int finalize) {
  SDB_REQUIRE(x && qa && work && (upd_packed || !finalize), "sdb_slot_attend_fused: null argument");
  SDB_REQUIRE(B > 0 && B <= 65535 && N > 0 && N < (1 << 24), "sdb_slot_attend_fused: bad B=%lld N=%lld", (long long)B,
              (long long)N);
  SDB_REQUIRE(sdb_slot_attend_fused_supported(S, Din), "sdb_slot_attend_fused: unsupported num_slots=%lld in_features=%lld",
              (long long)S, (long long)Din);
  SDB_REQUIRE(ldq >= Din + 1 && ldq % 4 == 0, "sdb_slot_attend_fused: ldq=%lld must be a multiple of 4 and > in_features",
              (long long)ldq);
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(qa) & 15) == 0,
              "sdb_slot_attend_fused: x and qa must be 16-byte aligned");
  const int chunks = sf_chunks(B, N);
  float* part_upd = work;
  float* part_cs = work + B * chunks * S * Din;
  cudaStream_t st = as_stream(stream);
  int rc = 0;
#define SF_CASE(DD)                                                                                                  \
  if (Din == DD) {                                                                                                   \
    rc = (S <= 16) ? launch_fused<DD, 16>(x, qa, (int)ldq, seg_mask, part_upd, part_cs, B, (int)N, (int)S, chunks,   \
                                          ln_eps, eps, g_sf_dbg, st)                                                 \
                   : launch_fused<DD, 32>(x, qa, (int)ldq, seg_mask, part_upd, part_cs, B, (int)N, (int)S, chunks,   \
                                          ln_eps, eps, g_sf_dbg, st);                                                \
  }
  SF_CASE(128) else SF_CASE(192) else SF_CASE(256)
#undef SF_CASE
  if (rc || !finalize) return rc;
  slot_attend_fused_finalize_kernel<<<(unsigned)(B * S), 64, 0, st>>>(part_upd, part_cs, (__half*)upd_packed, upd32,
                                                                      B * S, (int)S, (int)Din, chunks, SF_ASCALE);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_slot_attend_fused(const float* x, const float* qa, int64_t ldq, float* seg_mask, void* upd_packed,
                                     float* upd32, float* work, int64_t B, int64_t N, int64_t S, int64_t Din,
                                     float ln_eps, float eps, void* stream) {
  return sf_run(x, qa, ldq, seg_mask, upd_packed, upd32, work, B, N, S, Din, ln_eps, eps, stream, 1);
}

/* token chunks per sample the attend kernel uses for (B, N) on this device, and the scale of its partial sums:
 * work = part_upd [B, chunks, S, Din] (= ascale * sum_n a[n,s] n[n,:]) followed by part_cs [B, chunks, S] (= sum_n a) */
extern "C" int64_t sdb_slot_attend_fused_chunks(int64_t B, int64_t N) { return sf_chunks(B, N); }
extern "C" float sdb_slot_attend_fused_ascale(void) { return SF_ASCALE; }

extern "C" int sdb_slot_attend_fused_partials(const float* x, const float* qa, int64_t ldq, float* seg_mask, float* work,
                                              int64_t B, int64_t N, int64_t S, int64_t Din, float ln_eps, float eps,
                                              void* stream) {
  return sf_run(x, qa, ldq, seg_mask, nullptr, nullptr, work, B, N, S, Din, ln_eps, eps, stream, 0);
}
