// Slot Attention inner loop, tensor-core form (slot_attention.py:67-91, sa_diffusion.py:43-56) -- inference path.
//
// The k / v projections are never materialised.  With n = (x - mean) * rstd (LayerNorm without its affine part):
//   logits[n,s] = scale * k[n] . q[s]          = n[n] . qa[s,:Din] + qa[s,Din]          (qa = LN_q(slots) W_qa^T, host-folded
//                                                                                        W_qa = scale * diag(gamma) Wk^T Wq | beta row)
//   updates[s]  = (sum_n a[n,s] v[n]) / sum_n a = (gamma * U[s] + beta) Wv^T,  U[s] = (sum_n a[n,s] n[n]) / sum_n a[n,s]
// so one pass over the RAW features per iteration does LayerNorm statistics, both N x S contractions and the softmax:
//
//   warp 13      TMA producer: cp.async.bulk of 32-token fp32 chunks straight into the operand slots (mbarrier full/empty)
//   warps 4..11  converters: LayerNorm (8 lanes per token, warp-shuffle reductions) and fp16 hi/lo split IN PLACE: the
//                6 KB of eight fp32 token rows become the 6 one-KB 128B-swizzled UMMA atoms (2 planes x Din/64 blocks)
//                of the same eight tokens.  One operand tile [128 tokens][Din] serves BOTH contractions (K-major for
//                the logits, MN-major for the weighted sum); tiles are double-buffered (two sets of four chunk slots)
//   warp 12      MMA issuer (one thread): logits[128 x SP] = X Q^T, then U^T[Din x SP] += X^T A, tcgen05.mma kind::f16.
//                fp32-faithful products hi*hi + lo*hi + hi*lo in TWO instructions per k-step: the hi and lo planes of
//                the small operand are stacked along N (x_hi * [b_hi ; b_lo], then x_lo * b_hi) -- small MMAs are
//                issue-bound (~51 cycles each, tools/probes/umma_probe.cu), so instruction count is what matters
//   warps 0..3   softmax over slots (thread <-> token <-> TMEM lane), seg-mask store, a = softmax + eps as the B operand
//                of the second contraction, column sums by warp shuffles; finally drain U from TMEM
//
// Grid (chunks, B): a CTA owns a contiguous range of 128-token tiles of one sample; partial sums are combined by
// slot_attend_fused_finalize_kernel (deterministic, no atomics).
#include "common.cuh"
#include "ptx.cuh"

namespace sdb {

constexpr int SF_TILE = 128;          // tokens per tile (UMMA M of the logits product, K extent of the update product)
constexpr int SF_CT = 32;             // tokens per TMA chunk
constexpr int SF_CPT = SF_TILE / SF_CT;   // chunks per tile
constexpr int SF_SOFT_WARPS = 4;
constexpr int SF_CONV_WARPS = 8;
constexpr int SF_MMA_WARP = SF_SOFT_WARPS + SF_CONV_WARPS;   // 12
constexpr int SF_TMA_WARP = SF_MMA_WARP + 1;                 // 13
constexpr int SF_THREADS = 32 * (SF_TMA_WARP + 1);           // 448
constexpr int SF_TMEM_COLS = 256;     // logits set s at [64 s, 64 s + 2 SP) | U half h at [128 + 64 h, ... + 2 SP)
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
  static constexpr int GROUP_BYTES = 2 * NKB * 1024;         // 8 tokens: fp32 rows == 2 planes x NKB atoms of 1 KB
  static constexpr int CHUNK_BYTES = (SF_CT / 8) * GROUP_BYTES;
  static constexpr int TILE_BYTES = SF_CPT * CHUNK_BYTES;
  static constexpr int QBYTES = NKB * 2 * SP * 128;          // per k-block: [q_hi rows | q_lo rows] x 128 B
  static constexpr int ABYTES = (SF_TILE / 64) * 2 * SP * 128;   // per set: 2 token blocks x [a_hi rows | a_lo rows] x 128 B
  static constexpr int CTL = 1024;
  static constexpr int NSETS = (1024 + 2 * TILE_BYTES + QBYTES + 2 * ABYTES + CTL <= 227 * 1024) ? 2 : 1;
  static constexpr int SMEM = 1024 + NSETS * TILE_BYTES + QBYTES + NSETS * ABYTES + CTL;
  static_assert(GROUP_BYTES == 8 * DIN * 4, "in-place conversion needs equal fp32 and fp16x2 footprints");
  static_assert(SMEM <= 227 * 1024, "slot_attend_fused: tile does not fit shared memory");
  static_assert(sizeof(SfCtl) <= CTL, "control block too large");
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

// one lane polls, the warp follows (keeps hundreds of threads from hammering the barrier)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
  mbar_wait(bar, parity);   // already complete: one try_wait per thread = its own acquire
}

template <int DIN, int SP>
__global__ void __launch_bounds__(SF_THREADS, 1)
slot_attend_fused_kernel(const float* __restrict__ x, const float* __restrict__ qa, int ldq,
                         float* __restrict__ seg_mask, float* __restrict__ part_upd, float* __restrict__ part_cs,
                         int N, int S, int chunks, float ln_eps, float eps) {
  using C = SfCfg<DIN, SP>;
  constexpr int NKB = C::NKB;
  constexpr int NSETS = C::NSETS;
  constexpr int GB = C::GROUP_BYTES;
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

  // ------------------------------------------------------------------ prologue
  if (tid == 0) {
    for (int i = 0; i < 2 * SF_CPT; ++i) mbar_init(&ctl.sfull[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl.xfull[i], SF_CONV_WARPS);
      mbar_init(&ctl.xempty[i], 1);
      mbar_init(&ctl.lfull[i], 1);
      mbar_init(&ctl.afull[i], SF_SOFT_WARPS);
    }
    mbar_init(&ctl.ufull, 1);
    fence_mbar_init();
  }
  if (warp == SF_MMA_WARP) {
    tmem_alloc(&ctl.tmem_base, SF_TMEM_COLS);
    tmem_relinquish();
  }
  // slot-side operand of the logits product: qa[s, 0:DIN] split into fp16 planes stacked along N, K-major swizzled rows
  for (int i = tid; i < SP * (DIN / 4); i += SF_THREADS) {
    const int s = i / (DIN / 4), c = (i % (DIN / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < S) v = *reinterpret_cast<const float4*>(qa + (b * S + s) * ldq + c);
    uint2 hi, lo;
    split2(v.x, v.y, hi.x, lo.x);
    split2(v.z, v.w, hi.y, lo.y);
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
  const uint32_t tmem = ctl.tmem_base;

  if (warp == SF_TMA_WARP) {
    // ================================================================ TMA producer
    if (lane == 0) {
      for (int i = 0; i < ntiles; ++i) {
        const int set = i % NSETS;
        mbar_wait(&ctl.xempty[set], ((i / NSETS) & 1) ^ 1);    // the MMAs of the tile that used this set have retired
        for (int cc = 0; cc < SF_CPT; ++cc) {
          const int c = i * SF_CPT + cc;
          if (c >= nchunks) break;
          const int rows = min(SF_CT, ntok - c * SF_CT);
          const uint32_t bytes = (uint32_t)rows * DIN * 4;
          uint64_t* bar = &ctl.sfull[set * SF_CPT + cc];
          mbar_arrive_expect_tx(bar, bytes);
          bulk_load(xop + set * C::TILE_BYTES + cc * C::CHUNK_BYTES, x + (b * N + n_begin + c * SF_CT) * DIN, bytes, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == SF_MMA_WARP) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t id1w = umma_idesc_f16(SF_TILE, 2 * SP), id1n = umma_idesc_f16(SF_TILE, SP);   // A, B K-major
      const uint32_t id2w = id1w | (1u << 15), id2n = id1n | (1u << 15);                          // A MN-major (features^T)
      const uint32_t x0 = smem_u32(xop), q0 = smem_u32(qop), a0 = smem_u32(aop);
      constexpr int NH = NKB > 2 ? 2 : 1;       // 128-channel halves of U (overlapping when DIN = 192)
      int g1 = 0, g2 = 0;                       // next tile of the logits product / of the update product
      long long t_idle = 0;
      while (g2 < ntiles) {
        bool progress = false;
        if (g1 < ntiles && g1 - g2 < NSETS && mbar_try_wait(&ctl.xfull[g1 % NSETS], (g1 / NSETS) & 1)) {
          // ---- logits[128 tokens x SP] = X[128 x DIN] * Qa[SP x DIN]^T       (cols [0,SP): hi*hi + lo*hi, [SP,2SP): hi*lo)
          tc_fence_after();
          const int set = g1 % NSETS;
          const uint32_t xs = x0 + set * C::TILE_BYTES;
          const uint32_t d_tmem = tmem + 64u * set;
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = uint64_t((k * 32) >> 4);
              const uint64_t dxh = umma_desc_k_sw128(xs + kb * 1024, GB) + adv;
              const uint64_t dxl = umma_desc_k_sw128(xs + (NKB + kb) * 1024, GB) + adv;
              const uint64_t dq = umma_desc_k_sw128(q0 + kb * (2 * SP * 128), 1024) + adv;
              umma_f16(d_tmem, dxh, dq, id1w, (kb | k) ? 1u : 0u);
              umma_f16(d_tmem, dxl, dq, id1n, 1u);
            }
          }
          umma_commit(&ctl.lfull[set]);
          ++g1;
          progress = true;
        }
        if (g2 < g1 && mbar_try_wait(&ctl.afull[g2 % NSETS], (g2 / NSETS) & 1)) {
          // ---- U^T[128 channels x SP] += X^T[channels x 128 tokens] * A[SP x 128 tokens]^T
          tc_fence_after();
          const int set = g2 % NSETS;
          const uint32_t xs = x0 + set * C::TILE_BYTES;
          const uint32_t as = a0 + set * C::ABYTES;
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const int m0 = h == 0 ? 0 : NKB - 2;  // first 64-channel block of this half
            const uint32_t d_tmem = tmem + 128u + 64u * h;
#pragma unroll
            for (int k2 = 0; k2 < SF_TILE / 16; ++k2) {
              const uint64_t xadv = uint64_t((k2 * 2 * GB) >> 4);             // 16 tokens = two 8-token groups
              const uint64_t dxh = umma_desc_mn_sw128(xs + m0 * 1024, 1024, GB) + xadv;
              const uint64_t dxl = umma_desc_mn_sw128(xs + (NKB + m0) * 1024, 1024, GB) + xadv;
              const uint64_t da = umma_desc_k_sw128(as + (k2 >> 2) * (2 * SP * 128), 1024) + uint64_t(((k2 & 3) * 32) >> 4);
              umma_f16(d_tmem, dxh, da, id2w, (g2 | k2) ? 1u : 0u);
              umma_f16(d_tmem, dxl, da, id2n, 1u);
            }
          }
          umma_commit(&ctl.xempty[set]);    // operand tile, a tile and the logits columns of this set may be overwritten
          ++g2;
          progress = true;
        }
        if (progress) {
          t_idle = 0;
        } else {
          const long long now = clock64();
          if (t_idle == 0) t_idle = now;
          else if (now - t_idle > 4000000000LL) {
            printf("sdb200: slot_attend_fused MMA issuer stalled (block %d,%d g1 %d g2 %d)\n", blockIdx.x, blockIdx.y, g1, g2);
            __trap();
          }
        }
      }
      umma_commit(&ctl.ufull);
    }
    __syncwarp();
  } else if (warp >= SF_SOFT_WARPS) {
    // ================================================================ converters: LayerNorm + fp16 split, in place
    // warp cw owns token group (cw & 3) of every second chunk; 8 lanes per token, two rounds of 4 tokens
    const int cw = warp - SF_SOFT_WARPS, sub = lane >> 3, j = lane & 7;
    const int g = cw & 3;
    constexpr int NV = DIN / 32;                          // float4 per lane per token
    for (int i = 0; i < ntiles; ++i) {
      const int set = i % NSETS;
      for (int cc = cw >> 2; cc < SF_CPT; cc += 2) {
        const int c = i * SF_CPT + cc;
        uint8_t* grp = xop + set * C::TILE_BYTES + cc * C::CHUNK_BYTES + g * GB;   // fp32 rows in, UMMA atoms out
        float4 v[2][NV];
        bool valid[2];
        if (c < nchunks) mbar_wait_warp(&ctl.sfull[set * SF_CPT + cc], (i / NSETS) & 1, lane);   // CTA-uniform branch
        else mbar_wait_warp(&ctl.xempty[set], ((i / NSETS) & 1) ^ 1, lane);   // zero-filled chunk: the set must still be free
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int tg = (sub & 1) + 4 * (sub >> 1) + 2 * q;          // token within the group: rows {0,1,4,5} / {2,3,6,7}
          valid[q] = (c * SF_CT + g * 8 + tg) < ntok;
          const float* row = reinterpret_cast<const float*>(grp) + tg * DIN;
#pragma unroll
          for (int k = 0; k < NV; ++k)
            v[q][k] = valid[q] ? *reinterpret_cast<const float4*>(row + 4 * (j + 8 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();                                       // all fp32 rows of the group are in registers
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int tg = (sub & 1) + 4 * (sub >> 1) + 2 * q;
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
          const float rstd = valid[q] ? 1.f / sqrtf(sq * (1.f / DIN) + ln_eps) : 0.f;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            uint2 hi, lo;
            split2(v[q][k].x * rstd, v[q][k].y * rstd, hi.x, lo.x);
            split2(v[q][k].z * rstd, v[q][k].w * rstd, hi.y, lo.y);
            const int kb = k >> 1;
            const int col16 = (j >> 1) + 4 * (k & 1);
            const uint32_t off = kb * 1024 + tg * 128 + ((col16 ^ tg) << 4) + (j & 1) * 8;
            *reinterpret_cast<uint2*>(grp + off) = hi;
            *reinterpret_cast<uint2*>(grp + NKB * 1024 + off) = lo;
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.xfull[set]);
    }
  } else {
    // ================================================================ softmax over slots; thread <-> token <-> TMEM lane
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    float cs[SP];
#pragma unroll
    for (int s = 0; s < SP; ++s) cs[s] = 0.f;
    for (int i = 0; i < ntiles; ++i) {
      const int set = i % NSETS;
      mbar_wait_warp(&ctl.lfull[set], (i / NSETS) & 1, lane);
      tc_fence_after();
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
      const int kb = r >> 6, col = r & 63;
      uint8_t* arow = aop + set * C::ABYTES + kb * (2 * SP * 128) + (col & 7) * 2;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const float p = l[s] * inv;
        if (seg_mask && valid && s < S) seg_mask[(b * S + s) * N + n] = p;
        const float a = (valid && s < S) ? p + eps : 0.f;
        cs[s] += a;
        const float as = a * SF_ASCALE;
        const __half h = __float2half_rn(as);
        const __half lo = __float2half_rn(as - __half2float(h));
        const uint32_t off = s * 128 + (((col >> 3) ^ (s & 7)) << 4);
        *reinterpret_cast<__half*>(arow + off) = h;
        *reinterpret_cast<__half*>(arow + SP * 128 + off) = lo;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.afull[set]);
    }
    // ---- drain U^T (lane <-> channel, column <-> slot) and the column sums
    mbar_wait_warp(&ctl.ufull, 0, lane);
    tc_fence_after();
    constexpr int NH = NKB > 2 ? 2 : 1;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int ch = (h == 0 ? 0 : (NKB - 2) * 64) + r;
      const bool need = h == 0 || ch >= 128;
      float u[SP];
      tmem_ld_folded<SP>(lane_addr + 128u + 64u * h, u);
      if (need) {
#pragma unroll
        for (int s = 0; s < SP; ++s)
          if (s < S) my_upd[s * DIN + ch] = u[s];
      }
    }
#pragma unroll
    for (int s = 0; s < SP; ++s) {
      const float v = warp_sum(cs[s]);
      if (lane == 0) ctl.cs_scr[warp][s] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SF_SOFT_WARPS * 32) : "memory");
    if (tid < S) my_cs[tid] = (ctl.cs_scr[0][tid] + ctl.cs_scr[1][tid]) + (ctl.cs_scr[2][tid] + ctl.cs_scr[3][tid]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SF_MMA_WARP) {
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
                        int64_t B, int N, int S, int chunks, float ln_eps, float eps, cudaStream_t st) {
  using C = SfCfg<DIN, SP>;
  auto kern = slot_attend_fused_kernel<DIN, SP>;
  static bool attr = false;
  if (!attr) {
    SDB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  dim3 grid(chunks, (unsigned)B);
  kern<<<grid, SF_THREADS, C::SMEM, st>>>(x, qa, ldq, seg_mask, part_upd, part_cs, N, S, chunks, ln_eps, eps);
  SDB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sdb

using namespace sdb;

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

extern "C" int sdb_slot_attend_fused(const float* x, const float* qa, int64_t ldq, float* seg_mask, void* upd_packed,
                                     float* upd32, float* work, int64_t B, int64_t N, int64_t S, int64_t Din,
                                     float ln_eps, float eps, void* stream) {
  SDB_REQUIRE(x && qa && upd_packed && work, "sdb_slot_attend_fused: null argument");
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
                                          ln_eps, eps, st)                                                           \
                   : launch_fused<DD, 32>(x, qa, (int)ldq, seg_mask, part_upd, part_cs, B, (int)N, (int)S, chunks,   \
                                          ln_eps, eps, st);                                                          \
  }
  SF_CASE(128) else SF_CASE(192) else SF_CASE(256)
#undef SF_CASE
  if (rc) return rc;
  slot_attend_fused_finalize_kernel<<<(unsigned)(B * S), 64, 0, st>>>(part_upd, part_cs, (__half*)upd_packed, upd32,
                                                                      B * S, (int)S, (int)Din, chunks, SF_ASCALE);
  SDB_LAUNCH_CHECK();
  return 0;
}
