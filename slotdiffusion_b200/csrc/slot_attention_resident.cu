// Slot Attention, the WHOLE iterative update (slot_attention.py:67-104, sa_diffusion.py:28-70) as ONE persistent kernel.
//
// A thread-block cluster of CL = 4 / 8 CTAs owns one sample at a time (persistent loop over the batch).  The sample's
// raw features are read from HBM exactly ONCE: TMA (cp.async.bulk) lands the fp32 rows in shared memory, LayerNorm + fp16
// hi/lo split rewrite them IN PLACE as 128B-swizzled UMMA tiles (NT tiles of 128 tokens per CTA, CL * NT * 128 tokens per
// cluster), and all iterations run from that resident copy:
//
//   attend   logits[128 x SP] = X Q^T (tcgen05, TMEM) -> softmax over slots (thread <-> token <-> TMEM lane) -> seg mask ->
//            U^T[Din x SP] += X^T A (tcgen05; same operand tile, MN-major view); column sums by warp shuffles
//   reduce   per-CTA partial U, column sums -> reduce-scatter / all-gather over distributed shared memory
//   update   GRU (folded input projection), LayerNorm, MLP + residual, LayerNorm_q, folded slot-side projection of the next
//            iteration: fp32 FMA on the CUDA cores, every stage column-split over the CL CTAs of the cluster (each CTA streams
//            1/CL of the 1.6 MB of weights from L2 per iteration), results all-gathered by DSMEM stores
//
// so a forward is one launch, no k / v tensors, no intermediate in HBM; the next sample's feature load is issued as soon as
// the last token contraction of the current one has retired and overlaps its last update.
// The algebra (k / v projections and the LayerNorm affine folded into the slot side) is that of slot_attention_fused.cu;
// weights arrive fp32, TRANSPOSED and k-quad interleaved: w4[K/4][ncols][4] (ops.WeightCache.slot_resident_weights).
//
// Warps: 0..NWORK-1 workers (conversion, update; the first 4*NT of them also softmax / TMEM drain), then the two MMA
// issuers (the first one also issues the TMA loads).
#include <cstddef>

#include "common.cuh"
#include "ptx.cuh"

namespace sdb {
namespace sr {

constexpr int TILE = 128;             // tokens per tile (UMMA M of the logits product, K extent of the update product)
constexpr int GROUPS = TILE / 8;
constexpr int CT = 32;                // tokens per TMA chunk
constexpr int CPT = TILE / CT;        // chunks per tile
constexpr int TMEM_COLS = 512;        // logits of tile t at [64 t, 64 t + 2 SP) | U block kb at [128 + 2 SP kb, ... + 2 SP)
constexpr float ASCALE = 4096.f;      // a = softmax + eps is scaled before the fp16 split (keeps 1e-6 out of fp16 subnormals)
constexpr int MAX_SMEM = 227 * 1024;

struct Ctl {
  uint64_t sfull[2 * CPT];            // TMA chunk landed (per sample)
  uint64_t xfull[2];                  // tile converted (per sample)
  uint64_t lfull[2], afull[2];        // logits in TMEM / a operand written (per iteration)
  uint64_t ufull;                     // weighted sums complete (per iteration)
  uint32_t tmem_base;
  uint32_t pad;
  float cb[32];                       // logit bias per slot (beta row of the folded projection)
  float cs_scr[128];                  // per-warp column sums [softmax warp][SP] (8 x 16 or 4 x 32)
  float cs_part[32];                  // this CTA's column sums (read by the whole cluster)
};

template <int DIN, int SP, int NT>
struct Cfg {
  static constexpr int NKB = DIN / 64;
  static constexpr int GROUP_BYTES = 2 * NKB * 1024;         // 8 tokens: fp32 rows == 2 planes x NKB swizzle atoms of 1 KB
  static constexpr int CHUNK_BYTES = (CT / 8) * GROUP_BYTES;
  static constexpr int TILE_BYTES = GROUPS * GROUP_BYTES;
  static constexpr int QBYTES = NKB * 2 * SP * 128;          // per k-block: [q_hi rows | q_lo rows] x 128 B
  static constexpr int ABYTES = (TILE / 64) * 2 * SP * 128;  // per tile: 2 token blocks x [a_hi rows | a_lo rows] x 128 B
  static constexpr int CTL = 1024;
  static constexpr int SCR = ((MAX_SMEM - CTL - NT * TILE_BYTES) / 1024) * 1024;   // scratch: operands q | a during attend, update regions after
  static constexpr int SMEM = NT * TILE_BYTES + SCR + CTL;
  static_assert(GROUP_BYTES == 8 * DIN * 4, "in-place conversion needs equal fp32 and fp16x2 footprints");
  static_assert(NT >= 1 && NT <= 2, "one or two resident tiles per CTA");
  static_assert(QBYTES + NT * ABYTES <= SCR, "operand tiles do not fit the scratch region");
  static_assert(sizeof(Ctl) <= CTL, "control block too large");
  static_assert(128 + NKB * 2 * SP <= TMEM_COLS, "TMEM columns");
  static_assert(SMEM <= MAX_SMEM, "shared memory");
};

// ---------------------------------------------------------------- small PTX helpers (cluster / DSMEM)
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t a, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t a, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// full cluster barrier (every thread of every CTA of the cluster), release / acquire at cluster scope
__device__ __forceinline__ void cluster_barrier() {
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
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

// bounded waits: a protocol bug must trap (visible CUDA error), never hang the GPU box
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
    if ((++polls & 0xfffu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        printf("sdb200: slot_attention_resident mbarrier spin timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) {
    long long t0 = 0;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
      if ((++polls & 0xffu) == 0) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) {
          printf("sdb200: slot_attention_resident mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
          __trap();
        }
      }
    }
  }
  __syncwarp();
  mbar_wait(bar, parity);   // already complete: one try_wait per thread = its own acquire
}

struct Args {
  SdbSlotAttentionResident p;
  int rsz;                  // bytes per update region (4 regions at the start of the scratch area)
  long long* dbg;           // optional timeline of thread 0 of CTA 0 (tools/sa_resident_timeline.py); NULL = off
};

// ---------------------------------------------------------------- update stage: column-sliced matrix product on the CUDA cores
// A warp owns 8 consecutive output columns (lane & 7) and splits its K range four ways (ks = lane >> 3, k-quads interleaved:
// the four 16-byte activation reads of a warp are one conflict-free 64-byte wavefront, every lane loads DISTINCT weights --
// 32 lanes x 16 B = four full 128-byte lines per load instruction):
//   acc[g][r] = sum_k x[r][k] * w4[k / 4][col_g][k % 4]       for NG column sets (the GRU gates r | z | n of a unit), RS rows.
// Weight loads are staged PD bodies ahead in registers.  Measured on the B200 (profiles/README.md, round 2): the stage time
// does not react to the staging depth -- ptxas gives every global load of a loop the same scoreboard and sinks the loads to
// the end of the trip, so one L2 round trip (~1 k cycles) per trip stays exposed -- and the loop is then bound by the
// shared-memory pipe: RS broadcast LDS.128 (512 B of write-back each) per 4 NG RS FMAs.  All 32 lanes hold the full sums
// on return.  Rows >= S read whatever follows the region (inside the CTA's allocation, checked by the entry point); their
// sums are discarded by the caller.
#ifndef SR_FENCE
#define SR_FENCE() __syncwarp()   /* a warp barrier: ptxas does not sink the loads above it towards their uses */
#endif
template <int RS, int NG, int BATCH, int RH, int LDX>
__device__ __forceinline__ void dot_ks(const float* __restrict__ x, int ks, const float4* __restrict__ w4, int ncols,
                                       const int (&col)[NG], int kq0, int kq1, float (&acc)[NG][RS]) {
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int r = 0; r < RS; ++r) acc[g][r] = 0.f;
  const float4* wl = w4 + (size_t)(kq0 + ks) * ncols;
  const float* xl = x + 4 * (kq0 + ks);
  const int nbatch = (kq1 - kq0) / (4 * BATCH);      // lane quad index of (batch, u): kq0 + 4 (batch BATCH + u) + ks
#pragma unroll 1
  for (int bt = 0; bt < nbatch; ++bt) {
    float4 buf[BATCH][NG];
#pragma unroll
    for (int u = 0; u < BATCH; ++u)
#pragma unroll
      for (int g = 0; g < NG; ++g) buf[u][g] = __ldg(wl + (size_t)(4 * (bt * BATCH + u)) * ncols + col[g]);
    SR_FENCE();
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const float* xk = xl + 16 * (bt * BATCH + u);
#pragma unroll
      for (int r0 = 0; r0 < RS; r0 += RH) {
        float4 xv[RH];
#pragma unroll
        for (int r = 0; r < RH; ++r) xv[r] = *reinterpret_cast<const float4*>(xk + (r0 + r) * LDX);
        SR_FENCE();
#define SR_FMA_COMP(C)                                                                             \
  _Pragma("unroll") for (int r = 0; r < RH; ++r) _Pragma("unroll") for (int g = 0; g < NG; ++g)    \
      acc[g][r0 + r] = fmaf(xv[r].C, buf[u][g].C, acc[g][r0 + r]);
        SR_FMA_COMP(x) SR_FMA_COMP(y) SR_FMA_COMP(z) SR_FMA_COMP(w)
#undef SR_FMA_COMP
      }
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int r = 0; r < RS; ++r) {
      acc[g][r] += __shfl_xor_sync(0xffffffffu, acc[g][r], 8);
      acc[g][r] += __shfl_xor_sync(0xffffffffu, acc[g][r], 16);
    }
}

// a[4 i + ks] without dynamic register indexing
template <int RS>
__device__ __forceinline__ float sel_row(const float (&a)[RS], int i, int ks) {
  float v = a[4 * i];
  if (ks == 1) v = a[4 * i + 1];
  if (ks == 2) v = a[4 * i + 2];
  if (ks == 3) v = a[4 * i + 3];
  return v;
}

// LayerNorm of rows [S][D] (row stride ld; biased variance, two passes), one warp per row
template <int D>
__device__ __forceinline__ void layernorm_rows(const float* __restrict__ src, float* __restrict__ dst, int ld, int S,
                                               const float* __restrict__ g, const float* __restrict__ b, float eps, int warp,
                                               int nwarps, int lane) {
  constexpr int PER = D / 32;
  for (int r = warp; r < S; r += nwarps) {
    float v[PER];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      v[i] = src[r * ld + lane + 32 * i];
      sum += v[i];
    }
    const float mean = warp_sum(sum) * (1.f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      v[i] -= mean;
      sq += v[i] * v[i];
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / D) + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int d = lane + 32 * i;
      dst[r * ld + d] = v[i] * rstd * g[d] + b[d];
    }
  }
}

template <int DIN, int SP, int NT, int RS, int NWORK>
__global__ void __launch_bounds__(32 * (NWORK + 2), 1)
slot_attention_resident_kernel(const Args args) {
  using C = Cfg<DIN, SP, NT>;
  constexpr int NKB = C::NKB;
  constexpr int GB = C::GROUP_BYTES;
  constexpr int D = DIN;                             // slot_size == in_features (checked by the entry point)
  constexpr int LD = D + 4;                          // row stride of the update regions (floats)
  constexpr int RL = RS / 4;                         // rows per lane in the update stages
  constexpr int G1_WARP = NWORK, G2_WARP = NWORK + 1;
  constexpr int NWT = NWORK * 32;                    // worker threads
  constexpr int SOFT_WARPS = 4 * NT;
  constexpr int NV = DIN / 32;                       // float4 per lane per token (conversion)
  // output columns per lane in the MLP / projection stages (c, c + 8, ...).  Measured (profiles/README.md, round 2): a warp's
  // product is a serial chain of shared-memory and L2 latencies (~10 k cycles for K = 192 whatever runs beside it), so more
  // warps with one column each beat fewer warps with three.
  constexpr int NGC = 1;
  constexpr int GW0 = SOFT_WARPS;                    // first GRU warp (the softmax warps carry no GRU state through the attend phase)
  const SdbSlotAttentionResident& p = args.p;
  constexpr int M = 2 * D;                           // mlp_hidden_size == 2 * slot_size (checked by the entry point)
  constexpr int LDM = M + 4;                         // row stride of the MLP hidden rows
  const int S = p.S, N = p.N, ldq = p.ldq;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* xop = smem_raw;                           // NT tiles: [16 token groups][2 planes][NKB][8 tokens][128 B]
  uint8_t* scr = xop + NT * C::TILE_BYTES;           // attend: qop | aop ; update: four regions of rsz bytes
  uint8_t* qop = scr;                                // [NKB][q_hi SP rows | q_lo SP rows][128 B]
  uint8_t* aop = scr + C::QBYTES;                    // NT x [2 token blocks][a_hi SP rows | a_lo SP rows][128 B]
  Ctl& ctl = *reinterpret_cast<Ctl*>(scr + C::SCR);
  float* R0 = reinterpret_cast<float*>(scr);                     // own partial U [S][DIN] | MLP hidden [S][LDM] (spans regions 0 and 1)
  float* R2 = reinterpret_cast<float*>(scr + 2 * args.rsz);      // U (cluster sum) | LayerNorm(h') | split-K scratch | new slots, all [S][LD]
  float* R3 = reinterpret_cast<float*>(scr + 3 * args.rsz);      // h' | LayerNorm_q(slots) [S][LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank(), CL = cluster_nctarank();
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int tok0 = rank * NT * TILE;
  const int ntok = max(0, min(N - tok0, NT * TILE));
  const int nchunks = (ntok + CT - 1) / CT;
  const bool is_worker = warp < NWORK;
  const int JH = D / CL;                             // hidden units / slot channels per CTA
  const int MC = M / CL;                             // MLP hidden columns per CTA
  const uint32_t scr_s = smem_u32(scr), ctl_s = smem_u32(&ctl);
  const int j8 = lane & 7, ks = lane >> 3;           // update stages: column within the warp's group, K split (rows ks, ks + 4, ... in the epilogues)

  if ((smem_u32(smem_raw) & 1023u) != 0) {
    if (tid == 0) printf("sdb200: slot_attention_resident: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }

  // TMA chunk loads of sample b (this CTA's token range); issued by one thread
  auto issue_loads = [&](long long b) {
    for (int c = 0; c < nchunks; ++c) {
      const int rows = min(CT, ntok - c * CT);
      const uint32_t bytes = (uint32_t)rows * DIN * 4;
      mbar_arrive_expect_tx(&ctl.sfull[c], bytes);
      bulk_load(xop + c * C::CHUNK_BYTES, p.x + ((size_t)b * N + tok0 + c * CT) * DIN, bytes, &ctl.sfull[c]);
    }
  };

  // ------------------------------------------------------------------ prologue
  if (warp == G1_WARP && lane == 0) {
    for (int i = 0; i < 2 * CPT; ++i) mbar_init(&ctl.sfull[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl.xfull[i], NWORK);
      mbar_init(&ctl.lfull[i], 1);
      mbar_init(&ctl.afull[i], 4);
    }
    mbar_init(&ctl.ufull, 1);
    fence_mbar_init();
    if (cid < p.B) issue_loads(cid);                 // the feature stream starts before the rest of the prologue
  }
  if (warp == G2_WARP) {
    tmem_alloc(&ctl.tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  if (is_worker) {                                   // padding slot rows of the q operand stay zero for the whole kernel
    for (int i = tid; i < C::QBYTES / 16; i += NWT) reinterpret_cast<uint4*>(qop)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 32) ctl.cb[tid] = 0.f;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;
  cluster_barrier();                                 // every CTA of the cluster is resident before the first DSMEM access

  // GRU state of warp GW0 + jg (hidden units 8 jg .. 8 jg + 7 of this CTA's slice), rows ks, ks + 4, ... per lane: the hidden-side
  // projection gh = slots W_hh^T + b_hh (r | z | n) and the previous slots.  Both are taken when the slots are produced,
  // one stage BEFORE the attend phase whose operands overwrite the regions, so the GRU stage only adds the input side.
  float ghv[3][RL], hprev[RL];
#pragma unroll
  for (int i = 0; i < RL; ++i) ghv[0][i] = ghv[1][i] = ghv[2][i] = hprev[i] = 0.f;
  const int gw = warp - GW0;                         // GRU unit of this warp (valid: 0 <= gw < JH / 8)
  const bool gru_warp = gw >= 0 && gw < JH / 8;
  const int jl = 8 * gw + j8;                        // hidden unit within the CTA's slice ...
  const int jgl = (int)rank * JH + (gru_warp ? jl : 0);   // ... and in the layer

  uint32_t itc = 0, smp = 0;                         // iteration / sample counters (mbarrier phase parities)
  const long long t_entry = clock64();
  int nstamp = 0;
#define SR_T()                                                                                   \
  do {                                                                                           \
    if (args.dbg && blockIdx.x == 0 && tid == 0 && nstamp < 120) args.dbg[1 + nstamp++] = clock64() - t_entry; \
  } while (0)
#define SR_TW(k)                                                                                               \
  do {                                                                                                         \
    if (args.dbg && blockIdx.x == 0 && lane == 0 && itc == 0 && is_worker) args.dbg[128 + warp * 4 + (k)] = clock64() - t_entry; \
  } while (0)

  for (long long b = cid; b < p.B; b += ncl, ++smp) {
    // it = -1: the initial slots are projected to the q operand of iteration 0 and the tiles are converted
    for (int it = -1; it < p.iterations; ++it) {
      const bool last = it == p.iterations - 1;
      if (it < 0) {
        if (is_worker) {
          const float* s_in = p.slots_in + (size_t)b * S * D;
          for (int i = tid; i < S * D; i += NWT) {
            const int r = i / D;
            R2[r * LD + (i - r * D)] = s_in[i];
          }
        }
      } else {
        // ================================================================ attend
        if (warp == G1_WARP) {
          if (lane == 0) {
            fence_proxy_async_all();
            tc_fence_after();
            const uint32_t id1 = umma_idesc_f16(TILE, 2 * SP);   // A, B K-major
            const uint32_t x0 = smem_u32(xop), q0 = smem_u32(qop);
            for (int t = 0; t < NT; ++t) {
              // logits[128 tokens x 2SP] = X[128 x DIN] * [Q_hi ; Q_lo]^T for X = hi plane, then lo plane
              if (it == 0) {
                mbar_spin(&ctl.xfull[t], smp & 1);
                tc_fence_after();
              }
              const uint32_t xs = x0 + t * C::TILE_BYTES;
              const uint32_t d_tmem = tmem + 64u * t;
              const uint64_t dxh0 = desc_k_sw128(xs, GB), dxl0 = desc_k_sw128(xs + NKB * 1024, GB);
              const uint64_t dq0 = desc_k_sw128(q0, 1024);
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
              umma_commit(&ctl.lfull[t]);
            }
          }
          __syncwarp();
        } else if (warp == G2_WARP) {
          if (lane == 0) {
            const uint32_t id2 = umma_idesc_f16(TILE, 2 * SP) | (1u << 15);     // A MN-major (features^T), B K-major
            const uint32_t x0 = smem_u32(xop), a0 = smem_u32(aop);
            for (int t = 0; t < NT; ++t) {
              // per 64-channel block kb: D_kb[128 x 2SP] += [X_hi^T ; X_lo^T][(64+64) x 128 tokens] * [A_hi ; A_lo]^T
              mbar_spin(&ctl.afull[t], itc & 1);
              tc_fence_after();
              const uint32_t xs = x0 + t * C::TILE_BYTES;
              const uint64_t da0 = desc_k_sw128(a0 + t * C::ABYTES, 1024);
              const uint64_t dx0 = desc_mn_sw128(xs, NKB * 1024, GB);
              const uint32_t first = t ? 1u : 0u;
#pragma unroll
              for (int kb = 0; kb < NKB; ++kb) {
                const uint32_t d_tmem = tmem + 128u + uint32_t(2 * SP * kb);
#pragma unroll
                for (int k2 = 0; k2 < TILE / 16; ++k2) {
                  const uint64_t xadv = uint64_t((kb * 1024 + k2 * 2 * GB) >> 4);   // 16 tokens = two 8-token groups
                  const uint64_t aadv = uint64_t(((k2 >> 2) * (2 * SP * 128) + (k2 & 3) * 32) >> 4);
                  umma_f16(d_tmem, dx0 + xadv, da0 + aadv, id2, k2 ? 1u : first);
                }
              }
            }
            umma_commit(&ctl.ufull);
          }
          __syncwarp();
        } else if (warp < SOFT_WARPS) {
          // ---- softmax over slots; thread <-> token <-> TMEM lane; warps 4t..4t+3 own tile t
          const int t = warp >> 2, wq = warp & 3;
          const int r = wq * 32 + lane;
          const uint32_t lane_addr = tmem + (uint32_t(wq * 32) << 16);
          const int colx = r & 63, odd = lane & 1;
          const uint32_t a_lane = (r >> 6) * (2 * SP * 128) + (odd ? 4 * 128 : 0) + (((colx >> 3) ^ (odd ? 4 : 0)) << 4) +
                                  ((colx & 7) >> 1) * 4;
          const uint32_t sel_send = odd ? 0x5410u : 0x7632u;
          const uint32_t sel_hi = odd ? 0x3254u : 0x5410u, sel_lo = odd ? 0x3276u : 0x7610u;
          mbar_wait_warp(&ctl.lfull[t], itc & 1, lane);
          tc_fence_after();
          float l[SP];
          tmem_ld_folded<SP>(lane_addr + 64u * t, l);
          const int n = tok0 + t * TILE + r;
          const bool valid = (t * TILE + r) < ntok;
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
          if (last && p.seg_mask && valid) {
            float* mrow = p.seg_mask + ((size_t)b * S) * N + n;
#pragma unroll
            for (int s = 0; s < SP; ++s)
              if (s < S) mrow[(size_t)s * N] = l[s];
          }
          uint8_t* abase = aop + t * C::ABYTES;
#pragma unroll
          for (int s = 0; s < SP; ++s) l[s] = (valid && s < S) ? l[s] + p.attn_eps : 0.f;   // a = attn + eps (zero for padding)
#pragma unroll
          for (int pq = 0; pq < SP / 2; ++pq) {
            const int s = (pq & 3) + 8 * (pq >> 2);                      // slots s and s + 4
            uint32_t H, L;
            split2(l[s] * ASCALE, l[s + 4] * ASCALE, H, L);
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, __byte_perm(H, L, sel_send), 1);
            uint8_t* dst = abase + ((a_lane ^ ((s & 3) << 4)) + s * 128);
            *reinterpret_cast<uint32_t*>(dst) = __byte_perm(H, recv, sel_hi);
            *reinterpret_cast<uint32_t*>(dst + SP * 128) = __byte_perm(L, recv, sel_lo);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl.afull[t]);
          SR_T();                                      // it+0: softmax of tile 0 done (a operand written)
#pragma unroll
          for (int s = 0; s < SP; ++s) {
            const float w = warp_sum(l[s]);
            if (lane == 0) ctl.cs_scr[warp * SP + s] = w;
          }
          named_barrier(1, SOFT_WARPS * 32);
          if (tid < SP) {
            float cs = 0.f;
            for (int w = 0; w < SOFT_WARPS; ++w) cs += ctl.cs_scr[w * SP + tid];
            ctl.cs_part[tid] = cs;
          }
          if (warp < 4) {
            // ---- drain U^T: block kb, TMEM lane r < 64: hi-plane part of channel 64 kb + r; lane 64 + r: its lo-plane part
            mbar_wait_warp(&ctl.ufull, itc & 1, lane);
            tc_fence_after();
            SR_T();                                    // it+1: weighted sums complete
#pragma unroll 1
            for (int kb = 0; kb < NKB; ++kb) {
              float u[SP];
              tmem_ld_folded<SP>(lane_addr + 128u + uint32_t(2 * SP * kb), u);
              float* dst = R0 + kb * 64 + (r & 63);
              if (r >= 64) {
#pragma unroll
                for (int s = 0; s < SP; ++s)
                  if (s < S) dst[s * DIN] = u[s];
              }
              named_barrier(3, 128);
              if (r < 64) {
#pragma unroll
                for (int s = 0; s < SP; ++s)
                  if (s < S) dst[s * DIN] += u[s];
              }
            }
            tc_fence_before();
          }
        }
        tc_fence_before();
        SR_T();
        cluster_barrier();                                               // (1) partial sums of every CTA are in its R0 / cs_part
        SR_T();                                                          // it+2,3: drained | barrier 1

        if (last && warp == G1_WARP && lane == 0 && b + ncl < p.B) issue_loads(b + ncl);   // tiles are free: next sample's features

        // ================================================================ U = sum over CTAs / (ascale * column sums): reduce-scatter + all-gather
        if (is_worker) {
          const int DC = DIN / CL, q4 = DC / 4;
          for (int i = tid; i < S * q4; i += NWT) {
            const int s = i / q4, c = (int)rank * DC + 4 * (i % q4);
            float cs = 0.f;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t q = 0; q < CL; ++q) {
              cs += ld_cluster_f32(mapa(ctl_s + (uint32_t)offsetof(Ctl, cs_part) + 4u * s, q));
              const float4 v = ld_cluster_v4(mapa(scr_s + 4u * (s * DIN + c), q));
              a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            const float inv = 1.f / (cs * ASCALE);
            a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
            for (uint32_t q = 0; q < CL; ++q) st_cluster_v4(mapa(scr_s + 2u * args.rsz + 4u * (s * LD + c), q), a);
          }
        }
        SR_T();
        cluster_barrier();                                               // (2) U complete in every CTA
        SR_T();                                                          // it+4,5: reduce | barrier 2

        // ================================================================ GRU: this CTA's JH hidden units, input side + gates
        SR_TW(0);
        if (gru_warp) {
          float bia[3];
#pragma unroll
          for (int g = 0; g < 3; ++g) bia[g] = __ldg(p.b_iv + g * D + jgl);   // issued before the product: off its critical path
          float acc[3][RS];
          const int col[3] = {jgl, D + jgl, 2 * D + jgl};
          dot_ks<RS, 3, 2, RS / 2, LD>(R2, ks, reinterpret_cast<const float4*>(p.w_iv4), 3 * D, col, 0, D / 4, acc);
          SR_TW(1);
#ifdef SR_REPEAT_GRU      // experiment: the same product again, now with its code in the instruction cache
#pragma unroll 1
          for (int rep = 0; rep < 2; ++rep) {
            float acc2[3][RS];
            dot_ks<RS, 3, 2, RS / 2, LD>(R2, ks, reinterpret_cast<const float4*>(p.w_iv4) + rep * (args.dbg ? 0 : 1), 3 * D, col, 0, D / 4, acc2);
            if (acc2[0][0] == 12345.678f) acc[0][0] += acc2[1][1];
            SR_TW(2);
          }
#endif
#pragma unroll
          for (int i = 0; i < RL; ++i) {
            const int r = ks + 4 * i;
            const float ar = sel_row<RS>(acc[0], i, ks), az = sel_row<RS>(acc[1], i, ks), an = sel_row<RS>(acc[2], i, ks);
            if (r >= S) continue;
            const float rgate = 1.f / (1.f + expf(-(ar + bia[0] + ghv[0][i])));       // GRUCell, PyTorch gate order r | z | n
            const float zgate = 1.f / (1.f + expf(-(az + bia[1] + ghv[1][i])));
            const float ngate = tanhf(an + bia[2] + rgate * ghv[2][i]);
            const float hv = (1.f - zgate) * ngate + zgate * hprev[i];
            const uint32_t dst = scr_s + 3u * args.rsz + 4u * (r * LD + jgl);
            for (uint32_t q = 0; q < CL; ++q) st_cluster_f32(mapa(dst, q), hv);
          }
        }
        SR_TW(3);
        SR_T();
        cluster_barrier();                                               // (3) h' complete in every CTA (R3)
        SR_T();                                                          // it+6,7: GRU | barrier 3

        // ================================================================ MLP hidden: relu(LN(h') W_1^T + b_1), this CTA's MC columns
        if (is_worker) {
          layernorm_rows<D>(R3, R2, LD, S, p.ln_m_g, p.ln_m_b, p.ln_m_eps, warp, NWORK, lane);
          named_barrier(2, NWT);
          for (int u = warp; u < MC / (8 * NGC); u += NWORK) {
            int col[NGC];
            float bias[NGC];
#pragma unroll
            for (int g = 0; g < NGC; ++g) {
              col[g] = (int)rank * MC + 8 * (NGC * u + g) + j8;
              bias[g] = __ldg(p.b1 + col[g]);
            }
            float acc[NGC][RS];
            dot_ks<RS, NGC, 4, RS, LD>(R2, ks, reinterpret_cast<const float4*>(p.w1_4), M, col, 0, D / 4, acc);
#pragma unroll
            for (int g = 0; g < NGC; ++g) {
#pragma unroll
              for (int i = 0; i < RL; ++i) {
                const int r = ks + 4 * i;
                const float v = fmaxf(sel_row<RS>(acc[g], i, ks) + bias[g], 0.f);
                if (r >= S) continue;
                const uint32_t dst = scr_s + 4u * (r * LDM + col[g]);
                for (uint32_t q = 0; q < CL; ++q) st_cluster_f32(mapa(dst, q), v);
              }
            }
          }
        }
        SR_T();
        cluster_barrier();                                               // (4) MLP hidden complete in every CTA (R0 | R1)
        SR_T();                                                          // it+8,9: MLP1 | barrier 4

        // ================================================================ slots = h' + y1 W_2^T + b_2, this CTA's JH channels (split-K over warps)
        if (is_worker) {
          constexpr int nq = M / 4;
          const int ngrp = JH / (8 * NGC);                               // column units of 8 NGC channels
          int ksp = 4;                                                   // warps per column unit; partial sums meet in R2
          while (ksp > 1 && (ngrp * ksp > NWORK || nq % (16 * ksp) || ksp * RS * JH * 4 > args.rsz)) --ksp;
          const int unit = warp, u = unit / ksp, kpart = unit - u * ksp;
          const bool has = unit < ngrp * ksp;                            // (ngrp <= NWORK checked by the entry point)
          int col[NGC];
          float bias2[NGC];
#pragma unroll
          for (int g = 0; g < NGC; ++g) {
            col[g] = (int)rank * JH + (has ? 8 * (NGC * u + g) + j8 : 0);
            bias2[g] = __ldg(p.b2 + col[g]);
          }
          float mine[NGC][RL];                                           // rows ks, ks + 4, ... of this lane
          if (has) {
            float acc[NGC][RS];
            const int span = nq / ksp;
            dot_ks<RS, NGC, 4, RS, LDM>(R0, ks, reinterpret_cast<const float4*>(p.w2_4), D, col, kpart * span, (kpart + 1) * span, acc);
#pragma unroll
            for (int g = 0; g < NGC; ++g)
#pragma unroll
              for (int i = 0; i < RL; ++i) {
                mine[g][i] = sel_row<RS>(acc[g], i, ks);
                if (kpart) R2[((kpart - 1) * RS + ks + 4 * i) * JH + (col[g] - (int)rank * JH)] = mine[g][i];
              }
          }
          named_barrier(2, NWT);
          if (has && !kpart) {
#pragma unroll
            for (int g = 0; g < NGC; ++g)
#pragma unroll
              for (int i = 0; i < RL; ++i) {
                const int r = ks + 4 * i;
                float v = mine[g][i];
                if (r < S)
                  for (int k = 1; k < ksp; ++k) v += R2[((k - 1) * RS + r) * JH + (col[g] - (int)rank * JH)];
                mine[g][i] = v + bias2[g] + R3[(r < S ? r : 0) * LD + col[g]];
              }
          }
          named_barrier(2, NWT);                                         // every partial sum in R2 has been consumed
          if (has && !kpart) {
            // own slice of the new slots: global result, or this CTA's R2 (the peers pull it after the cluster barrier --
            // a push could land in a peer's split-K scratch before that peer has consumed it)
            float* s_out = p.slots_out + (size_t)b * S * D;
#pragma unroll
            for (int g = 0; g < NGC; ++g)
#pragma unroll
              for (int i = 0; i < RL; ++i) {
                const int r = ks + 4 * i;
                if (r >= S) continue;
                if (last) s_out[r * D + col[g]] = mine[g][i];
                else R2[r * LD + col[g]] = mine[g][i];
              }
          }
        }
        SR_T();
        cluster_barrier();                                               // (5) new slots complete in every CTA (R2)
        SR_T();                                                          // it+10,11: MLP2 | barrier 5
        ++itc;
      }

      if (!last) {
        // ================================================================ LayerNorm_q(R2) -> R3, then this CTA's slice of the folded slot-side
        // projection, written as the fp16 hi/lo operand rows of EVERY CTA of the cluster (+ the logit bias); the slots are
        // parked in registers (the attend operands overwrite R1..R3)
        if (is_worker) {
          if (it >= 0) {                                                 // the other CTAs' channel slices of the new slots
            for (int i = tid; i < S * (D / 4); i += NWT) {
              const int r = i / (D / 4), c = 4 * (i - r * (D / 4));
              const uint32_t owner = (uint32_t)(c / JH);
              if (owner != rank)
                *reinterpret_cast<float4*>(R2 + r * LD + c) = ld_cluster_v4(mapa(scr_s + 2u * args.rsz + 4u * (r * LD + c), owner));
            }
          }
          named_barrier(2, NWT);                                         // slots (initial / new) complete in R2
          layernorm_rows<D>(R2, R3, LD, S, p.ln_q_g, p.ln_q_b, p.ln_q_eps, warp, NWORK, lane);
          named_barrier(2, NWT);
          const int DC = DIN / CL;                       // channels of the projection per CTA
          const int nun = DC / (8 * NGC);                // units of 8 NGC channels; rank 0 also owns the bias column (Din)
          if (gru_warp) {
            // hidden-side GRU projection of the NEXT iteration and the slots it will gate, into registers
            float bia[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) bia[g] = __ldg(p.b_hh + g * D + jgl);
            float acc[3][RS];
            const int col[3] = {jgl, D + jgl, 2 * D + jgl};
            dot_ks<RS, 3, 2, RS / 2, LD>(R2, ks, reinterpret_cast<const float4*>(p.w_hh4), 3 * D, col, 0, D / 4, acc);
#pragma unroll
            for (int i = 0; i < RL; ++i) {
              const int r = ks + 4 * i;
#pragma unroll
              for (int g = 0; g < 3; ++g) ghv[g][i] = sel_row<RS>(acc[g], i, ks) + bia[g];
              hprev[i] = R2[(r < S ? r : 0) * LD + jgl];
            }
          } else if ((warp < GW0 ? warp : warp - JH / 8) < nun + (rank == 0 ? 1 : 0)) {   // projection units on the other workers
            const int un = warp < GW0 ? warp : warp - JH / 8;
            const bool bias_unit = un == nun;
            int c[NGC], col[NGC];
#pragma unroll
            for (int g = 0; g < NGC; ++g) {
              c[g] = bias_unit ? DIN + 8 * g + j8 : (int)rank * DC + 8 * (NGC * un + g) + j8;
              col[g] = c[g] < ldq ? c[g] : ldq - 1;
            }
            float acc[NGC][RS];
            dot_ks<RS, NGC, 4, RS, LD>(R3, ks, reinterpret_cast<const float4*>(p.w_qa4), ldq, col, 0, D / 4, acc);
#pragma unroll
            for (int g = 0; g < NGC; ++g) {
#pragma unroll
              for (int i = 0; i < RL; ++i) {
                const int r = ks + 4 * i;                // (warp-uniform trip count: the shuffle is executed by all lanes)
                const float v = sel_row<RS>(acc[g], i, ks);
                const __half h = __float2half_rn(v);
                const __half l = __float2half_rn(v - __half2float(h));
                const uint32_t mine = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
                const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                if (r >= S) continue;
                if (bias_unit) {
                  if (g == 0 && j8 == 0)
                    for (uint32_t q = 0; q < CL; ++q) st_cluster_f32(mapa(ctl_s + (uint32_t)offsetof(Ctl, cb) + 4u * r, q), v);
                } else {
                  // even lanes store the hi-plane word of channels (c, c + 1), odd lanes the lo-plane word of (c - 1, c)
                  const int ce = c[g] & ~1;
                  const int kb = ce >> 6, cc = ce & 63;
                  const uint32_t word = (j8 & 1) ? ((other >> 16) | (mine & 0xffff0000u)) : ((mine & 0xffffu) | (other << 16));
                  const uint32_t off = kb * (2 * SP * 128) + (((cc >> 3) ^ (r & 7)) << 4) + (cc & 7) * 2 + r * 128 +
                                       ((j8 & 1) ? SP * 128 : 0);
                  for (uint32_t q = 0; q < CL; ++q) st_cluster_u32(mapa(scr_s + off, q), word);
                }
              }
            }
          }
        }
        fence_proxy_async_all();                         // operand rows were written through the generic proxy (also remotely)
        SR_T();
        cluster_barrier();                                               // (6) q operand of the next iteration in place
        SR_T();                                                          // it+12,13: project_q | barrier 6
      }

      if (it < 0) {
        // ---------------------------------------------------------------- LayerNorm + fp16 split of the resident tiles, in place
        if (is_worker) {
          const int sub = lane >> 3, j = lane & 7;
          auto tok_in_group = [&](int q) { return 4 * (sub & 1) + (sub >> 1) + 2 * q; };
          for (int t = 0; t < NT; ++t) {
            int g = t * GROUPS + ((warp - t * GROUPS) % NWORK + NWORK) % NWORK;     // first group of tile t owned by this warp
            for (; g < (t + 1) * GROUPS; g += NWORK) {
              const int c = g >> 2;
              uint8_t* grp = xop + g * GB;               // fp32 rows in, UMMA atoms out
              float4 v[2][NV];
              if (c < nchunks) mbar_wait_warp(&ctl.sfull[c], smp & 1, lane);      // CTA-uniform branch
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int tg = tok_in_group(q);
                const bool valid = (g * 8 + tg) < ntok;
                const float* row = reinterpret_cast<const float*>(grp) + tg * DIN + 4 * j;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                  v[q][k] = *reinterpret_cast<const float4*>(row + 32 * k);       // stale bytes if !valid: zeroed, no branch
                  if (!valid) v[q][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
              __syncwarp();                              // all fp32 rows of the group are in registers
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
                const float rstd = rsqrtf(sq * (1.f / DIN) + p.ln_in_eps);        // zero rows stay zero
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
            if (lane == 0) mbar_arrive(&ctl.xfull[t]);
          }
        }
        SR_T();                                          // conversion done (this warp)
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == G2_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
  cluster_barrier();                                 // no CTA leaves while a peer may still address its shared memory
}

// ---------------------------------------------------------------- host side
static long long* g_dbg = nullptr;

struct Geometry {
  int nt, cl, rs, nwork, rsz, sp;
};

static bool geometry(int64_t N, int64_t S, int64_t Din, int64_t D, int64_t M, Geometry& g) {
  if (D != Din || M != 2 * D || !(Din == 128 || Din == 192 || Din == 256)) return false;
  if (S < 1 || S > 32 || N < 1) return false;
  g.sp = S <= 16 ? 16 : 32;
  g.rs = S <= 8 ? 8 : S <= 12 ? 12 : S <= 16 ? 16 : S <= 24 ? 24 : 32;
  g.nwork = g.sp == 16 ? 18 : 14;
  const int tile_bytes = GROUPS * 8 * (int)Din * 4;
  const int qbytes = (int)(Din / 64) * 2 * g.sp * 128, abytes = 2 * 2 * g.sp * 128;
  auto scr_of = [&](int nt) { return ((MAX_SMEM - 1024 - nt * tile_bytes) / 1024) * 1024; };
  auto rsz_of = [&](int cl) {
    int r = (int)(S * (D + 4) * 4);                                         // rows are D + 4 floats apart (bank spread)
    const int y = (int)(S * (M + 4) * 4 + 1) / 2;                           // MLP hidden [S][M + 4] spans regions 0 and 1
    r = r > y ? r : y;
    const int gh = 3 * g.rs * (int)(D / cl) * 4;                            // GRU hidden-side scratch in region 0
    r = r > gh ? r : gh;
    r = r > qbytes / 2 ? r : qbytes / 2;      // peers push q operand rows into [0, QBYTES) while regions 2 / 3 are still read
    return (r + 15) / 16 * 16;
  };
  // NT = 2 (Din <= 192) keeps twice the tokens per CTA; fall back to one tile per CTA when the regions do not fit
  for (int nt = (Din <= 192 ? 2 : 1); nt >= 1; --nt) {
    const int scr = scr_of(nt);
    if (scr < qbytes + nt * abytes) continue;
    const int cl = N <= 4 * nt * TILE ? 4 : 8;
    if (N > (int64_t)cl * nt * TILE) continue;
    // one GRU warp per 8 hidden units after the softmax warps; the projection units (+ bias unit) take the other workers
    if (D % (8 * cl) || 4 * nt + D / (8 * cl) > g.nwork || 2 * (D / (8 * cl)) + 1 > g.nwork || M % 16) continue;
    const int rsz = rsz_of(cl);
    if (4 * rsz > scr) continue;
    if (3 * rsz + g.rs * (int)(D + 4) * 4 > scr + 1024) continue;   // rows S..RS-1 of region 3 are read (and discarded)
    g.nt = nt;
    g.cl = cl;
    g.rsz = rsz;
    return true;
  }
  return false;
}

// cluster launch configuration + the number of clusters of g.cl CTAs that are resident at the same time (cached)
template <int DIN, int SP, int NT, int RS, int NWORK>
static int configure(const Geometry& g, cudaStream_t st, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int& waves_of) {
  using C = Cfg<DIN, SP, NT>;
  auto kern = slot_attention_resident_kernel<DIN, SP, NT, RS, NWORK>;
  static int max_clusters[2] = {0, 0};               // per cluster size 4 / 8
  static bool attr = false;
  if (!attr) {
    SDB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  cfg = cudaLaunchConfig_t{};
  cfg.blockDim = dim3(32 * (NWORK + 2));
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = g.cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int& mc = max_clusters[g.cl == 8];
  if (mc == 0) {
    cfg.gridDim = dim3(g.cl * 64);
    int n = 0;
    SDB_CHECK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    SDB_REQUIRE(n > 0, "sdb_slot_attention_resident: no cluster of %d CTAs fits the device", g.cl);
    mc = n;
  }
  waves_of = mc;
  return 0;
}

// a == nullptr: only report the number of concurrently resident clusters (= samples per wave) in *wave
template <int DIN, int SP, int NT, int RS, int NWORK>
static int launch(const SdbSlotAttentionResident* a, const Geometry& g, cudaStream_t st, int* wave) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  int mc = 0;
  const int rc = configure<DIN, SP, NT, RS, NWORK>(g, st, cfg, at, mc);
  if (rc) return rc;
  if (wave) *wave = mc;
  if (!a) return 0;
  const int64_t ncl = a->B < mc ? a->B : mc;
  cfg.gridDim = dim3((unsigned)(ncl * g.cl));
  Args args;
  args.p = *a;
  args.rsz = g.rsz;
  args.dbg = g_dbg;
  SDB_CHECK(cudaLaunchKernelEx(&cfg, slot_attention_resident_kernel<DIN, SP, NT, RS, NWORK>, args));
  return 0;
}

static int dispatch(const SdbSlotAttentionResident* a, int Din, const Geometry& g, cudaStream_t st, int* wave) {
#define SR_CASE(DD, SPP, NTT, RSS, NW)                                                        \
  if (Din == DD && g.sp == SPP && g.nt == NTT && g.rs == RSS) return launch<DD, SPP, NTT, RSS, NW>(a, g, st, wave);
  SR_CASE(192, 16, 2, 12, 18)
#ifndef SR_FAST_BUILD
  SR_CASE(192, 16, 2, 8, 18) SR_CASE(192, 16, 1, 8, 18) SR_CASE(192, 16, 1, 12, 18)
  SR_CASE(192, 16, 1, 16, 18) SR_CASE(192, 32, 1, 24, 14) SR_CASE(192, 32, 1, 32, 14)
  SR_CASE(128, 16, 2, 8, 18) SR_CASE(128, 16, 2, 12, 18) SR_CASE(128, 16, 2, 16, 18)
  SR_CASE(256, 16, 1, 8, 18) SR_CASE(256, 16, 1, 12, 18) SR_CASE(256, 16, 1, 16, 18)
#endif
#undef SR_CASE
  set_error("sdb_slot_attention_resident: no kernel instance for Din=%d rows=%d (tiles %d)", Din, g.rs, g.nt);
  return SDB_ERR_INVALID;
}

}  // namespace sr
}  // namespace sdb

using namespace sdb;

/* debug: device buffer of 256 int64 receiving the timeline (SM cycles since kernel entry) of thread 0 of CTA 0 */
extern "C" int sdb_slot_attention_resident_debug(void* buf) {
  sr::g_dbg = reinterpret_cast<long long*>(buf);
  return 0;
}

extern "C" int sdb_slot_attention_resident_supported(int64_t N, int64_t S, int64_t Din, int64_t D, int64_t M) {
  sr::Geometry g;
  return sr::geometry(N, S, Din, D, M, g) ? 1 : 0;
}

extern "C" int sdb_slot_attention_resident(const SdbSlotAttentionResident* a, void* stream) {
  SDB_REQUIRE(a && a->x && a->slots_in && a->slots_out && a->w_iv4 && a->b_iv && a->w_hh4 && a->b_hh && a->ln_m_g &&
                  a->ln_m_b && a->w1_4 && a->b1 && a->w2_4 && a->b2 && a->ln_q_g && a->ln_q_b && a->w_qa4,
              "sdb_slot_attention_resident: null argument");
  SDB_REQUIRE(a->B > 0 && a->iterations > 0, "sdb_slot_attention_resident: bad B=%lld iterations=%d", (long long)a->B,
              a->iterations);
  sr::Geometry g;
  SDB_REQUIRE(sr::geometry(a->N, a->S, a->Din, a->D, a->M, g),
              "sdb_slot_attention_resident: unsupported geometry N=%d S=%d Din=%d D=%d M=%d", a->N, a->S, a->Din, a->D, a->M);
  SDB_REQUIRE(a->ldq > a->Din && a->ldq % 4 == 0, "sdb_slot_attention_resident: ldq=%d must be a multiple of 4 and > in_features",
              a->ldq);
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0, "sdb_slot_attention_resident: x must be 16-byte aligned");
  return sr::dispatch(a, a->Din, g, as_stream(stream), nullptr);
}

/* samples the device processes at the same time (resident clusters) for this geometry; 0 = unsupported / query failed */
extern "C" int64_t sdb_slot_attention_resident_wave(int64_t N, int64_t S, int64_t Din, int64_t D, int64_t M) {
  sr::Geometry g;
  if (!sr::geometry(N, S, Din, D, M, g)) return 0;
  int wave = 0;
  if (sr::dispatch(nullptr, (int)Din, g, nullptr, &wave)) return 0;
  return wave;
}
