// Slot Attention inner loop (slot_attention.py:84-91, sa_diffusion.py:45-56): one fused pass over the tokens
//   logits -> softmax over slots -> (seg mask) -> +eps -> weighted sum of V and column sums
// The spatial renormalisation a / sum_n a is folded out algebraically (updates = (sum_n a v) / (sum_n a)), so
// the N x S attention matrix is never materialised in HBM and K,V are streamed exactly once per iteration.
//
// Grid (chunks, B): each CTA owns a contiguous chunk of one sample's tokens and streams [T x 2D] fp32 tiles of
// the fused k|v projection with cp.async.bulk (TMA bulk copies, mbarrier completion, double-buffered).
// Partial sums go to a small workspace; slot_attend_finalize_kernel reduces the chunks, divides by the column
// sums and emits the GRU input in packed GEMM-operand format.
#include "common.cuh"
#include "ptx.cuh"

namespace sdb {

constexpr int SA_THREADS = 256;
constexpr int SA_SP = 36;   // padded slot stride of the per-tile attention scratch (16-B aligned rows)

template <int D, int T>
struct SaSmem {
  static constexpr int KV_STRIDE = 2 * D + 4;   // floats; (2D+4) mod 32 == 4 -> conflict-free 128-bit row reads
  static constexpr int Q_STRIDE = D + 4;
  float kv[2][T][KV_STRIDE];
  float a[T][SA_SP];      // logits -> a = softmax + eps
  float p[T][SA_SP];      // softmax (seg mask values)
  float cs[32];           // column sums of a
  uint64_t bar[2];
  // q [S][Q_STRIDE] follows (dynamic S)
};

template <int D, int T, int SMAX>
__global__ void __launch_bounds__(SA_THREADS, 1)
slot_attend_kernel(const float* __restrict__ kv, const float* __restrict__ q, float* __restrict__ seg_mask,
                   float* __restrict__ part_upd, float* __restrict__ part_cs, int64_t N, int S, int chunks,
                   float scale, float eps) {
  using SM = SaSmem<D, T>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  float* sq = reinterpret_cast<float*>(smem_raw + sizeof(SM));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x;
  const int64_t b = blockIdx.y;
  // token range of this chunk (multiples of T except the last)
  const int64_t tiles_total = (N + T - 1) / T;
  const int64_t tiles_per_chunk = (tiles_total + chunks - 1) / chunks;
  const int64_t n_begin = chunk * tiles_per_chunk * T;
  const int64_t n_end = min(N, n_begin + tiles_per_chunk * T);
  const int ntiles = n_begin < n_end ? (int)((n_end - n_begin + T - 1) / T) : 0;

  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < S * (D / 4); i += SA_THREADS) {
    const int s = i / (D / 4), c = i % (D / 4);
    float4 t = reinterpret_cast<const float4*>(q + (b * S + s) * D)[c];
    t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
    reinterpret_cast<float4*>(sq + s * SM::Q_STRIDE)[c] = t;
  }
  if (tid < 32) sm.cs[tid] = 0.f;
  __syncthreads();

  auto issue_tile = [&](int ti) {   // warp 0
    const int buf = ti & 1;
    const int64_t n0 = n_begin + (int64_t)ti * T;
    const int rows = (int)min((int64_t)T, n_end - n0);
    if (lane == 0) mbar_arrive_expect_tx(&sm.bar[buf], (uint32_t)rows * 2 * D * 4);
    __syncwarp();
    for (int r = lane; r < rows; r += 32)
      bulk_load(&sm.kv[buf][r][0], kv + ((b * N + n0 + r) * 2 * D), 2 * D * 4, &sm.bar[buf]);
  };
  if (warp == 0 && ntiles > 0) issue_tile(0);

  // update accumulators: thread <-> (channel pair, token group)
  constexpr int PAIRS = D / 2;
  constexpr int DGROUPS = SA_THREADS / PAIRS;       // 2
  constexpr int TOK_PER_G = T / DGROUPS;
  const int pair = tid % PAIRS, dgroup = tid / PAIRS;
  const bool upd_active = dgroup < DGROUPS;
  float acc[SMAX][2];
#pragma unroll
  for (int s = 0; s < SMAX; ++s) acc[s][0] = acc[s][1] = 0.f;
  float cs_lane = 0.f;   // lane <-> slot partial column sum (softmax step)

  // logits: thread <-> (token, slot group)
  constexpr int SGROUPS = SA_THREADS / T;
  const int tok = tid % T, sgroup = tid / T;
  const int spt = (S + SGROUPS - 1) / SGROUPS;      // slots per thread

  for (int ti = 0; ti < ntiles; ++ti) {
    const int buf = ti & 1;
    const int64_t n0 = n_begin + (int64_t)ti * T;
    const int rows = (int)min((int64_t)T, n_end - n0);
    if (warp == 0 && ti + 1 < ntiles) issue_tile(ti + 1);   // buffer (ti+1)&1 was released by the sync ending tile ti-1
    mbar_wait(&sm.bar[buf], (ti >> 1) & 1);

    // ---- A: logits[tok][s] = (scale q[s]) . k[tok]
    for (int sb = 0; sb < spt; sb += 4) {
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      const int s0 = sgroup * spt + sb;
      const float* q0 = sq + min(s0 + 0, S - 1) * SM::Q_STRIDE;
      const float* q1 = sq + min(s0 + 1, S - 1) * SM::Q_STRIDE;
      const float* q2 = sq + min(s0 + 2, S - 1) * SM::Q_STRIDE;
      const float* q3 = sq + min(s0 + 3, S - 1) * SM::Q_STRIDE;
      const float* kr = &sm.kv[buf][tok][0];
      const int cnt = min(4, spt - sb);
#pragma unroll 4
      for (int c = 0; c < D; c += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + c);
        const float4 a0 = *reinterpret_cast<const float4*>(q0 + c);
        d0 += kk.x * a0.x + kk.y * a0.y + kk.z * a0.z + kk.w * a0.w;
        if (cnt > 1) {
          const float4 a1 = *reinterpret_cast<const float4*>(q1 + c);
          d1 += kk.x * a1.x + kk.y * a1.y + kk.z * a1.z + kk.w * a1.w;
        }
        if (cnt > 2) {
          const float4 a2 = *reinterpret_cast<const float4*>(q2 + c);
          d2 += kk.x * a2.x + kk.y * a2.y + kk.z * a2.z + kk.w * a2.w;
        }
        if (cnt > 3) {
          const float4 a3 = *reinterpret_cast<const float4*>(q3 + c);
          d3 += kk.x * a3.x + kk.y * a3.y + kk.z * a3.z + kk.w * a3.w;
        }
      }
      if (s0 + 0 < S && cnt > 0) sm.a[tok][s0 + 0] = d0;
      if (s0 + 1 < S && cnt > 1) sm.a[tok][s0 + 1] = d1;
      if (s0 + 2 < S && cnt > 2) sm.a[tok][s0 + 2] = d2;
      if (s0 + 3 < S && cnt > 3) sm.a[tok][s0 + 3] = d3;
    }
    __syncthreads();

    // ---- B: softmax over slots (lane <-> slot), warp w handles tokens w, w+8, ...
    for (int t = warp; t < T; t += SA_THREADS / 32) {
      const bool valid = t < rows;
      const float x = (lane < S && valid) ? sm.a[t][lane] : -INFINITY;
      const float m = warp_max(x);
      const float e = (lane < S && valid) ? expf(x - m) : 0.f;
      const float sum = warp_sum(e);
      const float p = valid ? e / sum : 0.f;
      const float a = (lane < S && valid) ? p + eps : 0.f;
      if (lane < SA_SP) {
        sm.a[t][lane] = a;      // zero beyond S and for invalid tokens
        sm.p[t][lane] = p;
      }
      cs_lane += a;
    }
    __syncthreads();

    // ---- C: seg mask (last iteration only), coalesced over tokens: seg_mask[b][s][n]
    if (seg_mask) {
      for (int i = tid; i < S * T; i += SA_THREADS) {
        const int s = i / T, t = i % T;
        if (t < rows) seg_mask[(b * S + s) * N + n0 + t] = sm.p[t][s];
      }
    }
    // ---- D: acc[s][d] += a[t][s] * v[t][d]
    if (upd_active) {
      const int t_begin = dgroup * TOK_PER_G;
      const int t_end = min(t_begin + TOK_PER_G, rows);   // stale smem rows beyond `rows` may hold NaN patterns
#pragma unroll 2
      for (int t = t_begin; t < t_end; ++t) {
        const float2 vv = *reinterpret_cast<const float2*>(&sm.kv[buf][t][D + 2 * pair]);
#pragma unroll
        for (int s4 = 0; s4 < SMAX; s4 += 4) {
          const float4 aa = *reinterpret_cast<const float4*>(&sm.a[t][s4]);
          acc[s4 + 0][0] += aa.x * vv.x; acc[s4 + 0][1] += aa.x * vv.y;
          acc[s4 + 1][0] += aa.y * vv.x; acc[s4 + 1][1] += aa.y * vv.y;
          acc[s4 + 2][0] += aa.z * vv.x; acc[s4 + 2][1] += aa.z * vv.y;
          acc[s4 + 3][0] += aa.w * vv.x; acc[s4 + 3][1] += aa.w * vv.y;
        }
      }
    }
    __syncthreads();   // everyone done with kv[buf], a, p -> buffer may be refilled
  }

  // ---- reduce the token groups through shared memory (reuse kv[0]) and write the chunk partials
  atomicAdd(&sm.cs[lane], cs_lane);
  float* red = &sm.kv[0][0][0];   // needs DGROUPS*S*D floats <= 2*32*256 = 16K floats = 64 KB  (kv[0] is >= T*KV_STRIDE)
  if (upd_active && dgroup > 0) {
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
      if (s < S) {
        red[((dgroup - 1) * SMAX + s) * D + 2 * pair] = acc[s][0];
        red[((dgroup - 1) * SMAX + s) * D + 2 * pair + 1] = acc[s][1];
      }
    }
  }
  __syncthreads();
  if (upd_active && dgroup == 0) {
    float* dst = part_upd + ((b * chunks + chunk) * S) * D;
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
      if (s < S) {
        float v0 = acc[s][0], v1 = acc[s][1];
        for (int g = 1; g < DGROUPS; ++g) {
          v0 += red[((g - 1) * SMAX + s) * D + 2 * pair];
          v1 += red[((g - 1) * SMAX + s) * D + 2 * pair + 1];
        }
        *reinterpret_cast<float2*>(dst + s * D + 2 * pair) = make_float2(v0, v1);
      }
    }
  }
  if (tid < S) part_cs[(b * chunks + chunk) * S + tid] = sm.cs[tid];
}

// updates[b,s,:] = sum_chunks part_upd / sum_chunks part_cs  -> packed (+ fp32)
__global__ void slot_attend_finalize_kernel(const float* __restrict__ part_upd, const float* __restrict__ part_cs,
                                            __half* __restrict__ out, float* __restrict__ upd32,
                                            float* __restrict__ colsum, int64_t BS, int S, int D, int chunks) {
  const int64_t bs = blockIdx.x;
  const int64_t b = bs / S;
  const int s = (int)(bs % S);
  float cs = 0.f;
  for (int c = 0; c < chunks; ++c) cs += part_cs[(b * chunks + c) * S + s];
  const float inv = 1.f / cs;
  if (colsum && threadIdx.x == 0) colsum[bs] = cs;
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

static int sa_chunks(int64_t B, int64_t N, int T) {
  int64_t tiles = cdiv(N, T);
  int64_t c = num_sms() / (B > 0 ? B : 1);
  if (c < 1) c = 1;
  if (c > tiles) c = tiles;
  if (c > 64) c = 64;
  return (int)c;
}
static int sa_tile(int64_t D) { return D <= 128 ? 64 : 32; }

template <int D, int T, int SMAX>
static int launch_attend(const float* kv, const float* q, float* seg_mask, float* part_upd, float* part_cs, int64_t B,
                         int64_t N, int S, int chunks, float scale, float eps, cudaStream_t st) {
  const size_t smem = sizeof(SaSmem<D, T>) + (size_t)S * (D + 4) * 4;
  auto kern = slot_attend_kernel<D, T, SMAX>;
  static size_t attr = 0;
  if (smem > attr) {
    SDB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid(chunks, (unsigned)B);
  kern<<<grid, SA_THREADS, smem, st>>>(kv, q, seg_mask, part_upd, part_cs, N, S, chunks, scale, eps);
  SDB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sdb

using namespace sdb;

extern "C" int64_t sdb_slot_attend_workspace(int64_t B, int64_t N, int64_t S, int64_t D) {
  const int chunks = sa_chunks(B, N, sa_tile(D));
  return (B * chunks * S * D + B * chunks * S) * (int64_t)sizeof(float);
}

extern "C" int sdb_slot_attend(const float* kv, const float* q, float* seg_mask, void* upd_packed, float* upd32,
                               float* work, int64_t B, int64_t N, int64_t S, int64_t D, float scale, float eps,
                               void* stream) {
  return sdb_slot_attend_train(kv, q, seg_mask, upd_packed, upd32, nullptr, work, B, N, S, D, scale, eps, stream);
}

extern "C" int sdb_slot_attend_train(const float* kv, const float* q, float* seg_mask, void* upd_packed, float* upd32,
                                     float* colsum, float* work, int64_t B, int64_t N, int64_t S, int64_t D, float scale,
                                     float eps, void* stream) {
  SDB_REQUIRE(kv && q && upd_packed && work, "sdb_slot_attend: null argument");
  SDB_REQUIRE(B > 0 && B <= 65535 && N > 0, "sdb_slot_attend: bad B=%lld N=%lld", (long long)B, (long long)N);
  SDB_REQUIRE(S >= 1 && S <= 32, "sdb_slot_attend: num_slots=%lld must be in 1..32", (long long)S);
  SDB_REQUIRE(D == 64 || D == 128 || D == 192 || D == 256, "sdb_slot_attend: slot_size=%lld unsupported (64/128/192/256)",
              (long long)D);
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(kv) & 15) == 0, "sdb_slot_attend: kv must be 16-byte aligned");
  const int T = sa_tile(D);
  const int chunks = sa_chunks(B, N, T);
  float* part_upd = work;
  float* part_cs = work + B * chunks * S * D;
  cudaStream_t st = as_stream(stream);
  int rc = 0;
#define SA_CASE(DD, TT)                                                                                          \
  if (D == DD) {                                                                                                 \
    rc = (S <= 16) ? launch_attend<DD, TT, 16>(kv, q, seg_mask, part_upd, part_cs, B, N, (int)S, chunks, scale,  \
                                               eps, st)                                                          \
                   : launch_attend<DD, TT, 32>(kv, q, seg_mask, part_upd, part_cs, B, N, (int)S, chunks, scale,  \
                                               eps, st);                                                         \
  }
  SA_CASE(64, 64) else SA_CASE(128, 64) else SA_CASE(192, 32) else SA_CASE(256, 32)
#undef SA_CASE
  if (rc) return rc;
  slot_attend_finalize_kernel<<<(unsigned)(B * S), 128, 0, st>>>(part_upd, part_cs, (__half*)upd_packed, upd32, colsum,
                                                                 B * S, (int)S, (int)D, chunks);
  SDB_LAUNCH_CHECK();
  return 0;
}
