// Kernels of the slot transition function (TransformerPredictor, /root/reference/slotdiffusion/video_based/models/
// predictor.py:20-44: nn.TransformerEncoder over the [B, S, D] slots between two video frames).
// Its linears are sdb_gemm launches; what is left is multi-head self-attention over S <= 32 slot tokens per sample (head
// dims 32 / 48 / 64 -- far below a tensor-core tile) with nn.MultiheadAttention's dropout on the attention probabilities,
// and the dropout + residual adds of nn.TransformerEncoderLayer.  All dropout masks are counter based (seed, element
// index, device-resident step counter: common.cuh dropout_scale), so the backward regenerates them.
#include "common.cuh"

namespace sdb {

constexpr int TA_MAXS = 32;

__device__ __forceinline__ unsigned long long mix_seed(unsigned long long seed, const unsigned long long* seed_dev) {
  return seed + (seed_dev ? *seed_dev * 0x9E3779B97F4A7C15ull : 0ull);
}

// One warp per (head, sample); lane i owns query row i.  qkv rows [B*S, ld]: q at column h*DH, k at D + h*DH, v at 2D + h*DH
// (the fused in_proj of nn.MultiheadAttention).  out [B*S, D] fp32.
template <int DH>
__global__ void __launch_bounds__(32) token_attention_kernel(const float* __restrict__ qkv, int64_t ld, float* __restrict__ out,
                                                             int S, int heads, float scale, float drop_p,
                                                             unsigned long long seed, const unsigned long long* seed_dev) {
  __shared__ float ks[TA_MAXS][DH + 1], vs[TA_MAXS][DH + 1];
  const int h = blockIdx.x, lane = threadIdx.x;
  const int64_t b = blockIdx.y;
  const int D = heads * DH;
  const unsigned long long sd = mix_seed(seed, seed_dev);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int j = 0; j < S; ++j)
    for (int c = lane; c < DH; c += 32) {
      const float* row = qkv + (b * S + j) * ld + h * DH + c;
      ks[j][c] = row[D];
      vs[j][c] = row[2 * D];
    }
  __syncwarp();
  if (lane >= S) return;
  float q[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) q[c] = qkv[(b * S + lane) * ld + h * DH + c];
  float s[TA_MAXS];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < TA_MAXS; ++j) {
    float a = 0.f;
    if (j < S) {
#pragma unroll
      for (int c = 0; c < DH; ++c) a = fmaf(q[c], ks[j][c], a);
      a *= scale;
      mx = fmaxf(mx, a);
    }
    s[j] = a;
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < TA_MAXS; ++j) {
    s[j] = j < S ? expf(s[j] - mx) : 0.f;
    sum += s[j];
  }
  const float inv = 1.f / sum;
  float o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < TA_MAXS; ++j) {
    if (j >= S) continue;
    float p = s[j] * inv;
    if (drop_p > 0.f)
      p *= dropout_scale(sd, (unsigned long long)(((b * heads + h) * S + lane) * S + j), drop_p, inv_keep);
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = fmaf(p, vs[j][c], o[c]);
  }
#pragma unroll
  for (int c = 0; c < DH; ++c) out[(b * S + lane) * D + h * DH + c] = o[c];
}

// Backward by recomputation.  Phase A (lane = query row i): softmax statistics, rowsum(P o dP), dQ.  Phase B (lane = key row
// j): dK, dV -- the same probabilities recomputed column-wise, so no cross-lane reductions are needed.
template <int DH>
__global__ void __launch_bounds__(32) token_attention_bwd_kernel(const float* __restrict__ qkv, int64_t ld,
                                                                 const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                 int S, int heads, float scale, float drop_p,
                                                                 unsigned long long seed, const unsigned long long* seed_dev) {
  __shared__ float qs[TA_MAXS][DH + 1], ks[TA_MAXS][DH + 1], vs[TA_MAXS][DH + 1], ds[TA_MAXS][DH + 1];
  __shared__ float s_m[TA_MAXS], s_l[TA_MAXS], s_d[TA_MAXS];
  const int h = blockIdx.x, lane = threadIdx.x;
  const int64_t b = blockIdx.y;
  const int D = heads * DH;
  const unsigned long long sd = mix_seed(seed, seed_dev);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int j = 0; j < S; ++j)
    for (int c = lane; c < DH; c += 32) {
      const float* row = qkv + (b * S + j) * ld + h * DH + c;
      qs[j][c] = row[0];
      ks[j][c] = row[D];
      vs[j][c] = row[2 * D];
      ds[j][c] = dout[(b * S + j) * D + h * DH + c];
    }
  __syncwarp();
  const int64_t mbase = ((b * heads + h) * S) * (int64_t)S;
  if (lane < S) {
    // ---- phase A
    const int i = lane;
    float s[TA_MAXS];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TA_MAXS; ++j) {
      float a = 0.f;
      if (j < S) {
#pragma unroll
        for (int c = 0; c < DH; ++c) a = fmaf(qs[i][c], ks[j][c], a);
        a *= scale;
        mx = fmaxf(mx, a);
      }
      s[j] = a;
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < TA_MAXS; ++j) {
      s[j] = j < S ? expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
    const float inv = 1.f / sum;
    float dp[TA_MAXS];
    float dsum = 0.f;
#pragma unroll
    for (int j = 0; j < TA_MAXS; ++j) {
      float a = 0.f;
      if (j < S) {
#pragma unroll
        for (int c = 0; c < DH; ++c) a = fmaf(ds[i][c], vs[j][c], a);
        if (drop_p > 0.f) a *= dropout_scale(sd, (unsigned long long)(mbase + i * S + j), drop_p, inv_keep);
        s[j] *= inv;                                  // p_ij
        dsum = fmaf(s[j], a, dsum);
      }
      dp[j] = a;                                      // gradient w.r.t. p_ij
    }
    float dq[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) dq[c] = 0.f;
#pragma unroll
    for (int j = 0; j < TA_MAXS; ++j) {
      if (j >= S) continue;
      const float dsc = s[j] * (dp[j] - dsum) * scale;
#pragma unroll
      for (int c = 0; c < DH; ++c) dq[c] = fmaf(dsc, ks[j][c], dq[c]);
    }
#pragma unroll
    for (int c = 0; c < DH; ++c) dqkv[(b * S + i) * (3 * D) + h * DH + c] = dq[c];
    s_m[i] = mx;
    s_l[i] = inv;
    s_d[i] = dsum;
  }
  __syncwarp();
  if (lane < S) {
    // ---- phase B
    const int j = lane;
    float dk[DH], dv[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) dk[c] = dv[c] = 0.f;
    for (int i = 0; i < S; ++i) {
      float a = 0.f, g = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        a = fmaf(qs[i][c], ks[j][c], a);
        g = fmaf(ds[i][c], vs[j][c], g);
      }
      const float p = expf(a * scale - s_m[i]) * s_l[i];
      const float mk = drop_p > 0.f ? dropout_scale(sd, (unsigned long long)(mbase + i * S + j), drop_p, inv_keep) : 1.f;
      const float pm = p * mk;                                        // probability after dropout
      const float dsc = p * (g * mk - s_d[i]) * scale;
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        dv[c] = fmaf(pm, ds[i][c], dv[c]);
        dk[c] = fmaf(dsc, qs[i][c], dk[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < DH; ++c) {
      dqkv[(b * S + j) * (3 * D) + D + h * DH + c] = dk[c];
      dqkv[(b * S + j) * (3 * D) + 2 * D + h * DH + c] = dv[c];
    }
  }
}

// out = (res ? res : 0) + x * dropout_scale(idx)      (the backward is the same call on dy with res = NULL)
__global__ void dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ out,
                                   int64_t n, float drop_p, unsigned long long seed, const unsigned long long* seed_dev) {
  const unsigned long long sd = mix_seed(seed, seed_dev);
  const float inv_keep = 1.f / (1.f - drop_p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i] * dropout_scale(sd, (unsigned long long)i, drop_p, inv_keep);
    out[i] = res ? res[i] + v : v;
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_token_attention_supported(int64_t S, int64_t dh) {
  return (S >= 1 && S <= TA_MAXS && (dh == 32 || dh == 48 || dh == 64)) ? 1 : 0;
}

extern "C" int sdb_token_attention(const float* qkv, int64_t ld, float* out, int64_t B, int64_t S, int heads, int dh,
                                   float scale, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  SDB_REQUIRE(qkv && out && B > 0 && heads > 0, "sdb_token_attention: bad args");
  SDB_REQUIRE(sdb_token_attention_supported(S, dh), "sdb_token_attention: unsupported S=%lld head dim %d", (long long)S, dh);
  SDB_REQUIRE(ld >= 3 * heads * dh && B <= 65535 && drop_p >= 0.f && drop_p < 1.f, "sdb_token_attention: bad ld / B / p");
  dim3 grid((unsigned)heads, (unsigned)B);
  const unsigned long long* sdv = reinterpret_cast<const unsigned long long*>(seed_dev);
  cudaStream_t st = as_stream(stream);
  if (dh == 32) token_attention_kernel<32><<<grid, 32, 0, st>>>(qkv, ld, out, (int)S, heads, scale, drop_p, seed, sdv);
  else if (dh == 48) token_attention_kernel<48><<<grid, 32, 0, st>>>(qkv, ld, out, (int)S, heads, scale, drop_p, seed, sdv);
  else token_attention_kernel<64><<<grid, 32, 0, st>>>(qkv, ld, out, (int)S, heads, scale, drop_p, seed, sdv);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_token_attention_bwd(const float* qkv, int64_t ld, const float* dout, float* dqkv, int64_t B, int64_t S,
                                       int heads, int dh, float scale, float drop_p, uint64_t seed,
                                       const uint64_t* seed_dev, void* stream) {
  SDB_REQUIRE(qkv && dout && dqkv && B > 0 && heads > 0, "sdb_token_attention_bwd: bad args");
  SDB_REQUIRE(sdb_token_attention_supported(S, dh), "sdb_token_attention_bwd: unsupported S=%lld head dim %d", (long long)S, dh);
  SDB_REQUIRE(ld >= 3 * heads * dh && B <= 65535 && drop_p >= 0.f && drop_p < 1.f, "sdb_token_attention_bwd: bad ld / B / p");
  dim3 grid((unsigned)heads, (unsigned)B);
  const unsigned long long* sdv = reinterpret_cast<const unsigned long long*>(seed_dev);
  cudaStream_t st = as_stream(stream);
  if (dh == 32) token_attention_bwd_kernel<32><<<grid, 32, 0, st>>>(qkv, ld, dout, dqkv, (int)S, heads, scale, drop_p, seed, sdv);
  else if (dh == 48) token_attention_bwd_kernel<48><<<grid, 32, 0, st>>>(qkv, ld, dout, dqkv, (int)S, heads, scale, drop_p, seed, sdv);
  else token_attention_bwd_kernel<64><<<grid, 32, 0, st>>>(qkv, ld, dout, dqkv, (int)S, heads, scale, drop_p, seed, sdv);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_dropout_add(const float* x, const float* res, float* out, int64_t n, float drop_p, uint64_t seed,
                               const uint64_t* seed_dev, void* stream) {
  SDB_REQUIRE(x && out && n > 0 && drop_p >= 0.f && drop_p < 1.f, "sdb_dropout_add: bad args");
  int64_t blocks = cdiv(n, 256);
  if (blocks > 4096) blocks = 4096;
  dropout_add_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, res, out, n, drop_p, seed,
                                                                      reinterpret_cast<const unsigned long long*>(seed_dev));
  SDB_LAUNCH_CHECK();
  return 0;
}
