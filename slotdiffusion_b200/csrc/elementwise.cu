// Memory-bound producers of GEMM operands and the small fused pointwise kernels of the hot path.
// All are HBM/L2-bandwidth kernels: coalesced 128-bit accesses, grid sized from the element count.
#include "common.cuh"

namespace sdb {

static inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = cdiv(work_items, threads);
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------ weights
SDB_DEFINE_PACK_MODE_SETTER(set_pack_mode_elementwise)

__global__ void pack_weight_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t n4, int64_t plane,
                                   int fmt, int wexp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(w)[i];
    store_split4_w(out, out + plane, i * 4, v, fmt, wexp);
  }
}

// [Cout][Cin][3][3] -> [Cout][tap][Cin], four input channels per thread (SDB_FMT_F8C needs 4-byte e4m3 stores)
__global__ void pack_weight_conv3_v4_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t Cout,
                                            int64_t Cin, int fmt, int wexp) {
  const int64_t c4n = Cin / 4, total4 = Cout * 9 * c4n, plane = Cout * 9 * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = (i % c4n) * 4;
    const int64_t tap = (i / c4n) % 9;
    const int64_t o = i / (9 * c4n);
    const float* src = w + (o * Cin + c) * 9 + tap;
    store_split4_w(out, out + plane, (o * 9 + tap) * Cin + c, make_float4(src[0], src[9], src[18], src[27]), fmt, wexp);
  }
}

// [Cout][Cin][3][3] -> [Cout][tap][Cin]
__global__ void pack_weight_conv3_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t Cout,
                                         int64_t Cin) {
  const int64_t total = Cout * 9 * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % Cin;
    const int64_t tap = (i / Cin) % 9;
    const int64_t o = i / (9 * Cin);
    const float v = w[(o * Cin + c) * 9 + tap];
    __half h, l;
    split_f16(v, h, l);
    out[i] = h;
    out[total + i] = l;
  }
}

// ------------------------------------------------------------------ rows
template <int PM>
__global__ void pack_rows_kernel(const float* __restrict__ x, int64_t ldx, __half* __restrict__ out, int64_t M,
                                 int64_t K, int act) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int64_t k4 = K / 4;
  const int64_t total = M * k4;
  const int64_t plane = M * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / k4, j = i % k4;
    float4 v = *reinterpret_cast<const float4*>(x + m * ldx + j * 4);
    if (act == 1) {
      v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
    } else if (act == 2) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    store_split4(out, out + plane, m * K + j * 4, v, pmode);
  }
}

// ------------------------------------------------------------------ LayerNorm (one warp per row)
template <int PM>
__global__ void __launch_bounds__(256, 4) layernorm_pack_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, __half* __restrict__ out,
                                      float* __restrict__ y, int64_t M, int C) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t plane = M * (int64_t)C;
  for (int64_t row = blockIdx.x * (int64_t)warps_per_block + (threadIdx.x >> 5); row < M;
       row += (int64_t)gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    const int c4 = C / 4;
    float4 v[4];   // C <= 512
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        v[j] = xr[idx];
        s += v[j].x + v[j].y + v[j].z + v[j].w;
      }
    }
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        ss += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        const float4 gm = reinterpret_cast<const float4*>(gamma)[idx];
        const float4 bt = reinterpret_cast<const float4*>(beta)[idx];
        float4 o;
        o.x = (v[j].x - mean) * rstd * gm.x + bt.x;
        o.y = (v[j].y - mean) * rstd * gm.y + bt.y;
        o.z = (v[j].z - mean) * rstd * gm.z + bt.z;
        o.w = (v[j].w - mean) * rstd * gm.w + bt.w;
        if (out) store_split4(out, out + plane, row * C + idx * 4, o, pmode);
        if (y) reinterpret_cast<float4*>(y + row * C)[idx] = o;
      }
    }
  }
}

// ------------------------------------------------------------------ GroupNorm
// stats: one CTA per (b, group); two-pass in registers is not possible (HW*cpg up to 32K elements), so use
// a shifted single pass (Welford-free: subtract the first element as pivot) followed by an exact second pass
// from L2 for the variance -- the activation of one sample-group (<= 128 KB) stays L2/L1 resident.
__global__ void groupnorm_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                       float* __restrict__ stats, int64_t HW, int G, float eps, int vec) {
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int C = C1 + C2;
  const int cpg = C / G;
  const int c_begin = g * cpg;
  const int64_t n = HW * cpg;
  __shared__ float red[32];
  __shared__ float s_mean;
  auto load = [&](int64_t i) -> float {
    const int64_t p = i / cpg;
    const int c = c_begin + int(i % cpg);
    return (c < C1) ? x1[(b * HW + p) * C1 + c] : x2[(b * HW + p) * C2 + (c - C1)];
  };
  // vec (launcher): groups of a multiple of 4 channels that do not straddle x1 | x2, 16-byte aligned rows, HW * cpg / 4 < 2^31
  // -- float4 loads and 32-bit index arithmetic (the scalar form runs a 64-bit division and a modulo per ELEMENT)
  const int q4 = cpg >> 2;
  const unsigned n4 = vec ? (unsigned)(HW * q4) : 0u;
  const bool in1 = c_begin < C1;
  const float* vbase = (in1 ? x1 + c_begin : x2 + (c_begin - C1)) + (int64_t)b * HW * (in1 ? C1 : C2);
  const int vld = in1 ? C1 : C2;
  auto load4 = [&](unsigned i) -> float4 {
    const unsigned p = i / (unsigned)q4, q = i - p * (unsigned)q4;
    return *reinterpret_cast<const float4*>(vbase + (int64_t)p * vld + q * 4);
  };
  float s = 0.f;
  if (vec) {
#pragma unroll 4
    for (unsigned i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = load4(i);
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += load(i);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) s_mean = t / n;
  }
  __syncthreads();
  const float mean = s_mean;
  float ss = 0.f;
  if (vec) {
#pragma unroll 4
    for (unsigned i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = load4(i);
      const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
      ss += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const float d = load(i) - mean;
      ss += d * d;
    }
  }
  ss = warp_sum(ss);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      stats[(b * G + g) * 2 + 0] = mean;
      stats[(b * G + g) * 2 + 1] = rsqrtf(t / n + eps);
    }
  }
}

template <int PM>
__global__ void groupnorm_apply_pack_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                            int C2, const float* __restrict__ stats,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            __half* __restrict__ out, int64_t B, int64_t HW, int G, int silu) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int C = C1 + C2;
  const int c4n = C / 4;
  const int cpg = C / G;
  const int64_t total = B * HW * c4n;
  const int64_t plane = B * HW * (int64_t)C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = int(i % c4n) * 4;
    const int64_t row = i / c4n;   // b*HW + p
    const int64_t b = row / HW;
    float4 v = (c < C1) ? *reinterpret_cast<const float4*>(x1 + row * C1 + c)
                        : *reinterpret_cast<const float4*>(x2 + row * C2 + (c - C1));
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    const float4 bt = *reinterpret_cast<const float4*>(beta + c);
    float r[4] = {v.x, v.y, v.z, v.w};
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w};
    const float btv[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;   // cpg may not be a multiple of 4 (e.g. 28)
      const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
      float o = (r[j] - mean) * rstd * gmv[j] + btv[j];
      r[j] = silu == 1 ? silu_f(o) : (silu == 2 ? fmaxf(o, 0.f) : o);
    }
    store_split4(out, out + plane, row * C + c, make_float4(r[0], r[1], r[2], r[3]), pmode);
  }
}

// Block end of the ResNet18-GN encoder (resnet.py:72-90): out = relu(GN_h(h) + identity), identity = the block input
// (idn, stats_i == NULL), its normalised 1x1 projection (GN_i(idn), the `downsample` branch), or nothing (idn == NULL: the
// stem conv1 -> bn1 -> relu, resnet.py:288-291).  Emits the fp32 rows (next block's identity) and / or the packed operand
// of the next convolution in one pass.
template <int PM>
__global__ void groupnorm_add_relu_kernel(const float* __restrict__ h, const float* __restrict__ stats_h,
                                          const float* __restrict__ gamma_h, const float* __restrict__ beta_h,
                                          const float* __restrict__ idn, const float* __restrict__ stats_i,
                                          const float* __restrict__ gamma_i, const float* __restrict__ beta_i,
                                          float* __restrict__ out, __half* __restrict__ out_packed, int64_t B, int64_t HW,
                                          int C, int G) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int c4n = C / 4, cpg = C / G;
  const int64_t total = B * HW * c4n;
  const int64_t plane = B * HW * (int64_t)C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = int(i % c4n) * 4;
    const int64_t row = i / c4n;
    const int64_t b = row / HW;
    const float4 hv = *reinterpret_cast<const float4*>(h + row * C + c);
    const float4 gm = *reinterpret_cast<const float4*>(gamma_h + c);
    const float4 bt = *reinterpret_cast<const float4*>(beta_h + c);
    float r[4] = {hv.x, hv.y, hv.z, hv.w};
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
    float id[4] = {0.f, 0.f, 0.f, 0.f};
    if (idn) {
      const float4 iv = *reinterpret_cast<const float4*>(idn + row * C + c);
      id[0] = iv.x; id[1] = iv.y; id[2] = iv.z; id[3] = iv.w;
      if (stats_i) {
        const float4 gi = *reinterpret_cast<const float4*>(gamma_i + c);
        const float4 bi = *reinterpret_cast<const float4*>(beta_i + c);
        const float giv[4] = {gi.x, gi.y, gi.z, gi.w}, biv[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = (c + j) / cpg;
          id[j] = (id[j] - stats_i[(b * G + g) * 2]) * stats_i[(b * G + g) * 2 + 1] * giv[j] + biv[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;
      const float z = (r[j] - stats_h[(b * G + g) * 2]) * stats_h[(b * G + g) * 2 + 1] * gmv[j] + btv[j];
      r[j] = fmaxf(z + id[j], 0.f);
    }
    const float4 o = make_float4(r[0], r[1], r[2], r[3]);
    if (out) *reinterpret_cast<float4*>(out + row * C + c) = o;
    if (out_packed) store_split4(out_packed, out_packed + plane, row * C + c, o, pmode);
  }
}

// GroupNorm apply (+SiLU) + pack, statistics either given ([B,G,2] mean/rstd) or derived on the fly from the
// per-(sample, 4-channel block) partial sums that the producing GEMM epilogues accumulated (sdb_gemm `gsum`).
// Grid (chunks, B); every thread owns 4 fixed channels (scale/shift live in registers) and strides over rows.
// (256, 6): the launcher sizes the grid for six resident CTAs per SM.  ACT (0 none / 1 SiLU / 2 ReLU) and DROP are
// compile-time: the kernel is instruction-bound, and the UNet-inference instance <PM, 1, false> carries none of the other forms
template <int PM, int ACT, bool DROP>
__global__ void __launch_bounds__(256, 6)
groupnorm_apply_pack_fused_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ gsum1,
                                  const float* __restrict__ x2, int C2, const float* __restrict__ gsum2,
                                  const float* __restrict__ stats, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, __half* __restrict__ out, int64_t B, int HW, int G,
                                  float eps, int silu, int rows_per_chunk, float drop_p, unsigned long long seed,
                                  const unsigned long long* __restrict__ seed_dev) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  if (seed_dev) seed += *seed_dev * 0x9E3779B97F4A7C15ull;   // per-step counter living in device memory (graph replays)
  __shared__ float s_mean[64], s_rstd[64];
  const int C = C1 + C2, c4n = C >> 2, cpg = C / G;
  const int64_t b = blockIdx.y;
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    if (stats) {
      s_mean[g] = stats[(b * G + g) * 2];
      s_rstd[g] = stats[(b * G + g) * 2 + 1];
    } else {
      float s = 0.f, ss = 0.f;
      for (int kb = g * cpg / 4; kb < (g + 1) * cpg / 4; ++kb) {
        const float* p = (kb * 4 < C1) ? gsum1 + (b * (C1 >> 2) + kb) * 2 : gsum2 + (b * (C2 >> 2) + kb - (C1 >> 2)) * 2;
        s += p[0];
        ss += p[1];
      }
      const float inv_n = 1.f / ((float)HW * cpg);
      const float mean = s * inv_n;
      const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
      s_mean[g] = mean;
      s_rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  const int tx = threadIdx.x % c4n, ty = threadIdx.x / c4n, rpb = blockDim.x / c4n;
  if (ty >= rpb) return;   // launch width was rounded up for the statistics threads
  const int c = tx * 4;
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
  const float4 bt = *reinterpret_cast<const float4*>(beta + c);
  float sc[4], sh[4];
  {
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;
      sc[j] = s_rstd[g] * gmv[j];
      sh[j] = btv[j] - s_mean[g] * sc[j];
    }
  }
  const bool from1 = c < C1;
  const float* src = from1 ? x1 + c : x2 + (c - C1);
  const int ld = from1 ? C1 : C2;
  const int64_t plane = B * (int64_t)HW * C;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(HW, r_begin + rows_per_chunk);
#pragma unroll 4
  for (int r = r_begin + ty; r < r_end; r += rpb) {
    const int64_t row = b * HW + r;
    const float4 v = *reinterpret_cast<const float4*>(src + row * ld);
    float o[4] = {v.x * sc[0] + sh[0], v.y * sc[1] + sh[1], v.z * sc[2] + sh[2], v.w * sc[3] + sh[3]};
    if (ACT == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = silu_f(o[j]);
    } else if (ACT == 2) {        // ReLU (ResNet18-GN encoder: conv -> GN -> ReLU, resnet.py:76-78)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    if (DROP) {
      const float inv_keep = 1.f / (1.f - drop_p);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] *= dropout_scale(seed, (unsigned long long)(row * C + c + j), drop_p, inv_keep);
    }
    store_split4(out, out + plane, row * C + c, make_float4(o[0], o[1], o[2], o[3]), pmode);
  }
}

// GroupNorm partial sums of an activation that no GEMM epilogue produced (the 3 -> C input convolution): per (sample,
// 4-channel block) sum and sum of squares in the `gsum` format, so every consumer takes the fused statistics path
// instead of a full two-pass statistics kernel each.  Grid (chunks, B); thread owns 4 channels and strides over rows.
__global__ void __launch_bounds__(256)
channel_block_sums_kernel(const float* __restrict__ x, int C, float* __restrict__ gsum, int HW, int rows_per_chunk) {
  __shared__ float red[256][2];
  const int c4n = C >> 2;
  const int64_t b = blockIdx.y;
  const int tx = threadIdx.x % c4n, ty = threadIdx.x / c4n, rpb = blockDim.x / c4n;
  float s = 0.f, ss = 0.f;
  if (ty < rpb) {
    const int r_begin = blockIdx.x * rows_per_chunk, r_end = min(HW, r_begin + rows_per_chunk);
#pragma unroll 4
    for (int r = r_begin + ty; r < r_end; r += rpb) {
      const float4 v = *reinterpret_cast<const float4*>(x + (b * HW + r) * C + tx * 4);
      s += (v.x + v.y) + (v.z + v.w);
      ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  }
  red[threadIdx.x][0] = s;
  red[threadIdx.x][1] = ss;
  __syncthreads();
  if (threadIdx.x < c4n) {
    for (int y = 1; y < rpb; ++y) {
      s += red[y * c4n + threadIdx.x][0];
      ss += red[y * c4n + threadIdx.x][1];
    }
    float* p = gsum + (b * c4n + threadIdx.x) * 2;
    atomicAdd(p, s);
    atomicAdd(p + 1, ss);
  }
}

// partial sums -> [B,G,2] (mean, rstd) for consumers that want final statistics (output head)
__global__ void groupnorm_finalize_kernel(const float* __restrict__ gsum1, int C1, const float* __restrict__ gsum2,
                                          int C2, float* __restrict__ stats, int64_t B, int HW, int G, float eps) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B * G) return;
  const int64_t b = i / G;
  const int g = (int)(i % G);
  const int cpg = (C1 + C2) / G;
  float s = 0.f, ss = 0.f;
  for (int kb = g * cpg / 4; kb < (g + 1) * cpg / 4; ++kb) {
    const float* p = (kb * 4 < C1) ? gsum1 + (b * (C1 >> 2) + kb) * 2 : gsum2 + (b * (C2 >> 2) + kb - (C1 >> 2)) * 2;
    s += p[0];
    ss += p[1];
  }
  const float inv_n = 1.f / ((float)HW * cpg);
  const float mean = s * inv_n;
  stats[i * 2] = mean;
  stats[i * 2 + 1] = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + eps);
}

// the same for sums kept in blocks of cb channels (2: the 64-channel levels with 32 groups)
__global__ void groupnorm_finalize_cb_kernel(const float* __restrict__ gsum, int C, float* __restrict__ stats, int64_t B,
                                             int HW, int G, float eps, int cb) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B * G) return;
  const int64_t b = i / G;
  const int g = (int)(i % G);
  const int cpg = C / G, nb = C / cb;
  float s = 0.f, ss = 0.f;
  for (int kb = g * cpg / cb; kb < (g + 1) * cpg / cb; ++kb) {
    s += gsum[(b * nb + kb) * 2];
    ss += gsum[(b * nb + kb) * 2 + 1];
  }
  const float inv_n = 1.f / ((float)HW * cpg);
  const float mean = s * inv_n;
  stats[i * 2] = mean;
  stats[i * 2 + 1] = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + eps);
}

// GEGLU weight layout: output row r' (chunk j = r'/32) <- a-row 16j + r'%32 (r'%32 < 16) or g-row F + 16j + r'%32 - 16
template <int PM>
__global__ void pack_weight_geglu_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t F, int64_t K) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int64_t k4 = K / 4, total = 2 * F * k4, plane = 2 * F * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k4, j = i % k4;
    const int64_t chunk = r / 32, within = r % 32;
    const int64_t srow = within < 16 ? chunk * 16 + within : F + chunk * 16 + within - 16;
    const float4 v = reinterpret_cast<const float4*>(w + srow * K)[j];
    store_split4(out, out + plane, r * K + j * 4, v, pmode);
  }
}
__global__ void permute_geglu_bias_kernel(const float* __restrict__ bsrc, float* __restrict__ out, int64_t F) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < 2 * F; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t chunk = r / 32, within = r % 32;
    out[r] = bsrc[within < 16 ? chunk * 16 + within : F + chunk * 16 + within - 16];
  }
}

// ------------------------------------------------------------------ raw NHWC packing with layout transforms
template <int PM>
__global__ void __launch_bounds__(256) pack_nhwc_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                 __half* __restrict__ out, float* __restrict__ ycat, int64_t B, int H, int W,
                                 int mode) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int C = C1 + C2;
  const int c4n = C / 4;
  const int Ho = (mode == SDB_PACK_UP2) ? 2 * H : H;
  const int Wo = (mode == SDB_PACK_UP2) ? 2 * W : W;
  const int64_t total = B * Ho * Wo * c4n;     // iterate over OUTPUT elements
  const int64_t plane = B * (int64_t)Ho * Wo * C;
  if (total < (1ll << 31)) {
    // 32-bit index arithmetic (64-bit div/mod was most of this kernel's instruction stream)
    const unsigned tot = (unsigned)total, stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) {
      const unsigned pq = i / (unsigned)c4n;
      const int c = int(i - pq * (unsigned)c4n) * 4;
      const unsigned pr = pq / (unsigned)Wo;
      const int xo = int(pq - pr * (unsigned)Wo);
      const unsigned bb = pr / (unsigned)Ho;
      const int yo = int(pr - bb * (unsigned)Ho);
      int yi = yo, xi = xo;
      if (mode == SDB_PACK_UP2) { yi = yo >> 1; xi = xo >> 1; }
      const int64_t src = ((int64_t)bb * H + yi) * W + xi;
      float4 v = (c < C1) ? *reinterpret_cast<const float4*>(x1 + src * C1 + c)
                          : *reinterpret_cast<const float4*>(x2 + src * C2 + (c - C1));
      int64_t dst;
      if (mode == SDB_PACK_PHASE2) {
        const int ph = (yo & 1) * 2 + (xo & 1);
        dst = ((((int64_t)bb * 4 + ph) * (H / 2) + (yo >> 1)) * (W / 2) + (xo >> 1)) * C + c;
      } else {
        dst = (((int64_t)bb * Ho + yo) * Wo + xo) * C + c;
      }
      store_split4(out, out + plane, dst, v, pmode);
      if (ycat) *reinterpret_cast<float4*>(ycat + src * C + c) = v;
    }
    return;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = int(i % c4n) * 4;
    int64_t p = i / c4n;
    const int xo = int(p % Wo);
    p /= Wo;
    const int yo = int(p % Ho);
    const int64_t b = p / Ho;
    int yi = yo, xi = xo;
    if (mode == SDB_PACK_UP2) { yi = yo >> 1; xi = xo >> 1; }
    const int64_t src = (b * H + yi) * W + xi;
    float4 v = (c < C1) ? *reinterpret_cast<const float4*>(x1 + src * C1 + c)
                        : *reinterpret_cast<const float4*>(x2 + src * C2 + (c - C1));
    int64_t dst;
    if (mode == SDB_PACK_PHASE2) {
      const int ph = (yo & 1) * 2 + (xo & 1);
      dst = (((b * 4 + ph) * (H / 2) + (yo >> 1)) * (W / 2) + (xo >> 1)) * C + c;
    } else {
      dst = ((b * Ho + yo) * Wo + xo) * C + c;
    }
    store_split4(out, out + plane, dst, v, pmode);
    if (ycat) *reinterpret_cast<float4*>(ycat + src * C + c) = v;
  }
}

// ------------------------------------------------------------------ GEGLU
template <int PM>
__global__ void geglu_pack_kernel(const float* __restrict__ u, __half* __restrict__ out, int64_t M, int64_t F) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int64_t f4 = F / 4;
  const int64_t total = M * f4;
  const int64_t plane = M * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / f4, j = (i % f4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(u + m * 2 * F + j);
    const float4 g = *reinterpret_cast<const float4*>(u + m * 2 * F + F + j);
    float4 o;
    o.x = a.x * gelu_erf_f(g.x);
    o.y = a.y * gelu_erf_f(g.y);
    o.z = a.z * gelu_erf_f(g.z);
    o.w = a.w * gelu_erf_f(g.w);
    store_split4(out, out + plane, m * F + j, o, pmode);
  }
}

// ------------------------------------------------------------------ timestep embedding
template <int PM>
__global__ void timestep_embedding_pack_kernel(const float* __restrict__ t, __half* __restrict__ out, int64_t B,
                                               int dim) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int half = dim / 2;
  const int64_t total = B * dim;
  // four consecutive columns per thread (dim % 8 == 0: a group never straddles the cos | sin halves), written in the
  // operand format of the current pack mode (common.cuh)
  for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < total / 4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i4 * 4;
    const int64_t b = i / dim;
    const int j = int(i % dim);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = j < half ? j + e : j + e - half;
      // freqs = exp(-ln(1e4) * f / half) in fp32 (unet/utils.py:81-84)
      const float freq = expf(-9.210340371976184f * (float)f / (float)half);
      const float arg = t[b] * freq;
      v[e] = j < half ? cosf(arg) : sinf(arg);
    }
    store_split4(out, out + total, i, make_float4(v[0], v[1], v[2], v[3]), pmode);
  }
}

// ------------------------------------------------------------------ row softmax (+ scale) -> packed operand
// one warp per row, N <= 4096: the row lives in registers (up to 32 float4 per lane)
template <int PM>
__global__ void softmax_pack_kernel(const float* __restrict__ x, int64_t ldx, float scale, float out_scale,
                                    __half* __restrict__ out, int64_t M, int N) {
  constexpr int pmode = PM;        // operand format of the consumer GEMM, chosen by the launcher (common.cuh)
  pdl_wait();       // launched with the PDL attribute (common.cuh): nothing global before this line
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int n4 = N >> 2;
  float4 v[32];
  float mx = -3.0e38f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = i * 32 + lane;
    if (j < n4) {
      v[i] = *reinterpret_cast<const float4*>(x + row * ldx + j * 4);
      v[i].x *= scale; v[i].y *= scale; v[i].z *= scale; v[i].w *= scale;
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = i * 32 + lane;
    if (j < n4) {
      v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx); v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = warp_sum(sum);
  const float inv = out_scale / sum;
  const int64_t plane = M * (int64_t)N;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = i * 32 + lane;
    if (j < n4)
      store_split4(out, out + plane, row * N + j * 4, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv), pmode);
  }
}

// ------------------------------------------------------------------ GRU gates
__global__ void gru_gates_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                 const float* __restrict__ h, float* __restrict__ hn, int64_t R, int64_t D) {
  const int64_t total = R * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D, d = i % D;
    const float* a = gi + r * 3 * D;
    const float* b = gh + r * 3 * D;
    const float rg = 1.f / (1.f + expf(-(a[d] + b[d])));
    const float zg = 1.f / (1.f + expf(-(a[D + d] + b[D + d])));
    const float ng = tanhf(a[2 * D + d] + rg * b[2 * D + d]);
    hn[i] = (1.f - zg) * ng + zg * h[i];
  }
}

// ------------------------------------------------------------------ small-channel convs
// input conv: x NCHW [B,Cin,H,W] -> y NHWC [B,H,W,Cout]; one thread per (pixel, 4 output channels)
__global__ void conv3_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ y, int64_t B, int Cin, int H,
                                int W, int Cout) {
  extern __shared__ float sw[];   // [Cin*9][Cout]
  for (int i = threadIdx.x; i < Cin * 9 * Cout; i += blockDim.x) {
    const int o = i % Cout, r = i / Cout;    // r = c*9 + tap
    sw[i] = w[o * Cin * 9 + r];
  }
  __syncthreads();
  // A thread keeps its 4 output channels and walks over pixels (blockDim is a multiple of Cout / 4: sdb_conv3_in): the only
  // divisions left are the two 32-bit ones of the pixel decomposition (the i / o4n form of this loop ran a 64-bit division per
  // element -- most of the kernel's 134 M warp instructions, ncu: profiles/README.md section 23).  The bias and the 27
  // weight float4 of the thread's channels are loop invariants read through shared memory.
  const int o4n = Cout / 4;
  const int o = int(threadIdx.x % o4n) * 4;
  const unsigned ppb = blockDim.x / o4n;                      // pixels per CTA and trip
  const unsigned npix = (unsigned)(B * H * W);
  const float4 bias4 = *reinterpret_cast<const float4*>(bias + o);
  for (unsigned pq = blockIdx.x * ppb + threadIdx.x / o4n; pq < npix; pq += gridDim.x * ppb) {
    const unsigned pr = pq / (unsigned)W;
    const int xo = int(pq - pr * (unsigned)W);
    const unsigned bb = pr / (unsigned)H;
    const int yo = int(pr - bb * (unsigned)H);
    const int64_t b = bb;
    float4 acc = bias4;
    for (int c = 0; c < Cin; ++c) {
      const float* xc = x + (b * Cin + c) * (int64_t)H * W;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        // unconditional load from a clamped address, then select: the nine loads of a channel issue back to back (behind a
        // per-tap branch they were serialised: long-scoreboard stall 8.4 per issue, ncu)
        const int yi = yo + tap / 3 - 1, xi = xo + tap % 3 - 1;
        const bool inb = yi >= 0 && yi < H && xi >= 0 && xi < W;
        const float ld = xc[min(max(yi, 0), H - 1) * W + min(max(xi, 0), W - 1)];
        const float v = inb ? ld : 0.f;
        const float4 ww = *reinterpret_cast<const float4*>(sw + (c * 9 + tap) * Cout + o);
        acc.x += v * ww.x; acc.y += v * ww.y; acc.z += v * ww.z; acc.w += v * ww.w;
      }
    }
    *reinterpret_cast<float4*>(y + (int64_t)pq * Cout + o) = acc;
  }
}

// Same conv, four adjacent output pixels per thread (W % 4 == 0): the 3 x 6 input window of a channel is loaded once
// (clamped addresses, zero select) and every weight float4 is reused by the four pixels -- 160 instead of 510 warp
// instructions per pixel; the one-pixel form above was issue-bound (134 M warp instructions for 33.5 M outputs, ncu).
__global__ void __launch_bounds__(256)
conv3_in_px4_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ y, int64_t B, int Cin, int H, int W, int Cout) {
  extern __shared__ float sw[];   // [Cin*9][Cout]
  for (int i = threadIdx.x; i < Cin * 9 * Cout; i += blockDim.x) {
    const int o = i % Cout, r = i / Cout;    // r = c*9 + tap
    sw[i] = w[o * Cin * 9 + r];
  }
  __syncthreads();
  const int o4n = Cout / 4;
  const int o = int(threadIdx.x % o4n) * 4;
  const unsigned qpb = blockDim.x / o4n;                      // pixel quads per CTA and trip
  const unsigned wq = (unsigned)W / 4, nquads = (unsigned)(B * H) * wq;
  const float4 bias4 = *reinterpret_cast<const float4*>(bias + o);
  for (unsigned qd = blockIdx.x * qpb + threadIdx.x / o4n; qd < nquads; qd += gridDim.x * qpb) {
    const unsigned pr = qd / wq;                              // b * H + yo
    const int x0 = int(qd - pr * wq) * 4;
    const unsigned bb = pr / (unsigned)H;
    const int yo = int(pr - bb * (unsigned)H);
    float4 acc[4] = {bias4, bias4, bias4, bias4};
    for (int c = 0; c < Cin; ++c) {
      const float* xc = x + ((int64_t)bb * Cin + c) * (int64_t)H * W;
      float win[3][6];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yi = yo + ky - 1;
        const bool yin = yi >= 0 && yi < H;
        const float* row = xc + min(max(yi, 0), H - 1) * W;
#pragma unroll
        for (int kx = 0; kx < 6; ++kx) {
          const int xi = x0 + kx - 1;
          const float ld = row[min(max(xi, 0), W - 1)];
          win[ky][kx] = (yin && xi >= 0 && xi < W) ? ld : 0.f;
        }
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4 ww = *reinterpret_cast<const float4*>(sw + (c * 9 + tap) * Cout + o);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float v = win[tap / 3][px + tap % 3];
          acc[px].x += v * ww.x; acc[px].y += v * ww.y; acc[px].z += v * ww.z; acc[px].w += v * ww.w;
        }
      }
    }
    float* dst = y + ((int64_t)pr * W + x0) * Cout + o;
#pragma unroll
    for (int px = 0; px < 4; ++px) *reinterpret_cast<float4*>(dst + (int64_t)px * Cout) = acc[px];
  }
}

// output head: y NCHW [B,Cout,H,W] = conv3x3(SiLU(GN(h))), h NHWC [B,H,W,C]; one warp per output pixel,
// lanes split the C channels, Cout (<= 4) accumulators per lane, warp-shuffle reduction.
template <int COUT_MAX>
__global__ void conv3_out_kernel(const float* __restrict__ h, const float* __restrict__ stats,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                                 int64_t B, int H, int W, int C, int G, int Cout) {
  extern __shared__ float sw[];   // [Cout][9][C]
  for (int i = threadIdx.x; i < Cout * 9 * C; i += blockDim.x) {
    const int c = i % C, tap = (i / C) % 9, o = i / (9 * C);
    sw[i] = w[(o * C + c) * 9 + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int cpg = C / G;
  const int64_t total = B * H * W;
  for (int64_t p = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); p < total; p += (int64_t)gridDim.x * wpb) {
    const int xo = int(p % W), yo = int((p / W) % H);
    const int64_t b = p / ((int64_t)H * W);
    float acc[COUT_MAX];
#pragma unroll
    for (int o = 0; o < COUT_MAX; ++o) acc[o] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int yi = yo + tap / 3 - 1, xi = xo + tap % 3 - 1;
      if (yi < 0 || yi >= H || xi < 0 || xi >= W) continue;
      const float* hp = h + ((b * H + yi) * W + xi) * C;
      for (int c = lane; c < C; c += 32) {
        const int g = c / cpg;
        const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
        const float a = silu_f((hp[c] - mean) * rstd * gamma[c] + beta[c]);
#pragma unroll
        for (int o = 0; o < COUT_MAX; ++o)
          if (o < Cout) acc[o] += a * sw[(o * 9 + tap) * C + c];
      }
    }
#pragma unroll
    for (int o = 0; o < COUT_MAX; ++o) {
      if (o < Cout) {
        const float s = warp_sum(acc[o]);
        if (lane == 0) y[((b * Cout + o) * H + yo) * W + xo] = s + bias[o];
      }
    }
  }
}

// Tiled form of the output head: a CTA owns R output rows of one image, evaluates SiLU(GN(h)) ONCE per input pixel of
// the R + 2 rows it needs (zero border = the conv padding) into shared memory, then every warp computes 4 adjacent
// output pixels at a time (lanes split the channels, float4 smem reads, weight reads shared by the 4 pixels).
template <int COUT_MAX>
__global__ void __launch_bounds__(256)
conv3_out_tile_kernel(const float* __restrict__ h, const float* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ y, int H, int W, int C, int G, int Cout, int R) {
  extern __shared__ __align__(16) float smem_f[];
  float* sw = smem_f;                         // [Cout][9][C]
  float* sa = smem_f + Cout * 9 * C;          // [R + 2][W + 2][C]
  const int64_t b = blockIdx.y;
  const int y0 = blockIdx.x * R;
  const int Wp = W + 2, cpg = C / G;
  // Stage 1: raw tile and weights as asynchronous copies, all in flight together
  auto cp4 = [](void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  auto cp16 = [](void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  // Index arithmetic: a warp owns (output channel, tap) rows of the weights and pixels of the tile, lanes run over the
  // channels -- no per-element division (the i % c4n, (i / c4n) % Wp, i / (c4n Wp) form of these loops was 60 % of the
  // kernel's 279 M warp instructions: ncu, profiles/README.md section 23).
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int ot = warp; ot < Cout * 9; ot += nwarps) {
    const int o = ot / 9, tap = ot - o * 9;
    for (int c = lane; c < C; c += 32) cp4(sw + ot * C + c, w + (o * C + c) * 9 + tap);
  }
  const int npix = (R + 2) * Wp;
  for (int pix = warp; pix < npix; pix += nwarps) {
    const int rr = pix / Wp, xp = pix - rr * Wp;
    const int yi = y0 - 1 + rr, xi = xp - 1;
    const bool inb = yi >= 0 && yi < H && xi >= 0 && xi < W;
    float* dst = sa + pix * C;
    const float* src = h + ((b * H + yi) * W + xi) * C;
    for (int c = lane * 4; c < C; c += 128) {
      if (inb) cp16(dst + c, src + c);
      else *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);      // zero border = the conv padding
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // Stage 2: SiLU(GN(.)) in place, once per input pixel; a lane keeps its channels, so scale / shift are loop invariants
  for (int c = lane * 4; c < C; c += 128) {
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const int g = c / cpg;                  // C / G is a multiple of 4: the four channels share a group
    const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
    const float4 sc = make_float4(rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w);
    const float4 sh = make_float4(be.x - mean * sc.x, be.y - mean * sc.y, be.z - mean * sc.z, be.w - mean * sc.w);
    for (int pix = warp; pix < npix; pix += nwarps) {
      const int rr = pix / Wp, xp = pix - rr * Wp;
      const int yi = y0 - 1 + rr, xi = xp - 1;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
        float4* pa = reinterpret_cast<float4*>(sa + pix * C + c);
        const float4 v = *pa;
        *pa = make_float4(silu_f(fmaf(v.x, sc.x, sh.x)), silu_f(fmaf(v.y, sc.y, sh.y)), silu_f(fmaf(v.z, sc.z, sh.z)),
                          silu_f(fmaf(v.w, sc.w, sh.w)));
      }
    }
  }
  __syncthreads();
  const int rows = min(R, H - y0);
  const int quads = rows * (W / 4);
  for (int qd = warp; qd < quads; qd += nwarps) {
    const int yr = qd / (W / 4), x0 = (qd % (W / 4)) * 4;
    float acc[4][COUT_MAX];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int o = 0; o < COUT_MAX; ++o) acc[px][o] = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        float4 wv[COUT_MAX];
#pragma unroll
        for (int o = 0; o < COUT_MAX; ++o)
          wv[o] = o < Cout ? *reinterpret_cast<const float4*>(sw + (o * 9 + tap) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float4 a = *reinterpret_cast<const float4*>(sa + ((yr + ky) * Wp + x0 + px + kx) * C + c);
#pragma unroll
          for (int o = 0; o < COUT_MAX; ++o)
            acc[px][o] += (a.x * wv[o].x + a.y * wv[o].y) + (a.z * wv[o].z + a.w * wv[o].w);
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int o = 0; o < COUT_MAX; ++o) {
        if (o < Cout) {
          const float sum = warp_sum(acc[px][o]);
          if (lane == 0) y[((b * Cout + o) * H + y0 + yr) * W + x0 + px] = sum + bias[o];
        }
      }
  }
}

// ------------------------------------------------------------------ DPM-Solver glue
// x0 = (x - sigma*eps)/alpha, then nearest codebook row.  One thread per latent pixel; codebook in smem.
__global__ void dpm_x0_kernel(const float* __restrict__ x, const float* __restrict__ eps, float alpha, float sigma,
                              const float* __restrict__ codebook, int ncodes, float* __restrict__ x0,
                              int* __restrict__ idx, int64_t B, int C, int64_t HW) {
  extern __shared__ float scb[];   // [ncodes][C] + [ncodes] squared norms
  float* snorm = scb + (size_t)ncodes * C;
  if (codebook) {
    for (int i = threadIdx.x; i < ncodes * C; i += blockDim.x) scb[i] = codebook[i];
    __syncthreads();
    for (int i = threadIdx.x; i < ncodes; i += blockDim.x) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += scb[i * C + c] * scb[i * C + c];
      snorm[i] = s;
    }
    __syncthreads();
  }
  const int64_t total = B * HW;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p % HW;
    float z[8];
    float zz = 0.f;
    for (int c = 0; c < C; ++c) {
      const int64_t o = (b * C + c) * HW + q;
      z[c] = (x[o] - sigma * eps[o]) / alpha;
      zz += z[c] * z[c];
    }
    if (codebook) {
      // d = |z|^2 + |e|^2 - 2 z.e  (the reference's expanded form, quantize.py:89-91); first minimum wins
      float best = INFINITY;
      int bi = 0;
      for (int j = 0; j < ncodes; ++j) {
        float dot = 0.f;
        for (int c = 0; c < C; ++c) dot += z[c] * scb[j * C + c];
        const float d = zz + snorm[j] - 2.f * dot;
        if (d < best) { best = d; bi = j; }
      }
      for (int c = 0; c < C; ++c) z[c] = scb[bi * C + c];
      if (idx) idx[p] = bi;
    }
    for (int c = 0; c < C; ++c) x0[(b * C + c) * HW + q] = z[c];
  }
}

// Three-channel latents (every shipped VQ-VAE, quantize.py:80-94): register-blocked nearest-code search.  A block stages
// the codebook as (e0, e1, e2, |e|^2) float4 rows in shared memory once; every thread scores PX pixels against each
// broadcast row (one LDS.128 feeds PX * 5 FMA-pipe operations), strict '<' in increasing code order = argmin's first
// minimum.  Same expanded distance |z|^2 + |e|^2 - 2 z.e as the reference, with the contraction order pinned by fmaf.
constexpr int VQ_PX = 5;          // 262 144 latent pixels (B = 256): 410 CTAs for the 444 resident slots = one wave (4: 512 CTAs, 1.15 waves)
constexpr int VQ_THREADS = 128;
constexpr int VQ_G = 8;          // codes per comparison group
__global__ void __launch_bounds__(VQ_THREADS)
dpm_x0_vq3_kernel(const float* __restrict__ x, const float* __restrict__ eps, float alpha, float sigma,
                  const float* __restrict__ codebook, int ncodes, float* __restrict__ x0, int* __restrict__ idx,
                  int64_t B, int64_t HW) {
  extern __shared__ __align__(16) float4 scb4[];   // [ncodes]
  for (int i = threadIdx.x; i < ncodes; i += VQ_THREADS) {
    const float e0 = codebook[i * 3], e1 = codebook[i * 3 + 1], e2 = codebook[i * 3 + 2];
    float nrm = 0.f;
    nrm += e0 * e0;
    nrm += e1 * e1;
    nrm += e2 * e2;
    scb4[i] = make_float4(e0, e1, e2, nrm);
  }
  __syncthreads();
  const int64_t total = B * HW;
  const int64_t p0 = (int64_t)blockIdx.x * (VQ_THREADS * VQ_PX) + threadIdx.x;
  float z0[VQ_PX], z1[VQ_PX], z2[VQ_PX], zz[VQ_PX], best[VQ_PX];
  int bi[VQ_PX];
#pragma unroll
  for (int i = 0; i < VQ_PX; ++i) {
    const int64_t p = p0 + (int64_t)i * VQ_THREADS;
    z0[i] = z1[i] = z2[i] = 0.f;
    if (p < total) {
      const int64_t b = p / HW, q = p % HW;
      const int64_t o = b * 3 * HW + q;
      z0[i] = (x[o] - sigma * eps[o]) / alpha;
      z1[i] = (x[o + HW] - sigma * eps[o + HW]) / alpha;
      z2[i] = (x[o + 2 * HW] - sigma * eps[o + 2 * HW]) / alpha;
    }
    float t = 0.f;
    t += z0[i] * z0[i];
    t += z1[i] * z1[i];
    t += z2[i] * z2[i];
    zz[i] = t;
    best[i] = INFINITY;
    bi[i] = 0;
  }
  // Codes in groups of VQ_G: per pixel only the group minimum is compared with the running best (strict '<', groups in
  // increasing order: the FIRST group that holds the global minimum wins); the winning group is rescored afterwards to
  // find its first code at that distance -- the same expression, so bit-identical.  3 instead of 3 * VQ_G compare / select
  // instructions per group and pixel (the per-code form spent 48 % of its issue slots on them: ncu, profiles/README 23).
  auto dist = [&](const float4& e, int i) {
    const float dot = fmaf(z2[i], e.z, fmaf(z1[i], e.y, z0[i] * e.x));
    return fmaf(-2.f, dot, zz[i] + e.w);
  };
  const int ngroups = ncodes / VQ_G;
  for (int g = 0; g < ngroups; ++g) {
    float m[VQ_PX];
#pragma unroll
    for (int i = 0; i < VQ_PX; ++i) m[i] = INFINITY;
#pragma unroll
    for (int jj = 0; jj < VQ_G; ++jj) {
      const float4 e = scb4[g * VQ_G + jj];
#pragma unroll
      for (int i = 0; i < VQ_PX; ++i) m[i] = fminf(m[i], dist(e, i));
    }
#pragma unroll
    for (int i = 0; i < VQ_PX; ++i)
      if (m[i] < best[i]) { best[i] = m[i]; bi[i] = g; }
  }
#pragma unroll
  for (int i = 0; i < VQ_PX; ++i) {                 // first code of the winning group at the best distance
    const int g = bi[i];
    int first = g * VQ_G;
    for (int jj = VQ_G - 1; jj >= 0; --jj)
      if (dist(scb4[g * VQ_G + jj], i) == best[i]) first = g * VQ_G + jj;
    bi[i] = ngroups > 0 ? first : 0;
  }
  for (int j = ngroups * VQ_G; j < ncodes; ++j) {   // codebook sizes that are not a multiple of VQ_G
    const float4 e = scb4[j];
#pragma unroll
    for (int i = 0; i < VQ_PX; ++i) {
      const float d = dist(e, i);
      if (d < best[i]) { best[i] = d; bi[i] = j; }
    }
  }
#pragma unroll
  for (int i = 0; i < VQ_PX; ++i) {
    const int64_t p = p0 + (int64_t)i * VQ_THREADS;
    if (p < total) {
      const int64_t b = p / HW, q = p % HW;
      const int64_t o = b * 3 * HW + q;
      const float4 e = scb4[bi[i]];
      x0[o] = e.x;
      x0[o + HW] = e.y;
      x0[o + 2 * HW] = e.z;
      if (idx) idx[p] = bi[i];
    }
  }
}

__global__ void lincomb_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ m0,
                               const float* __restrict__ m1, float a, float b, float c, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = a * x[i] + b * m0[i];
    if (m1) v += c * (m1[i] - m0[i]);
    out[i] = v;
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_pack_weight(const float* w, void* out, int64_t N, int64_t K, void* stream) {
  SDB_REQUIRE(w && out && N > 0 && K > 0 && K % 4 == 0, "sdb_pack_weight: bad args N=%lld K=%lld", (long long)N,
              (long long)K);
  const int64_t n4 = N * K / 4;
  pack_weight_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(w, (__half*)out, n4, N * K, SDB_FMT_F16X2, 0);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_weight_fmt(const float* w, void* out, int64_t N, int64_t K, int fmt, int wexp, void* stream) {
  SDB_REQUIRE(w && out && N > 0 && K > 0 && K % 4 == 0, "sdb_pack_weight_fmt: bad args N=%lld K=%lld", (long long)N,
              (long long)K);
  SDB_REQUIRE(fmt == SDB_FMT_F16X2 || fmt == SDB_FMT_F8C, "sdb_pack_weight_fmt: bad fmt %d", fmt);
  SDB_REQUIRE(wexp >= -24 && wexp <= 40, "sdb_pack_weight_fmt: weight exponent %d out of range", wexp);
  const int64_t n4 = N * K / 4;
  pack_weight_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(w, (__half*)out, n4, N * K, fmt, wexp);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_weight_conv3_fmt(const float* w, void* out, int64_t Cout, int64_t Cin, int fmt, int wexp,
                                         void* stream) {
  SDB_REQUIRE(w && out && Cout > 0 && Cin > 0 && Cin % 4 == 0, "sdb_pack_weight_conv3_fmt: bad args (Cin %% 4 == 0)");
  SDB_REQUIRE(fmt == SDB_FMT_F16X2 || fmt == SDB_FMT_F8C, "sdb_pack_weight_conv3_fmt: bad fmt %d", fmt);
  SDB_REQUIRE(wexp >= -24 && wexp <= 40, "sdb_pack_weight_conv3_fmt: weight exponent %d out of range", wexp);
  pack_weight_conv3_v4_kernel<<<grid_for(Cout * 9 * Cin / 4, 256), 256, 0, as_stream(stream)>>>(w, (__half*)out, Cout, Cin,
                                                                                                fmt, wexp);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_weight_conv3(const float* w, void* out, int64_t Cout, int64_t Cin, void* stream) {
  SDB_REQUIRE(w && out && Cout > 0 && Cin > 0, "sdb_pack_weight_conv3: bad args");
  pack_weight_conv3_kernel<<<grid_for(Cout * 9 * Cin, 256), 256, 0, as_stream(stream)>>>(w, (__half*)out, Cout, Cin);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_rows(const float* x, int64_t ldx, void* out, int64_t M, int64_t K, int act, void* stream) {
  SDB_REQUIRE(x && out && M > 0 && K > 0 && K % 4 == 0 && ldx % 4 == 0, "sdb_pack_rows: bad args M=%lld K=%lld ldx=%lld",
              (long long)M, (long long)K, (long long)ldx);
  SDB_LAUNCH_PM(pack_rows_kernel, (grid_for(M * K / 4, 256)), (256), 0, as_stream(stream), x, ldx, (__half*)out, M, K, act);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_layernorm_pack(const float* x, const float* gamma, const float* beta, float eps, void* out,
                                  float* y, int64_t M, int64_t C, void* stream) {
  SDB_REQUIRE(x && gamma && beta && (out || y) && M > 0, "sdb_layernorm_pack: null argument");
  SDB_REQUIRE(C % 4 == 0 && C <= 512, "sdb_layernorm_pack: C=%lld must be a multiple of 4 and <= 512", (long long)C);
  const int threads = 256;
  SDB_LAUNCH_PM(layernorm_pack_kernel, (grid_for(M, threads / 32)), (threads), 0, as_stream(stream), x, gamma, beta, eps,
                                                                                      (__half*)out, y, M, (int)C);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_stats(const float* x1, int64_t C1, const float* x2, int64_t C2, float* stats, int64_t B,
                                   int64_t HW, int G, float eps, void* stream) {
  SDB_REQUIRE(x1 && stats && B > 0 && HW > 0 && G > 0, "sdb_groupnorm_stats: bad args");
  SDB_REQUIRE((C1 + C2) % G == 0 && (C2 == 0 || x2), "sdb_groupnorm_stats: C=%lld not divisible by G=%d",
              (long long)(C1 + C2), G);
  const int64_t cpg = (C1 + C2) / G;
  const bool vec = cpg % 4 == 0 && C1 % 4 == 0 && C2 % 4 == 0 && C1 % cpg == 0 && HW * (cpg / 4) < (1ll << 31) &&
                   (reinterpret_cast<uintptr_t>(x1) & 15) == 0 && (!x2 || (reinterpret_cast<uintptr_t>(x2) & 15) == 0);
  groupnorm_stats_kernel<<<(int)(B * G), 256, 0, as_stream(stream)>>>(x1, (int)C1, x2, (int)C2, stats, HW, G, eps, vec ? 1 : 0);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_apply_pack(const float* x1, int64_t C1, const float* x2, int64_t C2, const float* stats,
                                        const float* gamma, const float* beta, void* out, int64_t B, int64_t HW,
                                        int G, int silu, void* stream) {
  SDB_REQUIRE(x1 && stats && gamma && beta && out, "sdb_groupnorm_apply_pack: null argument");
  SDB_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0 && (C1 + C2) % G == 0, "sdb_groupnorm_apply_pack: bad channels");
  SDB_LAUNCH_PM(groupnorm_apply_pack_kernel, (grid_for(B * HW * (C1 + C2) / 4, 256)), (256), 0, as_stream(stream), 
      x1, (int)C1, x2, (int)C2, stats, gamma, beta, (__half*)out, B, HW, G, silu);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_nhwc(const float* x1, int64_t C1, const float* x2, int64_t C2, void* out, float* y_cat,
                             int64_t B, int64_t H, int64_t W, int mode, void* stream) {
  SDB_REQUIRE(x1 && out && B > 0 && H > 0 && W > 0, "sdb_pack_nhwc: bad args");
  SDB_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0 && (C2 == 0 || x2), "sdb_pack_nhwc: channels must be multiples of 4");
  SDB_REQUIRE(mode == SDB_PACK_PLAIN || mode == SDB_PACK_UP2 || mode == SDB_PACK_PHASE2, "sdb_pack_nhwc: bad mode");
  SDB_REQUIRE(mode != SDB_PACK_PHASE2 || (H % 2 == 0 && W % 2 == 0), "sdb_pack_nhwc: phase split needs even H, W");
  SDB_REQUIRE(!y_cat || mode != SDB_PACK_UP2, "sdb_pack_nhwc: y_cat unsupported with upsample");
  const int64_t mult = mode == SDB_PACK_UP2 ? 4 : 1;
  SDB_LAUNCH_PM(pack_nhwc_kernel, (grid_for(B * H * W * mult * (C1 + C2) / 4, 256)), (256), 0, as_stream(stream), 
      x1, (int)C1, x2, (int)C2, (__half*)out, y_cat, B, (int)H, (int)W, mode);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_geglu_pack(const float* u, void* out, int64_t M, int64_t F, void* stream) {
  SDB_REQUIRE(u && out && M > 0 && F > 0 && F % 4 == 0, "sdb_geglu_pack: bad args");
  SDB_LAUNCH_PM(geglu_pack_kernel, (grid_for(M * F / 4, 256)), (256), 0, as_stream(stream), u, (__half*)out, M, F);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_timestep_embedding_pack(const float* t, void* out, int64_t B, int dim, void* stream) {
  SDB_REQUIRE(t && out && B > 0 && dim > 0 && dim % 2 == 0, "sdb_timestep_embedding_pack: bad args");
  SDB_REQUIRE(dim % 8 == 0, "sdb_timestep_embedding_pack: dim %% 8 == 0 required (got %d)", dim);
  SDB_LAUNCH_PM(timestep_embedding_pack_kernel, (grid_for(B * dim / 4, 128)), (128), 0, as_stream(stream), t, (__half*)out, B, dim);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_finalize_cb(const float* gsum, int64_t C, float* stats, int64_t B, int64_t HW, int G, float eps,
                                         int cb, void* stream) {
  SDB_REQUIRE(gsum && stats && B > 0 && HW > 0 && G > 0 && C % G == 0 && (cb == 2 || cb == 4) && (C / G) % cb == 0,
              "sdb_groupnorm_finalize_cb: bad args C=%lld G=%d cb=%d", (long long)C, G, cb);
  groupnorm_finalize_cb_kernel<<<(unsigned)cdiv(B * G, 128), 128, 0, as_stream(stream)>>>(gsum, (int)C, stats, B, (int)HW, G,
                                                                                          eps, cb);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_add_relu(const float* h, const float* stats_h, const float* gamma_h, const float* beta_h,
                                      const float* idn, const float* stats_i, const float* gamma_i, const float* beta_i,
                                      float* out, void* out_packed, int64_t B, int64_t HW, int64_t C, int G, void* stream) {
  SDB_REQUIRE(h && stats_h && gamma_h && beta_h && (out || out_packed), "sdb_groupnorm_add_relu: null argument");
  SDB_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 4 == 0 && G > 0 && C % G == 0, "sdb_groupnorm_add_relu: bad shape C=%lld G=%d",
              (long long)C, G);
  SDB_REQUIRE(!stats_i || (idn && gamma_i && beta_i), "sdb_groupnorm_add_relu: stats_i needs idn, gamma_i, beta_i");
  SDB_LAUNCH_PM(groupnorm_add_relu_kernel, (grid_for(B * HW * C / 4, 256)), (256), 0, as_stream(stream), 
      h, stats_h, gamma_h, beta_h, idn, stats_i, gamma_i, beta_i, out, (__half*)out_packed, B, HW, (int)C, G);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_softmax_pack(const float* x, int64_t ldx, float scale, float out_scale, void* out, int64_t M, int64_t N,
                                void* stream) {
  SDB_REQUIRE(x && out && M > 0 && N > 0 && N % 4 == 0 && N <= 4096 && ldx % 4 == 0,
              "sdb_softmax_pack: bad args M=%lld N=%lld (N %% 4 == 0, N <= 4096)", (long long)M, (long long)N);
  const int wpb = 4;
  SDB_LAUNCH_PM(softmax_pack_kernel, ((unsigned)cdiv(M, wpb)), (wpb * 32), 0, as_stream(stream), x, ldx, scale, out_scale, (__half*)out, M,
                                                                                  (int)N);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_gru_gates(const float* gi, const float* gh, const float* h, float* h_new, int64_t R, int64_t D,
                             void* stream) {
  SDB_REQUIRE(gi && gh && h && h_new && R > 0 && D > 0, "sdb_gru_gates: bad args");
  gru_gates_kernel<<<grid_for(R * D, 256), 256, 0, as_stream(stream)>>>(gi, gh, h, h_new, R, D);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_conv3_in(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t Cin,
                            int64_t H, int64_t W, int64_t Cout, void* stream) {
  SDB_REQUIRE(x && w && bias && y && B > 0, "sdb_conv3_in: null argument");
  SDB_REQUIRE(B * H * W < (1ll << 31), "sdb_conv3_in: B*H*W must be below 2^31");
  SDB_REQUIRE(Cout % 4 == 0 && Cin * 9 * Cout * 4 <= 96 * 1024, "sdb_conv3_in: Cin=%lld Cout=%lld unsupported",
              (long long)Cin, (long long)Cout);
  const size_t smem = (size_t)Cin * 9 * Cout * 4;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    SDB_CHECK(cudaFuncSetAttribute(conv3_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int o4n = (int)(Cout / 4);
  SDB_REQUIRE(o4n <= 256, "sdb_conv3_in: Cout=%lld too large", (long long)Cout);
  const int threads = (256 / o4n) * o4n;          // a thread keeps its 4 output channels: blockDim is a multiple of Cout / 4
  if (W % 4 == 0) {
    if (smem > 48 * 1024) SDB_CHECK(cudaFuncSetAttribute(conv3_in_px4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3_in_px4_kernel<<<grid_for(B * H * (W / 4) * o4n, threads, 4), threads, smem, as_stream(stream)>>>(
        x, w, bias, y, B, (int)Cin, (int)H, (int)W, (int)Cout);
  } else {
    conv3_in_kernel<<<grid_for(B * H * W * o4n, threads, 4), threads, smem, as_stream(stream)>>>(x, w, bias, y, B, (int)Cin,
                                                                                          (int)H, (int)W, (int)Cout);
  }
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_conv3_out(const float* h, const float* stats, const float* gamma, const float* beta,
                             const float* w, const float* bias, float* y, int64_t B, int64_t H, int64_t W, int64_t C,
                             int G, int64_t Cout, void* stream) {
  SDB_REQUIRE(h && stats && gamma && beta && w && bias && y, "sdb_conv3_out: null argument");
  SDB_REQUIRE(Cout >= 1 && Cout <= 4 && C % G == 0, "sdb_conv3_out: Cout=%lld must be in 1..4", (long long)Cout);
  const size_t smem = (size_t)Cout * 9 * C * 4;
  SDB_REQUIRE(smem <= 48 * 1024, "sdb_conv3_out: C=%lld too large", (long long)C);
  // tiled kernel: 2 output rows per CTA when the shapes allow it (activation tile + weights in shared memory)
  const int R = 2;
  const size_t smem_t = smem + (size_t)(R + 2) * (W + 2) * C * 4;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (W % 4 == 0 && C % 4 == 0 && (C / G) % 4 == 0 && smem_t <= 200 * 1024 && B <= 65535 && al16(h) && al16(gamma) &&
      al16(beta)) {
    static size_t attr = 0;
    if (smem_t > attr) {
      SDB_CHECK(cudaFuncSetAttribute(conv3_out_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      attr = smem_t;
    }
    dim3 grid((unsigned)cdiv(H, R), (unsigned)B);
    conv3_out_tile_kernel<4><<<grid, 256, smem_t, as_stream(stream)>>>(h, stats, gamma, beta, w, bias, y, (int)H, (int)W,
                                                                       (int)C, G, (int)Cout, R);
    SDB_LAUNCH_CHECK();
    return 0;
  }
  conv3_out_kernel<4><<<grid_for(B * H * W, 8, 4), 256, smem, as_stream(stream)>>>(h, stats, gamma, beta, w, bias, y, B,
                                                                                  (int)H, (int)W, (int)C, G, (int)Cout);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_dpm_x0(const float* x, const float* eps, float alpha, float sigma, const float* codebook,
                          int64_t ncodes, float* x0, int32_t* idx, int64_t B, int64_t C, int64_t HW, void* stream) {
  SDB_REQUIRE(x && eps && x0 && B > 0 && C > 0 && C <= 8 && HW > 0, "sdb_dpm_x0: bad args (C <= 8)");
  if (codebook && C == 3 && ncodes > 0 && (size_t)ncodes * 16 <= 200 * 1024) {
    const size_t smem4 = (size_t)ncodes * 16;
    static size_t attr4 = 0;
    if (smem4 > 48 * 1024 && smem4 > attr4) {
      SDB_CHECK(cudaFuncSetAttribute(dpm_x0_vq3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
      attr4 = smem4;
    }
    const int64_t blocks = cdiv(B * HW, VQ_THREADS * VQ_PX);
    SDB_REQUIRE(blocks < (1ll << 31), "sdb_dpm_x0: grid too large");
    dpm_x0_vq3_kernel<<<(unsigned)blocks, VQ_THREADS, smem4, as_stream(stream)>>>(x, eps, alpha, sigma, codebook,
                                                                                 (int)ncodes, x0, idx, B, HW);
    SDB_LAUNCH_CHECK();
    return 0;
  }
  size_t smem = codebook ? (size_t)ncodes * (C + 1) * 4 : 0;
  SDB_REQUIRE(smem <= 200 * 1024, "sdb_dpm_x0: codebook too large for shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    SDB_CHECK(cudaFuncSetAttribute(dpm_x0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int threads = 128;
  dpm_x0_kernel<<<grid_for(B * HW, threads, 1), threads, smem, as_stream(stream)>>>(x, eps, alpha, sigma, codebook,
                                                                                    (int)ncodes, x0, idx, B, (int)C, HW);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_lincomb(float* out, const float* x, const float* m0, const float* m1, float a, float b, float c,
                           int64_t n, void* stream) {
  SDB_REQUIRE(out && x && m0 && n > 0, "sdb_lincomb: bad args");
  lincomb_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(out, x, m0, m1, a, b, c, n);
  SDB_LAUNCH_CHECK();
  return 0;
}

static thread_local float g_drop_p = 0.f;                 // set only by sdb_groupnorm_apply_pack_dropout
static thread_local unsigned long long g_drop_seed = 0;
static thread_local const unsigned long long* g_drop_seed_dev = nullptr;

extern "C" int sdb_groupnorm_apply_pack_fused(const float* x1, int64_t C1, const float* gsum1, const float* x2,
                                              int64_t C2, const float* gsum2, const float* stats, const float* gamma,
                                              const float* beta, void* out, int64_t B, int64_t HW, int G, float eps,
                                              int silu, void* stream) {
  SDB_REQUIRE(x1 && gamma && beta && out && B > 0 && HW > 0, "sdb_groupnorm_apply_pack_fused: null argument");
  SDB_REQUIRE(stats || (gsum1 && (C2 == 0 || gsum2)), "sdb_groupnorm_apply_pack_fused: need stats or partial sums");
  SDB_REQUIRE(C2 == 0 || x2, "sdb_groupnorm_apply_pack_fused: x2 missing");
  const int64_t C = C1 + C2;
  SDB_REQUIRE(G >= 1 && G <= 64 && C % G == 0 && C1 % 4 == 0 && C2 % 4 == 0 && C <= 1024 && B <= 65535,
              "sdb_groupnorm_apply_pack_fused: unsupported channels C1=%lld C2=%lld G=%d", (long long)C1, (long long)C2, G);
  SDB_REQUIRE(stats || (C / G) % 4 == 0, "sdb_groupnorm_apply_pack_fused: partial sums need (C/G) %% 4 == 0");
  const int c4n = (int)(C / 4);
  const int rpb = 256 / c4n > 0 ? 256 / c4n : 1;
  int threads = c4n * rpb;
  if (threads < 64) threads = 64;       // the first G threads also derive the statistics (G <= 64)
  // note: threads % c4n may be != 0 only in the clamp case; extra threads then map to ty >= rpb rows, still valid rows
  // ONE wave: chunks * B CTAs must not exceed the resident slots (6 CTAs per SM) -- rounding UP left a second wave of a few
  // CTAs that ran alone at a fraction of the bandwidth (ncu: 1.08 waves, SMs active 70 % of the kernel, profiles/README 23)
  int64_t chunks = ((int64_t)num_sms() * 6) / B;
  if (chunks > cdiv(HW, rpb)) chunks = cdiv(HW, rpb);
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (int)cdiv(HW, chunks);
  chunks = cdiv(HW, rows_per_chunk);
  dim3 grid((unsigned)chunks, (unsigned)B);
  {
    cudaStream_t st = as_stream(stream);
    const bool f8 = host_pack_mode() == SDB_FMT_F8C;
    const bool drop = g_drop_p > 0.f;
#define SDB_GN_LAUNCH(PMv, ACTv, DROPv)                                                                                    \
    (void)launch_k(groupnorm_apply_pack_fused_kernel<PMv, ACTv, DROPv>, grid, dim3(threads), 0, st,                         \
        x1, (int)C1, gsum1, x2, (int)C2, gsum2, stats, gamma, beta, (__half*)out, B, (int)HW, G, eps, silu, rows_per_chunk, \
        g_drop_p, g_drop_seed, g_drop_seed_dev)
#define SDB_GN_ACT(PMv, DROPv)                          \
    do {                                                \
      if (silu == 1) SDB_GN_LAUNCH(PMv, 1, DROPv);      \
      else if (silu == 2) SDB_GN_LAUNCH(PMv, 2, DROPv); \
      else SDB_GN_LAUNCH(PMv, 0, DROPv);                \
    } while (0)
    if (f8) {
      if (drop) SDB_GN_ACT(SDB_FMT_F8C, true);
      else SDB_GN_ACT(SDB_FMT_F8C, false);
    } else {
      if (drop) SDB_GN_ACT(SDB_FMT_F16X2, true);
      else SDB_GN_ACT(SDB_FMT_F16X2, false);
    }
#undef SDB_GN_ACT
#undef SDB_GN_LAUNCH
  }
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_apply_pack_dropout(const float* x1, int64_t C1, const float* x2, int64_t C2,
                                                const float* stats, const float* gamma, const float* beta, void* out,
                                                int64_t B, int64_t HW, int G, int silu, float drop_p, uint64_t seed,
                                                const uint64_t* seed_dev, void* stream) {
  SDB_REQUIRE(stats, "sdb_groupnorm_apply_pack_dropout: stats required");
  SDB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "sdb_groupnorm_apply_pack_dropout: bad p");
  g_drop_p = drop_p;
  g_drop_seed = seed;
  g_drop_seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  const int rc = sdb_groupnorm_apply_pack_fused(x1, C1, nullptr, x2, C2, nullptr, stats, gamma, beta, out, B, HW, G, 0.f,
                                                silu, stream);
  g_drop_p = 0.f;
  g_drop_seed_dev = nullptr;
  return rc;
}

extern "C" int sdb_channel_block_sums(const float* x, int64_t C, float* gsum, int64_t B, int64_t HW, void* stream) {
  SDB_REQUIRE(x && gsum && B > 0 && HW > 0, "sdb_channel_block_sums: null argument");
  SDB_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024 && B <= 65535 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "sdb_channel_block_sums: C=%lld unsupported", (long long)C);
  const int c4n = (int)(C / 4);
  const int rpb = 256 / c4n > 0 ? 256 / c4n : 1;
  // ONE wave: chunks * B CTAs must not exceed the resident slots (4 CTAs per SM) -- rounding UP left a second wave of a few
  // CTAs that ran alone at a fraction of the bandwidth (ncu: 1.08 waves, SMs active 70 % of the kernel, profiles/README 23)
  int64_t chunks = ((int64_t)num_sms() * 4) / B;
  if (chunks > cdiv(HW, rpb)) chunks = cdiv(HW, rpb);
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (int)cdiv(HW, chunks);
  chunks = cdiv(HW, rows_per_chunk);
  dim3 grid((unsigned)chunks, (unsigned)B);
  channel_block_sums_kernel<<<grid, c4n * rpb, 0, as_stream(stream)>>>(x, (int)C, gsum, (int)HW, rows_per_chunk);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_finalize(const float* gsum1, int64_t C1, const float* gsum2, int64_t C2, float* stats,
                                      int64_t B, int64_t HW, int G, float eps, void* stream) {
  SDB_REQUIRE(gsum1 && stats && B > 0 && (C2 == 0 || gsum2), "sdb_groupnorm_finalize: null argument");
  SDB_REQUIRE((C1 + C2) % G == 0 && ((C1 + C2) / G) % 4 == 0 && C1 % 4 == 0, "sdb_groupnorm_finalize: bad channels");
  groupnorm_finalize_kernel<<<(unsigned)cdiv(B * G, 128), 128, 0, as_stream(stream)>>>(gsum1, (int)C1, gsum2, (int)C2,
                                                                                      stats, B, (int)HW, G, eps);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_weight_geglu(const float* w, const float* bias, void* out, float* bias_out, int64_t F, int64_t K,
                                     void* stream) {
  SDB_REQUIRE(w && out && F > 0 && F % 16 == 0 && K > 0 && K % 4 == 0, "sdb_pack_weight_geglu: bad args F=%lld K=%lld",
              (long long)F, (long long)K);
  SDB_LAUNCH_PM(pack_weight_geglu_kernel, (grid_for(2 * F * K / 4, 256)), (256), 0, as_stream(stream), w, (__half*)out, F, K);
  SDB_LAUNCH_CHECK();
  if (bias && bias_out) {
    permute_geglu_bias_kernel<<<grid_for(2 * F, 256), 256, 0, as_stream(stream)>>>(bias, bias_out, F);
    SDB_LAUNCH_CHECK();
  }
  return 0;
}
