// Backward (training) kernels of the hot path: everything around the tensor-core GEMMs of the backward pass.
// The contractions themselves (dgrad, wgrad) are sdb_gemm calls on operands produced here:
//   dX = dY W          -> sdb_gemm(pack(dY),            pack_T(W))                        (linear / 1x1 conv)
//   dX = conv3(dY, W') -> sdb_gemm(pack(dY) conv mode,  pack_conv3_dgrad(W))              (flipped taps, in/out swapped)
//   dW = dY^T X        -> sdb_gemm(pack_T(dY),          transpose_packed(X))  split-K     (linear)
//   dW = conv wgrad    -> sdb_gemm(transpose_packed(X) SDB_A_WGRAD, pack_T(dY)) split-K   (3x3, contraction over pixels)
// Reference semantics: torch autograd of the forward lines cited in sdb200.h.
#include "common.cuh"

namespace sdb {

static inline int grid_for_bw(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = cdiv(work_items, threads);
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float silu_grad_f(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}
__device__ __forceinline__ float gelu_grad_f(float g) {
  return 0.5f * (1.f + erff(g * 0.70710678118654752440f)) + g * 0.39894228040143267794f * __expf(-0.5f * g * g);
}
// ------------------------------------------------------------------ gradient packing
// dy [M,N] fp32 (row stride ld) -> BF16-split packed rows [2][M][N] (dgrad operand), packed transpose [2][N][M] (wgrad operand),
// column sums (bias gradient, atomically accumulated) and per-group column sums (timestep-embedding gradient:
// group = row / rows_per_group).  64x64 tiles through shared memory; one read of dy feeds all four products.
constexpr int GP_T = 64;
__global__ void __launch_bounds__(256)
grad_pack_kernel(const float* __restrict__ dy, int64_t ld, __half* __restrict__ out_rows, __half* __restrict__ out_T,
                 float* __restrict__ bias_grad, float* __restrict__ group_grad, int64_t ldg, int64_t M, int64_t N,
                 int rows_per_group, int64_t ldt) {
  __shared__ float tile[GP_T][GP_T + 1];
  const int64_t m0 = (int64_t)blockIdx.y * GP_T, n0 = (int64_t)blockIdx.x * GP_T;
  const int tid = threadIdx.x;
  {
    const int c4 = (tid & 15) * 4, r0 = tid >> 4;
    const bool vec = (N % 4 == 0) && (ld % 4 == 0);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = r0 + p * 16;
      const int64_t m = m0 + r, n = n0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        if (vec && n + 3 < N) {
          v = *reinterpret_cast<const float4*>(dy + m * ld + n);
          if (out_rows) store_split4_bf16(out_rows, out_rows + M * N, m * N + n, v);
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j)
            if (n + j < N) {
              t[j] = dy[m * ld + n + j];
              if (out_rows) {
                __half h, l;
                split_bf16(t[j], h, l);
                out_rows[m * N + n + j] = h;
                out_rows[M * N + m * N + n + j] = l;
              }
            }
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
    }
  }
  __syncthreads();
  if (tid < GP_T && n0 + tid < N && (bias_grad || group_grad)) {
    float tot = 0.f, run = 0.f;
    int64_t cur_g = m0 / (rows_per_group > 0 ? rows_per_group : 1);
    for (int r = 0; r < GP_T; ++r) {
      const int64_t m = m0 + r;
      if (m >= M) break;
      if (group_grad) {
        const int64_t gidx = m / rows_per_group;
        if (gidx != cur_g) {
          atomicAdd(group_grad + cur_g * ldg + n0 + tid, run);
          run = 0.f;
          cur_g = gidx;
        }
      }
      const float v = tile[r][tid];
      run += v;
      tot += v;
    }
    if (group_grad) atomicAdd(group_grad + cur_g * ldg + n0 + tid, run);
    if (bias_grad) atomicAdd(bias_grad + n0 + tid, tot);
  }
  if (out_T) {
    // thread <-> (column n = tid / 4, 16 consecutive rows): 32-byte contiguous writes per plane
    const int n = tid >> 2, mq = (tid & 3) * 16;
    if (n0 + n < N) {
      const int64_t base = (n0 + n) * ldt + m0 + mq;
      const bool full = (m0 + mq + 15 < M) && (ldt % 8 == 0);
      __align__(16) __half hi[16];
      __align__(16) __half lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_bf16(tile[mq + j][n], hi[j], lo[j]);
      if (full) {
        *reinterpret_cast<uint4*>(out_T + base) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(out_T + base + 8) = *reinterpret_cast<uint4*>(hi + 8);
        *reinterpret_cast<uint4*>(out_T + N * ldt + base) = *reinterpret_cast<uint4*>(lo);
        *reinterpret_cast<uint4*>(out_T + N * ldt + base + 8) = *reinterpret_cast<uint4*>(lo + 8);
      } else {
        for (int j = 0; j < 16; ++j)
          if (m0 + mq + j < M) {
            out_T[base + j] = hi[j];
            out_T[N * ldt + base + j] = lo[j];
          }
      }
    }
  }
}

// packed [2][M][K] -> packed [2][K][M] (both planes), 64x64 tiles of halves
__global__ void __launch_bounds__(256)
transpose_packed_kernel(const __half* __restrict__ in, __half* __restrict__ out, int64_t M, int64_t K, int64_t ldt,
                        int to_bf16) {
  __shared__ __half tile[2][GP_T][GP_T + 2];
  const int64_t m0 = (int64_t)blockIdx.y * GP_T, k0 = (int64_t)blockIdx.x * GP_T;
  const int tid = threadIdx.x;
  const int64_t plane = M * K;
  {
    const int c8 = (tid & 7) * 8, r0 = tid >> 3;    // 8 threads x 8 halves per row, 32 rows per pass
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int r = r0 + p * 32;
      const int64_t m = m0 + r, k = k0 + c8;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        __align__(16) __half v[8];
        if (m < M && k + 7 < K && (K % 8 == 0)) {
          *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(in + pl * plane + m * K + k);
        } else {
          for (int j = 0; j < 8; ++j) v[j] = (m < M && k + j < K) ? in[pl * plane + m * K + k + j] : __float2half(0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[pl][r][c8 + j] = v[j];
      }
      if (to_bf16) {   // re-split the fp16 (hi, lo) pair as a bf16 (hi, lo) pair: both GEMM operands must share a format
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = __half2float(tile[0][r][c8 + j]) + __half2float(tile[1][r][c8 + j]);
          split_bf16(x, tile[0][r][c8 + j], tile[1][r][c8 + j]);
        }
      }
    }
  }
  __syncthreads();
  {
    const int k = tid >> 2, mq = (tid & 3) * 16;
    if (k0 + k < K) {
      const int64_t base = (k0 + k) * ldt + m0 + mq;
      const bool full = (m0 + mq + 15 < M) && (ldt % 8 == 0);
      const int64_t oplane = K * ldt;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        __align__(16) __half v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tile[pl][mq + j][k];
        if (full) {
          *reinterpret_cast<uint4*>(out + pl * oplane + base) = *reinterpret_cast<uint4*>(v);
          *reinterpret_cast<uint4*>(out + pl * oplane + base + 8) = *reinterpret_cast<uint4*>(v + 8);
        } else {
          for (int j = 0; j < 16; ++j)
            if (m0 + mq + j < M) out[pl * oplane + base + j] = v[j];
        }
      }
    }
  }
}

// fp16-split operand -> bf16-split operand of the same shape (conv wgrad reads the forward activation operand next to a
// bf16 gradient operand; tcgen05 kind::f16 wants one format for both)
__global__ void repack_bf16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int64_t n8) {
  const int64_t plane = n8 * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(in + i * 8);
    *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(in + plane + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16(__half2float(hi[j]) + __half2float(lo[j]), hi[j], lo[j]);
    *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(out + plane + i * 8) = *reinterpret_cast<uint4*>(lo);
  }
}

// conv3x3 weight [Cout][Cin][3][3] -> dgrad operand packed [Cin][9*Cout], k = tap'*Cout + co, tap' = 8 - tap
// (180-degree rotation), i.e. the weight of the transposed convolution in the implicit-GEMM layout of sdb_gemm.
__global__ void pack_weight_conv3_dgrad_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t Cout,
                                               int64_t Cin, int bf16) {
  const int64_t total = Cin * 9 * Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = i % Cout;
    const int64_t tapp = (i / Cout) % 9;
    const int64_t ci = i / (9 * Cout);
    const float v = w[(co * Cin + ci) * 9 + (8 - tapp)];
    __half h, l;
    if (bf16) split_bf16(v, h, l);
    else split_f16(v, h, l);
    out[i] = h;
    out[total + i] = l;
  }
}

// wgrad GEMM result c9 [9*Cin][Cout] (row = tap*Cin + ci) -> dW [Cout][Cin_w][3][3] (first Cin_w input channels kept)
__global__ void wgrad_conv3_scatter_kernel(const float* __restrict__ c9, int64_t ldc, float* __restrict__ dw,
                                           int64_t Cout, int64_t Cin, int64_t Cin_w, int accumulate) {
  const int64_t total = Cout * Cin_w * 9;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tap = i % 9, ci = (i / 9) % Cin_w, co = i / (9 * Cin_w);
    const float v = c9[(tap * Cin + ci) * ldc + co];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// out = a (+ b) (+ c): gradient accumulation of residual / skip paths
__global__ void add3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                            const float* __restrict__ c, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(a)[i];
    if (b) { const float4 t = reinterpret_cast<const float4*>(b)[i]; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    if (c) { const float4 t = reinterpret_cast<const float4*>(c)[i]; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// dx = dy * act'(pre): act 1 SiLU, 2 ReLU.  Strided rows.
__global__ void act_bwd_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ pre, int64_t ldp,
                               float* __restrict__ dx, int64_t ldx, int64_t M, int64_t N, int act) {
  const int64_t total = M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N, n = i % N;
    const float z = pre[m * ldp + n];
    const float g = act == 1 ? silu_grad_f(z) : (z > 0.f ? 1.f : 0.f);
    dx[m * ldx + n] = dy[m * ldy + n] * g;
  }
}

// same, four elements per thread with 32-bit index arithmetic (N, the strides and the base addresses multiples of 4 floats,
// M * N / 4 < 2^31): the scalar form above runs a 64-bit division and a modulo per element
__global__ void act_bwd_vec4_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ pre, int64_t ldp,
                                    float* __restrict__ dx, int64_t ldx, unsigned total4, unsigned n4, int act) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const unsigned m = i / n4, n = (i - m * n4) * 4;
    const float4 z = *reinterpret_cast<const float4*>(pre + (int64_t)m * ldp + n);
    const float4 d = *reinterpret_cast<const float4*>(dy + (int64_t)m * ldy + n);
    float4 o;
    if (act == 1) {
      o = make_float4(d.x * silu_grad_f(z.x), d.y * silu_grad_f(z.y), d.z * silu_grad_f(z.z), d.w * silu_grad_f(z.w));
    } else {
      o = make_float4(z.x > 0.f ? d.x : 0.f, z.y > 0.f ? d.y : 0.f, z.z > 0.f ? d.z : 0.f, z.w > 0.f ? d.w : 0.f);
    }
    *reinterpret_cast<float4*>(dx + (int64_t)m * ldx + n) = o;
  }
}

// ------------------------------------------------------------------ GroupNorm (+SiLU, +dropout) backward
// forward: xh = (x - mean_g) rstd_g ; z = xh gamma + beta ; a = act(z) * dropmask.   Given da:
//   dz = da * dropmask * act'(z);  A_bc = sum_hw dz;  B_bc = sum_hw dz xh   (reduce kernel, atomics into sums[B][C][2])
//   dx = rstd (dz gamma - S1_g / n - xh S2_g / n),  S1_g = sum_{c in g} gamma_c A_bc,  S2_g = sum_{c in g} gamma_c B_bc
//   dgamma_c = sum_b B_bc ; dbeta_c = sum_b A_bc                                      (param kernel)
struct GnBwdArgs {
  const float* x1; const float* x2; const float* da; const float* stats; const float* gamma; const float* beta;
  float* sums;            // [B][C][2]
  float* dx1; float* dx2; // outputs of the apply pass
  const float* add1; const float* add2;   // optional gradients to accumulate into dx1 / dx2
  int64_t B; int C1, C2, HW, G, silu, rows_per_chunk;
  float drop_p; unsigned long long seed; const unsigned long long* seed_dev;
};

// ACT (0 none / 1 SiLU / 2 ReLU) and DROP are compile-time (instances, not flags: DESIGN.md section 4); the row loop is
// unrolled so that the loads of several rows are in flight together (one row per trip left the kernel latency bound).
template <bool APPLY, int ACT, bool DROP>
__global__ void __launch_bounds__(256) groupnorm_bwd_kernel(const GnBwdArgs a) {
  __shared__ float s_mean[64], s_rstd[64], s_s1[64], s_s2[64];
  const int C = a.C1 + a.C2, c4n = C >> 2, cpg = C / a.G;
  const int64_t b = blockIdx.y;
  const float inv_n = 1.f / ((float)a.HW * cpg);
  if (threadIdx.x < a.G) {
    const int g = threadIdx.x;
    s_mean[g] = a.stats[(b * a.G + g) * 2];
    s_rstd[g] = a.stats[(b * a.G + g) * 2 + 1];
    if (APPLY) {
      float s1 = 0.f, s2 = 0.f;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s1 += a.gamma[c] * a.sums[(b * C + c) * 2];
        s2 += a.gamma[c] * a.sums[(b * C + c) * 2 + 1];
      }
      s_s1[g] = s1 * inv_n;
      s_s2[g] = s2 * inv_n;
    }
  }
  __syncthreads();
  const int tx = threadIdx.x % c4n, ty = threadIdx.x / c4n, rpb = blockDim.x / c4n;
  if (ty >= rpb) return;
  const int c = tx * 4;
  float gm[4], bt[4], mu[4], rs[4], m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g = (c + j) / cpg;
    gm[j] = a.gamma[c + j]; bt[j] = a.beta[c + j]; mu[j] = s_mean[g]; rs[j] = s_rstd[g];
    m1[j] = APPLY ? s_s1[g] : 0.f; m2[j] = APPLY ? s_s2[g] : 0.f;
  }
  const bool from1 = c < a.C1;
  const float* src = from1 ? a.x1 + c : a.x2 + (c - a.C1);
  const int ld = from1 ? a.C1 : a.C2;
  float* dst = from1 ? a.dx1 + c : (a.dx2 ? a.dx2 + (c - a.C1) : nullptr);
  const float* add = from1 ? (a.add1 ? a.add1 + c : nullptr) : (a.add2 ? a.add2 + (c - a.C1) : nullptr);
  const float inv_keep = DROP ? 1.f / (1.f - a.drop_p) : 1.f;
  const unsigned long long seed = a.seed + ((DROP && a.seed_dev) ? *a.seed_dev * 0x9E3779B97F4A7C15ull : 0ull);
  float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
  const int r_begin = blockIdx.x * a.rows_per_chunk;
  const int r_end = min(a.HW, r_begin + a.rows_per_chunk);
#pragma unroll 4
  for (int r = r_begin + ty; r < r_end; r += rpb) {
    const int64_t row = b * a.HW + r;
    const float4 xv = *reinterpret_cast<const float4*>(src + row * ld);
    const float4 dv = *reinterpret_cast<const float4*>(a.da + row * C + c);
    const float x[4] = {xv.x, xv.y, xv.z, xv.w};
    const float d[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (x[j] - mu[j]) * rs[j];
      float dz = d[j];
      if (DROP) dz *= dropout_scale(seed, (unsigned long long)(row * C + c + j), a.drop_p, inv_keep);
      if (ACT == 1) dz *= silu_grad_f(xh * gm[j] + bt[j]);
      else if (ACT == 2) {
        // ReLU (ResNet18-GN encoder, resnet.py:76-78).  The mask must be the FORWARD's: the same expression as the operand
        // producer (elementwise.cu: v * (rstd gamma) + (beta - mean rstd gamma)), not an algebraically equal one -- an
        // input within round-off of zero would otherwise pass in one direction and be blocked in the other
        const float scj = rs[j] * gm[j];
        const float shj = bt[j] - mu[j] * scj;
        // ... and "passed" means the value that reached the next convolution is non-zero: the operand is stored as fp16 hi + lo,
        // which rounds everything up to 2^-25 to zero.  With `> 0` an input in (0, 2^-25] was blocked in the forward and let
        // through here: ~0.2 such elements per 7.7 M ReLU inputs, each moving upstream gradients by O(1e-3) -- the
        // one-in-four failure of tests/test_resnet_gpu.py forced-pattern checks
        dz = (x[j] * scj + shj > 2.98023223876953125e-8f) ? dz : 0.f;
      }
      if (APPLY) {
        o[j] = rs[j] * (dz * gm[j] - m1[j] - xh * m2[j]);
      } else {
        accA[j] += dz;
        accB[j] += dz * xh;
      }
    }
    if (APPLY && dst) {
      float4 ov = make_float4(o[0], o[1], o[2], o[3]);
      if (add) {
        const float4 t = *reinterpret_cast<const float4*>(add + row * ld);
        ov.x += t.x; ov.y += t.y; ov.z += t.z; ov.w += t.w;
      }
      *reinterpret_cast<float4*>(dst + row * ld) = ov;
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(a.sums + (b * C + c + j) * 2, accA[j]);
      atomicAdd(a.sums + (b * C + c + j) * 2 + 1, accB[j]);
    }
  }
}

__global__ void groupnorm_bwd_param_kernel(const float* __restrict__ sums, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta, int64_t B, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, bsum = 0.f;
  for (int64_t b = 0; b < B; ++b) {
    a += sums[(b * C + c) * 2];
    bsum += sums[(b * C + c) * 2 + 1];
  }
  dbeta[c] += a;
  dgamma[c] += bsum;
}

// ------------------------------------------------------------------ LayerNorm backward (one warp per row, C <= 512)
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dn, const float* __restrict__ gamma,
                     float eps, const float* __restrict__ add, float* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int64_t M, int C) {
  __shared__ float s_dg[512], s_db[512];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int c4 = C / 4;
  float4 gm[4];
  float4 ag[4], ab[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = lane + j * 32;
    gm[j] = idx < c4 ? reinterpret_cast<const float4*>(gamma)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); row < M; row += (int64_t)gridDim.x * wpb) {
    float4 v[4], d[4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        v[j] = reinterpret_cast<const float4*>(x + row * C)[idx];
        d[j] = reinterpret_cast<const float4*>(dn + row * C)[idx];
        s += v[j].x + v[j].y + v[j].z + v[j].w;
      } else {
        v[j] = d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
        ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    float t1 = 0.f, t2 = 0.f;   // sum dxh, sum dxh*xh
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        v[j].x = (v[j].x - mean) * rstd; v[j].y = (v[j].y - mean) * rstd;
        v[j].z = (v[j].z - mean) * rstd; v[j].w = (v[j].w - mean) * rstd;
        ag[j].x += d[j].x * v[j].x; ag[j].y += d[j].y * v[j].y; ag[j].z += d[j].z * v[j].z; ag[j].w += d[j].w * v[j].w;
        ab[j].x += d[j].x; ab[j].y += d[j].y; ab[j].z += d[j].z; ab[j].w += d[j].w;
        d[j].x *= gm[j].x; d[j].y *= gm[j].y; d[j].z *= gm[j].z; d[j].w *= gm[j].w;
        t1 += d[j].x + d[j].y + d[j].z + d[j].w;
        t2 += d[j].x * v[j].x + d[j].y * v[j].y + d[j].z * v[j].z + d[j].w * v[j].w;
      }
    }
    t1 = warp_sum(t1) / C;
    t2 = warp_sum(t2) / C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + j * 32;
      if (idx < c4) {
        float4 o;
        o.x = rstd * (d[j].x - t1 - v[j].x * t2); o.y = rstd * (d[j].y - t1 - v[j].y * t2);
        o.z = rstd * (d[j].z - t1 - v[j].z * t2); o.w = rstd * (d[j].w - t1 - v[j].w * t2);
        if (add) {
          const float4 t = reinterpret_cast<const float4*>(add + row * C)[idx];
          o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
        }
        reinterpret_cast<float4*>(dx + row * C)[idx] = o;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = lane + j * 32;
    if (idx < c4) {
      atomicAdd(&s_dg[idx * 4], ag[j].x); atomicAdd(&s_dg[idx * 4 + 1], ag[j].y);
      atomicAdd(&s_dg[idx * 4 + 2], ag[j].z); atomicAdd(&s_dg[idx * 4 + 3], ag[j].w);
      atomicAdd(&s_db[idx * 4], ab[j].x); atomicAdd(&s_db[idx * 4 + 1], ab[j].y);
      atomicAdd(&s_db[idx * 4 + 2], ab[j].z); atomicAdd(&s_db[idx * 4 + 3], ab[j].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, s_dg[i]);
    atomicAdd(dbeta + i, s_db[i]);
  }
}

// ------------------------------------------------------------------ attention core backward (attention.py:188-205)
// pass Q: thread <-> query row: recompute the softmax row (online), LSE_i, D_i = do_i . o_i, then dq_i.
// pass KV: thread <-> key row: dv_j = sum_i p_ij do_i ; dk_j = scale sum_i p_ij (do_i . v_j - D_i) q_i.
constexpr int AB_THREADS = 64;
constexpr int AB_TILE = 64;

template <int D>
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_q_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                       const float* __restrict__ v, int64_t ldv, const float* __restrict__ dout, int64_t ldo,
                       float* __restrict__ dq, int64_t lddq, float* __restrict__ lse, float* __restrict__ dsum,
                       int64_t Lq, int64_t Lk, int heads, float scale) {
  __shared__ __align__(16) float sk[AB_TILE][D];
  __shared__ __align__(16) float sv[AB_TILE][D];
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t row = blockIdx.x * (int64_t)AB_THREADS + threadIdx.x;
  const bool active = row < Lq;
  float qr[D], dor[D], acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    qr[i] = active ? q[(b * Lq + row) * ldq + h * D + i] * scale : 0.f;
    dor[i] = active ? dout[(b * Lq + row) * ldo + h * D + i] : 0.f;
    acc[i] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int64_t j0 = 0; j0 < Lk; j0 += AB_TILE) {
    const int nk = (int)min((int64_t)AB_TILE, Lk - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * (D / 4); i += AB_THREADS) {
      const int r = i / (D / 4), c = i % (D / 4);
      reinterpret_cast<float4*>(&sk[r][0])[c] = reinterpret_cast<const float4*>(k + (b * Lk + j0 + r) * ldk + h * D)[c];
      reinterpret_cast<float4*>(&sv[r][0])[c] = reinterpret_cast<const float4*>(v + (b * Lk + j0 + r) * ldv + h * D)[c];
    }
    __syncthreads();
    for (int j = 0; j < nk; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) s += qr[i] * sk[j][i];
      const float mnew = fmaxf(mrun, s);
      const float corr = expf(mrun - mnew), p = expf(s - mnew);
      lrun = lrun * corr + p;
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] = acc[i] * corr + p * sv[j][i];
      mrun = mnew;
    }
  }
  const float L = mrun + logf(lrun);
  float Dm = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) Dm += dor[i] * acc[i];
  Dm /= lrun;
  if (active) {
    lse[(b * heads + h) * Lq + row] = L;
    dsum[(b * heads + h) * Lq + row] = Dm;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.f;   // now dq
  for (int64_t j0 = 0; j0 < Lk; j0 += AB_TILE) {
    const int nk = (int)min((int64_t)AB_TILE, Lk - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * (D / 4); i += AB_THREADS) {
      const int r = i / (D / 4), c = i % (D / 4);
      reinterpret_cast<float4*>(&sk[r][0])[c] = reinterpret_cast<const float4*>(k + (b * Lk + j0 + r) * ldk + h * D)[c];
      reinterpret_cast<float4*>(&sv[r][0])[c] = reinterpret_cast<const float4*>(v + (b * Lk + j0 + r) * ldv + h * D)[c];
    }
    __syncthreads();
    for (int j = 0; j < nk; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) { s += qr[i] * sk[j][i]; dp += dor[i] * sv[j][i]; }
      const float ds = expf(s - L) * (dp - Dm) * scale;
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] += ds * sk[j][i];
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < D; ++i) dq[(b * Lq + row) * lddq + h * D + i] = acc[i];
  }
}

template <int D>
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_kv_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                        const float* __restrict__ v, int64_t ldv, const float* __restrict__ dout, int64_t ldo,
                        const float* __restrict__ lse, const float* __restrict__ dsum, float* __restrict__ dk,
                        int64_t lddk, float* __restrict__ dv, int64_t lddv, int64_t Lq, int64_t Lk, int heads,
                        float scale) {
  __shared__ __align__(16) float sq[AB_TILE][D];
  __shared__ __align__(16) float sdo[AB_TILE][D];
  __shared__ float sl[AB_TILE], sd[AB_TILE];
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t row = blockIdx.x * (int64_t)AB_THREADS + threadIdx.x;   // key row
  const bool active = row < Lk;
  float kr[D], vr[D], dkr[D], dvr[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    kr[i] = active ? k[(b * Lk + row) * ldk + h * D + i] : 0.f;
    vr[i] = active ? v[(b * Lk + row) * ldv + h * D + i] : 0.f;
    dkr[i] = dvr[i] = 0.f;
  }
  for (int64_t i0 = 0; i0 < Lq; i0 += AB_TILE) {
    const int nq = (int)min((int64_t)AB_TILE, Lq - i0);
    __syncthreads();
    for (int i = threadIdx.x; i < nq * (D / 4); i += AB_THREADS) {
      const int r = i / (D / 4), c = i % (D / 4);
      float4 t = reinterpret_cast<const float4*>(q + (b * Lq + i0 + r) * ldq + h * D)[c];
      t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
      reinterpret_cast<float4*>(&sq[r][0])[c] = t;
      reinterpret_cast<float4*>(&sdo[r][0])[c] =
          reinterpret_cast<const float4*>(dout + (b * Lq + i0 + r) * ldo + h * D)[c];
    }
    for (int i = threadIdx.x; i < nq; i += AB_THREADS) {
      sl[i] = lse[(b * heads + h) * Lq + i0 + i];
      sd[i] = dsum[(b * heads + h) * Lq + i0 + i];
    }
    __syncthreads();
    for (int i = 0; i < nq; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) { s += sq[i][d] * kr[d]; dp += sdo[i][d] * vr[d]; }
      const float p = expf(s - sl[i]);
      const float ds = p * (dp - sd[i]);
#pragma unroll
      for (int d = 0; d < D; ++d) { dvr[d] += p * sdo[i][d]; dkr[d] += ds * sq[i][d]; }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      dk[(b * Lk + row) * lddk + h * D + i] = dkr[i];     // sq already carries `scale`
      dv[(b * Lk + row) * lddv + h * D + i] = dvr[i];
    }
  }
}

// quotient / remainder of an element index: 32-bit when the whole index space fits (`small`, CTA-uniform) -- a 64-bit division
// is ~100 instructions, and these element-wise kernels were issue-bound on them
__device__ __forceinline__ void divmod_idx(int64_t i, int64_t d, bool small, int64_t& q, int64_t& r) {
  if (small) {
    const unsigned qq = (unsigned)i / (unsigned)d;
    q = qq;
    r = (unsigned)i - qq * (unsigned)d;
  } else {
    q = i / d;
    r = i - q * d;
  }
}

// ------------------------------------------------------------------ GEGLU backward (attention.py:46-48)
__global__ void geglu_bwd_kernel(const float* __restrict__ u, const float* __restrict__ dg, float* __restrict__ du,
                                 int64_t M, int64_t F) {
  const int64_t f4 = F / 4, total = M * f4;
  const bool small = total < (1ll << 31);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m, j;
    divmod_idx(i, f4, small, m, j);
    j *= 4;
    const float4 a = *reinterpret_cast<const float4*>(u + m * 2 * F + j);
    const float4 g = *reinterpret_cast<const float4*>(u + m * 2 * F + F + j);
    const float4 d = *reinterpret_cast<const float4*>(dg + m * F + j);
    float4 da, dgate;
    da.x = d.x * gelu_erf_f(g.x); da.y = d.y * gelu_erf_f(g.y); da.z = d.z * gelu_erf_f(g.z); da.w = d.w * gelu_erf_f(g.w);
    dgate.x = d.x * a.x * gelu_grad_f(g.x); dgate.y = d.y * a.y * gelu_grad_f(g.y);
    dgate.z = d.z * a.z * gelu_grad_f(g.z); dgate.w = d.w * a.w * gelu_grad_f(g.w);
    *reinterpret_cast<float4*>(du + m * 2 * F + j) = da;
    *reinterpret_cast<float4*>(du + m * 2 * F + F + j) = dgate;
  }
}

// ------------------------------------------------------------------ resampling adjoints
// nearest x2 upsample adjoint: dx[b,y,x,:] = sum of the 2x2 block of dup [B,2H,2W,C]
__global__ void up2_adjoint_kernel(const float* __restrict__ dup, float* __restrict__ dx, int64_t B, int H, int W,
                                   int C) {
  const int c4n = C / 4;
  const int64_t total = B * H * W * c4n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = int(i % c4n) * 4;
    int64_t p = i / c4n;
    const int x = int(p % W);
    p /= W;
    const int y = int(p % H);
    const int64_t b = p / H;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx_ = 0; dx_ < 2; ++dx_) {
        const float4 t = *reinterpret_cast<const float4*>(dup + ((b * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx_) * C + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
    *reinterpret_cast<float4*>(dx + ((b * H + y) * W + x) * C + c) = s;
  }
}

// zero-insertion: z[b,2y,2x,:] = dy[b,y,x,:], 0 elsewhere -> packed [2][B*2H*2W][C]  (dgrad of a stride-2 conv as a
// stride-1 conv of z with the rotated kernel)
__global__ void pack_zero_up2_kernel(const float* __restrict__ dy, __half* __restrict__ out, int64_t B, int H, int W,
                                     int C) {
  const int c4n = C / 4;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t total = B * Ho * Wo * c4n;
  const int64_t plane = B * (int64_t)Ho * Wo * C;
  const bool small = total < (1ll << 31);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p, c64, p2, xo64, b, yo64;
    divmod_idx(i, c4n, small, p, c64);
    divmod_idx(p, Wo, small, p2, xo64);
    divmod_idx(p2, Ho, small, b, yo64);
    const int c = int(c64) * 4, xo = int(xo64), yo = int(yo64);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(yo & 1) && !(xo & 1)) v = *reinterpret_cast<const float4*>(dy + ((b * H + (yo >> 1)) * W + (xo >> 1)) * C + c);
    store_split4_bf16(out, out + plane, ((b * Ho + yo) * Wo + xo) * C + c, v);
  }
}

// NCHW [B,Cs,H,W] fp32 -> NHWC packed [2][B*H*W][Cp] with channels zero-padded to Cp (input conv as a GEMM in training)
__global__ void pack_nchw_pad_kernel(const float* __restrict__ x, __half* __restrict__ out, int64_t B, int Cs, int HW,
                                     int Cp) {
  const int64_t total = B * HW * Cp;
  const bool small = total < (1ll << 31);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p, c64, b, pix;           // p = b*HW + pix
    divmod_idx(i, Cp, small, p, c64);
    divmod_idx(p, HW, small, b, pix);
    const int c = int(c64);
    const float v = c < Cs ? x[(b * Cs + c) * HW + pix] : 0.f;
    __half h, l;
    split_f16(v, h, l);
    out[i] = h;
    out[total + i] = l;
  }
}
// NHWC rows [B*HW, ld] (first Cs columns) <-> NCHW [B,Cs,HW]
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int64_t ld, float* __restrict__ out, int64_t B, int Cs,
                                    int HW) {
  const int64_t total = B * Cs * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i % HW, c = (i / HW) % Cs, b = i / ((int64_t)HW * Cs);
    out[i] = in[(b * HW + pix) * ld + c];
  }
}
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t B, int Cs, int HW,
                                        int Cp) {
  const int64_t total = B * HW * Cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = int(i % Cp);
    const int64_t p = i / Cp;
    const int64_t b = p / HW, pix = p % HW;
    out[i] = c < Cs ? in[(b * Cs + c) * HW + pix] : 0.f;
  }
}

// explicit transposed im2col for tiny feature maps (W < 8, where a TMA box row would be < 16 bytes):
// in packed [2][B*H*W][C] NHWC -> out packed [2][9*C][B*H*W], row = tap*C + c, zero padding
__global__ void im2col_T_kernel(const __half* __restrict__ in, __half* __restrict__ out, int64_t B, int H, int W, int C) {
  const int64_t Mpix = B * H * W;
  const int64_t total = 9 * (int64_t)C * Mpix;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i % Mpix;
    const int64_t rc = i / Mpix;
    const int c = int(rc % C), tap = int(rc / C);
    const int x = int(m % W), y = int((m / W) % H);
    const int64_t b = m / ((int64_t)H * W);
    const int yi = y + tap / 3 - 1, xi = x + tap % 3 - 1;
    __half h = __float2half(0.f), l = h;
    if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
      const int64_t src = ((b * H + yi) * W + xi) * C + c;
      h = in[src];
      l = in[Mpix * C + src];
    }
    out[i] = h;
    out[total + i] = l;
  }
}

// ------------------------------------------------------------------ GRUCell backward, pointwise part
// given dh' [R,D]: dgi, dgh [R,3D] and the direct path dh = dh' * z.   (forward: sdb_gru_gates)
__global__ void gru_gates_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                     const float* __restrict__ h, const float* __restrict__ dhn,
                                     float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh,
                                     int64_t R, int64_t D) {
  const int64_t total = R * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D, d = i % D;
    const float* a = gi + r * 3 * D;
    const float* b = gh + r * 3 * D;
    const float rg = 1.f / (1.f + expf(-(a[d] + b[d])));
    const float zg = 1.f / (1.f + expf(-(a[D + d] + b[D + d])));
    const float ng = tanhf(a[2 * D + d] + rg * b[2 * D + d]);
    const float g = dhn[i];
    const float dn = g * (1.f - zg);
    const float dz = g * (h[i] - ng);
    const float dpn = dn * (1.f - ng * ng);
    const float dpr = dpn * b[2 * D + d] * rg * (1.f - rg);
    const float dpz = dz * zg * (1.f - zg);
    dgi[r * 3 * D + d] = dpr; dgi[r * 3 * D + D + d] = dpz; dgi[r * 3 * D + 2 * D + d] = dpn;
    dgh[r * 3 * D + d] = dpr; dgh[r * 3 * D + D + d] = dpz; dgh[r * 3 * D + 2 * D + d] = dpn * rg;
    dh[i] = g * zg;
  }
}

// ------------------------------------------------------------------ Slot-Attention iteration backward
// forward (sdb_slot_attend): l = scale k q^T ; P = softmax_s(l) ; a = P + eps ; cs_s = sum_n a ; U_s = sum_n a v / cs_s
// given dU [B,S,D], U, cs:   G_s = dU_s / cs_s ;  c_s = (U_s . dU_s) / cs_s
//   da[n,s] = v_n . G_s - c_s ;  dl = P (da - sum_s' P da) ;  dv_n (+)= sum_s a G_s ;  dk_n (+)= scale sum_s dl q_s
//   dq_s += scale sum_n dl[n,s] k_n   (atomics over token chunks)
// One warp per token; lanes <-> slots for the softmax algebra, lanes <-> channels for the row outputs.
template <int D>
__global__ void __launch_bounds__(256)
slot_attend_bwd_kernel(const float* __restrict__ kv, const float* __restrict__ q, const float* __restrict__ U,
                       const float* __restrict__ cs, const float* __restrict__ dU, float* __restrict__ dkv,
                       float* __restrict__ dq, int64_t N, int S, float scale, float eps, int accumulate,
                       int tokens_per_cta) {
  extern __shared__ __align__(16) float smem[];
  float* sq = smem;                 // [S][D]  scaled q
  float* sG = sq + S * D;           // [S][D]
  float* sdq = sG + S * D;          // [S][D]  per-CTA dq accumulator
  float* sc = sdq + S * D;          // [32]
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < S * D; i += blockDim.x) {
    const int s = i / D;
    const float inv = 1.f / cs[b * S + s];
    sq[i] = q[b * S * D + i] * scale;
    sG[i] = dU[b * S * D + i] * inv;
    sdq[i] = 0.f;
  }
  __syncthreads();
  if (warp == 0) {
    float c = 0.f;
    if (lane < S) {
      for (int d = 0; d < D; ++d) c += U[(b * S + lane) * D + d] * sG[lane * D + d];
    }
    sc[lane] = c;
  }
  __syncthreads();
  constexpr int PER = D / 32;       // channels per lane
  const int64_t n_begin = (int64_t)blockIdx.x * tokens_per_cta;
  const int64_t n_end = min(N, n_begin + tokens_per_cta);
  for (int64_t n = n_begin + warp; n < n_end; n += (blockDim.x >> 5)) {
    const float* kr = kv + (b * N + n) * 2 * D;
    float kx[PER], vx[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) { kx[j] = kr[lane + 32 * j]; vx[j] = kr[D + lane + 32 * j]; }
    // logits and v.G for every slot: lane-partial dot products, butterfly-reduced; lane s keeps slot s
    float logit = -INFINITY, vg = 0.f;
    for (int s = 0; s < S; ++s) {
      float p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        p1 += kx[j] * sq[s * D + lane + 32 * j];
        p2 += vx[j] * sG[s * D + lane + 32 * j];
      }
      p1 = warp_sum(p1);
      p2 = warp_sum(p2);
      if (lane == s) { logit = p1; vg = p2; }
    }
    const float mx = warp_max(logit);
    const float e = lane < S ? expf(logit - mx) : 0.f;
    const float P = e / warp_sum(e);
    const float a = lane < S ? P + eps : 0.f;
    const float da = lane < S ? vg - sc[lane] : 0.f;
    const float dl = P * (da - warp_sum(P * da));      // d logits (before scale; sq carries scale for dk)
    // row outputs: dv = sum_s a_s G_s ; dk = sum_s dl_s (scale q_s)
    float dvx[PER], dkx[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) { dvx[j] = 0.f; dkx[j] = 0.f; }
    for (int s = 0; s < S; ++s) {
      const float as = __shfl_sync(0xffffffffu, a, s), dls = __shfl_sync(0xffffffffu, dl, s);
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        dvx[j] += as * sG[s * D + lane + 32 * j];
        dkx[j] += dls * sq[s * D + lane + 32 * j];
        atomicAdd(&sdq[s * D + lane + 32 * j], dls * scale * kx[j]);
      }
    }
    float* dr = dkv + (b * N + n) * 2 * D;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (accumulate) {
        dr[lane + 32 * j] += dkx[j];
        dr[D + lane + 32 * j] += dvx[j];
      } else {
        dr[lane + 32 * j] = dkx[j];
        dr[D + lane + 32 * j] = dvx[j];
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < S * D; i += blockDim.x) atomicAdd(dq + b * S * D + i, sdq[i]);
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_grad_pack(const float* dy, int64_t ld, void* out_rows, void* out_T, int64_t ldt, float* bias_grad,
                             float* group_grad, int64_t ldg, int64_t M, int64_t N, int rows_per_group, void* stream) {
  SDB_REQUIRE(!out_T || ldt >= M, "sdb_grad_pack: ldt < M");
  SDB_REQUIRE(dy && M > 0 && N > 0 && ld >= N, "sdb_grad_pack: bad args M=%lld N=%lld ld=%lld", (long long)M,
              (long long)N, (long long)ld);
  SDB_REQUIRE(!group_grad || rows_per_group > 0, "sdb_grad_pack: group_grad needs rows_per_group");
  SDB_REQUIRE(cdiv(M, GP_T) <= 65535, "sdb_grad_pack: M too large");
  dim3 grid((unsigned)cdiv(N, GP_T), (unsigned)cdiv(M, GP_T));
  grad_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(dy, ld, (__half*)out_rows, (__half*)out_T, bias_grad, group_grad,
                                                        ldg, M, N, rows_per_group, ldt);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_transpose_packed(const void* in, void* out, int64_t ldt, int64_t M, int64_t K, int to_bf16,
                                    void* stream) {
  SDB_REQUIRE(in && out && M > 0 && K > 0 && ldt >= M && cdiv(M, GP_T) <= 65535, "sdb_transpose_packed: bad args");
  dim3 grid((unsigned)cdiv(K, GP_T), (unsigned)cdiv(M, GP_T));
  transpose_packed_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __half*)in, (__half*)out, M, K, ldt, to_bf16);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_weight_conv3_dgrad(const float* w, void* out, int64_t Cout, int64_t Cin, int bf16, void* stream) {
  SDB_REQUIRE(w && out && Cout > 0 && Cin > 0, "sdb_pack_weight_conv3_dgrad: bad args");
  pack_weight_conv3_dgrad_kernel<<<grid_for_bw(Cout * 9 * Cin, 256), 256, 0, as_stream(stream)>>>(w, (__half*)out, Cout,
                                                                                                  Cin, bf16);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_repack_bf16(const void* in, void* out, int64_t n, void* stream) {
  SDB_REQUIRE(in && out && n > 0 && n % 8 == 0, "sdb_repack_bf16: element count per plane must be a multiple of 8");
  repack_bf16_kernel<<<grid_for_bw(n / 8, 256), 256, 0, as_stream(stream)>>>((const __half*)in, (__half*)out, n / 8);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_wgrad_conv3_scatter(const float* c9, int64_t ldc, float* dw, int64_t Cout, int64_t Cin, int64_t Cin_w,
                                       int accumulate, void* stream) {
  SDB_REQUIRE(c9 && dw && Cout > 0 && Cin > 0 && Cin_w > 0 && Cin_w <= Cin && ldc >= Cout,
              "sdb_wgrad_conv3_scatter: bad args");
  wgrad_conv3_scatter_kernel<<<grid_for_bw(Cout * Cin_w * 9, 256), 256, 0, as_stream(stream)>>>(c9, ldc, dw, Cout, Cin,
                                                                                                Cin_w, accumulate);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_add3(float* out, const float* a, const float* b, const float* c, int64_t n, void* stream) {
  SDB_REQUIRE(out && a && n > 0 && n % 4 == 0, "sdb_add3: bad args (n %% 4 == 0)");
  add3_kernel<<<grid_for_bw(n / 4, 256), 256, 0, as_stream(stream)>>>(out, a, b, c, n / 4);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_act_bwd(const float* dy, int64_t ldy, const float* pre, int64_t ldp, float* dx, int64_t ldx, int64_t M,
                           int64_t N, int act, void* stream) {
  SDB_REQUIRE(dy && pre && dx && M > 0 && N > 0 && (act == 1 || act == 2), "sdb_act_bwd: bad args");
  const bool vec = N % 4 == 0 && ldy % 4 == 0 && ldp % 4 == 0 && ldx % 4 == 0 && M * N / 4 < (1ll << 31) &&
                   ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (vec)
    act_bwd_vec4_kernel<<<grid_for_bw(M * N / 4, 256), 256, 0, as_stream(stream)>>>(dy, ldy, pre, ldp, dx, ldx,
                                                                                  (unsigned)(M * N / 4), (unsigned)(N / 4), act);
  else
    act_bwd_kernel<<<grid_for_bw(M * N, 256), 256, 0, as_stream(stream)>>>(dy, ldy, pre, ldp, dx, ldx, M, N, act);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_groupnorm_bwd(const float* x1, int64_t C1, const float* x2, int64_t C2, const float* da,
                                 const float* stats, const float* gamma, const float* beta, float* sums_work,
                                 float* dx1, float* dx2, const float* add1, const float* add2, float* dgamma,
                                 float* dbeta, int64_t B, int64_t HW, int G, int silu, float drop_p, uint64_t seed,
                                 const uint64_t* seed_dev, void* stream) {
  SDB_REQUIRE(x1 && da && stats && gamma && beta && sums_work && dx1 && dgamma && dbeta, "sdb_groupnorm_bwd: null argument");
  const int64_t C = C1 + C2;
  SDB_REQUIRE(G >= 1 && G <= 64 && C % G == 0 && C1 % 4 == 0 && C2 % 4 == 0 && C <= 1024 && B <= 65535 &&
                  (C2 == 0 || x2), "sdb_groupnorm_bwd: unsupported channels");
  GnBwdArgs a{};
  a.x1 = x1; a.x2 = x2; a.da = da; a.stats = stats; a.gamma = gamma; a.beta = beta; a.sums = sums_work;
  a.dx1 = dx1; a.dx2 = dx2; a.add1 = add1; a.add2 = add2;
  a.B = B; a.C1 = (int)C1; a.C2 = (int)C2; a.HW = (int)HW; a.G = G; a.silu = silu; a.drop_p = drop_p; a.seed = seed;
  a.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  const int c4n = (int)(C / 4);
  const int rpb = 256 / c4n > 0 ? 256 / c4n : 1;
  int threads = c4n * rpb;
  if (threads < 64) threads = 64;
  // ONE wave: chunks * B CTAs must not exceed the resident slots (4 CTAs per SM) -- rounding UP left a second wave of a few
  // CTAs that ran alone at a fraction of the bandwidth (ncu: 1.08 waves, SMs active 70 % of the kernel, profiles/README 23)
  int64_t chunks = ((int64_t)num_sms() * 4) / B;
  if (chunks > cdiv(HW, rpb)) chunks = cdiv(HW, rpb);
  if (chunks < 1) chunks = 1;
  a.rows_per_chunk = (int)cdiv(HW, chunks);
  chunks = cdiv(HW, a.rows_per_chunk);
  cudaStream_t st = as_stream(stream);
  SDB_CHECK(cudaMemsetAsync(sums_work, 0, (size_t)B * C * 2 * sizeof(float), st));
  dim3 grid((unsigned)chunks, (unsigned)B);
  const bool drop = drop_p > 0.f;
#define SDB_GNB_LAUNCH(APPLYv, ACTv, DROPv) groupnorm_bwd_kernel<APPLYv, ACTv, DROPv><<<grid, threads, 0, st>>>(a)
#define SDB_GNB(APPLYv)                                                                             \
  do {                                                                                              \
    if (silu == 1) { if (drop) SDB_GNB_LAUNCH(APPLYv, 1, true); else SDB_GNB_LAUNCH(APPLYv, 1, false); }       \
    else if (silu == 2) { if (drop) SDB_GNB_LAUNCH(APPLYv, 2, true); else SDB_GNB_LAUNCH(APPLYv, 2, false); }  \
    else { if (drop) SDB_GNB_LAUNCH(APPLYv, 0, true); else SDB_GNB_LAUNCH(APPLYv, 0, false); }                 \
  } while (0)
  SDB_GNB(false);
  SDB_LAUNCH_CHECK();
  SDB_GNB(true);
  SDB_LAUNCH_CHECK();
#undef SDB_GNB
#undef SDB_GNB_LAUNCH
  groupnorm_bwd_param_kernel<<<(unsigned)cdiv(C, 128), 128, 0, st>>>(sums_work, dgamma, dbeta, B, (int)C);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_layernorm_bwd(const float* x, const float* dn, const float* gamma, float eps, const float* add,
                                 float* dx, float* dgamma, float* dbeta, int64_t M, int64_t C, void* stream) {
  SDB_REQUIRE(x && dn && gamma && dx && dgamma && dbeta && M > 0, "sdb_layernorm_bwd: null argument");
  SDB_REQUIRE(C % 4 == 0 && C <= 512, "sdb_layernorm_bwd: C=%lld must be a multiple of 4 and <= 512", (long long)C);
  int grid = grid_for_bw(M, 8, 2);
  layernorm_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dn, gamma, eps, add, dx, dgamma, dbeta, M, (int)C);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_attention_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                 const float* dout, int64_t ldo, float* dq, int64_t lddq, float* dk, int64_t lddk,
                                 float* dv, int64_t lddv, float* work, int64_t B, int64_t Lq, int64_t Lk, int heads,
                                 int d, float scale, void* stream) {
  SDB_REQUIRE(q && k && v && dout && dq && dk && dv && work, "sdb_attention_bwd: null argument");
  SDB_REQUIRE(d == 32, "sdb_attention_bwd: head dim %d unsupported (32)", d);
  SDB_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "sdb_attention_bwd: strides must be multiples of 4");
  SDB_REQUIRE(B <= 65535 && heads <= 65535, "sdb_attention_bwd: grid too large");
  float* lse = work;
  float* dsum = work + B * heads * Lq;
  cudaStream_t st = as_stream(stream);
  dim3 gq((unsigned)cdiv(Lq, AB_THREADS), (unsigned)heads, (unsigned)B);
  attention_bwd_q_kernel<32><<<gq, AB_THREADS, 0, st>>>(q, ldq, k, ldk, v, ldv, dout, ldo, dq, lddq, lse, dsum, Lq, Lk,
                                                        heads, scale);
  SDB_LAUNCH_CHECK();
  dim3 gk((unsigned)cdiv(Lk, AB_THREADS), (unsigned)heads, (unsigned)B);
  attention_bwd_kv_kernel<32><<<gk, AB_THREADS, 0, st>>>(q, ldq, k, ldk, v, ldv, dout, ldo, lse, dsum, dk, lddk, dv, lddv,
                                                         Lq, Lk, heads, scale);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_geglu_bwd(const float* u, const float* dg, float* du, int64_t M, int64_t F, void* stream) {
  SDB_REQUIRE(u && dg && du && M > 0 && F > 0 && F % 4 == 0, "sdb_geglu_bwd: bad args");
  geglu_bwd_kernel<<<grid_for_bw(M * F / 4, 256), 256, 0, as_stream(stream)>>>(u, dg, du, M, F);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_up2_adjoint(const float* dup, float* dx, int64_t B, int64_t H, int64_t W, int64_t C, void* stream) {
  SDB_REQUIRE(dup && dx && B > 0 && C % 4 == 0, "sdb_up2_adjoint: bad args");
  up2_adjoint_kernel<<<grid_for_bw(B * H * W * C / 4, 256), 256, 0, as_stream(stream)>>>(dup, dx, B, (int)H, (int)W, (int)C);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_zero_up2(const float* dy, void* out, int64_t B, int64_t H, int64_t W, int64_t C, void* stream) {
  SDB_REQUIRE(dy && out && B > 0 && C % 4 == 0, "sdb_pack_zero_up2: bad args");
  pack_zero_up2_kernel<<<grid_for_bw(B * H * W * C, 256), 256, 0, as_stream(stream)>>>(dy, (__half*)out, B, (int)H, (int)W,
                                                                                       (int)C);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_pack_nchw_pad(const float* x, void* out, int64_t B, int64_t Cs, int64_t HW, int64_t Cp, void* stream) {
  SDB_REQUIRE(x && out && B > 0 && Cs > 0 && Cp >= Cs, "sdb_pack_nchw_pad: bad args");
  pack_nchw_pad_kernel<<<grid_for_bw(B * HW * Cp, 256), 256, 0, as_stream(stream)>>>(x, (__half*)out, B, (int)Cs, (int)HW,
                                                                                     (int)Cp);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_nhwc_to_nchw(const float* in, int64_t ld, float* out, int64_t B, int64_t Cs, int64_t HW, void* stream) {
  SDB_REQUIRE(in && out && B > 0 && Cs > 0 && ld >= Cs, "sdb_nhwc_to_nchw: bad args");
  nhwc_to_nchw_kernel<<<grid_for_bw(B * Cs * HW, 256), 256, 0, as_stream(stream)>>>(in, ld, out, B, (int)Cs, (int)HW);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_nchw_to_nhwc_pad(const float* in, float* out, int64_t B, int64_t Cs, int64_t HW, int64_t Cp,
                                    void* stream) {
  SDB_REQUIRE(in && out && B > 0 && Cs > 0 && Cp >= Cs, "sdb_nchw_to_nhwc_pad: bad args");
  nchw_to_nhwc_pad_kernel<<<grid_for_bw(B * HW * Cp, 256), 256, 0, as_stream(stream)>>>(in, out, B, (int)Cs, (int)HW,
                                                                                        (int)Cp);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_im2col_t(const void* in, void* out, int64_t B, int64_t H, int64_t W, int64_t C, void* stream) {
  SDB_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0, "sdb_im2col_t: bad args");
  im2col_T_kernel<<<grid_for_bw(9 * C * B * H * W, 256), 256, 0, as_stream(stream)>>>((const __half*)in, (__half*)out, B,
                                                                                      (int)H, (int)W, (int)C);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_gru_gates_bwd(const float* gi, const float* gh, const float* h, const float* dh_new, float* dgi,
                                 float* dgh, float* dh, int64_t R, int64_t D, void* stream) {
  SDB_REQUIRE(gi && gh && h && dh_new && dgi && dgh && dh && R > 0 && D > 0, "sdb_gru_gates_bwd: bad args");
  gru_gates_bwd_kernel<<<grid_for_bw(R * D, 256), 256, 0, as_stream(stream)>>>(gi, gh, h, dh_new, dgi, dgh, dh, R, D);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_slot_attend_bwd(const float* kv, const float* q, const float* upd, const float* colsum,
                                   const float* d_upd, float* dkv, float* dq, int64_t B, int64_t N, int64_t S, int64_t D,
                                   float scale, float eps, int accumulate, void* stream) {
  SDB_REQUIRE(kv && q && upd && colsum && d_upd && dkv && dq, "sdb_slot_attend_bwd: null argument");
  SDB_REQUIRE(S >= 1 && S <= 32 && B <= 65535, "sdb_slot_attend_bwd: num_slots must be in 1..32");
  SDB_REQUIRE(D == 64 || D == 128 || D == 192 || D == 256, "sdb_slot_attend_bwd: slot_size unsupported");
  cudaStream_t st = as_stream(stream);
  SDB_CHECK(cudaMemsetAsync(dq, 0, (size_t)B * S * D * sizeof(float), st));
  int64_t chunks = cdiv((int64_t)num_sms() * 2, B);
  if (chunks > cdiv(N, 8)) chunks = cdiv(N, 8);
  if (chunks < 1) chunks = 1;
  const int tpc = (int)cdiv(N, chunks);
  chunks = cdiv(N, tpc);
  const size_t smem = ((size_t)3 * S * D + 32) * sizeof(float);
  dim3 grid((unsigned)chunks, (unsigned)B);
#define SAB_CASE(DD)                                                                                              \
  if (D == DD) {                                                                                                  \
    static size_t attr = 0;                                                                                       \
    if (smem > 48 * 1024 && smem > attr) {                                                                        \
      SDB_CHECK(cudaFuncSetAttribute(slot_attend_bwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)smem));                                                                 \
      attr = smem;                                                                                                \
    }                                                                                                             \
    slot_attend_bwd_kernel<DD><<<grid, 256, smem, st>>>(kv, q, upd, colsum, d_upd, dkv, dq, N, (int)S, scale, eps, \
                                                        accumulate, tpc);                                         \
  }
  SAB_CASE(64) else SAB_CASE(128) else SAB_CASE(192) else SAB_CASE(256)
#undef SAB_CASE
  SDB_LAUNCH_CHECK();
  return 0;
}
