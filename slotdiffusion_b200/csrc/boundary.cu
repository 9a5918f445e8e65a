// Small fused kernels at the boundary of the hot path:
//   * q_sample            x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps        (ddpm.py:161-165, called from ldm.py:68-69)
//   * eps-MSE loss        mean((pred - target)^2) and its gradient            (ldm.py:76-77: F.mse_loss)
//   * mask upsample + argmax: bilinear (align_corners = False) resize of the slot masks to the image resolution and the
//                         per-pixel argmax over slots                          (sa_diffusion.py:172-180, test_seg.py:27)
// All HBM-bound: one read of each input, 128-bit accesses, grid = a multiple of the SM count.
#include "common.cuh"

namespace sdb {

static inline int grid_for_b(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = cdiv(work_items, threads);
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// x0, eps, out: [B, n] fp32 (n = C*H*W, n % 4 == 0); t [B] int64; ca / cs: [T] fp32 tables
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ eps,
                                const long long* __restrict__ t, const float* __restrict__ ca,
                                const float* __restrict__ cs, float* __restrict__ out, int64_t B, int64_t n4) {
  const int64_t total = B * n4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / n4;
    const long long tb = t[b];
    const float a = ca[tb], s = cs[tb];
    const float4 x = reinterpret_cast<const float4*>(x0)[i];
    const float4 e = reinterpret_cast<const float4*>(eps)[i];
    // the reference evaluates a * x0 + s * eps as two rounded products and one addition (no FMA contraction in eager
    // PyTorch): __fmul_rn / __fadd_rn keep the same roundings, so x_t is bit-identical to ddpm.py:163-165
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(s, e.x));
    o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(s, e.y));
    o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(s, e.z));
    o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(s, e.w));
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// loss[0] += sum((p - t)^2) * inv_n  (loss zeroed by the caller);  optional diff = p - t kept for the backward
__global__ void mse_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ diff,
                               float* __restrict__ loss, int64_t n4, float inv_n) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(p)[i];
    const float4 b = reinterpret_cast<const float4*>(t)[i];
    const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    if (diff) reinterpret_cast<float4*>(diff)[i] = d;
    acc += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(loss, v * inv_n);
  }
}

// dpred = diff * (2 / n) * gout[0]
__global__ void mse_bwd_kernel(const float* __restrict__ diff, const float* __restrict__ gout, float* __restrict__ dp,
                               int64_t n4, float two_inv_n) {
  const float g = gout[0] * two_inv_n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 d = reinterpret_cast<const float4*>(diff)[i];
    reinterpret_cast<float4*>(dp)[i] = make_float4(d.x * g, d.y * g, d.z * g, d.w * g);
  }
}

// masks [B, S, h, w] fp32 -> up [B, S, H, W] (optional) and idx [B, H, W] int64 = argmax over S of the upsampled value
// (first maximum wins, like torch.argmax).  Bilinear weights and evaluation order follow ATen's upsample_bilinear2d
// (UpSampleBilinear2d.cu: area_pixel_compute_source_index with align_corners = false, then
//  h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)), written with explicit roundings (no FMA contraction).
__global__ void mask_upsample_argmax_kernel(const float* __restrict__ m, float* __restrict__ up,
                                            long long* __restrict__ idx, int64_t B, int S, int h, int w, int H, int W,
                                            float rh, float rw) {
  const int64_t total = B * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = int(i % W);
    const int y = int((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    float sy = __fmul_rn(rh, (float)y + 0.5f) - 0.5f;
    float sx = __fmul_rn(rw, (float)x + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < h - 1) ? 1 : 0, xp = (x0 < w - 1) ? 1 : 0;
    const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    float best = 0.f;
    int bi = 0;
    for (int s = 0; s < S; ++s) {
      const float* p = m + ((b * S + s) * h + y0) * (int64_t)w + x0;
      const float v00 = p[0], v01 = p[xp], v10 = p[yp * w], v11 = p[yp * w + xp];
      const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
      const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
      const float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
      if (up) up[((b * S + s) * H + y) * (int64_t)W + x] = v;
      if (s == 0 || v > best) { best = v; bi = s; }
    }
    if (idx) idx[i] = bi;
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_q_sample(const float* x0, const float* eps, const int64_t* t, const float* sqrt_abar,
                            const float* sqrt_1m_abar, float* out, int64_t B, int64_t n, void* stream) {
  SDB_REQUIRE(x0 && eps && t && sqrt_abar && sqrt_1m_abar && out && B > 0 && n > 0 && n % 4 == 0,
              "sdb_q_sample: bad args B=%lld n=%lld (n %% 4 == 0)", (long long)B, (long long)n);
  q_sample_kernel<<<grid_for_b(B * n / 4, 256), 256, 0, as_stream(stream)>>>(
      x0, eps, reinterpret_cast<const long long*>(t), sqrt_abar, sqrt_1m_abar, out, B, n / 4);
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_mse_loss_fwd(const float* pred, const float* target, float* diff, float* loss, int64_t n,
                                void* stream) {
  SDB_REQUIRE(pred && target && loss && n > 0 && n % 4 == 0, "sdb_mse_loss_fwd: bad args n=%lld (n %% 4 == 0)",
              (long long)n);
  SDB_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), as_stream(stream)));
  mse_fwd_kernel<<<grid_for_b(n / 4, 256, 4), 256, 0, as_stream(stream)>>>(pred, target, diff, loss, n / 4,
                                                                            (float)(1.0 / (double)n));
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_mse_loss_bwd(const float* diff, const float* grad_out, float* dpred, int64_t n, void* stream) {
  SDB_REQUIRE(diff && grad_out && dpred && n > 0 && n % 4 == 0, "sdb_mse_loss_bwd: bad args n=%lld", (long long)n);
  mse_bwd_kernel<<<grid_for_b(n / 4, 256), 256, 0, as_stream(stream)>>>(diff, grad_out, dpred, n / 4,
                                                                         (float)(2.0 / (double)n));
  SDB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sdb_mask_upsample_argmax(const float* masks, float* up, int64_t* idx, int64_t B, int S, int h, int w,
                                        int H, int W, void* stream) {
  SDB_REQUIRE(masks && (up || idx) && B > 0 && S > 0 && h > 0 && w > 0 && H > 0 && W > 0,
              "sdb_mask_upsample_argmax: bad args");
  // area_pixel_compute_scale (align_corners = false, no explicit scale factor): input / output in fp32
  const float rh = (float)h / (float)H, rw = (float)w / (float)W;
  mask_upsample_argmax_kernel<<<grid_for_b(B * H * W, 256), 256, 0, as_stream(stream)>>>(
      masks, up, reinterpret_cast<long long*>(idx), B, S, h, w, H, W, rh, rw);
  SDB_LAUNCH_CHECK();
  return 0;
}
