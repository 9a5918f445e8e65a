// Attention core of the UNet transformer blocks (attention.py:188-205): softmax(scale * q k^T) v per head.
// fp32 CUDA-core kernel with online softmax; one thread owns one query row of one head (d = 32 registers),
// K/V tiles of the (batch, head) are staged in shared memory and read as warp-wide broadcasts.
// The QK^T/AV cores are 0.4 GF/sample (1.8 % of the UNet) -- the projections around them run on
// the tensor cores through sdb_gemm; the output is written directly in packed GEMM-operand format for to_out.
#include "common.cuh"

namespace sdb {

SDB_DEFINE_PACK_MODE_SETTER(set_pack_mode_attention)

constexpr int ATT_THREADS = 64;
constexpr int ATT_KV_TILE = 64;

template <int D>
__global__ void __launch_bounds__(ATT_THREADS)
attention_pack_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                      const float* __restrict__ v, int64_t ldv, __half* __restrict__ out, int64_t Lq, int64_t Lk,
                      int heads, float scale, int64_t plane) {
  const int pmode = g_pack_mode;   // operand format of the consumer GEMM: read once (common.cuh)
  __shared__ __align__(16) float sk[ATT_KV_TILE][D];
  __shared__ __align__(16) float sv[ATT_KV_TILE][D];
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t row = blockIdx.x * (int64_t)ATT_THREADS + threadIdx.x;
  const bool active = row < Lq;
  const int C = heads * D;

  float qr[D], acc[D];
  if (active) {
    const float4* qp = reinterpret_cast<const float4*>(q + (b * Lq + row) * ldq + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 t = qp[i];
      qr[4 * i] = t.x * scale; qr[4 * i + 1] = t.y * scale; qr[4 * i + 2] = t.z * scale; qr[4 * i + 3] = t.w * scale;
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) qr[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;

  for (int64_t j0 = 0; j0 < Lk; j0 += ATT_KV_TILE) {
    const int nk = (int)min((int64_t)ATT_KV_TILE, Lk - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * (D / 4); i += ATT_THREADS) {
      const int r = i / (D / 4), c = i % (D / 4);
      reinterpret_cast<float4*>(&sk[r][0])[c] =
          reinterpret_cast<const float4*>(k + (b * Lk + j0 + r) * ldk + h * D)[c];
      reinterpret_cast<float4*>(&sv[r][0])[c] =
          reinterpret_cast<const float4*>(v + (b * Lk + j0 + r) * ldv + h * D)[c];
    }
    __syncthreads();
    // process the tile in register chunks of 16 keys: scores -> chunk max -> one rescale -> accumulate
    for (int c0 = 0; c0 < nk; c0 += 16) {
      float s[16];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = c0 + jj;
        float d = -INFINITY;
        if (j < nk) {
          d = 0.f;
          const float4* kp = reinterpret_cast<const float4*>(&sk[j][0]);
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            const float4 t = kp[i];
            d += qr[4 * i] * t.x + qr[4 * i + 1] * t.y + qr[4 * i + 2] * t.z + qr[4 * i + 3] * t.w;
          }
        }
        s[jj] = d;
        cmax = fmaxf(cmax, d);
      }
      const float mnew = fmaxf(mrun, cmax);
      const float corr = expf(mrun - mnew);     // exp(-inf) = 0 on the first chunk
      lrun *= corr;
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] *= corr;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = c0 + jj;
        if (j < nk) {
          const float p = expf(s[jj] - mnew);
          lrun += p;
          const float4* vp = reinterpret_cast<const float4*>(&sv[j][0]);
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            const float4 t = vp[i];
            acc[4 * i] += p * t.x; acc[4 * i + 1] += p * t.y; acc[4 * i + 2] += p * t.z; acc[4 * i + 3] += p * t.w;
          }
        }
      }
      mrun = mnew;
    }
  }
  if (active) {
    const float inv = 1.f / lrun;
    const int64_t o = (b * Lq + row) * C + h * D;
#pragma unroll
    for (int i = 0; i < D; i += 4)     // operand format of the consumer GEMM follows the stream-ordered pack mode (common.cuh)
      store_split4(out, out + plane, o + i, make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv), pmode);
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_attention_pack(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                                  int64_t ldv, void* out, int64_t B, int64_t Lq, int64_t Lk, int heads, int d,
                                  float scale, void* stream) {
  SDB_REQUIRE(q && k && v && out && B > 0 && Lq > 0 && Lk > 0 && heads > 0, "sdb_attention_pack: bad args");
  SDB_REQUIRE(d == 32 || d == 64, "sdb_attention_pack: head dim %d unsupported (32 or 64)", d);
  SDB_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "sdb_attention_pack: row strides must be multiples of 4");
  SDB_REQUIRE(B <= 65535 && heads <= 65535, "sdb_attention_pack: grid too large");
  dim3 grid((unsigned)cdiv(Lq, ATT_THREADS), (unsigned)heads, (unsigned)B);
  const int64_t plane = B * Lq * heads * d;
  if (d == 32)
    attention_pack_kernel<32><<<grid, ATT_THREADS, 0, as_stream(stream)>>>(q, ldq, k, ldk, v, ldv, (__half*)out, Lq,
                                                                           Lk, heads, scale, plane);
  else
    attention_pack_kernel<64><<<grid, ATT_THREADS, 0, as_stream(stream)>>>(q, ldq, k, ldk, v, ldv, (__half*)out, Lq,
                                                                           Lk, heads, scale, plane);
  SDB_LAUNCH_CHECK();
  return 0;
}
