// Attention core for FEW keys: the slot cross-attention of the UNet transformer blocks (attention.py:188-205 with
// context = slots: Lk = num_slots = 7 .. 24 keys per sample, head dim 32).  0.1 % of the UNet's FLOPs but 16 launches per
// evaluation, and neither general kernel suits it: the tensor-core kernel (attention_tc.cu) pays a TMEM / mbarrier / two
// MMA round trips per 128 queries for a [128 x 16] product, the CUDA-core kernel (attention.cu) reads and writes 128-byte
// row segments at row stride (a quarter of every sector used).  Both took 84 us at L = 256, B = 256 where the data moved
// (q in, packed o out: 134 MB) is 21 us of HBM time.
//
// Here a CTA owns 32 query rows of ONE sample with ALL heads: warp w <-> head w, lane <-> row.
//   * the [32 x C] fp32 q tile is loaded with fully coalesced 16-byte loads into shared memory (XOR-swizzled chunks so the
//     per-thread 128-byte reads are conflict free), K and V of the sample ([Lk x C] each) likewise;
//   * every lane of a warp reads the SAME k / v address (one head): all shared-memory reads of the products are
//     broadcasts, one wavefront per 16 bytes per warp;
//   * scores, softmax and the weighted sum are fp32 FMAs in registers (Lk <= 32 scores per thread);
//   * the result is split into the fp16 hi / lo planes of the consumer GEMM's operand, staged through the same tile and
//     written as whole 512-byte rows.
#include "common.cuh"

namespace sdb {

constexpr int FK_ROWS = 32;      // query rows per CTA (= lanes)
constexpr int FK_D = 32;         // head dim
constexpr int FK_MAXK = 32;      // keys

// 16-byte chunk `c` of row `r` of a [32][C floats] tile whose chunks are XOR-swizzled with the row (8 rows = 8 bank groups)
__device__ __forceinline__ void fk_cp16(void* smem_dst, const void* gsrc) {     // 16-byte asynchronous global -> shared copy
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}

// (mask 7; the output planes of an odd head count have rows of 4 (mod 8) chunks: mask 3 keeps the XOR inside the row)
__device__ __forceinline__ int fk_chunk(int r, int c, int chunks_per_row, int mask = 7) {
  return r * chunks_per_row + (c ^ (r & mask));
}

// MAXT: block-size class (256: up to 8 heads, three CTAs resident -- the kernel is issue / latency bound, 24 % active warps
// with two: 66 -> 55 us; 384: up to 12 heads, two CTAs; 512: up to 16 heads)
template <int MAXK, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 3 : (MAXT == 384 ? 2 : 1))
attention_fewkeys_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                         const float* __restrict__ v, int64_t ldv, __half* __restrict__ out, int Lq, int Lk, int heads,
                         float scale, int64_t plane) {
  extern __shared__ __align__(16) uint8_t fk_smem[];
  const int C = heads * FK_D, cpr = C / 4;                  // 16-byte chunks per row
  float4* qt = reinterpret_cast<float4*>(fk_smem);          // [32][cpr] swizzled; later the hi | lo output planes
  float4* kt = qt + FK_ROWS * cpr;                          // [Lk][cpr]
  float4* vt = kt + Lk * cpr;                               // [Lk][cpr]
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, h = tid >> 5;
  const int64_t b = blockIdx.y;
  const int r0 = blockIdx.x * FK_ROWS;
  const int rows = min(FK_ROWS, Lq - r0);
  pdl_wait();        // PDL (common.cuh)
  pdl_trigger();

  // all tile loads are asynchronous copies in flight together (a register-staged loop exposed one DRAM round trip per
  // trip: 11 per CTA, 80 us per launch instead of the ~25 us the bytes take)
  for (int i = tid; i < rows * cpr; i += nthr) {
    const int r = i / cpr, c = i - r * cpr;
    fk_cp16(qt + fk_chunk(r, c, cpr), q + (b * Lq + r0 + r) * ldq + c * 4);
  }
  for (int i = tid; i < Lk * cpr; i += nthr) {
    const int j = i / cpr, c = i - j * cpr;
    fk_cp16(kt + i, k + (b * Lk + j) * ldk + c * 4);
    fk_cp16(vt + i, v + (b * Lk + j) * ldv + c * 4);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const int r = lane;
  float o[FK_D];
  if (r < rows) {
    float qr[FK_D];
#pragma unroll
    for (int c = 0; c < FK_D / 4; ++c) {
      const float4 t = qt[fk_chunk(r, h * (FK_D / 4) + c, cpr)];
      qr[4 * c] = t.x * scale; qr[4 * c + 1] = t.y * scale; qr[4 * c + 2] = t.z * scale; qr[4 * c + 3] = t.w * scale;
    }
    float s[MAXK];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      float d = -INFINITY;
      if (j < Lk) {
        const float4* kp = kt + j * cpr + h * (FK_D / 4);
        float d0 = 0.f, d1 = 0.f;                               // two chains
#pragma unroll
        for (int c = 0; c < FK_D / 4; c += 2) {
          const float4 t0 = kp[c], t1 = kp[c + 1];
          d0 = fmaf(qr[4 * c], t0.x, d0); d0 = fmaf(qr[4 * c + 1], t0.y, d0);
          d0 = fmaf(qr[4 * c + 2], t0.z, d0); d0 = fmaf(qr[4 * c + 3], t0.w, d0);
          d1 = fmaf(qr[4 * c + 4], t1.x, d1); d1 = fmaf(qr[4 * c + 5], t1.y, d1);
          d1 = fmaf(qr[4 * c + 6], t1.z, d1); d1 = fmaf(qr[4 * c + 7], t1.w, d1);
        }
        d = d0 + d1;
        mx = fmaxf(mx, d);
      }
      s[j] = d;
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      s[j] = j < Lk ? expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < FK_D; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      if (j < Lk) {
        const float p = s[j];
        const float4* vp = vt + j * cpr + h * (FK_D / 4);
#pragma unroll
        for (int c = 0; c < FK_D / 4; ++c) {
          const float4 t = vp[c];
          o[4 * c] = fmaf(p, t.x, o[4 * c]); o[4 * c + 1] = fmaf(p, t.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(p, t.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(p, t.w, o[4 * c + 3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < FK_D; ++i) o[i] *= inv;
  }
  __syncthreads();      // every q chunk has been read: the tile becomes the output staging area
  // hi plane: rows of C halves = cpr / 2 chunks of 16 bytes; lo plane behind it.  Same XOR swizzle (4 chunks per head).
  uint4* ht = reinterpret_cast<uint4*>(qt);
  const int opr = cpr / 2;                                   // 16-byte chunks per output row and plane
  uint4* lt = ht + FK_ROWS * opr;
  const int om = (heads & 1) ? 3 : 7;
  if (r < rows) {
#pragma unroll
    for (int c = 0; c < FK_D / 8; ++c) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_f16x2(o[8 * c + 2 * e], o[8 * c + 2 * e + 1], hi[e], lo[e]);
      const int idx = fk_chunk(r, h * (FK_D / 8) + c, opr, om);
      ht[idx] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      lt[idx] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  __syncthreads();
  for (int i = tid; i < rows * opr; i += nthr) {
    const int rr = i / opr, c = i - rr * opr;
    const int64_t dst = (b * Lq + r0 + rr) * (int64_t)C + c * 8;
    *reinterpret_cast<uint4*>(out + dst) = ht[fk_chunk(rr, c, opr, om)];
    *reinterpret_cast<uint4*>(out + plane + dst) = lt[fk_chunk(rr, c, opr, om)];
  }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_attention_fewkeys_supported(int64_t heads, int64_t d, int64_t Lk, int64_t ldq, int64_t ldk, int64_t ldv) {
  // the packed output is written in the default fp16 hi / lo format only (FP8C mode: the general kernels)
  return d == FK_D && heads >= 1 && heads <= 16 && Lk >= 1 && Lk <= FK_MAXK && ldq % 4 == 0 && ldk % 4 == 0 &&
         ldv % 4 == 0 && host_pack_mode() == SDB_FMT_F16X2;
}

extern "C" int sdb_attention_fewkeys(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                     void* out, int64_t B, int64_t Lq, int64_t Lk, int heads, int d, float scale,
                                     void* stream) {
  SDB_REQUIRE(q && k && v && out && B > 0 && Lq > 0 && Lk > 0 && heads > 0, "sdb_attention_fewkeys: bad args");
  SDB_REQUIRE(sdb_attention_fewkeys_supported(heads, d, Lk, ldq, ldk, ldv),
              "sdb_attention_fewkeys: unsupported (heads %d, head dim %d, %lld keys, strides %lld %lld %lld, pack mode %d)",
              heads, d, (long long)Lk, (long long)ldq, (long long)ldk, (long long)ldv, host_pack_mode());
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "sdb_attention_fewkeys: operands must be 16-byte aligned");
  SDB_REQUIRE(B <= 65535 && Lq < (1ll << 30), "sdb_attention_fewkeys: grid too large");
  const int C = heads * d;
  const size_t smem = (size_t)(FK_ROWS + 2 * Lk) * C * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    SDB_CHECK(cudaFuncSetAttribute(attention_fewkeys_kernel<16, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_fewkeys_kernel<32, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_fewkeys_kernel<16, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_fewkeys_kernel<32, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_fewkeys_kernel<16, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDB_CHECK(cudaFuncSetAttribute(attention_fewkeys_kernel<32, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((unsigned)cdiv(Lq, FK_ROWS), (unsigned)B);
  const dim3 block(32 * heads);
  cudaStream_t st = as_stream(stream);
  const int64_t plane = B * Lq * (int64_t)C;
#define SDB_FK_LAUNCH(K, T) \
  (void)launch_k(attention_fewkeys_kernel<K, T>, grid, block, smem, st, q, ldq, k, ldk, v, ldv, (__half*)out, (int)Lq, (int)Lk, heads, scale, plane)
  if (heads <= 8) { if (Lk <= 16) SDB_FK_LAUNCH(16, 256); else SDB_FK_LAUNCH(32, 256); }
  else if (heads <= 12) { if (Lk <= 16) SDB_FK_LAUNCH(16, 384); else SDB_FK_LAUNCH(32, 384); }
  else { if (Lk <= 16) SDB_FK_LAUNCH(16, 512); else SDB_FK_LAUNCH(32, 512); }
#undef SDB_FK_LAUNCH
  SDB_LAUNCH_CHECK();
  return 0;
}
