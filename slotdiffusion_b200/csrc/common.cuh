// Shared host/device helpers for libsdb200.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sdb200.h"

namespace sdb {

// ---- error plumbing (no C++ exceptions cross the ABI) ----
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define SDB_CHECK(expr)                                   \
  do {                                                    \
    int _rc = ::sdb::check_cuda((expr), #expr);           \
    if (_rc) return _rc;                                  \
  } while (0)
#define SDB_REQUIRE(cond, ...)                            \
  do {                                                    \
    if (!(cond)) {                                        \
      ::sdb::set_error(__VA_ARGS__);                      \
      return SDB_ERR_INVALID;                             \
    }                                                     \
  } while (0)
#define SDB_LAUNCH_CHECK() SDB_CHECK(cudaGetLastError())

int num_sms();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- fp32 -> (fp16 hi, fp16 lo) split: x ~= hi + lo with ~22 mantissa bits ----
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  // saturate instead of producing inf for |x| > 65504 (never reached by normalised activations)
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// packed form: two elements per conversion instruction (F2FP.PACK_AB / HADD2.F32); same roundings as split_f16
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- operand format of the second plane (sdb200.h "packed") ----
//   SDB_FMT_F16X2 (0): fp16 lo plane -- three kind::f16 passes per product (hi*hi + lo*hi + hi*lo)
//   SDB_FMT_F8C   (2): the SAME bytes hold two e4m3 half-planes, h8 = e4m3(hi * 2^eh) then l8 = e4m3(lo * 2^el): one
//                      kind::f16 pass (hi*hi) plus two kind::f8f6f4 correction products (l8*h8' + h8*l8') in a second
//                      accumulator = 2 pass-equivalents.  Activations use the static exponents below; weights a
//                      per-tensor exponent chosen at pack time.
// The format the operand PRODUCERS write is a stream-ordered device flag (one copy per translation unit: the library is
// built without relocatable device code), flipped by sdb_set_pack_mode() -- a 1-thread kernel, so it is CUDA-graph
// capturable and ordered with the producers on the same stream.
static __device__ int g_pack_mode = 0;
constexpr int F8_ACT_HI_EXP = 2;     // |x| < 112 before the e4m3 copy of hi saturates (448 / 4)
constexpr int F8_ACT_LO_EXP = 12;    // lo = x - fp16(x) <= 2^-12 |x|

__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
  const uint32_t p0 = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);   // low byte = a
  const uint32_t p1 = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return p0 | (p1 << 16);
}

// fp16 hi plane + two e4m3 half-planes in place of the lo plane; `lo - hi` is the plane size in elements (= bytes of one
// e4m3 half-plane), sh / sl = 2^eh / 2^el
__device__ __forceinline__ void store_split4_f8(__half* hi, __half* lo, long long idx, float4 v, float sh, float sl) {
  v.x = fminf(fmaxf(v.x, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
  v.z = fminf(fmaxf(v.z, -65504.f), 65504.f); v.w = fminf(fmaxf(v.w, -65504.f), 65504.f);
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  uint2 ph;
  ph.x = *reinterpret_cast<const uint32_t*>(&h0);
  ph.y = *reinterpret_cast<const uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(hi + idx) = ph;
  uint8_t* p8 = reinterpret_cast<uint8_t*>(lo);
  const long long plane = lo - hi;
  *reinterpret_cast<uint32_t*>(p8 + idx) = e4m3x4(f0.x * sh, f0.y * sh, f1.x * sh, f1.y * sh);
  *reinterpret_cast<uint32_t*>(p8 + plane + idx) =
      e4m3x4((v.x - f0.x) * sl, (v.y - f0.y) * sl, (v.z - f1.x) * sl, (v.w - f1.y) * sl);
}

// `mode`: the kernel's copy of g_pack_mode, read ONCE per thread (`const int pmode = g_pack_mode;` at the top of the
// kernel) -- re-reading the global at every store costs a dependent global load per 16 bytes written, because the stores
// in between may alias it (measured: the operand producers 15-29 % slower, a sampling step 6 %).
__device__ __forceinline__ void store_split4(__half* hi, __half* lo, long long idx, float4 v, int mode) {
  if (mode == SDB_FMT_F8C) {
    store_split4_f8(hi, lo, idx, v, float(1 << F8_ACT_HI_EXP), float(1 << F8_ACT_LO_EXP));
    return;
  }
  uint2 ph, pl;
  split_f16x2(v.x, v.y, ph.x, pl.x);
  split_f16x2(v.z, v.w, ph.y, pl.y);
  *reinterpret_cast<uint2*>(hi + idx) = ph;
  *reinterpret_cast<uint2*>(lo + idx) = pl;
}

// weights: fmt / exponent are explicit arguments of the pack calls (wexp: e4m3(hi * 2^wexp), e4m3(lo * 2^(wexp+10)))
__device__ __forceinline__ void store_split4_w(__half* hi, __half* lo, long long idx, float4 v, int fmt, int wexp) {
  if (fmt == SDB_FMT_F8C) {
    store_split4_f8(hi, lo, idx, v, exp2f((float)wexp), exp2f((float)(wexp + 10)));
    return;
  }
  uint2 ph, pl;
  split_f16x2(v.x, v.y, ph.x, pl.x);
  split_f16x2(v.z, v.w, ph.y, pl.y);
  *reinterpret_cast<uint2*>(hi + idx) = ph;
  *reinterpret_cast<uint2*>(lo + idx) = pl;
}

#define SDB_DEFINE_PACK_MODE_SETTER(name)                                         \
  __global__ void name##_kernel(int m) { g_pack_mode = m; }                        \
  int name(int m, cudaStream_t st) {                                              \
    name##_kernel<<<1, 1, 0, st>>>(m);                                            \
    return check_cuda(cudaGetLastError(), #name);                                 \
  }
// Host mirror of the pack mode: sdb_set_pack_mode() is called in program order with the launches it governs, so the
// launchers pick the kernel INSTANCE for the current format (template <int PM>) instead of branching on the device flag --
// the default-format instances then carry none of the FP8C code (registers / instructions of instruction-bound producers).
// Kernels that are templates already (attention cores) read the device flag once per thread.
// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// The kernels of an evaluation run back to back on one stream (430 launches per UNet evaluation, replayed from a CUDA graph);
// between two of them the GPU drains, resolves the dependency and ramps the next grid up -- a few microseconds each.  Kernels
// launched through launch_k() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become resident while
// the previous kernel is still finishing (its last wave), run their prologue (barrier init, TMEM allocation, descriptor
// prefetch, index arithmetic) and then block in pdl_wait() until the previous grid has COMPLETED and its writes are visible.
// Rules: (1) every kernel launched this way calls pdl_wait() in every CTA before its first global access and before any
// exit -- completion of a grid that skipped the wait would release its own dependents early; (2) pdl_trigger() right after
// it lets the next grid start its prologue; (3) launches without the attribute (cudaMemsetAsync, torch kernels, everything not
// converted) keep full stream-order semantics, so mixing is safe.  SDB_PDL=0 switches the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

int host_pack_mode();
void set_host_pack_mode(int m);
#define SDB_LAUNCH_PM(kernel, grid, block, smem, st, ...)                                                            \
  do {                                                                                                               \
    if (host_pack_mode() == SDB_FMT_F8C) (void)launch_k(kernel<SDB_FMT_F8C>, dim3 grid, dim3 block, smem, st, __VA_ARGS__); \
    else (void)launch_k(kernel<SDB_FMT_F16X2>, dim3 grid, dim3 block, smem, st, __VA_ARGS__);                        \
  } while (0)
int set_pack_mode_elementwise(int m, cudaStream_t st);
int set_pack_mode_gemm(int m, cudaStream_t st);
int set_pack_mode_attention(int m, cudaStream_t st);
int set_pack_mode_attention_tc(int m, cudaStream_t st);

// Gradient operands: bf16 hi/lo split (x ~= hi + lo, ~16 mantissa bits, fp32 exponent range).  Loss gradients are
// routinely 1e-6 and smaller, far inside fp16's subnormal range where the fp16 split would keep only a few bits.
// Stored in the same 16-bit planes; sdb_gemm is told per operand which format the planes hold.
__device__ __forceinline__ void split_bf16(float x, __half& hi, __half& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __ushort_as_half(__bfloat16_as_ushort(h));
  lo = __ushort_as_half(__bfloat16_as_ushort(l));
}
__device__ __forceinline__ void store_split4_bf16(__half* hi, __half* lo, long long idx, float4 v) {
  __half h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(v.x, h0, l0);
  split_bf16(v.y, h1, l1);
  split_bf16(v.z, h2, l2);
  split_bf16(v.w, h3, l3);
  __half2 a = __halves2half2(h0, h1), b = __halves2half2(h2, h3);
  __half2 cc = __halves2half2(l0, l1), d = __halves2half2(l2, l3);
  uint2 ph, pl;
  ph.x = *reinterpret_cast<uint32_t*>(&a);
  ph.y = *reinterpret_cast<uint32_t*>(&b);
  pl.x = *reinterpret_cast<uint32_t*>(&cc);
  pl.y = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint2*>(hi + idx) = ph;
  *reinterpret_cast<uint2*>(lo + idx) = pl;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// x * sigmoid(x) with the fast reciprocal (MUFU.RCP, <= 2 ulp): these run once per activation element in the
// memory-bound operand producers, where the IEEE division was a third of the instruction stream (profiles/README 9)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// erf-GELU for the GEGLU epilogue of the GEMM, where 64 evaluations per thread and tile compete with the tensor pipe for a
// K = 256 main loop: Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7 ABSOLUTE, which is what 1 + erf needs; the result is then
// stored as fp16 hi + lo = 22 bits anyway) -- one MUFU.RCP, one MUFU.EX2, 8 FMA-pipe instructions, no branch (erff: two
// divergent polynomial branches, ~3x the instructions).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = x * 0.70710678118654752440f, az = fabsf(z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = fmaf(-p * t, __expf(-az * az), 1.f);      // erf(|z|)
  return 0.5f * x * (1.f + copysignf(r, z));
}

// counter-based dropout mask: keep-scale (1/(1-p)) or 0 for element `idx` of the tensor identified by `seed`
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long idx, float p, float inv_keep) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u = (float)(unsigned)(z >> 40) * (1.f / 16777216.f);   // 24 random bits -> [0,1)
  return u < p ? 0.f : inv_keep;
}

}  // namespace sdb
