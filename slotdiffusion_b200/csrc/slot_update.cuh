// Slot update ("tail" of a Slot-Attention iteration) as ONE kernel: partial-sum finalize -> GRU input/hidden
// projections -> GRU gates -> LayerNorm -> MLP (+ residual) -> LayerNorm_q -> folded slot-side projection of the NEXT
// iteration (slot_attention.py:82, :97-102 of the reference; weight fold in ops.slot_attention_fold_math).
//
// Written as a sequence of PHASES separated by CTA barriers; phase(ph, t, ...) is what thread t does in phase ph and
// touches shared scratch only through the Lay offsets below.  The same function is compiled
//   * by nvcc into slot_update_kernel (slot_update.cu): for (ph) { phase(ph, threadIdx.x, ...); __syncthreads(); }
//   * by g++ into the host emulation used by tests/test_slot_update_emulation_cpu.py: for (ph) for (t) phase(ph, t, ...)
// so the index arithmetic and the fp32 operation order (explicit fmaf, fixed summation orders) of the kernel are
// checked against the oracle on the CPU; no warp-level primitives are used on purpose.
//
// Arithmetic is plain fp32 FMA on the CUDA cores: with rows = B*S <= ~1k the ~0.8 MFLOP/row of the tail are launch-
// and latency-bound as ten separate tensor-core launches (profiles/README.md section 4: ~9 us floor per small GEMM);
// one CTA per RT rows streams the 1.6 MB of transposed weights from L2 once and reuses each element RT times.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/sdb200.h"

#ifdef __CUDACC__
#define SU_HD __host__ __device__ __forceinline__
#else
#define SU_HD inline
#endif

namespace sdb {
namespace su {

constexpr int NUM_PHASES = 17;

struct Lay {   // float offsets into the CTA's scratch; every array starts 16-byte aligned
  int u, h, gi, gh, hn, ln, y1, so, stat, red, total;
};

// SU_LAYOUT_PAD (host emulation only): unused floats after every array, so that the emulation can check that no phase
// writes outside the array it owns; the kernel is compiled with 0.
#ifndef SU_LAYOUT_PAD
#define SU_LAYOUT_PAD 0
#endif

SU_HD Lay layout(int RT, int Din, int D, int M, int nt) {
  Lay l;
  int o = 0;
  l.u = o;    o += RT * Din + SU_LAYOUT_PAD;     // normalised weighted feature means U (GRU input before the folded projection)
  l.h = o;    o += RT * D + SU_LAYOUT_PAD;       // previous slots
  l.gi = o;   o += RT * 3 * D + SU_LAYOUT_PAD;   // GRU input projection  (r | z | n)
  l.gh = o;   o += RT * 3 * D + SU_LAYOUT_PAD;   // GRU hidden projection (r | z | n)
  l.hn = o;   o += RT * D + SU_LAYOUT_PAD;       // GRU output
  l.ln = o;   o += RT * D + SU_LAYOUT_PAD;       // LayerNorm output (MLP input, later the q-projection input)
  l.y1 = o;   o += RT * M + SU_LAYOUT_PAD;       // MLP hidden
  l.so = o;   o += RT * D + SU_LAYOUT_PAD;       // new slots
  l.stat = o; o += 2 * RT + SU_LAYOUT_PAD;       // (mean, rstd) per row
  l.red = o;  o += (nt + 3) / 4 * 4 + SU_LAYOUT_PAD;   // per-thread partials of the row reductions
  l.total = o;
  return l;
}

struct F4 { float x, y, z, w; };
SU_HD F4 ld4(const float* p) {
#ifdef __CUDA_ARCH__
  const float4 v = *reinterpret_cast<const float4*>(p);
  return F4{v.x, v.y, v.z, v.w};
#else
  return F4{p[0], p[1], p[2], p[3]};
#endif
}

// out[r][j] = act(bias[j] + sum_k x[r][k] * wT[k][j] (+ resid[r][j])) for the columns j this thread owns; rows r < nr are
// stored.  x: scratch [RT][K]; wT: global, row stride ldw, consecutive j contiguous (coalesced across threads).
// A thread owns NC columns (j0, j0 + nt, ...) and walks K in chunks of KU: all NC * KU weight loads of a chunk are issued
// before the first FMA, so a thread keeps 16-24 L2 requests in flight (the kernel lives on L2 latency: 1.6 MB of weights per
// CTA), and one broadcast LDS.128 of x feeds 4 * NC FMAs.  Accumulation order per output is k ascending whatever NC / KU.
template <int RT, int NC, int KU>
SU_HD void matvec_nc(int t, int nt, int nr, const float* x, int K, const float* wT, int ldw, int N, const float* bias,
                     float* out, int64_t ldo, const float* resid, int ldres, bool relu) {
  for (int j0 = t; j0 < N; j0 += NC * nt) {
    float acc[NC][RT];
    bool ok[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      ok[c] = j0 + c * nt < N;
      const float b = (bias && ok[c]) ? bias[j0 + c * nt] : 0.f;
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[c][r] = b;
    }
    for (int k = 0; k < K; k += KU) {
      float w[NC][KU];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int u = 0; u < KU; ++u) w[c][u] = ok[c] ? wT[(int64_t)(k + u) * ldw + j0 + c * nt] : 0.f;
#pragma unroll
      for (int u = 0; u < KU; u += 4) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const F4 xv = ld4(x + r * K + k + u);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            acc[c][r] = fmaf(xv.x, w[c][u + 0], acc[c][r]);
            acc[c][r] = fmaf(xv.y, w[c][u + 1], acc[c][r]);
            acc[c][r] = fmaf(xv.z, w[c][u + 2], acc[c][r]);
            acc[c][r] = fmaf(xv.w, w[c][u + 3], acc[c][r]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (!ok[c]) continue;
      const int j = j0 + c * nt;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        if (r < nr) {
          float v = acc[c][r];
          if (resid) v += resid[r * ldres + j];
          if (relu) v = fmaxf(v, 0.f);
          out[(int64_t)r * ldo + j] = v;
        }
      }
    }
  }
}

// K % 16 == 0 (checked by the entry point)
template <int RT>
SU_HD void matvec(int t, int nt, int nr, const float* x, int K, const float* wT, int ldw, int N, const float* bias,
                  float* out, int64_t ldo, const float* resid, int ldres, bool relu) {
  if (N > 2 * nt) matvec_nc<RT, 3, 8>(t, nt, nr, x, K, wT, ldw, N, bias, out, ldo, resid, ldres, relu);
  else if (N > nt) matvec_nc<RT, 2, 8>(t, nt, nr, x, K, wT, ldw, N, bias, out, ldo, resid, ldres, relu);
  else matvec_nc<RT, 1, 16>(t, nt, nr, x, K, wT, ldw, N, bias, out, ldo, resid, ldres, relu);
}

// Row reductions for LayerNorm over src[RT][D], two-pass (mean, then centred squares), nt % RT == 0:
// thread t owns row t % RT and the columns d = t / RT, t / RT + nt / RT, ...; partials are combined in slice order.
template <int RT>
SU_HD void ln_partial(int t, int nt, const float* src, int D, const float* stat, float* red, bool centred) {
  const int r = t % RT, s0 = t / RT, ns = nt / RT;
  const float mean = centred ? stat[2 * r] : 0.f;
  float a = 0.f;
  for (int d = s0; d < D; d += ns) {
    const float v = src[r * D + d] - mean;
    a += centred ? v * v : v;
  }
  red[t] = a;
}
template <int RT>
SU_HD void ln_combine(int t, int nt, int D, float* stat, const float* red, bool centred, float eps) {
  if (t >= RT) return;
  const int ns = nt / RT;
  float a = 0.f;
  for (int s = 0; s < ns; ++s) a += red[s * RT + t];
  if (!centred) stat[2 * t] = a / (float)D;
  else stat[2 * t + 1] = 1.f / sqrtf(a / (float)D + eps);
}
template <int RT>
SU_HD void ln_apply(int t, int nt, const float* src, int D, const float* stat, const float* g, const float* b,
                    float* dst) {
  for (int i = t; i < RT * D; i += nt) {
    const int r = i / D, d = i % D;
    dst[i] = (src[i] - stat[2 * r]) * stat[2 * r + 1] * g[d] + b[d];
  }
}

// What thread t of the CTA that owns rows [tile * RT, tile * RT + RT) does in phase ph.  sm: the CTA's scratch.
template <int RT>
SU_HD void phase(int ph, int t, int nt, int64_t tile, const SdbSlotUpdate& p, float* sm) {
  const int Din = p.Din, D = p.D, M = p.M, S = p.S;
  const Lay l = layout(RT, Din, D, M, nt);
  const int64_t r0 = tile * RT;
  const int nr = (int)((p.rows - r0) < RT ? (p.rows - r0) : RT);
  float *u = sm + l.u, *h = sm + l.h, *gi = sm + l.gi, *gh = sm + l.gh, *hn = sm + l.hn, *ln = sm + l.ln;
  float *y1 = sm + l.y1, *so = sm + l.so, *stat = sm + l.stat, *red = sm + l.red;
  const bool upd = p.do_update != 0, wq = p.qa_out != nullptr;
  switch (ph) {
    case 0: {   // previous slots; U = sum_chunks part_upd / (ascale * sum_chunks part_cs)   (finalize of the attend kernel)
      for (int i = t; i < RT * D; i += nt) {
        const int r = i / D;
        const float v = r < nr ? p.slots_in[(r0 + r) * D + (i % D)] : 0.f;
        h[i] = v;
        if (!upd) so[i] = v;       // q-only call: project the given slots
      }
      if (upd) {
        for (int i = t; i < RT * Din; i += nt) {
          const int r = i / Din, c = i % Din;
          float v = 0.f;
          if (r < nr) {
            const int64_t row = r0 + r, b = row / S, s = row % S;
            float cs = 0.f;
            for (int ch = 0; ch < p.chunks; ++ch) cs += p.part_cs[(b * p.chunks + ch) * S + s];
            const float inv = 1.f / (cs * p.ascale);
            for (int ch = 0; ch < p.chunks; ++ch) v += p.part_upd[((b * p.chunks + ch) * S + s) * Din + c];
            v *= inv;
          }
          u[i] = v;
        }
      }
    } break;
    case 1:     // gi = U W_iv^T + b_iv (v projection and norm_inputs affine folded in), gh = h W_hh^T + b_hh
      if (upd) {
        matvec<RT>(t, nt, RT, u, Din, p.w_ivT, 3 * D, 3 * D, p.b_iv, gi, 3 * D, nullptr, 0, false);
        matvec<RT>(t, nt, RT, h, D, p.w_hhT, 3 * D, 3 * D, p.b_hh, gh, 3 * D, nullptr, 0, false);
      }
      break;
    case 2:     // GRUCell gates, PyTorch order (r, z, n)
      if (upd) {
        for (int i = t; i < RT * D; i += nt) {
          const int r = i / D, d = i % D;
          const float* a = gi + r * 3 * D;
          const float* b = gh + r * 3 * D;
          const float rg = 1.f / (1.f + expf(-(a[d] + b[d])));
          const float zg = 1.f / (1.f + expf(-(a[D + d] + b[D + d])));
          const float ng = tanhf(a[2 * D + d] + rg * b[2 * D + d]);
          hn[i] = (1.f - zg) * ng + zg * h[i];
        }
      }
      break;
    case 3: if (upd) ln_partial<RT>(t, nt, hn, D, stat, red, false); break;
    case 4: if (upd) ln_combine<RT>(t, nt, D, stat, red, false, p.ln_m_eps); break;
    case 5: if (upd) ln_partial<RT>(t, nt, hn, D, stat, red, true); break;
    case 6: if (upd) ln_combine<RT>(t, nt, D, stat, red, true, p.ln_m_eps); break;
    case 7: if (upd) ln_apply<RT>(t, nt, hn, D, stat, p.ln_m_g, p.ln_m_b, ln); break;
    case 8:     // MLP hidden: relu(LN(h') W_1^T + b_1)
      if (upd) matvec<RT>(t, nt, RT, ln, D, p.w1T, M, M, p.b1, y1, M, nullptr, 0, true);
      break;
    case 9:     // slots = h' + y1 W_2^T + b_2  -> scratch (for the q projection) and global
      if (upd) {
        matvec<RT>(t, nt, RT, y1, M, p.w2T, D, D, p.b2, so, D, hn, D, false);
      }
      break;
    case 10:
      if (upd) {
        for (int i = t; i < nr * D; i += nt) p.slots_out[r0 * D + i] = so[i];
      }
      if (wq) ln_partial<RT>(t, nt, so, D, stat, red, false);
      break;
    case 11: if (wq) ln_combine<RT>(t, nt, D, stat, red, false, p.ln_q_eps); break;
    case 12: if (wq) ln_partial<RT>(t, nt, so, D, stat, red, true); break;
    case 13: if (wq) ln_combine<RT>(t, nt, D, stat, red, true, p.ln_q_eps); break;
    case 14: if (wq) ln_apply<RT>(t, nt, so, D, stat, p.ln_q_g, p.ln_q_b, ln); break;
    case 15:    // qa[:, :Din] = LN_q(slots) W_qa^T; the logit-bias column Din is a split-K reduction over all threads
      if (wq) {
        matvec<RT>(t, nt, nr, ln, D, p.w_qaT, p.ldq, Din, nullptr, p.qa_out + r0 * p.ldq, p.ldq, nullptr, 0, false);
        const int r = t % RT, s0 = t / RT, ns = nt / RT;
        float a = 0.f;
        for (int k = s0; k < D; k += ns) a = fmaf(ln[r * D + k], p.w_qaT[(int64_t)k * p.ldq + Din], a);
        red[t] = a;
      }
      break;
    case 16:
      if (wq && t < nr) {
        const int ns = nt / RT;
        float a = 0.f;
        for (int s = 0; s < ns; ++s) a += red[s * RT + t];
        float* q = p.qa_out + (r0 + t) * p.ldq;
        q[Din] = a;
        for (int c = Din + 1; c < p.ldq; ++c) q[c] = 0.f;
      }
      break;
    default: break;
  }
}

}  // namespace su
}  // namespace sdb
