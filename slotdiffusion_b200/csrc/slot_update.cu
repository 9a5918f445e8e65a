// sdb_slot_update: the whole per-iteration "tail" of Slot Attention in one launch (see slot_update.cuh for the phases
// and for why this is CUDA-core fp32 rather than ten tensor-core launches), plus the attend entry point that leaves its
// per-chunk partial sums for it (no finalize launch).
#include "common.cuh"
#include "slot_update.cuh"

namespace sdb {

template <int RT>
__global__ void __launch_bounds__(256) slot_update_kernel(const SdbSlotUpdate p) {
  extern __shared__ float4 su_smem4[];
  float* sm = reinterpret_cast<float*>(su_smem4);
  for (int ph = 0; ph < su::NUM_PHASES; ++ph) {
    su::phase<RT>(ph, (int)threadIdx.x, (int)blockDim.x, (int64_t)blockIdx.x, p, sm);
    __syncthreads();
  }
}

template <int RT>
static int launch_slot_update(const SdbSlotUpdate& p, cudaStream_t st) {
  const int nt = p.D;                                     // one thread per slot channel: 3D / 2D / D outputs = 3 / 2 / 1 rounds
  const su::Lay l = su::layout(RT, p.Din, p.D, p.M, nt);
  const size_t smem = (size_t)l.total * sizeof(float);
  auto kern = slot_update_kernel<RT>;
  static size_t attr = 0;
  if (smem > attr) {
    SDB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  kern<<<(unsigned)cdiv(p.rows, RT), nt, smem, st>>>(p);
  SDB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_slot_update_supported(int64_t S, int64_t Din, int64_t D, int64_t M) {
  if (S < 1 || Din < 16 || Din % 16 || M < 16 || M % 16) return 0;   // K loops run in chunks of 8 / 16
  if (!(D == 128 || D == 192 || D == 256)) return 0;      // block size = D, a multiple of the row tile
  const su::Lay l = su::layout(8, (int)Din, (int)D, (int)M, (int)D);
  return (size_t)l.total * sizeof(float) <= 200 * 1024 ? 1 : 0;
}

extern "C" int sdb_slot_update(const SdbSlotUpdate* pp, void* stream) {
  SDB_REQUIRE(pp, "sdb_slot_update: null argument");
  const SdbSlotUpdate& p = *pp;
  SDB_REQUIRE(p.rows > 0 && p.rows < (1ll << 31), "sdb_slot_update: bad rows=%lld", (long long)p.rows);
  SDB_REQUIRE(sdb_slot_update_supported(p.S, p.Din, p.D, p.M), "sdb_slot_update: unsupported S=%d Din=%d D=%d M=%d", p.S,
              p.Din, p.D, p.M);
  SDB_REQUIRE(p.slots_in, "sdb_slot_update: slots_in is null");
  SDB_REQUIRE(p.do_update || p.qa_out, "sdb_slot_update: nothing to do (do_update = 0 and qa_out = NULL)");
  if (p.do_update) {
    SDB_REQUIRE(p.part_upd && p.part_cs && p.chunks >= 1 && p.ascale > 0.f, "sdb_slot_update: attend partials missing");
    SDB_REQUIRE(p.rows % p.S == 0, "sdb_slot_update: rows=%lld is not a multiple of num_slots=%d", (long long)p.rows, p.S);
    SDB_REQUIRE(p.w_ivT && p.b_iv && p.w_hhT && p.b_hh && p.ln_m_g && p.ln_m_b && p.w1T && p.b1 && p.w2T && p.b2 &&
                    p.slots_out,
                "sdb_slot_update: null weight / output pointer");
  }
  if (p.qa_out) {
    SDB_REQUIRE(p.ln_q_g && p.ln_q_b && p.w_qaT, "sdb_slot_update: null q-projection weight");
    SDB_REQUIRE(p.ldq >= p.Din + 1, "sdb_slot_update: ldq=%d must exceed in_features=%d", p.ldq, p.Din);
  }
  cudaStream_t st = as_stream(stream);
  // small row counts: 4-row tiles spread the work over more SMs; otherwise 8 rows per CTA halve the L2 weight traffic
  if (p.rows <= 4ll * num_sms()) return launch_slot_update<4>(p, st);
  return launch_slot_update<8>(p, st);
}
