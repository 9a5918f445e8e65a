// Host emulation of slot_update_kernel (slotdiffusion_b200/csrc/slot_update.cuh): the kernel is a sequence of
// barrier-separated phases; here every phase is executed for every thread index of every CTA, sequentially.
// TEST INFRASTRUCTURE (built by tests/test_slot_update_emulation_cpu.py with g++); never part of the product path.
#include <cmath>
#include <limits>
#include <vector>

#include "../../slotdiffusion_b200/csrc/slot_update.cuh"

static long g_canary_violations = 0;

template <int RT>
static void run(const SdbSlotUpdate& p, int nt, int order) {
  const sdb::su::Lay l = sdb::su::layout(RT, p.Din, p.D, p.M, nt);
  const int64_t tiles = (p.rows + RT - 1) / RT;
  // shared memory is uninitialised on the device: NaN-fill so that a read-before-write shows up in the outputs
  std::vector<float> sm((size_t)l.total);
  for (int64_t tile = 0; tile < tiles; ++tile) {
    for (auto& v : sm) v = std::numeric_limits<float>::quiet_NaN();
    for (int ph = 0; ph < sdb::su::NUM_PHASES; ++ph) {
      // order 0: threads 0..nt-1; order 1: reversed (a result that depends on the order within a phase is a race)
      for (int i = 0; i < nt; ++i) sdb::su::phase<RT>(ph, order ? nt - 1 - i : i, nt, tile, p, sm.data());
    }
#if SU_LAYOUT_PAD > 0
    // the SU_LAYOUT_PAD floats after every array must still hold their NaN fill: nothing wrote past an array's end
    const int starts[] = {l.u, l.h, l.gi, l.gh, l.hn, l.ln, l.y1, l.so, l.stat, l.red, l.total};
    for (int a = 1; a <= 10; ++a)
      for (int i = starts[a] - SU_LAYOUT_PAD; i < starts[a]; ++i)
        if (!std::isnan(sm[(size_t)i])) ++g_canary_violations;
#endif
  }
}

extern "C" long su_canary_violations() { return g_canary_violations; }
extern "C" int su_layout_pad() { return SU_LAYOUT_PAD; }
extern "C" int su_scratch_floats(int RT, int Din, int D, int M, int nt) { return sdb::su::layout(RT, Din, D, M, nt).total; }

extern "C" int su_emulate(const SdbSlotUpdate* p, int RT, int nt, int order) {
  if (!p || nt % RT) return 1;
  if (RT == 4) run<4>(*p, nt, order);
  else if (RT == 8) run<8>(*p, nt, order);
  else return 1;
  return 0;
}
