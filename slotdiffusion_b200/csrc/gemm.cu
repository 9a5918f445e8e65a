// sdb_gemm: C[M,N] = A[M,K] * W[N,K]^T with fused epilogue, on the 5th-gen tensor cores.
//
// Persistent, warp-specialised kernel (one CTA per SM):
//   warp 0      TMA producer   cp.async.bulk.tensor (5-D map for A: plain / implicit-im2col 3x3 / stride-2
//                              phase-split; 2-D map for W), SWIZZLE_128B, 3..8-stage mbarrier ring
//   warp 1      MMA issuer     tcgen05.mma.kind::f16, (128*CG) x BN x 16 per instruction (CG = 2: CTA pair,
//                              cta_group::2), fp32 accumulators in TMEM (2 accumulator stages so the epilogue
//                              of tile i overlaps the main loop of tile i+1)
//   warps 2..5  epilogue       tcgen05.ld -> registers -> smem staging -> 128-bit coalesced fp32 stores (or
//                              red.global.add for split-K) with bias / timestep-embedding row vector /
//                              residual / ReLU fused; all global loads of a 32x32 block are issued before use
// Operands are fp16 hi/lo planes (see sdb200.h "packed"); passes=3 issues hi*hi + lo*hi + hi*lo per k-step,
// which reproduces the fp32 product to ~2^-22 while running on the fp16 tensor pipe.
// passes=2 (SDB_FMT_F8C operands, inference): hi*hi as kind::f16 into accumulator D1, the two correction products
// l8*h8' + h8*l8' as kind::f8f6f4 (e4m3, K = 32 per instruction, twice the fp16 rate) into a second accumulator D2 in
// the same TMEM stage, C = D1 + corr_scale * D2 in the epilogue: 8 instead of 12 instructions per k-block and
// 1 + 1/2 + 1/2 = 2 pass-equivalents of tensor time (measured 0.669 of the 3-pass issue time, profiles/README 13).
// The stage holds the same bytes (fp16 tile + two half-size e4m3 tiles, 64-byte rows, SWIZZLE_64B); bn <= 128 so that
// D1 | D2 fit the 256 columns of an accumulator stage.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sdb {

// Issuer wait accounting, compiled in only with -DSDB_GEMM_TIMING=1 (SDB_GEMM_TIMING=1 python -m slotdiffusion_b200.build):
// cycles the MMA-issuing thread spends waiting for operands (full[stage]: TMA / L2 fill) and for a free accumulator
// (acc_empty: epilogue), summed over all CTAs and launches since the last reset; read with sdb_gemm_timing().
// The default build contains none of it.
#ifndef SDB_GEMM_TIMING
#define SDB_GEMM_TIMING 0
#endif
#if SDB_GEMM_TIMING
__device__ unsigned long long g_gemm_timing[4];   // wait full | wait acc_empty | issuer lifetime | tiles issued
#endif

SDB_DEFINE_PACK_MODE_SETTER(set_pack_mode_gemm)

constexpr int BM = 128;          // rows of A per CTA (UMMA M = 128 * CG)
constexpr int BK = 64;           // fp16 elements per stage row = 128 B = one swizzle-128B row
constexpr int UK = 16;           // UMMA K for 16-bit operands
constexpr int MAX_STAGES = 8;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;    // TMEM columns per accumulator stage (max UMMA N)
constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;
constexpr int EPI_WARPS = 8;     // two warps per TMEM lane quarter, interleaved over the 32-column chunks
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TILE_A_BYTES = BM * BK * 2;       // 16 KB
constexpr int SMEM_LIMIT = 227 * 1024;
// epilogue flavours (template parameter)
constexpr int EPI_F32 = 0;      // fp32 C (+ optional packed copy), bias / rowvec / residual / relu, split-K, GN partial sums
constexpr int EPI_GEGLU = 1;    // packed out[:, f] = a * gelu(g), W rows interleaved [16 a | 16 g] per 32-column chunk

struct GemmCtl {
  float epi[EPI_WARPS][32 * 32];   // per-warp staging tile, 16-B chunks XOR-swizzled by (row & 7): conflict-free both ways
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t acc_full[ACC_STAGES];
  uint64_t acc_empty[ACC_STAGES];
  uint32_t tmem_base;
};

struct GemmArgs {
  float* c;
  const float* bias;
  const float* rowvec;
  const float* residual;
  __half* out_packed;   // optional packed copy of the result (GEGLU: the only output), planes out_plane halves apart
  float* gsum;          // optional GroupNorm partial sums [M / rows_per_group][N / 4][2] (sum, sum of squares)
  long long ldc, ldv, ldr, out_plane;
  int out_act;          // activation applied to the packed copy: 0 none, 1 SiLU, 2 ReLU
  int M, N, K;
  int mode;
  int tile_rows;        // valid rows per CTA M tile (<= 128)
  int bn;               // N tile of the CTA group (multiple of 16, <= 256)
  int bnl;              // rows of W each CTA loads per stage (= bn / CG)
  int n_tiles_m;        // M tiles of the CTA GROUP (each covers CG * tile_rows rows)
  int n_tiles_n;
  int kblocks;          // K blocks of 64 per tap (plain: ceil(K/64); conv: C/64)
  int ntaps;            // 1 or 9
  int splits;           // split-K factor (partial sums combined with red.global.add)
  int stages;
  uint32_t stage_bytes;
  int passes;
  int relu;
  int vec_ok;           // 128-bit epilogue accesses allowed (alignment / N % 4)
  unsigned rows_per_group;
  int H, W;             // output H, W (conv modes)
  int ctiles;           // WGRAD modes: channel blocks per tap
  int tap_pair;         // WGRAD modes with C == 64: an M tile holds TWO taps (rows 0..63 = tap 2 tm, 64..127 = tap 2 tm + 1)
  int a_bf16, w_bf16;   // operand planes hold bf16 (gradient operands) instead of fp16
  float corr_scale;     // passes == 2: C = D1 + corr_scale * D2
  float alpha;          // accumulator scale (1 unless the caller pre-scaled an operand by a power of two)
  int gsum_cb;          // channels per GroupNorm partial-sum block (4, or 2 for the 64-channel / 32-group layers)
  int batch_rows, w_row_step, w_k_step;   // block-diagonal batching (sdb200.h)
  int acc_two;          // two accumulator stages (epilogue of tile i overlaps the main loop of tile i+1); 0: passes == 2
                        // with bn > 128, where D1 | D2 fill all 512 TMEM columns (long-K tiles: the exposed epilogue costs
                        // less than the shared-memory traffic of narrower tiles)
  int d2_off;           // passes == 2: TMEM column offset of the correction accumulator inside a stage
  int debug;            // SDB_GEMM_DEBUG (launch-floor experiments): 1 = exit at entry, 2 = prologue + teardown only
};

// Warp-specialised persistent GEMM.  CG = 1: one CTA per tile (UMMA 128 x bn).  CG = 2: a CTA pair (cluster of 2)
// per tile, tcgen05 cta_group::2 (UMMA 256 x bn): each CTA stages its own 128 rows of A and HALF of the W tile,
// which halves the shared-memory fill and read traffic per flop -- with three MMA passes per product the single-CTA
// form is shared-memory-bandwidth bound well below the tensor peak.
// VAR: compile-time feature set of an instance, so that the instance the UNet forward runs carries none of the code of the
// rarer forms (measured on one B200, one evaluation at B=256: the runtime-flag version of these features cost the 144
// pair-GEMM launches 15.4 -> 16.7 ms).  Bit 0 (VAR_F8C): FP8-corrected operands / packed outputs in the stream's pack mode
// (passes == 2, D2 accumulator, one-stage accumulators); bit 1 (VAR_EXT): alpha != 1, 2-channel GroupNorm sums, block-
// diagonal batching, WGRAD tap pairing.  The host picks the smallest instance that covers the call.
constexpr int VAR_F8C = 1, VAR_EXT = 2;

template <int CG, int EPI, int VAR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
            const __grid_constant__ CUtensorMap map_a_l8, const __grid_constant__ CUtensorMap map_b_l8,
            const GemmArgs g) {
  constexpr bool kF8 = (VAR & VAR_F8C) != 0, kExt = (VAR & VAR_EXT) != 0;
  const int acc_two = kF8 ? g.acc_two : 1;
  const int gsum_cb = kExt ? g.gsum_cb : 4;
  const int batch_rows = kExt ? g.batch_rows : 0;
  const int tap_pair = kExt ? g.tap_pair : 0;
  extern __shared__ uint8_t smem_raw[];
  if (g.debug == 1) { pdl_wait(); return; }
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmCtl& ctl = *reinterpret_cast<GemmCtl*>(ring + (size_t)g.stages * g.stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int group = blockIdx.x / CG, ngroups = gridDim.x / CG;
  const int num_items = g.n_tiles_m * g.n_tiles_n * g.splits;
  const int ksteps = g.kblocks * g.ntaps;   // k-blocks (stages) of a whole tile
  const bool three = g.passes >= 2;    // a second operand plane is staged (fp16 lo, or the two e4m3 half-planes)
  const bool f8c = kF8 && g.passes == 2;      // map_*_lo then address the h8 half-plane, map_*_l8 the l8 half-plane
  const uint32_t b_tile_bytes = uint32_t(g.bnl) * BK * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (three) {
      tma_prefetch_desc(&map_a_lo);
      tma_prefetch_desc(&map_b_lo);
    }
    if (f8c) {
      tma_prefetch_desc(&map_a_l8);
      tma_prefetch_desc(&map_b_l8);
    }
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&ctl.full[s], 1);
      mbar_init(&ctl.empty[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&ctl.acc_full[a], 1);
      mbar_init(&ctl.acc_empty[a], EPI_WARPS * CG);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_pair(&ctl.tmem_base, TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(&ctl.tmem_base, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;
  // PDL (common.cuh): everything above -- descriptor prefetch, mbarrier init, TMEM allocation, the cluster barrier -- may
  // overlap the tail of the previous kernel; nothing below may: operands, bias / residual, the pack-mode flag and C belong to it
  pdl_wait();
  pdl_trigger();
  const int pmode = kF8 ? g_pack_mode : SDB_FMT_F16X2;   // operand format of the consumer GEMM: read once (common.cuh)

  if (g.debug == 2) {
    // skip the work
  } else if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      const uint32_t tx = (three ? 2u : 1u) * (uint32_t(g.tile_rows) * BK * 2 + b_tile_bytes) * CG;
      int stage = 0;
      uint32_t phase = 0;
      for (int item = group; item < num_items; item += ngroups) {
        const int tile = item / g.splits, split = item - tile * g.splits;
        const int tm = (tile / g.n_tiles_n) * CG + (int)rank, tn = tile % g.n_tiles_n;
        const int ks_begin = (int)((long long)ksteps * split / g.splits);
        const int ks_end = (int)((long long)ksteps * (split + 1) / g.splits);
        int x0 = 0, y0 = 0, b0 = 0;
        const bool wgrad = (g.mode == SDB_A_WGRAD || g.mode == SDB_A_WGRAD_S2);
        int wtap = 0, wc0 = 0;
        if (g.mode == SDB_A_PLAIN) {
          x0 = tm * BM;   // row index lives in dim 1
        } else if (wgrad) {
          // output row block tm = (tap, channel block): A rows are CHANNELS of the transposed activation, K = pixels
          wtap = tm / g.ctiles;
          wc0 = (tm - wtap * g.ctiles) * g.tile_rows;
        } else {
          const int rows_per_img = g.H * g.W;
          const long long m0 = (long long)tm * g.tile_rows;
          b0 = int(m0 / rows_per_img);
          y0 = int((m0 % rows_per_img) / g.W);
        }
        const int bidx = batch_rows ? (int)(((long long)tm * g.tile_rows) / batch_rows) : 0;
        const int nrow = tn * g.bn + (int)rank * g.bnl + bidx * g.w_row_step;
        const int wk0 = bidx * g.w_k_step;
        for (int ks = ks_begin; ks < ks_end; ++ks) {
          mbar_wait(&ctl.empty[stage], phase ^ 1);
          uint8_t* base = ring + (size_t)stage * g.stage_bytes;
          if (rank == 0) mbar_arrive_expect_tx(&ctl.full[stage], tx);
          int tap = ks / g.kblocks;
          int c0 = (ks - tap * g.kblocks) * BK;
          int cx = x0, cy = y0, cp = 0;
          int c3 = cp, c4 = b0;
          if (wgrad) {
            // k-block ks = 64 consecutive output pixels (64/W rows of one image, or 64/(H*W) whole images).  A is the
            // NHWC activation itself: box {64 channels, W, rows, 1, images} shifted by the tap in the OUTER dims only.
            const int m0 = ks * BK;
            const int hw = g.H * g.W;
            const int bimg = m0 / hw;
            const int yrow = (m0 - bimg * hw) / g.W;
            const int xoff = (m0 - bimg * hw) - yrow * g.W;   // non-zero only for W > 64 (a k-block is a 64-pixel row segment)
            tap = wtap;
            c0 = wc0;
            if (g.mode == SDB_A_WGRAD) {
              cx = tap % 3 - 1 + xoff;
              cy = yrow + tap / 3 - 1;
              c3 = 0;
            } else {
              const int ky = tap / 3, kx = tap % 3;
              cx = ((kx == 0) ? -1 : 0) + xoff;
              cy = yrow + ((ky == 0) ? -1 : 0);
              c3 = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
            }
            c4 = bimg;
          } else if (g.mode == SDB_A_CONV3) {
            cx = tap % 3 - 1;
            cy = y0 + tap / 3 - 1;
          } else if (g.mode == SDB_A_CONV3S2) {
            // input pixel (2y+ky-1, 2x+kx-1): ky=0 -> odd phase, row y-1; ky=1 -> even phase, row y; ky=2 -> odd, row y
            const int ky = tap / 3, kx = tap % 3;
            cp = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
            cx = (kx == 0) ? -1 : 0;
            cy = y0 + ((ky == 0) ? -1 : 0);
          } else if (g.mode == SDB_A_CONV3S2A) {
            // input pixel (2y+ky, 2x+kx), zero row / column past the bottom / right edge: ky=0 -> even phase, row y;
            // ky=1 -> odd phase, row y; ky=2 -> even phase, row y+1 (out of bounds at y = H-1: TMA zero fill = the pad)
            const int ky = tap / 3, kx = tap % 3;
            cp = ((ky == 1) ? 2 : 0) + ((kx == 1) ? 1 : 0);
            cx = (kx == 2) ? 1 : 0;
            cy = y0 + ((ky == 2) ? 1 : 0);
          }
          if (!wgrad) { c3 = cp; c4 = b0; }
          uint8_t* a_hi = base, * a_lo = base + TILE_A_BYTES;
          uint8_t* b_hi = base + 2 * TILE_A_BYTES, * b_lo = b_hi + b_tile_bytes;
          if (wgrad) {
            // MN-major operands: 64-wide (128-byte) column blocks of [64 pixels] x [64 channels], 8 KB apart
            const int na = g.tile_rows / 64, nb = g.bnl / 64;
            for (int pl = 0; pl < (three ? 2 : 1); ++pl) {
              const CUtensorMap* ma = pl ? &map_a_lo : &map_a_hi;
              const CUtensorMap* mb = pl ? &map_b_lo : &map_b_hi;
              uint8_t* ap = pl ? a_lo : a_hi;
              uint8_t* bp = pl ? b_lo : b_hi;
              for (int j = 0; j < na; ++j) {
                int jc0 = c0 + j * 64, jcx = cx, jcy = cy, jc3 = c3;
                if (tap_pair) {      // block j = tap 2 tm + j of the 64 channels (output row = tap * 64 + ci stays contiguous);
                  const int tj = min(2 * tm + j, 8);     // the odd ninth tap repeats: its rows lie beyond M and are dropped
                  const int m0 = ks * BK, hw = g.H * g.W, bimg = m0 / hw;
                  const int yrow = (m0 - bimg * hw) / g.W, xoff = (m0 - bimg * hw) - yrow * g.W;
                  jc0 = 0;
                  if (g.mode == SDB_A_WGRAD) {
                    jcx = tj % 3 - 1 + xoff; jcy = yrow + tj / 3 - 1; jc3 = 0;
                  } else {
                    const int ky = tj / 3, kx = tj % 3;
                    jcx = ((kx == 0) ? -1 : 0) + xoff; jcy = yrow + ((ky == 0) ? -1 : 0);
                    jc3 = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
                  }
                }
                if (CG == 2) tma_load_5d_pair(ap + j * 8192, ma, &ctl.full[stage], jc0, jcx, jcy, jc3, c4);
                else tma_load_5d(ap + j * 8192, ma, &ctl.full[stage], jc0, jcx, jcy, jc3, c4);
              }
              for (int j = 0; j < nb; ++j) {
                if (CG == 2) tma_load_2d_pair(bp + j * 8192, mb, &ctl.full[stage], nrow + j * 64, ks * BK);
                else tma_load_2d(bp + j * 8192, mb, &ctl.full[stage], nrow + j * 64, ks * BK);
              }
            }
          } else if (CG == 2) {
            tma_load_5d_pair(a_hi, &map_a_hi, &ctl.full[stage], c0, cx, cy, c3, c4);
            tma_load_2d_pair(b_hi, &map_b_hi, &ctl.full[stage], wk0 + ks * BK, nrow);
            if (three) {
              tma_load_5d_pair(a_lo, &map_a_lo, &ctl.full[stage], c0, cx, cy, c3, c4);
              tma_load_2d_pair(b_lo, &map_b_lo, &ctl.full[stage], wk0 + ks * BK, nrow);
            }
            if (f8c) {   // second e4m3 half-plane: half a 16-bit tile further on
              tma_load_5d_pair(a_lo + TILE_A_BYTES / 2, &map_a_l8, &ctl.full[stage], c0, cx, cy, c3, c4);
              tma_load_2d_pair(b_lo + b_tile_bytes / 2, &map_b_l8, &ctl.full[stage], wk0 + ks * BK, nrow);
            }
          } else {
            tma_load_5d(a_hi, &map_a_hi, &ctl.full[stage], c0, cx, cy, c3, c4);
            tma_load_2d(b_hi, &map_b_hi, &ctl.full[stage], wk0 + ks * BK, nrow);
            if (three) {
              tma_load_5d(a_lo, &map_a_lo, &ctl.full[stage], c0, cx, cy, c3, c4);
              tma_load_2d(b_lo, &map_b_lo, &ctl.full[stage], wk0 + ks * BK, nrow);
            }
            if (f8c) {
              tma_load_5d(a_lo + TILE_A_BYTES / 2, &map_a_l8, &ctl.full[stage], c0, cx, cy, c3, c4);
              tma_load_2d(b_lo + b_tile_bytes / 2, &map_b_l8, &ctl.full[stage], wk0 + ks * BK, nrow);
            }
          }
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread of the leader CTA) =====================
    if (lane == 0 && rank == 0) {
      const bool mn_major = (g.mode == SDB_A_WGRAD || g.mode == SDB_A_WGRAD_S2);
      // wgrad: both operands are MN-major in shared memory (pixels = K run along the 128-byte rows' ROW index)
      const uint32_t idesc = umma_idesc_f16(BM * CG, g.bn) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
                             (g.a_bf16 ? (1u << 7) : 0u) | (g.w_bf16 ? (1u << 10) : 0u);   // a_format / b_format = BF16
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
#if SDB_GEMM_TIMING
      long long tw_full = 0, tw_acc = 0;
      const long long t_begin = clock64();
#endif
      for (int item = group; item < num_items; item += ngroups, ++it) {
        const int tile = item / g.splits, split = item - tile * g.splits;
        const int ks_begin = (int)((long long)ksteps * split / g.splits);
        const int ks_end = (int)((long long)ksteps * (split + 1) / g.splits);
        const int as = acc_two ? (it & 1) : 0;
        const uint32_t aphase = acc_two ? ((it >> 1) & 1) : (it & 1);
#if SDB_GEMM_TIMING
        const long long ta = clock64();
#endif
        mbar_wait(&ctl.acc_empty[as], aphase ^ 1);
#if SDB_GEMM_TIMING
        tw_acc += clock64() - ta;
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int ks = ks_begin; ks < ks_end; ++ks) {
#if SDB_GEMM_TIMING
          const long long tf = clock64();
#endif
          mbar_wait(&ctl.full[stage], phase);
#if SDB_GEMM_TIMING
          tw_full += clock64() - tf;
#endif
          tc_fence_after();
          const uint32_t a_hi = smem_u32(ring + (size_t)stage * g.stage_bytes);
          const uint32_t a_lo = a_hi + TILE_A_BYTES;
          const uint32_t b_hi = a_hi + 2 * TILE_A_BYTES;
          const uint32_t b_lo = b_hi + b_tile_bytes;
          const uint64_t da_hi = mn_major ? umma_desc_mnmajor_sw128(a_hi) : umma_desc_kmajor_sw128(a_hi);
          const uint64_t da_lo = mn_major ? umma_desc_mnmajor_sw128(a_lo) : umma_desc_kmajor_sw128(a_lo);
          const uint64_t db_hi = mn_major ? umma_desc_mnmajor_sw128(b_hi) : umma_desc_kmajor_sw128(b_hi);
          const uint64_t db_lo = mn_major ? umma_desc_mnmajor_sw128(b_lo) : umma_desc_kmajor_sw128(b_lo);
          if (f8c) {
            // D1 += A_h W_h (4 x kind::f16, K = 16);  D2 += A_l8 W_h8 + A_h8 W_l8 (2 x 2 x kind::f8f6f4, K = 32)
            const uint32_t d2 = d_tmem + g.d2_off;
            const uint32_t first = (ks != ks_begin) ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t adv = uint64_t((k * UK * 2) >> 4);
              if (CG == 2) umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, k ? 1u : first);
              else umma_f16(d_tmem, da_hi + adv, db_hi + adv, idesc, k ? 1u : first);
            }
            const uint64_t da_h8 = umma_desc_kmajor_sw64(a_lo), da_l8 = umma_desc_kmajor_sw64(a_lo + TILE_A_BYTES / 2);
            const uint64_t db_h8 = umma_desc_kmajor_sw64(b_lo), db_l8 = umma_desc_kmajor_sw64(b_lo + b_tile_bytes / 2);
#pragma unroll
            for (int k = 0; k < BK / 32; ++k) {
              const uint64_t adv = uint64_t((k * 32) >> 4);     // 32 e4m3 = 32 B inside the 64-B swizzle row
              if (CG == 2) {
                umma_f8_pair(d2, da_l8 + adv, db_h8 + adv, idesc, k ? 1u : first);
                umma_f8_pair(d2, da_h8 + adv, db_l8 + adv, idesc, 1u);
              } else {
                umma_f8(d2, da_l8 + adv, db_h8 + adv, idesc, k ? 1u : first);
                umma_f8(d2, da_h8 + adv, db_l8 + adv, idesc, 1u);
              }
            }
          } else
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // K-major: 32 B per k-step inside the 128-B swizzle row; MN-major: 16 K rows = two 1024-B atoms per k-step
            const uint64_t adv = mn_major ? uint64_t((k * UK * 128) >> 4) : uint64_t((k * UK * 2) >> 4);
            const uint32_t acc = (ks != ks_begin || k != 0) ? 1u : 0u;
            if (CG == 2) {
              umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, acc);
              if (g.passes == 3) {
                umma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1);
                umma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
              }
            } else {
              umma_f16(d_tmem, da_hi + adv, db_hi + adv, idesc, acc);
              if (g.passes == 3) {
                umma_f16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1);
                umma_f16(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
              }
            }
          }
          // smem slot reusable (in both CTAs) once these MMAs retire
          if (CG == 2) umma_commit_pair(&ctl.empty[stage]); else umma_commit(&ctl.empty[stage]);
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair(&ctl.acc_full[as]); else umma_commit(&ctl.acc_full[as]);   // accumulator complete
      }
#if SDB_GEMM_TIMING
      atomicAdd(&g_gemm_timing[0], (unsigned long long)tw_full);
      atomicAdd(&g_gemm_timing[1], (unsigned long long)tw_acc);
      atomicAdd(&g_gemm_timing[2], (unsigned long long)(clock64() - t_begin));
      atomicAdd(&g_gemm_timing[3], (unsigned long long)it);
#endif
    }
  } else {
    // ===================== epilogue warps (both CTAs: each drains its own 128 TMEM lanes) =====================
    const int ew = warp - 2;                // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int par = ew >> 2;                // parity of the 32-column chunks this warp handles
    float* stg = ctl.epi[ew];
    const int c4i = lane & 7, rsub = lane >> 3;
    const unsigned n4 = (unsigned)g.N >> 2;
    int it = 0;
    for (int item = group; item < num_items; item += ngroups, ++it) {
      const int tile = item / g.splits, split = item - tile * g.splits;
      const int tm = (tile / g.n_tiles_n) * CG + (int)rank, tn = tile % g.n_tiles_n;
      const bool first = split == 0;       // the split that also adds bias / rowvec / residual
      const int as = acc_two ? (it & 1) : 0;
      const uint32_t aphase = acc_two ? ((it >> 1) & 1) : (it & 1);
      const long long row0l = (long long)tm * g.tile_rows + q * 32;   // first output row of this warp
      const int rows_valid = (int)min((long long)min(g.tile_rows - q * 32, 32), (long long)g.M - row0l);
      const int row0 = (int)min(row0l, (long long)g.M);
      const int ncols = min(g.bn, g.N - tn * g.bn);
      const int nch = (ncols + 31) >> 5;

      // global loads feeding chunk j (residual, timestep row vector): issued one chunk ahead of their use
      auto prefetch = [&](int j, float4 (&dst)[8]) {
        const int n = tn * g.bn + j * 32 + c4i * 4;
        const bool col_ok = (j * 32 + c4i * 4 < ncols) && first && EPI == EPI_F32 && g.vec_ok;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rsub;
          dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && rr < rows_valid) {
            const unsigned m = (unsigned)(row0 + rr);
            if (g.residual) dst[i] = *reinterpret_cast<const float4*>(g.residual + (long long)m * g.ldr + n);
            if (g.rowvec) {
              const float4 t =
                  *reinterpret_cast<const float4*>(g.rowvec + (long long)(m / g.rows_per_group) * g.ldv + n);
              dst[i].x += t.x; dst[i].y += t.y; dst[i].z += t.z; dst[i].w += t.w;
            }
          }
        }
      };
      float4 pre[8];
      if (par < nch) prefetch(par, pre);
      mbar_wait(&ctl.acc_full[as], aphase);
      tc_fence_after();

      for (int j = par; j < nch; j += 2) {
        const int cb = j * 32;
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * ACC_COLS + cb);
        if (g.bn - cb >= 32) {
          tmem_ld_32x32(taddr, r);
        } else {
          uint32_t r16[16];
          tmem_ld_32x16(taddr, r16);
#pragma unroll
          for (int k = 0; k < 16; ++k) { r[k] = r16[k]; r[k + 16] = 0; }
        }
        float4 cur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = pre[i];
        if (j + 2 < nch) prefetch(j + 2, pre);
        if (f8c) {      // correction accumulator: same lanes, ACC_COLS / 2 columns further on
          uint32_t r2[32];
          if (g.bn - cb >= 32) {
            tmem_ld_32x32(taddr + g.d2_off, r2);
          } else {
            uint32_t r16[16];
            tmem_ld_32x16(taddr + g.d2_off, r16);
#pragma unroll
            for (int k = 0; k < 16; ++k) { r2[k] = r16[k]; r2[k + 16] = 0; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(fmaf(g.corr_scale, __uint_as_float(r2[k]), __uint_as_float(r[k])));
        }
        tmem_ld_wait();
        if (kExt && g.alpha != 1.f) {
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * g.alpha);
        }
        if (rows_valid <= 0) continue;       // warp-uniform
        // thread `lane` holds row (q*32+lane), columns cb..cb+31 -> stage (swizzled) so that 8 lanes cover one
        // 128-B row segment on the way out
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(stg + lane * 32 + ((k ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                          __uint_as_float(r[4 * k + 3]));
        __syncwarp();
        if (EPI == EPI_GEGLU) {
          // chunk = [16 a columns | 16 g columns]; lane <-> (row = i*8 + lane/4, 4 output columns cc)
          const int cc = lane & 3, rq = lane >> 2;
          const int na = tn * g.bn + cb + cc * 4;                 // column of a in the interleaved N space
          const int fo = (tn * g.bn + cb) / 2 + cc * 4;           // output column
          const int F = g.N >> 1;
          float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
          if (g.bias) {
            ba = *reinterpret_cast<const float4*>(g.bias + na);
            bg = *reinterpret_cast<const float4*>(g.bias + na + 16);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + rq;
            if (rr < rows_valid && cb + cc * 4 < ncols) {
              const float4 a = *reinterpret_cast<const float4*>(stg + rr * 32 + ((cc ^ (rr & 7)) << 2));
              const float4 gg = *reinterpret_cast<const float4*>(stg + rr * 32 + (((cc + 4) ^ (rr & 7)) << 2));
              float4 o;
              o.x = (a.x + ba.x) * gelu_erf_fast(gg.x + bg.x);
              o.y = (a.y + ba.y) * gelu_erf_fast(gg.y + bg.y);
              o.z = (a.z + ba.z) * gelu_erf_fast(gg.z + bg.z);
              o.w = (a.w + ba.w) * gelu_erf_fast(gg.w + bg.w);
              store_split4(g.out_packed, g.out_packed + g.out_plane, (long long)(row0 + rr) * F + fo, o, pmode);
            }
          }
        } else if (g.vec_ok) {
          const int n = tn * g.bn + cb + c4i * 4;
          const bool col_ok = (cb + c4i * 4 < ncols);
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && first && g.bias) bias4 = *reinterpret_cast<const float4*>(g.bias + n);
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;   // GroupNorm partial sums of rows 0..15 / 16..31
          float s0b = 0.f, q0b = 0.f, s1b = 0.f, q1b = 0.f;   // gsum_cb == 2: second channel pair (.z, .w) of the float4
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            if (col_ok && rr < rows_valid) {
              float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + ((c4i ^ (rr & 7)) << 2));
              v.x += cur[i].x + bias4.x; v.y += cur[i].y + bias4.y; v.z += cur[i].z + bias4.z; v.w += cur[i].w + bias4.w;
              if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              const long long m = row0 + rr;
              if (g.c) {
                float* dst = g.c + m * g.ldc + n;
                if (g.splits > 1) red_add_v4(dst, v);
                else *reinterpret_cast<float4*>(dst) = v;
              }
              if (g.gsum) {
                if (gsum_cb == 2) {
                  const float sa = v.x + v.y, qa = v.x * v.x + v.y * v.y, sb = v.z + v.w, qb = v.z * v.z + v.w * v.w;
                  if (i < 4) { s0 += sa; q0 += qa; s0b += sb; q0b += qb; } else { s1 += sa; q1 += qa; s1b += sb; q1b += qb; }
                } else {
                  const float s = (v.x + v.y) + (v.z + v.w);
                  const float qq = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                  if (i < 4) { s0 += s; q0 += qq; } else { s1 += s; q1 += qq; }
                }
              }
              if (g.out_packed) {
                if (g.out_act == 1) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
                else if (g.out_act == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                store_split4(g.out_packed, g.out_packed + g.out_plane, m * g.N + n, v, pmode);
              }
            }
          }
          if (g.gsum) {
            // rows rr = i*4 + rsub: sum over the 4 row sub-lanes (lanes l, l+8, l+16, l+24 share the columns)
            s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  q0 += __shfl_xor_sync(0xffffffffu, q0, 8);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 8);  q1 += __shfl_xor_sync(0xffffffffu, q1, 8);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
            if (gsum_cb == 2) {
              s0b += __shfl_xor_sync(0xffffffffu, s0b, 8);  q0b += __shfl_xor_sync(0xffffffffu, q0b, 8);
              s1b += __shfl_xor_sync(0xffffffffu, s1b, 8);  q1b += __shfl_xor_sync(0xffffffffu, q1b, 8);
              s0b += __shfl_xor_sync(0xffffffffu, s0b, 16); q0b += __shfl_xor_sync(0xffffffffu, q0b, 16);
              s1b += __shfl_xor_sync(0xffffffffu, s1b, 16); q1b += __shfl_xor_sync(0xffffffffu, q1b, 16);
            }
            if (rsub == 0 && col_ok) {
              const unsigned b0 = (unsigned)row0 / g.rows_per_group;
              const unsigned b1 = (unsigned)(row0 + 16) / g.rows_per_group;
              // block index of this thread's first channel: n / cb; blocks per row: N / cb
              const unsigned nb = gsum_cb == 2 ? ((unsigned)g.N >> 1) : n4;
              const unsigned ib = gsum_cb == 2 ? ((unsigned)n >> 1) : ((unsigned)n >> 2);
              if (b1 == b0 || rows_valid <= 16) {
                float* p = g.gsum + ((long long)b0 * nb + ib) * 2;
                atomicAdd(p, s0 + s1);
                atomicAdd(p + 1, q0 + q1);
                if (gsum_cb == 2) { atomicAdd(p + 2, s0b + s1b); atomicAdd(p + 3, q0b + q1b); }
              } else {
                float* p = g.gsum + ((long long)b0 * nb + ib) * 2;
                atomicAdd(p, s0);
                atomicAdd(p + 1, q0);
                if (gsum_cb == 2) { atomicAdd(p + 2, s0b); atomicAdd(p + 3, q0b); }
                p = g.gsum + ((long long)b1 * nb + ib) * 2;
                atomicAdd(p, s1);
                atomicAdd(p + 1, q1);
                if (gsum_cb == 2) { atomicAdd(p + 2, s1b); atomicAdd(p + 3, q1b); }
              }
            }
          }
        } else {
          // scalar fallback (unaligned views / N % 4 != 0): lane <-> column, rows in sequence
          const int nn = tn * g.bn + cb + lane;
          const bool col_ok = (cb + lane < ncols);
          const float bias = (col_ok && first && g.bias) ? g.bias[nn] : 0.f;
          for (int rr = 0; rr < rows_valid; ++rr) {
            if (!col_ok) break;
            const unsigned m = (unsigned)(row0 + rr);
            float v = stg[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))] + bias;
            if (first && g.rowvec) v += g.rowvec[(long long)(m / g.rows_per_group) * g.ldv + nn];
            if (first && g.residual) v += g.residual[(long long)m * g.ldr + nn];
            if (g.relu) v = fmaxf(v, 0.f);
            float* dst = g.c + (long long)m * g.ldc + nn;
            if (g.splits > 1) atomicAdd(dst, v);
            else *dst = v;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&ctl.acc_empty[as]); else mbar_arrive(&ctl.acc_empty[as]);
      }
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 tensor map, up to rank 5, SWIZZLE_128B, zero OOB fill. dims/box innermost first; strides in bytes for dims 1..
static int make_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool e4m3 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return SDB_ERR_CUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  // e4m3 planes: one byte per element, 64-element (64-byte) box rows, SWIZZLE_64B
  CUresult r = enc(m, e4m3 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                   const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   e4m3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]", (int)r,
              rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
              (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
              rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return SDB_ERR_CUDA;
  }
  return 0;
}

// N tile: multiple of 16 (32 for CTA pairs so that each CTA's half is a whole number of 8-row swizzle atoms... 16 rows),
// <= max_bn, minimising padded columns first and the tile count second.
static int pick_bn(int N, int max_bn, int step) {
  int best = step;
  long long best_cost = -1;
  for (int bn = max_bn; bn >= step; bn -= step) {
    long long tiles = cdiv(N, bn);
    long long cost = tiles * bn * 1000 + tiles * 40;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

template <int CG, int EPI, int VAR>
static int launch_gemm(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mb_hi,
                       const CUtensorMap& mb_lo, const CUtensorMap& ma_l8, const CUtensorMap& mb_l8, const GemmArgs& g,
                       int groups, size_t smem, cudaStream_t st) {
  static size_t attr = 0;
  if (smem > attr) {
    SDB_CHECK(cudaFuncSetAttribute(gemm_kernel<CG, EPI, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // PDL, see common.cuh
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SDB_CHECK(cudaLaunchKernelEx(&cfg, gemm_kernel<CG, EPI, VAR>, ma_hi, ma_lo, mb_hi, mb_lo, ma_l8, mb_l8, g));
  SDB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_gemm(const SdbGemm* p, void* stream) {
  SDB_REQUIRE(p && p->a && p->w && (p->c || p->out_packed), "sdb_gemm: null operand");
  SDB_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "sdb_gemm: empty problem M=%d N=%d K=%d", p->M, p->N, p->K);
  SDB_REQUIRE(p->passes == 1 || p->passes == 2 || p->passes == 3, "sdb_gemm: passes must be 1, 2 or 3");
  const bool f8c = p->passes == 2;
  if (f8c) {
    SDB_REQUIRE(p->mode == SDB_A_PLAIN || p->mode == SDB_A_CONV3 || p->mode == SDB_A_CONV3S2 || p->mode == SDB_A_CONV3S2A,
                "sdb_gemm: passes = 2 (SDB_FMT_F8C operands) is a forward-only path (mode %d)", p->mode);
    SDB_REQUIRE(p->w_plane_stride == 0, "sdb_gemm: passes = 2 takes whole packed W tensors");
    SDB_REQUIRE(!p->a_bf16 && !p->w_bf16, "sdb_gemm: passes = 2 takes fp16 + e4m3 operands, not bf16");
    SDB_REQUIRE(p->K % 16 == 0, "sdb_gemm: passes = 2 needs K %% 16 == 0 (16-byte rows of the e4m3 planes), K=%d", p->K);
    SDB_REQUIRE(p->corr_scale > 0.f, "sdb_gemm: passes = 2 needs corr_scale = 2^-(12 + wexp)");
  }
  SDB_REQUIRE((p->a_bf16 != 0) == (p->w_bf16 != 0), "sdb_gemm: A and W must use the same 16-bit format (fp16 or bf16)");
  SDB_REQUIRE(p->K % 8 == 0 || p->mode == SDB_A_WGRAD || p->mode == SDB_A_WGRAD_S2,
              "sdb_gemm: K=%d must be a multiple of 8 (16-byte TMA rows)", p->K);
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(p->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
              "sdb_gemm: operands must be 16-byte aligned");
  SDB_REQUIRE(!p->rowvec || p->rows_per_group > 0, "sdb_gemm: rowvec needs rows_per_group");
  GemmArgs g{};
  g.c = p->c; g.bias = p->bias; g.rowvec = p->rowvec; g.residual = p->residual;
  g.out_packed = reinterpret_cast<__half*>(p->out_packed); g.gsum = p->gsum;
  g.out_plane = p->out_plane_stride; g.out_act = p->out_act;
  g.a_bf16 = p->a_bf16 != 0; g.w_bf16 = p->w_bf16 != 0;
  g.corr_scale = p->corr_scale;
  g.alpha = p->alpha == 0.f ? 1.f : p->alpha;
  g.gsum_cb = p->gsum_cb == 2 ? 2 : 4;
  SDB_REQUIRE(p->gsum_cb == 0 || p->gsum_cb == 2 || p->gsum_cb == 4, "sdb_gemm: gsum_cb must be 0, 2 or 4");
  g.batch_rows = p->batch_rows; g.w_row_step = p->w_row_step; g.w_k_step = p->w_k_step;
  if (p->batch_rows) {
    SDB_REQUIRE(p->mode == SDB_A_PLAIN && p->batch_rows % 256 == 0 && p->M % p->batch_rows == 0 && p->w_rows > 0 && p->w_cols > 0 &&
                    p->w_cols % 8 == 0 && p->passes != 2 && p->w_plane_stride == 0,
                "sdb_gemm: batched form needs plain mode, batch_rows %% 256 == 0, M %% batch_rows == 0, w_rows / w_cols");
    const long long nb = p->M / p->batch_rows;
    SDB_REQUIRE((nb - 1) * p->w_row_step + p->N <= p->w_rows && (nb - 1) * p->w_k_step + p->K <= p->w_cols,
                "sdb_gemm: batched W extents exceed the [w_rows, w_cols] tensor");
  }
  g.debug = env_int("SDB_GEMM_DEBUG", 0);
  g.ldc = p->ldc; g.ldv = p->ldv; g.ldr = p->ldr;
  g.M = p->M; g.N = p->N; g.K = p->K; g.mode = p->mode; g.passes = p->passes; g.relu = p->relu;
  g.rows_per_group = p->rows_per_group > 0 ? (unsigned)p->rows_per_group : 1u;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  g.vec_ok = (p->N % 4 == 0) && (!p->c || (p->ldc % 4 == 0 && al16(p->c))) && (!p->bias || al16(p->bias)) &&
             (!p->residual || (al16(p->residual) && p->ldr % 4 == 0)) &&
             (!p->rowvec || (al16(p->rowvec) && p->ldv % 4 == 0));

  const bool geglu = p->geglu != 0;
  if (geglu) {
    SDB_REQUIRE(p->out_packed && !p->c && !p->residual && !p->rowvec && !p->relu && !p->gsum,
                "sdb_gemm: GEGLU epilogue writes only the packed output (no c / residual / rowvec / relu / gsum)");
    SDB_REQUIRE(p->N % 32 == 0 && (!p->bias || al16(p->bias)), "sdb_gemm: GEGLU needs N %% 32 == 0");
  }
  if (p->out_packed) {
    SDB_REQUIRE(al16(p->out_packed) && p->out_plane_stride % 8 == 0 && p->N % 4 == 0 && g.vec_ok,
                "sdb_gemm: packed output needs 16-byte aligned planes and vectorisable operands");
    SDB_REQUIRE(p->out_act >= 0 && p->out_act <= 2, "sdb_gemm: bad out_act %d", p->out_act);
  }
  if (p->gsum) {
    SDB_REQUIRE(g.vec_ok && p->rows_per_group > 0 && p->rows_per_group % 16 == 0,
                "sdb_gemm: gsum needs vectorisable operands and rows_per_group %% 16 == 0 (got %d)", p->rows_per_group);
  }

  // ---- A-side tiling (per CTA: up to 128 rows)
  int n_tiles_m1;   // M tiles of ONE CTA
  int box_h = 1, box_b = 1, box_w = 0;
  if (p->mode == SDB_A_PLAIN) {
    g.tile_rows = BM; g.ntaps = 1; g.kblocks = (int)cdiv(p->K, BK);
    n_tiles_m1 = (int)cdiv(p->M, BM);
    g.H = 1; g.W = 1;
  } else if (p->mode == SDB_A_WGRAD || p->mode == SDB_A_WGRAD_S2) {
    // C[9*Cin, N] (row = tap*Cin + ci) = sum over output pixels of X_tap[pix, ci] * dY[pix, n];  K = B*H*W
    SDB_REQUIRE(p->C % 64 == 0, "sdb_gemm: wgrad C=%d must be a multiple of 64", p->C);
    SDB_REQUIRE(p->M == 9 * p->C, "sdb_gemm: wgrad M=%d != 9*C", p->M);
    SDB_REQUIRE((long long)p->K == (long long)p->B * p->H * p->W, "sdb_gemm: wgrad K != B*H*W");
    SDB_REQUIRE(((p->W <= 64 && 64 % p->W == 0) || p->W % 64 == 0) && ((p->H * p->W) % 64 == 0 || 64 % (p->H * p->W) == 0),
                "sdb_gemm: wgrad needs W | 64 (or 64 | W) and H*W a multiple or divisor of 64 (got %dx%d)", p->H, p->W);
    SDB_REQUIRE(p->N % 8 == 0, "sdb_gemm: wgrad N=%d must be a multiple of 8", p->N);
    g.H = p->H; g.W = p->W;
    if (p->W > 64) {        // 128-wide feature maps (ResNet stem / layer1): a k-block is a 64-pixel segment of one row
      box_w = 64; box_h = 1; box_b = 1;
    } else {
      box_w = p->W;
      box_h = 64 / p->W < p->H ? 64 / p->W : p->H;
      box_b = 64 / (p->W * box_h);
    }
    g.tile_rows = (p->C % 128 == 0) ? 128 : 64;   // 64: the upper half of the UMMA tile is computed on stale rows and dropped
    g.ctiles = p->C / g.tile_rows;
    g.ntaps = 1; g.kblocks = (int)cdiv(p->K, BK);
    n_tiles_m1 = 9 * g.ctiles;
    if (p->C == 64 && env_int("SDB_WGRAD_TAP_PAIR", 1)) {   // ResNet stem / layer1, VQ-VAE level 0: two taps per 128-row tile
      g.tap_pair = 1;
      g.tile_rows = 128;
      g.ctiles = 1;
      n_tiles_m1 = 5;
    }
  } else {
    SDB_REQUIRE(p->mode == SDB_A_CONV3 || p->mode == SDB_A_CONV3S2 || p->mode == SDB_A_CONV3S2A, "sdb_gemm: bad mode %d",
                p->mode);
    SDB_REQUIRE(p->C % BK == 0, "sdb_gemm: conv C=%d must be a multiple of 64", p->C);
    SDB_REQUIRE(p->K == 9 * p->C, "sdb_gemm: conv K=%d != 9*C", p->K);
    SDB_REQUIRE((long long)p->M == (long long)p->B * p->H * p->W, "sdb_gemm: conv M != B*H*W");
    SDB_REQUIRE(p->W <= 128, "sdb_gemm: conv W=%d > 128 unsupported", p->W);
    const int H = p->H, W = p->W, B = p->B;   // output geometry
    g.H = H; g.W = W; g.ntaps = 9; g.kblocks = p->C / BK;
    if (W * H <= BM) {            // whole images per tile
      box_h = H;
      box_b = BM / (W * H);
      if (box_b > B) box_b = B;
    } else {
      box_b = 1;
      box_h = BM / W;             // rows per tile
      while (H % box_h) --box_h;  // tiles must not straddle images
    }
    g.tile_rows = W * box_h * box_b;
    n_tiles_m1 = (int)cdiv((long long)B * H * W, g.tile_rows);
  }

  // ---- CTA grouping, N tile, split-K, pipeline depth
  const int sms = num_sms();
  // CTA pairs pay ~3 us of extra fixed latency (cluster barriers, remote arrives: 12.0 vs 8.8 us for a 3-k-block
  // problem) and win on large problems through halved shared-memory traffic: crossover measured near 6 GFLOP
  const double flop = 2.0 * (double)p->M * (double)p->N * (double)p->K;
  int cg = (n_tiles_m1 >= 2 && flop >= 6e9) ? 2 : 1;
  const int force_cg = env_int("SDB_GEMM_CG", 0);
  if (force_cg == 1 || force_cg == 2) cg = force_cg;
  const bool wgrad_mode = (p->mode == SDB_A_WGRAD || p->mode == SDB_A_WGRAD_S2);
  if (wgrad_mode) g.bn = pick_bn(p->N, 128 * cg, 64 * cg);   // each CTA stages whole 64-column (128-byte) blocks of dY
  else {
    // f8c: D1 | D2 share an accumulator stage -> bn <= 128 with two stages; long-K pair tiles take bn up to 256 with ONE
    // stage instead (each staged byte is read once, so shared-memory traffic per flop falls with the tile width:
    // 256 * (256 + 1.5 bn) bytes per k-block and CTA against 64 * bn / 128 * 8 tensor cycles -- profiles/README 13)
    const bool wide = f8c && cg == 2 && p->N >= 256 && (long long)cdiv(p->K, BK) >= 16 && env_int("SDB_GEMM_F8_WIDE", 1);
    g.bn = (cg == 2) ? pick_bn(p->N, (f8c && !wide) ? 128 : 256, 32) : pick_bn(p->N, 128, geglu ? 32 : 16);
  }
  g.n_tiles_m = (int)cdiv(n_tiles_m1, cg);
  const int ksteps = g.kblocks * g.ntaps;
  const int units = sms / cg;
  const bool can_split = !p->relu && !p->out_packed && !p->gsum && !p->batch_rows;   // those epilogues need the complete sum
  if (!wgrad_mode && !can_split && env_int("SDB_GEMM_NARROW", 1)) {
    // under-filled grid that cannot use split-K (the GroupNorm-sum / packed / ReLU epilogues need complete sums): narrower
    // N tiles put more SMs to work.  Cost model: waves * (bn + fixed per-tile overhead worth ~64 columns).
    const int step = (cg == 2) ? 32 : (geglu ? 32 : 16);
    long long best_cost = -1;
    int best_bn = g.bn;
    for (int bn = g.bn; bn >= 32 * cg && bn >= step; bn /= 2) {
      if (bn % step) break;
      const long long t = (long long)g.n_tiles_m * cdiv(p->N, bn);
      const long long waves = cdiv(t, units);
      const long long cost = waves * (bn + 64);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bn = bn; }
    }
    g.bn = best_bn;
  }
  g.bnl = g.bn / cg;
  g.acc_two = !(f8c && g.bn > 128);
  g.d2_off = g.bn > 128 ? ACC_COLS : ACC_COLS / 2;
  g.n_tiles_n = (int)cdiv(p->N, g.bn);
  const long long tiles = (long long)g.n_tiles_m * g.n_tiles_n;
  g.splits = 1;
  if (can_split && tiles * 4 < (long long)units * 3 && ksteps >= 8) {
    long long s = units / tiles;
    if (s > ksteps / 4) s = ksteps / 4;
    if (s > 32) s = 32;
    if (s >= 2) g.splits = (int)s;
  }
  const int force_split = env_int("SDB_GEMM_SPLITK", 0);
  if (force_split >= 1 && can_split) g.splits = force_split > ksteps ? ksteps : force_split;
  g.stage_bytes = 2 * TILE_A_BYTES + 2 * (uint32_t)g.bnl * BK * 2;
  const int avail = SMEM_LIMIT - 1024 - (int)sizeof(GemmCtl);
  g.stages = avail / (int)g.stage_bytes;
  if (g.stages > MAX_STAGES) g.stages = MAX_STAGES;
  SDB_REQUIRE(g.stages >= 2, "sdb_gemm: tile does not fit shared memory");

  // ---- tensor maps
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, ma_l8, mb_l8;   // *_lo: fp16 lo plane, or (passes 2) the e4m3 h8 half-plane
  const __half* a = reinterpret_cast<const __half*>(p->a);
  const __half* w = reinterpret_cast<const __half*>(p->w);
  const long long w_rows = p->batch_rows ? (long long)p->w_rows : (long long)p->N;
  const long long w_cols = p->batch_rows ? (long long)p->w_cols : (long long)p->K;
  const long long w_plane = p->w_plane_stride > 0 ? (long long)p->w_plane_stride : w_rows * w_cols;
  int rc;
  if (p->mode == SDB_A_PLAIN) {
    uint64_t dims[5] = {(uint64_t)p->K, (uint64_t)p->M, 1, 1, 1};
    uint64_t st[4] = {(uint64_t)p->K * 2, (uint64_t)p->K * 2 * p->M, (uint64_t)p->K * 2 * p->M,
                      (uint64_t)p->K * 2 * p->M};
    uint32_t box[5] = {BK, BM, 1, 1, 1};
    if ((rc = make_map(&ma_hi, a, 5, dims, st, box))) return rc;
    if (f8c) {   // e4m3 half-planes: 1 byte per element, a_plane_stride BYTES each
      uint64_t st8[4] = {(uint64_t)p->K, (uint64_t)p->K * p->M, (uint64_t)p->K * p->M, (uint64_t)p->K * p->M};
      const uint8_t* a8 = reinterpret_cast<const uint8_t*>(a + p->a_plane_stride);
      if ((rc = make_map(&ma_lo, a8, 5, dims, st8, box, true))) return rc;
      if ((rc = make_map(&ma_l8, a8 + p->a_plane_stride, 5, dims, st8, box, true))) return rc;
    } else {
      if ((rc = make_map(&ma_lo, a + p->a_plane_stride, 5, dims, st, box))) return rc;
      ma_l8 = ma_lo;
    }
  } else if (wgrad_mode) {
    // the NHWC activation (or its stride-2 phase split): box = 64 channels x 64 pixels
    const int H = p->H, W = p->W, B = p->B, C = p->C;
    const uint64_t phases = (p->mode == SDB_A_WGRAD_S2) ? 4 : 1;
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, phases, (uint64_t)B};
    uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H, (uint64_t)C * 2 * W * H * phases};
    uint32_t box[5] = {64, (uint32_t)box_w, (uint32_t)box_h, 1, (uint32_t)box_b};
    if ((rc = make_map(&ma_hi, a, 5, dims, st, box))) return rc;
    if ((rc = make_map(&ma_lo, a + p->a_plane_stride, 5, dims, st, box))) return rc;
    ma_l8 = ma_lo;
  } else {
    const int H = p->H, W = p->W, B = p->B, C = p->C;
    const uint64_t phases = (p->mode == SDB_A_CONV3S2 || p->mode == SDB_A_CONV3S2A) ? 4 : 1;   // phase-split input [B][4][H][W][C]
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, phases, (uint64_t)B};
    uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H, (uint64_t)C * 2 * W * H * phases};
    uint32_t box[5] = {BK, (uint32_t)W, (uint32_t)box_h, 1, (uint32_t)box_b};
    if ((rc = make_map(&ma_hi, a, 5, dims, st, box))) return rc;
    if (f8c) {
      uint64_t st8[4] = {(uint64_t)C, (uint64_t)C * W, (uint64_t)C * W * H, (uint64_t)C * W * H * phases};
      const uint8_t* a8 = reinterpret_cast<const uint8_t*>(a + p->a_plane_stride);
      if ((rc = make_map(&ma_lo, a8, 5, dims, st8, box, true))) return rc;
      if ((rc = make_map(&ma_l8, a8 + p->a_plane_stride, 5, dims, st8, box, true))) return rc;
    } else {
      if ((rc = make_map(&ma_lo, a + p->a_plane_stride, 5, dims, st, box))) return rc;
      ma_l8 = ma_lo;
    }
  }
  if (wgrad_mode) {   // dY rows [K = pixels][N], N contiguous: 64 x 64 boxes
    uint64_t dims[2] = {(uint64_t)p->N, (uint64_t)p->K};
    uint64_t st[1] = {(uint64_t)p->N * 2};
    uint32_t box[2] = {64, BK};
    if ((rc = make_map(&mb_hi, w, 2, dims, st, box))) return rc;
    if ((rc = make_map(&mb_lo, w + w_plane, 2, dims, st, box))) return rc;
    mb_l8 = mb_lo;
  } else {
    uint64_t dims[2] = {(uint64_t)w_cols, (uint64_t)w_rows};
    uint64_t st[1] = {(uint64_t)w_cols * 2};
    uint32_t box[2] = {BK, (uint32_t)g.bnl};
    if ((rc = make_map(&mb_hi, w, 2, dims, st, box))) return rc;
    if (f8c) {
      uint64_t st8[1] = {(uint64_t)p->K};
      const uint8_t* w8 = reinterpret_cast<const uint8_t*>(w + (long long)p->N * p->K);
      if ((rc = make_map(&mb_lo, w8, 2, dims, st8, box, true))) return rc;
      if ((rc = make_map(&mb_l8, w8 + (long long)p->N * p->K, 2, dims, st8, box, true))) return rc;
    } else {
      if ((rc = make_map(&mb_lo, w + w_plane, 2, dims, st, box))) return rc;
      mb_l8 = mb_lo;
    }
  }
  cudaStream_t st = as_stream(stream);
  if (g.splits > 1) {   // partial sums are accumulated with red.global.add: C starts at zero
    if (p->ldc == p->N) SDB_CHECK(cudaMemsetAsync(p->c, 0, (size_t)p->M * p->N * sizeof(float), st));
    else SDB_CHECK(cudaMemset2DAsync(p->c, (size_t)p->ldc * sizeof(float), 0, (size_t)p->N * sizeof(float), p->M, st));
  }
  const size_t smem = (size_t)g.stages * g.stage_bytes + sizeof(GemmCtl) + 1024;
  const long long items = tiles * g.splits;
  const int groups = (int)(items < units ? items : units);
  // smallest instance that covers the call (see VAR_F8C / VAR_EXT at the kernel)
  const bool v_f8 = f8c || (p->out_packed && host_pack_mode() == SDB_FMT_F8C);
  const bool v_ext = g.alpha != 1.f || (p->gsum && g.gsum_cb == 2) || g.batch_rows != 0 || g.tap_pair != 0;
#define SDB_GEMM_LAUNCH(CGv, EPIv, VARv) \
  launch_gemm<CGv, EPIv, VARv>(ma_hi, ma_lo, mb_hi, mb_lo, ma_l8, mb_l8, g, groups, smem, st)
  if (geglu) {
    SDB_REQUIRE(!v_ext, "sdb_gemm: the GEGLU epilogue takes no alpha / 2-channel sums / batching");
    if (v_f8) return cg == 2 ? SDB_GEMM_LAUNCH(2, EPI_GEGLU, VAR_F8C) : SDB_GEMM_LAUNCH(1, EPI_GEGLU, VAR_F8C);
    return cg == 2 ? SDB_GEMM_LAUNCH(2, EPI_GEGLU, 0) : SDB_GEMM_LAUNCH(1, EPI_GEGLU, 0);
  }
  if (v_f8 && v_ext) return cg == 2 ? SDB_GEMM_LAUNCH(2, EPI_F32, VAR_F8C | VAR_EXT) : SDB_GEMM_LAUNCH(1, EPI_F32, VAR_F8C | VAR_EXT);
  if (v_f8) return cg == 2 ? SDB_GEMM_LAUNCH(2, EPI_F32, VAR_F8C) : SDB_GEMM_LAUNCH(1, EPI_F32, VAR_F8C);
  if (v_ext) return cg == 2 ? SDB_GEMM_LAUNCH(2, EPI_F32, VAR_EXT) : SDB_GEMM_LAUNCH(1, EPI_F32, VAR_EXT);
  return cg == 2 ? SDB_GEMM_LAUNCH(2, EPI_F32, 0) : SDB_GEMM_LAUNCH(1, EPI_F32, 0);
#undef SDB_GEMM_LAUNCH
}

/* issuer wait accounting of the SDB_GEMM_TIMING build: out4 = {cycles waiting for operands, cycles waiting for a free
 * accumulator, issuer lifetime cycles, tiles issued} summed over CTAs and launches since the last reset (synchronises the
 * device).  Default build: returns SDB_ERR_UNSUPPORTED and zeros. */
extern "C" int sdb_gemm_timing(uint64_t* out4, int reset) {
  SDB_REQUIRE(out4, "sdb_gemm_timing: null argument");
  for (int i = 0; i < 4; ++i) out4[i] = 0;
#if SDB_GEMM_TIMING
  unsigned long long h[4] = {0, 0, 0, 0};
  SDB_CHECK(cudaDeviceSynchronize());
  SDB_CHECK(cudaMemcpyFromSymbol(h, sdb::g_gemm_timing, sizeof(h)));
  for (int i = 0; i < 4; ++i) out4[i] = h[i];
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    SDB_CHECK(cudaMemcpyToSymbol(sdb::g_gemm_timing, z, sizeof(z)));
  }
  return 0;
#else
  (void)reset;
  sdb::set_error("sdb_gemm_timing: library built without SDB_GEMM_TIMING=1");
  return SDB_ERR_UNSUPPORTED;
#endif
}
