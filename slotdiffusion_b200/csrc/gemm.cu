// sdb_gemm: C[M,N] = A[M,K] * W[N,K]^T with fused epilogue, on the 5th-gen tensor cores.
//
// Persistent, warp-specialised kernel (one CTA per SM):
//   warp 0      TMA producer   cp.async.bulk.tensor (5-D map for A: plain / implicit-im2col 3x3 / stride-2
//                              phase-split; 2-D map for W), SWIZZLE_128B, 3-stage mbarrier ring
//   warp 1      MMA issuer     tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction, fp32
//                              accumulators in TMEM (2 accumulator stages so the epilogue of tile i overlaps
//                              the main loop of tile i+1)
//   warps 2..5  epilogue       tcgen05.ld -> registers -> smem transpose -> coalesced fp32 stores with
//                              bias / timestep-embedding row vector / residual / ReLU fused
// Operands are fp16 hi/lo planes (see sdb200.h "packed"); passes=3 issues hi*hi + lo*hi + hi*lo per k-step,
// which reproduces the fp32 product to ~2^-22 while running on the fp16 tensor pipe.
#include "common.cuh"
#include "ptx.cuh"

namespace sdb {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // fp16 elements per stage row = 128 B = one swizzle-128B row
constexpr int UK = 16;           // UMMA K for 16-bit operands
constexpr int MAX_BN = 128;
constexpr int STAGES = 3;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 128;    // TMEM columns per accumulator stage
constexpr int GEMM_THREADS = 192;
constexpr int EPI_WARPS = 4;
constexpr uint32_t TILE_A_BYTES = BM * BK * 2;       // 16 KB
constexpr uint32_t TILE_B_BYTES = MAX_BN * BK * 2;   // 16 KB (allocated for the max BN)
constexpr uint32_t STAGE_BYTES = 2 * TILE_A_BYTES + 2 * TILE_B_BYTES;  // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t EPI_STAGE_FLOATS = 32 * 33;

struct GemmSmem {
  // operand ring first: every tile must be 1024-B aligned for SWIZZLE_128B
  uint8_t ring[STAGES][STAGE_BYTES];
  float epi[EPI_WARPS][EPI_STAGE_FLOATS];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full[ACC_STAGES];
  uint64_t acc_empty[ACC_STAGES];
  uint32_t tmem_base;
};

struct GemmArgs {
  float* c;
  const float* bias;
  const float* rowvec;
  const float* residual;
  long long ldc, ldv, ldr;
  int M, N, K;
  int mode;
  int tile_rows;        // valid rows per M tile (<= 128)
  int bn;               // N tile (multiple of 16, <= 128)
  int n_tiles_m, n_tiles_n;
  int kblocks;          // K blocks of 64 per tap (plain: ceil(K/64); conv: C/64)
  int ntaps;            // 1 or 9
  int passes;
  int relu;
  int rows_per_group;
  // conv geometry for A coordinates
  int box_w, box_h, box_b;   // box extents (rows = box_w*box_h*box_b = tile_rows)
  int H, W;                  // output H, W (conv modes)
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
            const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = g.n_tiles_m * g.n_tiles_n;
  const int ksteps = g.kblocks * g.ntaps;   // stages consumed per tile
  const bool three = g.passes == 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (three) {
      tma_prefetch_desc(&map_a_lo);
      tma_prefetch_desc(&map_b_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&sm.acc_full[a], 1);
      mbar_init(&sm.acc_empty[a], EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, ACC_STAGES * ACC_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx = (three ? 2u : 1u) * (uint32_t(g.tile_rows) * BK * 2 + uint32_t(g.bn) * BK * 2);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / g.n_tiles_n, tn = tile % g.n_tiles_n;
        // A tile origin
        int x0 = 0, y0 = 0, b0 = 0;
        if (g.mode == SDB_A_PLAIN) {
          x0 = tm * BM;   // row index lives in dim 1
        } else {
          const int rows_per_img = g.H * g.W;
          const long long m0 = (long long)tm * g.tile_rows;
          b0 = int(m0 / rows_per_img);
          y0 = int((m0 % rows_per_img) / g.W);
        }
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          uint8_t* base = sm.ring[stage];
          mbar_arrive_expect_tx(&sm.full[stage], tx);
          const int tap = ks / g.kblocks;
          const int c0 = (ks % g.kblocks) * BK;
          int cx = x0, cy = y0, cp = 0;
          if (g.mode == SDB_A_CONV3) {
            cx = tap % 3 - 1;
            cy = y0 + tap / 3 - 1;
          } else if (g.mode == SDB_A_CONV3S2) {
            // input pixel (2y+ky-1, 2x+kx-1): ky=0 -> odd phase, row y-1; ky=1 -> even phase, row y; ky=2 -> odd, row y
            const int ky = tap / 3, kx = tap % 3;
            cp = ((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0);
            cx = (kx == 0) ? -1 : 0;
            cy = y0 + ((ky == 0) ? -1 : 0);
          }
          tma_load_5d(base, &map_a_hi, &sm.full[stage], c0, cx, cy, cp, b0);
          tma_load_2d(base + 2 * TILE_A_BYTES, &map_b_hi, &sm.full[stage], ks * BK, tn * g.bn);
          if (three) {
            tma_load_5d(base + TILE_A_BYTES, &map_a_lo, &sm.full[stage], c0, cx, cy, cp, b0);
            tma_load_2d(base + 2 * TILE_A_BYTES + TILE_B_BYTES, &map_b_lo, &sm.full[stage], ks * BK, tn * g.bn);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BM, g.bn);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&sm.acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(sm.ring[stage]);
          const uint32_t a_lo = a_hi + TILE_A_BYTES;
          const uint32_t b_hi = a_hi + 2 * TILE_A_BYTES;
          const uint32_t b_lo = b_hi + TILE_B_BYTES;
          const uint64_t da_hi = umma_desc_kmajor_sw128(a_hi), da_lo = umma_desc_kmajor_sw128(a_lo);
          const uint64_t db_hi = umma_desc_kmajor_sw128(b_hi), db_lo = umma_desc_kmajor_sw128(b_lo);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t adv = uint64_t((k * UK * 2) >> 4);   // 32 B per k-step inside the 128-B swizzle row
            umma_f16(d_tmem, da_hi + adv, db_hi + adv, idesc, (ks | k) != 0);
            if (three) {
              umma_f16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1);
              umma_f16(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
            }
          }
          umma_commit(&sm.empty[stage]);   // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&sm.acc_full[as]);     // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    float* stg = sm.epi[warp - 2];
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / g.n_tiles_n, tn = tile % g.n_tiles_n;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&sm.acc_full[as], aphase);
      tc_fence_after();
      const long long row0 = (long long)tm * g.tile_rows + q * 32;   // first output row of this warp
      const int rows_valid = min(g.tile_rows - q * 32, 32);          // rows of this warp inside the tile
      for (int cb = 0; cb < g.bn; cb += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * ACC_COLS + cb);
        if (g.bn - cb >= 32) {
          tmem_ld_32x32(taddr, r);
        } else {
          uint32_t r16[16];
          tmem_ld_32x16(taddr, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { r[j] = r16[j]; r[j + 16] = 0; }
        }
        tmem_ld_wait();
        // thread `lane` holds row (q*32+lane), columns cb..cb+31 -> transpose through smem
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int n = tn * g.bn + cb + lane;
        const bool ncol_ok = (cb + lane < g.bn) && (n < g.N);
        const float bias = (ncol_ok && g.bias) ? g.bias[n] : 0.f;
        for (int i = 0; i < rows_valid; ++i) {
          const long long m = row0 + i;
          if (m >= g.M) break;
          if (ncol_ok) {
            float v = stg[i * 33 + lane] + bias;
            if (g.rowvec) v += g.rowvec[(m / g.rows_per_group) * g.ldv + n];
            if (g.residual) v += g.residual[m * g.ldr + n];
            if (g.relu) v = fmaxf(v, 0.f);
            g.c[m * g.ldc + n] = v;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&sm.acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ACC_STAGES * ACC_COLS);
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
                    const uint32_t* box) {
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
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

static int pick_bn(int N) {
  // largest multiple of 16 <= 128 that minimises padded work
  int best = 16;
  long long best_cost = -1;
  for (int bn = 128; bn >= 16; bn -= 16) {
    long long tiles = cdiv(N, bn);
    long long cost = tiles * bn * 1000 + tiles * 40;   // padded columns dominate, then tile count
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_gemm(const SdbGemm* p, void* stream) {
  SDB_REQUIRE(p && p->a && p->w && p->c, "sdb_gemm: null operand");
  SDB_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "sdb_gemm: empty problem M=%d N=%d K=%d", p->M, p->N, p->K);
  SDB_REQUIRE(p->passes == 1 || p->passes == 3, "sdb_gemm: passes must be 1 or 3");
  SDB_REQUIRE(p->K % 8 == 0, "sdb_gemm: K=%d must be a multiple of 8 (16-byte TMA rows)", p->K);
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(p->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
              "sdb_gemm: operands must be 16-byte aligned");
  SDB_REQUIRE(!p->rowvec || p->rows_per_group > 0, "sdb_gemm: rowvec needs rows_per_group");
  GemmArgs g{};
  g.c = p->c; g.bias = p->bias; g.rowvec = p->rowvec; g.residual = p->residual;
  g.ldc = p->ldc; g.ldv = p->ldv; g.ldr = p->ldr;
  g.M = p->M; g.N = p->N; g.K = p->K; g.mode = p->mode; g.passes = p->passes; g.relu = p->relu;
  g.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : 1;
  g.bn = pick_bn(p->N);
  g.n_tiles_n = (int)cdiv(p->N, g.bn);

  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  const __half* a = reinterpret_cast<const __half*>(p->a);
  const __half* w = reinterpret_cast<const __half*>(p->w);
  int rc;
  if (p->mode == SDB_A_PLAIN) {
    g.tile_rows = BM; g.ntaps = 1; g.kblocks = (int)cdiv(p->K, BK);
    g.n_tiles_m = (int)cdiv(p->M, BM);
    g.H = 1; g.W = 1; g.box_w = BM; g.box_h = 1; g.box_b = 1;
    uint64_t dims[5] = {(uint64_t)p->K, (uint64_t)p->M, 1, 1, 1};
    uint64_t st[4] = {(uint64_t)p->K * 2, (uint64_t)p->K * 2 * p->M, (uint64_t)p->K * 2 * p->M,
                      (uint64_t)p->K * 2 * p->M};
    uint32_t box[5] = {BK, BM, 1, 1, 1};
    if ((rc = make_map(&ma_hi, a, 5, dims, st, box))) return rc;
    if ((rc = make_map(&ma_lo, a + p->a_plane_stride, 5, dims, st, box))) return rc;
  } else {
    SDB_REQUIRE(p->mode == SDB_A_CONV3 || p->mode == SDB_A_CONV3S2, "sdb_gemm: bad mode %d", p->mode);
    SDB_REQUIRE(p->C % BK == 0, "sdb_gemm: conv C=%d must be a multiple of 64", p->C);
    SDB_REQUIRE(p->K == 9 * p->C, "sdb_gemm: conv K=%d != 9*C", p->K);
    SDB_REQUIRE((long long)p->M == (long long)p->B * p->H * p->W, "sdb_gemm: conv M != B*H*W");
    SDB_REQUIRE(p->W <= 128, "sdb_gemm: conv W=%d > 128 unsupported", p->W);
    const int H = p->H, W = p->W, B = p->B, C = p->C;   // output geometry
    g.H = H; g.W = W; g.ntaps = 9; g.kblocks = C / BK;
    // tile = box_b images x box_h rows x full width
    int box_h, box_b;
    if (W * H <= BM) {            // whole images per tile
      box_h = H;
      box_b = BM / (W * H);
      if (box_b > B) box_b = B;
    } else {
      box_b = 1;
      box_h = BM / W;             // rows per tile
      while (H % box_h) --box_h;  // tiles must not straddle images
    }
    g.box_w = W; g.box_h = box_h; g.box_b = box_b;
    g.tile_rows = W * box_h * box_b;
    g.n_tiles_m = (int)cdiv((long long)B * H * W, g.tile_rows);
    if (p->mode == SDB_A_CONV3) {
      uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, 1, (uint64_t)B};
      uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H, (uint64_t)C * 2 * W * H};
      uint32_t box[5] = {BK, (uint32_t)W, (uint32_t)box_h, 1, (uint32_t)box_b};
      if ((rc = make_map(&ma_hi, a, 5, dims, st, box))) return rc;
      if ((rc = make_map(&ma_lo, a + p->a_plane_stride, 5, dims, st, box))) return rc;
    } else {
      // phase-split input [B][4][H][W][C]
      uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, 4, (uint64_t)B};
      uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H, (uint64_t)C * 2 * W * H * 4};
      uint32_t box[5] = {BK, (uint32_t)W, (uint32_t)box_h, 1, (uint32_t)box_b};
      if ((rc = make_map(&ma_hi, a, 5, dims, st, box))) return rc;
      if ((rc = make_map(&ma_lo, a + p->a_plane_stride, 5, dims, st, box))) return rc;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
    uint64_t st[1] = {(uint64_t)p->K * 2};
    uint32_t box[2] = {BK, (uint32_t)g.bn};
    if ((rc = make_map(&mb_hi, w, 2, dims, st, box))) return rc;
    if ((rc = make_map(&mb_lo, w + (long long)p->N * p->K, 2, dims, st, box))) return rc;
  }
  const size_t smem = sizeof(GemmSmem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    SDB_CHECK(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int tiles = g.n_tiles_m * g.n_tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_kernel<<<grid, GEMM_THREADS, smem, as_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, g);
  SDB_LAUNCH_CHECK();
  return 0;
}
