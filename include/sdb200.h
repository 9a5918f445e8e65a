/* libsdb200 -- C ABI of the B200-native SlotDiffusion hot path.
 *
 * The reference (Wuziyi616/SlotDiffusion) is pure Python/PyTorch and has no FFI; every entry
 * point below replaces the ATen/cuDNN/cuBLAS call sequence that the cited reference lines
 * issue from eager PyTorch.  The Python host (slotdiffusion_b200/*.py)
 * binds these with ctypes and mirrors the reference nn.Module interfaces on top.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; the library never
 *     allocates or frees device memory and keeps no pointer after the call returns);
 *   - `stream` is a cudaStream_t; calls only enqueue work (no sync, no host read-back), so a
 *     sequence of calls is CUDA-graph capturable;
 *   - return 0 on success, SDB_ERR_* otherwise; sdb_last_error() gives the message (thread-local);
 *   - fp32 tensors are row-major contiguous unless a leading dimension is passed;
 *   - "packed" = GEMM operand format: 16-bit [2][rows][K], plane 0 = hi = fp16(x), plane 1 = lo =
 *     fp16(x - hi) (gradient operands of the backward pass use the same layout with bf16, see SdbGemm.a_bf16).  hi*hi + hi*lo + lo*hi on the fp16 tensor pipe with fp32 accumulation
 *     reproduces an fp32 product to ~2^-22 (DESIGN.md "precision").
 *     SDB_FMT_F8C ("fp8-corrected", inference): the bytes of plane 1 hold two e4m3 half-planes instead,
 *     [rows][K] h8 = e4m3(hi * 2^eh) then [rows][K] l8 = e4m3(lo * 2^el); a product is then one fp16 pass hi*hi
 *     plus the two correction products l8*h8' + h8*l8' on the fp8 pipe (twice the fp16 rate) in a second
 *     accumulator: C = D1 + 2^-s D2 -- two pass-equivalents instead of three.  Activations use static exponents
 *     (eh = 2, el = 12), weights a per-tensor exponent given to the sdb_pack_weight*_fmt calls.
 */
#ifndef SDB200_H_
#define SDB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_FMT_F16X2 0 /* plane 1 = fp16 lo plane */
#define SDB_FMT_F8C 2   /* plane 1 = e4m3(hi) | e4m3(lo) half-planes */

#define SDB_OK 0
#define SDB_ERR_INVALID 1 /* bad shape / alignment / argument */
#define SDB_ERR_CUDA 2    /* a CUDA runtime/driver call failed */
#define SDB_ERR_UNSUPPORTED 3

int sdb_version(void);
const char* sdb_last_error(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t sdb_launch_count(void);

/* ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA)
 * C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ rowvec[m / rows_per_group, N]) (+ residual[M,N])
 * A and W are packed operands.  Replaces nn.Linear / nn.Conv2d(1x1, 3x3) calls:
 *   unet.py:222,249,258 (ResBlock convs), :408 (input conv), :161-168 (Downsample), :111-112
 *   (Upsample conv), attention.py:175-180,278-295 (to_q/k/v/out, proj_in/out), :44,60 (GEGLU FF),
 *   unet.py:400-403,240 (time_embed, emb_layers), slot_attention.py:41,43-44,47,50-52.
 * mode SDB_A_PLAIN : A is [M,K] row-major.
 * mode SDB_A_CONV3 : A is an NHWC activation [B,H,W,C]; implicit im2col, 3x3, pad 1, stride 1;
 *                    M = B*H*W, K = 9*C with k = (ky*3+kx)*C + c  (W packed by sdb_pack_weight_conv3).
 * mode SDB_A_CONV3S2: A is a phase-split NHWC activation [B,4,H,W,C] (phase = (iy&1)*2 + (ix&1),
 *                    H,W = OUTPUT size = input/2) produced by sdb_pack_nhwc(..., SDB_PACK_PHASE2);
 *                    3x3, pad 1, stride 2.
 */
#define SDB_A_PLAIN 0
#define SDB_A_CONV3 1
#define SDB_A_CONV3S2 2
/* weight-gradient forms (training, nn.Conv2d backward of the lines above): the contraction runs over the OUTPUT PIXELS.
 * A is the SAME packed NHWC activation the forward conv consumed ([B,H_in,W_in,C]; mode 4: the stride-2 phase split),
 * W is the packed output gradient rows [B*H*W][N] (N contiguous): both are MN-major tensor-core operands, the 3x3 taps
 * and zero padding are again TMA box shifts.  H, W = OUTPUT size.  C[9*C, N] with row = tap*C + ci;  M = 9*C, K = B*H*W. */
#define SDB_A_WGRAD 3
#define SDB_A_WGRAD_S2 4
/* 3x3, stride 2 with the ASYMMETRIC padding of the VQ-VAE encoder's Downsample (modules.py:44-48: F.pad (0,1,0,1) then
 * conv(stride 2, padding 0)): output (y,x) reads input (2y+ky, 2x+kx); same phase-split operand as SDB_A_CONV3S2. */
#define SDB_A_CONV3S2A 5

typedef struct SdbGemm {
  const void* a;         /* packed A, planes are a_plane_stride halves apart */
  const void* w;         /* packed W [2][N][K] */
  float* c;              /* [M, ldc] */
  const float* bias;     /* [N] or NULL */
  const float* rowvec;   /* [M / rows_per_group, ldv] or NULL (ResBlock timestep-embedding add, unet.py:280-283) */
  const float* residual; /* [M, ldr] or NULL */
  int64_t a_plane_stride; /* halves between the hi and lo plane of A */
  int64_t ldc, ldv, ldr;
  int32_t M, N, K;
  int32_t mode;
  int32_t B, H, W, C;     /* conv geometry (modes 1,2) */
  int32_t rows_per_group; /* H*W for rowvec */
  int32_t passes;         /* 3 = hi*hi + lo*hi + hi*lo (fp32-faithful), 1 = hi*hi only, 2 = both operands are SDB_FMT_F8C:
                             hi*hi (kind::f16) + corr_scale * (l8*h8' + h8*l8') (kind::f8f6f4, second accumulator) */
  int32_t relu;           /* apply max(.,0) to the result (slot-attention MLP, slot_attention.py:51) */
  void* out_packed;       /* optional: also emit act(result) as a packed operand [2][M][N] for the next GEMM (c may then be NULL) */
  float* gsum;            /* optional: atomically accumulate GroupNorm partial sums [M / rows_per_group][N / 4][2] = (sum, sum of
                             squares) of the result per sample and 4-channel block (consumed by sdb_groupnorm_apply_pack) */
  int64_t out_plane_stride; /* halves between the hi and lo plane of out_packed */
  int32_t out_act;        /* activation of the packed copy: 0 none, 1 SiLU, 2 ReLU */
  int32_t geglu;          /* GEGLU epilogue (attention.py:46-48): W / bias rows interleaved by sdb_pack_weight_geglu; the only
                             output is out_packed [2][M][N/2] = a * gelu_erf(g) */
  int32_t a_bf16, w_bf16; /* the planes of A / W hold a BF16 hi/lo split (gradient operands from sdb_grad_pack /
                             sdb_pack_zero_up2: fp32 exponent range, ~16 mantissa bits) instead of the FP16 split.
                             tcgen05 kind::f16 takes ONE input format: a_bf16 must equal w_bf16 */
  float corr_scale;       /* passes == 2: 2^-s with s = el(A) + eh(W) = eh(A) + el(W) (= 12 + wexp of the packed weight) */
  int32_t gsum_cb;        /* channels per GroupNorm partial-sum block of `gsum`: 0 or 4 = [M/rows_per_group][N/4][2]; 2 =
                             [..][N/2][2] (64-channel layers with 32 groups: ResNet stem / layer1, VQ-VAE 64-channel levels) */
  int64_t w_plane_stride; /* halves between the hi and lo plane of W; 0 = N*K (a whole packed tensor).  Non-zero when W is
                             a row range of a larger packed tensor (per-sample K / V^T operands of the VQ-VAE AttnBlock) */
  /* block-diagonal ("batched") products in one launch: rows [b*batch_rows, (b+1)*batch_rows) of A meet the W operand
   * shifted by b*w_row_step rows and b*w_k_step columns; W is then a [w_rows, w_cols] packed tensor (N, K = the per-batch
   * extents).  S_b = q_b k_b^T: w_row_step = N; O_b = P_b v_b with v^T stored [C, B*L]: w_k_step = K.  batch_rows = 0: off;
   * batch_rows % 256 == 0. */
  int32_t batch_rows, w_row_step, w_k_step;
  float alpha;            /* C = alpha * (A W^T) + bias + ...; 0 means 1.  Power-of-two operand scalings are undone here: the
                             softmax probabilities of the VQ-VAE AttnBlock are packed as 2^12 P (sdb_softmax_pack out_scale) so
                             that their fp16 lo plane stays out of the subnormal range, O = 2^-12 (2^12 P) V */
  int64_t w_rows, w_cols;
} SdbGemm;

int sdb_gemm(const SdbGemm* p, void* stream);

/* ------------------------------------------------------------------ operand producers */
/* W [N,K] fp32 (nn.Linear / 1x1 conv weight) -> packed [2][N][K] */
int sdb_pack_weight(const float* w, void* out, int64_t N, int64_t K, void* stream);
/* W [Cout,Cin,3,3] -> packed [2][Cout][9*Cin], k = (ky*3+kx)*Cin + c */
int sdb_pack_weight_conv3(const float* w, void* out, int64_t Cout, int64_t Cin, void* stream);

/* the same with an explicit operand format: fmt SDB_FMT_F16X2 (wexp ignored) or SDB_FMT_F8C with h8 = e4m3(hi * 2^wexp),
 * l8 = e4m3(lo * 2^(wexp+10)); the caller picks wexp = floor(log2(448 / max|w|)) per tensor */
int sdb_pack_weight_fmt(const float* w, void* out, int64_t N, int64_t K, int fmt, int wexp, void* stream);
int sdb_pack_weight_conv3_fmt(const float* w, void* out, int64_t Cout, int64_t Cin, int fmt, int wexp, void* stream);
/* Format written by every ACTIVATION operand producer (sdb_pack_rows, sdb_layernorm_pack, sdb_groupnorm_apply_pack*,
 * sdb_pack_nhwc, sdb_geglu_pack, the packed epilogues of sdb_gemm, ...) from this point of `stream` on: a one-thread kernel
 * flips a device flag, so the switch is stream-ordered and CUDA-graph capturable.  Default SDB_FMT_F16X2. */
int sdb_set_pack_mode(int fmt, void* stream);

/* rows x [M,K] (ldx) -> packed [2][M][K]; act: 0 none, 1 SiLU (unet.py:237-238), 2 ReLU */
int sdb_pack_rows(const float* x, int64_t ldx, void* out, int64_t M, int64_t K, int act, void* stream);

/* LayerNorm over the last dim (eps, affine) then pack: nn.LayerNorm at attention.py:232-234,
 * slot_attention.py:36,40,49.  y (optional, may be NULL) also receives the fp32 result. */
int sdb_layernorm_pack(const float* x, const float* gamma, const float* beta, float eps, void* out, float* y,
                       int64_t M, int64_t C, void* stream);

/* GroupNorm statistics of an NHWC activation that is the channel-concatenation of x1 [B,HW,C1] and
 * x2 [B,HW,C2] (x2 may be NULL, C2 = 0): stats[b][g] = (mean, rstd); groups of (C1+C2)/G channels.
 * GroupNorm32 (unet/utils.py:136-139, eps 1e-5) and Normalize (attention.py:77-79, eps 1e-6). */
int sdb_groupnorm_stats(const float* x1, int64_t C1, const float* x2, int64_t C2, float* stats, int64_t B,
                        int64_t HW, int G, float eps, void* stream);
/* apply GN (+ optional SiLU) and pack: out = packed [2][B*HW][C1+C2] */
int sdb_groupnorm_apply_pack(const float* x1, int64_t C1, const float* x2, int64_t C2, const float* stats,
                             const float* gamma, const float* beta, void* out, int64_t B, int64_t HW, int G,
                             int silu, void* stream);

/* Same normalisation with the statistics taken either from `stats` [B,G,2] or (stats == NULL) derived on the fly from
 * the partial sums gsum1 [B][C1/4][2] (+ gsum2 [B][C2/4][2]) accumulated by the sdb_gemm epilogues that produced x1 / x2
 * (biased variance as E[x^2] - mean^2 in fp32).  Removes the separate statistics pass over the activation. */
int sdb_groupnorm_apply_pack_fused(const float* x1, int64_t C1, const float* gsum1, const float* x2, int64_t C2,
                                   const float* gsum2, const float* stats, const float* gamma, const float* beta,
                                   void* out, int64_t B, int64_t HW, int G, float eps, int silu, void* stream);
/* training form: nn.Dropout(p) after the SiLU (ResBlock out_layers, unet.py:245-246) with a counter-based mask
 * (seed, element index) that sdb_groupnorm_bwd regenerates; stats [B,G,2] required.  seed_dev (optional): device
 * pointer to a per-step counter mixed into the seed, so a CUDA-graph-captured step draws a new mask per replay. */
int sdb_groupnorm_apply_pack_dropout(const float* x1, int64_t C1, const float* x2, int64_t C2, const float* stats,
                                     const float* gamma, const float* beta, void* out, int64_t B, int64_t HW, int G,
                                     int silu, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* per-(sample, 4-channel block) [sum, sum of squares] of an NHWC activation x [B*HW, C], ACCUMULATED into gsum
 * [B, C/4, 2] (zero it first) -- the same partial sums sdb_gemm's `gsum` epilogue emits, for activations that no GEMM
 * produced (output of sdb_conv3_in, unet.py:408), so their GroupNorm consumers can use the fused statistics path */
int sdb_channel_block_sums(const float* x, int64_t C, float* gsum, int64_t B, int64_t HW, void* stream);
/* partial sums -> stats [B,G,2] (mean, rstd); sdb_groupnorm_finalize_cb: sums in blocks of cb channels (2 or 4) */
int sdb_groupnorm_finalize(const float* gsum1, int64_t C1, const float* gsum2, int64_t C2, float* stats, int64_t B,
                           int64_t HW, int G, float eps, void* stream);
int sdb_groupnorm_finalize_cb(const float* gsum, int64_t C, float* stats, int64_t B, int64_t HW, int G, float eps, int cb,
                              void* stream);
/* GEGLU.proj weight [2F,K] (+ bias [2F]) -> packed rows interleaved [16 a | 16 g] per 32-row chunk (+ permuted bias),
 * the operand layout of sdb_gemm's `geglu` epilogue (attention.py:39-48) */
int sdb_pack_weight_geglu(const float* w, const float* bias, void* out, float* bias_out, int64_t F, int64_t K,
                          void* stream);

/* raw NHWC -> packed, with layout transform.  x1 [B,H,W,C1] (+ x2 [B,H,W,C2] concatenated on C).
 *  SDB_PACK_PLAIN : out [2][B*H*W][C]                      (skip_connection 1x1 conv input, unet.py:256-259)
 *  SDB_PACK_UP2   : nearest x2 upsample, out [2][B*2H*2W][C]   (Upsample, unet.py:118)
 *  SDB_PACK_PHASE2: out [2][B][4][H/2][W/2][C], phase = (y&1)*2+(x&1)   (Downsample stride-2 conv, unet.py:161-168)
 *  y_cat (optional): also write the fp32 concatenation [B,H,W,C1+C2] (th.cat at unet.py:572). */
#define SDB_PACK_PLAIN 0
#define SDB_PACK_UP2 1
#define SDB_PACK_PHASE2 2
int sdb_pack_nhwc(const float* x1, int64_t C1, const float* x2, int64_t C2, void* out, float* y_cat, int64_t B,
                  int64_t H, int64_t W, int mode, void* stream);

/* GEGLU: u [M, 2F] -> packed [2][M][F] of u[:, :F] * gelu_erf(u[:, F:])   (attention.py:46-48) */
int sdb_geglu_pack(const float* u, void* out, int64_t M, int64_t F, void* stream);

/* sinusoidal timestep embedding, [cos | sin] order, fp32 freqs (unet/utils.py:79-86) -> packed [2][B][dim] */
int sdb_timestep_embedding_pack(const float* t, void* out, int64_t B, int dim, void* stream);

/* Diagnostic (only in a library built with SDB_GEMM_TIMING=1; otherwise SDB_ERR_UNSUPPORTED): where the MMA-issuing
 * thread of sdb_gemm waits.  out4 = {cycles waiting for operand stages, cycles waiting for a free accumulator, issuer
 * lifetime cycles, tiles issued}, summed over CTAs and launches since the last reset.  Synchronises the device. */
int sdb_gemm_timing(uint64_t* out4, int reset);

/* ------------------------------------------------------------------ attention core (attention.py:188-205)
 * out[b, i, h*d:(h+1)*d] = softmax_j(scale * q[b,i,h,:] . k[b,j,h,:]) v[b,j,h,:],   d = head dim (32)
 * q [B,Lq,*] with row stride ldq, k/v [B,Lk,*] with row strides ldk/ldv (views into fused projections).
 * out: packed [2][B*Lq][heads*d] (feeds to_out) */
int sdb_attention_pack(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                       void* out, int64_t B, int64_t Lq, int64_t Lk, int heads, int d, float scale, void* stream);
/* same contract on the tensor cores (tcgen05, fp32-faithful split-fp16 products, online softmax over 64-key chunks);
 * head dim 32 only (sdb_attention_tc_supported) */
int sdb_attention_tc_supported(int64_t heads, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv);
int sdb_attention_tc(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, void* out,
                     int64_t B, int64_t Lq, int64_t Lk, int heads, int d, float scale, void* stream);
/* Few keys (Lk <= 32: the slot cross-attention, attention.py:188-205 with context = slots): fp32 CUDA-core kernel, one CTA per
 * 32 query rows with all heads, coalesced tile loads / stores; head dim 32, heads <= 16, default operand format only
 * (sdb_attention_fewkeys_supported) */
int sdb_attention_fewkeys_supported(int64_t heads, int64_t d, int64_t Lk, int64_t ldq, int64_t ldk, int64_t ldv);
int sdb_attention_fewkeys(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, void* out,
                          int64_t B, int64_t Lq, int64_t Lk, int heads, int d, float scale, void* stream);


/* ------------------------------------------------------------------ small-channel convolutions
 * input conv (unet.py:408): x NCHW [B,Cin,H,W] (Cin small, e.g. 3) -> NHWC fp32 [B,H,W,Cout], 3x3 pad 1 */
int sdb_conv3_in(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t Cin, int64_t H,
                 int64_t W, int64_t Cout, void* stream);
/* output head (unet.py:537-542): y NCHW [B,Cout,H,W] = conv3x3(SiLU(GN(h))) for NHWC h [B,H,W,C], Cout small */
int sdb_conv3_out(const float* h, const float* stats, const float* gamma, const float* beta, const float* w,
                  const float* bias, float* y, int64_t B, int64_t H, int64_t W, int64_t C, int G, int64_t Cout,
                  void* stream);

/* ------------------------------------------------------------------ Slot Attention (slot_attention.py:55-104,
 * sa_diffusion.py:15-70).  One attention iteration over the tokens of every sample, fused:
 *   logits = scale * k q^T ; attn = softmax over slots ; (optional) seg_mask[b,s,n] = attn ;
 *   a = attn + eps ; updates[b,s,:] = sum_n a[n,s] v[n,:] / sum_n a[n,s]
 * kv [B,N,2D] (k | v fused projection of LayerNorm(inputs)), q [B,S,D].
 * updates: packed [2][B*S][D] (feeds the GRU input GEMM) and fp32 upd32 (optional, may be NULL).
 * work: fp32 scratch of sdb_slot_attend_workspace(B,N,S,D) bytes. */
int64_t sdb_slot_attend_workspace(int64_t B, int64_t N, int64_t S, int64_t D);
int sdb_slot_attend(const float* kv, const float* q, float* seg_mask, void* upd_packed, float* upd32, float* work,
                    int64_t B, int64_t N, int64_t S, int64_t D, float scale, float eps, void* stream);
/* same, also returning colsum[b,s] = sum_n (attn + eps) (needed by sdb_slot_attend_bwd) */
int sdb_slot_attend_train(const float* kv, const float* q, float* seg_mask, void* upd_packed, float* upd32,
                          float* colsum, float* work, int64_t B, int64_t N, int64_t S, int64_t D, float scale, float eps,
                          void* stream);
/* Tensor-core form of one Slot-Attention iteration for the inference path (slot_attention.py:67-91 with the k / v
 * projections folded away): x [B,N,Din] RAW encoder features (the LayerNorm of slot_attention.py:68 is applied
 * in-kernel without its affine part, which the caller folds into qa and into the GRU input weight);
 * qa [B*S, ldq]: columns [0,Din) = scale * gamma * (Wk^T Wq LN_q(slots)), column Din = the beta term of the logits.
 * Returns U[b,s,:] = sum_n a[n,s] n[n,:] / sum_n a[n,s] (a = softmax_s(logits) + eps, n = normalised features) as a
 * packed GEMM operand [2][B*S][Din] (+ optional fp32 copy) and the seg mask [B,S,N] (optional).
 * Supported: Din in {128,192,256}, S <= 32 (sdb_slot_attend_fused_supported); work: _workspace() bytes. */
int sdb_slot_attend_fused_supported(int64_t S, int64_t Din);
/* debug aid: device buffer of 6*64 int64 receiving the per-role timeline (SM cycles) of CTA (0,0); NULL switches it off */
int sdb_slot_attend_fused_debug(void* buf);
int64_t sdb_slot_attend_fused_workspace(int64_t B, int64_t N, int64_t S, int64_t Din);
int sdb_slot_attend_fused(const float* x, const float* qa, int64_t ldq, float* seg_mask, void* upd_packed,
                          float* upd32, float* work, int64_t B, int64_t N, int64_t S, int64_t Din, float ln_eps,
                          float eps, void* stream);
/* The attend kernel alone: leaves its per-chunk partial sums in `work` for sdb_slot_update (no finalize launch).
 * work = part_upd [B, chunks, S, Din] (= ascale * sum_n a[n,s] n[n,:]) | part_cs [B, chunks, S] (= sum_n a[n,s]) with
 * chunks = sdb_slot_attend_fused_chunks(B, N), ascale = sdb_slot_attend_fused_ascale(). */
int64_t sdb_slot_attend_fused_chunks(int64_t B, int64_t N);
float sdb_slot_attend_fused_ascale(void);
int sdb_slot_attend_fused_partials(const float* x, const float* qa, int64_t ldq, float* seg_mask, float* work,
                                   int64_t B, int64_t N, int64_t S, int64_t Din, float ln_eps, float eps, void* stream);

/* ------------------------------------------------------------------ slot update: everything of one Slot-Attention
 * iteration that is not the token contraction, in ONE launch (slot_attention.py:82 and :97-102):
 *   U      = sum_chunks part_upd / (ascale * sum_chunks part_cs)          (finalize of the attend kernel)
 *   h'     = GRUCell(updates, slots_in)   with gi = U W_iv^T + b_iv  (W_iv = W_ih Wv diag(gamma), b_iv = W_ih Wv beta + b_ih)
 *   slots  = h' + W_2 relu(W_1 LN_m(h') + b_1) + b_2                       -> slots_out [rows, D]
 *   qa     = LN_q(slots) W_qa^T  (next iteration's folded slot-side operand, sdb_slot_attend_fused) -> qa_out [rows, ldq]
 * do_update = 0: only the last line, applied to slots_in (the projection of the initial slots).
 * All weights fp32 and TRANSPOSED (input-major, [K][N] contiguous in N): w_ivT [Din][3D], w_hhT [D][3D], w1T [D][M],
 * w2T [M][D], w_qaT [D][ldq] (column Din = logit bias row, columns > Din ignored).  rows = B * S.
 * Plain fp32 FMA arithmetic (no operand splitting); intended for rows <~ 1k where the tail is launch-latency bound. */
typedef struct SdbSlotUpdate {
  const float* part_upd;
  const float* part_cs;
  const float* slots_in;
  const float* w_ivT;
  const float* b_iv;
  const float* w_hhT;
  const float* b_hh;
  const float* ln_m_g;
  const float* ln_m_b;
  const float* w1T;
  const float* b1;
  const float* w2T;
  const float* b2;
  const float* ln_q_g;
  const float* ln_q_b;
  const float* w_qaT;
  float* slots_out;
  float* qa_out;          /* NULL: no projection (last iteration) */
  int64_t rows;
  int32_t S, Din, D, M, ldq, chunks;
  float ascale, ln_m_eps, ln_q_eps;
  int32_t do_update;
} SdbSlotUpdate;
int sdb_slot_update_supported(int64_t S, int64_t Din, int64_t D, int64_t M);
int sdb_slot_update(const SdbSlotUpdate* args, void* stream);

/* ------------------------------------------------------------------ Slot Attention, the whole iterative update as ONE
 * persistent kernel (slot_attention.py:67-104, SlotAttentionWMask: sa_diffusion.py:28-70; inference path).
 * A thread-block cluster of 4 / 8 CTAs owns a sample: its raw features x [N, Din] are read from HBM once, LayerNorm-ed and
 * split into fp16 hi/lo tcgen05 operand tiles in shared memory, and all `iterations` run from that resident copy (logits and
 * weighted-sum contractions on the tensor cores, softmax over slots, GRU / MLP / next q projection on the CUDA cores,
 * column-split over the CTAs of the cluster and exchanged through distributed shared memory).  Same folded algebra as
 * sdb_slot_attend_fused + sdb_slot_update.  Weights fp32, TRANSPOSED and k-quad interleaved, w4[K/4][ncols][4]:
 * w_iv4 [Din/4][3D][4], w_hh4 [D/4][3D][4], w1_4 [D/4][M][4], w2_4 [M/4][D][4], w_qa4 [D/4][ldq][4] (column Din = logit bias).
 * slots_in / slots_out [B, S, D]; seg_mask [B, S, N] (last-iteration softmax over slots) or NULL.
 * Supported (sdb_slot_attention_resident_supported): D == Din in {128,192,256}, S <= 32, N up to 8 x 256 (Din <= 192) /
 * 8 x 128 tokens, M % 64 == 0. */
typedef struct SdbSlotAttentionResident {
  const float* x;
  const float* slots_in;
  float* slots_out;
  float* seg_mask;
  const float* w_iv4;
  const float* b_iv;
  const float* w_hh4;
  const float* b_hh;
  const float* ln_m_g;
  const float* ln_m_b;
  const float* w1_4;
  const float* b1;
  const float* w2_4;
  const float* b2;
  const float* ln_q_g;
  const float* ln_q_b;
  const float* w_qa4;
  int64_t B;
  int32_t N, S, Din, D, M, ldq, iterations;
  float ln_in_eps, attn_eps, ln_m_eps, ln_q_eps;
} SdbSlotAttentionResident;
int sdb_slot_attention_resident_supported(int64_t N, int64_t S, int64_t Din, int64_t D, int64_t M);
int sdb_slot_attention_resident(const SdbSlotAttentionResident* args, void* stream);
/* number of samples processed concurrently (resident clusters) for this geometry on the current device; the persistent
 * kernel loops over the batch in waves of this size.  0 = unsupported geometry. */
int64_t sdb_slot_attention_resident_wave(int64_t N, int64_t S, int64_t Din, int64_t D, int64_t M);
/* debug aid: device buffer of 256 int64 receiving the phase timeline (SM cycles) of thread 0 of CTA 0; NULL switches it off */
int sdb_slot_attention_resident_debug(void* buf);

/* GRUCell pointwise part (PyTorch gate order r,z,n; slot_attention.py:97-100): gi = x W_ih^T + b_ih and
 * gh = h W_hh^T + b_hh come from sdb_gemm; h_new = (1-z) n + z h.  All [R, 3D] / [R, D]. */
int sdb_gru_gates(const float* gi, const float* gh, const float* h, float* h_new, int64_t R, int64_t D,
                  void* stream);

/* ------------------------------------------------------------------ DPM-Solver++ glue (dpm_solver.py:523-534,
 * :804-831; quantize.py:84-94).  Latents are NCHW fp32 [B,C,HW]. */
/* x0 = (x - sigma*eps)/alpha ; if codebook != NULL: x0 = nearest codebook row (squared distance, first min) */
int sdb_dpm_x0(const float* x, const float* eps, float alpha, float sigma, const float* codebook, int64_t ncodes,
               float* x0, int32_t* idx /*optional*/, int64_t B, int64_t C, int64_t HW, void* stream);
/* out = a*x + b*m0 + c*(m1 - m0)   (m1 may be NULL -> c ignored) */
int sdb_lincomb(float* out, const float* x, const float* m0, const float* m1, float a, float b, float c,
                int64_t n, void* stream);


/* ------------------------------------------------------------------ slot transition function (TransformerPredictor,
 * video_based/models/predictor.py:20-44: nn.TransformerEncoder over the [B, S, D] slots between two video frames).
 * Multi-head self-attention over S <= 32 slot tokens per sample on the fused in_proj rows qkv [B*S, ld >= 3 D]
 * (q | k | v, head h at columns h*dh of each third; nn.MultiheadAttention), head dims 32 / 48 / 64, with dropout on the
 * attention probabilities (counter-based mask: seed, element index, optional device-resident step counter); out [B*S, D].
 * The backward recomputes the probabilities and the mask: dqkv [B*S, 3 D] (dense). */
int sdb_token_attention_supported(int64_t S, int64_t dh);
int sdb_token_attention(const float* qkv, int64_t ld, float* out, int64_t B, int64_t S, int heads, int dh, float scale,
                        float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);
int sdb_token_attention_bwd(const float* qkv, int64_t ld, const float* dout, float* dqkv, int64_t B, int64_t S, int heads,
                            int dh, float scale, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* out = (res ? res : 0) + dropout(x)  (nn.TransformerEncoderLayer dropout / dropout1 / dropout2 + residual add); the
 * backward of the dropout is the same call on dy with res = NULL */
int sdb_dropout_add(const float* x, const float* res, float* out, int64_t n, float drop_p, uint64_t seed,
                    const uint64_t* seed_dev, void* stream);

/* ================================================================== backward (training) entry points
 * Replace torch autograd of the forward lines cited above (LDM.loss_function, ldm.py:59-83 -> UNet backward;
 * SlotAttention backward through all iterations, slot_attention.py:78-102).  The contractions of the backward pass are
 * sdb_gemm calls (dgrad: pack(dY) x pack_T(W); wgrad: SDB_A_WGRAD / plain on transposed operands, split-K). */

/* dy [M,N] (row stride ld) -> any of: packed rows [2][M][N]; packed transpose [2][N][ldt >= M] (the caller zero-fills
 * the padding columns when ldt > M; ldt % 8 == 0 makes it a valid GEMM operand); bias_grad[N] += column sums;
 * group_grad[row / rows_per_group][n] (row stride ldg) += column sums per group (timestep-embedding gradient). */
int sdb_grad_pack(const float* dy, int64_t ld, void* out_rows, void* out_T, int64_t ldt, float* bias_grad,
                  float* group_grad, int64_t ldg, int64_t M, int64_t N, int rows_per_group, void* stream);
/* packed [2][M][K] -> packed [2][K][ldt >= M]; to_bf16: re-split an fp16 operand as bf16 on the way */
int sdb_transpose_packed(const void* in, void* out, int64_t ldt, int64_t M, int64_t K, int to_bf16, void* stream);
/* fp16-split packed operand (n elements per plane) -> bf16-split, same layout */
int sdb_repack_bf16(const void* in, void* out, int64_t n, void* stream);
/* conv weight [Cout,Cin,3,3] -> dgrad operand packed [2][Cin][9*Cout] (taps rotated by 180 degrees), fp16 or bf16 split */
int sdb_pack_weight_conv3_dgrad(const float* w, void* out, int64_t Cout, int64_t Cin, int bf16, void* stream);
/* wgrad GEMM result c9 [9*Cin][ldc >= Cout] -> dw [Cout][Cin_w][3][3] (= or +=) */
int sdb_wgrad_conv3_scatter(const float* c9, int64_t ldc, float* dw, int64_t Cout, int64_t Cin, int64_t Cin_w,
                            int accumulate, void* stream);
/* out = a + b + c (b, c optional), n % 4 == 0 */
int sdb_add3(float* out, const float* a, const float* b, const float* c, int64_t n, void* stream);
/* dx = dy * act'(pre); act 1 SiLU, 2 ReLU */
int sdb_act_bwd(const float* dy, int64_t ldy, const float* pre, int64_t ldp, float* dx, int64_t ldx, int64_t M, int64_t N,
                int act, void* stream);
/* GroupNorm(+SiLU)(+dropout) backward.  da [B*HW, C1+C2] = gradient w.r.t. the normalised/activated operand;
 * dx1/dx2 = gradients of the two concatenated sources (+ add1/add2 if given); dgamma/dbeta are accumulated (+=).
 * sums_work: fp32 scratch [B][C][2]. */
int sdb_groupnorm_bwd(const float* x1, int64_t C1, const float* x2, int64_t C2, const float* da, const float* stats,
                      const float* gamma, const float* beta, float* sums_work, float* dx1, float* dx2, const float* add1,
                      const float* add2, float* dgamma, float* dbeta, int64_t B, int64_t HW, int G, int silu, float drop_p,
                      uint64_t seed, const uint64_t* seed_dev, void* stream);
/* LayerNorm backward: dx = LN'(x) dn (+ add); dgamma/dbeta accumulated (+=) */
int sdb_layernorm_bwd(const float* x, const float* dn, const float* gamma, float eps, const float* add, float* dx,
                      float* dgamma, float* dbeta, int64_t M, int64_t C, void* stream);
/* attention core backward; work: fp32 scratch of 2*B*heads*Lq floats */
int sdb_attention_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                      const float* dout, int64_t ldo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                      int64_t lddv, float* work, int64_t B, int64_t Lq, int64_t Lk, int heads, int d, float scale,
                      void* stream);
/* GEGLU backward: u [M,2F], dg [M,F] -> du [M,2F] */
int sdb_geglu_bwd(const float* u, const float* dg, float* du, int64_t M, int64_t F, void* stream);
/* adjoint of the nearest x2 upsample: dup [B,2H,2W,C] -> dx [B,H,W,C] */
int sdb_up2_adjoint(const float* dup, float* dx, int64_t B, int64_t H, int64_t W, int64_t C, void* stream);
/* zero insertion dy [B,H,W,C] -> packed [2][B*2H*2W][C] (dgrad of a stride-2 conv = stride-1 conv of it, rotated taps) */
int sdb_pack_zero_up2(const float* dy, void* out, int64_t B, int64_t H, int64_t W, int64_t C, void* stream);
/* NCHW [B,Cs,HW] -> NHWC packed with channels zero-padded to Cp; NHWC <-> NCHW converters for the 3-channel ends */
int sdb_pack_nchw_pad(const float* x, void* out, int64_t B, int64_t Cs, int64_t HW, int64_t Cp, void* stream);
int sdb_nhwc_to_nchw(const float* in, int64_t ld, float* out, int64_t B, int64_t Cs, int64_t HW, void* stream);
int sdb_nchw_to_nhwc_pad(const float* in, float* out, int64_t B, int64_t Cs, int64_t HW, int64_t Cp, void* stream);
/* explicit transposed im2col (feature maps narrower than 8): packed [2][B*H*W][C] -> packed [2][9*C][B*H*W] */
int sdb_im2col_t(const void* in, void* out, int64_t B, int64_t H, int64_t W, int64_t C, void* stream);
/* GRUCell pointwise backward */
int sdb_gru_gates_bwd(const float* gi, const float* gh, const float* h, const float* dh_new, float* dgi, float* dgh,
                      float* dh, int64_t R, int64_t D, void* stream);
/* one Slot-Attention iteration backward: dkv [B,N,2D] (= or +=), dq [B,S,D] (overwritten) */
int sdb_slot_attend_bwd(const float* kv, const float* q, const float* upd, const float* colsum, const float* d_upd,
                        float* dkv, float* dq, int64_t B, int64_t N, int64_t S, int64_t D, float scale, float eps,
                        int accumulate, void* stream);

/* ResNet18-GN encoder block end (resnet.py:72-90, stem :288-291): out = relu(GN(h; stats_h, gamma_h, beta_h) + identity),
 * identity = idn (stats_i NULL), GN(idn; stats_i, gamma_i, beta_i) (the 1x1 `downsample` branch) or 0 (idn NULL).
 * h, idn: NHWC rows [B*HW, C]; stats: [B,G,2] (mean, rstd).  Writes fp32 rows `out` and / or the packed operand. */
int sdb_groupnorm_add_relu(const float* h, const float* stats_h, const float* gamma_h, const float* beta_h,
                           const float* idn, const float* stats_i, const float* gamma_i, const float* beta_i, float* out,
                           void* out_packed, int64_t B, int64_t HW, int64_t C, int G, void* stream);

/* softmax over the last dim of x [M, N] (ldx) * scale, then pack: the P operand of the single-head attention of the
 * VQ-VAE AttnBlock (modules.py:136-139: w_ = softmax(q k^T * C^-0.5)), multiplied by out_scale (a power of two) before
 * the fp16 hi/lo split; N <= 4096, N % 4 == 0 */
int sdb_softmax_pack(const float* x, int64_t ldx, float scale, float out_scale, void* out, int64_t M, int64_t N,
                     void* stream);

/* ------------------------------------------------------------------ boundary fusions (csrc/boundary.cu)
 * q_sample: x_t[b] = sqrt_abar[t[b]] * x0[b] + sqrt_1m_abar[t[b]] * eps[b], rows of n = C*H*W floats (n % 4 == 0); the two
 * products and the sum are rounded separately, so x_t is bit-identical to eager PyTorch (ddpm.py:161-165). */
int sdb_q_sample(const float* x0, const float* eps, const int64_t* t, const float* sqrt_abar, const float* sqrt_1m_abar,
                 float* out, int64_t B, int64_t n, void* stream);
/* F.mse_loss(pred, target) (ldm.py:76-77): loss[0] = mean((pred - target)^2); diff (optional) keeps pred - target for
 * sdb_mse_loss_bwd: dpred = diff * 2/n * grad_out[0]. */
int sdb_mse_loss_fwd(const float* pred, const float* target, float* diff, float* loss, int64_t n, void* stream);
int sdb_mse_loss_bwd(const float* diff, const float* grad_out, float* dpred, int64_t n, void* stream);
/* eval-time slot masks (sa_diffusion.py:172-180, savi_diffusion.py:205-213 + test_seg.py:27): bilinear resize
 * (align_corners = False, ATen's weights and evaluation order) of masks [B,S,h,w] to up [B,S,H,W] (optional) and the
 * per-pixel argmax over slots idx [B,H,W] int64 (optional), first maximum wins. */
int sdb_mask_upsample_argmax(const float* masks, float* up, int64_t* idx, int64_t B, int S, int h, int w, int H, int W,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H_ */
