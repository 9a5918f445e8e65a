"""Tensor-level wrappers over the C ABI (include/sdb200.h).

PyTorch is used only as the owner of device memory and streams: every function takes CUDA tensors,
allocates outputs with torch.empty and enqueues hand-written kernels from libsdb200.so on
torch.cuda.current_stream().  No function here computes anything with torch ops.
"""
import ctypes

import torch

from . import _lib
from ._lib import (SDB_A_CONV3, SDB_A_CONV3S2, SDB_A_CONV3S2A, SDB_A_PLAIN, SDB_PACK_PHASE2, SDB_PACK_PLAIN, SDB_PACK_UP2,
                   SdbGemm, SdbSlotAttentionResident, SdbSlotUpdate, check, lib)

import contextlib
import math
import os

SDB_FMT_F16X2, SDB_FMT_F8C = 0, 2

_PASSES = 3   # 3: hi*hi + lo*hi + hi*lo (fp32-faithful, default); 1: single fp16 pass
_FMT = SDB_FMT_F16X2          # format the activation producers write NOW on the current stream / weights get packed in
# Format of the UNet INFERENCE path (no-grad forward, DPM-Solver sampling): 'fp8c' = SDB_FMT_F8C operands, two
# pass-equivalents per product (fp16 main term + two e4m3 correction terms, sdb200.h), UNet output 2.8e-5 rel-L2 from the
# reference golden (three passes: 3.9e-6; contract 1e-3).  Measured on the B200 (profiles/README 13): the GEMMs are
# shared-memory-bandwidth bound, not tensor bound, so the format buys 6-8 % of GEMM time and 1-3 % of a sampling step --
# it stays OPT-IN (set_precision('fp8c') / SDB_UNET_PRECISION=fp8c); the default is the fp32-faithful three-pass format.
# Training and Slot Attention (bit-exact argmax masks) always use three passes.
_UNET_INFERENCE = os.environ.get('SDB_UNET_PRECISION', 'fp32')


def set_precision(mode):
    """'fp32' (default): 3-pass split-fp16 tensor-core products everywhere, ~2^-22 relative product error.
    'fp8c': the same, except that UNet inference runs SDB_FMT_F8C operands (2 pass-equivalents, ~2^-15).
    'fp16': single pass (hi planes only), ~2^-11 -- the accuracy class of the reference under TF32/AMP."""
    global _PASSES, _UNET_INFERENCE
    if mode not in ('fp32', 'fp16', 'fp8c'):
        raise ValueError(mode)
    _PASSES = 1 if mode == 'fp16' else 3
    _UNET_INFERENCE = 'fp8c' if mode == 'fp8c' else 'fp32'


def get_passes():
    return _PASSES


def precision_key():
    """What a captured graph depends on (sampler graph cache key)."""
    return (_PASSES, _UNET_INFERENCE, _FMT)


@contextlib.contextmanager
def pack_format(fmt):
    """Stream-ordered scope in which every activation producer writes `fmt` and weights are packed in `fmt`
    (sdb_set_pack_mode is a one-thread kernel: CUDA-graph capturable)."""
    global _FMT
    prev = _FMT
    if fmt != prev:
        check(lib().sdb_set_pack_mode(fmt, _stream()), 'sdb_set_pack_mode')
        _FMT = fmt
    try:
        yield
    finally:
        if fmt != prev:
            check(lib().sdb_set_pack_mode(prev, _stream()), 'sdb_set_pack_mode')
            _FMT = prev


_TRAIN_DEPTH = [0]


@contextlib.contextmanager
def training_scope():
    """Marks the forward / backward schedules of the autograd Functions (backward.py, resnet.py): WeightCache.nocache is
    honoured only here (grad mode cannot tell: autograd.Function.forward runs under no_grad like the inference paths)."""
    _TRAIN_DEPTH[0] += 1
    try:
        yield
    finally:
        _TRAIN_DEPTH[0] -= 1


def unet_inference_format():
    return SDB_FMT_F8C if (_UNET_INFERENCE == 'fp8c' and _PASSES == 3) else SDB_FMT_F16X2


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t, name='tensor'):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f'{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}')
    return t


class Packed:
    """GEMM operand: 16-bit [2][rows][K] (hi plane, lo plane); fp16 split, or bf16 split for gradient operands; fmt
    SDB_FMT_F8C: plane 1 holds two e4m3 half-planes (sdb200.h), exponents (eh, el) = (2, 12) for activations,
    (wexp, wexp + 10) for weights."""
    __slots__ = ('t', 'rows', 'K', 'bf16', 'fmt', 'wexp', 'plane', 'off')

    def __init__(self, t, rows, K, bf16=False, fmt=SDB_FMT_F16X2, wexp=None, plane=None, off=0):
        self.t, self.rows, self.K, self.bf16, self.fmt, self.wexp = t, rows, K, bf16, fmt, wexp
        self.plane = rows * K if plane is None else plane      # halves between the hi and lo plane
        self.off = off                                          # first half of the hi plane inside t (row-range views)

    def ptr(self):
        return self.t.data_ptr() + 2 * self.off

    def row_range(self, r0, r1):
        """Rows [r0, r1) as an operand of its own (no copy): same planes, shifted start -- e.g. the tokens of one sample."""
        assert self.fmt == SDB_FMT_F16X2 and 0 <= r0 < r1 <= self.rows
        return Packed(self.t, r1 - r0, self.K, self.bf16, self.fmt, self.wexp, plane=self.plane, off=self.off + r0 * self.K)

    @staticmethod
    def empty(rows, K, device, bf16=False, fmt=None):
        """fmt None: whatever the activation producers write on the current stream (pack_format scope)."""
        if fmt is None:
            fmt = SDB_FMT_F16X2 if bf16 else _FMT
        return Packed(torch.empty(2 * rows * K, dtype=torch.float16, device=device), rows, K, bf16, fmt)

    def unpack(self):
        """fp32 value of the operand (tests only)."""
        if self.fmt == SDB_FMT_F8C:
            n = self.rows * self.K
            hi = self.t[:n].view(self.rows, self.K).float()
            f8 = self.t[n:].view(torch.float8_e4m3fn).view(2, self.rows, self.K).float()
            eh, el = (2, 12) if self.wexp is None else (self.wexp, self.wexp + 10)
            assert torch.allclose(f8[0] * 2.0 ** -eh, hi, rtol=2.0 ** -3, atol=2.0 ** (-9 - eh))    # e4m3 copy of hi
            return hi + f8[1] * 2.0 ** -el
        t = self.t.view(torch.bfloat16) if self.bf16 else self.t
        t = t.view(2, self.rows, self.K).float()
        return t[0] + t[1]


def slot_attention_fold_math(wq, wk, wv, gamma, beta, w_ih, b_ih, scale):
    """fp64 algebra of the Slot-Attention weight fold (pure torch, any device; unit-tested on CPU against the oracle):
    returns (W_qa [Din+4, D], W_iv [3D, Din], b_iv [3D]) such that with n = LayerNorm-without-affine(inputs),
      logits = n @ W_qa[:Din] @ LN_q(slots)^T + W_qa[Din] @ LN_q(slots)^T      ( = scale * k q^T of slot_attention.py:84)
      gi     = U @ W_iv^T + b_iv,  U = (a^T n) / sum_n a                        ( = GRU input projection of updates, :91-97)"""
    wq, wk, wv, g, b, wih, bih = [t.double() for t in (wq, wk, wv, gamma, beta, w_ih, b_ih)]
    Din = wk.shape[1]
    wqk = (wk.t() @ wq) * scale                               # [Din, D]
    wqa = torch.zeros(Din + 4, wq.shape[1], dtype=torch.float64, device=wq.device)
    wqa[:Din] = wqk * g[:, None]
    wqa[Din] = b @ wqk
    wiv = (wih @ wv) * g[None, :]                             # [3D, Din]
    biv = wih @ (wv @ b) + bih
    return wqa, wiv, biv


# ---------------------------------------------------------------- weights (cached per parameter version)
class WeightCache:
    """Packed copies of parameters, re-packed when the parameter storage or version changes
    (an optimizer step bumps `_version`)."""

    def __init__(self):
        self._c = {}
        self.nocache = False        # graphed training (graphed.py): inside a training schedule, pack every time -- the
        #                             packing kernels become part of the captured forward / backward, so a replay after
        #                             optimizer.step() sees the new parameter values

    def _get(self, key, tensors, fn):
        if self.nocache and _TRAIN_DEPTH[0] > 0:
            return fn()
        key = (key, _FMT)           # packed weights exist per operand format
        sig = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
        hit = self._c.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        val = fn()
        self._c[key] = (sig, val)
        return val

    def linear(self, key, *ws):
        """Row-concatenation of nn.Linear / 1x1-conv weights [N_i, K] -> Packed [sum N_i, K]."""
        def make():
            w2 = [w.detach().reshape(w.shape[0], -1) for w in ws]
            w = w2[0] if len(w2) == 1 else torch.cat(w2, 0)
            return pack_weight(w.contiguous())
        return self._get(key, ws, make)

    def geglu(self, key, w, b):
        """(Packed interleaved weight, permuted bias) of a GEGLU projection."""
        return self._get(key, (w, b), lambda: pack_weight_geglu(w.detach().contiguous(), b.detach().contiguous()))

    def conv3(self, key, w):
        return self._get(key, (w,), lambda: pack_weight_conv3(w.detach().contiguous()))

    def slot_attention_fold(self, mod):
        """Parameter-only preprocessing for the tensor-core Slot-Attention path (done once per parameter version, in
        fp64): the k / v projections and the affine part of norm_inputs are folded into the slot-side projection and
        into the GRU input weight (see csrc/slot_attention_fused.cu):
          W_qa [Din+4, D]: rows c < Din = scale * gamma_c * (Wk^T Wq)[c, :], row Din = scale * beta^T Wk^T Wq, rest 0
          W_iv [3D, Din]  = W_ih Wv diag(gamma);   b_iv = W_ih (Wv beta) + b_ih"""
        ps = (mod.project_q[1].weight, mod.project_k.weight, mod.project_v.weight, mod.norm_inputs.weight,
              mod.norm_inputs.bias, mod.gru.weight_ih, mod.gru.bias_ih)

        def make():
            wqa, wiv, biv = slot_attention_fold_math(*[p.detach() for p in ps], mod.attn_scale)
            return dict(w_qa=pack_weight(wqa.float().contiguous()), w_iv=pack_weight(wiv.float().contiguous()),
                        b_iv=biv.float().contiguous())
        return self._get('sa_fold', ps, make)

    def slot_update_weights(self, mod):
        """fp32 TRANSPOSED ([K][N]) weights of the one-launch slot update (csrc/slot_update.cuh): the same fold as
        slot_attention_fold (fp64 algebra, once per parameter version), then input-major copies so that consecutive
        threads of the kernel read consecutive output columns."""
        ps = (mod.project_q[1].weight, mod.project_k.weight, mod.project_v.weight, mod.norm_inputs.weight,
              mod.norm_inputs.bias, mod.gru.weight_ih, mod.gru.bias_ih)
        rest = (mod.gru.weight_hh, mod.gru.bias_hh, mod.mlp[0].weight, mod.mlp[0].bias, mod.mlp[1].weight,
                mod.mlp[1].bias, mod.mlp[3].weight, mod.mlp[3].bias, mod.project_q[0].weight, mod.project_q[0].bias)

        def make():
            wqa, wiv, biv = slot_attention_fold_math(*[p.detach() for p in ps], mod.attn_scale)
            whh, bhh, gm, bm, w1, b1, w2, b2, gq, bq = [p.detach().float() for p in rest]

            def T(w):
                return w.float().t().contiguous()
            return dict(w_ivT=T(wiv), b_iv=biv.float().contiguous(), w_hhT=T(whh), b_hh=bhh.contiguous(),
                        ln_m_g=gm.contiguous(), ln_m_b=bm.contiguous(), w1T=T(w1), b1=b1.contiguous(), w2T=T(w2),
                        b2=b2.contiguous(), ln_q_g=gq.contiguous(), ln_q_b=bq.contiguous(), w_qaT=T(wqa),
                        ldq=int(wqa.shape[0]), ln_m_eps=float(mod.mlp[0].eps), ln_q_eps=float(mod.project_q[0].eps))
        return self._get('sa_update', ps + rest, make)

    def slot_resident_weights(self, mod):
        """Weights of the persistent Slot-Attention kernel (csrc/slot_attention_resident.cu): the fold of
        slot_update_weights, each matrix additionally k-quad interleaved -- w4[K/4][ncols][4] -- so that a thread reads
        four consecutive K elements of its output column with one 16-byte load and a warp reads 128 contiguous bytes."""
        base = self.slot_update_weights(mod)

        def make():
            def k4(wT):                               # [K][N] -> [K/4][N][4]
                K, N = wT.shape
                return wT.reshape(K // 4, 4, N).permute(0, 2, 1).contiguous()
            out = dict(base)
            for src, dst in (('w_ivT', 'w_iv4'), ('w_hhT', 'w_hh4'), ('w1T', 'w1_4'), ('w2T', 'w2_4'), ('w_qaT', 'w_qa4')):
                out[dst] = k4(base[src])
            return out
        return self._get('sa_resident', tuple(base[k] for k in ('w_ivT', 'w_hhT', 'w1T', 'w2T', 'w_qaT')), make)

    def cat(self, key, *vs):
        """Concatenation of fp32 vectors (fused biases)."""
        return self._get(key, vs, lambda: torch.cat([v.detach().reshape(-1) for v in vs]).contiguous())

    def clear(self):
        self._c.clear()


def weight_exponent(w):
    """Per-tensor exponent of an SDB_FMT_F8C weight: the largest b with max|w| * 2^b <= 448 (e4m3 max).  Parameter
    preprocessing, once per parameter version (host read of one scalar)."""
    m = float(w.detach().abs().max())
    if not math.isfinite(m) or m <= 0.0:
        return 0
    return max(-24, min(40, int(math.floor(math.log2(448.0 / m)))))


def pack_weight(w):
    _f32(w, 'weight')
    N, K = w.shape
    out = Packed.empty(N, K, w.device)
    if out.fmt == SDB_FMT_F8C:
        out.wexp = weight_exponent(w)
        check(lib().sdb_pack_weight_fmt(_p(w), _p(out.t), N, K, SDB_FMT_F8C, out.wexp, _stream()), 'sdb_pack_weight_fmt')
        return out
    check(lib().sdb_pack_weight(_p(w), _p(out.t), N, K, _stream()), 'sdb_pack_weight')
    return out


def pack_weight_conv3(w):
    _f32(w, 'weight')
    Cout, Cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    out = Packed.empty(Cout, 9 * Cin, w.device)
    if out.fmt == SDB_FMT_F8C:
        out.wexp = weight_exponent(w)
        check(lib().sdb_pack_weight_conv3_fmt(_p(w), _p(out.t), Cout, Cin, SDB_FMT_F8C, out.wexp, _stream()),
              'sdb_pack_weight_conv3_fmt')
        return out
    check(lib().sdb_pack_weight_conv3(_p(w), _p(out.t), Cout, Cin, _stream()), 'sdb_pack_weight_conv3')
    return out


# ---------------------------------------------------------------- operand producers
def pack_rows(x, act=0):
    """x [M,K] fp32 (last dim contiguous) -> Packed; act 0 none / 1 SiLU / 2 ReLU."""
    _f32(x)
    assert x.dim() == 2 and x.stride(1) == 1
    M, K = x.shape
    out = Packed.empty(M, K, x.device)
    check(lib().sdb_pack_rows(_p(x), x.stride(0), _p(out.t), M, K, act, _stream()), 'sdb_pack_rows')
    return out


def groupnorm_add_relu(h, stats_h, gn_h, B, HW, idn=None, stats_i=None, gn_i=None, want_f32=True, want_packed=True):
    """relu(GN(h) + identity) -> (fp32 rows | None, Packed | None); identity = idn, GN_i(idn) or nothing."""
    C = h.shape[-1]
    out = torch.empty_like(h) if want_f32 else None
    pk = Packed.empty(B * HW, C, h.device) if want_packed else None
    check(lib().sdb_groupnorm_add_relu(_p(h), _p(stats_h), _p(gn_h.weight), _p(gn_h.bias), _p(idn), _p(stats_i),
                                       _p(gn_i.weight) if gn_i is not None else None,
                                       _p(gn_i.bias) if gn_i is not None else None, _p(out),
                                       _p(pk.t) if pk is not None else None, B, HW, C, gn_h.num_groups, _stream()),
          'sdb_groupnorm_add_relu')
    return out, pk


SOFTMAX_PACK_SCALE = 4096.0     # probabilities of ~1e-3 would put their fp16 lo plane (2^-11 of the value) into subnormals


def softmax_pack(x, scale=1.0, out_scale=SOFTMAX_PACK_SCALE):
    """out_scale * softmax(x * scale) over the last dim of x [M, N] (strided rows ok) -> Packed [M, N]; the consumer GEMM
    undoes the power-of-two out_scale with alpha = 1 / out_scale."""
    _f32(x)
    assert x.dim() == 2 and x.stride(1) == 1
    M, N = x.shape
    out = Packed.empty(M, N, x.device)
    check(lib().sdb_softmax_pack(_p(x), x.stride(0), float(scale), float(out_scale), _p(out.t), M, N, _stream()),
          'sdb_softmax_pack')
    return out


def layernorm_pack(x, gamma, beta, eps=1e-5, want_fp32=False):
    _f32(x)
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    assert x2.is_contiguous()
    M = x2.shape[0]
    out = Packed.empty(M, C, x.device)
    y = torch.empty_like(x2) if want_fp32 else None
    check(lib().sdb_layernorm_pack(_p(x2), _p(gamma), _p(beta), eps, _p(out.t), _p(y), M, C, _stream()),
          'sdb_layernorm_pack')
    return (out, y) if want_fp32 else out


def groupnorm_stats(x1, x2, B, HW, G, eps):
    """x1 [B*HW, C1] (+ x2 [B*HW, C2]) NHWC rows -> stats [B, G, 2] (mean, rstd)."""
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    stats = torch.empty(B, G, 2, dtype=torch.float32, device=x1.device)
    check(lib().sdb_groupnorm_stats(_p(x1), C1, _p(x2), C2, _p(stats), B, HW, G, eps, _stream()),
          'sdb_groupnorm_stats')
    return stats


def groupnorm_pack(x1, x2, gamma, beta, B, HW, G=32, eps=1e-5, silu=True, stats=None):
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    if stats is None:
        stats = groupnorm_stats(x1, x2, B, HW, G, eps)
    out = Packed.empty(B * HW, C1 + C2, x1.device)
    check(lib().sdb_groupnorm_apply_pack(_p(x1), C1, _p(x2), C2, _p(stats), _p(gamma), _p(beta), _p(out.t), B, HW,
                                         G, int(silu), _stream()), 'sdb_groupnorm_apply_pack')
    return out


def pack_nhwc(x1, x2, B, H, W, mode=SDB_PACK_PLAIN, want_cat=False):
    """raw NHWC rows (+ channel concat) -> Packed in plain / nearest-x2 / stride-2 phase-split layout."""
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    mult = 4 if mode == SDB_PACK_UP2 else 1
    out = Packed.empty(B * H * W * mult, C1 + C2, x1.device)
    ycat = torch.empty(B * H * W, C1 + C2, dtype=torch.float32, device=x1.device) if want_cat else None
    check(lib().sdb_pack_nhwc(_p(x1), C1, _p(x2), C2, _p(out.t), _p(ycat), B, H, W, mode, _stream()), 'sdb_pack_nhwc')
    return (out, ycat) if want_cat else out


def geglu_pack(u):
    M, F2 = u.shape
    out = Packed.empty(M, F2 // 2, u.device)
    check(lib().sdb_geglu_pack(_p(u), _p(out.t), M, F2 // 2, _stream()), 'sdb_geglu_pack')
    return out


def timestep_embedding_pack(t, dim):
    t = t.to(torch.float32).contiguous()
    B = t.shape[0]
    out = Packed.empty(B, dim, t.device)
    check(lib().sdb_timestep_embedding_pack(_p(t), _p(out.t), B, dim, _stream()), 'sdb_timestep_embedding_pack')
    return out


# ---------------------------------------------------------------- GEMM
_ACTS = {None: 0, 'none': 0, 'silu': 1, 'relu': 2}


def gemm(a, w, bias=None, rowvec=None, rows_per_group=0, residual=None, relu=False, conv=None, passes=None,
         out=None, pack_out=None, keep_c=True, gsum=None, geglu=False, gsum_cb=4, batch=None, alpha=1.0):
    """C = A W^T (+bias)(+rowvec[row // rows_per_group])(+residual); A, W Packed.
    conv: None (plain) or (mode, B, H, W, C) with H, W the OUTPUT size (see sdb200.h).
    pack_out: None, or 'none' / 'silu' / 'relu' -- the epilogue ALSO emits act(C) as a Packed operand for the next
              GEMM; returns (C or None, Packed); keep_c=False drops the fp32 copy.
    gsum:     fp32 [M // rows_per_group, N // 4, 2] zero-initialised buffer that receives GroupNorm partial sums of C.
    geglu:    W/bias packed by pack_weight_geglu; returns Packed [M, N/2] = a * gelu(g).
    gsum_cb:  channels per partial-sum block of gsum (4, or 2: buffer [M // rows_per_group, N // 2, 2]).
    batch:    (batch_rows, N, w_row_step, w_k_step): block-diagonal product in one launch -- rows [b*batch_rows, ...) of A
              meet the [N, a.K] block of the packed tensor w at row b*w_row_step, column b*w_k_step (sdb200.h)."""
    N, K = w.rows, w.K
    if batch is not None:
        N, K = batch[1], a.K
    if conv is None:
        M, mode, geo = a.rows, SDB_A_PLAIN, (0, 0, 0, 0)
        assert a.K == K, (a.K, K)
    else:
        mode, B, H, W, C = conv
        M, geo = B * H * W, (B, H, W, C)
        assert K == 9 * C and a.K == C
    dev = a.t.device
    packed = None
    if geglu:
        packed = Packed.empty(M, N // 2, dev)
        keep_c = False
    elif pack_out is not None:
        packed = Packed.empty(M, N, dev)
    if keep_c and out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
    g = SdbGemm()
    g.a, g.w = a.ptr(), w.ptr()
    g.w_plane_stride = 0 if w.plane == w.rows * w.K else w.plane
    g.c = out.data_ptr() if keep_c else None
    g.bias = bias.data_ptr() if bias is not None else None
    g.rowvec = rowvec.data_ptr() if rowvec is not None else None
    g.residual = residual.data_ptr() if residual is not None else None
    g.a_plane_stride = a.plane
    g.ldc = out.stride(0) if keep_c else N
    g.ldv = rowvec.stride(0) if rowvec is not None else 0
    g.ldr = residual.stride(0) if residual is not None else 0
    g.M, g.N, g.K, g.mode = M, N, K, mode
    g.B, g.H, g.W, g.C = geo
    g.rows_per_group = rows_per_group
    g.passes = passes or _PASSES
    if a.fmt == SDB_FMT_F8C or w.fmt == SDB_FMT_F8C:
        if not (a.fmt == SDB_FMT_F8C and w.fmt == SDB_FMT_F8C and w.wexp is not None):
            raise RuntimeError('sdb_gemm: mixed operand formats (activation fmt %d, weight fmt %d)' % (a.fmt, w.fmt))
        g.passes = 2
        g.corr_scale = 2.0 ** -(12 + w.wexp)
    g.relu = int(relu)
    if packed is not None:
        g.out_packed = packed.t.data_ptr()
        g.out_plane_stride = packed.rows * packed.K
        g.out_act = _ACTS[pack_out]
    g.gsum = gsum.data_ptr() if gsum is not None else None
    g.gsum_cb = gsum_cb
    g.alpha = float(alpha)
    if batch is not None:
        g.batch_rows, g.w_row_step, g.w_k_step = batch[0], batch[2], batch[3]
        g.w_rows, g.w_cols = w.rows, w.K
    g.geglu = int(geglu)
    g.a_bf16, g.w_bf16 = int(a.bf16), int(w.bf16)
    check(lib().sdb_gemm(ctypes.byref(g), _stream()), 'sdb_gemm')
    if geglu:
        return packed
    if pack_out is not None:
        return (out if keep_c else None), packed
    return out


def pack_weight_geglu(w, bias):
    """GEGLU.proj weight [2F, K] (+bias [2F]) -> (Packed interleaved [16 a | 16 g], permuted bias)."""
    _f32(w, 'weight')
    F2, K = w.shape
    out = Packed.empty(F2, K, w.device)
    bout = torch.empty_like(bias) if bias is not None else None
    check(lib().sdb_pack_weight_geglu(_p(w), _p(bias), _p(out.t), _p(bout), F2 // 2, K, _stream()),
          'sdb_pack_weight_geglu')
    if out.fmt == SDB_FMT_F8C:
        # the interleave [16 a | 16 g] is a row permutation: applied here (parameter preprocessing), then the generic
        # per-tensor-exponent packer
        F = F2 // 2
        wp = torch.stack([w[:F].view(F // 16, 16, K), w[F:].view(F // 16, 16, K)], 1).reshape(F2, K).contiguous()
        out.wexp = weight_exponent(w)
        check(lib().sdb_pack_weight_fmt(_p(wp), _p(out.t), F2, K, SDB_FMT_F8C, out.wexp, _stream()), 'sdb_pack_weight_fmt')
    return out, bout


def groupnorm_pack_fused(x1, x2, gamma, beta, B, HW, G=32, eps=1e-5, silu=True, gsum1=None, gsum2=None, stats=None):
    """GroupNorm(+SiLU)+pack with statistics from the producers' partial sums (gsum*) or from `stats` [B,G,2]."""
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    out = Packed.empty(B * HW, C1 + C2, x1.device)
    check(lib().sdb_groupnorm_apply_pack_fused(_p(x1), C1, _p(gsum1), _p(x2), C2, _p(gsum2), _p(stats), _p(gamma),
                                               _p(beta), _p(out.t), B, HW, G, eps, int(silu), _stream()),
          'sdb_groupnorm_apply_pack_fused')
    return out


def channel_block_sums(x, gsum, B, HW):
    """Accumulate GroupNorm partial sums [B, C/4, 2] of the NHWC rows x [B*HW, C] into the (zeroed) buffer gsum."""
    C = x.shape[-1]
    check(lib().sdb_channel_block_sums(_p(x), C, _p(gsum), B, HW, _stream()), 'sdb_channel_block_sums')
    return gsum


def groupnorm_finalize(gsum1, gsum2, C1, C2, B, HW, G, eps):
    stats = torch.empty(B, G, 2, dtype=torch.float32, device=gsum1.device)
    check(lib().sdb_groupnorm_finalize(_p(gsum1), C1, _p(gsum2), C2, _p(stats), B, HW, G, eps, _stream()),
          'sdb_groupnorm_finalize')
    return stats


def groupnorm_finalize_cb(gsum, C, B, HW, G, eps, cb):
    """partial sums in blocks of cb channels (sdb_gemm gsum_cb) -> stats [B, G, 2] (mean, rstd)"""
    stats = torch.empty(B, G, 2, dtype=torch.float32, device=gsum.device)
    check(lib().sdb_groupnorm_finalize_cb(_p(gsum), C, _p(stats), B, HW, G, eps, cb, _stream()), 'sdb_groupnorm_finalize_cb')
    return stats


# ---------------------------------------------------------------- attention / convs / slot attention / sampler
ATTENTION_TC = True   # tensor-core attention cores (csrc/attention_tc.cu) when the shape is supported


ATTENTION_FEWKEYS = os.environ.get('SDB_ATTENTION_FEWKEYS', '1') == '1'


def attention_pack(q, k, v, B, Lq, Lk, heads, d, scale, tc=None):
    """q/k/v: 2-D strided views [B*L, heads*d] (last dim contiguous) -> Packed [B*Lq, heads*d]."""
    out = Packed.empty(B * Lq, heads * d, q.device)
    use_tc = ATTENTION_TC if tc is None else tc
    if tc is None and ATTENTION_FEWKEYS and _FMT == SDB_FMT_F16X2 and lib().sdb_attention_fewkeys_supported(
            heads, d, Lk, q.stride(0), k.stride(0), v.stride(0)):
        # slot cross-attention (Lk = num_slots): coalesced fp32 kernel, see csrc/attention_fewkeys.cu
        check(lib().sdb_attention_fewkeys(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out.t), B, Lq, Lk,
                                          heads, d, scale, _stream()), 'sdb_attention_fewkeys')
        return out
    if use_tc and lib().sdb_attention_tc_supported(heads, d, q.stride(0), k.stride(0), v.stride(0)):
        check(lib().sdb_attention_tc(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out.t), B, Lq, Lk,
                                     heads, d, scale, _stream()), 'sdb_attention_tc')
        return out
    check(lib().sdb_attention_pack(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out.t), B, Lq, Lk,
                                   heads, d, scale, _stream()), 'sdb_attention_pack')
    return out


def conv3_in(x, w, bias):
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    y = torch.empty(B * H * W, Cout, dtype=torch.float32, device=x.device)
    check(lib().sdb_conv3_in(_p(x.contiguous()), _p(w), _p(bias), _p(y), B, Cin, H, W, Cout, _stream()), 'sdb_conv3_in')
    return y


def conv3_out(h, stats, gamma, beta, w, bias, B, H, W, G=32):
    C = h.shape[-1]
    Cout = w.shape[0]
    y = torch.empty(B, Cout, H, W, dtype=torch.float32, device=h.device)
    check(lib().sdb_conv3_out(_p(h), _p(stats), _p(gamma), _p(beta), _p(w), _p(bias), _p(y), B, H, W, C, G, Cout,
                              _stream()), 'sdb_conv3_out')
    return y


def slot_attend(kv, q, B, N, S, D, scale, eps, want_mask, want_fp32=False):
    ws = lib().sdb_slot_attend_workspace(B, N, S, D)
    work = torch.empty(ws // 4, dtype=torch.float32, device=kv.device)
    mask = torch.empty(B, S, N, dtype=torch.float32, device=kv.device) if want_mask else None
    upd = Packed.empty(B * S, D, kv.device)
    upd32 = torch.empty(B * S, D, dtype=torch.float32, device=kv.device) if want_fp32 else None
    check(lib().sdb_slot_attend(_p(kv), _p(q), _p(mask), _p(upd.t), _p(upd32), _p(work), B, N, S, D, scale, eps,
                                _stream()), 'sdb_slot_attend')
    return upd, mask, upd32


def slot_attend_fused_supported(S, Din):
    return bool(lib().sdb_slot_attend_fused_supported(S, Din))


def slot_attend_fused(x, qa, B, N, S, Din, ln_eps, eps, want_mask, want_fp32=False):
    """Tensor-core Slot-Attention iteration on the RAW features x [B,N,Din] (LayerNorm folded in-kernel);
    qa [B*S, ldq] = folded slot-side projection (logit weights | logit bias).  Returns (U Packed [B*S,Din], mask, U fp32)."""
    _f32(x), _f32(qa)
    assert x.is_contiguous() and qa.stride(1) == 1
    ws = lib().sdb_slot_attend_fused_workspace(B, N, S, Din)
    work = torch.empty(ws // 4, dtype=torch.float32, device=x.device)
    mask = torch.empty(B, S, N, dtype=torch.float32, device=x.device) if want_mask else None
    upd = Packed.empty(B * S, Din, x.device)
    upd32 = torch.empty(B * S, Din, dtype=torch.float32, device=x.device) if want_fp32 else None
    check(lib().sdb_slot_attend_fused(_p(x), _p(qa), qa.stride(0), _p(mask), _p(upd.t), _p(upd32), _p(work), B, N, S,
                                      Din, ln_eps, eps, _stream()), 'sdb_slot_attend_fused')
    return upd, mask, upd32


def slot_attend_fused_partials(x, qa, B, N, S, Din, ln_eps, eps, want_mask):
    """The attend kernel without its finalize launch.  Returns ((part_upd, part_cs, chunks, ascale), mask): the
    per-chunk partial sums that sdb_slot_update consumes."""
    _f32(x), _f32(qa)
    assert x.is_contiguous() and qa.stride(1) == 1
    ws = lib().sdb_slot_attend_fused_workspace(B, N, S, Din)
    chunks = int(lib().sdb_slot_attend_fused_chunks(B, N))
    work = torch.empty(ws // 4, dtype=torch.float32, device=x.device)
    mask = torch.empty(B, S, N, dtype=torch.float32, device=x.device) if want_mask else None
    check(lib().sdb_slot_attend_fused_partials(_p(x), _p(qa), qa.stride(0), _p(mask), _p(work), B, N, S, Din, ln_eps,
                                               eps, _stream()), 'sdb_slot_attend_fused_partials')
    n_upd = B * chunks * S * Din
    return (work[:n_upd], work[n_upd:n_upd + B * chunks * S], chunks, float(lib().sdb_slot_attend_fused_ascale())), mask


def slot_update_supported(S, Din, D, M):
    return bool(lib().sdb_slot_update_supported(S, Din, D, M))


def slot_update_args(w, parts, slots_in, slots_out, qa_out, S, Din, D, M):
    """SdbSlotUpdate for one call (device-agnostic: only data_ptr()s; the CPU emulation test fills the same struct).
    parts = (part_upd, part_cs, chunks, ascale) or None for the projection of the initial slots."""
    def ptr(t):
        return t.data_ptr() if t is not None else None
    a = SdbSlotUpdate()
    if parts is not None:
        a.part_upd, a.part_cs, a.chunks, a.ascale = ptr(parts[0]), ptr(parts[1]), int(parts[2]), float(parts[3])
    for k in ('w_ivT', 'b_iv', 'w_hhT', 'b_hh', 'ln_m_g', 'ln_m_b', 'w1T', 'b1', 'w2T', 'b2', 'ln_q_g', 'ln_q_b',
              'w_qaT'):
        setattr(a, k, ptr(w[k]))
    a.slots_in, a.slots_out, a.qa_out = ptr(slots_in), ptr(slots_out), ptr(qa_out)
    a.rows, a.S, a.Din, a.D, a.M, a.ldq = slots_in.shape[0], S, Din, D, M, w['ldq']
    a.ln_m_eps, a.ln_q_eps = w['ln_m_eps'], w['ln_q_eps']
    a.do_update = 1 if parts is not None else 0
    return a


def slot_update(w, parts, slots_in, S, Din, D, M, want_q):
    """One launch: (optional) GRU + MLP update of slots_in [B*S, D] from the attend partials, and (optional) the
    folded slot-side projection qa [B*S, ldq] of the result.  Returns (slots_out or slots_in, qa or None)."""
    _f32(slots_in)
    rows = slots_in.shape[0]
    slots_out = torch.empty_like(slots_in) if parts is not None else None
    qa = torch.empty(rows, w['ldq'], dtype=torch.float32, device=slots_in.device) if want_q else None
    a = slot_update_args(w, parts, slots_in, slots_out, qa, S, Din, D, M)
    check(lib().sdb_slot_update(ctypes.byref(a), _stream()), 'sdb_slot_update')
    return (slots_out if parts is not None else slots_in), qa


def slot_attention_resident_supported(N, S, Din, D, M):
    return bool(lib().sdb_slot_attention_resident_supported(N, S, Din, D, M))


_RESIDENT_WAVE = {}


def slot_attention_resident_wave(N, S, Din, D, M):
    """Samples the persistent kernel processes at the same time on this device (0: geometry unsupported)."""
    key = (torch.cuda.current_device(), N, S, Din, D, M)
    if key not in _RESIDENT_WAVE:
        _RESIDENT_WAVE[key] = int(lib().sdb_slot_attention_resident_wave(N, S, Din, D, M))
    return _RESIDENT_WAVE[key]


def slot_attention_resident_args(w, x, slots_in, slots_out, mask, iterations, ln_in_eps, attn_eps, M):
    """SdbSlotAttentionResident for one call (only data_ptr()s and shapes)."""
    B, N, Din = x.shape
    S, D = slots_in.shape[1], slots_in.shape[2]
    a = SdbSlotAttentionResident()
    for k in ('w_iv4', 'b_iv', 'w_hh4', 'b_hh', 'ln_m_g', 'ln_m_b', 'w1_4', 'b1', 'w2_4', 'b2', 'ln_q_g', 'ln_q_b', 'w_qa4'):
        setattr(a, k, w[k].data_ptr())
    a.x, a.slots_in, a.slots_out = x.data_ptr(), slots_in.data_ptr(), slots_out.data_ptr()
    a.seg_mask = mask.data_ptr() if mask is not None else None
    a.B, a.N, a.S, a.Din, a.D, a.M, a.ldq, a.iterations = B, N, S, Din, D, M, w['ldq'], iterations
    a.ln_in_eps, a.attn_eps, a.ln_m_eps, a.ln_q_eps = ln_in_eps, attn_eps, w['ln_m_eps'], w['ln_q_eps']
    return a


def slot_attention_resident(w, x, slots_in, iterations, ln_in_eps, attn_eps, M, want_mask):
    """The whole Slot-Attention forward in one launch: x [B,N,Din] raw features, slots_in [B,S,D] ->
    (slots [B,S,D], seg mask [B,S,N] or None)."""
    _f32(x)
    _f32(slots_in)
    B, N, _ = x.shape
    S = slots_in.shape[1]
    slots_out = torch.empty_like(slots_in)
    mask = torch.empty(B, S, N, dtype=torch.float32, device=x.device) if want_mask else None
    a = slot_attention_resident_args(w, x, slots_in, slots_out, mask, iterations, ln_in_eps, attn_eps, M)
    check(lib().sdb_slot_attention_resident(ctypes.byref(a), _stream()), 'sdb_slot_attention_resident')
    return slots_out, mask


def token_attention_supported(S, dh):
    return bool(lib().sdb_token_attention_supported(S, dh))


def token_attention(qkv, B, S, heads, drop_p=0.0, seed=0):
    """Multi-head self-attention over S slot tokens per sample on the fused in_proj rows qkv [B*S, 3D] -> [B*S, D]."""
    _f32(qkv)
    D = qkv.shape[1] // 3
    out = torch.empty(B * S, D, dtype=torch.float32, device=qkv.device)
    dh = D // heads
    check(lib().sdb_token_attention(_p(qkv), qkv.stride(0), _p(out), B, S, heads, dh, dh ** -0.5, float(drop_p), int(seed),
                                    _p(dropout_step_counter(qkv.device)) if drop_p > 0 else None, _stream()),
          'sdb_token_attention')
    return out


def token_attention_bwd(qkv, dout, B, S, heads, drop_p=0.0, seed=0):
    D = qkv.shape[1] // 3
    dqkv = torch.empty(B * S, 3 * D, dtype=torch.float32, device=qkv.device)
    dh = D // heads
    assert dout.is_contiguous()
    check(lib().sdb_token_attention_bwd(_p(qkv), qkv.stride(0), _p(dout), _p(dqkv), B, S, heads, dh, dh ** -0.5,
                                        float(drop_p), int(seed),
                                        _p(dropout_step_counter(qkv.device)) if drop_p > 0 else None, _stream()),
          'sdb_token_attention_bwd')
    return dqkv


def dropout_add(x, res, drop_p, seed):
    """res + dropout(x) (res may be None); counter-based mask (seed, element index, device step counter)."""
    assert x.is_contiguous() and (res is None or res.is_contiguous())
    out = torch.empty_like(x)
    check(lib().sdb_dropout_add(_p(x), _p(res), _p(out), x.numel(), float(drop_p), int(seed),
                                _p(dropout_step_counter(x.device)), _stream()), 'sdb_dropout_add')
    return out


def gru_gates(gi, gh, h):
    R, D = h.shape
    out = torch.empty_like(h)
    check(lib().sdb_gru_gates(_p(gi), _p(gh), _p(h), _p(out), R, D, _stream()), 'sdb_gru_gates')
    return out


def dpm_x0(x, eps, alpha, sigma, codebook=None, want_idx=False):
    B, C = x.shape[:2]
    HW = x[0, 0].numel()
    x0 = torch.empty_like(x)
    idx = torch.empty(B, HW, dtype=torch.int32, device=x.device) if want_idx else None
    nc = codebook.shape[0] if codebook is not None else 0
    check(lib().sdb_dpm_x0(_p(x), _p(eps), float(alpha), float(sigma), _p(codebook), nc, _p(x0), _p(idx), B, C, HW,
                           _stream()), 'sdb_dpm_x0')
    return (x0, idx) if want_idx else x0


def lincomb(x, m0, m1, a, b, c=0.0, out=None):
    if out is None:
        out = torch.empty_like(x)
    check(lib().sdb_lincomb(_p(out), _p(x), _p(m0), _p(m1), float(a), float(b), float(c), x.numel(), _stream()),
          'sdb_lincomb')
    return out


# ================================================================== backward (training) wrappers
from ._lib import SDB_A_WGRAD, SDB_A_WGRAD_S2  # noqa: E402


def grad_pack(dy, want_rows=True, want_T=True, bias_grad=None, group_grad=None, rows_per_group=0):
    """dy [M,N] fp32 (last dim contiguous) -> (Packed [M,N] | None, Packed [N,M] | None); bias_grad[N] += colsum(dy);
    group_grad[M // rows_per_group, N] += per-group column sums."""
    assert dy.dim() == 2 and dy.stride(1) == 1
    M, N = dy.shape
    rows = Packed.empty(M, N, dy.device, bf16=True) if want_rows else None
    tr = _packed_T(N, M, dy.device, bf16=True) if want_T else None
    check(lib().sdb_grad_pack(_p(dy), dy.stride(0), _p(rows.t) if rows else None, _p(tr.t) if tr else None,
                              tr.K if tr else 0, _p(bias_grad), _p(group_grad),
                              group_grad.stride(0) if group_grad is not None else 0, M, N, rows_per_group, _stream()),
          'sdb_grad_pack')
    return rows, tr


def _packed_T(rows, M, device, bf16=False):
    """Transposed operand [rows, M]: the contraction length is padded to a multiple of 8 (16-byte TMA rows); padding
    columns are zero in BOTH operands of the wgrad GEMM, so they contribute nothing."""
    Kp = (M + 7) // 8 * 8
    if Kp == M:
        return Packed.empty(rows, M, device, bf16)
    return Packed(torch.zeros(2 * rows * Kp, dtype=torch.float16, device=device), rows, Kp, bf16)


def transpose_packed(a, to_bf16=False):
    """Packed [M,K] -> Packed [K, M(+pad)]; to_bf16 re-splits an fp16 operand as bf16 (backward GEMMs run bf16 x bf16)."""
    conv = to_bf16 and not a.bf16
    out = _packed_T(a.K, a.rows, a.t.device, a.bf16 or to_bf16)
    check(lib().sdb_transpose_packed(_p(a.t), _p(out.t), out.K, a.rows, a.K, int(conv), _stream()),
          'sdb_transpose_packed')
    return out


def repack_bf16(a):
    """fp16-split Packed -> bf16-split Packed (same shape)."""
    if a.bf16:
        return a
    out = Packed.empty(a.rows, a.K, a.t.device, bf16=True)
    check(lib().sdb_repack_bf16(_p(a.t), _p(out.t), a.rows * a.K, _stream()), 'sdb_repack_bf16')
    return out


def pack_weight_T(w):
    """W [N,K] fp32 -> Packed [K,N] bf16 split (operand of dX = dY W next to the bf16 gradient operand)."""
    w2 = w.reshape(w.shape[0], -1).contiguous()
    return transpose_packed(pack_weight(w2), to_bf16=True)


def pack_weight_conv3_dgrad(w):
    Cout, Cin = w.shape[0], w.shape[1]
    out = Packed.empty(Cin, 9 * Cout, w.device, bf16=True)
    check(lib().sdb_pack_weight_conv3_dgrad(_p(w), _p(out.t), Cout, Cin, 1, _stream()), 'sdb_pack_weight_conv3_dgrad')
    return out


def wgrad_conv3_scatter(c9, dw, Cin, accumulate=False):
    Cout, Cin_w = dw.shape[0], dw.shape[1]
    check(lib().sdb_wgrad_conv3_scatter(_p(c9), c9.stride(0), _p(dw), Cout, Cin, Cin_w, int(accumulate), _stream()),
          'sdb_wgrad_conv3_scatter')


def gemm_wgrad_conv(x, dy, B, H, W, C, stride2=False, passes=None):
    """Conv weight gradient: x Packed NHWC activation [B*H_in*W_in, C] (the operand the forward conv consumed; the phase
    split for stride 2), dy Packed rows [B*H*W, Cout] -> c9 fp32 [9*C, Cout] (row = tap*C + ci).  H, W = OUTPUT size."""
    Mpix, Cout = dy.rows, dy.K
    assert Mpix == B * H * W and x.K == C
    if dy.bf16 and not x.bf16:
        x = repack_bf16(x)
    out = torch.empty(9 * C, Cout, dtype=torch.float32, device=dy.t.device)
    g = SdbGemm()
    g.a, g.w, g.c = x.t.data_ptr(), dy.t.data_ptr(), out.data_ptr()
    g.a_plane_stride = x.rows * x.K
    g.ldc = Cout
    g.M, g.N, g.K = 9 * C, Cout, Mpix
    g.mode = SDB_A_WGRAD_S2 if stride2 else SDB_A_WGRAD
    g.B, g.H, g.W, g.C = B, H, W, C
    g.passes = passes or _PASSES
    g.a_bf16, g.w_bf16 = int(x.bf16), int(dy.bf16)
    check(lib().sdb_gemm(ctypes.byref(g), _stream()), 'sdb_gemm(wgrad)')
    return out


def add3(a, b=None, c=None, out=None):
    if out is None:
        out = torch.empty_like(a)
    assert a.is_contiguous() and (b is None or b.is_contiguous()) and (c is None or c.is_contiguous())
    check(lib().sdb_add3(_p(out), _p(a), _p(b), _p(c), a.numel(), _stream()), 'sdb_add3')
    return out


def act_bwd(dy, pre, act):
    """dx = dy * act'(pre); dy, pre 2-D (strided rows ok); act 'silu' | 'relu'."""
    M, N = dy.shape
    dx = torch.empty(M, N, dtype=torch.float32, device=dy.device)
    check(lib().sdb_act_bwd(_p(dy), dy.stride(0), _p(pre), pre.stride(0), _p(dx), N, M, N, _ACTS[act], _stream()),
          'sdb_act_bwd')
    return dx


_dropout_step = {}


def dropout_step_counter(device):
    """Device-resident int64 step counter mixed into every dropout seed.  A training loop that replays a captured CUDA
    graph advances it inside the graph (`dropout_step_counter(dev).add_(1)`) to get a fresh mask per replay."""
    t = _dropout_step.get(device)
    if t is None:
        t = _dropout_step[device] = torch.zeros(1, dtype=torch.int64, device=device)
    return t


def groupnorm_pack_dropout(x1, x2, gamma, beta, stats, B, HW, G, silu, drop_p, seed):
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    out = Packed.empty(B * HW, C1 + C2, x1.device)
    check(lib().sdb_groupnorm_apply_pack_dropout(_p(x1), C1, _p(x2), C2, _p(stats), _p(gamma), _p(beta), _p(out.t), B,
                                                 HW, G, int(silu), float(drop_p), int(seed),
                                                 _p(dropout_step_counter(x1.device)), _stream()),
          'sdb_groupnorm_apply_pack_dropout')
    return out


def groupnorm_bwd(x1, x2, da, stats, gamma, beta, dgamma, dbeta, B, HW, G, silu, add1=None, add2=None, drop_p=0.0,
                  seed=0):
    """-> (dx1, dx2 | None); dgamma/dbeta accumulated in place."""
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    work = torch.empty(B * (C1 + C2) * 2, dtype=torch.float32, device=x1.device)
    dx1 = torch.empty_like(x1)
    dx2 = torch.empty_like(x2) if x2 is not None else None
    check(lib().sdb_groupnorm_bwd(_p(x1), C1, _p(x2), C2, _p(da), _p(stats), _p(gamma), _p(beta), _p(work), _p(dx1),
                                  _p(dx2), _p(add1), _p(add2), _p(dgamma), _p(dbeta), B, HW, G, int(silu),
                                  float(drop_p), int(seed), _p(dropout_step_counter(x1.device)) if drop_p > 0 else None,
                                  _stream()), 'sdb_groupnorm_bwd')
    return dx1, dx2


def layernorm_bwd(x, dn, gamma, eps, dgamma, dbeta, add=None):
    M, C = x.shape
    dx = torch.empty_like(x)
    check(lib().sdb_layernorm_bwd(_p(x), _p(dn), _p(gamma), eps, _p(add), _p(dx), _p(dgamma), _p(dbeta), M, C,
                                  _stream()), 'sdb_layernorm_bwd')
    return dx


def attention_bwd(q, k, v, dout, dq, dk, dv, B, Lq, Lk, heads, d, scale):
    """All 2-D strided views [B*L, heads*d]; dq/dk/dv are written (views into the caller's gradient buffers)."""
    work = torch.empty(2 * B * heads * Lq, dtype=torch.float32, device=q.device)
    check(lib().sdb_attention_bwd(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(dout), dout.stride(0),
                                  _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), _p(work), B, Lq, Lk,
                                  heads, d, scale, _stream()), 'sdb_attention_bwd')


def geglu_bwd(u, dg):
    M, F2 = u.shape
    du = torch.empty_like(u)
    check(lib().sdb_geglu_bwd(_p(u), _p(dg), _p(du), M, F2 // 2, _stream()), 'sdb_geglu_bwd')
    return du


def up2_adjoint(dup, B, H, W, C):
    dx = torch.empty(B * H * W, C, dtype=torch.float32, device=dup.device)
    check(lib().sdb_up2_adjoint(_p(dup), _p(dx), B, H, W, C, _stream()), 'sdb_up2_adjoint')
    return dx


def pack_zero_up2(dy, B, H, W, C):
    out = Packed.empty(B * 4 * H * W, C, dy.device, bf16=True)
    check(lib().sdb_pack_zero_up2(_p(dy), _p(out.t), B, H, W, C, _stream()), 'sdb_pack_zero_up2')
    return out


def pack_nchw_pad(x, Cp):
    B, Cs = x.shape[:2]
    HW = x[0, 0].numel()
    out = Packed.empty(B * HW, Cp, x.device)
    check(lib().sdb_pack_nchw_pad(_p(x.contiguous()), _p(out.t), B, Cs, HW, Cp, _stream()), 'sdb_pack_nchw_pad')
    return out


def nhwc_to_nchw(rows, B, Cs, H, W):
    out = torch.empty(B, Cs, H, W, dtype=torch.float32, device=rows.device)
    check(lib().sdb_nhwc_to_nchw(_p(rows), rows.stride(0), _p(out), B, Cs, H * W, _stream()), 'sdb_nhwc_to_nchw')
    return out


def nchw_to_nhwc_pad(x, Cp):
    B, Cs = x.shape[:2]
    HW = x[0, 0].numel()
    out = torch.empty(B * HW, Cp, dtype=torch.float32, device=x.device)
    check(lib().sdb_nchw_to_nhwc_pad(_p(x.contiguous()), _p(out), B, Cs, HW, Cp, _stream()), 'sdb_nchw_to_nhwc_pad')
    return out


def im2col_T(a, B, H, W, C):
    out = Packed.empty(9 * C, B * H * W, a.t.device)
    check(lib().sdb_im2col_t(_p(a.t), _p(out.t), B, H, W, C, _stream()), 'sdb_im2col_t')
    return out


def gru_gates_bwd(gi, gh, h, dh_new):
    dgi, dgh, dh = torch.empty_like(gi), torch.empty_like(gh), torch.empty_like(h)
    R, D = h.shape
    check(lib().sdb_gru_gates_bwd(_p(gi), _p(gh), _p(h), _p(dh_new), _p(dgi), _p(dgh), _p(dh), R, D, _stream()),
          'sdb_gru_gates_bwd')
    return dgi, dgh, dh


def slot_attend_train(kv, q, B, N, S, D, scale, eps, want_mask):
    ws = lib().sdb_slot_attend_workspace(B, N, S, D)
    work = torch.empty(ws // 4, dtype=torch.float32, device=kv.device)
    mask = torch.empty(B, S, N, dtype=torch.float32, device=kv.device) if want_mask else None
    upd = Packed.empty(B * S, D, kv.device)
    upd32 = torch.empty(B * S, D, dtype=torch.float32, device=kv.device)
    cs = torch.empty(B * S, dtype=torch.float32, device=kv.device)
    check(lib().sdb_slot_attend_train(_p(kv), _p(q), _p(mask), _p(upd.t), _p(upd32), _p(cs), _p(work), B, N, S, D, scale,
                                      eps, _stream()), 'sdb_slot_attend_train')
    return upd, mask, upd32, cs


def slot_attend_bwd(kv, q, upd32, cs, d_upd, dkv, B, N, S, D, scale, eps, accumulate):
    dq = torch.empty(B * S, D, dtype=torch.float32, device=kv.device)
    check(lib().sdb_slot_attend_bwd(_p(kv), _p(q), _p(upd32), _p(cs), _p(d_upd), _p(dkv), _p(dq), B, N, S, D, scale, eps,
                                    int(accumulate), _stream()), 'sdb_slot_attend_bwd')
    return dq
