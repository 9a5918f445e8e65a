"""Slot transition function of the video models on libsdb200.

Mirrors /root/reference/slotdiffusion/video_based/models/predictor.py:20-44 (`TransformerPredictor`: an
nn.TransformerEncoder over the [B, S, D] slots, called between two frames by the per-frame Slot-Attention recurrence,
savi_diffusion.py:183-196 / savi.py:341-367): same constructor, same parameter tree (`transformer_encoder.layers.i.
self_attn.in_proj_weight` ...), so reference checkpoints load unchanged.  The nn modules only own the parameters; forward
and backward are ONE kernel schedule on the C ABI:

  per layer (pre-LN, the shipped configs; post-LN supported):
    LayerNorm+pack -> in_proj GEMM (q|k|v fused, tcgen05) -> token attention over the S <= 32 slots of a sample
    (csrc/predictor.cu, probabilities dropout inside) -> pack -> out_proj GEMM (+bias, +residual in the epilogue)
    LayerNorm+pack -> linear1 GEMM (+bias, ReLU, packed operand out) -> linear2 GEMM (+bias, +residual)
  training mode adds the three nn.Dropout(0.1) sites of nn.TransformerEncoderLayer as counter-based masks regenerated in
  the backward (no mask tensor is stored).

There is no PyTorch fallback: CPU tensors raise.
"""
import itertools

import torch
from torch import nn

from . import ops, parallel
from .backward import GradBuffer, Tape, _cat_T, layernorm_node, linear_node

_call_counter = itertools.count(1)


class _NoTape:
    """Inference: the schedule is the same, nothing is recorded."""
    G = None

    def push(self, fn):
        pass


def _seed(site):
    return (site * 0x632BE5AB) & 0x7FFFFFFFFFFFFFFF


def _dropout_add(tp, x, res, p, seed):
    """res + dropout(x) with its backward (mask regenerated from the counters)."""
    y = ops.dropout_add(x, res, p, seed)

    def bw():
        dy = tp.pop(y)
        if dy is None:
            return
        tp.acc(x, ops.dropout_add(dy, None, p, seed))
        if res is not None:
            tp.acc(res, dy)
    tp.push(bw)
    return y


def _pack(tp, x):
    xp = ops.pack_rows(x)

    def bw():
        tp.acc(x, tp.pop(xp))
    tp.push(bw)
    return xp


def _ln(tp, x, ln, G, want_fp32=False):
    if G is None:
        return ops.layernorm_pack(x, ln.weight, ln.bias, ln.eps, want_fp32=want_fp32)
    return layernorm_node(tp, x, ln, G, want_fp32=want_fp32)


def _linear(tp, wc, key, a, lin_w, lin_b, G, **kw):
    w = wc.linear(key, lin_w)
    if G is None:
        return ops.gemm(a, w, bias=lin_b, **kw)
    return linear_node(tp, a, w, lambda: _cat_T(wc, key, lin_w), G.view(lin_w), G.view(lin_b), bias=lin_b, **kw)


def encoder_layer(tp, wc, li, layer, x, B, S, G, p, seed0):
    """One nn.TransformerEncoderLayer (torch/nn/modules/transformer.py `_sa_block` / `_ff_block`) on rows x [B*S, D]."""
    sa = layer.self_attn
    heads = sa.num_heads
    pre = layer.norm_first
    train = G is not None

    def attn_block(xin_packed, res):
        qkv = _linear(tp, wc, (li, 'in'), xin_packed, sa.in_proj_weight, sa.in_proj_bias, G)
        pa = float(sa.dropout) if p > 0 else 0.0
        sd = _seed(seed0 + 4 * li)
        a = ops.token_attention(qkv, B, S, heads, pa, sd)
        if train:
            def bw():
                da = tp.pop(a)
                if da is not None:
                    tp.acc(qkv, ops.token_attention_bwd(qkv, da, B, S, heads, pa, sd))
            tp.push(bw)
        ap = _pack(tp, a) if train else ops.pack_rows(a)
        ow, ob = sa.out_proj.weight, sa.out_proj.bias
        if p > 0:
            o = _linear(tp, wc, (li, 'out'), ap, ow, ob, G)
            return _dropout_add(tp, o, res, p, _seed(seed0 + 4 * li + 1))
        return _linear(tp, wc, (li, 'out'), ap, ow, ob, G, residual=res)

    def ff_block(xin_packed, res):
        l1, l2 = layer.linear1, layer.linear2
        if p > 0:
            y1 = _linear(tp, wc, (li, 'l1'), xin_packed, l1.weight, l1.bias, G, relu=True)
            d1 = _dropout_add(tp, y1, None, p, _seed(seed0 + 4 * li + 2))
            y2 = _linear(tp, wc, (li, 'l2'), _pack(tp, d1), l2.weight, l2.bias, G)
            return _dropout_add(tp, y2, res, p, _seed(seed0 + 4 * li + 3))
        _, y1p = _linear(tp, wc, (li, 'l1'), xin_packed, l1.weight, l1.bias, G, relu=True, pack_out='none')
        return _linear(tp, wc, (li, 'l2'), y1p, l2.weight, l2.bias, G, residual=res)

    if pre:                                     # x = x + sa(norm1(x)); x = x + ff(norm2(x))
        x = attn_block(_ln(tp, x, layer.norm1, G), x)
        return ff_block(_ln(tp, x, layer.norm2, G), x)
    # post-LN: x = norm1(x + sa(x)); x = norm2(x + ff(x))
    h = attn_block(_pack(tp, x) if train else ops.pack_rows(x), x)
    hp, x = _ln(tp, h, layer.norm1, G, want_fp32=True)
    if train:
        _fp32_of_packed(tp, hp, x)
    h = ff_block(hp, x)
    hp, x = _ln(tp, h, layer.norm2, G, want_fp32=True)
    if train:
        _fp32_of_packed(tp, hp, x)
    return x


def _fp32_of_packed(tp, packed, y):
    """layernorm_node keys its gradient on the packed operand; the fp32 copy of the same values feeds it too."""
    def bw():
        tp.acc(packed, tp.pop(y))
    tp.push(bw)


def _check(mod, x):
    if not x.is_cuda:
        raise RuntimeError('slotdiffusion_b200.TransformerPredictor runs on CUDA (sm_100a) only; no CPU fallback')
    enc = mod.transformer_encoder
    l0 = enc.layers[0]
    D = x.shape[-1]
    S = x.shape[-2]
    dh = D // l0.self_attn.num_heads
    if not ops.token_attention_supported(S, dh):
        raise RuntimeError(f'TransformerPredictor: unsupported geometry S={S}, head dim {dh} (S <= 32, head dim 32/48/64)')
    for l in enc.layers:
        act = l.activation
        if not (act is torch.nn.functional.relu or isinstance(act, nn.ReLU)):
            raise RuntimeError('TransformerPredictor: only the ReLU feed-forward of the reference is implemented')


def predictor_schedule(mod, tp, x2, B, S, G, p, seed0):
    enc = mod.transformer_encoder
    for li, layer in enumerate(enc.layers):
        x2 = encoder_layer(tp, mod._wcache, li, layer, x2, B, S, G, p, seed0)
    if enc.norm is not None:
        hp, x2n = _ln(tp, x2, enc.norm, G, want_fp32=True)
        if G is not None:
            _fp32_of_packed(tp, hp, x2n)
        x2 = x2n
    return x2


class PredictorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        tp = Tape()
        layout = getattr(mod, '_gradbuf', None)
        ids = tuple(id(q) for q in mod.parameters())
        if layout is None or layout.params_ids != ids:
            layout = GradBuffer(mod)
            layout.params_ids = ids
            mod._gradbuf = layout
        G = tp.G = layout.instance(x.device)          # per call: the recurrence calls the predictor once per frame
        shape = tuple(x.shape)
        D = shape[-1]
        S = shape[-2]
        x2 = x.detach().contiguous().float().reshape(-1, D)
        B = x2.shape[0] // S
        p = float(mod.transformer_encoder.layers[0].dropout.p) if mod.training else 0.0
        seed0 = ((torch.initial_seed() + 0x9E3779B1 * parallel.rank()) * 1000003 + next(_call_counter) * 7919) & 0x3FFFFFFFFFFF
        with ops.training_scope():
            y = predictor_schedule(mod, tp, x2, B, S, G, p, seed0)
        ctx.tape, ctx.x2, ctx.y, ctx.shape, ctx.params = tp, x2, y, shape, params
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        tp = ctx.tape
        if tp is None:
            raise RuntimeError('slotdiffusion_b200: TransformerPredictor backward called twice on the same graph (retain_graph is not supported)')
        G = tp.G
        tp.set(ctx.y, dy.contiguous().float().reshape(-1, ctx.shape[-1]))
        with ops.training_scope():
            tp.run()
        dx = tp.pop(ctx.x2)
        ctx.tape = None
        if parallel.enabled():
            parallel.allreduce_flat(G.flat, async_op=True)
            parallel.wait_all()
        grads = tuple(G.view(q) if q.requires_grad else None for q in ctx.params)
        dx = dx.view(ctx.shape) if dx is not None and ctx.needs_input_grad[1] else None
        return (None, dx) + grads


class Predictor(nn.Module):
    """Base class of the transition functions (predictor.py:7-17)."""

    def forward(self, x):
        raise NotImplementedError

    def burnin(self, x):
        pass

    def reset(self):
        pass


class TransformerPredictor(Predictor):
    """Transformer encoder over the slots (same constructor and parameter names as the reference)."""

    def __init__(self, d_model=128, num_layers=1, num_heads=4, ffn_dim=256, norm_first=True):
        super().__init__()
        transformer_enc_layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=num_heads, dim_feedforward=ffn_dim,
                                                           norm_first=norm_first, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer=transformer_enc_layer, num_layers=num_layers,
                                                         enable_nested_tensor=False)      # (no parameters; silences a warning)
        self._wcache = ops.WeightCache()

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_wcache', None)
        d.pop('_gradbuf', None)
        d.pop('_sdb_graphs', None)
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self._wcache = ops.WeightCache()

    def invalidate_caches(self):
        self._wcache._c.clear()

    def forward(self, x):
        """x [B, S, D] (any leading dims) -> same shape."""
        _check(self, x)
        params = list(self.parameters())
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(q.requires_grad for q in params))
        if needs_grad:
            # never routed through graphed.py: the recurrence calls the predictor once per frame before ONE backward, and
            # a graphed callable cannot be replayed twice before its backward (its saved tensors are static buffers) --
            # video training is captured as a whole step instead (bench.py train_video)
            return PredictorFn.apply(self, x, *params)
        shape = x.shape
        D, S = shape[-1], shape[-2]
        with torch.no_grad(), ops.pack_format(ops.SDB_FMT_F16X2):
            x2 = x.contiguous().float().reshape(-1, D)
            p = float(self.transformer_encoder.layers[0].dropout.p) if self.training else 0.0
            seed0 = ((torch.initial_seed() + 0x9E3779B1 * parallel.rank()) * 1000003 + next(_call_counter) * 7919) & 0x3FFFFFFFFFFF
            return predictor_schedule(self, _NoTape(), x2, x2.shape[0] // S, S, None, p, seed0).view(shape)
