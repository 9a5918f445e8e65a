"""Kernel schedule of the UNet forward pass (reference: unet.py:551-576 and the modules it calls).

Activation layout: NHWC fp32 rows [B*H*W, C].  Every dense contraction (3x3 / 1x1 conv, linear) is one
sdb_gemm call (tcgen05, implicit im2col through TMA); normalisation + activation are fused into the kernel
that produces the GEMM operand; bias, timestep-embedding add and residual adds are fused into the GEMM
epilogue; skip-connection concats are never materialised (the consumers read two sources).

Per forward the timestep-embedding projections of all ResBlocks are one GEMM, and the cross-attention K/V
projections of the slots for all transformer layers are one GEMM (`context_kv`), which the DPM-Solver driver
computes once per sampling run because the slots do not change across the 20 evaluations.
"""
import torch

from . import ops
from .ops import SDB_A_CONV3, SDB_A_CONV3S2, SDB_PACK_PHASE2, SDB_PACK_PLAIN, SDB_PACK_UP2


GS_MIN_ROWS = 4096     # below this the producing GEMMs want split-K, which excludes the fused GroupNorm sums


class Act:
    """NHWC activation: rows [B*H*W, C]; gs = GroupNorm partial sums [B, C/4, 2] accumulated by its producer."""
    __slots__ = ('t', 'H', 'W', 'C', 'gs')

    def __init__(self, t, H, W, C, gs=None):
        self.t, self.H, self.W, self.C, self.gs = t, H, W, C, gs


class UNetExecutor:
    def __init__(self, net):
        # keep a plain reference without registering `net` as a sub-module of itself
        object.__setattr__(self, 'net', net)
        self.wc = ops.WeightCache()
        from . import unet as U
        self.U = U
        self.resblocks = [m for m in net.modules() if isinstance(m, U.ResBlock)]
        self.tblocks = [m for m in net.modules() if isinstance(m, U.BasicTransformerBlock)]
        # column offsets inside the fused timestep-embedding GEMM output
        self.emb_off, off = {}, 0
        for m in self.resblocks:
            self.emb_off[id(m)] = off
            off += m.out_channels
        self.emb_total = off
        # column offsets inside the fused context K|V GEMM output
        self.kv_off, off = {}, 0
        for m in self.tblocks:
            self.kv_off[id(m)] = off
            off += 2 * m.attn2.to_k.weight.shape[0]
        self.kv_total = off
        self._gs_arena = None
        self._gs_retired = []      # outgrown arenas: captured CUDA graphs keep zeroing / accumulating into them
        self._gs_off = 0

    # ------------------------------------------------------------------ GroupNorm partial-sum arena
    def begin(self, B, device):
        """Start of a forward: one memset clears every partial-sum buffer the epilogues will accumulate into."""
        need = B * 2 * 16384          # floats: sum over producing sites of C/4 * 2 (~11k) per sample
        if self._gs_arena is None or self._gs_arena.numel() < need or self._gs_arena.device != device:
            # A sampler graph captured at a smaller batch has this arena's ADDRESS baked into its memset and epilogue
            # atomics and stays cached by shape: the old block must never return to the caching allocator (a replay
            # would zero / accumulate into whatever tensor owns it by then), so it is retired, not freed (33 MB at
            # B = 256; growth events are rare).
            if self._gs_arena is not None:
                self._gs_retired.append(self._gs_arena)
            self._gs_arena = torch.zeros(need, dtype=torch.float32, device=device)
        else:
            self._gs_arena.zero_()
        self._gs_off = 0

    def _gs(self, B, HW, C):
        """Partial-sum buffer for an activation produced by a GEMM epilogue, or None when not applicable."""
        if B * HW < GS_MIN_ROWS or HW % 16 or C % 128:     # (C / 32 groups) % 4 == 0
            return None
        n = B * (C // 4) * 2
        if self._gs_arena is None or self._gs_off + n > self._gs_arena.numel():
            return None
        v = self._gs_arena[self._gs_off:self._gs_off + n]
        self._gs_off += n
        return v

    def input_conv_sums(self, h, B):
        """The input convolution is not a GEMM, so nothing accumulated GroupNorm sums for its output; it feeds three
        GroupNorms (first ResBlock + two skip concats): one pass over it replaces three exact-statistics kernels."""
        gs = self._gs(B, h.H * h.W, h.C)
        if gs is not None:
            h.gs = ops.channel_block_sums(h.t, gs, B, h.H * h.W)

    def group_norm(self, x1, x2, gn, B, HW, silu):
        """GroupNorm(+SiLU) -> packed operand; statistics from the producers' epilogue sums when available."""
        x2t = x2.t if x2 is not None else None
        if x1.gs is not None and (x2 is None or x2.gs is not None):
            return ops.groupnorm_pack_fused(x1.t, x2t, gn.weight, gn.bias, B, HW, gn.num_groups, gn.eps, silu,
                                            gsum1=x1.gs, gsum2=x2.gs if x2 is not None else None)
        stats = ops.groupnorm_stats(x1.t, x2t, B, HW, gn.num_groups, gn.eps)
        return ops.groupnorm_pack_fused(x1.t, x2t, gn.weight, gn.bias, B, HW, gn.num_groups, gn.eps, silu, stats=stats)

    # ------------------------------------------------------------------ pieces
    def time_embedding(self, t, B):
        net, wc = self.net, self.wc
        if t.numel() == 1 and B > 1:
            t = t.reshape(1).expand(B)
        tp = ops.timestep_embedding_pack(t, net.model_channels)                          # unet.py:560-561
        _, e1 = ops.gemm(tp, wc.linear('te0', net.time_embed[0].weight), bias=net.time_embed[0].bias,
                         pack_out='silu', keep_c=False)
        _, emb = ops.gemm(e1, wc.linear('te2', net.time_embed[2].weight), bias=net.time_embed[2].bias,
                          pack_out='silu', keep_c=False)                                 # unet.py:562
        # every ResBlock's Linear(SiLU(emb)) in one GEMM                                   unet.py:279
        w = wc.linear('emb_all', *[m.emb_layers[1].weight for m in self.resblocks])
        b = wc.cat('emb_all_b', *[m.emb_layers[1].bias for m in self.resblocks])
        return ops.gemm(emb, w, bias=b)                                                  # [B, emb_total]

    def context_kv(self, context):
        """to_k | to_v of every cross-attention layer applied to the slots: [B*S, kv_total]."""
        B, S, Dc = context.shape
        cp = ops.pack_rows(context.reshape(B * S, Dc).contiguous().float())
        ws = []
        for m in self.tblocks:
            ws += [m.attn2.to_k.weight, m.attn2.to_v.weight]
        return ops.gemm(cp, self.wc.linear('ctx_kv_all', *ws))

    def res_block(self, m, x1, x2, emb_all, B):
        wc, key = self.wc, id(m)
        H, W = x1.H, x1.W
        HW = H * W
        Cin = x1.C + (x2.C if x2 is not None else 0)
        Cout = m.out_channels
        x2t = x2.t if x2 is not None else None
        gn1, gn2 = m.in_layers[0], m.out_layers[0]
        p = self.group_norm(x1, x2, gn1, B, HW, silu=True)
        off = self.emb_off[key]
        gs_h = self._gs(B, HW, Cout)
        h = ops.gemm(p, wc.conv3((key, 'c1'), m.in_layers[2].weight), bias=m.in_layers[2].bias,
                     rowvec=emb_all[:, off:off + Cout], rows_per_group=HW, conv=(SDB_A_CONV3, B, H, W, Cin), gsum=gs_h)
        p2 = self.group_norm(Act(h, H, W, Cout, gs_h), None, gn2, B, HW, silu=True)
        if isinstance(m.skip_connection, torch.nn.Identity):
            assert x2 is None
            xs = x1.t
        else:
            xp = ops.pack_nhwc(x1.t, x2t, B, H, W, SDB_PACK_PLAIN)
            xs = ops.gemm(xp, wc.linear((key, 'skip'), m.skip_connection.weight), bias=m.skip_connection.bias)
        gs_o = self._gs(B, HW, Cout)
        out = ops.gemm(p2, wc.conv3((key, 'c2'), m.out_layers[3].weight), bias=m.out_layers[3].bias, residual=xs,
                       conv=(SDB_A_CONV3, B, H, W, Cout), gsum=gs_o, rows_per_group=HW)
        return Act(out, H, W, Cout, gs_o)

    def attention(self, a, xn, B, L, kv=None, S=None, residual=None, key=None):
        """attention.py:182-206.  xn: packed LayerNorm output."""
        wc = self.wc
        C = a.to_q.weight.shape[0]
        heads, d = a.heads, C // a.heads
        if kv is None:      # self attention: fused q|k|v projection
            qkv = ops.gemm(xn, wc.linear((key, 'qkv'), a.to_q.weight, a.to_k.weight, a.to_v.weight))
            o = ops.attention_pack(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, L, L, heads, d, a.scale)
        else:
            q = ops.gemm(xn, wc.linear((key, 'q'), a.to_q.weight))
            o = ops.attention_pack(q, kv[0], kv[1], B, L, S, heads, d, a.scale)
        return ops.gemm(o, wc.linear((key, 'o'), a.to_out[0].weight), bias=a.to_out[0].bias, residual=residual)

    def spatial_transformer(self, m, x, ctx_kv, B, S):
        wc, key = self.wc, id(m)
        H, W, C = x.H, x.W, x.C
        L = H * W
        pn = self.group_norm(x, None, m.norm, B, L, silu=False)
        t = ops.gemm(pn, wc.linear((key, 'pin'), m.proj_in.weight), bias=m.proj_in.bias)
        nblk = len(m.transformer_blocks)
        for i, blk in enumerate(m.transformer_blocks):
            bk = id(blk)
            n1 = ops.layernorm_pack(t, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
            t = self.attention(blk.attn1, n1, B, L, residual=t, key=(bk, 'a1'))
            n2 = ops.layernorm_pack(t, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            off = self.kv_off[bk]
            kv = (ctx_kv[:, off:off + C], ctx_kv[:, off + C:off + 2 * C])
            t = self.attention(blk.attn2, n2, B, L, kv=kv, S=S, residual=t, key=(bk, 'a2'))
            n3 = ops.layernorm_pack(t, blk.norm3.weight, blk.norm3.bias, blk.norm3.eps)
            proj = blk.ff.net[0].proj
            wg, bg = wc.geglu((bk, 'ff0'), proj.weight, proj.bias)
            u = ops.gemm(n3, wg, bias=bg, geglu=True)            # a * gelu(g) straight from the accumulators
            w2 = wc.linear((bk, 'ff2'), blk.ff.net[2].weight)
            if i == nblk - 1:   # only proj_out consumes it: emit the packed operand, skip the fp32 copy
                _, tp = ops.gemm(u, w2, bias=blk.ff.net[2].bias, residual=t, pack_out='none', keep_c=False)
            else:
                t = ops.gemm(u, w2, bias=blk.ff.net[2].bias, residual=t)
        gs_o = self._gs(B, L, C)
        out = ops.gemm(tp, wc.linear((key, 'pout'), m.proj_out.weight), bias=m.proj_out.bias, residual=x.t, gsum=gs_o,
                       rows_per_group=L)
        return Act(out, H, W, C, gs_o)

    def downsample(self, m, x, B):
        Ho, Wo = x.H // 2, x.W // 2
        xp = ops.pack_nhwc(x.t, None, B, x.H, x.W, SDB_PACK_PHASE2)
        gs = self._gs(B, Ho * Wo, m.out_channels)
        out = ops.gemm(xp, self.wc.conv3((id(m), 'op'), m.op.weight), bias=m.op.bias,
                       conv=(SDB_A_CONV3S2, B, Ho, Wo, x.C), gsum=gs, rows_per_group=Ho * Wo)
        return Act(out, Ho, Wo, m.out_channels, gs)

    def upsample(self, m, x, B):
        Ho, Wo = 2 * x.H, 2 * x.W
        xp = ops.pack_nhwc(x.t, None, B, x.H, x.W, SDB_PACK_UP2)
        gs = self._gs(B, Ho * Wo, m.out_channels)
        out = ops.gemm(xp, self.wc.conv3((id(m), 'conv'), m.conv.weight), bias=m.conv.bias,
                       conv=(SDB_A_CONV3, B, Ho, Wo, x.C), gsum=gs, rows_per_group=Ho * Wo)
        return Act(out, Ho, Wo, m.out_channels, gs)

    def run_block(self, block, x1, x2, emb_all, ctx_kv, B, S):
        U = self.U
        h = x1
        for layer in block:
            if isinstance(layer, U.ResBlock):
                h = self.res_block(layer, h, x2, emb_all, B)
                x2 = None
            elif isinstance(layer, U.SpatialTransformer):
                h = self.spatial_transformer(layer, h, ctx_kv, B, S)
            elif isinstance(layer, U.Downsample):
                h = self.downsample(layer, h, B)
            elif isinstance(layer, U.Upsample):
                h = self.upsample(layer, h, B)
            else:
                raise RuntimeError(f'unexpected layer {type(layer)}')
        return h

    # ------------------------------------------------------------------ whole forward
    def forward_inference(self, x, timesteps, context, ctx_kv=None):
        net = self.net
        B, Cin, H, W = x.shape
        S = context.shape[1]
        self.begin(B, x.device)
        emb_all = self.time_embedding(timesteps, B)
        if ctx_kv is None:
            ctx_kv = self.context_kv(context)
        conv_in = net.input_blocks[0][0]
        h = Act(ops.conv3_in(x.float(), conv_in.weight, conv_in.bias), H, W, net.model_channels)   # unet.py:408
        self.input_conv_sums(h, B)
        hs = [h]
        for block in list(net.input_blocks)[1:]:                                                   # unet.py:566-568
            h = self.run_block(block, h, None, emb_all, ctx_kv, B, S)
            hs.append(h)
        h = self.run_block(net.middle_block, h, None, emb_all, ctx_kv, B, S)
        for block in net.output_blocks:                                                            # unet.py:570-572
            h = self.run_block(block, h, hs.pop(), emb_all, ctx_kv, B, S)
        return self.head(h, B)

    def head(self, h, B):
        """unet.py:537-542: conv3x3(SiLU(GN(h))) -> NCHW."""
        net = self.net
        gn, conv = net.out[0], net.out[2]
        if h.gs is not None:
            stats = ops.groupnorm_finalize(h.gs, None, h.C, 0, B, h.H * h.W, gn.num_groups, gn.eps)
        else:
            stats = ops.groupnorm_stats(h.t, None, B, h.H * h.W, gn.num_groups, gn.eps)
        return ops.conv3_out(h.t, stats, gn.weight, gn.bias, conv.weight, conv.bias, B, h.H, h.W, gn.num_groups)

    def __call__(self, x, timesteps, context, ctx_kv=None):
        net = self.net
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or (context is not None and context.requires_grad)
            or any(p.requires_grad for p in net.parameters()))
        if context is None:
            raise RuntimeError('UNetModel: conditioning context (slots) is required (conditioning_key="crossattn")')
        if needs_grad:
            from .backward import unet_forward_train
            return unet_forward_train(self, x, timesteps, context)
        if net.training and net.dropout > 0 and torch.is_grad_enabled():
            raise RuntimeError('training-mode forward without gradients is not supported; use torch.no_grad()/eval()')
        with torch.no_grad(), ops.pack_format(ops.unet_inference_format()):
            return self.forward_inference(x, timesteps, context, ctx_kv)
