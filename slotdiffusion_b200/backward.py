"""Training (forward + backward) schedules of the hot modules on the C-ABI kernels.

The reference obtains gradients from torch autograd over eager ops (LDM.loss_function,
/root/reference/slotdiffusion/video_based/models/ddpm/ldm.py:59-83 -> ddpm.py:270-273 -> unet.py:551-576; Slot Attention
through all iterations, img_based/models/slot_attention.py:78-102).  Here each module is ONE torch.autograd.Function:
its forward runs the hand-written kernel schedule while recording a tape of backward closures; its backward replays
the tape in reverse.  Every contraction of the backward pass is an sdb_gemm launch (dgrad on rotated / transposed
weight operands, wgrad as an MN-major tcgen05 GEMM over the pixels with split-K); normalisation / attention /
activation backward are the kernels in csrc/backward.cu.  Parameter gradients are written straight into one flat fp32
buffer (fused projections occupy contiguous ranges), which is also the bucket layout of the data-parallel all-reduce.
"""
import itertools

import torch

from . import ops, parallel
from .ops import (SDB_A_CONV3, SDB_A_CONV3S2, SDB_PACK_PHASE2, SDB_PACK_PLAIN, SDB_PACK_UP2)

_step_counter = itertools.count(1)


class Tape:
    """Reverse-mode tape: closures + a gradient table keyed by object identity (tensors / Packed operands)."""

    def __init__(self):
        self.fns = []
        self.g = {}
        self.keep = []
        # per-call state of the module schedules (never kept on the shared trainer / module: re-entrancy)
        self.G = None          # GradBuffer.instance of this call
        self.dctx = None       # UNet: gradient of the context rows
        self.xpk = None        # UNet: packed input operand (key of dx)
        self.ykey = None       # UNet: key of the output gradient
        self.reduced_lo = None  # data parallel: flat[reduced_lo:] has already been handed to the all-reduce

    def bucket(self, start_off, force=False):
        """Data-parallel gradient exchange overlapped with backward: the flat buffer is laid out in module
        order, the tape runs in reverse module order, so when the backward of a top-level block has finished,
        flat[start_off:] is final.  Ranges are handed to NCCL (its own stream, ordered after the kernels enqueued so far)
        once they reach parallel.BUCKET_BYTES; the remaining backward kernels keep running on the compute stream."""
        if not parallel.enabled():
            return
        hi = self.G.total if self.reduced_lo is None else self.reduced_lo
        if start_off >= hi:
            return
        if force or (hi - start_off) * 4 >= parallel.BUCKET_BYTES:
            parallel.allreduce_flat(self.G.flat[start_off:hi], async_op=True)
            self.reduced_lo = start_off

    def acc(self, key, grad):
        if grad is None:
            return
        k = id(key)
        cur = self.g.get(k)
        if cur is None:
            self.g[k] = grad if grad.is_contiguous() else grad.contiguous()
        else:
            self.g[k] = ops.add3(cur, grad if grad.is_contiguous() else grad.contiguous())

    def set(self, key, grad):
        self.g[id(key)] = grad

    def get(self, key):
        return self.g.get(id(key))

    def pop(self, key):
        return self.g.pop(id(key), None)

    def push(self, fn):
        self.fns.append(fn)

    def run(self):
        for fn in reversed(self.fns):
            fn()
        self.fns = []


class GradBuffer:
    """One flat fp32 buffer holding the gradient of every parameter of a module; `groups` lists parameters that must
    be contiguous (the row-concatenated weights of fused GEMMs) so their wgrad GEMM writes in place."""

    def __init__(self, module, groups=()):
        order, seen = [], set()
        for grp in groups:
            for p in grp:
                if id(p) not in seen:
                    order.append((p, False))
                    seen.add(id(p))
            order[-1] = (order[-1][0], True)          # pad only after a whole group
        for p in module.parameters():
            if id(p) not in seen:
                order.append((p, True))
                seen.add(id(p))
        self.off, off = {}, 0
        for p, pad in order:
            self.off[id(p)] = off
            off += p.numel()
            if pad:
                off = (off + 3) // 4 * 4            # 16-byte aligned starts for vectorised epilogues
        self.total = off
        self.flat = None
        self.params = [p for p, _ in order]

    def instance(self, device):
        """Storage for ONE forward/backward pair: same layout, its own zeroed flat buffer.  The module may be called
        several times before a single backward (the video models run Slot Attention once per frame,
        savi_diffusion.py:183-196; two UNet evaluations can share one graph), so gradient storage belongs to the call,
        never to the module; autograd may also adopt the returned views as .grad, so a buffer is never reused."""
        g = GradBuffer.__new__(GradBuffer)
        g.off, g.total, g.params = self.off, self.total, self.params
        g.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        return g

    def view(self, p):
        o = self.off[id(p)]
        return self.flat[o:o + p.numel()].view(p.shape)

    def span(self, plist, rows, cols):
        o = self.off[id(plist[0])]
        assert sum(p.numel() for p in plist) == rows * cols
        e = o
        for p in plist:                       # fused projections must really be back to back
            assert self.off[id(p)] == e, 'GradBuffer.span: parameters are not contiguous in the flat buffer'
            e += p.numel()
        return self.flat[o:o + rows * cols].view(rows, cols)


def _cat_T(wc, key, *ws):
    """Packed transpose [K, sum N_i] of row-concatenated weights (operand of dX = dY W)."""
    def make():
        w2 = [w.detach().reshape(w.shape[0], -1) for w in ws]
        w = w2[0] if len(w2) == 1 else torch.cat(w2, 0)
        return ops.pack_weight_T(w.contiguous())
    return wc._get(('T',) + (key,), ws, make)


def _conv_dgrad(wc, key, w):
    return wc._get(('dg', key), (w,), lambda: ops.pack_weight_conv3_dgrad(w.detach().contiguous()))


def _pad_conv(w, cout_pad, cin_pad):
    wp = torch.zeros(cout_pad, cin_pad, 3, 3, dtype=torch.float32, device=w.device)
    wp[:w.shape[0], :w.shape[1]] = w.detach()
    return wp


# =============================================================================================== generic nodes
def linear_node(tp, a, w, wT_fn, dW, db, bias=None, residual=None, relu=False, need_da=True, pack_out=None, keep_c=True,
                rowvec=None, rows_per_group=0):
    """y = a W^T (+bias)(+residual); returns y (or (y, packed)).  dW [N,K] / db [N] are gradient views (written /
    accumulated).  wT_fn() -> Packed [K,N]."""
    res = ops.gemm(a, w, bias=bias, residual=residual, relu=relu, pack_out=pack_out, keep_c=True, rowvec=rowvec,
                   rows_per_group=rows_per_group)
    y, yp = (res if pack_out is not None else (res, None))

    def bw():
        dy = tp.pop(y)
        if yp is not None:
            dpk = tp.pop(yp)
            if dpk is not None:
                if pack_out in ('silu', 'relu'):
                    dpk = ops.act_bwd(dpk, y, pack_out)
                dy = dpk if dy is None else ops.add3(dy, dpk)
        if dy is None:
            return
        if relu:
            dy = ops.act_bwd(dy, y, 'relu')
        dyp, dyT = ops.grad_pack(dy, want_rows=need_da, bias_grad=db)
        if need_da:
            tp.acc(a, ops.gemm(dyp, wT_fn()))
        ops.gemm(dyT, ops.transpose_packed(a, to_bf16=True), out=dW)
        if residual is not None:
            tp.acc(residual, dy)
    tp.push(bw)
    return (y, yp) if pack_out is not None else y


def conv3_node(tp, a, w, wdg_fn, dWp, db, geo, bias=None, rowvec=None, demb=None, residual=None, need_da=True,
               stride2=False, gsum=None, cin_w=None, gsum_cb=4):
    """3x3 conv as implicit GEMM.  a: packed NHWC operand (phase split for stride 2); geo = (B, H, W, C) OUTPUT size and
    input channels.  dWp: gradient view of the conv weight [Cout, Cin_w, 3, 3]."""
    B, H, W, C = geo
    mode = SDB_A_CONV3S2 if stride2 else SDB_A_CONV3
    y = ops.gemm(a, w, bias=bias, rowvec=rowvec, rows_per_group=H * W, residual=residual, conv=(mode, B, H, W, C),
                 gsum=gsum, gsum_cb=gsum_cb)

    def bw():
        dy = tp.pop(y)
        if dy is None:
            return
        dyp, _ = ops.grad_pack(dy, want_T=False, bias_grad=db, group_grad=demb, rows_per_group=H * W)
        if need_da:
            if stride2:    # transposed conv = stride-1 conv of the zero-inserted gradient with the rotated kernel
                z = ops.pack_zero_up2(dy, B, H, W, dy.shape[1])
                tp.acc(a, ops.gemm(z, wdg_fn(), conv=(SDB_A_CONV3, B, 2 * H, 2 * W, dy.shape[1])))
            else:
                tp.acc(a, ops.gemm(dyp, wdg_fn(), conv=(SDB_A_CONV3, B, H, W, dy.shape[1])))
        c9 = ops.gemm_wgrad_conv(a, dyp, B, H, W, C, stride2=stride2)
        ops.wgrad_conv3_scatter(c9, dWp, C)
        if residual is not None:
            tp.acc(residual, dy)
    tp.push(bw)
    return y


def groupnorm_node(tp, x1, x2, gn, G_, B, HW, silu, stats, drop_p=0.0, seed=0):
    """x1/x2: fp32 NHWC rows (tensors).  Returns the packed GN(+SiLU)(+dropout) operand."""
    if drop_p > 0:
        p = ops.groupnorm_pack_dropout(x1, x2, gn.weight, gn.bias, stats, B, HW, gn.num_groups, silu, drop_p, seed)
    else:
        p = ops.groupnorm_pack_fused(x1, x2, gn.weight, gn.bias, B, HW, gn.num_groups, gn.eps, silu, stats=stats)

    def bw():
        da = tp.pop(p)
        if da is None:
            return
        dx1, dx2 = ops.groupnorm_bwd(x1, x2, da, stats, gn.weight, gn.bias, G_.view(gn.weight), G_.view(gn.bias), B, HW,
                                     gn.num_groups, silu, add1=tp.pop(x1), add2=tp.pop(x2) if x2 is not None else None,
                                     drop_p=drop_p, seed=seed)
        tp.set(x1, dx1)
        if x2 is not None:
            tp.set(x2, dx2)
    tp.push(bw)
    return p


def layernorm_node(tp, x, ln, G_, want_fp32=False):
    res = ops.layernorm_pack(x, ln.weight, ln.bias, ln.eps, want_fp32=want_fp32)
    n = res[0] if want_fp32 else res

    def bw():
        dn = tp.pop(n)
        if dn is None:
            return
        tp.set(x, ops.layernorm_bwd(x, dn, ln.weight, ln.eps, G_.view(ln.weight), G_.view(ln.bias), add=tp.pop(x)))
    tp.push(bw)
    return res


# =============================================================================================== UNet
class UNetTrainer:
    """Training schedule of UNetModel.forward (unet.py:551-576) around a UNetExecutor (shares its weight cache)."""

    def __init__(self, ex):
        self.ex = ex
        self.net = ex.net
        self.wc = ex.wc
        net = self.net
        # Flat gradient layout: the GLOBAL fused groups first (cross-attention K|V of every layer, the timestep
        # projections of every ResBlock: their wgrad runs at the very end of backward), then every other parameter in
        # module order -- so the backward of top-level block i completes flat[block_off[i]:] (see Tape.bucket).
        # attn1.to_q / to_k / to_v of a block are consecutive in module order (attention.py:171-173), C*C % 4 == 0: the
        # fused QKV wgrad still writes one contiguous [3C, C] range (GradBuffer.span asserts it).
        groups = []
        kv = []
        for blk in ex.tblocks:
            kv += [blk.attn2.to_k.weight, blk.attn2.to_v.weight]
        groups.append(kv)
        groups.append([m.emb_layers[1].weight for m in ex.resblocks])
        groups.append([m.emb_layers[1].bias for m in ex.resblocks])
        self.G = GradBuffer(net, groups)
        self.kv_params = kv
        front = {id(p) for g in groups for p in g}

        def first_off(block):
            offs = [self.G.off[id(p)] for p in block.parameters() if id(p) not in front]
            return min(offs) if offs else None
        self.block_off = {id(b): first_off(b) for b in
                          list(net.input_blocks)[1:] + [net.middle_block] + list(net.output_blocks)}
        self.block_off[id(net.out)] = first_off(net.out)

    # ------------------------------------------------------------------ pieces
    def _stats(self, x1, x2, gs1, gs2, gn, B, HW):
        if gs1 is not None and (x2 is None or gs2 is not None):
            C1 = x1.shape[1]
            C2 = x2.shape[1] if x2 is not None else 0
            return ops.groupnorm_finalize(gs1, gs2, C1, C2, B, HW, gn.num_groups, gn.eps)
        return ops.groupnorm_stats(x1, x2, B, HW, gn.num_groups, gn.eps)

    def res_block(self, tp, m, x1, x2, emb_all, demb_all, B, drop_p, seed):
        ex, wc, G, key = self.ex, self.wc, tp.G, id(m)
        H, W = x1.H, x1.W
        HW = H * W
        Cin = x1.C + (x2.C if x2 is not None else 0)
        Cout = m.out_channels
        x2t = x2.t if x2 is not None else None
        gn1, gn2 = m.in_layers[0], m.out_layers[0]
        st1 = self._stats(x1.t, x2t, x1.gs, x2.gs if x2 is not None else None, gn1, B, HW)
        p = groupnorm_node(tp, x1.t, x2t, gn1, G, B, HW, True, st1)
        off = ex.emb_off[key]
        c1, c2 = m.in_layers[2], m.out_layers[3]
        gs_h = ex._gs(B, HW, Cout)
        h = conv3_node(tp, p, wc.conv3((key, 'c1'), c1.weight), lambda: _conv_dgrad(wc, (key, 'c1'), c1.weight),
                       G.view(c1.weight), G.view(c1.bias), (B, H, W, Cin), bias=c1.bias,
                       rowvec=emb_all[:, off:off + Cout], demb=demb_all[:, off:off + Cout], gsum=gs_h)
        st2 = self._stats(h, None, gs_h, None, gn2, B, HW)
        p2 = groupnorm_node(tp, h, None, gn2, G, B, HW, True, st2, drop_p, seed)
        if isinstance(m.skip_connection, torch.nn.Identity):
            xs = x1.t
        else:
            sk = m.skip_connection
            xp = ops.pack_nhwc(x1.t, x2t, B, H, W, SDB_PACK_PLAIN)
            C1 = x1.C

            def bw_pack():
                da = tp.pop(xp)
                if da is None:
                    return
                if x2t is None:
                    tp.acc(x1.t, da)
                else:
                    tp.acc(x1.t, da[:, :C1])
                    tp.acc(x2t, da[:, C1:])
            tp.push(bw_pack)
            xs = linear_node(tp, xp, wc.linear((key, 'skip'), sk.weight), lambda: _cat_T(wc, (key, 'skip'), sk.weight),
                             G.view(sk.weight).view(Cout, Cin), G.view(sk.bias), bias=sk.bias)
        gs_o = ex._gs(B, HW, Cout)
        out = conv3_node(tp, p2, wc.conv3((key, 'c2'), c2.weight), lambda: _conv_dgrad(wc, (key, 'c2'), c2.weight),
                         G.view(c2.weight), G.view(c2.bias), (B, H, W, Cout), bias=c2.bias, residual=xs, gsum=gs_o)
        return self.ex_act(out, H, W, Cout, gs_o)

    def ex_act(self, t, H, W, C, gs=None):
        from .unet_exec import Act
        return Act(t, H, W, C, gs)

    def attention(self, tp, a, xn, B, L, t_res, key, kv=None, dkv=None, S=None):
        wc, G = self.wc, tp.G
        C = a.to_q.weight.shape[0]
        heads, d = a.heads, C // a.heads
        if kv is None:
            ws = (a.to_q.weight, a.to_k.weight, a.to_v.weight)
            qkv = linear_node(tp, xn, wc.linear((key, 'qkv'), *ws), lambda: _cat_T(wc, (key, 'qkv'), *ws),
                              G.span(ws, 3 * C, C), None)
            q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
            o = ops.attention_pack(q, k, v, B, L, L, heads, d, a.scale)

            def bw():
                do = tp.pop(o)
                if do is None:
                    return
                dqkv = torch.empty_like(qkv)
                ops.attention_bwd(q, k, v, do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], B, L, L, heads, d, a.scale)
                tp.acc(qkv, dqkv)
            tp.push(bw)
        else:
            qt = linear_node(tp, xn, wc.linear((key, 'q'), a.to_q.weight), lambda: _cat_T(wc, (key, 'q'), a.to_q.weight),
                             G.view(a.to_q.weight), None)
            o = ops.attention_pack(qt, kv[0], kv[1], B, L, S, heads, d, a.scale)

            def bw():
                do = tp.pop(o)
                if do is None:
                    return
                dq = torch.empty_like(qt)
                ops.attention_bwd(qt, kv[0], kv[1], do, dq, dkv[0], dkv[1], B, L, S, heads, d, a.scale)
                tp.acc(qt, dq)
            tp.push(bw)
        lo = a.to_out[0]
        return linear_node(tp, o, wc.linear((key, 'o'), lo.weight), lambda: _cat_T(wc, (key, 'o'), lo.weight),
                           G.view(lo.weight), G.view(lo.bias), bias=lo.bias, residual=t_res)

    def spatial_transformer(self, tp, m, x, ctx_kv, d_ctx_kv, B, S):
        ex, wc, G, key = self.ex, self.wc, tp.G, id(m)
        H, W, C = x.H, x.W, x.C
        L = H * W
        st = self._stats(x.t, None, x.gs, None, m.norm, B, L)
        pn = groupnorm_node(tp, x.t, None, m.norm, G, B, L, False, st)
        t = linear_node(tp, pn, wc.linear((key, 'pin'), m.proj_in.weight),
                        lambda: _cat_T(wc, (key, 'pin'), m.proj_in.weight), G.view(m.proj_in.weight).view(C, C),
                        G.view(m.proj_in.bias), bias=m.proj_in.bias)
        for blk in m.transformer_blocks:
            bk = id(blk)
            n1 = layernorm_node(tp, t, blk.norm1, G)
            t = self.attention(tp, blk.attn1, n1, B, L, t, (bk, 'a1'))
            n2 = layernorm_node(tp, t, blk.norm2, G)
            off = ex.kv_off[bk]
            kv = (ctx_kv[:, off:off + C], ctx_kv[:, off + C:off + 2 * C])
            dkv = (d_ctx_kv[:, off:off + C], d_ctx_kv[:, off + C:off + 2 * C])
            t = self.attention(tp, blk.attn2, n2, B, L, t, (bk, 'a2'), kv=kv, dkv=dkv, S=S)
            n3 = layernorm_node(tp, t, blk.norm3, G)
            proj, l2 = blk.ff.net[0].proj, blk.ff.net[2]
            u = linear_node(tp, n3, wc.linear((bk, 'ff0u'), proj.weight), lambda p=proj: _cat_T(wc, (id(p), 'ff0u'), p.weight),
                            G.view(proj.weight), G.view(proj.bias), bias=proj.bias)
            gp = ops.geglu_pack(u)

            def bw_geglu(u=u, gp=gp):
                dg = tp.pop(gp)
                if dg is not None:
                    tp.acc(u, ops.geglu_bwd(u, dg))
            tp.push(bw_geglu)
            t = linear_node(tp, gp, wc.linear((bk, 'ff2'), l2.weight), lambda l=l2: _cat_T(wc, (id(l), 'ff2'), l.weight),
                            G.view(l2.weight), G.view(l2.bias), bias=l2.bias, residual=t)
        tpk = ops.pack_rows(t)

        def bw_pack(t=t, tpk=tpk):
            tp.acc(t, tp.pop(tpk))
        tp.push(bw_pack)
        gs_o = ex._gs(B, L, C)
        out = ops.gemm(tpk, wc.linear((key, 'pout'), m.proj_out.weight), bias=m.proj_out.bias, residual=x.t, gsum=gs_o,
                       rows_per_group=L)
        po = m.proj_out

        def bw_out():
            dy = tp.pop(out)
            if dy is None:
                return
            dyp, dyT = ops.grad_pack(dy, bias_grad=G.view(po.bias))
            tp.acc(tpk, ops.gemm(dyp, _cat_T(wc, (key, 'pout'), po.weight)))
            ops.gemm(dyT, ops.transpose_packed(tpk, to_bf16=True), out=G.view(po.weight).view(C, C))
            tp.acc(x.t, dy)
        tp.push(bw_out)
        return self.ex_act(out, H, W, C, gs_o)

    def downsample(self, tp, m, x, B):
        ex, wc, G = self.ex, self.wc, tp.G
        Ho, Wo = x.H // 2, x.W // 2
        xp = ops.pack_nhwc(x.t, None, B, x.H, x.W, SDB_PACK_PHASE2)
        gs = ex._gs(B, Ho * Wo, m.out_channels)
        op = m.op
        y = ops.gemm(xp, wc.conv3((id(m), 'op'), op.weight), bias=op.bias, conv=(SDB_A_CONV3S2, B, Ho, Wo, x.C), gsum=gs,
                     rows_per_group=Ho * Wo)

        def bw():
            dy = tp.pop(y)
            if dy is None:
                return
            dyp, _ = ops.grad_pack(dy, want_T=False, bias_grad=G.view(op.bias))
            z = ops.pack_zero_up2(dy, B, Ho, Wo, m.out_channels)
            tp.acc(x.t, ops.gemm(z, _conv_dgrad(wc, (id(m), 'op'), op.weight),
                                 conv=(SDB_A_CONV3, B, x.H, x.W, m.out_channels)))
            c9 = ops.gemm_wgrad_conv(xp, dyp, B, Ho, Wo, x.C, stride2=True)
            ops.wgrad_conv3_scatter(c9, G.view(op.weight), x.C)
        tp.push(bw)
        return self.ex_act(y, Ho, Wo, m.out_channels, gs)

    def upsample(self, tp, m, x, B):
        ex, wc, G = self.ex, self.wc, tp.G
        Ho, Wo = 2 * x.H, 2 * x.W
        xp = ops.pack_nhwc(x.t, None, B, x.H, x.W, SDB_PACK_UP2)

        def bw_up():
            da = tp.pop(xp)
            if da is not None:
                tp.acc(x.t, ops.up2_adjoint(da, B, x.H, x.W, x.C))
        tp.push(bw_up)
        gs = ex._gs(B, Ho * Wo, m.out_channels)
        cv = m.conv
        y = conv3_node(tp, xp, wc.conv3((id(m), 'conv'), cv.weight), lambda: _conv_dgrad(wc, (id(m), 'conv'), cv.weight),
                       G.view(cv.weight), G.view(cv.bias), (B, Ho, Wo, x.C), bias=cv.bias, gsum=gs)
        return self.ex_act(y, Ho, Wo, m.out_channels, gs)

    def run_block(self, tp, block, x1, x2, emb_all, demb_all, ctx_kv, d_ctx_kv, B, S, drop_p, seed):
        U = self.ex.U
        h = x1
        for i, layer in enumerate(block):
            if isinstance(layer, U.ResBlock):
                h = self.res_block(tp, layer, h, x2, emb_all, demb_all, B, drop_p, seed + 7919 * (i + 1))
                x2 = None
            elif isinstance(layer, U.SpatialTransformer):
                h = self.spatial_transformer(tp, layer, h, ctx_kv, d_ctx_kv, B, S)
            elif isinstance(layer, U.Downsample):
                h = self.downsample(tp, layer, h, B)
            elif isinstance(layer, U.Upsample):
                h = self.upsample(tp, layer, h, B)
            else:
                raise RuntimeError(f'unexpected layer {type(layer)}')
        return h

    # ------------------------------------------------------------------ whole forward (recording) and backward
    def forward(self, tp, x, timesteps, context, need_dx):
        ex, net, wc = self.ex, self.net, self.wc
        B, Cin, H, W = x.shape
        S, Dc = context.shape[1], context.shape[2]
        dev = x.device
        G = tp.G = self.G.instance(dev)
        ex.begin(B, dev)
        drop_p = float(net.dropout) if net.training else 0.0
        seed0 = ((torch.initial_seed() + 0x9E3779B1 * parallel.rank()) * 1000003 + next(_step_counter) * 7907) & 0x7FFFFFFFFFFF
        # ---- timestep embedding chain (unet.py:560-564 + every ResBlock's emb_layers in one GEMM)
        t = timesteps
        if t.numel() == 1 and B > 1:
            t = t.reshape(1).expand(B)
        tpk = ops.timestep_embedding_pack(t, net.model_channels)
        te0, te2 = net.time_embed[0], net.time_embed[2]
        _, e1p = linear_node(tp, tpk, wc.linear('te0', te0.weight), None, G.view(te0.weight), G.view(te0.bias),
                             bias=te0.bias, need_da=False, pack_out='silu')
        _, embp = linear_node(tp, e1p, wc.linear('te2', te2.weight), lambda: _cat_T(wc, 'te2', te2.weight),
                              G.view(te2.weight), G.view(te2.bias), bias=te2.bias, pack_out='silu')
        ew = [m.emb_layers[1].weight for m in ex.resblocks]
        eb = [m.emb_layers[1].bias for m in ex.resblocks]
        emb_all = ops.gemm(embp, wc.linear('emb_all', *ew), bias=wc.cat('emb_all_b', *eb))
        demb_all = torch.zeros_like(emb_all)

        def bw_emb():
            dyp, dyT = ops.grad_pack(demb_all, bias_grad=G.span(eb, 1, ex.emb_total).view(-1))
            tp.acc(embp, ops.gemm(dyp, _cat_T(wc, 'emb_all', *ew)))
            ops.gemm(dyT, ops.transpose_packed(embp, to_bf16=True), out=G.span(ew, ex.emb_total, ew[0].shape[1]))
        tp.push(bw_emb)
        # ---- cross-attention K|V of the slots for all layers in one GEMM
        ctx2 = context.reshape(B * S, Dc).contiguous().float()
        cp = ops.pack_rows(ctx2)
        ctx_kv = ops.gemm(cp, wc.linear('ctx_kv_all', *self.kv_params))
        d_ctx_kv = torch.zeros_like(ctx_kv)

        def bw_ctx():
            dyp, dyT = ops.grad_pack(d_ctx_kv)
            tp.dctx = ops.gemm(dyp, _cat_T(wc, 'ctx_kv_all', *self.kv_params))
            ops.gemm(dyT, ops.transpose_packed(cp, to_bf16=True), out=G.span(self.kv_params, ex.kv_total, Dc))
        tp.push(bw_ctx)
        # ---- input conv (3 -> model_channels) as a GEMM on channels zero-padded to 64
        conv_in = net.input_blocks[0][0]
        xpk = ops.pack_nchw_pad(x.float(), 64)
        w_in = wc._get('conv_in_pad', (conv_in.weight,), lambda: ops.pack_weight_conv3(_pad_conv(conv_in.weight, conv_in.weight.shape[0], 64)))
        h0 = conv3_node(tp, xpk, w_in, lambda: wc._get('conv_in_pad_dg', (conv_in.weight,), lambda: ops.pack_weight_conv3_dgrad(
                            _pad_conv(conv_in.weight, conv_in.weight.shape[0], 64))),
                        G.view(conv_in.weight), G.view(conv_in.bias), (B, H, W, 64), bias=conv_in.bias, need_da=need_dx)
        tp.xpk = xpk
        h = self.ex_act(h0, H, W, net.model_channels)
        hs = [h]
        blocks = list(net.input_blocks)[1:]
        def mark(block):
            # pushed BEFORE the block's nodes: runs AFTER all of them in the reversed replay
            o = self.block_off[id(block)]
            if o is not None:
                tp.push(lambda: tp.bucket(o))
        for bi, block in enumerate(blocks):
            mark(block)
            h = self.run_block(tp, block, h, None, emb_all, demb_all, ctx_kv, d_ctx_kv, B, S, drop_p, seed0 + 104729 * bi)
            hs.append(h)
        mark(net.middle_block)
        h = self.run_block(tp, net.middle_block, h, None, emb_all, demb_all, ctx_kv, d_ctx_kv, B, S, drop_p, seed0 + 15485863)
        for bi, block in enumerate(net.output_blocks):
            mark(block)
            h = self.run_block(tp, block, h, hs.pop(), emb_all, demb_all, ctx_kv, d_ctx_kv, B, S, drop_p,
                               seed0 + 32452843 + 104729 * bi)
        mark(net.out)
        # ---- output head: GN + SiLU + conv3x3 (C -> out_channels), NHWC -> NCHW
        gn, conv = net.out[0], net.out[2]
        HW = h.H * h.W
        st = self._stats(h.t, None, h.gs, None, gn, B, HW)
        pk = groupnorm_node(tp, h.t, None, gn, G, B, HW, True, st)
        Co = conv.weight.shape[0]
        y_rows = ops.gemm(pk, wc.conv3('conv_out', conv.weight), bias=conv.bias, conv=(SDB_A_CONV3, B, h.H, h.W, h.C))
        y = ops.nhwc_to_nchw(y_rows, B, Co, h.H, h.W)
        hC, hH, hW = h.C, h.H, h.W

        ykey = tp.ykey = object()     # NOT the output tensor: ctx -> output -> grad_fn -> ctx would be a reference
        #                                  cycle that keeps stale AccumulateGrad nodes alive (breaks CUDA-graph capture)

        def bw_head():
            dy = tp.pop(ykey)                                # [B, Co, H, W]
            dyr = ops.nchw_to_nhwc_pad(dy, 64)               # [M, 64] zero-padded channels
            btmp = torch.zeros(64, dtype=torch.float32, device=dev)
            dyp, _ = ops.grad_pack(dyr, want_T=False, bias_grad=btmp)
            G.view(conv.bias).copy_(btmp[:Co])
            wdg = wc._get('conv_out_dg', (conv.weight,), lambda: ops.pack_weight_conv3_dgrad(_pad_conv(conv.weight, 64, hC)))
            tp.acc(pk, ops.gemm(dyp, wdg, conv=(SDB_A_CONV3, B, hH, hW, 64)))
            c9 = ops.gemm_wgrad_conv(pk, dyp, B, hH, hW, hC)
            ops.wgrad_conv3_scatter(c9, G.view(conv.weight), hC)
        tp.push(bw_head)
        return y

    def backward(self, tp, dy, x_shape, need_dx):
        tp.set(tp.ykey, dy.contiguous().float())
        tp.run()
        dctx = tp.dctx
        dx = None
        if need_dx:
            da = tp.pop(tp.xpk)                           # [M, 64]
            B, Cin, H, W = x_shape
            dx = ops.nhwc_to_nchw(da, B, Cin, H, W)
        return dx, dctx


class UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer, x, timesteps, context, *params):
        tp = Tape()
        need_dx = x.requires_grad
        with ops.training_scope():
            y = trainer.forward(tp, x.detach(), timesteps, context.detach(), need_dx)
        ctx.trainer, ctx.tape, ctx.need_dx = trainer, tp, need_dx
        ctx.x_shape, ctx.ctx_shape = tuple(x.shape), tuple(context.shape)
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        tr, tp = ctx.trainer, ctx.tape
        if tp is None:
            raise RuntimeError('slotdiffusion_b200: UNet backward called twice on the same graph (retain_graph is not supported)')
        with ops.training_scope():
            dx, dctx = tr.backward(tp, dy, ctx.x_shape, ctx.need_dx)
        G = tp.G
        ctx.tape = None
        if parallel.enabled():              # data parallel: what the buckets have not covered yet, then join
            tp.bucket(0, force=True)
            parallel.wait_all()
        grads = tuple(G.view(p) if p.requires_grad else None for p in ctx.params)
        dcontext = dctx.view(ctx.ctx_shape) if dctx is not None and ctx.needs_input_grad[3] else None
        return (None, dx, None, dcontext) + grads


def unet_forward_train(ex, x, timesteps, context):
    tr = getattr(ex, '_trainer', None)
    if tr is None:
        tr = UNetTrainer(ex)
        ex._trainer = tr
    params = tuple(ex.net.parameters())
    return UNetFn.apply(tr, x, timesteps, context, *params)


# =============================================================================================== Slot Attention
class SlotAttentionFn(torch.autograd.Function):
    """Forward = the kernel chain of autograd.slot_attention_forward (recorded); backward through all iterations
    (slot_attention.py:78-102 has no stop-gradient)."""

    @staticmethod
    def forward(ctx, mod, want_mask, inputs, slots, *params):
        with ops.training_scope():
            return SlotAttentionFn._forward(ctx, mod, want_mask, inputs, slots, *params)

    @staticmethod
    def _forward(ctx, mod, want_mask, inputs, slots, *params):
        tp = Tape()
        layout = getattr(mod, '_gradbuf', None)
        if layout is None or layout.params_ids != tuple(id(p) for p in mod.parameters()):
            layout = GradBuffer(mod, [[mod.project_k.weight, mod.project_v.weight]])
            layout.params_ids = tuple(id(p) for p in mod.parameters())
            mod._gradbuf = layout
        B, N, Din = inputs.shape
        S, D = slots.shape[1], slots.shape[2]
        dev = inputs.device
        G = tp.G = layout.instance(dev)      # per call: T frames -> T buffers, summed by autograd
        wc = mod._wcache
        x = inputs.detach().contiguous().float().reshape(B * N, Din)
        s0 = slots.detach().contiguous().float().reshape(B * S, D)
        xn = layernorm_node(tp, x, mod.norm_inputs, G)
        kvw = (mod.project_k.weight, mod.project_v.weight)
        kv = linear_node(tp, xn, wc.linear('kv', *kvw), lambda: _cat_T(wc, 'kv', *kvw), G.span(kvw, 2 * D, Din), None)
        dkv = torch.empty_like(kv)
        state = {'first': True}
        wq, gru = mod.project_q[1].weight, mod.gru
        w1, w2 = mod.mlp[1], mod.mlp[3]
        acc_w = {}

        def dW_acc(p, shape=None):
            """Weights reused in every iteration: accumulate their wgrad across iterations."""
            v = G.view(p) if shape is None else G.view(p).view(shape)
            return v

        cur = s0
        mask = None
        for it in range(mod.num_iterations):
            last = it == mod.num_iterations - 1
            prev = cur
            sn = layernorm_node(tp, prev, mod.project_q[0], G)
            q = _linear_acc(tp, sn, wc.linear('q', wq), lambda: _cat_T(wc, 'q', wq), G.view(wq), None)
            upd, m, upd32, cs = ops.slot_attend_train(kv, q, B, N, S, D, mod.attn_scale, mod.eps, want_mask and last)
            if m is not None:
                mask = m

            def bw_attend(q=q, upd=upd, upd32=upd32, cs=cs):
                dU = tp.pop(upd)
                if dU is None:
                    return
                dq = ops.slot_attend_bwd(kv, q, upd32, cs, dU, dkv, B, N, S, D, mod.attn_scale, mod.eps, not state['first'])
                state['first'] = False
                tp.acc(q, dq)
            tp.push(bw_attend)
            gi = _linear_acc(tp, upd, wc.linear('ih', gru.weight_ih), lambda: _cat_T(wc, 'ih', gru.weight_ih),
                             G.view(gru.weight_ih), G.view(gru.bias_ih), bias=gru.bias_ih)
            hp = ops.pack_rows(prev)

            def bw_hp(prev=prev, hp=hp):
                tp.acc(prev, tp.pop(hp))
            tp.push(bw_hp)
            gh = _linear_acc(tp, hp, wc.linear('hh', gru.weight_hh), lambda: _cat_T(wc, 'hh', gru.weight_hh),
                             G.view(gru.weight_hh), G.view(gru.bias_hh), bias=gru.bias_hh)
            h = ops.gru_gates(gi, gh, prev)

            def bw_gru(gi=gi, gh=gh, prev=prev, h=h):
                dh = tp.pop(h)
                if dh is None:
                    return
                dgi, dgh, dprev = ops.gru_gates_bwd(gi, gh, prev, dh)
                tp.acc(gi, dgi)
                tp.acc(gh, dgh)
                tp.acc(prev, dprev)
            tp.push(bw_gru)
            hn = layernorm_node(tp, h, mod.mlp[0], G)
            y1, y1p = _linear_acc(tp, hn, wc.linear('m1', w1.weight), lambda: _cat_T(wc, 'm1', w1.weight), G.view(w1.weight),
                                  G.view(w1.bias), bias=w1.bias, relu=True, pack_out='none')
            cur = _linear_acc(tp, y1p, wc.linear('m2', w2.weight), lambda: _cat_T(wc, 'm2', w2.weight), G.view(w2.weight),
                              G.view(w2.bias), bias=w2.bias, residual=h)

        def bw_kv_seed():
            # runs after every iteration's attend backward has accumulated into dkv (tape is replayed in reverse)
            if not state['first']:
                tp.acc(kv, dkv)
        # must execute BEFORE the kv linear's backward and AFTER all iterations: insert right after the kv node
        tp.fns.insert(2, bw_kv_seed)
        ctx.tape, ctx.mod, ctx.G, ctx.out, ctx.x, ctx.s0 = tp, mod, G, cur, x, s0
        ctx.shapes = (tuple(inputs.shape), tuple(slots.shape))
        ctx.params = params
        out = cur.view(B, S, D)
        if want_mask:
            ctx.mark_non_differentiable(mask)
            return out, mask
        return out, torch.empty(0, device=dev)

    @staticmethod
    def backward(ctx, dslots, _dmask):
        tp, G = ctx.tape, ctx.G
        if tp is None:
            raise RuntimeError('slotdiffusion_b200: Slot Attention backward called twice on the same graph (retain_graph is not supported)')
        B, S, D = ctx.shapes[1]
        tp.set(ctx.out, dslots.contiguous().float().reshape(B * S, D))
        with ops.training_scope():
            tp.run()
        dx = tp.pop(ctx.x)
        ds0 = tp.pop(ctx.s0)
        ctx.tape = None
        if parallel.enabled():
            parallel.allreduce_flat(G.flat, async_op=True)
            parallel.wait_all()
        grads = tuple(G.view(p) if p.requires_grad else None for p in ctx.params)
        dx = dx.view(ctx.shapes[0]) if dx is not None and ctx.needs_input_grad[2] else None
        ds0 = ds0.view(ctx.shapes[1]) if ds0 is not None and ctx.needs_input_grad[3] else None
        return (None, None, dx, ds0) + grads


def _linear_acc(tp, a, w, wT_fn, dW, db, bias=None, residual=None, relu=False, pack_out=None):
    """linear_node for weights shared by several applications (slot-attention iterations): wgrad accumulates (+=)."""
    res = ops.gemm(a, w, bias=bias, residual=residual, relu=relu, pack_out=pack_out, keep_c=True)
    y, yp = (res if pack_out is not None else (res, None))

    def bw():
        dy = tp.pop(y)
        if yp is not None:
            dpk = tp.pop(yp)
            if dpk is not None:
                dy = dpk if dy is None else ops.add3(dy, dpk)
        if dy is None:
            return
        if relu:
            dy = ops.act_bwd(dy, y, 'relu')
        dyp, dyT = ops.grad_pack(dy, bias_grad=db)
        tp.acc(a, ops.gemm(dyp, wT_fn()))
        dw = ops.gemm(dyT, ops.transpose_packed(a, to_bf16=True))
        ops.add3(dW.view(dw.shape), dw, out=dW.view(dw.shape))
        if residual is not None:
            tp.acc(residual, dy)
    tp.push(bw)
    return (y, yp) if pack_out is not None else y
