"""ResNet18/34-GN image encoder of SlotDiffusion on the B200 kernels, forward AND backward.

Mirrors the reference module -- factory names and keywords, attribute / parameter names and shapes (checkpoints load with
strict=True), forward(x [B,3,H,W]) -> feature map [B,C,H/4,W/4] (use_layer4=False) -- so dropin.install() rebinds the names
the reference's model constructor evaluates:
  resnet18 / resnet34 / ResNet   /root/reference/slotdiffusion/video_based/models/resnet.py:150-340 (img_based has a copy),
  constructed at img_based/models/slot_attention.py:185-188 (`eval(enc_dict['resnet'])(small_inputs=True, use_layer4=...)`)
  and video_based/models/savi.py (same code); ~32 % of the training-step FLOPs of the image model.
Only the configuration the repository ships is built: BasicBlock, GroupNorm(32) normalisation, small_inputs=True (3x3
stride-1 stem, no max-pool); anything else raises.

Schedule (post-activation residual blocks, resnet.py:72-90): every 3x3 convolution is an implicit-GEMM tcgen05 launch
(stride-2 ones on the phase-split operand); `conv -> GN -> ReLU` packs the next operand straight from the GroupNorm
statistics of the producing epilogue; the block end `relu(GN(h) + identity)` is one pass that emits the fp32 rows (next
identity) and the packed operand; the stride-2 1x1 `downsample` projection runs as the centre tap of a 3x3 stride-2 launch
on the same phase-split operand as conv1.  Backward = backward.py's tape (dgrad / wgrad GEMMs, GroupNorm backward kernels),
parameter gradients in one flat buffer, data-parallel all-reduce when parallel.enable_grad_allreduce() is active.
No PyTorch fallback: CPU tensors raise.
"""
import torch
from torch import nn

from . import ops, parallel
from .backward import GradBuffer, Tape, _conv_dgrad, _pad_conv, conv3_node, groupnorm_node
from .ops import SDB_PACK_PHASE2


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.GroupNorm(32, planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.GroupNorm(32, planes)
        self.downsample = downsample
        self.stride = stride


class _NoTape:
    """Inference: the same schedule with nothing recorded."""
    G = None

    def push(self, fn):
        pass


class ResNet(nn.Module):
    """resnet.py:150-315 restricted to (BasicBlock, GroupNorm(32), small_inputs=True)."""

    def __init__(self, block, layers, small_inputs=True, use_layer4=True, zero_init_residual=False, groups=1,
                 width_per_group=64, replace_stride_with_dilation=None, norm_layer=nn.GroupNorm):
        super().__init__()
        if block is not BasicBlock or norm_layer is not nn.GroupNorm or not small_inputs or groups != 1 \
                or width_per_group != 64 or (replace_stride_with_dilation and any(replace_stride_with_dilation)):
            raise NotImplementedError('slotdiffusion_b200.resnet: only the shipped encoder configuration is built '
                                      '(BasicBlock, GroupNorm(32), small_inputs=True, no dilation)')
        self.small_inputs, self.use_layer4 = small_inputs, use_layer4
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.maxpool = nn.Identity()
        self.bn1 = nn.GroupNorm(32, 64)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        if use_layer4:
            self.layer4 = self._make_layer(512, layers[3], stride=2)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.GroupNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)
        self._wc = ops.WeightCache()
        self._layout = None

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                       nn.GroupNorm(32, planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        layers += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_wc', None)
        d.pop('_layout', None)
        d.pop('_sdb_graphs', None)
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self._wc, self._layout = ops.WeightCache(), None

    def invalidate_caches(self):
        self._wc.clear()

    # ------------------------------------------------------------------------------------------------ schedule
    def _stats(self, t, gs, gn, B, HW):
        if gs is not None:
            return ops.groupnorm_finalize_cb(gs[0], t.shape[1], B, HW, gn.num_groups, gn.eps, gs[1])
        return ops.groupnorm_stats(t, None, B, HW, gn.num_groups, gn.eps)

    @staticmethod
    def _gs(B, HW, C, dev):
        """(GroupNorm partial sums accumulated by the producing GEMM epilogue, channels per block): blocks of 4 channels,
        or of 2 for the 64-channel stem / layer1 (32 groups of 2 channels).  None: the exact two-pass statistics kernel."""
        if B * HW < 4096 or HW % 16 or C % 64:
            return None
        cb = 4 if C % 128 == 0 else 2
        return torch.zeros(B * (C // cb) * 2, dtype=torch.float32, device=dev), cb

    def _block_end(self, tp, h, st_h, gn_h, idn, st_i, gn_i, B, HW, want_packed):
        """out = relu(GN(h) + identity) as fp32 rows (+ packed operand); backward through both normalisations."""
        out, pk = ops.groupnorm_add_relu(h, st_h, gn_h, B, HW, idn=idn, stats_i=st_i, gn_i=gn_i, want_packed=want_packed)
        G = tp.G

        def bw():
            dout = tp.pop(out)
            dpk = tp.pop(pk) if pk is not None else None
            if dout is None and dpk is None:
                return
            dout = dpk if dout is None else (dout if dpk is None else ops.add3(dout, dpk))
            dz = ops.act_bwd(dout, out, 'relu')
            dh, _ = ops.groupnorm_bwd(h, None, dz, st_h, gn_h.weight, gn_h.bias, G.view(gn_h.weight), G.view(gn_h.bias), B, HW,
                                      gn_h.num_groups, 0, add1=tp.pop(h))
            tp.set(h, dh)
            if idn is not None:
                if st_i is not None:
                    di, _ = ops.groupnorm_bwd(idn, None, dz, st_i, gn_i.weight, gn_i.bias, G.view(gn_i.weight),
                                              G.view(gn_i.bias), B, HW, gn_i.num_groups, 0, add1=tp.pop(idn))
                    tp.set(idn, di)
                else:
                    tp.acc(idn, dz)
        tp.push(bw)
        return out, pk

    def _conv(self, tp, a, conv, key, B, H, W, Cin, stride2=False):
        """3x3 convolution (no bias) -> (fp32 rows [B*H*W, Cout], GroupNorm partial sums | None); H, W = output size."""
        wc, G = self._wc, tp.G
        Cout = conv.weight.shape[0]
        gs = self._gs(B, H * W, Cout, a.t.device)
        if G is None:
            mode = ops.SDB_A_CONV3S2 if stride2 else ops.SDB_A_CONV3
            y = ops.gemm(a, wc.conv3(key, conv.weight), conv=(mode, B, H, W, Cin), gsum=gs[0] if gs else None,
                         gsum_cb=gs[1] if gs else 4, rows_per_group=H * W)
        else:
            y = conv3_node(tp, a, wc.conv3(key, conv.weight), lambda: _conv_dgrad(wc, key, conv.weight), G.view(conv.weight),
                           None, (B, H, W, Cin), stride2=stride2, gsum=gs[0] if gs else None, gsum_cb=gs[1] if gs else 4)
        return y, gs

    def _downsample(self, tp, xp, ds, key, B, H, W, Cin):
        """1x1 stride-2 projection (resnet.py:264-268) = centre tap of a 3x3 stride-2 launch on the phase-split operand."""
        wc, G = self._wc, tp.G
        conv, gn = ds[0], ds[1]
        Cout = conv.weight.shape[0]

        def expand():
            w3 = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float32, device=conv.weight.device)
            w3[:, :, 1, 1] = conv.weight.detach()[:, :, 0, 0]
            return w3
        w_fwd = wc._get((key, 'ds'), (conv.weight,), lambda: ops.pack_weight_conv3(expand()))
        gs = self._gs(B, H * W, Cout, xp.t.device)
        if G is None:
            y = ops.gemm(xp, w_fwd, conv=(ops.SDB_A_CONV3S2, B, H, W, Cin), gsum=gs[0] if gs else None,
                         gsum_cb=gs[1] if gs else 4, rows_per_group=H * W)
        else:
            dw3 = torch.empty(Cout, Cin, 3, 3, dtype=torch.float32, device=conv.weight.device)
            # pushed BEFORE the conv node: runs after its wgrad in the reversed replay
            tp.push(lambda: G.view(conv.weight).view(Cout, Cin).copy_(dw3[:, :, 1, 1]))
            y = conv3_node(tp, xp, w_fwd, lambda: wc._get((key, 'ds_dg'), (conv.weight,),
                                                           lambda: ops.pack_weight_conv3_dgrad(expand())),
                           dw3, None, (B, H, W, Cin), stride2=True, gsum=gs[0] if gs else None, gsum_cb=gs[1] if gs else 4)
        return y, self._stats(y, gs, gn, B, H * W), gn

    def _run(self, tp, x):
        wc, G = self._wc, tp.G
        B, _, H, W = x.shape
        dev = x.device
        # ---- stem: conv1 -> bn1 -> relu (resnet.py:288-291; maxpool is the identity for small inputs)
        gs0 = None
        if G is None:
            zero_b = wc._get('stem_b', (self.conv1.weight,), lambda: torch.zeros(64, dtype=torch.float32, device=dev))
            h = ops.conv3_in(x.float(), self.conv1.weight, zero_b)
        else:       # training: channels zero-padded to 64 so that the stem's wgrad is the common implicit-GEMM launch
            xpk = ops.pack_nchw_pad(x.float(), 64)
            w_in = wc._get('stem_pad', (self.conv1.weight,), lambda: ops.pack_weight_conv3(_pad_conv(self.conv1.weight, 64, 64)))
            gs0 = self._gs(B, H * W, 64, dev)
            h = conv3_node(tp, xpk, w_in, None, G.view(self.conv1.weight), None, (B, H, W, 64), need_da=False,
                           gsum=gs0[0] if gs0 else None, gsum_cb=gs0[1] if gs0 else 4)
        st = self._stats(h, gs0, self.bn1, B, H * W)
        layers = [self.layer1, self.layer2, self.layer3] + ([self.layer4] if self.use_layer4 else [])
        blocks = [b for layer in layers for b in layer]
        nxt_s2 = blocks[0].stride == 2
        cur, cur_p = self._block_end(tp, h, st, self.bn1, None, None, None, B, H * W, want_packed=not nxt_s2)
        trace = getattr(self, '_relu_trace', None)      # tests: the activation pattern of every ReLU, in forward order
        if trace is not None:
            trace.append((cur > 0, (B, 64, H, W)))
        C = 64
        for bi, blk in enumerate(blocks):
            key = id(blk)
            Cout = blk.conv1.weight.shape[0]
            if blk.stride == 2:
                Ho, Wo = H // 2, W // 2
                xp = ops.pack_nhwc(cur, None, B, H, W, SDB_PACK_PHASE2)

                def bw_split(cur=cur, xp=xp):
                    tp.acc(cur, tp.pop(xp))         # stride-2 dgrads arrive in the un-split layout (backward.conv3_node)
                tp.push(bw_split)
                h1, gs1 = self._conv(tp, xp, blk.conv1, (key, 'c1'), B, Ho, Wo, C, stride2=True)
                idn, st_i, gn_i = self._downsample(tp, xp, blk.downsample, key, B, Ho, Wo, C)
                H, W = Ho, Wo
            else:
                if blk.downsample is not None:
                    raise NotImplementedError('slotdiffusion_b200.resnet: stride-1 projection shortcuts are not built')
                h1, gs1 = self._conv(tp, cur_p, blk.conv1, (key, 'c1'), B, H, W, C)
                idn, st_i, gn_i = cur, None, None
            HW = H * W
            st1 = self._stats(h1, gs1, blk.bn1, B, HW)
            if G is None:
                p1 = ops.groupnorm_pack_fused(h1, None, blk.bn1.weight, blk.bn1.bias, B, HW, 32, blk.bn1.eps, 2, stats=st1)
            else:
                p1 = groupnorm_node(tp, h1, None, blk.bn1, G, B, HW, 2, st1)          # act 2 = ReLU
            if trace is not None:
                trace.append((p1.unpack() > 0, (B, Cout, H, W)))
            h2, gs2 = self._conv(tp, p1, blk.conv2, (key, 'c2'), B, H, W, Cout)
            st2 = self._stats(h2, gs2, blk.bn2, B, HW)
            last = bi == len(blocks) - 1
            nxt_s2 = (not last) and blocks[bi + 1].stride == 2
            cur, cur_p = self._block_end(tp, h2, st2, blk.bn2, idn, st_i, gn_i, B, HW, want_packed=not (last or nxt_s2))
            if trace is not None:
                trace.append((cur > 0, (B, Cout, H, W)))
            C = Cout
        return cur, (B, C, H, W)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('slotdiffusion_b200.resnet.ResNet runs on CUDA (sm_100a) only; no CPU fallback')
        params = tuple(self.parameters())
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            if x.requires_grad:
                raise NotImplementedError('slotdiffusion_b200.resnet: gradient w.r.t. the input image is not built '
                                          '(the reference never asks for it)')
            from . import graphed
            if graphed.enabled(self):       # eager training loops: schedules replayed from CUDA graphs (graphed.py)
                g = graphed.graphs_of(self, lambda a: _ResNetFn.apply(self, a, *tuple(self.parameters())), lambda: [self._wc])
                return g(x)
            return _ResNetFn.apply(self, x, *params)
        with torch.no_grad(), ops.pack_format(ops.SDB_FMT_F16X2):
            rows, (B, C, H, W) = self._run(_NoTape(), x)
            return ops.nhwc_to_nchw(rows, B, C, H, W)

    _forward_impl = forward


class _ResNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        tp = Tape()
        if net._layout is None or net._layout.params_ids != tuple(id(p) for p in params):
            net._layout = GradBuffer(net)
            net._layout.params_ids = tuple(id(p) for p in params)
        tp.G = net._layout.instance(x.device)        # per call (never on the module)
        with ops.pack_format(ops.SDB_FMT_F16X2), ops.training_scope():
            rows, geo = net._run(tp, x.detach())
        ctx.tape, ctx.rows, ctx.geo, ctx.params = tp, rows, geo, params
        B, C, H, W = geo
        return ops.nhwc_to_nchw(rows, B, C, H, W)

    @staticmethod
    def backward(ctx, dy):
        tp = ctx.tape
        if tp is None:
            raise RuntimeError('slotdiffusion_b200: ResNet backward called twice on the same graph (retain_graph is not supported)')
        B, C, H, W = ctx.geo
        with ops.pack_format(ops.SDB_FMT_F16X2), ops.training_scope():
            tp.set(ctx.rows, ops.nchw_to_nhwc_pad(dy.contiguous().float(), C))
            tp.run()
        G = tp.G
        ctx.tape = None
        if parallel.enabled():
            parallel.allreduce_flat(G.flat, async_op=True)
            parallel.wait_all()
        return (None, None) + tuple(G.view(p) if p.requires_grad else None for p in ctx.params)


def resnet18(small_inputs=True, use_layer4=True, **kwargs):
    return ResNet(BasicBlock, [2, 2, 2, 2], small_inputs, use_layer4, **kwargs)


def resnet34(small_inputs=True, use_layer4=True, **kwargs):
    return ResNet(BasicBlock, [3, 4, 6, 3], small_inputs, use_layer4, **kwargs)
