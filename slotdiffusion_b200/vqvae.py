"""Frozen VQ-VAE encoder / decoder of the LDM on the B200 kernels.

Mirrors the reference modules -- constructor keywords, attribute / parameter names and shapes, forward() signatures --
so VQ-VAE checkpoints load with strict=True and dropin.install() can rebind the two names the reference constructs:
  Encoder   /root/reference/slotdiffusion/video_based/models/vqvae/modules.py:168-262   (VQVAE.encode, VQVAE.py:94-99;
            called in EVERY training step: x0 = vae.encode(img), ldm.py:62-64)
  Decoder   modules.py:265-362                                                           (VQVAE.decode, VQVAE.py:108-113;
            called after sampling: log_images, ldm.py:112-124)
The nn.Modules own the parameters; all arithmetic runs in libsdb200 (same kernel families as the UNet: implicit-GEMM 3x3
convolutions on tcgen05 with bias / residual / GroupNorm-sum epilogues, GroupNorm+swish fused into the operand producer).
The LDM uses the VQ-VAE frozen and under torch.no_grad(): these modules are INFERENCE ONLY (a forward that needs gradients
raises; training the VQ-VAE itself -- scripts/train.py --task ... vqvae -- stays the reference's PyTorch code).
There is no PyTorch fallback: CPU tensors raise.
"""
import torch
from torch import nn

from . import ops
from .ops import SDB_A_CONV3, SDB_A_CONV3S2A, SDB_PACK_PHASE2, SDB_PACK_PLAIN, SDB_PACK_UP2
from .unet_exec import Act


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1)


def make_attn(in_channels, attn_type='vanilla'):
    assert attn_type in ('vanilla', 'none'), f'attn_type {attn_type} unknown'
    return AttnBlock(in_channels) if attn_type == 'vanilla' else nn.Identity(in_channels)


# ======================================================================================================= kernel schedule
class _Exec:
    """Kernel schedule shared by Encoder and Decoder.  Activations are NHWC fp32 rows [B*H*W, C] (unet_exec.Act); `gs` are
    the GroupNorm partial sums accumulated by the producing GEMM epilogue (4-channel blocks: levels with C % 128 == 0),
    otherwise the exact two-pass statistics kernel runs (the 64-channel levels: 2 channels per group)."""

    def __init__(self):
        self.wc = ops.WeightCache()
        self._arena = None
        self._retired = []
        self._off = 0

    def begin(self, B, device):
        need = B * 16384
        if self._arena is None or self._arena.numel() < need or self._arena.device != device:
            if self._arena is not None:
                self._retired.append(self._arena)        # a captured graph may still address it (see unet_exec.begin)
            self._arena = torch.zeros(need, dtype=torch.float32, device=device)
        else:
            self._arena.zero_()
        self._off = 0

    def gs(self, B, HW, C):
        """(partial-sum buffer, channels per block) for an activation a GEMM epilogue produces: blocks of 4 channels, or of 2
        for the 64-channel levels (32 groups of 2 channels)."""
        if B * HW < 4096 or HW % 16 or C % 64:
            return None
        cb = 4 if C % 128 == 0 else 2
        n = B * (C // cb) * 2
        if self._off + n > self._arena.numel():
            return None
        v = self._arena[self._off:self._off + n]
        self._off += n
        return v, cb

    def stats(self, x, gn, B):
        HW = x.H * x.W
        if x.gs is not None:
            return ops.groupnorm_finalize_cb(x.gs[0], x.C, B, HW, gn.num_groups, gn.eps, x.gs[1])
        return ops.groupnorm_stats(x.t, None, B, HW, gn.num_groups, gn.eps)

    def norm(self, x, gn, B, silu):
        HW = x.H * x.W
        if x.gs is not None and x.gs[1] == 4:
            return ops.groupnorm_pack_fused(x.t, None, gn.weight, gn.bias, B, HW, gn.num_groups, gn.eps, silu, gsum1=x.gs[0])
        return ops.groupnorm_pack_fused(x.t, None, gn.weight, gn.bias, B, HW, gn.num_groups, gn.eps, silu,
                                        stats=self.stats(x, gn, B))

    def conv3(self, p, conv, key, B, H, W, Cin, residual=None, mode=SDB_A_CONV3):
        Cout = conv.weight.shape[0]
        gs = self.gs(B, H * W, Cout)
        y = ops.gemm(p, self.wc.conv3(key, conv.weight), bias=conv.bias, residual=residual, conv=(mode, B, H, W, Cin),
                     gsum=gs[0] if gs else None, gsum_cb=gs[1] if gs else 4, rows_per_group=H * W)
        return Act(y, H, W, Cout, gs)

    def conv1(self, p, conv, key, residual=None):
        return ops.gemm(p, self.wc.linear(key, conv.weight), bias=conv.bias, residual=residual)

    def resblock(self, m, x, B):
        """modules.py:99-116: x + conv2(swish(GN(conv1(swish(GN(x)))))), 1x1 / 3x3 shortcut when the width changes
        (dropout is 0 in every shipped config and the VQ-VAE runs in eval mode)."""
        key, H, W = id(m), x.H, x.W
        h = self.conv3(self.norm(x, m.norm1, B, True), m.conv1, (key, 'c1'), B, H, W, m.in_channels)
        if m.in_channels == m.out_channels:
            xs = x.t
        elif m.use_conv_shortcut:
            xs = self.conv3(ops.pack_nhwc(x.t, None, B, H, W, SDB_PACK_PLAIN), m.conv_shortcut, (key, 'cs'), B, H, W,
                            m.in_channels).t
        else:
            xs = self.conv1(ops.pack_nhwc(x.t, None, B, H, W, SDB_PACK_PLAIN), m.nin_shortcut, (key, 'nin'))
        return self.conv3(self.norm(h, m.norm2, B, True), m.conv2, (key, 'c2'), B, H, W, m.out_channels, residual=xs)

    def attn(self, m, x, B):
        """modules.py:130-153: single-head attention over the H*W tokens with head dim C.  q | k | v in one GEMM over the
        whole batch, then S = q k^T, softmax(S C^-0.5) and O = P v as tcgen05 GEMMs."""
        if not isinstance(m, AttnBlock):
            return x
        key, C, L = id(m), x.C, x.H * x.W
        pn = self.norm(x, m.norm, B, False)
        wqkv = self.wc.linear((key, 'qkv'), m.q.weight, m.k.weight, m.v.weight)
        qkv = ops.gemm(pn, wqkv, bias=self.wc.cat((key, 'qkvb'), m.q.bias, m.k.bias, m.v.bias))
        scale = float(C) ** -0.5
        if L % 256 == 0:
            # all samples in ONE launch per product: block-diagonal tcgen05 GEMMs (sdb200.h SdbGemm.batch_rows) --
            # S_b = q_b k_b^T (the W operand advances L rows per sample), O_b = P_b v_b with v^T stored [C, B*L] (the W
            # operand advances L columns per sample)
            qp, kp = ops.pack_rows(qkv[:, :C]), ops.pack_rows(qkv[:, C:2 * C])
            vt = ops.transpose_packed(ops.pack_rows(qkv[:, 2 * C:]))                 # [C, B*L]
            s = ops.gemm(qp, kp, batch=(L, L, L, 0))                                 # [B*L, L]
            o = ops.gemm(ops.softmax_pack(s, scale), vt, batch=(L, C, 0, L), alpha=1.0 / ops.SOFTMAX_PACK_SCALE)   # [B*L, C]
        else:       # ragged token counts: one sample at a time, the other sample's packed rows as the 'weight' operand
            qp, kp = ops.pack_rows(qkv[:, :C]), ops.pack_rows(qkv[:, C:2 * C])
            o = torch.empty(B * L, C, dtype=torch.float32, device=x.t.device)
            s = torch.empty(L, L, dtype=torch.float32, device=x.t.device)
            for b in range(B):
                r0, r1 = b * L, (b + 1) * L
                ops.gemm(qp.row_range(r0, r1), kp.row_range(r0, r1), out=s)
                vt = ops.transpose_packed(ops.pack_rows(qkv[r0:r1, 2 * C:]))         # [C, L]
                ops.gemm(ops.softmax_pack(s, scale), vt, out=o[r0:r1], alpha=1.0 / ops.SOFTMAX_PACK_SCALE)
        y = self.conv1(ops.pack_rows(o), m.proj_out, (key, 'po'), residual=x.t)
        return Act(y, x.H, x.W, C, None)

    def head(self, x, gn, conv, B):
        """norm_out -> swish -> conv_out (3 output channels: too thin for tensor tiles) -> NCHW"""
        return ops.conv3_out(x.t, self.stats(x, gn, B), gn.weight, gn.bias, conv.weight, conv.bias, B, x.H, x.W, gn.num_groups)


def _check_input(mod, x):
    if not x.is_cuda:
        raise RuntimeError(f'slotdiffusion_b200.vqvae.{type(mod).__name__} runs on CUDA (sm_100a) only; no CPU fallback')
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters())):
        raise RuntimeError(f'slotdiffusion_b200.vqvae.{type(mod).__name__} is the FROZEN first stage of the LDM (inference '
                           'only): call it under torch.no_grad() with requires_grad_(False) parameters (VQVAEWrapper does, '
                           'VQVAE.py:172-176); training the VQ-VAE itself stays the reference PyTorch code')


class _ExecMixin:
    def __getstate__(self):                       # deepcopy / pickle: the schedule and its packed weights are derived state
        d = self.__dict__.copy()
        d.pop('_ex', None)
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self._ex = _Exec()

    def invalidate_caches(self):
        self._ex.wc.clear()


class Encoder(_ExecMixin, nn.Module):
    """modules.py:168-262 (same constructor keywords)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=False, attn_type='vanilla',
                 **ignore_kwargs):
        super().__init__()
        if not resamp_with_conv:
            raise NotImplementedError('slotdiffusion_b200.vqvae: resamp_with_conv=False (avg-pool resampling) is not built')
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)
        self._ex = _Exec()

    def forward(self, x):
        """x [B, in_channels, H, W] -> pre-quantisation features [B, z_channels, H / 2^(levels-1), W / 2^(levels-1)]."""
        _check_input(self, x)
        with torch.no_grad(), ops.pack_format(ops.SDB_FMT_F16X2):
            ex = self._ex
            B, _, H, W = x.shape
            ex.begin(B, x.device)
            h = Act(ops.conv3_in(x.float(), self.conv_in.weight, self.conv_in.bias), H, W, self.ch)      # modules.py:239
            for i_level, lvl in enumerate(self.down):
                for i_block, blk in enumerate(lvl.block):
                    h = ex.resblock(blk, h, B)
                    if len(lvl.attn) > 0:
                        h = ex.attn(lvl.attn[i_block], h, B)
                if i_level != self.num_resolutions - 1:
                    # F.pad (0,1,0,1) + conv(stride 2, padding 0), modules.py:44-48: phase split + asymmetric-tap implicit GEMM
                    xp = ops.pack_nhwc(h.t, None, B, h.H, h.W, SDB_PACK_PHASE2)
                    h = ex.conv3(xp, lvl.downsample.conv, (id(lvl), 'down'), B, h.H // 2, h.W // 2, h.C, mode=SDB_A_CONV3S2A)
            h = ex.resblock(self.mid.block_1, h, B)
            h = ex.attn(self.mid.attn_1, h, B)
            h = ex.resblock(self.mid.block_2, h, B)
            return ex.head(h, self.norm_out, self.conv_out, B)


class Decoder(_ExecMixin, nn.Module):
    """modules.py:265-362 (same constructor keywords)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, attn_type='vanilla', **ignorekwargs):
        super().__init__()
        if not resamp_with_conv:
            raise NotImplementedError('slotdiffusion_b200.vqvae: resamp_with_conv=False is not built')
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self._ex = _Exec()

    def forward(self, z):
        """z [B, z_channels, h, w] (post_quant_conv output) -> image [B, out_ch, h * 2^(levels-1), w * 2^(levels-1)]."""
        _check_input(self, z)
        self.last_z_shape = z.shape
        with torch.no_grad(), ops.pack_format(ops.SDB_FMT_F16X2):
            ex = self._ex
            B, _, H, W = z.shape
            ex.begin(B, z.device)
            h = Act(ops.conv3_in(z.float(), self.conv_in.weight, self.conv_in.bias), H, W, self.conv_in.weight.shape[0])
            h = ex.resblock(self.mid.block_1, h, B)
            h = ex.attn(self.mid.attn_1, h, B)
            h = ex.resblock(self.mid.block_2, h, B)
            for i_level in reversed(range(self.num_resolutions)):
                lvl = self.up[i_level]
                for i_block, blk in enumerate(lvl.block):
                    h = ex.resblock(blk, h, B)
                    if len(lvl.attn) > 0:
                        h = ex.attn(lvl.attn[i_block], h, B)
                if i_level != 0:                                                    # nearest x2 + conv3x3, modules.py:27-31
                    xp = ops.pack_nhwc(h.t, None, B, h.H, h.W, SDB_PACK_UP2)
                    h = ex.conv3(xp, lvl.upsample.conv, (id(lvl), 'up'), B, 2 * h.H, 2 * h.W, h.C)
            return ex.head(h, self.norm_out, self.conv_out, B)
