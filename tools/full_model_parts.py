"""The reference's modules AROUND the hot path, as plain PyTorch (cuDNN / cuBLAS through ATen) -- NOT product code.

SURVEY 8d asks for the train-step metric "of the full model with the new modules dropped in".  On the GPU box the
reference package is absent, so the parts of the full model that are out of scope for this repository
are restated here the way the reference runs them -- eager PyTorch -- for tools/full_model_train_bench.py only:

  * ImageEncoder: ResNet18 with GroupNorm(32), stride-1 stem, no layer4  (video_based/models/resnet.py:150-315,
    img_based/models/slot_attention.py:182-194), SoftPositionEmbed (utils.py:37-63), LayerNorm + 2-layer MLP
    (slot_attention.py:238-245, :294-304)  ->  [B, (H/4)(W/4), enc_out_channels]
  * VQVAEEncoder: taming-style Encoder (vqvae/modules.py:168-262: GroupNorm(32, eps 1e-6) + swish ResnetBlocks, asymmetric-
    pad stride-2 Downsample, single-head AttnBlock in the middle) + quant_conv (VQVAE.py:94-100)  ->  latents [B, 3, H/4, W/4]

Module / parameter names equal the reference's (so a reference state_dict loads with strict=True, which is how
tests/test_full_model_parts_cpu.py checks the arithmetic against the reference itself); the code is this repository's own.
"""
import torch
import torch.nn.functional as F
from torch import nn


def _gn(c, eps=1e-5):
    return nn.GroupNorm(32, c, eps=eps)


class _BasicBlock(nn.Module):
    """conv3x3 -> GN -> ReLU -> conv3x3 -> GN -> (+ identity | 1x1-conv+GN shortcut) -> ReLU   (resnet.py:37-88)"""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = _gn(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = _gn(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), _gn(cout))

    def forward(self, x):
        y = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _ResNet18GN(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)       # small_inputs: 3x3 stride-1 stem, no max-pool
        self.bn1 = _gn(64)
        self.layer1 = nn.Sequential(_BasicBlock(64, 64, 1), _BasicBlock(64, 64, 1))
        self.layer2 = nn.Sequential(_BasicBlock(64, 128, 2), _BasicBlock(128, 128, 1))
        self.layer3 = nn.Sequential(_BasicBlock(128, 256, 2), _BasicBlock(256, 256, 1))

    def forward(self, x):
        return self.layer3(self.layer2(self.layer1(F.relu(self.bn1(self.conv1(x))))))


class _SoftPositionEmbed(nn.Module):
    def __init__(self, channels, resolution):
        super().__init__()
        self.dense = nn.Linear(4, channels)
        ys, xs = torch.meshgrid(torch.linspace(0, 1, resolution[0]), torch.linspace(0, 1, resolution[1]), indexing='ij')
        g = torch.stack([ys, xs], -1)[None]
        self.register_buffer('grid', torch.cat([g, 1 - g], -1))     # [1, H, W, 4]

    def forward(self, x):
        return x + self.dense(self.grid).permute(0, 3, 1, 2)


class ImageEncoder(nn.Module):
    """img [B,3,H,W] in [-1,1] -> Slot-Attention inputs [B, (H/4)*(W/4), out_channels]."""

    def __init__(self, resolution=(128, 128), out_channels=192):
        super().__init__()
        vis = (resolution[0] // 4, resolution[1] // 4)
        self.encoder = _ResNet18GN()
        self.encoder_pos_embedding = _SoftPositionEmbed(256, vis)
        self.encoder_out_layer = nn.Sequential(nn.LayerNorm(256), nn.Linear(256, out_channels), nn.ReLU(),
                                               nn.Linear(out_channels, out_channels))

    def forward(self, img):
        f = self.encoder_pos_embedding(self.encoder(img))
        return self.encoder_out_layer(f.flatten(2).transpose(1, 2).contiguous())


def _swish(x):
    return x * torch.sigmoid(x)


class _VResBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = _gn(cin, 1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2 = _gn(cout, 1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)
        self.has_shortcut = cin != cout

    def forward(self, x):
        h = self.conv2(_swish(self.norm2(self.conv1(_swish(self.norm1(x))))))
        return (self.nin_shortcut(x) if self.has_shortcut else x) + h


class _VAttn(nn.Module):
    """single-head attention over the h*w positions with 1x1-conv projections (modules.py:113-154)"""

    def __init__(self, c):
        super().__init__()
        self.norm = _gn(c, 1e-6)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        B, C, H, W = x.shape
        n = self.norm(x)
        q, k, v = (m(n).flatten(2).transpose(1, 2) for m in (self.q, self.k, self.v))     # [B, HW, C]
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]        # scale C^-0.5
        return x + self.proj_out(o.transpose(1, 2).reshape(B, C, H, W))


class _VDown(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 2, 0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))                  # asymmetric padding, modules.py:45-48


class _Level(nn.Module):
    def __init__(self, cin, cout, nblocks, down):
        super().__init__()
        self.block = nn.ModuleList([_VResBlock(cin if i == 0 else cout, cout) for i in range(nblocks)])
        self.attn = nn.ModuleList()
        if down:
            self.downsample = _VDown(cout)
        self.has_down = down

    def forward(self, x):
        for b in self.block:
            x = b(x)
        return self.downsample(x) if self.has_down else x


class _Mid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.block_1 = _VResBlock(c, c)
        self.attn_1 = _VAttn(c)
        self.block_2 = _VResBlock(c, c)

    def forward(self, x):
        return self.block_2(self.attn_1(self.block_1(x)))


class _VEncoder(nn.Module):
    def __init__(self, ch=64, ch_mult=(1, 2, 4), num_res_blocks=2, z_channels=3):
        super().__init__()
        self.conv_in = nn.Conv2d(3, ch, 3, 1, 1)
        chans = [ch] + [ch * m for m in ch_mult]
        self.down = nn.ModuleList([_Level(chans[i], chans[i + 1], num_res_blocks, i != len(ch_mult) - 1)
                                   for i in range(len(ch_mult))])
        self.mid = _Mid(chans[-1])
        self.norm_out = _gn(chans[-1], 1e-6)
        self.conv_out = nn.Conv2d(chans[-1], z_channels, 3, 1, 1)

    def forward(self, x):
        h = self.conv_in(x)
        for lvl in self.down:
            h = lvl(h)
        return self.conv_out(_swish(self.norm_out(self.mid(h))))


class VQVAEEncoder(nn.Module):
    """img [B,3,H,W] -> pre-quantisation latents x0 [B, 3, H/4, W/4] (VQVAE.encode, VQVAE.py:94-100); frozen in LDM training."""

    def __init__(self, ch=64, ch_mult=(1, 2, 4), num_res_blocks=2, z_channels=3, embed_dim=3):
        super().__init__()
        self.encoder = _VEncoder(ch, ch_mult, num_res_blocks, z_channels)
        self.quant_conv = nn.Conv2d(z_channels, embed_dim, 1)

    def forward(self, img):
        return self.quant_conv(self.encoder(img))
