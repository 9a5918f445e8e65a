"""Drop-in slot-conditioned LDM UNet denoiser backed by libsdb200 (hand-written sm_100a kernels).

Mirrors the reference module tree -- constructor arguments, attribute / parameter names and shapes, so
`unet_dict` configs and reference checkpoints (`dm_decoder.model.diffusion_model.*`) load unchanged:
  * UNetModel / ResBlock / Downsample / Upsample / TimestepEmbedSequential
        /root/reference/slotdiffusion/video_based/models/unet/unet.py:344-584, :182-285, :140-179, :84-121, :67-81
  * SpatialTransformer / BasicTransformerBlock / CrossAttention / FeedForward / GEGLU
        video_based/models/unet/attention.py:254-308, :209-251, :157-206, :51-65, :39-48
  * GroupNorm32 (eps 1e-5), Normalize (eps 1e-6), zero_module, timestep_embedding
        video_based/models/unet/utils.py:120-139, attention.py:77-79, utils.py:95-102, :70-92
The nn.Module objects only own parameters (reference initialisation included); forward() executes the
kernel schedule in unet_exec.py.  Internal activation layout is NHWC fp32; GEMM operands are split-fp16.
"""
import math

import torch
from torch import nn


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    """Parameter holder (32 groups, eps 1e-5, computed in fp32 by the kernels)."""


def normalization(channels):
    return GroupNorm32(32, channels)


class TimestepBlock(nn.Module):
    pass


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    pass


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert dims == 2 and use_conv, 'only conv_resample=True, dims=2 (all shipped configs)'
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=padding)


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert dims == 2 and use_conv, 'only conv_resample=True, dims=2 (all shipped configs)'
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, dims=2,
                 use_checkpoint=False, up=False, down=False):
        super().__init__()
        assert dims == 2 and not up and not down and not use_conv, \
            'resblock_updown / use_conv skip are not used by any SlotDiffusion config'
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.updown = False
        self.in_layers = nn.Sequential(
            normalization(channels), nn.SiLU(), nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        assert glu, 'BasicTransformerBlock always uses gated_ff=True'
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, use_checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.use_checkpoint = use_checkpoint


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, use_checkpoint=False):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        assert inner_dim == in_channels
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim,
                                  use_checkpoint=use_checkpoint) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))


class UNetModel(nn.Module):
    """The full UNet with attention and timestep embedding (same ctor as the reference)."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, use_checkpoint=False, num_head_channels=32,
                 resblock_updown=False, transformer_depth=1, context_dim=None, n_embed=None):
        super().__init__()
        if dims != 2 or resblock_updown or n_embed is not None or not conv_resample:
            raise NotImplementedError('slotdiffusion_b200.UNetModel supports dims=2, conv_resample=True, '
                                      'resblock_updown=False, n_embed=None (every SlotDiffusion config)')
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.use_checkpoint = use_checkpoint
        self.num_head_channels = num_head_channels
        self.context_dim = context_dim
        self.predict_codebook_ids = False

        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))

        def res(cin, cout):
            return ResBlock(cin, ted, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint)

        def attn(ch):
            return SpatialTransformer(ch, ch // num_head_channels, num_head_channels, depth=transformer_depth,
                                      context_dim=context_dim, use_checkpoint=use_checkpoint)

        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        self._feature_size = model_channels
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(
                    TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                chans.append(ch)
                ds *= 2
                self._feature_size += ch
        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch), res(ch, ch))
        self._feature_size += ch
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [res(ch + ich, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
        self.out = nn.Sequential(
            normalization(ch), nn.SiLU(), zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

        from .unet_exec import UNetExecutor
        self._exec = UNetExecutor(self)

    # copy.deepcopy / pickle (EMA copies, torch.save(model)): the kernel schedule, its packed-weight cache and any
    # captured sampler graphs are derived state -- dropped here and rebuilt for the copy
    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_exec', None)
        d.pop('_sdb_samplers', None)
        d.pop('_sdb_graphs', None)
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        from .unet_exec import UNetExecutor
        self._exec = UNetExecutor(self)

    def invalidate_caches(self):
        """Drop packed weights and captured sampler graphs.  Parameter changes are detected through
        (data_ptr, _version); updates made through `p.data` (EMA swaps, some optimizers, manual weight surgery) bump
        neither -- call this after such an update."""
        self._exec.wc._c.clear()
        tr = getattr(self._exec, '_trainer', None)
        if tr is not None:
            tr.wc._c.clear()
        self.__dict__.pop('_sdb_samplers', None)

    def forward(self, x, timesteps=None, context=None, **kwargs):
        """x [N,C,h,w], timesteps [N] (int or fractional float), context [N,S,Dc] -> [N,C,h,w]."""
        if not x.is_cuda:
            raise RuntimeError('slotdiffusion_b200.UNetModel runs on CUDA (sm_100a) only; no CPU fallback')
        from . import graphed
        if graphed.enabled(self) and torch.is_grad_enabled() and context is not None and torch.is_tensor(timesteps) and (
                x.requires_grad or context.requires_grad or any(p.requires_grad for p in self.parameters())):
            # eager training loops: forward / backward schedules replayed from CUDA graphs (graphed.py)
            g = graphed.graphs_of(self, lambda a, t, c: self._exec(a, t, c), lambda: [self._exec.wc])
            return g(x, timesteps, context)
        return self._exec(x, timesteps, context)

    @property
    def device(self):
        return self.time_embed[0].weight.device

    @property
    def dtype(self):
        return self.time_embed[0].weight.dtype
