"""Oracle: slot-conditioned LDM UNet denoiser (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional restatement over a reference-keyed state_dict of
  * UNetModel.forward            video_based/models/unet/unet.py:551-576 (structure :344-549)
  * ResBlock._forward            unet.py:271-285
  * Downsample / Upsample        unet.py:140-179 / :84-121
  * SpatialTransformer.forward   video_based/models/unet/attention.py:297-308
  * BasicTransformerBlock        attention.py:247-251
  * CrossAttention.forward       attention.py:182-206
  * GEGLU / FeedForward          attention.py:39-65
  * timestep_embedding           video_based/models/unet/utils.py:70-92
  * GroupNorm32 (eps 1e-5) / Normalize (eps 1e-6)   utils.py:120-139 / attention.py:77-79
(all under /root/reference/slotdiffusion/).  Eval-mode semantics (Dropout off).
Works in the dtype of `x` (fp32 or fp64).
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(  # img_based/configs/sa_ldm/sa_ldm_clevrtex_params-res128.py:79-95
    in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
    attention_resolutions=(8, 4, 2), channel_mult=(1, 2, 3, 4),
    num_head_channels=32, context_dim=192)


def timestep_embedding(t, dim, max_period=10000.0):
    # utils.py:79-86: freqs in fp32, [cos, sin] order; t may be fractional
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def block_plan(cfg):
    """Enumerate the module tree the way unet.py:407-538 builds it.

    Returns (input_blocks, middle, output_blocks); each block is a list of
    ('res', cin, cout) | ('attn', c) | ('down', c) | ('up', c) | ('conv_in',).
    """
    mc, mult, nrb = cfg['model_channels'], cfg['channel_mult'], cfg['num_res_blocks']
    attn_res = set(cfg['attention_resolutions'])
    inp = [[('conv_in',)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            blk = [('res', ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                blk.append(('attn', ch))
            inp.append(blk)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([('down', ch)])
            chans.append(ch)
            ds *= 2
    mid = [('res', ch, ch), ('attn', ch), ('res', ch, ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            blk = [('res', ch + ich, m * mc)]
            ch = m * mc
            if ds in attn_res:
                blk.append(('attn', ch))
            if level and i == nrb:
                blk.append(('up', ch))
                ds //= 2
            out.append(blk)
    return inp, mid, out


class _W:
    """state_dict view with a key prefix and dtype cast."""

    def __init__(self, sd, prefix, dtype):
        self.sd, self.prefix, self.dtype = sd, prefix, dtype

    def __call__(self, name):
        return self.sd[self.prefix + name].to(self.dtype)

    def has(self, name):
        return (self.prefix + name) in self.sd

    def sub(self, name):
        return _W(self.sd, self.prefix + name, self.dtype)


def _gn(x, w, name, eps):
    return F.group_norm(x, 32, w(name + '.weight'), w(name + '.bias'), eps)


def res_block(w, x, emb):
    # unet.py:271-285 (no up/down inside ResBlock: resblock_updown=False everywhere)
    h = F.conv2d(F.silu(_gn(x, w, 'in_layers.0', 1e-5)), w('in_layers.2.weight'), w('in_layers.2.bias'), padding=1)
    e = F.linear(F.silu(emb), w('emb_layers.1.weight'), w('emb_layers.1.bias'))
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(h, w, 'out_layers.0', 1e-5)), w('out_layers.3.weight'), w('out_layers.3.bias'), padding=1)
    if w.has('skip_connection.weight'):
        x = F.conv2d(x, w('skip_connection.weight'), w('skip_connection.bias'))
    return x + h


def attention(w, x, ctx, heads):
    # attention.py:182-206; ctx=None -> self attention
    ctx = x if ctx is None else ctx
    q = F.linear(x, w('to_q.weight'))
    k = F.linear(ctx, w('to_k.weight'))
    v = F.linear(ctx, w('to_v.weight'))
    B, L, C = q.shape
    d = C // heads

    def split(t):
        return t.reshape(B, -1, heads, d).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum('bhid,bhjd->bhij', q, k) * (float(d) ** -0.5)
    p = torch.softmax(sim, dim=-1)
    o = torch.einsum('bhij,bhjd->bhid', p, v).permute(0, 2, 1, 3).reshape(B, L, C)
    return F.linear(o, w('to_out.0.weight'), w('to_out.0.bias'))


def transformer_block(w, x, ctx, heads):
    # attention.py:247-251 (pre-LN, eps 1e-5); GEGLU with exact erf GELU (attention.py:46-48)
    C = x.shape[-1]

    def ln(t, n):
        return F.layer_norm(t, (C,), w(n + '.weight'), w(n + '.bias'), 1e-5)
    x = attention(w.sub('attn1.'), ln(x, 'norm1'), None, heads) + x
    x = attention(w.sub('attn2.'), ln(x, 'norm2'), ctx, heads) + x
    u = F.linear(ln(x, 'norm3'), w('ff.net.0.proj.weight'), w('ff.net.0.proj.bias'))
    a, g = u.chunk(2, dim=-1)
    x = F.linear(a * F.gelu(g), w('ff.net.2.weight'), w('ff.net.2.bias')) + x
    return x


def spatial_transformer(w, x, ctx, head_ch=32):
    # attention.py:297-308
    B, C, H, W = x.shape
    h = F.conv2d(_gn(x, w, 'norm', 1e-6), w('proj_in.weight'), w('proj_in.bias'))
    h = h.flatten(2).transpose(1, 2)                     # b (h w) c
    h = transformer_block(w.sub('transformer_blocks.0.'), h, ctx, C // head_ch)
    h = h.transpose(1, 2).reshape(B, C, H, W)
    h = F.conv2d(h, w('proj_out.weight'), w('proj_out.bias'))
    return h + x


def _run_block(w, blk, h, emb, ctx, head_ch):
    for j, layer in enumerate(blk):
        lw = w.sub(f'{j}.')
        kind = layer[0]
        if kind == 'conv_in':
            h = F.conv2d(h, lw('weight'), lw('bias'), padding=1)
        elif kind == 'res':
            h = res_block(lw, h, emb)
        elif kind == 'attn':
            h = spatial_transformer(lw, h, ctx, head_ch)
        elif kind == 'down':
            h = F.conv2d(h, lw('op.weight'), lw('op.bias'), stride=2, padding=1)     # unet.py:140-179
        elif kind == 'up':
            h = F.interpolate(h, scale_factor=2, mode='nearest')                      # unet.py:118
            h = F.conv2d(h, lw('conv.weight'), lw('conv.bias'), padding=1)
    return h


def unet_forward(sd, x, t, context, cfg=None, prefix=''):
    """x [B,Cin,h,w], t [B] (int or float), context [B,S,Dc] -> [B,Cout,h,w]."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    dt = x.dtype
    w = _W(sd, prefix, dt)
    inp, mid, out = block_plan(cfg)
    head_ch = cfg['num_head_channels']
    context = context.to(dt)
    temb = timestep_embedding(t, cfg['model_channels']).to(dt)            # unet.py:560-561
    emb = F.linear(F.silu(F.linear(temb, w('time_embed.0.weight'), w('time_embed.0.bias'))),
                   w('time_embed.2.weight'), w('time_embed.2.bias'))      # unet.py:399-404
    hs = []
    h = x
    for i, blk in enumerate(inp):                                         # unet.py:566-568
        h = _run_block(w.sub(f'input_blocks.{i}.'), blk, h, emb, context, head_ch)
        hs.append(h)
    h = _run_block(w.sub('middle_block.'), mid, h, emb, context, head_ch)
    for i, blk in enumerate(out):                                         # unet.py:570-572
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(w.sub(f'output_blocks.{i}.'), blk, h, emb, context, head_ch)
    h = F.silu(_gn(h, w, 'out.0', 1e-5))
    return F.conv2d(h, w('out.2.weight'), w('out.2.bias'), padding=1)     # unet.py:537-542


def random_state_dict(cfg=None, seed=0, zero_init_std=0.02):
    """Random UNet weights with reference key names/shapes (torch default-like init;
    the reference's zero-initialised tensors are drawn from N(0, zero_init_std) so
    that outputs are non-trivial, SURVEY.md section 7 'zero-init layers')."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True, std=None):
        b = in_f ** -0.5
        if std is None:
            sd[name + '.weight'] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        else:
            sd[name + '.weight'] = torch.randn(out_f, in_f, generator=g) * std
        if bias:
            sd[name + '.bias'] = (torch.rand(out_f, generator=g) * 2 - 1) * b

    def conv(name, cout, cin, k, std=None):
        b = (cin * k * k) ** -0.5
        if std is None:
            sd[name + '.weight'] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * b
        else:
            sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * std
        sd[name + '.bias'] = (torch.rand(cout, generator=g) * 2 - 1) * b

    def norm(name, c):
        sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g)

    mc, ted, dc = cfg['model_channels'], cfg['model_channels'] * 4, cfg['context_dim']
    lin('time_embed.0', ted, mc)
    lin('time_embed.2', ted, ted)

    def make(prefix, blk):
        for j, layer in enumerate(blk):
            p = f'{prefix}{j}.'
            if layer[0] == 'conv_in':
                conv(p[:-1], mc, cfg['in_channels'], 3)
            elif layer[0] == 'res':
                _, cin, cout = layer
                norm(p + 'in_layers.0', cin)
                conv(p + 'in_layers.2', cout, cin, 3)
                lin(p + 'emb_layers.1', cout, ted)
                norm(p + 'out_layers.0', cout)
                conv(p + 'out_layers.3', cout, cout, 3, std=zero_init_std)
                if cin != cout:
                    conv(p + 'skip_connection', cout, cin, 1)
            elif layer[0] == 'attn':
                c = layer[1]
                norm(p + 'norm', c)
                conv(p + 'proj_in', c, c, 1)
                t = p + 'transformer_blocks.0.'
                for a, kd in (('attn1', c), ('attn2', dc)):
                    lin(t + a + '.to_q', c, c, bias=False)
                    lin(t + a + '.to_k', c, kd, bias=False)
                    lin(t + a + '.to_v', c, kd, bias=False)
                    lin(t + a + '.to_out.0', c, c)
                lin(t + 'ff.net.0.proj', 8 * c, c)
                lin(t + 'ff.net.2', c, 4 * c)
                for n in ('norm1', 'norm2', 'norm3'):
                    norm(t + n, c)
                conv(p + 'proj_out', c, c, 1, std=zero_init_std)
            elif layer[0] == 'down':
                conv(p + 'op', layer[1], layer[1], 3)
            elif layer[0] == 'up':
                conv(p + 'conv', layer[1], layer[1], 3)

    inp, mid, out = block_plan(cfg)
    for i, blk in enumerate(inp):
        make(f'input_blocks.{i}.', blk)
    make('middle_block.', mid)
    for i, blk in enumerate(out):
        make(f'output_blocks.{i}.', blk)
    norm('out.0', mc)
    conv('out.2', cfg['out_channels'], mc, 3, std=zero_init_std)
    return sd
