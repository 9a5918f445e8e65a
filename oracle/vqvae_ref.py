"""Oracle: VQ-VAE encoder / decoder of the LDM first stage (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional restatement (dict of tensors keyed like the reference state_dict, plain torch ops, any dtype) of
  Encoder.forward   /root/reference/slotdiffusion/video_based/models/vqvae/modules.py:236-262
  Decoder.forward   modules.py:338-362
  ResnetBlock       modules.py:99-116      AttnBlock  modules.py:130-153
  Downsample        modules.py:44-52 (asymmetric pad (0,1,0,1), conv stride 2)      Upsample  modules.py:27-31
Pinned to the unmodified reference by tests/golden/vqvae.npz (tools/make_golden.py vqvae; tests/test_oracle_golden.py).
"""
import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(resolution=128, in_channels=3, z_channels=3, ch=64, ch_mult=(1, 2, 4), num_res_blocks=2,
                   attn_resolutions=(), out_ch=3)        # sa_ldm_clevrtex_params-res128.py:59-69


def _gn(x, sd, name):
    return F.group_norm(x, 32, sd[name + '.weight'], sd[name + '.bias'], 1e-6)       # Normalize, modules.py:12-14


def _swish(x):
    return x * torch.sigmoid(x)                                                        # modules.py:7-9


def _conv(x, sd, name, stride=1, padding=1):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], stride=stride, padding=padding)


def resnet_block(sd, pre, x):
    h = _conv(_swish(_gn(x, sd, pre + 'norm1')), sd, pre + 'conv1')
    h = _conv(_swish(_gn(h, sd, pre + 'norm2')), sd, pre + 'conv2')
    if pre + 'nin_shortcut.weight' in sd:
        x = _conv(x, sd, pre + 'nin_shortcut', padding=0)
    elif pre + 'conv_shortcut.weight' in sd:
        x = _conv(x, sd, pre + 'conv_shortcut')
    return x + h


def attn_block(sd, pre, x):
    h = _gn(x, sd, pre + 'norm')
    q, k, v = (_conv(h, sd, pre + n, padding=0) for n in 'qkv')
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    o = torch.bmm(v.reshape(b, c, hh * ww), w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(o, sd, pre + 'proj_out', padding=0)


def _mid(sd, h):
    h = resnet_block(sd, 'mid.block_1.', h)
    if 'mid.attn_1.norm.weight' in sd:
        h = attn_block(sd, 'mid.attn_1.', h)
    return resnet_block(sd, 'mid.block_2.', h)


def encoder_forward(sd, x, cfg=None):
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    sd = {k: v.to(x.dtype) for k, v in sd.items()}
    n_res = len(cfg['ch_mult'])
    h = _conv(x, sd, 'conv_in')
    for i in range(n_res):
        for j in range(cfg['num_res_blocks']):
            h = resnet_block(sd, f'down.{i}.block.{j}.', h)
            if f'down.{i}.attn.{j}.norm.weight' in sd:
                h = attn_block(sd, f'down.{i}.attn.{j}.', h)
        if i != n_res - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1)), sd, f'down.{i}.downsample.conv', stride=2, padding=0)
    h = _mid(sd, h)
    return _conv(_swish(_gn(h, sd, 'norm_out')), sd, 'conv_out')


def decoder_forward(sd, z, cfg=None):
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    sd = {k: v.to(z.dtype) for k, v in sd.items()}
    n_res = len(cfg['ch_mult'])
    h = _mid(sd, _conv(z, sd, 'conv_in'))
    for i in reversed(range(n_res)):
        for j in range(cfg['num_res_blocks'] + 1):
            h = resnet_block(sd, f'up.{i}.block.{j}.', h)
            if f'up.{i}.attn.{j}.norm.weight' in sd:
                h = attn_block(sd, f'up.{i}.attn.{j}.', h)
        if i != 0:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode='nearest'), sd, f'up.{i}.upsample.conv')
    return _conv(_swish(_gn(h, sd, 'norm_out')), sd, 'conv_out')


def random_state_dicts(cfg=None, seed=0):
    """(encoder state_dict, decoder state_dict) with nn-default-like init and non-trivial GroupNorm affines, keyed and
    shaped like the reference modules built from `cfg` (shapes come from a throw-away instance of OUR parameter holders,
    whose layout tests/test_dropin_cpu.py pins to the reference)."""
    from slotdiffusion_b200 import vqvae
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    kw = dict(cfg, dropout=0.0)
    g = torch.Generator().manual_seed(seed)
    out = []
    for cls in (vqvae.Encoder, vqvae.Decoder):
        sd = {}
        for k, v in cls(**kw).state_dict().items():
            if k.endswith('.weight') and v.dim() == 1:
                sd[k] = 1 + 0.1 * torch.randn(v.shape, generator=g)
            elif v.dim() == 1:
                sd[k] = 0.1 * torch.randn(v.shape, generator=g)
            else:
                fan_in = v[0].numel()
                sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) * fan_in ** -0.5
        out.append(sd)
    return tuple(out)
