#!/usr/bin/env python
"""Debug helper: run one UNet evaluation in the three-pass and in the fp8-corrected operand format, record the result of
every ops.* call in order, and print the first calls whose results diverge (a format bug is O(1), the format's own error
is ~3e-5)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import unet_ref  # noqa: E402
from slotdiffusion_b200 import ops  # noqa: E402
from slotdiffusion_b200.unet import UNetModel  # noqa: E402

NAMES = ['gemm', 'pack_rows', 'layernorm_pack', 'groupnorm_pack_fused', 'pack_nhwc', 'attention_pack', 'geglu_pack',
         'timestep_embedding_pack', 'conv3_in', 'conv3_out', 'groupnorm_stats', 'groupnorm_finalize', 'channel_block_sums']


def val(o):
    if isinstance(o, ops.Packed):
        return o.unpack().double()
    if torch.is_tensor(o):
        return o.double().clone()
    return None


def trace(net, x, t, ctx, mode):
    ops.set_precision(mode)
    rec = []
    saved = {n: getattr(ops, n) for n in NAMES}

    def wrap(n, f):
        def g(*a, **k):
            out = f(*a, **k)
            outs = out if isinstance(out, tuple) else (out,)
            desc = n
            if n == 'gemm':
                A, W = a[0], a[1]
                desc = 'gemm A[%d,%d] W[%d,%d] conv=%s %s' % (A.rows, A.K, W.rows, W.K, k.get('conv'),
                                                              [kk for kk in k if k[kk] is not None and kk != 'conv'])
            for o in outs:
                v = val(o)
                if v is not None:
                    rec.append((desc, v))
            return out
        return g
    for n in NAMES:
        setattr(ops, n, wrap(n, saved[n]))
    try:
        with torch.no_grad():
            y = net(x, t, context=ctx)
    finally:
        for n in NAMES:
            setattr(ops, n, saved[n])
    return y, rec


def main():
    small = os.environ.get('SMALL', '1') == '1'
    over = dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1, context_dim=64) if small else {}
    cfg = dict(unet_ref.DEFAULT_CFG, **over)
    sd = unet_ref.random_state_dict(cfg, seed=31)
    net = UNetModel(dropout=0.0, dims=2, use_checkpoint=False, resblock_updown=False, conv_resample=True,
                    transformer_depth=1, n_embed=None, **cfg).cuda().eval()
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    B, R = 2, (16 if small else 32)
    x = torch.randn(B, 3, R, R, generator=g).cuda()
    ctx = torch.randn(B, 5, cfg['context_dim'], generator=g).cuda()
    t = torch.tensor([7.0, 503.0]).cuda()
    y3, r3 = trace(net, x, t, ctx, 'fp32')
    y2, r2 = trace(net, x, t, ctx, 'fp8c')
    print('final rel', ((y2 - y3).norm() / y3.norm()).item(), 'calls', len(r3), len(r2))
    shown = 0
    for i, ((d3, v3), (d2, v2)) in enumerate(zip(r3, r2)):
        if v3.shape != v2.shape:
            print(i, 'SHAPE', d3, tuple(v3.shape), tuple(v2.shape))
            break
        e = ((v3 - v2).norm() / v3.norm().clamp_min(1e-30)).item()
        if e > 1e-3 or i < 3:
            print('%4d rel %.3e  %s' % (i, e, d3))
            shown += e > 1e-3
            if shown >= 6:
                break


if __name__ == '__main__':
    main()
