"""CPU experiment: can the two correction passes of the split-fp16 product run in FP8?

Today  A*W ~= A_h*W_h + A_l*W_h + A_h*W_l  with fp16 planes (3 tensor-core passes at the fp16 rate).
Idea   keep A_h*W_h in fp16 (kind::f16) and compute the two O(2^-11) correction terms with e4m3 operands
       (kind::f8f6f4, twice the rate, half the shared-memory bytes) into a second accumulator with power-of-two
       per-tensor scales:  C = D1 + 2^-s * D2,
       D2 = q8(A_l * 2^a) * q8(W_h * 2^b) + q8(A_h * 2^c) * q8(W_l * 2^d),   a + b = c + d = s.
       Cost: 1 + 0.5 + 0.5 = 2 pass-equivalents instead of 3.
This script emulates exactly that operand rounding inside the oracle UNet (conv2d / linear only) and prints the error
of one evaluation against the fp32 oracle, next to the 3-pass fp16 split and the reduced-pass modes.
Run: python tools/experiments/fp8_correction_numerics.py [batch]
"""
import math
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as TF
from oracle import unet_ref

F8 = torch.float8_e4m3fn


def h16(t):
    return t.half().float()


def q8(t, target=192.0):
    """e4m3 rounding with a power-of-two per-tensor scale that puts max|t| just below `target` (e4m3 max 448)."""
    m = t.abs().max().item()
    if m == 0:
        return t
    k = math.floor(math.log2(target / m))
    s = 2.0 ** k
    return (t * s).to(F8).float() / s


def q8s(t, k):
    """e4m3 rounding with a STATIC power-of-two scale 2^k (activations: no per-tensor max pass), saturating at +-448."""
    s = 2.0 ** k
    return (t * s).clamp(-448.0, 448.0).to(F8).float() / s


def shim(mode):
    ns = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith('__')})

    def prod(op, x, w, b, **kw):
        xh, wh = h16(x), h16(w)
        xl, wl = x - xh, w - wh
        if mode == 'split3':       # today's kernel (lo planes rounded to fp16 as well)
            y = op(xh, wh, None, **kw) + op(h16(xl), wh, None, **kw) + op(xh, h16(wl), None, **kw)
        elif mode == 'f8corr':
            y = op(xh, wh, None, **kw) + op(q8(xl), q8(wh), None, **kw) + op(q8(xh), q8(wl), None, **kw)
        elif mode == 'f8static':   # activations: fixed scales (|A| <= 112 unsaturated), weights: per-tensor (pack time)
            y = op(xh, wh, None, **kw) + op(q8s(xl, 12), q8(wh), None, **kw) + op(q8s(xh, 2), q8(wl), None, **kw)
        elif mode == 'hh_only':
            y = op(xh, wh, None, **kw)
        else:
            raise ValueError(mode)
        if b is not None:
            y = y + (b.view(1, -1, 1, 1) if y.dim() == 4 else b)
        return y
    ns.conv2d = lambda x, w, b=None, **kw: prod(TF.conv2d, x, w, b, **kw)
    ns.linear = lambda x, w, b=None: prod(lambda a, ww, bb: TF.linear(a, ww, bb), x, w, b)
    return ns


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.set_num_threads(16)
    for seed in (0, 1):
        sd = unet_ref.random_state_dict(seed=seed)
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(B, 3, 32, 32, generator=g)
        t = torch.rand(B, generator=g) * 999
        ctx = torch.randn(B, 11, 192, generator=g)
        unet_ref.F = TF
        ref = unet_ref.unet_forward(sd, x.double(), t, ctx.double()).float() if False else unet_ref.unet_forward(sd, x, t, ctx)
        for mode in ('split3', 'f8corr', 'f8static', 'hh_only'):
            unet_ref.F = shim(mode)
            out = unet_ref.unet_forward(sd, x, t, ctx)
            unet_ref.F = TF
            d = (out - ref).double()
            print(f'seed {seed} {mode:8s} rel_l2 {d.norm().item() / ref.double().norm().item():.3e} '
                  f'max|d|/max|ref| {d.abs().max().item() / ref.abs().max().item():.3e}', flush=True)


if __name__ == '__main__':
    main()
