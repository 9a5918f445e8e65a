"""CPU experiment: end-to-end error of reduced-pass split-fp16 GEMM modes on the UNet (oracle arithmetic).

The CUDA GEMM computes A*W as A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (3 tensor-core passes).  Dropping a pass is
the same as rounding one operand to fp16:
  w16   : W -> fp16(W)             (2 passes: A_hi*W_hi + A_lo*W_hi)
  a16   : A -> fp16(A)             (2 passes: A_hi*W_hi + A_hi*W_lo)
  both  : both rounded             (1 pass)
This script emulates that in the oracle (fp32 accumulate stays exact-ish on CPU) and prints the relative L2 /
max-normalised error of one UNet evaluation against the unmodified fp32 oracle, for the contract 1e-3.
Only conv2d / linear operands are rounded (attention cores and norms stay fp32, as in the CUDA path where
the attention kernels do their own 3-term products).
Run: python tools/experiments/pass_numerics.py [batch]
"""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as TF
from oracle import unet_ref


def r16(t):
    return t.half().float()


def shim(mode):
    ns = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith('__')})
    ra = r16 if mode in ('a16', 'both') else (lambda t: t)
    rw = r16 if mode in ('w16', 'both') else (lambda t: t)
    ns.conv2d = lambda x, w, b=None, **kw: TF.conv2d(ra(x), rw(w), b, **kw)
    ns.linear = lambda x, w, b=None: TF.linear(ra(x), rw(w), b)
    return ns


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.manual_seed(0)
    torch.set_num_threads(16)
    for seed in (0, 1):
        sd = unet_ref.random_state_dict(seed=seed)
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(B, 3, 32, 32, generator=g)
        t = torch.rand(B, generator=g) * 999
        ctx = torch.randn(B, 11, 192, generator=g)
        unet_ref.F = TF
        ref = unet_ref.unet_forward(sd, x, t, ctx)
        for mode in ('w16', 'a16', 'both'):
            unet_ref.F = shim(mode)
            out = unet_ref.unet_forward(sd, x, t, ctx)
            unet_ref.F = TF
            d = (out - ref).double()
            print(f'seed {seed} {mode:5s} rel_l2 {d.norm().item() / ref.double().norm().item():.3e} '
                  f'max|d|/max|ref| {d.abs().max().item() / ref.abs().max().item():.3e}', flush=True)


if __name__ == '__main__':
    main()
