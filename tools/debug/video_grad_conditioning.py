"""How well conditioned are the video model's recurrence gradients?  The stock reference model (SAViDiffusion, MOVi-D config)
is run twice in fp32: as is, and with the predictor's OUTPUT perturbed by a relative 1e-6 (the size of fp32 round-off
differences between two correct implementations).  The relative change of the gradient groups is the amplification the
recurrence applies to round-off -- the floor for any stock-vs-drop-in comparison of those groups
(tests/test_reference_dropin_gpu.py::test_reference_savidiffusion_video_train_step_stock_vs_dropin)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import test_reference_dropin_gpu as T  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda')
B, Tn = 2, 3
m = T._build(*T.VCFG)
T._nonzero_init(m)
m = m.to(dev)
img = torch.randn(B, Tn, 3, 128, 128, generator=torch.Generator().manual_seed(9)).clamp(-1, 1).to(dev)
_, _, g0 = T._train_pass(m, img, 321)
_, _, g0b = T._train_pass(m, img, 321)


def groups(ga, gb):
    out = {}
    for name in ('slot_attention.', 'predictor.', 'encoder', 'dm_decoder', ''):
        d = sum((ga[k].double() - gb[k].double()).norm().item() ** 2 for k in gb if k.startswith(name))
        r = sum(gb[k].double().norm().item() ** 2 for k in gb if k.startswith(name))
        out[name or 'all'] = (d / max(r, 1e-300)) ** 0.5
    return out


print('repeatability (same run twice):', groups(g0b, g0))
for rel in (1e-7, 1e-6, 1e-5):
    gen = torch.Generator(device='cuda').manual_seed(1)

    def hook(mod, inp, out, rel=rel):
        return out + rel * out.abs().mean() * torch.randn(out.shape, device=out.device, generator=gen)
    h = m.predictor.register_forward_hook(hook)
    _, _, g1 = T._train_pass(m, img, 321)
    h.remove()
    print('predictor output perturbed by rel %.0e:' % rel, groups(g1, g0))
