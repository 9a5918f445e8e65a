"""Per-parameter comparison of graphed vs eager Slot-Attention training gradients (debug aid for tests/test_graphed_gpu.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
from helpers import rel_l2, seeded  # noqa: E402
from oracle import slot_attention_ref as sa_ref  # noqa: E402
from slotdiffusion_b200 import graphed, ops  # noqa: E402
from slotdiffusion_b200.slot_attention import SlotAttentionWMask  # noqa: E402

ops.set_precision('fp32')
p = sa_ref.random_params(192, 192, 384, seed=11)


def make():
    m = SlotAttentionWMask(192, 2, 5, 192, 384).cuda().train()
    m.load_state_dict(p)
    return m


def run(m, step):
    x = seeded((2, 96, 192), 10 + step).cuda().requires_grad_(True)
    s0 = seeded((2, 5, 192), 20 + step).cuda().requires_grad_(True)
    slots, mask = m(x, s0)
    return (slots * seeded((2, 5, 192), 30 + step).cuda()).sum(), x, s0


a, b = make(), make()
graphed.enable(b)
for step in range(3):
    la, xa, sa = run(a, step)
    lb, xb, sb = run(b, step)
    print('step', step, 'loss rel', rel_l2(lb, la))
    a.zero_grad(set_to_none=True)
    b.zero_grad(set_to_none=True)
    la.backward()
    lb.backward()
    print('  dx', rel_l2(xb.grad, xa.grad), 'ds0', rel_l2(sb.grad, sa.grad))
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        print('  %-28s %.3e   |ga| %.3e |gb| %.3e' % (k, rel_l2(pb.grad, pa.grad), float(pa.grad.norm()), float(pb.grad.norm())))
