"""Error of the persistent Slot-Attention kernel vs the reference goldens for every geometry of tests/helpers.SA_CASES."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import SA_CASES, golden, rel_l2, sa_case
from slotdiffusion_b200 import autograd, ops
from slotdiffusion_b200.slot_attention import SlotAttentionWMask
ops.set_precision('fp32')
for name, (B, N, Din, S, D, M, I) in SA_CASES.items():
    p, x, s0, gw, iters = sa_case(name)
    mod = SlotAttentionWMask(Din, I, S, D, M).cuda(); mod.load_state_dict(p)
    g = golden(name)
    out = {}
    for res in (True, False):
        autograd.RESIDENT = res
        with torch.no_grad():
            out[res] = mod(x.cuda(), s0.cuda())
    print(name, 'resident vs golden: slots %.2e mask %.2e | vs per-iteration path: slots %.2e mask %.2e | nan %s' % (
        rel_l2(out[True][0], g['slots']), rel_l2(out[True][1], g['mask']), rel_l2(out[True][0], out[False][0]),
        rel_l2(out[True][1], out[False][1]), bool(torch.isnan(out[True][0]).any())))
