#!/usr/bin/env python
"""Phase timeline of the persistent Slot-Attention kernel (csrc/slot_attention_resident.cu): SM-cycle stamps of thread 0 of
CTA 0 at every phase boundary of its first sample, printed as durations.

    python tools/sa_resident_timeline.py [--batch 4] [--slots 11] [--iters 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from slotdiffusion_b200 import _lib, autograd  # noqa: E402
from slotdiffusion_b200.slot_attention import SlotAttentionWMask  # noqa: E402

ITER_LABELS = ['softmax tile 0 done', 'weighted sums complete', 'drained (partials in smem)', 'barrier 1',
               'U reduce-scatter done', 'barrier 2', 'GRU done', 'barrier 3', 'LN + MLP1 done', 'barrier 4', 'MLP2 done',
               'barrier 5', 'LN_q + q projection done', 'barrier 6']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--slots', type=int, default=11)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--tokens', type=int, default=1024)
    ap.add_argument('--dim', type=int, default=192)
    a = ap.parse_args()
    dev = torch.device('cuda')
    torch.manual_seed(0)
    mod = SlotAttentionWMask(a.dim, a.iters, a.slots, a.dim, 2 * a.dim).to(dev).eval()
    x = torch.randn(a.batch, a.tokens, a.dim, device=dev)
    s0 = torch.randn(a.batch, a.slots, a.dim, device=dev)
    autograd.RESIDENT, autograd.RESIDENT_WAVES = True, 1 << 20
    with torch.no_grad():
        mod(x, s0)
        torch.cuda.synchronize()
        buf = torch.zeros(256, dtype=torch.int64, device=dev)
        _lib.lib().sdb_slot_attention_resident_debug(buf.data_ptr())
        mod(x, s0)
        torch.cuda.synchronize()
        _lib.lib().sdb_slot_attention_resident_debug(None)
    t = buf.cpu().tolist()[1:]
    labels = ['q projection of the initial slots done', 'barrier 6', 'conversion done (warp 0)']
    for it in range(a.iters):
        n = len(ITER_LABELS) - (2 if it == a.iters - 1 else 0)
        labels += ['it%d: %s' % (it, l) for l in ITER_LABELS[:n]]
    full = buf.cpu().tolist()
    print('GRU phase of iteration 0, per worker warp: start | input-side product done | - | gates + pushes done')
    for w in range(18):
        print('  warp %2d  ' % w + '  '.join('%7d' % full[128 + 4 * w + k] for k in range(4)))
    prev = 0
    for lab, v in zip(labels, t):
        print('%8d  +%7d  %s' % (v, v - prev, lab))
        prev = v


if __name__ == '__main__':
    main()
