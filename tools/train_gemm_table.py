#!/usr/bin/env python
"""Per-shape table of every sdb_gemm launch of ONE full-model training step (forward + backward), each launch bracketed by
CUDA events in an eager run: where the GEMM time of the training step goes.

    python tools/train_gemm_table.py [--batch 64]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from slotdiffusion_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    a = ap.parse_args()
    dev = torch.device('cuda')
    torch.manual_seed(0)
    model = bench.FullImageModel(dev)
    model.train()
    img = torch.randn(a.batch, 3, 128, 128, device=dev).clamp_(-1, 1)
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        model.loss(img).backward()
    torch.cuda.synchronize()
    real = ops.gemm
    recs = []

    def rec_gemm(x, w, *pa, **kw):
        conv = kw.get('conv')
        M = x.rows if conv is None else conv[1] * conv[2] * conv[3]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real(x, w, *pa, **kw)
        e1.record()
        mode = kw.get('mode', None)
        kind = 'conv' if conv else 'lin'
        recs.append((e0, e1, (M, w.rows, w.K, kind, str(mode) if mode is not None else '',
                              'bf16' if getattr(x, 'bf16', False) else 'f16'), 2.0 * M * w.rows * w.K, phase[0]))
        return out
    phase = ['fwd']
    ops.gemm = rec_gemm
    try:
        model.zero_grad(set_to_none=True)
        loss = model.loss(img)
        phase[0] = 'bwd'
        loss.backward()
    finally:
        ops.gemm = real
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for e0, e1, key, fl, ph in recs:
        d = agg.setdefault((ph,) + key, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1) * 1e3
        d[2] += fl
    tot = sum(d[1] for d in agg.values())
    print('%d GEMM launches, %.1f ms in total (eager, per-launch events)' % (len(recs), tot / 1e3))
    for ph in ('fwd', 'bwd'):
        print('%s: %.1f ms' % (ph, sum(d[1] for k, d in agg.items() if k[0] == ph) / 1e3))
    print('%-58s %4s %9s %8s %9s' % ('phase, M, N, K, kind, mode, fmt', 'n', 'total us', 'avg us', 'alg TF/s'))
    for k, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print('%-58s %4d %9.1f %8.1f %9.1f' % (str(k), n, us, us / n, fl / us / 1e6))


if __name__ == '__main__':
    main()
