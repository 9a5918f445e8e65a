#!/usr/bin/env python
"""A/B of the UNet inference operand formats in ONE process (same box, same clocks): 20-NFE sampling step at batch B under
'fp32' (three fp16 passes) and 'fp8c' (fp16 + two e4m3 correction products), alternating, CUDA events."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from slotdiffusion_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--rounds', type=int, default=3)
    args = ap.parse_args()
    dev = torch.device('cuda')
    sa, unet, sampler, init_slots = bench.build_models(dev)
    B = args.batch
    noise = torch.randn(B, 3, 32, 32, device=dev)
    slots = torch.randn(B, bench.S, bench.D, device=dev)
    res = {'fp32': [], 'fp8c': []}
    for mode in ('fp32', 'fp8c'):
        ops.set_precision(mode)
        for _ in range(2):
            sampler.sample(noise, slots)
    torch.cuda.synchronize()
    for r in range(args.rounds):
        for mode in ('fp32', 'fp8c'):
            ops.set_precision(mode)
            sampler.sample(noise, slots)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                sampler.sample(noise, slots)
            e1.record()
            torch.cuda.synchronize()
            res[mode].append(e0.elapsed_time(e1) / 2)
    out = {'batch': B, 'ms_per_20nfe': res,
           'median_fp32': sorted(res['fp32'])[len(res['fp32']) // 2], 'median_fp8c': sorted(res['fp8c'])[len(res['fp8c']) // 2]}
    out['speedup'] = out['median_fp32'] / out['median_fp8c']
    print(json.dumps(out))


if __name__ == '__main__':
    main()
