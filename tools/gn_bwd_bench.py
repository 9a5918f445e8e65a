#!/usr/bin/env python
"""GroupNorm backward timing (CUDA-graph replay, L2 flushed) at the shapes of the full-model training step.
SDB_GN_BWD_FUSED=0 selects the two-launch form, default the fused two-pass kernel.

    python tools/gn_bwd_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from slotdiffusion_b200 import ops  # noqa: E402
from tools.sa_bench import graph_time  # noqa: E402

dev = torch.device('cuda')
flush = torch.empty(64 * 1024 * 1024, device=dev)
B = 64
res = {}
for HW, C, silu in ((16384, 64, 2), (4096, 128, 2), (4096, 128, 1), (1024, 256, 1), (256, 384, 1), (64, 512, 1), (1024, 384, 1)):
    G = 32
    x = torch.randn(B * HW, C, device=dev)
    da = torch.randn(B * HW, C, device=dev)
    stats = torch.stack([torch.zeros(B * G, device=dev), torch.ones(B * G, device=dev)], -1).contiguous()
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    us = graph_time(lambda: ops.groupnorm_bwd(x, None, da, stats, gamma, beta, dg, db, B, HW, G, silu), flush=flush)
    byts = 3 * B * HW * C * 4
    res['%dx%d' % (HW, C)] = {'us': round(us, 1), 'min_traffic_GBps': round(byts / us / 1e3, 1)}
print(json.dumps({'fused': os.environ.get('SDB_GN_BWD_FUSED', '1'), 'B': B, 'shapes': res}))
