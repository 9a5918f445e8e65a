#!/usr/bin/env python
"""Attention-core timing at the UNet shapes (B=64, or SDB_B): tensor-core kernel vs the CUDA-core kernel, CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slotdiffusion_b200 import ops

def graph_us(fn, reps=10):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

B, S, d = int(os.environ.get('SDB_B', 64)), 11, 32
tot = {True: 0.0, False: 0.0}
print(f'{"shape":34s} {"count":>5s} {"tc us":>8s} {"cuda-core us":>13s} {"default route us":>17s}')
for (L, heads, n) in [(256, 8, 5), (64, 12, 5), (16, 16, 6)]:
    C = heads * d
    qkv = torch.randn(B * L, 3 * C, device='cuda')
    kv = torch.randn(B * S, 2 * C, device='cuda')
    for name, (q, k, v, Lk) in {'self': (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], L),
                                'cross': (qkv[:, :C], kv[:, :C], kv[:, C:], S)}.items():
        t = {}
        for tc in (True, False, None):        # None: what the modules call (few keys -> csrc/attention_fewkeys.cu)
            t[tc] = graph_us(lambda: ops.attention_pack(q, k, v, B, L, Lk, heads, d, d ** -0.5, tc=tc))
            tot[tc] = tot.get(tc, 0.0) + n * t[tc]
        fl = 4.0 * L * Lk * d * heads * B
        print(f'{name:5s} L={L:4d} Lk={Lk:4d} heads={heads:3d}      {n:5d} {t[True]:8.1f} {t[False]:13.1f} {t[None]:17.1f}   tc: {fl / t[True] / 1e6:7.1f} TFLOP/s algorithmic')
print(f'sum over the 32 attention cores of one UNet evaluation: tc {tot[True]:.0f} us, cuda-core {tot[False]:.0f} us, default route {tot[None]:.0f} us')
