#!/usr/bin/env python
"""Micro-benchmark of sdb_gemm on the UNet / Slot-Attention shapes at B=64 (CUDA events, back-to-back launches).

    python tools/gemm_bench.py [--reps 20] [--cg 0|1|2]

Prints one line per shape: time per launch, algorithmic TFLOP/s, issued tensor TFLOP/s (x3 passes).
Operands of different launches rotate through a pool larger than L2 so weights/activations come from HBM.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from slotdiffusion_b200 import ops  # noqa: E402

# (name, kind, M-geometry, N, K/C): kind 'lin' (M, N, K) or 'conv' (B, H, W, C, N)
SHAPES = [
    ('sa q-proj 704x192x192', 'lin', 704, 192, 192),
    ('sa mlp2 704x192x384', 'lin', 704, 192, 384),
    ('sa kv 65536x384x192', 'lin', 65536, 384, 192),
    ('conv 32x32 128->128', 'conv', (64, 32, 32, 128), 128),
    ('conv 32x32 384->128', 'conv', (64, 32, 32, 384), 128),
    ('conv 32x32 256->256 (up)', 'conv', (64, 32, 32, 256), 256),
    ('conv 16x16 256->256', 'conv', (64, 16, 16, 256), 256),
    ('conv 16x16 640->256', 'conv', (64, 16, 16, 640), 256),
    ('conv 8x8 384->384', 'conv', (64, 8, 8, 384), 384),
    ('conv 8x8 896->384', 'conv', (64, 8, 8, 896), 384),
    ('conv 4x4 512->512', 'conv', (64, 4, 4, 512), 512),
    ('conv 4x4 1024->512', 'conv', (64, 4, 4, 1024), 512),
    ('lin 16384x768x256 (qkv)', 'lin', 16384, 768, 256),
    ('lin 16384x2048x256 (ff0)', 'lin', 16384, 2048, 256),
    ('lin 16384x256x1024 (ff2)', 'lin', 16384, 256, 1024),
    ('lin 16384x256x256 (proj)', 'lin', 16384, 256, 256),
    ('lin 4096x3072x384 (ff0)', 'lin', 4096, 3072, 384),
    ('lin 4096x384x384', 'lin', 4096, 384, 384),
    ('lin 1024x4096x512 (ff0)', 'lin', 1024, 4096, 512),
    ('lin 1024x512x512', 'lin', 1024, 512, 512),
    ('lin 1024x512x2048 (ff2)', 'lin', 1024, 512, 2048),
    ('lin 64x512x512 (temb)', 'lin', 64, 512, 512),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--cg', type=int, default=0)
    ap.add_argument('--splitk', type=int, default=0)
    ap.add_argument('--passes', type=int, default=3)
    ap.add_argument('--only', default='', help='substring filter on shape names')
    args = ap.parse_args()
    if args.cg:
        os.environ['SDB_GEMM_CG'] = str(args.cg)
    if args.splitk:
        os.environ['SDB_GEMM_SPLITK'] = str(args.splitk)
    dev = torch.device('cuda')
    tot_ms, tot_fl = 0.0, 0.0
    print(f'{"shape":34s} {"us":>8s} {"alg TF/s":>9s} {"issued TF/s":>11s}')
    for spec in SHAPES:
        name, kind = spec[0], spec[1]
        if args.only and args.only not in name:
            continue
        if kind == 'lin':
            M, N, K = spec[2:]
            conv = None
            arows, aK = M, K
        else:
            (B, H, W, C), N = spec[2], spec[3]
            M, K = B * H * W, 9 * C
            conv = (ops.SDB_A_CONV3, B, H, W, C)
            arows, aK = M, C
        nbuf = max(2, min(8, int(200e6 // (arows * aK * 4 + N * K * 4) + 1)))
        As = [ops.Packed(torch.randn(2 * arows * aK, device=dev).half(), arows, aK) for _ in range(nbuf)]
        Ws = [ops.Packed((torch.randn(2 * N * K, device=dev) * K ** -0.5).half(), N, K) for _ in range(nbuf)]
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if not any(t in name for t in ('(qkv)', '(ff0)', ' kv ', 'q-proj', '(up)')) else None
        out = torch.empty(M, N, device=dev)
        def run():
            for i in range(args.reps):
                ops.gemm(As[i % nbuf], Ws[i % nbuf], bias=bias, residual=res, conv=conv, out=out, passes=args.passes)
        run()
        torch.cuda.synchronize()
        # replay from a CUDA graph so the host-side cost (ctypes + 4 cuTensorMapEncode per call) is not in the number
        sidestream = torch.cuda.Stream()
        sidestream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(sidestream):
            run()
        torch.cuda.current_stream().wait_stream(sidestream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        fl = 2.0 * M * N * K
        tot_ms += us / 1e3
        tot_fl += fl
        print(f'{name:34s} {us:8.1f} {fl / us / 1e6:9.1f} {args.passes * fl / us / 1e6:11.1f}')
    print(f'{"sum":34s} {tot_ms * 1e3:8.1f} {tot_fl / tot_ms / 1e9:9.1f}')


if __name__ == '__main__':
    main()
