#!/usr/bin/env python
"""Where does the GEMM's MMA-issuing thread wait?  Needs the diagnostic build:

    SDB_GEMM_TIMING=1 python -m slotdiffusion_b200.build          # -> slotdiffusion_b200/libsdb200_gtiming.so (next to the product .so)
    SDB_LIB=$PWD/slotdiffusion_b200/libsdb200_gtiming.so python tools/gemm_wait_split.py [--batch 256]

Per shape: share of the issuer's lifetime spent waiting for operand stages (full[stage]: TMA / L2 fill behind), waiting
for a free accumulator (acc_empty: epilogue behind), and issuing (the rest ~ tensor pipe busy or issue-bound).
DESIGN.md section 8 item 3: decides what to do about the N = 128 convolutions (66 % tensor-pipe activity)."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from slotdiffusion_b200 import ops  # noqa: E402
from slotdiffusion_b200._lib import check, lib  # noqa: E402


def timing(reset):
    out = (ctypes.c_uint64 * 4)()
    check(lib().sdb_gemm_timing(ctypes.cast(out, ctypes.c_void_p), 1 if reset else 0), 'sdb_gemm_timing')
    return [int(v) for v in out]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--fmt', default='f16x3', choices=['f16x3', 'f8c'], help='operand format: three fp16 passes or fp8-corrected (2 pass-equivalents)')
    args = ap.parse_args()
    B = args.batch
    dev = torch.device('cuda')
    shapes = [
        ('conv 32x32 128->128 (+res)', 'conv', (B, 32, 32, 128), 128, True),
        ('conv 32x32 256->128 (+res)', 'conv', (B, 32, 32, 256), 128, True),
        ('conv 16x16 256->256 (+res)', 'conv', (B, 16, 16, 256), 256, True),
        ('conv 16x16 512->256 (+res)', 'conv', (B, 16, 16, 512), 256, True),
        ('lin  L=256 256->768 (qkv)', 'lin', (B * 256, 256), 768, False),
        ('lin  L=256 1024->256 (+res)', 'lin', (B * 256, 1024), 256, True),
        ('lin  L=256 256->256 (+res)', 'lin', (B * 256, 256), 256, True),
    ]
    print(f'{"shape":32s} {"us":>8s} {"wait operands":>14s} {"wait accum":>11s} {"issuing":>8s} {"tiles/CTA":>9s}')
    for name, kind, geo, N, with_res in shapes:
        if kind == 'conv':
            Bc, H, W, C = geo
            M, K, arows, aK, conv = Bc * H * W, 9 * C, Bc * H * W, C, (ops.SDB_A_CONV3, Bc, H, W, C)
        else:
            M, K = geo
            arows, aK, conv = M, K, None
        fmt = ops.SDB_FMT_F8C if args.fmt == 'f8c' else ops.SDB_FMT_F16X2      # timing only: plane contents are random bits
        a = ops.Packed(torch.randn(2 * arows * aK, device=dev).half(), arows, aK, fmt=fmt)
        w = ops.Packed((torch.randn(2 * N * K, device=dev) * K ** -0.5).half(), N, K, fmt=fmt, wexp=0)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if with_res else None
        out = torch.empty(M, N, device=dev)
        ops.gemm(a, w, bias=bias, residual=res, conv=conv, out=out)
        timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            ops.gemm(a, w, bias=bias, residual=res, conv=conv, out=out)
        e1.record()
        torch.cuda.synchronize()
        wf, wa, life, tiles = timing(True)
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        life = max(life, 1)
        print(f'{name:32s} {us:8.1f} {100 * wf / life:13.1f}% {100 * wa / life:10.1f}% {100 * (life - wf - wa) / life:7.1f}% '
              f'{tiles / args.reps / 148:9.2f}')


if __name__ == '__main__':
    main()
