#!/usr/bin/env python
"""Slot-Attention timing on the GPU (CUDA-graph replay, CUDA events): whole module forward with the tensor-core
attend (default) and with the materialised-k/v path, and the attend kernel alone against the HBM roofline.

    python tools/sa_bench.py [--batch 64 256] [--slots 11] [--iters 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from slotdiffusion_b200 import autograd, ops  # noqa: E402
from slotdiffusion_b200.slot_attention import SlotAttentionWMask  # noqa: E402


def graph_time(fn, reps=20, flush=None):
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()          # evict L2 (256 MB write)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, nargs='+', default=[64, 256])
    ap.add_argument('--slots', type=int, default=11)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--tokens', type=int, default=1024)
    ap.add_argument('--dim', type=int, default=192)
    args = ap.parse_args()
    dev = torch.device('cuda')
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    hbm = peaks['hbm_gbs']
    S, D, N = args.slots, args.dim, args.tokens
    torch.manual_seed(0)
    mod = SlotAttentionWMask(D, args.iters, S, D, 2 * D).to(dev).eval()
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for B in args.batch:
        x = torch.randn(B, N, D, device=dev)
        s0 = torch.randn(B, S, D, device=dev)
        with torch.no_grad():
            res = {}
            t_res = None
            if ops.slot_attention_resident_supported(N, S, D, D, 2 * D):
                autograd.RESIDENT, autograd.RESIDENT_WAVES = True, 1 << 20
                t_res = graph_time(lambda: mod(x, s0), flush=flush)
            autograd.RESIDENT = False
            for fused in (True, False):
                autograd.FUSED_ATTEND = fused
                res[fused] = graph_time(lambda: mod(x, s0), flush=flush)
            autograd.FUSED_ATTEND = True
            # one-launch slot update instead of the GEMM tail (opt-in path; forced here whatever the row count)
            t_tail = None
            if ops.slot_update_supported(S, D, D, 2 * D):
                saved = autograd.FUSED_TAIL, autograd.FUSED_TAIL_MAX_ROWS
                autograd.FUSED_TAIL, autograd.FUSED_TAIL_MAX_ROWS = True, 1 << 30
                t_tail = graph_time(lambda: mod(x, s0), flush=flush)
                autograd.FUSED_TAIL, autograd.FUSED_TAIL_MAX_ROWS = saved
            # attend kernel alone (one iteration): cold (L2 flushed) and warm
            qa = torch.randn(B * S, D + 4, device=dev) * D ** -0.5
            t_cold = graph_time(lambda: ops.slot_attend_fused(x, qa, B, N, S, D, 1e-5, 1e-6, True), flush=flush)
            t_warm = graph_time(lambda: ops.slot_attend_fused(x, qa, B, N, S, D, 1e-5, 1e-6, True))
            kv = torch.randn(B * N, 2 * D, device=dev)
            q = torch.randn(B * S, D, device=dev)
            t_old = graph_time(lambda: ops.slot_attend(kv, q, B, N, S, D, D ** -0.5, 1e-6, True), flush=flush)
        mod_bytes = B * 4 * (N * D + 2 * S * D + S * N) + 1928448
        it_bytes = B * 4 * (N * D + S * N + 2 * S * D)
        print(json.dumps({
            'B': B, 'N': N, 'S': S, 'D': D, 'iters': args.iters,
            'module_us_resident': None if t_res is None else round(t_res, 1),
            'module_hbm_frac_resident': None if t_res is None else round(mod_bytes / t_res / 1e3 / hbm, 4),
            'module_us_fused': round(res[True], 1), 'module_us_kv_path': round(res[False], 1),
            'module_us_fused_one_launch_tail': None if t_tail is None else round(t_tail, 1),
            'module_hbm_frac_fused': round(mod_bytes / res[True] / 1e3 / hbm, 4),
            'attend_us_cold': round(t_cold, 1), 'attend_us_warm': round(t_warm, 1), 'attend_us_old_kernel_cold': round(t_old, 1),
            'attend_GBps_cold': round(it_bytes / t_cold / 1e3, 1), 'attend_hbm_frac_cold': round(it_bytes / t_cold / 1e3 / hbm, 4),
            'hbm_peak_GBps': hbm}))


if __name__ == '__main__':
    main()
