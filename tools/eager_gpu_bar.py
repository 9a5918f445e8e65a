#!/usr/bin/env python
"""The honest GPU bar: the reference's arithmetic as eager PyTorch ON THE SAME B200 -- cuBLAS/cuDNN through
ATen, the way the reference runs today -- timed next to the B200-native path on the bench workload.

    python tools/eager_gpu_bar.py [--batch 64] [--nfe 20]

The reference package itself cannot travel to the GPU box (un-vendored `nerv`), so the eager arm is the oracle's
functional restatement of the same modules (oracle/unet_ref.py, oracle/slot_attention_ref.py: plain torch ops, pinned
to the reference by tests/golden) moved to CUDA.  This is a measurement tool, not a product path and not part of
bench.py; it reports ms per UNet evaluation and per Slot-Attention forward for
  eager fp32 (TF32 off: the accuracy class of the B200 path) | eager TF32 (PyTorch's default for cuDNN convolutions) | libsdb200.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import slot_attention_ref as sa_ref  # noqa: E402
from oracle import unet_ref  # noqa: E402


def timed(fn, reps):
    if not torch.cuda.is_available():        # --cpu-selftest: exercise the code path only
        import time
        t0 = time.perf_counter()
        fn()
        return (time.perf_counter() - t0) * 1e3
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--cpu-selftest', action='store_true', help='run the eager arm once on the CPU (no libsdb200 arm): script check')
    args = ap.parse_args()
    dev = torch.device('cpu' if args.cpu_selftest else 'cuda')
    B, S, D, N = args.batch, 11, 192, 1024
    sd = {k: v.to(dev) for k, v in unet_ref.random_state_dict(seed=0).items()}
    p = {k: v.to(dev) for k, v in sa_ref.random_params(D, D, 2 * D, seed=0).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    t = (torch.rand(B, generator=g) * 999).to(dev)
    ctx = torch.randn(B, S, D, generator=g).to(dev)
    feats = torch.randn(B, N, D, generator=g).to(dev)
    s0 = torch.randn(B, S, D, generator=g).to(dev)
    out = {'batch': B}
    with torch.no_grad():
        for name, tf32 in (('eager_fp32', False), ('eager_tf32', True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            out[name + '_unet_ms'] = round(timed(lambda: unet_ref.unet_forward(sd, x, t, ctx), args.reps), 3)
            out[name + '_slot_attention_ms'] = round(timed(lambda: sa_ref.slot_attention_forward(p, feats, s0, 3), args.reps), 3)
        ref = unet_ref.unet_forward(sd, x, t, ctx)          # TF32 result
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ref32 = unet_ref.unet_forward(sd, x, t, ctx)
        out['eager_tf32_vs_fp32_rel_l2'] = float((ref - ref32).norm() / ref32.norm())
        if args.cpu_selftest:
            print(json.dumps(out))
            return
        from slotdiffusion_b200.slot_attention import SlotAttentionWMask
        from slotdiffusion_b200.unet import UNetModel
        net = UNetModel(dropout=0.1, **unet_ref.DEFAULT_CFG).to(dev).eval()
        net.load_state_dict(sd)
        sa = SlotAttentionWMask(D, 3, S, D, 2 * D).to(dev).eval()
        sa.load_state_dict(p)
        out['sdb200_unet_ms'] = round(timed(lambda: net(x, t, context=ctx), args.reps), 3)
        out['sdb200_slot_attention_ms'] = round(timed(lambda: sa(feats, s0), args.reps), 3)
        out['sdb200_vs_fp32_rel_l2'] = float((net(x, t, context=ctx) - ref32).norm() / ref32.norm())
    print(json.dumps(out))


if __name__ == '__main__':
    main()
