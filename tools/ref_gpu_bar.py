#!/usr/bin/env python
"""The honest GPU bar (SURVEY 8d last line): the UNMODIFIED reference (baseline/_ref, eager PyTorch,
cuDNN / cuBLAS through ATen -- what `python scripts/train.py` gives a user today) timed ON THE SAME B200, next to the same
reference model with the B200 hot path dropped in (slotdiffusion_b200.dropin.install()).

    python tools/ref_gpu_bar.py [--out gpurun_out/ref_gpu_bar.json] [--quick]

TF32 modes of the stock arm: 'stock' = PyTorch defaults (cuDNN convolutions TF32, matmul fp32), 'off' = full fp32 (the
accuracy class of the B200 path), 'all' = TF32 everywhere.  CUDA events, 3 warm-ups, inputs resident on the device.
A measurement tool: not a product path, not imported by tests.
"""
import argparse
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402

import ref_import  # noqa: E402

CFG = ('img_based', 'sa_ldm/sa_ldm_clevrtex_params-res128.py')
TF32 = {'stock': (False, True), 'off': (False, False), 'all': (True, True)}     # (matmul, cudnn)


def set_tf32(mode):
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = TF32[mode]


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def build(dev, seed=0):
    if ref_import.box_copy_available():
        ref_import.use_box_copy()
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = ref_import.img_models().build_model(ref_import.fresh_params(*CFG))
    with torch.no_grad():
        for p in m.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.02)
    return m.to(dev)


def measure(model, dev, batches, train_batch, reps, sample_batches):
    out = {}
    g = torch.Generator().manual_seed(0)
    unet = model.dm_decoder.model.diffusion_model
    sa = model.slot_attention
    model.eval()
    with torch.no_grad():
        for B in batches:
            feats = torch.randn(B, 1024, 192, generator=g).to(dev)
            s0 = torch.randn(B, 11, 192, generator=g).to(dev)
            x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
            t = (torch.rand(B, generator=g) * 999).to(dev)
            ctx = torch.randn(B, 11, 192, generator=g).to(dev)
            out[f'slot_attention_fwd_ms_b{B}'] = round(timed(lambda: sa(feats, s0), reps), 4)
            out[f'unet_fwd_ms_b{B}'] = round(timed(lambda: unet(x, t, context=ctx), reps), 3)
        for B in sample_batches:
            ctx = torch.randn(B, 11, 192, generator=g).to(dev)
            ms = timed(lambda: model.dm_decoder.generate_imgs(cond=ctx, batch_size=B, use_dpm=True, verbose=False), 2, warm=2)
            out[f'sample_20nfe_ms_b{B}'] = round(ms, 2)
            out[f'denoise_sample_steps_per_s_b{B}'] = round(B * 20 / (ms / 1e3), 1)
    # full training step: the reference's forward -> calc_train_loss -> backward + Adam over all trainable parameters
    model.train()
    B = train_batch
    img = torch.randn(B, 3, 128, 128, generator=g).clamp(-1, 1).to(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)

    def step():
        data = {'img': img}
        loss = model.calc_train_loss(data, model(data))['denoise_loss']
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    ms = timed(step, reps, warm=3)
    out[f'train_step_ms_b{B}'] = round(ms, 2)
    out[f'train_samples_per_s_b{B}'] = round(B / (ms / 1e3), 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'ref_gpu_bar.json'))
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--train-batch', type=int, default=64)
    args = ap.parse_args()
    dev = torch.device('cuda')
    batches = (4, 64) if args.quick else (4, 64, 256)
    sample_batches = (4,) if args.quick else (4, 64, 256)
    reps = 3 if args.quick else 5
    res = {'gpu': torch.cuda.get_device_name(0), 'torch': torch.__version__, 'config': CFG[1],
           'what': 'unmodified reference (baseline/_ref) eager on this GPU; dropin = same model with libsdb200 modules'}
    ref = build(dev)
    for mode in (('stock',) if args.quick else ('stock', 'off', 'all')):
        set_tf32(mode)
        res[f'reference_tf32_{mode}'] = measure(ref, dev, batches, args.train_batch, reps, sample_batches)
        print(mode, json.dumps(res[f'reference_tf32_{mode}']), flush=True)
    sd = ref.state_dict()
    del ref
    torch.cuda.empty_cache()
    set_tf32('stock')
    from slotdiffusion_b200 import dropin
    dropin.install()
    new = build(dev)
    new.load_state_dict(sd, strict=True)
    res['dropin_eager'] = measure(new, dev, batches, args.train_batch, reps, sample_batches)
    print('dropin', json.dumps(res['dropin_eager']), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
