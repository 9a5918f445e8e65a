#!/usr/bin/env python
"""Train-step throughput of the FULL SlotDiffusion image model (CLEVRTex 128x128 config) with the B200 modules dropped in:

  img --ResNet18-GN encoder + pos-embed + MLP (eager PyTorch, tools/full_model_parts.py)--> features
      --SlotAttentionWMask (libsdb200)--> slots
  img --frozen VQ-VAE encoder (eager PyTorch, no grad)--> x0 --q_sample--> x_t --UNetModel (libsdb200)--> eps_hat
  loss = mse(eps_hat, eps); backward through UNet, Slot Attention AND the encoder; fused Adam over all trainable parameters

i.e. LDM.loss_function / SADiffusion.forward of the reference (img_based/models/ddpm/ldm.py:58-83, sa_diffusion.py:155-200)
with only the two hot modules replaced -- the metric SURVEY 8d(i) names.  bench.py's `train` block times the hot modules
alone (synthetic features / latents); this tool adds the reference's own eager parts around them and reports both the
whole step and the share of each part.  Eager mode (the reference's trainer does not capture graphs).

    python tools/full_model_train_bench.py [--batch 64] [--steps 5]
    python tools/full_model_train_bench.py --frames 3 --slots 15 --iters 2 --batch 16      # SAVi video model (MOVi-D config)

--frames T > 0 runs the video model (video_based/models/savi_diffusion.py:169-216, :256-...): the encoder sees all B*T
frames, Slot Attention runs once per frame starting from init_latents (t = 0) or TransformerPredictor(previous slots)
(predictor.py:20-44: nn.TransformerEncoder, 2 layers, 4 heads, ffn 4*D, norm_first -- the reference's own eager module), and
the LDM loss is taken over the B*T flattened frames; samples/s then counts frames (BASELINE configs[2]).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402

from full_model_parts import ImageEncoder, VQVAEEncoder  # noqa: E402


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--frames', type=int, default=0, help='> 0: SAVi video model with this many frames per clip')
    ap.add_argument('--slots', type=int, default=11)
    ap.add_argument('--iters', type=int, default=3)
    args = ap.parse_args()
    from slotdiffusion_b200 import ops
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from slotdiffusion_b200.unet import UNetModel
    dev = torch.device('cuda')
    B, S, D, T = args.batch, args.slots, 192, args.frames
    F_ = B * max(T, 1)            # frames per step
    torch.manual_seed(0)
    enc = ImageEncoder((128, 128), D).to(dev).train()
    vae = VQVAEEncoder().to(dev).eval().requires_grad_(False)
    sa = SlotAttentionWMask(D, args.iters, S, D, 2 * D).to(dev).train()
    predictor = None
    if T > 0:
        layer = torch.nn.TransformerEncoderLayer(d_model=D, nhead=4, dim_feedforward=4 * D, norm_first=True, batch_first=True)
        predictor = torch.nn.TransformerEncoder(layer, num_layers=2).to(dev).train()
    unet = UNetModel(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(8, 4, 2),
                     dropout=0.1, channel_mult=(1, 2, 3, 4), num_head_channels=32, context_dim=D).to(dev).train()
    with torch.no_grad():
        for p in unet.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.02)
    init_slots = torch.nn.Parameter(torch.randn(1, S, D, device=dev))
    params = list(enc.parameters()) + list(sa.parameters()) + list(unet.parameters()) + [init_slots]
    if predictor is not None:
        params += list(predictor.parameters())
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    betas = torch.linspace(0.0015 ** 0.5, 0.0195 ** 0.5, 1000, dtype=torch.float64) ** 2
    acp = torch.cumprod(1 - betas, 0).float().to(dev)
    img_h = torch.randn(F_, 3, 128, 128).clamp_(-1, 1).pin_memory()         # [B*T, 3, H, W], clip-major
    out = {}

    def step():
        img = img_h.to(dev, non_blocking=True)
        ops.dropout_step_counter(dev).add_(1)
        with torch.no_grad():
            x0 = vae(img)
        t = torch.randint(0, 1000, (F_,), device=dev)
        eps = torch.randn_like(x0)
        a = acp[t].view(F_, 1, 1, 1)
        xt = a.sqrt() * x0 + (1 - a).sqrt() * eps
        feats = enc(img)
        if T > 0:                                                    # savi_diffusion.py:183-196
            feats = feats.unflatten(0, (B, T))
            prev, per_frame = None, []
            for f in range(T):
                latents = init_slots.expand(B, -1, -1) if prev is None else predictor(prev)
                prev, _ = sa(feats[:, f].contiguous(), latents)
                per_frame.append(prev)
            slots = torch.stack(per_frame, 1).flatten(0, 1)          # [B*T, S, D], matches the clip-major frames
        else:
            slots, _ = sa(feats, init_slots.expand(B, -1, -1))
        loss = torch.nn.functional.mse_loss(unet(xt, t, context=slots), eps)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        out['loss'] = loss

    ms = timed(step, args.steps)
    img = img_h.to(dev)

    def enc_only():
        enc(img).sum().backward()
        enc.zero_grad(set_to_none=True)

    def vae_only():
        with torch.no_grad():
            vae(img)
    ms_enc, ms_vae = timed(enc_only, args.steps), timed(vae_only, args.steps)
    print(json.dumps({
        'metric': 'train_step_samples_per_sec (full model, hot modules = libsdb200, rest = eager PyTorch)',
        'value': F_ / (ms / 1e3), 'unit': 'frames/s' if T > 0 else 'samples/s', 'ms_per_step': ms, 'per_gpu_batch': B,
        'frames_per_clip': T, 'num_slots': S, 'mode': 'eager',
        'loss': float(out['loss']),
        'parts_ms': {'resnet_encoder_fwd_bwd_eager_torch': ms_enc, 'vqvae_encoder_fwd_eager_torch': ms_vae,
                     'slot_attention_unet_adam_and_rest': ms - ms_enc - ms_vae},
        'cudnn_allow_tf32': torch.backends.cudnn.allow_tf32, 'matmul_allow_tf32': torch.backends.cuda.matmul.allow_tf32}))


if __name__ == '__main__':
    main()
