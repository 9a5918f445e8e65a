"""Generate tests/golden/*.npz from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tools/make_golden.py
Inputs and weights are produced by seeded generators that the tests re-run
(oracle.*.random_params / random_state_dict / torch.Generator), so the fixtures
hold only the reference OUTPUTS plus input checksums.  The reference classes
used: img_based SlotAttentionWMask (sa_diffusion.py:9-70), video_based
SlotAttentionWMask (savi_diffusion.py:10-71), UNetModel (unet.py:344-584),
NoiseScheduleVP/model_wrapper/DPM_Solver (dpm_solver.py), VectorQuantizer2
distance/argmin (quantize.py:84-94), DDPM schedule buffers (ddpm.py:69-131).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_import  # noqa: E402

ref_import.setup()
from oracle import slot_attention_ref as sa_ref  # noqa: E402
from oracle import unet_ref, dpm_ref  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def seeded(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


def checksum(t):
    t = t.double().flatten()
    return np.array([t.sum().item(), (t * t).sum().item(), t[0].item(), t[-1].item()])


SA_CASES = {
    # name: (B, N, Din, S, D, M, iters, which reference class)
    'sa_img_clevrtex': (2, 1024, 192, 11, 192, 384, 3, 'img'),      # BASELINE configs[0]/[1] shape
    'sa_vid_movid': (2, 1024, 192, 15, 192, 384, 2, 'vid'),         # shipped MOVi-D: 15 slots, 2 iters
    'sa_movie_24slots': (1, 1024, 192, 24, 192, 384, 2, 'vid'),     # BASELINE configs[3]: 24 slots
    'sa_coco_vitb16': (2, 196, 256, 7, 256, 512, 3, 'img'),         # BASELINE configs[4]: N=196, D=256
    'sa_ragged_small': (3, 77, 192, 5, 192, 384, 1, 'img'),         # ragged N, 1 iteration
}


def gen_sa():
    from slotdiffusion.img_based.models.sa_diffusion import SlotAttentionWMask as ImgSA
    from slotdiffusion.video_based.models.savi_diffusion import SlotAttentionWMask as VidSA
    for name, (B, N, Din, S, D, M, I, which) in SA_CASES.items():
        p = sa_ref.random_params(Din, D, M, seed=11)
        mod = (ImgSA if which == 'img' else VidSA)(Din, I, S, D, M).eval()
        mod.load_state_dict(p)
        x = seeded((B, N, Din), 21)
        s0 = seeded((B, S, D), 22)
        with torch.no_grad():
            slots, mask = mod(x, s0)
            slots64, mask64 = mod.double()(x.double(), s0.double())
        # gradient fixture: d/d(inputs, slots, params) of a fixed linear functional of slots
        mod = mod.float()
        xg = x.clone().requires_grad_(True)
        sg = s0.clone().requires_grad_(True)
        so, _ = mod(xg, sg)
        gw = seeded(tuple(so.shape), 23)
        (so * gw).sum().backward()
        grads = {('grad.' + k): (v.grad.numpy() if v.grad.dim() == 1 else checksum(v.grad))
                 for k, v in mod.named_parameters()}
        np.savez_compressed(
            os.path.join(OUT, name + '.npz'),
            cfg=np.array([B, N, Din, S, D, M, I]), x_sum=checksum(x), s0_sum=checksum(s0),
            slots=slots.numpy(), mask=mask.numpy().astype(np.float32),
            slots64=slots64.numpy(), argmax64=mask64.argmax(1).numpy().astype(np.uint8),
            margin64=(mask64.topk(2, dim=1).values[:, 0] - mask64.topk(2, dim=1).values[:, 1]).numpy().astype(np.float32),
            grad_inputs_sum=checksum(xg.grad), grad_slots=sg.grad.numpy(),
            grad_inputs_head=xg.grad[:, :8].numpy(), **grads)
        print(name, 'ok', slots.abs().max().item())


def gen_unet():
    from slotdiffusion.video_based.models.unet.unet import UNetModel
    cases = {
        'unet_clevrtex': (dict(), 2, 11, 32),                                   # shipped config
        'unet_small': (dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,),
                            num_res_blocks=1, context_dim=64), 3, 5, 16),       # fast CPU case
    }
    for name, (over, B, S, hw) in cases.items():
        cfg = dict(unet_ref.DEFAULT_CFG, **over)
        sd = unet_ref.random_state_dict(cfg, seed=31)
        net = UNetModel(dropout=0.1, dims=2, use_checkpoint=False, resblock_updown=False,
                        conv_resample=True, transformer_depth=1, n_embed=None, **cfg).eval()
        net.load_state_dict(sd)
        x = seeded((B, cfg['in_channels'], hw, hw), 41)
        ctx = seeded((B, S, cfg['context_dim']), 42)
        t_int = torch.tensor([7, 503, 999][:B])
        t_flt = torch.tensor([0.0, 333.25, 998.999][:B])
        with torch.no_grad():
            y_int = net(x, t_int, context=ctx)
            y_flt = net(x, t_flt, context=ctx)
        # gradient fixture (training use): d loss / d context and a few weights
        ctxg = ctx.clone().requires_grad_(True)
        gw = seeded(tuple(y_int.shape), 43)
        (net(x, t_int, context=ctxg) * gw).sum().backward()
        sdn = dict(net.named_parameters())
        keys = ['time_embed.0.weight', 'input_blocks.0.0.weight', 'out.2.weight',
                'middle_block.1.transformer_blocks.0.attn2.to_k.weight',
                'middle_block.0.in_layers.2.weight', 'output_blocks.0.0.skip_connection.weight']
        g = {('gsum.' + k): checksum(sdn[k].grad) for k in keys}
        np.savez_compressed(os.path.join(OUT, name + '.npz'), x_sum=checksum(x), ctx_sum=checksum(ctx),
                            y_int=y_int.numpy(), y_flt=y_flt.numpy(), grad_ctx=ctxg.grad.numpy(), **g)
        print(name, 'ok', y_int.abs().max().item())


def gen_dpm():
    from slotdiffusion.video_based.models.ddpm.dpm_solver import NoiseScheduleVP, model_wrapper, DPM_Solver
    from slotdiffusion.video_based.models.unet.unet import UNetModel
    betas = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
    ns = NoiseScheduleVP(betas=betas)
    t = torch.linspace(1e-3, 1, 101)
    lam = ns.marginal_lambda(t)
    out = dict(t=t.numpy(), log_alpha=ns.marginal_log_mean_coeff(t).numpy(), lam=lam.numpy(),
               inv=ns.inverse_lambda(lam).numpy(), betas_sum=checksum(betas))
    # full sampler with the real UNet + a seeded codebook standing in for the frozen VQ-VAE
    cfg = dict(unet_ref.DEFAULT_CFG)
    sd = unet_ref.random_state_dict(cfg, seed=31)
    net = UNetModel(dropout=0.1, dims=2, use_checkpoint=False, resblock_updown=False,
                    conv_resample=True, transformer_depth=1, n_embed=None, **cfg).eval()
    net.load_state_dict(sd)
    cb = seeded((4096, 3), 51)

    class VAE:
        def quantize(self, h):
            # VectorQuantizer.forward distance + argmin (quantize.py:84-94), straight-through value
            from slotdiffusion.video_based.models.vqvae.quantize import VectorQuantizer2
            vq = VectorQuantizer2(4096, 3, beta=0.25)
            vq.embedding.weight.data.copy_(cb)
            return vq(h)[0]
    calls = []

    class Wrap(torch.nn.Module):
        def forward(self, x, t, context=None):
            calls.append(t[0].item())
            return net(x, t, context=context)
    model = Wrap()
    model.vae = VAE()
    B = 1
    ctx = seeded((B, 11, 192), 52)
    xT = seeded((B, 3, 32, 32), 53)
    for vq in (True, False):
        calls.clear()
        fn = model_wrapper(model=model, noise_schedule=ns, model_type='noise',
                           guidance_type='classifier-free', condition=ctx)
        solver = DPM_Solver(fn, ns, algorithm_type='dpmsolver++', correcting_x0_fn=False, vq_denoised=vq)
        with torch.no_grad():
            y = solver.sample(xT, steps=20, order=3, method='singlestep')
        out['sample_vq' if vq else 'sample_novq'] = y.numpy()
        out['t_model'] = np.array(calls)
    # q_sample
    x0 = seeded((4, 3, 32, 32), 54)
    noise = seeded((4, 3, 32, 32), 55)
    tt = torch.tensor([0, 10, 500, 999])
    from slotdiffusion.video_based.models.ddpm.ddpm import DDPM
    from slotdiffusion.video_based.models.ddpm.utils import extract_to
    bufs = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())
    xt = extract_to(bufs['sqrt_alphas_bar'], tt, x0.shape) * x0 + \
        extract_to(bufs['sqrt_one_minus_alphas_bar'], tt, x0.shape) * noise
    out['q_sample'] = xt.numpy()
    np.savez_compressed(os.path.join(OUT, 'dpm.npz'), **out)
    print('dpm ok', len(calls))


PLAN_CASES = [(20, 3), (10, 3), (15, 3), (21, 3), (22, 3), (50, 3), (7, 3), (12, 2), (9, 2), (6, 1)]   # (steps, order)


def toy_eps(x, t_model, cond):
    """cheap analytic stand-in for the UNet (tests/test_dpm_plan_cpu.py uses the same function)"""
    return 0.3 * torch.sin(1.7 * x + cond.mean(dim=(1, 2)).view(-1, 1, 1, 1)) + 1e-4 * t_model.view(-1, 1, 1, 1) * x.roll(1, -1)


def gen_dpm_plan():
    """Reference DPM_Solver.sample (singlestep, dpmsolver++) for several (steps, order) with a toy noise model:
    pins the host-precomputed plan of slotdiffusion_b200.dpm_solver.build_plan for more than the 20-step case."""
    from slotdiffusion.video_based.models.ddpm.dpm_solver import NoiseScheduleVP, model_wrapper, DPM_Solver
    betas = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
    ns = NoiseScheduleVP(betas=betas)
    cb = seeded((64, 3), 58)

    class Toy(torch.nn.Module):
        def forward(self, x, t, context=None, quantize=False):
            return toy_eps(x, t, context)
    model = Toy()

    class VAE:
        def quantize(self, h):
            from slotdiffusion.video_based.models.vqvae.quantize import VectorQuantizer2     # the reference's quantiser
            vq = VectorQuantizer2(64, 3, beta=0.25)
            vq.embedding.weight.data.copy_(cb)
            return vq(h)[0]
    model.vae = VAE()
    ctx = seeded((2, 11, 192), 56)
    xT = seeded((2, 3, 16, 16), 57)
    out = {}
    for steps, order in PLAN_CASES:
        for vq in (False, True):
            fn = model_wrapper(model=model, noise_schedule=ns, model_type='noise', guidance_type='classifier-free',
                               condition=ctx)
            solver = DPM_Solver(fn, ns, algorithm_type='dpmsolver++', correcting_x0_fn=False, vq_denoised=vq)
            with torch.no_grad():
                y = solver.sample(xT, steps=steps, order=order, method='singlestep')
            out[f's{steps}_o{order}_{"vq" if vq else "novq"}'] = y.numpy()
    np.savez_compressed(os.path.join(OUT, 'dpm_plan.npz'), **out)
    print('dpm_plan ok', len(out))


def gen_vqvae():
    """Encoder / Decoder of the frozen VQ-VAE (modules.py:168-362) at the shipped CLEVRTex config (128x128, ch 64,
    ch_mult [1,2,4], mid attention at 32x32) and a small ragged config; weights from oracle.vqvae_ref.random_state_dicts."""
    import contextlib
    import io
    from slotdiffusion.video_based.models.vqvae.modules import Encoder, Decoder
    from oracle import vqvae_ref
    out = {}
    for tag, over, B in (('full', {}, 2), ('small', dict(resolution=32, ch_mult=(1, 2)), 3)):
        cfg = dict(vqvae_ref.DEFAULT_CFG, **over)
        esd, dsd = vqvae_ref.random_state_dicts(cfg, seed=61)
        with contextlib.redirect_stdout(io.StringIO()):
            enc, dec = Encoder(dropout=0.0, **cfg).eval(), Decoder(dropout=0.0, **cfg).eval()
        enc.load_state_dict(esd, strict=True)
        dec.load_state_dict(dsd, strict=True)
        R = cfg['resolution']
        r = R // 2 ** (len(cfg['ch_mult']) - 1)
        x = seeded((B, 3, R, R), 62).clamp(-1, 1)
        z = seeded((B, 3, r, r), 63)
        with torch.no_grad():
            out[tag + '_enc'] = enc(x).numpy()
            out[tag + '_dec'] = dec(z).numpy()
        out[tag + '_x_sum'], out[tag + '_z_sum'] = checksum(x), checksum(z)
    np.savez_compressed(os.path.join(OUT, 'vqvae.npz'), **out)
    print('vqvae.npz', {k: v.shape for k, v in out.items()})


def gen_resnet():
    """ResNet18-GN encoder (resnet.py:150-315, small_inputs=True, use_layer4=False): forward at 128x128 (B=1) and
    forward + parameter gradients of a fixed linear functional at 64x64 (B=2); weights: oracle.resnet_ref.random_state_dict."""
    from slotdiffusion.video_based.models.resnet import resnet18
    from oracle import resnet_ref
    sd = resnet_ref.random_state_dict('resnet18', False, seed=71)
    net = resnet18(small_inputs=True, use_layer4=False)
    net.load_state_dict(sd, strict=True)
    out = {}
    x128 = seeded((1, 3, 128, 128), 72).clamp(-1, 1)
    with torch.no_grad():
        out['y128'] = net(x128).numpy()
    x64 = seeded((2, 3, 64, 64), 73).clamp(-1, 1)
    y = net(x64)
    gw = seeded(tuple(y.shape), 74)
    (y * gw).sum().backward()
    out['y64'] = y.detach().numpy()
    for k, v in net.named_parameters():
        out['grad.' + k] = v.grad.numpy() if v.grad.dim() == 1 else checksum(v.grad)
    np.savez_compressed(os.path.join(OUT, 'resnet.npz'), **out)
    print('resnet.npz', out['y128'].shape, out['y64'].shape, len(out))


PRED_CASES = {
    # name: (B, S, D, layers, heads, ffn, norm_first)
    'movid': (3, 15, 192, 2, 4, 768, True),        # shipped MOVi-D / MOVi-E SAVi-diffusion predictor (head dim 48)
    'clevrer': (2, 7, 128, 2, 4, 512, True),       # savi defaults (head dim 32)
    'postln': (2, 11, 256, 1, 4, 512, False),      # post-LN variant, head dim 64
}


def gen_predictor():
    """TransformerPredictor (video_based/models/predictor.py:20-44), eval mode (dropout off): outputs, input gradient and
    parameter gradients of a fixed linear functional; weights: oracle.predictor_ref.random_state_dict."""
    from slotdiffusion.video_based.models.predictor import TransformerPredictor
    from oracle import predictor_ref
    out = {}
    for name, (B, S, D, L, Hh, F, nf) in PRED_CASES.items():
        sd = predictor_ref.random_state_dict(D, L, F, seed=900 + len(name))
        net = TransformerPredictor(d_model=D, num_layers=L, num_heads=Hh, ffn_dim=F, norm_first=nf).eval()
        net.load_state_dict(sd, strict=True)
        x = seeded((B, S, D), 91).requires_grad_(True)
        y = net(x)
        gw = seeded((B, S, D), 92)
        (y * gw).sum().backward()
        out[name + '.y'] = y.detach().numpy()
        out[name + '.dx'] = x.grad.numpy()
        for k, v in net.named_parameters():
            out[name + '.grad.' + k] = v.grad.numpy() if v.grad.dim() == 1 else checksum(v.grad)
    np.savez_compressed(os.path.join(OUT, 'predictor.npz'), **out)
    print('predictor.npz', len(out))


def gen_layout():
    """state_dict layout (ordered key -> shape) of the hot-path sub-modules inside the full reference models
    (build_model of the shipped configs): the checkpoint contract of the drop-in modules."""
    import json
    import runpy
    out = {}
    for task, rel, get in (('img_based', 'sa_ldm/sa_ldm_clevrtex_params-res128.py', ref_import.img_models),
                           ('video_based', 'savi_ldm/savi_ldm_movid_params-res128.py', ref_import.video_models),
                           ('video_based', 'savi_ldm/savi_ldm_movie_params-res128.py', ref_import.video_models)):
        cfg = os.path.join(ref_import.REF_ROOT, 'slotdiffusion', task, 'configs', rel)
        if not os.path.exists(cfg):
            continue
        params = runpy.run_path(cfg)['SlotAttentionParams']()      # fresh class: build_model pops from its dicts
        model = get().build_model(params)
        name = os.path.basename(rel)[:-3]
        out[name] = {
            'slot_attention': [[k, list(v.shape)] for k, v in model.slot_attention.state_dict().items()],
            'unet': [[k, list(v.shape)] for k, v in model.dm_decoder.model.diffusion_model.state_dict().items()],
            'slot_attention_ctor': dict(in_features=model.slot_attention.in_features,
                                        num_iterations=model.slot_attention.num_iterations,
                                        num_slots=model.slot_attention.num_slots,
                                        slot_size=model.slot_attention.slot_size,
                                        mlp_hidden_size=model.slot_attention.mlp_hidden_size),
            'unet_dict': {k: (list(v) if isinstance(v, tuple) else v) for k, v in params.unet_dict.items()},
            'n_model_state': len(model.state_dict()),
        }
        print(name, 'layout ok', len(out[name]['unet']), len(out[name]['slot_attention']))
    # COCO (DINO encoder: ViT weights are not constructible offline) -- the two hot-path modules built directly
    # from the config dicts, the way sa_diffusion.py:132-139 and ddpm.py:342 do
    from slotdiffusion.img_based.models.sa_diffusion import SlotAttentionWMask
    from slotdiffusion.video_based.models.unet.unet import UNetModel
    rel = 'sa_ldm/sa_ldm_dino_coco_params-res224.py'
    params = runpy.run_path(os.path.join(ref_import.REF_ROOT, 'slotdiffusion', 'img_based', 'configs', rel))[
        'SlotAttentionParams']()
    sd_, ed_ = params.slot_dict, params.enc_dict
    ctor = dict(in_features=ed_['enc_out_channels'], num_iterations=sd_['num_iterations'], num_slots=sd_['num_slots'],
                slot_size=sd_['slot_size'], mlp_hidden_size=sd_['slot_mlp_size'])
    sa = SlotAttentionWMask(eps=1e-6, **ctor)
    un = UNetModel(**params.unet_dict)
    out[os.path.basename(rel)[:-3]] = {
        'slot_attention': [[k, list(v.shape)] for k, v in sa.state_dict().items()],
        'unet': [[k, list(v.shape)] for k, v in un.state_dict().items()],
        'slot_attention_ctor': ctor,
        'unet_dict': {k: (list(v) if isinstance(v, tuple) else v) for k, v in params.unet_dict.items()},
        'n_model_state': None,
    }
    print('coco layout ok')
    with open(os.path.join(OUT, 'state_dict_layout.json'), 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['sa', 'unet', 'dpm', 'layout', 'dpm_plan', 'vqvae', 'resnet', 'predictor']
    if 'predictor' in which:
        gen_predictor()
    if 'vqvae' in which:
        gen_vqvae()
    if 'resnet' in which:
        gen_resnet()
    if 'sa' in which:
        gen_sa()
    if 'unet' in which:
        gen_unet()
    if 'dpm' in which:
        gen_dpm()
    if 'layout' in which:
        gen_layout()
    if 'dpm_plan' in which:
        gen_dpm_plan()
