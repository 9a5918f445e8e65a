"""GPU parity of the drop-in modules (through the C ABI) against the CPU oracle and the committed golden
fixtures generated from the reference modules.  north_star tolerance: <= 1e-3 relative fp32, bit-exact
argmax slot masks (near-ties of the fp64 reference excluded and counted)."""
import numpy as np
import pytest
import torch

from slotdiffusion_b200 import ops as ops_mod

from helpers import SA_CASES, argmax_mismatch, golden, rel_l2, sa_case, seeded
from oracle import dpm_ref, unet_ref
from oracle import slot_attention_ref as sa_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 1e-3          # north_star: 1e-3 relative fp32
TIGHT = 5e-5        # what the 3-pass split-fp16 path actually achieves


def make_sa(name):
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, Din, S, D, M, I = SA_CASES[name]
    p, x, s0, gw, iters = sa_case(name)
    mod = SlotAttentionWMask(Din, I, S, D, M).cuda()
    mod.load_state_dict(p)       # reference key names / shapes
    return mod, p, x, s0, gw, iters


@pytest.mark.parametrize('name', list(SA_CASES))
def test_slot_attention_matches_reference_golden(name):
    g = golden(name)
    mod, p, x, s0, gw, iters = make_sa(name)
    with torch.no_grad():
        slots, mask = mod(x.cuda(), s0.cuda())
    assert rel_l2(slots, g['slots']) < TIGHT
    assert rel_l2(mask, g['mask']) < TIGHT
    assert rel_l2(slots, g['slots64']) < TIGHT
    real, near = argmax_mismatch(mask, g['argmax64'], g['margin64'], 1e-5)
    assert real == 0, (real, near)


def test_slot_attention_full_size_properties():
    """BASELINE full size (B=64, N=1024, S=11): size-independent properties + oracle on a sub-batch."""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, S, D = 64, 1024, 11, 192
    p = sa_ref.random_params(D, D, 2 * D, seed=5)
    mod = SlotAttentionWMask(D, 3, S, D, 2 * D).cuda()
    mod.load_state_dict(p)
    x, s0 = seeded((B, N, D), 61), seeded((B, S, D), 62)
    with torch.no_grad():
        slots, mask = mod(x.cuda(), s0.cuda())
        assert torch.isfinite(slots).all()
        assert (mask.sum(1) - 1).abs().max().item() < 1e-5           # softmax over slots sums to 1 per token
        assert mask.min().item() >= 0
        # batch independence (the token-chunking of the attend kernel depends on B, so not bit-identical)
        s2, m2 = mod(x[5:9].cuda(), s0[5:9].cuda())
        assert rel_l2(s2, slots[5:9]) < 1e-5 and rel_l2(m2, mask[5:9]) < 1e-5
        assert (m2.argmax(1) != mask[5:9].argmax(1)).sum().item() <= 2
        # slot-permutation equivariance
        perm = torch.randperm(S)
        s3, m3 = mod(x[:4].cuda(), s0[:4, perm].cuda())
        assert rel_l2(s3, slots[:4, perm]) < 1e-5
    ref_s, ref_m = sa_ref.slot_attention_forward(p, x[:3].double(), s0[:3].double(), 3)
    assert rel_l2(slots[:3], ref_s) < TIGHT
    margin = ref_m.topk(2, dim=1).values
    real, near = argmax_mismatch(mask[:3], ref_m.argmax(1).numpy(), (margin[:, 0] - margin[:, 1]).numpy(), 1e-5)
    assert real == 0


def make_unet(cfg_over=None, seed=31):
    from slotdiffusion_b200.unet import UNetModel
    cfg = dict(unet_ref.DEFAULT_CFG, **(cfg_over or {}))
    sd = unet_ref.random_state_dict(cfg, seed=seed)
    net = UNetModel(dropout=0.1, dims=2, use_checkpoint=False, resblock_updown=False, conv_resample=True,
                    transformer_depth=1, n_embed=None, **cfg).cuda().eval()
    net.load_state_dict(sd)      # reference key names / shapes, strict
    return net, sd, cfg


def test_unet_state_dict_layout_matches_reference():
    net, sd, cfg = make_unet()
    assert set(net.state_dict().keys()) == set(sd.keys())
    assert sum(p.numel() for p in net.parameters()) == sum(v.numel() for v in sd.values())
    assert abs(sum(p.numel() for p in net.parameters()) / 1e6 - 134.24) < 0.01


def test_unet_fresh_init_is_zero_like_reference():
    """zero_module on out-convs / proj_out / final conv => a freshly built UNet outputs exactly 0."""
    from slotdiffusion_b200.unet import UNetModel
    torch.manual_seed(0)
    net = UNetModel(**dict(unet_ref.DEFAULT_CFG, model_channels=64, channel_mult=(1, 2), num_res_blocks=1,
                           attention_resolutions=(2,), context_dim=64)).cuda().eval()
    with torch.no_grad():
        y = net(torch.randn(2, 3, 16, 16).cuda(), torch.tensor([1, 2]).cuda(), context=torch.randn(2, 5, 64).cuda())
    assert y.abs().max().item() == 0.0


def test_unet_small_matches_reference_golden():
    g = golden('unet_small')
    net, sd, cfg = make_unet(dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,),
                                  num_res_blocks=1, context_dim=64))
    x, ctx = seeded((3, 3, 16, 16), 41).cuda(), seeded((3, 5, 64), 42).cuda()
    with torch.no_grad():
        y = net(x, torch.tensor([7, 503, 999]).cuda(), context=ctx)
        assert rel_l2(y, g['y_int']) < TIGHT
        y = net(x, torch.tensor([0.0, 333.25, 998.999]).cuda(), context=ctx)
        assert rel_l2(y, g['y_flt']) < TIGHT


def test_unet_full_matches_reference_golden():
    g = golden('unet_clevrtex')
    net, sd, cfg = make_unet()
    x, ctx = seeded((2, 3, 32, 32), 41).cuda(), seeded((2, 11, 192), 42).cuda()
    with torch.no_grad():
        y = net(x, torch.tensor([7, 503]).cuda(), context=ctx)
        e = rel_l2(y, g['y_int'])
        assert e < TOL and e < TIGHT, e
        y = net(x, torch.tensor([0.0, 333.25]).cuda(), context=ctx)
        assert rel_l2(y, g['y_flt']) < TIGHT
        # single-pass fp16 mode is the AMP/TF32 accuracy class, reported not gated at 1e-3
        from slotdiffusion_b200 import ops
        ops.set_precision('fp16')
        try:
            y1 = net(x, torch.tensor([7, 503]).cuda(), context=ctx)
        finally:
            ops.set_precision('fp32')
        assert rel_l2(y1, g['y_int']) < 5e-3


def test_unet_batch64_properties():
    """BASELINE batch size: batch independence + agreement with the oracle on a sub-batch."""
    net, sd, cfg = make_unet()
    B = 64
    x, ctx = seeded((B, 3, 32, 32), 71).cuda(), seeded((B, 11, 192), 72).cuda()
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(73)).cuda()
    with torch.no_grad():
        y = net(x, t, context=ctx)
        assert torch.isfinite(y).all()
        y2 = net(x[10:13], t[10:13], context=ctx[10:13])
        # the batch size selects tiling / CTA pairing / split-K / fused-vs-separate GroupNorm statistics, i.e. a different
        # fp32 summation order: agreement is at accumulation round-off, not bit level
        assert rel_l2(y2, y[10:13]) < 3e-5
        ref = unet_ref.unet_forward(sd, x[10:12].cpu(), t[10:12].cpu(), ctx[10:12].cpu())
    assert rel_l2(y[10:12], ref) < TIGHT


def test_unet_coco_latent56_ctx256_matches_oracle():
    """BASELINE configs[4] geometry (DINOSAUR+SlotDiffusion COCO): 56x56 latents, 7 slots of size 256 -- non-power-of-2
    widths (56/28/14/7) through the implicit-GEMM convolutions and 3136/784/196/49-token attention."""
    cfg_over = dict(context_dim=256)
    net, sd, cfg = make_unet(cfg_over, seed=33)
    x, ctx = seeded((1, 3, 56, 56), 81).cuda(), seeded((1, 7, 256), 82).cuda()
    t = torch.tensor([421.5]).cuda()
    with torch.no_grad():
        y = net(x, t, context=ctx)
    ref = unet_ref.unet_forward(sd, x.cpu(), t.cpu(), ctx.cpu(), cfg)
    assert rel_l2(y, ref) < TIGHT


def test_unet_movie_24slots_matches_oracle():
    """BASELINE configs[3] geometry (MOVi-E): 24 slots of size 192 as cross-attention context."""
    net, sd, cfg = make_unet(seed=34)
    x, ctx = seeded((2, 3, 32, 32), 83).cuda(), seeded((2, 24, 192), 84).cuda()
    t = torch.tensor([3, 977]).cuda()
    with torch.no_grad():
        y = net(x, t, context=ctx)
    ref = unet_ref.unet_forward(sd, x.cpu(), t.cpu(), ctx.cpu(), cfg)
    assert rel_l2(y, ref) < TIGHT


def test_savi_per_frame_driver_matches_oracle():
    """a6 (savi_diffusion.py:183-196): the module is called once per frame with the previous frame's slots; T = 3 frames,
    15 slots, 2 iterations (the shipped MOVi config)."""
    mod, p, x, s0, gw, iters = make_sa('sa_vid_movid')
    T = 3
    frames = [seeded(tuple(x.shape), 90 + f) for f in range(T)]
    slots, ref = s0.cuda(), s0.double()
    with torch.no_grad():
        for f in range(T):
            slots, mask = mod(frames[f].cuda(), slots)
            ref, ref_m = sa_ref.slot_attention_forward(p, frames[f].double(), ref, iters)
            assert rel_l2(slots, ref) < TIGHT
            margin = ref_m.topk(2, dim=1).values
            real, near = argmax_mismatch(mask, ref_m.argmax(1).numpy(), (margin[:, 0] - margin[:, 1]).numpy(), 1e-5)
            assert real == 0, (f, real, near)


def test_dpm_sampler_matches_reference_golden():
    from slotdiffusion_b200.dpm_solver import DPMSolverSampler
    g = golden('dpm')
    net, sd, cfg = make_unet()
    betas = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
    ctx, xT = seeded((1, 11, 192), 52).cuda(), seeded((1, 3, 32, 32), 53).cuda()
    for graph in (False, True):
        smp = DPMSolverSampler(net, betas, codebook=None, use_cuda_graph=graph)
        assert smp.nfe == 20
        np.testing.assert_allclose(np.array(smp.t_model), g['t_model'], rtol=1e-6)
        y = smp.sample(xT, ctx)
        assert rel_l2(y, g['sample_novq']) < 2e-4       # 20 chained UNet evaluations
    # vq_denoised, free-running: ONE flipped nearest-code decision changes the trajectory of every later evaluation, so
    # the end-to-end comparison can only be statistical; the exact statement is the teacher-forced test below
    cb = seeded((4096, 3), 51).cuda()
    y = DPMSolverSampler(net, betas, codebook=cb, use_cuda_graph=True).sample(xT, ctx)
    diff = (y.cpu() - torch.as_tensor(g['sample_vq'])).abs().amax(1)
    assert (diff > 1e-3).float().mean().item() < 0.02


def test_dpm_vq_decisions_teacher_forced_exact_outside_near_ties():
    """Every nearest-code decision of a vq_denoised sampling run (dpm_solver.py:523-534, quantize.py:84-94), checked one
    evaluation at a time on the ORACLE's trajectory: the B200 UNet + fused x0/VQ kernel see the oracle's latent of
    evaluation k, and their 1024 code indices must equal the fp64 nearest codes of the oracle's x0 -- except where the
    bisector distance is below what an eps error of 1e-4 (a tenth of the 1e-3 contract) moves x0: dz = sigma/alpha * 1e-4.
    real == 0 over all decisions; the count of near-tie flips is reported."""
    from helpers import vq_mismatch
    net, sd, cfg = make_unet()
    betas = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
    ctx, xT = seeded((1, 11, 192), 52), seeded((1, 3, 32, 32), 53)
    cb = seeded((4096, 3), 51)
    steps = 8                                                     # 8 NFE: keeps the CPU oracle to ~20 s
    ns = dpm_ref.NoiseScheduleVP(betas)
    evals = [e for st in dpm_ref.dpm_coefficients(ns, steps) for e in st['evals']]
    _, trace = dpm_ref.dpm_sample(lambda x, t, c: unet_ref.unet_forward(sd, x, t, c, cfg), betas, xT, ctx, codebook=cb,
                                  steps=steps, return_trace=True)
    assert len(trace) == len(evals) == steps
    tot_real = tot_near = 0
    with torch.no_grad():
        for e, tr in zip(evals, trace):
            alpha, sigma = float(e['alpha']), float(e['sigma'])
            x = tr['x'].cuda()
            eps = net(x, e['t_model'].expand(1).float().cuda(), context=ctx.cuda())
            assert rel_l2(eps, tr['eps']) < TIGHT
            x0, idx = ops_mod.dpm_x0(x, eps, alpha, sigma, cb.cuda(), want_idx=True)
            z = ((tr['x'].double() - sigma * tr['eps'].double()) / alpha).permute(0, 2, 3, 1).reshape(-1, 3)
            real, near, ref = vq_mismatch(idx, z, cb, sigma / alpha * 1e-4)
            tot_real += real
            tot_near += near
            ok = (idx.cpu().flatten().long() == ref)
            assert torch.equal(x0.cpu().permute(0, 2, 3, 1).reshape(-1, 3)[ok], cb[ref[ok]])
    print('VQ decisions:', steps * 1024, 'near-tie flips:', tot_near)
    assert tot_real == 0, (tot_real, tot_near)
    assert tot_near <= 0.02 * steps * 1024
