"""GPU parity of the TRAINING path (forward + backward through the C-ABI kernels) against autograd over the CPU oracle
and the golden gradient fixtures generated from the reference modules.  Gradients: <= 1e-3 relative (north_star),
the 3-pass split-fp16 path achieves ~1e-5."""
import numpy as np
import pytest
import torch

from helpers import SA_CASES, checksum, golden, rel_l2, sa_case, seeded
from oracle import unet_ref
from oracle import slot_attention_ref as sa_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GTOL = 2e-4


@pytest.mark.parametrize('name', ['sa_ragged_small', 'sa_coco_vitb16', 'sa_img_clevrtex'])
def test_slot_attention_gradients_match_oracle(name):
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, Din, S, D, M, I = SA_CASES[name]
    p, x, s0, gw, iters = sa_case(name)
    mod = SlotAttentionWMask(Din, I, S, D, M).cuda().train()
    mod.load_state_dict(p)
    xg, sg = x.cuda().requires_grad_(True), s0.cuda().requires_grad_(True)
    slots, mask = mod(xg, sg)
    assert not mask.requires_grad
    (slots * gw.cuda()).sum().backward()
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    x64, s64 = x.double().requires_grad_(True), s0.double().requires_grad_(True)
    ref, ref_mask = sa_ref.slot_attention_forward(p64, x64, s64, iters)
    (ref * gw.double()).sum().backward()
    assert rel_l2(slots, ref) < 5e-5 and rel_l2(mask, ref_mask) < 5e-5
    assert rel_l2(xg.grad, x64.grad) < GTOL, 'd inputs'
    assert rel_l2(sg.grad, s64.grad) < GTOL, 'd slots'
    for k, v in mod.named_parameters():
        r = p64[k].grad
        if r.norm().item() < 1e-9 * max(1.0, p64[k].norm().item()):       # mathematically zero (softmax shift invariance)
            assert v.grad.norm().item() < 1e-5, k
        else:
            assert rel_l2(v.grad, r) < GTOL, k
    if name in ('sa_ragged_small', 'sa_coco_vitb16'):                      # reference-generated fixture
        g = golden(name)
        assert rel_l2(sg.grad, g['grad_slots']) < GTOL


def _unet(cfg_over=None, seed=31, dropout=0.0, train=False):
    from slotdiffusion_b200.unet import UNetModel
    cfg = dict(unet_ref.DEFAULT_CFG, **(cfg_over or {}))
    sd = unet_ref.random_state_dict(cfg, seed=seed)
    net = UNetModel(dropout=dropout, dims=2, use_checkpoint=False, resblock_updown=False, conv_resample=True,
                    transformer_depth=1, n_embed=None, **cfg).cuda()
    net.load_state_dict(sd)
    net.train(train)
    return net, sd, cfg


def test_unet_small_gradients_match_oracle():
    cfg_over = dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1, context_dim=64)
    net, sd, cfg = _unet(cfg_over)
    x, ctx = seeded((3, 3, 16, 16), 41), seeded((3, 5, 64), 42)
    t = torch.tensor([7, 503, 999])
    gw = seeded((3, 3, 16, 16), 43)
    xg, cg = x.cuda().requires_grad_(True), ctx.cuda().requires_grad_(True)
    y = net(xg, t.cuda(), context=cg)
    (y * gw.cuda()).sum().backward()
    sdg = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    x64, c64 = x.double().requires_grad_(True), ctx.double().requires_grad_(True)
    ref = unet_ref.unet_forward(sdg, x64, t, c64, cfg)
    (ref * gw.double()).sum().backward()
    assert rel_l2(y, ref) < 5e-5
    assert rel_l2(cg.grad, c64.grad) < GTOL, 'd context'
    assert rel_l2(xg.grad, x64.grad) < GTOL, 'd x'
    worst = 0.0
    for k, v in net.named_parameters():
        assert v.grad is not None, k
        e = rel_l2(v.grad, sdg[k].grad)
        worst = max(worst, e)
        assert e < GTOL, (k, e)
    print('worst parameter-gradient rel err', worst)


def test_unet_full_gradients_match_reference_golden():
    g = golden('unet_clevrtex')
    net, sd, cfg = _unet()
    x = seeded((2, 3, 32, 32), 41).cuda()
    ctx = seeded((2, 11, 192), 42).cuda().requires_grad_(True)
    y = net(x, torch.tensor([7, 503]).cuda(), context=ctx)
    assert rel_l2(y, g['y_int']) < 5e-5
    (y * seeded(tuple(y.shape), 43).cuda()).sum().backward()
    assert rel_l2(ctx.grad, g['grad_ctx']) < GTOL
    params = dict(net.named_parameters())
    n = 0
    for k in g.files:
        if k.startswith('gsum.'):
            np.testing.assert_allclose(checksum(params[k[5:]].grad)[1], g[k][1], rtol=1e-3, err_msg=k)
            n += 1
    assert n > 0


def test_video_recurrence_gradients_match_oracle():
    """SAViDiffusion.encode (savi_diffusion.py:183-196) calls the SAME module once per frame before a single backward:
    every call owns its gradient buffer (a module-level buffer returned the last frame's gradient T times)."""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, T, N, S, D, I = 2, 3, 96, 5, 192, 2
    p = sa_ref.random_params(D, D, 2 * D, seed=11)
    mod = SlotAttentionWMask(D, I, S, D, 2 * D).cuda().train()
    mod.load_state_dict(p)
    frames, s0, gw = seeded((B, T, N, D), 61), seeded((B, S, D), 62), seeded((B, T, S, D), 63)
    wp = seeded((D, D), 64) * D ** -0.5                      # stand-in for the predictor: a linear map of the slots
    fg, sg = frames.cuda().requires_grad_(True), s0.cuda().requires_grad_(True)
    prev, outs = None, []
    for t in range(T):
        init = sg if prev is None else prev @ wp.cuda()
        prev, mask = mod(fg[:, t], init)
        outs.append(prev)
    (torch.stack(outs, 1) * gw.cuda()).sum().backward()
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    f64, s64 = frames.double().requires_grad_(True), s0.double().requires_grad_(True)
    ref, _ = sa_ref.slot_attention_video(p64, f64, s64, I, predictor=lambda s: s @ wp.double())
    (ref * gw.double()).sum().backward()
    assert rel_l2(torch.stack(outs, 1), ref) < 5e-5
    assert rel_l2(fg.grad, f64.grad) < GTOL and rel_l2(sg.grad, s64.grad) < GTOL
    for k, v in mod.named_parameters():
        if p64[k].grad.norm().item() > 1e-9 * max(1.0, p64[k].norm().item()):
            assert rel_l2(v.grad, p64[k].grad) < GTOL, k


def test_unet_two_forwards_before_one_backward():
    """two evaluations in one autograd graph (two loss terms / two timesteps): per-call gradient storage"""
    cfg_over = dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1, context_dim=64)
    net, sd, cfg = _unet(cfg_over)
    xa, xb, ctx = seeded((2, 3, 16, 16), 41), seeded((2, 3, 16, 16), 44), seeded((2, 5, 64), 42)
    ta, tb = torch.tensor([7, 503]), torch.tensor([950, 20])
    gwa, gwb = seeded((2, 3, 16, 16), 43), seeded((2, 3, 16, 16), 45)
    xag, xbg, cg = xa.cuda().requires_grad_(True), xb.cuda().requires_grad_(True), ctx.cuda().requires_grad_(True)
    ya = net(xag, ta.cuda(), context=cg)
    yb = net(xbg, tb.cuda(), context=cg)
    ((ya * gwa.cuda()).sum() + (yb * gwb.cuda()).sum()).backward()
    sdg = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    xa64, xb64, c64 = xa.double().requires_grad_(True), xb.double().requires_grad_(True), ctx.double().requires_grad_(True)
    ra, rb = unet_ref.unet_forward(sdg, xa64, ta, c64, cfg), unet_ref.unet_forward(sdg, xb64, tb, c64, cfg)
    ((ra * gwa.double()).sum() + (rb * gwb.double()).sum()).backward()
    assert rel_l2(xag.grad, xa64.grad) < GTOL and rel_l2(xbg.grad, xb64.grad) < GTOL and rel_l2(cg.grad, c64.grad) < GTOL
    for k, v in net.named_parameters():
        assert rel_l2(v.grad, sdg[k].grad) < GTOL, k


def test_unet_dropout_backward_uses_the_forward_mask(monkeypatch):
    """Model-level mask consistency: with the dropout seed stream pinned (same masks in every forward), the analytic
    directional derivative of the train-mode loss equals its central finite difference.  A backward that regenerated a
    DIFFERENT mask would be off by O(1).  (Keep rate and the 1/(1-p) scaling are checked at op level:
    test_backward_ops_gpu.py::test_groupnorm_bwd with p = 0.1.)"""
    import itertools
    import slotdiffusion_b200.backward as bwd
    cfg_over = dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1, context_dim=64)
    net, sd, cfg = _unet(cfg_over, dropout=0.3, train=True)
    x, ctx = seeded((4, 3, 16, 16), 41).cuda(), seeded((4, 5, 64), 42).cuda()
    t = torch.tensor([7, 503, 999, 1]).cuda()
    gw = seeded((4, 3, 16, 16), 48).cuda()

    def loss_at(c):
        monkeypatch.setattr(bwd, '_step_counter', itertools.count(1234))       # same seed -> same masks
        return (net(x, t, context=c).double() * gw).sum()
    cg = ctx.clone().requires_grad_(True)
    loss_at(cg).backward()
    d = cg.grad / cg.grad.norm()                  # steepest direction: the derivative along it is |grad|
    analytic = (cg.grad * d).sum().item()
    h = 1e-2
    with torch.enable_grad():
        lp = loss_at((ctx + h * d).requires_grad_(True)).item()
        lm = loss_at((ctx - h * d).requires_grad_(True)).item()
    fd = (lp - lm) / (2 * h)
    assert abs(analytic) > 1e-3
    assert abs(fd - analytic) / abs(analytic) < 3e-2, (fd, analytic)


def test_unet_dropout_train_mode():
    """nn.Dropout(p) in every ResBlock (unet.py:245-246) is active in train mode: finite output and gradients, and the
    output differs from eval mode by a plausible amount (mask statistics: test_backward_ops_gpu.py, mask consistency
    between forward and backward: test_unet_dropout_backward_uses_the_forward_mask)."""
    cfg_over = dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1, context_dim=64)
    net, sd, cfg = _unet(cfg_over, dropout=0.1, train=True)
    x, ctx = seeded((4, 3, 16, 16), 41).cuda(), seeded((4, 5, 64), 42).cuda().requires_grad_(True)
    t = torch.tensor([7, 503, 999, 1]).cuda()
    y1 = net(x, t, context=ctx)
    y1.square().mean().backward()
    assert torch.isfinite(y1).all() and all(torch.isfinite(p.grad).all() for p in net.parameters())
    with torch.no_grad():
        net.eval()
        y0 = net(x, t, context=ctx)
    d = rel_l2(y1, y0)
    assert 1e-3 < d < 1.0, d


def test_denoise_loss_end_to_end():
    """LDM.loss_function shape (ldm.py:59-83): slots = SlotAttention(features); x_t = q_sample(x0, t, eps);
    loss = mse(UNet(x_t, t, slots), eps); gradients reach the UNet, Slot Attention and the encoder features."""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from oracle import dpm_ref
    cfg_over = dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1, context_dim=64)
    net, sd, cfg = _unet(cfg_over)
    B, N, S, D = 3, 64, 5, 64
    p = sa_ref.random_params(D, D, 2 * D, seed=11)
    sa = SlotAttentionWMask(D, 2, S, D, 2 * D).cuda().train()
    sa.load_state_dict(p)
    feats, s0 = seeded((B, N, D), 1), seeded((B, S, D), 2)
    x0, eps = seeded((B, 3, 16, 16), 3), seeded((B, 3, 16, 16), 4)
    t = torch.tensor([10, 500, 900])
    buf = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())
    a, s = buf['sqrt_alphas_bar'][t].view(B, 1, 1, 1), buf['sqrt_one_minus_alphas_bar'][t].view(B, 1, 1, 1)
    xt = a * x0 + s * eps
    fg = feats.cuda().requires_grad_(True)
    slots, _ = sa(fg, s0.cuda())
    loss = torch.nn.functional.mse_loss(net(xt.cuda(), t.cuda(), context=slots), eps.cuda())
    loss.backward()
    # oracle
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    sdg = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    f64 = feats.double().requires_grad_(True)
    rs, _ = sa_ref.slot_attention_forward(p64, f64, s0.double(), 2)
    rl = torch.nn.functional.mse_loss(unet_ref.unet_forward(sdg, xt.double(), t, rs, cfg), eps.double())
    rl.backward()
    assert abs(loss.item() - rl.item()) / rl.item() < 1e-4
    assert rel_l2(fg.grad, f64.grad) < 5e-4
    for k, v in sa.named_parameters():
        if p64[k].grad.norm().item() > 1e-12:
            assert rel_l2(v.grad, p64[k].grad) < 5e-4, k
    for k, v in net.named_parameters():
        assert rel_l2(v.grad, sdg[k].grad) < 5e-4, k
