"""Graphed training (slotdiffusion_b200/graphed.py): the modules' training forward / backward replayed from CUDA graphs inside
an eager loop give the same outputs and gradients as the launch-by-launch path -- also after optimizer steps (weights are
re-packed inside the captured forward) and with new inputs (static input copies), with fresh dropout masks per replay."""
import pytest
import torch

from helpers import rel_l2, seeded
from oracle import resnet_ref, unet_ref
from oracle import slot_attention_ref as sa_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _grads(mod):
    return {k: p.grad.detach().clone() for k, p in mod.named_parameters() if p.grad is not None}


def _compare_three_steps(make, run, lr=1e-4):
    """two identically initialised modules, one graphed: three SGD steps on different inputs stay in lock-step"""
    from slotdiffusion_b200 import graphed
    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    graphed.enable(b)
    oa, ob = torch.optim.SGD(a.parameters(), lr=lr), torch.optim.SGD(b.parameters(), lr=lr)
    for step in range(3):
        la, lb = run(a, step), run(b, step)
        assert rel_l2(lb, la) < 1e-4, (step, rel_l2(lb, la))        # atomics make two runs differ by ~1e-7; SGD steps amplify it
        oa.zero_grad(set_to_none=True)
        ob.zero_grad(set_to_none=True)
        la.backward()
        lb.backward()
        ga, gb = _grads(a), _grads(b)
        assert set(ga) == set(gb)
        # parameters whose gradient is zero up to round-off (LayerNorm_q's bias: the softmax over slots ignores a per-token
        # shift of the logits) carry only noise -- 1e-6 against gradients of 1e+1 elsewhere -- and are not compared
        floor = 1e-5 * max(float(g.norm()) for g in ga.values())
        worst = max((rel_l2(gb[k], ga[k]) for k in ga if ga[k].norm() > floor), default=0.0)
        assert worst < 1e-3, (step, worst)
        oa.step()
        ob.step()
    assert b.__dict__['_sdb_graphs'].graphs, 'the graphed path was not taken'
    return a, b


def test_unet_graphed_training_matches_eager():
    from slotdiffusion_b200.unet import UNetModel
    cfg = dict(unet_ref.DEFAULT_CFG, model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1,
               context_dim=64)
    sd = unet_ref.random_state_dict(cfg, seed=31)

    def make():
        net = UNetModel(dropout=0.0, dims=2, use_checkpoint=False, resblock_updown=False, conv_resample=True,
                        transformer_depth=1, n_embed=None, **cfg).cuda().train()
        net.load_state_dict(sd)
        return net

    def run(net, step):
        x, ctx = seeded((3, 3, 16, 16), 100 + step).cuda(), seeded((3, 5, 64), 200 + step).cuda().requires_grad_(True)
        t = torch.tensor([7, 503, 999]).cuda() + step
        return (net(x, t, context=ctx) * seeded((3, 3, 16, 16), 300 + step).cuda()).sum()
    _compare_three_steps(make, run)


def test_slot_attention_graphed_training_matches_eager():
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    p = sa_ref.random_params(192, 192, 384, seed=11)

    def make():
        m = SlotAttentionWMask(192, 2, 5, 192, 384).cuda().train()
        m.load_state_dict(p)
        return m

    def run(m, step):
        x, s0 = seeded((2, 96, 192), 10 + step).cuda().requires_grad_(True), seeded((2, 5, 192), 20 + step).cuda().requires_grad_(True)
        slots, mask = m(x, s0)
        assert mask.shape == (2, 5, 96) and not mask.requires_grad
        return (slots * seeded((2, 5, 192), 30 + step).cuda()).sum()
    _compare_three_steps(make, run, lr=1e-4)


def test_resnet_graphed_training_matches_eager():
    from slotdiffusion_b200 import resnet
    sd = resnet_ref.random_state_dict('resnet18', False, seed=5)

    def make():
        n = resnet.resnet18(small_inputs=True, use_layer4=False).cuda().train()
        n.load_state_dict(sd)
        return n

    def run(n, step):
        y = n(seeded((2, 3, 32, 32), 40 + step).clamp(-1, 1).cuda())
        return (y * seeded(tuple(y.shape), 50 + step).cuda()).sum() * 1e-3
    _compare_three_steps(make, run, lr=1e-4)


def test_graphed_dropout_draws_new_masks_per_replay():
    from slotdiffusion_b200 import graphed
    from slotdiffusion_b200.unet import UNetModel
    cfg = dict(unet_ref.DEFAULT_CFG, model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1,
               context_dim=64)
    net = UNetModel(dropout=0.3, dims=2, use_checkpoint=False, resblock_updown=False, conv_resample=True,
                    transformer_depth=1, n_embed=None, **cfg).cuda().train()
    net.load_state_dict(unet_ref.random_state_dict(cfg, seed=31))
    graphed.enable(net)
    x, ctx, t = seeded((2, 3, 16, 16), 1).cuda(), seeded((2, 5, 64), 2).cuda().requires_grad_(True), torch.tensor([5, 900]).cuda()
    y1 = net(x, t, context=ctx).detach().clone()
    y2 = net(x, t, context=ctx).detach().clone()
    assert rel_l2(y2, y1) > 1e-3                   # same inputs, same graph, different dropout masks
    net.eval()
    with torch.no_grad():
        e1, e2 = net(x, t, context=ctx), net(x, t, context=ctx)
    assert rel_l2(e2, e1) < 1e-6                   # eval: no dropout (floating-point atomics in the GroupNorm sums only)
