"""GPU parity of the slot transition function: csrc/predictor.cu ops through the C ABI vs fp64 torch math,
slotdiffusion_b200.predictor.TransformerPredictor vs the golden outputs / gradients of the unmodified reference module
(tests/golden/predictor.npz) and vs the fp64 oracle, dropout statistics and forward/backward mask consistency, and the
per-frame recurrence (predictor -> Slot Attention over T frames, savi_diffusion.py:183-196) against the oracle."""
import itertools

import numpy as np
import pytest
import torch

from helpers import checksum, golden, rel_l2, seeded
from oracle import predictor_ref
from oracle import slot_attention_ref as sa_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GTOL = 2e-4          # contract: 1e-3 relative fp32 (north_star)

PRED_CASES = {
    'movid': (3, 15, 192, 2, 4, 768, True),
    'clevrer': (2, 7, 128, 2, 4, 512, True),
    'postln': (2, 11, 256, 1, 4, 512, False),
}


def _attn64(qkv, B, S, heads, mask=None):
    D = qkv.shape[1] // 3
    dh = D // heads
    q, k, v = (t.reshape(B, S, heads, dh).transpose(1, 2) for t in qkv.double().split(D, dim=-1))
    a = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
    if mask is not None:
        a = a * mask
    return (a @ v).transpose(1, 2).reshape(B * S, D), a


@pytest.mark.parametrize('B,S,heads,dh', [(3, 11, 4, 48), (2, 32, 4, 32), (5, 1, 2, 64), (64, 15, 4, 48), (2, 24, 8, 32)])
def test_token_attention_forward_and_backward(B, S, heads, dh):
    from slotdiffusion_b200 import ops
    D = heads * dh
    qkv = seeded((B * S, 3 * D), 5).cuda()
    out = ops.token_attention(qkv, B, S, heads)
    q64 = qkv.double().requires_grad_(True)
    ref, _ = _attn64(q64, B, S, heads)
    assert rel_l2(out, ref) < 2e-6
    dout = seeded((B * S, D), 6).cuda()
    (ref * dout.double()).sum().backward()
    dqkv = ops.token_attention_bwd(qkv, dout, B, S, heads)
    assert rel_l2(dqkv, q64.grad) < 5e-6
    # strided rows (the fused projection may live inside a wider buffer)
    wide = torch.zeros(B * S, 3 * D + 16, device='cuda')
    wide[:, :3 * D] = qkv
    assert torch.equal(ops.token_attention(wide[:, :3 * D], B, S, heads), out)


def test_token_attention_dropout_mask_is_shared_by_forward_and_backward():
    """nn.MultiheadAttention drops attention probabilities (p = 0.1 in nn.TransformerEncoderLayer).  With v = one-hot rows the
    output IS the masked probability matrix, which recovers the mask; the backward must differentiate through the same mask."""
    from slotdiffusion_b200 import ops
    B, S, heads, dh, p, seed = 40, 16, 4, 32, 0.25, 12345
    D = heads * dh
    qkv = seeded((B * S, 3 * D), 7)
    v = torch.zeros(B, S, heads, dh)
    for j in range(S):
        v[:, j, :, j] = 1.0
    qkv[:, 2 * D:] = v.reshape(B * S, D)
    qkv = qkv.cuda()
    out = ops.token_attention(qkv, B, S, heads, p, seed)
    _, a = _attn64(qkv, B, S, heads)                                    # [B, heads, S, S] unmasked probabilities
    pm = out.double().reshape(B, S, heads, dh).transpose(1, 2)[..., :S]     # masked probabilities
    ratio = pm / a
    kept = ratio > 0.5
    assert torch.allclose(ratio[kept], torch.full_like(ratio[kept], 1 / (1 - p)), rtol=1e-5)      # survivors scaled by 1/(1-p)
    assert (ratio[~kept].abs() < 1e-12).all()
    keep_rate = kept.double().mean().item()
    n = kept.numel()
    assert abs(keep_rate - (1 - p)) < 5 * np.sqrt(p * (1 - p) / n), keep_rate
    assert torch.equal(ops.token_attention(qkv, B, S, heads, p, seed), out)                 # counter based: repeatable
    assert not torch.equal(ops.token_attention(qkv, B, S, heads, p, seed + 1), out)
    # backward against fp64 autograd through the recovered mask
    mask = kept.double() / (1 - p)
    q64 = qkv.double().requires_grad_(True)
    ref, _ = _attn64(q64, B, S, heads, mask)
    dout = seeded((B * S, D), 8).cuda()
    (ref * dout.double()).sum().backward()
    dqkv = ops.token_attention_bwd(qkv, dout, B, S, heads, p, seed)
    assert rel_l2(dqkv, q64.grad) < 5e-6


def test_dropout_add_statistics_and_adjoint():
    from slotdiffusion_b200 import ops
    n, p, seed = 1 << 20, 0.1, 77
    x, res = seeded((n,), 9).cuda(), seeded((n,), 10).cuda()
    y = ops.dropout_add(x, None, p, seed)
    kept = y != 0
    assert torch.allclose(y[kept], x[kept] / (1 - p), rtol=1e-6)                    # survivors scaled by 1/(1-p)
    assert abs(kept.double().mean().item() - (1 - p)) < 5 * np.sqrt(p * (1 - p) / n)
    assert torch.allclose(ops.dropout_add(x, res, p, seed), res + y, rtol=1e-6, atol=1e-6)      # + residual (one FMA in the kernel)
    dy = seeded((n,), 11).cuda()
    dx = ops.dropout_add(dy, None, p, seed)                      # backward = the same mask applied to dy
    assert torch.equal(dx != 0, kept)
    assert torch.allclose(dx[kept], dy[kept] / (1 - p), rtol=1e-6)
    assert torch.equal(ops.dropout_add(x, None, 0.0, seed), x)
    assert not torch.equal(ops.dropout_add(x, None, p, seed + 1) != 0, kept)


def _module(name, train=False):
    from slotdiffusion_b200.predictor import TransformerPredictor
    B, S, D, L, Hh, F, nf = PRED_CASES[name]
    sd = predictor_ref.random_state_dict(D, L, F, seed=900 + len(name))
    net = TransformerPredictor(d_model=D, num_layers=L, num_heads=Hh, ffn_dim=F, norm_first=nf)
    net.load_state_dict(sd, strict=True)                     # reference-keyed state_dict
    net = net.cuda()
    net.train(train)
    return net, sd


@pytest.mark.parametrize('name', sorted(PRED_CASES))
def test_predictor_matches_reference_golden(name):
    g = golden('predictor')
    B, S, D, L, Hh, F, nf = PRED_CASES[name]
    net, sd = _module(name)
    x = seeded((B, S, D), 91)
    with torch.no_grad():
        y0 = net(x.cuda())                                   # inference schedule
    assert rel_l2(y0, g[name + '.y']) < 5e-6
    xg = x.cuda().requires_grad_(True)
    y = net(xg)                                              # training schedule (tape)
    assert rel_l2(y, g[name + '.y']) < 5e-6
    (y * seeded((B, S, D), 92).cuda()).sum().backward()
    assert rel_l2(xg.grad, g[name + '.dx']) < GTOL
    for k, v in net.named_parameters():
        ref = g[name + '.grad.' + k]
        if v.grad.dim() == 1:
            assert rel_l2(v.grad, ref) < GTOL, k
        else:
            got = checksum(v.grad)
            assert abs(got[0] - ref[0]) <= 1e-3 * max(1.0, np.sqrt(ref[1])) and abs(got[1] - ref[1]) <= 1e-3 * ref[1], k


def test_predictor_full_size_matches_fp64_oracle():
    """MOVi-D geometry at the bench's clip batch: B = 64 clips, 11 slots, D = 192; every gradient tensor in full."""
    from slotdiffusion_b200.predictor import TransformerPredictor
    B, S, D, L, Hh, F = 64, 11, 192, 2, 4, 768
    sd = predictor_ref.random_state_dict(D, L, F, seed=17)
    net = TransformerPredictor(D, L, Hh, F, True).cuda().eval()
    net.load_state_dict(sd)
    x, gw = seeded((B, S, D), 93), seeded((B, S, D), 94)
    xg = x.cuda().requires_grad_(True)
    y = net(xg)
    (y * gw.cuda()).sum().backward()
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    ref = predictor_ref.predictor_forward(p64, x64, L, Hh, True)
    (ref * gw.double()).sum().backward()
    assert rel_l2(y, ref) < 5e-6
    assert rel_l2(xg.grad, x64.grad) < GTOL
    for k, v in net.named_parameters():
        assert rel_l2(v.grad, p64[k].grad) < GTOL, k
    # batch independence: a clip's result does not depend on its neighbours (the GEMM tiling follows the row count, so
    # equal to fp32 round-off, not bit for bit)
    with torch.no_grad():
        assert rel_l2(net(x[5:6].cuda()), net(x.cuda())[5:6]) < 1e-6


def test_predictor_train_mode_dropout(monkeypatch):
    """nn.TransformerEncoderLayer(dropout=0.1) is active in train mode (three sites + attention probabilities).  With the
    seed stream pinned the masks repeat, and the analytic directional derivative equals the central finite difference --
    a backward that regenerated different masks would be off by O(1)."""
    import slotdiffusion_b200.predictor as pr
    net, sd = _module('movid', train=True)
    B, S, D = 16, 15, 192
    x, gw = seeded((B, S, D), 95).cuda(), seeded((B, S, D), 96).cuda().double()

    def loss_at(xx):
        monkeypatch.setattr(pr, '_call_counter', itertools.count(4321))
        return (net(xx).double() * gw).sum()
    xg = x.clone().requires_grad_(True)
    l0 = loss_at(xg)
    l0.backward()
    assert all(torch.isfinite(q.grad).all() for q in net.parameters())
    with torch.enable_grad():
        l1 = loss_at(x.clone().requires_grad_(True)).item()
    assert abs(l1 - l0.item()) < 1e-6 * abs(l0.item())          # pinned seed -> identical masks (a different mask moves it by O(1))
    d = xg.grad / xg.grad.norm()
    analytic = (xg.grad * d).sum().item()
    h = 1e-2
    lp = loss_at((x + h * d).requires_grad_(True)).item()
    lm = loss_at((x - h * d).requires_grad_(True)).item()
    fd = (lp - lm) / (2 * h)
    assert abs(analytic) > 1e-3
    assert abs(fd - analytic) / abs(analytic) < 3e-2, (fd, analytic)
    with torch.no_grad():
        y1 = net(x)                                          # train mode without grad: dropout still applies
        net.eval()
        y0 = net(x)
    assert 1e-3 < rel_l2(y1, y0) < 1.0


def test_video_recurrence_with_predictor_matches_oracle():
    """SAViDiffusion.encode (savi_diffusion.py:183-196): slots of frame t+1 start from predictor(slots of frame t); one backward
    through T Slot-Attention calls and T-1 predictor calls of the SAME modules (per-call gradient storage)."""
    from slotdiffusion_b200.predictor import TransformerPredictor
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, T, N, S, D, I, L, Hh = 2, 4, 96, 7, 192, 2, 2, 4
    p = sa_ref.random_params(D, D, 2 * D, seed=11)
    sa = SlotAttentionWMask(D, I, S, D, 2 * D).cuda().train()
    sa.load_state_dict(p)
    psd = predictor_ref.random_state_dict(D, L, 4 * D, seed=19)
    pred = TransformerPredictor(D, L, Hh, 4 * D, True).cuda().eval()       # eval: dropout off, gradients still flow
    pred.load_state_dict(psd)
    frames, s0, gw = seeded((B, T, N, D), 61), seeded((B, S, D), 62), seeded((B, T, S, D), 63)
    fg, sg = frames.cuda().requires_grad_(True), s0.cuda().requires_grad_(True)
    prev, outs = None, []
    for t in range(T):
        init = sg if prev is None else pred(prev)
        prev, _ = sa(fg[:, t], init)
        outs.append(prev)
    (torch.stack(outs, 1) * gw.cuda()).sum().backward()
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    q64 = {k: v.double().clone().requires_grad_(True) for k, v in psd.items()}
    f64, s64 = frames.double().requires_grad_(True), s0.double().requires_grad_(True)
    ref, _ = sa_ref.slot_attention_video(p64, f64, s64, I,
                                         predictor=lambda s: predictor_ref.predictor_forward(q64, s, L, Hh, True))
    (ref * gw.double()).sum().backward()
    assert rel_l2(torch.stack(outs, 1), ref) < 5e-5
    assert rel_l2(fg.grad, f64.grad) < GTOL and rel_l2(sg.grad, s64.grad) < GTOL
    for k, v in sa.named_parameters():
        if p64[k].grad.norm().item() > 1e-9 * max(1.0, p64[k].norm().item()):
            assert rel_l2(v.grad, p64[k].grad) < GTOL, k
    for k, v in pred.named_parameters():
        assert rel_l2(v.grad, q64[k].grad) < GTOL, k


def test_predictor_rejects_cpu_and_unsupported_geometry():
    from slotdiffusion_b200.predictor import TransformerPredictor
    net = TransformerPredictor(128, 1, 4, 256)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.zeros(1, 5, 128))
    net = net.cuda()
    with pytest.raises(RuntimeError, match='unsupported geometry'):
        net(torch.zeros(1, 33, 128, device='cuda'))
