"""Host logic of the kernel schedules without a GPU: every schedule (UNet forward, UNet training step, Slot Attention
inference / one-launch tail / training, the 20-NFE sampler program) is executed on CPU tensors against a stand-in for
libsdb200 that answers the host-only queries with the real library, checks the ctypes marshalling of every launch
(argument count and types against _lib.SIGNATURES) and enqueues nothing.  Values are garbage by construction; what is
checked is that the Python side wires shapes, views, workspaces and argument lists consistently -- the class of error
that would otherwise cost GPU minutes to find.  Numerics live in the -m gpu tests."""
import collections
import ctypes

import pytest
import torch

from oracle import dpm_ref, unet_ref
from oracle import slot_attention_ref as sa_ref

QUERIES = {'sdb_version', 'sdb_last_error', 'sdb_launch_count', 'sdb_attention_tc_supported', 'sdb_attention_fewkeys_supported', 'sdb_slot_attend_workspace',
           'sdb_slot_attend_fused_supported', 'sdb_slot_attend_fused_workspace', 'sdb_slot_attend_fused_chunks',
           'sdb_slot_attend_fused_ascale', 'sdb_slot_update_supported', 'sdb_token_attention_supported'}


class DryLib:
    def __init__(self, real):
        self.real = real
        self.calls = collections.Counter()

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if name in QUERIES:
            return fn

        def launch(*args):
            assert len(args) == len(fn.argtypes), (name, len(args), len(fn.argtypes))
            for i, (a, t) in enumerate(zip(args, fn.argtypes)):
                try:
                    t.from_param(a)
                except (TypeError, ctypes.ArgumentError) as e:
                    raise AssertionError(f'{name}: argument {i} ({a!r}) does not marshal as {t.__name__}') from e
            self.calls[name] += 1
            return 0
        return launch


@pytest.fixture
def dry(monkeypatch):
    from slotdiffusion_b200 import _lib, ops
    lib = DryLib(_lib.lib())
    monkeypatch.setattr(ops, 'lib', lambda: lib)
    monkeypatch.setattr(ops, '_stream', lambda: None)

    def f32(t, name='tensor'):
        if t.dtype != torch.float32:
            raise RuntimeError(f'{name}: expected float32, got {t.dtype}')
        return t
    monkeypatch.setattr(ops, '_f32', f32)
    return lib


SMALL = dict(in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1, attention_resolutions=(2, 1),
             channel_mult=(1, 2), num_head_channels=32, context_dim=64)


def small_unet(dropout=0.0):
    from slotdiffusion_b200.unet import UNetModel
    net = UNetModel(dropout=dropout, **SMALL)
    for p in net.parameters():          # zero-initialised tensors of the reference init -> something non-trivial
        if p.abs().sum() == 0:
            torch.nn.init.normal_(p, std=0.02)
    return net


def test_unet_forward_schedule(dry):
    net = small_unet().eval()
    x, ctx = torch.randn(3, 3, 16, 16), torch.randn(3, 5, 64)
    with torch.no_grad():
        for t in (torch.tensor([7, 503, 999]), torch.tensor([0.0, 333.25, 998.999])):
            y = net._exec(x, t, ctx)
            assert y.shape == x.shape and y.dtype == torch.float32
    assert dry.calls['sdb_gemm'] > 20 and dry.calls['sdb_conv3_in'] == 2 and dry.calls['sdb_conv3_out'] == 2


def test_unet_training_step_schedule(dry):
    """forward with a tape + backward replay (backward.py): every parameter and the context receive a gradient"""
    net = small_unet(dropout=0.1).train()
    x, ctx = torch.randn(2, 3, 16, 16), torch.randn(2, 5, 64, requires_grad=True)
    out = net._exec(x, torch.tensor([3, 700]), ctx)
    assert out.shape == x.shape and out.requires_grad
    fwd_gemms = dry.calls['sdb_gemm']
    out.sum().backward()
    assert ctx.grad is not None and ctx.grad.shape == ctx.shape
    missing = [n for n, p in net.named_parameters() if p.grad is None]
    assert not missing, missing[:5]
    assert all(p.grad.shape == p.shape for p in net.parameters())
    assert dry.calls['sdb_gemm'] > 2.5 * fwd_gemms            # dgrad + wgrad per forward contraction
    assert dry.calls['sdb_grad_pack'] > 0 and dry.calls['sdb_groupnorm_bwd'] > 0 and dry.calls['sdb_attention_bwd'] > 0


def test_dry_library_rejects_bad_marshalling(dry):
    with pytest.raises(AssertionError):
        dry.sdb_gru_gates(None, None)                          # wrong argument count
    with pytest.raises(AssertionError):
        dry.sdb_lincomb(None, None, None, None, 'x', 1.0, 1.0, 4, None)      # a str where a float is declared


@pytest.mark.parametrize('tail', [False, True])
def test_slot_attention_inference_schedules(dry, tail, monkeypatch):
    from slotdiffusion_b200 import autograd
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    monkeypatch.setattr(autograd, 'FUSED_TAIL', tail)
    B, N, S, D = 3, 200, 7, 192
    mod = SlotAttentionWMask(D, 3, S, D, 2 * D)
    mod.load_state_dict(sa_ref.random_params(D, D, 2 * D, seed=1))
    with torch.no_grad():
        slots, mask = autograd.slot_attention_apply(mod, torch.randn(B, N, D), torch.randn(B, S, D), True)
    assert slots.shape == (B, S, D) and mask.shape == (B, S, N)
    if tail:
        assert dry.calls['sdb_slot_update'] == 4 and dry.calls['sdb_slot_attend_fused_partials'] == 3
        assert dry.calls['sdb_gemm'] == 0
    else:
        assert dry.calls['sdb_slot_attend_fused'] == 3 and dry.calls['sdb_gemm'] == 15


def test_slot_attention_training_schedule(dry):
    from slotdiffusion_b200 import autograd
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, S, D = 2, 96, 5, 192
    mod = SlotAttentionWMask(D, 2, S, D, 2 * D)
    x = torch.randn(B, N, D, requires_grad=True)
    s0 = torch.randn(B, S, D, requires_grad=True)
    slots, mask = autograd.slot_attention_apply(mod, x, s0, True)
    assert slots.requires_grad and not mask.requires_grad            # sa_diffusion.py:50: the mask is detached
    slots.sum().backward()
    assert x.grad.shape == x.shape and s0.grad.shape == s0.shape
    assert all(p.grad is not None and p.grad.shape == p.shape for p in mod.parameters())


def test_sampler_program_schedule(dry):
    from slotdiffusion_b200.dpm_solver import DPMSolverSampler
    net = small_unet().eval()
    betas = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
    smp = DPMSolverSampler(net, betas, codebook=torch.randn(64, 3), steps=20, use_cuda_graph=False)
    with torch.no_grad():
        y = smp.sample(torch.randn(2, 3, 16, 16), torch.randn(2, 5, 64))
    assert y.shape == (2, 3, 16, 16)
    assert smp.nfe == 20 and dry.calls['sdb_dpm_x0'] == 20 and dry.calls['sdb_conv3_out'] == 20


@pytest.mark.parametrize('name,hw,S,dc', [('clevrtex', 32, 11, 192), ('movie_24slots', 32, 24, 192), ('coco_dino', 56, 7, 256)])
def test_shipped_unet_geometries_forward_and_backward(dry, name, hw, S, dc):
    """the shipped 134 M-parameter UNet (sa_ldm_clevrtex_params-res128.py:79-95) at the latent sizes / slot counts of the
    BASELINE configs: forward and one training step wire up without a GPU"""
    from slotdiffusion_b200.unet import UNetModel
    cfg = dict(unet_ref.DEFAULT_CFG, context_dim=dc)
    net = UNetModel(dropout=0.1, **cfg).train()
    x, ctx = torch.randn(1, 3, hw, hw), torch.randn(1, S, dc, requires_grad=True)
    out = net._exec(x, torch.tensor([500]), ctx)
    assert out.shape == x.shape
    out.sum().backward()
    assert all(p.grad is not None for p in net.parameters()) and ctx.grad.shape == ctx.shape
    with torch.no_grad():
        assert net.eval()._exec(x, torch.tensor([12.5]), ctx.detach()).shape == x.shape


def test_full_model_training_chain_wires_up(dry):
    """tools/full_model_train_bench.py in miniature: eager-PyTorch encoder -> Slot Attention (tape) -> UNet (tape) -> loss;
    gradients reach the encoder through both hand-written backward passes"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    from full_model_parts import ImageEncoder, VQVAEEncoder
    from slotdiffusion_b200 import autograd
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, S, D = 2, 4, 64
    enc = ImageEncoder((32, 32), D).train()
    vae = VQVAEEncoder().eval().requires_grad_(False)
    sa = SlotAttentionWMask(D, 2, S, D, 2 * D).train()
    net = small_unet(dropout=0.1).train()
    init = torch.nn.Parameter(torch.randn(1, S, D))
    img = torch.randn(B, 3, 32, 32).clamp(-1, 1)
    with torch.no_grad():
        x0 = vae(img)
    assert x0.shape == (B, 3, 8, 8)
    feats = enc(img)
    assert feats.shape == (B, 64, D)
    slots, mask = autograd.slot_attention_apply(sa, feats, init.expand(B, -1, -1), True)
    eps = torch.randn_like(x0)
    loss = torch.nn.functional.mse_loss(net._exec(0.7 * x0 + 0.3 * eps, torch.tensor([10, 900]), slots), eps)
    loss.backward()
    for name, mod in (('encoder', enc), ('slot attention', sa), ('unet', net)):
        assert all(p.grad is not None for p in mod.parameters()), name
    assert init.grad is not None and init.grad.shape == init.shape


def test_video_recurrence_chain_wires_up(dry):
    """SAViDiffusion.encode (savi_diffusion.py:183-196) in miniature: the same Slot-Attention module applied once per frame,
    frame t+1 initialised by a TransformerPredictor of frame t's slots; backward runs through all T applications (a fresh
    flat gradient buffer per application, accumulated by autograd)"""
    from slotdiffusion_b200 import autograd
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, T, N, S, D = 2, 3, 64, 5, 64
    sa = SlotAttentionWMask(D, 2, S, D, 2 * D).train()
    layer = torch.nn.TransformerEncoderLayer(d_model=D, nhead=4, dim_feedforward=4 * D, norm_first=True, batch_first=True)
    predictor = torch.nn.TransformerEncoder(layer, num_layers=2).train()
    init = torch.nn.Parameter(torch.randn(1, S, D))
    feats = torch.randn(B, T, N, D, requires_grad=True)
    prev, frames = None, []
    for f in range(T):
        lat = init.expand(B, -1, -1) if prev is None else predictor(prev)
        prev, mask = autograd.slot_attention_apply(sa, feats[:, f].contiguous(), lat, True)
        assert mask.shape == (B, S, N) and not mask.requires_grad
        frames.append(prev)
    slots = torch.stack(frames, 1)
    assert slots.shape == (B, T, S, D)
    slots.sum().backward()
    assert feats.grad.shape == feats.shape and init.grad is not None
    assert all(p.grad is not None for p in sa.parameters()) and all(p.grad is not None for p in predictor.parameters())
    assert dry.calls['sdb_slot_attend_bwd'] == T * 2            # iterations x frames


@pytest.mark.parametrize('norm_first,train', [(True, False), (True, True), (False, True)])
def test_predictor_schedules(dry, monkeypatch, norm_first, train):
    """TransformerPredictor (predictor.py:20-44): inference, training (tape + backward) with and without dropout, pre-/post-LN;
    T-1 calls of the same module before one backward (savi_diffusion.py:183-196) each own their gradient buffer."""
    from slotdiffusion_b200 import predictor as pr
    monkeypatch.setattr(pr, '_check', lambda mod, x: None)        # the CUDA-only guard; geometry support is a real query
    B, S, D = 3, 11, 192
    net = pr.TransformerPredictor(D, 2, 4, 4 * D, norm_first).train(train)
    assert net._wcache is not None and pr.ops.token_attention_supported(S, D // 4)
    with torch.no_grad():
        y = net(torch.randn(B, S, D))
    assert y.shape == (B, S, D)
    n_attn = dry.calls['sdb_token_attention']
    assert n_attn == 2 and dry.calls['sdb_gemm'] == 8
    assert dry.calls['sdb_dropout_add'] == (6 if train else 0)
    x = torch.randn(B, S, D, requires_grad=True)
    a = net(x)
    b = net(a)                                                      # second application before the backward
    assert a.requires_grad and b.shape == (B, S, D)
    (a.sum() + b.sum()).backward()
    assert x.grad is not None and x.grad.shape == x.shape
    missing = [n for n, q in net.named_parameters() if q.grad is None]
    assert not missing, missing
    assert all(q.grad.shape == q.shape for q in net.parameters())
    assert dry.calls['sdb_token_attention_bwd'] == 4
    assert dry.calls['sdb_dropout_add'] == (6 + 2 * 12 if train else 0)
