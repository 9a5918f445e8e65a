"""Drop-in boundary: the B200 modules keep the reference's constructor arguments and state_dict layout,
and slotdiffusion_b200.dropin.install() puts them into the UNMODIFIED reference without editing it.

Two layers: (1) against the committed fixture tests/golden/state_dict_layout.json (ordered key -> shape of the
hot-path sub-modules inside build_model() of the shipped configs, written by tools/make_golden.py layout from the
reference) -- runs anywhere; (2) against the reference itself when /root/reference is present (build container)."""
import json
import os
import runpy
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYOUT = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_layout.json')))
REF = os.environ.get('SDB_REFERENCE_ROOT', '/root/reference')
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'slotdiffusion')), reason='reference not present')


def _layout(mod):
    return [[k, list(v.shape)] for k, v in mod.state_dict().items()]


@pytest.mark.parametrize('cfg', sorted(LAYOUT))
def test_modules_reproduce_reference_state_dict_layout(cfg):
    """same keys, same order, same shapes as the reference modules built from the shipped config dicts"""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from slotdiffusion_b200.unet import UNetModel
    L = LAYOUT[cfg]
    sa = SlotAttentionWMask(eps=1e-6, **L['slot_attention_ctor'])
    assert _layout(sa) == L['slot_attention']
    ud = {k: (tuple(v) if isinstance(v, list) else v) for k, v in L['unet_dict'].items()}
    net = UNetModel(**ud)            # the reference's unet_dict, unchanged (ddpm.py:342)
    assert _layout(net) == L['unet']
    assert [n for n, _ in net.named_parameters()] == [k for k, _ in L['unet']]      # no buffers, parameter order kept


def test_unsupported_unet_options_are_rejected_loudly():
    from slotdiffusion_b200.unet import UNetModel
    ud = {k: (tuple(v) if isinstance(v, list) else v) for k, v in LAYOUT['sa_ldm_clevrtex_params-res128']['unet_dict'].items()}
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(ud, resblock_updown=True))
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(ud, dims=3))


def test_no_cpu_fallback():
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    sa = SlotAttentionWMask(192, 3, 11, 192, 384)
    with pytest.raises(RuntimeError, match='CUDA'):
        sa(torch.zeros(1, 16, 192), torch.zeros(1, 11, 192))


# ------------------------------------------------------------------------------------------------ with the reference
def _ref_setup():
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import ref_import
    ref_import.setup()
    return ref_import


def _fresh_params(ref_import, task, rel):
    # build_model pops entries from the class-level dicts of the params object: a fresh class per build
    return runpy.run_path(os.path.join(ref_import.REF_ROOT, 'slotdiffusion', task, 'configs', rel))['SlotAttentionParams']()


CASES = [('img_based', 'sa_ldm/sa_ldm_clevrtex_params-res128.py'),
         ('video_based', 'savi_ldm/savi_ldm_movid_params-res128.py')]


@pytest.fixture
def dropin():
    from slotdiffusion_b200 import dropin as d
    d.uninstall()
    yield d
    d.uninstall()


@needs_ref
@pytest.mark.parametrize('task,rel', CASES)
def test_install_swaps_hot_path_inside_unmodified_reference(dropin, task, rel):
    ri = _ref_setup()
    mods = ri.img_models() if task == 'img_based' else ri.video_models()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref_model = mods.build_model(_fresh_params(ri, task, rel))
        dropin.install()
        new_model = mods.build_model(_fresh_params(ri, task, rel))
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from slotdiffusion_b200.unet import UNetModel
    assert type(new_model.slot_attention) is SlotAttentionWMask
    assert type(new_model.dm_decoder.model.diffusion_model) is UNetModel
    assert type(ref_model.slot_attention) is not SlotAttentionWMask
    # whole-model checkpoint compatibility: ordered keys, shapes, strict load; optimizer grouping substring intact
    a = [(k, tuple(v.shape)) for k, v in ref_model.state_dict().items()]
    b = [(k, tuple(v.shape)) for k, v in new_model.state_dict().items()]
    assert a == b and len(a) > 900
    new_model.load_state_dict(ref_model.state_dict(), strict=True)
    assert [n for n, _ in new_model.named_parameters()] == [n for n, _ in ref_model.named_parameters()]
    assert any('dm_decoder' in n for n, _ in new_model.named_parameters())      # img_based/method.py:251-257
    # the image encoder and the frozen VQ-VAE first stage are the B200 modules too ...
    from slotdiffusion_b200 import resnet as b200_resnet, vqvae as b200_vqvae
    assert type(new_model.encoder) is b200_resnet.ResNet and type(ref_model.encoder) is not b200_resnet.ResNet
    assert type(new_model.dm_decoder.vae.vqvae.encoder) is b200_vqvae.Encoder
    assert type(new_model.dm_decoder.vae.vqvae.decoder) is b200_vqvae.Decoder
    assert not any(p.requires_grad for p in new_model.dm_decoder.vae.parameters())      # frozen (VQVAE.py:172-176)
    if task == 'video_based':      # the slot transition function between frames (savi.py:331-336)
        from slotdiffusion_b200.predictor import TransformerPredictor
        assert type(new_model.predictor) is TransformerPredictor and type(ref_model.predictor) is not TransformerPredictor
    # ... everything else is still the reference's own code
    assert type(new_model.dm_decoder) is type(ref_model.dm_decoder)
    assert type(new_model.dm_decoder.vae) is type(ref_model.dm_decoder.vae)
    assert type(new_model.dm_decoder.vae.vqvae.quantize) is type(ref_model.dm_decoder.vae.vqvae.quantize)
    assert type(new_model.encoder_out_layer) is type(ref_model.encoder_out_layer)
    # a partial install leaves the rest alone
    dropin.uninstall()
    dropin.install(encoder=False, vqvae=False, boundary=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        part = mods.build_model(_fresh_params(ri, task, rel))
    assert type(part.encoder) is type(ref_model.encoder) and type(part.slot_attention) is SlotAttentionWMask
    assert type(part.dm_decoder.vae.vqvae.encoder) is type(ref_model.dm_decoder.vae.vqvae.encoder)
    dropin.uninstall()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        again = mods.build_model(_fresh_params(ri, task, rel))
    assert type(again.slot_attention) is type(ref_model.slot_attention)


class _RecordingSampler:
    made = []

    def __init__(self, unet, betas, codebook=None, steps=20, order=3, use_cuda_graph=True):
        self.unet, self.betas, self.codebook, self.steps, self.order = unet, betas, codebook, steps, order
        self.calls = []
        _RecordingSampler.made.append(self)

    def sample(self, x_T, context):
        self.calls.append((x_T, context))
        return torch.full_like(x_T, 0.5)


@needs_ref
def test_generate_imgs_use_dpm_routes_to_the_b200_sampler(dropin, monkeypatch):
    """cond_ddpm.py:155-189 (the only sampler configuration the repo uses) reaches DPMSolverSampler with the model's
    betas, the VQ codebook and the slots; other requests run the reference loop (with a warning) around the B200 UNet."""
    ri = _ref_setup()
    from slotdiffusion_b200 import dpm_solver as ds
    monkeypatch.setattr(ds, 'DPMSolverSampler', _RecordingSampler)
    _RecordingSampler.made.clear()
    dropin.install()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = ri.img_models().build_model(_fresh_params(ri, *CASES[0]))
    dec = model.dm_decoder
    cond = torch.randn(2, 11, 192)
    out = dec.generate_imgs(cond, batch_size=2, use_dpm=True, verbose=False)
    assert out.shape == (2, 3, 32, 32) and (out == 0.5).all()
    (smp,) = _RecordingSampler.made
    assert smp.unet is dec.model.diffusion_model and smp.steps == 20 and smp.order == 3
    assert torch.equal(torch.as_tensor(smp.betas), dec.betas)
    emb = dec.vae.vqvae.quantize.embedding.weight
    assert smp.codebook.shape == (4096, 3) and torch.allclose(smp.codebook, emb / dec.vae.scale_factor)
    (x_T, ctx), = smp.calls
    assert x_T.shape == (2, 3, 32, 32) and ctx is cond
    # second call reuses the sampler (its CUDA graph) instead of building another
    dec.generate_imgs(cond, batch_size=2, use_dpm=True, verbose=False)
    assert len(_RecordingSampler.made) == 1 and len(smp.calls) == 2
    # a request outside the plan: reference loop + warning, and the B200 UNet refuses CPU tensors (no silent fallback)
    from slotdiffusion.img_based.models.ddpm import cond_ddpm
    ns = cond_ddpm.NoiseScheduleVP(betas=dec.betas)
    fn = cond_ddpm.model_wrapper(model=dec.model, noise_schedule=ns, model_type='noise',
                                 guidance_type='classifier-free', condition=cond)
    solver = cond_ddpm.DPM_Solver(fn, ns, algorithm_type='dpmsolver++', correcting_x0_fn=False, vq_denoised=False)
    with pytest.warns(UserWarning, match='outside the B200 plan'):
        with pytest.raises(RuntimeError, match='CUDA'):
            solver.sample(torch.randn(2, 3, 32, 32), steps=10, order=2, method='multistep')


def test_sampler_graph_cache_is_invalidated_by_parameter_updates():
    """train-then-sample: a captured graph reads the packed weights of the parameter version it was captured with"""
    from slotdiffusion_b200.dpm_solver import DPMSolverSampler
    from slotdiffusion_b200.unet import UNetModel
    from oracle import dpm_ref
    net = UNetModel(3, 32, 3, 1, (2,), channel_mult=(1, 2), num_head_channels=32, context_dim=32)
    smp = DPMSolverSampler(net, dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas'])
    s0 = smp._param_signature()
    assert smp._param_signature() == s0
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    s1 = smp._param_signature()
    assert s1 != s0
    net.load_state_dict(net.state_dict())
    assert smp._param_signature() != s1


def test_modules_survive_deepcopy_and_pickle():
    """EMA copies / torch.save(model): derived state (kernel schedule, packed-weight cache) is rebuilt for the copy"""
    import copy
    import io
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from slotdiffusion_b200.unet import UNetModel
    net = UNetModel(3, 32, 3, 1, (2,), channel_mult=(1, 2), num_head_channels=32, context_dim=32)
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    for c in (copy.deepcopy(net), torch.load(buf, weights_only=False)):
        assert c._exec.net is c and c._exec is not net._exec
        ids = {id(m) for m in c.modules()}
        assert all(k in ids for k in c._exec.emb_off) and all(k in ids for k in c._exec.kv_off)
        assert [(k, v.shape) for k, v in c.state_dict().items()] == [(k, v.shape) for k, v in net.state_dict().items()]
        assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), net.state_dict().values()))
    sa = SlotAttentionWMask(192, 3, 11, 192, 384)
    c = copy.deepcopy(sa)
    assert c._wcache is not sa._wcache and torch.equal(c.project_k.weight, sa.project_k.weight)
