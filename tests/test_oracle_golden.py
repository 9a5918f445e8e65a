"""Pin the oracle (oracle/) against outputs of the UNMODIFIED reference modules
(tests/golden/*.npz, made by tools/make_golden.py in the build container)."""
import numpy as np
import pytest
import torch

from helpers import SA_CASES, argmax_mismatch, checksum, golden, rel_l2, sa_case, seeded
from oracle import dpm_ref, unet_ref
from oracle import slot_attention_ref as sa_ref


@pytest.mark.parametrize('name', list(SA_CASES))
def test_slot_attention_oracle_matches_reference(name):
    g = golden(name)
    p, x, s0, gw, iters = sa_case(name)
    np.testing.assert_allclose(checksum(x), g['x_sum'], rtol=1e-12)      # same seeded inputs as the fixture
    np.testing.assert_allclose(checksum(s0), g['s0_sum'], rtol=1e-12)
    slots, mask = sa_ref.slot_attention_forward(p, x, s0, iters)
    assert rel_l2(slots, g['slots']) < 2e-6
    assert rel_l2(mask, g['mask']) < 2e-6
    s64, m64 = sa_ref.slot_attention_forward(p, x.double(), s0.double(), iters)
    assert rel_l2(s64, g['slots64']) < 1e-12
    assert (m64.argmax(1).numpy() == g['argmax64']).all()
    # fp32 oracle argmax vs fp64 reference argmax: only near-ties may differ
    real, near = argmax_mismatch(mask, g['argmax64'], g['margin64'], 1e-5)
    assert real == 0


@pytest.mark.parametrize('name', ['sa_ragged_small', 'sa_coco_vitb16'])
def test_slot_attention_oracle_gradients(name):
    g = golden(name)
    p, x, s0, gw, iters = sa_case(name)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    x = x.requires_grad_(True)
    s0 = s0.requires_grad_(True)
    slots, _ = sa_ref.slot_attention_forward(p, x, s0, iters)
    (slots * gw).sum().backward()
    assert rel_l2(s0.grad, g['grad_slots']) < 1e-5
    assert rel_l2(x.grad[:, :8], g['grad_inputs_head']) < 1e-5
    np.testing.assert_allclose(checksum(x.grad)[1], g['grad_inputs_sum'][1], rtol=1e-4)
    for k, v in p.items():
        ref = g['grad.' + k]
        if v.dim() == 1:
            if np.linalg.norm(ref) < 1e-6:      # mathematically-zero gradient (softmax shift invariance)
                assert v.grad.norm().item() < 1e-6, k
            else:
                assert rel_l2(v.grad, ref) < 2e-5, k
        else:
            np.testing.assert_allclose(checksum(v.grad)[1], ref[1], rtol=1e-4, err_msg=k)


def test_unet_oracle_matches_reference_small():
    g = golden('unet_small')
    cfg = dict(unet_ref.DEFAULT_CFG, model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,),
               num_res_blocks=1, context_dim=64)
    sd = unet_ref.random_state_dict(cfg, seed=31)
    x = seeded((3, 3, 16, 16), 41)
    ctx = seeded((3, 5, 64), 42)
    np.testing.assert_allclose(checksum(x), g['x_sum'], rtol=1e-12)
    y = unet_ref.unet_forward(sd, x, torch.tensor([7, 503, 999]), ctx, cfg)
    assert rel_l2(y, g['y_int']) < 1e-5
    y = unet_ref.unet_forward(sd, x, torch.tensor([0.0, 333.25, 998.999]), ctx, cfg)
    assert rel_l2(y, g['y_flt']) < 1e-5


def test_unet_oracle_matches_reference_full_and_grad():
    g = golden('unet_clevrtex')
    sd = unet_ref.random_state_dict(seed=31)
    x = seeded((2, 3, 32, 32), 41)
    ctx = seeded((2, 11, 192), 42).requires_grad_(True)
    np.testing.assert_allclose(checksum(ctx), g['ctx_sum'], rtol=1e-12)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y = unet_ref.unet_forward(sdg, x, torch.tensor([7, 503]), ctx)
    assert rel_l2(y, g['y_int']) < 1e-5
    (y * seeded(tuple(y.shape), 43)).sum().backward()
    assert rel_l2(ctx.grad, g['grad_ctx']) < 1e-4
    for k in g.files:
        if k.startswith('gsum.'):
            np.testing.assert_allclose(checksum(sdg[k[5:]].grad)[1], g[k][1], rtol=1e-3, err_msg=k)
    with torch.no_grad():
        y = unet_ref.unet_forward(sd, x, torch.tensor([0.0, 333.25]), ctx)
    assert rel_l2(y, g['y_flt']) < 1e-5


def test_noise_schedule_and_qsample():
    g = golden('dpm')
    bufs = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())
    np.testing.assert_allclose(checksum(bufs['betas']), g['betas_sum'], rtol=1e-12)
    ns = dpm_ref.NoiseScheduleVP(bufs['betas'])
    t = torch.as_tensor(g['t'])
    np.testing.assert_array_equal(ns.log_mean_coeff(t).numpy(), g['log_alpha'])
    np.testing.assert_array_equal(ns.lam(t).numpy(), g['lam'])
    np.testing.assert_array_equal(ns.inverse_lambda(torch.as_tensor(g['lam'])).numpy(), g['inv'])
    xt = dpm_ref.q_sample(bufs, seeded((4, 3, 32, 32), 54), torch.tensor([0, 10, 500, 999]), seeded((4, 3, 32, 32), 55))
    np.testing.assert_array_equal(xt.numpy(), g['q_sample'])


def test_dpm_sampler_oracle_matches_reference():
    g = golden('dpm')
    sd = unet_ref.random_state_dict(seed=31)
    betas = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
    cb = seeded((4096, 3), 51)
    ctx = seeded((1, 11, 192), 52)
    xT = seeded((1, 3, 32, 32), 53)
    calls = []

    def fn(x, t, c):
        calls.append(t[0].item())
        return unet_ref.unet_forward(sd, x, t, c)
    with torch.no_grad():
        y = dpm_ref.dpm_sample(fn, betas, xT, ctx, None)
        assert len(calls) == 20
        np.testing.assert_allclose(np.array(calls), g['t_model'], rtol=0, atol=0)
        assert rel_l2(y, g['sample_novq']) < 1e-5
        y = dpm_ref.dpm_sample(fn, betas, xT, ctx, cb)
        assert rel_l2(y, g['sample_vq']) < 1e-5


def test_vqvae_oracle_matches_reference_golden():
    """oracle/vqvae_ref.py pinned to the unmodified reference Encoder / Decoder (tests/golden/vqvae.npz); small config on
    every run, the full 128x128 config too (a few seconds of CPU)."""
    import numpy as np
    from helpers import golden, seeded
    from oracle import vqvae_ref
    g = golden('vqvae')
    for tag, over, B in (('small', dict(resolution=32, ch_mult=(1, 2)), 3), ('full', {}, 2)):
        cfg = dict(vqvae_ref.DEFAULT_CFG, **over)
        esd, dsd = vqvae_ref.random_state_dicts(cfg, seed=61)
        R = cfg['resolution']
        r = R // 2 ** (len(cfg['ch_mult']) - 1)
        x = seeded((B, 3, R, R), 62).clamp(-1, 1)
        z = seeded((B, 3, r, r), 63)
        with torch.no_grad():
            np.testing.assert_allclose(vqvae_ref.encoder_forward(esd, x, cfg).numpy(), g[tag + '_enc'], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(vqvae_ref.decoder_forward(dsd, z, cfg).numpy(), g[tag + '_dec'], rtol=1e-4, atol=1e-5)


def test_resnet_oracle_matches_reference_golden():
    """oracle/resnet_ref.py pinned to the unmodified reference resnet18(small_inputs=True, use_layer4=False)"""
    from helpers import golden, seeded
    from oracle import resnet_ref
    g = golden('resnet')
    sd = resnet_ref.random_state_dict('resnet18', False, seed=71)
    x = seeded((2, 3, 64, 64), 73).clamp(-1, 1)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y = resnet_ref.resnet_forward(sdg, x)
    assert rel_l2(y, g['y64']) < 1e-6
    (y * seeded(tuple(y.shape), 74)).sum().backward()
    for k in ('bn1.weight', 'layer2.0.downsample.1.bias', 'layer3.1.bn2.weight'):
        assert rel_l2(sdg[k].grad, g['grad.' + k]) < 1e-5, k
    assert abs(checksum(sdg['layer1.0.conv1.weight'].grad)[1] - g['grad.layer1.0.conv1.weight'][1]) \
        / g['grad.layer1.0.conv1.weight'][1] < 1e-5


PRED_CASES = {
    'movid': (3, 15, 192, 2, 4, 768, True),
    'clevrer': (2, 7, 128, 2, 4, 512, True),
    'postln': (2, 11, 256, 1, 4, 512, False),
}


@pytest.mark.parametrize('name', sorted(PRED_CASES))
def test_predictor_oracle_matches_reference_golden(name):
    """oracle/predictor_ref.py pinned to the unmodified reference TransformerPredictor (predictor.py:20-44): outputs, input
    gradient and parameter gradients of a fixed linear functional."""
    from oracle import predictor_ref
    g = golden('predictor')
    B, S, D, L, Hh, F, nf = PRED_CASES[name]
    sd = predictor_ref.random_state_dict(D, L, F, seed=900 + len(name))
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = seeded((B, S, D), 91).requires_grad_(True)
    y = predictor_ref.predictor_forward(sdg, x, L, Hh, nf)
    (y * seeded((B, S, D), 92)).sum().backward()
    assert rel_l2(y, g[name + '.y']) < 2e-6
    assert rel_l2(x.grad, g[name + '.dx']) < 1e-5
    for k, v in sdg.items():
        ref = g[name + '.grad.' + k]
        if v.grad.dim() == 1:
            assert rel_l2(v.grad, ref) < 2e-5, k
        else:
            got = checksum(v.grad)
            assert abs(got[0] - ref[0]) <= 2e-4 * max(1.0, np.sqrt(ref[1])) and abs(got[1] - ref[1]) <= 1e-4 * ref[1], k
