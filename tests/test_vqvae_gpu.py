"""GPU parity of the frozen VQ-VAE encoder / decoder (slotdiffusion_b200/vqvae.py) against the outputs of
the UNMODIFIED reference modules (tests/golden/vqvae.npz, tools/make_golden.py vqvae) and the CPU oracle."""
import pytest
import torch

from helpers import golden, rel_l2, seeded
from oracle import vqvae_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TIGHT = 5e-5


def _build(over, seed=61):
    from slotdiffusion_b200 import vqvae
    cfg = dict(vqvae_ref.DEFAULT_CFG, **over)
    esd, dsd = vqvae_ref.random_state_dicts(cfg, seed=seed)
    enc, dec = vqvae.Encoder(dropout=0.0, **cfg).cuda().eval(), vqvae.Decoder(dropout=0.0, **cfg).cuda().eval()
    enc.load_state_dict(esd, strict=True)
    dec.load_state_dict(dsd, strict=True)
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.requires_grad_(False)                              # VQVAEWrapper freezes the first stage (VQVAE.py:172-176)
    return enc, dec, esd, dsd, cfg


@pytest.mark.parametrize('tag,over,B', [('small', dict(resolution=32, ch_mult=(1, 2)), 3), ('full', {}, 2)])
def test_encoder_decoder_match_reference_golden(tag, over, B):
    g = golden('vqvae')
    enc, dec, esd, dsd, cfg = _build(over)
    R = cfg['resolution']
    r = R // 2 ** (len(cfg['ch_mult']) - 1)
    x = seeded((B, 3, R, R), 62).clamp(-1, 1).cuda()
    z = seeded((B, 3, r, r), 63).cuda()
    with torch.no_grad():
        ye, yd = enc(x), dec(z)
    assert ye.shape == g[tag + '_enc'].shape and yd.shape == g[tag + '_dec'].shape
    ee, ed = rel_l2(ye, g[tag + '_enc']), rel_l2(yd, g[tag + '_dec'])
    print(tag, 'encoder rel', ee, 'decoder rel', ed)
    assert ee < TIGHT and ed < TIGHT


def test_encoder_batch_independence_and_oracle_at_other_shapes():
    """rectangular input, batch 5, different weights: vs the oracle; per-sample results do not depend on the batch"""
    enc, dec, esd, dsd, cfg = _build(dict(resolution=64, ch_mult=(1, 2, 4)), seed=7)
    x = seeded((5, 3, 64, 96), 71).clamp(-1, 1)
    with torch.no_grad():
        y = enc(x.cuda())
        y1 = enc(x[3:4].cuda())
    ref = vqvae_ref.encoder_forward(esd, x.double(), cfg)
    assert rel_l2(y, ref) < TIGHT
    assert rel_l2(y[3:4], y1) < 1e-5
    z = seeded((2, 3, 16, 24), 72)
    with torch.no_grad():
        yd = dec(z.cuda())
    assert rel_l2(yd, vqvae_ref.decoder_forward(dsd, z.double(), cfg)) < TIGHT


def test_inference_only_and_no_cpu_fallback():
    from slotdiffusion_b200 import vqvae
    cfg = dict(vqvae_ref.DEFAULT_CFG, resolution=32, ch_mult=(1, 2))
    enc = vqvae.Encoder(dropout=0.0, **cfg)
    with pytest.raises(RuntimeError, match='CUDA'):
        enc(torch.zeros(1, 3, 32, 32))
    enc = enc.cuda()
    with pytest.raises(RuntimeError, match='inference'):
        enc(torch.zeros(1, 3, 32, 32, device='cuda'))           # parameters still require grad, grad mode on
