"""tools/full_model_parts.py (the reference's eager modules around the hot path, restated for the full-model train-step
tool) against the reference itself: strict state_dict load and equal outputs.  Runs where /root/reference is present."""
import os
import runpy
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('SDB_REFERENCE_ROOT', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'slotdiffusion')), reason='reference not present')


def test_encoder_and_vqvae_encoder_match_the_reference():
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import ref_import
    from full_model_parts import ImageEncoder, VQVAEEncoder
    ref_import.setup()
    cfg = os.path.join(ref_import.REF_ROOT, 'slotdiffusion/img_based/configs/sa_ldm/sa_ldm_clevrtex_params-res128.py')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = ref_import.img_models().build_model(runpy.run_path(cfg)['SlotAttentionParams']()).eval()
    torch.manual_seed(0)
    img = torch.randn(1, 3, 128, 128).clamp(-1, 1)
    enc = ImageEncoder((128, 128), 192).eval()
    enc.load_state_dict({k: v for k, v in model.state_dict().items()
                         if k.split('.')[0] in ('encoder', 'encoder_pos_embedding', 'encoder_out_layer')}, strict=True)
    vae = VQVAEEncoder().eval()
    vq = model.dm_decoder.vae.vqvae
    vae.load_state_dict({k: v for k, v in vq.state_dict().items() if k.startswith(('encoder.', 'quant_conv.'))}, strict=True)
    with torch.no_grad():
        a, b = model._get_encoder_out(img), enc(img)
        assert a.shape == (1, 1024, 192) and ((a - b).norm() / a.norm()).item() < 1e-6
        a, b = model.dm_decoder.vae.encode(img), vae(img)
        assert a.shape == (1, 3, 32, 32) and ((a - b).norm() / a.norm()).item() < 1e-5
