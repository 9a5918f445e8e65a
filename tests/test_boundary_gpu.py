"""GPU parity of the boundary fusions (csrc/boundary.cu) against the eager PyTorch ops the reference
calls: q_sample (ddpm.py:161-165) bit-exact, F.mse_loss (ldm.py:76-77) value + gradient, bilinear mask resize + argmax
(sa_diffusion.py:172-180, test_seg.py:27) with indices exact outside near-ties."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from oracle import dpm_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


def rnd(*shape, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(*shape, generator=g, device='cuda')


@pytest.mark.parametrize('B,shape', [(5, (3, 32, 32)), (64, (3, 32, 32)), (3, (3, 56, 56)), (2, (4, 8, 8))])
def test_q_sample_bit_exact(B, shape):
    from slotdiffusion_b200 import boundary
    buf = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())
    ca, cs = buf['sqrt_alphas_bar'].float().cuda(), buf['sqrt_one_minus_alphas_bar'].float().cuda()
    x0, eps = rnd(B, *shape, seed=1), rnd(B, *shape, seed=2)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(3)).cuda()
    out = boundary.q_sample(x0, t, eps, ca, cs)
    v = (B,) + (1,) * len(shape)
    ref = ca[t].view(v) * x0 + cs[t].view(v) * eps           # extract_to(...) * x0 + extract_to(...) * noise
    assert torch.equal(out, ref)
    ref_cpu = dpm_ref.q_sample(buf, x0.cpu(), t.cpu(), eps.cpu())
    assert rel_l2(out, ref_cpu) < 1e-6


@pytest.mark.parametrize('shape', [(64, 3, 32, 32), (5, 3, 32, 32), (3, 7, 4)])
def test_mse_loss_value_and_gradient(shape):
    from slotdiffusion_b200 import boundary
    p = rnd(*shape, seed=4).requires_grad_(True)
    t = rnd(*shape, seed=5)
    loss = boundary.mse_loss(p, t)
    (loss * 3.0).backward()
    p64 = p.detach().double().requires_grad_(True)
    ref = F.mse_loss(p64, t.double())
    (ref * 3.0).backward()
    assert loss.shape == () and abs(loss.item() - ref.item()) / ref.item() < 1e-6
    assert rel_l2(p.grad, p64.grad) < 1e-6
    with torch.no_grad():
        assert abs(boundary.mse_loss(p.detach(), t).item() - ref.item()) / ref.item() < 1e-6


def test_mse_proxy_falls_back_for_other_signatures():
    from slotdiffusion_b200 import boundary
    Fp = boundary.FunctionalProxy()
    p, t = rnd(8, 12, seed=6), rnd(8, 12, seed=7)
    assert torch.equal(Fp.mse_loss(p, t, reduction='none'), F.mse_loss(p, t, reduction='none'))
    assert Fp.silu is F.silu
    assert abs(Fp.mse_loss(p, t).item() - F.mse_loss(p, t).item()) < 1e-6


@pytest.mark.parametrize('B,S,h,w,H,W', [(4, 11, 32, 32, 128, 128), (2, 24, 32, 32, 128, 128), (3, 7, 14, 14, 224, 224),
                                         (2, 5, 9, 7, 31, 50), (1, 15, 32, 32, 64, 64)])
def test_mask_upsample_and_argmax(B, S, h, w, H, W):
    from slotdiffusion_b200 import boundary
    m = torch.softmax(rnd(B, S, h, w, seed=8) * 3, dim=1)
    up, idx = boundary.mask_upsample(m, (H, W), want_up=True, want_argmax=True)
    # the reference's call shape: [B*S, 1, h, w] (sa_diffusion.py:173-180)
    ref = F.interpolate(m.flatten(0, 1).unsqueeze(1), (H, W), mode='bilinear', align_corners=False).squeeze(1).unflatten(0, (B, S))
    assert (up - ref).abs().max().item() < 6e-7           # <= 2 ulp at 0.25: ATen's kernel may contract the blend into FMAs
    ref64 = F.interpolate(m.double().flatten(0, 1).unsqueeze(1), (H, W), mode='bilinear', align_corners=False) \
        .squeeze(1).unflatten(0, (B, S))
    top = ref64.topk(2, dim=1).values
    near = (top[:, 0] - top[:, 1]) < 1e-6
    bad = idx != ref64.argmax(1)
    assert int((bad & ~near).sum()) == 0, (int(bad.sum()), int(near.sum()))
    assert torch.equal(idx, up.argmax(1))                    # first maximum wins, like torch.argmax
    assert float((idx == ref.argmax(1)).float().mean()) > 0.9999
    # proxy route (dropin): F.interpolate call shape of the reference
    Fp = boundary.FunctionalProxy()
    out = Fp.interpolate(m.flatten(0, 1).unsqueeze(1), (H, W), mode='bilinear', align_corners=False)
    assert out.shape == (B * S, 1, H, W) and torch.equal(out.squeeze(1).unflatten(0, (B, S)), up)
    # anything else is torch's own
    assert torch.equal(Fp.interpolate(m, scale_factor=2.0, mode='nearest'), F.interpolate(m, scale_factor=2.0, mode='nearest'))
