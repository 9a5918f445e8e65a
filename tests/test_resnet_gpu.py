"""GPU parity of the ResNet18-GN encoder (slotdiffusion_b200/resnet.py), forward and backward, against
the outputs / gradients of the UNMODIFIED reference module (tests/golden/resnet.npz, tools/make_golden.py resnet) and
fp64 autograd over the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import checksum, golden, rel_l2, seeded
from oracle import resnet_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TIGHT, GTOL = 5e-5, 2e-4


def _net(arch='resnet18', use_layer4=False, seed=71):
    from slotdiffusion_b200 import resnet
    sd = resnet_ref.random_state_dict(arch, use_layer4, seed=seed)
    net = getattr(resnet, arch)(small_inputs=True, use_layer4=use_layer4).cuda()
    net.load_state_dict(sd, strict=True)
    return net, sd


def test_forward_matches_reference_golden_128():
    g = golden('resnet')
    net, sd = _net()
    x = seeded((1, 3, 128, 128), 72).clamp(-1, 1).cuda()
    with torch.no_grad():
        y = net(x)
    assert y.shape == (1, 256, 32, 32)
    assert rel_l2(y, g['y128']) < TIGHT
    # the training-mode forward (stem as a zero-padded implicit GEMM) gives the same features
    y2 = net(x)
    assert y2.requires_grad and rel_l2(y2, g['y128']) < TIGHT


def test_gradients_match_reference_golden_64():
    g = golden('resnet')
    net, sd = _net()
    x = seeded((2, 3, 64, 64), 73).clamp(-1, 1).cuda()
    y = net(x)
    assert rel_l2(y, g['y64']) < TIGHT
    (y * seeded(tuple(y.shape), 74).cuda()).sum().backward()
    worst = ('', 0.0)
    for k, p in net.named_parameters():
        ref = g['grad.' + k]
        if p.grad.dim() == 1:
            e = rel_l2(p.grad, ref)
        else:
            e = abs(checksum(p.grad)[1] - ref[1]) / ref[1]           # sum of squares of the gradient tensor
        worst = max(worst, (k, e), key=lambda t: t[1])
        assert e < 1e-3, (k, e)
    print('worst parameter gradient vs reference golden', worst)


def test_gradients_match_fp64_oracle_full_tensors():
    """every parameter-gradient tensor element-wise against fp64 autograd of the oracle (128x128 stem / layer1 included:
    the 128-wide feature maps take the 64-pixel-segment form of the wgrad GEMM)"""
    net, sd = _net(seed=5)
    x = seeded((2, 3, 128, 128), 81).clamp(-1, 1)
    y = net(x.cuda())
    gw = seeded(tuple(y.shape), 82)
    (y * gw.cuda()).sum().backward()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    ref = resnet_ref.resnet_forward(sd64, x.double())
    (ref * gw.double()).sum().backward()
    assert rel_l2(y, ref) < TIGHT
    for k, p in net.named_parameters():
        assert rel_l2(p.grad, sd64[k].grad) < GTOL, (k, rel_l2(p.grad, sd64[k].grad))


def test_resnet34_layer4_and_two_calls_one_backward():
    """the other factory / use_layer4=True, and two forward calls before one backward (per-call gradient buffers)"""
    net, sd = _net('resnet34', True, seed=9)
    xa, xb = seeded((1, 3, 64, 64), 91).clamp(-1, 1), seeded((1, 3, 64, 64), 92).clamp(-1, 1)
    ya, yb = net(xa.cuda()), net(xb.cuda())
    assert ya.shape == (1, 512, 8, 8)
    ga, gb = seeded(tuple(ya.shape), 93), seeded(tuple(ya.shape), 94)
    ((ya * ga.cuda()).sum() + (yb * gb.cuda()).sum()).backward()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    ra = resnet_ref.resnet_forward(sd64, xa.double(), 'resnet34', True)
    rb = resnet_ref.resnet_forward(sd64, xb.double(), 'resnet34', True)
    ((ra * ga.double()).sum() + (rb * gb.double()).sum()).backward()
    assert rel_l2(ya, ra) < TIGHT and rel_l2(yb, rb) < TIGHT
    for k, p in net.named_parameters():
        assert rel_l2(p.grad, sd64[k].grad) < GTOL, k


def test_no_cpu_fallback_and_unsupported_configs():
    from slotdiffusion_b200 import resnet
    net = resnet.resnet18(small_inputs=True, use_layer4=False)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.zeros(1, 3, 32, 32))
    with pytest.raises(NotImplementedError):
        resnet.resnet18(small_inputs=False)
