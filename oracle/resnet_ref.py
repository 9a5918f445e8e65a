"""Oracle: ResNet18/34-GN image encoder (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional restatement (dict of tensors keyed like the reference state_dict, plain torch ops, any dtype -- fp64 for
gradient checks) of
  ResNet._forward_impl   /root/reference/slotdiffusion/video_based/models/resnet.py:286-306
  BasicBlock.forward     resnet.py:72-90           downsample = conv1x1(stride) + GroupNorm, resnet.py:264-268
for the shipped configuration (GroupNorm(32), small_inputs=True: 3x3 stride-1 stem, identity max-pool).
Pinned to the unmodified reference by tests/golden/resnet.npz (tools/make_golden.py resnet).
"""
import torch
import torch.nn.functional as F

LAYERS = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3)}


def _gn(x, sd, name):
    return F.group_norm(x, 32, sd[name + '.weight'], sd[name + '.bias'], 1e-5)       # _gn, resnet.py:8-9


def _relu(z, rec):
    """ReLU; `rec` (optional) = dict(masks=iterator | None, seen=list): with `masks` the activation pattern is FORCED
    (z * mask) -- gradient checks against an fp32 implementation compare like with like: an element within round-off of
    zero may legitimately fall on either side, and a single flipped element moves every upstream gradient by O(1e-3) --
    and `seen` collects this run's own patterns so the caller can count the disagreements."""
    if rec is None:
        return F.relu(z)
    rec['seen'].append(z.detach() > 0)
    if rec.get('masks') is None:
        return F.relu(z)
    return z * next(rec['masks']).to(z.dtype)


def basic_block(sd, pre, x, stride, rec=None):
    out = _relu(_gn(F.conv2d(x, sd[pre + 'conv1.weight'], None, stride=stride, padding=1), sd, pre + 'bn1'), rec)
    out = _gn(F.conv2d(out, sd[pre + 'conv2.weight'], None, stride=1, padding=1), sd, pre + 'bn2')
    if pre + 'downsample.0.weight' in sd:
        x = _gn(F.conv2d(x, sd[pre + 'downsample.0.weight'], None, stride=stride), sd, pre + 'downsample.1')
    return _relu(out + x, rec)


def resnet_forward(sd, x, arch='resnet18', use_layer4=False, rec=None):
    sd = {k: v.to(x.dtype) for k, v in sd.items()}
    h = _relu(_gn(F.conv2d(x, sd['conv1.weight'], None, stride=1, padding=1), sd, 'bn1'), rec)
    for li, n in enumerate(LAYERS[arch][:4 if use_layer4 else 3]):
        for bi in range(n):
            h = basic_block(sd, f'layer{li + 1}.{bi}.', h, 2 if (li > 0 and bi == 0) else 1, rec)
    return h


def random_state_dict(arch='resnet18', use_layer4=False, seed=0):
    """kaiming-like convolution weights, non-trivial GroupNorm affines; keys / shapes of the reference module."""
    from slotdiffusion_b200 import resnet
    g = torch.Generator().manual_seed(seed)
    net = getattr(resnet, arch)(small_inputs=True, use_layer4=use_layer4)
    sd = {}
    for k, v in net.state_dict().items():
        if v.dim() == 1:
            sd[k] = (1 + 0.2 * torch.randn(v.shape, generator=g)) if k.endswith('.weight') else 0.2 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = torch.randn(v.shape, generator=g) * (2.0 / (v.shape[0] * v[0, 0].numel())) ** 0.5
    return sd
