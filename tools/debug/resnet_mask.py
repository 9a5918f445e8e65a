import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import rel_l2, seeded
from oracle import resnet_ref
from slotdiffusion_b200 import resnet
sd = resnet_ref.random_state_dict('resnet18', False, seed=71)
net = resnet.resnet18(small_inputs=True, use_layer4=False).cuda(); net.load_state_dict(sd)
x = seeded((2, 3, 64, 64), 81).clamp(-1, 1)
y = net(x.cuda()); gw = seeded(tuple(y.shape), 82)
(y * gw.cuda()).sum().backward()
sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
ref = resnet_ref.resnet_forward(sd64, x.double()); (ref * gw.double()).sum().backward()
yc = y.detach().cpu().double()
print('fwd rel', rel_l2(y, ref), 'max abs', (yc - ref.detach()).abs().max().item(), 'ref absmax', ref.abs().max().item())
mm = ((yc > 0) != (ref.detach() > 0))
print('mask mismatches', int(mm.sum()), 'of', mm.numel())
# last-layer bias gradient three ways
k = 'layer3.1.bn2.bias'
ours = dict(net.named_parameters())[k].grad.cpu().double()
manual = (gw.double() * (yc > 0)).sum((0, 2, 3))
print('ours vs manual(our mask)', rel_l2(ours, manual), ' ours vs oracle', rel_l2(ours, sd64[k].grad), ' manual vs oracle', rel_l2(manual, sd64[k].grad))
d = (ours - sd64[k].grad).abs(); print('top channel errors', d.topk(5).values.tolist(), d.topk(5).indices.tolist(), 'grad scale', sd64[k].grad.abs().mean().item())
