import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import rel_l2, seeded
from oracle import resnet_ref
from slotdiffusion_b200 import resnet
R = int(os.environ.get('R', 64)); Bn = int(os.environ.get('B', 2)); SEED = int(os.environ.get('SEED', 5))
sd = resnet_ref.random_state_dict('resnet18', False, seed=SEED)
net = resnet.resnet18(small_inputs=True, use_layer4=False).cuda(); net.load_state_dict(sd)
x = seeded((Bn, 3, R, R), 81).clamp(-1, 1)
y = net(x.cuda()); gw = seeded(tuple(y.shape), 82)
(y * gw.cuda()).sum().backward()
sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
ref = resnet_ref.resnet_forward(sd64, x.double()); (ref * gw.double()).sum().backward()
print('fwd', rel_l2(y, ref))
errs = [(k, rel_l2(p.grad, sd64[k].grad)) for k, p in net.named_parameters()]
print('R', R, 'B', Bn, 'seed', SEED, 'worst', max(errs, key=lambda t: t[1]), 'bad', [(k, '%.1e' % e) for k, e in errs if e > 2e-4][:8])
