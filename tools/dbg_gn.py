import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from helpers import seeded, rel_l2
from oracle import unet_ref
import test_modules_gpu as T
import slotdiffusion_b200.unet_exec as ue
net, sd, cfg = T.make_unet()
B=64
x, ctx = seeded((B, 3, 32, 32), 71).cuda(), seeded((B, 11, 192), 72).cuda()
t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(73)).cuda()
with torch.no_grad():
    y = net(x, t, context=ctx)
    y2 = net(x[10:13], t[10:13], context=ctx[10:13])
    print('fused vs small-batch(fallback)', rel_l2(y2, y[10:13]))
    ref = unet_ref.unet_forward(sd, x[10:12].cpu(), t[10:12].cpu(), ctx[10:12].cpu())
    print('fused vs oracle', rel_l2(y[10:12], ref), 'fallback vs oracle', rel_l2(y2[:2], ref))
    ue.GS_MIN_ROWS = 10**9
    y3 = net(x, t, context=ctx)
    print('nofuse B64 vs oracle', rel_l2(y3[10:12], ref), 'nofuse vs fused', rel_l2(y3,y))
