"""ncu target: a few launches of the fused Slot-Attention attend kernel at B=64 (and the old kernel for comparison)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slotdiffusion_b200 import ops
B, N, S, D = int(os.environ.get('SA_B', 64)), 1024, 11, 192
x = torch.randn(B, N, D, device='cuda')
qa = torch.randn(B * S, D + 4, device='cuda') * D ** -0.5
for _ in range(3):
    ops.slot_attend_fused(x, qa, B, N, S, D, 1e-5, 1e-6, True)
torch.cuda.synchronize()
