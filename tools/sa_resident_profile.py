"""ncu target: a few launches of the persistent Slot-Attention kernel (whole module forward in one launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slotdiffusion_b200 import autograd
from slotdiffusion_b200.slot_attention import SlotAttentionWMask
B, N, S, D = int(os.environ.get('SA_B', 33)), 1024, int(os.environ.get('SA_S', 11)), 192
torch.manual_seed(0)
mod = SlotAttentionWMask(D, 3, S, D, 2 * D).cuda().eval()
x = torch.randn(B, N, D, device='cuda')
s0 = torch.randn(B, S, D, device='cuda')
autograd.RESIDENT, autograd.RESIDENT_WAVES = True, 1 << 20
with torch.no_grad():
    for _ in range(3):
        mod(x, s0)
torch.cuda.synchronize()
