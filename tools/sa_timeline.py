"""Per-role timeline (SM cycles) of CTA (0,0) of the fused Slot-Attention attend kernel (debug buffer)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slotdiffusion_b200 import ops
from slotdiffusion_b200._lib import lib
B, N, S, D = int(os.environ.get('SA_B', 64)), 1024, 11, 192
x = torch.randn(B, N, D, device='cuda')
qa = torch.randn(B * S, D + 4, device='cuda') * D ** -0.5
for _ in range(3):
    ops.slot_attend_fused(x, qa, B, N, S, D, 1e-5, 1e-6, True)
buf = torch.zeros(6 * 64, dtype=torch.int64, device='cuda')
lib().sdb_slot_attend_fused_debug(ctypes.c_void_p(buf.data_ptr()))
ops.slot_attend_fused(x, qa, B, N, S, D, 1e-5, 1e-6, True)
torch.cuda.synchronize()
lib().sdb_slot_attend_fused_debug(None)
t = buf.cpu().view(6, 16, 4)
names = ['conv0 (round0 stored, round1 stored, fence done)', 'g1    (xfull ok, issued)', 'g2    (afull ok, issued)', 'conv  (g0 data ok, w0 done, g15 data ok, w15 done)',
         'soft  (lfull ok, a written)']
ntiles = -(-N // 128) // max(1, min(-(-N // 128), 148 // B))
print('B', B, 'tiles per CTA', ntiles, ' prologue done', int(t[5, 0, 2]), ' ufull', int(t[5, 0, 0]), ' end', int(t[5, 0, 1]))
for r, nm in enumerate(names):
    print(nm)
    for i in range(ntiles):
        print('   tile', i, [int(v) for v in t[r, i]])
