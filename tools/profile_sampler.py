"""ncu target: warm-up, then the first evaluations of the un-captured DPM-Solver++ loop (UNet + x0/VQ + combinations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(os.environ.get('SDB_B', 256))
dev = torch.device('cuda', 0)
sa, unet, sampler, init_slots = bench.build_models(dev)
sampler.use_cuda_graph = False
x = torch.randn(B, 3, 32, 32, device=dev)
slots = torch.randn(B, bench.S, bench.D, device=dev)
with torch.no_grad():
    sampler.sample(x, slots)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    sampler.sample(x, slots)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
