"""Slot transition function (predictor.py:20-44, MOVi-D geometry: 2 layers, 4 heads, D = 192, ffn 768, 11 slots): forward +
backward of slotdiffusion_b200.predictor.TransformerPredictor against torch's nn.TransformerEncoder on the same GPU, eager and
replayed from a CUDA graph (the form the training step uses).  usage: python tools/predictor_bench.py [B]"""
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slotdiffusion_b200.predictor import TransformerPredictor  # noqa: E402

warnings.simplefilter('ignore')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, D = 11, 192
dev = torch.device('cuda')
ours = TransformerPredictor(D, 2, 4, 4 * D, True).to(dev).train()
layer = torch.nn.TransformerEncoderLayer(d_model=D, nhead=4, dim_feedforward=4 * D, norm_first=True, batch_first=True)
ref = torch.nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False).to(dev).train()
x = torch.randn(B, S, D, device=dev, requires_grad=True)
gw = torch.randn(B, S, D, device=dev)


def step(net):
    for p in net.parameters():
        p.grad = None
    x.grad = None
    (net(x) * gw).sum().backward()


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def graphed(net):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step(net)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step(net)
    return g.replay


res = {'B': B, 'S': S, 'D': D, 'unit': 'us per forward+backward'}
for name, net in (('ours', ours), ('torch', ref)):
    for tf32 in ((False, True) if name == 'torch' else (False,)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        key = name + ('_tf32' if tf32 else '')
        res[key + '_eager'] = round(timed(lambda: step(net)), 1)
        res[key + '_graph'] = round(timed(graphed(net)), 1)
print(json.dumps(res))
