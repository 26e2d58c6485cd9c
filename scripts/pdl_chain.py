"""Chain of dependent launches of one subnet convolution inside a CUDA graph: time per launch with and without programmatic
dependent launch (IBINN_PDL=0/1, separate processes).  Usage: IBINN_PDL=1 python scripts/pdl_chain.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ibinn_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
for (B, C, H) in ((128, 32, 16), (128, 64, 8)):
    x = torch.randn(B, C, H, H, device=dev)
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    N = 50

    def chain():
        h = x
        for _ in range(N):
            h = ops.conv_forward(h, w, 3)
        return h
    chain()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = chain()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / N)
    print('IBINN_PDL=%s  conv3x3 %d->%d @%dx%d B=%d: %.2f us per dependent launch' % (os.environ.get('IBINN_PDL', '1'), C, C, H, H, B, best))
