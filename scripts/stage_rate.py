"""Steady-state cost of one coupling block of the fused AIO-run kernel (csrc/block_tc.cu): kernel time at 1, 2 and 14 images per CTA,
differences divided by the images and blocks added (host overhead of the Python call stays outside: the launch is bracketed inside a
busy stream).  With IBINN_B200_LIB pointing at a `python ibinn_b200/csrc/build.py --variant nomma IBINN_EXP_NOMMA` (or
`--variant skiplo IBINN_EXP_SKIPLO`) build the differences give the marginal cost of the conv1 + conv2 MMAs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
from test_fused_stage import make_run       # noqa: E402
from ibinn_b200 import ops                  # noqa: E402

dev = torch.device('cuda:0')
spin = torch.empty(64 << 20, device=dev)


def timed(plan, x, train, n):
    best = 1e9
    for it in range(4):
        for _ in range(20):
            spin.add_(1.0)                          # ~50 us each: the stream is busy while the host prepares the launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.aio_stage_forward(plan, x, train, [None] * n, save_intermediates=train)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


for (C, H, width) in ((12, 16, 32), (48, 8, 64)):
    res = {}
    for nblk in (4, 24):
        blocks = make_run(C, H, width, nblk)
        for b in blocks:
            b.to(dev).eval()
        plan = ops.AioStagePlan(blocks)
        for B in (148, 296, 2072):
            with torch.no_grad():
                res[(nblk, B)] = timed(plan, torch.randn(B, C, H, H, device=dev), False, nblk)
        for b in blocks:
            b.train()
        res[(nblk, 'train')] = timed(plan, torch.randn(128, C, H, H, device=dev), True, nblk)
    print('C=%d H=%d width=%d' % (C, H, width), {k: round(v, 1) for k, v in res.items()})
    print('  eval: first image %.2f us/block, later images %.2f us/block (1 CTA = 1 image at a time, 148 CTAs) -> %.0f img/s per stage-run of 24'
          % ((res[(24, 148)] - res[(4, 148)]) / 20, (res[(24, 2072)] - res[(24, 296)]) / 12 / 24, 148e6 / ((res[(24, 2072)] - res[(24, 296)]) / 12)))
    print('  train (B=128): %.2f us/block' % ((res[(24, 'train')] - res[(4, 'train')]) / 20))
