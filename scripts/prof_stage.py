"""A few launches of the fused AIO-run kernel (csrc/block_tc.cu) at the bench geometries, for `ncu --set full -k regex:aio_stage_kernel`:
the 24-block runs of the default CIFAR-10 model, eval mode at the scoring batch (2048) and train mode at the training batch (128)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
from test_fused_stage import make_run       # noqa: E402
from ibinn_b200 import ops                  # noqa: E402

dev = torch.device('cuda:0')
for (C, H, width) in ((12, 16, 32), (48, 8, 64)):
    for train, B in ((False, 2048), (True, 128)):
        blocks = make_run(C, H, width, 24)
        for b in blocks:
            b.to(dev).train(train)
        plan = ops.AioStagePlan(blocks)
        x = torch.randn(B, C, H, H, device=dev)
        for it in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.aio_stage_forward(plan, x, train, [None] * 24)
            e1.record()
            torch.cuda.synchronize()
        print('C=%d H=%d width=%d %s B=%d: %.1f us' % (C, H, width, 'train' if train else 'eval', B, e0.elapsed_time(e1) * 1e3))
