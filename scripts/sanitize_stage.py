"""Small launches of the fused kernels for `compute-sanitizer --tool memcheck` (one tool per gpurun call, B200_PROFILING.md):
the fused AIO run in eval and train mode, the tensor-core GMM head, the 1x1 backward, the augmentation kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
from test_fused_stage import make_run       # noqa: E402
from ibinn_b200 import config, ops          # noqa: E402
from ibinn_b200.data_gpu import GpuAugmentor  # noqa: E402

dev = torch.device('cuda:0')
for (C, H, width) in ((12, 16, 32), (48, 8, 64), (4, 14, 32)):
    for train, B in ((False, 150), (True, 9)):
        blocks = make_run(C, H, width, 2)
        for b in blocks:
            b.to(dev).train(train)
        res = ops.aio_stage_forward(ops.AioStagePlan(blocks), torch.randn(B, C, H, H, device=dev), train, [None] * 2)
        torch.cuda.synchronize()
        assert torch.isfinite(res['out']).all()
B, Cc, D = 1100, 100, 3072
ops.gmm_head_forward(torch.randn(B, D, device=dev), torch.randn(B, device=dev), torch.randn(Cc, D, device=dev), torch.zeros(Cc, device=dev), None)
ops.FUSED_1X1_BWD = True
from ibinn_b200.ops import BN               # noqa: E402
g, h = torch.randn(3, 12, 16, 16, device=dev), torch.randn(3, 32, 16, 16, device=dev)
bn = BN(torch.zeros(32, device=dev), torch.ones(32, device=dev), torch.ones(32, device=dev), torch.zeros(32, device=dev), 1e-4, False)
ops.conv1x1_backward(g, h, bn, None, torch.randn(12, 32, 1, 1, device=dev), torch.empty(2, 32, device=dev), None, None,
                     torch.zeros(12, 32, 1, 1, device=dev), torch.zeros(12, device=dev))
aug = GpuAugmentor(config.make_args(), dev)
aug(torch.randint(0, 256, (7, 32, 32, 3), dtype=torch.uint8, device=dev), train=True)
torch.cuda.synchronize()
print('sanitize_stage: done')
