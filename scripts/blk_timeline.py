"""Phase time line of the fused AIO-run kernel (csrc/block_tc.cu built with -DIBINN_BLK_DBG; globaltimer stamps of CTA 0,
first image, first two blocks).  Build the instrumented library here, run on the GPU box:
    python ibinn_b200/csrc/build.py --variant dbg IBINN_BLK_DBG
    IBINN_B200_LIB=$PWD/ibinn_b200/csrc/libibinn_b200_dbg.so python scripts/blk_timeline.py"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
from test_fused_stage import make_run       # noqa: E402
from ibinn_b200 import _lib, ops            # noqa: E402

E = ['top', 'x1+tables', 'conv1 done(t0)', 'B done', 'stats2 done', 'conv2 done(t0)', 'D0 done', 'conv3 done(t0)', 'F0 done', 'tail end', 'perm done(t0)']
M = ['start', 'x1_full seen', 'conv1 issued', 'h1 seen', 'conv2 issued', 'tail issued']


def main():
    dev = torch.device('cuda:0')
    L = ctypes.CDLL(_lib.LIB_PATH)
    for (C, H, width) in ((12, 16, 32), (48, 8, 64)):
        for train, B in ((False, 296), (True, 128)):
            blocks = make_run(C, H, width, 4)
            for b in blocks:
                b.to(dev).train(train)
            plan = ops.AioStagePlan(blocks)
            x = torch.randn(B, C, H, H, device=dev)
            for it in range(3):
                L.ibinn_blk_debug_clear()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.aio_stage_forward(plan, x, train, [None] * 4)
                e1.record()
                torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * 256)()
            L.ibinn_blk_debug_read(buf)
            v = np.array(list(buf), dtype=np.float64)
            t0 = v[0]
            print('\n== C=%d H=%d width=%d %s B=%d: kernel %.1f us (4 blocks)' % (C, H, width, 'train' if train else 'eval', B, e0.elapsed_time(e1) * 1e3))
            for k in range(2):
                print(' block %d epilogue: ' % k + ', '.join('%s %.2f' % (n, (v[32 * k + i] - t0) / 1e3) for i, n in enumerate(E) if v[32 * k + i] > 0))
                print(' block %d mma     : ' % k + ', '.join('%s %.2f' % (n, (v[128 + 32 * k + i] - t0) / 1e3) for i, n in enumerate(M) if v[128 + 32 * k + i] > 0))
            print(' ring-wait cycles (MMA issuer): conv1 %s, conv2 %s; first conv2 chunk seen at %s us' % ([int(v[200 + k]) for k in range(2)], [int(v[204 + k]) for k in range(2)], ['%.2f' % ((v[208 + k] - t0) / 1e3) for k in range(2)]))
            for k in range(2):
                dt = v[128 + 32 * k + 4] - v[128 + 32 * k + 3]
                if dt > 0:
                    print(' block %d conv2 issue: %.0f SM cycles in %.0f ns -> %.2f GHz' % (k, v[216 + k] - v[212 + k], dt, (v[216 + k] - v[212 + k]) / dt))
            base = v[213]
            print(' conv2 chunks of block 1 (cycles from start: after ring wait / after issue):', ' '.join('%d/%d' % (v[220 + 2 * c] - base, v[221 + 2 * c] - base) for c in range(9) if v[220 + 2 * c] > 0))
            print(' conv2 of block 1, issuer cycles in: mbarrier wait %d, tcgen05 fence %d, tcgen05.commit %d' % (v[240], v[241], v[242]))
            if train:
                S = ['pass1a', 'pass1b', 'records', 'barrier', 'combine']
                for ph in range(2):
                    print(' block 1 stats %d  : ' % ph + ', '.join('%s %.2f' % (n, (v[16 + 8 * ph + i] - t0) / 1e3) for i, n in enumerate(S)))


if __name__ == '__main__':
    main()
