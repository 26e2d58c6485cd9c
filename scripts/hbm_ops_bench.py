#!/usr/bin/env python
"""Achieved HBM bandwidth of the HBM-bound kernels of the path (couplings, downsampling, DCT, head, optimiser) at a batch
where they are bandwidth- rather than launch-latency-limited (BASELINE.json configs[3]: large-batch scoring, B up to 8192).

  python scripts/hbm_ops_bench.py [--batch 8192] [--reps 20] [--out gpurun_out/hbm_ops.json]

Bytes are the ALGORITHMIC bytes of SURVEY.md section 8(d) / DESIGN.md section 3 (every operand read once, every result written
once).  Time = CUDA events around one CUDA graph of `reps` back-to-back launches over ROTATING operand sets whose
total size exceeds the 126 MB L2, after 3 warm-up launches.  Peak = MEASURED_PEAKS.json `hbm_gbs`.
At the training batch (128) the same kernels move 3-8 MB per launch and are latency-bound; that case is bench.py's."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--out', default=None)
    ap.add_argument('--only', default=None, help='substring filter on the kernel names')
    a = ap.parse_args()
    import torch
    from scipy.stats import special_ortho_group
    import __graft_entry__ as ge
    ge.build()
    from ibinn_b200 import ops
    assert torch.cuda.is_available(), 'needs a B200'
    dev = torch.device('cuda:0')
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        src = 'measured'
    except Exception:
        peak, src = 6650., 'fallback'
    B, reps = a.batch, a.reps
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    rows = []

    def run(name, nbytes, make, call, nsets=None):
        """make() -> operand tuple; call(*operands) launches once."""
        if a.only and a.only not in name:
            return
        nsets = nsets or max(2, int(160e6 // max(nbytes, 1)) + 1)
        sets = [make() for _ in range(nsets)]
        for i in range(3):
            call(*sets[i % nsets])
        torch.cuda.synchronize()
        # the reps are captured into one CUDA graph so that Python / allocator overhead is outside the timed region
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for i in range(reps):
                call(*sets[i % nsets])
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        gbs = nbytes / us * 1e-3
        rows.append(dict(kernel=name, batch=B, bytes_per_launch=int(nbytes), us_per_launch=us, achieved_gbs=gbs, frac=gbs / peak,
                         operand_sets=nsets))
        print('%-46s %8.1f MB %8.1f us %8.0f GB/s  %5.1f %%' % (name, nbytes / 1e6, us, gbs, 100 * gbs / peak), flush=True)
        del sets, gr
        torch.cuda.empty_cache()

    # ---- AIO coupling blocks (all_in_one_block.py:61-114): stage shapes of the CIFAR network
    for C, H in ((3, 32), (12, 16), (48, 8)):
        c1 = C - C // 2
        c2 = C // 2
        w = torch.tensor(special_ortho_group.rvs(C, random_state=0), dtype=torch.float32, device=dev).contiguous()
        an, ao = torch.full((C,), -0.28, device=dev), torch.zeros(C, device=dev)
        HW = H * H
        mk = lambda: (rn(B, C, H, H), 0.5 * rn(B, 2 * c2, H, H))
        run('aio_coupling_forward C=%d %dx%d' % (C, H, H), 4 * HW * (2 * C + 2 * c2) * B + 4 * B, mk,
            lambda x, a_: ops.aio_coupling_forward(x, a_, w, an, ao, 2.0))

        def mkb():
            x, a_ = rn(B, C, H, H), 0.5 * rn(B, 2 * c2, H, H)
            out, _ = ops.aio_coupling_forward(x, a_, w, an, ao, 2.0)
            return rn(B, C, H, H), rn(B), x, a_, out
        gn, go = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        ops.DEFERRED_AIO = []          # as in engine.TrainStep: the ActNorm-gradient records of all blocks are summed by one launch per step
        run('aio_coupling_backward C=%d %dx%d' % (C, H, H), 4 * HW * (3 * C + 4 * c2) * B, mkb,
            lambda go_, gj, x, a_, out: ops.aio_coupling_backward(go_, gj, x, a_, out, w, an, ao, 2.0, 0.1, gn, go))
        ops.DEFERRED_AIO = None
        del c1

    # ---- space-to-depth coupling (downsampling_coupling_block.py:49-97): DOWN_0 and DOWN_1 input shapes
    for cp, H in ((2, 32), (6, 16)):
        def mks(cp=cp, H=H):
            return rn(B, cp, H, H), 0.5 * rn(B, 8 * cp, H // 2, H // 2), torch.empty(B, 4 * cp, H // 2, H // 2, device=dev), torch.zeros(B, device=dev)
        run('s2d_coupling_forward C=%d %dx%d' % (cp, H, H), 4 * 4 * cp * H * H * B, mks,
            lambda x, a_, out, jac: ops.s2d_coupling_forward(x, a_, out, jac, False, 2.0))

    # ---- DCT pooling (dct_transform.py:64-104)
    M = torch.linalg.qr(rn(8, 8))[0].contiguous()
    run('dct2d_forward C=48 8x8', 8 * 3072 * B, lambda: (rn(B, 48, 8, 8),), lambda x: ops.dct2d_forward(x, M, 1.0))
    # ---- FrEIA PermuteRandom + GLOW affine halves (inn_architecture.py:160-161)
    idx = torch.randperm(3072, device=dev, generator=g).to(torch.int32)
    run('permute_features D=3072', 8 * 3072 * B, lambda: (rn(B, 3072),), lambda x: ops.permute_features(x, idx))
    run('glow_affine_forward L=1536', 4 * 1536 * 4 * B,
        lambda: (rn(B, 1536), 0.5 * rn(B, 3072), torch.empty(B, 1536, device=dev), torch.zeros(B, device=dev)),
        lambda x, r, y, jac: ops.glow_affine_forward(x, r, y, jac, False, 2.0))
    # ---- GMM / IB head (model.py:113-126,144-163)
    for C in (10, 100):
        mu, phi = rn(C, 3072), torch.zeros(C, device=dev)

        def mkh(C=C):
            lab = torch.randint(0, C, (B,), device=dev, generator=g)
            y = torch.zeros(B, C, device=dev).scatter_(1, lab.view(-1, 1), 1.) * 0.98 + 0.02 / C
            return rn(B, 3072), rn(B), y
        run('gmm_head_forward C=%d' % C, 4 * 3072 * B + 4 * C * 3072 + 4 * (C + 3) * B, mkh,
            lambda z, jac, y, _ws={}: ops.gmm_head_forward(z, jac, mu, phi, y, head_ws=_ws))      # scoring loop: class means prepared once
    # ---- optimiser (train.py:109-110): 11.19 M parameters, read p,g,buf / write p,buf; norm reads g
    n = 11_193_154

    def mko():
        return torch.zeros(n, device=dev), rn(n), torch.zeros(n, device=dev)
    hyper = torch.tensor([0.07, 0.9, 1e-4], device=dev)
    run('sgd_step 11.19 M params', 4 * 5 * n, mko, lambda p, gr, buf: ops.sgd_step_(p, gr, buf, hyper), nsets=2)
    run('grad_norm_clip 11.19 M params', 4 * n, lambda: (rn(n),), lambda gr: ops.grad_norm_clip(gr, 8.0), nsets=4)

    res = dict(peak_gbs=peak, peak_source=src + ' (MEASURED_PEAKS.json hbm_gbs)', batch=B, reps=reps, rows=rows)
    if a.out:
        json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
