"""Data-parallel training over NCCL on real GPUs (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_nccl.py -m gpu`):
N ranks == one process that runs the N micro-batches one after the other with gradient accumulation, with the
eager schedule (early all-reduce of the FC / mu / phi tail overlapped with the backward pass) and with the whole step captured as one
CUDA graph (collectives inside the graph)."""
import os
import sys

import pytest
import torch

from util import ROOT, rel

pytestmark = pytest.mark.gpu
OV = {'model': {'n_coupling_blocks_conv': '[1, 2, 2]', 'fc_width': '32', 'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}}
B = 8


def _batch(n, gen):
    x = torch.randn(n, 3, 32, 32, generator=gen)
    lab = torch.randint(0, 10, (n,), generator=gen)
    y = torch.zeros(n, 10).scatter_(1, lab.view(-1, 1), 1.) * 0.98 + 0.002
    return x, y


def _model(dev, seed):
    import numpy as np
    from ibinn_b200 import config
    from ibinn_b200.model import GenerativeClassifier
    torch.manual_seed(seed)
    np.random.seed(seed)
    return GenerativeClassifier(config.make_args(overrides=OV)).to(dev)


def _worker(rank, world, port, use_graph, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), NCCL_DEBUG='WARN')
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
    import torch.distributed as dist
    from ibinn_b200.engine import TrainStep
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    m = _model(dev, seed=rank)             # differently seeded replicas: TrainStep broadcasts rank 0's
    step = TrainStep(m, batch_size=B, beta=1.0, clip=8., use_graph=use_graph)
    assert step.early_start is not None and 0 < step.early_start < step.flat.grad.numel()
    gen = torch.Generator().manual_seed(5)
    for it in range(3):
        xs, ys = _batch(B * world, gen)
        step.step(xs[rank * B:(rank + 1) * B].to(dev), ys[rank * B:(rank + 1) * B].to(dev))
    torch.cuda.synchronize()
    sd = {k: v.cpu() for k, v in m.state_dict().items() if 'running' not in k and 'num_batches' not in k and 'tmp_var' not in k}
    torch.save({'sd': sd, 'one_graph': bool(getattr(step, '_one_graph', False))}, os.path.join(tmp, 'r%d_%d.pt' % (rank, int(use_graph))))
    dist.destroy_process_group()


@pytest.mark.parametrize('use_graph', [False, True])
def test_nccl_data_parallel_equals_gradient_accumulation(tmp_path, use_graph):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from ibinn_b200.engine import FlatParams
    world = 2
    port = 29600 + os.getpid() % 2000 + int(use_graph)
    mp.spawn(_worker, args=(world, port, use_graph, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / ('r%d_%d.pt' % (i, int(use_graph))), weights_only=False) for i in range(world)]
    for k in r[0]['sd']:
        assert torch.equal(r[0]['sd'][k], r[1]['sd'][k]), k          # replicas stay bit-identical
    if use_graph:
        print('collectives captured inside the step graph:', r[0]['one_graph'])
    dev = torch.device('cuda:0')
    m = _model(dev, seed=0)
    flat = FlatParams(m)
    gen = torch.Generator().manual_seed(5)
    for it in range(3):
        xs, ys = _batch(B * world, gen)
        flat.zero_grad()
        for q in range(world):
            out = m(xs[q * B:(q + 1) * B].to(dev), ys[q * B:(q + 1) * B].to(dev))
            (out['L_x_tr'] - out['L_y_tr']).backward()
        flat.clip_and_step(8., grad_scale=1. / world)
    sd = m.state_dict()
    num = den = 0.
    for k, v in r[0]['sd'].items():
        if v.is_floating_point():
            d = v.double() - sd[k].cpu().double()
            num += float((d ** 2).sum())
            den += float((sd[k].double() ** 2).sum())
            assert rel(v, sd[k]) < 2e-2, k      # per tensor loose (ReLU-flip noise of the split-K atomics, see test_engine), whole vector tight
    assert (num / den) ** 0.5 < 1e-4
