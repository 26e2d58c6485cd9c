"""Fused clip + SGD and the whole training step (flat buffers, direct gradient accumulation) against
torch.optim.SGD / clip_grad_norm_ and the fp64 oracle; data-parallel equivalence on 2 gloo ranks."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from util import O, ROOT, rel

from ibinn_b200 import config, ops
from ibinn_b200.engine import FlatParams, TrainStep
from ibinn_b200.model import GenerativeClassifier

SMALL = {'model': {'n_coupling_blocks_conv': '[1, 1, 1]', 'fc_width': '8', 'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}}


def test_clip_and_sgd_kernels(dev):
    g = torch.Generator().manual_seed(0)
    n = 10007
    p = torch.randn(n, generator=g).to(dev)
    gr = (torch.randn(n, generator=g) * 3).to(dev)
    buf = torch.randn(n, generator=g).to(dev)
    pad = torch.zeros(16 - n % 16 if n % 16 else 0)     # the kernels only need 16-byte aligned starts
    hyper = torch.tensor([0.07, 0.9, 1e-4]).to(dev)
    p0, g0, b0 = p.clone(), gr.clone(), buf.clone()
    nc = ops.grad_norm_clip(gr, 8.0)
    ops.sgd_step_(p, gr, buf, hyper, nc[1:], 1.0)
    norm = g0.double().norm()
    coef = min(1., 8.0 / (norm.item() + 1e-6))
    assert rel(nc[0:1], norm.view(1)) < 1e-6 and abs(nc[1].item() - coef) < 1e-6
    d = g0.double() * coef + 1e-4 * p0.double()
    b1 = 0.9 * b0.double() + d
    assert rel(buf, b1) < 1e-6
    assert rel(p, p0.double() - 0.07 * b1) < 1e-6


def _model(dev, seed=0, overrides=SMALL):
    args = config.make_args(overrides=overrides)
    torch.manual_seed(seed)
    np.random.seed(seed)
    return GenerativeClassifier(args).to(dev)


def _batch(B, C, gen):
    x = torch.randn(B, 3, 32, 32, generator=gen)
    lab = torch.randint(0, C, (B,), generator=gen)
    y = torch.zeros(B, C).scatter_(1, lab.view(-1, 1), 1.) * 0.98 + 0.02 / C
    return x, y


def test_train_step_matches_torch_loop(dev):
    """TrainStep (flat buffers, in-place gradient accumulation, fused clip+SGD) == the reference's loop
    (loss.backward(); clip_grad_norm_; optimizer.step()) run on the same ibinn_b200 model."""
    gen = torch.Generator().manual_seed(1)
    m1 = _model(dev)
    m2 = copy.deepcopy(m1)
    step = TrainStep(m2, batch_size=4, beta=1.0, clip=8., use_graph=False)
    for it in range(2):
        x, y = _batch(4, 10, gen)
        x, y = x.to(dev), y.to(dev)
        out = m1(x, y)
        loss = 1.0 * out['L_x_tr'] - 1.0 * out['L_y_tr']
        loss.backward()
        n1 = torch.nn.utils.clip_grad_norm_(m1.trainable_params, 8.)
        m1.optimizer.step()
        m1.optimizer.zero_grad()
        o2 = step.step(x, y)
        assert rel(o2['L_x_tr'], out['L_x_tr']) < 1e-6
        assert rel(o2['grad_norm'], n1) < 1e-4
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    for k in sd1:
        if sd1[k].is_floating_point():
            assert rel(sd2[k], sd1[k]) < 2e-4, k
    # the torch optimiser's momentum state and ours agree too
    p1 = m1.optimizer.param_groups[0]['params'][0]
    p2 = m2.optimizer.param_groups[0]['params'][0]
    b2 = step.flat.buf[:p2.numel()].view(p2.shape)
    assert rel(b2, m1.optimizer.state[p1]['momentum_buffer']) < 2e-4


def test_train_step_follows_lr_schedule(dev):
    m = _model(dev)
    step = TrainStep(m, batch_size=2, use_graph=False)
    sched = torch.optim.lr_scheduler.MultiStepLR(m.optimizer, milestones=[1], gamma=0.1)
    gen = torch.Generator().manual_seed(2)
    x, y = _batch(2, 10, gen)
    step.step(x.to(dev), y.to(dev))
    m.optimizer.step = lambda *a, **k: None      # the scheduler only needs the call order to be sane
    sched.step()
    step.step(x.to(dev), y.to(dev))
    assert step.flat.segments[0]['hyper'][0].item() == pytest.approx(0.007)
    assert step.flat.segments[1]['hyper'][0].item() == pytest.approx(0.021)


def _ddp_worker(rank, world, port, lib_path, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
    import torch.distributed as dist
    from ibinn_b200 import _lib
    _lib._use_library_for_tests(lib_path, True)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cpu')
    m = _model(dev, seed=rank)          # replicas are seeded differently on purpose: TrainStep broadcasts rank 0's model
    step = TrainStep(m, batch_size=3, beta=1.0, clip=8., use_graph=False)
    gen = torch.Generator().manual_seed(5)
    for it in range(2):
        xs, ys = _batch(3 * world, 10, gen)
        step.step(xs[rank * 3:(rank + 1) * 3], ys[rank * 3:(rank + 1) * 3])
    torch.save({k: v for k, v in m.state_dict().items() if 'running' not in k and 'num_batches' not in k and 'tmp_var' not in k}, os.path.join(tmp, 'r%d.pt' % rank))
    dist.destroy_process_group()


def test_data_parallel_equals_gradient_accumulation(tmp_path):
    """2 gloo ranks (simulator kernels) == 1 process that runs the two micro-batches one after the other with
    gradient accumulation (BatchNorm statistics local to each micro-batch), SURVEY.md section 8e."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
    import build_emu
    lib_path = build_emu.build()
    from ibinn_b200 import _lib
    _lib._use_library_for_tests(lib_path, True)
    try:
        port = 29500 + os.getpid() % 2000
        mp.spawn(_ddp_worker, args=(2, port, lib_path, str(tmp_path)), nprocs=2, join=True)
        r0, r1 = torch.load(tmp_path / 'r0.pt'), torch.load(tmp_path / 'r1.pt')
        for k in r0:
            assert torch.equal(r0[k], r1[k]), k          # replicas stay bit-identical
        dev = torch.device('cpu')
        m = _model(dev)
        flat = FlatParams(m)
        gen = torch.Generator().manual_seed(5)
        for it in range(2):
            xs, ys = _batch(6, 10, gen)
            flat.zero_grad()
            for r in range(2):
                out = m(xs[r * 3:(r + 1) * 3], ys[r * 3:(r + 1) * 3])
                (out['L_x_tr'] - out['L_y_tr']).backward()
            flat.clip_and_step(8., grad_scale=0.5)
        sd = m.state_dict()
        for k in r0:
            if r0[k].is_floating_point():
                assert rel(r0[k], sd[k]) < 1e-5, k
    finally:
        _lib._use_library_for_tests(None, False)


@pytest.mark.gpu
def test_train_step_schedules_agree():
    """The same step with every stream / graph schedule: eager on one stream, weight gradients on a side stream, grouped
    record reductions on a third stream, and the captured CUDA graphs: the same parameters after two steps.  The schedules
    only reorder independent launches; what remains is the fp32 atomic order of the split-K FC GEMMs (1e-7), which now and
    then flips one ReLU mask element and moves a small tensor by 1e-3 (`scripts/schedule_noise.py`: the same happens between
    two runs of ONE schedule) - hence a tight gate on the whole parameter vector and a loose one per tensor."""
    if not torch.cuda.is_available():
        pytest.skip('needs a B200')
    dev = torch.device('cuda:0')
    base = _model(dev, overrides={'model': {'n_coupling_blocks_conv': '[1, 2, 2]', 'fc_width': '16'}})
    gen = torch.Generator().manual_seed(5)
    batches = [_batch(8, 10, gen) for _ in range(2)]
    results = []
    for env, graph in (({'IBINN_WGRAD_STREAMS': '0'}, False), ({'IBINN_WGRAD_STREAMS': '1'}, False),
                       ({'IBINN_WGRAD_STREAMS': '1', 'IBINN_WGRAD_REDUCE_GROUP': '3'}, False), ({'IBINN_WGRAD_STREAMS': '1'}, True)):
        old = {k: os.environ.get(k) for k in ('IBINN_WGRAD_STREAMS', 'IBINN_WGRAD_REDUCE_GROUP')}
        os.environ.pop('IBINN_WGRAD_REDUCE_GROUP', None)
        os.environ.update(env)
        try:
            m = copy.deepcopy(base)
            m.eval()                                   # no dropout masks to draw
            step = TrainStep(m, batch_size=8, beta=1.0, clip=8., use_graph=graph)
            if graph:
                step.load(*[t.to(dev) for t in batches[0]])
                step.capture()                         # side-effect free (test_capture_has_no_training_side_effects)
            for x, y in batches:
                step.step(x.to(dev), y.to(dev))
            torch.cuda.synchronize()
            results.append({k: v.clone() for k, v in m.state_dict().items() if v.is_floating_point()})
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    flat = lambda r: torch.cat([r[k].flatten().double() for k in results[0]])
    for other in results[1:]:
        assert float((flat(other) - flat(results[0])).norm() / flat(results[0]).norm()) < 1e-4
        for k in results[0]:
            assert rel(other[k], results[0][k].double()) < 2e-2, k


EMP = {'model': dict(SMALL['model']), 'training': {'empirical_mu': 'True'}}


def test_empirical_mu_updates_in_place(dev):
    """`empirical_mu = True` (reference experiments_configs/cifar10/misc/empiric_mu.ini, model.py:115-120): the running
    class means are blended into `mu` IN PLACE, so mu stays a view of the flat parameter buffer and the step runs."""
    m = _model(dev, overrides=EMP)
    assert m.mu_empirical
    step = TrainStep(m, batch_size=4, beta=1.0, use_graph=False)
    gen = torch.Generator().manual_seed(3)
    ptr0 = m.mu.data_ptr()
    mu0 = m.mu.detach().clone()
    x, y = _batch(4, 10, gen)
    z = m.inn(x.to(dev)).detach()
    emp = (z.t().double() @ y.to(dev).round().double() / y.to(dev).double().sum(0, keepdim=True)).t().reshape(1, 10, -1)
    out = step.step(x.to(dev), y.to(dev))
    assert m.mu.data_ptr() == ptr0 and m.mu.is_contiguous()
    assert np.isfinite(float(out['L_x_tr']))
    # the blend happened before the optimiser touched mu: undo the SGD update (mu group: momentum 0, weight decay 0)
    o = (m.mu.data_ptr() - step.flat.flat.data_ptr()) // 4
    lr_mu = m.optimizer.param_groups[1]['lr']
    coef = min(1., 8. / (float(out['grad_norm']) + 1e-6))
    blended = m.mu.detach().double() + lr_mu * coef * step.flat.grad[o:o + m.mu.numel()].view_as(m.mu).double()
    assert rel(blended, 0.995 * mu0.double() + 0.005 * emp) < 1e-5


@pytest.mark.gpu
def test_empirical_mu_under_cuda_graph(cuda_dev):
    dev = cuda_dev
    gen = torch.Generator().manual_seed(3)
    batches = [_batch(4, 10, gen) for _ in range(3)]
    res = []
    for graph in (False, True):
        m = _model(dev, overrides=EMP)
        m.eval()
        step = TrainStep(m, batch_size=4, beta=1.0, use_graph=graph)
        for x, y in batches:
            step.step(x.to(dev), y.to(dev))
        torch.cuda.synchronize()
        res.append(m.mu.detach().clone())
    mu_init = _model(dev, overrides=EMP).mu.detach()
    assert rel(res[0], mu_init) > 5e-3          # three blends + three SGD updates moved mu ...
    assert rel(res[1], res[0]) < 1e-3           # ... and the replayed graph accumulated the same (a frozen address would not)


@pytest.mark.gpu
def test_capture_has_no_training_side_effects(cuda_dev):
    """TrainStep.capture() warms up with real steps; parameters, momentum, BatchNorm buffers and the RNG state must come
    back, so the first replayed batch gets exactly ONE update (the reference trajectory)."""
    dev = cuda_dev
    m = _model(dev, overrides={'model': {'n_coupling_blocks_conv': '[1, 1, 1]', 'fc_width': '8'}})     # dropout live
    step = TrainStep(m, batch_size=4, beta=1.0, use_graph=True)
    gen = torch.Generator().manual_seed(4)
    x, y = _batch(4, 10, gen)
    step.load(x.to(dev), y.to(dev))
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    buf0, rng0 = step.flat.buf.clone(), torch.cuda.get_rng_state(dev)
    step.capture()
    torch.cuda.synchronize()
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    assert torch.equal(step.flat.buf, buf0)
    assert torch.equal(torch.cuda.get_rng_state(dev), rng0)
    step.run()
    torch.cuda.synchronize()
    nbt = [v for k, v in m.state_dict().items() if k.endswith('num_batches_tracked')]
    assert all(int(v) == 1 for v in nbt)


def test_checkpoint_carries_momentum_and_load_is_strict(dev, tmp_path):
    m = _model(dev)
    step = TrainStep(m, batch_size=2, use_graph=False)
    gen = torch.Generator().manual_seed(6)
    x, y = _batch(2, 10, gen)
    step.step(x.to(dev), y.to(dev))
    f = str(tmp_path / 'ck.pt')
    m.save(f)
    ck = torch.load(f, weights_only=False)
    st = ck['opt']['state']
    n_inn = len(m.optimizer.param_groups[0]['params'])
    assert len(st) >= n_inn and all(st[i]['momentum_buffer'] is not None for i in range(n_inn))
    p0 = m.optimizer.param_groups[0]['params'][0]
    assert rel(st[0]['momentum_buffer'], step.flat.buf[:p0.numel()].view(p0.shape)) == 0.
    m2 = _model(dev, seed=1)
    step2 = TrainStep(m2, batch_size=2, use_graph=False)
    m2.load(f)
    assert torch.equal(step2.flat.buf.cpu(), step.flat.buf.cpu()) and torch.equal(step2.flat.flat.cpu(), step.flat.flat.cpu())
    # a checkpoint of another architecture is refused (reference model.py:246 loads strictly)
    other = _model(dev, overrides={'model': dict(SMALL['model'], n_coupling_blocks_conv='[1, 2, 1]')})
    with pytest.raises(RuntimeError):
        other.load(f)


@pytest.mark.gpu
def test_two_piece_backward_equals_single_pass(cuda_dev):
    """engine.TrainStep splits the backward pass after the fully connected block (that is where the gradient all-reduce of the
    FC / mu / phi tail starts on multi-GPU runs); forced on one GPU, eager and as [graph | graph | graph], it must give the
    parameters of the single-pass step."""
    dev = cuda_dev
    base = _model(dev, overrides={'model': {'n_coupling_blocks_conv': '[1, 2, 2]', 'fc_width': '16'}})
    gen = torch.Generator().manual_seed(5)
    batches = [_batch(8, 10, gen) for _ in range(2)]
    results = []
    for force, graph in (('0', False), ('1', False), ('1', True)):
        old = os.environ.get('IBINN_FORCE_SPLIT_BACKWARD')
        os.environ['IBINN_FORCE_SPLIT_BACKWARD'] = force
        try:
            m = copy.deepcopy(base)
            m.eval()
            step = TrainStep(m, batch_size=8, beta=1.0, clip=8., use_graph=graph)
            assert (step.early_start is not None) == (force == '1')
            for x, y in batches:
                step.step(x.to(dev), y.to(dev))
            torch.cuda.synchronize()
            if graph:
                assert step._g1b is not None
            results.append({k: v.clone() for k, v in m.state_dict().items() if v.is_floating_point()})
        finally:
            if old is None:
                os.environ.pop('IBINN_FORCE_SPLIT_BACKWARD', None)
            else:
                os.environ['IBINN_FORCE_SPLIT_BACKWARD'] = old
    flat = lambda r: torch.cat([r[k].flatten().double() for k in results[0]])
    for other in results[1:]:
        assert float((flat(other) - flat(results[0])).norm() / flat(results[0]).norm()) < 1e-4


@pytest.mark.gpu
def test_training_step_is_bit_reproducible(cuda_dev):
    """two engines from the same seed, the same batches, dropout on, CUDA graph, weight gradients on the side stream: every
    parameter, momentum buffer and BatchNorm statistic ends up bit-identical (all reductions run in a fixed order; the split-K
    GEMMs of the FC blocks sum their partial tiles in split order)"""
    dev = cuda_dev
    gen = torch.Generator().manual_seed(11)
    batches = [_batch(16, 10, gen) for _ in range(3)]
    res = []
    for rep in range(2):
        torch.cuda.manual_seed(0)
        m = _model(dev, overrides={})                      # the default CIFAR-10 network, dropout as shipped
        step = TrainStep(m, batch_size=16, beta=1.0, use_graph=True)
        for x, y in batches:
            step.step(x.to(dev), y.to(dev))
        torch.cuda.synchronize()
        res.append(({k: v.detach().clone() for k, v in m.state_dict().items()}, step.flat.buf.clone()))
    for k in res[0][0]:
        assert torch.equal(res[0][0][k], res[1][0][k]), k
    assert torch.equal(res[0][1], res[1][1])
