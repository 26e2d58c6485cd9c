"""The fused run-of-AIO-blocks kernel (ibinn_aio_stage_forward, csrc/block_tc.cu: one CTA per image, activations resident in
shared memory, four tcgen05 contractions per block incl. the soft permutation) against the per-op path it replaces and
against the fp64 oracle: eval mode (persistent over images), train mode (grid-wide BatchNorm statistics, saved tensors,
running statistics, backward through functions.AIOStageFn), and the whole default model with / without fusion.  GPU only."""
import copy
import os
from functools import partial

import numpy as np
import pytest
import torch

from util import O, randomize_, rel, state_of

from ibinn_b200 import config, functions as Fn, ops
from ibinn_b200.all_in_one_block import AIO_Block
from ibinn_b200.inn_architecture import _weights_init
from ibinn_b200.model import GenerativeClassifier
from ibinn_b200.subnets import ConvSubnet

pytestmark = pytest.mark.gpu
BN_ARGS = {'track_running_stats': True, 'momentum': 0.999, 'eps': 1e-4}


def conv_constr(width, dropout, relu_first, cin, cout):
    net = ConvSubnet(cin, cout, width, dropout, relu_first, False, BN_ARGS)
    net.apply(_weights_init)
    return net


def make_run(C, H, width, nblk, dropout=0., seed=0, first_relu=True):
    torch.manual_seed(seed)
    np.random.seed(seed)
    g = torch.Generator().manual_seed(seed + 1)
    blocks = []
    for k in range(nblk):
        blk = AIO_Block([(C, H, H)], subnet_constructor=partial(conv_constr, width, dropout, first_relu or k > 0), clamp=0.7,
                        act_norm=0.75, permute_soft=True)
        randomize_(blk, g)
        blocks.append(blk)
    return blocks


def oracle_run(blocks, x, train, masks=None):
    """fp64 oracle of the run: out, jac and per-block outputs"""
    h = x.double()
    jac = torch.zeros(x.shape[0], dtype=torch.float64)
    outs = []
    new_stats = []
    for k, blk in enumerate(blocks):
        C, H = blk.in_channels, int(round(blk.n_pixels ** 0.5))
        spec = O.NodeSpec('aio', 1, 'aio', (C, H, H), (C, H, H), dict(width=blk.s.conv1.weight.shape[0], dropout=0.,
                                                                       relu_first=blk.s.relu_first, clamp=0.7))
        P = state_of(blk, spec.key + '.')
        mk = None
        if masks is not None and masks[k] is not None:
            mk = {'%s.s.%d' % (spec.key, 7 if blk.s.relu_first else 6): masks[k].double().cpu()}
        st = O._State(train, mk, False)
        with torch.no_grad():
            h, j = O.node_forward(h, P, spec, st)
        jac = jac + j
        outs.append(h)
        new_stats.append(st.new_stats)
    return h, jac, outs, new_stats


SHAPES = [(12, 16, 32), (48, 8, 64), (4, 14, 32), (16, 7, 64)]          # CIFAR stages 1 / 2, MNIST stages


@pytest.mark.parametrize('C,H,width', SHAPES)
@pytest.mark.parametrize('nblk,B', [(1, 3), (3, 5), (2, 300)])
def test_fused_run_eval(cuda_dev, C, H, width, nblk, B):
    dev = cuda_dev
    blocks = make_run(C, H, width, nblk)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, H, generator=g)
    ref, jref, _, _ = oracle_run(blocks, x, False)
    for b in blocks:
        b.to(dev).eval()
    plan = ops.AioStagePlan(blocks)
    assert plan.supported(x.to(dev), False)
    with torch.no_grad():
        res = ops.aio_stage_forward(plan, x.to(dev), False)
        h = x.to(dev)
        jun = 0.
        for b in blocks:                       # the per-op path
            h = b([h])[0]
            jun = jun + b.jacobian([h])
    torch.cuda.synchronize()
    assert rel(res['out'], ref) < 1e-5 * nblk and rel(res['jac'], jref) < 2e-5
    assert rel(res['out'], h) < 1e-5 * nblk and rel(res['jac'], jun) < 2e-5


@pytest.mark.parametrize('C,H,width', SHAPES)
@pytest.mark.parametrize('nblk,B,dropout', [(1, 4, 0.), (3, 6, 0.), (2, 128, 0.25)])
def test_fused_run_train(cuda_dev, C, H, width, nblk, B, dropout):
    """forward (batch statistics through the grid barrier), saved tensors, running statistics and the backward"""
    dev = cuda_dev
    blocks = make_run(C, H, width, nblk, dropout=dropout)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, C, H, H, generator=g)
    masks = [((torch.rand(B, width, generator=g) >= dropout).float() / (1. - dropout)) if dropout > 0 else None for _ in blocks]
    ref, jref, outs64, stats64 = oracle_run(blocks, x, True, masks)
    unf = [copy.deepcopy(b).to(dev).train() for b in blocks]
    for b in blocks:
        b.to(dev).train()
    plan = ops.AioStagePlan(blocks)
    assert plan.supported(x.to(dev), True)
    # ---- fused forward + backward
    for b, m in zip(blocks, masks):
        if m is not None:
            b.s._ibinn_mask, b.s._ibinn_mask_fresh = m.to(dev), True
    xf = x.to(dev).requires_grad_()
    out, jac = Fn.AIOStageFn.apply(xf, plan, *Fn.aio_stage_params(plan))
    gout = torch.randn(out.shape, generator=g).to(dev)
    gjac = (0.01 * torch.randn(B, generator=g)).to(dev)
    ((out * gout).sum() + (jac * gjac).sum()).backward()
    # ---- per-op path on copies of the same blocks
    xu = x.to(dev).requires_grad_()
    h, ju = xu, 0.
    for b, m in zip(unf, masks):
        h = b([h], mask=None if m is None else m.to(dev))[0]
        ju = ju + b.jacobian([h])
    ((h * gout).sum() + (ju * gjac).sum()).backward()
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-5 * nblk and rel(jac, jref) < 2e-5
    assert rel(out, h) < 1e-5 * nblk and rel(jac, ju) < 2e-5
    for k, (bf, bu) in enumerate(zip(blocks, unf)):
        sf, su = bf.state_dict(), bu.state_dict()
        for key in sf:
            if 'running' in key:
                assert rel(sf[key], su[key]) < 1e-5, (k, key)
                assert rel(sf[key], stats64[k]['module_list.1.' + key]) < 1e-5, (k, key)
            if 'num_batches' in key:
                assert int(sf[key]) == int(su[key]) == 1
    # gradients: both paths run the same backward kernels on (nearly) the same saved tensors
    # (the two forwards differ in the last bits of the batch statistics: at B = 128 a ReLU input within 1e-7 of zero can fall on
    # either side, which changes the gradient inside that position's receptive field - a handful of elements - and nothing else)
    off = ((xf.grad - xu.grad).abs() > 1e-3 * xu.grad.abs().max()).float().mean().item()
    assert rel(xf.grad, xu.grad) < 1e-3 or off < 5e-3, (rel(xf.grad, xu.grad), off)
    for bf, bu in zip(blocks, unf):
        for (n1, p1), (n2, p2) in zip(bf.named_parameters(), bu.named_parameters()):
            if p1.requires_grad:
                assert p1.grad is not None, n1
                assert rel(p1.grad, p2.grad) < (1e-2 if B >= 128 else 2e-3), n1      # (B = 128: one flipped ReLU moves a sum by ~3e-3)


NO_DROP = {'model': {'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}}


@pytest.mark.parametrize('mode,B', [('eval', 64), ('train', 128)])
def test_default_model_fused_equals_unfused(cuda_dev, mode, B):
    """the whole default CIFAR-10 network with the stage runs fused and with one launch group per node"""
    dev = cuda_dev
    args = config.make_args(overrides=dict(NO_DROP, **{'model': dict(NO_DROP['model'], weight_init='0.5')}))
    torch.manual_seed(0)
    np.random.seed(0)
    m = GenerativeClassifier(args).to(dev)
    m.train(mode == 'train')
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    res = {}
    for fused in (True, False):
        mm = copy.deepcopy(m)
        mm.inn.fused_stages = fused
        with torch.no_grad():
            z = mm.inn(x)
            jac = mm.inn.log_jacobian(run_forward=False)
        res[fused] = (z, jac, {k: v.clone() for k, v in mm.inn.state_dict().items() if 'running' in k})
        if fused:
            assert len(mm.inn._node_inputs) < 20          # the 48 stage-1 / stage-2 blocks went through two launches
    assert rel(res[True][0], res[False][0]) < 2e-4 and rel(res[True][1], res[False][1]) < 1e-5
    if mode == 'train':
        for k, v in res[True][2].items():
            assert rel(v, res[False][2][k]) < 1e-4, k


@pytest.mark.parametrize('C,H,width,B', [(12, 16, 32, 128), (48, 8, 64, 100)])
def test_fused_run_train_is_deterministic(cuda_dev, C, H, width, B):
    """two launches on the same inputs give bit-identical outputs, saved tensors and statistics (fixed-order reductions, no atomics)"""
    dev = cuda_dev
    blocks = make_run(C, H, width, 3)
    for b in blocks:
        b.to(dev).train()
    plan = ops.AioStagePlan(blocks)
    x = torch.randn(B, C, H, H, generator=torch.Generator().manual_seed(5)).to(dev)
    runs = []
    for _ in range(3):
        res = ops.aio_stage_forward(plan, x, True, [None] * 3)
        torch.cuda.synchronize()
        runs.append({k: res[k].clone() for k in ('out', 'jac', 'outs', 'h1', 'h2', 'a', 'save')})
    for r in runs[1:]:
        for k, v in r.items():
            assert torch.equal(v, runs[0][k]), k


@pytest.mark.parametrize('C,H,width,dropout', [(12, 16, 32, 0.), (48, 8, 64, 0.25)])
def test_fused_run_invertible_backward(cuda_dev, C, H, width, dropout):
    """Backward that exploits invertibility (functions.INVERTIBLE_BACKWARD): the forward keeps only the run's output and the batch
    statistics; every block's input, subnet output and raw conv outputs are recomputed from its OUTPUT (all_in_one_block.py:93-103).
    Gradients must agree with the stored-activation backward up to the rounding of the reconstruction."""
    dev, nblk, B = cuda_dev, 4, 64
    blocks = make_run(C, H, width, nblk, dropout=dropout)
    for b in blocks:
        b.to(dev).train()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, C, H, H, generator=g)
    masks = [((torch.rand(B, width, generator=g) >= dropout).float() / (1. - dropout)).to(dev) if dropout > 0 else None for _ in blocks]
    gout = torch.randn(B, C, H, H, generator=g).to(dev)
    gjac = (0.01 * torch.randn(B, generator=g)).to(dev)
    res = {}
    for inv in (False, True):
        bl = [copy.deepcopy(b) for b in blocks]
        plan = ops.AioStagePlan(bl)
        for b, m in zip(bl, masks):
            if m is not None:
                b.s._ibinn_mask, b.s._ibinn_mask_fresh = m, True
        Fn.INVERTIBLE_BACKWARD = inv
        try:
            xf = x.to(dev).requires_grad_()
            out, jac = Fn.AIOStageFn.apply(xf, plan, *Fn.aio_stage_params(plan))
            kept = sum(t.numel() for t in out.grad_fn.saved_tensors) if hasattr(out.grad_fn, 'saved_tensors') else 0
            ((out * gout).sum() + (jac * gjac).sum()).backward()
        finally:
            Fn.INVERTIBLE_BACKWARD = False
        torch.cuda.synchronize()
        res[inv] = (out.detach(), xf.grad, [p.grad for b in bl for p in b.parameters() if p.requires_grad], kept)
    assert torch.equal(res[True][0], res[False][0])
    assert res[True][3] * 6 < res[False][3]                  # the forward keeps less than a sixth of the floats (anchor every 2nd block)
    # the reconstruction is good to ~1e-6 per inverted block; what shows in the gradients are the few ReLU decisions it flips
    assert rel(res[True][1], res[False][1]) < 2e-2
    for a, b in zip(res[True][2], res[False][2]):
        assert rel(a, b) < 5e-2
    print('saved floats: stored %d, invertible %d' % (res[False][3], res[True][3]))
