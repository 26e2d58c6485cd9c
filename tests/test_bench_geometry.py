"""Parity at the geometries bench.py times (BASELINE.json configs[1]: CIFAR-10, batch 128 per GPU), not only at the toy
batches of the other files: the multi-tile persistent loop of conv_tc_kernel (16->16 @32x32 = 1089 tiles), its two
co-resident single-tile CTAs branch (32->32 @16x16 = 289 tiles), the 64 / 80 / 148 CTA caps of wgrad_tc_kernel with the
engine's deferred record reduction, aio_bwd_kernel's sample splitting, and a teacher-forced sweep of every node of the DEFAULT
[8, 24, 24] model at B = 128 against the fp64 oracle.  GPU only (`-m gpu`); everything goes through the C ABI."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import test_conv as TC
from util import O, rel

from ibinn_b200 import config, ops
from ibinn_b200.model import GenerativeClassifier
from ibinn_b200.ops import BN

pytestmark = pytest.mark.gpu

B = int(os.environ.get('IBINN_TEST_BENCH_B', '128'))      # 128 = bench.py's per-GPU batch; the override is for simulator dry runs
# every subnet convolution of the default CIFAR model: (Cin, Cout, H, ks, stride)
LAYERS = [
    (2, 16, 32, 3, 1), (16, 16, 32, 3, 1), (16, 2, 32, 1, 1),        # stage 0 AIO subnet
    (6, 32, 16, 3, 1), (32, 32, 16, 3, 1), (32, 12, 16, 1, 1),       # stage 1
    (24, 64, 8, 3, 1), (64, 64, 8, 3, 1), (64, 48, 8, 1, 1),         # stage 2
    (1, 32, 32, 3, 1), (32, 32, 32, 3, 2), (32, 16, 16, 1, 1),       # DOWN_0 s_hi
    (8, 32, 16, 3, 1), (32, 8, 16, 1, 1),                            # DOWN_0 s_lo
    (6, 64, 16, 3, 1), (64, 64, 16, 3, 2),                           # DOWN_1 s_hi (its 1x1 and s_lo equal stage 2's shapes)
]
IDS = ['%d-%d@%d_k%ds%d' % c for c in LAYERS]


@pytest.mark.parametrize('Cin,Cout,H,ks,st', LAYERS, ids=IDS)
def test_conv_forward_pipeline_at_bench_batch(cuda_dev, Cin, Cout, H, ks, st):
    """pre = relu(bn(.)) * dropout mask, epilogue = batch statistics + running-stat update (forward of conv2 / conv3)"""
    TC.test_conv_plain_bias_addend(cuda_dev, B, Cin, Cout, H, ks, st)
    TC.test_conv_bn_pipeline(cuda_dev, B, Cin, Cout, H, ks, st)


@pytest.mark.parametrize('Cin,Cout,H,ks,st', LAYERS, ids=IDS)
def test_conv_backward_pieces_at_bench_batch(cuda_dev, Cin, Cout, H, ks, st):
    """dgrad with the BatchNorm-backward pre-op and the mask / statistics epilogues, in-place ReLU-mask accumulate, wgrad"""
    TC.test_conv_backward_pieces(cuda_dev, B, Cin, Cout, H, ks, st)


@pytest.mark.parametrize('Cin,Cout,H,ks,st', [c for c in LAYERS if c[4] == 1], ids=[i for i, c in zip(IDS, LAYERS) if c[4] == 1])
def test_wgrad_deferred_records_at_bench_batch(cuda_dev, Cin, Cout, H, ks, st):
    """The engine's schedule: ibinn_conv_wgrad leaves per-CTA records (defer_reduce), ONE ibinn_wgrad_reduce_many sums the
    records of several layers; dw / dbias are accumulated into."""
    dev = cuda_dev
    g = torch.Generator().manual_seed(Cin * 13 + Cout)
    x = torch.randn(B, Cin, H, H, generator=g)
    gy = torch.randn(B, Cout, H, H, generator=g)
    w = (torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2).double().requires_grad_()
    b = torch.zeros(Cout).double().requires_grad_()
    (F.conv2d(F.relu(x.double()), w, b, padding=ks // 2) * gy.double()).sum().backward()
    dw0 = torch.randn(Cout, Cin, ks, ks, generator=g)
    db0 = torch.randn(Cout, generator=g)
    dws = [dw0.clone().to(dev) for _ in range(2)]
    dbs = [db0.clone().to(dev) for _ in range(2)]
    ops.DEFERRED_WGRAD = entries = []
    try:
        for dw, db in zip(dws, dbs):          # two layers' worth of records in one reduction launch
            ops.conv_wgrad(gy.to(dev), x.to(dev), dw, ks, 1, x_pre=ops.PRE_RELU, dbias=db if ks == 1 else None)
    finally:
        ops.DEFERRED_WGRAD = None
    live = [e for e in entries if e[1] is not None]
    if live:                                   # tensor-core path: nothing has been added yet
        assert torch.equal(dws[0].cpu(), dw0)
        ops.flush_deferred_wgrad(entries, dws[0], None)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    for dw, db in zip(dws, dbs):
        assert rel(dw, dw0.double() + w.grad) < 5e-5
        if ks == 1:
            assert rel(db, db0.double() + b.grad) < 5e-5


@pytest.mark.parametrize('C,H', [(3, 32), (12, 16), (48, 8)])
def test_aio_coupling_backward_at_bench_batch(cuda_dev, C, H):
    """aio_bwd_kernel at B = 128 (several CTAs per sample, deferred ActNorm records) against fp64 autograd of
    all_in_one_block.py:61-114"""
    from scipy.stats import special_ortho_group
    dev = cuda_dev
    c1, c2 = C - C // 2, C // 2
    g = torch.Generator().manual_seed(31 + C)
    x, a = torch.randn(B, C, H, H, generator=g), 0.7 * torch.randn(B, 2 * c2, H, H, generator=g)
    w = torch.tensor(special_ortho_group.rvs(C, random_state=5), dtype=torch.float32)
    an, ao = np.log(0.75) + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    go, gj = torch.randn(B, C, H, H, generator=g), 0.01 * torch.randn(B, generator=g)
    xd, ad = x.double().requires_grad_(), a.double().requires_grad_()
    and_, aod = an.double().requires_grad_(), ao.double().requires_grad_()
    sig = 0.7 * torch.tanh(0.1 * ad[:, :c2])
    u = torch.cat([xd[:, :c1], xd[:, c1:] * torch.exp(sig) + ad[:, c2:]], 1)
    out64 = torch.einsum('oi,bihw->bohw', w.double(), u) * and_.exp().view(1, C, 1, 1) + aod.view(1, C, 1, 1)
    jac64 = sig.sum((1, 2, 3)) + H * H * and_.sum()
    ((out64 * go.double()).sum() + (jac64 * gj.double()).sum()).backward()
    d = lambda t: t.to(dev)
    out, jac = ops.aio_coupling_forward(d(x), d(a), d(w), d(an.float()), d(ao), 0.7)
    assert rel(out, out64) < 1e-5 and rel(jac, jac64) < 2e-6
    for defer in (False, True):
        g_an, g_ao = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        ops.DEFERRED_AIO = entries = [] if defer else None
        try:
            gx, ga = ops.aio_coupling_backward(d(go), d(gj), d(x), d(a), out, d(w), d(an.float()), d(ao), 0.7, 0.1, g_an, g_ao)
        finally:
            ops.DEFERRED_AIO = None
        if defer:
            ops.flush_deferred_aio(entries, gx, None)
        if dev.type == 'cuda':
            torch.cuda.synchronize()
        assert rel(gx, xd.grad) < 1e-5 and rel(ga, ad.grad) < 1e-5
        assert rel(g_an, and_.grad) < 2e-5 and rel(g_ao, aod.grad) < 2e-5


NO_DROP = {'model': {'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}}


@pytest.fixture(scope='module')
def default_model():
    if not torch.cuda.is_available() and os.environ.get('IBINN_TEST_EMU_AS_CUDA') != '1':
        pytest.skip('no GPU')
    args = config.make_args(overrides=NO_DROP)            # the reference's default.ini: [8, 24, 24] blocks, fc 1024
    torch.manual_seed(0)
    np.random.seed(0)
    m = GenerativeClassifier(args)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, 32, 32, generator=g)
    return m, args, x


def test_default_model_every_node_teacher_forced_at_bench_batch(cuda_dev, default_model):
    """The bench's own model and batch: every one of the 61 nodes gets the input OUR pipeline produced for it (train-mode
    BatchNorm, B = 128) and is compared with the fp64 oracle on exactly that input (SURVEY.md section 8c gates: out 1e-5 of
    abs-max, log-det 2e-5).  The end-to-end z of this chaotic default init is reported as a ratio to the reference-fp32 error
    by test_full_depth (tests/test_golden_full.py)."""
    m, args, x = default_model
    dev = cuda_dev
    m = copy.deepcopy(m).to(dev)
    m.train()
    m.inn.fused_stages = False            # one launch group per node: every node's input / output / log-det is observable
    P = {k: v.detach().cpu().double().clone() for k, v in m.inn.state_dict().items() if 'tmp_var' not in k}
    spec = O.graph_spec(O.cfg_from_args(args))
    with torch.no_grad():
        m.inn(x.to(dev))
        inputs = {k: v.detach().cpu() for k, v in m.inn._node_inputs.items()}
        outs = dict(m.inn._jacs)
        jacs = {k: m.inn.module_list[k].jacobian([m.inn._node_inputs[k]]) for k in inputs}
    worst = (0., None)
    for n in spec:
        st = O._State(True, None, False)
        with torch.no_grad():
            o64, j64 = O.node_forward(inputs[n.idx].double(), P, n, st)
        e = rel(outs[n.idx], o64)
        worst = max(worst, (e, n.name))
        assert e < 1e-5, (n.name, e)
        if torch.is_tensor(j64):
            assert rel(jacs[n.idx], j64) < 2e-5, n.name
        else:
            assert abs(float(jacs[n.idx]) - float(j64)) <= 1e-6 * max(1., abs(float(j64))), n.name
    print('worst teacher-forced node error at B=128: %.2e (%s)' % worst)


@pytest.mark.parametrize('name', ['CONV_0_0', 'CONV_0_5', 'DOWN_0', 'CONV_1_11', 'DOWN_1', 'CONV_2_17', 'FC_0'])
def test_default_model_node_backward_at_bench_batch(cuda_dev, default_model, name):
    """forward + backward of single nodes of the default model at B = 128, flip-free (the oracle takes our ReLU decisions):
    input / parameter gradients within 3e-5 of the fp64 oracle's autograd, BatchNorm running statistics within 1e-5."""
    from test_blocks import run_node_flip_free
    m, args, x = default_model
    spec = {n.name: n for n in O.graph_spec(O.cfg_from_args(args))}[name]
    blk = copy.deepcopy(m.inn.module_list[spec.idx])
    g = torch.Generator().manual_seed(spec.idx)
    with torch.no_grad():                                   # exercise every gradient path: move biases / offsets off zero
        for k, p in blk.named_parameters():
            if p.requires_grad and (k.endswith('bias') or k.endswith('act_offset')):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    one = O.NodeSpec(spec.kind, 1, spec.name, spec.dims_in, spec.dims_out, spec.extra)
    xin = torch.randn(B, *spec.dims_in, generator=g)
    run_node_flip_free(blk, one, xin, cuda_dev, True, seed=spec.idx)
