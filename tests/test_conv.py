"""Subnet convolution kernel (ibinn_conv_forward / ibinn_conv_wgrad) against torch in fp64: plain,
with every pre-op / epilogue, dgrad (transposed weights, zero-insert for stride 2) and wgrad.
On a B200 stride-1 cases run on the tcgen05 3xTF32 path, the rest on the CUDA-core kernel; the
simulator build runs the CUDA-core kernel only."""
import pytest
import torch
import torch.nn.functional as F

from util import rel

from ibinn_b200 import ops
from ibinn_b200.ops import BN

CASES = [  # B, Cin, Cout, H, ks, stride
    (3, 2, 16, 32, 3, 1), (2, 16, 16, 32, 3, 1), (3, 6, 32, 16, 3, 1), (2, 32, 32, 16, 3, 1), (3, 24, 64, 8, 3, 1),
    (5, 64, 64, 8, 3, 1), (2, 32, 32, 32, 3, 2), (2, 64, 64, 14, 3, 2), (3, 8, 64, 7, 3, 1), (3, 64, 48, 8, 1, 1),
    (2, 16, 2, 32, 1, 1), (2, 32, 12, 16, 1, 1), (2, 64, 128, 8, 1, 1), (4, 64, 16, 7, 1, 1), (130, 64, 64, 8, 3, 1),
]
TOL = 2e-5


@pytest.mark.parametrize('B,Cin,Cout,H,ks,st', CASES)
def test_conv_plain_bias_addend(dev, B, Cin, Cout, H, ks, st):
    if dev.type == 'cpu' and B > 100:
        pytest.skip('multi-wave case: B200 only')
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn(B, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=st, padding=ks // 2)
    y = ops.conv_forward(x.to(dev), w.to(dev), ks, st, bias=b.to(dev))
    assert rel(y, ref) < TOL
    add = torch.randn(ref.shape, generator=g)
    y2 = ops.conv_forward(x.to(dev), w.to(dev), ks, st, addend=add.to(dev))
    assert rel(y2, ref - b.double().view(1, -1, 1, 1) + add.double()) < TOL


@pytest.mark.parametrize('B,Cin,Cout,H,ks,st', [c for c in CASES if c[2] <= 64])
def test_conv_bn_pipeline(dev, B, Cin, Cout, H, ks, st):
    if dev.type == 'cpu' and B > 100:
        pytest.skip('multi-wave case: B200 only')
    """pre = relu(bn(.))*dropmask on a channel slice, epilogue = batch statistics + running-stat update."""
    g = torch.Generator().manual_seed(Cin + 3 * Cout)
    xx = torch.randn(B, Cin + 3, H, H, generator=g)
    x = xx[:, 2:2 + Cin]
    w = torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2
    mean, var = torch.randn(Cin, generator=g) * 0.3, torch.rand(Cin, generator=g) + 0.5
    gam, bet = torch.randn(Cin, generator=g), torch.randn(Cin, generator=g) * 0.2
    mask = (torch.rand(B, Cin, generator=g) > 0.25).float() / 0.75
    xa = F.relu(F.batch_norm(x.double(), mean.double(), var.double(), gam.double(), bet.double(), False, 0., 1e-4)) * mask.double().view(B, Cin, 1, 1)
    ref = F.conv2d(xa, w.double(), stride=st, padding=ks // 2)
    d = lambda t: t.to(dev)
    sm, si = torch.empty(Cout, device=dev), torch.empty(Cout, device=dev)
    rm, rv = torch.randn(Cout, generator=g), torch.rand(Cout, generator=g) + 0.5
    rmd, rvd = d(rm.clone()), d(rv.clone())
    nbt = torch.tensor(3, dtype=torch.int64, device=dev)
    y = ops.conv_forward(d(xx)[:, 2:2 + Cin], d(w), ks, st, pre=ops.PRE_BN_RELU, pre_bn=BN(d(mean), d(var), d(gam), d(bet), 1e-4, True),
                         pre_mask=d(mask), epi=ops.EPI_BNSTATS,
                         stats=dict(save_mean=sm, save_invstd=si, running_mean=rmd, running_var=rvd, nbt=nbt, momentum=0.999, eps=1e-4))
    assert rel(y, ref) < TOL
    bm, bv = ref.mean(dim=(0, 2, 3)), ref.var(dim=(0, 2, 3), unbiased=False)
    n = ref.numel() // Cout
    assert rel(sm, bm) < 1e-5 and rel(si, torch.rsqrt(bv + 1e-4)) < 1e-5
    assert rel(rmd, 0.001 * rm.double() + 0.999 * bm) < 1e-5
    assert rel(rvd, 0.001 * rv.double() + 0.999 * bv * n / (n - 1)) < 1e-5
    assert int(nbt) == 4
    # cumulative-average mode (momentum=None), as evaluation/__init__.py:18-53 uses it
    rmd2 = d(rm.clone())
    rvd2 = d(rv.clone())
    ops.conv_forward(d(x.contiguous()), d(w), ks, st, epi=ops.EPI_BNSTATS,
                     stats=dict(save_mean=sm, save_invstd=si, running_mean=rmd2, running_var=rvd2, nbt=nbt, momentum=None, eps=1e-4))
    ref2 = F.conv2d(x.double(), w.double(), stride=st, padding=ks // 2)
    assert rel(rmd2, 0.8 * rm.double() + 0.2 * ref2.mean(dim=(0, 2, 3))) < 1e-5 and int(nbt) == 5


@pytest.mark.parametrize('B,Cin,Cout,H,ks,st', [c for c in CASES if c[1] <= 64])
def test_conv_backward_pieces(dev, B, Cin, Cout, H, ks, st):
    if dev.type == 'cpu' and B > 100:
        pytest.skip('multi-wave case: B200 only')
    """The three backward uses: dgrad with BN-backward pre-op + mask/statistics epilogue, dgrad with
    ReLU mask + in-place accumulate, and wgrad -- against autograd of conv(relu(bn_train(h)))."""
    g = torch.Generator().manual_seed(11 * Cin + Cout)
    d = lambda t: t.to(dev)
    Ho = (H + 2 * (ks // 2) - ks) // st + 1
    # forward graph in fp64: u --bn_u,relu--> a --conv--> h --bn_h (batch stats)--> ...  with upstream grad on bn_h output
    u = torch.randn(B, Cin, H, H, generator=g).double().requires_grad_()
    gu, bu = (torch.randn(Cin, generator=g).double() + 1.5).requires_grad_(), torch.randn(Cin, generator=g).double().requires_grad_()
    w = (torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2).double().requires_grad_()
    gh_, bh_ = (torch.randn(Cout, generator=g).double() + 1.5).requires_grad_(), torch.randn(Cout, generator=g).double().requires_grad_()
    mu_u, var_u = u.mean(dim=(0, 2, 3)), u.var(dim=(0, 2, 3), unbiased=False)
    xhat_u = (u - mu_u.view(1, -1, 1, 1)) * torch.rsqrt(var_u + 1e-4).view(1, -1, 1, 1)
    zu = xhat_u * gu.view(1, -1, 1, 1) + bu.view(1, -1, 1, 1)
    a = F.relu(zu)
    h = F.conv2d(a, w, stride=st, padding=ks // 2)
    mu_h, var_h = h.mean(dim=(0, 2, 3)), h.var(dim=(0, 2, 3), unbiased=False)
    zh = (h - mu_h.view(1, -1, 1, 1)) * torch.rsqrt(var_h + 1e-4).view(1, -1, 1, 1) * gh_.view(1, -1, 1, 1) + bh_.view(1, -1, 1, 1)
    gz = torch.randn(zh.shape, generator=g).double()          # gradient w.r.t. bn_h output (already masked upstream)
    zu.retain_grad()
    (zh * gz).sum().backward()
    # ---- ours
    sums_h = torch.stack([gz.sum(dim=(0, 2, 3)), (gz * ((h - mu_h.view(1, -1, 1, 1)) * torch.rsqrt(var_h + 1e-4).view(1, -1, 1, 1))).sum(dim=(0, 2, 3))]).float()
    bn_h = BN(d(mu_h.detach().float()), d(torch.rsqrt(var_h + 1e-4).detach().float()), d(gh_.detach().float()), d(bh_.detach().float()), 1e-4, False)
    bn_u = BN(d(mu_u.detach().float()), d(torch.rsqrt(var_u + 1e-4).detach().float()), d(gu.detach().float()), d(bu.detach().float()), 1e-4, False)
    wd, hd, ud, gzd = d(w.detach().float()), d(h.detach().float()), d(u.detach().float()), d(gz.float())
    dw = torch.zeros(Cout, Cin, ks, ks, device=dev)
    ops.conv_wgrad(gzd, ud, dw, ks, st, g_pre=ops.PRE_BNBWD, gh=hd, g_bn=bn_h, g_sums=d(sums_h), x_pre=ops.PRE_BN_RELU, x_bn=bn_u)
    assert rel(dw, w.grad) < 5e-5
    sums_u = torch.empty(2, Cin, device=dev)
    dgam, dbet = torch.zeros(Cin, device=dev), torch.zeros(Cin, device=dev)
    gzu = ops.conv_forward(gzd, wd, ks, 1, cout=Cin, w_transposed=True, pre=ops.PRE_BNBWD, x2=hd, pre_bn=bn_h, pre_sums=d(sums_h),
                           upsample_to=(H, H) if st == 2 else None, epi=ops.EPI_MASKSTATS,
                           mask=dict(h=ud, bn=bn_u, sums_out=sums_u, dgamma=dgam, dbeta=dbet))
    assert rel(gzu, zu.grad) < 5e-5
    assert rel(dgam, gu.grad) < 5e-5 and rel(dbet, bu.grad) < 5e-5
    assert rel(sums_u[0], bu.grad) < 5e-5
    # relu-mask epilogue with in-place accumulation: y += dgrad * [u > 0]
    acc0 = torch.randn(B, Cin, H, H, generator=g)
    accd = d(acc0.clone())
    ops.conv_forward(gzd, wd, ks, 1, cout=Cin, w_transposed=True, upsample_to=(H, H) if st == 2 else None, epi=ops.EPI_RELUMASK,
                     mask=dict(h=ud), addend=accd, out=accd)
    full = torch.nn.grad.conv2d_input((B, Cin, H, H), w.detach(), gz, stride=st, padding=ks // 2)
    assert rel(accd, acc0.double() + full * (u.detach() > 0)) < 5e-5


@pytest.mark.parametrize('B,Cin,Cout,H,ks', [(3, 16, 16, 16, 3), (2, 32, 12, 16, 1), (3, 24, 40, 8, 3), (2, 64, 48, 8, 1)])
def test_conv_wgrad_with_bias(dev, B, Cin, Cout, H, ks):
    """plain wgrad + dbias (the 1x1 convs of the subnets carry a bias; for 3x3 the bias gradient must come from the
    unshifted copy of the gradient operand only)."""
    g = torch.Generator().manual_seed(5 * Cin + Cout + ks)
    x = torch.randn(B, Cin, H, H, generator=g).double()
    w = (torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2).double().requires_grad_()
    b = torch.randn(Cout, generator=g).double().requires_grad_()
    gy = torch.randn(B, Cout, H, H, generator=g).double()
    (F.conv2d(F.relu(x), w, b, padding=ks // 2) * gy).sum().backward()
    dw0, db0 = torch.randn(Cout, Cin, ks, ks, generator=g), torch.randn(Cout, generator=g)
    dw, db = dw0.clone().to(dev), db0.clone().to(dev)
    ops.conv_wgrad(gy.float().to(dev), x.float().to(dev), dw, ks, 1, x_pre=ops.PRE_RELU, dbias=db)      # accumulates
    assert rel(dw, dw0.double() + w.grad) < 5e-5
    assert rel(db, db0.double() + b.grad) < 5e-5


def test_bn_statistics_with_large_mean(dev):
    """batch statistics of channels whose mean dwarfs their spread (mean 50, std 0.1): the per-tile, per-CTA and
    last-CTA sums are all taken about a nearby shift, so fp32 is enough."""
    g = torch.Generator().manual_seed(3)
    B, C, H = 6, 16, 16
    x = (50. + 5. * torch.arange(C).view(1, C, 1, 1) + 0.1 * torch.randn(B, C, H, H, generator=g)).float()
    w = torch.eye(C).view(C, C, 1, 1)
    sm, si = torch.empty(C, device=dev), torch.empty(C, device=dev)
    y = ops.conv_forward(x.to(dev), w.to(dev), 1, 1, epi=ops.EPI_BNSTATS, stats=dict(save_mean=sm, save_invstd=si, momentum=0.999, eps=1e-4))
    assert rel(y, x.double()) < 1e-6
    xd = x.double()
    assert rel(sm, xd.mean(dim=(0, 2, 3))) < 1e-6
    assert rel(si, torch.rsqrt(xd.var(dim=(0, 2, 3), unbiased=False) + 1e-4)) < 2e-4


@pytest.mark.parametrize('B,Cin,Cout,H,drop', [(3, 32, 12, 16, False), (2, 64, 48, 8, True), (2, 16, 2, 32, False), (130, 64, 48, 8, True), (5, 32, 4, 14, False)])
def test_conv1x1_backward_fused(dev, monkeypatch, B, Cin, Cout, H, drop):
    """ibinn_conv1x1_backward (input gradient with mask + BatchNorm-backward sums AND weight / bias gradient records in one launch)
    against autograd of conv1x1(relu(bn_train(h)) * dropmask) in fp64."""
    import ctypes
    from ibinn_b200 import _lib
    if dev.type == 'cpu' and B > 100:
        pytest.skip('large case: B200 only')
    g = torch.Generator().manual_seed(17 * Cin + Cout)
    d = lambda t: t.to(dev)
    hh = torch.randn(B, Cin + 2, H, H, generator=g)
    h = hh[:, 1:1 + Cin].double().requires_grad_()
    gam, bet = (torch.randn(Cin, generator=g).double() + 1.5).requires_grad_(), torch.randn(Cin, generator=g).double().requires_grad_()
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) * 0.3).double().requires_grad_()
    bias = torch.zeros(Cout).double().requires_grad_()
    mask = ((torch.rand(B, Cin, generator=g) > 0.25).float() / 0.75) if drop else None
    mu, var = h.mean(dim=(0, 2, 3)), h.var(dim=(0, 2, 3), unbiased=False)
    xhat = (h - mu.view(1, -1, 1, 1)) * torch.rsqrt(var + 1e-4).view(1, -1, 1, 1)
    z = xhat * gam.view(1, -1, 1, 1) + bet.view(1, -1, 1, 1)
    z.retain_grad()
    act = F.relu(z)
    if drop:
        act = act * mask.double().view(B, Cin, 1, 1)
    out = F.conv2d(act, w, bias)
    gy = torch.randn(out.shape, generator=g)
    (out * gy.double()).sum().backward()
    # ---- ours, straight through the C ABI (the record reduction is done here with torch so that the simulator can run it too)
    bn = BN(d(mu.detach().float()), d(torch.rsqrt(var + 1e-4).detach().float()), d(gam.detach().float()), d(bet.detach().float()), 1e-4, False)
    a = _lib.Conv1x1BwdArgs()
    a.B, a.Cin, a.Cout, a.H, a.W = B, Cin, Cout, H, H
    L = _lib.lib()
    ncta, rec = ctypes.c_int32(), ctypes.c_int32()
    assert L.ibinn_conv1x1_bwd_geometry(ctypes.byref(a), ctypes.byref(ncta), ctypes.byref(rec)) == 0
    gd, hd = d(gy.float()), d(hh.float())[:, 1:1 + Cin]
    wd, md = d(w.detach().float().reshape(Cout, Cin).contiguous()), (d(mask) if drop else None)
    g_in = torch.empty(B, Cin, H, H, device=dev)
    sums, dgam, dbet = torch.empty(2, Cin, device=dev), torch.zeros(Cin, device=dev), torch.zeros(Cin, device=dev)
    part = torch.empty(ncta.value * 2 * Cin, device=dev)
    wrec = torch.zeros(ncta.value, rec.value, device=dev)
    a.g = _lib.ptr(gd)
    a.h, a.h_bs = ops._nchw(hd)
    a.bn = bn.ref()
    a.drop, a.w = _lib.ptr(md), _lib.ptr(wd)
    a.g_in, a.sums_out, a.dgamma, a.dbeta = _lib.ptr(g_in), _lib.ptr(sums), _lib.ptr(dgam), _lib.ptr(dbet)
    a.partials, a.counter, a.wrecords = _lib.ptr(part), ops.counter_ptr(gd), _lib.ptr(wrec)
    _lib.check(L.ibinn_conv1x1_backward(ctypes.byref(a), _lib.stream(gd)), 'conv1x1_backward')
    assert rel(g_in, z.grad) < 2e-5
    assert rel(sums[0], bet.grad) < 2e-5 and rel(sums[1], gam.grad) < 2e-5
    assert rel(dbet, bet.grad) < 2e-5 and rel(dgam, gam.grad) < 2e-5
    tot = wrec.double().sum(0)
    assert rel(tot[:Cout * Cin].view(Cout, Cin), w.grad.view(Cout, Cin)) < 2e-5
    assert rel(tot[Cout * Cin:Cout * Cin + Cout], bias.grad) < 2e-5
    if dev.type == 'cuda':
        # and through ops.conv1x1_backward (records summed by ibinn_wgrad_reduce_many, accumulated into dw / dbias)
        dw0, db0 = torch.randn(Cout, Cin, 1, 1, generator=g), torch.randn(Cout, generator=g)
        dw, db = d(dw0.clone()), d(db0.clone())
        s2 = torch.empty(2, Cin, device=dev)
        monkeypatch.setattr(ops, 'FUSED_1X1_BWD', True)        # (opt-in: IBINN_FUSED_1X1_BWD=1)
        g2 = ops.conv1x1_backward(gd, hd, bn, md, d(w.detach().float()), s2, None, None, dw, db)
        assert g2 is not None and torch.equal(g2, g_in)
        assert rel(dw, dw0.double() + w.grad) < 2e-5 and rel(db, db0.double() + bias.grad) < 2e-5
