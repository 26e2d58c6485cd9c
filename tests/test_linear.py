"""FC-subnet GEMMs (y = x W^T + b, dx = g W, dW = g^T x with the ReLU / dropout pre-ops) against fp64.
On the GPU these run on tcgen05 with 3xTF32 operands: the gate is fp32-grade accuracy (1e-5 of abs-max;
single-pass TF32 would sit at ~1e-3)."""
import pytest
import torch

from util import rel

from ibinn_b200 import ops

SHAPES = [(128, 1536, 1024), (128, 1024, 3072), (5, 392, 1024), (64, 1024, 784), (3, 48, 8), (130, 100, 70)]


@pytest.mark.parametrize('B,K,N', SHAPES)
def test_linear_forward_dgrad_wgrad(dev, B, K, N):
    if dev.type == 'cpu' and B * K * N > 3e7:
        pytest.skip('full-size FC GEMM: B200 only (the simulator covers the small shapes)')
    g = torch.Generator().manual_seed(B + K + N)
    x = torch.randn(B, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.1
    b = torch.randn(N, generator=g)
    mask = (torch.rand(B, K, generator=g) > 0.3).float() * 1.5
    gy = torch.randn(B, N, generator=g)
    xd, wd, bd, md, gd = (t.to(dev) for t in (x, w, b, mask, gy))
    # forward, plain and with relu*mask on the input
    y = ops.linear_forward(xd, wd, bd)
    assert rel(y, x.double() @ w.double().t() + b.double()) < 1e-5
    y2 = ops.linear_forward(xd, wd, bd, x_pre=1, x_mask=md)
    xa = torch.relu(x.double()) * mask.double()
    assert rel(y2, xa @ w.double().t() + b.double()) < 1e-5
    # dgrad: plain (fresh output) and masked, accumulating into a strided slice
    dx = torch.empty(B, K, device=dev)
    ops.linear_dgrad(gd, wd, dx, accumulate=False)
    assert rel(dx, gy.double() @ w.double()) < 1e-5
    h = torch.randn(B, N, generator=g)
    m2 = (torch.rand(B, N, generator=g) > 0.5).float() * 2
    big = torch.randn(B, K + 8, generator=g)
    bigd = big.clone().to(dev)
    ops.linear_dgrad(gd, wd, bigd[:, 4:4 + K], accumulate=True, g_pre=2, h=h.to(dev), g_mask=m2.to(dev))
    gm = gy.double() * (h.double() > 0) * m2.double()
    ref = big.double().clone()
    ref[:, 4:4 + K] += gm @ w.double()
    assert rel(bigd, ref) < 1e-5
    # wgrad: both pre-op placements, accumulating
    dw = torch.ones(N, K, device=dev)
    db = torch.ones(N, device=dev)
    ops.linear_wgrad(gd, xd, dw, db, x_pre=1, x_mask=md)
    assert rel(dw, 1 + gy.double().t() @ xa) < 1e-5
    assert rel(db, 1 + gy.double().sum(0)) < 1e-5
    dw2 = torch.zeros(N, K, device=dev)
    db2 = torch.zeros(N, device=dev)
    ops.linear_wgrad(gd, bigd[:, 4:4 + K], dw2, db2, g_pre=2, h=h.to(dev), g_mask=m2.to(dev))
    assert rel(dw2, gm.t() @ bigd[:, 4:4 + K].double().cpu()) < 1e-5
    assert rel(db2, gm.sum(0)) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('B,K,N', [(128, 1536, 1024), (16, 1024, 3072)])
def test_split_k_is_bit_reproducible(cuda_dev, B, K, N):
    """small-batch GEMMs are split along K; the partial tiles are summed in split order by the last CTA of a tile (workspace +
    tickets), not with fp32 atomics: repeated launches give identical bits"""
    dev = cuda_dev
    g = torch.Generator().manual_seed(7)
    x, w, b = torch.randn(B, K, generator=g).to(dev), (torch.randn(N, K, generator=g) * 0.1).to(dev), torch.randn(N, generator=g).to(dev)
    gy = torch.randn(B, N, generator=g).to(dev)
    from ibinn_b200 import _lib
    assert _lib.lib().ibinn_linear_workspace_floats(B, N, K) > 0          # this shape does split
    y0 = ops.linear_forward(x, w, b)
    dx0 = torch.empty(B, K, device=dev)
    ops.linear_dgrad(gy, w, dx0, accumulate=False)
    for _ in range(8):
        assert torch.equal(ops.linear_forward(x, w, b), y0)
        dx = torch.empty(B, K, device=dev)
        ops.linear_dgrad(gy, w, dx, accumulate=False)
        assert torch.equal(dx, dx0)


@pytest.mark.gpu
def test_linear_without_workspace_uses_atomics(cuda_dev):
    """the C ABI keeps its workspace-free mode (fp32 atomics on the result, zeroed by the library): same values to fp32 accuracy"""
    from ibinn_b200 import _lib
    dev = cuda_dev
    B, K, N = 128, 1536, 1024
    g = torch.Generator().manual_seed(9)
    x, w, b = torch.randn(B, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    y = torch.full((B, N), 7.0, device=dev)
    _lib.check(_lib.lib().ibinn_linear_forward(_lib.ptr(xd), K, None, 0, _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(y), B, K, N, None, None,
                                               _lib.stream(xd)), 'linear_forward')
    assert rel(y, x.double() @ w.double().t() + b.double()) < 1e-5
    assert torch.equal(ops.linear_forward(xd, wd, bd) - y < 1e-3, torch.ones_like(y, dtype=torch.bool))
