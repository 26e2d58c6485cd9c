"""Run a few representative hot-path kernel calls in isolation (for ncu / timing on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibinn_b200 import ops
from ibinn_b200.ops import BN

dev = torch.device('cuda:0')
B = 128
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def bnref(C):
    return BN(torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev), 1e-4, False)


def conv_case(Cin, Cout, H, ks, pre, epi):
    x = torch.randn(B, Cin, H, H, device=dev)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) * 0.1
    kw = {}
    if pre == ops.PRE_BN_RELU:
        kw.update(pre=pre, pre_bn=bnref(Cin))
    if epi == ops.EPI_BNSTATS:
        kw.update(epi=epi, stats=dict(save_mean=torch.empty(Cout, device=dev), save_invstd=torch.empty(Cout, device=dev), momentum=0.999, eps=1e-4))
    return lambda: ops.conv_forward(x, w, ks, 1, **kw)


def wgrad_case(Cin, Cout, H, ks):
    g = torch.randn(B, Cout, H, H, device=dev)
    x = torch.randn(B, Cin, H, H, device=dev)
    dw = torch.zeros(Cout, Cin, ks, ks, device=dev)
    return lambda: ops.conv_wgrad(g, x, dw, ks, 1, x_pre=ops.PRE_RELU)


cases = {
    'conv3 64->64 @8 bn_relu/bnstats': conv_case(64, 64, 8, 3, ops.PRE_BN_RELU, ops.EPI_BNSTATS),
    'conv3 32->32 @16 bn_relu/bnstats': conv_case(32, 32, 16, 3, ops.PRE_BN_RELU, ops.EPI_BNSTATS),
    'conv3 16->16 @32 bn_relu/bnstats': conv_case(16, 16, 32, 3, ops.PRE_BN_RELU, ops.EPI_BNSTATS),
    'conv1 64->48 @8 bn_relu/store': conv_case(64, 48, 8, 1, ops.PRE_BN_RELU, ops.EPI_STORE),
    'conv3 24->64 @8 relu/bnstats': conv_case(24, 64, 8, 3, ops.PRE_NONE, ops.EPI_BNSTATS),
    'wgrad3 64->64 @8': wgrad_case(64, 64, 8, 3),
    'wgrad3 32->32 @16': wgrad_case(32, 32, 16, 3),
    'wgrad3 16->16 @32': wgrad_case(16, 16, 32, 3),
}
for name, fn in cases.items():
    fn()
torch.cuda.synchronize()
for name, fn in cases.items():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print('%-36s %8.2f us per call (graph of %d)' % (name, 1000 * e0.elapsed_time(e1) / reps, reps), flush=True)
