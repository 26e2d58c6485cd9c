import sys, os, ctypes, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ibinn_b200 import ops, _lib
from ibinn_b200.ops import BN
dev = torch.device('cuda:0'); B = 128
def run(Cin, Cout, H, ks):
    x = torch.randn(B, Cin, H, H, device=dev); w = torch.randn(Cout, Cin, ks, ks, device=dev)*0.1
    bn = BN(torch.zeros(Cin, device=dev), torch.ones(Cin, device=dev), torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev), 1e-4, False)
    st = dict(save_mean=torch.empty(Cout, device=dev), save_invstd=torch.empty(Cout, device=dev), momentum=0.999, eps=1e-4)
    for _ in range(3): ops.conv_forward(x, w, ks, 1, pre=ops.PRE_BN_RELU, pre_bn=bn, epi=ops.EPI_BNSTATS, stats=st)
    torch.cuda.synchronize()
    L = ctypes.CDLL(_lib.LIB_PATH); buf = (ctypes.c_ulonglong*64)(); L.ibinn_debug_read(buf)
    t0 = buf[0]; r = lambda i: (buf[i]-t0)/1000.
    print('conv%d %d->%d @%d: setup %.2f | end %.2f | last CTA reaches ticket %.2f, has ticket %.2f, records summed %.2f, combined %.2f, finalize-end %.2f' % (ks, Cin, Cout, H, r(1), r(2), r(5), r(4), r(6), r(7), r(3)))
    print('   tile 0 mma phase: %d cycles, %.2f us -> %.2f GHz' % (buf[51]-buf[50], r(11)-r(10), (buf[51]-buf[50])/1000./max(r(11)-r(10),1e-9)))
    print('   tile 0 epilogue: columns done %.2f, synced %.2f, stats done %.2f' % (r(60), r(61), r(62)))
    for j in range(2):
        print('   tile %d: load %.2f-%.2f  mma %.2f-%.2f  epi %.2f-%.2f' % (j, r(20+2*j), r(21+2*j), r(10+2*j), r(11+2*j), r(30+2*j), r(31+2*j)))
run(64,64,8,3); run(32,32,16,3); run(6,32,16,3); run(16,16,32,3)
