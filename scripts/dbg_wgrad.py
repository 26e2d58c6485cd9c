import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ibinn_b200 import ops, _lib
dev = torch.device('cuda:0'); B = 128
def run(Cin, Cout, H, ks):
    g = torch.randn(B, Cout, H, H, device=dev); x = torch.randn(B, Cin, H, H, device=dev); dw = torch.zeros(Cout, Cin, ks, ks, device=dev)
    for _ in range(3): ops.conv_wgrad(g, x, dw, ks, 1, x_pre=ops.PRE_RELU)
    torch.cuda.synchronize()
    L = ctypes.CDLL(_lib.LIB_PATH); buf = (ctypes.c_ulonglong*64)(); L.ibinn_wdebug_read(buf)
    t0 = buf[0]; r = lambda i: (buf[i]-t0)/1000.
    print('wgrad%d %d->%d @%d: setup %.2f | mma-done %.2f epi %.2f-%.2f end %.2f' % (ks, Cin, Cout, H, r(1), r(2), r(3), r(4), r(5)))
    print('   mma block starts:', ' '.join('%.2f' % r(10+i) for i in range(6)))
    print('   block0 detail: start %.2f | B loads issued %.2f | A loads issued %.2f | A stored %.2f | B stored %.2f | fence %.2f | arrive %.2f' % (r(20), r(40), r(41), r(42), r(43), r(44), r(21)))
    print('   loader blocks   :', ' '.join('%.2f-%.2f' % (r(20+2*i), r(21+2*i)) for i in range(6)))
run(64,64,8,3); run(32,32,16,3); run(16,16,32,3)
