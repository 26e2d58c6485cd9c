import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibinn_b200 import ops
from ibinn_b200.ops import BN
dev = torch.device('cuda:0'); B = 128
Cin, Cout, H = [int(v) for v in sys.argv[1:4]]
x = torch.randn(B, Cin, H, H, device=dev); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.1
bn = BN(torch.zeros(Cin, device=dev), torch.ones(Cin, device=dev), torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev), 1e-4, False)
st = dict(save_mean=torch.empty(Cout, device=dev), save_invstd=torch.empty(Cout, device=dev), momentum=0.999, eps=1e-4)
for _ in range(3):
    ops.conv_forward(x, w, 3, 1, pre=ops.PRE_BN_RELU, pre_bn=bn, epi=ops.EPI_BNSTATS, stats=st)
torch.cuda.synchronize()
print('ok')
