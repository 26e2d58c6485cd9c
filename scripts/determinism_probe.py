"""Is the training step bit-reproducible?  Two engines from the same seed, the same three batches (CUDA graph, dropout on);
prints which parameter tensors differ bitwise after the steps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ibinn_b200 import config                      # noqa: E402
from ibinn_b200.engine import TrainStep            # noqa: E402
from ibinn_b200.model import GenerativeClassifier  # noqa: E402

dev = torch.device('cuda:0')
B = 128
gen = torch.Generator().manual_seed(5)
batches = []
for _ in range(3):
    x = torch.randn(B, 3, 32, 32, generator=gen)
    lab = torch.randint(0, 10, (B,), generator=gen)
    batches.append((x.to(dev), (torch.zeros(B, 10).scatter_(1, lab.view(-1, 1), 1.) * 0.98 + 0.002).to(dev)))
res = []
for rep in range(2):
    args = config.make_args(overrides={})
    torch.manual_seed(0)
    np.random.seed(0)
    torch.cuda.manual_seed(0)
    m = GenerativeClassifier(args).to(dev)
    step = TrainStep(m, batch_size=B, beta=1.0, use_graph=True)
    outs = []
    for x, y in batches:
        o = step.step(x, y)
        outs.append({k: float(v) for k, v in o.items() if torch.is_tensor(v) and v.numel() == 1})
    torch.cuda.synchronize()
    res.append(({k: v.detach().clone() for k, v in m.state_dict().items()}, outs))
print('losses run 0:', res[0][1][-1])
print('losses run 1:', res[1][1][-1])
bad = [k for k in res[0][0] if not torch.equal(res[0][0][k], res[1][0][k])]
print('%d of %d state tensors differ bitwise' % (len(bad), len(res[0][0])))
for k in bad[:12]:
    a, b = res[0][0][k].float(), res[1][0][k].float()
    print('  ', k, float((a - b).abs().max() / (a.abs().max() + 1e-30)))
