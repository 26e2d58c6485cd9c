import sys, os, copy, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_engine import _model, _batch
from util import rel
from ibinn_b200.engine import TrainStep
dev = torch.device('cuda:0')
base = _model(dev, overrides={'model': {'n_coupling_blocks_conv': '[1, 2, 2]', 'fc_width': '16'}})
gen = torch.Generator().manual_seed(5)
batches = [_batch(8, 10, gen) for _ in range(2)]
def run(streams, group):
    os.environ['IBINN_WGRAD_STREAMS'] = streams
    os.environ['IBINN_WGRAD_REDUCE_GROUP'] = group
    m = copy.deepcopy(base); m.eval()
    step = TrainStep(m, batch_size=8, beta=1.0, clip=8., use_graph=False)
    for x, y in batches: step.step(x.to(dev), y.to(dev))
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in m.state_dict().items() if v.is_floating_point()}
ref = run('0', '0')
for name, (s, g) in {'same schedule again': ('0', '0'), 'again': ('0', '0'), 'side stream': ('1', '0'), 'side stream again': ('1', '0'), 'grouped': ('1', '3')}.items():
    r = run(s, g)
    worst = max(((rel(r[k], ref[k].double()), k) for k in ref), key=lambda t: t[0])
    tot = torch.cat([(r[k] - ref[k]).flatten() for k in ref]).norm() / torch.cat([ref[k].flatten() for k in ref]).norm()
    print('%-22s worst per-tensor rel %.2e (%s)   global rel %.2e' % (name, worst[0], worst[1], float(tot)))
