"""CPU baseline / reference arm of bench.py: the oracle's fp32 restatement of the reference's training
step (train.py:98-111: forward, IB loss, backward, clip_grad_norm_ 8.0, SGD) on the host cores.
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by ibinn_b200/."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ibinn_oracle as O  # noqa: E402


class CpuTrainer:
    def __init__(self, cfg: 'O.Cfg', batch, beta=1.0, seed=0, threads=None, lr=0.07):
        self.threads = threads or os.cpu_count()
        torch.set_num_threads(self.threads)
        self.cfg, self.spec = cfg, O.graph_spec(cfg)
        self.P = O.init_state(cfg, seed=seed, dtype=torch.float32)
        self.keys = O.trainable_keys(self.P)
        for k in self.keys:
            self.P[k].requires_grad_()
        inn = [k for k in self.keys if k not in ('mu', 'phi')]
        self.groups = [dict(keys=inn, lr=lr, momentum=0.9, weight_decay=1e-4),
                       dict(keys=['mu'], lr=lr * 3.0, momentum=0., weight_decay=0.),
                       dict(keys=['phi'], lr=lr, momentum=0.8, weight_decay=0.)]
        self.bufs = {}
        self.bx, self.by = O.beta_weights(beta)
        g = torch.Generator().manual_seed(1)
        self.x = torch.randn(batch, *cfg.dims, generator=g)
        lab = torch.randint(0, cfg.n_classes, (batch,), generator=g)
        eps = 0.02
        self.y = torch.zeros(batch, cfg.n_classes).scatter_(1, lab.view(-1, 1), 1.) * (1 - eps) + eps / cfg.n_classes
        self.batch = batch
        self.gen = torch.Generator().manual_seed(2)

    def _masks(self):
        m = {}
        for n in self.spec:
            p = n.extra.get('dropout', 0.)
            if n.kind == 'aio' and p > 0:
                j = 7 if n.extra['relu_first'] else 6
                m['%s.s.%d' % (n.key, j)] = (torch.rand(self.batch, n.extra['width'], generator=self.gen) >= p).float() / (1 - p)
            if n.kind == 'glow' and p > 0:
                for s in ('s1', 's2'):
                    m['%s.%s.2' % (n.key, s)] = (torch.rand(self.batch, n.extra['width'], generator=self.gen) >= p).float() / (1 - p)
        return m

    def step(self):
        out, z, jac, st = O.classifier_forward(self.spec, self.P, self.x, self.y, loss_mean=True, train=True, masks=self._masks())
        loss = self.bx * out['L_x_tr'] - self.by * out['L_y_tr']
        grads = torch.autograd.grad(loss, [self.P[k] for k in self.keys])
        with torch.no_grad():
            O.sgd_clip_step(self.P, dict(zip(self.keys, grads)), self.bufs, self.groups, 8.0)
            for k, v in st.new_stats.items():
                self.P[k].copy_(v)
        return float(loss.detach())

    def forward_only(self):
        with torch.no_grad():
            out, z, jac, st = O.classifier_forward(self.spec, self.P, self.x, self.y, loss_mean=True, train=True, masks=self._masks())
        return float(out['L_x_tr'])


def time_steps(trainer, steps, warmup, fn='step'):
    f = getattr(trainer, fn)
    for _ in range(warmup):
        f()
    t0 = time.perf_counter()
    for _ in range(steps):
        f()
    return (time.perf_counter() - t0) / steps
