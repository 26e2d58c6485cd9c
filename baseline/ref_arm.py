"""The reference arm of bench.py: the UNMODIFIED reference files, copied by `install()` (run by __graft_entry__.build() in the
build container, where /root/reference exists) into baseline/_ref/ -- git-ignored, shipped to the GPU box with the snapshot --
and driven through their own public API (`model.GenerativeClassifier`, the loop body of reference train.py:98-111).
MEASUREMENT INFRASTRUCTURE ONLY: nothing under ibinn_b200/ imports this.

The reference is not a package (no setup.py / pyproject.toml: `pip install /root/reference` has nothing to build) and needs
four shims on this image (oracle/ref_shims): torch.rfft/irfft aliases, a FrEIA stand-in (the pinned
FrEIA commit is not vendored), a matplotlib stub, and -- for the CPU arm -- `.cuda()` neutralised.
"""
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, '_ref')
FILES = ['model.py', 'inn_architecture.py', 'all_in_one_block.py', 'downsampling_coupling_block.py', 'dct_transform.py',
         'feed_forward_architecture.py', 'data.py', 'default.ini']


def install(src='/root/reference'):
    """copy the reference's own files (verbatim) next to this module; returns True if baseline/_ref is complete"""
    if os.path.isfile(os.path.join(src, 'model.py')):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    return available()


def available():
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in FILES)


def load(device):
    """the reference's `model` module, importable on this image; device = 'cpu' | 'cuda'"""
    os.environ['IBINN_REFERENCE_DIR'] = REF_DIR
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shims'))
    import load_reference as lr
    lr.REFERENCE_DIR = REF_DIR
    ref = lr.load(force_cpu=(device == 'cpu'))
    return ref, lr


WORKLOADS = {
    # name: (ini overrides, per-GPU batch, kind)
    'cifar10_train': ({}, 128, 'train'),
    'cifar100_train': ({'data': {'dataset': 'CIFAR100'}}, 128, 'train'),
    'cifar10_beta_sweep': ({}, 128, 'train'),
    'mnist_fwd': ({'data': {'dataset': 'MNIST', 'label_smoothing': '0.01'}, 'model': {'act_norm': '0.80'}}, 64, 'fwd'),
    'cifar10_infer': ({}, 2048, 'infer'),
}


class RefRunner:
    """one workload of bench.py on the reference's own code path"""

    def __init__(self, workload, device='cpu', batch=None, beta=1.0, threads=None, seed=0):
        import numpy as np
        import torch
        self.torch = torch
        ov, b, self.kind = WORKLOADS[workload]
        self.batch = batch or b
        self.device = device
        if device == 'cpu':
            self.threads = threads or os.cpu_count()
            torch.set_num_threads(self.threads)
        else:
            self.threads = None
            torch.backends.cuda.matmul.allow_tf32 = False       # fp32 like the reference (its torch <= 1.7 had no TF32 default)
            torch.backends.cudnn.allow_tf32 = False
        ref, lr = load(device)
        args = lr.make_args(ov)
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.model = ref.GenerativeClassifier(args)
        if device == 'cuda':
            self.model.cuda()
        C = self.model.n_classes
        g = torch.Generator().manual_seed(1)
        x = torch.randn(self.batch, *self.model.dims, generator=g)
        lab = torch.randint(0, C, (self.batch,), generator=g)
        eps = float(args['data']['label_smoothing'])
        y = torch.zeros(self.batch, C).scatter_(1, lab.view(-1, 1), 1.) * (1 - eps) + eps / C
        self.x, self.y = (x.cuda(), y.cuda()) if device == 'cuda' else (x, y)
        self.bx, self.by = 2. / (1 + beta), 2. * beta / (1 + beta)        # train.py:88-92
        self.clip = float(args['training']['clip_grad_norm'])
        if self.kind == 'infer':
            self.model.eval()

    def step(self):
        torch, m = self.torch, self.model
        if self.kind == 'train':                                 # reference train.py:100-111
            losses = m(self.x, self.y)
            loss = self.bx * losses['L_x_tr'] - self.by * losses['L_y_tr']
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.trainable_params, self.clip)
            m.optimizer.step()
            m.optimizer.zero_grad()
            return float(loss.detach())
        if self.kind == 'fwd':                                   # forward + IB loss (BASELINE.json configs[0])
            with torch.no_grad():
                losses = m(self.x, self.y)
            return float(losses['L_x_tr'])
        with torch.no_grad():                                    # reference evaluation/ood.py:43 (eval mode, per-sample scores)
            losses = m(self.x, y=None, loss_mean=False)
            scores = losses['L_x_tr'].cpu()
        return float(scores[0])

    def time(self, steps, warmup, budget_s=None):
        """seconds per step; with budget_s the number of steps is clipped so that the run ends within about that time"""
        torch = self.torch
        sync = torch.cuda.synchronize if self.device == 'cuda' else (lambda: None)
        t0 = time.perf_counter()
        self.step()
        sync()
        first = time.perf_counter() - t0
        if budget_s is not None:
            steps = max(2, min(steps, int(budget_s / max(first, 1e-3))))
            warmup = min(warmup, 1)
        for _ in range(max(warmup - 1, 0)):
            self.step()
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        sync()
        return (time.perf_counter() - t0) / steps, steps
