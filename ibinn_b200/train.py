"""Training driver with the reference's semantics (reference train.py:14-168): same .ini keys, loss mix,
clip + SGD, MultiStepLR / exponential schedule, `losses.dat` log format, divergence guard, checkpoints.
The per-batch work (train.py:100-111) goes through engine.TrainStep (CUDA graphs, fused clip + SGD); the
per-batch `.item()` syncs of the reference (train.py:120-121) are replaced by one sync per `interval_log`.

`dataset` is any object with the reference data.Dataset surface that the loop touches: `train_loader`,
`val_x`, `val_y`, `onehot(labels, smoothing)`.  The reference's own `data.py` (torchvision loaders) is outside
the hot path; `SyntheticDataset` below stands in for it where no data is on disk."""
from os.path import join
from time import time

import numpy as np
import torch

from .engine import TrainStep
from .model import GenerativeClassifier

PLOT_COLUMNS = ['time', 'epoch', 'iteration', 'L_x_tr', 'L_x_val', 'L_y_tr', 'L_y_val', 'acc_tr', 'acc_val', 'delta_mu_val']


class SyntheticDataset:
    """N(0,1) images with random labels in the shapes of the configured dataset (throughput / smoke runs)."""

    def __init__(self, args, n_batches=8, device='cuda', seed=1):
        self.batch_size = eval(args['data']['batch_size'])
        ds = args['data']['dataset']
        pad = eval(args['data']['pad_noise_channels'])
        self.dims = (28, 28) if ds == 'MNIST' else (3 + pad, 32, 32)
        self.n_classes = 100 if ds == 'CIFAR100' else 10
        g = torch.Generator().manual_seed(seed)
        self.device = torch.device(device)
        self.train_loader = [(torch.randn(self.batch_size, *self.dims, generator=g),
                              torch.randint(0, self.n_classes, (self.batch_size,), generator=g)) for _ in range(n_batches)]
        self.val_x = torch.randn(self.batch_size, *self.dims, generator=g).to(self.device)
        self.val_y = self.onehot(torch.randint(0, self.n_classes, (self.batch_size,), generator=g).to(self.device),
                                 eval(args['data']['label_smoothing']))

    def onehot(self, l, label_smooth=0):
        y = torch.zeros(l.shape[0], self.n_classes, device=l.device)
        y.scatter_(1, l.view(-1, 1), 1.)
        if label_smooth:
            y = y * (1 - label_smooth) + label_smooth / self.n_classes
        return y


def train(args, dataset=None, device='cuda', use_graph=True, max_iterations=None, synthetic=False):
    t, c = args['training'], args['checkpoints']
    N_epochs = eval(t['n_epochs'])
    label_smoothing = eval(args['data']['label_smoothing'])
    interval_log = eval(c['interval_log'])
    interval_checkpoint = eval(c['interval_checkpoint'])
    save_on_crash = eval(c['checkpoint_when_crash'])
    output_dir, resume = c['output_dir'], c['resume_checkpoint']
    ensemble_index = eval(c['ensemble_index'])
    ensemble_str = '' if ensemble_index is None else '.%.2i' % ensemble_index
    if eval(args['ablations']['vib']):
        raise NotImplementedError('the VIB ablation (reference VIB.py) is outside the ibinn_b200 hot path')

    dev = torch.device(device)
    inn = GenerativeClassifier(args).to(dev)
    if dataset is None:
        if synthetic:
            print('ibinn_b200.train: training on SYNTHETIC N(0,1) images with random labels (throughput / smoke run only)', flush=True)
            dataset = SyntheticDataset(args, device=device)
        else:
            # the reference's own loader (reference data.py:93-191; outside the hot path, not restated here) must be importable
            try:
                import data
                dataset = data.Dataset(args)
            except (ImportError, AttributeError) as e:
                raise RuntimeError('ibinn_b200.train: no dataset given and the reference data.py is not importable (%s); put the '
                                   'reference on sys.path, pass dataset=..., or opt in to noise with synthetic=True' % e) from e
    if resume:
        inn.load(resume)
    step = TrainStep(inn, batch_size=eval(args['data']['batch_size']), use_graph=use_graph)

    logfile = open(join(output_dir, f'losses{ensemble_str}.dat'), 'w')

    def log_write(line, endline='\n'):
        print(line, flush=True)
        logfile.write(line)
        logfile.write(endline)

    train_names = [l for l in PLOT_COLUMNS if l[-3:] == '_tr']
    val_names = [l for l in PLOT_COLUMNS if l[-4:] == '_val']
    header_fmt = '{:>15}' * len(PLOT_COLUMNS)
    output_fmt = '{:15.1f}      {:04d}/{:04d}      {:04d}/{:04d}' + '{:15.5f}' * (len(PLOT_COLUMNS) - 3)
    if eval(t['exponential_scheduler']):
        sched = torch.optim.lr_scheduler.StepLR(inn.optimizer, gamma=0.002 ** (1 / N_epochs), step_size=1)
    else:
        sched = torch.optim.lr_scheduler.MultiStepLR(inn.optimizer, gamma=0.1, milestones=eval(t['scheduler_milestones']))
    log_write(header_fmt.format(*PLOT_COLUMNS))

    t_start = time()
    high_loss, n_iter, val_losses = False, 0, None
    try:
        for i_epoch in range(N_epochs):
            acc = torch.zeros(len(train_names), device=dev)       # running sums stay on the device
            n_acc = 0
            for i_batch, (x, l) in enumerate(dataset.train_loader):
                y = dataset.onehot(l.to(dev, non_blocking=True), label_smoothing)
                losses = step.step(x, y)
                acc += torch.stack([losses[k] for k in train_names])
                n_acc += 1
                n_iter += 1
                if not i_batch % interval_log:
                    val_losses = inn.validate(dataset.val_x, dataset.val_y)
                    row = dict(zip(train_names, (acc / n_acc).tolist()))          # the only host sync of the interval
                    row.update({k: val_losses[k].item() for k in val_names})
                    log_write(output_fmt.format((time() - t_start) / 60., i_epoch, N_epochs, i_batch, len(dataset.train_loader),
                                                *[row[k] for k in PLOT_COLUMNS[3:]]))
                    acc.zero_()
                    n_acc = 0
                if max_iterations is not None and n_iter >= max_iterations:
                    break
            # torch's scheduler wants optimizer.step() to have been called: the fused kernel did the step
            inn.optimizer._opt_called = True
            sched.step()
            if val_losses is not None and i_epoch > 2 and (val_losses['L_x_val'].item() > 1e5 or not np.isfinite(val_losses['L_x_val'].item())):
                if high_loss:
                    raise RuntimeError("loss is astronomical")
                high_loss = True
            else:
                high_loss = False
            if i_epoch > 0 and (i_epoch % interval_checkpoint) == 0:
                inn.save(join(output_dir, f'model_{i_epoch}{ensemble_str}.pt'))
            if max_iterations is not None and n_iter >= max_iterations:
                break
    except BaseException:
        if save_on_crash:
            inn.save(join(output_dir, f'model_ABORT{ensemble_str}.pt'))
        raise
    finally:
        logfile.close()
    for k in list(inn.inn._buffers.keys()):
        if 'tmp_var' in k:
            del inn.inn._buffers[k]
    inn.save(join(output_dir, f'model{ensemble_str}.pt'))
    return inn
