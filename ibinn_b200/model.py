"""GenerativeClassifier with the reference's constructor, attributes, methods and loss dict
(reference model.py:12-252).  The INN and the GMM / Information-Bottleneck head run on ibinn_b200's
CUDA kernels; optimiser, checkpoint format and everything else stay plain torch.  The ablation
models (feed-forward ResNet, VIB) are outside the hot path and not provided."""
import numpy as np
import torch
import torch.nn as nn

from . import functions as Fn
from . import inn_architecture


class GenerativeClassifier(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        init_latent_scale = eval(self.args['model']['mu_init'])
        weight_init = eval(self.args['model']['weight_init'])
        self.dataset = self.args['data']['dataset']
        self.ch_pad = eval(self.args['data']['pad_noise_channels'])
        self.feed_forward = eval(self.args['ablations']['feed_forward_resnet'])
        self.feed_forward_revnet = eval(self.args['ablations']['feed_forward_irevnet'])

        if self.dataset == 'MNIST':
            self.dims = (28, 28)
            self.input_channels = 1
            self.n_classes = 10
        elif self.dataset in ['CIFAR10', 'CIFAR100']:
            self.dims = (3 + self.ch_pad, 32, 32)
            self.input_channels = 3 + self.ch_pad
            self.n_classes = 10 if self.dataset == 'CIFAR10' else 100
        else:
            raise ValueError(f"what is this dataset, {args['data']['dataset']}?")
        self.ndim_tot = int(np.prod(self.dims))

        if self.feed_forward or self.feed_forward_revnet:
            raise NotImplementedError('ibinn_b200 covers the INN path only; the feed-forward / i-RevNet ablations '
                                      '(reference feed_forward_architecture.py) are out of scope')
        self.inn = inn_architecture.constuct_inn(self)

        mu_populate_dims = self.ndim_tot
        init_scale = init_latent_scale / np.sqrt(2 * mu_populate_dims // self.n_classes)
        self.mu = nn.Parameter(torch.zeros(1, self.n_classes, self.ndim_tot))
        self.mu_empirical = eval(self.args['training']['empirical_mu'])
        for k in range(mu_populate_dims // self.n_classes):
            self.mu.data[0, :, self.n_classes * k: self.n_classes * (k + 1)] = init_scale * torch.eye(self.n_classes)
        self.phi = nn.Parameter(torch.zeros(self.n_classes))

        self.trainable_params = [p for p in self.inn.parameters() if p.requires_grad]
        self.train_mu = eval(self.args['training']['train_mu'])
        self.train_phi = eval(self.args['training']['train_mu'])      # sic: reference model.py:62 reads train_mu
        self.train_inn = True
        for p in self.trainable_params:
            p.data *= weight_init
        self.trainable_params += [self.mu, self.phi]

        t = self.args['training']
        optimizer = t['optimizer']
        base_lr = float(t['lr'])
        groups = [{'params': [p for p in self.inn.parameters() if p.requires_grad]}]
        for flag, par, key in ((self.train_mu, self.mu, 'mu'), (self.train_phi, self.phi, 'phi')):
            if not flag:
                continue
            g = {'params': [par], 'lr': base_lr * float(t['lr_' + key]), 'weight_decay': 0.}
            if optimizer == 'SGD':
                g['momentum'] = float(t['sgd_momentum_' + key])
            if optimizer == 'ADAM':
                g['betas'] = eval(t['adam_betas_' + key])
            if optimizer == 'AGGMO':
                g['betas'] = eval(t['aggmo_betas_' + key])
            groups.append(g)
        if optimizer == 'SGD':
            self.optimizer = torch.optim.SGD(groups, base_lr, momentum=float(t['sgd_momentum']), weight_decay=float(t['weight_decay']))
        elif optimizer == 'ADAM':
            self.optimizer = torch.optim.Adam(groups, base_lr, betas=eval(t['adam_betas']), weight_decay=float(t['weight_decay']))
        elif optimizer == 'AGGMO':
            import aggmo
            self.optimizer = aggmo.AggMo(groups, base_lr, betas=eval(t['aggmo_betas']), weight_decay=float(t['weight_decay']))
        else:
            raise ValueError(f'what is this optimizer, {optimizer}?')

    # ------------------------------------------------------------------ GMM helpers (plain torch, off the hot path)
    def cluster_distances(self, z, y=None):
        if y is not None:
            self._update_empirical_mu(z, y)
        z_i_z_i = torch.sum(z ** 2, dim=1, keepdim=True)
        mu_j_mu_j = torch.sum(self.mu ** 2, dim=2)
        z_i_mu_j = torch.mm(z, self.mu.squeeze().t())
        return -2 * z_i_mu_j + z_i_z_i + mu_j_mu_j

    def _update_empirical_mu(self, z, y):
        mu = torch.mm(z.t().detach(), y.round())
        mu = mu / torch.sum(y, dim=0, keepdim=True)
        # In place (the reference rebinds `.data`, model.py:115-120): engine.FlatParams keeps mu as a view of its flat buffer
        # and a captured CUDA graph reads that address, so the tensor must never move.
        self.mu.data.mul_(0.995).add_(mu.t().contiguous().view_as(self.mu), alpha=0.005)

    def mu_pairwise_dist(self):
        m = self.mu.squeeze()
        mu_i_mu_j = m.mm(m.t())
        mu_i_mu_i = torch.sum(m ** 2, 1, keepdim=True).expand(self.n_classes, self.n_classes)
        dist = mu_i_mu_i + mu_i_mu_i.t() - 2 * mu_i_mu_j
        off_diag = (1 - torch.eye(self.n_classes, device=dist.device)).bool()
        return torch.masked_select(dist, off_diag).clamp(min=0.)

    # ------------------------------------------------------------------ the hot path
    def forward(self, x, y=None, loss_mean=True):
        z = self.inn(x)
        jac = self.inn.log_jacobian(run_forward=False)
        if self.mu_empirical and y is not None and self.inn.training:
            self._update_empirical_mu(z, y)
        lx, logits, lc, ly, acc = Fn.GMMHeadFn.apply(z, jac, self.mu, self.phi, y, loss_mean)
        losses = {'L_x_tr': lx, 'logits_tr': logits}
        if y is not None:
            losses['L_cNLL_tr'] = lc
            losses['L_y_tr'] = ly
            losses['acc_tr'] = acc
        return losses

    def validate(self, x, y, eval_mode=True):
        is_train = self.inn.training
        if eval_mode:
            self.inn.eval()
        with torch.no_grad():
            losses = self.forward(x, y, loss_mean=False)
            l_x, class_nll, l_y, logits, acc = (losses['L_x_tr'].mean(), losses['L_cNLL_tr'].mean(), losses['L_y_tr'].mean(),
                                                losses['logits_tr'], losses['acc_tr'])
            mu_dist = torch.mean(torch.sqrt(self.mu_pairwise_dist()))
        if is_train:
            self.inn.train()
        return {'L_x_val': l_x, 'L_cNLL_val': class_nll, 'logits_val': logits, 'L_y_val': l_y, 'acc_val': acc,
                'delta_mu_val': mu_dist}

    def reset_mu(self, dataset):
        mu = torch.zeros(1, self.n_classes, self.ndim_tot, device=self.mu.device)
        counter = 0
        with torch.no_grad():
            for x, l in dataset.train_loader:
                x, y = x.to(self.mu.device), dataset.onehot(l.to(self.mu.device), 0.05)
                z = self.inn(x)
                mu_batch = torch.mm(z.t().detach(), y.round())
                mu_batch = mu_batch / torch.sum(y, dim=0, keepdim=True)
                mu += mu_batch.t().view(1, self.n_classes, -1)
                counter += 1
            mu /= counter
        self.mu.data.copy_(mu)          # in place: see _update_empirical_mu

    def sample(self, y, temperature=1., temperature_class=1.):
        z = temperature * torch.randn(y.shape[0], self.ndim_tot, device=self.mu.device)
        mu = temperature_class * torch.sum(y.round().view(-1, self.n_classes, 1) * self.mu, dim=1)
        return self.inn(z + mu, rev=True)

    def save(self, fname):
        """reference model.py:237-241.  The fused step keeps the SGD momentum in engine.FlatParams.buf rather than in
        the torch optimiser; `_ibinn_flat` (set by engine.TrainStep) exposes it as the optimiser's `momentum_buffer`
        entries first, so the `opt` entry is what torch.optim.SGD would have written."""
        flat = getattr(self, '_ibinn_flat', None)
        if flat is not None:
            flat.export_momentum()
        torch.save({'inn': self.inn.state_dict(), 'mu': self.mu, 'phi': self.phi, 'opt': self.optimizer.state_dict()}, fname)

    def load(self, fname):
        """reference model.py:243-248: strict on everything but the `tmp_var` scratch buffers."""
        data = torch.load(fname, weights_only=False)
        inn_state = {k: v for k, v in data['inn'].items() if 'tmp_var' not in k}
        own = {k for k in self.inn.state_dict() if 'tmp_var' not in k}
        missing, unexpected = sorted(own - set(inn_state)), sorted(set(inn_state) - own)
        if missing or unexpected:
            raise RuntimeError('Error(s) in loading state_dict for %s: missing keys %s, unexpected keys %s'
                               % (type(self.inn).__name__, missing[:8], unexpected[:8]))
        self.inn.load_state_dict(inn_state, strict=False)          # only the tmp_var buffers may be absent
        self.mu.data.copy_(data['mu'].data)
        self.phi.data.copy_(data['phi'].data)
        flat = getattr(self, '_ibinn_flat', None)
        if flat is not None and data.get('opt') is not None:
            flat.import_momentum(data['opt'])
