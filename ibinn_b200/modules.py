"""The five FrEIA-v0.2 modules IB-INN uses (reference inn_architecture.py:126,127,157,160,161),
re-provided on top of the ibinn_b200 kernels.  Behaviour restated in SURVEY.md Appendix C."""
import numpy as np
import torch
import torch.nn as nn

from . import functions as Fn
from .subnets import FCSubnet


class Reshape(nn.Module):
    def __init__(self, dims_in, target_dim):
        super().__init__()
        self.size = tuple(dims_in[0])
        self.target_dim = tuple(target_dim)

    def forward(self, x, rev=False):
        return [x[0].reshape(x[0].shape[0], *(self.size if rev else self.target_dim))]

    def jacobian(self, x, rev=False):
        return 0.

    def output_dims(self, dims):
        return [self.target_dim]


class Flatten(nn.Module):
    def __init__(self, dims_in):
        super().__init__()
        self.size = tuple(dims_in[0])

    def forward(self, x, rev=False):
        if rev:
            return [x[0].reshape(x[0].shape[0], *self.size)]
        return [x[0].reshape(x[0].shape[0], -1)]

    def jacobian(self, x, rev=False):
        return 0.

    def output_dims(self, dims):
        return [(int(np.prod(dims[0])),)]


class HaarDownsampling(nn.Module):
    def __init__(self, dims_in, order_by_wavelet=False, rebalance=1.):
        super().__init__()
        if order_by_wavelet:
            raise ValueError('order_by_wavelet is not used by IB-INN and not supported')
        self.in_channels = dims_in[0][0]
        self.fac_fwd, self.fac_rev = 0.5 * rebalance, 0.5 / rebalance
        k = torch.ones(4, 1, 2, 2)
        k[1, 0, :, 1] = -1
        k[2, 0, 1, :] = -1
        k[3, 0, 1, 0] = -1
        k[3, 0, 0, 1] = -1
        # kept only so that the state_dict matches FrEIA checkpoints; the kernel hard-wires the signs
        self.haar_weights = nn.Parameter(torch.cat([k] * self.in_channels, 0), requires_grad=False)
        elements = dims_in[0][0] * dims_in[0][1] * dims_in[0][2]
        self.last_jac = elements / 4 * (np.log(16.) + 4 * np.log(self.fac_fwd))

    def forward(self, x, rev=False):
        if rev:
            return [Fn.ops.haar_inverse(x[0], self.fac_rev)]
        return [Fn.HaarFn.apply(x[0], self.fac_fwd)]

    def jacobian(self, x, rev=False):
        return -self.last_jac if rev else self.last_jac

    def output_dims(self, dims):
        c, h, w = dims[0]
        return [(c * 4, h // 2, w // 2)]


class PermuteRandom(nn.Module):
    def __init__(self, dims_in, seed):
        super().__init__()
        self.in_channels = dims_in[0][0]
        np.random.seed(seed)
        perm = np.random.permutation(self.in_channels)
        np.random.seed()
        perm_inv = np.zeros_like(perm)
        perm_inv[perm] = np.arange(len(perm))
        self.perm = torch.LongTensor(perm)
        self.perm_inv = torch.LongTensor(perm_inv)
        self._dev = {}

    def _index(self, dev):
        key = (dev.type, dev.index)
        if key not in self._dev:
            self._dev[key] = (self.perm.to(dev, torch.int32), self.perm_inv.to(dev, torch.int32))
        return self._dev[key]

    def forward(self, x, rev=False):
        p, pi = self._index(x[0].device)
        if rev:
            return [Fn.PermuteFn.apply(x[0], pi, p)]
        return [Fn.PermuteFn.apply(x[0], p, pi)]

    def jacobian(self, x, rev=False):
        return 0.

    def output_dims(self, dims):
        return dims


class GLOWCouplingBlock(nn.Module):
    def __init__(self, dims_in, dims_c=[], subnet_constructor=None, clamp=5.):
        super().__init__()
        if dims_c:
            raise ValueError('conditioning is not used by IB-INN and not supported')
        channels = dims_in[0][0]
        if len(dims_in[0]) != 1:
            raise ValueError('GLOWCouplingBlock: only the fully connected (flattened) form is supported')
        self.split_len1 = channels // 2
        self.split_len2 = channels - channels // 2
        self.clamp = clamp
        self.s1 = subnet_constructor(self.split_len1, self.split_len2 * 2)
        self.s2 = subnet_constructor(self.split_len2, self.split_len1 * 2)
        for s in (self.s1, self.s2):
            if not isinstance(s, FCSubnet):
                raise TypeError('ibinn_b200: subnet_constructor must build an ibinn_b200.subnets.FCSubnet (no generic fallback)')
        self.last_jac = None

    def forward(self, x, c=[], rev=False, masks=None):
        if rev:
            out, self.last_jac = Fn.glow_block_reverse(self, x[0])
        else:
            params = Fn.fc_subnet_params(self.s1) + Fn.fc_subnet_params(self.s2)
            out, self.last_jac = Fn.GlowBlockFn.apply(x[0], self, masks, *params)
        return [out]

    def jacobian(self, x, c=[], rev=False):
        return self.last_jac

    def output_dims(self, dims):
        return dims
