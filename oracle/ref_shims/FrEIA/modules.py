"""Stand-in for FrEIA.modules (v0.2): the five modules the reference uses
(/root/reference/inn_architecture.py:126,127,157,160,161).  Restated from the
published behaviour.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Reshape(nn.Module):
    def __init__(self, dims_in, target_dim):
        super().__init__()
        self.size = dims_in[0]
        self.target_dim = tuple(target_dim)

    def forward(self, x, rev=False):
        if not rev:
            return [x[0].reshape(x[0].shape[0], *self.target_dim)]
        return [x[0].reshape(x[0].shape[0], *self.size)]

    def jacobian(self, x, rev=False):
        return 0.

    def output_dims(self, dims):
        return [self.target_dim]


class Flatten(nn.Module):
    def __init__(self, dims_in):
        super().__init__()
        self.size = dims_in[0]

    def forward(self, x, rev=False):
        if not rev:
            return [x[0].reshape(x[0].shape[0], -1)]
        return [x[0].reshape(x[0].shape[0], *self.size)]

    def jacobian(self, x, rev=False):
        return 0.

    def output_dims(self, dims):
        return [(int(np.prod(dims[0])),)]


class HaarDownsampling(nn.Module):
    def __init__(self, dims_in, order_by_wavelet=False, rebalance=1.):
        super().__init__()
        self.in_channels = dims_in[0][0]
        self.fac_fwd = 0.5 * rebalance
        self.fac_rev = 0.5 / rebalance
        k = torch.ones(4, 1, 2, 2)
        k[1, 0, 0, 1] = -1
        k[1, 0, 1, 1] = -1
        k[2, 0, 1, 0] = -1
        k[2, 0, 1, 1] = -1
        k[3, 0, 1, 0] = -1
        k[3, 0, 0, 1] = -1
        self.haar_weights = nn.Parameter(torch.cat([k] * self.in_channels, 0), requires_grad=False)
        self.elements = dims_in[0][0] * dims_in[0][1] * dims_in[0][2]
        self.last_jac = self.elements / 4 * (np.log(16.) + 4 * np.log(self.fac_fwd))

    def forward(self, x, rev=False):
        if not rev:
            out = F.conv2d(x[0], self.haar_weights, bias=None, stride=2, groups=self.in_channels)
            return [out * self.fac_fwd]
        return [F.conv_transpose2d(x[0] * self.fac_rev, self.haar_weights, bias=None, stride=2,
                                   groups=self.in_channels)]

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
        self.perm = np.random.permutation(self.in_channels)
        np.random.seed()
        self.perm_inv = np.zeros_like(self.perm)
        for i, p in enumerate(self.perm):
            self.perm_inv[p] = i
        self.perm = torch.LongTensor(self.perm)
        self.perm_inv = torch.LongTensor(self.perm_inv)

    def forward(self, x, rev=False):
        if not rev:
            return [x[0][:, self.perm]]
        return [x[0][:, self.perm_inv]]

    def jacobian(self, x, rev=False):
        return 0.

    def output_dims(self, dims):
        return dims


class GLOWCouplingBlock(nn.Module):
    def __init__(self, dims_in, dims_c=[], subnet_constructor=None, clamp=5.):
        super().__init__()
        channels = dims_in[0][0]
        self.split_len1 = channels // 2
        self.split_len2 = channels - channels // 2
        self.clamp = clamp
        self.s1 = subnet_constructor(self.split_len1, self.split_len2 * 2)
        self.s2 = subnet_constructor(self.split_len2, self.split_len1 * 2)
        self.last_jac = None

    def log_e(self, s):
        return self.clamp * 0.636 * torch.atan(s / self.clamp)

    def forward(self, x, c=[], rev=False):
        x1 = x[0].narrow(1, 0, self.split_len1)
        x2 = x[0].narrow(1, self.split_len1, self.split_len2)
        if not rev:
            r2 = self.s2(x2)
            s2, t2 = r2[:, :self.split_len1], r2[:, self.split_len1:]
            y1 = torch.exp(self.log_e(s2)) * x1 + t2
            r1 = self.s1(y1)
            s1, t1 = r1[:, :self.split_len2], r1[:, self.split_len2:]
            y2 = torch.exp(self.log_e(s1)) * x2 + t1
            self.last_jac = self.log_e(s1).sum(dim=1) + self.log_e(s2).sum(dim=1)
        else:
            r1 = self.s1(x1)
            s1, t1 = r1[:, :self.split_len2], r1[:, self.split_len2:]
            y2 = (x2 - t1) / torch.exp(self.log_e(s1))
            r2 = self.s2(y2)
            s2, t2 = r2[:, :self.split_len1], r2[:, self.split_len1:]
            y1 = (x1 - t2) / torch.exp(self.log_e(s2))
            self.last_jac = -self.log_e(s1).sum(dim=1) - self.log_e(s2).sum(dim=1)
        return [torch.cat((y1, y2), 1)]

    def jacobian(self, x, c=[], rev=False):
        return self.last_jac

    def output_dims(self, dims):
        return dims
