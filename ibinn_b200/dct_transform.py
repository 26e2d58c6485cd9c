"""DCTPooling2d with the reference's constructor, parameters and declared log-det
(reference dct_transform.py:64-110).  The weight is the analytic DCT-II matrix
W[k,n] = 2 cos(pi (2n+1) k / (2N)) that the reference obtains by pushing the identity through its
FFT-based dct_1d (dct_transform.py:7-31,75); the inverse is W^-1 (dct_transform.py:33-62,76)."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import functions as Fn


def dct_matrix(n):
    k = torch.arange(n, dtype=torch.float64).view(-1, 1)
    m = torch.arange(n, dtype=torch.float64).view(1, -1)
    return 2. * torch.cos(math.pi * (2 * m + 1) * k / (2 * n))


class DCTPooling2d(nn.Module):
    def __init__(self, dims_in, rebalance=1.):
        super().__init__()
        self.ch = dims_in[0][0]
        self.N = dims_in[0][1]
        if dims_in[0][1] != dims_in[0][2]:
            raise ValueError('DCTPooling2d needs square feature maps')
        self.rebalance = 2 * (self.N + 1) / rebalance
        self.jac = (self.N ** 2 * self.ch) * np.log(rebalance)
        W = dct_matrix(self.N)
        self.weight = nn.Parameter(W.float().contiguous(), requires_grad=False)
        self.inv_weight = nn.Parameter(torch.linalg.inv(W).float().contiguous(), requires_grad=False)

    def forward(self, x, rev=False):
        x = x[0]
        if rev:
            return [Fn.ops.dct2d_inverse(x, self.inv_weight, self.ch, self.N, float(self.rebalance), False)]
        return [Fn.DCTFn.apply(x, self.weight, float(1. / self.rebalance))]

    def jacobian(self, x, rev=False):
        return self.jac

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Can only use 1 input"
        c, w, h = input_dims[0]
        return [(c * w * h,)]
