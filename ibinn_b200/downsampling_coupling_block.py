"""DownsampleCouplingBlock with the reference's constructor, attributes and state_dict keys
(reference downsampling_coupling_block.py:7-104); space-to-depth and the affine couplings are fused
ibinn_b200 kernels instead of grouped stride-2 convolutions with one-hot kernels."""
from math import exp

import torch
import torch.nn as nn

from . import functions as Fn
from .subnets import ConvSubnet


class DownsampleCouplingBlock(nn.Module):
    def __init__(self, dims_in, dims_c=[], subnet_constructor_strided=None, subnet_constructor_low_res=None, clamp=2.):
        super().__init__()
        channels = dims_in[0][0]
        if dims_c:
            raise ValueError('does not support conditioning yet')
        self.split_len1 = channels // 2
        self.split_len2 = channels - channels // 2
        self.splits = [self.split_len1, self.split_len2]
        self.in_channels = channels
        self.clamp = clamp
        self.max_s, self.min_s = exp(clamp), exp(-clamp)
        self.conditional = False

        self.s_hi = subnet_constructor_strided(self.split_len1, 8 * self.split_len2)
        self.s_lo = subnet_constructor_low_res(4 * self.split_len2, self.split_len1 * 8)
        for s in (self.s_hi, self.s_lo):
            if not isinstance(s, ConvSubnet):
                raise TypeError('ibinn_b200: subnet constructors must build ibinn_b200.subnets.ConvSubnet (no generic fallback)')
        self.last_jac = None

        # one-hot 2x2 kernels: only kept so that state_dicts interchange with the reference
        rw = torch.zeros(4, 1, 2, 2)
        rw[0, 0, 0, 0] = 1
        rw[1, 0, 0, 1] = 1
        rw[2, 0, 1, 0] = 1
        rw[3, 0, 1, 1] = 1
        self.reshape_kernels = nn.ParameterList()
        for split in self.splits:
            self.reshape_kernels.append(nn.Parameter(torch.cat([rw] * split, 0), requires_grad=False))

    def forward(self, x, c=[], rev=False, mask=None):
        if rev:
            out, self.last_jac = Fn.down_block_reverse(self, x[0])
        else:
            params = Fn.conv_subnet_params(self.s_hi) + Fn.conv_subnet_params(self.s_lo)
            out, self.last_jac = Fn.DownBlockFn.apply(x[0], self, mask, *params)
        return [out]

    def jacobian(self, x, c=[], rev=False):
        return self.last_jac

    def output_dims(self, input_dims):
        c, h, w = input_dims[0]
        return [[c * 4, h // 2, w // 2]]
