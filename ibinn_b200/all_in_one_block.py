"""AIO_Block with the reference's constructor, attributes, state_dict keys and FrEIA node protocol
(reference all_in_one_block.py:9-117); the arithmetic runs in ibinn_b200's fused CUDA kernels."""
import numpy as np
import torch
import torch.nn as nn
from scipy.stats import special_ortho_group

from . import functions as Fn
from .subnets import ConvSubnet


class AIO_Block(nn.Module):
    def __init__(self, dims_in, dims_c=[], subnet_constructor=None, clamp=2., gin_block=False, act_norm=1.,
                 permute_soft=False):
        super().__init__()
        channels = dims_in[0][0]
        if dims_c:
            raise ValueError('does not support conditioning yet')
        if gin_block:
            raise ValueError('gin_block=True is never used by IB-INN and not supported')
        self.split_len1 = channels - channels // 2
        self.split_len2 = channels // 2
        self.splits = [self.split_len1, self.split_len2]
        self.n_pixels = dims_in[0][1] * dims_in[0][2]
        self.in_channels = channels
        self.clamp = clamp
        self.GIN = gin_block

        self.act_norm = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.act_offset = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.act_norm_trigger = True
        if act_norm:
            self.act_norm.data += np.log(act_norm)
            self.act_norm_trigger = False

        if permute_soft:
            w = special_ortho_group.rvs(channels)
        else:
            w = np.zeros((channels, channels))
            for i, j in enumerate(np.random.permutation(channels)):
                w[i, j] = 1.
        w_inv = np.linalg.inv(w)
        self.w = nn.Parameter(torch.FloatTensor(w).view(channels, channels, 1, 1), requires_grad=False)
        self.w_inv = nn.Parameter(torch.FloatTensor(w_inv).view(channels, channels, 1, 1), requires_grad=False)
        self.conditional = False

        self.s = subnet_constructor(self.split_len1, 2 * self.split_len2)
        if not isinstance(self.s, ConvSubnet):
            raise TypeError('ibinn_b200: subnet_constructor must build an ibinn_b200.subnets.ConvSubnet (no generic fallback)')
        self.last_jac = None

    def forward(self, x, c=[], rev=False, mask=None):
        if self.act_norm_trigger:
            # data-dependent ActNorm initialisation (reference :84-91); dead for every shipped config
            with torch.no_grad():
                print('ActNorm triggered')
                self.act_norm_trigger = False
                x_out = self.forward(x)[0]
                x_out = x_out.transpose(0, 1).contiguous().view(self.in_channels, -1)
                self.act_norm.data -= x_out.std(dim=1, unbiased=False).log().view(1, self.in_channels, 1, 1)
                self.act_offset.data -= x_out.mean(dim=1).view(1, self.in_channels, 1, 1)
        if rev:
            out, j = Fn.aio_block_reverse(self, x[0])
            self.last_jac = j
        else:
            params = Fn.conv_subnet_params(self.s) + [self.act_norm, self.act_offset]
            out, j = Fn.AIOBlockFn.apply(x[0], self, mask, *params)
            self.last_jac = j
        return [out]

    def jacobian(self, x, c=[], rev=False):
        # the kernels already add (-1)**rev * act_norm.sum() * n_pixels (reference :113-114)
        return self.last_jac

    def output_dims(self, input_dims):
        return input_dims
