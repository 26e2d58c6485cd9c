"""Shared helpers for the parity tests: build ibinn_b200 nodes and the matching oracle state."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import ibinn_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def randomize_(module, gen, scale=0.1):
    """Move every trainable tensor / BN buffer off its symmetric init so all gradient paths are exercised."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.requires_grad and (n.endswith('bias') or n.endswith('act_offset') or n.endswith('act_norm')):
                p.add_(scale * torch.randn(p.shape, generator=gen))
            if p.requires_grad and n.endswith('weight') and p.dim() == 1:
                p.add_(scale * torch.randn(p.shape, generator=gen))
        for n, b in module.named_buffers():
            if n.endswith('running_mean'):
                b.add_(scale * torch.randn(b.shape, generator=gen))
            if n.endswith('running_var'):
                b.mul_(1. + 0.3 * torch.rand(b.shape, generator=gen))


def state_of(module, prefix, dtype=torch.float64):
    """Oracle parameter dict from a module's state_dict, keys prefixed like `module_list.<i>.`"""
    return {prefix + k: v.detach().cpu().to(dtype).clone() if v.is_floating_point() else v.detach().cpu().clone()
            for k, v in module.state_dict().items()}
