"""Training-data pipeline on the GPU: the reference's `data.Augmentor` (data.py:47-90) and the torchvision
transform chain of its CIFAR training set (data.py:152-160) as one CUDA kernel per batch (csrc/augment.cu), fed from the raw uint8
images held in HBM.  `GpuDataset` offers the surface of the reference's `data.Dataset` that train.py touches (`train_loader`,
`val_x`, `val_y`, `onehot`), so `ibinn_b200.train.train(args, dataset=GpuDataset(...))` runs without the 6-worker PIL loader
(which caps the reference at a few thousand images per second).  Reading the dataset files themselves stays the caller's job."""
import ctypes
import math

import torch

from . import _lib
from ._lib import check, ptr, stream


class GpuAugmentor:
    def __init__(self, args, device):
        d = args['data']
        self.dataset = d['dataset']
        self.sigma = float(eval(d['noise_amplitde']))
        self.unif = bool(eval(d['dequantize_uniform']))
        self.tanh = bool(eval(d['tanh_augmentation']))
        self.pad = int(eval(d['pad_noise_channels']))
        self.pad_sigma = float(eval(d['pad_noise_std']))
        if self.dataset == 'MNIST':                                   # data.py:103-105
            self.C, self.H, self.W = 1, 28, 28
            self.beta, self.gamma = [0.5], [2.]
            if self.pad:
                raise ValueError('needs to be fixed, channel padding does not work with mnist')     # data.py:117
        else:                                                         # data.py:107-110
            self.C, self.H, self.W = 3, 32, 32
            self.beta = [0.4914, 0.4822, 0.4465]
            self.gamma = [1. / 0.247, 1. / 0.243, 1. / 0.261]
        self.geometric = self.dataset != 'MNIST'                      # the MNIST training set only gets ToTensor + Augmentor (data.py:125-127)
        self.device = torch.device(device)

    def draw_params(self, B, generator=None):
        """(B,8) random decisions of data.py:152-157: flip p = 0.5, ColorJitter(0.1, 0.1, 0.05), RandomRotation(12), RandomCrop(32) of 36"""
        r = torch.rand(B, 8, device=self.device, generator=generator)
        p = torch.empty(B, 8, device=self.device)
        p[:, 0] = (r[:, 0] < 0.5).float()
        p[:, 1] = 0.9 + 0.2 * r[:, 1]
        p[:, 2] = 0.9 + 0.2 * r[:, 2]
        p[:, 3] = 0.95 + 0.1 * r[:, 3]
        p[:, 4] = (r[:, 4] * 2. - 1.) * (12. * math.pi / 180.)
        p[:, 5] = torch.floor(r[:, 5] * 5.).clamp_(max=4.)
        p[:, 6] = torch.floor(r[:, 6] * 5.).clamp_(max=4.)
        p[:, 7] = 0.
        return p

    def __call__(self, images, index=None, train=True, generator=None, params=None, unif=None, gauss=None):
        """images (N,H,W,C) uint8 on the device; index (B) int32 picks the batch (None: all N in order).  train=False is the
        reference's deterministic test transform.  params / unif / gauss may be injected (tests); otherwise drawn here."""
        N, H, W, C = images.shape
        assert (H, W, C) == (self.H, self.W, self.C) and images.dtype == torch.uint8
        B = N if index is None else index.shape[0]
        dev = images.device
        if train:
            if params is None and self.geometric:
                params = self.draw_params(B, generator)
            if unif is None and self.unif:
                unif = torch.rand(B, C, H, W, device=dev, generator=generator)
            if gauss is None and (self.sigma > 0. or self.pad):
                gauss = torch.randn(B, C + self.pad, H, W, device=dev, generator=generator)
        else:
            params = unif = None
            gauss = torch.randn(B, C + self.pad, H, W, device=dev, generator=generator) if self.pad and gauss is None else gauss
        out = torch.empty(B, C + self.pad, H, W, dtype=torch.float32, device=dev)
        beta = (ctypes.c_float * C)(*self.beta)
        gamma = (ctypes.c_float * C)(*self.gamma)
        check(_lib.lib().ibinn_augment_batch(ptr(images, torch.uint8), ptr(index, torch.int32), ptr(params), ptr(unif), ptr(gauss), ptr(out),
                                             B, H, W, C, self.pad, beta, gamma, self.sigma if train else 0., self.pad_sigma, int(self.tanh),
                                             stream(images)), 'augment_batch')
        return out


class GpuDataset:
    """`data.Dataset` as train.py uses it, fed from device-resident uint8 images: `train_loader` yields (x, labels) batches already
    augmented on the GPU (shuffled, drop_last, data.py:172-173), `val_x` / `val_y` are the last 1024 training images through the
    test transform (data.py:164-168), `onehot` is data.py:186-191."""

    def __init__(self, args, images, labels, device='cuda', seed=0, n_val=1024):
        self.device = torch.device(device)
        self.batch_size = eval(args['data']['batch_size'])
        self.label_smoothing = eval(args['data']['label_smoothing'])
        self.aug = GpuAugmentor(args, self.device)
        self.n_classes = 100 if self.aug.dataset == 'CIFAR100' else 10
        self.dims = (self.aug.C + self.aug.pad, self.aug.H, self.aug.W) if self.aug.dataset != 'MNIST' else (28, 28)
        images = images.to(self.device)
        labels = labels.to(self.device).long()
        g = torch.Generator(device='cpu').manual_seed(seed)
        perm = torch.randperm(images.shape[0], generator=g).to(self.device)          # random_split (data.py:162)
        self.train_idx, val_idx = perm[:-n_val], perm[-n_val:]
        self.images, self.labels = images, labels
        self.gen = torch.Generator(device=self.device).manual_seed(seed + 1) if self.device.type == 'cuda' else None
        self.val_x = self._shape(self.aug(images, val_idx.int(), train=False))
        self.val_y = self.onehot(labels[val_idx], self.label_smoothing)
        self.train_loader = self

    def _shape(self, x):
        return x.view(x.shape[0], 28, 28) if self.aug.dataset == 'MNIST' else x

    def __len__(self):
        return self.train_idx.shape[0] // self.batch_size

    def __iter__(self):
        order = self.train_idx[torch.randperm(self.train_idx.shape[0], device=self.device, generator=self.gen)]
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            yield self._shape(self.aug(self.images, idx.int(), train=True, generator=self.gen)), self.labels[idx]

    def onehot(self, l, label_smooth=0):
        y = torch.zeros(l.shape[0], self.n_classes, device=l.device)
        y.scatter_(1, l.view(-1, 1), 1.)
        if label_smooth:
            y = y * (1 - label_smooth) + label_smooth / self.n_classes
        return y
