"""GPU data pipeline (csrc/augment.cu, ibinn_b200/data_gpu.py) against the reference's transform chain (reference data.py:47-90,
152-160) restated with torch / torchvision tensor ops on the CPU, with the random decisions injected on both sides."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import rel

from ibinn_b200 import config
from ibinn_b200.data_gpu import GpuAugmentor, GpuDataset

BETA = torch.tensor([0.4914, 0.4822, 0.4465]).view(1, 3, 1, 1)
GAMMA = 1. / torch.tensor([0.247, 0.243, 0.261]).view(1, 3, 1, 1)


def _images(n, gen, C=3, H=32):
    return torch.randint(0, 256, (n, H, H, C), generator=gen, dtype=torch.uint8)


def test_test_transform_is_totensor_plus_normalisation(dev):
    g = torch.Generator().manual_seed(0)
    img = _images(5, g)
    aug = GpuAugmentor(config.make_args(), dev)
    out = aug(img.to(dev), train=False)
    ref = GAMMA * (img.permute(0, 3, 1, 2).float() / 255. - BETA)
    assert torch.allclose(out.cpu(), ref, atol=1e-6, rtol=1e-6)


def test_flip_pad_crop_geometry_is_exact(dev):
    """angle 0, identity colour: flip -> Pad(8, edge) -> CenterCrop(36) -> RandomCrop(32) is a pure gather"""
    g = torch.Generator().manual_seed(1)
    B = 6
    img = _images(B, g)
    aug = GpuAugmentor(config.make_args(overrides={'data': {'dequantize_uniform': 'False', 'noise_amplitde': '0.'}}), dev)
    p = torch.zeros(B, 8)
    p[:, 0] = torch.tensor([0., 1., 0., 1., 1., 0.])
    p[:, 1:4] = 1.
    p[:, 5] = torch.tensor([0., 4., 2., 1., 3., 0.])
    p[:, 6] = torch.tensor([4., 0., 2., 3., 1., 0.])
    out = aug(img.to(dev), train=True, params=p.to(dev))
    x = img.permute(0, 3, 1, 2).float()
    for b in range(B):
        t = x[b:b + 1]
        if p[b, 0] > 0.5:
            t = t.flip(-1)
        t = F.pad(t, (8, 8, 8, 8), mode='replicate')[:, :, 6:42, 6:42]
        cy, cx = int(p[b, 6]), int(p[b, 5])
        t = t[:, :, cy:cy + 32, cx:cx + 32]
        ref = GAMMA * (t / 255. - BETA)
        assert torch.allclose(out[b:b + 1].cpu(), ref, atol=1e-6, rtol=1e-6), b


def test_rotation_matches_torchvision_nearest(dev):
    tvf = pytest.importorskip('torchvision.transforms.functional')
    from torchvision.transforms import InterpolationMode
    g = torch.Generator().manual_seed(2)
    B = 4
    img = _images(B, g)
    aug = GpuAugmentor(config.make_args(overrides={'data': {'dequantize_uniform': 'False', 'noise_amplitde': '0.'}}), dev)
    angles = torch.tensor([12., -7.5, 3., -11.])
    p = torch.zeros(B, 8)
    p[:, 1:4] = 1.
    p[:, 4] = angles * math.pi / 180.
    p[:, 5:7] = 2.
    out = aug(img.to(dev), train=True, params=p.to(dev)).cpu()
    x = img.permute(0, 3, 1, 2).float()
    for b in range(B):
        t = F.pad(x[b:b + 1], (8, 8, 8, 8), mode='replicate')
        t = tvf.rotate(t, float(angles[b]), interpolation=InterpolationMode.NEAREST)[:, :, 6:42, 6:42][:, :, 2:34, 2:34]
        ref = GAMMA * (t / 255. - BETA)
        mismatch = ((out[b:b + 1] - ref).abs() > 1e-5).any(dim=1).float().mean().item()
        assert mismatch < 0.03, (b, mismatch)          # nearest-neighbour ties at pixel boundaries


def test_colour_jitter_noise_and_padding_channels(dev):
    g = torch.Generator().manual_seed(3)
    B = 3
    img = _images(B, g)
    ov = {'data': {'noise_amplitde': '0.02', 'pad_noise_channels': '2', 'pad_noise_std': '0.3', 'tanh_augmentation': 'False'}}
    aug = GpuAugmentor(config.make_args(overrides=ov), dev)
    p = torch.zeros(B, 8)
    p[:, 1] = torch.tensor([0.93, 1.08, 1.0])
    p[:, 2] = torch.tensor([1.07, 0.91, 1.0])
    p[:, 3] = torch.tensor([0.96, 1.04, 1.0])
    p[:, 5:7] = 2.
    unif, gauss = torch.rand(B, 3, 32, 32, generator=g), torch.randn(B, 5, 32, 32, generator=g)
    out = aug(img.to(dev), train=True, params=p.to(dev), unif=unif.to(dev), gauss=gauss.to(dev)).cpu()
    x = img.permute(0, 3, 1, 2).double()
    w = torch.tensor([0.299, 0.587, 0.114], dtype=torch.float64).view(1, 3, 1, 1)
    for b in range(B):
        t = (p[b, 1].double() * x[b:b + 1]).clamp(0, 255)
        m = torch.floor((t * w).sum(1, keepdim=True).mean() + 0.5)
        t = (p[b, 2].double() * (t - m) + m).clamp(0, 255)
        gy = (t * w).sum(1, keepdim=True)
        t = (p[b, 3].double() * (t - gy) + gy).clamp(0, 255)
        t = t / 255. + unif[b:b + 1].double() / 256. + 0.02 * gauss[b:b + 1, :3].double()
        t = GAMMA.double() * (t - BETA.double())
        padc = torch.cat([t, t], 1)[:, :2] * math.sqrt(1. - 0.3 ** 2) + 0.3 * gauss[b:b + 1, 3:].double()
        ref = torch.cat([t, padc], 1)
        assert rel(out[b:b + 1], ref) < 1e-5, b


def test_gpu_dataset_feeds_the_training_loop(dev):
    """the reference data.Dataset surface: shuffled drop_last batches, 1024 validation images, one-hot labels"""
    g = torch.Generator().manual_seed(4)
    n = 1024 + 64 + 5
    img, lab = _images(n, g), torch.randint(0, 10, (n,), generator=g)
    args = config.make_args(overrides={'data': {'batch_size': '32'}})
    ds = GpuDataset(args, img, lab, device=str(dev))
    assert ds.val_x.shape == (1024, 3, 32, 32) and ds.val_y.shape == (1024, 10) and len(ds.train_loader) == 2
    seen = 0
    for x, l in ds.train_loader:
        assert x.shape == (32, 3, 32, 32) and l.shape == (32,) and torch.isfinite(x).all()
        y = ds.onehot(l, 0.02)
        assert torch.allclose(y.sum(1).cpu(), torch.ones(32))
        seen += 1
    assert seen == 2
