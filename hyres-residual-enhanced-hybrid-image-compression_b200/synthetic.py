"""Synthetic inputs for benchmarks and examples: smooth noise snapped to
k/255 (so the JPEG stage is meaningful) and a synthetic residual.  Seed 1926 is the reference's
(src/training.py:114)."""
import torch
import torch.nn.functional as F


def synthetic_image(B, H, W, seed=1926):
    """Bicubic x8 upsample of uniform noise + 0.02 randn, clamped to [0,1], snapped to k/255."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3, H // 8, W // 8, generator=g)
    x = F.interpolate(low, scale_factor=8, mode="bicubic", align_corners=False)
    x = x + 0.02 * torch.randn(B, 3, H, W, generator=g)
    return torch.round(x.clamp(0, 1) * 255) / 255


def synthetic_residual(B, H, W, seed=1926):
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(B, 3, H, W, generator=g)).clamp(-1, 1)
