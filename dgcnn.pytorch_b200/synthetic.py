"""Synthetic inputs of the shapes the reference trains on; no datasets exist offline.
Pure host-side generators (CPU tensors); the benchmark and the tools copy them to the device."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def synthetic_xyz(B: int, N: int, seed: int = 1) -> torch.Tensor:
    """ModelNet40 / ShapeNet-shape clouds: Gaussian points, centred, scaled into the unit ball
    (as the pre-normalised HDF5 the reference's loaders read, data.py:86-95) -> [B,3,N] fp32, the
    layout main_cls.py:91 feeds the model."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, N, 3, generator=g)
    p = p - p.mean(dim=1, keepdim=True)
    p = p / p.norm(dim=2).amax(dim=1).view(B, 1, 1)
    return p.permute(0, 2, 1).contiguous()


def synthetic_features(B: int, C: int, N: int, seed: int = 1) -> torch.Tensor:
    """Post-activation-like feature clouds [B,C,N] (inputs of the feature-space layers)."""
    g = torch.Generator().manual_seed(seed)
    return F.leaky_relu(torch.randn(B, C, N, generator=g), 0.2)


def synthetic_s3dis(B: int, N: int, seed: int = 1) -> torch.Tensor:
    """S3DIS-shape 9-channel blocks [B,9,N]: block xyz U[-0.5,0.5]^2 x U[0,3], rgb U[0,1], room-
    normalised xyz U[0,1] (prepare_data/indoor3d_util.py:243-260)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(B, 9, N, generator=g)
    u[:, 0:2] -= 0.5
    u[:, 2] *= 3.0
    return u.contiguous()
