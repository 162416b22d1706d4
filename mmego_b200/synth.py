"""Deterministic synthetic inputs and stand-in weights (numpy PCG64 streams are version-stable, so every box
regenerates the same tensors).  Shapes follow Config/config.py of the reference; value distributions follow the
statistics of Resource/Sample_data measured in SURVEY.md section 8(d)."""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from .Net import _layout

_SKELETON_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "skeleton.npy")


def default_skeleton() -> np.ndarray:
    """The single calibration skeleton of Resource/Sample_data: 20 bone vectors parent - child in skeleton_all order
    (Util/Universal_Util/Dataset_sample.py:167-169 of the reference)."""
    return np.load(_SKELETON_PATH).astype(np.float32)


def imu_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for the IMU_Net checkpoint, which is absent from the reference mount
    (.MISSING_LARGE_BLOBS).  U(-1/sqrt(fan), 1/sqrt(fan)) like torch's default init."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, kind, init in _layout.imu_layout(15, 9, 512, 2, True):
        out[name] = torch.from_numpy(rng.uniform(-float(init), float(init), size=shape).astype(np.float32))
    return out


def _random_rotations(rng: np.random.Generator, n: int, sigma_deg: float) -> np.ndarray:
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = np.abs(rng.normal(0.0, np.deg2rad(sigma_deg), size=n))
    K = np.zeros((n, 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -axis[:, 2], axis[:, 1]
    K[:, 1, 0], K[:, 1, 2] = axis[:, 2], -axis[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -axis[:, 1], axis[:, 0]
    s, c = np.sin(ang)[:, None, None], np.cos(ang)[:, None, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def batch(B: int, L: int = 20, N: int = 128, n_imu: int = 20, seed: int = 1234,
          skeleton: Optional[np.ndarray] = None, distinct_skeletons: bool = False) -> Dict[str, torch.Tensor]:
    """imu [B,L,n_imu,15], data [B,L,N,6] (x,y,z,range,velocity,intensity; ~40 % zero-padded slots), skl [B,20,3],
    plus a plausible head pose R,t for runs that bypass IMU_Net."""
    rng = np.random.default_rng(seed)
    F_ = B * L
    Rm = _random_rotations(rng, F_ * n_imu, 15.0).reshape(F_ * n_imu, 9)
    gyr = rng.normal([-4.1, -1.2, -1.1], [1.5, 1.0, 1.4], size=(F_ * n_imu, 3))
    acc = rng.normal(0.0, [0.29, 0.33, 0.61], size=(F_ * n_imu, 3))
    imu = np.concatenate([Rm, gyr, acc], axis=1).reshape(B, L, n_imu, 15).astype(np.float32)
    data = np.zeros((F_, N, 6), dtype=np.float32)
    nv = np.clip(np.rint(rng.normal(77.0, 21.0, size=F_) * (N / 128.0)), 3, N).astype(np.int64)
    order = np.argsort(rng.random((F_, N)), axis=1)
    valid = order < nv[:, None]
    px = rng.uniform(0.01, 2.0, size=(F_, N))
    py = rng.normal(0.04, 0.29, size=(F_, N))
    pz = rng.normal(0.19, 0.38, size=(F_, N))
    vel = rng.normal(0.0, 0.39, size=(F_, N))
    inten = 10.1 + rng.exponential(7.7, size=(F_, N))
    pts = np.stack([px, py, pz, np.sqrt(px * px + py * py + pz * pz), vel, inten], axis=-1).astype(np.float32)
    data[valid] = pts[valid]
    data = data.reshape(B, L, N, 6)
    if skeleton is None:
        skeleton = default_skeleton()
    skl = np.repeat(skeleton[None].astype(np.float32), B, axis=0)
    if distinct_skeletons:
        skl = skl * rng.uniform(0.85, 1.15, size=(B, 1, 1)).astype(np.float32)
    R = _random_rotations(rng, F_, 15.0).reshape(B, L, 3, 3).astype(np.float32)
    t = rng.normal([0.0, 0.0, 0.6], [0.1, 0.1, 0.05], size=(B, L, 3)).astype(np.float32)
    return dict(imu=torch.from_numpy(imu), data=torch.from_numpy(data), skl=torch.from_numpy(skl),
                R=torch.from_numpy(R), t=torch.from_numpy(t))


def target_like(pred: torch.Tensor, seed: int = 99, sigma: float = 0.03) -> torch.Tensor:
    """Synthetic ground truth for the metrics kernel: a prediction plus N(0, 3 cm) noise."""
    g = torch.Generator().manual_seed(seed)
    return pred.detach().cpu() + sigma * torch.randn(pred.shape, generator=g)
