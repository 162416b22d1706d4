"""Drop-in for the two rigid-transform helpers of the reference's Util/Universal_Util/Utils.py:274-292."""
from __future__ import annotations

import torch

from ...Config.config import Config
from ...engine import MMEgoError, get_handle


def Transform2R(points, batch_size, length_size, N, R, t):
    """points [B, L, N, 3] head frame -> reference frame, R^T p + t (out of place).  Utils.py:274-281."""
    if not Config.IMU_used:
        return points
    if not points.is_cuda:
        raise MMEgoError("Transform2R: mmego_b200 has no CPU path")
    h = get_handle(points.device)
    out = h.transform2r(points.reshape(batch_size * length_size, N, 3).contiguous(),
                        R.reshape(batch_size * length_size, 3, 3).contiguous(),
                        t.reshape(batch_size * length_size, 3).contiguous())
    return out.view(batch_size * length_size, N, 3)


def Transform2H(points, batch_size, length_size, N, R, t):
    """points [B*L, N, D>=3]: xyz <- R (xyz - t) IN PLACE on the caller's storage (Utils.py:284-292)."""
    if not Config.IMU_used:
        return points
    if not points.is_cuda:
        raise MMEgoError("Transform2H: mmego_b200 has no CPU path")
    if not points.is_contiguous():
        raise MMEgoError("Transform2H works in place and needs a contiguous tensor")
    h = get_handle(points.device)
    h.transform2h_(points.view(batch_size * length_size, N, points.shape[-1]),
                   R.reshape(batch_size * length_size, 3, 3).contiguous(),
                   t.reshape(batch_size * length_size, 3).contiguous())
    return points
