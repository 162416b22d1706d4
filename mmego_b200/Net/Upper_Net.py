"""Drop-in for the reference's Net/Upper_Net.py (class UpperNet, :367-404).  The forward pass runs in
libmmego_b200 (mmego_upper_forward): fused Transform2H + PointNet + GlobalPointNet + attention pooling kernel,
persistent bi-LSTM kernel, head GEMMs, and the 6D->rotation / forward-kinematics / Transform2R decode kernel."""
from __future__ import annotations

import torch

from .. import _capi
from ..engine import MMEgoError, NativeNet
from . import _layout


class UpperNet(NativeNet):
    _net_id = _capi.NET_UPPER
    # "ref" reproduces initial_body[r % B] of the reference's ForKinematics (.repeat(L,1,1,1), Net/Upper_Net.py:134);
    # "per_snippet" uses initial_body[r // L].
    body_index_mode = "ref"

    def __init__(self):
        super().__init__()
        _layout.populate(self, _layout.upper_layout())

    def forward(self, x, h0_g, c0_g, initial_body, R, t):
        """x [B,L,N,6] (xyz OVERWRITTEN in place with R(xyz - t), exactly like the reference's Transform2H),
        h0_g/c0_g [6,B,64], initial_body [B,20,3], R [B,L,3,3], t [B,L,3]
        -> (l [B,L,15,3], q [B,L,14,3,3], global_weights [B*L,N,1], hn_g, cn_g [6,B,64])."""
        x = self._cuda_f32(x, "x")
        if x.dim() != 4 or x.shape[-1] != 6:
            raise MMEgoError(f"x must be [B, L, N, 6] (got {tuple(x.shape)})")
        if not x.is_contiguous():
            raise MMEgoError("x must be contiguous: UpperNet.forward transforms its xyz channels in place")
        h = self._sync(x.device)
        args = [self._cuda_f32(v, n).contiguous() for v, n in ((h0_g, "h0_g"), (c0_g, "c0_g"),
                                                              (initial_body, "initial_body"), (R, "R"), (t, "t"))]
        mode = _capi.BODY_REF if self.body_index_mode == "ref" else _capi.BODY_PER_SNIPPET
        return h.upper_forward(x, *args, body_index_mode=mode)
