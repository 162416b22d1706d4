"""Drop-in for the reference's Net/Lower_Net.py (class LowerNet, :170-258).  The forward pass runs in
libmmego_b200 (mmego_lower_forward): ST-GCN key encoder, fused per-frame kernel (second in-place Transform2H,
top-64 select, BasePointNet, cross-attention, pooling), persistent bi-LSTM kernel, head GEMMs and decode kernel."""
from __future__ import annotations

import torch

from .. import _capi
from ..engine import MMEgoError, NativeNet
from . import _layout


class LowerNet(NativeNet):
    _net_id = _capi.NET_LOWER
    body_index_mode = "ref"          # see UpperNet.body_index_mode

    def __init__(self, hidden_dim):
        super().__init__()
        if hidden_dim != 64:
            raise MMEgoError("libmmego_b200 implements LowerNet(64) -- the configuration of the shipped checkpoint")
        self.hidden_dim = hidden_dim
        _layout.populate(self, _layout.lower_layout(hidden_dim))

    def forward(self, upper_l, x, h0_p, c0_p, h0_k, c0_k, initial_body, R, t):
        """upper_l [B,L,15,3] (not modified), x [B,L,N,6] (xyz transformed IN PLACE a second time, as
        Net/Lower_Net.py:191-192 does), four unused state arguments (the reference ignores them too),
        initial_body [B,20,3], R, t -> (l [B,L,8,3], q [B,L,6,3,3])."""
        x = self._cuda_f32(x, "x")
        if x.dim() != 4 or x.shape[-1] != 6:
            raise MMEgoError(f"x must be [B, L, N, 6] (got {tuple(x.shape)})")
        if not x.is_contiguous():
            raise MMEgoError("x must be contiguous: LowerNet.forward transforms its xyz channels in place")
        h = self._sync(x.device)
        upper_l, initial_body, R, t = (self._cuda_f32(v, n).contiguous() for v, n in (
            (upper_l, "upper_l"), (initial_body, "initial_body"), (R, "R"), (t, "t")))
        mode = _capi.BODY_REF if self.body_index_mode == "ref" else _capi.BODY_PER_SNIPPET
        return h.lower_forward(upper_l, x, initial_body, R, t, body_index_mode=mode)
