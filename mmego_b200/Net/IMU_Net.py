"""Drop-in for the reference's Net/IMU_Net.py (class IMUNet, :50-114): same constructor, forward signature,
state_dict keys and save/load; the forward pass runs in libmmego_b200 (mmego_imu_forward)."""
from __future__ import annotations

import torch

from .. import _capi
from ..engine import MMEgoError, NativeNet
from . import _layout


class IMUNet(NativeNet):
    _net_id = _capi.NET_IMU

    def __init__(self, input_n, output_n, hidden_n, n_rnn_layer, bidirectional=True, dropout=0):
        super().__init__()
        if (input_n, output_n, hidden_n, n_rnn_layer, bool(bidirectional)) != (15, 9, 512, 2, True):
            raise MMEgoError("libmmego_b200 implements IMUNet(15, 9, 512, 2, True, .) -- the configuration used at "
                             "Processor/Test/Demo_test.py:54 of the reference")
        self.dropout = dropout          # inactive at inference
        _layout.populate(self, _layout.imu_layout(input_n, output_n, hidden_n, n_rnn_layer, bidirectional))

    def forward(self, imu, h0_i=None):
        """imu [B, L, n, 15] -> R [B, L, 3, 3], t [B, L, 3]   (Net/IMU_Net.py:67-94).
        h0_i must be None: the reference passes it to two LSTMs with different batch sizes, so nothing else works."""
        if h0_i is not None:
            raise MMEgoError("IMUNet.forward: only h0_i=None is supported (as in the reference's eval loop)")
        imu = self._cuda_f32(imu, "imu")
        if imu.dim() != 4 or imu.shape[-1] != 15:
            raise MMEgoError(f"imu must be [B, L, n, 15] (got {tuple(imu.shape)})")
        h = self._sync(imu.device)
        return h.imu_forward(imu.contiguous())
