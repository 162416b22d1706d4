"""Drop-in for the feature-extractor surface of the reference's Net/GCN.py (class Model, :301-355) as used by
KeyEncoder (Net/Lower_Net.py:149-167): Model(3, 64, {'layout': 'kinect_upper', 'strategy': 'distance'}).extract_feature."""
from __future__ import annotations

import torch

from .. import _capi
from ..engine import MMEgoError, NativeNet
from . import _layout


class Graph:
    """Adjacency of the 15-node upper-body graph, strategy 'distance', max_hop 1 (Net/GCN.py:150-278)."""

    def __init__(self, layout="kinect_upper", strategy="distance", max_hop=1, dilation=1):
        if (layout, strategy, max_hop, dilation) != ("kinect_upper", "distance", 1, 1):
            raise MMEgoError("only Graph('kinect_upper', 'distance', max_hop=1) is implemented (the one LowerNet uses)")
        self.num_node = 15
        self.A = _layout.graph_adjacency().numpy()


class Model(NativeNet):
    """Stand-alone ST-GCN feature extractor.  Its state_dict uses the un-prefixed key names of GCN.Model; the
    library consumes them under the Lower_Net prefix `keyEncoder.gcn.`."""
    _net_id = _capi.NET_LOWER

    def __init__(self, in_channels, hidden_dim, graph_args={}, edge_importance_weighting=True, **kwargs):
        super().__init__()
        if in_channels != 3 or hidden_dim != 64 or not edge_importance_weighting:
            raise MMEgoError("libmmego_b200 implements GCN.Model(3, 64, ..., edge_importance_weighting=True)")
        self.graph = Graph(**graph_args)
        _layout.populate(self, _layout.gcn_layout("", in_channels, hidden_dim))
        # the rest of a Lower_Net state dict is needed by mmego_set_weights(NET_LOWER); keep neutral placeholders
        self._rest = {k: torch.zeros(s) if kind != "counter" else torch.tensor(0)
                      for k, s, kind, _ in _layout.lower_layout(hidden_dim) if not k.startswith("keyEncoder.gcn.")}
        for k in self._rest:
            if k.endswith("running_var"):
                self._rest[k] = torch.ones_like(self._rest[k])

    def _sync(self, device):
        from ..engine import get_handle
        h = get_handle(device)
        key = (id(self), self._weights_key())
        if h.weights_owner.get(self._net_id) != key:         # the NET_LOWER slot may hold a LowerNet's (or another GCN's) set
            sd = dict(self._rest)
            sd.update({"keyEncoder.gcn." + k: v for k, v in self.state_dict().items()})
            h.set_weights(self._net_id, sd)
            h.weights_owner[self._net_id] = key
        return h

    def extract_feature(self, x):
        """x [B, 3, T, 15, 1] -> [B, T, 15, 64] (the raw reinterpretation of the [B,64,T,15] block, Net/GCN.py:352-353)."""
        x = self._cuda_f32(x, "x")
        h = self._sync(x.device)
        return h.gcn_extract_feature(x.contiguous())
