"""State-dict layouts of the three networks (the "Pretrained_model checkpoint layout" contract, SURVEY.md
Appendix A) and the nn.Module scaffolding that exposes them.

The drop-in classes in this package own *parameters only*: a tree of bare ``nn.Module`` containers whose
``state_dict()`` has exactly the key names, shapes, dtypes and order of the reference's modules, so the shipped
``.pth`` files load with ``strict=True`` and ``save()`` writes files the reference can read.  There is no
PyTorch compute graph behind them -- ``forward`` hands raw device pointers to libmmego_b200.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Tuple

import torch
from torch import nn

# entry = (dotted name, shape, kind, init) ; kind in {"param", "buffer", "counter"}
Entry = Tuple[str, Tuple[int, ...], str, object]


def linear(name: str, out_f: int, in_f: int) -> List[Entry]:
    b = 1.0 / math.sqrt(in_f)
    return [(f"{name}.weight", (out_f, in_f), "param", b), (f"{name}.bias", (out_f,), "param", b)]


def conv(name: str, out_c: int, in_c: int, *kernel: int) -> List[Entry]:
    fan = in_c
    for k in kernel:
        fan *= k
    b = 1.0 / math.sqrt(fan)
    return [(f"{name}.weight", (out_c, in_c, *kernel), "param", b), (f"{name}.bias", (out_c,), "param", b)]


def batchnorm(name: str, c: int) -> List[Entry]:
    return [(f"{name}.weight", (c,), "param", "ones"), (f"{name}.bias", (c,), "param", "zeros"),
            (f"{name}.running_mean", (c,), "buffer", "zeros"), (f"{name}.running_var", (c,), "buffer", "ones"),
            (f"{name}.num_batches_tracked", (), "counter", 0)]


def lstm(name: str, in_f: int, hidden: int, layers: int, bidirectional: bool = True) -> List[Entry]:
    """torch.nn.LSTM parameter naming/order: per layer, per direction: weight_ih, weight_hh, bias_ih, bias_hh;
    gate row blocks i, f, g, o."""
    out: List[Entry] = []
    b = 1.0 / math.sqrt(hidden)
    dirs = ("", "_reverse") if bidirectional else ("",)
    for l in range(layers):
        k = in_f if l == 0 else hidden * len(dirs)
        for sfx in dirs:
            out += [(f"{name}.weight_ih_l{l}{sfx}", (4 * hidden, k), "param", b),
                    (f"{name}.weight_hh_l{l}{sfx}", (4 * hidden, hidden), "param", b),
                    (f"{name}.bias_ih_l{l}{sfx}", (4 * hidden,), "param", b),
                    (f"{name}.bias_hh_l{l}{sfx}", (4 * hidden,), "param", b)]
    return out


def imu_layout(input_n: int, output_n: int, hidden_n: int, n_rnn_layer: int, bidirectional: bool) -> List[Entry]:
    """IMUNet (Net/IMU_Net.py:51-65 of the reference): fc1, fc2, fc3, rnn_fast, rnn_slow, attn."""
    d = 2 if bidirectional else 1
    return (linear("fc1", hidden_n, input_n) + linear("fc2", output_n, hidden_n * d) + linear("fc3", 3, output_n)
            + lstm("rnn_fast", hidden_n, hidden_n, n_rnn_layer, bidirectional)
            + lstm("rnn_slow", hidden_n * d, hidden_n, n_rnn_layer, bidirectional)
            + linear("attn", 1, hidden_n * d))


def upper_layout() -> List[Entry]:
    """UpperNet (Net/Upper_Net.py:367-372): module0 = PointNet, module1 = GlobalModule, mlpHead."""
    e: List[Entry] = []
    for i, (ci, co) in enumerate(((6, 8), (8, 16), (16, 24)), 1):
        e += conv(f"module0.conv{i}", co, ci, 1) + batchnorm(f"module0.cb{i}", co)
    for i, (ci, co) in enumerate(((28, 32), (32, 48), (48, 64)), 1):
        e += conv(f"module1.gpointnet.conv{i}", co, ci, 1) + batchnorm(f"module1.gpointnet.cb{i}", co)
    e += linear("module1.gpointnet.attn", 1, 64)
    e += lstm("module1.grnn", 64, 64, 3)
    e += linear("mlpHead.fc1", 128, 128) + linear("mlpHead.fc2", 87, 128)
    return e


def gcn_layout(prefix: str, in_channels: int, hidden_dim: int) -> List[Entry]:
    """GCN.Model (Net/GCN.py:301-330): A, data_bn, 3 st_gcn blocks, edge_importance, fcn."""
    p = prefix
    e: List[Entry] = [(f"{p}A", (2, 15, 15), "buffer", "graph")]
    e += batchnorm(f"{p}data_bn", in_channels * 15)
    cin = in_channels
    for i, co in enumerate((32, 64, 128)):
        g = f"{p}gcn_networks.{i}."
        e += conv(g + "gcn.conv", 2 * co, cin, 1, 1) + batchnorm(g + "tcn.0", co) + conv(g + "tcn.2", co, co, 9, 1)
        e += batchnorm(g + "tcn.3", co) + conv(g + "residual.0", co, cin, 1, 1) + batchnorm(g + "residual.1", co)
        cin = co
    e += [(f"{p}edge_importance.{i}", (2, 15, 15), "param", "ones") for i in range(3)]
    e += conv(f"{p}fcn", hidden_dim, 128, 1, 1)
    return e


def lower_layout(hidden_dim: int) -> List[Entry]:
    """LowerNet(hidden_dim) (Net/Lower_Net.py:170-176): pointEncoder, keyEncoder, fusion."""
    e: List[Entry] = []
    for i, (ci, co) in enumerate(((6, 16), (16, 32), (32, hidden_dim - 3)), 1):
        e += conv(f"pointEncoder.module0.conv{i}", co, ci, 1) + batchnorm(f"pointEncoder.module0.cb{i}", co)
    e += gcn_layout("keyEncoder.gcn.", 3, hidden_dim)
    h = hidden_dim
    e += linear("fusion.fc0", 128, 2 * h + 45) + linear("fusion.fc1", 64, 128)
    e += linear("fusion.to_q", h, h) + linear("fusion.to_k", h, h) + linear("fusion.to_v", h, h)
    e += linear("fusion.fc2", 42, 64) + linear("fusion.attn", 1, 2 * h)
    e += lstm("fusion.rnn_pk", 3 * h, h, 3)
    return e


def graph_adjacency() -> torch.Tensor:
    """A [2,15,15] of Graph(layout='kinect_upper', strategy='distance', max_hop=1) (Net/GCN.py:189-214, 270-278):
    A[0] = self loops, A[1] = 1-hop neighbours, both taken from D^-1/2 (I + Adj) D^-1/2."""
    from ..Config.config import Config
    V = 15
    adj = torch.zeros(V, V, dtype=torch.float64)
    for i, j in Config.kinect_upper_gragh:
        adj[i, j] = adj[j, i] = 1.0
    full = adj + torch.eye(V, dtype=torch.float64)
    d = full.sum(0).pow(-0.5)
    norm = d[:, None] * full * d[None, :]
    return torch.stack((norm * torch.eye(V, dtype=torch.float64), norm * adj)).float()


class _Node(nn.Module):
    """Bare container; children/parameters are attached by name."""


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, kind: str):
    *path, leaf = dotted.split(".")
    node = root
    for part in path:
        nxt = node._modules.get(part)
        if nxt is None:
            nxt = _Node()
            node.add_module(part, nxt)
        node = nxt
    if kind == "param":
        node.register_parameter(leaf, nn.Parameter(tensor, requires_grad=False))
    else:
        node.register_buffer(leaf, tensor)


def populate(root: nn.Module, layout: Iterable[Entry]):
    for name, shape, kind, init in layout:
        if kind == "counter":
            t = torch.tensor(int(init), dtype=torch.long)
        elif init == "ones":
            t = torch.ones(shape)
        elif init == "zeros":
            t = torch.zeros(shape)
        elif init == "graph":
            t = graph_adjacency()
        else:
            t = torch.empty(shape).uniform_(-float(init), float(init))
        _attach(root, name, t, kind)
