"""CPU oracle for the mmEgo inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this file.  The product path (``mmego_b200``) never does; it fails loudly when
the CUDA library is missing.

This is an independent restatement (explicit tensor algebra on CPU torch tensors, fp32 or fp64)
of the reference's forward pass.  Every function cites the reference file:line it follows
(paths relative to the reference checkout).  It consumes plain ``state_dict``-style mappings
(``name -> tensor``) with the exact key names of the shipped checkpoints.

Parity pinning: ``oracle/make_golden.py`` runs the *reference's own* ``Net/*.py`` classes (imported
from /root/reference in the build container) on seeded inputs and freezes their outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------
# constants (Config/config.py:16-24, 37-55)
# ----------------------------------------------------------------------------------------------
FRAME_NO = 20
PC_NO = 128
LOWER_PC_NO = 64
JOINT_ALL = 21
JOINT_UPPER = 15
JOINT_LOWER = 8
SKELETON_ALL = [[20, 3], [3, 2], [2, 1], [2, 4], [2, 8], [4, 5], [5, 6], [6, 7], [8, 9], [9, 10], [10, 11],
                [1, 0], [0, 12], [0, 16], [12, 13], [13, 14], [14, 15], [16, 17], [17, 18], [18, 19]]
SKELETON_UPPER = SKELETON_ALL[:14]
SKELETON_LOWER = SKELETON_ALL[14:]
KINECT_UPPER_GRAPH = [(0, 12), (0, 13), (0, 1), (1, 2), (2, 3), (2, 4), (2, 8), (3, 14), (4, 5), (5, 6), (6, 7),
                      (8, 9), (9, 10), (10, 11)]
UPPER_JOINT_MAP = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 20]
LOWER_JOINT_MAP = [12, 13, 14, 15, 16, 17, 18, 19]
LOWER_ROT_MAP = [13, 14, 15, 17, 18, 19]  # Net/Lower_Net.py:29
BN_EPS = 1e-5


def _sd(sd: Mapping[str, Tensor], dtype) -> Dict[str, Tensor]:
    return {k: (v.detach().to("cpu", dtype) if v.is_floating_point() else v.detach().cpu()) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def lstm_direction(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                   h0: Tensor, c0: Tensor, reverse: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """One direction of one nn.LSTM layer, batch_first.  Gate row blocks are i, f, g, o
    (torch.nn.LSTM semantics used at Net/IMU_Net.py:58-62, Net/Upper_Net.py:333, Net/Lower_Net.py:91).
    x [S, T, In] -> y [S, T, H], h_T, c_T."""
    S, T, _ = x.shape
    H = w_hh.shape[1]
    gx = x @ w_ih.t() + (b_ih + b_hh)            # [S, T, 4H]
    h, c = h0, c0
    ys = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = gx[:, t] + h @ w_hh.t()
        i = torch.sigmoid(g[:, 0:H])
        f = torch.sigmoid(g[:, H:2 * H])
        gg = torch.tanh(g[:, 2 * H:3 * H])
        o = torch.sigmoid(g[:, 3 * H:4 * H])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        ys[t] = h
    return torch.stack(ys, dim=1), h, c


def bilstm(x: Tensor, sd: Mapping[str, Tensor], prefix: str, num_layers: int,
           h0: Optional[Tensor] = None, c0: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Multi-layer bidirectional LSTM (eval mode: inter-layer dropout inactive).
    h0/c0: [2*num_layers, S, H] ordered (l0 fwd, l0 bwd, l1 fwd, ...) as in torch.nn.LSTM."""
    S = x.shape[0]
    H = sd[prefix + "weight_hh_l0"].shape[1]
    hn, cn = [], []
    inp = x
    for layer in range(num_layers):
        outs = []
        for d, sfx in enumerate(("", "_reverse")):
            k = f"l{layer}{sfx}"
            idx = 2 * layer + d
            h_init = h0[idx] if h0 is not None else x.new_zeros(S, H)
            c_init = c0[idx] if c0 is not None else x.new_zeros(S, H)
            y, h, c = lstm_direction(inp, sd[prefix + "weight_ih_" + k], sd[prefix + "weight_hh_" + k],
                                     sd[prefix + "bias_ih_" + k], sd[prefix + "bias_hh_" + k],
                                     h_init, c_init, reverse=(d == 1))
            outs.append(y)
            hn.append(h)
            cn.append(c)
        inp = torch.cat(outs, dim=-1)
    return inp, torch.stack(hn), torch.stack(cn)


def ortho6d_to_matrix(a: Tensor, b: Tensor, eps: float) -> Tensor:
    """Gram-Schmidt 6D -> rotation with columns (x, y, z).
    eps=1e-8, max(norm, eps): Net/IMU_Net.py:7-47.  eps=1e-12 (F.normalize): Net/Upper_Net.py:356-363,
    Net/Lower_Net.py:126-133."""
    x = a / a.norm(dim=-1, keepdim=True).clamp_min(eps)
    z = torch.linalg.cross(x, b, dim=-1)
    z = z / z.norm(dim=-1, keepdim=True).clamp_min(eps)
    y = torch.linalg.cross(z, x, dim=-1)
    return torch.stack((x, y, z), dim=-1)


def transform2h(p: Tensor, R: Tensor, t: Tensor) -> Tensor:
    """world -> head frame, R (p - t).  Util/Universal_Util/Utils.py:284-292 (out-of-place here).
    p [F, n, 3], R [F, 3, 3], t [F, 3]."""
    return torch.einsum("fij,fnj->fni", R, p - t[:, None, :])


def transform2r(p: Tensor, R: Tensor, t: Tensor) -> Tensor:
    """head -> world frame, R^T p + t.  Util/Universal_Util/Utils.py:274-281."""
    return torch.einsum("fji,fnj->fni", R, p) + t[:, None, :]


def conv_bn_relu(x: Tensor, sd: Mapping[str, Tensor], conv: str, bn: str) -> Tensor:
    """Conv1d(k=1) + BatchNorm1d(eval) + ReLU on channel-last x [F, n, Cin]
    (Net/Upper_Net.py:261-263, 293-295; Net/Lower_Net.py:65-67)."""
    w = sd[conv + ".weight"][:, :, 0]
    y = x @ w.t() + sd[conv + ".bias"]
    y = (y - sd[bn + ".running_mean"]) / torch.sqrt(sd[bn + ".running_var"] + BN_EPS) * sd[bn + ".weight"] + sd[bn + ".bias"]
    return torch.relu(y)


def bn_affine(sd: Mapping[str, Tensor], bn: str) -> Tuple[Tensor, Tensor]:
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
    return s, sd[bn + ".bias"] - sd[bn + ".running_mean"] * s


# ----------------------------------------------------------------------------------------------
# stage 1: IMUNet.forward (Net/IMU_Net.py:67-94)
# ----------------------------------------------------------------------------------------------
def imu_forward(sd: Mapping[str, Tensor], imu: Tensor, dtype=torch.float32, taps: Optional[dict] = None):
    sd = _sd(sd, dtype)
    imu = imu.detach().to("cpu", dtype)
    B, L, n, _ = imu.shape
    u = torch.relu(imu.reshape(B * L, n, -1) @ sd["fc1.weight"].t() + sd["fc1.bias"])            # :79
    f, _, _ = bilstm(u, sd, "rnn_fast.", 2)                                                       # :80
    a = torch.softmax(f @ sd["attn.weight"].t() + sd["attn.bias"], dim=1)                         # :82
    s = (f * a).sum(dim=1).reshape(B, L, -1)                                                      # :83-84
    g, _, _ = bilstm(s, sd, "rnn_slow.", 2)                                                       # :85
    T = (g @ sd["fc2.weight"].t() + sd["fc2.bias"]).reshape(B * L, -1)                            # :87-88
    R = ortho6d_to_matrix(T[:, 0:3], T[:, 3:6], 1e-8).reshape(B, L, 3, 3)                         # :89-92
    t = T[:, 6:9].reshape(B, L, 3)
    if taps is not None:
        taps.update(u=u, f=f, a=a, s=s, g=g, T=T)
    return R, t


# ----------------------------------------------------------------------------------------------
# stage 2: UpperNet.forward (Net/Upper_Net.py:374-388)
# ----------------------------------------------------------------------------------------------
def forward_kinematics_upper(q: Tensor, body: Tensor, head: Tensor, B: int, L: int,
                             ref_body_index: bool = True, b_offset: int = 0, B_global: Optional[int] = None) -> Tensor:
    """Net/Upper_Net.py:122-144.  q [B*L,14,3,3], body [Bg,20,3], head [B*L,3].
    Flat row r uses body[r % B] (the .repeat(L,1,1,1) quirk, F8) when ref_body_index else body[r // L].
    b_offset/B_global let a shard reproduce the quirk of the unsharded batch."""
    Bg = B_global if B_global is not None else B
    r = torch.arange(B * L) + b_offset * L
    bi = (r % Bg) if ref_body_index else (r // L)
    J = q.new_zeros(B * L, 15, 3)
    J[:, 14] = head
    for i, (p, c) in enumerate(SKELETON_UPPER):
        ci, pi = UPPER_JOINT_MAP.index(c), UPPER_JOINT_MAP.index(p)
        J[:, ci] = J[:, pi] + torch.einsum("fij,fj->fi", q[:, ci], body[bi, i])
    return J


def upper_forward(sd: Mapping[str, Tensor], x: Tensor, h0: Tensor, c0: Tensor, initial_body: Tensor,
                  R: Tensor, t: Tensor, dtype=torch.float32, ref_body_index: bool = True,
                  taps: Optional[dict] = None, b_offset: int = 0, B_global: Optional[int] = None):
    """Returns (l, q, global_weights, hn, cn, x_after) -- x_after is the caller's tensor after the
    in-place Transform2H side effect (F5); the input tensor itself is NOT modified here."""
    sd = _sd(sd, dtype)
    x = x.detach().to("cpu", dtype).clone()
    R = R.detach().to("cpu", dtype)
    t = t.detach().to("cpu", dtype)
    body = initial_body.detach().to("cpu", dtype)
    B, L, N, D = x.shape
    Rf, tf = R.reshape(B * L, 3, 3), t.reshape(B * L, 3)
    xf = x.reshape(B * L, N, D)
    xf[:, :, :3] = transform2h(xf[:, :, :3], Rf, tf)                                              # :379
    p = conv_bn_relu(xf, sd, "module0.conv1", "module0.cb1")                                      # :261
    p = conv_bn_relu(p, sd, "module0.conv2", "module0.cb2")
    p = conv_bn_relu(p, sd, "module0.conv3", "module0.cb3")
    F = torch.cat((xf[:, :, :4], p), dim=-1)                                                      # :266
    s = conv_bn_relu(F, sd, "module1.gpointnet.conv1", "module1.gpointnet.cb1")                   # :293
    s = conv_bn_relu(s, sd, "module1.gpointnet.conv2", "module1.gpointnet.cb2")
    s = conv_bn_relu(s, sd, "module1.gpointnet.conv3", "module1.gpointnet.cb3")
    w = torch.softmax(s @ sd["module1.gpointnet.attn.weight"].t() + sd["module1.gpointnet.attn.bias"], dim=1)  # :299
    g = (s * w).sum(dim=1)                                                                        # :300
    Hs, hn, cn = bilstm(g.reshape(B, L, -1), sd, "module1.grnn.", 3,
                        h0.detach().to("cpu", dtype), c0.detach().to("cpu", dtype))               # :339
    o = torch.relu(Hs @ sd["mlpHead.fc1.weight"].t() + sd["mlpHead.fc1.bias"])                    # :351-353
    o = (o @ sd["mlpHead.fc2.weight"].t() + sd["mlpHead.fc2.bias"]).reshape(B * L, 87)
    q6 = o[:, :84].reshape(B * L, 14, 6)
    q = ortho6d_to_matrix(q6[..., 0:3], q6[..., 3:6], 1e-12)                                      # :355-362
    head = o[:, 84:87]
    J = forward_kinematics_upper(q, body, head, B, L, ref_body_index, b_offset, B_global)         # :385
    l = transform2r(J, Rf, tf).reshape(B, L, 15, 3)                                               # :386
    if taps is not None:
        taps.update(g=g, lstm=Hs, o=o, J=J)
    return l, q.reshape(B, L, 14, 3, 3), w, hn, cn, xf.reshape(B, L, N, D)


# ----------------------------------------------------------------------------------------------
# ST-GCN feature extractor (Net/GCN.py)
# ----------------------------------------------------------------------------------------------
def graph_adjacency() -> np.ndarray:
    """Graph(layout='kinect_upper', strategy='distance', max_hop=1): Net/GCN.py:189-214, 242-278.
    A[0] = normalised self loops, A[1] = normalised 1-hop neighbours, D^-1/2 (I+Adj) D^-1/2."""
    V = 15
    adj = np.zeros((V, V))
    for i, j in KINECT_UPPER_GRAPH:
        adj[i, j] = adj[j, i] = 1
    full = adj + np.eye(V)
    d = full.sum(0) ** -0.5
    norm = (d[:, None] * full) * d[None, :]
    A = np.zeros((2, V, V))
    A[0] = norm * np.eye(V)
    A[1] = norm * adj
    return A


def gcn_extract_feature(sd: Mapping[str, Tensor], x: Tensor, prefix: str = "") -> Tensor:
    """GCN.Model.extract_feature (Net/GCN.py:332-355) on x [B, 3, T, 15, 1] -> [B, T, 15, 64]
    (raw reinterpretation of the contiguous [B, 64, T, 15] block, F6)."""
    B, C, T, V, M = x.shape
    p = prefix
    s, o = bn_affine(sd, p + "data_bn")
    y = x[..., 0].permute(0, 3, 1, 2).reshape(B, V * C, T)                                        # :339-340
    y = y * s[None, :, None] + o[None, :, None]                                                   # :341
    y = y.reshape(B, V, C, T).permute(0, 2, 3, 1)                                                 # [B, C, T, V]  :342-344
    A = sd[p + "A"]
    for i in range(3):
        g = f"{p}gcn_networks.{i}."
        Ai = A * sd[f"{p}edge_importance.{i}"]                                                    # :347
        cw = sd[g + "gcn.conv.weight"][:, :, 0, 0]
        Cout = cw.shape[0] // 2
        z = torch.einsum("oc,bctv->botv", cw, y) + sd[g + "gcn.conv.bias"][None, :, None, None]   # :58
        z = z.reshape(B, 2, Cout, T, V)
        m = torch.einsum("bkctv,kvw->bctw", z, Ai)                                                # :62
        sa, oa = bn_affine(sd, g + "tcn.0")
        u = torch.relu(m * sa[None, :, None, None] + oa[None, :, None, None])                     # :107-108
        tw = sd[g + "tcn.2.weight"][:, :, :, 0]                                                   # [Cout, Cout, 9]
        up = torch.nn.functional.pad(u, (0, 0, 4, 4))                                             # temporal zero pad 4
        tn = sum(torch.einsum("oc,bctv->botv", tw[:, :, k], up[:, :, k:k + T]) for k in range(9))
        tn = tn + sd[g + "tcn.2.bias"][None, :, None, None]                                       # :109-115
        sb, ob = bn_affine(sd, g + "tcn.3")
        tn = tn * sb[None, :, None, None] + ob[None, :, None, None]                               # :116
        rw = sd[g + "residual.0.weight"][:, :, 0, 0]
        res = torch.einsum("oc,bctv->botv", rw, y) + sd[g + "residual.0.bias"][None, :, None, None]
        sr, orr = bn_affine(sd, g + "residual.1")
        res = res * sr[None, :, None, None] + orr[None, :, None, None]                            # :128-136
        y = torch.relu(tn + res)                                                                  # :142-147
    fw = sd[p + "fcn.weight"][:, :, 0, 0]
    e = torch.einsum("oc,bctv->botv", fw, y) + sd[p + "fcn.bias"][None, :, None, None]            # :352
    return e.contiguous().reshape(B, T, V, -1)                                                    # :353 (F6)


# ----------------------------------------------------------------------------------------------
# stage 3: LowerNet.forward (Net/Lower_Net.py:177-239)
# ----------------------------------------------------------------------------------------------
def forward_kinematics_lower(q: Tensor, hip_l: Tensor, hip_r: Tensor, body: Tensor, B: int, L: int,
                             ref_body_index: bool = True, b_offset: int = 0, B_global: Optional[int] = None) -> Tensor:
    """Net/Lower_Net.py:12-37."""
    Bg = B_global if B_global is not None else B
    r = torch.arange(B * L) + b_offset * L
    bi = (r % Bg) if ref_body_index else (r // L)
    J = q.new_zeros(B * L, 8, 3)
    J[:, 0] = hip_l
    J[:, 4] = hip_r
    for i, (p, c) in enumerate(SKELETON_LOWER):
        J[:, LOWER_JOINT_MAP.index(c)] = J[:, LOWER_JOINT_MAP.index(p)] + torch.einsum(
            "fij,fj->fi", q[:, LOWER_ROT_MAP.index(c)], body[bi, i + 14])
    return J


def lower_forward(sd: Mapping[str, Tensor], upper_l: Tensor, x: Tensor, initial_body: Tensor, R: Tensor, t: Tensor,
                  dtype=torch.float32, ref_body_index: bool = True, taps: Optional[dict] = None,
                  b_offset: int = 0, B_global: Optional[int] = None, tie_rule: str = "lowest_slot"):
    """x is the cloud as LEFT BEHIND by UpperNet.forward (already transformed once, F5).
    Returns (l, q, x_after).

    tie_rule: the sample data holds DISTINCT radar points with bit-identical xyz (same range/angle bin,
    different doppler/intensity), so ties in the top-64 key are real and the reference's unstable
    ``torch.sort`` (Net/Lower_Net.py:218) decides which of them survives -- an implementation detail of
    the sort backend (CPU introsort vs CUDA bitonic/radix).  "lowest_slot" is the documented rule of this
    framework (= torch.sort(stable=True)); "torch_cpu" replays torch's CPU unstable sort and reproduces
    the CPU reference exactly."""
    sd = _sd(sd, dtype)
    x = x.detach().to("cpu", dtype).clone()
    upper_l = upper_l.detach().to("cpu", dtype)
    R = R.detach().to("cpu", dtype)
    t = t.detach().to("cpu", dtype)
    body = initial_body.detach().to("cpu", dtype)
    B, L, N, D = x.shape
    F_ = B * L
    Rf, tf = R.reshape(F_, 3, 3), t.reshape(F_, 3)
    xf = x.reshape(F_, N, D)
    xf[:, :, :3] = transform2h(xf[:, :, :3], Rf, tf)                                              # :191-192 (second transform)
    # top-64 by transformed x, descending (:216-227)
    if tie_rule == "torch_cpu":
        idx = torch.sort(xf[:, :, 0].float(), dim=1, descending=True).indices[:, :LOWER_PC_NO]
    else:
        idx = torch.sort(xf[:, :, 0], dim=1, descending=True, stable=True).indices[:, :LOWER_PC_NO]
    sel = torch.gather(xf, 1, idx[:, :, None].expand(-1, -1, D))
    uh = transform2h(upper_l.reshape(F_, 15, 3), Rf, tf)                                          # :229
    p = conv_bn_relu(sel, sd, "pointEncoder.module0.conv1", "pointEncoder.module0.cb1")           # :65-67
    p = conv_bn_relu(p, sd, "pointEncoder.module0.conv2", "pointEncoder.module0.cb2")
    p = conv_bn_relu(p, sd, "pointEncoder.module0.conv3", "pointEncoder.module0.cb3")
    P = torch.cat((sel[:, :, :3], p), dim=-1)                                                     # :70  [F,64,64]
    gx = uh.reshape(B, L, 15, 3).permute(0, 3, 1, 2).unsqueeze(-1)                                # :161-162
    K = gcn_extract_feature(sd, gx, "keyEncoder.gcn.").reshape(F_, 15, 64)                        # :163-164
    tq = P @ sd["fusion.to_q.weight"].t() + sd["fusion.to_q.bias"]                                # :104-106
    tk = K @ sd["fusion.to_k.weight"].t() + sd["fusion.to_k.bias"]
    tv = K @ sd["fusion.to_v.weight"].t() + sd["fusion.to_v.bias"]
    att = torch.softmax(tq @ tk.transpose(-2, -1) * (64 ** -0.5), dim=-1)                         # :107-108
    tx = att @ tv                                                                                 # :109
    a = torch.cat((P, tx), dim=-1).sum(dim=1)            # softmax over a size-1 dim == 1 (F7)    # :111-113
    kbar = K.mean(dim=1)                                                                          # :114-115
    ak = torch.cat((a, kbar), dim=-1).reshape(B, L, 192)                                          # :116
    V, _, _ = bilstm(ak, sd, "fusion.rnn_pk.", 3)                                                 # :117
    o = torch.cat((V, uh.reshape(B, L, 45)), dim=-1)                                              # :119
    o = torch.relu(o @ sd["fusion.fc0.weight"].t() + sd["fusion.fc0.bias"])
    o = torch.relu(o @ sd["fusion.fc1.weight"].t() + sd["fusion.fc1.bias"])
    o = (o @ sd["fusion.fc2.weight"].t() + sd["fusion.fc2.bias"]).reshape(F_, 42)                 # :120-124
    q6 = o[:, :36].reshape(F_, 6, 6)
    q = ortho6d_to_matrix(q6[..., 0:3], q6[..., 3:6], 1e-12)                                      # :126-133
    hip_l, hip_r = o[:, 36:39], o[:, 39:42]                                                       # :134-135
    J = forward_kinematics_lower(q, hip_l, hip_r, body, B, L, ref_body_index, b_offset, B_global) # :235
    l = transform2r(J, Rf, tf).reshape(B, L, 8, 3)                                                # :237
    if taps is not None:
        taps.update(idx=idx, P=P, uh=uh, K=K, ak=ak, lstm=V, o=o, J=J)
    return l, q.reshape(B, L, 6, 3, 3), xf.reshape(B, L, N, D)


# ----------------------------------------------------------------------------------------------
# assembly + metrics (Processor/Test/Demo_test.py:64-69, 121-123, 150-180)
# ----------------------------------------------------------------------------------------------
def assemble(upper_l: Tensor, lower_l: Tensor) -> Tensor:
    B, L = upper_l.shape[:2]
    pred = upper_l.new_zeros(B, L, JOINT_ALL, 3)
    pred[:, :, UPPER_JOINT_MAP] = upper_l
    pred[:, :, LOWER_JOINT_MAP] = lower_l          # lower wins on joints 12, 16
    return pred


def metric_sums(pred: Tensor, upper_l: Tensor, lower_l: Tensor, target: Tensor) -> Dict[str, np.ndarray]:
    """Sums (float64) from which every printed line of Demo_test.py:176-180 follows.  All batches of the
    reference loop are equally sized, so its mean-of-batch-means equals the global mean."""
    pred, upper_l, lower_l, target = (v.detach().to("cpu", torch.float32) for v in (pred, upper_l, lower_l, target))
    e = (pred - target).square().sum(-1).sqrt()                                                    # :157
    eu = (upper_l - target[:, :, UPPER_JOINT_MAP]).square().sum(-1).sqrt()                         # :150
    el = (lower_l - target[:, :, LOWER_JOINT_MAP]).square().sum(-1).sqrt()                         # :153
    root = [b[0] for b in SKELETON_ALL]
    leaf = [b[1] for b in SKELETON_ALL]
    pv = pred[:, :, leaf] - pred[:, :, root]
    tv = target[:, :, leaf] - target[:, :, root]
    cos = torch.nn.functional.cosine_similarity(pv, tv, dim=-1)
    ang = torch.abs(torch.acos(torch.clamp(cos, -1.0, 1.0)) / 3.14159265358 * 180.0)               # :64-69
    F_ = pred.shape[0] * pred.shape[1]
    return dict(frames=np.float64(F_),
                err_joint=e.double().sum((0, 1)).numpy(),         # [21]
                err_upper=np.float64(eu.double().sum()),
                err_lower=np.float64(el.double().sum()),
                angle_bone=ang.double().sum((0, 1)).numpy())      # [20]


def report_from_sums(s: Mapping[str, np.ndarray]) -> Dict[str, object]:
    F_ = float(s["frames"])
    return dict(mpjpe_cm=float(s["err_joint"].sum() / (F_ * JOINT_ALL) * 100.0),
                upper_cm=float(s["err_upper"] / (F_ * JOINT_UPPER) * 100.0),
                lower_cm=float(s["err_lower"] / (F_ * JOINT_LOWER) * 100.0),
                angle_deg=float((s["angle_bone"] / F_).mean()),
                per_joint_cm=(s["err_joint"] / F_ * 100.0))


# ----------------------------------------------------------------------------------------------
# full pipeline as chained at Processor/Test/Demo_test.py:111-123
# ----------------------------------------------------------------------------------------------
def pipeline(sd_imu, sd_upper, sd_lower, imu: Tensor, data: Tensor, skl: Tensor, dtype=torch.float32,
             R_t: Optional[Tuple[Tensor, Tensor]] = None, ref_body_index: bool = True):
    B = data.shape[0]
    if R_t is None:
        R, t = imu_forward(sd_imu, imu, dtype)
    else:
        R, t = R_t
    h0 = torch.zeros(6, B, 64)
    c0 = torch.zeros(6, B, 64)
    up, qu, gw, hn, cn, x1 = upper_forward(sd_upper, data, h0, c0, skl, R, t, dtype, ref_body_index)
    lo, ql, x2 = lower_forward(sd_lower, up, x1, skl, R, t, dtype, ref_body_index)
    pred = assemble(up, lo)
    return dict(R=R, t=t, upper_l=up, q_upper=qu, global_weights=gw, hn=hn, cn=cn, x_after_upper=x1,
                lower_l=lo, q_lower=ql, x_after_lower=x2, pred=pred)


# ----------------------------------------------------------------------------------------------
# deterministic synthetic inputs / weights shared by tests, smoke() and bench.py
# ----------------------------------------------------------------------------------------------
def imu_state_dict_shapes(input_n=15, output_n=9, hidden=512) -> Dict[str, Tuple[int, ...]]:
    """State-dict layout of IMUNet(15, 9, 512, 2, True, .) (Net/IMU_Net.py:51-65), in module order."""
    H = hidden
    shp: Dict[str, Tuple[int, ...]] = {}
    shp["fc1.weight"], shp["fc1.bias"] = (H, input_n), (H,)
    shp["fc2.weight"], shp["fc2.bias"] = (output_n, 2 * H), (output_n,)
    shp["fc3.weight"], shp["fc3.bias"] = (3, output_n), (3,)
    for name, in0 in (("rnn_fast", H), ("rnn_slow", 2 * H)):
        for layer in range(2):
            for sfx in ("", "_reverse"):
                k = f"l{layer}{sfx}"
                shp[f"{name}.weight_ih_{k}"] = (4 * H, in0 if layer == 0 else 2 * H)
                shp[f"{name}.weight_hh_{k}"] = (4 * H, H)
                shp[f"{name}.bias_ih_{k}"] = (4 * H,)
                shp[f"{name}.bias_hh_{k}"] = (4 * H,)
    shp["attn.weight"], shp["attn.bias"] = (1, 2 * H), (1,)
    return shp


def synth_imu_state_dict(seed: int = 0, hidden: int = 512) -> Dict[str, Tensor]:
    """Seeded stand-in for the IMU_Net checkpoint that is missing from the reference mount
    (.MISSING_LARGE_BLOBS).  numpy PCG64 streams are version-stable, so the GPU box regenerates the
    same weights.  Scale follows torch's default U(-1/sqrt(fan), 1/sqrt(fan)) init."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, shape in imu_state_dict_shapes(hidden=hidden).items():
        if k.startswith("rnn"):
            bound = 1.0 / math.sqrt(hidden)
        else:
            base = k.rsplit(".", 1)[0] + ".weight"
            bound = 1.0 / math.sqrt(imu_state_dict_shapes(hidden=hidden)[base][1])
        out[k] = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))
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


# The single calibration skeleton of Resource/Sample_data (bone vectors parent - child in skeleton_all order,
# Util/Universal_Util/Dataset_sample.py:167-169); frozen by oracle/make_golden.py into tests/golden/skeleton.npy.
def synth_batch(B: int, L: int = FRAME_NO, N: int = PC_NO, n_imu: int = 20, seed: int = 1234,
                skeleton: Optional[np.ndarray] = None, distinct_skeletons: bool = False) -> Dict[str, Tensor]:
    """Synthetic radar/IMU snippets matched to the sample-data statistics."""
    rng = np.random.default_rng(seed)
    F_ = B * L
    # IMU: ch 0-8 row-major rotation, 9-11 ~ N(mu, sd), 12-14 ~ N(0, sd)
    Rm = _random_rotations(rng, F_ * n_imu, 15.0).reshape(F_ * n_imu, 9)
    gyr = rng.normal([-4.1, -1.2, -1.1], [1.5, 1.0, 1.4], size=(F_ * n_imu, 3))
    acc = rng.normal(0.0, [0.29, 0.33, 0.61], size=(F_ * n_imu, 3))
    imu = np.concatenate([Rm, gyr, acc], axis=1).reshape(B, L, n_imu, 15).astype(np.float32)
    # radar cloud: n_valid points in random slots, the rest exact zeros
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
    # plausible head pose (for runs that bypass IMU_Net)
    R = _random_rotations(rng, F_, 15.0).reshape(B, L, 3, 3).astype(np.float32)
    t = rng.normal([0.0, 0.0, 0.6], [0.1, 0.1, 0.05], size=(B, L, 3)).astype(np.float32)
    return dict(imu=torch.from_numpy(imu), data=torch.from_numpy(data), skl=torch.from_numpy(skl),
                R=torch.from_numpy(R), t=torch.from_numpy(t))


def default_skeleton() -> np.ndarray:
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "skeleton.npy")
    return np.load(p)


# ------------------------------------------------------------------------------------------------ snippet builder
# CPU restatement of what the reference's loader COMPUTES per frame (Util/Universal_Util/Dataset_sample.py:153-231) and
# of its snippet windows (:233-260), on the packed raw cache written by scripts/pack_sample_data.py.  TEST
# INFRASTRUCTURE like the rest of this file.  Pinned by tests/test_oracle_golden.py against tensors produced by the
# reference's own PosePC class (tests/golden/raw_subset.npz).
R_RI = np.array([[0, 0, 1], [0, -1, 0], [1, 0, 0]], dtype=np.float64)        # Dataset_sample.py:18
R_TTB = np.array([[0, -1, 0], [-1, 0, 0], [0, 0, -1]], dtype=np.float64)     # Dataset_sample.py:19


def snippet_windows(rec_start: np.ndarray, frame_no: int = FRAME_NO) -> np.ndarray:
    """First-frame index of every snippet, in the loader's order: recording by recording, windows cut from the END of
    the recording backwards (Dataset_sample.py:233-260)."""
    starts = []
    for r in range(len(rec_start) - 1):
        s, e = int(rec_start[r]), int(rec_start[r + 1])
        while e - s >= frame_no:
            starts.append(e - frame_no)
            e -= frame_no
    return np.asarray(starts, dtype=np.int64)


def slot_hash(seed: int, frame: int, i: int) -> int:
    """Counter-based 32-bit hash used for the random slot placement (the reference uses the unseeded global numpy RNG,
    Dataset_sample.py:215-223, which cannot be reproduced; any uniformly random injective placement is equivalent)."""
    x = (seed * 0x9E3779B1 + frame * 0x85EBCA77 + i * 0xC2B2AE3D + 0x27D4EB2F) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
    x ^= x >> 12
    x = (x * 0x297A2D39) & 0xFFFFFFFF
    x ^= x >> 15
    return x


def slot_assignment(n: int, pc_no: int, seed: int, frame: int) -> np.ndarray:
    """slot -> source point (or -1).  n < pc_no: every point lands in a distinct random slot; n >= pc_no: pc_no
    distinct random points.  Keys are ranked ascending, ties by index."""
    src = np.full(pc_no, -1, dtype=np.int32)
    if n < pc_no:
        keys = [(slot_hash(seed, frame, s), s) for s in range(pc_no)]
        order = [s for _, s in sorted(keys)]
        for i in range(n):
            src[order[i]] = i
    else:
        keys = [(slot_hash(seed, frame, p), p) for p in range(n)]
        order = [p for _, p in sorted(keys)]
        src[:] = order[:pc_no]
    return src


def build_frame(raw: Mapping[str, np.ndarray], f: int, slot_src: np.ndarray, pc_no: int = PC_NO):
    """One frame of the loader: radar cloud [pc_no,6] (x,y,z,range,velocity,intensity), re-framed IMU block [20,15],
    R_R0R [3,3] -- all float32 as Demo_test.py:95-109 casts them."""
    p = raw["points"][int(raw["pt_start"][f]):int(raw["pt_start"][f + 1])].astype(np.float64)
    xyzrvi = np.zeros((len(p), 6), dtype=np.float32)
    xyzrvi[:, 0:3] = p[:, :3]
    xyzrvi[:, 3] = np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) + p[:, 2] * p[:, 2])      # :204
    xyzrvi[:, 4] = p[:, 4]                                                                   # :208 ([4:2:-1] = v, i)
    xyzrvi[:, 5] = p[:, 3]
    cloud = np.zeros((pc_no, 6), dtype=np.float32)
    live = slot_src >= 0
    cloud[live] = xyzrvi[slot_src[live]]
    imu = raw["imu"][f].copy()
    R_NI = np.stack([imu[:, :3], imu[:, 3:6], imu[:, 6:9]], axis=2)                           # :186
    m = R_RI @ (raw["orientation_ref"].T @ R_NI) @ R_RI.T                                    # :187-188
    imu[:, :3], imu[:, 3:6], imu[:, 6:9] = m[:, 0, :], m[:, 1, :], m[:, 2, :]
    imu[:, 11] = imu[:, 11] + 9.8                                                            # :192
    imu[:, 10:12] = -1 * imu[:, 10:12]
    imu[:, 13:] = -1 * imu[:, 13:]
    R = R_TTB @ raw["R_ref"] @ raw["R_btc"][f].T @ R_TTB.T                                   # :182
    return cloud, imu.astype(np.float32), R.astype(np.float32)


def build_snippets(raw: Mapping[str, np.ndarray], starts: np.ndarray, slot_src: Optional[np.ndarray] = None,
                   seed: int = 0, frame_no: int = FRAME_NO, pc_no: int = PC_NO) -> Dict[str, np.ndarray]:
    B = len(starts)
    out = dict(data=np.zeros((B, frame_no, pc_no, 6), np.float32), imu=np.zeros((B, frame_no, 20, 15), np.float32),
               key=np.zeros((B, frame_no, 21, 3), np.float32), R=np.zeros((B, frame_no, 3, 3), np.float32),
               t=np.zeros((B, frame_no, 3), np.float32), skl=np.repeat(raw["skl"][None].astype(np.float32), B, 0))
    for b in range(B):
        for l in range(frame_no):
            f = int(starts[b]) + l
            n = int(raw["pt_start"][f + 1] - raw["pt_start"][f])
            src = slot_src[b, l] if slot_src is not None else slot_assignment(n, pc_no, seed, f)
            out["data"][b, l], out["imu"][b, l], out["R"][b, l] = build_frame(raw, f, src, pc_no)
            out["key"][b, l] = raw["key"][f]
            out["t"][b, l] = raw["t_R0R"][f]
    return out


def recover_slots(raw: Mapping[str, np.ndarray], starts: np.ndarray, data: np.ndarray) -> np.ndarray:
    """slot -> source point of reference-built clouds (matches xyz + intensity + velocity exactly): lets a test feed the
    reference's own random placement to the builder."""
    B, Lf, pc_no, _ = data.shape
    out = np.full((B, Lf, pc_no), -1, dtype=np.int32)
    for b in range(B):
        for l in range(Lf):
            f = int(starts[b]) + l
            p = raw["points"][int(raw["pt_start"][f]):int(raw["pt_start"][f + 1])]
            used = np.zeros(len(p), bool)
            for s in range(pc_no):
                row = data[b, l, s]
                if not row.any():
                    continue
                hit = np.nonzero((p[:, 0] == row[0]) & (p[:, 1] == row[1]) & (p[:, 2] == row[2]) & (p[:, 4] == row[4]) &
                                 (p[:, 3] == row[5]) & ~used)[0]
                if len(hit) == 0:
                    raise ValueError(f"snippet {b} frame {l} slot {s}: no raw point matches")
                out[b, l, s] = hit[0]
                used[hit[0]] = True
    return out
