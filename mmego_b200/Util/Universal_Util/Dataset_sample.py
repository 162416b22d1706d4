"""Device-side counterpart of the reference's Util/Universal_Util/Dataset_sample.py (class PosePC).

The reference walks ~19k .mat files and builds every snippet on the host with numpy (12.8 s for the sample set, most of
an evaluation's wall clock).  Here the raw per-frame sensor data are read once from the packed cache written by
scripts/pack_sample_data.py, kept resident on the GPU, and `mmego_build_snippets` turns any list of snippets into the
batch tensors of Dataset_sample.py:73-78 (data_ti, data_key, imu, skl, R_R0R, t_R0R) in one kernel launch.

Host logic kept from the reference: the snippet windows (cut from the end of every recording, Dataset_sample.py:
233-260), the seeded shuffle and the 80/20 train/test split (:37-70).  Ground planes, foot contacts and R_RtW (never read by
the inference path) are served by `__getitem__` from the small side file Resource/Sample_data_packed/extras.npz, on the
host.  The random pad-slot placement is seeded here (the reference uses numpy's
global RNG and is not repeatable from run to run).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from ...Config.config import Config
from ...engine import MMEgoError, get_handle

RAW_ARRAYS = ("points", "pt_start", "key", "imu", "R_btc", "t_R0R", "R_ref", "orientation_ref")
R_TTB = np.array([[0, -1, 0], [-1, 0, 0], [0, 0, -1]], dtype=np.float64)       # Dataset_sample.py:19
R_CTW = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float64)         # Dataset_sample.py:20


def snippet_windows(rec_start: np.ndarray, frame_no: int) -> np.ndarray:
    """First source frame of every snippet in the loader's order (Dataset_sample.py:233-260)."""
    starts = []
    for r in range(len(rec_start) - 1):
        s, e = int(rec_start[r]), int(rec_start[r + 1])
        while e - s >= frame_no:
            starts.append(e - frame_no)
            e -= frame_no
    return np.asarray(starts, dtype=np.int64)


class PosePC:
    """`PosePC(train, vis, batch_length)` as in the reference, backed by the packed cache and the GPU builder.

    len(ds) / ds.starts follow the reference's ordering rules: vis=True keeps recording order; otherwise snippets are
    shuffled with RandomState(Config.dataset_random_seed) and split 80/20 (train / test).  `ds.batch(indices)` builds
    the tensors of those snippets on the device; `ds[i]` returns one snippet's tuple like the reference's __getitem__
    (numpy, via a device round trip -- for compatibility, not for speed)."""

    def __init__(self, train=True, vis=False, batch_length=None, packed_path: Optional[str] = None, device=None,
                 seed: int = 0, lib_handle=None, extras_path: Optional[str] = None):
        self.frame_no = int(batch_length or Config.frame_no)
        self.vis, self.train = bool(vis), bool(train)
        self.pc_no = Config.pc_no
        self.seed = int(seed)
        path = packed_path or Config.sample_packed_path
        z = np.load(path)
        missing = [k for k in RAW_ARRAYS + ("rec_start", "skl") if k not in z]
        if missing:
            raise MMEgoError(f"{path} is not a packed raw cache (missing {missing}); build it with scripts/pack_sample_data.py")
        self.handle = lib_handle or get_handle(device or Config.device)
        dev = self.handle.device
        dt = dict(points=np.float32, pt_start=np.int64)
        self.raw: Dict[str, torch.Tensor] = {k: torch.from_numpy(np.ascontiguousarray(z[k], dtype=dt.get(k, np.float64))).to(dev)
                                             for k in RAW_ARRAYS}
        self.skl_row = torch.from_numpy(z["skl"].astype(np.float32))
        # host-side fields of __getitem__ that the networks never read (ground plane, foot contact, R_RtW)
        import os
        self.R_btc_host = np.ascontiguousarray(z["R_btc"], dtype=np.float64)
        ex = extras_path or os.path.join(os.path.dirname(path), "extras.npz")
        self.extras = dict(np.load(ex)) if os.path.exists(ex) else None
        starts = snippet_windows(z["rec_start"], self.frame_no)
        if not vis:
            np.random.RandomState(Config.dataset_random_seed).shuffle(starts)     # Dataset_sample.py:37-53
            cut = int(len(starts) * 0.8)
            starts = starts[:cut] if train else starts[cut:]                       # :54-70
        self.starts = starts

    def __len__(self):
        return len(self.starts)

    def batch(self, indices=None, slot_src: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        idx = np.arange(len(self.starts)) if indices is None else np.asarray(indices)
        st = torch.from_numpy(np.ascontiguousarray(self.starts[idx])).to(self.handle.device)
        out = self.handle.build_snippets(self.raw, st, slot_src, self.seed, self.frame_no, self.pc_no)
        out["skl"] = self.skl_row.to(self.handle.device).unsqueeze(0).repeat(len(idx), 1, 1).contiguous()
        return out

    def __getitem__(self, index):
        """One snippet in the order of the reference's __getitem__ (Dataset_sample.py:73-94):
        (ti, label, skl, imu, ground, foot_contact, R_R0R, t_R0R[, R_RtW with vis=True]).  ground / foot_contact come from the
        side file written by scripts/pack_sample_data.py --extras (zeros of the right shape if it is absent); R_RtW =
        R_ttb R_btc R_ctw (:182) is formed here from the cached R_btc."""
        o = self.batch([index])
        g = {k: v[0].cpu().numpy() for k, v in o.items()}
        L = self.frame_no
        fr = np.arange(int(self.starts[index]), int(self.starts[index]) + L)
        if self.extras is not None:
            ground = self.extras["ground"][fr]                                       # [L,1,4] float64, sign-normalised (:198-200)
            raw = self.extras["foot_contact_raw"][fr].astype(bool)                   # [L,2]
            foot = np.where(raw[:, :, None], np.array([0, 1]), np.array([1, 0])).astype(np.int64)      # (:195-197)
        else:
            ground, foot = np.zeros((L, 1, 4), np.float64), np.zeros((L, 2, 2), np.int64)
        item = (g["data"], g["key"], g["skl"], g["imu"], ground, foot, g["R"], g["t"].reshape(L, 1, 3))
        if not self.vis:
            return item
        return item + (R_TTB @ self.R_btc_host[fr] @ R_CTW,)
