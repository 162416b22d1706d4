#!/usr/bin/env python
"""One-off .mat -> packed-binary cache of the raw per-frame sensor data.

    python scripts/pack_sample_data.py --root <mmEgo checkout>/Resource/Sample_data --out Resource/Sample_data_packed/raw.npz
    python scripts/pack_sample_data.py --root ... --fixture tests/golden/raw_subset.npz --recordings 3

Walks the recordings exactly as the reference loader does (Util/Universal_Util/Dataset_sample.py:106-151: numeric
order of the action folders, the first recording skipped, frames in numeric file order, empty clouds dropped) and stores
what the loader reads from every .mat file, untouched: the radar points (x, y, z, intensity, velocity; float32), the 21
selected Kinect joints, the raw 20x15 IMU block, R_btc and t_R0R (float64).  Everything the loader COMPUTES from them
(range channel, channel reorder, padding / sub-sampling to 128 slots, IMU re-framing, R_R0R, snippet windows) is done
on the GPU by mmego_build_snippets.

--fixture additionally runs the reference's own PosePC class (imported from the checkout, plotting modules stubbed)
with np.random.seed(0) and stores its tensors for the snippets of the first recordings: the parity pin of the builder.
"""
import argparse
import glob
import os
import re
import sys
import types

import numpy as np
import scipy.io as scio

JOINTS = [0, 1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14, 18, 19, 20, 21, 22, 23, 24, 25, 26]      # Config/config.py:49
SKELETON = [[20, 3], [3, 2], [2, 1], [2, 4], [2, 8], [4, 5], [5, 6], [6, 7], [8, 9], [9, 10], [10, 11], [1, 0], [0, 12],
            [0, 16], [12, 13], [13, 14], [14, 15], [16, 17], [17, 18], [18, 19]]                    # Config/config.py:37-39


def walk(root):
    """Yields (recording index, list of .mat paths) in the loader's order (Dataset_sample.py:120-139)."""
    acts = sorted(os.listdir(root), key=lambda x: int(x))
    rec = 0
    for a, act in enumerate(acts):
        sub = os.path.join(root, act)
        for j, name in enumerate(sorted(os.listdir(sub))):
            path = os.path.join(sub, name)
            if not os.path.isdir(path):
                continue
            regex = re.compile(r"\d+")
            mats = sorted(glob.glob(os.path.join(path, "*.mat")),
                          key=lambda x: [int(y) for y in regex.findall(os.path.basename(x))])
            if not mats or (a == 0 and j == 0):
                continue
            yield rec, mats
            rec += 1


def pack(root, max_recordings=None):
    rec_start, pt_start = [0], [0]
    pts, key, imu, rbtc, t0 = [], [], [], [], []
    ref = None
    for rec, mats in walk(root):
        if max_recordings is not None and rec >= max_recordings:
            break
        for f in mats:
            d = scio.loadmat(f)
            p = np.asarray(d["pc_xyziv_ti2"][:, 0:5], dtype=np.float32)
            if len(p) == 0:
                continue
            k = np.asarray(d["pc_xyz_key_2"][:, 0:3], dtype=np.float64)[JOINTS]
            if ref is None:      # the loader's st == 0 branch (:165-181): global reference pose and bone vectors
                ref = dict(R_ref=np.asarray(d["R_btc"], np.float64), orientation_ref=np.asarray(d["orientation_imu_img"], np.float64),
                           skl=np.asarray([k[a] - k[b] for a, b in SKELETON], np.float64))
            pts.append(p)
            pt_start.append(pt_start[-1] + len(p))
            key.append(k)
            imu.append(np.asarray(d["imu_save_l"], np.float64))
            rbtc.append(np.asarray(d["R_btc"], np.float64))
            t0.append(np.asarray(d["t_R0R"], np.float64).reshape(3))
        rec_start.append(len(key))
    return dict(rec_start=np.asarray(rec_start, np.int64), pt_start=np.asarray(pt_start, np.int64),
                points=np.concatenate(pts, 0), key=np.asarray(key), imu=np.asarray(imu), R_btc=np.asarray(rbtc),
                t_R0R=np.asarray(t0), **ref)


def pack_extras(root, max_recordings=None):
    """The three loader fields the inference path never reads (Dataset_sample.py:182, 195-202): per frame the raw
    foot-contact flags and the ground plane (sign-normalised as the loader does); R_RtW is a product of R_btc (already in
    the main cache) with two constant matrices and is formed on the host.  Kept in a separate small file so that the
    57 MB main cache stays untouched."""
    foot, ground = [], []
    for rec, mats in walk(root):
        if max_recordings is not None and rec >= max_recordings:
            break
        for f in mats:
            d = scio.loadmat(f)
            if len(d["pc_xyziv_ti2"]) == 0:
                continue
            fc = np.asarray(d["foot_contact"])
            foot.append([1 if fc[0, 0] else 0, 1 if fc[0, 1] else 0])
            g = np.asarray(d["abcd_ground_2"], dtype=np.float64)
            if g[0, 0] > 0:
                g = -1 * g
            ground.append(g.reshape(1, 4))
    return dict(foot_contact_raw=np.asarray(foot, np.uint8), ground=np.asarray(ground, np.float64))


def reference_tensors(checkout, n_snippets):
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.animation", "seaborn", "imageio",
              "imageio.v2", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, checkout)
    from Util.Universal_Util.Dataset_sample import PosePC
    np.random.seed(0)
    ds = PosePC(train=False, vis=True, batch_length=20)
    s = slice(0, n_snippets)
    return dict(exp_data=ds.data_ti_[s].astype(np.float32), exp_key=ds.data_key_[s].astype(np.float32),
                exp_imu=ds.imu_[s].astype(np.float32), exp_skl=ds.skl_[s].astype(np.float32),
                exp_R=ds.R_R0R_[s].astype(np.float32), exp_t=ds.t_R0R_[s].reshape(-1, 20, 3).astype(np.float32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", required=True, help="<mmEgo checkout>/Resource/Sample_data")
    ap.add_argument("--out")
    ap.add_argument("--fixture")
    ap.add_argument("--recordings", type=int, default=3)
    ap.add_argument("--extras", help="write the ground / foot-contact side file (Resource/Sample_data_packed/extras.npz)")
    ap.add_argument("--extras-fixture", help="side file for the first --recordings recordings + the reference loader's own fields")
    a = ap.parse_args()
    if a.extras:
        d = pack_extras(a.root)
        np.savez_compressed(a.extras, **d)
        print(a.extras, "frames", len(d["ground"]), os.path.getsize(a.extras), "bytes")
    if a.extras_fixture:
        d = pack_extras(a.root, a.recordings)
        main_d = pack(a.root, a.recordings)
        nsn = int(sum((main_d["rec_start"][i + 1] - main_d["rec_start"][i]) // 20 for i in range(len(main_d["rec_start"]) - 1)))
        for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.animation", "seaborn", "imageio",
                  "imageio.v2", "mpl_toolkits", "mpl_toolkits.mplot3d"):
            sys.modules.setdefault(m, types.ModuleType(m))
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(a.root))))
        from Util.Universal_Util.Dataset_sample import PosePC
        np.random.seed(0)
        ds = PosePC(train=False, vis=True, batch_length=20)
        d.update(exp_ground=np.asarray(ds.ground_[:nsn]), exp_foot_contact=np.asarray(ds.foot_contact_[:nsn]),
                 exp_R_RtW=np.asarray(ds.R_RtW_[:nsn]))
        np.savez_compressed(a.extras_fixture, **d)
        print(a.extras_fixture, "snippets", nsn, os.path.getsize(a.extras_fixture), "bytes")
    if a.out:
        d = pack(a.root)
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        np.savez_compressed(a.out, **d)
        print(a.out, "frames", len(d["key"]), "recordings", len(d["rec_start"]) - 1, "points", len(d["points"]))
    if a.fixture:
        d = pack(a.root, a.recordings)
        nsn = int(sum((d["rec_start"][i + 1] - d["rec_start"][i]) // 20 for i in range(len(d["rec_start"]) - 1)))
        d.update(reference_tensors(os.path.dirname(os.path.dirname(os.path.abspath(a.root))), nsn))
        np.savez_compressed(a.fixture, **d)
        print(a.fixture, "frames", len(d["key"]), "snippets", nsn, os.path.getsize(a.fixture), "bytes")


if __name__ == "__main__":
    main()
