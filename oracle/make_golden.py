"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own classes.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python oracle/make_golden.py [--full-sample]

What it does (SURVEY.md section 8c recipe):
  * stubs matplotlib/seaborn/imageio (absent in this image, imported by Util/Universal_Util/Utils.py),
  * imports Net.IMU_Net / Net.Upper_Net / Net.Lower_Net from /root/reference,
  * loads the shipped Upper/Lower checkpoints with map_location='cpu',
  * IMU_Net checkpoint is missing from the mount -> seeded numpy weights (oracle.synth_imu_state_dict),
  * builds the 835-snippet sample set with np.random.seed(0) (the loader uses the unseeded global RNG),
  * freezes inputs + reference outputs as .npz.

Nothing from the reference's sources is copied; only tensors it computes are stored.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.animation", "seaborn", "imageio",
                 "imageio.v2", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    from Net.IMU_Net import IMUNet
    from Net.Upper_Net import UpperNet
    from Net.Lower_Net import LowerNet
    from Config.config import Config
    return IMUNet, UpperNet, LowerNet, Config


def ref_metrics(pred, upper_l, lower_l, target, Config):
    """Per-batch quantities of Processor/Test/Demo_test.py:64-69,150-163 (module not importable: its
    import of Dataset_action.py hits a SyntaxError), evaluated with the same torch ops."""
    um, lm = Config.upper_joint_map, Config.lower_joint_map
    leaf = [int(v) for v in Config.skeleton_all[:, 1]]
    root = [int(v) for v in Config.skeleton_all[:, 0]]
    accu_upper = torch.sqrt(torch.sum(torch.square(upper_l - target[:, :, um]), dim=-1)).mean().item()
    accu_lower = torch.sqrt(torch.sum(torch.square(lower_l - target[:, :, lm]), dim=-1)).mean().item()
    pv = pred[:, :, leaf] - pred[:, :, root]
    tv = target[:, :, leaf] - target[:, :, root]
    cos = torch.nn.functional.cosine_similarity(pv, tv, dim=-1)
    ang = torch.abs(torch.acos(torch.clamp(cos, min=-1.0, max=1.0)) / 3.14159265358 * 180.0)
    angle_l = ang.mean(0).mean(0).numpy()
    accu_a = torch.sqrt(torch.sum(torch.square(pred - target), dim=-1))
    return accu_a.mean().item(), accu_upper, accu_lower, accu_a.mean(0).mean(0).numpy(), angle_l


def run_reference_chain(up_net, lo_net, Config, data, skl, R, t, bs):
    """Chain exactly as Demo_test.py:106-123 in batches of `bs` snippets (the reference uses bs=1)."""
    outs = {k: [] for k in ("upper_l", "q_upper", "gw", "hn", "cn", "x1", "lower_l", "q_lower", "x2", "pred")}
    with torch.no_grad():
        for s in range(0, data.shape[0], bs):
            d = data[s:s + bs].clone()
            b = d.shape[0]
            h0 = torch.zeros(6, b, 64)
            c0 = torch.zeros(6, b, 64)
            up, qu, gw, hn, cn = up_net(d, h0, c0, skl[s:s + bs], R[s:s + bs], t[s:s + bs])
            outs["x1"].append(d.clone())
            upper_l = up.clone()
            lo, ql = lo_net(upper_l, d, h0, c0, h0, c0, skl[s:s + bs], R[s:s + bs], t[s:s + bs])
            outs["x2"].append(d.clone())
            pred = torch.zeros(b, d.shape[1], 21, 3)
            pred[:, :, Config.upper_joint_map] = upper_l
            pred[:, :, Config.lower_joint_map] = lo
            for k, v in (("upper_l", up), ("q_upper", qu), ("gw", gw.reshape(b, d.shape[1], -1)), ("lower_l", lo),
                         ("q_lower", ql), ("pred", pred)):
                outs[k].append(v)
            outs["hn"].append(hn.permute(1, 0, 2))
            outs["cn"].append(cn.permute(1, 0, 2))
    return {k: torch.cat(v).numpy() for k, v in outs.items()}


SWEEP_SHAPES = [(2, 40, 256), (1, 80, 128), (2, 20, 512)]       # (B, L, N): config 5 of BASELINE.json (N, L up to 4x)
SWEEP_GCN_T = [(2, 40), (1, 80)]


def sweep_pins(up_net, lo_net, Config, O):
    """Pins F/G: the reference's classes accept any L and N (Net/GCN.py:103-117 pads by kernel size only; the point
    encoders are 1x1 convolutions; Net/Lower_Net.py:216-227 keeps Config.lower_pc_no = 64 of N points), so the sweep
    shapes get reference-generated vectors of their own.  One reference call per shape (B = 2: the initial_body[r % B]
    quirk is visible); (R, t) are the synthetic generator's poses."""
    skeleton = np.load(os.path.join(GOLD, "skeleton.npy"))
    for B, L, N in SWEEP_SHAPES:
        sb = O.synth_batch(B, L=L, N=N, n_imu=1, seed=500 + L + N, skeleton=skeleton, distinct_skeletons=True)
        o = run_reference_chain(up_net, lo_net, Config, sb["data"], sb["skl"], sb["R"], sb["t"], B)
        keep = {k: o[k] for k in ("upper_l", "q_upper", "gw", "hn", "cn", "lower_l", "q_lower", "pred")}
        np.savez_compressed(os.path.join(GOLD, f"sweep_L{L}_N{N}.npz"), data=sb["data"].numpy(), skl=sb["skl"].numpy(),
                            R=sb["R"].numpy(), t=sb["t"].numpy(), **keep)
        print("sweep pin", (B, L, N), "upper_l", keep["upper_l"].shape)
    for B, T in SWEEP_GCN_T:
        g = torch.Generator().manual_seed(40 + T)
        xg = torch.randn(B, 3, T, 15, 1, generator=g)
        with torch.no_grad():
            kf = lo_net.keyEncoder.gcn.extract_feature(xg)
        np.savez_compressed(os.path.join(GOLD, f"gcn_T{T}.npz"), x=xg.numpy(), out=kf.numpy())
        print("gcn pin T =", T, tuple(kf.shape))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-sample", action="store_true", help="also freeze all 835 sample snippets (large)")
    ap.add_argument("--sweep-only", action="store_true", help="only (re)generate the non-config-shape pins (F, G)")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    IMUNet, UpperNet, LowerNet, Config = import_reference()
    from oracle import mmego_oracle as O

    up_net, lo_net = UpperNet(), LowerNet(64)
    up_net.load_state_dict(torch.load(Config.model_upper_path, map_location="cpu", weights_only=True))
    lo_net.load_state_dict(torch.load(Config.model_lower_path, map_location="cpu", weights_only=True))
    up_net.eval(), lo_net.eval()

    if args.sweep_only:
        sweep_pins(up_net, lo_net, Config, O)
        return

    # ---- sample data (seeded; F9) --------------------------------------------------------------
    from Util.Universal_Util.Dataset_sample import PosePC
    np.random.seed(0)
    ds = PosePC(train=False, vis=True, batch_length=20)
    data = torch.tensor(np.asarray(ds.data_ti_), dtype=torch.float32)
    target = torch.tensor(np.asarray(ds.data_key_), dtype=torch.float32)
    skl = torch.tensor(np.asarray(ds.skl_), dtype=torch.float32)
    imu = torch.tensor(np.asarray(ds.imu_), dtype=torch.float32)
    R_sur = torch.tensor(np.asarray(ds.R_R0R_), dtype=torch.float32)
    t_sur = target[:, :, 20].clone()
    print("sample set", tuple(data.shape), tuple(imu.shape), tuple(skl.shape))
    np.save(os.path.join(GOLD, "skeleton.npy"), skl[0].numpy())
    assert float((skl - skl[0]).abs().max()) == 0.0

    # ---- pin A: full 835-snippet evaluation with the IMU surrogate, reference batch_size=1 ------
    accs = []
    with torch.no_grad():
        for i in range(data.shape[0]):
            o = run_reference_chain(up_net, lo_net, Config, data[i:i + 1], skl[i:i + 1], R_sur[i:i + 1], t_sur[i:i + 1], 1)
            accs.append(ref_metrics(torch.from_numpy(o["pred"]), torch.from_numpy(o["upper_l"]),
                                    torch.from_numpy(o["lower_l"]), target[i:i + 1], Config))
    pin = dict(mpjpe_cm=np.mean([a[0] for a in accs]) * 100, upper_cm=np.mean([a[1] for a in accs]) * 100,
               lower_cm=np.mean([a[2] for a in accs]) * 100,
               per_joint_cm=np.mean([a[3] for a in accs], axis=0) * 100,
               angle_deg=float(np.mean(np.mean([a[4] for a in accs], axis=0))),
               angle_bone_deg=np.mean([a[4] for a in accs], axis=0))
    print("surrogate pin:", pin["mpjpe_cm"], pin["upper_cm"], pin["lower_cm"], pin["angle_deg"])
    np.savez(os.path.join(GOLD, "sample835_pin.npz"), **pin)

    # ---- pin B: a 16-snippet slice of the real data with every stage output --------------------
    sel = np.linspace(0, data.shape[0] - 1, 16).astype(int)
    o = run_reference_chain(up_net, lo_net, Config, data[sel], skl[sel], R_sur[sel], t_sur[sel], 1)
    np.savez_compressed(os.path.join(GOLD, "sample16.npz"), sel=sel, data=data[sel].numpy(), target=target[sel].numpy(),
                        skl=skl[sel].numpy(), imu=imu[sel].numpy(), R=R_sur[sel].numpy(), t=t_sur[sel].numpy(), **o)

    # ---- pin C: synthetic batch, B=3 in ONE reference call, distinct skeletons (exercises F8) ---
    sb = O.synth_batch(3, seed=77, skeleton=skl[0].numpy(), distinct_skeletons=True)
    o = run_reference_chain(up_net, lo_net, Config, sb["data"], sb["skl"], sb["R"], sb["t"], 3)
    np.savez_compressed(os.path.join(GOLD, "synth3.npz"), data=sb["data"].numpy(), skl=sb["skl"].numpy(),
                        R=sb["R"].numpy(), t=sb["t"].numpy(), **o)

    # ---- pin D: IMU_Net with seeded weights ----------------------------------------------------
    sd_imu = O.synth_imu_state_dict(0)
    imu_net = IMUNet(15, 9, 512, 2, True, 0.1)
    imu_net.load_state_dict(sd_imu)
    imu_net.eval()
    sb = O.synth_batch(2, seed=5)
    with torch.no_grad():
        R1, t1 = imu_net(sb["imu"])
        R2, t2 = imu_net(imu[sel[:2]])
    np.savez_compressed(os.path.join(GOLD, "imu_seed0.npz"), imu_synth=sb["imu"].numpy(), R_synth=R1.numpy(),
                        t_synth=t1.numpy(), imu_real=imu[sel[:2]].numpy(), R_real=R2.numpy(), t_real=t2.numpy())

    # ---- pin E: GCN.Model.extract_feature standalone ------------------------------------------
    g = torch.Generator().manual_seed(3)
    xg = torch.randn(2, 3, 20, 15, 1, generator=g)
    with torch.no_grad():
        kf = lo_net.keyEncoder.gcn.extract_feature(xg)
    np.savez_compressed(os.path.join(GOLD, "gcn2.npz"), x=xg.numpy(), out=kf.numpy())

    sweep_pins(up_net, lo_net, Config, O)

    if args.full_sample:
        np.savez_compressed(os.path.join(ROOT, "Resource", "Sample_data_frozen", "sample835_seed0.npz"),
                            data=data.numpy(), target=target.numpy(), skl=skl.numpy(), imu=imu.numpy(),
                            R_sur=R_sur.numpy(), t_sur=t_sur.numpy())
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
