"""Configs 2 and 5 of BASELINE.json on one B200: IMU_Net standalone at B=4096, and the point-count / sequence-length
sweep (N up to 4x, L up to 4x the Config/config.py shape) for the point encoders and the LSTM kernels, each with its
algorithmic work and the fraction of the measured roofline.  Writes gpurun_out/sweep.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mmego_b200 import _capi, synth
from mmego_b200.pipeline import MMEgoPipeline

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else dict(hbm_gbs=6650.0, bf16_tflops_sustained=1400.0)
H = 512


def imu_flops(frames, n_imu, L):
    fast = 2 * 2 * (4 * H * (H + H) + 4 * H * (2 * H + H)) * n_imu            # per frame
    slow = 2 * 2 * 2 * 4 * H * (2 * H + H)                                     # per frame
    return frames * (fast + slow + 2 * 15 * H * n_imu + 2 * 9 * 2 * H)


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    pipe = MMEgoPipeline(dev)
    h = pipe.handle
    h.set_option("imu_resident", 0)          # the sweep measures the throughput kernels
    out = {"peaks": {"hbm_gbs": PEAKS["hbm_gbs"], "bf16_tflops_sustained": PEAKS["bf16_tflops_sustained"]}, "config2": [],
           "config5": []}
    # ---- config 2: IMU_Net standalone, B=4096
    sb = synth.batch(4096, seed=1234)
    imu = sb["imu"].to(dev)
    for mode, name in ((1, "fp16x3 (fp32-grade)"), (2, "fp16")):
        h.set_option("imu_gemm", mode)
        ms = timed(lambda: h.imu_forward(imu))
        fl = imu_flops(4096 * 20, 20, 20)
        out["config2"].append({"mode": name, "B": 4096, "ms": ms, "frames_per_s": 4096 * 20 / ms * 1e3,
                               "algorithmic_tflops": fl / ms / 1e9,
                               "frac_of_measured_bf16_sustained": fl / ms / 1e9 / PEAKS["bf16_tflops_sustained"]})
    h.set_option("imu_gemm", 1)
    del imu, sb
    # ---- config 5: sweep, ~20k frames per point
    for L in (20, 40, 80):
        for N in (128, 256, 512):
            B = 20480 // L
            sb = synth.batch(B, L=L, N=N, seed=7)
            data0 = sb["data"].to(dev)
            skl, R, t = sb["skl"].to(dev), sb["R"].to(dev), sb["t"].to(dev)
            h0 = torch.zeros(6, B, 64, device=dev)
            imu = sb["imu"].to(dev)
            up = h.upper_forward(data0.clone(), h0, h0.clone(), skl, R, t)[0]
            x = data0.clone()

            def step():
                x.copy_(data0)
                u = h.upper_forward(x, h0, h0, skl, R, t, want_q=False, want_weights=True, want_state=False)[0]
                h.lower_forward(u, x, skl, R, t, want_q=False)
                h.imu_forward(imu)

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            h.profile_begin()
            reps = 3
            for _ in range(reps):
                step()
            prof = h.profile_read()
            h.profile_end()
            F = B * L
            row = {"L": L, "N": N, "B": B, "frames": F, "ms": {k: v["ms"] / reps for k, v in prof.items()}}
            pt = row["ms"]["upper.point"]
            # upper point encoder: 12,256 FLOP/point (six 1x1 convs + attention score), 5.4 KB/frame at N=128
            row["upper_point"] = {"gflops": F * N * 12256 / pt / 1e6, "bytes_per_frame": N * (24 + 12 + 4) + 256 + 48,
                                  "gbs": F * (N * 40 + 304) / pt / 1e6,
                                  "frac_hbm": F * (N * 40 + 304) / pt / 1e6 / PEAKS["hbm_gbs"]}
            fr = row["ms"]["lower.frame"]
            row["lower_frame"] = {"gbs": F * (N * 36 + 15 * 64 * 4 + 192 * 4 + 48) / fr / 1e6}
            lst = row["ms"]["imu.lstm_fast"] + row["ms"]["imu.lstm_slow"]
            fl = imu_flops(F, 20, L)
            row["imu_lstm"] = {"algorithmic_tflops": fl / lst / 1e9,
                               "frac_of_measured_bf16_sustained": fl / lst / 1e9 / PEAKS["bf16_tflops_sustained"]}
            sm = row["ms"]["small_lstm"]
            row["small_lstm"] = {"gflops": F * (524288 + 655360) / sm / 1e6}
            # ---- roofline fractions against the measured peaks (MEASURED_PEAKS.json; mma.sync peak from
            # profiles/r01_ubench_mma_sync_rate.txt: 557 TFLOP/s f16 on this pool's B200).  "issued" = 3 fp16 MMA passes per
            # algorithmic multiply (fp32-grade split products).
            MMA_SYNC_PEAK = 557.0
            bf16 = PEAKS["bf16_tflops_sustained"]

            def tensor(name, alg_flops, ms_, passes=3, peak=MMA_SYNC_PEAK, peak_name="mma.sync f16 measured 557 TFLOP/s"):
                alg = alg_flops / ms_ / 1e9
                return {"stage": name, "bound": "tensor", "algorithmic_tflops": round(alg, 2), "issued_tflops": round(alg * passes, 2),
                        "peak": peak, "peak_name": peak_name, "frac_issued": round(alg * passes / peak, 4),
                        "frac_algorithmic": round(alg / peak, 4), "frac_of_bf16_sustained": round(alg / bf16, 4)}

            def hbm(name, nbytes, ms_):
                g = nbytes / ms_ / 1e6
                return {"stage": name, "bound": "hbm", "gbs": round(g, 1), "peak": PEAKS["hbm_gbs"], "frac": round(g / PEAKS["hbm_gbs"], 4)}

            gcn = row["ms"]["lower.gcn"]
            row["rooflines"] = [
                tensor("upper.point (mma.sync, 12,256 FLOP/point)", F * N * 12256, pt),
                hbm("upper.point (cloud in, xyz back, weights out, g)", F * (N * 40 + 304), pt),
                tensor("lower.frame (mma.sync, 1.40 MFLOP/frame, top-64 of N)", F * 1.40e6, fr),
                hbm("lower.frame (cloud in, xyz back, K in, ak out)", F * (N * 36 + 15 * 64 * 4 + 192 * 4 + 48), fr),
                tensor("small_lstm: grnn + rnn_pk (mma.sync, H=64, 1.18 MFLOP/frame)", F * (524288 + 655360), sm),
                tensor("lower.gcn (tcgen05, 7.18 MFLOP/frame)", F * 7.18e6, gcn, peak=bf16,
                       peak_name="dense bf16 sustained (MEASURED_PEAKS.json)"),
                tensor("imu LSTMs (tcgen05, rnn_fast + rnn_slow)", fl, lst, peak=bf16,
                       peak_name="dense bf16 sustained (MEASURED_PEAKS.json)"),
                hbm("imu.pool (y1 planes in, s planes out)", F * (20 * 1024 * 4 + 1024 * 4), row["ms"]["imu.pool"]),
                hbm("imu.fc1 (imu in, u planes out)", F * 20 * (15 * 4 + 512 * 4), row["ms"]["imu.fc1"]),
            ]
            out["config5"].append(row)
            print(json.dumps(row), flush=True)
            del sb, data0, x, imu
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
    # compact table
    for row in out["config5"]:
        print(f"L={row['L']:3d} N={row['N']:3d} B={row['B']:4d}: " + "; ".join(
            f"{r['stage'].split(' ')[0]} {r.get('frac_issued', r.get('frac'))}" for r in row["rooflines"]))
    print(json.dumps(out["config2"]))


if __name__ == "__main__":
    main()
