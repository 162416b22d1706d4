"""GPU: stage-by-stage check of the tcgen05 IMU_Net path against the oracle (run under `timeout`)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mmego_b200 import _capi
from oracle import mmego_oracle as O

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n = int(sys.argv[4]) if len(sys.argv) > 4 else 20
h = _capi.Handle()
sd = O.synth_imu_state_dict(0)
h.set_weights(_capi.NET_IMU, sd)
h.set_option("imu_gemm", mode)
h.set_option("tc_precise_act", int(os.environ.get("TC_PRECISE", "0")))
h.set_option("tc_kb_chunk", int(os.environ.get("TC_CHUNK", "4")))
h.set_option("tc_kb_chunk0", int(os.environ.get("TC_CHUNK0", "8")))
h.set_option("tc_cta_pair", int(os.environ.get("TC_PAIR", "1")))
sb = O.synth_batch(B, L=L, n_imu=n, seed=5)
imu = sb["imu"]
taps = {}
R0, t0 = O.imu_forward(sd, imu, taps=taps)
sdf = {k: v.float() for k, v in sd.items()}
y0, _, _ = O.bilstm(taps["u"], sdf, "rnn_fast.", 1)
S = B * L
dst = dict(u=torch.zeros(S, n, 512, device="cuda"), y0=torch.zeros(S, n, 1024, device="cuda"),
           f=torch.zeros(S, n, 1024, device="cuda"), s=torch.zeros(S, 1024, device="cuda"),
           g=torch.zeros(S, 1024, device="cuda"))
for k, v in dst.items():
    h.tap("imu." + k, v)
t1 = time.time()
R, t = h.imu_forward(imu.cuda())
torch.cuda.synchronize()
print(f"mode {mode} B={B} L={L} n={n}: forward {time.time() - t1:.3f}s")
ref = dict(u=taps["u"], y0=y0, f=taps["f"], s=taps["s"].reshape(S, 1024), g=taps["g"].reshape(S, 1024))
for k in ("u", "y0", "f", "s", "g"):
    d = (dst[k].cpu() - ref[k]).abs()
    print(f"  {k:3s} max|err| {float(d.max()):.3e}  mean {float(d.mean()):.3e}  ref max {float(ref[k].abs().max()):.3f}")
print(f"  R   max|err| {float((R.cpu() - R0).abs().max()):.3e}")
print(f"  t   max|err| {float((t.cpu() - t0).abs().max()):.3e}")
