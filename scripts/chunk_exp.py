"""GPU experiment: accuracy of IMU_Net (R, t vs the float64 oracle) against the TMEM drain interval tc_kb_chunk
(K blocks accumulated in TMEM before the partial sum is drained into fp32 registers; 0 = whole tile)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmego_b200 import _capi
from oracle import mmego_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = O.synth_imu_state_dict(0)
h = _capi.Handle()
h.set_weights(_capi.NET_IMU, sd)
for seed in (5, 6):
    imu = O.synth_batch(B, seed=seed)["imu"]
    R0, t0 = O.imu_forward(sd, imu, dtype=torch.float64)
    for chunk in (4, 6, 8, 12, 0):
        h.set_option("tc_kb_chunk", chunk)
        R, t = h.imu_forward(imu.cuda())
        torch.cuda.synchronize()
        print(f"seed {seed} tc_kb_chunk={chunk}: max|err| R {float((R.cpu().double() - R0).abs().max()):.2e}  t {float((t.cpu().double() - t0).abs().max()):.2e}   (tolerance R 2e-5)", flush=True)
