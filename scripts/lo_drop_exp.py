"""GPU experiment: accuracy of IMU_Net (R, t vs the float64 oracle) as low mantissa bits of the residual planes are
rounded away (tc_lo_drop), on the seeded stand-in weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmego_b200 import _capi
from oracle import mmego_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = O.synth_imu_state_dict(0)
imu = O.synth_batch(B, seed=5)["imu"]
taps = {}
R0, t0 = O.imu_forward(sd, imu, dtype=torch.float64, taps=taps)
for drop in (0, 2, 3, 4, 5, 6):
    h = _capi.Handle()
    h.set_option("tc_lo_drop", drop)
    h.set_weights(_capi.NET_IMU, sd)
    S = B * 20
    f = torch.zeros(S, 20, 1024, device="cuda")
    h.tap("imu.f", f)
    R, t = h.imu_forward(imu.cuda())
    torch.cuda.synchronize()
    print(f"tc_lo_drop={drop}: max|err| R {float((R.cpu().double() - R0).abs().max()):.2e}  t {float((t.cpu().double() - t0).abs().max()):.2e}"
          f"  rnn_fast out {float((f.cpu().double() - taps['f']).abs().max()):.2e}   (tolerance R 2e-5)", flush=True)
