"""IMU_Net alone at B snippets (latency path): timing and ncu target for lstm_resident_kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
from oracle import mmego_oracle as O
h = P.make_handle()
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    h.set_option(k, int(v))
for B in [int(b) for b in sys.argv[1].split(",")]:
    imu = O.synth_batch(B, seed=B)["imu"].cuda()
    for _ in range(3):
        h.imu_forward(imu)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        h.imu_forward(imu)
    e1.record()
    torch.cuda.synchronize()
    print(f"IMU_Net B={B}: {e0.elapsed_time(e1) / 20:.3f} ms per call; error flag {h.debug_stats(reset=False)[7]}")
