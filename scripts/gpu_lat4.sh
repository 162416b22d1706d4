#!/bin/bash
# where does the resident (latency) IMU path stop paying?  IMU_Net alone, B = 1..12 snippets, both paths, device time per call
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/lat4.log | tail -30
import sys, torch
sys.path.insert(0, ".")
from tests import _parity as P
from oracle import mmego_oracle as O
h = P.make_handle()
for B in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16):
    imu = O.synth_batch(B, seed=B)["imu"].cuda()
    row = []
    for mx in (4096, 0):
        h.set_option("imu_res_max_seq", mx)
        for _ in range(3): h.imu_forward(imu)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): h.imu_forward(imu)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 10)
    print(f"B={B:3d}: resident {row[0]:.3f} ms per call ({row[0]/B:.3f} per snippet), tcgen05 {row[1]:.3f} ms per call ({row[1]/B:.3f} per snippet)")
PY
