#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "latency or graph_replay" 2>&1 | tail -3
{ for i in 1 2; do python scripts/imu_small.py 1,2,3,4,5 imu_res_direct=0; python scripts/imu_small.py 1,2,3,4,5 imu_res_direct=1; done; python scripts/imu_small.py 4,5,6 imu_res_max_seq=0; } 2>&1 | tee gpurun_out/lat7.log
timeout 200 python - <<'PY' 2>&1 | tee -a gpurun_out/lat7.log | tail -6
import sys
sys.path.insert(0, ".")
from mmego_b200.Processor.Test.Demo_test import MMEgo
for bs in (1, 3, 4):
    m = MMEgo(batch_size=bs, imu_surrogate=False, quiet=True)
    m.eval_model()
    best = 1e9
    for _ in range(2):
        m.eval_model(); best = min(best, m.seconds)
    n = m.data.shape[0]
    print(f"MMEgo batch={bs} graphed={m.graphed}: {best / n * 1e3:.3f} ms per snippet, {n / best:.0f} it/s, mpjpe {m.report['mpjpe_cm']:.6f}")
PY
