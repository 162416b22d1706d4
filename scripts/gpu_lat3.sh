#!/bin/bash
# tensor-core form of the resident LSTM kernel: parity, then batch-1/3 A/B against the fp32 form on one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "latency or graph_replay" 2>&1 | grep -v "^$" | tail -16
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/lat3.log | tail -20
import sys, time, torch
sys.path.insert(0, ".")
from mmego_b200.Processor.Test.Demo_test import MMEgo
for rnd in (0, 1):
    for tc in (0, 1):
        for bs in (1, 3):
            m = MMEgo(batch_size=bs, imu_surrogate=False, quiet=True)
            m.pipe.handle.set_option("imu_res_tc", tc)
            m.eval_model()
            best = 1e9
            for _ in range(2):
                m.eval_model(); best = min(best, m.seconds)
            n = m.data.shape[0]
            print(f"imu_res_tc={tc} batch={bs} graphed={m.graphed}: {best / n * 1e3:.3f} ms per snippet, {n / best:.0f} it/s, mpjpe {m.report['mpjpe_cm']:.6f}")
PY
