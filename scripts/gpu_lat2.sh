#!/bin/bash
# latency path round: parity of the resident LSTM changes + graph replay test, then the batch-1 numbers (A/B of imu_res_pre and graph)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "latency or graph_replay or eval_driver or surrogate" 2>&1 | tail -6
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/lat2.log | tail -20
import sys, time, torch
sys.path.insert(0, ".")
from mmego_b200.Processor.Test.Demo_test import MMEgo
for pre in (0, 1):
    for kw in (dict(use_graph=False, fused=False), dict(use_graph=False), dict(use_graph=True)):
        for bs in (1, 3):
            m = MMEgo(batch_size=bs, imu_surrogate=False, quiet=True, **kw)
            m.pipe.handle.set_option("imu_res_pre", pre)
            m.eval_model()
            best = 1e9
            for _ in range(2):
                m.eval_model(); best = min(best, m.seconds)
            n = m.data.shape[0]
            print(f"imu_res_pre={pre} {kw} batch={bs} graphed={m.graphed}: {best / n * 1e3:.3f} ms per snippet, {n / best:.0f} it/s, mpjpe {m.report['mpjpe_cm']:.6f}")
PY
