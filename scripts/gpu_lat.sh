#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-lat}
timeout 600 python -m pytest tests -m gpu -q -x -s -k "latency or default_mode" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python - <<'PY' 2>&1 | tail -20
import sys, time, torch
sys.path.insert(0, ".")
from mmego_b200.Processor.Test.Demo_test import MMEgo
for res in (1, 0):
    for bs in (1, 4, 8):
        m = MMEgo(batch_size=bs, imu_surrogate=False, quiet=True)
        m.pipe.handle.set_option("imu_resident", res)
        m.eval_model(); m.eval_model()
        n = m.data.shape[0]
        print(f"imu_resident={res} batch={bs}: {m.seconds / n * 1e3:.3f} ms per snippet, {n / m.seconds:.0f} it/s")
PY
