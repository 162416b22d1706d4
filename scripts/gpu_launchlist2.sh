#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "default_mode or persistent" 2>&1 | tail -2
CMD="python bench.py --batch 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-half --no-config1"
timeout 200 $CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || exit 1
PER=$(python -c "import json;print(json.load(open('gpurun_out/ncu_plain_${TAG}.log'))['gpu_launches'])")
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * (PER + 1))) -c $((PER + 1)) --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
grep -c mmego gpurun_out/launches_${TAG}.csv
