#!/bin/bash
# One GPU-box visit: parity tests, point-encoder error report, bench (A/B of an option), all logs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-run}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
python scripts/point_err.py --gpu --n 16 > gpurun_out/point_err_${TAG}.log 2>&1; cat gpurun_out/point_err_${TAG}.log | tail -8
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-half > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], d["stage_ms_per_step"])
PY
