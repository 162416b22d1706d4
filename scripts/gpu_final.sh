#!/bin/bash
# Round-end pass on one B200: full GPU suite, bench (+ reference arm), ncu launch list + top-kernel capture, sweep.
TAG=${1:-r02f}
bash scripts/gpu_r02.sh $TAG
bash scripts/gpu_ncu_r02.sh $TAG
timeout 600 python scripts/sweep.py > gpurun_out/sweep_${TAG}.log 2>&1; tail -2 gpurun_out/sweep_${TAG}.log | cut -c1-300
