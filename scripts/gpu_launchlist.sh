#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01final}
CMD="python bench.py --batch 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-half ${EXTRA}"
timeout 200 $CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 160 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
