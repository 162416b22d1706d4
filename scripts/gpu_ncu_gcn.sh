#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01gcn}
CMD="python bench.py --batch 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-half"
timeout 200 $CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel' -s 28 -c 7 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
