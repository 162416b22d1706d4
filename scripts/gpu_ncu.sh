#!/bin/bash
# ncu evidence: launch list of one small bench pass + one full capture of the top kernel (selected by $1 regex, $2 skip)
set -x
mkdir -p gpurun_out
KREGEX=${1:-gemm_ffma}
SKIP=${2:-200}
TAG=${3:-r01}
CMD="python bench.py --batch 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_launches_${TAG}.log
$CMD > gpurun_out/ncu_plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP} -c 3 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out
