#!/bin/bash
# ncu evidence at the bench's own shapes (round 2): launch list of one bench step + one full capture of the top kernel + SASS excerpt
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python bench.py --batch 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-half --no-config1"
timeout 200 $CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_plain_${TAG}.log; exit 1; }
# skip the three warm-up steps, list the timed one (launches per step read from the plain run's own count)
PER=$(python -c "import json;print(json.load(open('gpurun_out/ncu_plain_${TAG}.log'))['gpu_launches'])")
echo "launches per step: $PER"
# one more launch per step than the library counts: torch's fill of the step's float64 sums
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * (PER + 1))) -c $((PER + 1)) --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log
timeout 200 $CMD > gpurun_out/ncu_plain2_${TAG}.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_step -s 100 -c 3 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out | grep ${TAG}
