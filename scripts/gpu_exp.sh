#!/bin/bash
mkdir -p gpurun_out
run() { timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-half "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); sm=d['stage_ms_per_step']; print('value',round(d['value']),'ms',round(d['ms_per_step'],1),'lstm fast/slow ms',sm['imu.lstm_fast'],sm['imu.lstm_slow'],'issued TF',round(d['roofline']['tensor_pipe_tflops_issued']),'clk',d['clocks']['sm_mhz'])"; }
{
set -e
for bn in 128; do for pair in 1 0; do for a in "1 2 20 20" "1 3 5 3" "1 200 20 20" "2 2 20 20"; do
  echo "bn=$bn pair=$pair $a"; TC_BN=$bn TC_PAIR=$pair timeout 60 python scripts/tc_check.py $a 2>&1 | tail -3
done; done; done
for opts in "--opt tc_bn=128" "--opt tc_bn=128 --opt tc_kb_chunk=2" "--opt tc_bn=128 --opt tc_kb_chunk=8" "--opt tc_bn=256" "--opt tc_bn=128 --imu-gemm 2" "--opt tc_bn=256 --imu-gemm 2" "--opt tc_bn=128 --opt tc_cta_pair=0"; do
  echo "$opts"; run $opts
done
} 2>&1 | tee gpurun_out/exp7.log
