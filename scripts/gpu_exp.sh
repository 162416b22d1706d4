#!/bin/bash
mkdir -p gpurun_out
run() { timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-half "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],1),'lstm ms',d['stage_ms_per_step']['imu.lstm_step'],'issued TF',round(d['roofline']['tensor_pipe_tflops_issued']),'clk',d['clocks']['sm_mhz'])"; }
{
set -e
for c0 in 6 8; do for a in "1 3 5 3" "1 200 20 20"; do
  echo "chunk0=$c0 chunk=4 $a"; TC_CHUNK0=$c0 TC_CHUNK=4 timeout 60 python scripts/tc_check.py $a 2>&1 | tail -3
done; done
for opts in "--opt tc_kb_chunk0=6" "--opt tc_kb_chunk0=8" "--opt tc_kb_chunk0=4"; do
  echo "$opts"; run $opts
done
} 2>&1 | tee gpurun_out/exp6.log
