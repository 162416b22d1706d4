#!/bin/bash
mkdir -p gpurun_out
{
set -e
echo "pair=0 sanity"; TC_PAIR=0 timeout 40 python scripts/tc_check.py 1 2 20 20 2>&1 | tail -3
for a in "1 2 20 20" "1 3 5 3" "1 40 20 20" "2 2 20 20"; do
  echo "pair=1 $a"; TC_PAIR=1 timeout 40 python scripts/tc_check.py $a 2>&1 | tail -8
done
for m in 1 2; do
timeout 120 python bench.py --steps 3 --warmup 3 --imu-gemm $m --no-cpu-baseline --no-e2e --no-half --opt tc_cta_pair=1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PAIR mode',d['dtype'],'value',d['value'],'ms',d['ms_per_step'],'lstm ms',d['stage_ms_per_step']['imu.lstm_step'],'roof',d['roofline']['frac'],d['clocks'])"
done
} 2>&1 | tee gpurun_out/tc_check5.log
