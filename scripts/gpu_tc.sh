#!/bin/bash
mkdir -p gpurun_out
{
set -e
for a in "1 2 20 20" "1 3 5 3" "1 40 20 20" "2 2 20 20"; do
  timeout 60 python scripts/tc_check.py $a 2>&1 | tail -8
done
} 2>&1 | tee gpurun_out/tc_check4.log
