#!/bin/bash
# ncu --set full of rnn_fast SECOND-layer step launches (K = 1536: the launch with the most DRAM traffic)
TAG=${1:-r02f}
CMD="python bench.py --batch 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-half --no-config1"
timeout 200 $CMD > gpurun_out/ncu_plain3_${TAG}.log 2>&1 || exit 1
# 44 H=512 launches per step (2 x 20 rnn_fast + 2 x 2 rnn_slow); step 3 starts at 88, its second layer at 108
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_step -s 110 -c 2 -f -o gpurun_out/prof_${TAG}_l2 $CMD > gpurun_out/ncu_full_${TAG}_l2.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}_l2.log
# and one rnn_slow persistent launch (19 timesteps in one launch): index 88 + 41
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_step -s 129 -c 1 -f -o gpurun_out/prof_${TAG}_slow $CMD > gpurun_out/ncu_full_${TAG}_slow.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}_slow.log
