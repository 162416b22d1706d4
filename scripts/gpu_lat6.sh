#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "latency or graph_replay" 2>&1 | tail -3
{ python scripts/imu_small.py 1,2,3,4 imu_res_xchg=0; python scripts/imu_small.py 1,2,3,4 imu_res_xchg=1; python scripts/imu_small.py 1,2,3,4 imu_res_xchg=0; python scripts/imu_small.py 1,2,3,4 imu_res_xchg=1; } 2>&1 | tee gpurun_out/lat6.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_resident -s 12 -c 4 -f -o gpurun_out/prof_r02g_resident python scripts/imu_small.py 1 > gpurun_out/ncu_r02g_resident.log 2>&1
tail -3 gpurun_out/ncu_r02g_resident.log | cut -c1-200
