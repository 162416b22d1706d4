#!/bin/bash
# last evidence of the round: break-even of the two IMU paths with the final resident kernel, ncu --set full of its four launches
# (after the same command ran clean without ncu), and the ncu launch list of one bench step of the final build
mkdir -p gpurun_out
{ python scripts/imu_small.py 1 && python scripts/imu_small.py 4,5,6,7 imu_res_max_seq=4096 && python scripts/imu_small.py 4,5,6,7 imu_res_max_seq=0; } 2>&1 | tee gpurun_out/lat8.log || exit 1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lstm_resident -s 12 -c 4 -f -o gpurun_out/prof_r02h_resident python scripts/imu_small.py 1 > gpurun_out/ncu_r02h_resident.log 2>&1
tail -2 gpurun_out/ncu_r02h_resident.log | cut -c1-160
bash scripts/gpu_launchlist2.sh r02h
