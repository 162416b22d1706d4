#!/bin/bash
# full ncu capture of one step's worth of the non-LSTM-tc kernels (mma.sync stages and the HBM-bound stages)
mkdir -p gpurun_out
TAG=${1:-r01small}
CMD="python bench.py --batch 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-half"
timeout 200 $CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_plain_${TAG}.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'imu_pool|imu_fc1|imu_decode|upper_point_mma|lower_frame_mma|lstm_rec_mma|lstm_proj_mma|gcn_agg8|assemble_metrics' \
  -s 80 -c 20 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}
