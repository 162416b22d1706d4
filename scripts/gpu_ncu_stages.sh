#!/bin/bash
# ncu --set full on the per-frame / small-LSTM stage kernels at the bench's shapes (one launch each)
mkdir -p gpurun_out
TAG=${1:-r02stages}
CMD="python bench.py --batch 2048 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-half --no-config1"
timeout 200 $CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_plain_${TAG}.log; exit 1; }
for K in lower_frame_mma upper_point_mma imu_pool_split imu_fc1_mma 'lstm_rec_mma_kernel' 'lstm_proj_mma_kernel'; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_${K} $CMD > gpurun_out/ncu_full_${TAG}_${K}.log 2>&1
  tail -1 gpurun_out/ncu_full_${TAG}_${K}.log | cut -c1-160
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "bitonic or sweep" 2>&1 | tail -3
timeout 600 python scripts/sweep.py > gpurun_out/sweep_${TAG}.json 2> gpurun_out/sweep_${TAG}.err; tail -c 600 gpurun_out/sweep_${TAG}.json
ls -la gpurun_out | grep ${TAG}
