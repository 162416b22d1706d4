#!/bin/bash
# first GPU contact: parity tests, smoke, a short and a full bench, launch list (ncu) of one small step
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
cat gpurun_out/smoke.log
timeout 600 python bench.py --batch 512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b512.json 2> gpurun_out/bench_b512.err; echo "exit $?"
cat gpurun_out/bench_b512.json; tail -5 gpurun_out/bench_b512.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b4096.json 2> gpurun_out/bench_b4096.err; echo "exit $?"
cat gpurun_out/bench_b4096.json; tail -5 gpurun_out/bench_b4096.err
