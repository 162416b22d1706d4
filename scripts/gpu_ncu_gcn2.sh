#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-gcnsnip}
python scripts/gcn_only.py 2048 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tconv_snip|gemm_tc_kernel' -s 20 -c 7 -f -o gpurun_out/prof_${TAG} python scripts/gcn_only.py 2048 > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}
